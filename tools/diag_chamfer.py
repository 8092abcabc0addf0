"""Mismatch counts of the tensor-core Chamfer forward against the generic exact kernel (run on the GPU box)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "volumetric-primitives-net_b200")); sys.path.insert(0, REPO)
import torch
import vpn_b200

def main():
    lib = vpn_b200._lib.load()
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(3)
    for (b, p, m, kind) in ((2, 16384, 4096, "patches"), (2, 4096, 2048, "uniform"), (1, 2048, 8192, "uniform"), (2, 65536, 8192, "patches")):
        if kind == "uniform":
            p1 = torch.rand(b, p, 3, generator=g) - 0.5; p2 = torch.rand(b, m, 3, generator=g) - 0.5
        else:
            k = 16; n = p // k
            c = (torch.rand(b, k, 1, 3, generator=g) - 0.5) * 0.8
            p1 = (c + (torch.rand(b, k, n, 3, generator=g) - 0.5) * 0.2).reshape(b, p, 3)
            p2 = (torch.rand(b, m, 3, generator=g) - 0.5)
        p1 = p1.to(dev).contiguous(); p2 = p2.to(dev).contiguous()
        ref = [t.clone() for t in vpn_b200.chamfer_nn(p1, p2, 1)]
        for prune, half in ((0, 0), (0, 1), (0, 2), (0, 3), (0, 5), (2, 0)):
            for nb in (0, 8, 4):
                lib.vpn_set_tuning(b"tc_prune", prune); lib.vpn_set_tuning(b"tc_nb", nb); lib.vpn_set_tuning(b"tc_hunits", half)   # half: 1 + units on the FP16 pipe, 0 = default
                out = vpn_b200.chamfer_nn(p1, p2, 5)
                torch.cuda.synchronize()
                bad = [int((o != r).sum()) if o.dtype != torch.float32 else int((o.view(torch.int32) != r.view(torch.int32)).sum()) for o, r in zip(out, ref)]
                line = f"{kind} B={b} P={p} M={m} prune={prune} hunits+1={half} nb={nb}: mismatches min1 {bad[0]} idx1 {bad[1]} min2 {bad[2]} idx2 {bad[3]}"
                if bad[1]:
                    w = (out[1] != ref[1]).nonzero()[:4]
                    line += "  first idx1: " + str([(int(a), int(bb), int(out[1][a, bb]), int(ref[1][a, bb]), float(out[0][a, bb]), float(ref[0][a, bb])) for a, bb in w])
                if bad[3]:
                    w = (out[3] != ref[3]).nonzero()[:4]
                    line += "  first idx2: " + str([(int(a), int(bb), int(out[3][a, bb]), int(ref[3][a, bb]), float(out[2][a, bb]), float(ref[2][a, bb])) for a, bb in w])
                print(line, flush=True)
        lib.vpn_set_tuning(b"tc_prune", 0); lib.vpn_set_tuning(b"tc_nb", 0); lib.vpn_set_tuning(b"tc_hunits", 0)

if __name__ == "__main__":
    main()
