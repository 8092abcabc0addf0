"""Quick GPU check of the tensor-core Chamfer filter against the C oracle + stage timings (run on the GPU box)."""
import ctypes, json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "volumetric-primitives-net_b200")); sys.path.insert(0, REPO)
import torch
import vpn_b200
from bench import synthetic, WORKLOADS

lib = ctypes.CDLL(os.path.join(REPO, "oracle", "_build", "libvpn_oracle.so"))
def oracle(p1, p2):
    p1 = np.ascontiguousarray(p1, np.float32); p2 = np.ascontiguousarray(p2, np.float32)
    b, p, _ = p1.shape; m = p2.shape[1]
    m1 = np.empty((b, p), np.float32); i1 = np.empty((b, p), np.int64); m2 = np.empty((b, m), np.float32); i2 = np.empty((b, m), np.int64)
    f = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.vpn_oracle_chamfer_nn(f(p1), f(p2), b, p, m, f(m1), f(i1), f(m2), f(i2), 0) == 0
    return m1, i1, m2, i2

gen = torch.Generator().manual_seed(5)
for (b, p, m) in ((1, 512, 128), (2, 2048, 640), (1, 16384, 2048), (2, 65536, 8192)):
    p1 = torch.rand(b, p, 3, generator=gen) - 0.5; p2 = torch.rand(b, m, 3, generator=gen) - 0.5
    ref = oracle(p1.numpy(), p2.numpy())
    got = vpn_b200.chamfer_nn(p1.cuda(), p2.cuda(), 5)
    torch.cuda.synchronize()
    ok = [bool((g.cpu().numpy().astype(r.dtype) == r).all()) for g, r in zip(got, ref)]
    print("shape", (b, p, m), "min1/idx1/min2/idx2 equal:", ok, flush=True)
    if not all(ok):
        for name, g, r in zip(("min1", "idx1", "min2", "idx2"), got, ref):
            g = g.cpu().numpy().astype(r.dtype)
            print("  ", name, "mismatches", int((g != r).sum()), "of", r.size)

wl = "c2"
kind, b, k, n, m, res = WORKLOADS[wl]
dev = torch.device("cuda")
s = {kk: (vv.to(dev) if vv is not None else None) for kk, vv in synthetic(wl, "cpu")[0].items()}
u = torch.rand((b, k, n, 3), device=dev)
pts = vpn_b200.sample_primitives(kind, s["v"], s["q"], s["t"], u)
flops = 8.0 * b * k * n * m
for name, impl in (("expand", 4), ("tc", 5)):
    vpn_b200.chamfer_nn_stage_ms(pts, s["target"], impl, reps=2)
    st = vpn_b200.chamfer_nn_stage_ms(pts, s["target"], impl, reps=10)
    st["tflops_main"] = flops / (st["main"] * 1e-3) / 1e12
    st["tflops_total"] = flops / (st["total"] * 1e-3) / 1e12
    print(name, json.dumps(st), flush=True)
a = vpn_b200.chamfer_nn(pts, s["target"], 4); c = vpn_b200.chamfer_nn(pts, s["target"], 5)
print("c2 tc == expand:", [bool(torch.equal(x, y)) for x, y in zip(a, c)])
