"""A/B probe: stage timings of vpn_chamfer_fwd_timed (impl 5) from two builds of libvpn_b200.so on the same data.
usage: python tools/ab_chamfer.py libA.so libB.so"""
import ctypes, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "volumetric-primitives-net_b200")); sys.path.insert(0, REPO)
import torch
from bench import synthetic, WORKLOADS
import vpn_b200

kind, b, k, n, m, res = WORKLOADS["c2"]
dev = torch.device("cuda")
s = {kk: (vv.to(dev) if vv is not None else None) for kk, vv in synthetic("c2", "cpu")[0].items()}
pts = vpn_b200.sample_primitives(kind, s["v"], s["q"], s["t"], torch.rand((b, k, n, 3), device=dev)).contiguous()
tgt = s["target"].contiguous()
P = k * n
for path in sys.argv[1:]:
    lib = ctypes.CDLL(path)
    nb = ctypes.c_size_t(0)
    lib.vpn_chamfer_workspace_bytes(b, P, m, 5, ctypes.byref(nb))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    o = [torch.empty(b, P, device=dev), torch.empty(b, P, dtype=torch.int32, device=dev), torch.empty(b, m, device=dev), torch.empty(b, m, dtype=torch.int32, device=dev)]
    ms = (ctypes.c_float * 4)()
    vp = ctypes.c_void_p
    lib.vpn_chamfer_fwd_timed.argtypes = [vp] * 6 + [ctypes.c_int] * 3 + [vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), vp]
    for reps in (2, 10):
        rc = lib.vpn_chamfer_fwd_timed(pts.data_ptr(), tgt.data_ptr(), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(),
                                       b, P, m, ws.data_ptr(), nb.value, 5, reps, ms, None)
    print(os.path.basename(path), "rc", rc, "main %.3f fallback %.3f rows %.3f cols %.3f" % tuple(ms), flush=True)
    if hasattr(lib, "vpn_set_tuning"):
        lib.vpn_set_tuning(b"tc_prune", 2)
        for reps in (2, 10):
            lib.vpn_chamfer_fwd_timed(pts.data_ptr(), tgt.data_ptr(), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(),
                                      b, P, m, ws.data_ptr(), nb.value, 5, reps, ms, None)
        print(os.path.basename(path), "unpruned: main %.3f fallback %.3f rows %.3f cols %.3f" % tuple(ms), flush=True)
        lib.vpn_set_tuning(b"tc_prune", 0)
