"""Per-feature-map time of the pooling backward kernels at the train_gcn.py shape (B = 64, N = 2048).
VPN_POOL_BWD_ATOMIC=1 forces the shared-memory-atomic kernel for every map.  Run on the GPU box."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "volumetric-primitives-net_b200")); sys.path.insert(0, REPO)
import torch
import vpn_b200
from vpn_b200 import _lib
from vpn_b200._lib import ptr, stream_ptr, check

b, n = 64, 2048
maps = [(64, 35, 35), (128, 18, 18), (256, 9, 9), (512, 5, 5)]
dev = torch.device("cuda")
lib = _lib.load()
pts = (torch.rand(b, n, 3, device=dev) - 0.5) * 0.9
bounds = torch.tensor([[-0.8, 0.75, -0.7, 0.6]], device=dev).repeat(b, 1).contiguous()
rng = torch.empty(b, 4, device=dev); arg = torch.empty(b, 4, dtype=torch.int32, device=dev)
check(lib.vpn_points_yz_range(ptr(pts), ptr(rng), ptr(arg), b, n, stream_ptr(dev)), "range")
ctot = sum(c for c, _, _ in maps)
gout = torch.randn(b, n, ctot, device=dev)
ggrid = torch.zeros(b, n, 2, device=dev)
coff = 0
for c, h, w in maps:
    f = torch.randn(b, c, h, w, device=dev); gf = torch.empty_like(f)
    import ctypes
    nws = ctypes.c_size_t(0)
    check(lib.vpn_feature_pool_bwd_workspace_bytes(b, n, ctypes.byref(nws)), "ws")
    ws = torch.empty(nws.value, dtype=torch.uint8, device=dev)
    def run_old():
        check(lib.vpn_feature_pool_bwd(ptr(f), ptr(pts), ptr(bounds), ptr(rng), ptr(gout), ptr(gf), ptr(ggrid), b, c, h, w, n, ctot, coff,
                                       stream_ptr(dev)), "bwd")
    def run_new():
        check(lib.vpn_feature_pool_bwd_sorted(ptr(f), ptr(pts), ptr(bounds), ptr(rng), ptr(gout), ptr(gf), ptr(ggrid), ptr(ws), nws.value,
                                              b, c, h, w, n, ctot, coff, stream_ptr(dev)), "bwd sorted")
    res = []
    for run in (run_old, run_new):
        for _ in range(3): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 20)
    print(f"map {c}x{h}x{w}: bwd shared-memory-atomic kernel {res[0]:.4f} ms, cell-sorted kernels {res[1]:.4f} ms", flush=True)
    coff += c
