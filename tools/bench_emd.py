"""EMD auction forward+backward time at the reference's training setting (eps 0.005, iters 50; train.py:188-195)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "volumetric-primitives-net_b200"))
import torch
import vpn_b200
g = torch.Generator().manual_seed(0)
for b, n in ((32, 2048), (32, 1024), (64, 2048), (8, 4096), (4, 16384)):
    x1 = torch.rand(b, n, 3, generator=g).cuda().requires_grad_(); x2 = torch.rand(b, n, 3, generator=g).cuda()
    for _ in range(2):
        d, a = vpn_b200.emd_auction(x1, x2, 0.005, 50); d.sqrt().mean().backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        d, a = vpn_b200.emd_auction(x1, x2, 0.005, 50); d.sqrt().mean().backward()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"emd B={b} n={n}: {ms:.3f} ms fwd+bwd = {b / ms * 1e3:.0f} samples/s, mean L2 {d.sqrt().mean().item():.4f}", flush=True)
