"""Small driver for ncu: a few forwards of one Chamfer implementation at a reduced batch."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "volumetric-primitives-net_b200")); sys.path.insert(0, REPO)
import torch
import vpn_b200
from bench import synthetic, WORKLOADS
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 4
b = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kind, _, k, n, m, res = WORKLOADS["c2"]
dev = torch.device("cuda")
s = {kk: (vv[:b].to(dev) if vv is not None else None) for kk, vv in synthetic("c2", "cpu")[0].items()}
u = torch.rand((b, k, n, 3), device=dev)
pts = vpn_b200.sample_primitives(kind, s["v"], s["q"], s["t"], u)
for _ in range(3):
    vpn_b200.chamfer_nn(pts, s["target"], impl)
torch.cuda.synchronize()
print("ok")
