"""GCN vertex-feature pooling at the train_gcn.py shape (B = 64, N = 2048 vertices, ResNet-18 maps of a 137 x 137 view):
forward / backward time and achieved HBM bandwidth on the algorithmic bytes.  Run on the GPU box."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "volumetric-primitives-net_b200")); sys.path.insert(0, REPO)
import torch
import vpn_b200

def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n, maps = 2048, [(64, 35, 35), (128, 18, 18), (256, 9, 9), (512, 5, 5)]
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    pts = ((torch.rand(b, n, 3, device=dev, generator=g) - 0.5) * 0.9).requires_grad_()
    imgs = torch.zeros(b, 3, 137, 137, device=dev); imgs[:, :, 20:110, 15:120] = 0.5
    feats = [torch.randn(b, c, h, w, device=dev, generator=g).requires_grad_() for c, h, w in maps]
    ctot = sum(c for c, _, _ in maps)
    t_bounds = timed(lambda: vpn_b200.image_bounds(imgs))
    bounds = vpn_b200.image_bounds(imgs)
    t_fwd = timed(lambda: vpn_b200.perceptual_feature_pooling([f.detach() for f in feats], pts.detach(), bounds))
    out = vpn_b200.perceptual_feature_pooling(feats, pts, bounds)
    up = torch.randn_like(out)
    t_fb = timed(lambda: torch.autograd.grad(vpn_b200.perceptual_feature_pooling(feats, pts, bounds), [pts] + feats, up))
    fbytes = sum(f.numel() for f in feats) * 4
    obytes = out.numel() * 4
    res = {"B": b, "N": n, "channels": ctot, "bounds_ms": t_bounds, "fwd_ms": t_fwd, "fwd_bwd_ms": t_fb,
           "fwd_algorithmic_bytes": obytes + fbytes + pts.numel() * 4,
           "fwd_GBps": (obytes + fbytes + pts.numel() * 4) / (t_fwd * 1e-3) / 1e9,
           "bwd_algorithmic_bytes": obytes + 2 * fbytes + 2 * pts.numel() * 4,
           "bwd_GBps": (obytes + 2 * fbytes + 2 * pts.numel() * 4) / (max(t_fb - t_fwd, 1e-6) * 1e-3) / 1e9}
    # torch's own ops on the same device, the reference's formulation (grid_sample per map + cat + permute), for scale
    def torch_ref():
        mx, mn = pts.detach().max(1)[0], pts.detach().min(1)[0]
        sz = (pts.detach()[..., 2] - mn[:, None, 2]) / (mx[:, None, 2] - mn[:, None, 2])
        sy = (pts.detach()[..., 1] - mn[:, None, 1]) / (mx[:, None, 1] - mn[:, None, 1])
        gx = bounds[:, None, 0] + (1 - sz) * (bounds[:, None, 1] - bounds[:, None, 0])
        gy = bounds[:, None, 2] + (1 - sy) * (bounds[:, None, 3] - bounds[:, None, 2])
        grid = torch.stack([gx, gy], -1)[:, None]
        pooled = [torch.nn.functional.grid_sample(f.detach(), grid, align_corners=True) for f in feats]
        return torch.cat(pooled, 1).view(b, -1, n).permute(0, 2, 1).contiguous()
    res["torch_ops_fwd_ms"] = timed(torch_ref)
    res["max_abs_diff_vs_torch_ops"] = float((torch_ref() - out.detach()).abs().max())
    print(json.dumps(res))

if __name__ == "__main__":
    main()
