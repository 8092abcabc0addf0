"""Sweep of the pruning-bounds knobs (near blocks x representatives, both directions): stage timings and the share of
128 x 128 blocks skipped.  Run on the GPU box: python tools/prep_sweep.py [c2|c3|c5]"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "volumetric-primitives-net_b200")); sys.path.insert(0, REPO)
import torch
import vpn_b200
from bench import synthetic, WORKLOADS, VERTEX_MODE


def points_of(wl, dev):
    kind, b, k, n, m, res = WORKLOADS[wl]
    s = {kk: (vv.to(dev) if vv is not None else None) for kk, vv in synthetic(wl, "cpu")[0].items()}
    if wl in VERTEX_MODE:
        from vpn_b200 import ops, templates
        tv, _ = templates.template(kind, dev)
        return ops.mesh_vertices(tv, s["v"], s["q"], s["t"]).detach(), s["target"]
    u = torch.rand((b, k, n, 2 if kind == "sphere" else 3), device=dev)
    return vpn_b200.sample_primitives(kind, s["v"], s["q"], s["t"], u), s["target"]


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    dev = torch.device("cuda")
    pts, tgt = points_of(wl, dev)
    lib = vpn_b200._lib.load()
    lib.vpn_set_tuning(b"prep_probe", 1)
    vpn_b200.chamfer_nn_stage_ms(pts, tgt, 5, reps=2)
    st = vpn_b200.chamfer_nn_stage_ms(pts, tgt, 5, reps=3)
    lib.vpn_set_tuning(b"prep_probe", 0)
    c = st["tc_counters"]
    print(wl, "sort kernel phases, clk: target CTA box %d cells %d sort %d copies %d chunk boxes %d | row CTA stage+box %d cells %d sort %d block boxes %d copy %d" % tuple(c[2:12]), flush=True)
    if len(sys.argv) > 2 and sys.argv[2] == "probe":
        return
    keys = (b"prep_near_rows", b"prep_reps_rows", b"prep_near_cols", b"prep_reps_cols")
    for cfg in ((0, 0, 0, 0), (8, 4, 16, 4), (8, 4, 16, 2), (8, 4, 8, 4), (8, 4, 8, 2), (8, 4, 4, 4), (4, 4, 8, 4), (8, 2, 8, 4), (4, 2, 8, 2), (16, 4, 32, 4)):
        for k, v in zip(keys, cfg):
            lib.vpn_set_tuning(k, v)
        vpn_b200.chamfer_nn_stage_ms(pts, tgt, 5, reps=2)
        st = vpn_b200.chamfer_nn_stage_ms(pts, tgt, 5, reps=10)
        print(wl, "near/reps rows %d x %d cols %d x %d" % cfg, "main %.4f rows %.4f cols %.4f total %.4f skipped %.4f" % (
            st["main"], st["rows"], st["cols"], st["total"], st["stages_skipped"] / max(1, st["stages"])), flush=True)
    for k in keys:
        lib.vpn_set_tuning(k, 0)


if __name__ == "__main__":
    main()
