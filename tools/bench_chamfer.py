"""Stage timings of the Chamfer forward for every implementation (run on the GPU box)."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "volumetric-primitives-net_b200")); sys.path.insert(0, REPO)
import torch
import vpn_b200
from bench import synthetic, WORKLOADS

def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    kind, b, k, n, m, res = WORKLOADS[wl]
    dev = torch.device("cuda")
    s = {kk: (vv.to(dev) if vv is not None else None) for kk, vv in synthetic(wl, "cpu")[0].items()}
    u = torch.rand((b, k, n, 2 if kind == "sphere" else 3), device=dev)
    pts = vpn_b200.sample_primitives(kind, s["v"], s["q"], s["t"], u)
    flops = 8.0 * b * k * n * m
    peak = vpn_b200.fp32_peak_tflops(dev)
    out = {"workload": wl, "peak": peak}
    lib = vpn_b200._lib.load()
    for name, impl, prune, nb, half in (("tc_pruned", 5, 0, 0, 0), ("tc_pruned_whole_stages", 5, 0, 0, 1), ("tc_pruned_nb8", 5, 0, 8, 0),
                                        ("tc_unpruned", 5, 2, 0, 0), ("exact", 2, 0, 0, 0), ("expand", 4, 0, 0, 0)):
        lib.vpn_set_tuning(b"tc_prune", prune); lib.vpn_set_tuning(b"tc_nb", nb)
        vpn_b200.chamfer_nn_stage_ms(pts, s["target"], impl, reps=2)
        st = vpn_b200.chamfer_nn_stage_ms(pts, s["target"], impl, reps=10)
        st["tflops_total"] = flops / (st["total"] * 1e-3) / 1e12
        st["tflops_main"] = flops / (st["main"] * 1e-3) / 1e12
        st["frac_total"] = st["tflops_total"] / peak["ffma2"]
        out[name] = st
        print(name, json.dumps(st), flush=True)
    lib.vpn_set_tuning(b"tc_prune", 0); lib.vpn_set_tuning(b"tc_nb", 0)
    # uniform random clouds of the same size (worst case for the expansion filter's slack)
    p1 = torch.rand(b, k * n, 3, device=dev) - 0.5
    for name, impl in (("tc_uniform", 5), ("exact_uniform", 2), ("expand_uniform", 4)):
        st = vpn_b200.chamfer_nn_stage_ms(p1, s["target"], impl, reps=5)
        print(name, json.dumps(st), flush=True)
    print(json.dumps(out))

if __name__ == "__main__":
    main()
