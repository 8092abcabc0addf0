"""Stage timings of the tensor-core Chamfer forward against the number of 32-column units reduced on the FP16 pipe
(vpn_set_tuning("tc_hunits", 1 + units)).  Run on the GPU box: python tools/tc_sweep.py [c2|c3|c5]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "volumetric-primitives-net_b200")); sys.path.insert(0, REPO)
import torch
import vpn_b200
sys.path.insert(0, os.path.join(REPO, "tools"))
from prep_sweep import points_of


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    dev = torch.device("cuda")
    pts, tgt = points_of(wl, dev)
    lib = vpn_b200._lib.load()
    sweep = ((1, 0), (3, 0), (5, 0)) if (len(sys.argv) > 2 and sys.argv[2] == "hunits") else ((0, 0), (0, 8), (0, 16), (0, 0))
    for h, nb in sweep:
        lib.vpn_set_tuning(b"tc_hunits", h); lib.vpn_set_tuning(b"tc_nb", nb)
        vpn_b200.chamfer_nn_stage_ms(pts, tgt, 5, reps=2)
        st = vpn_b200.chamfer_nn_stage_ms(pts, tgt, 5, reps=10)
        print(wl, "tc_hunits %d tc_nb %d: main %.4f rows %.4f cols %.4f total %.4f skipped %.4f" % (
            h, nb, st["main"], st["rows"], st["cols"], st["total"], st["stages_skipped"] / max(1, st["stages"])), flush=True)
    lib.vpn_set_tuning(b"tc_hunits", 0); lib.vpn_set_tuning(b"tc_nb", 0)


if __name__ == "__main__":
    main()
