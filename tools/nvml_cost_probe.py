"""How much does an NVML query cost the GPU it samples?  Replays the C2 step graph 60 times under a polling thread
that issues one kind of query every `period` ms, and reports ms per step (CUDA events).  Run on the GPU box."""
import os, sys, threading, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "volumetric-primitives-net_b200")); sys.path.insert(0, REPO)
import torch, pynvml
import vpn_b200
from bench import synthetic, WORKLOADS

def main():
    dev = torch.device("cuda")
    kind, b, k, n, m, res = WORKLOADS["c2"]
    s = {kk: (vv.to(dev) if vv is not None else None) for kk, vv in synthetic("c2", "cpu")[0].items()}
    cfg = vpn_b200.PrimitiveLossConfig(kind=kind)
    g = vpn_b200.GraphedPrimitiveLoss(cfg, s["v"], s["q"], s["t"], s["target"], None, n_samples=n)
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    queries = {
        "none": None,
        "clock": lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
        "reasons": lambda: pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h),
        "power": lambda: pynvml.nvmlDeviceGetPowerUsage(h),
        "clock+reasons": lambda: (pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)),
    }
    steps = 60
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream()
    # extras between the steps, as bench.py has them: per-step timing events / the L2-flush memset / a side-stream hop
    for extra in ("none", "events", "events+flush", "events+flush+sidestream"):
      print("--- between steps:", extra, flush=True)
      for name, q in (("none", None), ("clock+reasons", queries["clock+reasons"])):
        for period in ((0.0,) if q is None else (0.03,)):
            stop, count, lat = [False], [0], [0.0]
            def poll():
                while not stop[0]:
                    t0 = time.perf_counter(); q(); lat[0] += time.perf_counter() - t0; count[0] += 1
                    time.sleep(period)
            for _ in range(3):
                g(s["v"], s["q"], s["t"], s["target"])
            torch.cuda.synchronize()
            th = threading.Thread(target=poll, daemon=True) if q else None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            evs = []
            for _ in range(steps):
                if "flush" in extra:
                    flush.zero_()
                if "events" in extra:
                    a = torch.cuda.Event(enable_timing=True); a.record(); evs.append(a)
                g(s["v"], s["q"], s["t"], s["target"])
                if "sidestream" in extra:
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        flush[:1024].zero_()
                    torch.cuda.current_stream().wait_stream(side)
                if "events" in extra:
                    a = torch.cuda.Event(enable_timing=True); a.record(); evs.append(a)
            e1.record()
            if th: th.start()                      # like bench.py: poll only once everything is enqueued
            torch.cuda.synchronize()
            stop[0] = True
            if th: th.join()
            ms = e0.elapsed_time(e1) / steps
            print(f"{name:14s} period {period*1e3:5.1f} ms: {ms:.4f} ms/step, {count[0]} queries, {1e3*lat[0]/max(count[0],1):.2f} ms host time per query", flush=True)

if __name__ == "__main__":
    main()
