"""Where the end-to-end loop's time goes: graphed C2 step on device inputs vs vpn_b200.HostPipeline, with / without the L2
flush and the result read-back.  Run on the GPU box."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "volumetric-primitives-net_b200")); sys.path.insert(0, REPO)
import torch
import vpn_b200 as vpn
from bench import synthetic, WORKLOADS

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
kind, b, k, n, m, res = WORKLOADS[wl]
dev = torch.device("cuda")
host = synthetic(wl, "cpu", sets=4)
pinned = [{kk: (vv.pin_memory() if vv is not None else None) for kk, vv in s.items()} for s in host]
devs = [{kk: (vv.to(dev) if vv is not None else None) for kk, vv in s.items()} for s in host]
cfg = vpn.PrimitiveLossConfig(kind=kind, l_sil=(1.0 if res else 0.0), vertex_chamfer=(wl == "c5"))
d0 = devs[0]
gr = vpn.GraphedPrimitiveLoss(cfg, d0["v"], d0["q"], d0["t"], d0["target"], d0["sil"], n_samples=n)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
step = lambda s: gr(s["v"], s["q"], s["t"], s["target"], s["sil"])
N = 20


def timed(fn, label):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / N
    print("%-58s %.4f ms / step" % (label, dt * 1e3), flush=True)


def direct(do_flush):
    def f():
        for i in range(N):
            if do_flush: flush.zero_()
            step(devs[i % 4])
    return f


def piped(do_flush, batches):
    pipe = vpn.HostPipeline(step, pinned[0], dev, pre_step=(flush.zero_ if do_flush else None))
    def f():
        pipe.submit(batches[0])
        for i in range(1, N):
            pipe.submit(batches[i % 4]); pipe.result()
        pipe.result()
    return f


def serial(do_flush):
    outs = None
    def f():
        for i in range(N):
            if do_flush: flush.zero_()
            o = step(pinned[i % 4])
            h = [x.detach().to("cpu", non_blocking=True) for x in o]
            torch.cuda.synchronize()
    return f

timed(direct(False), "device inputs, no flush, no read-back")
timed(direct(True), "device inputs, flush")
timed(piped(False, pinned), "HostPipeline, no flush")
timed(piped(True, pinned), "HostPipeline, flush")
timed(piped(True, devs), "HostPipeline fed device tensors (no PCIe input), flush")
timed(serial(True), "serial: copy, step, read back, synchronise; flush")
