#!/usr/bin/env python
"""Headline benchmark: primitive-loss forward+backward samples/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c1]

A step = one pass of the hot path over one synthetic batch: draw uniforms -> fused sample+pose of K
primitives -> Chamfer NN (both directions) vs the targets (+ VP-diverse Chamfer, + mesh vertices ->
soft silhouette -> L1 when the workload has a render term) -> backward to (v, q, t).
Default workload = BASELINE.json configs[1] ("c2": B=32/GPU, 16 cuboids x 4096 samples, 8192 targets).
Multi-GPU: one process per GPU under torchrun, batch sharded (weak scaling, 32 samples per GPU), plus one
NCCL all-reduce of a VPNetOneRes-sized gradient buffer (22 875 848 fp32) per step.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "volumetric-primitives-net_b200")
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    # name: (kind, B per GPU, K, N, M, render resolution or 0)          BASELINE.json configs[0..4]
    "c1": ("sphere", 1, 16, 1024, 2048, 64),
    "c2": ("cuboid", 32, 16, 4096, 8192, 0),
    "c3": ("cuboid", 32, 32, 4096, 8192, 128),
    "c4": ("cuboid", 256, 32, 4096, 8192, 128),       # B = 256 in TOTAL, split over the ranks (strong scaling)
    "c5": ("sphere", 8, 128, 128, 16384, 256),        # train_gcn.py shape: the K*128 mesh vertices are the predicted points
    # c2 as train.py:152-163 runs it: BOTH big Chamfers (view frame + canonical frame through view_to_obj_points; the
    # reference computes the second even at its default weight L_CAN_CD = 0) + VP-diverse
    "c2f": ("cuboid", 32, 16, 4096, 8192, 0),
}
STRONG = {"c4"}            # workloads whose B is the global batch
VERTEX_MODE = {"c5"}       # Chamfer on the composed mesh vertices (train_gcn.py:127-130) instead of surface samples
FAITHFUL = {"c2f"}         # canonical-frame Chamfer included (l_can_cd = 1 so that it is computed AND differentiated)


def per_gpu_batch(workload, world):
    b = WORKLOADS[workload][1]
    return max(1, b // world) if workload in STRONG else b

GRAD_NUMEL = 22_875_848          # VPNetOneRes parameters at K = 16
METRIC = "primitive-loss fwd+bwd samples/sec"


def synthetic(workload, device, seed=1234, sets=1, batch=None):
    """Network-output-shaped primitives + ShapeNet-shaped targets, fp32."""
    kind, b, k, n, m, res = WORKLOADS[workload]
    b = batch or b
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(sets):
        v = (torch.sigmoid(torch.randn(b, k, 3, generator=g)) + 0.1) / torch.tensor([8.0, 10.0, 10.0])
        q = torch.sigmoid(torch.randn(b, k, 4, generator=g))
        t = torch.tanh(torch.randn(b, k, 3, generator=g)) * 0.4
        # targets: points on the surface of a random union of boxes inside [-0.5, 0.5]^3
        centre = (torch.rand(b, 8, 3, generator=g) - 0.5) * 0.7
        half = torch.rand(b, 8, 3, generator=g) * 0.12 + 0.03
        which = torch.randint(0, 8, (b, m), generator=g)
        p = (torch.rand(b, m, 3, generator=g) * 2 - 1)
        axis = torch.randint(0, 3, (b, m), generator=g)
        sign = torch.randint(0, 2, (b, m), generator=g).float() * 2 - 1
        p.scatter_(2, axis[..., None], sign[..., None])
        bi = torch.arange(b)[:, None]
        tgt = centre[bi, which] + p * half[bi, which]
        sil = (torch.rand(b, 1, res, res, generator=g) > 0.5).float() if res else None
        d = dict(v=v, q=q, t=t, target=tgt.contiguous(), sil=sil)
        if workload in FAITHFUL:
            # object-frame copy of the targets + the view parameters that map one frame to the other (dataset.py ranges)
            d["dists"] = 1.0 + 0.5 * torch.rand(b, generator=g)
            d["elevs"] = 20.0 + 20.0 * torch.rand(b, generator=g)
            d["azims"] = 360.0 * torch.rand(b, generator=g)
            d["angles"] = 360.0 * torch.rand(b, generator=g)
            d["canon"] = (tgt * d["dists"][:, None, None]).contiguous()
        out.append(d)
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent clock / throttle sampling (NVML) during the timed region, on rank 0 only; arm() starts it.
    In a multi-rank job, sampling a GPU makes that rank reach the following all-reduces late and every other rank waits
    for it; a single process on one GPU shows no cost at all (profiles/r02_nvml_cost_probe_1gpu.txt).  Measured, ms per
    C2 step (2.94 without any sampling on 2 or 8 GPUs with the NVLS all-reduce; single runs, different boxes):
      all ranks, every 10 ms, from before the loop   8 GPUs 3.22 - 3.72     (2 GPUs, NCCL: 2.87 vs 2.86 unsampled)
      rank 0,    every 20 ms, from before the loop   8 GPUs 3.16, 3.16      <- what bench.py does
      rank 0,    every 20 ms, once the steps are enqueued      8 GPUs 3.45; 2 GPUs 3.15
      rank 0,    every 200 ms, once enqueued (one sample)      2 GPUs 3.17; with NVML warmed up first 3.34 / 4.13
    (profiles/r01w_scale_8gpu.txt, r01y_*, r02_scale_2gpu_sampling_variants.txt).  The delay does not scale with the
    number of samples and its mechanism is not understood; VPN_BENCH_NO_CLOCKS=1 switches the sampling off."""

    def __init__(self, index, enabled=True, period_s=0.03):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.period_s, self.nv = period_s, None
        self.armed, self.ready = threading.Event(), threading.Event()
        if not enabled:
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            self.ready.set()
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        self.ready.set()
        self.armed.wait()
        while True:                                  # at least one sample, taken right after arm()
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            if self.stop_flag:
                break
            time.sleep(self.period_s)

    def arm(self):
        self.armed.set()

    def stop(self):
        self.stop_flag = True
        self.armed.set()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_step(workload, rows_per_sample, threads, rows_per_slice=8192):
    """The reference's own formulation on the host (oracle port, op for op: dense (B,P,M,3) broadcast,
    modules/loss/chamfer_distance.py:14-23), fwd + bwd to (v,q,t), on a bounded sample of the workload:
    one sample, `rows_per_sample` of its K*N predicted points, taken as slices of whole primitives (the dense
    temporaries of one slice are ~1 GB each; a whole C2 sample at once needs 15 GB)."""
    from oracle import vpn_oracle as O
    kind, b, k, n, m, res = WORKLOADS[workload]
    per = max(1, rows_per_slice // n)                    # primitives per slice
    kk = min(k, max(per, (rows_per_sample // n) // per * per))
    torch.set_num_threads(threads)
    data = synthetic(workload, "cpu")[0]
    tgt = data["target"][:1]
    g = torch.Generator().manual_seed(1)
    dt = 0.0
    for k0 in range(0, kk, per):
        v, q, t = (data[x][:1, k0:k0 + per].clone().requires_grad_() for x in ("v", "q", "t"))
        u = torch.rand(1, v.shape[1], n, 2 if kind == "sphere" else 3, generator=g)
        t0 = time.perf_counter()
        pts = O.sample_predict_points(kind, v, q, t, u)
        loss = O.chamfer_dense(pts, tgt) + 0.1 * O.chamfer_dense(t, tgt, w1=0.5, w2=1.0)
        loss.backward()
        dt += time.perf_counter() - t0
    frac = (kk * n) / float(k * n)          # share of one sample's pair work that was timed
    return dt, frac, kk * n


def run_reference(args):
    """`--impl reference`: the reference's CPU path (oracle port) on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    kind, b, k, n, m, res = WORKLOADS[args.workload]
    rows = min(k * n, 32768)
    for _ in range(args.warmup):
        cpu_reference_step(args.workload, min(rows, 8192), threads)
    times = []
    for _ in range(args.steps):
        dt, frac, used = cpu_reference_step(args.workload, rows, threads)
        times.append(dt)
    mean = sum(times) / len(times)
    per_sample = mean / frac
    value = 1.0 / per_sample
    sample = (f"1 sample, {used} of {k * n} predicted points x {m} targets per step (dense torch-CPU formulation of the "
              f"reference, fwd+bwd), scaled linearly to a full sample; render term not included")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": describe(args.workload)},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def describe(workload, world=1):
    kind, b, k, n, m, res = WORKLOADS[workload]
    b = per_gpu_batch(workload, world)
    what = "template vertices" if workload in VERTEX_MODE else "samples"
    s = f"{workload}: B={b}/GPU, {k} {kind} primitives x {n} {what} (P={k * n}), Chamfer vs M={m} targets + VP-diverse"
    if workload in FAITHFUL:
        s += " + canonical-frame Chamfer through view_to_obj_points (train.py:152-163)"
    if res:
        s += f" + {res}x{res} soft-silhouette L1"
    return s + ", fwd+bwd to (v,q,t)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chamfer-impl", type=int, default=0)
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step from a CUDA graph (vpn_b200.GraphedPrimitiveLoss); auto falls back to eager launches")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import vpn_b200
    from vpn_b200 import _lib, dist as vdist
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    kind, b, k, n, m, res = WORKLOADS[args.workload]
    b = per_gpu_batch(args.workload, world)
    vertex_mode = args.workload in VERTEX_MODE
    width = 2 if kind == "sphere" else 3
    nsets = 4
    host = synthetic(args.workload, "cpu", seed=1234 + rank, sets=nsets, batch=b)
    devsets = [{kk: (vv.to(dev) if vv is not None else None) for kk, vv in s.items()} for s in host]
    pinned = [{kk: (vv.pin_memory() if vv is not None else None) for kk, vv in s.items()} for s in host]
    faithful = args.workload in FAITHFUL
    cfg = vpn_b200.PrimitiveLossConfig(kind=kind, l_sil=(1.0 if res else 0.0), chamfer_impl=args.chamfer_impl,
                                       vertex_chamfer=vertex_mode, l_can_cd=(1.0 if faithful else 0.0))

    def cams_of(s):
        return (s["dists"], s["elevs"], s["azims"], s["angles"]) if faithful else None
    step_fn = vpn_b200.PrimitiveLoss(cfg)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2
    sync = vdist.GradientAllReduce(GRAD_NUMEL, dev) if world > 1 else None

    # the bench joins the all-reduce right after launching it (nothing of this path can overlap it), so it is queued on the
    # step's own stream: no cross-stream event waits (VPN_BENCH_AR_STREAM=side restores the dedicated stream)
    ar_inline = os.environ.get("VPN_BENCH_AR_STREAM", "inline") != "side"
    graphed, graph_error = None, None
    if args.graph != "off":
        try:
            d0 = devsets[0]
            graphed = vpn_b200.GraphedPrimitiveLoss(cfg, d0["v"], d0["q"], d0["t"], d0["target"], d0["sil"], n_samples=n,
                                                    canonical_points=d0.get("canon"), cameras=cams_of(d0))
        except Exception as e:                               # noqa: BLE001 - eager launches are always available
            graph_error = repr(e)[:200]
            if args.graph == "on":
                raise

    def one_step(s, grads_out=None):
        if graphed is not None:
            loss, gv, gq, gt = graphed(s["v"], s["q"], s["t"], s["target"], s["sil"], canonical_points=s.get("canon"),
                                       cameras=cams_of(s))
            if sync is not None:
                sync.launch(inline=ar_inline)
            return loss, gv, gq, gt
        v, q, t = (s[x].detach().requires_grad_() for x in ("v", "q", "t"))
        u = None if vertex_mode else torch.rand((b, k, n, width), device=dev)      # drawn on device, like the reference
        c4 = cams_of(s) or (None, None, None, None)
        out = step_fn(v, q, t, u, s["target"], silhouettes=s["sil"], canonical_points=s.get("canon"),
                      dists=c4[0], elevs=c4[1], azims=c4[2], angles=c4[3])
        out["total"].backward()
        if sync is not None:
            sync.launch(inline=ar_inline)
        return out["total"], v.grad, q.grad, t.grad

    # ---- device-resident throughput ("value") ---------------------------------------------------
    for i in range(args.warmup):
        one_step(devsets[i % nsets])
    if sync is not None:
        sync.join()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # clock sampling: rank 0 only, every 20 ms, started just before the timed loop (VPN_BENCH_CLOCKS=armed starts it only
    # once the steps are enqueued; that measured worse on 2 and 8 GPUs - see the class docstring)
    sampler = ClockSampler(local, enabled=(rank == 0 and os.environ.get("VPN_BENCH_NO_CLOCKS", "0") != "1"), period_s=0.02)
    sampler.start()
    sampler.ready.wait(timeout=10.0)
    if os.environ.get("VPN_BENCH_CLOCKS", "early") != "armed":
        sampler.arm()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    mids = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]      # end of the rank's own compute, before the join
    do_flush = os.environ.get("VPN_BENCH_FLUSH", "1") != "0"
    launches0 = lib.vpn_launch_count()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        if do_flush:
            flush.zero_()                                           # L2 flush between timed iterations
        evs[i][0].record()
        one_step(devsets[i % nsets])
        mids[i].record()
        if sync is not None:
            sync.join()
        evs[i][1].record()
    sampler.arm()                       # no-op unless VPN_BENCH_CLOCKS=armed
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    launches = graphed.launches_per_step if graphed is not None else (lib.vpn_launch_count() - launches0) // args.steps
    sampler.stop()
    sampler.join(timeout=5.0)
    step_ms = [a.elapsed_time(bb) for a, bb in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    # per rank: mean step time and the part of it spent in the rank's own kernels (the rest is the all-reduce + waiting
    # for the slowest rank)
    own_ms = sum(a.elapsed_time(m_) for (a, _), m_ in zip(evs, mids)) / args.steps
    rank_ms = torch.tensor([sum(step_ms) / args.steps, own_ms], dtype=torch.float64, device=dev)
    rank_all = [rank_ms]
    if world > 1:
        dist.barrier()
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        rank_all = [torch.zeros_like(rank_ms) for _ in range(world)]
        dist.all_gather(rank_all, rank_ms)
    rank_all = [[round(float(x), 4) for x in r_.tolist()] for r_ in rank_all]
    total_ms = float(total_ms.item())
    value = world * b * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API with host buffers ("e2e") ------------------------------
    # results land in pinned host buffers with asynchronous copies and ONE synchronisation per step
    host_out = [torch.empty((), pin_memory=True), torch.empty((b, k, 3), pin_memory=True),
                torch.empty((b, k, 4), pin_memory=True), torch.empty((b, k, 3), pin_memory=True)]
    done = torch.cuda.Event()

    def e2e_step(ps):
        # with the graph the pinned inputs are copied straight into its static buffers; eager: fresh device tensors
        s = ps if graphed is not None else {kk: (vv.to(dev, non_blocking=True) if vv is not None else None) for kk, vv in ps.items()}
        res = one_step(s)
        if sync is not None:
            sync.join()
        for h, d in zip(host_out, res):
            h.copy_(d.detach(), non_blocking=True)
        done.record()
        done.synchronize()
        return host_out

    for i in range(2):
        e2e_step(pinned[i % nsets])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, args.steps // 2)
    t0 = time.perf_counter()
    e0.record()
    for i in range(e2e_steps):
        e2e_step(pinned[i % nsets])
    e1.record()
    torch.cuda.synchronize()
    e2e_wall = time.perf_counter() - t0
    e2e_ms = torch.tensor([max(e0.elapsed_time(e1), e2e_wall * 1e3)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * b * e2e_steps / (float(e2e_ms.item()) * 1e-3)
    h2d = sum(vv.numel() * vv.element_size() for vv in pinned[0].values() if vv is not None)
    d2h = 4 + b * k * 10 * 4

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel group: Chamfer forward (FP32 pipe) ----------------------
        s = devsets[0]
        with torch.no_grad():
            if vertex_mode:
                from vpn_b200 import templates as _tpl
                pts = vpn_b200.mesh_vertices(_tpl.template(kind, dev)[0], s["v"], s["q"], s["t"])
            else:
                pts = vpn_b200.sample_primitives(kind, s["v"], s["q"], s["t"], torch.rand((b, k, n, width), device=dev))
            reps = 10
            vpn_b200.chamfer_nn_stage_ms(pts, s["target"], args.chamfer_impl, reps=3)
            stage = vpn_b200.chamfer_nn_stage_ms(pts, s["target"], args.chamfer_impl, reps=reps)
            cham_ms = stage["total"]
            # sampling kernel: a stream of launches over rotating uniform buffers larger than L2 in total (every
            # launch reads cold inputs; no write-flush, whose dirty lines would be written back during the launch)
            sbytes = b * k * n * (12 + 4 * width)
            nrot = min(64, max(2, -(-300_000_000 // sbytes)))
            nl = 4 * nrot
            us = [torch.rand((b, k, n, width), device=dev) for _ in range(nrot)]
            samp_ms = vpn_b200.sample_primitives_ms(kind, s["v"], s["q"], s["t"], us, reps=nl)
            del us
        peak = vpn_b200.fp32_peak_tflops(dev)
        flops = 8.0 * b * (k * n) * m                      # 8 flop per (predicted, target) pair, both directions
        main_ms = stage["main"] if stage["main"] > 0 else cham_ms
        achieved = flops / (main_ms * 1e-3) / 1e12
        sm_mhz_max = sampler.summary()["sm_max_mhz"] or 1965
        nominal = 148 * 128 * 2 * sm_mhz_max * 1e6 / 1e12
        peak_tf = max(peak["ffma2"], peak["ffma"])
        main_kernel = vpn_b200.chamfer_main_kernel_name(b, k * n, m, args.chamfer_impl)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))[args.workload][main_kernel]
        except Exception:
            pass
        roofline = {"kernel": main_kernel + " (main kernel of vpn_chamfer_fwd; both Chamfer directions in one launch)",
                    "bound": "fp32", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "peak_source": "live FFMA2 stream probe (vpn_fp32_peak_probe), burst; MEASURED_PEAKS.json has no FP32 "
                                   "entry; nominal 148 SM x 128 lanes x 2 x max clock given beside it",
                    "peak_nominal": nominal, "frac_of_nominal": achieved / nominal, "ms": main_ms,
                    "algorithmic_flops": flops, "traffic": traffic,
                    "traffic_source": "profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full "
                                      "capture of this launch)" if traffic else None,
                    "note": "achieved = 8 flop per (predicted, target) pair / kernel time, against the FP32 FMA peak the "
                            "north star names; chamfer_tc_kernel evaluates the pairs on the tensor cores (fp16-split "
                            "operands, fp32 accumulate) and is bounded by TMEM reads + FMNMX on the ALU pipe, see DESIGN.md 4.1",
                    "forward_total": {"ms": cham_ms, "stages_ms": stage, "achieved": flops / (cham_ms * 1e-3) / 1e12,
                                      "frac": flops / (cham_ms * 1e-3) / 1e12 / peak_tf,
                                      "note": "main kernel + exact recovery kernels: the time to the final min / arg-min"}}
        hbm_peak, tensor_peak = 6536.7, None
        try:
            mp = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
            hbm_peak, hbm_src = mp["hbm_gbs"], "MEASURED_PEAKS.json"
            tensor_peak = mp.get("bf16_tflops")
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback"
        if main_kernel == "chamfer_tc_kernel":
            # what the tensor pipe actually executes: a 16-term fp16 dot product per pair and direction (2 x 16 MACs = 64 flop)
            executed = 64.0 * b * (k * n) * m / (main_ms * 1e-3) / 1e12
            roofline["tensor_executed"] = {"tflops": executed, "peak_dense_16bit": tensor_peak,
                                           "frac": (executed / tensor_peak) if tensor_peak else None,
                                           "note": "fp16 operands, fp32 accumulate, K = 16; padded tiles not counted"}
        roof_s = {"kernel": "pose_fwd_kernel (fused sample+pose)", "bound": "hbm",
                  "achieved": sbytes / (samp_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                  "frac": sbytes / (samp_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src, "ms": samp_ms,
                  "algorithmic_bytes": sbytes, "traffic": None,
                  "how": f"{nl} back-to-back launches over {nrot} rotating uniform buffers ({nrot * sbytes // 2 >> 20} MB of inputs > L2), "
                         "one CUDA-event pair around the stream (vpn_pose_points_fwd_timed)"}
        roof_r = None
        if res:
            from vpn_b200 import templates
            with torch.no_grad():
                tv, tf = templates.template(kind, dev)
                verts = vpn_b200.mesh_vertices(tv, s["v"], s["q"], s["t"])
                faces = step_fn.composed_faces(k, dev)
                rot, pos = vpn_b200.look_at_cameras(torch.zeros(b, device=dev), torch.zeros(b, device=dev), torch.ones(b, device=dev))
                for _ in range(2):
                    vpn_b200.soft_silhouette(verts, faces, rot, pos, res, res)
                re_ = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
                for i in range(reps):
                    flush.zero_()
                    re_[i][0].record()
                    vpn_b200.soft_silhouette(verts, faces, rot, pos, res, res)
                    re_[i][1].record()
                torch.cuda.synchronize()
                r_ms = sum(a.elapsed_time(bb) for a, bb in re_) / reps
            rbytes = 12 * b * verts.shape[1] + 12 * faces.shape[0] + 4 * b * res * res
            roof_r = {"kernel": "sil_project_kernel + sil_faces_kernel + sil_raster_fwd_kernel (vpn_silhouette_fwd, whole batch)",
                      "bound": "hbm", "achieved": rbytes / (r_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                      "frac": rbytes / (r_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src, "ms": r_ms,
                      "algorithmic_bytes": rbytes, "traffic": None, "faces": int(faces.shape[0]),
                      "pixel_face_tests_per_s": b * res * res * float(faces.shape[0]) / (r_ms * 1e-3),
                      "note": "algorithmic bytes are ~1 us of HBM time: the rasteriser is latency / ALU bound (DESIGN.md 4.4)"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rows = min(k * n, 65536)
            cpu_reference_step(args.workload, min(rows, n), threads)      # warm-up on one primitive
            dt, frac, used = cpu_reference_step(args.workload, rows, threads)
            cpu = {"value": 1.0 / (dt / frac), "unit": "samples/s", "cores": threads, "kind": "port",
                   "sample": f"1 sample, {used} of {k * n} predicted points x {m} targets, dense torch-CPU formulation "
                             f"of the reference (fwd+bwd, no render term), {dt:.1f} s, scaled linearly to a full sample"}
        line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.workload in STRONG else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": describe(args.workload, world), "global_batch": world * b,
                           "parallelism": f"dp{world}", "l2": "256 MB L2 flush between timed iterations, outside the "
                           "per-step CUDA-event pairs; 4 rotating input sets",
                           "cuda_graph": graphed is not None, "cuda_graph_error": graph_error,
                           "allreduce_numel": GRAD_NUMEL if world > 1 else 0,
                           "allreduce": (sync.mode if sync is not None else None),
                           "allreduce_stream": (("step stream" if ar_inline else "dedicated stream") if sync is not None else None),
                           "allreduce_trial_ms": (sync.trial_ms if sync is not None else None),
                           "wall_ms_per_step_incl_flush": wall * 1e3 / args.steps,
                           "rank_ms_step_and_own_kernels": rank_all, "l2_flush": do_flush},
                "clocks": sampler.summary(), "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "steps": e2e_steps},
                "roofline": roofline, "roofline_sampling": roof_s, "roofline_raster": roof_r, "cpu_baseline": cpu,
                "fp32_peak_probe_tflops": peak}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
