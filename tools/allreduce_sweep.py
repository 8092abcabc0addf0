"""Sweep of the NVLS all-reduce kernel variants (torchrun --nproc-per-node N tools/allreduce_sweep.py): 91.5 MB gradient
buffer, 30 launches per variant, max over ranks; NCCL beside it.  Every variant is checked against the expected sum."""
import os, sys, torch, torch.distributed as dist
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "volumetric-primitives-net_b200"))
from vpn_b200 import dist as vd, _lib
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
lib = _lib.load()
n = 22_875_848
ar = vd.GradientAllReduce(n, dev, prefer="nvls")
assert ar.graph_capturable, ar.nvls_error
scratch = torch.zeros(n, device=dev)


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


results = []
# variant 8 = loads-then-stores (0 selects the library default), grid div 1 = every SM
combos = [(v, c, 512, 1) for v in (8, 1, 2, 4) for c in (4, 2, 1)]
combos += [(v, 1, th, dv) for v in (8, 1, 2) for th in (512, 256, 128) for dv in (1, 2, 4) if not (th == 512 and dv == 1)]
combos += [(0, 0, 0, 0)]                # the library default
for variant, ctas, threads, div in combos:
    if True:
        lib.vpn_set_tuning(b"ar_variant", variant); lib.vpn_set_tuning(b"ar_ctas", ctas)
        lib.vpn_set_tuning(b"ar_threads", threads); lib.vpn_set_tuning(b"ar_grid_div", div)
        ar.buf.fill_(float(rank + 1)); torch.cuda.synchronize(); dist.barrier()
        ar.launch(inline=True); torch.cuda.synchronize()
        ok = bool((ar.buf == float(world * (world + 1) // 2)).all()) and not ar.nvls_timed_out()
        ar.buf.zero_()
        ms = timed(lambda: ar.launch(inline=True))
        results.append((variant, ctas, ms, ok))
        if rank == 0:
            print(f"variant {variant} ctas/SM {ctas} threads {threads} grid/{div}: {ms:.4f} ms  busbw {n * 4 / ms / 1e6 * 2 * (world - 1) / world:.0f} GB/s  correct={ok}", flush=True)
for key in (b"ar_variant", b"ar_ctas", b"ar_threads", b"ar_grid_div"):
    lib.vpn_set_tuning(key, 0)
ms = timed(lambda: dist.all_reduce(scratch))
if rank == 0:
    print(f"NCCL all_reduce: {ms:.4f} ms", flush=True)
dist.barrier(); dist.destroy_process_group()
