"""Time the per-step gradient all-reduce (22 875 848 fp32) alone: torchrun --nproc-per-node N tools/allreduce_probe.py"""
import os, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
buf = torch.zeros(22_875_848, device=dev)
for numel in (22_875_848,):
    x = buf[:numel]
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if dist.get_rank() == 0:
        w = dist.get_world_size()
        print(f"NCCL_ALGO={os.environ.get('NCCL_ALGO','default')} PROTO={os.environ.get('NCCL_PROTO','default')} world {w}: {numel*4/1e6:.1f} MB all-reduce {ms:.3f} ms, "
              f"algbw {numel*4/ms/1e6:.1f} GB/s, busbw {numel*4/ms/1e6*2*(w-1)/w:.1f} GB/s", flush=True)
dist.destroy_process_group()
