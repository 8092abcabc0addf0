"""2+ ranks: the NVLS multimem gradient all-reduce kernel against NCCL's all-reduce, values and time."""
import os, sys, torch, torch.distributed as dist
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "volumetric-primitives-net_b200"))
from vpn_b200 import dist as vd
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
n = 22_875_848
out = {}
for prefer in ("nvls", "nccl"):
    ar = vd.GradientAllReduce(n, dev, prefer=prefer)
    g = torch.Generator(device=dev).manual_seed(rank)
    src = torch.randn(n, device=dev, generator=g)
    ar.buf.copy_(src); ar.launch(); ar.join(); torch.cuda.synchronize()
    out[prefer] = ar.buf.clone()
    for _ in range(3):
        ar.launch(); ar.join()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ar.launch(); ar.join()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"prefer={prefer}: mode={ar.mode} err={ar.nvls_error} {e0.elapsed_time(e1)/20:.3f} ms per all-reduce", flush=True)
if rank == 0:
    d = (out["nvls"] - out["nccl"]).abs().max().item()
    print("max |nvls - nccl| =", d, "of max |x| =", out["nccl"].abs().max().item(), flush=True)
dist.destroy_process_group()
