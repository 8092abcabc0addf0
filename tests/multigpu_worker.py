"""Worker of tests/test_gpu_multigpu.py, launched by torchrun with one process per GPU (NCCL over NVLink).

Checks, on every rank, and exits non-zero on any failure:
  1. 1-vs-G parity (SURVEY.md section 4 layer 4): the batch-sharded primitive-loss step, local loss scaled by
     local_loss_scale and the (v,q,t)-gradient chained through a shared 'network' parameter vector and SUM-all-reduced,
     equals the single-GPU step on the whole batch (loss to 1e-6 relative, parameter gradient to 1e-5).
  2. The gradient all-reduce: our self-synchronising NVLS multimem kernel against NCCL's all-reduce on the same
     (integer-valued, hence order-independent) data - bit-identical, also when the kernel is replayed from a CUDA
     graph - and no cross-rank wait timed out.
"""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "volumetric-primitives-net_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def heads(w, b, k):
    """Stand-in for the network heads: parameter vector w (10,) -> (v, q, t) of shape (B,K,.), differentiable."""
    z = torch.linspace(-1, 1, b * k * 10, device=w.device).view(b, k, 10)
    h = z * w[None, None, :] + w.flip(0)[None, None, :] * 0.1
    v = (torch.sigmoid(h[..., 0:3]) + 0.1) / torch.tensor([8.0, 10.0, 10.0], device=w.device)
    return v.contiguous(), torch.sigmoid(h[..., 3:7]).contiguous(), (torch.tanh(h[..., 7:10]) * 0.3).contiguous()


def main():
    import vpn_b200
    from vpn_b200 import dist as vd
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    fails = []

    # ---- 1. sharded step == single-GPU step -------------------------------------------------------------
    gb, k, n, m, res = 8 * world, 4, 256, 512, 32
    g = torch.Generator().manual_seed(11)
    w0 = torch.randn(10, generator=g)
    u = torch.rand(gb, k, n, 2, generator=g)
    tgt = (torch.rand(gb, m, 3, generator=g) - 0.5) * 0.8
    sil = (torch.rand(gb, 1, res, res, generator=g) > 0.5).float()
    cfg = vpn_b200.PrimitiveLossConfig(kind="sphere", l_sil=1.0)
    step = vpn_b200.PrimitiveLoss(cfg)
    # whole batch on this GPU (every rank computes it: the reference is single process)
    w = w0.clone().to(dev).requires_grad_()
    v, q, t = heads(w, gb, k)
    full = step(v, q, t, u.to(dev), tgt.to(dev), silhouettes=sil.to(dev))["total"]
    full.backward()
    g_full = w.grad.clone()
    # this rank's shard
    lo, hi = vd.shard_range(gb, rank, world)
    w = w0.clone().to(dev).requires_grad_()
    v, q, t = heads(w, gb, k)
    part = step(v[lo:hi].contiguous(), q[lo:hi].contiguous(), t[lo:hi].contiguous(), u[lo:hi].to(dev), tgt[lo:hi].to(dev),
                silhouettes=sil[lo:hi].to(dev))["total"] * vd.local_loss_scale(gb, rank, world)
    part.backward()
    ar = vd.GradientAllReduce(10, dev, prefer="nccl")
    ar.buf.copy_(w.grad)
    ar.launch(); ar.join()
    loss_sum = part.detach().clone()
    dist.all_reduce(loss_sum)
    torch.cuda.synchronize()
    if not torch.allclose(loss_sum, full.detach(), rtol=1e-6, atol=1e-8):
        fails.append(f"sharded loss {loss_sum.item()} != full-batch loss {full.item()}")
    if not torch.allclose(ar.buf, g_full, rtol=1e-5, atol=1e-7 * float(g_full.abs().max())):
        fails.append(f"sharded gradient differs: max abs {(ar.buf - g_full).abs().max().item()}")

    # ---- 2. NVLS kernel == NCCL, eager and from a CUDA graph ----------------------------------------------
    numel = 22_875_848
    nv = vd.GradientAllReduce(numel, dev, prefer="nvls")
    nc = vd.GradientAllReduce(numel, dev, prefer="nccl")
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    # integer-valued fp32 data: every partial sum is exact, so the result is independent of the reduction order and the
    # two implementations must agree bit for bit whatever algorithm NCCL picks at this world size
    src = torch.randint(-1000, 1000, (numel,), device=dev, generator=gen).float()
    if nv.graph_capturable:
        nc.buf.copy_(src); nc.launch(inline=True)
        nv.buf.copy_(src); nv.launch(inline=True)
        torch.cuda.synchronize()
        if not torch.equal(nv.buf, nc.buf):
            fails.append(f"NVLS != NCCL (eager): max abs {(nv.buf - nc.buf).abs().max().item()}")
        want = nc.buf.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                nv.launch(inline=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        for it in range(3):
            nv.buf.copy_(src)
            graph.replay()
            torch.cuda.synchronize()
            if not torch.equal(nv.buf, want):
                fails.append(f"NVLS graph replay {it} != NCCL: max abs {(nv.buf - want).abs().max().item()}")
        if nv.nvls_timed_out():
            fails.append("a cross-rank wait inside the NVLS kernel timed out")
        mode = nv.mode
    else:
        mode = f"nvls unavailable ({nv.nvls_error}); NCCL only"
    ok = torch.tensor([0 if fails else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"multigpu_worker: world={world} allreduce={mode} ok={int(ok.item())}", flush=True)
    for f in fails:
        print(f"[rank {rank}] FAIL: {f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
