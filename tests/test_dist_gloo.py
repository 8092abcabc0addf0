"""Data-parallel host logic on CPU: world_size 2, gloo, 127.0.0.1.  The per-sample arithmetic is the
oracle's (no GPU here); what is tested is the sharding contract of vpn_b200.dist: per-rank shard ranges,
the local-loss scale, and that the SUM all-reduce of the shard gradients equals the single-process gradient
(1-vs-G parity)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _model_outputs(w, b, k):
    """A stand-in for the network heads: parameters w -> (v, q, t) of shape (B,K,.)."""
    z = torch.linspace(-1, 1, b * k * 10).view(b, k, 10)
    h = z * w[None, None, :] + w.flip(0)[None, None, :] * 0.1
    v = (torch.sigmoid(h[..., 0:3]) + 0.1) / torch.tensor([8.0, 10.0, 10.0])
    return v, torch.sigmoid(h[..., 3:7]), torch.tanh(h[..., 7:10]) * 0.3


def _loss(v, q, t, u, tgt):
    from oracle import vpn_oracle as O
    pts = O.sample_predict_points("sphere", v, q, t, u)
    return O.chamfer_dense(pts, tgt) + 0.1 * O.chamfer_dense(t, tgt, w1=0.5, w2=1.0)


def _worker(rank, world, port, gb, k, n, m, out):
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (repo, os.path.join(repo, "volumetric-primitives-net_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from vpn_b200 import dist as vd
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    w = torch.randn(10, requires_grad=True)
    g = torch.Generator().manual_seed(3)
    u = torch.rand(gb, k, n, 2, generator=g)            # drawn globally, then sliced: shard-invariant randomness
    tgt = torch.rand(gb, m, 3, generator=g) - 0.5
    lo, hi = vd.shard_range(gb, rank, world)
    v, q, t = _model_outputs(w, gb, k)
    loss = _loss(v[lo:hi], q[lo:hi], t[lo:hi], u[lo:hi], tgt[lo:hi]) * vd.local_loss_scale(gb, rank, world)
    loss.backward()
    sync = vd.GradientAllReduce(10, "cpu")
    sync.buf.copy_(w.grad)
    sync.launch(inline=(gb % 2 == 1)); sync.join()          # both launch modes (dedicated stream / caller's stream)
    tot = loss.detach().clone()
    dist.all_reduce(tot)
    if rank == 0:
        torch.save({"grad": sync.buf.clone(), "loss": tot, "range": (lo, hi)}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("gb", [4, 5])
def test_two_rank_gradients_match_single_process(tmp_path, gb):
    from vpn_b200 import dist as vd
    k, n, m = 2, 16, 24
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), gb, k, n, m, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    w = torch.randn(10, requires_grad=True)
    g = torch.Generator().manual_seed(3)
    u = torch.rand(gb, k, n, 2, generator=g)
    tgt = torch.rand(gb, m, 3, generator=g) - 0.5
    v, q, t = _model_outputs(w, gb, k)
    loss = _loss(v, q, t, u, tgt)
    loss.backward()
    torch.testing.assert_close(got["loss"], loss.detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(got["grad"], w.grad, rtol=1e-4, atol=1e-7)
    assert got["range"] == vd.shard_range(gb, 0, 2)


def test_shard_ranges_partition_the_batch():
    from vpn_b200 import dist as vd
    for gb in (1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            r = [vd.shard_range(gb, i, world) for i in range(world)]
            assert r[0][0] == 0 and r[-1][1] == gb and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert abs(sum(vd.local_loss_scale(gb, i, world) for i in range(world)) - 1.0) < 1e-12
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
