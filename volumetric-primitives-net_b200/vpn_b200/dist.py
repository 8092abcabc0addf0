"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The hot path shards by batch with no data-path collective: every sample's
sampling, Chamfer, mesh and silhouette is independent and the only cross-sample operation is the final
mean over B (chamfer_distance.py:30; L1Loss mean, silhouette.py:11,22).  The one collective per step is
the all-reduce of the network-parameter gradients, which the reference does not have (it is single
process); it is issued on a side stream so it overlaps whatever the caller runs next.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Rank r of G owns samples [lo, hi); shards differ by at most one sample."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_loss_scale(global_batch: int, rank: int, world: int) -> float:
    """Factor that turns a rank's local batch-mean into its share of the global batch-mean, so that the
    SUM all-reduce of the gradients equals the single-process gradient."""
    lo, hi = shard_range(global_batch, rank, world)
    return (hi - lo) / float(global_batch)


class GradientAllReduce:
    """Sum all-reduce of one flat gradient buffer, on the current stream or on a dedicated one, joined on demand.

    mode "nvls": the buffer is symmetric memory with a multicast mapping and the reduction is ONE launch of
    vpn_allreduce_nvls_sync (csrc/allreduce.cu: multimem.ld_reduce + multimem.st through the NVSwitch, with the two
    cross-rank barriers inside the kernel and the epoch in device memory, so the launch can be captured in a CUDA
    graph).  Every rank must take the same path: the ranks agree with a MIN all-reduce FIRST that the symmetric-memory
    set-up worked everywhere, then run a trial (the kernel's waits are bounded, a missing peer cannot hang the GPU) and
    agree again on its result; otherwise NCCL's all-reduce is used (mode "nccl").  With "auto" both are timed once at
    construction (max over ranks) and the faster one is kept.  VPN_ALLREDUCE=nccl|nvls|auto overrides."""

    def __init__(self, numel: int, device, dtype=torch.float32, prefer: Optional[str] = None):
        prefer = prefer or os.environ.get("VPN_ALLREDUCE", "auto")
        self.cuda = torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.done: Optional[torch.cuda.Event] = None
        self.mode, self.nvls_error, self._full, self._handle = "none", None, None, None
        self._prefer, self.trial_ms = prefer, None
        self._padded = 0
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.buf = None
        if distributed and self.cuda and dtype == torch.float32 and prefer in ("auto", "nvls"):
            self._try_nvls(numel, device)
        if self.buf is None:
            self.buf = torch.zeros(numel, dtype=dtype, device=device)
            self.mode = "nccl all_reduce" if distributed and self.cuda else ("gloo all_reduce" if distributed else "none")

    @property
    def graph_capturable(self) -> bool:
        """True when launch(inline=True) is a single self-synchronising kernel of ours (no NCCL call, no host state)."""
        return self._handle is not None

    def _nvls_launch(self):
        from . import _lib
        h = self._handle
        _lib.check(_lib.load().vpn_allreduce_nvls_sync(h.multicast_ptr, self._full.data_ptr(), self._padded, h.rank,
                                                       h.world_size, _lib.stream_ptr(self._full.device)), "vpn_allreduce_nvls_sync")

    def nvls_timed_out(self) -> bool:
        """True if a cross-rank wait inside the kernel ever timed out on this rank (synchronises; call outside timed regions)."""
        if self._handle is None:
            return False
        from . import _lib
        word = self._padded + _lib.load().vpn_allreduce_nvls_error_word()
        return bool(self._full[word:word + 1].view(torch.int32).item() != 0)

    def _agree(self, ok: bool, device) -> bool:
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return int(flag.item()) == 1

    def _try_nvls(self, numel, device):
        import ctypes
        from . import _lib
        ok = False
        try:
            import torch.distributed._symmetric_memory as symm_mem
            nflag = ctypes.c_size_t(0)
            _lib.check(_lib.load().vpn_allreduce_nvls_flag_floats(ctypes.byref(nflag)), "vpn_allreduce_nvls_flag_floats")
            self._padded = -(-numel // 4096) * 4096           # 16-byte aligned per-rank slices for any world size
            self._full = symm_mem.empty(self._padded + nflag.value, dtype=torch.float32, device=device)
            self._handle = symm_mem.rendezvous(self._full, dist.group.WORLD.group_name)
            if not getattr(self._handle, "multicast_ptr", 0):
                raise RuntimeError("no multicast (NVLS) mapping")
            if int(getattr(self._handle, "offset", 0)) != 0 or int(self._handle.buffer_ptrs[self._handle.rank]) != self._full.data_ptr():
                raise RuntimeError("symmetric buffer does not start at its allocation: multicast address unknown")
            self._full.zero_()                                # payload and flag words
            torch.cuda.synchronize(device)
            ok = True
        except Exception as e:                                # noqa: BLE001 - any failure means "use NCCL"
            self.nvls_error = repr(e)[:300]
        # agreement BEFORE any cross-rank kernel: a rank whose set-up failed must not leave the others waiting for it;
        # the all-reduce is also the barrier that makes every rank's zeroed flag words visible before the first launch
        if not self._agree(ok, device):
            self._full, self._handle = None, None
            return
        ok = False
        try:
            # trial: sum of (rank + 1) over the ranks
            self._full[:self._padded].fill_(float(self._handle.rank + 1))
            self._nvls_launch()
            torch.cuda.synchronize(device)
            w = self._handle.world_size
            if self.nvls_timed_out():
                raise RuntimeError("NVLS trial all-reduce: cross-rank wait timed out")
            if not bool((self._full[:self._padded] == float(w * (w + 1) // 2)).all()):
                raise RuntimeError("NVLS trial all-reduce gave a wrong sum")
            ok = True
        except Exception as e:                                # noqa: BLE001
            self.nvls_error = repr(e)[:300]
        if not self._agree(ok, device):
            self._full, self._handle = None, None
            return
        if self._prefer == "auto":
            # time both paths (max over ranks) and keep the faster one
            scratch = torch.zeros(numel, dtype=torch.float32, device=device)
            times = []
            for fn in (self._nvls_launch, lambda: dist.all_reduce(scratch, op=dist.ReduceOp.SUM)):
                for _ in range(2):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    fn()
                e1.record()
                torch.cuda.synchronize(device)
                times.append(e0.elapsed_time(e1) / 5)
            t = torch.tensor(times, dtype=torch.float32, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            self.trial_ms = {"nvls": float(t[0].item()), "nccl": float(t[1].item())}
            if self.trial_ms["nccl"] <= self.trial_ms["nvls"]:
                self._full, self._handle = None, None
                return
        self._full[:self._padded].zero_()
        self.buf = self._full[:numel]
        self.mode = "nvls multimem kernel, barriers in-kernel (vpn_allreduce_nvls_sync)"

    def launch(self, inline: bool = False):
        """Start the all-reduce of `buf`.  Default: on the dedicated stream, after everything already queued on the
        current stream, so that later work on the current stream overlaps it until join().  inline=True queues it on the
        current stream itself - for callers that would join() immediately anyway: no cross-stream event waits at all."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        if not self.cuda:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
            return
        if inline:
            self.done = None
            if self._handle is not None:
                self._nvls_launch()
            else:
                dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
            return
        self.stream.wait_stream(torch.cuda.current_stream(self.buf.device))
        with torch.cuda.stream(self.stream):
            if self._handle is not None:
                self._nvls_launch()
            else:
                dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
            self.done = torch.cuda.Event()
            self.done.record(self.stream)

    def join(self):
        if self.done is not None:
            torch.cuda.current_stream(self.buf.device).wait_event(self.done)
            self.done = None
