"""Collectives of the env-sharded data-parallel trainer.  Backend-agnostic (NCCL on GPUs, gloo in the CPU tests):
one all-reduce per optimiser step carrying the flat gradient AND the 4-scalar tail (KL / loss / entropy sums) so the
KL-adaptive learning rate stays identical on every rank; one tiny all-reduce per iteration for the advantage statistics."""
import ctypes as C
import os
import warnings

import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_grads_and_tail(flat_grads, tail, comm_buf=None):
    """flat_grads [n] already scaled by 1 / (global mini-batch rows); tail [t] partial sums.  SUM all-reduce, in place.
    Returns the staging buffer (reuse it across calls)."""
    n, t = flat_grads.numel(), tail.numel()
    if comm_buf is None or comm_buf.numel() != n + t:
        comm_buf = torch.empty(n + t, device=flat_grads.device, dtype=flat_grads.dtype)
    comm_buf[:n].copy_(flat_grads)
    comm_buf[n:].copy_(tail)
    dist.all_reduce(comm_buf)
    flat_grads.copy_(comm_buf[:n])
    tail.copy_(comm_buf[n:])
    return comm_buf


def allreduce_adv_stats(stats, local_count):
    """stats = [sum, sum of squares] of the un-normalised advantages (float64).  -> global sample count."""
    dist.all_reduce(stats)
    return local_count * world_size()


# ---- gradient exchange over NVLink peer memory (csrc/dist_kernels.cu) ----------------------------------------------------------------------------
TAIL = 32      # floats in front of the flat gradient: the loss kernel's scalar block (KL / loss / entropy sums, std gradient)
_FLAGS = 32    # floats behind it: 16 uint32 flag words (ready[8], done[8]) + padding
_reducers = {}


class P2PReducer:
    """The flat gradient of one model as a SYMMETRIC buffer [tail | gradient | flags] (torch.distributed._symmetric_memory: same allocation on every
    GPU of the node, mapped into every process) + the library's one-shot all-reduce kernel over it.  `grads` is what the weight-gradient kernels
    write (no staging copy), `tail` what the loss kernel writes; allreduce(off, n) leaves the rank-ordered sums in `out` (same layout), which the
    learning-rate and Adam kernels read.  Everything is a plain kernel launch: the optimiser steps INCLUDING their exchanges replay as one CUDA graph."""

    def __init__(self, n, device):
        import torch.distributed._symmetric_memory as symm
        assert n % 4 == 0
        self.n = n
        group = dist.group.WORLD
        m = TAIL + n
        self.buf = symm.empty(2 * m + _FLAGS, dtype=torch.float32, device=device)      # [tail | gradient] [result, same layout] [flags]
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group.group_name)
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self._data = (C.c_void_p * self.world)(*ptrs)
        self._out = (C.c_void_p * self.world)(*[p + 4 * m for p in ptrs])
        self._flags = (C.c_void_p * self.world)(*[p + 8 * m for p in ptrs])
        self.out = self.buf[m:2 * m]                    # symmetric too: with >= 4 ranks every rank sums one slice and stores it into all results
        self.ctr = torch.zeros(2, dtype=torch.int32, device=device)
        self.tail, self.grads = self.buf[:TAIL], self.buf[TAIL:m]
        self.out_tail, self.out_grads = self.out[:TAIL], self.out[TAIL:m]
        torch.cuda.synchronize(device)
        dist.barrier()          # every rank's flag words are zero before any peer can write them

    def allreduce(self, off, n):
        """out[off : off + n] = sum over ranks of buf[off : off + n] (float offsets into [tail | gradient]); collective, capturable."""
        from . import _ops
        _ops.call("go2_allreduce_p2p2", self._data, self._flags, self._out, self.out.data_ptr(), off, n, self.rank, self.world, self.ctr.data_ptr())


def new_flat_grad(n, device):
    """The flat gradient vector of a model (n floats, n % 4 == 0).  One process: a plain zero tensor.  Env-sharded over the GPUs of a node (NCCL
    backend): the gradient region of a P2PReducer's symmetric buffer; the algorithm picks the reducer up with reducer_for().  GO2_DIST_P2P=0 or a
    failed symmetric allocation (no peer access between the GPUs) keeps the NCCL all-reduce of allreduce_grads_and_tail()."""
    dev = torch.device(device)
    if world_size() > 1 and dev.type == "cuda" and dist.get_backend() == "nccl" and os.environ.get("GO2_DIST_P2P", "1") != "0":
        try:
            red = P2PReducer(n, dev)
            _reducers[red.grads.data_ptr()] = red
            return red.grads
        except Exception as e:      # noqa: BLE001 - anything the symmetric allocator / rendezvous rejects
            warnings.warn(f"symmetric-memory gradient exchange unavailable ({type(e).__name__}: {e}); using NCCL all-reduce")
    return torch.zeros(n, device=device)


def reducer_for(flat_grads):
    return _reducers.get(flat_grads.data_ptr())


class SmallSum:
    """SUM of a short vector over the ranks in the middle of a backward pass (the mixture layers' mean gate usage): the NVLink kernel over its own
    tiny symmetric buffer when the gradient exchange uses it too (capturable), a plain dist.all_reduce with gloo (CPU tests).  `inp` is written by
    the caller, reduce() returns the tensor that holds the sums.  make_small_sum() -> None when there is nothing to sum (one process) or when the
    only collective available is NCCL between graph segments (each rank then keeps its own shard's usage, as before)."""

    def __init__(self, n, device, red):
        self.n, self.red = n, red
        self.inp = red.grads[:n] if red is not None else torch.zeros(n, device=device)

    def reduce(self):
        if self.red is not None:
            self.red.allreduce(TAIL, (self.n + 3) // 4 * 4)
            return self.red.out_grads[:self.n]
        dist.all_reduce(self.inp)
        return self.inp


def make_small_sum(n, device):
    if world_size() == 1:
        return None
    dev = torch.device(device)
    if dev.type == "cuda" and dist.get_backend() == "nccl":
        if os.environ.get("GO2_DIST_P2P", "1") == "0":
            return None
        try:
            return SmallSum(n, dev, P2PReducer((n + 3) // 4 * 4, dev))
        except Exception as e:      # noqa: BLE001
            warnings.warn(f"symmetric-memory exchange unavailable for the gate usage ({type(e).__name__}: {e}); per-rank usage")
            return None
    return SmallSum(n, dev, None)
