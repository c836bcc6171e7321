"""Collectives of the env-sharded data-parallel trainer.  Backend-agnostic (NCCL on GPUs, gloo in the CPU tests):
one all-reduce per optimiser step carrying the flat gradient AND the 4-scalar tail (KL / loss / entropy sums) so the
KL-adaptive learning rate stays identical on every rank; one tiny all-reduce per iteration for the advantage statistics."""
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
