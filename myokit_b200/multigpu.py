"""
Sharding a simulation over the GPUs of one box.

The reference has no multi-device path at all (no NCCL / MPI /
threads anywhere near the step loop). Here a regular 2-d grid is cut into row
slabs, one per GPU: GPU r owns rows ``[r*ny/N, (r+1)*ny/N)`` of the x-fastest
layout (``cid = ix + iy*nx``, ``myokit/_sim/openclsim.cl:316``), which are
contiguous in every state plane. Per time step each slab needs ONE row of V
from each neighbour. That exchange is not a collective call: the step kernel
of GPU r stores its first / last new V row straight into the neighbour's
ghost buffer (peer stores over NVLink; the buffer is mapped with CUDA IPC
when the neighbour is another process), fences at system scope and bumps an
arrival flag; the neighbour's next step kernel waits on that flag only in the
thread blocks that touch the ghost row, which are scheduled first so the rest
of the slab overlaps the transfer. ``32 KiB`` per neighbour per step at
``nx = 8192`` fp32 is latency-, not bandwidth-bound, so there is no host
round trip and no NCCL launch on the step path. Uncoupled populations
(``diffusion=False``) shard as contiguous cell blocks with no communication.

What the communicator is for: exchanging the 64-byte IPC handles once per
run, a barrier before the first step, and agreeing on NaN halts after every
back-end call (>= 1000 steps). Anything with ``rank``, ``size``,
``allgather(obj) -> list`` and ``barrier()`` works:

``TorchComm``
    one process per GPU under ``torchrun`` (``torch.distributed``, NCCL or
    gloo): the layout ``bench.py`` uses.
``ThreadComm``
    one thread per GPU inside one process (direct peer pointers instead of
    IPC): what the tests use on a multi-GPU box.
"""
import threading

import numpy as np


class TorchComm:
    """
    Communicator over an initialised ``torch.distributed`` process group.

    Host-side agreement (handles, NaN flags, barriers) goes over a ``gloo``
    side group when the default group is NCCL: an ``all_gather_object`` over
    NCCL costs a device synchronisation, two collectives and pickling on
    every call (~0.5 ms each, measured as 1.6 ms of fixed cost per run at 8
    GPUs), and nothing here is device data.
    """

    def __init__(self, group=None, host_group=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError('torch.distributed is not initialised.')
        self._dist = dist
        self._torch = torch
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self._nccl_group = None
        if host_group is None:
            if dist.get_backend(group) == 'gloo':
                host_group = group
            else:
                ranks = (None if group is None
                         else dist.get_process_group_ranks(group))
                host_group = dist.new_group(ranks=ranks, backend='gloo')
                if dist.get_backend(group) == 'nccl' and torch.cuda.is_available():
                    # one-word agreements go over NVLink: ~50 us instead of
                    # the ~0.4 ms of a gloo all-reduce among 8 processes
                    self._nccl_group = (group,)
        self._group = host_group

    def allgather(self, obj):
        out = [None] * self.size
        self._dist.all_gather_object(out, obj, group=self._group)
        return out

    def any(self, flag):
        """True on every rank if ``flag`` is true on any."""
        torch = self._torch
        if self._nccl_group is not None:
            t = torch.tensor([1 if flag else 0], dtype=torch.int32, device='cuda')
            self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX,
                                  group=self._nccl_group[0])
            return bool(t.item())
        t = torch.tensor([1 if flag else 0], dtype=torch.int32)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX, group=self._group)
        return bool(t.item())

    def barrier(self):
        self._dist.barrier(group=self._group)


class ThreadComm:
    """
    Communicators for ``size`` threads of one process::

        comms = ThreadComm.create(2)
        # thread r: SimulationCUDA(..., device=r, comm=comms[r])
    """

    class _Shared:
        def __init__(self, size):
            self.size = size
            self.barrier = threading.Barrier(size)
            self.slots = [None] * size

    def __init__(self, shared, rank):
        self._shared = shared
        self.rank = rank
        self.size = shared.size

    @classmethod
    def create(cls, size):
        shared = cls._Shared(size)
        return [cls(shared, r) for r in range(size)]

    def allgather(self, obj):
        sh = self._shared
        sh.slots[self.rank] = obj
        sh.barrier.wait()
        out = list(sh.slots)
        sh.barrier.wait()
        return out

    def any(self, flag):
        return any(self.allgather(bool(flag)))

    def barrier(self):
        self._shared.barrier.wait()


def slab_rows(ny, size):
    """Row ranges ``[(y0, y1), ...]`` of the ``size`` slabs of ``ny`` rows."""
    return [(r * ny // size, (r + 1) * ny // size) for r in range(size)]


def gather_rows(comm, local, axis=-2):
    """
    Assembles per-rank row slabs (arrays that differ along ``axis``, the y
    axis of ``run_fields`` output) into the global array, on every rank.
    """
    parts = comm.allgather(np.ascontiguousarray(local))
    return np.concatenate(parts, axis=axis)


def run_threads(size, target):
    """
    Runs ``target(comm)`` on ``size`` threads (one per GPU); returns the list
    of results in rank order and re-raises the first exception.
    """
    comms = ThreadComm.create(size)
    results = [None] * size
    errors = [None] * size

    def work(r):
        try:
            results[r] = target(comms[r])
        except BaseException as e:     # noqa
            errors[r] = e
            try:
                comms[r]._shared.barrier.abort()
            except Exception:
                pass
    threads = [threading.Thread(target=work, args=(r,)) for r in range(size)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in errors:
        if e is not None:
            raise e
    return results
