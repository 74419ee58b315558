"""Gradient exchange of the data-parallel PPO update over peer-mapped memory.

The flat gradient buffer of every rank is allocated in symmetric memory (`torch.distributed._symmetric_memory`: the same
allocation mapped into every process of the node over NVLink / NVSwitch).  `reduce()` is barrier -> ONE kernel
(`b200gym_grad_reduce_peers`: every rank sums all ranks' buffers itself, in rank order, and accumulates the squared gradient norm
in the same pass) -> barrier.  Against `dist.all_reduce` + `b200gym_grad_sumsq` this saves a collective launch per minibatch,
makes the sums bit-identical on all ranks by construction, and — being plain kernels — lets the whole minibatch step be
captured in a CUDA graph across ranks (the NCCL path runs eagerly).  torch owns the memory and the rendezvous: plumbing; the
reduction itself is ours.
"""
import torch

from . import _lib


class PeerGradReducer:
    def __init__(self, numel, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > 16:
            raise RuntimeError("b200gym_grad_reduce_peers takes at most 16 ranks")
        self.numel = int(numel)
        self.buf = symm.empty(self.numel, dtype=torch.float32, device=device)      # this rank's flat gradient buffer
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group=dist.group.WORLD.group_name)
        self.peers = _lib.PeerPtrsPOD()
        for r in range(self.world):
            self.peers.ptr[r] = int(self.hdl.buffer_ptrs[r])
        self.out = torch.zeros(self.numel, dtype=torch.float32, device=device)       # the reduced gradients (+ piggy-backed tail)
        self.lib = _lib.lib()
        self.device = torch.device(device)

    def reduce(self, n_params, sumsq):
        """out[:] = sum over ranks of their buffers; sumsq += |out[:n_params]|^2.  Stream-ordered, graph-capturable."""
        self.hdl.barrier()
        _lib.check(self.lib.b200gym_grad_reduce_peers(self.peers, self.world, _lib.ptr(self.out), self.numel, int(n_params),
                                                      _lib.ptr(sumsq), _lib.stream_ptr(self.device)), "grad_reduce_peers")
        self.hdl.barrier()
        return self.out
