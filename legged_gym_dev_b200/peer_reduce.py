"""Gradient exchange of the data-parallel PPO update over NVLink peer memory.

`PeerExchange` (default): every rank owns a symmetric-memory buffer (`torch.distributed._symmetric_memory`: one allocation
per rank, mapped into every process of the node over NVLink / NVSwitch) holding 2 x world gradient slots + world flags.
The optimiser kernel of each rank (`b200gym_ppo_optimizer_step_peers`, csrc/ppo_train.cu) pushes its gradients into every
rank's slot, signals with a system-scope flag, waits for the peers' flags, sums the slots in rank order and performs the
clip + Adam step — ONE launch per rank and minibatch, no host barrier, no torch barrier kernel, no NCCL call; bit-identical
parameters on all ranks by construction.  torch owns the memory and the rendezvous (plumbing); the exchange is ours.

`PeerGradReducer` is the round-1 form (barrier -> reduce kernel -> barrier, then separate clip/Adam launches), kept as the
A/B baseline of tools/ddp_train_check.py.
"""
import torch

from . import _lib


class PeerExchange:
    def __init__(self, n_params, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > 16:
            raise RuntimeError("b200gym_ppo_optimizer_step_peers takes at most 16 ranks")
        self.device = torch.device(device)
        self.n = int(n_params)
        self.n_pad = (self.n + 2 + 3) // 4 * 4
        floats = 2 * self.world * self.n_pad + 32              # slots [parity][source rank][n_pad] + flags (uint32 per rank)
        self.buf = symm.empty(floats, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group=dist.group.WORLD.group_name)
        torch.cuda.synchronize(self.device)
        dist.barrier()                                           # nobody pushes before every rank's buffer is zeroed
        self.peers = _lib.PeerBasesPOD()
        for r in range(self.world):
            self.peers.base[r] = int(self.hdl.buffer_ptrs[r])
        self.grad_sum = torch.zeros(self.n_pad, dtype=torch.float32, device=device)
        self.ws = torch.zeros(_lib.PEER_WS_BYTES // 4, dtype=torch.int32, device=device)

    def error(self):
        """True if a launch gave up waiting for a peer (host sync)."""
        return bool(self.ws[4].item())


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
