"""Probe: does torch symmetric memory (peer-mapped buffers over NVLink) work between the ranks of this box?  torchrun --nproc-per-node 2"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
t = symm.empty(1024, dtype=torch.float32, device="cuda")
t.fill_(float(rank + 1))
hdl = symm.rendezvous(t, group=dist.group.WORLD.group_name)
print(rank, "rendezvous ok", [hex(p) for p in hdl.buffer_ptrs], hdl.rank, hdl.world_size, flush=True)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
print(rank, "peer value", float(peer[0]), flush=True)
g = torch.cuda.CUDAGraph()
out = torch.zeros(1024, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    hdl.barrier()
torch.cuda.synchronize()
try:
    with torch.cuda.graph(g):
        hdl.barrier()
        out.copy_(peer)
        hdl.barrier()
    g.replay(); g.replay()
    torch.cuda.synchronize()
    print(rank, "graph-captured barrier + peer read ok", float(out[0]), flush=True)
except Exception as e:
    print(rank, "graph capture failed:", type(e).__name__, e, flush=True)
dist.destroy_process_group()
