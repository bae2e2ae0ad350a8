"""Probe: torch symmetric memory (peer pointers over NVLink) on this box.  torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem

print(rank, "symm_mem api:", [n for n in dir(symm_mem) if not n.startswith("_")][:40], flush=True)
t = symm_mem.empty(1024, dtype=torch.float64, device=dev)
t.fill_(float(rank))
try:
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
except Exception as e:
    print(rank, "rendezvous(group_name) failed:", repr(e), flush=True)
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, "handle attrs:", [n for n in dir(hdl) if not n.startswith("_")], flush=True)
print(rank, "buffer_ptrs:", [hex(p) for p in hdl.buffer_ptrs], "rank", hdl.rank, "world", hdl.world_size, flush=True)
hdl.barrier(channel=0)
peer = (rank + 1) % world
buf = hdl.get_buffer(peer, (1024,), torch.float64)
buf[rank * 8:(rank + 1) * 8] = 100.0 + rank   # store into the PEER's buffer
hdl.barrier(channel=0)
torch.cuda.synchronize()
print(rank, "my buffer after peers wrote:", t[:24].tolist(), flush=True)
dist.destroy_process_group()
