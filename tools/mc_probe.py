"""2-GPU+ probe (torchrun): does torch symmetric memory expose an NVLS multicast alias here?
Prints the multicast pointer and the per-rank buffer addresses."""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank=int(os.environ["RANK"]); torch.cuda.set_device(rank); dev=torch.device("cuda",rank)
dist.init_process_group("nccl", device_id=dev)
buf=symm.empty(1<<20, dtype=torch.uint8, device=dev)
hdl=symm.rendezvous(buf, dist.group.WORLD.group_name)
print(rank, "mc_ptr", hex(hdl.multicast_ptr), "ptrs", [hex(p) for p in hdl.buffer_ptrs], "backend", symm.get_backend(dev) if hasattr(symm,'get_backend') else None, flush=True)
try:
    print(rank, "has_multicast_support", type(hdl).has_multicast_support(torch.device('cuda').type if False else dev.type, rank), flush=True)
except Exception as e:
    print("hms err", e)
dist.barrier(); os._exit(0)
