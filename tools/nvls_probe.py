"""Development probe (not on the product path): does torch's symmetric memory give this process group a multicast (NVLS)
address on this box?  torchrun --nproc-per-node N tools/nvls_probe.py"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def _mc(h):
    for dt in (getattr(getattr(torch._C, "_autograd", None), "DeviceType", None), getattr(torch._C, "DeviceType", None)):
        try:
            return bool(h.has_multicast_support(dt.CUDA, local))
        except Exception:
            continue
    return None


t = symm.empty(1 << 20, dtype=torch.uint8, device=f"cuda:{local}")
h = symm.rendezvous(t, dist.group.WORLD.group_name)
info = dict(rank=rank, world=h.world_size, backend=symm.get_backend(torch.device("cuda", local)) if hasattr(symm, "get_backend") else None,
            multicast=_mc(h),
            multicast_ptr=hex(h.multicast_ptr) if getattr(h, "multicast_ptr", 0) else None,
            buffer_ptrs=[hex(p) for p in h.buffer_ptrs], signal_pad=[hex(p) for p in h.signal_pad_ptrs],
            signal_pad_size=h.signal_pad_size)
print(info, flush=True)
dist.barrier()
dist.destroy_process_group()
