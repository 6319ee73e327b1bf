"""Tensor-parallel group: maps every rank's exchange buffer into every other rank with CUDA IPC and builds the
native NVLink communicator (b200_comm).  One process per GPU; torch.distributed is used ONLY to exchange the
64-byte IPC handles at start-up -- no NCCL call is on the decode path.
Stands where DeviceCommImpl allocation/registration stands in the reference
(src/turbomind/comm/cuda_ipc/cuda_ipc_comm.cu:165-291)."""
import ctypes as C
import os

import torch
import torch.distributed as dist

from .. import _lib as L


class TpGroup:
    def __init__(self, model, rank, world, group=None):
        self.lib = L.load()
        self.rank, self.world = rank, world
        spec = model.spec
        nbytes = self.lib.b200_comm_buffer_bytes(model.max_batch, spec.hidden, model.vocab_local)
        fbytes = self.lib.b200_comm_flags_bytes()
        self.local, self.opened, self.symm = [], [], None
        self.nvls = False
        # NVLS (B200_TP_NVLS=1, or =auto: used when the fabric offers it): the exchange buffers come from torch's symmetric
        # memory, whose rendezvous also yields a MULTICAST address; the fused all-reduce then reduces inside the NVSwitch.
        # Default: "auto" from 4 ranks up (measured on B200s, profiles/r02_scale.md: C2 at TP 4 2.70 ms with NVLS vs 2.75 ms with
        # peer loads, the fused exchange alone 8.0 vs 8.5 us; at TP 2 peer loads win, 6.9 vs 7.6 us).
        want = os.environ.get("B200_TP_NVLS", "auto" if world >= 4 else "0")
        if want not in ("", "0") and world > 1:
            try:
                self._init_symmetric(model, rank, world, nbytes, fbytes, group)
                return
            except Exception as e:  # noqa: BLE001 -- any failure of the optional path falls back to CUDA IPC peer loads
                if want != "auto":
                    raise
                self.symm = None
                import warnings
                warnings.warn(f"NVLS exchange unavailable ({e}); using CUDA-IPC peer loads")
        handles = []
        for size in (nbytes, fbytes):
            ptr = C.c_void_p()
            h = C.create_string_buffer(64)
            L.check(self.lib.b200_ipc_alloc(size, C.byref(ptr), h), "b200_ipc_alloc")
            self.local.append(ptr)
            handles.append(bytes(h.raw))
        gathered = [None] * world
        dist.all_gather_object(gathered, handles, group=group)
        data = (C.c_void_p * world)()
        flags = (C.c_void_p * world)()
        for r in range(world):
            if r == rank:
                data[r], flags[r] = self.local[0], self.local[1]
                continue
            ptrs = []
            for h in gathered[r]:
                p = C.c_void_p()
                L.check(self.lib.b200_ipc_open(h, C.byref(p)), "b200_ipc_open")
                ptrs.append(p)
                self.opened.append(p)
            data[r], flags[r] = ptrs
        self.handle = C.c_void_p()
        L.check(self.lib.b200_comm_create(C.byref(self.handle), rank, world, data, flags), "b200_comm_create")
        model.set_comm(self.handle)
        dist.barrier(group=group)

    def _init_symmetric(self, model, rank, world, nbytes, fbytes, group):
        import torch.distributed._symmetric_memory as symm
        dev = model.device
        total = (nbytes + 4095) // 4096 * 4096 + fbytes
        gname = (group or dist.group.WORLD).group_name
        buf = symm.empty(total, dtype=torch.uint8, device=dev)
        buf.zero_()
        torch.cuda.synchronize()
        h = symm.rendezvous(buf, gname)
        mc = int(getattr(h, "multicast_ptr", 0) or 0)
        if not mc:
            raise RuntimeError("symmetric memory rendezvous returned no multicast address")
        foff = (nbytes + 4095) // 4096 * 4096
        data = (C.c_void_p * world)(*[int(p) for p in h.buffer_ptrs])
        flags = (C.c_void_p * world)(*[int(p) + foff for p in h.buffer_ptrs])
        self.handle = C.c_void_p()
        L.check(self.lib.b200_comm_create(C.byref(self.handle), rank, world, data, flags), "b200_comm_create")
        L.check(self.lib.b200_comm_set_multicast(self.handle, C.c_void_p(mc)), "b200_comm_set_multicast")
        self.symm, self.nvls = (buf, h), True
        model.set_comm(self.handle)
        dist.barrier(group=group)

    def close(self):
        torch.cuda.synchronize()
        if self.handle:
            self.lib.b200_comm_destroy(self.handle)
            self.handle = None
        for p in self.opened:
            self.lib.b200_ipc_close(p)
        self.opened = []
