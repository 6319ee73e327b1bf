"""Tensor-parallel group: maps every rank's exchange buffer into every other rank with CUDA IPC and builds the
native NVLink communicator (b200_comm).  One process per GPU; torch.distributed is used ONLY to exchange the
64-byte IPC handles at start-up -- no NCCL call is on the decode path.
Stands where DeviceCommImpl allocation/registration stands in the reference
(src/turbomind/comm/cuda_ipc/cuda_ipc_comm.cu:165-291)."""
import ctypes as C

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
        self.local = []
        handles = []
        for size in (nbytes, fbytes):
            ptr = C.c_void_p()
            h = C.create_string_buffer(64)
            L.check(self.lib.b200_ipc_alloc(size, C.byref(ptr), h), "b200_ipc_alloc")
            self.local.append(ptr)
            handles.append(bytes(h.raw))
        gathered = [None] * world
        dist.all_gather_object(gathered, handles, group=group)
        self.opened = []
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

    def close(self):
        torch.cuda.synchronize()
        if self.handle:
            self.lib.b200_comm_destroy(self.handle)
            self.handle = None
        for p in self.opened:
            self.lib.b200_ipc_close(p)
        self.opened = []
