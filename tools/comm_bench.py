"""Development tool (measurement only; torchrun, one rank per GPU): CUDA-graph time of the fused all-reduce + residual +
RMSNorm at the decode shapes (tokens x hidden partial per rank), alternating the two halves of the exchange buffer as a
layer does, with CUDA-IPC peer loads or -- B200_TP_NVLS=1 -- NVLS.  Prints the max over ranks."""
import ctypes as C
import os
import sys
import types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from lmdeploy_b200 import _lib as L
from lmdeploy_b200.turbomind.tp import TpGroup

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
lib = L.load()


class RawBuf:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


SHAPES = ((64, 4096), (32, 8192)) if "--short" in sys.argv else ((64, 4096), (64, 6144), (32, 8192), (8, 4096))
for tokens, hidden in SHAPES:
    model = types.SimpleNamespace(spec=types.SimpleNamespace(hidden=hidden), max_batch=64, vocab_local=92544 // world,
                                  device=torch.device("cuda", torch.cuda.current_device()), set_comm=lambda h: None)
    tp = TpGroup(model, rank, world)
    L.check(lib.b200_comm_set_geometry(tp.handle, 64, hidden), "geometry")
    halves = []
    for h in (0, 1):
        ptr = lib.b200_comm_local_buffer(tp.handle, h)
        t = torch.as_tensor(RawBuf(ptr, tokens * hidden * 2), device="cuda").view(torch.float16).view(tokens, hidden)
        t.copy_((torch.randn(tokens, hidden, device="cuda") * 0.1).half())
        halves.append((ptr, t))
    out = torch.zeros(tokens, hidden, dtype=torch.float16, device="cuda")
    res = torch.zeros(tokens, hidden, dtype=torch.float16, device="cuda")
    w = torch.ones(hidden, dtype=torch.float16, device="cuda")
    st = torch.cuda.Stream()
    calls = 64

    def body():
        for i in range(calls):
            L.check(lib.b200_comm_allreduce_residual_rmsnorm(tp.handle, out.data_ptr(), 0, res.data_ptr(), halves[i & 1][0],
                                                             w.data_ptr(), None, tokens, hidden, 1e-5, 0, st.cuda_stream), "ar")
    torch.cuda.synchronize()
    dist.barrier()
    with torch.cuda.stream(st):
        body()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            body()
        for _ in range(3):
            g.replay()
        st.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record(st)
        for _ in range(reps):
            g.replay()
        e1.record(st)
        st.synchronize()
    us = torch.tensor([e0.elapsed_time(e1) * 1e3 / (reps * calls)], device="cuda")
    dist.all_reduce(us, op=dist.ReduceOp.MAX)
    # single-GPU yardstick: the plain residual RMSNorm kernel on the same rows
    if rank == 0:
        print(f"world {world} {'NVLS' if tp.nvls else 'IPC peer loads'} tokens {tokens} hidden {hidden}: "
              f"{us.item():.2f} us per all-reduce+residual+RMSNorm (graph of {calls}, max over ranks); "
              f"payload {tokens * hidden * 2 / 1024:.0f} KiB per rank", flush=True)
    dist.barrier()
    tp.close()
dist.destroy_process_group()
