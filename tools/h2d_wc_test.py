"""Raw host->device rate of N ranks at once with different kinds of pinned host memory (development aid for the e2e-at-N>1
question: profiles/r2_topology_8gpu.txt).  torchrun --nproc-per-node N tools/h2d_wc_test.py"""
import ctypes, os, time
import torch
import torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
NB, SZ = 16, 3840 * 2 * 2160 * 3 // 2


def alloc(kind):
    bufs = []
    for _ in range(NB):
        if kind == "torch":
            bufs.append(torch.empty(SZ, dtype=torch.uint8).pin_memory())
        else:
            p = ctypes.c_void_p()
            flags = {"default": 0, "wc": 4, "portable": 1, "wc+portable": 5}[kind]
            assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(SZ), ctypes.c_uint(flags)) == 0
            t = torch.frombuffer((ctypes.c_uint8 * SZ).from_address(p.value), dtype=torch.uint8)
            t.fill_(7)
            bufs.append(t)
    return bufs


stage = [torch.empty(SZ, dtype=torch.uint8, device="cuda") for _ in range(4)]
for kind in ("torch", "default", "wc", "wc+portable"):
    bufs = alloc(kind)
    def run(n):
        for i in range(n):
            stage[i % 4].copy_(bufs[i % NB], non_blocking=True)
    run(8)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record(); run(n); e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{kind:12s}: {world} ranks, {world * n * SZ / (ms.item() / 1e3) / 1e9:7.1f} GB/s aggregate ({n * SZ / (ms.item() / 1e3) / 1e9:5.1f} per GPU), pinned seen by torch: {bufs[0].is_pinned()}", flush=True)
    del bufs
