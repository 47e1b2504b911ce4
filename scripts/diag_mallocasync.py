"""How long does a 4-byte cudaMallocAsync take right after a synchronisation (default pool: release threshold 0)
and once the pool is told to keep its memory?  Background of csrc/api.cu::scratch_flag."""
import time
import torch
from cuda.bindings import runtime as rt

torch.zeros(1, device="cuda")
big = torch.empty(int(8e9), dtype=torch.uint8, device="cuda")      # some mapped memory, like a real job
err, stream = rt.cudaStreamCreate()

def cycle(n=12):
    out = []
    for _ in range(n):
        rt.cudaDeviceSynchronize()
        t0 = time.perf_counter()
        err, ptr = rt.cudaMallocAsync(4, stream)
        rt.cudaMemsetAsync(ptr, 0, 4, stream)
        rt.cudaFreeAsync(ptr, stream)
        rt.cudaStreamSynchronize(stream)
        out.append(round(1e3 * (time.perf_counter() - t0), 3))
    return out

print("default pool (threshold 0), ms per malloc+memset+free+sync:", cycle())
err, pool = rt.cudaDeviceGetDefaultMemPool(0)
from cuda.bindings import driver as drv
val = drv.cuuint64_t(2**64 - 1)
print("set threshold:", rt.cudaMemPoolSetAttribute(pool, rt.cudaMemPoolAttr.cudaMemPoolAttrReleaseThreshold, val))
print("pool keeps its memory, ms:", cycle())
