"""Host -> device copy rate from default pinned memory vs write-combined pinned memory (cudaHostAllocWriteCombined),
alone and while kernels run (diagnostic for the end-to-end arm, which is upload-bound)."""
import time

import torch
from cuda.bindings import runtime as rt

n = 384 * 1024 * 1024
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
res = {}
for name, flags in (("default", rt.cudaHostAllocDefault), ("write-combined", rt.cudaHostAllocWriteCombined),
                    ("mapped", rt.cudaHostAllocMapped), ("wc+mapped", rt.cudaHostAllocWriteCombined | rt.cudaHostAllocMapped)):
    err, ptr = rt.cudaHostAlloc(n, flags)
    assert int(err) == 0, err
    rt.cudaMemset(dev.data_ptr(), 0, n)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        rt.cudaMemcpyAsync(dev.data_ptr(), ptr, n, rt.cudaMemcpyKind.cudaMemcpyHostToDevice, st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        rt.cudaMemcpyAsync(dev.data_ptr(), ptr, n, rt.cudaMemcpyKind.cudaMemcpyHostToDevice, st)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:15s} H2D {20 * n / dt / 1e9:6.1f} GB/s", flush=True)
    rt.cudaFreeHost(ptr)
