"""Can the copy engine skip the hydrogens?  ("heavy-atoms-first upload")

MDAnalysis hands positions as float32 [N][3]; a water is O, H, H = 36 bytes of which the neighbour search
needs the first 12.  This probe times, from pinned host memory to the device,
  (a) the contiguous copy of W waters (36 W bytes)                       cudaMemcpyAsync
  (b) the strided copy of the oxygens only (width 12, pitch 36, W rows)  cudaMemcpy2DAsync
and prints waters per second for both.  (b) only pays if it moves the rows faster than (a) moves everything.
"""
import ctypes
import time

import torch


def cudart():
    with open("/proc/self/maps") as fh:
        for line in fh:
            if "libcudart" in line:
                return ctypes.CDLL(line.split()[-1])
    return ctypes.CDLL("libcudart.so.12")


def main():
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    rt = cudart()
    rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                     ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    W = 3_200_000  # waters (32 C4 frames)
    host = torch.empty(W * 9, dtype=torch.float32).pin_memory()
    host.normal_()
    dev = torch.empty(W * 9, dtype=torch.float32, device="cuda")
    H2D = 1

    def timed(fn, reps=10):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) / reps

    full = timed(lambda: rt.cudaMemcpyAsync(dev.data_ptr(), host.data_ptr(), W * 36, H2D, None))
    print(f"contiguous  36 B/water : {full * 1e3:8.3f} ms  {W * 36 / full / 1e9:6.1f} GB/s  {W / full / 1e9:6.3f} G waters/s", flush=True)
    for width, pitch, label in ((12, 36, "oxygens only, packed"), (24, 36, "hydrogens only, packed"), (36, 36, "2D call, full rows")):
        t = timed(lambda: rt.cudaMemcpy2DAsync(dev.data_ptr(), width, host.data_ptr(), pitch, width, W, H2D, None), reps=3)
        print(f"2D width {width:2d} pitch {pitch:2d} ({label}): {t * 1e3:8.3f} ms  {W * width / t / 1e9:6.1f} GB/s payload  {W / t / 1e9:6.3f} G waters/s", flush=True)
    # rows of 4 waters' worth (is it the row count or the row width?)
    for width, pitch in ((48, 144), (384, 1152), (3072, 9216)):
        rows = W * 36 // pitch
        t = timed(lambda: rt.cudaMemcpy2DAsync(dev.data_ptr(), width, host.data_ptr(), pitch, width, rows, H2D, None), reps=3)
        print(f"2D width {width:4d} pitch {pitch:4d} rows {rows}: {t * 1e3:8.3f} ms  {rows * width / t / 1e9:6.1f} GB/s payload", flush=True)


if __name__ == "__main__":
    main()
