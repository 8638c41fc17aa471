"""Host<->device copy bandwidth from pinned memory; under torchrun all ranks copy at the same time
(shows what the host can feed to N GPUs at once):  python -m torch.distributed.run --nproc-per-node 8 tools/pcie_bw.py"""
import os, time
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
x = torch.empty(256 * 1024 * 1024, dtype=torch.uint8).pin_memory(); x.fill_(1); y = torch.empty_like(x, device="cuda")
z = torch.empty_like(x).pin_memory()
def run(dst, src, label):
    for _ in range(3): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    g = torch.tensor([20 * x.numel() / dt / 1e9], device="cuda")
    if world > 1:
        lst = [torch.zeros_like(g) for _ in range(world)]; dist.all_gather(lst, g)
        if rank == 0: print(label, "GB/s per rank", [round(float(v), 1) for v in lst], "sum", round(sum(float(v) for v in lst), 1), flush=True)
    else:
        print(label, "GB/s", round(float(g), 1), flush=True)
run(y, x, "H2D pinned"); run(z, y, "D2H pinned")
if world > 1: dist.destroy_process_group()
