"""Does the host side of the end-to-end step scale?  Every rank repeats only the COPIES of one C3 step (0.32 GB of
pinned input rows to its GPU, 0.70 GB of rows back, on two streams at once, no kernels) between barriers.  If this
alone slows down as ranks are added, the e2e leg is bound by the host's memory system / PCIe topology, not by the
host threads.  torchrun --nproc-per-node N scripts/copy_scaling.py [h2d_MB] [d2h_MB]"""
import os, sys, time
import torch, torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h2d_mb = float(sys.argv[1]) if len(sys.argv) > 1 else 320.0
d2h_mb = float(sys.argv[2]) if len(sys.argv) > 2 else 700.0
src = torch.empty(int(h2d_mb * 1e6), dtype=torch.uint8).pin_memory()
dst = torch.empty(int(d2h_mb * 1e6), dtype=torch.uint8).pin_memory()
d_in = torch.empty_like(src, device="cuda")
d_out = torch.empty_like(dst, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def step(which):
    if which in ("both", "h2d"):
        with torch.cuda.stream(s_in):
            d_in.copy_(src, non_blocking=True)
    if which in ("both", "d2h"):
        with torch.cuda.stream(s_out):
            dst.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()


for which in ("h2d", "d2h", "both"):
    for _ in range(3):
        step(which)
    barrier()
    t0 = time.perf_counter()
    K = 10
    for _ in range(K):
        step(which)
    barrier()
    ms = (time.perf_counter() - t0) / K * 1e3
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        gb = {"h2d": h2d_mb, "d2h": d2h_mb, "both": h2d_mb + d2h_mb}[which] / 1e3
        print("N=%d %-4s %.1f ms per step (max over ranks)  %.1f GB/s per GPU, %.1f GB/s host total" % (
            world, which, t.item(), gb / (t.item() / 1e3), world * gb / (t.item() / 1e3)), flush=True)
if world > 1:
    dist.destroy_process_group()
