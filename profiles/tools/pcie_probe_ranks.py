"""Copy-only PCIe probe on N ranks of one box (one rank per GPU): what the host side of the link sustains when every GPU
copies at once.  No kernels of the path run here; this states the ceiling the end-to-end numbers live under.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 \
        profiles/tools/pcie_probe_ranks.py > profiles/r2/pcie_probe_nN.txt

Per pattern: every rank repeats its copies `reps` times between two barriers; reported is the aggregate over ranks
(total bytes / max-over-ranks wall time) and the slowest / fastest rank.
"""
import os
import time

import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
try:
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
except Exception:
    pass
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

n = 1 << 21


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(label, fn, nbytes, reps=10):
    fn()
    barrier()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    barrier()
    ts = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        all_t = [torch.zeros_like(ts) for _ in range(world)]
        dist.all_gather(all_t, ts)
        all_t = [float(x.item()) for x in all_t]
    else:
        all_t = [dt]
    if rank == 0:
        worst, best = max(all_t), min(all_t)
        print(f"{label:58s} {worst * 1e3:8.3f} ms  aggregate {world * nbytes / worst / 1e9:7.1f} GB/s  "
              f"per rank {nbytes / worst / 1e9:6.1f} .. {nbytes / best / 1e9:6.1f} GB/s", flush=True)


def bufs(bytes_in, bytes_out):
    return (torch.empty(bytes_in, dtype=torch.uint8).pin_memory(), torch.empty(bytes_in, dtype=torch.uint8, device=dev),
            torch.empty(bytes_out, dtype=torch.uint8).pin_memory(), torch.empty(bytes_out, dtype=torch.uint8, device=dev))


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
if rank == 0:
    print(f"# {world} rank(s), one per GPU; 2^21 items per rank per repetition; pinned host buffers; {os.cpu_count()} host cpus")
for name, bi, bo in (("struct API      27 B in, 36 B out per item", 27, 36), ("compact output  27 B in, 22.4 B out per item", 27, 22.4),
                     ("packed wire v2  16 B in, 14.2 B out per item", 16, 14.2)):
    hin, din, hout, dout = bufs(int(n * bi), int(n * bo))

    def h2d():
        din.copy_(hin, non_blocking=True)

    def d2h():
        hout.copy_(dout, non_blocking=True)

    def both():
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
    if rank == 0:
        print(f"## {name}")
    run("H2D only", h2d, int(n * bi))
    run("D2H only", d2h, int(n * bo))
    run("H2D + D2H concurrent (bytes = both directions)", both, int(n * bi) + int(n * bo))
    del hin, din, hout, dout
if world > 1:
    dist.destroy_process_group()
