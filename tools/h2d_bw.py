"""Aggregate pinned host<->device copy bandwidth with all N ranks copying at once (the ceiling of the end-to-end batched path):
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/h2d_bw.py [bind]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
place = bench.bind_to_gpu_cpus(lr) if "bind" in sys.argv else None
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cpu = dist.new_group(backend="gloo")
MB = 64
h_in = torch.empty(MB << 20, dtype=torch.uint8).pin_memory(); h_in.fill_(1)
h_out = torch.empty(MB << 20, dtype=torch.uint8).pin_memory(); h_out.fill_(0)
d = torch.empty(MB << 20, dtype=torch.uint8, device="cuda"); d2 = torch.ones(MB << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(mode, reps=20):
    torch.cuda.synchronize(); dist.barrier(group=cpu)
    t0 = time.perf_counter()
    for _ in range(reps):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1): d.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2): h_out.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nbytes = reps * (MB << 20) * (2 if mode == "both" else 1)
    return nbytes / float(t.item()) / 1e9
for mode in ("h2d", "d2h", "both"):
    run(mode, 3)
    bw = run(mode)
    if rank == 0:
        print("N=%d %s %s: %.1f GB/s per GPU, %.1f GB/s aggregate (slowest rank)" % (world, "bound" if place else "unbound", mode, bw, bw * world), flush=True)
if rank == 0 and place: print(place)
dist.destroy_process_group()
