"""Developer probe (torchrun, 2+ GPUs): cost of the halo exchange and the all-reduce."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fem-libraries_b200")]
import torch, torch.distributed as td
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 2897 * 2
v = torch.zeros(4 * n + 1000000, dtype=torch.float64, device="cuda")
peer = rank ^ 1
def halo():
    ops = [td.P2POp(td.isend, v[:2 * n], peer), td.P2POp(td.irecv, v[2 * n:4 * n], peer)]
    for r in td.batch_isend_irecv(ops):
        r.wait()
s1 = torch.zeros(1, dtype=torch.float64, device="cuda")
def ar():
    td.all_reduce(s1)
def timeit(fn, k=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); td.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k * 1e3, (time.perf_counter() - t0) / k * 1e6
if rank == 0:
    print("p2p access 0->1:", torch.cuda.can_device_access_peer(0, 1))
for name, fn in (("halo p2p 93KB", halo), ("all_reduce 8B", ar)):
    dev_us, host_us = timeit(fn)
    if rank == 0:
        print(f"{name}: device {dev_us:.1f} us/call, host {host_us:.1f} us/call", flush=True)
td.destroy_process_group()
