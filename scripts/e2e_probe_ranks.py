"""Bare PCIe probe under torchrun (one rank per GPU): every rank copies the bench's per-step volume pinned-host -> device and device ->
pinned-host CONCURRENTLY, all ranks at the same time (barrier before and after), with no compute -- the ceiling any host-buffer front end
can reach at N ranks on this box.  Then the same ranks run pgb_glwe_keyswitch_host (bench.py's e2e leg) for comparison.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/e2e_probe_ranks.py

Rank 0 prints one JSON line: per-rank and aggregate GB/s per direction, the key-switch/s ceiling they imply (393 216 B each way per
key-switch) and the measured e2e key-switch/s."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poulpy_b200 as pb

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, k, B = 4096, 18, 4096
nbytes = B * 3 * 2 * n * 8
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda"); d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

def max_ranks(x):
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

res = {"n_gpus": world, "bytes_per_direction_per_rank": nbytes}
for mode in ("h2d", "d2h", "both"):
    for it in range(2):  # first pass warms the pages
        barrier(); t0 = time.perf_counter()
        for _ in range(4):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        barrier(); dt = max_ranks((time.perf_counter() - t0) / 4)
    res[mode] = {"gbs_per_rank_per_direction": nbytes / dt / 1e9, "gbs_aggregate_per_direction": world * nbytes / dt / 1e9,
                 "keyswitch_per_s_ceiling": world * B / dt}
m = pb.Module(n, pb.NTT120, device=local)
rng = np.random.default_rng(1)
mat = rng.integers(-(1 << 17), 1 << 17, size=(3, 1, 4, 2, n), dtype=np.int64)
pm = m.vmp_pmat_alloc(3, 1, 2, 4); m.vmp_prepare(pm, m.mat_znx_from_numpy(mat)); m.gadget_key_pin(pm)
a = pb.pinned_empty((B, 3, 2, n)); a[:] = rng.integers(-(1 << 17), 1 << 17, size=a.shape, dtype=np.int64)
r = pb.pinned_empty((B, 3, 2, n))
m.glwe_keyswitch_host(r, k, a, k, pm, k, 1)
barrier(); t0 = time.perf_counter()
for _ in range(5): m.glwe_keyswitch_host(r, k, a, k, pm, k, 1)
barrier(); dt = max_ranks((time.perf_counter() - t0) / 5)
res["keyswitch_host"] = {"keyswitch_per_s": world * B / dt, "frac_of_probe_ceiling": world * B / dt / res["both"]["keyswitch_per_s_ceiling"]}
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
