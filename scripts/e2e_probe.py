"""PCIe ceiling vs the host front end: concurrent pinned H2D + D2H copies of the bench's per-step volume, then
pgb_glwe_keyswitch_host at several staging chunk sizes (PGB_HOST_CHUNK_MB is read once per process: run one size per process)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poulpy_b200 as pb
n, k, B = 4096, 18, 4096
rng = np.random.default_rng(1)
if len(sys.argv) > 1 and sys.argv[1] == "pcie":
    nbytes = B * 3 * 2 * n * 8
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda"); d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for mode in ("h2d", "d2h", "both"):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        print(mode, "GB/s per direction", round(nbytes / dt / 1e9, 1), "-> key-switch/s ceiling", round(B / dt))
    sys.exit(0)
m = pb.Module(n, pb.NTT120)
mat = rng.integers(-(1 << 17), 1 << 17, size=(3, 1, 4, 2, n), dtype=np.int64)
pm = m.vmp_pmat_alloc(3, 1, 2, 4); m.vmp_prepare(pm, m.mat_znx_from_numpy(mat))
a = pb.pinned_empty((B, 3, 2, n)); a[:] = rng.integers(-(1 << 17), 1 << 17, size=a.shape, dtype=np.int64)
r = pb.pinned_empty((B, 3, 2, n))
m.glwe_keyswitch_host(r, k, a, k, pm, k, 1)
t0 = time.perf_counter()
for _ in range(5): m.glwe_keyswitch_host(r, k, a, k, pm, k, 1)
dt = (time.perf_counter() - t0) / 5
print("chunk MB", os.environ.get("PGB_HOST_CHUNK_MB", "64"), "e2e key-switch/s", round(B / dt))
