"""Per-category device time of one batched CKKS multiplication (tensor_apply + relinearize), N = 2^15, base2k = 52, 14 limbs."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poulpy_b200 as pb
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 15
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n, k, size = 1 << log_n, 52, 14
rng = np.random.default_rng(1)
m = pb.Module(n, pb.NTT120)
tsk = m.vmp_pmat_alloc(size, 1, 2, size + 1)
a = m.vec_znx_from_numpy(rng.integers(-(1 << 51), 1 << 51, size=(B, size, 2, n), dtype=np.int64))
b = m.vec_znx_from_numpy(rng.integers(-(1 << 51), 1 << 51, size=(B, size, 2, n), dtype=np.int64))
tensor, r = m.vec_znx_alloc(3, size, B), m.vec_znx_alloc(2, size, B)
s0 = s1 = None
for _ in range(2):
    s0 = m.glwe_tensor_apply(size * k, tensor, k, a, size * k, b, size * k, k, s0)
    s1 = m.glwe_tensor_relinearize(r, k, tensor, k, tsk, k, 1, s1)
m.sync()
lib = pb.lib()
lib.pgb_profile_category_name.restype = C.c_char_p
for phase in ("tensor_apply", "relinearize"):
    lib.pgb_profile_enable(m._h, 1)
    for _ in range(3):
        if phase == "tensor_apply":
            s0 = m.glwe_tensor_apply(size * k, tensor, k, a, size * k, b, size * k, k, s0)
        else:
            s1 = m.glwe_tensor_relinearize(r, k, tensor, k, tsk, k, 1, s1)
    ms = (C.c_double * 7)(); cnt = (C.c_uint64 * 7)()
    lib.pgb_profile_read(m._h, ms, cnt, 1)
    lib.pgb_profile_enable(m._h, 0)
    print(phase, "total ms/batch", round(sum(ms) / 3, 3), {lib.pgb_profile_category_name(i).decode(): (round(ms[i] / 3, 3), cnt[i] // 3) for i in range(7) if cnt[i]})
