"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root: python scripts/make_golden.py).  See tests/golden_cases.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from golden_cases import CASES  # noqa: E402

import zlib  # noqa: E402

# the eight round-1 fixtures keep their seeds (7000 + index in that sorted list); later cases take a seed from their name, so adding a case
# never changes an existing fixture
ROUND1 = sorted(["cggi_blind_rotate_fft64", "cggi_blind_rotate_ntt120", "glwe_external_product_fft64", "glwe_external_product_ntt120",
                 "glwe_keyswitch_fft64", "glwe_keyswitch_ntt120", "glwe_trace_fft64", "glwe_trace_ntt120"])
for name, (mk, orc, _) in sorted(CASES.items()):
    seed = 7000 + ROUND1.index(name) if name in ROUND1 else 8000 + zlib.crc32(name.encode()) % 1000
    inp = mk(np.random.default_rng(seed))
    out = orc(inp)
    # digits are far below 2^31: int32 storage keeps the fixtures small (loaded back as int64)
    assert all(np.abs(v).max() < 2**31 for v in list(inp.values()) + [out])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), want=out.astype(np.int32),
                        **{"in_" + k: v.astype(np.int32) for k, v in inp.items()})
    print(name, out.shape, int(np.abs(out).max()))
