"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root: python scripts/make_golden.py).  See tests/golden_cases.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from golden_cases import CASES  # noqa: E402

for i, (name, (mk, orc, _)) in enumerate(sorted(CASES.items())):
    inp = mk(np.random.default_rng(7000 + i))
    out = orc(inp)
    # digits are far below 2^31: int32 storage keeps the fixtures small (loaded back as int64)
    assert all(np.abs(v).max() < 2**31 for v in list(inp.values()) + [out])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), want=out.astype(np.int32),
                        **{"in_" + k: v.astype(np.int32) for k, v in inp.items()})
    print(name, out.shape, int(np.abs(out).max()))
