"""The committed golden vectors ARE what the real reference produces: when /root/reference is
present (the build container), regenerate every .npz with the committed generator scripts into a
scratch directory and compare bit for bit.  Skipped where the reference is absent (GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.mark.skipif(not os.path.isdir("/root/reference/pygho"), reason="reference not mounted")
def test_committed_goldens_are_reproducible_from_the_reference(tmp_path):
    env = dict(os.environ, PYGHO_GOLDEN_OUT=str(tmp_path), PYTHONDONTWRITEBYTECODE="1")
    for script, extra in (("make_golden.py", []), ("make_golden.py", ["spmamm"]),
                          ("make_golden_hodata.py", [])):
        r = subprocess.run([sys.executable, os.path.join(GOLDEN, script)] + extra, env=env, cwd=ROOT,
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
    committed = sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz"))
    fresh = sorted(f for f in os.listdir(tmp_path) if f.endswith(".npz"))
    assert fresh == committed
    for f in committed:
        a, b = dict(np.load(os.path.join(GOLDEN, f))), dict(np.load(os.path.join(tmp_path, f)))
        assert set(a) == set(b), f
        for k in a:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, (f, k)
            assert np.array_equal(a[k], b[k], equal_nan=a[k].dtype.kind == "f"), (f, k)
