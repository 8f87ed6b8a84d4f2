"""GPU: every kernel family on the bounds-check build (build/libnw_check.so, `make check`: every global-memory index of the
strip kernels is asserted against its allocation).  compute-sanitizer is closed on the pool this runs on; this is the
in-tree substitute.  Runs in a subprocess because the library is chosen at import time (NW_CUDA_LIB)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_every_kernel_family_on_the_bounds_check_build(gpu):
    lib = os.path.join(ROOT, "build", "libnw_check.so")
    if not os.path.exists(lib):
        pytest.fail(f"{lib} is missing: run `make check` (or __graft_entry__.build())")
    env = dict(os.environ, NW_CUDA_LIB=lib)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanity_small.py")], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.strip().endswith("ok")
