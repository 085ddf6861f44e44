"""compute-sanitizer over every kernel family on small cases (the reference has no race detector; this is the
CUDA one).  memcheck: out-of-bounds and misaligned accesses; racecheck: shared-memory hazards (radix sort, exchange scatter, tile
kernel).  Runs tools/sanitize_small.py under the sanitizer in a subprocess; an error makes the sanitizer exit with code 9."""
import os
import shutil
import subprocess
import sys

import pytest

import helpers as H

pytestmark = pytest.mark.gpu
SAN = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"


@pytest.mark.skipif(not os.path.exists(SAN), reason="compute-sanitizer not installed")
@pytest.mark.parametrize("tool,what", [("memcheck", "dbc,mdbc,float,slab,2d"), ("racecheck", "dbc,slab")])
def test_sanitizer_clean(tool, what, built_lib):
    cmd = [SAN, "--tool", tool, "--error-exitcode", "9", sys.executable, os.path.join(H.ROOT, "tools", "sanitize_small.py"), "1", what]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout[-3000:] + r.stderr[-2000:])
    clean = "ERROR SUMMARY: 0 errors" if tool == "memcheck" else "RACECHECK SUMMARY: 0 hazards displayed (0 errors, 0 warnings)"
    assert r.returncode == 0 and "sanitize_small done" in r.stdout and clean in r.stdout + r.stderr, tail
