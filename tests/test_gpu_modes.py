"""The matcher picks kernel shapes by workload (thread- or warp-per-row gating, K6b for combination-heavy pairs).
Parity must not depend on the choice: the parity suite is re-run in subprocesses with every shape forced."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("env", [
    {"QMS_GATE_MODE": "1", "QMS_HEAVY_MIN": "2"},      # warp-per-row gate, K6b for every pair with >= 2 combinations
    {"QMS_GATE_MODE": "0", "QMS_HEAVY_MIN": "0"},      # thread-per-row gate, K6 only
    {"QMS_PROBE_VARIANT": "3"},                        # thread-strided probe, 8 rows per thread, evict-first loads
    {"QMS_PROBE_VARIANT": "0", "QMS_PROBE_CTAS": "3", "QMS_GATE_CTAS": "1"},   # thread-strided probe, small grids
    {"QMS_PROBE_VARIANT": "5", "QMS_PROBE_CTAS": "128"},                       # warp-contiguous probe, 8 loads per lane
])
def test_parity_with_forced_kernel_shapes(env):
    cmd = [sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-m", "gpu", "-q", "-x", "-k",
           "dense or bsa_pct or overlap or crafted or huge or c1_basic or edge or multi_label"]
    done = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    assert done.returncode == 0, done.stdout[-3000:] + done.stderr[-2000:]
    assert " passed" in done.stdout
