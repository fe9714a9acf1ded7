"""bench.py's contract that can be checked without a GPU: the reference arm prints one JSON line with the keys the driver
reads, both arms share one `config` dict, and a planted-peak table reproduces a synthetic run without a library."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_planted_table_reproduces_the_run():
    from pyqms_b200.synthetic import synth_run
    rng = np.random.default_rng(0)
    mz = [np.sort(rng.uniform(150, 2000, 5)) for _ in range(400)]
    ci = [rng.uniform(1, 100, 5) for _ in range(400)]
    peaks, offsets, table = synth_run(mz, ci, n_spectra=40, seed=3, entry_fraction=0.1, return_planted=True)
    again, offsets_again = synth_run(None, None, n_spectra=40, seed=3, entry_fraction=0.1, planted=table)
    assert np.array_equal(peaks, again) and np.array_equal(offsets, offsets_again)
    assert len(table["lens"]) == 40 and table["n_entries"] == 400


def test_bench_fixture_matches_its_declared_shape():
    sys.path.insert(0, str(ROOT))
    import bench
    z = np.load(bench.FIXTURE)
    assert int(z["n_entries"]) == 1585058 and len(z["lens"]) == 31701 and int(z["lens"].sum()) == len(z["mz"]) == len(z["ci"])
    args = type("A", (), {"workload": "proteome", "proteome_size": 500000})()
    assert bench.static_config(args) == bench.static_config(args)
    assert bench.static_config(args)["library_peptides"] == 500000


@pytest.mark.timeout(300)
def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` on the small configuration: the unmodified reference from baseline/_ref when it is
    installed (build container, GPU box), else the oracle port."""
    done = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "bsa", "--steps", "1",
                           "--warmup", "0", "--ref-spectra-per-core", "2"], capture_output=True, text=True, cwd=ROOT, timeout=280)
    assert done.returncode == 0, done.stderr[-2000:]
    lines = [line for line in done.stdout.splitlines() if line.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert (ROOT / "baseline" / "_ref" / "pyqms").exists() == (line["cpu_baseline"]["kind"] == "reference")
