#!/usr/bin/env python
"""Install the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box with the snapshot).

    python tools/install_reference.py [--force]

1. `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref` from a copy of /root/reference
   (the source tree is read-only and setuptools writes build/ and egg-info next to setup.py);
2. the package data pip's file list misses (pyqms/kb/ext/usermod.xml) is copied so that the package directory is
   byte-identical to /root/reference/pyqms;
3. the two parser modules the reference imports but does not vendor (`chemical_composition`, `unimod_mapper`,
   PyPI, not in the offline wheelhouse) are provided by the ~100-line stand-in of tests/golden/_standin
   (SURVEY.md 8c): it only turns molecule strings into element counts, every arithmetic line of the two hot paths
   that then runs is the reference's own.
bench.py --impl reference and bench.py's cpu_baseline import pyqms from there.
"""
import filecmp
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference")
TARGET = ROOT / "baseline" / "_ref"


def main(force=False):
    if not REFERENCE.exists():
        print("no /root/reference here: baseline/_ref is used as shipped" if TARGET.exists() else "no reference and no baseline/_ref")
        return TARGET.exists()
    if (TARGET / "pyqms" / "__init__.py").exists() and not force:
        return True
    shutil.rmtree(TARGET, ignore_errors=True)
    TARGET.mkdir(parents=True)
    with tempfile.TemporaryDirectory() as tmp:
        src = Path(tmp) / "reference"
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns(".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", str(TARGET), str(src)]
        done = subprocess.run(cmd, capture_output=True, text=True)
        if done.returncode != 0:                   # no pip / no setuptools: the package is pure Python, copy it
            shutil.copytree(REFERENCE / "pyqms", TARGET / "pyqms", dirs_exist_ok=True)
    for path in (REFERENCE / "pyqms").rglob("*"):   # package data the wheel's file list leaves out
        rel = path.relative_to(REFERENCE / "pyqms")
        if path.is_file() and "__pycache__" not in rel.parts and not (TARGET / "pyqms" / rel).exists():
            (TARGET / "pyqms" / rel).parent.mkdir(parents=True, exist_ok=True)
            shutil.copy2(path, TARGET / "pyqms" / rel)
    cmp = filecmp.dircmp(REFERENCE / "pyqms", TARGET / "pyqms", ignore=["__pycache__"])
    assert not cmp.left_only and not cmp.diff_files, (cmp.left_only, cmp.diff_files)
    for name in ("chemical_composition", "unimod_mapper"):
        shutil.copytree(ROOT / "tests" / "golden" / "_standin" / name, TARGET / name, dirs_exist_ok=True)
    print("reference installed in", TARGET)
    return True


if __name__ == "__main__":
    main(force="--force" in sys.argv[1:])
