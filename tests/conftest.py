import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # On a GPU box import torch before the first test touches the CUDA library, the order bench.py and a torchrun worker
    # have anyway: test_gpu_distributed.py needs it to count the devices, and a run that reached that file last -- after
    # 80 s of contexts, pinned pools and reader threads in the same process -- once saw that first import fail
    # (ImportError; not reproducible in isolation, where `Context(0)` followed by `import torch` works).  A CPU-only
    # container has no /proc/driver/nvidia and pays nothing.
    if Path("/proc/driver/nvidia/gpus").is_dir():
        import torch  # noqa: F401


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"
