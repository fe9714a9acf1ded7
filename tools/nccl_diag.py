"""Staged 2-rank NCCL diagnostic (prints after every stage; run under a short timeout)."""
import datetime
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist


def say(*a):
    print("[rank %s %.1fs]" % (os.environ.get("RANK"), time.time() - T0), *a, flush=True)


T0 = time.time()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=90), device_id=torch.device("cuda", local))
say("init done")
x = torch.ones(1024, device="cuda") * (rank + 1)
dist.all_reduce(x)
torch.cuda.synchronize()
say("all_reduce ok", float(x[0]))
pieces = [torch.empty(1000, device="cuda", dtype=torch.float64) for _ in range(world)]
dist.all_gather(pieces, torch.full((1000,), float(rank), device="cuda", dtype=torch.float64))
torch.cuda.synchronize()
say("all_gather ok", [float(p[0]) for p in pieces])
import pyqms_b200
from pyqms_b200.synthetic import BSA_MOLECULES, CAM_FIXED_LABEL
kw = dict(molecules=BSA_MOLECULES, charges=[1, 2, 3], fixed_labels=CAM_FIXED_LABEL, metabolic_labels={"15N": [0, 0.5, 0.99]}, verbose=False,
          device=local)
single = pyqms_b200.IsotopologueLibrary(distributed=False, **kw)
say("single-GPU library built", single._n_entries)
sharded = pyqms_b200.IsotopologueLibrary(distributed=True, **kw)
say("sharded library built", sharded._n_entries)
dist.barrier()
say("barrier ok")
dist.destroy_process_group()
say("done")
