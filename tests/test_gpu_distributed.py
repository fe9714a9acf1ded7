"""Multi-GPU path on real devices (skipped on a 1-GPU box): sharded envelope build + NCCL all-gather must
give a library bit-identical to the single-GPU build; spectra shards matched per rank and gathered on rank 0
must equal the single-GPU hit list."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    import pyqms_b200
    from pyqms_b200 import distributed as qd
    from pyqms_b200.synthetic import BSA_MOLECULES, CAM_FIXED_LABEL, random_tryptic_peptides, synth_run
    torch.cuda.set_device(rank)
    import datetime
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=120))
    kwargs = dict(molecules=BSA_MOLECULES + random_tryptic_peptides(300, seed=9), charges=[1, 2, 3], fixed_labels=CAM_FIXED_LABEL,
                  metabolic_labels={"15N": [0, 0.5, 0.99]}, verbose=False, device=rank)
    sharded = pyqms_b200.IsotopologueLibrary(distributed=True, **kwargs)
    single = pyqms_b200.IsotopologueLibrary(distributed=False, **kwargs)
    for name in ("_ent_env", "_ent_z", "_ent_lower", "_ent_upper", "_ent_win_off", "_win_lo", "_win_hi", "_win_cmz", "_win_ri", "_win_ci"):
        assert np.array_equal(getattr(sharded, name), getattr(single, name)), name
    a, b = sharded._export_envelopes(), single._export_envelopes()
    for name in ("mass", "relabun", "abun", "pos", "c_flag", "len", "n_c"):
        assert np.array_equal(a[name], b[name]), name
    # spectra sharded by contiguous ranges, library replicated, hits gathered in spectrum order
    mz, ci = single.entry_peaks()
    peaks, offsets = synth_run(mz, ci, n_spectra=64, seed=31, mean_peaks=300, elution_sigma=2.0, entry_fraction=0.05)
    whole = single.match_packed(peaks, offsets)
    begin, end = qd.shard_range(64, rank, world)
    mine = sharded.match_packed(peaks[offsets[begin]:offsets[end]], offsets[begin:end + 1] - offsets[begin])
    merged = qd.gather_packed_hits(mine, begin, world)
    if rank == 0:
        nh, nw = whole["n_hits"], whole["n_windows"]
        assert merged["n_hits"] == nh and nh > 20
        for name in ("spec", "entry", "score", "scaling"):
            assert np.array_equal(merged[name], whole[name][:nh]), name
        assert np.array_equal(merged["win_off"], whole["win_off"][:nh + 1])
        assert np.array_equal(merged["mmz"], whole["mmz"][:nw], equal_nan=True)
    assert sharded._build_stats["sharded"] and sharded._build_stats["how"].startswith("one ncclAllGather")
    assert sharded._build_stats["bytes_received"] > 0
    # the same through the shared-memory arena: the library copies each rank's hits into the rank's page-locked
    # segment, rank 0 maps the segments of all ranks (no pickling, no collective on the data path)
    arena = qd.SharedHitArena()
    sharded.use_hit_arena(arena)
    for step in range(2):
        mine = sharded.match_packed(peaks[offsets[begin]:offsets[end]], offsets[begin:end + 1] - offsets[begin])
        batches = arena.collect(mine, begin)
        if rank == 0:
            nh, nw = whole["n_hits"], whole["n_windows"]
            assert [b["rank"] for b in batches] == list(range(world))
            assert np.array_equal(np.concatenate([b["spec"] + b["spec_base"] for b in batches]), whole["spec"][:nh])
            for name in ("entry", "score", "scaling"):
                assert np.array_equal(np.concatenate([b[name] for b in batches]), whole[name][:nh]), name
            assert np.array_equal(np.concatenate([b["mmz"] for b in batches]), whole["mmz"][:nw], equal_nan=True)
            shift = np.cumsum([0] + [b["n_windows"] for b in batches])
            assert np.array_equal(np.concatenate([b["win_off"][:-1] + shift[k] for k, b in enumerate(batches)]), whole["win_off"][:nh])
            # a peak row counts from the first row of the rank's shard
            row_shift = [offsets[qd.shard_range(64, b["rank"], world)[0]] for b in batches]
            rows = np.concatenate([np.where(b["peak"] >= 0, b["peak"] + row_shift[k], -1) for k, b in enumerate(batches)])
            assert np.array_equal(rows, whole["peak"][:nw])
        else:
            assert batches is None
        dist.barrier()
    del batches, mine
    sharded.use_hit_arena(None)
    arena.close()
    dist.barrier()
    dist.destroy_process_group()
    Path(out_dir, "ok%d" % rank).write_text("ok")


def test_two_gpu_build_and_match(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
