"""N>1 host path on CPU: world_size-2 gloo processes exercise shard assembly and hit gathering."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from pyqms_b200 import distributed as qd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert qd.world_size() == world and qd.rank() == rank
    # --- envelope shard assembly: ragged slot ranges, several dtypes
    rng = np.random.default_rng(5)
    n_env = 11
    lens = rng.integers(1, 9, size=n_env)
    slot_off = np.concatenate([[0], np.cumsum(lens)])
    truth = {np.float64: rng.normal(size=slot_off[-1]), np.int32: rng.integers(0, 1 << 20, size=slot_off[-1]).astype(np.int32),
             np.uint8: rng.integers(0, 2, size=slot_off[-1]).astype(np.uint8)}
    env_ranges = [qd.shard_range(n_env, r, world) for r in range(world)]
    slot_ranges = [(int(slot_off[b]), int(slot_off[e])) for b, e in env_ranges]
    for dtype, full_truth in truth.items():
        mine = np.zeros_like(full_truth)
        b, e = slot_ranges[rank]
        mine[b:e] = full_truth[b:e]                     # this rank computed only its shard
        t = torch.from_numpy(mine)
        qd.allgather_ranges(t, slot_ranges)
        assert np.array_equal(t.numpy(), full_truth), dtype
    # --- hits of spectrum shards come back in spectrum order on rank 0
    n_spec = 9
    begin, end = qd.shard_range(n_spec, rank, world)
    nh = 2 * (end - begin)
    local = {"spec": np.repeat(np.arange(end - begin, dtype=np.int32), 2), "entry": np.tile(np.array([3, 7], dtype=np.int32), end - begin),
             "score": np.linspace(0.5, 0.9, nh), "scaling": np.ones(nh), "win_off": np.arange(0, 3 * nh + 1, 3, dtype=np.int64),
             "mmz": np.arange(3 * nh, dtype=np.float64) + 1000 * rank, "mi": np.ones(3 * nh), "n_hits": nh, "n_windows": 3 * nh,
             "peak": np.where(np.arange(3 * nh) % 3 == 2, -1, np.arange(3 * nh)).astype(np.int64)}
    merged = qd.gather_packed_hits(local, begin, world, row_begin=500 * rank)
    if rank == 0:
        assert merged["n_hits"] == 2 * n_spec and merged["n_windows"] == 6 * n_spec
        assert list(merged["spec"]) == [s for s in range(n_spec) for _ in range(2)]
        assert list(merged["win_off"]) == list(range(0, 6 * n_spec + 1, 3))
        first_of_rank_1 = 6 * (qd.shard_range(n_spec, 1, world)[0])
        assert merged["mmz"][first_of_rank_1] == 1000.0
        assert merged["peak"][first_of_rank_1] == 500 and merged["peak"][first_of_rank_1 + 2] == -1 and merged["peak"][1] == 1
    else:
        assert merged is None
    # --- the same through the shared-memory arena: every rank's hits land in its own segment, rank 0 maps them all
    # (register=False: no CUDA here; on a GPU box the owning rank page-locks its segment and the library DMAs into it)
    from pyqms_b200.isotopologue_library import PackedHits
    arena = qd.SharedHitArena(tag="qmstest_%d" % port, register=False)
    for step in range(2):                               # the second step reuses the segments (and peers' mappings)
        target = arena.arrays(nh, 3 * nh, True)
        out = PackedHits()
        for field in ("spec", "entry", "score", "scaling", "win_off", "mmz", "mi"):
            target[field][:len(local[field])] = local[field] if field != "score" else local[field] + step
            out[field] = target[field][:len(local[field])]
        target["peak"][:3 * nh] = np.arange(3 * nh) + 7 * rank
        out["peak"] = target["peak"][:3 * nh]
        out["n_hits"], out["n_windows"] = nh, 3 * nh
        batches = arena.collect(out, begin)
        if rank == 0:
            assert [b["rank"] for b in batches] == list(range(world))
            assert [b["spec_base"] for b in batches] == [qd.shard_range(n_spec, r, world)[0] for r in range(world)]
            spectra = np.concatenate([b["spec"] + b["spec_base"] for b in batches])
            assert list(spectra) == [s for s in range(n_spec) for _ in range(2)]
            assert batches[1]["mmz"][0] == 1000.0 and batches[1]["peak"][1] == 8 and batches[1]["score"][0] == 0.5 + step
            assert sum(b["n_hits"] for b in batches) == 2 * n_spec
        else:
            assert batches is None
        dist.barrier()                                  # rank 0 has read the segments before anybody writes them again
    # --- a step in the compact form (what a deferred match_many fetches): entry / score / scaling per hit, uint16 rows per
    # window, hit offsets per spectrum; rank 0 widens spec / win_off / peak of every rank from them
    n_local = end - begin
    row_off = 100 * rank + 40 * np.arange(n_local + 1, dtype=np.int64)            # first row of every spectrum of the shard
    entry_windows = np.full(16, 3, dtype=np.int64)
    arena.entry_windows = entry_windows
    target = arena.arrays(nh, 3 * nh, "compact", n_local)
    assert set(target) == {"entry", "score", "scaling", "peak16", "spec_off"}
    out = PackedHits(None, row_off, entry_windows)
    for field in ("entry", "score", "scaling"):
        target[field][:nh] = local[field]
        out[field] = target[field][:nh]
    target["peak16"][:3 * nh] = np.where(np.arange(3 * nh) % 3 == 2, 0xFFFF, np.arange(3 * nh) % 7)
    target["spec_off"][:n_local + 1] = 2 * np.arange(n_local + 1)
    out["peak16"], out["spec_off"] = target["peak16"][:3 * nh], target["spec_off"][:n_local + 1]
    out["n_hits"], out["n_windows"] = nh, 3 * nh
    assert list(out["spec"]) == list(local["spec"]) and list(out["win_off"]) == list(local["win_off"])
    assert out["peak"][2] == -1 and out["peak"][4] == row_off[0] + 4 and out["peak"][6] == row_off[1] + 6
    batches = arena.collect(out, begin)
    if rank == 0:
        assert all("spec" not in dict.keys(b) and "win_off" not in dict.keys(b) for b in batches)      # widened on access, not shipped
        spectra = np.concatenate([b["spec"] + b["spec_base"] for b in batches])
        assert list(spectra) == [s for s in range(n_spec) for _ in range(2)]
        for b in batches:
            assert list(b["win_off"]) == list(range(0, 3 * b["n_hits"] + 1, 3))
            assert b["peak"][2] == -1 and b["peak"][6] == 40 + 6          # rows counted from the first row of the rank's shard
    dist.barrier()
    del batches, out, target
    arena.close()
    dist.barrier()
    dist.destroy_process_group()
    Path(out_dir, "ok%d" % rank).write_text("ok")


@pytest.mark.timeout(300)
def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
