"""Multi-GPU path on CPU: world_size-2 gloo processes shard a sequence of frame pairs by contiguous
blocks (no data-path collective), process their block, and do the single final gather.  The per-pair
work is done by the oracle here (tests/ may use it); on the GPU box bench.py plugs the CUDA path in."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions(pkg):
    import importlib
    seq = importlib.import_module("rs-aware-differential-sfm_b200.sequence")
    for num, world in ((1000, 8), (7, 2), (3, 4), (0, 2), (1000, 1)):
        got = []
        for r in range(world):
            lo, hi = seq.shard_range(num, r, world)
            assert 0 <= lo <= hi <= num and (hi - lo) in (num // world, num // world + 1)
            got.extend(range(lo, hi))
        assert got == list(range(num))
    assert seq.shard_range(1000, 3, 8) == (375, 500)          # 125 pairs per GPU (BASELINE config 5)


def _worker(rank, world, port, num_pairs, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq = importlib.import_module("rs-aware-differential-sfm_b200.sequence")
    synth = importlib.import_module("rs-aware-differential-sfm_b200.synth")
    from oracle import pyoracle as O
    import helpers

    def process_pair(p):
        c = helpers.make_case(O, synth, 24, 32, (40.0, 40.0, 16.0, 12.0), H=4, tol=0.02, seed=100 + p)
        r = O.refine_rectify(c["flow"], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"], c["ransac"]["v"], c["ransac"]["w"],
                             c["ransac"]["k"], False, False, c["P"]["image"], c["K4"], c["gamma"])
        return r["v"], r["w"], r["k"], r["summary"]["iterations"]

    lo, rec = seq.run_shard(process_pair, num_pairs, rank, world)
    full = seq.gather_records(rec, num_pairs, dist)
    t = seq.max_over_ranks(10.0 + rank, dist)
    np.save(os.path.join(out_dir, "full_%d.npy" % rank), full)
    np.save(os.path.join(out_dir, "t_%d.npy" % rank), np.array([t, lo]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_final_gather(tmp_path, oracle, synth):
    import importlib
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    seq = importlib.import_module("rs-aware-differential-sfm_b200.sequence")
    num_pairs, world, port = 5, 2, 29731
    mp.spawn(_worker, args=(world, port, num_pairs, str(tmp_path)), nprocs=world, join=True)
    # single-process result for comparison
    def process_pair(p):
        c = helpers.make_case(oracle, synth, 24, 32, (40.0, 40.0, 16.0, 12.0), H=4, tol=0.02, seed=100 + p)
        r = oracle.refine_rectify(c["flow"], c["inliers3"], c["alpha_in"], c["alpha_k_in"], c["m"], c["ransac"]["v"], c["ransac"]["w"],
                                  c["ransac"]["k"], False, False, c["P"]["image"], c["K4"], c["gamma"])
        return r["v"], r["w"], r["k"], r["summary"]["iterations"]
    _, ref = seq.run_shard(process_pair, num_pairs, 0, 1)
    f0 = np.load(tmp_path / "full_0.npy"); f1 = np.load(tmp_path / "full_1.npy")
    assert np.array_equal(f0, f1) and np.array_equal(f0, ref)          # same records, in pair order, on every rank
    t0 = np.load(tmp_path / "t_0.npy"); t1 = np.load(tmp_path / "t_1.npy")
    assert t0[0] == 11.0 and t1[0] == 11.0                             # max over ranks
    assert (t0[1], t1[1]) == (0, 3)                                    # contiguous blocks [0,3) and [3,5)


def test_run_shard_sequence_batches_and_records(pkg):
    """Host logic of sequence.run_shard_sequence with a stand-in context: every pair of the rank's block is
    submitted exactly once, in order, in batches of at most `batch`, and the records carry v, w, k, iterations."""
    import importlib
    seq = importlib.import_module("rs-aware-differential-sfm_b200.sequence")

    class FakeCtx:
        def __init__(self):
            self.calls = []

        def refine_rectify_sequence(self, pairs, const_acc, gs_mode, K4, gamma, layout=0):
            self.calls.append([p["id"] for p in pairs])
            return [dict(v=np.full(3, p["id"]), w=np.full(3, -p["id"]), k=0.5 * p["id"], summary=dict(iterations=10 + p["id"])) for p in pairs]

    for world in (1, 2, 3):
        seen = []
        for rank in range(world):
            ctx = FakeCtx()
            lo, rec, res = seq.run_shard_sequence(ctx, lambda i: dict(id=i), 11, rank, world, True, False, None, 0.95, batch=4)
            assert all(len(c) <= 4 for c in ctx.calls)
            ids = [i for c in ctx.calls for i in c]
            assert ids == list(range(lo, lo + rec.shape[0])) and len(res) == rec.shape[0]
            assert np.array_equal(rec[:, 0], np.array(ids, dtype=float)) and np.array_equal(rec[:, 6], 0.5 * np.array(ids))
            assert np.array_equal(rec[:, 7], 10.0 + np.array(ids))
            seen += ids
        assert seen == list(range(11))
