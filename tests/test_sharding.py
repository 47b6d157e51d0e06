"""CPU suite: pair sharding + the single sum all-reduce, world_size 2 over gloo."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_plan_covers_every_pair_once(tcl):
    frames = tcl.synth.SINTEL_TRAIN_FRAMES
    pairs = tcl.sharding.pairs_per_sequence(frames)
    assert sum(pairs) == 1041 and sum(tcl.sharding.pairs_per_sequence(frames, gap=5)) == 949
    for world in (1, 2, 3, 4, 8):
        plans = [tcl.plan_shards(pairs, world, r) for r in range(world)]
        assert plans[0].start == 0 and plans[-1].stop == sum(pairs)
        for a, b in zip(plans, plans[1:]):
            assert a.stop == b.start
        sizes = [p.n_local for p in plans]
        assert max(sizes) - min(sizes) <= 1
        glob_seq = sum((p.seq_of_pair for p in plans), [])
        assert glob_seq == [s for s, n in enumerate(pairs) for _ in range(n)]
    with pytest.raises(ValueError):
        tcl.plan_shards(pairs, 2, 2)


def _fake_pair_results(n_pairs, seed=0):
    rng = np.random.default_rng(seed)
    sums = rng.random(n_pairs) * 1e4
    return sums


def _worker(rank, world, port, pairs, elems, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tcl_b200 as tcl
    plan = tcl.plan_shards(pairs, world, rank)
    sums = torch.from_numpy(_fake_pair_results(sum(pairs)))[plan.start:plan.stop]
    vals = (sums / elems).sqrt().float()
    packed = tcl.sharding.pack_local(vals, sums.sum(), torch.tensor(plan.seq_of_pair, dtype=torch.long), len(pairs), elems)
    out = tcl.sharding.unpack(tcl.allreduce_sums(packed), len(pairs))
    torch.save({k: v.clone() for k, v in out.items()}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_allreduce_of_packed_sums_gloo(tcl, tmp_path, world):
    pairs = [7, 3, 0, 12, 5]
    elems = 3 * 8 * 8
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, pairs, elems, str(tmp_path)), nprocs=world, join=True)
    sums = _fake_pair_results(sum(pairs))
    vals = np.sqrt(sums / elems).astype(np.float32).astype(np.float64)
    seq = np.array([s for s, n in enumerate(pairs) for _ in range(n)])
    per_seq = np.array([vals[seq == s].mean() if (seq == s).any() else 0.0 for s in range(len(pairs))])
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert np.allclose(got["per_sequence_mean"].numpy(), per_seq, rtol=1e-12)
        present = np.array([n > 0 for n in pairs])
        assert np.isclose(float(got["mean_over_sequences"]), per_seq[present].mean(), rtol=1e-12)
        assert np.isclose(float(got["mean_over_pairs"]), vals.mean(), rtol=1e-12)
        assert np.isclose(float(got["pooled_rmse"]), np.sqrt(sums.sum() / (sum(pairs) * elems)), rtol=1e-12)
        assert int(got["n_pairs"]) == sum(pairs)


def test_single_process_is_identity(tcl):
    packed = torch.arange(8, dtype=torch.float64)
    assert torch.equal(tcl.allreduce_sums(packed.clone()), packed)


def test_aggregate_means_matches_reference_formula(tcl):
    # utils/sintel_eval.py:112-126 : mean over keys, per-style means over len/3 keys
    d = {"alley_1_s1": 0.1, "alley_1_s2": 0.2, "alley_1_s3": 0.4, "bamboo_2_s1": 0.3, "bamboo_2_s2": 0.5, "bamboo_2_s3": 0.9}
    out = tcl.aggregate_means("TCL-ST", d, 4)
    assert np.isclose(out["TCL-ST_mean"], np.mean(list(d.values())))
    assert np.isclose(out["TCL-ST_mean_s1"], 0.2) and np.isclose(out["TCL-ST_mean_s2"], 0.35) and np.isclose(out["TCL-ST_mean_s3"], 0.65)
