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


def test_band_rows_cover_every_row_once_in_whole_tile_rows(tcl):
    """Band mode (fewer pairs than GPUs): contiguous bands of whole 32-row tile rows, every target row in
    exactly one band, balanced to within one tile row, empty bands when there are more GPUs than tile rows."""
    for H in (1, 31, 32, 33, 256, 436, 1080, 2160, 16384):
        for world in (1, 2, 3, 4, 8, 16):
            bands = [tcl.band_rows(H, world, r) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == H
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0 and a0 <= a1
            for r0, r1 in bands:
                assert r0 == r1 or (r0 % 32 == 0 and (r1 % 32 == 0 or r1 == H))      # (an empty band evaluates nothing)
            tile_rows = [-(-(r1 - r0) // 32) for r0, r1 in bands]
            assert max(tile_rows) - min(tile_rows) <= 1
            if world > -(-H // 32):
                assert any(r0 == r1 for r0, r1 in bands)
    assert tcl.band_rows(2160, 8, 0) == (0, 288) and tcl.band_rows(2160, 8, 7) == (1920, 2160)


def test_window_index_arrays_describe_every_directed_evaluation(tcl):
    """Window mode (BASELINE config 4; utils/sintel_eval.py:84-86,216-222 generalised to both directions): every target
    frame t and source s = t-1 .. t-(window-1) appears once per direction, field 2j is flow(t -> s) and 2j+1 its opposite,
    complete windows come first so that pair_group keeps a target frame's evaluations together."""
    for T, window in ((2, 4), (4, 4), (9, 4), (7, 2), (6, 3)):
        idx = tcl.window_evaluations(T, window)
        prev, cur, bfi, ffi = (idx[k].tolist() for k in ("prev_index", "cur_index", "bf_index", "ff_index"))
        ft, fs = idx["field_t"].tolist(), idx["field_s"].tolist()
        want = {(s, t) for t in range(1, T) for s in range(max(0, t - window + 1), t)}
        assert {(min(p, c), max(p, c)) for p, c in zip(prev, cur)} == want
        assert len(prev) == 2 * len(want) == 2 * len(ft) and idx["group"] == 2 * (window - 1)
        assert sorted(bfi) == sorted(ffi) == list(range(2 * len(ft)))      # every field is the bf of one evaluation and the ff of one
        for p, c, b, f in zip(prev, cur, bfi, ffi):
            j = b // 2
            assert f == (b ^ 1) and {p, c} == {ft[j], fs[j]}
            # the bf of "warp prev into cur" is sampled on cur's grid: field 2j = flow(t -> s) serves cur = t, field 2j+1 cur = s
            assert c == (ft[j] if b % 2 == 0 else fs[j])
        n_full, G = idx["n_complete"], idx["group"]
        assert n_full % G == 0 and n_full == G * max(0, T - (window - 1))
        for g in range(n_full // G):                                          # one complete window = one target frame
            grp = [max(p, c) for p, c in zip(prev[g * G:(g + 1) * G], cur[g * G:(g + 1) * G])]
            assert len(set(grp)) == 1


def _band_worker(rank, world, port, H, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tcl_b200 as tcl
    rows = torch.from_numpy(_fake_pair_results(H, seed=3))          # a per-row sum of squares stands in for the kernel's band sum
    r0, r1 = tcl.band_rows(H, world, rank)
    sums = tcl.allreduce_sums(rows[r0:r1].sum().reshape(1).clone())
    torch.save(sums, os.path.join(out_dir, f"band{rank}.pt"))
    dist.destroy_process_group()


def test_band_sums_add_up_over_gloo(tcl, tmp_path):
    """The banded evaluation's exchange step: every rank contributes the fp64 sum of its band, one all-reduce, every rank
    holds the whole frame's sum (the GPU suite checks the kernel's band sums themselves: test_gpu_parity / test_gpu_multirank)."""
    H, world = 436, 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_band_worker, args=(world, port, H, str(tmp_path)), nprocs=world, join=True)
    rows = _fake_pair_results(H, seed=3)
    r0, r1 = tcl.band_rows(H, world, 0)
    want = rows[r0:r1].sum() + rows[r1:].sum()
    for r in range(world):
        assert float(torch.load(os.path.join(tmp_path, f"band{r}.pt"))[0]) == want
