"""Device input pipeline (data.DeviceBatcher over rlg_batch_prepare) against the restated reference functions
(oracle.ref_port_create_incomplete / _augment / _normalize / _pad, pinned bit for bit on utils/dataset.py and
utils/data_utils.py in tests/test_oracle.py) fed the same random draws."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _expected(cache, plan):
    B = len(plan["method"])
    comp, inc = [], []
    for b in range(B):
        raw = cache[plan["item"][b]].astype(np.float64)        # the reference works on the loaded float64 array
        if plan["method"][b] == 0:
            draws = {"method": 0, "keep_idx": plan["keep_idx"][b, :plan["n_keep"][b]]}
        else:
            draws = {"method": 1, "center": int(plan["center"][b]), "ratio": float(plan["ratio"][b])}
        part = O.ref_port_create_incomplete(raw, draws)
        aug = [None, None]
        for which, pc in enumerate((raw, part)):
            rot = plan["rot"][which, b].reshape(3, 3) if "rot" in plan else None
            noise = plan["jitter"][which, b, :len(pc)] if "jitter" in plan else None
            scale = float(plan["scale"][which, b]) if "scale" in plan else None
            aug[which] = O.ref_port_normalize(O.ref_port_augment(pc, rot, noise, scale))
        comp.append(aug[0])
        inc.append(aug[1])
    m = max(len(p) for p in inc)
    pad = [plan["pad_idx"][b, :m - len(inc[b])] for b in range(B)]
    return np.stack(comp), O.ref_port_pad(inc, pad), np.array([len(p) for p in inc])


@pytest.mark.parametrize("N,B,augment", [(2048, 16, True), (2048, 5, False), (700, 7, True), (33, 4, True)])
def test_batches_match_the_restated_reference(rlg, N, B, augment):
    rng = np.random.default_rng(N + B)
    cache = rlg.build_cache([rng.normal(size=(N + (7 if k % 2 else -5), 3)) * rng.uniform(0.3, 3.0) + rng.normal(size=3)
                             for k in range(B + 3)], num_points=N, seed=1)
    assert cache.shape == (B + 3, N, 3) and cache.dtype == np.float32
    plan = rlg.draw_plan(rng, B, N, items=rng.permutation(B + 3)[:B], augment=augment, host_jitter=True)
    assert set(np.unique(plan["method"])) <= {0, 1}
    batch = rlg.DeviceBatcher(cache, DEV).make_batch(plan)
    want_c, want_i, want_len = _expected(cache, plan)
    assert np.array_equal(batch["lengths"].cpu().numpy(), want_len)            # the kept sets have the reference's sizes
    got_c, got_i = batch["complete_pc"].cpu().numpy(), batch["incomplete_pc"].cpu().numpy()
    assert got_i.shape == want_i.shape
    # normalised coordinates live in the unit ball: fp32 rounding of the rotation / centroid / norm differs in the last bits
    assert np.abs(got_c - want_c).max() <= 2e-6
    assert np.abs(got_i - want_i).max() <= 2e-6
    # padding rows repeat kept rows bit for bit
    for b in range(B):
        n = int(want_len[b])
        for s in range(n, got_i.shape[1]):
            assert np.array_equal(got_i[b, s], got_i[b, int(plan["pad_idx"][b, s - n]) % n])


def test_spatial_removal_keeps_exactly_the_reference_points(rlg):
    """Every cloud through the percentile branch, duplicated points included (ties at the radius): the kept index sets must be
    the reference's, which shows in the un-augmented, un-normalised ORDER of the kept points."""
    N, B = 512, 24
    rng = np.random.default_rng(3)
    base = rng.normal(size=(B, N, 3)).astype(np.float32)
    base[:, N // 2:] = base[:, : N // 2]                                        # every point twice: distance ties
    plan = rlg.draw_plan(rng, B, N, augment=False)
    plan["method"][:] = 1
    plan["center"] = rng.integers(0, N, B).astype(np.int32)
    for b in range(B):
        k, g = O.percentile_parts(N, float(plan["ratio"][b]))
        plan["q_index"][b], plan["q_gamma"][b] = k, g
    batch = rlg.DeviceBatcher(base, DEV).make_batch(plan)
    want_c, want_i, want_len = _expected(base, plan)
    assert np.array_equal(batch["lengths"].cpu().numpy(), want_len)
    assert np.abs(batch["incomplete_pc"].cpu().numpy() - want_i).max() <= 2e-6


def test_device_drawn_jitter_is_clipped_noise_on_the_flagged_clouds(rlg):
    """Without host noise in the plan the batcher draws the jitter on the device: clouds whose flag is off come out exactly as
    without augmentation noise, flagged ones differ by at most the clip (before normalisation the noise is <= 0.05)."""
    N, B = 256, 6
    rng = np.random.default_rng(11)
    cache = rng.normal(size=(B, N, 3)).astype(np.float32)
    plan = rlg.draw_plan(rng, B, N)
    plan["rot"][:] = np.eye(3, dtype=np.float32).reshape(9)
    plan["scale"][:] = 1.0
    plan["jitter_on"][0] = [True, False, True, False, True, False]
    plan["jitter_on"][1] = False
    noisy = rlg.DeviceBatcher(cache, DEV).make_batch(plan)["complete_pc"].cpu().numpy()
    quiet_plan = dict(plan)
    quiet_plan["jitter_on"] = np.zeros((2, B), np.bool_)
    quiet = rlg.DeviceBatcher(cache, DEV).make_batch(quiet_plan)["complete_pc"].cpu().numpy()
    for b in range(B):
        if plan["jitter_on"][0, b]:
            assert 0 < np.abs(noisy[b] - quiet[b]).max() < 0.2
        else:
            assert np.array_equal(noisy[b], quiet[b])


def test_device_drawn_subsets_and_padding(rlg):
    """host_indices=False: the random subsets' permutations and the padding indices are drawn on the device.  Read back, they
    are permutations / valid indices, and the batch equals the restated reference fed with exactly those draws."""
    N, B = 512, 8
    rng = np.random.default_rng(5)
    cache = rng.normal(size=(B, N, 3)).astype(np.float32)
    plan = rlg.draw_plan(rng, B, N, host_jitter=True, host_indices=False)
    plan["method"][::2] = 0                                             # both branches in the batch
    plan["n_keep"] = np.array([int(N * (1 - r)) for r in plan["ratio"]], np.int32)
    assert plan["keep_idx"] is None and plan["pad_idx"] is None
    batch = rlg.DeviceBatcher(cache, DEV).make_batch(plan)
    keep_idx = batch["draws"]["keep_idx"].cpu().numpy()
    pad_idx = batch["draws"]["pad_idx"].cpu().numpy()
    assert all(np.array_equal(np.sort(keep_idx[b]), np.arange(N)) for b in range(B)) and pad_idx.min() >= 0
    full = dict(plan, keep_idx=keep_idx, pad_idx=pad_idx)
    want_c, want_i, want_len = _expected(cache, full)
    assert np.array_equal(batch["lengths"].cpu().numpy(), want_len)
    assert np.abs(batch["complete_pc"].cpu().numpy() - want_c).max() <= 2e-6
    assert np.abs(batch["incomplete_pc"].cpu().numpy() - want_i).max() <= 2e-6
