"""Property tests (hypothesis) of the Chamfer forward on the GPU: whatever the shapes, offsets, scales, duplicated points or
near-ties, every sweep kernel returns the direct-form oracle's bits (distances and arg-min under the reference's tie rule,
utils/losses.py:29-33), and the backward matches the float64 closed form."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SWEEPS = [("fp32", False), ("tensor", False), ("tensor", True)]


def _cloud(rng, b, n, kind, offset, scale, dup_frac, lattice):
    if kind == "sphere":
        x = rng.standard_normal((b, n, 3))
        x /= np.maximum(np.linalg.norm(x, axis=2, keepdims=True), 1e-12)
    elif kind == "uniform":
        x = rng.uniform(-1.0, 1.0, (b, n, 3))
    else:                                   # clustered: a few tight blobs -> many near-ties inside a candidate group
        centres = rng.uniform(-1.0, 1.0, (b, 4, 3))
        x = centres[:, rng.integers(0, 4, n)] + 1e-3 * rng.standard_normal((b, n, 3))
    if lattice:                             # coordinates on a coarse grid: exact ties and sqrt collisions
        x = np.round(x * lattice) / lattice
    x = x * scale + offset
    if dup_frac > 0 and n > 1:              # the dataset pads clouds by duplicating random points (utils/dataset.py:398-421)
        k = max(1, int(dup_frac * n))
        dst = rng.integers(0, n, k)
        src = rng.integers(0, n, k)
        x[:, dst] = x[:, src]
    return torch.from_numpy(x.astype(np.float32))


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(seed=st.integers(0, 2 ** 31 - 1), b=st.integers(1, 5), n=st.integers(1, 700), m=st.integers(1, 700),
       kind=st.sampled_from(["sphere", "uniform", "clustered"]), offset=st.sampled_from([0.0, 0.5, 30.0, 1e3, 1e4]),
       scale=st.sampled_from([1e-3, 1.0, 7.0, 1e4]), dup=st.sampled_from([0.0, 0.0, 0.25, 0.9]),
       lattice=st.sampled_from([0, 0, 4, 64]))
def test_forward_bits_equal_the_direct_oracle(rlg, seed, b, n, m, kind, offset, scale, dup, lattice):
    rng = np.random.default_rng(seed)
    pc1 = _cloud(rng, b, n, kind, offset, scale, dup, lattice)
    pc2 = _cloud(rng, b, m, kind, offset, scale, dup, lattice)
    o1, o2, j1, j2 = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
    old = rlg.get_default_sweep()
    try:
        for sweep, two in SWEEPS:
            rlg.set_default_sweep(sweep)
            kw = {"track_two": True} if two else {}
            d1, d2, i1, i2, m1, m2 = rlg.chamfer_nearest(pc1.to(DEV), pc2.to(DEV), **kw)
            assert np.array_equal(d1.cpu().numpy(), o1) and np.array_equal(d2.cpu().numpy(), o2), (sweep, two)
            assert np.array_equal(i1.cpu().numpy(), j1) and np.array_equal(i2.cpu().numpy(), j2), (sweep, two)
    finally:
        rlg.set_default_sweep(old)


@settings(max_examples=15, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(seed=st.integers(0, 2 ** 31 - 1), b=st.integers(1, 4), n=st.integers(2, 400), m=st.integers(2, 400),
       kind=st.sampled_from(["sphere", "uniform"]), offset=st.sampled_from([0.0, 2.0]), bidir=st.booleans())
def test_backward_matches_the_float64_closed_form(rlg, seed, b, n, m, kind, offset, bidir):
    rng = np.random.default_rng(seed)
    pc1 = _cloud(rng, b, n, kind, offset, 1.0, 0.0, 0)
    pc2 = _cloud(rng, b, m, kind, offset, 1.0, 0.0, 0)
    a = pc1.to(DEV).requires_grad_(True)
    c = pc2.to(DEV).requires_grad_(True)
    rlg.ChamferLoss(bidirectional=bidir)(a, c).backward()
    d1, d2, i1, i2 = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
    g1 = np.full(b, (0.5 if bidir else 1.0) / b)
    g2 = np.full(b, (0.5 if bidir else 0.0) / b)
    t1, t2 = O.chamfer_bwd_truth(pc1, pc2, d1, d2, i1, i2, g1, g2)
    assert O.rowwise_rel_err(a.grad.cpu().numpy(), t1) <= 1e-5
    assert O.rowwise_rel_err(c.grad.cpu().numpy(), t2) <= 1e-5
