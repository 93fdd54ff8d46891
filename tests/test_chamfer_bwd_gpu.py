"""GPU tests of the two Chamfer backward kernels beyond tests/test_chamfer_gpu.py.

The run-to-run reproducible variant (rlg_chamfer_bwd_det / rlg_chamfer_loss_bwd_det: partner terms summed in 64-bit fixed
point with integer atomics): the same gradient as the float-atomics kernel to the same 1e-5 row-wise tolerance against the
float64 closed form, and bit-identical results from call to call where the float-atomics kernel is free to differ (many
queries sharing one partner).  The float-atomics kernel's vector reductions: every alignment of the gradient buffers, cloud
sizes that put rows across 16-byte windows and warps across cloud boundaries, nothing written outside the buffers.
Nothing here reads /root/reference."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5       # north_star's row-wise gradient tolerance


@pytest.fixture(autouse=True)
def det_switch(rlg):
    rlg.set_deterministic_backward(True)
    yield
    rlg.set_deterministic_backward(None)


def _collision_heavy(B, N, M, seed):
    """pc2 is a tight cluster away from pc1: the N queries of a pair share a handful of partners, so their terms collide
    on the same gradient rows (hundreds of atomics per row)."""
    pc1 = O.make_clouds(B, N, "sphere", seed)
    pc2 = 1e-3 * O.make_clouds(B, M, "uniform", seed + 1) + torch.tensor([2.0, 0.5, -1.0])
    return pc1, pc2


@pytest.mark.parametrize("B,N,M", [(1, 1, 1), (2, 16, 25), (3, 300, 257), (2, 2048, 2048), (2, 2048, 1400), (1, 9000, 8200)])
def test_deterministic_backward_vs_float64_truth(rlg, B, N, M):
    pc1 = O.make_clouds(B, N, "sphere", 21)
    pc2 = O.pad_with_duplicates(O.make_clouds(B, M, "sphere", 22), 0.25, 23) if M > 100 else O.make_clouds(B, M, "sphere", 22)
    a = pc1.to(DEV).requires_grad_(True)
    b = pc2.to(DEV).requires_grad_(True)
    rlg.ChamferLoss()(a, b).backward()
    d1, d2, i1, i2 = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
    up = np.full((B,), 0.5 / B, np.float32)
    ga, gb = O.chamfer_bwd_truth(pc1, pc2, d1, d2, i1, i2, up, up)
    assert O.rowwise_rel_err(a.grad.cpu().numpy(), ga) < TOL
    assert O.rowwise_rel_err(b.grad.cpu().numpy(), gb) < TOL


def test_unidirectional_loss_and_per_pair_upstream(rlg):
    B, N, M = 4, 200, 150
    pc1, pc2 = O.make_clouds(B, N, "uniform", 31), O.make_clouds(B, M, "uniform", 32)
    d1, d2, i1, i2 = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
    a = pc1.to(DEV).requires_grad_(True)
    b = pc2.to(DEV).requires_grad_(True)
    rlg.ChamferLoss(bidirectional=False)(a, b).backward()          # rlg_chamfer_loss_bwd_det with a null second upstream
    up = np.full((B,), 1.0 / B, np.float32)
    ga, gb = O.chamfer_bwd_truth(pc1, pc2, d1, d2, i1, i2, up, np.zeros(B, np.float32))
    assert O.rowwise_rel_err(a.grad.cpu().numpy(), ga) < TOL
    assert O.rowwise_rel_err(b.grad.cpu().numpy(), gb) < TOL
    w = torch.tensor([0.5, -1.0, 2.0, 0.0])
    a.grad = b.grad = None
    (rlg.chamfer_distance(a, b) * w.to(DEV)).sum().backward()      # rlg_chamfer_bwd_det, per-pair upstreams
    ga, gb = O.chamfer_bwd_truth(pc1, pc2, d1, d2, i1, i2, (w / 2).numpy(), (w / 2).numpy())
    assert O.rowwise_rel_err(a.grad.cpu().numpy(), ga) < TOL
    assert O.rowwise_rel_err(b.grad.cpu().numpy(), gb) < TOL
    assert not a.grad[3].any() and not b.grad[3].any()


def test_bit_reproducible_where_terms_collide_and_over_the_exponent_range(rlg):
    """Upstream weights from 1e-30 to 1e20 (the quantum follows the weight's exponent), each pair judged on its own."""
    B, N, M = 6, 2048, 64
    pc1, pc2 = _collision_heavy(B, N, M, 51)
    g1 = torch.tensor([1.0, -3e-30, 7e19, 1e-3, 0.0, 0.37])
    g2 = torch.tensor([0.5, 2e-30, -1e20, 0.0, 4.0, 1.0])
    a, b = pc1.to(DEV), pc2.to(DEV)
    d1, d2, i1, i2, _, _ = rlg.chamfer_nearest(a, b)
    runs = []
    for _ in range(6):
        ga, gb = rlg.chamfer_backward(a, b, d1, d2, i1, i2, g1.to(DEV), g2.to(DEV), deterministic=True)
        torch.cuda.synchronize()
        runs.append((ga.clone(), gb.clone()))
        torch.empty(64 << 20, dtype=torch.uint8, device=DEV).zero_()        # perturb cache state / timing between runs
    for ga, gb in runs[1:]:
        assert torch.equal(ga, runs[0][0]) and torch.equal(gb, runs[0][1]), "deterministic backward differs run to run"
    o1, o2, j1, j2 = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
    ta, tb = O.chamfer_bwd_truth(pc1, pc2, o1, o2, j1, j2, g1.numpy(), g2.numpy())
    counts = np.bincount(j1[0], minlength=M)
    assert counts.max() >= 100, "the case is meant to pile many terms onto one row"
    for p in range(B):
        for got, want in ((runs[0][0][p], ta[p]), (runs[0][1][p], tb[p])):
            got = got.cpu().numpy()
            if not np.any(want):
                assert not got.any()
            else:
                assert O.rowwise_rel_err(got, want) < TOL, p
    # and the float-atomics kernel agrees to the same tolerance (it is allowed, not required, to differ in the last bits)
    fa, fb = rlg.chamfer_backward(a, b, d1, d2, i1, i2, g1.to(DEV), g2.to(DEV), deterministic=False)
    for p in (0, 3, 5):
        assert O.rowwise_rel_err(fa[p].cpu().numpy(), ta[p]) < TOL and O.rowwise_rel_err(fb[p].cpu().numpy(), tb[p]) < TOL


def test_accumulate_adds_the_finished_rows(rlg):
    B, N, M = 2, 500, 300
    pc1, pc2 = _collision_heavy(B, N, M, 61)
    a, b = pc1.to(DEV), pc2.to(DEV)
    d1, d2, i1, i2, _, _ = rlg.chamfer_nearest(a, b)
    g = torch.full((B,), 0.25, device=DEV)
    fresh = rlg.chamfer_backward(a, b, d1, d2, i1, i2, g, g, deterministic=True)
    base = (torch.randn_like(a), torch.randn_like(b))
    out = (base[0].clone(), base[1].clone())
    rlg.chamfer_backward(a, b, d1, d2, i1, i2, g, g, out=out, accumulate=True, deterministic=True)
    assert torch.equal(out[0], base[0] + fresh[0]) and torch.equal(out[1], base[1] + fresh[1])
    # out= without accumulate overwrites whatever the buffers held (no zero-fill needed)
    out2 = (torch.full_like(a, float("nan")), torch.full_like(b, float("nan")))
    rlg.chamfer_backward(a, b, d1, d2, i1, i2, g, g, out=out2, deterministic=True)
    assert torch.equal(out2[0], fresh[0]) and torch.equal(out2[1], fresh[1])


def test_follows_torch_deterministic_switch(rlg):
    rlg.set_deterministic_backward(None)
    assert not rlg.deterministic_backward()
    torch.use_deterministic_algorithms(True)
    try:
        assert rlg.deterministic_backward()
        pc1, pc2 = _collision_heavy(4, 2048, 32, 71)
        grads = []
        for _ in range(4):
            a = pc1.to(DEV).requires_grad_(True)
            b = pc2.to(DEV).requires_grad_(True)
            rlg.ChamferLoss()(a, b).backward()
            grads.append((a.grad.clone(), b.grad.clone()))
        for ga, gb in grads[1:]:
            assert torch.equal(ga, grads[0][0]) and torch.equal(gb, grads[0][1])
    finally:
        torch.use_deterministic_algorithms(False)
    rlg.set_deterministic_backward(False)
    assert not rlg.deterministic_backward()


def test_nonfinite_upstream_poisons_the_partner_cloud_of_that_pair_only(rlg):
    B, N, M = 2, 64, 48
    pc1, pc2 = O.make_clouds(B, N, "sphere", 81), O.make_clouds(B, M, "sphere", 82)
    a, b = pc1.to(DEV), pc2.to(DEV)
    d1, d2, i1, i2, _, _ = rlg.chamfer_nearest(a, b)
    g1 = torch.tensor([float("inf"), 1.0], device=DEV)
    g2 = torch.tensor([1.0, 1.0], device=DEV)
    ga, gb = rlg.chamfer_backward(a, b, d1, d2, i1, i2, g1, g2, deterministic=True)
    assert torch.isnan(gb[0]).all()                      # rows of pc2 receive the inf-weighted partner terms
    assert not torch.isfinite(ga[0]).any()               # own terms: inf * direction
    assert torch.isfinite(ga[1]).all() and torch.isfinite(gb[1]).all()


def test_captured_step_replays_to_the_same_bits(rlg):
    B, N, M = 4, 1024, 96
    pc1, pc2 = _collision_heavy(B, N, M, 91)
    a = pc1.to(DEV).requires_grad_(True)
    b = pc2.to(DEV).requires_grad_(True)
    rlg.ChamferLoss()(a, b).backward()
    eager = (a.grad.clone(), b.grad.clone())
    a.grad = b.grad = None
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):                       # warm the per-stream workspace and create .grad outside the capture
            rlg.ChamferLoss()(a, b).backward()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            a.grad.zero_(); b.grad.zero_()
            rlg.ChamferLoss()(a, b).backward()
    for _ in range(3):
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(a.grad, eager[0]) and torch.equal(b.grad, eager[1])


@pytest.mark.parametrize("deterministic", [False, True])
@pytest.mark.parametrize("offset,in_offset", [(0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (0, 1), (0, 2), (2, 3)])
@pytest.mark.parametrize("B,N,M", [(1, 5, 1), (3, 7, 3), (2, 64, 96), (3, 301, 257), (2, 2048, 1400)])
def test_any_buffer_alignment_and_nothing_outside_the_buffers(rlg, B, N, M, offset, in_offset, deterministic):
    """Gradient buffers `offset` floats into an allocation (16-, 8- and 4-byte aligned starts select the 4-float, 2-float and
    scalar accesses), guard floats either side.  (1, 5, 1) and (3, 7, 3): the last row starts a 16-byte window."""
    pc1, pc2 = O.make_clouds(B, N, "sphere", 101), O.make_clouds(B, M, "sphere", 102)
    # the clouds `in_offset` floats into their allocations (the 4-float loads need 16-byte aligned clouds as well)
    a, b = (torch.cat([torch.zeros(in_offset), c.reshape(-1)]).to(DEV)[in_offset:].view(c.shape) for c in (pc1, pc2))
    assert a.is_contiguous() and a.data_ptr() % 16 == (4 * in_offset) % 16
    d1, d2, i1, i2, _, _ = rlg.chamfer_nearest(a, b)
    g1 = torch.linspace(0.5, 1.5, B, device=DEV)
    g2 = torch.linspace(-1.0, 2.0, B, device=DEV)
    GUARD = 8
    flats = [torch.full((GUARD + offset + 3 * B * n + GUARD,), 7.0, device=DEV) for n in (N, M)]
    outs = tuple(f[GUARD + offset: GUARD + offset + 3 * B * n].view(B, n, 3) for f, n in zip(flats, (N, M)))
    assert all(o.data_ptr() % 16 == (4 * offset) % 16 for o in outs)
    rlg.chamfer_backward(a, b, d1, d2, i1, i2, g1, g2, out=outs, deterministic=deterministic)
    torch.cuda.synchronize()
    for f, n in zip(flats, (N, M)):
        assert (f[:GUARD + offset] == 7.0).all() and (f[GUARD + offset + 3 * B * n:] == 7.0).all(), "wrote outside the buffer"
    o1, o2, j1, j2 = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
    ta, tb = O.chamfer_bwd_truth(pc1, pc2, o1, o2, j1, j2, g1.cpu().numpy(), g2.cpu().numpy())
    assert O.rowwise_rel_err(outs[0].cpu().numpy(), ta) < TOL
    assert O.rowwise_rel_err(outs[1].cpu().numpy(), tb) < TOL
    # accumulate onto what the buffers hold: same rows on top of the first result
    first = (outs[0].clone(), outs[1].clone())
    rlg.chamfer_backward(a, b, d1, d2, i1, i2, g1, g2, out=outs, accumulate=True, deterministic=deterministic)
    assert O.rowwise_rel_err(outs[0].cpu().numpy(), 2 * ta) < TOL and O.rowwise_rel_err(outs[1].cpu().numpy(), 2 * tb) < TOL
    for f, n in zip(flats, (N, M)):
        assert (f[:GUARD + offset] == 7.0).all() and (f[GUARD + offset + 3 * B * n:] == 7.0).all()
    del first


@pytest.mark.parametrize("B,N,M,heavy", [(1, 1, 1, False), (3, 33, 25, False), (2, 300, 257, False), (6, 2048, 64, True),
                                         (2, 2048, 1400, False), (1, 9000, 8200, False)])
def test_reproducible_backward_is_bit_equal_to_its_restatement(rlg, B, N, M, heavy):
    """The fixed-point backward is defined operation for operation (oracle.chamfer_bwd_fixed_point: fp32 terms, integer
    quanta, integer sums, one rounding): the kernels must return exactly those bits, for per-pair upstream weights over the
    exponent range and for the fused ChamferLoss scaling."""
    pc1, pc2 = _collision_heavy(B, N, M, 131) if heavy else (O.make_clouds(B, N, "sphere", 131), O.make_clouds(B, M, "uniform", 132))
    a, b = pc1.to(DEV), pc2.to(DEV)
    d1, d2, i1, i2, _, _ = rlg.chamfer_nearest(a, b)
    weights = torch.tensor([1.0, -3e-30, 7e19, 1e-3, 0.0, 0.37])
    g1 = weights[torch.arange(B) % 6]
    g2 = weights[(torch.arange(B) + 2) % 6]
    ga, gb = rlg.chamfer_backward(a, b, d1, d2, i1, i2, g1.to(DEV), g2.to(DEV), deterministic=True)
    host = [t.cpu().numpy() for t in (d1, d2, i1, i2)]
    wa, wb = O.chamfer_bwd_fixed_point(pc1, pc2, *host, g1.numpy(), g2.numpy())
    assert np.array_equal(ga.cpu().numpy(), wa) and np.array_equal(gb.cpu().numpy(), wb)
    # through autograd: ChamferLoss scales the scalar upstream by 0.5 / B inside the kernels
    x = a.clone().requires_grad_(True)
    y = b.clone().requires_grad_(True)
    rlg.ChamferLoss()(x, y).backward()
    ones = np.ones(B, np.float32)
    wa, wb = O.chamfer_bwd_fixed_point(pc1, pc2, *host, ones, ones, 0.5 / B, 0.5 / B)
    assert np.array_equal(x.grad.cpu().numpy(), wa) and np.array_equal(y.grad.cpu().numpy(), wb)
