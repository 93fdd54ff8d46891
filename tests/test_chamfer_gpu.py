"""GPU parity tests of the Chamfer path, through the C ABI (librlg_b200.so), against the oracle and the golden
fixtures.  Bit-exact for distances and indices (vs the direct-form oracle); tolerances written where floating
point sums are involved.  Nothing here reads /root/reference."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


_TRACK_TWO = False


@pytest.fixture(autouse=True, params=["fp32", "tensor", "tensor_top3"])
def sweep(request, rlg):
    """Every test of this module runs once per pair-sweep kernel: the FP32-pipe filter (chamfer_filter.cu) and the
    tensor-core sweep with the fused refinement (chamfer_tcsweep.cu), the latter with and without the
    runner-up-group / third-value report the library otherwise enables by cloud size.  All must give the same bits."""
    global _TRACK_TWO
    old = rlg.get_default_sweep()
    rlg.set_default_sweep("fp32" if request.param == "fp32" else "tensor")
    _TRACK_TWO = request.param == "tensor_top3"
    yield request.param
    _TRACK_TWO = False
    rlg.set_default_sweep(old)


def _run(rlg, pc1, pc2, simple=False, **kw):
    if _TRACK_TWO and not simple:
        kw.setdefault("track_two", True)
    d1, d2, i1, i2, m1, m2 = rlg.chamfer_nearest(pc1.to(DEV), pc2.to(DEV), simple=simple, **kw)
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in (d1, d2, i1, i2, m1, m2)]


def _check_against_direct(rlg, pc1, pc2, simple=False, **kw):
    """Bit-equal to the direct-form oracle under the reference's tie rule (argmin over the sqrt-ed distances)."""
    d1, d2, i1, i2, m1, m2 = _run(rlg, pc1, pc2, simple, **kw)
    o1, o2, j1, j2 = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
    assert np.array_equal(d1, o1) and np.array_equal(d2, o2), "distances not bit-equal to the direct oracle"
    assert np.array_equal(i1, j1) and np.array_equal(i2, j2), "argmin not equal to the direct oracle"
    w1, w2 = O.chamfer_means(o1, o2)
    np.testing.assert_allclose(m1, w1, rtol=2e-7)
    np.testing.assert_allclose(m2, w2, rtol=2e-7)
    return d1, d2, i1, i2, m1, m2


SHAPES = [(1, 1, 1), (3, 1, 7), (2, 25, 1), (4, 16, 25), (2, 31, 33), (2, 32, 32), (3, 255, 257), (2, 256, 64),
          (1, 300, 1000), (2, 1400, 2048), (2, 2048, 2048), (1, 2049, 513), (5, 700, 90),
          # many clouds per CTA of the tensor sweep (several segments, ragged last query block / candidate tile)
          (300, 130, 70), (50, 257, 129), (7, 1030, 2050), (161, 64, 64)]


@pytest.mark.parametrize("B,N,M", SHAPES)
@pytest.mark.parametrize("kind", ["sphere", "uniform"])
def test_forward_bit_exact_vs_direct_oracle(rlg, B, N, M, kind):
    pc1 = O.make_clouds(B, N, kind, 1000 + N)
    pc2 = O.make_clouds(B, M, kind, 2000 + M)
    _check_against_direct(rlg, pc1, pc2)


@pytest.mark.parametrize("B,N,M", [(2, 31, 33), (2, 700, 2048), (1, 2048, 1400)])
def test_simple_kernel_agrees_with_tile_kernel(rlg, B, N, M):
    pc1, pc2 = O.make_clouds(B, N, "sphere", 5), O.make_clouds(B, M, "sphere", 6)
    a = _check_against_direct(rlg, pc1, pc2, simple=True)
    b = _check_against_direct(rlg, pc1, pc2, simple=False)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("offset,scale", [(0.0, 1e-3), (3.0, 1.0), (100.0, 1.0), (1e4, 10.0), (0.0, 1e4)])
def test_filter_margin_keeps_exactness_far_from_the_origin(rlg, offset, scale):
    """The filter |y|^2 - 2x.y loses precision when |p| >> nearest-neighbour distance; the rounding margin must then
    send (almost) every point through the exact scan.  Outputs stay bit-equal to the direct oracle."""
    pc1 = O.make_clouds(2, 700, "sphere", 91) * scale + offset
    pc2 = O.make_clouds(2, 900, "uniform", 92) * scale + offset
    _check_against_direct(rlg, pc1.contiguous(), pc2.contiguous())


def test_near_ties_across_candidate_groups(rlg):
    """Candidates at almost the same distance in different 32-column groups / row blocks: the filter cannot order
    them, the exact refinement must.  pc2 = jittered copies of a few anchor points spread over the whole cloud."""
    g = torch.Generator().manual_seed(5)
    anchors = O.make_clouds(2, 16, "sphere", 93)
    idx = torch.randint(0, 16, (2, 2048), generator=g)
    pc2 = torch.gather(anchors, 1, idx.unsqueeze(-1).expand(-1, -1, 3)) * (1 + 3e-7 * torch.randn(2, 2048, 1, generator=g))
    pc1 = O.make_clouds(2, 1500, "sphere", 94)
    _check_against_direct(rlg, pc1, pc2.contiguous())
    _check_against_direct(rlg, pc2.contiguous(), pc1)


def test_workspace_returns_to_clean_state(rlg):
    """Repeat calls reuse the cached workspace without a memset (RLG_CHAMFER_WS_CLEAN): results must not depend on
    what the previous call (other inputs, same shape) left behind."""
    a1, b1 = O.make_clouds(3, 1000, "sphere", 95), O.make_clouds(3, 777, "sphere", 96)
    a2, b2 = O.make_clouds(3, 1000, "uniform", 97) * 0.01, O.make_clouds(3, 777, "uniform", 98) * 0.01
    first = _run(rlg, a1, b1)
    _run(rlg, a2, b2)                      # much smaller distances: stale keys would win every atomicMin
    again = _run(rlg, a1, b1)
    assert all(np.array_equal(x, y) for x, y in zip(first, again))
    _check_against_direct(rlg, a2, b2)


def test_exact_ties_pick_lowest_index(rlg):
    """Duplicate padding of ragged partial clouds (utils/dataset.py:398-421) creates exact ties."""
    pc1 = O.make_clouds(3, 600, "sphere", 11)
    pc2 = O.pad_with_duplicates(O.make_clouds(3, 2048, "sphere", 12), 0.25, 13)
    d1, d2, i1, i2, *_ = _check_against_direct(rlg, pc1, pc2)
    t1 = O.chamfer_torch_direct(pc1, pc2)
    assert np.array_equal(i1, t1[2].numpy()) and np.array_equal(i2, t1[3].numpy())
    # identical clouds: zero distances, identity argmin
    same = O.make_clouds(2, 500, "uniform", 14)
    d1, d2, i1, i2, m1, m2 = _check_against_direct(rlg, same, same.clone())
    assert not d1.any() and not d2.any() and np.array_equal(i1[0], np.arange(500))
    # all points equal: every candidate ties -> index 0 everywhere
    ones = torch.ones(1, 70, 3)
    _, _, i1, i2, *_ = _check_against_direct(rlg, ones, ones.clone())
    assert not i1.any() and not i2.any()


def test_golden_small_cases_match_reference_bitwise(rlg, golden_chamfer):
    """N,M <= 25: the reference itself runs direct-mode cdist; distances bit-equal, means to 1 ulp-ish."""
    g = golden_chamfer
    for k in range(int(g["small_count"])):
        pc1, pc2 = torch.from_numpy(g[f"small{k}_pc1"]), torch.from_numpy(g[f"small{k}_pc2"])
        d1, d2, i1, i2, m1, m2 = _run(rlg, pc1, pc2)
        assert np.array_equal(d1, g[f"small{k}_d1"]) and np.array_equal(d2, g[f"small{k}_d2"])
        assert O.idx_mismatches_are_exact_ties(pc1, pc2, i1, g[f"small{k}_i1"])[1] == 0
        assert O.idx_mismatches_are_exact_ties(pc2, pc1, i2, g[f"small{k}_i2"])[1] == 0
        np.testing.assert_allclose(m1, g[f"small{k}_dist1"], rtol=2e-7)
        np.testing.assert_allclose(m2, g[f"small{k}_dist2"], rtol=2e-7)


def test_golden_big_cases_contract_vs_as_written_reference(rlg, golden_chamfer):
    """N or M > 25: the as-written reference is the noisy matmul path; the contract of SURVEY.md 8c."""
    g = golden_chamfer
    for k in range(int(g["big_count"])):
        B, N, M, s1, s2, sd = [int(v) for v in g[f"big{k}_meta"]]
        kind = str(g[f"big{k}_kind"])
        pc1, pc2 = O.make_clouds(B, N, kind, s1), O.make_clouds(B, M, kind, s2)
        if sd >= 0:
            pc2 = O.pad_with_duplicates(pc2, 0.25, sd)
        d1, d2, i1, i2, m1, m2 = _check_against_direct(rlg, pc1, pc2)
        assert O.idx_mismatches_are_near_ties(pc1, pc2, i1, g[f"big{k}_i1"])[1] == 0
        assert O.idx_mismatches_are_near_ties(pc2, pc1, i2, g[f"big{k}_i2"])[1] == 0
        f1, f2, *_ = O.chamfer_f64(pc1, pc2)
        t1, t2 = f1.mean(1).numpy(), f2.mean(1).numpy()
        assert O.rel_err(m1, t1) < 1e-6 and O.rel_err(m2, t2) < 1e-6           # vs float64 truth
        ref_err = max(O.rel_err(g[f"big{k}_dist1"], t1), O.rel_err(g[f"big{k}_dist2"], t2))
        tol = max(1e-5, 2 * ref_err)                                           # vs the reference: its own error
        assert O.rel_err(m1, g[f"big{k}_dist1"]) < tol and O.rel_err(m2, g[f"big{k}_dist2"]) < tol


@pytest.mark.parametrize("B,N,M", [(2, 16, 25), (3, 300, 257), (2, 2048, 2048), (2, 2048, 1400), (1, 9000, 8200)])
def test_backward_vs_float64_truth(rlg, B, N, M):
    pc1 = O.make_clouds(B, N, "sphere", 21)
    pc2 = O.pad_with_duplicates(O.make_clouds(B, M, "sphere", 22), 0.25, 23) if M > 100 else O.make_clouds(B, M, "sphere", 22)
    a = pc1.to(DEV).requires_grad_(True)
    b = pc2.to(DEV).requires_grad_(True)
    loss = rlg.ChamferLoss()(a, b)
    loss.backward()
    d1, d2, i1, i2 = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
    up = np.full((B,), 0.5 / B, np.float32)
    ga, gb = O.chamfer_bwd_truth(pc1, pc2, d1, d2, i1, i2, up, up)
    # tolerance: 1e-5 relative row-wise (north_star), rows measured against max(|row|, 1e-3 * largest row)
    assert O.rowwise_rel_err(a.grad.cpu().numpy(), ga) < 1e-5
    assert O.rowwise_rel_err(b.grad.cpu().numpy(), gb) < 1e-5
    m1, m2 = O.chamfer_means(d1, d2)
    want = float(np.mean((m1.astype(np.float64) + m2) / 2))
    assert abs(loss.item() - want) <= 1e-6 * abs(want)


def test_backward_matches_reference_autograd_on_small_golden(rlg, golden_chamfer):
    g = golden_chamfer
    for k in range(int(g["small_count"])):
        a = torch.from_numpy(g[f"small{k}_pc1"]).to(DEV).requires_grad_(True)
        b = torch.from_numpy(g[f"small{k}_pc2"]).to(DEV).requires_grad_(True)
        loss = rlg.ChamferLoss()(a, b)
        loss.backward()
        assert abs(loss.item() - float(g[f"small{k}_loss"])) <= 1e-6 * abs(float(g[f"small{k}_loss"])) + 1e-9
        # exact ties may route a gradient to a different (equally near) partner: compare only when indices agree
        _, _, i1, i2, *_ = _run(rlg, a.detach().cpu(), b.detach().cpu())
        if np.array_equal(i1, g[f"small{k}_i1"]) and np.array_equal(i2, g[f"small{k}_i2"]):
            assert O.rowwise_rel_err(a.grad.cpu().numpy(), g[f"small{k}_g1"]) < 1e-5
            assert O.rowwise_rel_err(b.grad.cpu().numpy(), g[f"small{k}_g2"]) < 1e-5


def test_per_pair_upstream_and_unidirectional(rlg):
    """chamfer_distance returns (B,) and may be weighted per pair (reward path); bidirectional=False uses
    dist1 only, so pc2 receives gradient only through the argmin partners."""
    B, N, M = 4, 200, 150
    pc1, pc2 = O.make_clouds(B, N, "uniform", 31), O.make_clouds(B, M, "uniform", 32)
    w = torch.tensor([0.5, -1.0, 2.0, 0.0])
    a = pc1.to(DEV).requires_grad_(True)
    b = pc2.to(DEV).requires_grad_(True)
    cd = rlg.chamfer_distance(a, b, bidirectional=False)
    assert cd.shape == (B,)
    (cd * w.to(DEV)).sum().backward()
    d1, d2, i1, i2 = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
    ga, gb = O.chamfer_bwd_truth(pc1, pc2, d1, d2, i1, i2, w.numpy(), np.zeros(B, np.float32))
    assert O.rowwise_rel_err(a.grad.cpu().numpy(), ga) < 1e-5
    assert O.rowwise_rel_err(b.grad.cpu().numpy(), gb) < 1e-5
    assert not a.grad[3].any() and not b.grad[3].any()


def test_no_grad_and_requires_grad_false_inputs(rlg):
    pc1, pc2 = O.make_clouds(2, 100, "sphere", 41).to(DEV), O.make_clouds(2, 90, "sphere", 42).to(DEV)
    with torch.no_grad():
        v = rlg.ChamferLoss()(pc1, pc2)
    assert v.dim() == 0 and not v.requires_grad
    a = pc1.clone().requires_grad_(True)
    rlg.ChamferLoss()(a, pc2).backward()           # target has no grad (the trainer's case, train:236)
    assert a.grad is not None and pc2.grad is None


def test_full_size_properties_cfg2(rlg):
    """B=32, N=M=2048 (BASELINE config 2), size-independent properties: symmetry under swapping the clouds,
    invariance of distances under a permutation of the candidates, idempotence (second call, clean workspace)."""
    pc1, pc2 = O.make_clouds(32, 2048, "sphere", 1236), O.make_clouds(32, 2048, "sphere", 2236)
    a, b = pc1.to(DEV), pc2.to(DEV)
    d1, d2, i1, i2, m1, m2 = rlg.chamfer_nearest(a, b)
    e2, e1, j2, j1, n2, n1 = rlg.chamfer_nearest(b, a)
    assert torch.equal(d1, e1) and torch.equal(d2, e2) and torch.equal(i1, j1) and torch.equal(i2, j2)
    assert torch.equal(m1, n1) and torch.equal(m2, n2)
    perm = torch.randperm(2048, generator=torch.Generator().manual_seed(0)).to(DEV)
    p1, _, pi1, *_ = rlg.chamfer_nearest(a, b[:, perm])
    assert torch.equal(p1, d1) and torch.equal(perm[pi1.long()], i1.long())     # no exact ties in this draw
    again = rlg.chamfer_nearest(a, b)
    assert all(torch.equal(x, y) for x, y in zip(again, (d1, d2, i1, i2, m1, m2)))
    # gathered partner really is at the reported distance
    gathered = torch.gather(b, 1, i1.long().unsqueeze(-1).expand(-1, -1, 3))
    assert torch.allclose((a - gathered).norm(dim=2), d1, rtol=1e-6, atol=1e-7)
    # the whole batch is bit-exact against the C oracle (B=32 x 2048^2 takes it well under a second per direction)
    o1, o2, j1, j2 = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
    assert np.array_equal(d1.cpu().numpy(), o1) and np.array_equal(d2.cpu().numpy(), o2)
    assert np.array_equal(i1.cpu().numpy(), j1) and np.array_equal(i2.cpu().numpy(), j2)


def test_large_cloud_16384(rlg):
    """BASELINE config 5 shape (N=M=16384), one pair against the oracle, plus memory stays O(N+M)."""
    pc1, pc2 = O.make_clouds(1, 16384, "sphere", 51), O.make_clouds(1, 16384, "sphere", 52)
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    _check_against_direct(rlg, pc1, pc2)
    assert torch.cuda.max_memory_allocated() - base < 64 << 20        # the reference needs 1 GiB per (N,M) matrix here


def test_error_paths_raise(rlg):
    with pytest.raises(ValueError):
        rlg.chamfer_distance_l2(torch.zeros(2, 4, 3, device=DEV), torch.zeros(3, 4, 3, device=DEV))
    with pytest.raises(ValueError):
        rlg.chamfer_distance_l2(torch.zeros(2, 0, 3, device=DEV), torch.zeros(2, 4, 3, device=DEV))
    with pytest.raises(ValueError):
        rlg.chamfer_distance_l2(torch.zeros(2, 4, 3, device=DEV, dtype=torch.float64), torch.zeros(2, 4, 3, device=DEV))
    out = rlg.chamfer_nearest(torch.zeros(0, 4, 3, device=DEV), torch.zeros(0, 5, 3, device=DEV))
    assert out[0].shape == (0, 4) and out[1].shape == (0, 5)


def test_cuda_graph_capture(rlg):
    """Every call is enqueued on the caller's stream without sync/alloc inside the library -> capturable."""
    pc1, pc2 = O.make_clouds(4, 512, "sphere", 61).to(DEV), O.make_clouds(4, 512, "sphere", 62).to(DEV)
    want = rlg.chamfer_nearest(pc1, pc2)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        rlg.chamfer_nearest(pc1, pc2)            # warm the per-stream workspace outside capture
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            got = rlg.chamfer_nearest(pc1, pc2)
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    assert all(torch.equal(x, y) for x, y in zip(got, want))


def _filter_values(a, b):
    """Per-query smallest filter value of the tensor sweep (RLG_CHAMFER_FILTER_ONLY diagnostic), both directions."""
    import importlib
    _lib = importlib.import_module("gan-rl_3d_b200._lib")
    lib = _lib.load()
    B, N, _ = a.shape
    M = b.shape[1]
    d1 = torch.empty(B, N, device=DEV); d2 = torch.empty(B, M, device=DEV)
    i1 = torch.empty(B, N, dtype=torch.int32, device=DEV); i2 = torch.empty(B, M, dtype=torch.int32, device=DEV)
    ws = torch.empty(lib.rlg_chamfer_ws_bytes(B, N, M), dtype=torch.uint8, device=DEV).fill_(0xFF)
    flags = _lib.CHAMFER_WS_CLEAN | _lib.CHAMFER_FILTER_ONLY | _lib.CHAMFER_ALGO_TENSOR
    rc = lib.rlg_chamfer_fwd(a.data_ptr(), b.data_ptr(), B, N, M, d1.data_ptr(), d2.data_ptr(), i1.data_ptr(),
                             i2.data_ptr(), None, None, ws.data_ptr(), ws.numel(), flags,
                             torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.rlg_last_error()
    torch.cuda.synchronize()
    keys = ws[: 8 * B * (N + M)].view(torch.int64)
    v1 = (keys[: B * N].view(B, N) >> 32).to(torch.int32).view(torch.float32).double()
    v2 = (keys[B * N:].view(B, M) >> 32).to(torch.int32).view(torch.float32).double()
    return v1, v2


def _filter_error_in_u(a, b):
    """Largest |filter value of the best candidate - float64 truth| in units of u (a^2 + b^2), u = 2^-24, with b^2 the
    largest squared norm of the candidate cloud (an upper bound of the winning candidate's)."""
    v1, v2 = _filter_values(a, b)
    a64, b64 = a.double(), b.double()
    D = ((a64[:, :, None, :] - b64[:, None, :, :]) ** 2).sum(-1)
    na, nb = (a64 ** 2).sum(-1), (b64 ** 2).sum(-1)
    u = 2.0 ** -24
    worst = 0.0
    for val, truth, nq, j in ((v1, D.min(2).values, na, D.argmin(2)), (v2, D.min(1).values, nb, D.argmin(1))):
        nc = torch.gather(nb if nq is na else na, 1, j)          # squared norm of the true nearest candidate
        worst = max(worst, float(((val - truth).abs() / (u * (nq + nc)).clamp_min(1e-300)).max()))
    return worst


def test_tensor_filter_error_is_far_inside_its_margin(rlg, sweep):
    """The tensor-core filter's value for the best candidate vs float64, in units of u (a^2 + b^2): the refinement's
    margin budgets 61 u for it (chamfer_tcsweep.cu header); the measured error is an order of magnitude smaller."""
    if sweep != "tensor":
        pytest.skip("one run is enough")
    B, N, M = 3, 1400, 2048
    worst = 0.0
    for kind, scale, shift in (("sphere", 1.0, 0.0), ("uniform", 1.0, 0.0), ("sphere", 0.1, 2.0), ("uniform", 50.0, 0.0)):
        a = (torch.as_tensor(O.make_clouds(B, N, kind, 11)) * scale + shift).to(DEV)
        b = (torch.as_tensor(O.make_clouds(B, M, kind, 12)) * scale + shift).to(DEV)
        worst = max(worst, _filter_error_in_u(a, b))
    assert worst < 16.0, f"tensor filter error {worst:.1f} u(a^2+b^2): the margin assumes < 61 u"


def test_tensor_filter_error_randomized_search(rlg, sweep):
    """Adversarial search for the largest filter error: random draws over scale (1e-3 .. 1e4), offset from the origin
    (0 .. 1e4, cancellation in |x|^2 + |y|^2 - 2x.y), anisotropy, clustered near-duplicates and coordinates with long
    mantissas; then hill-climbing on the worst draw.  Every evaluation must stay below the 61 u the margin budgets."""
    if sweep != "tensor":
        pytest.skip("one run is enough")
    g = torch.Generator().manual_seed(2024)
    B, N, M = 2, 384, 512

    def draw(scale, offset, aniso, cluster):
        a = torch.randn(B, N, 3, generator=g) * scale * torch.tensor([1.0, aniso, aniso * aniso]) + offset
        if cluster:
            b = a[:, torch.randint(0, N, (M,), generator=g)] * (1 + 2e-7 * torch.randn(B, M, 1, generator=g))
        else:
            b = torch.randn(B, M, 3, generator=g) * scale + offset
        return a.float().contiguous(), b.float().contiguous()

    worst, worst_cfg = 0.0, None
    for _ in range(60):
        cfg = (10.0 ** float(torch.empty(1).uniform_(-3, 4, generator=g)),
               float(torch.empty(1).uniform_(0, 1, generator=g) < 0.5) * 10.0 ** float(torch.empty(1).uniform_(-2, 4, generator=g)),
               10.0 ** float(torch.empty(1).uniform_(-2, 0, generator=g)),
               bool(torch.empty(1).uniform_(0, 1, generator=g) < 0.3))
        a, b = draw(*cfg)
        e = _filter_error_in_u(a.to(DEV), b.to(DEV))
        if e > worst:
            worst, worst_cfg = e, cfg
    # hill-climb around the worst configuration
    scale, offset, aniso, cluster = worst_cfg
    for _ in range(30):
        cand = (scale * 10.0 ** float(torch.empty(1).uniform_(-0.3, 0.3, generator=g)),
                offset * 10.0 ** float(torch.empty(1).uniform_(-0.3, 0.3, generator=g)),
                min(1.0, aniso * 10.0 ** float(torch.empty(1).uniform_(-0.3, 0.3, generator=g))), cluster)
        a, b = draw(*cand)
        e = _filter_error_in_u(a.to(DEV), b.to(DEV))
        if e > worst:
            worst, (scale, offset, aniso, cluster) = e, cand
    assert worst < 61.0, f"tensor filter error {worst:.1f} u(a^2+b^2) at {(scale, offset, aniso, cluster)}"


def _sqrt_collision_pair(seed=0):
    """(r, s): from the origin (r,s,0) has a squared distance ONE ULP ABOVE that of (r,0,0) and the same sqrtf."""
    rng = np.random.default_rng(seed)
    for _ in range(20000):
        r = np.float32(rng.uniform(1.42, 1.99))
        s_ = np.float32(2.0 ** -11 * rng.uniform(0.8, 1.2))
        tb = np.float32(r * r)
        ta = np.float32(np.float64(s_) * np.float64(s_) + np.float64(tb))      # fmaf(s,s,t), exact in float64
        if ta == np.nextafter(tb, np.float32(8), dtype=np.float32) and np.sqrt(ta, dtype=np.float32) == np.sqrt(tb, dtype=np.float32):
            return float(r), float(s_)
    raise AssertionError("no collision found")


@pytest.mark.parametrize("lo_idx,hi_idx,n", [(3, 9, 100), (3, 70, 100), (40, 1900, 2048), (100, 101, 300)])
def test_sqrt_ties_follow_the_reference_rule(rlg, lo_idx, hi_idx, n):
    """two candidates with different squared distances whose square roots collide in fp32 are an exact
    tie for torch.min on the sqrt-ed matrix, and the LOWER index wins even if it has the LARGER squared distance --
    inside one 32-candidate group (fused refinement) and across groups (ambiguous -> tail kernel)."""
    r, s_ = _sqrt_collision_pair()
    c = torch.full((1, n, 3), 5.0)
    c[0, lo_idx] = torch.tensor([r, s_, 0.0])        # larger squared distance, lower index
    c[0, hi_idx] = torch.tensor([r, 0.0, 0.0])
    q = torch.zeros(1, 70, 3)
    q[0, 1:] = O.make_clouds(1, 69, "uniform", 3)[0] * 0.01 + 7.0       # other queries: far away, nearest = a (5,5,5) point
    d1, d2, i1, i2, *_ = _check_against_direct(rlg, q, c.contiguous())
    assert i1[0, 0] == lo_idx
    ref_d, ref_i = torch.min(torch.cdist(q, c, compute_mode="donot_use_mm_for_euclid_dist"), dim=2)
    assert np.array_equal(i1, ref_i.numpy().astype(np.int32)) and np.array_equal(d1, ref_d.numpy())


def test_many_sqrt_collisions_radial_ladder(rlg):
    """Candidates on a fine radial ladder around queries near the origin: many distinct squared distances share a sqrtf."""
    g = torch.Generator().manual_seed(9)
    r = 1.5 + torch.arange(600).float() * 1.2e-7
    dirs = torch.nn.functional.normalize(torch.randn(600, 3, generator=g), dim=1)
    c2 = (dirs * r[:, None])[torch.randperm(600, generator=g)].unsqueeze(0).contiguous()
    q2 = torch.zeros(1, 65, 3)
    q2[0, 1:] = torch.randn(64, 3, generator=g) * 1e-4
    d1, d2, i1, i2, *_ = _check_against_direct(rlg, q2, c2)
    ref_d, ref_i = torch.min(torch.cdist(q2, c2, compute_mode="donot_use_mm_for_euclid_dist"), dim=2)
    assert np.array_equal(i1, ref_i.numpy().astype(np.int32)) and np.array_equal(d1, ref_d.numpy())


def test_workspace_cache_evicts_lru_and_keeps_graph_workspaces(rlg):
    """70 shapes overflow the 64-entry workspace cache; a CUDA graph captured before must still replay correctly
    (its workspace is pinned), and a capture without an eager warm-up on that stream must record its own memset."""
    pc1, pc2 = O.make_clouds(2, 300, "sphere", 71).to(DEV), O.make_clouds(2, 260, "sphere", 72).to(DEV)
    want = rlg.chamfer_nearest(pc1, pc2)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):       # no warm-up on this stream: the memset must be in the graph
            got = rlg.chamfer_nearest(pc1, pc2)
    for n in range(70):
        a = O.make_clouds(1, 40 + n, "uniform", n).to(DEV)
        rlg.chamfer_nearest(a, a)
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    assert all(torch.equal(x, y) for x, y in zip(got, want))


@pytest.mark.parametrize("case", ["identical", "apart", "mixed_scale", "collinear"])
def test_degenerate_geometries(rlg, case):
    """Zero distances everywhere, clouds far from each other (every filter value is large), three decades of scale in
    one cloud, points on a line (many near-ties): same bits as the direct oracle on every sweep."""
    g = torch.Generator().manual_seed(77)
    B, N, M = 2, 700, 900
    if case == "identical":
        pc1 = O.make_clouds(B, N, "sphere", 5)
        pc2 = pc1.clone()
    elif case == "apart":
        pc1 = O.make_clouds(B, N, "uniform", 6) + 5.0
        pc2 = O.make_clouds(B, M, "uniform", 7) - 5.0
    elif case == "mixed_scale":
        s1 = 10.0 ** torch.randint(-2, 2, (B, N, 1), generator=g).float()
        s2 = 10.0 ** torch.randint(-2, 2, (B, M, 1), generator=g).float()
        pc1 = O.make_clouds(B, N, "sphere", 8) * s1
        pc2 = O.make_clouds(B, M, "sphere", 9) * s2
    else:
        t1 = torch.rand(B, N, 1, generator=g)
        t2 = torch.rand(B, M, 1, generator=g)
        d = torch.tensor([0.3, -0.5, 0.8])
        pc1 = t1 * d + 0.1
        pc2 = t2 * d + 0.1
    _check_against_direct(rlg, pc1.contiguous(), pc2.contiguous())


def test_sweeps_alternate_on_one_workspace(rlg):
    """The FP32 and the tensor forward keep different scratch data in the shared workspace; alternating them on the same
    cached workspace (same shape, WS_CLEAN) must not disturb either -- many clouds, so that the regions are long."""
    pc1, pc2 = O.make_clouds(300, 130, "sphere", 81), O.make_clouds(300, 70, "sphere", 82)
    want = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
    w1, w2 = O.chamfer_means(want[0], want[1])
    for tensor in (False, True, False, True, True):
        d1, d2, i1, i2, m1, m2 = _run(rlg, pc1, pc2, tensor=tensor)
        assert np.array_equal(d1, want[0]) and np.array_equal(i2, want[3])
        np.testing.assert_allclose(m1, w1, rtol=2e-7)
        np.testing.assert_allclose(m2, w2, rtol=2e-7)


def test_reserved_sms_change_nothing_but_the_grid(rlg):
    """RLG_CHAMFER_RESERVE_SMS (set by distributed.init_from_env for multi-GPU runs): a smaller persistent grid, same bits."""
    pc1, pc2 = O.make_clouds(9, 700, "sphere", 31), O.make_clouds(9, 1300, "uniform", 32)
    want = _run(rlg, pc1, pc2)
    try:
        for n in (1, 7, 147, 200):
            rlg.set_reserved_sms(n)
            got = _run(rlg, pc1, pc2)
            assert all(np.array_equal(g, w) for g, w in zip(got, want)), n
    finally:
        rlg.set_reserved_sms(0)
    with pytest.raises(ValueError):
        rlg.set_reserved_sms(300)


def test_cfg4_shaped_batch_bit_exact(rlg):
    """BASELINE config 4's shape per GPU (128 of the 1024 episodes: 2048-point completions against 1400-point partial
    clouds, ragged N != M): the whole batch bit for bit against the C oracle."""
    pc1, pc2 = O.make_clouds(128, 2048, "sphere", 401), O.make_clouds(128, 1400, "uniform", 402)
    _check_against_direct(rlg, pc1, pc2)
