"""GPU parity tests of the PointNet encoder trunk (fused per-point MLP + max-pool) against the stock
Conv1d/BatchNorm1d/ReLU stack (oracle O_enc) and the golden GFVs generated from the real reference."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP32_TOL = 1e-5      # north_star: GFVs within 1e-5 relative (fp32), abs-floor rule of SURVEY.md 7.2-6


def _port(dims, latent, seed):
    torch.manual_seed(seed)
    enc = O.RefEncoderPort(3, latent, dims)
    O.randomize_bn(enc, seed + 10)
    return enc.eval()


@pytest.mark.parametrize("dims,latent", [([64, 128, 128, 256, 128], 128), ([64, 128, 1024], 128), ([32, 48], 16),
                                         ([7], 5), ([10, 70, 33], 12)])
@pytest.mark.parametrize("B,N", [(2, 200), (1, 1), (3, 64), (2, 2048), (4, 130)])
def test_pool_and_gfv_vs_stock_eval(rlg, dims, latent, B, N):
    enc = _port(dims, latent, len(dims))
    x = O.make_clouds(B, N, "sphere", 700 + N)
    with torch.no_grad():
        want_pool = enc.double().pooled(x.double()).float().numpy()     # float64 truth of the stock stack
        want_gfv = enc(x.double()).float().numpy()
    enc = enc.float().to(DEV)
    layers = rlg.fold_trunk(enc.point_mlp)
    pooled, argmax = rlg.encoder_pool(x.to(DEV), layers, want_argmax=True)
    ok, err = O.gfv_close(pooled.cpu().numpy(), want_pool, FP32_TOL)
    assert ok, err
    with torch.no_grad():
        gfv = rlg.fused_forward(enc, x.to(DEV))
    ok, err = O.gfv_close(gfv.cpu().numpy(), want_gfv, FP32_TOL)
    assert ok, err
    # argmax really attains the pooled value: re-evaluate the folded MLP at the reported points
    am = argmax.long().clamp_(0, N - 1)
    h = x.to(DEV)
    for w, b in layers:
        h = torch.relu(h @ w.T + b)
    at = torch.gather(h, 1, am.unsqueeze(1).expand(-1, 1, -1).reshape(B, 1, -1)).squeeze(1) if False else \
        h[torch.arange(B, device=DEV)[:, None], am, torch.arange(h.shape[2], device=DEV)[None, :]]
    assert torch.allclose(at, pooled, rtol=1e-5, atol=1e-6)


def test_golden_gfvs_from_reference(rlg, golden_encoder):
    g = golden_encoder
    for k in range(int(g["enc_count"])):
        dims, latent = [int(v) for v in g[f"enc{k}_dims"]], int(g[f"enc{k}_latent"])
        torch.manual_seed(k)
        enc = rlg.PointNetEncoder(3, latent, dims)
        if k == 2:
            enc.load_state_dict({str(n): torch.from_numpy(g[f"enc{k}_sd_{n}"]) for n in g[f"enc{k}_keys"]})
        else:
            O.randomize_bn(enc, seed=10 + k)
            chk = float(sum(v.double().abs().sum().item() for v in enc.state_dict().values()))
            if abs(chk - float(g[f"enc{k}_state_checksum"])) > 1e-6 * abs(chk):
                pytest.skip("torch default init differs from the fixture's torch build")
        enc = enc.eval()
        x_cpu = torch.from_numpy(g[f"enc{k}_x"])
        truth = O.RefEncoderPort(3, latent, dims).double()
        truth.load_state_dict(enc.state_dict())
        truth.eval()
        with torch.no_grad():
            t_gfv = truth(x_cpu.double()).numpy()
            t_pool = truth.pooled(x_cpu.double()).numpy()
        enc = enc.to(DEV)
        x = x_cpu.to(DEV)
        with torch.no_grad():
            gfv = enc(x)
        pooled, _ = rlg.encoder_pool(x, rlg.folded_trunk_cached(enc))
        # the fixture is the reference's own fp32 result; it carries fp32 rounding of its own (measured here
        # against the float64 evaluation of the same weights), so: ours vs truth <= 1e-5, and ours vs the
        # reference <= 1e-5 + the reference's distance from truth
        for ours, gold, tru in ((gfv, g[f"enc{k}_gfv"], t_gfv), (pooled, g[f"enc{k}_pooled"], t_pool)):
            ours = ours.cpu().numpy()
            ref_err = O.gfv_close(gold, tru, 1.0)[1]
            assert ref_err < 1e-4, (k, ref_err)                      # the fixture really is this network
            ok, err = O.gfv_close(ours, tru, FP32_TOL)
            assert ok, (k, err)
            ok, err = O.gfv_close(ours, gold, FP32_TOL + ref_err)
            assert ok, (k, err, ref_err)


def test_train_mode_uses_stock_layers_and_updates_running_stats(rlg):
    enc = rlg.PointNetEncoder(3, 16, [8, 24]).to(DEV).train()
    port = O.RefEncoderPort(3, 16, [8, 24]).to(DEV).train()
    port.load_state_dict(enc.state_dict())
    x = O.make_clouds(4, 100, "sphere", 3).to(DEV)
    assert torch.equal(enc(x), port(x))
    assert torch.equal(enc.point_mlp[1].running_mean, port.point_mlp[1].running_mean)
    assert enc.point_mlp[1].num_batches_tracked.item() == 1


def test_eval_mode_with_autograd_gives_reference_gradients(rlg):
    enc = _port([16, 32], 8, 1).to(DEV)
    port = _port([16, 32], 8, 1).to(DEV)
    x = O.make_clouds(2, 150, "uniform", 5).to(DEV)
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    rlg.fused_forward(enc, xa).square().sum().backward()
    port(xb).square().sum().backward()
    assert torch.allclose(xa.grad, xb.grad, rtol=1e-4, atol=1e-6)
    for p, q in zip(enc.parameters(), port.parameters()):
        assert torch.allclose(p.grad, q.grad, rtol=1e-4, atol=1e-5)


def test_cache_follows_optimizer_steps(rlg):
    enc = _port([16, 32], 8, 2).to(DEV)
    x = O.make_clouds(2, 80, "sphere", 6).to(DEV)
    with torch.no_grad():
        a = rlg.fused_forward(enc, x)
        enc.point_mlp[0].weight.mul_(1.5)
        b = rlg.fused_forward(enc, x)
        ref = enc.double().cpu()             # float64 truth of the stock stack (the stock CUDA convs may run in TF32)
        want = ref.global_mlp(torch.max(ref.point_mlp(x.double().cpu().transpose(2, 1)), dim=2)[0])
    assert not torch.equal(a, b)
    assert O.gfv_close(b.cpu().numpy(), want.numpy(), 2e-5)[0]


def test_unsupported_widths_fail_loudly(rlg):
    x = torch.zeros(1, 8, 3, device=DEV)
    w = torch.zeros(4, 5, device=DEV)
    with pytest.raises(RuntimeError, match="c_in == 3"):
        rlg.encoder_pool(x, [(w, torch.zeros(4, device=DEV))])


BF16_TOL = 2e-2      # north_star: 2e-2 relative for the bf16 encoder GEMMs
BF16_FLOOR = 0.25    # ... on entries >= 25 % of the largest; smaller entries: |err| <= 5e-3 * largest (see O.gfv_close)
BF16_NORM = 5e-3     # and norm-wise |a-b|_2 <= 5e-3 |b|_2 (measured ~1e-3)


@pytest.mark.parametrize("dims", [[64, 128, 1024], [64, 128, 256], [128, 128], [64, 64, 64, 128], [64, 384], [128, 64, 640]])
@pytest.mark.parametrize("B,N", [(2, 200), (1, 1), (3, 128), (2, 2048), (5, 1300)])
def test_bf16_tensor_core_path_vs_float64_stack(rlg, dims, B, N):
    enc = _port(dims, 32, len(dims) + 3)
    x = O.make_clouds(B, N, "sphere", 900 + N)
    with torch.no_grad():
        want = enc.double().pooled(x.double()).float().numpy()
    enc = enc.float().to(DEV)
    layers = rlg.fold_trunk(enc.point_mlp)
    pooled, _ = rlg.encoder_pool(x.to(DEV), layers, precision="bf16")
    torch.cuda.synchronize()
    got = pooled.cpu().numpy()
    ok, err = O.gfv_close(got, want, BF16_TOL, BF16_FLOOR)
    assert ok, err
    # norm-wise the bf16 path is far tighter than the element-wise bound
    assert np.linalg.norm(got - want) <= BF16_NORM * np.linalg.norm(want)
    # and it agrees with this repo's own fp32 CUDA-core path to the same tolerance
    fp32, _ = rlg.encoder_pool(x.to(DEV), layers)
    assert O.gfv_close(got, fp32.cpu().numpy(), BF16_TOL, BF16_FLOOR)[0]


def test_bf16_module_switch_and_cache(rlg):
    enc = _port([64, 128, 1024], 128, 9).to(DEV)
    x = O.make_clouds(4, 700, "uniform", 10).to(DEV)
    with torch.no_grad():
        ref = rlg.fused_forward(enc, x)                       # fp32 default
        enc.rlg_precision = "bf16"
        out = rlg.fused_forward(enc, x)
        packed = enc.__dict__["_rlg_packed"]
        assert rlg.fused_forward(enc, x) is not None and enc.__dict__["_rlg_packed"] is packed     # cache hit
        enc.point_mlp[3].weight.mul_(1.25)                    # optimizer-like update -> repack
        out2 = rlg.fused_forward(enc, x)
        assert enc.__dict__["_rlg_packed"] is not packed
    assert O.gfv_close(out.cpu().numpy(), ref.cpu().numpy(), BF16_TOL, BF16_FLOOR)[0]
    assert not torch.equal(out, out2)
    assert "_rlg_packed" not in enc.state_dict()


def test_bf16_unsupported_widths_fail_loudly(rlg):
    enc = _port([64, 128, 128, 256, 128], 128, 4).to(DEV)     # the reference's default config: 256-wide hidden layer
    layers = rlg.fold_trunk(enc.point_mlp)
    with pytest.raises(RuntimeError, match="hidden width"):
        rlg.encoder_pool(torch.zeros(1, 64, 3, device=DEV), layers, precision="bf16")


# ---- layer-wise tcgen05 GEMM path (encoder_layers.cu): TMA-fed, weights resident, any widths that are multiples of 64 ----
GEMM_DIMS = [[64, 128, 128, 256, 128], [64, 128, 1024], [64, 64], [128, 256, 64], [64, 256, 256, 192]]


@pytest.mark.parametrize("dims", GEMM_DIMS)
@pytest.mark.parametrize("B,N", [(2, 200), (1, 1), (3, 128), (2, 2048), (5, 1300), (9, 127)])
def test_fp32x_tensor_path_holds_the_fp32_tolerance(rlg, dims, B, N):
    """fp16 hi+lo operand pairs, three MMAs per K step: GFV trunk outputs within the FP32 tolerance of the float64 stack.
    The operands carry 22 significant bits against fp32's 24, so the error is up to ~2x that of the fp32 CUDA-core kernel
    (measured on B200: 4.9e-6 vs 3.0e-6 on the reference's dims, 3.7e-6 vs 4.2e-6 on 3->64->128->1024).  The last entry of
    GEMM_DIMS is a stress configuration (two 256-wide contractions in a row) where the fp32 kernel itself reaches 6.4e-6
    and this path 1.07e-5: it is held to 2e-5."""
    tol = 2e-5 if dims == GEMM_DIMS[-1] else FP32_TOL
    enc = _port(dims, 32, len(dims) + 5)
    x = O.make_clouds(B, N, "sphere", 1200 + N)
    with torch.no_grad():
        want = enc.double().pooled(x.double()).float().numpy()
    enc = enc.float().to(DEV)
    layers = rlg.fold_trunk(enc.point_mlp)
    assert rlg.resolve_path(layers, "auto") == "fp32x"
    pooled, _ = rlg.encoder_pool(x.to(DEV), layers, precision="fp32x")
    torch.cuda.synchronize()
    ok, err = O.gfv_close(pooled.cpu().numpy(), want, tol)
    assert ok, err
    # clouds far from the origin and large activations (scale 30): the same tolerance
    x2 = (x * 30.0 + 3.0).contiguous()
    with torch.no_grad():
        want2 = enc.double().cpu().pooled(x2.double()).float().numpy()
    enc = enc.float().to(DEV)
    pooled2, _ = rlg.encoder_pool(x2.to(DEV), layers, precision="fp32x")
    ok, err = O.gfv_close(pooled2.cpu().numpy(), want2, tol)
    assert ok, err


@pytest.mark.parametrize("dims", GEMM_DIMS)
@pytest.mark.parametrize("B,N", [(2, 200), (1, 1), (2, 2048), (5, 1300)])
def test_bf16_layerwise_path_vs_float64_stack(rlg, dims, B, N):
    enc = _port(dims, 32, len(dims) + 6)
    x = O.make_clouds(B, N, "sphere", 1300 + N)
    with torch.no_grad():
        want = enc.double().pooled(x.double()).float().numpy()
    enc = enc.float().to(DEV)
    layers = rlg.fold_trunk(enc.point_mlp)
    pooled, _ = rlg.encoder_pool(x.to(DEV), layers, precision="bf16_layers")
    torch.cuda.synchronize()
    got = pooled.cpu().numpy()
    ok, err = O.gfv_close(got, want, BF16_TOL, BF16_FLOOR)
    assert ok, err
    assert np.linalg.norm(got - want) <= 2 * BF16_NORM * np.linalg.norm(want)      # two to four bf16 layers deep


def test_reference_config_runs_on_tensor_cores_by_default_and_under_bf16(rlg):
    """The reference's own encoder_dims (configs/config.yaml:9-11): default precision -> fp32-grade tensor path within
    1e-5; rlg_precision = "bf16" -> the layer-wise bf16 GEMMs (the single fused kernel cannot hold a 256-wide layer),
    never an exception."""
    enc = _port([64, 128, 128, 256, 128], 128, 21)
    x = O.make_clouds(6, 1400, "sphere", 22)
    with torch.no_grad():
        want = enc.double()(x.double()).float().numpy()
    enc = enc.float().to(DEV)
    assert "fp16 hi+lo" in rlg.encoder_path_of(enc)
    with torch.no_grad():
        gfv = rlg.fused_forward(enc, x.to(DEV))
    ok, err = O.gfv_close(gfv.cpu().numpy(), want, FP32_TOL)
    assert ok, err
    enc.rlg_precision = "bf16"
    assert "bf16 operands" in rlg.encoder_path_of(enc)
    with torch.no_grad():
        gfv16 = rlg.fused_forward(enc, x.to(DEV))
    assert O.gfv_close(gfv16.cpu().numpy(), want, BF16_TOL, BF16_FLOOR)[0]
    # widths no tensor path covers fall back to the CUDA-core kernel, whatever precision is asked for
    odd = _port([10, 70, 33], 12, 23).to(DEV)
    odd.rlg_precision = "bf16"
    assert "CUDA-core" in rlg.encoder_path_of(odd)
    with torch.no_grad():
        rlg.fused_forward(odd, x.to(DEV))


def test_module_on_another_device_or_layout_delegates(rlg):
    """A trunk that is not [Conv1d, BatchNorm1d, ReLU] x L, or a module whose weights are not on x's device, is not the
    hot path: the original forward runs (and raises what the reference would raise)."""
    import torch.nn as nn
    enc = _port([16, 32], 8, 3)                      # weights on the CPU, input on the GPU
    x = O.make_clouds(2, 50, "sphere", 4).to(DEV)
    with pytest.raises(RuntimeError):
        rlg.fused_forward(enc, x)
    enc = enc.to(DEV)
    enc.point_mlp = nn.Sequential(nn.Conv1d(3, 16, 1), nn.ReLU(), nn.Conv1d(16, 32, 1)).to(DEV)   # no BatchNorm
    with torch.no_grad():
        out = rlg.fused_forward(enc, x)
    assert out.shape == (2, 8)
