"""GPU parity tests of the train-mode encoder trunk (BatchNorm batch statistics, running-stat update, max-pool, and the
whole backward) against the stock Conv1d/BatchNorm1d/ReLU stack in float64 (models/autoencoder.py:32-47,65-71 as run by
train_rl_gan_net.py:220-249), plus the golden train-step fixture generated from the real reference module.

The gradient of a ReLU network is discontinuous where a pre-activation crosses zero, and any fp32 forward (the stock CUDA
one included) puts a handful of the millions of pre-activations of a batch on the other side of zero than float64 does.
So the tests come in two kinds:
  * small batches whose float64 forward has NO pre-activation within 3e-6 of zero (the seed is searched for): every output
    and gradient is compared with plain float64 autograd of the stock stack;
  * batches of the reference's size: the branch our forward took (ReLU masks, arg-max points; read back with
    train.inspect_saved) must differ from float64's only at near-ties, and the gradients are compared with float64
    autograd of the stock layers ON THAT BRANCH."""
import copy

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OUT_TOL = 1e-5       # pooled features and running statistics, relative to the tensor's largest entry
GRAD_TOL = 1e-5      # parameter gradients, relative to the tensor's largest entry
SAFE_MARGIN = 3e-6  # a batch whose float64 pre-activations all stay this far (relative) from zero takes the same branch in fp32
MODULE_TOL = 2e-4   # gradients through the whole module: global_mlp's BatchNorm over a handful of GFVs (stock torch) amplifies
NEAR_ZERO = 1e-5     # |pre-activation| / (largest of its layer) below which fp32 and float64 may disagree on the ReLU


def _port(dims, latent, seed, train=True):
    torch.manual_seed(seed)
    enc = O.RefEncoderPort(3, latent, dims)
    O.randomize_bn(enc, seed + 10)
    return enc.train(train)


def _max_rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _forward64(e64, x, masks=None, argmax=None, positive=None):
    """float64 stock layers; with masks/argmax: on a prescribed branch.  Returns pooled and the pre-activations."""
    h = x.double().transpose(2, 1)
    mods = list(e64.point_mlp)
    L = len(mods) // 3
    ys = []
    for l in range(L):
        y = mods[3 * l + 1](mods[3 * l](h))
        ys.append(y)
        if masks is None or l == L - 1:
            h = torch.relu(y)
        else:
            h = y * masks[l].transpose(1, 2).double()
    if argmax is None:
        return torch.max(h, dim=2)[0], ys
    return torch.gather(ys[-1], 2, argmax.unsqueeze(2)).squeeze(2) * positive.double(), ys


def _truth(enc, x, g, train, branch=None):
    e64 = copy.deepcopy(enc).double().cpu().train(train)
    feat, ys = _forward64(e64, x, *(branch or (None, None, None)))
    feat.backward(g.double())
    grads = {n: p.grad.clone() for n, p in e64.point_mlp.named_parameters()}
    bufs = {n: b.clone() for n, b in e64.point_mlp.named_buffers()}
    return feat.detach(), grads, bufs, [y.detach() for y in ys]


def _smallest_margin(enc, x, train):
    with torch.no_grad():
        e64 = copy.deepcopy(enc).double().cpu().train(train)
        _, ys = _forward64(e64, x)
    return min(float(y.abs().min() / y.abs().max()) for y in ys)


def _run_ours(rlg, enc, x, g, train):
    enc = copy.deepcopy(enc).float().to(DEV).train(train)
    pooled = rlg.trunk_pool_autograd(enc, x.to(DEV))
    branch = tuple(t.cpu() for t in _flatten(rlg.train.inspect_saved(enc, x.shape, pooled.grad_fn.saved_tensors[1])))
    pooled.backward(g.to(DEV))
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().cpu() for n, p in enc.point_mlp.named_parameters()}
    bufs = {n: b.detach().cpu() for n, b in enc.point_mlp.named_buffers()}
    L = len(enc.point_mlp) // 3
    return (pooled.detach().cpu(), grads, bufs), (list(branch[:L - 1]), branch[L - 1], branch[L])


def _flatten(b):
    masks, argmax, positive = b
    return list(masks) + [argmax, positive]


def _compare(ours, truth, train):
    pooled, grads, bufs = ours
    pooled64, grads64, bufs64 = truth[:3]
    assert _max_rel(pooled, pooled64) <= OUT_TOL, ("pooled", _max_rel(pooled, pooled64))
    for n, b64 in bufs64.items():
        if n.endswith("num_batches_tracked"):
            assert int(bufs[n]) == int(b64), n
        else:
            assert _max_rel(bufs[n], b64) <= OUT_TOL, (n, _max_rel(bufs[n], b64))
    worst = {}
    gmax = max(float(v.abs().max()) for v in grads64.values())
    for n, g64 in grads64.items():
        if train and n.split(".")[-1] == "bias" and int(n.split(".")[0]) % 3 == 0:
            # conv bias under batch statistics: the true gradient is exactly zero (float64 autograd leaves rounding noise)
            assert float(grads[n].abs().max()) <= 1e-6 * gmax, n
            continue
        worst[n] = _max_rel(grads[n], g64)
    bad = {n: e for n, e in worst.items() if e > GRAD_TOL}
    assert not bad, bad
    return worst


def _branch_is_legitimate(branch, ys64):
    """Our ReLU masks / arg-max points may differ from float64's only where float64 itself is within NEAR_ZERO of a tie."""
    masks, argmax, positive = branch
    flips = 0
    for l, m in enumerate(masks):
        y = ys64[l].transpose(1, 2)
        diff = m != (y > 0)
        flips += int(diff.sum())
        if diff.any():
            assert float(y[diff].abs().max()) <= NEAR_ZERO * float(y.abs().max()), f"layer {l}: a ReLU mask differs away from zero"
    y = ys64[-1]
    top = y.max(dim=2)[0]
    ours = torch.gather(y, 2, argmax.unsqueeze(2)).squeeze(2)
    scale = float(y.abs().max())
    live = top > NEAR_ZERO * scale
    assert bool(((top - ours)[live] <= NEAR_ZERO * scale).all()), "an arg-max point is not a near-tie of the float64 maximum"
    assert bool((positive == (top > 0))[top.abs() > NEAR_ZERO * scale].all())
    return flips


@pytest.mark.parametrize("dims", [[64, 128, 128, 256, 128], [64, 64], [128, 256, 64], [64, 192, 128]])
@pytest.mark.parametrize("B,N", [(2, 200), (3, 64), (1, 33), (4, 130)])
def test_train_forward_backward_vs_float64_autograd(rlg, dims, B, N):
    """Plain float64 autograd of the stock stack, on batches without a pre-activation near zero."""
    for seed in range(200):
        enc = _port(dims, 32, 1000 * len(dims) + seed)
        x = O.make_clouds(B, N, "sphere", 900 + N + seed)
        if _smallest_margin(enc, x, True) > SAFE_MARGIN:
            break
    else:
        pytest.fail("no seed without a near-zero pre-activation")
    g = torch.randn(B, dims[-1], generator=torch.Generator().manual_seed(5))
    ours, branch = _run_ours(rlg, enc, x, g, True)
    truth = _truth(enc, x, g, True)
    assert _branch_is_legitimate(branch, truth[3]) == 0
    _compare(ours, truth, True)


@pytest.mark.parametrize("dims", [[64, 128, 128, 256, 128], [64, 64]])
def test_eval_mode_backward_vs_float64(rlg, dims):
    """Running statistics (a frozen-BatchNorm encoder inside an autograd graph): rlg_encoder_bwd of SURVEY 8(b)."""
    enc = _port(dims, 32, 3, train=False)
    x = O.make_clouds(3, 500, "sphere", 77)
    g = torch.randn(3, dims[-1], generator=torch.Generator().manual_seed(6))
    ours, branch = _run_ours(rlg, enc, x, g, False)
    natural = _truth(enc, x, g, False)
    _branch_is_legitimate(branch, natural[3])
    _compare(ours, _truth(enc, x, g, False, branch), False)


@pytest.mark.parametrize("B,N,gscale", [(16, 1400, 1e-6), (32, 1400, 1.0), (8, 2048, 1.0)])
def test_reference_config_batches_on_their_branch(rlg, B, N, gscale):
    """config.yaml / config_quick.yaml shapes (incomplete clouds of ~1400 points, encoder_dims [64,128,128,256,128]); tiny
    upstream gradients (what a mean-reduced loss hands back) must survive the fp16 operand split."""
    dims = [64, 128, 128, 256, 128]
    enc = _port(dims, 128, 21)
    x = O.make_clouds(B, N, "sphere", 5)
    g = gscale * torch.randn(B, 128, generator=torch.Generator().manual_seed(7))
    ours, branch = _run_ours(rlg, enc, x, g, True)
    natural = _truth(enc, x, g, True)
    flips = _branch_is_legitimate(branch, natural[3])
    worst = _compare(ours, _truth(enc, x, g, True, branch), True)
    print(f"B={B} N={N}: {flips} ReLU decisions differ from float64 (all within {NEAR_ZERO} of zero); worst relative "
          f"gradient error on the branch {max(worst.values()):.2e}")


def test_module_forward_routes_train_mode_through_the_kernels(rlg):
    """fused_forward in train mode == stock module in train mode: GFV, running statistics, every gradient."""
    dims = [64, 128, 128, 256, 128]
    for seed in range(200):
        enc = _port(dims, 128, 4 + seed)
        x = O.make_clouds(8, 100, "sphere", 11 + seed)
        if _smallest_margin(enc, x, True) > SAFE_MARGIN:
            break
    ours = copy.deepcopy(enc).to(DEV).train()
    stock = copy.deepcopy(enc).double().train()
    gfv = rlg.fused_forward(ours, x.to(DEV))
    want = stock(x.double())
    assert _max_rel(gfv.detach().cpu(), want.detach()) <= OUT_TOL
    gfv.square().sum().backward()
    want.square().sum().backward()
    gmax = max(float(q.grad.abs().max()) for q in stock.parameters())
    for (n, p), (_, q) in zip(ours.named_parameters(), stock.named_parameters()):
        if float(q.grad.abs().max()) < 1e-9 * gmax:
            # exactly zero in exact arithmetic: a bias in front of a batch-statistics BatchNorm (conv biases, and the last
            # trunk BatchNorm's beta, which global_mlp's BatchNorm removes)
            assert float(p.grad.abs().max()) <= 1e-5 * gmax, n
            continue
        # the global MLP's BatchNorm (stock torch, fp32) normalises over only 8 GFVs and amplifies rounding: routing test
        assert _max_rel(p.grad.cpu(), q.grad) <= MODULE_TOL, (n, _max_rel(p.grad.cpu(), q.grad))
    for (n, b), (_, c) in zip(ours.named_buffers(), stock.named_buffers()):
        assert _max_rel(b.cpu(), c) <= OUT_TOL, n
    # unsupported widths keep the stock layers (no error)
    odd = O.RefEncoderPort(3, 16, [32, 48]).to(DEV).train()
    out = rlg.fused_forward(odd, x.to(DEV))
    assert out.shape == (8, 16) and out.requires_grad


def test_golden_train_step_from_reference(rlg):
    """tests/golden/encoder_train_ref.npz: the reference's own PointNetEncoder (models/autoencoder.py), train mode, one forward
    + backward in float64 (gen_golden.py).  Our module with the same weights must reproduce GFV, gradients and buffers."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "encoder_train_ref.npz"))
    dims, latent = [int(v) for v in g["dims"]], int(g["latent"])
    enc = rlg.PointNetEncoder(3, latent, dims)
    enc.load_state_dict({str(n): torch.from_numpy(g[f"sd_{n}"]) for n in g["keys"]})
    enc = enc.to(DEV).train()
    assert rlg.train_supported(enc.point_mlp)
    x = torch.from_numpy(g["x"]).to(DEV)
    # the trunk alone: pooled features and their gradients at 1e-5
    trunk = copy.deepcopy(enc)
    pooled = rlg.trunk_pool_autograd(trunk, x)
    assert _max_rel(pooled.detach().cpu(), g["pooled"]) <= OUT_TOL
    (pooled * torch.from_numpy(g["pcoef"]).float().to(DEV)).sum().backward()
    pmax = max(float(np.abs(g[f"pgrad_{n}"]).max()) for n, _ in trunk.point_mlp.named_parameters())
    for n, p in trunk.point_mlp.named_parameters():
        want = g[f"pgrad_{n}"]
        if np.abs(want).max() < 1e-9 * pmax:          # conv biases under batch statistics: zero
            assert float(p.grad.abs().max()) <= 1e-6 * pmax, n
        else:
            assert _max_rel(p.grad.cpu(), want) <= GRAD_TOL, (n, _max_rel(p.grad.cpu(), want))
    # the whole module (global_mlp stays stock torch)
    gfv = enc(x)
    assert _max_rel(gfv.detach().cpu(), g["gfv"]) <= OUT_TOL
    (gfv * torch.from_numpy(g["coef"]).float().to(DEV)).sum().backward()
    gmax = max(float(np.abs(g[f"grad_{n}"]).max()) for n, _ in enc.named_parameters())
    for n, p in enc.named_parameters():
        want = g[f"grad_{n}"]
        if np.abs(want).max() < 1e-9 * gmax:
            assert float(p.grad.abs().max()) <= 1e-5 * gmax, n
        else:
            assert _max_rel(p.grad.cpu(), want) <= MODULE_TOL, (n, _max_rel(p.grad.cpu(), want))
    for n, b in enc.named_buffers():
        if "num_batches" in n:
            assert int(b) == int(g[f"after_{n}"]), n
        else:
            assert _max_rel(b.cpu(), g[f"after_{n}"]) <= OUT_TOL, n


def _custom_trunk(dims, bias=True, affine=True, momentum=0.1):
    import torch.nn as nn

    class Enc(nn.Module):
        def __init__(self):
            super().__init__()
            seq, c_in = [], 3
            for c in dims:
                seq += [nn.Conv1d(c_in, c, 1, bias=bias), nn.BatchNorm1d(c, affine=affine, momentum=momentum), nn.ReLU(inplace=True)]
                c_in = c
            self.point_mlp = nn.Sequential(*seq)
            self.global_mlp = nn.Sequential(nn.Linear(c_in, 16), nn.BatchNorm1d(16), nn.ReLU(inplace=True))
    return Enc()


@pytest.mark.parametrize("bias,affine", [(False, True), (True, False), (False, False)])
@pytest.mark.parametrize("B,N", [(5, 1), (2, 37), (40, 9)])
def test_modules_without_conv_bias_or_affine_batchnorm(rlg, bias, affine, B, N):
    """Conv1d(bias=False) and BatchNorm1d(affine=False) blocks, single-point and tiny clouds: same contract."""
    for seed in range(200):
        torch.manual_seed(70 + seed)
        enc = _custom_trunk([64, 128, 64], bias, affine).train()
        x = O.make_clouds(B, N, "uniform", 300 + seed)
        if _smallest_margin(enc, x, True) > SAFE_MARGIN:
            break
    assert rlg.train_supported(enc.to(DEV).point_mlp)
    enc = enc.cpu()
    g = torch.randn(B, 64, generator=torch.Generator().manual_seed(8))
    ours, branch = _run_ours(rlg, enc, x, g, True)
    truth = _truth(enc, x, g, True)
    assert _branch_is_legitimate(branch, truth[3]) == 0
    _compare(ours, truth, True)


def test_inputs_outside_the_train_kernels_keep_the_stock_layers(rlg):
    """x.requires_grad, BatchNorm(momentum=None) and widths the kernels do not cover go to the module's own layers: same
    results as the stock forward, gradients still reach x."""
    x = O.make_clouds(3, 50, "sphere", 1).to(DEV)
    enc = _port([64, 128, 64], 16, 2).to(DEV).train()
    xg = x.clone().requires_grad_(True)
    out = rlg.fused_forward(enc, xg)
    out.sum().backward()
    assert xg.grad is not None and float(xg.grad.abs().max()) > 0
    cum = _custom_trunk([64, 64], momentum=None).to(DEV).train()
    assert not rlg.train_supported(cum.point_mlp)
    assert rlg.fused_forward(cum, x).shape == (3, 16)
    assert int(cum.point_mlp[1].num_batches_tracked) == 1                  # the stock layers ran (cumulative average)
    wide = _custom_trunk([64, 512]).to(DEV).train()
    assert not rlg.train_supported(wide.point_mlp)
    assert rlg.fused_forward(wide, x).requires_grad
    rlg.set_train_path(False)
    try:
        a = _port([64, 128, 64], 16, 2).to(DEV).train()
        assert "EncoderTrainFn" not in type(rlg.fused_forward(a, x).grad_fn).__name__
    finally:
        rlg.set_train_path(True)
