"""GPU: the drop-in boundary end to end -- install() on reference-shaped modules, an autoencoder training step
with ChamferLoss as the loss, and the single-process path of the sharding helpers on a CUDA device."""
import importlib
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference_shaped_modules():
    """Modules with the reference's names and call structure (utils/losses.py:13-75, models/autoencoder.py),
    built from the oracle port -- the reference checkout is not on the GPU box."""
    losses = types.ModuleType("utils_losses_port")

    def chamfer_distance_l2(pc1, pc2):
        return O.ref_port_chamfer_l2(pc1, pc2)

    def chamfer_distance(pc1, pc2, bidirectional=True):
        d1, d2 = losses.chamfer_distance_l2(pc1, pc2)
        return (d1 + d2) / 2.0 if bidirectional else d1

    class ChamferLoss(nn.Module):
        def forward(self, pred, target):
            return torch.mean(losses.chamfer_distance(pred, target, True))

    losses.chamfer_distance_l2, losses.chamfer_distance, losses.ChamferLoss = chamfer_distance_l2, chamfer_distance, ChamferLoss
    ae = types.ModuleType("models_autoencoder_port")
    ae.PointNetEncoder = type("PointNetEncoder", (O.RefEncoderPort,), {})
    return losses, ae


def _graph_nodes(fn, seen=None):
    seen = set() if seen is None else seen
    if fn is None or fn in seen:
        return ""
    seen.add(fn)
    return type(fn).__name__ + " " + " ".join(_graph_nodes(n, seen) for n, _ in fn.next_functions)


def test_install_routes_cuda_inputs_through_the_kernels(rlg):
    losses, ae = _reference_shaped_modules()
    crit_before = losses.ChamferLoss()
    pc1, pc2 = O.make_clouds(4, 512, "sphere", 1).to(DEV), O.make_clouds(4, 400, "sphere", 2).to(DEV)
    stock = crit_before(pc1, pc2).item()                           # stock torch-CUDA reference path
    enc = ae.PointNetEncoder(3, 32, [16, 64]).to(DEV)
    O.randomize_bn(enc, 4)
    enc.eval()
    tf32_conv, tf32_mm = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                        # TF32 convolutions make the stock path itself 1e-2 noisy
    torch.backends.cuda.matmul.allow_tf32 = False                  # (and cuDNN's algorithm choice varies run to run)
    with torch.no_grad():
        stock_gfv = enc(pc1)                                       # stock CUDA path in true fp32
        truth_gfv = ae.PointNetEncoder(3, 32, [16, 64]).double()
        truth_gfv.load_state_dict(enc.state_dict())
        truth_gfv = truth_gfv.eval()(pc1.cpu().double()).float()   # float64 truth of the same module on the host
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32_conv, tf32_mm
    rlg.install(losses, ae)
    try:
        a = pc1.clone().requires_grad_(True)
        v = crit_before(a, pc2)                                    # created before the patch, follows it
        assert "ChamferFnBackward" in _graph_nodes(v.grad_fn)        # the CUDA Function is in the autograd graph
        v.backward()
        d1, d2, i1, i2 = O.chamfer_direct(pc1.cpu(), pc2.cpu(), O.TIE_FAITHFUL)
        m1, m2 = O.chamfer_means(d1, d2)
        want = float(np.mean((m1.astype(np.float64) + m2) / 2))
        assert abs(v.item() - want) <= 1e-6 * want
        assert abs(v.item() - stock) <= 1e-4 * want                # the stock matmul path is the noisy one
        with torch.no_grad():
            gfv = enc(pc1)
        assert O.gfv_close(gfv.cpu().numpy(), truth_gfv.numpy(), 1e-5)[0]
        assert O.gfv_close(gfv.cpu().numpy(), stock_gfv.cpu().numpy(), 1e-4)[0]   # fp32 summation-order noise of the stock path
        # 4-D input (validate_joint's broadcast defect, train_rl_gan_net.py:541) is NOT the hot path: the
        # original function must see it unchanged
        four_d = torch.zeros(2, 2, 8, 3, device=DEV)
        assert torch.equal(losses.chamfer_distance(four_d, pc2[:2, :8]), O.ref_port_chamfer(four_d, pc2[:2, :8]))
    finally:
        rlg.uninstall()


def test_check_finite_routes_nan_clouds_to_the_original(rlg):
    """torch.min lets a NaN candidate win; the kernels are specified for finite clouds only.  With
    install(check_finite=True) a cloud holding a NaN must reach the original function and come back bit for bit."""
    losses, ae = _reference_shaped_modules()
    pc1, pc2 = O.make_clouds(2, 300, "sphere", 1).to(DEV), O.make_clouds(2, 200, "sphere", 2).to(DEV)
    bad = pc2.clone()
    bad[1, 17, 2] = float("nan")
    want_ok = O.ref_port_chamfer_l2(pc1, pc2)
    want_bad = O.ref_port_chamfer_l2(pc1, bad)
    rlg.install(losses, ae, check_finite=True)
    try:
        got_bad = losses.chamfer_distance_l2(pc1, bad)
        assert all(torch.equal(torch.isnan(g), torch.isnan(w)) and torch.equal(torch.nan_to_num(g), torch.nan_to_num(w))
                   for g, w in zip(got_bad, want_bad))
        got_ok = losses.chamfer_distance_l2(pc1, pc2)               # finite clouds still take the kernels
        assert all(torch.allclose(g, w, rtol=1e-4) for g, w in zip(got_ok, want_ok))
    finally:
        rlg.uninstall()


def test_autoencoder_training_step_with_chamfer_loss(rlg):
    """A config_quick-shaped AE step (BASELINE config 1 on the GPU): encoder in train mode (stock layers),
    decoder, ChamferLoss from the CUDA path, Adam.  Loss must fall and gradients must match the stock path."""
    torch.manual_seed(0)

    class AE(nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder = rlg.PointNetEncoder(3, 32, [16, 32])
            self.decoder = nn.Sequential(nn.Linear(32, 64), nn.ReLU(), nn.Linear(64, 256 * 3))

        def forward(self, x):
            return self.decoder(self.encoder(x)).view(-1, 256, 3)

    model = AE().to(DEV).train()
    x = O.make_clouds(8, 180, "sphere", 7).to(DEV)
    target = O.make_clouds(8, 256, "sphere", 8).to(DEV)
    # gradient of the loss w.r.t. the decoder output against the float64 closed form
    pred = model(x)
    pred.retain_grad()
    rlg.ChamferLoss()(pred, target).backward()
    ours = [p.grad.clone() for p in model.parameters()]
    pc, tc = pred.detach().cpu(), target.cpu()
    d1, d2, i1, i2 = O.chamfer_direct(pc, tc, O.TIE_FAITHFUL)
    up = np.full((8,), 0.5 / 8, np.float32)
    g_pred, _ = O.chamfer_bwd_truth(pc, tc, d1, d2, i1, i2, up, up)      # float64 closed form for the oracle's indices
    assert O.rowwise_rel_err(pred.grad.cpu().numpy(), g_pred) < 1e-5
    # parameter gradients against the stock loss on the same graph: the stock path is the noisy matmul-expansion cdist and
    # may route near-ties differently (an untrained decoder emits a tight blob of points), so norm-wise only
    model.zero_grad()
    O.ref_port_chamfer_loss(model(x), target).backward()
    for g, p in zip(ours, model.parameters()):
        assert (g - p.grad).norm() <= 5e-2 * p.grad.norm() + 1e-6
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    crit = rlg.ChamferLoss()
    first = last = None
    for _ in range(30):
        opt.zero_grad()
        loss = crit(model(x), target)
        loss.backward()
        opt.step()
        first = loss.item() if first is None else first
        last = loss.item()
    assert last < first


def test_sharding_helpers_single_process_cuda(rlg):
    D = importlib.import_module("gan-rl_3d_b200.distributed")
    pred = O.make_clouds(6, 128, "sphere", 1).to(DEV).requires_grad_(True)
    target = O.make_clouds(6, 128, "sphere", 2).to(DEV)
    loss = D.sharded_chamfer_loss(pred, target, 6)
    loss.backward()
    want = rlg.ChamferLoss()(pred.detach(), target)
    assert abs(loss.item() - want.item()) < 1e-6 * want.item()
    # two "ranks" emulated in one process: shard, compute, add -> equals the full-batch loss and gradient
    g_full = pred.grad.clone()
    pred.grad = None
    total = 0.0
    for r in range(2):
        lo, hi = D.shard_bounds(6, r, 2)
        part = rlg.chamfer_distance(pred[lo:hi], target[lo:hi]).sum() / 6
        part.backward()
        total += part.item()
    assert abs(total - want.item()) < 1e-6 * want.item()
    assert torch.allclose(pred.grad, g_full, rtol=1e-6, atol=1e-9)


def test_decoder_output_and_gradient_are_zero_copy_around_the_loss(rlg):
    """the decoder's last Linear writes the (B, 6144) tensor the Chamfer kernels read in place as (B, 2048, 3),
    and the gradient the backward kernel writes is the very storage the Linear's backward GEMM consumes (no repack either way)."""
    torch.manual_seed(0)
    dec = rlg.PointNetDecoder(32, 2048, [64, 6144]).to(DEV)
    gfv = torch.randn(4, 32, device=DEV)
    flat = dec.mlp(gfv)                                            # (B, 6144): the last Linear's output
    recon = flat.view(-1, 2048, 3)                                 # what PointNetDecoder.forward returns (autoencoder.py:126)
    assert recon.data_ptr() == flat.data_ptr() and recon.is_contiguous()
    seen = {}
    flat.register_hook(lambda g: seen.__setitem__("flat", g))
    recon.register_hook(lambda g: seen.__setitem__("view", g))
    target = O.make_clouds(4, 2048, "sphere", 3).to(DEV)
    rlg.ChamferLoss()(recon, target).backward()
    assert seen["view"].shape == (4, 2048, 3) and seen["flat"].shape == (4, 6144)
    assert seen["flat"].data_ptr() == seen["view"].data_ptr()      # the kernel's output buffer, viewed: no copy
    assert dec.mlp[-1].weight.grad is not None and float(dec.mlp[-1].weight.grad.abs().max()) > 0
