"""CPU: host-side logic of the drop-ins -- BatchNorm folding, cache invalidation, install()/uninstall()
rebinding, input-contract routing.  Nothing here launches a kernel."""
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import oracle as O


def test_fold_trunk_equals_eval_batchnorm(rlg):
    torch.manual_seed(0)
    enc = O.RefEncoderPort(3, 16, [8, 24, 40])
    O.randomize_bn(enc, 3)
    enc.eval()
    layers = rlg.fold_trunk(enc.point_mlp)
    assert [tuple(w.shape) for w, _ in layers] == [(8, 3), (24, 8), (40, 24)]
    x = O.make_clouds(2, 50, "uniform", 4)
    h = x.double()
    for w, b in layers:                                   # folded affine + ReLU, in float64
        h = torch.relu(h @ w.double().T + b.double())
    got = h.max(dim=1)[0]
    with torch.no_grad():
        want = enc.double().pooled(x.double())
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-6, atol=1e-7)


def test_fold_handles_negative_gamma_and_no_conv_bias(rlg):
    seq = nn.Sequential(nn.Conv1d(3, 4, 1, bias=False), nn.BatchNorm1d(4), nn.ReLU())
    with torch.no_grad():
        seq[1].weight.copy_(torch.tensor([-1.0, 2.0, -0.5, 1.0]))
        seq[1].running_mean.copy_(torch.tensor([0.1, -0.2, 0.3, 0.0]))
        seq[1].running_var.copy_(torch.tensor([0.5, 1.5, 2.0, 1.0]))
    seq.eval()
    (w, b), = rlg.fold_trunk(seq)
    x = torch.randn(5, 3, 7)
    with torch.no_grad():
        want = seq(x)
    got = torch.relu(torch.einsum("oc,bcn->bon", w, x) + b[None, :, None])
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-6)


def test_fold_rejects_foreign_layouts(rlg):
    with pytest.raises(ValueError):
        rlg.fold_trunk(nn.Sequential(nn.Conv1d(3, 4, 3), nn.BatchNorm1d(4), nn.ReLU()))
    with pytest.raises(ValueError):
        rlg.fold_trunk(nn.Sequential(nn.Conv1d(3, 4, 1), nn.ReLU()))


def test_folded_cache_invalidates_on_parameter_change(rlg):
    enc = O.RefEncoderPort(3, 8, [4, 6]).eval()
    a = rlg.folded_trunk_cached(enc)
    assert rlg.folded_trunk_cached(enc) is a                        # cache hit
    with torch.no_grad():
        enc.point_mlp[0].weight.mul_(2.0)                           # what an optimizer step does
    b = rlg.folded_trunk_cached(enc)
    assert b is not a and torch.allclose(b[0][0], a[0][0] * 2.0)
    sd = {k: v.clone() for k, v in enc.state_dict().items()}
    sd["point_mlp.1.running_mean"] += 1.0
    enc.load_state_dict(sd)                                         # copy_ bumps the buffer version
    assert rlg.folded_trunk_cached(enc) is not b
    assert "_rlg_folded" not in enc.state_dict()                    # derived cache never leaks into checkpoints


def _fake_reference_modules():
    losses = types.ModuleType("fake_losses")

    def chamfer_distance_l2(pc1, pc2):
        return O.ref_port_chamfer_l2(pc1, pc2)

    def chamfer_distance(pc1, pc2, bidirectional=True):
        d1, d2 = losses.chamfer_distance_l2(pc1, pc2)               # global lookup, like utils/losses.py:54
        return (d1 + d2) / 2.0 if bidirectional else d1

    losses.chamfer_distance_l2 = chamfer_distance_l2
    losses.chamfer_distance = chamfer_distance
    ae = types.ModuleType("fake_autoencoder")
    ae.PointNetEncoder = type("PointNetEncoder", (O.RefEncoderPort,), {})
    return losses, ae


def test_install_rebinds_and_delegates_non_hot_path_inputs(rlg):
    losses, ae = _fake_reference_modules()
    orig_l2, orig_fwd = losses.chamfer_distance_l2, ae.PointNetEncoder.forward
    done = rlg.install(losses, ae)
    try:
        assert set(done) == {"chamfer_distance_l2", "PointNetEncoder.forward"} and rlg.is_installed()
        assert losses.chamfer_distance_l2 is not orig_l2 and losses.chamfer_distance_l2.__wrapped__ is orig_l2
        # CPU tensors are outside the CUDA contract -> the saved original runs, bit-identical results
        pc1, pc2 = O.make_clouds(2, 30, "sphere", 1), O.make_clouds(2, 40, "sphere", 2)
        assert torch.equal(losses.chamfer_distance(pc1, pc2), O.ref_port_chamfer(pc1, pc2))
        enc = ae.PointNetEncoder(3, 8, [4, 6])
        keys = list(enc.state_dict().keys())
        enc.eval()
        x = O.make_clouds(2, 20, "uniform", 3)
        assert torch.equal(enc(x), orig_fwd(enc, x))                # CPU input -> original forward
        assert list(enc.state_dict().keys()) == keys                # module tree untouched
        assert rlg.install(losses, ae) == {}                        # idempotent
    finally:
        rlg.uninstall()
    assert losses.chamfer_distance_l2 is orig_l2 and ae.PointNetEncoder.forward is orig_fwd
    assert not rlg.is_installed()


def test_install_on_real_reference_modules(rlg, ref_losses, ref_autoencoder):
    """With the reference mounted: instances created BEFORE the patch follow it (SURVEY.md 8b probe)."""
    loss_before = ref_losses.ChamferLoss()
    reward_before = ref_losses.RewardFunction()
    rlg.install(ref_losses, ref_autoencoder)
    try:
        calls = []
        patched = ref_losses.chamfer_distance_l2
        ref_losses.chamfer_distance_l2 = lambda a, b: (calls.append(1), patched(a, b))[1]
        pc1, pc2 = O.make_clouds(2, 30, "sphere", 1), O.make_clouds(2, 30, "sphere", 2)
        v = loss_before(pc1, pc2)
        reward_before.chamfer_loss(pc1, pc2)
        assert len(calls) == 2 and torch.equal(v, O.ref_port_chamfer_loss(pc1, pc2))
        ref_losses.chamfer_distance_l2 = patched
        model = ref_autoencoder.PointCloudAutoencoder()
        sd_keys = list(model.state_dict().keys())
        model.eval()
        x = O.make_clouds(2, 64, "sphere", 5)
        with torch.no_grad():
            rec, gfv = model(x)                                     # CPU -> delegated to the stock forward
        assert rec.shape == (2, 2048, 3) and gfv.shape == (2, 128)
        assert list(model.state_dict().keys()) == sd_keys
    finally:
        rlg.uninstall()


def test_standalone_api_refuses_cpu_tensors(rlg):
    pc = torch.zeros(1, 4, 3)
    with pytest.raises(ValueError, match="no CPU path"):
        rlg.chamfer_distance_l2(pc, pc)
    with pytest.raises(ValueError, match="no CPU path"):
        rlg.ChamferLoss()(pc, pc)
    with pytest.raises(ValueError, match="no CPU path"):
        rlg.encoder_pool(pc, [(torch.zeros(4, 3), torch.zeros(4))])


def test_standalone_encoder_has_reference_state_dict_layout(rlg):
    ours = rlg.PointNetEncoder(3, 128, [64, 128, 1024])
    port = O.RefEncoderPort(3, 128, [64, 128, 1024])
    assert list(ours.state_dict().keys()) == list(port.state_dict().keys())
    assert [tuple(v.shape) for v in ours.state_dict().values()] == [tuple(v.shape) for v in port.state_dict().values()]
    ours.train()
    x = O.make_clouds(4, 32, "sphere", 1)
    port.load_state_dict(ours.state_dict())
    port.train()
    assert torch.equal(ours(x), port(x))                            # train mode: stock layers, batch statistics


def test_input_pipeline_host_side(rlg):
    """data.build_cache / draw_plan / rotation_matrix: shapes, ranges and the reference's distributions (no GPU needed)."""
    rng = np.random.default_rng(0)
    cache = rlg.build_cache([rng.normal(size=(2500, 3)), rng.normal(size=(100, 4)), rng.normal(size=(2048, 3))], num_points=2048)
    assert cache.shape == (3, 2048, 3) and cache.dtype == np.float32
    short = rng.normal(size=(100, 4))
    padded = rlg.build_cache([short], num_points=256, seed=3)[0]
    assert np.array_equal(padded[:100], short[:, :3].astype(np.float32))          # padded with repeats of its own points
    assert all(any(np.array_equal(p, q) for q in padded[:100]) for p in padded[100:])
    B, N = 64, 2048
    plan = rlg.draw_plan(rng, B, N)
    assert 0.2 <= plan["ratio"].min() and plan["ratio"].max() <= 0.5               # utils/dataset.py:255
    for b in range(B):
        if plan["method"][b] == 0:
            k = int(plan["n_keep"][b])
            assert k == int(N * (1 - plan["ratio"][b])) and len(set(plan["keep_idx"][b, :k])) == k      # :256,260
        else:
            kq, g = O.percentile_parts(N, float(plan["ratio"][b]))
            assert plan["q_index"][b] == kq and abs(plan["q_gamma"][b] - g) < 1e-15 and 0 <= plan["center"][b] < N
    assert plan["rot"].shape == (2, B, 9) and plan["scale"].shape == (2, B) and plan["pad_idx"].min() >= 0
    assert ((plan["scale"] == 1.0) | ((plan["scale"] >= 0.8) & (plan["scale"] <= 1.2))).all()                 # :292-294
    for R in plan["rot"].reshape(-1, 3, 3):
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-6) and abs(np.linalg.det(R) - 1) < 1e-5
    assert "jitter" not in plan and plan["jitter_on"].dtype == np.bool_
    th = np.array([0.3, -1.2, 2.5])
    cx, sx = np.cos(th[0]), np.sin(th[0])
    assert np.allclose(rlg.data.rotation_matrix(th) @ np.array([1.0, 0, 0]),
                       rlg.data.rotation_matrix([0, th[1], th[2]]) @ np.array([1.0, 0, 0]))   # Rx leaves the x axis alone
    assert np.allclose(rlg.data.rotation_matrix([th[0], 0, 0]), [[1, 0, 0], [0, cx, -sx], [0, sx, cx]])


def test_train_path_selection_on_cpu_modules(rlg):
    """train_supported is False for modules that are not float32 CUDA modules of the covered widths; fused_forward then
    runs the module's own layers (train mode on the CPU == the stock forward)."""
    torch.manual_seed(0)
    enc = rlg.PointNetEncoder(3, 16, [64, 128]).train()
    assert not rlg.train_supported(enc.point_mlp)                      # CPU parameters
    x = torch.randn(4, 50, 3)
    want = enc.global_mlp(torch.max(enc.point_mlp(x.transpose(2, 1)), dim=2)[0])
    enc2 = rlg.PointNetEncoder(3, 16, [64, 128]).train()
    enc2.load_state_dict({k: v for k, v in enc.state_dict().items() if "num_batches" not in k and "running" not in k}, strict=False)
    got = enc2(x)
    assert torch.allclose(got, want, atol=1e-6)
    assert int(enc2.point_mlp[1].num_batches_tracked) == 1             # the stock BatchNorm ran in train mode


def test_deterministic_backward_switch_follows_torch(rlg):
    """set_deterministic_backward(None) follows torch.use_deterministic_algorithms; True / False override it."""
    rlg.set_deterministic_backward(None)
    was = torch.are_deterministic_algorithms_enabled()
    try:
        torch.use_deterministic_algorithms(False)
        assert rlg.deterministic_backward() is False
        torch.use_deterministic_algorithms(True)
        assert rlg.deterministic_backward() is True
        rlg.set_deterministic_backward(False)
        assert rlg.deterministic_backward() is False
        torch.use_deterministic_algorithms(False)
        rlg.set_deterministic_backward(True)
        assert rlg.deterministic_backward() is True
    finally:
        torch.use_deterministic_algorithms(was)
        rlg.set_deterministic_backward(None)
