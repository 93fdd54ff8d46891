"""The captured autoencoder training step (ae_step.AEStepGraph) against the same steps run eagerly on stock torch layers in
float64 (train_rl_gan_net.py:220-249: zero_grad, forward, ChamferLoss, backward, Adam step)."""
import copy

import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_state_dict_keys_match_the_reference_layout(rlg):
    ae = rlg.PointCloudAutoencoder()
    keys = list(ae.state_dict().keys())
    assert "encoder.point_mlp.0.weight" in keys and "encoder.point_mlp.13.num_batches_tracked" in keys
    assert "encoder.global_mlp.0.weight" in keys and "decoder.mlp.6.bias" in keys
    assert ae.state_dict()["encoder.point_mlp.0.weight"].shape == (64, 3, 1)
    assert abs(sum(p.numel() for p in ae.parameters()) * 4 / 1e6 - 7.15) < 0.05     # 7.15 MB of fp32 gradients


def test_captured_steps_follow_the_float64_trajectory(rlg):
    torch.manual_seed(3)
    model = rlg.PointCloudAutoencoder(3, 32, 256, [64, 128, 64], [64, 768]).to(DEV).train()
    ref = copy.deepcopy(model).double().cpu().train()
    eager = copy.deepcopy(model).train()
    S, B = 4, 6
    batches = [(O.make_clouds(B, 180, "sphere", 50 + k), O.make_clouds(B, 256, "sphere", 90 + k)) for k in range(S)]
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True)
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=1e-5)
    g = rlg.AEStepGraph(model, opt, [(a.to(DEV), b.to(DEV)) for a, b in batches])
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    g.replay()
    torch.cuda.synchronize()
    ref_losses = []
    for a, b in batches:
        ropt.zero_grad()
        recon = ref.decoder(ref.encoder.global_mlp(torch.max(ref.encoder.point_mlp(a.double().transpose(2, 1)), dim=2)[0]))
        loss = O.ref_port_chamfer_loss(recon, b.double())
        loss.backward()
        ropt.step()
        ref_losses.append(float(loss))
    ours = [float(l) for l in g.losses]
    # the same steps issued eagerly through the same kernels: the capture must not change the trajectory
    eager_losses = []
    eopt = torch.optim.Adam(eager.parameters(), lr=1e-3, weight_decay=1e-5)
    for a, b in batches:
        eopt.zero_grad()
        loss = rlg.ChamferLoss()(eager(a.to(DEV))[0], b.to(DEV))
        loss.backward()
        eopt.step()
        eager_losses.append(float(loss))
    assert abs(ours[0] - ref_losses[0]) <= 1e-5 * abs(ref_losses[0]), (ours, ref_losses)
    assert abs(ours[1] - ref_losses[1]) <= 1e-3 * abs(ref_losses[1]), (ours, ref_losses)
    for k, (o, r, e) in enumerate(zip(ours, ref_losses, eager_losses)):
        # Adam's g/sqrt(v) turns rounding noise in near-zero gradient entries into +-lr parameter moves, so trajectories of
        # different arithmetics drift apart after two steps (stock fp32 CUDA drifts from float64 by 3 % here, ours by 1.4 %)
        assert abs(o - r) <= 5e-2 * abs(r), (ours, ref_losses)
        # (the Chamfer backward's float atomics make even two runs of the same kernels differ in the last bits)
        assert abs(o - e) <= (1e-4 if k < 2 else 2e-2) * abs(e), (ours, eager_losses)
    moved = 0
    for (n, p), (_, q) in zip(model.named_parameters(), eager.named_parameters()):
        moved += int((p.detach() != before[n]).any())
        # after 4 Adam steps of lr 1e-3 a parameter has moved at most ~4e-3
        assert float((p.detach() - q.detach()).abs().max()) <= 4.5e-3, n
        assert float((p.detach() - before[n]).abs().max()) <= 4.5e-3, n
    assert moved >= len(before) - 12           # biases in front of a batch-statistics BatchNorm have exactly zero gradient
    for (n, b), (_, c), (_, e) in zip(model.named_buffers(), ref.named_buffers(), eager.named_buffers()):
        if n.endswith("num_batches_tracked"):
            assert int(b) == int(c) == int(e) == S, n
        else:
            assert torch.allclose(b, e, rtol=1e-2, atol=2e-3), n                       # capture == eager
            assert torch.allclose(b.cpu().double(), c, rtol=0.2, atol=2e-2), n         # and both stay near float64's
    # a second replay keeps training (the graph reuses its buffers)
    g.replay()
    torch.cuda.synchronize()
    assert all(torch.isfinite(l) for l in g.losses)
