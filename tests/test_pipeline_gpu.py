"""CUDA-graph step runners (gan-rl_3d_b200/pipeline.py): the captured steps give the same losses and gradients as
the same calls made eagerly, for device-resident batches and for pinned host batches (one-transfer and two-transfer)."""
import importlib

import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _eager(rlg, a, b):
    a = a.to(DEV).requires_grad_(True)
    loss = rlg.ChamferLoss()(a, b.to(DEV))
    loss.backward()
    return loss.detach(), a.grad


def test_device_step_graph_matches_eager(rlg):
    P = importlib.import_module("gan-rl_3d_b200.pipeline")
    batches = [(O.make_clouds(3, 500, "sphere", 10 + k).to(DEV), O.make_clouds(3, 400, "sphere", 20 + k).to(DEV)) for k in range(4)]
    g = P.ChamferStepGraph(batches)
    g.replay()
    g.replay()
    torch.cuda.synchronize()
    for k, (a, b) in enumerate(batches):
        loss, grad = _eager(rlg, a, b)
        assert torch.equal(g.losses[k], loss)
        assert torch.allclose(g.grads[k], grad, rtol=1e-6, atol=1e-9)      # atomics: summation order may differ
    assert g.kernel_launches_per_replay == P.launches_per_step(500, 400) * len(batches)


@pytest.mark.parametrize("one_transfer", [True, False])
def test_host_step_graph_matches_eager(rlg, one_transfer):
    P = importlib.import_module("gan-rl_3d_b200.pipeline")
    raw = [(O.make_clouds(2, 300, "uniform", 30 + k), O.make_clouds(2, 257, "uniform", 40 + k)) for k in range(5)]
    host = [P.pin_pair(a, b) if one_transfer else (a.pin_memory(), b.pin_memory()) for a, b in raw]
    g = P.HostChamferStepGraph(host, DEV)
    assert g.single_copy == one_transfer
    for _ in range(2):
        g.replay()
        torch.cuda.synchronize()
        for k, (a, b) in enumerate(raw):
            loss, _ = _eager(rlg, a, b)
            assert g.losses_host[k].item() == loss.item()
    assert g.h2d_bytes_per_step == (2 * 300 * 3 + 2 * 257 * 3) * 4 and g.d2h_bytes_per_step == 4


def test_host_step_graph_read_back_survives_reuse_of_the_capture_pool(rlg):
    """Regression: every step's loss is read back by a D2H copy on its own stream.  Its device scalar must stay
    allocated, or the capture pool hands the block to the next step's outputs and the copy reads garbage now and then.
    Fresh data per trial and a poisoned allocator make a stale or overwritten read visible."""
    P = importlib.import_module("gan-rl_3d_b200.pipeline")
    for trial in range(4):
        junk = [torch.randn(1 << 18, device=DEV) * 3 for _ in range(8)]
        del junk
        raw = [(O.make_clouds(2, 300, "uniform", 500 + 10 * trial + k), O.make_clouds(2, 257, "uniform", 700 + 10 * trial + k))
               for k in range(5)]
        g = P.HostChamferStepGraph([P.pin_pair(a, b) for a, b in raw], DEV)
        want = [_eager(rlg, a, b)[0].item() for a, b in raw]
        for _ in range(3):
            g.replay()
            torch.cuda.synchronize()
            assert [g.losses_host[k].item() for k in range(5)] == want
