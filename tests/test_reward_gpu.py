"""GPU parity of the batched reward evaluation against rewards the REAL reference RewardFunction
(utils/losses.py:209-246) produced episode by episode (tests/golden/reward_ref.npz) and against float64 truth."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _truth(pred, target, pg, tg, disc, wc=100.0, wg=10.0, wd=0.01):
    f1, f2, *_ = O.chamfer_f64(pred, target)
    cd = (f1.mean(1) + f2.mean(1)).numpy() / 2.0
    gfv = ((pg.double() - tg.double()) ** 2).mean(1).numpy()
    return -(wc * cd + wg * gfv + wd * (-disc.double().reshape(len(cd), -1).mean(1).numpy()))


def test_batched_rewards_match_reference_episode_by_episode(rlg, golden_reward):
    g = golden_reward
    for k in range(int(g["rw_count"])):
        E, N, M, s1, s2 = [int(v) for v in g[f"rw{k}_meta"]]
        kind = str(g[f"rw{k}_kind"])
        pred, target = O.make_clouds(E, N, kind, s1), O.make_clouds(E, M, kind, s2)
        pg, tg = torch.from_numpy(g[f"rw{k}_pred_gfv"]), torch.from_numpy(g[f"rw{k}_target_gfv"])
        disc = torch.from_numpy(g[f"rw{k}_disc"])
        got = rlg.batched_rewards(pred.to(DEV), target.to(DEV), pg.to(DEV), tg.to(DEV), disc.to(DEV)).cpu().numpy()
        assert got.shape == (E,)
        truth = _truth(pred, target, pg, tg, disc)
        assert O.rel_err(got, truth) < 1e-6                                   # vs float64
        ref = g[f"rw{k}_rewards"]
        tol = max(1e-5, 2 * O.rel_err(ref, truth))                            # the reference's own cdist noise (N > 25)
        assert O.rel_err(got, ref) < tol
        # the class mirror: one scalar for a B=1 call, like the environment's per-step use (rl_gan_net.py:316-324)
        rf = rlg.RewardFunction()
        one = rf.compute_reward(pred[:1].to(DEV), target[:1].to(DEV), pg[:1].to(DEV), tg[:1].to(DEV), disc[:1].to(DEV))
        assert one.dim() == 0 and abs(one.item() - got[0]) <= 1e-5 * abs(got[0])
        assert torch.allclose(rf.compute_rewards(pred.to(DEV), target.to(DEV), pg.to(DEV), tg.to(DEV), disc.to(DEV)).cpu(),
                              torch.from_numpy(got))


def test_rewards_do_not_track_gradients_and_accept_flat_logits(rlg):
    E = 3
    pred = O.make_clouds(E, 300, "sphere", 1).to(DEV).requires_grad_(True)
    target = O.make_clouds(E, 200, "sphere", 2).to(DEV)
    pg, tg = torch.rand(E, 16, device=DEV), torch.rand(E, 16, device=DEV)
    r1 = rlg.batched_rewards(pred, target, pg, tg, torch.zeros(E, device=DEV))
    r2 = rlg.batched_rewards(pred, target, pg, tg, torch.zeros(E, 1, device=DEV))
    assert not r1.requires_grad and torch.equal(r1, r2)
