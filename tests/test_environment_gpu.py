"""Batched RL environment step (environment.BatchedRLEnvironment) against tests/golden/environment_ref.npz: states, next
states and rewards of the reference's own RLGANNet + RLGANNetEnvironment (models/rl_gan_net.py:267-339), played one episode
at a time on the CPU by gen_golden.py."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Model:
    """The slice of RLGANNet's interface the environment uses (models/rl_gan_net.py:100-203)."""

    def __init__(self, rlg, g):
        latent, points = (int(v) for v in g["ae_dims"])
        self.autoencoder = rlg.PointCloudAutoencoder(3, latent, points, [64, 128, 64], [64, points * 3])
        self.autoencoder.load_state_dict({str(n): torch.from_numpy(g[f"ae_sd_{n}"]) for n in g["ae_keys"]})
        self.latent_gan = O.RefLatentGANPort(2, latent, [48, 32], [32, 16, 1])
        self.latent_gan.load_state_dict({str(n): torch.from_numpy(g[f"lgan_sd_{n}"]) for n in g["lgan_keys"]})
        self.autoencoder.to(DEV).eval()
        self.latent_gan.to(DEV).eval()
        self.device = torch.device(DEV)
        w = g["weights"]
        self.reward_function = rlg.RewardFunction(float(w[0]), float(w[1]), float(w[2]))

    def encode_point_cloud(self, pc):
        return self.autoencoder.encode(pc)

    def decode_gfv(self, gfv):
        return self.autoencoder.decode(gfv)

    def generate_clean_gfv(self, z):
        return self.latent_gan.generate(z)


@pytest.fixture(scope="module")
def golden_env():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "environment_ref.npz"))


@pytest.mark.parametrize("capture", [False, True])
def test_batched_step_reproduces_the_reference_episodes(rlg, golden_env, capture):
    g = golden_env
    model = _Model(rlg, g)
    Env = __import__("importlib").import_module("gan-rl_3d_b200.environment").BatchedRLEnvironment
    env = Env.from_model(model, capture=capture)
    batch = {"incomplete": torch.from_numpy(g["incomplete"]), "complete": torch.from_numpy(g["complete"])}
    states = env.reset(batch)
    assert states.is_cuda and states.shape == g["states"].shape
    assert np.abs(states.cpu().numpy() - g["states"]).max() <= 1e-5 * np.abs(g["states"]).max()
    for rep in range(2):                                    # the second call replays the captured graph
        next_states, rewards, dones, info = env.step(g["actions"])
        assert rewards.is_cuda and rewards.shape == (int(g["E"]),) and bool(dones.all())
        assert np.abs(next_states.cpu().numpy() - g["next_states"]).max() <= 1e-5
        r = rewards.cpu().numpy().astype(np.float64)
        assert np.abs(r - g["rewards"]).max() <= 1e-5 * np.abs(g["rewards"]).max(), (r, g["rewards"])
        assert info["completed_pc"].shape == (int(g["E"]), 256, 3)
    # a different action batch through the same (captured) step: rewards change, episode by episode as in an eager run
    other = -g["actions"]
    _, r_graph, _, _ = env.step(other)
    eager = Env.from_model(model, capture=False)
    eager.reset(batch)
    _, r_eager, _, _ = eager.step(other)
    assert torch.allclose(r_graph, r_eager, rtol=1e-6, atol=1e-6)
    assert not torch.allclose(r_graph, rewards)
