"""Generates tests/golden/*.npz from the REAL reference (phanich004/GAN-RL_3D), imported by file path from
/root/reference in the build container (the package import needs h5py, which is absent; utils/losses.py and
models/autoencoder.py themselves depend only on torch/numpy).  The reference holds no golden vectors for
this path, so these outputs are the pins.

    python tests/golden/gen_golden.py            # rewrites the fixtures (CPU, torch 2.11.0, ~10 s)

The fixtures carry inputs (or the seeds of oracle.make_clouds) and reference outputs, so the tests need
neither /root/reference nor this script at run time.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = os.environ.get("RLG_REFERENCE", "/root/reference")


def load_ref(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def chamfer_case(losses, pc1, pc2):
    """Everything the reference computes for one input pair, plus what autograd saves/returns."""
    a = pc1.clone().requires_grad_(True)
    b = pc2.clone().requires_grad_(True)
    dist1, dist2 = losses.chamfer_distance_l2(a, b)                 # utils/losses.py:13-39
    cd = losses.chamfer_distance(a, b)                              # :42-59
    cd_uni = losses.chamfer_distance(a, b, bidirectional=False)
    loss = losses.ChamferLoss()(a, b)                               # :62-75
    loss.backward()
    with torch.no_grad():
        D = torch.cdist(pc1, pc2, p=2)                              # :29 as written (mode picked by size)
        d1, i1 = torch.min(D, dim=2)                                # :32
        d2, i2 = torch.min(D, dim=1)                                # :33
    return dict(dist1=dist1.detach().numpy(), dist2=dist2.detach().numpy(), cd=cd.detach().numpy(),
                cd_uni=cd_uni.detach().numpy(), loss=np.float32(loss.item()),
                g1=a.grad.numpy(), g2=b.grad.numpy(),
                d1=d1.numpy(), d2=d2.numpy(), i1=i1.numpy().astype(np.int32), i2=i2.numpy().astype(np.int32))


def main():
    torch.set_num_threads(1)        # fixtures must not depend on the reduction split of a thread pool
    losses = load_ref("utils/losses.py", "ref_losses")
    ae = load_ref("models/autoencoder.py", "ref_autoencoder")
    out = {}

    # --- Chamfer, small shapes: the reference itself takes cdist's DIRECT path (N,M <= 25) -> bit-exact pins
    small = [(4, 16, 25, "sphere"), (3, 1, 7, "uniform"), (2, 25, 1, "sphere"), (5, 25, 25, "uniform"),
             (1, 2, 3, "sphere")]
    for k, (B, N, M, kind) in enumerate(small):
        pc1 = O.make_clouds(B, N, kind, seed=100 + k)
        pc2 = O.make_clouds(B, M, kind, seed=200 + k)
        if k == 3:   # exact duplicates -> exact ties
            pc2[:, 20:] = pc2[:, :5]
            pc1[:, 3] = pc2[:, 7]      # a zero distance -> zero-gradient rule of EuclideanDistBackward0
        case = chamfer_case(losses, pc1, pc2)
        for name, v in case.items():
            out[f"small{k}_{name}"] = v
        out[f"small{k}_pc1"] = pc1.numpy()
        out[f"small{k}_pc2"] = pc2.numpy()
    out["small_count"] = np.int32(len(small))

    # --- Chamfer, matmul-path shapes (N or M > 25): the as-written reference is noisy here;
    #     inputs are regenerated from seeds by the tests, outputs stored.
    big = [(2, 300, 257, "sphere", True), (2, 2048, 2048, "sphere", False), (2, 2048, 2048, "uniform", False),
           (2, 2048, 1400, "sphere", True)]
    for k, (B, N, M, kind, dup) in enumerate(big):
        pc1 = O.make_clouds(B, N, kind, seed=300 + k)
        pc2 = O.make_clouds(B, M, kind, seed=400 + k)
        if dup:
            pc2 = O.pad_with_duplicates(pc2, 0.25, seed=500 + k)
        case = chamfer_case(losses, pc1, pc2)
        for name in ("dist1", "dist2", "cd", "loss", "i1", "i2"):
            out[f"big{k}_{name}"] = case[name]
        out[f"big{k}_g1"] = case["g1"].astype(np.float32)
        out[f"big{k}_g2"] = case["g2"].astype(np.float32)
        out[f"big{k}_meta"] = np.array([B, N, M, 300 + k, 400 + k, 500 + k if dup else -1], np.int64)
        out[f"big{k}_kind"] = np.array(kind)
    out["big_count"] = np.int32(len(big))
    np.savez_compressed(os.path.join(HERE, "chamfer_ref.npz"), **out)

    # --- Encoder: reference PointNetEncoder in eval mode with non-trivial BatchNorm statistics
    enc_out = {}
    cfgs = [([64, 128, 128, 256, 128], 128), ([64, 128, 1024], 128), ([32, 48], 16)]
    for k, (dims, latent) in enumerate(cfgs):
        torch.manual_seed(k)
        enc = ae.PointNetEncoder(3, latent, dims)
        O.randomize_bn(enc, seed=10 + k)
        enc.eval()
        x = O.make_clouds(3, 200, "sphere", seed=600 + k)
        with torch.no_grad():
            gfv = enc(x)                                                        # autoencoder.py:56-76
            pooled = torch.max(enc.point_mlp(x.transpose(2, 1)), dim=2)[0]      # :65-71
        sd = enc.state_dict()
        checksum = float(sum(v.double().abs().sum().item() for v in sd.values()))
        enc_out[f"enc{k}_dims"] = np.array(dims, np.int32)
        enc_out[f"enc{k}_latent"] = np.int32(latent)
        enc_out[f"enc{k}_gfv"] = gfv.numpy()
        enc_out[f"enc{k}_pooled"] = pooled.numpy()
        enc_out[f"enc{k}_x"] = x.numpy()
        enc_out[f"enc{k}_state_checksum"] = np.float64(checksum)
        enc_out[f"enc{k}_keys"] = np.array(list(sd.keys()))
        if k == 2:   # tiny net: carry the weights themselves so the pin does not depend on torch's init RNG
            for name, v in sd.items():
                enc_out[f"enc{k}_sd_{name}"] = v.numpy()
    enc_out["enc_count"] = np.int32(len(cfgs))
    np.savez_compressed(os.path.join(HERE, "encoder_ref.npz"), **enc_out)

    # --- Encoder, TRAIN mode (train_rl_gan_net.py:220-249): the reference module itself, run in float64, one training
    #     forward + backward: GFV, every parameter gradient, the BatchNorm buffers after the step.  The seed is searched so
    #     that no pre-activation is within 3e-6 (relative) of zero: there the gradient is continuous and an fp32 run must
    #     reproduce it.
    tr_out = {}
    dims, latent, B, N = [64, 128, 64], 32, 4, 96
    for seed in range(500):
        torch.manual_seed(5000 + seed)
        enc = ae.PointNetEncoder(3, latent, dims).double()
        O.randomize_bn(enc, seed=40 + seed)
        enc.train()
        sd0 = {k: v.clone() for k, v in enc.state_dict().items()}
        x = O.make_clouds(B, N, "sphere", seed=650 + seed)
        margins = []
        hooks = [m.register_forward_hook(lambda mod, i, o: margins.append(float(o.abs().min() / o.abs().max())))
                 for m in enc.modules() if isinstance(m, torch.nn.BatchNorm1d)]
        gfv = enc(x.double())                                                    # autoencoder.py:56-76, train mode
        for h in hooks:
            h.remove()
        if min(margins) > 3e-6:
            break
    coef = torch.randn(B, latent, generator=torch.Generator().manual_seed(77)).double()
    (gfv * coef).sum().backward()
    # the trunk alone (autoencoder.py:65-71) from the same initial state: pooled features and their gradients
    import copy
    trunk = copy.deepcopy(enc)
    trunk.load_state_dict(sd0)
    trunk.train()
    trunk.zero_grad()
    pooled = torch.max(trunk.point_mlp(x.double().transpose(2, 1)), dim=2)[0]
    pcoef = torch.randn(B, dims[-1], generator=torch.Generator().manual_seed(78)).double()
    (pooled * pcoef).sum().backward()
    tr_out["pooled"], tr_out["pcoef"] = pooled.detach().numpy(), pcoef.numpy()
    for name, q in trunk.point_mlp.named_parameters():
        tr_out[f"pgrad_{name}"] = q.grad.numpy()
    tr_out["dims"], tr_out["latent"] = np.array(dims, np.int32), np.int32(latent)
    tr_out["x"], tr_out["coef"], tr_out["gfv"] = x.numpy(), coef.numpy(), gfv.detach().numpy()
    tr_out["margin"] = np.float64(min(margins))
    tr_out["keys"] = np.array(list(sd0.keys()))
    for name, v in sd0.items():
        tr_out[f"sd_{name}"] = v.numpy().astype(np.float32) if v.is_floating_point() else v.numpy()
    for name, v in enc.state_dict().items():
        if "running" in name or "num_batches" in name:
            tr_out[f"after_{name}"] = v.numpy()
    for name, q in enc.named_parameters():
        tr_out[f"grad_{name}"] = q.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "encoder_train_ref.npz"), **tr_out)

    # --- Reward: the reference's RewardFunction (utils/losses.py:209-246) called episode by episode
    #     with B=1 tensors, exactly as RLGANNetEnvironment.step does (models/rl_gan_net.py:316-324)
    rw_out = {}
    rf = losses.RewardFunction()
    shapes = [(6, 25, 25, "sphere"), (5, 2048, 2048, "sphere"), (4, 2048, 1400, "uniform")]
    for k, (E, N, M, kind) in enumerate(shapes):
        pred = O.make_clouds(E, N, kind, seed=700 + k)
        target = O.make_clouds(E, M, kind, seed=800 + k)
        g = torch.Generator().manual_seed(900 + k)
        pred_gfv, target_gfv = torch.rand(E, 128, generator=g), torch.rand(E, 128, generator=g)
        disc = torch.randn(E, 1, generator=g)
        with torch.no_grad():
            rewards = torch.stack([rf.compute_reward(pred[e:e + 1], target[e:e + 1], pred_gfv[e:e + 1], target_gfv[e:e + 1],
                                                     disc[e:e + 1]) for e in range(E)])
        rw_out[f"rw{k}_meta"] = np.array([E, N, M, 700 + k, 800 + k], np.int64)
        rw_out[f"rw{k}_kind"] = np.array(kind)
        rw_out[f"rw{k}_pred_gfv"] = pred_gfv.numpy()
        rw_out[f"rw{k}_target_gfv"] = target_gfv.numpy()
        rw_out[f"rw{k}_disc"] = disc.numpy()
        rw_out[f"rw{k}_rewards"] = rewards.numpy()
    rw_out["rw_count"] = np.int32(len(shapes))
    np.savez_compressed(os.path.join(HERE, "reward_ref.npz"), **rw_out)

    # --- Environment step: the reference's own RLGANNet + RLGANNetEnvironment (models/rl_gan_net.py:33-339)
    #     on a reduced configuration (so the fixture stays small), eval mode, one episode per reset/step pair exactly as
    #     train_rl_agent drives it (train_rl_gan_net.py:406-429): env-style batch keys 'incomplete' / 'complete'.
    import types
    import yaml
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))       # utils/__init__ imports it; nothing here uses it
    sys.path.insert(0, REF)
    import importlib
    rlmod = importlib.import_module("models.rl_gan_net")
    with open(os.path.join(REF, "configs", "config_quick.yaml")) as f:
        cfg = yaml.safe_load(f)

    def numeric(o):        # PyYAML reads 1e-4 as a string; the reference's trainer patches that (train_rl_gan_net.py:72-101)
        if isinstance(o, dict):
            return {k: numeric(v) for k, v in o.items()}
        if isinstance(o, list):
            return [numeric(v) for v in o]
        if isinstance(o, str):
            try:
                return float(o)
            except ValueError:
                return o
        return o
    cfg = numeric(cfg)
    cfg["training"]["device"] = "cpu"
    cfg["model"]["autoencoder"].update(latent_dim=32, num_points=256, encoder_dims=[64, 128, 64], decoder_dims=[64, 768])
    cfg["model"]["lgan"].update(z_dim=2, latent_dim=32, generator_dims=[48, 32], discriminator_dims=[32, 16, 1])
    cfg["model"]["rl_agent"].update(state_dim=32, action_dim=2, hidden_dims=[16, 16, 16, 16])
    torch.manual_seed(123)
    net = rlmod.RLGANNet(cfg)
    O.randomize_bn(net, seed=55)
    net.eval()
    env = rlmod.RLGANNetEnvironment(net, None)
    E = 6
    incomplete = O.make_clouds(E, 180, "sphere", seed=1100)
    complete = O.make_clouds(E, 256, "sphere", seed=1101)
    actions = torch.randn(E, 2, generator=torch.Generator().manual_seed(1102)).numpy().astype(np.float32)
    states, next_states, rewards = [], [], []
    for e in range(E):
        batch = {"incomplete": incomplete[e:e + 1], "complete": complete[e:e + 1]}
        states.append(env.reset(batch))                                        # rl_gan_net.py:279-297
        ns, r, done, info = env.step(actions[e])                               # :299-339
        assert done
        next_states.append(ns)
        rewards.append(r)
    ev_out = {"E": np.int32(E), "incomplete": incomplete.numpy(), "complete": complete.numpy(), "actions": actions,
              "states": np.stack(states), "next_states": np.stack(next_states), "rewards": np.array(rewards, np.float64),
              "weights": np.array([net.reward_function.w_chamfer, net.reward_function.w_gfv, net.reward_function.w_discriminator]),
              "ae_dims": np.array([32, 256], np.int32)}
    for prefix, mod in (("ae", net.autoencoder), ("lgan", net.latent_gan)):
        sd = mod.state_dict()
        ev_out[f"{prefix}_keys"] = np.array(list(sd.keys()))
        for name, v in sd.items():
            ev_out[f"{prefix}_sd_{name}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "environment_ref.npz"), **ev_out)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
