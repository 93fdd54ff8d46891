"""CPU: pins the oracle (oracle/) against the golden vectors generated from the real reference, and against
the reference itself where it is mounted.  The oracle is the checker of every GPU parity test."""
import numpy as np
import pytest
import torch

from oracle import oracle as O


def _small_cases(g):
    for k in range(int(g["small_count"])):
        yield k, torch.from_numpy(g[f"small{k}_pc1"]), torch.from_numpy(g[f"small{k}_pc2"])


def _big_cases(g):
    for k in range(int(g["big_count"])):
        B, N, M, s1, s2, sd = [int(v) for v in g[f"big{k}_meta"]]
        kind = str(g[f"big{k}_kind"])
        pc1 = O.make_clouds(B, N, kind, seed=s1)
        pc2 = O.make_clouds(B, M, kind, seed=s2)
        if sd >= 0:
            pc2 = O.pad_with_duplicates(pc2, 0.25, seed=sd)
        yield k, pc1, pc2


def test_direct_oracle_is_bit_exact_on_reference_direct_path(golden_chamfer):
    """N,M <= 25: the reference's own cdist takes the direct branch -> distances and indices bit-equal."""
    g = golden_chamfer
    for k, pc1, pc2 in _small_cases(g):
        d1, d2, i1, i2 = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
        assert np.array_equal(d1, g[f"small{k}_d1"]), k
        assert np.array_equal(d2, g[f"small{k}_d2"]), k
        assert np.array_equal(i1, g[f"small{k}_i1"]), k
        assert np.array_equal(i2, g[f"small{k}_i2"]), k
        # squared-distance tie rule: same distances; indices may differ only on exact ties of the sqrt-ed value
        s1, s2, j1, j2 = O.chamfer_direct(pc1, pc2, O.TIE_SQUARED)
        assert np.array_equal(s1, d1) and np.array_equal(s2, d2)
        assert O.idx_mismatches_are_exact_ties(pc1, pc2, j1, i1)[1] == 0
        assert O.idx_mismatches_are_exact_ties(pc2, pc1, j2, i2)[1] == 0
        m1, m2 = O.chamfer_means(d1, d2)
        np.testing.assert_allclose(m1, g[f"small{k}_dist1"], rtol=2e-7, atol=0)
        np.testing.assert_allclose(m2, g[f"small{k}_dist2"], rtol=2e-7, atol=0)
        np.testing.assert_allclose((m1.astype(np.float64) + m2) / 2, g[f"small{k}_cd"], rtol=3e-7)
        np.testing.assert_allclose(m1, g[f"small{k}_cd_uni"], rtol=2e-7)


def test_backward_closed_form_matches_reference_autograd(golden_chamfer):
    g = golden_chamfer
    for k, pc1, pc2 in _small_cases(g):
        B, N, M = pc1.shape[0], pc1.shape[1], pc2.shape[1]
        d1, d2, i1, i2 = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
        # ChamferLoss = mean_b (mean1+mean2)/2 -> upstream of mean1, mean2 is 1/(2B)
        up = np.full((B,), 0.5 / B, np.float32)
        ga, gb = O.chamfer_bwd_truth(pc1, pc2, d1, d2, i1, i2, up, up)
        assert O.rowwise_rel_err(g[f"small{k}_g1"], ga) < 2e-6, k
        assert O.rowwise_rel_err(g[f"small{k}_g2"], gb) < 2e-6, k


def test_port_matches_golden_everywhere(golden_chamfer):
    """The restated op sequence reproduces the reference outputs stored in the fixtures (same torch build:
    bit-exact; the tolerance only absorbs a different BLAS thread split on another host)."""
    g = golden_chamfer
    torch.set_num_threads(1)
    for k, pc1, pc2 in list(_small_cases(g)):
        m1, m2 = O.ref_port_chamfer_l2(pc1, pc2)
        assert np.array_equal(m1.numpy(), g[f"small{k}_dist1"]) and np.array_equal(m2.numpy(), g[f"small{k}_dist2"])
        assert np.float32(O.ref_port_chamfer_loss(pc1, pc2).item()) == g[f"small{k}_loss"]
    for k, pc1, pc2 in _big_cases(g):
        m1, m2 = O.ref_port_chamfer_l2(pc1, pc2)
        np.testing.assert_allclose(m1.numpy(), g[f"big{k}_dist1"], rtol=1e-6)
        np.testing.assert_allclose(m2.numpy(), g[f"big{k}_dist2"], rtol=1e-6)


def test_port_is_bit_identical_to_mounted_reference(ref_losses):
    torch.set_num_threads(1)
    for (B, N, M, kind) in [(3, 16, 25, "sphere"), (2, 300, 257, "uniform"), (1, 1024, 700, "sphere")]:
        pc1, pc2 = O.make_clouds(B, N, kind, 1), O.make_clouds(B, M, kind, 2)
        r1, r2 = ref_losses.chamfer_distance_l2(pc1, pc2)
        p1, p2 = O.ref_port_chamfer_l2(pc1, pc2)
        assert torch.equal(r1, p1) and torch.equal(r2, p2)
        assert torch.equal(ref_losses.chamfer_distance(pc1, pc2, False), O.ref_port_chamfer(pc1, pc2, False))
        assert torch.equal(ref_losses.ChamferLoss()(pc1, pc2), O.ref_port_chamfer_loss(pc1, pc2))


def test_direct_oracle_vs_noisy_reference_contract(golden_chamfer):
    """N or M > 25: the as-written reference uses the matmul expansion.  Contract: index
    disagreements are near-ties in float64; per-pair means agree to max(1e-5, the reference's own error)."""
    g = golden_chamfer
    for k, pc1, pc2 in _big_cases(g):
        d1, d2, i1, i2 = O.chamfer_direct(pc1, pc2, O.TIE_SQUARED)
        n1, bad1 = O.idx_mismatches_are_near_ties(pc1, pc2, i1, g[f"big{k}_i1"])
        n2, bad2 = O.idx_mismatches_are_near_ties(pc2, pc1, i2, g[f"big{k}_i2"])
        assert bad1 == 0 and bad2 == 0, (k, n1, bad1, n2, bad2)
        f1, f2, j1, j2, _ = O.chamfer_f64(pc1, pc2)
        # ours vs float64 truth: indices equal except exact fp32 ties, distances to 1e-6
        assert O.idx_mismatches_are_exact_ties(pc1, pc2, i1, j1.numpy())[1] == 0
        assert O.idx_mismatches_are_exact_ties(pc2, pc1, i2, j2.numpy())[1] == 0
        assert O.rel_err(d1, f1.numpy(), floor=1e-30) < 1e-6 and O.rel_err(d2, f2.numpy(), floor=1e-30) < 1e-6
        m1, m2 = O.chamfer_means(d1, d2)
        t1, t2 = f1.mean(1).numpy(), f2.mean(1).numpy()
        ref_err = max(O.rel_err(g[f"big{k}_dist1"], t1), O.rel_err(g[f"big{k}_dist2"], t2))
        tol = max(1e-5, 2 * ref_err)
        assert O.rel_err(m1, g[f"big{k}_dist1"]) < tol and O.rel_err(m2, g[f"big{k}_dist2"]) < tol


def test_sqrt_collision_tie_rules():
    """sqrtf is many-to-one: two squared distances one ulp apart can collide after the square root.  The
    faithful rule (torch.min over the sqrt-ed matrix) then picks the LOWER index, the squared rule the
    smaller t; both return the same distance (SURVEY.md 7.1 step 2 trap)."""
    rng = np.random.default_rng(0)
    found = 0
    for _ in range(2000):
        # from the origin, (r,0,0) has t = fl(r*r) in [2,4) where ulp(t) = 2^-22; adding a second coordinate
        # s ~ 2^-11 lifts t by one ulp, and sqrt compresses that spacing below one ulp of the result.
        r = np.float32(rng.uniform(1.42, 1.99))
        s_ = np.float32(2.0 ** -11 * rng.uniform(0.8, 1.2))
        tb = np.float32(r * r)
        ta = np.float32(np.float64(s_) * np.float64(s_) + np.float64(tb))      # fmaf(s,s,t), exact in float64
        if not (ta == np.nextafter(tb, np.float32(8), dtype=np.float32)
                and np.sqrt(ta, dtype=np.float32) == np.sqrt(tb, dtype=np.float32)):
            continue
        pc1 = np.zeros((1, 1, 3), np.float32)
        pc2 = np.array([[[r, s_, 0], [r, 0, 0]]], np.float32)     # index 0 has the LARGER squared distance
        df, _, jf, _ = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
        ds, _, js, _ = O.chamfer_direct(pc1, pc2, O.TIE_SQUARED)
        td, _, tj, _ = O.chamfer_torch_direct(pc1, pc2)
        assert jf[0, 0] == 0 and js[0, 0] == 1 and int(tj[0, 0]) == 0
        assert df[0, 0] == ds[0, 0] == td.numpy()[0, 0]
        assert O.idx_mismatches_are_exact_ties(pc1, pc2, js, jf) == (1, 0)
        found += 1
        if found >= 5:
            break
    assert found >= 1


def test_encoder_port_matches_golden(golden_encoder):
    g = golden_encoder
    torch.set_num_threads(1)
    for k in range(int(g["enc_count"])):
        dims, latent = [int(v) for v in g[f"enc{k}_dims"]], int(g[f"enc{k}_latent"])
        torch.manual_seed(k)
        enc = O.RefEncoderPort(3, latent, dims)
        assert list(enc.state_dict().keys()) == [str(s) for s in g[f"enc{k}_keys"]]
        if k == 2:
            enc.load_state_dict({str(n): torch.from_numpy(g[f"enc{k}_sd_{n}"]) for n in g[f"enc{k}_keys"]})
        else:
            O.randomize_bn(enc, seed=10 + k)
            chk = float(sum(v.double().abs().sum().item() for v in enc.state_dict().values()))
            if abs(chk - float(g[f"enc{k}_state_checksum"])) > 1e-6 * abs(chk):
                pytest.skip("torch default init differs from the fixture's torch build")
        enc.eval()
        x = torch.from_numpy(g[f"enc{k}_x"])
        with torch.no_grad():
            np.testing.assert_allclose(enc.pooled(x).numpy(), g[f"enc{k}_pooled"], rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(enc(x).numpy(), g[f"enc{k}_gfv"], rtol=1e-6, atol=1e-7)


def test_encoder_port_is_reference(ref_autoencoder):
    torch.manual_seed(3)
    ref = ref_autoencoder.PointNetEncoder(3, 32, [16, 64])
    O.randomize_bn(ref, 5)
    port = O.RefEncoderPort(3, 32, [16, 64])
    port.load_state_dict(ref.state_dict())          # identical keys and shapes
    x = O.make_clouds(2, 77, "uniform", 9)
    for mode in ("eval", "train"):
        getattr(ref, mode)(), getattr(port, mode)()
        assert torch.equal(ref(x), port(x))


def test_reward_golden_is_the_reference_formula(golden_reward):
    """The reward fixture (generated from the reference's RewardFunction) equals the closed form the batched API
    implements, evaluated in float64 -- within the reference's own cdist noise."""
    g = golden_reward
    for k in range(int(g["rw_count"])):
        E, N, M, s1, s2 = [int(v) for v in g[f"rw{k}_meta"]]
        pred, target = O.make_clouds(E, N, str(g[f"rw{k}_kind"]), s1), O.make_clouds(E, M, str(g[f"rw{k}_kind"]), s2)
        f1, f2, *_ = O.chamfer_f64(pred, target)
        cd = (f1.mean(1) + f2.mean(1)).numpy() / 2.0
        gfv = ((g[f"rw{k}_pred_gfv"].astype(np.float64) - g[f"rw{k}_target_gfv"]) ** 2).mean(1)
        truth = -(100.0 * cd + 10.0 * gfv + 0.01 * (-g[f"rw{k}_disc"].astype(np.float64).reshape(E)))
        assert O.rel_err(g[f"rw{k}_rewards"], truth) < 5e-5


def test_environment_ports_reproduce_the_reference_episodes():
    """oracle.RefEncoderPort / RefDecoderPort / RefLatentGANPort loaded with the reference's weights replay the golden
    environment episodes (tests/golden/environment_ref.npz, generated by the reference's RLGANNetEnvironment) on the CPU."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "environment_ref.npz"))
    latent, points = (int(v) for v in g["ae_dims"])
    enc = O.RefEncoderPort(3, latent, [64, 128, 64])
    dec = O.RefDecoderPort(latent, points, [64, points * 3])
    ae_sd = {str(n): torch.from_numpy(g[f"ae_sd_{n}"]) for n in g["ae_keys"]}
    enc.load_state_dict({k[len("encoder."):]: v for k, v in ae_sd.items() if k.startswith("encoder.")})
    dec.load_state_dict({k[len("decoder."):]: v for k, v in ae_sd.items() if k.startswith("decoder.")})
    gan = O.RefLatentGANPort(2, latent, [48, 32], [32, 16, 1])
    gan.load_state_dict({str(n): torch.from_numpy(g[f"lgan_sd_{n}"]) for n in g["lgan_keys"]})
    enc.eval(), dec.eval(), gan.eval()
    w = g["weights"]
    with torch.no_grad():
        for e in range(int(g["E"])):
            inc, comp = torch.from_numpy(g["incomplete"][e:e + 1]), torch.from_numpy(g["complete"][e:e + 1])
            assert np.allclose(enc(inc)[0].numpy(), g["states"][e], rtol=0, atol=1e-6)
            clean = gan.generate(torch.from_numpy(g["actions"][e]).unsqueeze(0))
            assert np.allclose(clean[0].numpy(), g["next_states"][e], rtol=0, atol=1e-6)
            cd = O.ref_port_chamfer_loss(dec(clean), comp)
            mse = torch.nn.functional.mse_loss(clean, enc(comp))
            reward = -(w[0] * cd + w[1] * mse + w[2] * (-gan.discriminate(clean).mean()))
            assert abs(float(reward) - float(g["rewards"][e])) <= 1e-5 * abs(float(g["rewards"][e]))


def _ref_dataset_module():
    """utils/dataset.py of the mounted reference (it imports h5py, absent here and unused by these functions)."""
    import importlib
    import os
    import sys
    import types
    ref = os.environ.get("RLG_REFERENCE", "/root/reference")
    if not os.path.exists(os.path.join(ref, "utils", "dataset.py")):
        pytest.skip("reference checkout not present")
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if ref not in sys.path:
        sys.path.insert(0, ref)
    return importlib.import_module("utils.dataset")


def test_input_pipeline_ports_are_bit_identical_to_the_mounted_reference():
    """oracle.ref_port_create_incomplete / _augment / _normalize / _pad with the draws the reference would make under the same
    seeds == utils/dataset.py:252-297,393-421 and utils/data_utils.py:15-60 themselves."""
    ds = _ref_dataset_module()
    rng = np.random.default_rng(5)
    for trial in range(12):
        n = 2048 if trial % 3 else 777
        complete = rng.normal(size=(n, 3)) * rng.uniform(0.5, 2.0) + rng.normal(size=3)
        np.random.seed(100 + trial)
        want = ds.ShapeNetDataset._create_incomplete_pc(None, complete)                 # the reference, its own draws
        np.random.seed(100 + trial)                                                     # the same draws, made explicit
        ratio = np.random.uniform(0.2, 0.5)
        num_keep = int(n * (1 - ratio))
        if np.random.random() < 0.5:
            draws = {"method": 0, "keep_idx": np.random.choice(n, num_keep, replace=False)}
        else:
            draws = {"method": 1, "center": np.random.randint(n), "ratio": ratio}
        got = O.ref_port_create_incomplete(complete, draws)
        assert got.shape == want.shape and np.array_equal(got, want)
        if draws["method"] == 1:                                                        # and the restated percentile
            d = np.sort(np.linalg.norm(complete - complete[draws["center"]], axis=1))
            k, g = O.percentile_parts(n, ratio)
            assert O.lerp_numpy(d[k], d[min(k + 1, n - 1)], g) == np.percentile(d, ratio * 100)
        # augmentation: np.random for the coin flips / angles / scale, torch's RNG for the jitter noise
        np.random.seed(200 + trial)
        torch.manual_seed(300 + trial)
        want_aug = ds.ShapeNetDataset._augment_point_cloud(None, want)
        np.random.seed(200 + trial)
        torch.manual_seed(300 + trial)
        rot = noise = scale = None
        if np.random.random() < 0.5:
            rot = ds.random_rotation_matrix()
        if np.random.random() < 0.5:
            noise = torch.clamp(torch.normal(0, 0.01, size=torch.FloatTensor(want).shape), -0.05, 0.05)
        if np.random.random() < 0.3:
            scale = np.random.uniform(0.8, 1.2)
        got_aug = O.ref_port_augment(want, rot, noise, scale)
        assert np.array_equal(got_aug, want_aug)
        assert np.array_equal(O.ref_port_normalize(got_aug), ds.normalize_point_cloud(want_aug))
    # collate: the padded batch
    items = [{"incomplete_pc": torch.randn(n, 3)} for n in (50, 64, 37)]
    torch.manual_seed(9)
    want = ds.shapenet_collate_fn(items)["incomplete_pc"]
    torch.manual_seed(9)
    pad_idx = [torch.randint(0, it["incomplete_pc"].shape[0], (64 - it["incomplete_pc"].shape[0],)).numpy()
               if it["incomplete_pc"].shape[0] < 64 else np.zeros(0, np.int64) for it in items]
    got = O.ref_port_pad([it["incomplete_pc"].numpy() for it in items], pad_idx)
    assert np.array_equal(got, want.numpy())


def test_fixed_point_backward_restatement_matches_the_float64_closed_form():
    """oracle.chamfer_bwd_fixed_point (what the reproducible backward kernels are held to bit for bit) against the float64
    closed form, each pair judged on its own over upstream weights from 1e-30 to 1e20; zero weights give exact zeros."""
    B, N, M = 6, 300, 40
    pc1 = O.make_clouds(B, N, "sphere", 7)
    pc2 = 0.05 * O.make_clouds(B, M, "uniform", 8) + torch.tensor([1.5, 0.0, -0.5])       # many queries per partner
    d1, d2, i1, i2 = O.chamfer_direct(pc1, pc2, O.TIE_FAITHFUL)
    g1 = np.array([1.0, -3e-30, 7e19, 1e-3, 0.0, 0.37], np.float32)
    g2 = np.array([0.5, 2e-30, -1e20, 0.0, 4.0, 1.0], np.float32)
    ga, gb = O.chamfer_bwd_fixed_point(pc1, pc2, d1, d2, i1, i2, g1, g2)
    ta, tb = O.chamfer_bwd_truth(pc1, pc2, d1, d2, i1, i2, g1, g2)
    assert ga.dtype == np.float32 and np.bincount(i1[0], minlength=M).max() >= 20
    for p in range(B):
        for got, want in ((ga[p], ta[p]), (gb[p], tb[p])):
            if not np.any(want):
                assert not got.any()
            else:
                assert O.rowwise_rel_err(got, want) < 1e-6
    # the order in which the terms are accumulated changes nothing (integer sums): renumber the points of cloud 1
    perm = np.random.default_rng(3).permutation(N)
    inv = np.argsort(perm)
    ga2, gb2 = O.chamfer_bwd_fixed_point(pc1[:, perm], pc2, d1[:, perm], d2, i1[:, perm], inv[i2].astype(np.int32), g1, g2)
    assert np.array_equal(ga2, ga[:, perm]) and np.array_equal(gb2, gb)
