"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU checkers for the hot path of phanich004/GAN-RL_3D (batched Chamfer distance + PointNet encoder).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under gan-rl_3d_b200/ does.

Oracles (citations into /root/reference):
  O_ref_port   `ref_port_*`     the reference's own op sequence restated in our words
                                (utils/losses.py:29-37,54-59,75; models/autoencoder.py:20-76).
                                tests/test_oracle.py asserts it is bit-identical to the real reference
                                functions whenever /root/reference is mounted (build container), and
                                against tests/golden/ everywhere.
  O_direct     `chamfer_direct` plain-C direct-difference restatement (oracle/chamfer_oracle.c), the
                                branch torch.cdist itself takes for N,M <= 25; bit-exact target of the
                                CUDA kernel for distances and indices.
  O_f64        `chamfer_f64`    the same in float64 = truth for tolerances.
  O_enc        `RefEncoderPort` stock Conv1d/BatchNorm1d/ReLU/Linear stack in eval mode.

Parity pinning: the reference has no golden vectors for this path; the pins are outputs of the real
reference functions run in the build container (tests/golden/gen_golden.py -> tests/golden/*.npz).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TIE_SQUARED = 0
TIE_FAITHFUL = 1


def build(force: bool = False) -> str:
    """Compile oracle/chamfer_oracle.c -> oracle/liborc.so with the committed Makefile."""
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "chamfer_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liborc.so"], check=True, capture_output=True)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        dp = ctypes.POINTER(ctypes.c_double)
        L.orc_chamfer_fwd.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      fp, fp, ip, ip]
        L.orc_chamfer_fwd.restype = ctypes.c_int
        L.orc_chamfer_means.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, fp]
        L.orc_chamfer_means.restype = ctypes.c_int
        L.orc_chamfer_bwd.argtypes = [fp, fp, fp, fp, ip, ip, fp, fp,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp]
        L.orc_chamfer_bwd.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


# --------------------------------------------------------------------------------------------
# O_direct: C restatement
# --------------------------------------------------------------------------------------------
def chamfer_direct(pc1, pc2, tie_rule: int = TIE_SQUARED):
    """Direct-form nearest neighbours in fp32.  Returns (d1 (B,N), d2 (B,M), i1, i2) numpy arrays.

    Follows utils/losses.py:29-33 with cdist in its direct mode."""
    a, b = _f32(pc1), _f32(pc2)
    B, N, _ = a.shape
    M = b.shape[1]
    d1 = np.empty((B, N), np.float32)
    d2 = np.empty((B, M), np.float32)
    i1 = np.empty((B, N), np.int32)
    i2 = np.empty((B, M), np.int32)
    rc = lib().orc_chamfer_fwd(_ptr(a, ctypes.c_float), _ptr(b, ctypes.c_float), B, N, M, tie_rule,
                               _ptr(d1, ctypes.c_float), _ptr(d2, ctypes.c_float),
                               _ptr(i1, ctypes.c_int32), _ptr(i2, ctypes.c_int32))
    if rc != 0:
        raise ValueError(f"orc_chamfer_fwd rc={rc}")
    return d1, d2, i1, i2


def chamfer_means(d1: np.ndarray, d2: np.ndarray):
    """utils/losses.py:36-37 (double accumulation, rounded to fp32)."""
    return (d1.astype(np.float64).mean(axis=1).astype(np.float32),
            d2.astype(np.float64).mean(axis=1).astype(np.float32))


def chamfer_bwd_truth(pc1, pc2, d1, d2, i1, i2, g1, g2):
    """Closed-form gradient of (mean1, mean2) for the given indices, accumulated in float64."""
    a, b = _f32(pc1), _f32(pc2)
    B, N, _ = a.shape
    M = b.shape[1]
    d1, d2 = _f32(d1), _f32(d2)
    i1 = np.ascontiguousarray(i1, np.int32)
    i2 = np.ascontiguousarray(i2, np.int32)
    g1, g2 = _f32(g1), _f32(g2)
    ga = np.empty((B, N, 3), np.float64)
    gb = np.empty((B, M, 3), np.float64)
    rc = lib().orc_chamfer_bwd(_ptr(a, ctypes.c_float), _ptr(b, ctypes.c_float),
                               _ptr(d1, ctypes.c_float), _ptr(d2, ctypes.c_float),
                               _ptr(i1, ctypes.c_int32), _ptr(i2, ctypes.c_int32),
                               _ptr(g1, ctypes.c_float), _ptr(g2, ctypes.c_float), B, N, M,
                               _ptr(ga, ctypes.c_double), _ptr(gb, ctypes.c_double))
    if rc != 0:
        raise ValueError(f"orc_chamfer_bwd rc={rc}")
    return ga, gb


def chamfer_bwd_fixed_point(pc1, pc2, d1, d2, i1, i2, g1, g2, scale1: float = 1.0, scale2: float = 1.0):
    """The reproducible backward (rlg_chamfer_bwd_det) restated operation for operation, so its result is defined bit for
    bit: fp32 terms u = (own - partner) * (w / d) with w = g[b]*scale/n, each partner term rounded to a 64-bit integer
    number of quanta 2^(ilogb|w| - s), s = min(40, 61 - ceil(log2 n)), integer sums, and ONE rounding of
    own + sum * quantum to fp32.  (A closed form of autograd's MinBackward0 / EuclideanDistBackward0 / MeanBackward
    behind utils/losses.py:29-37 for the saved arg-min indices; terms vanish where d == 0.)"""
    a, b = _f32(pc1), _f32(pc2)
    B, N, _ = a.shape
    M = b.shape[1]
    out = []
    terms = []
    for own, oth, d, idx, g, scale, n in ((a, b, _f32(d1), i1, g1, scale1, N), (b, a, _f32(d2), i2, g2, scale2, M)):
        w = (np.zeros(B, np.float32) if g is None else _f32(g) * np.float32(scale)) / np.float32(n)        # fp32 ops, in this order
        idx = np.asarray(idx, np.int64)
        partner = np.take_along_axis(oth, idx[:, :, None], axis=1)
        live = (d != 0) & (w != 0)[:, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            s = (w[:, None] / d).astype(np.float32)
            u = ((own - partner) * s[:, :, None]).astype(np.float32)
        u = np.where(live[:, :, None], u, np.float32(0))
        lg = int(np.ceil(np.log2(n))) if n > 1 else 0
        terms.append((u, idx, w, min(40, 61 - lg)))
    for k, (own_u, _, _, _) in enumerate(terms):
        u_o, idx_o, w_o, s_o = terms[1 - k]                 # the partner terms of these rows come from the other direction
        rows = own_u.shape[1]
        res = np.empty_like(own_u)
        for p in range(B):
            if w_o[p] == 0 or not np.isfinite(w_o[p]):
                part = np.zeros((rows, 3)) if w_o[p] == 0 else np.full((rows, 3), np.nan)
            else:
                e = int(np.frexp(np.float64(abs(w_o[p])))[1]) - 1                      # ilogb
                q = np.rint(-u_o[p].astype(np.float64) * np.ldexp(1.0, s_o - e)).astype(np.int64)
                acc = np.zeros((rows, 3), np.int64)
                np.add.at(acc, idx_o[p], q)
                part = acc.astype(np.float64) * np.ldexp(1.0, e - s_o)
            res[p] = (own_u[p].astype(np.float64) + part).astype(np.float32)
        out.append(res)
    return out[0], out[1]


# --------------------------------------------------------------------------------------------
# O_f64 (truth) and the torch direct-mode cross-check
# --------------------------------------------------------------------------------------------
def chamfer_f64(pc1, pc2):
    """Float64 direct-mode distances: (d1, d2, i1, i2, D) as torch tensors; D only for small inputs."""
    a = torch.as_tensor(pc1).double()
    b = torch.as_tensor(pc2).double()
    D = torch.cdist(a, b, p=2, compute_mode="donot_use_mm_for_euclid_dist")
    d1, i1 = torch.min(D, dim=2)
    d2, i2 = torch.min(D, dim=1)
    return d1, d2, i1, i2, D


def chamfer_torch_direct(pc1, pc2):
    """fp32 torch.cdist forced into its direct mode (the branch utils/losses.py:29 takes for N,M<=25)."""
    a = torch.as_tensor(pc1).float()
    b = torch.as_tensor(pc2).float()
    D = torch.cdist(a, b, p=2, compute_mode="donot_use_mm_for_euclid_dist")
    d1, i1 = torch.min(D, dim=2)
    d2, i2 = torch.min(D, dim=1)
    return d1, d2, i1, i2


# --------------------------------------------------------------------------------------------
# O_ref_port: the reference's op sequence as written
# --------------------------------------------------------------------------------------------
def ref_port_chamfer_l2(pc1: torch.Tensor, pc2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """utils/losses.py:13-39: cdist -> min over each axis -> mean over points."""
    dm = torch.cdist(pc1, pc2, p=2)
    near12 = torch.min(dm, dim=2)[0]
    near21 = torch.min(dm, dim=1)[0]
    return torch.mean(near12, dim=1), torch.mean(near21, dim=1)


def ref_port_chamfer(pc1, pc2, bidirectional: bool = True) -> torch.Tensor:
    """utils/losses.py:42-59."""
    m1, m2 = ref_port_chamfer_l2(pc1, pc2)
    return (m1 + m2) / 2.0 if bidirectional else m1


def ref_port_chamfer_loss(pred, target, bidirectional: bool = True) -> torch.Tensor:
    """utils/losses.py:62-75 (ChamferLoss.forward)."""
    return torch.mean(ref_port_chamfer(pred, target, bidirectional))


def ref_port_argmins(pc1: torch.Tensor, pc2: torch.Tensor):
    """The indices autograd saves for MinBackward0 in the as-written reference."""
    dm = torch.cdist(pc1, pc2, p=2)
    return torch.min(dm, dim=2)[1], torch.min(dm, dim=1)[1]


class RefEncoderPort(nn.Module):
    """models/autoencoder.py:13-76 restated: [Conv1d(k=1) -> BatchNorm1d -> ReLU] x L over (B,C,N),
    max over N, Linear -> BatchNorm1d -> ReLU.  Module/attribute names match the reference so a
    reference state_dict loads unchanged."""

    def __init__(self, input_dim: int = 3, latent_dim: int = 128,
                 hidden_dims: Sequence[int] = (64, 128, 128, 256, 128)):
        super().__init__()
        self.input_dim, self.latent_dim, self.hidden_dims = input_dim, latent_dim, list(hidden_dims)
        seq: List[nn.Module] = []
        c_in = input_dim
        for c_out in self.hidden_dims:
            seq += [nn.Conv1d(c_in, c_out, 1), nn.BatchNorm1d(c_out), nn.ReLU(inplace=True)]
            c_in = c_out
        self.point_mlp = nn.Sequential(*seq)
        self.global_mlp = nn.Sequential(nn.Linear(c_in, latent_dim), nn.BatchNorm1d(latent_dim),
                                        nn.ReLU(inplace=True))

    def pooled(self, x: torch.Tensor) -> torch.Tensor:
        return torch.max(self.point_mlp(x.transpose(2, 1)), dim=2)[0]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.global_mlp(self.pooled(x))


class RefDecoderPort(nn.Module):
    """models/autoencoder.py:79-129 restated: Linear+BatchNorm1d+ReLU blocks, a final Linear, view (B, num_points, 3)."""

    def __init__(self, latent_dim: int = 128, num_points: int = 2048, hidden_dims=(256, 256, 6144)):
        super().__init__()
        seq, c_in = [], latent_dim
        for c in hidden_dims[:-1]:
            seq += [nn.Linear(c_in, c), nn.BatchNorm1d(c), nn.ReLU(inplace=True)]
            c_in = c
        seq.append(nn.Linear(c_in, hidden_dims[-1]))
        self.mlp = nn.Sequential(*seq)
        self.num_points = num_points

    def forward(self, gfv):
        return self.mlp(gfv).view(-1, self.num_points, 3)


class _Seq(nn.Module):
    def __init__(self, name: str, seq: nn.Sequential):
        super().__init__()
        setattr(self, name, seq)
        self._name = name

    def forward(self, x):
        return getattr(self, self._name)(x)


class RefLatentGANPort(nn.Module):
    """models/latent_gan.py restated with the same module tree (state_dict keys generator.generator.N.*,
    discriminator.discriminator.N.*): generator = Linear+BatchNorm1d+ReLU blocks, Linear, Tanh (:14-62); improved
    discriminator = spectral-norm Linear + LayerNorm + LeakyReLU(0.2) + Dropout(0.3) blocks, spectral-norm Linear (:143-203)."""

    def __init__(self, z_dim=1, latent_dim=128, generator_dims=(256, 512, 512, 256, 128), discriminator_dims=(128, 256, 512, 256, 1)):
        super().__init__()
        seq, c_in = [], z_dim
        for c in generator_dims[:-1]:
            seq += [nn.Linear(c_in, c), nn.BatchNorm1d(c), nn.ReLU(inplace=True)]
            c_in = c
        seq += [nn.Linear(c_in, generator_dims[-1]), nn.Tanh()]
        self.generator = _Seq("generator", nn.Sequential(*seq))
        seq, c_in = [], latent_dim
        for c in discriminator_dims[:-1]:
            seq += [nn.utils.spectral_norm(nn.Linear(c_in, c)), nn.LayerNorm(c), nn.LeakyReLU(0.2, inplace=True), nn.Dropout(0.3)]
            c_in = c
        seq.append(nn.utils.spectral_norm(nn.Linear(c_in, discriminator_dims[-1])))
        self.discriminator = _Seq("discriminator", nn.Sequential(*seq))

    def generate(self, z):
        return self.generator(z)

    def discriminate(self, gfv):
        return self.discriminator(gfv)


# --------------------------------------------------------------------------------------------
# input pipeline (utils/dataset.py:135-187,252-297,393-421; utils/data_utils.py:15-60,74-142) with EXPLICIT random draws:
# the reference pulls its draws from np.random / torch's global RNG inside each function; here they are arguments, so the
# device pipeline and this restatement can be fed the same ones
# --------------------------------------------------------------------------------------------
def percentile_parts(n: int, removal_ratio: float):
    """(k, gamma) of np.percentile(distances, removal_ratio * 100) for n float64 values, method 'linear':
    virtual index (n - 1) * (q / 100), k = floor, gamma = fractional part (numpy/lib/_function_base_impl.py)."""
    q = removal_ratio * 100
    vi = (n - 1) * np.true_divide(q, 100.0)
    k = int(np.floor(vi))
    return k, float(vi - k)


def lerp_numpy(a: float, b: float, t: float) -> float:
    """numpy's _lerp in float64: a + (b - a) * t, and b - (b - a) * (1 - t) where t >= 0.5."""
    a, b, t = np.float64(a), np.float64(b), np.float64(t)
    d = b - a
    return float(b - d * (1 - t)) if t >= 0.5 else float(a + d * t)


def ref_port_create_incomplete(complete_pc: np.ndarray, draws: dict) -> np.ndarray:
    """utils/dataset.py:252-276 with its draws made explicit.  draws: {'method': 0, 'keep_idx': int array} (random subset,
    kept in the drawn order) or {'method': 1, 'center': int, 'ratio': float} (points outside a sphere around a point)."""
    if draws["method"] == 0:
        return complete_pc[np.asarray(draws["keep_idx"])]
    center = complete_pc[int(draws["center"])]
    distances = np.linalg.norm(complete_pc - center, axis=1)
    radius = np.percentile(distances, draws["ratio"] * 100)
    return complete_pc[distances > radius]


def ref_port_augment(pc: np.ndarray, rot=None, noise=None, scale=None) -> np.ndarray:
    """utils/dataset.py:278-297: float32 tensor; optional rotation pc @ R.T (data_utils.py:118), optional additive noise
    (already clipped, data_utils.py:140-142), optional scale."""
    t = torch.FloatTensor(pc)
    if rot is not None:
        t = t @ torch.tensor(rot, dtype=t.dtype).T
    if noise is not None:
        t = t + torch.as_tensor(noise, dtype=t.dtype)
    if scale is not None:
        t = t * scale
    return t.numpy()


def ref_port_normalize(pc: np.ndarray) -> np.ndarray:
    """utils/data_utils.py:15-60 for one (N, 3) cloud: centre on the centroid, divide by the largest norm (float32)."""
    t = torch.from_numpy(np.asarray(pc)).float()
    c = t - torch.mean(t, dim=0, keepdim=True)
    scale = torch.max(torch.norm(c, dim=1))
    return (c / scale if scale > 0 else c).numpy()


def ref_port_pad(pcs, pad_idx) -> np.ndarray:
    """shapenet_collate_fn (utils/dataset.py:393-421): incomplete clouds padded to the longest of the batch by repeating
    points; pad slot s of cloud b repeats point pad_idx[b][s] (the reference draws torch.randint(0, len_b))."""
    m = max(p.shape[0] for p in pcs)
    out = np.zeros((len(pcs), m, 3), np.float32)
    for b, p in enumerate(pcs):
        n = p.shape[0]
        out[b, :n] = p
        if n < m:
            out[b, n:] = p[np.asarray(pad_idx[b][:m - n]) % n]
    return out


def randomize_bn(module: nn.Module, seed: int = 0) -> None:
    """Non-trivial BatchNorm statistics/affine (a fresh BN is identity-like and hides folding bugs).
    SURVEY.md 8d: running_mean~N(0,.5), running_var~U(.3,2), gamma~N(1,.5), beta~N(0,.3)."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, nn.BatchNorm1d):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.5)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) * 1.7 + 0.3)
                m.weight.copy_(1.0 + torch.randn(m.num_features, generator=g) * 0.5)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.3)


# --------------------------------------------------------------------------------------------
# seeded synthetic inputs -- generated on the CPU so every backend sees the same bits
# --------------------------------------------------------------------------------------------
def make_clouds(B: int, N: int, kind: str = "sphere", seed: int = 1234) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    if kind == "sphere":
        x = torch.randn(B, N, 3, generator=g)
        x = x / x.norm(dim=2, keepdim=True).clamp_min(1e-12)
    elif kind == "uniform":
        x = torch.rand(B, N, 3, generator=g) * 2.0 - 1.0
    else:
        raise ValueError(kind)
    return x.contiguous()


def pad_with_duplicates(pc: torch.Tensor, frac: float = 0.25, seed: int = 99) -> torch.Tensor:
    """Overwrite the last `frac` of each cloud with copies of earlier points, as the reference's
    collate does for ragged partial clouds (utils/dataset.py:398-421) -> exact-tie stress."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    B, N, _ = pc.shape
    keep = N - int(N * frac)
    out = pc.clone()
    for b in range(B):
        src = torch.randint(0, keep, (N - keep,), generator=g)
        out[b, keep:] = pc[b, src]
    return out


# --------------------------------------------------------------------------------------------
# parity-contract helpers
# --------------------------------------------------------------------------------------------
def idx_mismatches_are_exact_ties(pc_q, pc_c, idx_ours, idx_ref) -> Tuple[int, int]:
    """For every query where idx_ours != idx_ref, check both candidates are at the SAME fp32 direct
    distance (an exact tie in the reference's sqrt-ed matrix).  Returns (n_mismatch, n_not_tie)."""
    q, c = _f32(pc_q), _f32(pc_c)
    io = np.asarray(idx_ours).astype(np.int64)
    ir = np.asarray(idx_ref).astype(np.int64)
    bb, ii = np.nonzero(io != ir)
    if len(bb) == 0:
        return 0, 0

    def dist(b, i, j):
        d = (q[b, i] - c[b, j]).astype(np.float32)
        t = np.float32(d[..., 0] * d[..., 0])
        t = (t.astype(np.float64) + d[..., 1].astype(np.float64) ** 2).astype(np.float32)  # fmaf
        t = (t.astype(np.float64) + d[..., 2].astype(np.float64) ** 2).astype(np.float32)
        return np.sqrt(t, dtype=np.float32)

    da, db = dist(bb, ii, io[bb, ii]), dist(bb, ii, ir[bb, ii])
    return len(bb), int(np.count_nonzero(da != db))


def idx_mismatches_are_near_ties(pc_q, pc_c, idx_ours, idx_ref, ulps: float = 8.0) -> Tuple[int, int]:
    """vs the as-written (matmul-path) reference: a disagreement must be a near-tie in float64:
    |d64(i,ours) - d64(i,ref)| <= ulps * eps32 * (|x|^2+|y|^2) / max(d,eps)  (SURVEY.md 8c)."""
    q = _f32(pc_q).astype(np.float64)
    c = _f32(pc_c).astype(np.float64)
    io = np.asarray(idx_ours).astype(np.int64)
    ir = np.asarray(idx_ref).astype(np.int64)
    bb, ii = np.nonzero(io != ir)
    if len(bb) == 0:
        return 0, 0
    x = q[bb, ii]
    ya, yb = c[bb, io[bb, ii]], c[bb, ir[bb, ii]]
    ta = ((x - ya) ** 2).sum(-1)
    tb = ((x - yb) ** 2).sum(-1)
    scale = (x ** 2).sum(-1) + np.maximum((ya ** 2).sum(-1), (yb ** 2).sum(-1))
    eps32 = float(np.finfo(np.float32).eps)
    bad = np.abs(ta - tb) > ulps * eps32 * scale          # compared on squared distances
    return len(bb), int(np.count_nonzero(bad))


def rel_err(a, b, floor: float = 0.0) -> float:
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def rowwise_rel_err(g, g_truth, floor_frac: float = 1e-3) -> float:
    """Row-wise gradient error: |g_i - t_i|_2 / max(|t_i|_2, floor_frac * max_i |t_i|_2)."""
    g = np.asarray(g, np.float64).reshape(-1, 3)
    t = np.asarray(g_truth, np.float64).reshape(-1, 3)
    tn = np.linalg.norm(t, axis=1)
    floor = floor_frac * (tn.max() if tn.size else 1.0)
    return float(np.max(np.linalg.norm(g - t, axis=1) / np.maximum(tn, max(floor, 1e-300))))


def gfv_close(a, b, rel: float, floor_frac: float = 1e-2) -> Tuple[bool, float]:
    """Encoder tolerance with the abs-floor rule of SURVEY.md 7.2-6:
    |a-b| <= rel * max(|b|, floor_frac * |b|_inf).  floor_frac is 1e-2 for the fp32 path; the bf16 tensor-core
    path is judged with floor_frac = 0.25 (bf16 operands carry 2^-9 relative rounding, so a K=64..128 dot product
    has an absolute error of ~1e-3 |b|_inf whatever the size of the entry) plus a norm-wise bound."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    floor = floor_frac * np.abs(b).max()
    err = np.abs(a - b) / np.maximum(np.abs(b), floor)
    return bool(err.max() <= rel), float(err.max())
