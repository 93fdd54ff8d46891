"""Accuracy and speed of every encoder kernel family on a few layer lists (diagnostic, not a benchmark value).
    python tools/encoder_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import gan_rl_3d_b200 as rlg  # noqa: E402
from oracle import oracle as O  # noqa: E402

dev = torch.device("cuda:0")
DIMS = [[64, 128, 128, 256, 128], [64, 128, 1024], [64, 256, 256, 192], [64, 64]]
for dims in DIMS:
    torch.manual_seed(len(dims))
    enc = O.RefEncoderPort(3, 32, dims)
    O.randomize_bn(enc, 3)
    enc.eval()
    worst = {}
    for B, N, seed in ((2, 2048, 1), (5, 1300, 2), (9, 127, 3), (3, 128, 4), (4, 700, 5)):
        x = O.make_clouds(B, N, "sphere", 100 + seed)
        with torch.no_grad():
            want = enc.double().pooled(x.double()).float().numpy()
        encd = enc.float().to(dev)
        layers = rlg.fold_trunk(encd.point_mlp)
        for prec in ("fp32", "fp32x", "bf16_layers"):
            got = rlg.encoder_pool(x.to(dev), layers, precision=prec)[0].cpu().numpy()
            err = O.gfv_close(got, want, 1.0, 1e-2 if prec != "bf16_layers" else 0.25)[1]
            worst[prec] = max(worst.get(prec, 0.0), err)
        enc = encd.cpu()
    # speed at the cfg3 batch
    B, N = 256, 2048
    xs = [O.make_clouds(B, N, "sphere", 300 + k).to(dev) for k in range(8)]
    encd = enc.float().to(dev)
    layers = rlg.fold_trunk(encd.point_mlp)
    speed = {}
    for prec in ("fp32", "fp32x", "bf16_layers") + (("bf16",) if rlg.encoder.fused_tc_supported(layers) else ()):
        packed = None
        if prec in ("fp32x", "bf16_layers"):
            packed = rlg.pack_gemm(layers, 2 if prec == "fp32x" else 1)
        elif prec == "bf16":
            packed = rlg.pack_bf16(layers)
        for k in range(3):
            rlg.encoder_pool(xs[k], layers, precision=prec, packed=packed)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 6 if prec == "fp32" else 40
        e0.record()
        for k in range(reps):
            rlg.encoder_pool(xs[k % 8], layers, precision=prec, packed=packed)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        speed[prec] = f"{B / ms * 1e3:9.0f} clouds/s ({ms * 1e3:7.1f} us)"
    print(dims, "max err (floor 1e-2 / 0.25):", {k: f"{v:.2e}" for k, v in worst.items()})
    print("    B=256 N=2048:", speed)
