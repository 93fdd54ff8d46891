"""Profiling driver: the eval-mode encoder trunk at the reference's own dims [64,128,128,256,128] (configs/config.yaml:11).
    python tools/run_encoder_cfg.py [reps] [B] [N] [precision: fp32x | bf16_layers | fp32]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import gan_rl_3d_b200 as rlg  # noqa: E402
from oracle import oracle as O  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
N = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
precision = sys.argv[4] if len(sys.argv) > 4 else "fp32x"
torch.manual_seed(0)
enc = O.RefEncoderPort(3, 128, [64, 128, 128, 256, 128])
O.randomize_bn(enc, 0)
enc = enc.eval().cuda()
layers = rlg.fold_trunk(enc.point_mlp)
x = O.make_clouds(B, N, "sphere", 5).cuda()
for _ in range(reps):
    pooled, _ = rlg.encoder_pool(x, layers, precision=precision)
torch.cuda.synchronize()
print("ok", float(pooled.sum()))
