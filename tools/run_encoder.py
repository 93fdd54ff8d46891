"""Profiling driver: a few encoder-trunk launches at BASELINE config 3 (B=256, N=2048, 3->64->128->1024).
    python tools/run_encoder.py [reps] [B] [N] [precision]
Used under `ncu` (see profiles/README.md); prints nothing that counts as a benchmark value."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import gan_rl_3d_b200 as rlg  # noqa: E402
from oracle import oracle as O  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
N = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
precision = sys.argv[4] if len(sys.argv) > 4 else "bf16"
torch.manual_seed(0)
enc = O.RefEncoderPort(3, 128, [64, 128, 1024])
O.randomize_bn(enc, 0)
enc = enc.eval().cuda()
layers = rlg.fold_trunk(enc.point_mlp)
packed = rlg.pack_bf16(layers) if precision == "bf16" else None
x = O.make_clouds(B, N, "sphere", 5).cuda()
for _ in range(reps):
    pooled, _ = rlg.encoder_pool(x, layers, precision=precision, packed=packed)
torch.cuda.synchronize()
print("ok", float(pooled.sum()))
