"""First-light check of the tcgen05 encoder path on small shapes (run under `timeout`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import gan_rl_3d_b200 as rlg  # noqa: E402
from oracle import oracle as O  # noqa: E402

for dims, B, N in [([64, 128], 1, 128), ([64, 128, 256], 1, 128), ([64, 128, 1024], 2, 300), ([64, 128, 1024], 256, 2048)]:
    torch.manual_seed(0)
    enc = O.RefEncoderPort(3, 32, dims)
    O.randomize_bn(enc, 1)
    enc.eval()
    x = O.make_clouds(B, N, "sphere", 5)
    with torch.no_grad():
        want = enc.pooled(x).numpy()
    layers = rlg.fold_trunk(enc.cuda().point_mlp)
    got, _ = rlg.encoder_pool(x.cuda(), layers, precision="bf16")
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    ok, err = O.gfv_close(got, want, 2e-2, 0.25)
    print(dims, B, N, "ok" if ok else "MISMATCH", "max rel err %.3e" % err, "norm err %.3e" % (np.linalg.norm(got - want) / np.linalg.norm(want)), flush=True)
    if not ok:
        bad = np.argwhere(np.abs(got - want) > 2e-2 * np.maximum(np.abs(want), 1e-2 * np.abs(want).max()))
        print("  first bad:", bad[:5].tolist(), got.flat[:4], want.flat[:4])
x = O.make_clouds(256, 2048, "sphere", 5).cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
packed = rlg.pack_bf16(layers)
for _ in range(3):
    rlg.encoder_pool(x, layers, precision="bf16", packed=packed)
e0.record()
for _ in range(20):
    rlg.encoder_pool(x, layers, precision="bf16", packed=packed)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("cfg3 bf16: %.3f ms  %.0f clouds/s  %.1f TFLOP/s" % (ms, 256 / ms * 1e3, 2 * 2048 * (3 * 64 + 64 * 128 + 128 * 1024) * 256 / ms / 1e9))
