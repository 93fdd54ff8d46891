"""One small call of every kernel family of librlg_b200.so, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import gan_rl_3d_b200 as rlg  # noqa: E402
from oracle import oracle as O  # noqa: E402

dev = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def chamfer():
    for sweep in ("fp32", "tensor"):
        rlg.set_default_sweep(sweep)
        for (b, n, m) in ((2, 300, 257), (3, 64, 130)):
            a = O.make_clouds(b, n, "sphere", 1).to(dev).requires_grad_(True)
            c = O.make_clouds(b, m, "sphere", 2).to(dev).requires_grad_(True)
            rlg.ChamferLoss()(a, c).backward()
            rlg.chamfer_nearest(a.detach(), c.detach(), track_two=(sweep == "tensor"))
    rlg.set_default_sweep("auto")
    torch.cuda.synchronize()
    print("chamfer ok")


def encoder():
    torch.manual_seed(0)
    for dims in ([64, 128, 64], [64, 128, 128]):
        enc = O.RefEncoderPort(3, 16, dims)
        O.randomize_bn(enc, 1)
        enc = enc.to(dev).eval()
        x = O.make_clouds(2, 200, "sphere", 3).to(dev)
        layers = rlg.fold_trunk(enc.point_mlp)
        for prec in ("fp32", "fp32x", "bf16_layers"):
            rlg.encoder_pool(x, layers, precision=prec)
        if dims[-1] % 128 == 0:
            rlg.encoder_pool(x, layers, precision="bf16")
    torch.cuda.synchronize()
    print("encoder ok")


def train():
    torch.manual_seed(0)
    enc = O.RefEncoderPort(3, 16, [64, 128, 64])
    O.randomize_bn(enc, 1)
    enc = enc.to(dev).train()
    x = O.make_clouds(2, 200, "sphere", 3).to(dev)
    pooled = rlg.trunk_pool_autograd(enc, x)
    pooled.square().sum().backward()
    enc.eval()
    rlg.trunk_pool_autograd(enc, x).sum().backward()
    torch.cuda.synchronize()
    print("train ok")


for name, fn in (("chamfer", chamfer), ("encoder", encoder), ("train", train)):
    if which in ("all", name):
        fn()
