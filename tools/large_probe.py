"""ChamferLoss forward + backward at the cfg5 shape (B pairs of N=M=16384), eager calls, CUDA events.  Diagnostic only.
    python tools/large_probe.py [B]          RLG_EXPERIMENTS_LIB=<variant .so> for A/B builds"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import gan_rl_3d_b200 as rlg  # noqa: E402
from oracle import oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = 16384
dev = torch.device("cuda:0")
ring = [(O.make_clouds(B, n, "sphere", 10 + k).to(dev).requires_grad_(True), O.make_clouds(B, n, "sphere", 90 + k).to(dev)) for k in range(3)]
crit = rlg.ChamferLoss()
one = torch.ones((), device=dev)


def step(k, backward=True):
    a, b = ring[k % len(ring)]
    a.grad = None
    loss = crit(a, b)
    if backward:
        loss.backward(gradient=one)


for name, bw in (("forward + backward", True), ("forward only (autograd on)", False)):
    for k in range(3):
        step(k, bw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(12):
        step(k, bw)
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B} N=M={n} {name}: {e0.elapsed_time(e1) / 12:.4f} ms per step")
