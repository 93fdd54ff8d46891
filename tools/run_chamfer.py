"""Profiling driver: a few forward+backward Chamfer steps at the headline shape (B=32, N=M=2048).
    python tools/run_chamfer.py [steps] [B] [N] [M]
Used under `ncu` (see profiles/README.md); prints nothing that counts as a benchmark value."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import gan_rl_3d_b200 as rlg  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
N = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
M = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
g = torch.Generator().manual_seed(0)


def sphere(b, n):
    x = torch.randn(b, n, 3, generator=g)
    return (x / x.norm(dim=2, keepdim=True)).cuda()


crit = rlg.ChamferLoss()
for s in range(steps):
    a = sphere(B, N).requires_grad_(True)
    b = sphere(B, M)
    loss = crit(a, b)
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
