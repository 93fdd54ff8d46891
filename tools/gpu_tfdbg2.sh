#!/bin/bash
for sp in 0 1; do
echo "== split $sp"
RLG_TF_SPLIT=$sp RLG_TF_DEBUG=2 timeout 120 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
import gan_rl_3d_b200 as rlg
g = torch.Generator().manual_seed(0)
def sphere(b, n):
    x = torch.randn(b, n, 3, generator=g)
    return (x / x.norm(dim=2, keepdim=True)).cuda()
a, b = sphere(32, 2048), sphere(32, 2048)
for _ in range(2):
    rlg.chamfer_nearest(a, b, tensor=True)
torch.cuda.synchronize()
PY
done
