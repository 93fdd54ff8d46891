#!/bin/bash
for ko in 0 3; do
echo "== RLG_TF_KO=$ko"
RLG_TF_DEBUG=1 RLG_TF_KO=$ko timeout 120 python - <<'PY'
import sys, os, importlib
sys.path.insert(0, os.getcwd())
import torch
import gan_rl_3d_b200 as rlg
g = torch.Generator().manual_seed(0)
def sphere(b, n):
    x = torch.randn(b, n, 3, generator=g)
    return (x / x.norm(dim=2, keepdim=True)).cuda()
a, b = sphere(32, 2048), sphere(32, 2048)
for _ in range(3):
    rlg.chamfer_nearest(a, b, tensor=True)
torch.cuda.synchronize()
PY
done
