"""A few eager autoencoder training steps at the reference's dims (for ncu launch lists / captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_rl_3d_b200 as rlg
dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
model = rlg.PointCloudAutoencoder().to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True, fused=True)
g = torch.Generator().manual_seed(1)
def sphere(b, n):
    x = torch.randn(b, n, 3, generator=g)
    return (x / x.norm(dim=2, keepdim=True)).to(dev)
x, y = sphere(B, 1400), sphere(B, 2048)
crit = rlg.ChamferLoss()
for k in range(steps):
    opt.zero_grad(set_to_none=True)
    loss = crit(model(x)[0], y)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss))
