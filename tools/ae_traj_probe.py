"""Loss trajectories of a few Adam steps: captured B200 step vs stock fp32 CUDA eager vs float64 CPU (diagnostic)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_rl_3d_b200 as rlg
from oracle import oracle as O
DEV = "cuda:0"
torch.manual_seed(3)
model = rlg.PointCloudAutoencoder(3, 32, 256, [64, 128, 64], [64, 768]).to(DEV).train()
ref = copy.deepcopy(model).double().cpu().train()
stock = copy.deepcopy(model).train()
S, B = 4, 6
batches = [(O.make_clouds(B, 180, "sphere", 50 + k), O.make_clouds(B, 256, "sphere", 90 + k)) for k in range(S)]

def run(m, opt, dt, dev, eager_rlg=False):
    out = []
    for a, b in batches:
        a, b = a.to(dev, dt), b.to(dev, dt)
        opt.zero_grad()
        if eager_rlg:
            recon, _ = m(a)
            loss = rlg.ChamferLoss()(recon, b)
        else:
            recon = m.decoder(m.encoder.global_mlp(torch.max(m.encoder.point_mlp(a.transpose(2, 1)), dim=2)[0]))
            loss = O.ref_port_chamfer_loss(recon, b)
        loss.backward(); opt.step(); out.append(float(loss))
    return out
eager = copy.deepcopy(model).train()
print("float64      ", run(ref, torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=1e-5), torch.float64, "cpu"))
print("stock fp32   ", run(stock, torch.optim.Adam(stock.parameters(), lr=1e-3, weight_decay=1e-5), torch.float32, DEV))
print("ours eager   ", run(eager, torch.optim.Adam(eager.parameters(), lr=1e-3, weight_decay=1e-5), torch.float32, DEV, True))
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True)
g = rlg.AEStepGraph(model, opt, [(a.to(DEV), b.to(DEV)) for a, b in batches])
g.replay(); torch.cuda.synchronize()
print("ours captured", [float(l) for l in g.losses])
