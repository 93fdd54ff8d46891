"""Where does an autoencoder training step spend its time on the B200 (diagnostic)?  config_quick.yaml shapes:
encoder_dims [64,128,128,256,128], latent 128, decoder [256,256,6144], incomplete (B,1400,3) -> complete (B,2048,3)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import gan_rl_3d_b200 as rlg  # noqa: E402
from oracle import oracle as O  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16


class Decoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(128, 256), nn.BatchNorm1d(256), nn.ReLU(True), nn.Linear(256, 256),
                                 nn.BatchNorm1d(256), nn.ReLU(True), nn.Linear(256, 6144))

    def forward(self, g):
        return self.mlp(g).view(-1, 2048, 3)


torch.manual_seed(0)
enc = O.RefEncoderPort(3, 128, [64, 128, 128, 256, 128]).to(dev).train()
dec = Decoder().to(dev).train()
params = list(enc.parameters()) + list(dec.parameters())
opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-5)
x = O.make_clouds(B, 1400, "sphere", 1).to(dev)
y = O.make_clouds(B, 2048, "sphere", 2).to(dev)
ours = rlg.ChamferLoss()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def enc_fwd_bwd():
    for p in enc.parameters():
        p.grad = None
    enc(x).square().sum().backward()


def enc_fwd():
    with torch.no_grad():
        enc(x)


def step(loss_fn):
    opt.zero_grad(set_to_none=True)
    loss = loss_fn(dec(enc(x)), y)
    loss.backward()
    opt.step()


print(f"B={B}: stock encoder train fwd {timed(enc_fwd):.3f} ms, fwd+bwd {timed(enc_fwd_bwd):.3f} ms")
print(f"      AE step, stock ChamferLoss {timed(lambda: step(O.ref_port_chamfer_loss)):.3f} ms")
print(f"      AE step, B200 ChamferLoss  {timed(lambda: step(ours)):.3f} ms")
