"""One batched environment step at the reference's dims, eager (for ncu launch lists): E episodes."""
import os, sys, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_rl_3d_b200 as rlg
from oracle import oracle as O
dev = "cuda:0"
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
ae = rlg.PointCloudAutoencoder().to(dev).eval()
gan = O.RefLatentGANPort().to(dev).eval()
env = rlg.BatchedRLEnvironment(ae.encode, gan.generate, ae.decode, gan.discriminate, dev)
env.reset({"incomplete": O.make_clouds(E, 1400, "sphere", 1), "complete": O.make_clouds(E, 2048, "sphere", 2)})
for k in range(3):
    ns, r, d, info = env.step(torch.randn(E, 1))
torch.cuda.synchronize()
print("mean reward", float(r.mean()))
