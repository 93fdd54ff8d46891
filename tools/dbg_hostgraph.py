import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_rl_3d_b200 as rlg
from oracle import oracle as O
P = importlib.import_module("gan-rl_3d_b200.pipeline")
DEV = torch.device("cuda:0")

def eager(a, b, sweep):
    rlg.set_default_sweep(sweep)
    a = a.to(DEV).requires_grad_(True)
    loss = rlg.ChamferLoss()(a, b.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    return loss.item()

seed = 1000
bad = 0
for trial in range(6):
    for sweep in ("tensor", "fp32"):
        for one in (True, False):
            # poison the allocator's free blocks so a stale read cannot look right
            junk = [torch.randn(1 << 18, device=DEV) * 3 for _ in range(8)]
            del junk
            seed += 10
            rlg.set_default_sweep(sweep)
            raw = [(O.make_clouds(2, 300, "uniform", seed + k), O.make_clouds(2, 257, "uniform", seed + 5 + k)) for k in range(5)]
            host = [P.pin_pair(a, b) if one else (a.pin_memory(), b.pin_memory()) for a, b in raw]
            g = P.HostChamferStepGraph(host, DEV)
            ref = [eager(a, b, "fp32") for a, b in raw]
            rlg.set_default_sweep(sweep)
            for rep in range(3):
                g.replay(); torch.cuda.synchronize()
                got = [g.losses_host[k].item() for k in range(5)]
                if got != ref:
                    bad += 1
                    print("MISMATCH", sweep, "one_transfer", one, "trial", trial, "rep", rep, got, ref, flush=True)
print("mismatches:", bad)
