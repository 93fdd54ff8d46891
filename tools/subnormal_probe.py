"""Does tcgen05 kind::f16 honour fp16 subnormal operands?  (activation 2^-20 through the fp32x layer GEMM)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_rl_3d_b200 as rlg
dev = "cuda:0"
x = torch.zeros(1, 128, 3, device=dev)
for ex in (-3, -10, -14, -15, -18, -20, -24):
    a = 2.0 ** ex
    layers = [(torch.zeros(64, 3, device=dev), torch.full((64,), a, device=dev)),
              (torch.ones(64, 64, device=dev), torch.zeros(64, device=dev))]
    out = rlg.encoder_pool_gemm(x, layers, 2)
    a2 = a * (1 + 2.0 ** -12)       # needs the lo piece: hi = a, lo = a * 2^-12
    layers[0] = (layers[0][0], torch.full((64,), a2, device=dev))
    out2 = rlg.encoder_pool_gemm(x, layers, 2)
    print(f"a = 2^{ex}: got {out[0,0].item():.6e} want {64*a:.6e};  with lo part: got {out2[0,0].item():.9e} want {64*a2:.9e}")
