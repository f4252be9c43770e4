"""One fp8 GEMM launch pattern for ncu capture (4096 x 28672 x 4096, the Llama-3-8B gate|up shape)."""
import torch

from colossalai_b200.ops import gemm_native

M, N, K = 4096, 28672, 4096
a = torch.randn(M, K, device="cuda").to(torch.float8_e4m3fn)
b = torch.randn(N, K, device="cuda").to(torch.float8_e4m3fn)
one = torch.ones(1, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(5):
    gemm_native.gemm_fp8_nt(a, b, one, one, torch.bfloat16, out)
torch.cuda.synchronize()
