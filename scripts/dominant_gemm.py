"""Runs the dominant kernel (2048x2048 res-block GEMM + bias + row-stat epilogue) alone, for `ncu --set full`.
Usage: python scripts/dominant_gemm.py [M_tokens] [cta_group]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smd_b200 import lib as L
lib = L.load_library()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
A = torch.randn(M, 2048, device="cuda").to(torch.bfloat16)
B = torch.randn(2048, 2048, device="cuda").to(torch.bfloat16)
out = torch.empty(M, 2048, device="cuda")
bias = torch.zeros(2048, device="cuda")
stats = torch.zeros(M, 2, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    flush.zero_()
    L.check(lib.smd_gemm_bf16(A.data_ptr(), B.data_ptr(), M, 2048, 2048, 0, 1, 256, cg, bias.data_ptr(), None, 0,
                              out.data_ptr(), None, stats.data_ptr(), None, None, st))
torch.cuda.synchronize()
print("ok")
