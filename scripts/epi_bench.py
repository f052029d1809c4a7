"""Micro-benchmark of the GEMM epilogue: K=64/128 GEMMs (negligible MMA time) with different fused epilogues."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smd_b200 import lib as L

lib = L.load_library()
st = torch.cuda.current_stream().cuda_stream


def run(M, N, K, cg, bias=False, res=False, act=0, f32=True, bf16=False, stats=False, ln=False, iters=20):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    o32 = torch.empty(M, N, device="cuda") if f32 else None
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if bf16 else None
    bi = torch.zeros(N, device="cuda") if bias else None
    rs = torch.randn(M, N, device="cuda") if res else None
    stt = torch.zeros(M, 2, device="cuda") if stats else None
    g = torch.ones(N, device="cuda") if ln else None
    b = torch.zeros(N, device="cuda") if ln else None
    p = lambda t: None if t is None else t.data_ptr()
    def call():
        L.check(lib.smd_gemm_bf16(A.data_ptr(), B.data_ptr(), M, N, K, 0, 0, 0, cg, p(bi), p(rs), act, p(o32), p(o16),
                                  p(stt), p(g), p(b), st))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for cg in (2, 1):
    for (M, N, K) in ((32000, 2048, 64), (32000, 2048, 128), (32000, 2048, 2048), (4096, 2048, 2048), (32000, 128, 2048)):
        cfgs = [("f32", dict()), ("f32+bias", dict(bias=True)), ("f32+bias+stats", dict(bias=True, stats=True)),
                ("f32+bias+res+stats", dict(bias=True, res=True, stats=True)),
                ("bf16", dict(f32=False, bf16=True)), ("bf16+bias+gelu", dict(f32=False, bf16=True, bias=True, act=1))]
        if N == 128:
            cfgs = [("f32", dict()), ("ln f32+bf16+res", dict(bias=True, res=True, bf16=True, ln=True))]
        for name, kw in cfgs:
            us = run(M, N, K, cg, **kw)
            tiles = (M + 128 * cg - 1) // (128 * cg) * ((N + 255) // 256 if N >= 256 else 1)
            per_cta = tiles / (148 // cg)
            print(f"cg{cg} M{M} N{N} K{K} {name:22s} {us:8.1f} us  ({us / per_cta:6.2f} us per tile-slot)")
