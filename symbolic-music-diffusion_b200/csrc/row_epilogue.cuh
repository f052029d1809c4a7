// Final epilogue shared by the fused FFN and the attention-block kernels:
//   x = acc + bias + residual  -> fp32 residual stream ;  LayerNorm(x) -> bf16 operand of the next GEMM
// (models/ncsn.py:160-166: "+ shortcut" followed by the next sub-block's LayerNorm; flax LayerNorm = E[x^2] - E[x]^2,
// eps 1e-6).  tcgen05.ld hands every thread one ROW of the accumulator, so reading the residual / writing the outputs
// straight from that layout makes each warp instruction touch 32 different 128-byte lines.  Instead every global
// access goes through a warp-private 32 x 32 fp32 transpose tile in shared memory (16-byte chunk j of row r at chunk
// slot j ^ (r & 7): 128-bit accesses are conflict-free both for "lane = row" and for "8 lanes = one row"), so a warp
// instruction covers whole 128-byte rows.
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

namespace smd {

struct RowEpiArgs {
  const float* bias;             // [128]
  const float* residual;         // fp32 [M][128] (may alias out_f32)
  float* out_f32;                // fp32 [M][128]
  const float* ln_gamma;         // [128]
  const float* ln_beta;
  __nv_bfloat16* out_bf16;       // bf16 [M][128]
  int M;
};

// coalesced load of the warp's 32 x 32 residual block (rows row0 .., columns c0 ..): lane -> row 4 it + lane / 8,
// columns 4 (lane % 8) .. + 3.  Issued before the accumulator is ready so the latency hides under the last MMAs.
__device__ __forceinline__ void row_epi_prefetch(const RowEpiArgs& p, int row0, int c0, uint32_t lane, float4 (&pre)[8]) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int grow = row0 + it * 4 + static_cast<int>(lane >> 3);
    pre[it] = grow < p.M ? *reinterpret_cast<const float4*>(p.residual + static_cast<size_t>(grow) * 128 + c0 + 4 * (lane & 7u))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// transpose tile addressing: row r is 8 chunks of 4 floats, chunk j of row r lives at float4 index r * 8 + (j ^ (r & 7))
__device__ __forceinline__ float4* row_epi_chunk(float* tsc, uint32_t r, uint32_t j) {
  return reinterpret_cast<float4*>(tsc) + (r * 8u + (j ^ (r & 7u)));
}

__device__ __forceinline__ void row_epi_stage(const float4 (&pre)[8], float* tsc, uint32_t lane) {
#pragma unroll
  for (int it = 0; it < 8; ++it) *row_epi_chunk(tsc, static_cast<uint32_t>(it) * 4u + (lane >> 3), lane & 7u) = pre[it];
  __syncwarp();
}

// acc: this thread's row (TMEM lane) of the accumulator, 32 columns starting at c0.  stat_mine: 64 floats of this warp,
// stat_base + (g * 4 + q) * stat_stride: the same of the warp that owns column quarter g of TMEM lane quadrant q.
__device__ __forceinline__ void row_epi_finish(const RowEpiArgs& p, const uint32_t (&acc)[32], float* tsc, float* stat_mine,
                                               const float* stat_base, int stat_stride, uint32_t q, int row0, int c0,
                                               uint32_t lane) {
  float v[32];
  float s1 = 0.f, s2 = 0.f;
  {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + c0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 bb = __ldg(b4 + i);
      const float4 rr = *row_epi_chunk(tsc, lane, static_cast<uint32_t>(i));
      v[4 * i] = __uint_as_float(acc[4 * i]) + bb.x + rr.x;
      v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + bb.y + rr.y;
      v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + bb.z + rr.z;
      v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + bb.w + rr.w;
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) { s1 += v[i]; s2 += v[i] * v[i]; }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *row_epi_chunk(tsc, lane, static_cast<uint32_t>(i)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const uint32_t rr = static_cast<uint32_t>(it) * 4u + (lane >> 3);
    const int grow = row0 + static_cast<int>(rr);
    if (grow < p.M)
      *reinterpret_cast<float4*>(p.out_f32 + static_cast<size_t>(grow) * 128 + c0 + 4u * (lane & 7u)) =
          *row_epi_chunk(tsc, rr, lane & 7u);
  }
  // row statistics: each of the four warps of a quadrant saw 32 of the 128 columns
  stat_mine[lane * 2] = s1; stat_mine[lane * 2 + 1] = s2;
  asm volatile("bar.sync %0, 128;" ::"r"(1u + q) : "memory");
  float t1 = 0.f, t2 = 0.f;
#pragma unroll
  for (int g4 = 0; g4 < 4; ++g4) {
    const float* sp = stat_base + (static_cast<uint32_t>(g4) * 4u + q) * stat_stride + lane * 2;
    t1 += sp[0]; t2 += sp[1];
  }
  asm volatile("bar.sync %0, 128;" ::"r"(1u + q) : "memory");
  const float mean = t1 * (1.0f / 128.0f);
  const float rstd = rsqrtf(t2 * (1.0f / 128.0f) - mean * mean + 1e-6f);
  {
    const float4* g4 = reinterpret_cast<const float4*>(p.ln_gamma + c0);
    const float4* b4 = reinterpret_cast<const float4*>(p.ln_beta + c0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 g = __ldg(g4 + i), e = __ldg(b4 + i);
      *row_epi_chunk(tsc, lane, static_cast<uint32_t>(i)) =
          make_float4((v[4 * i] - mean) * (rstd * g.x) + e.x, (v[4 * i + 1] - mean) * (rstd * g.y) + e.y,
                      (v[4 * i + 2] - mean) * (rstd * g.z) + e.z, (v[4 * i + 3] - mean) * (rstd * g.w) + e.w);
    }
  }
  __syncwarp();
  // bf16 rows are 64 bytes: lane -> row 8 it + lane / 4, columns 8 (lane % 4) .. + 7 (two chunks)
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const uint32_t rr = static_cast<uint32_t>(it) * 8u + (lane >> 2), cj = 2u * (lane & 3u);
    const int grow = row0 + static_cast<int>(rr);
    const float4 a = *row_epi_chunk(tsc, rr, cj), b = *row_epi_chunk(tsc, rr, cj + 1u);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    if (grow < p.M)
      *reinterpret_cast<uint4*>(p.out_bf16 + static_cast<size_t>(grow) * 128 + c0 + 4u * cj) =
          make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                     *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
  }
  __syncwarp();
}

}  // namespace smd
