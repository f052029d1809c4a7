// Fused self-attention block for sm_100a (inference):
//   h_mid = SelfAttention(a) + h_in ;  a2 = LayerNorm(h_mid)
// (models/ncsn.py:160-164 in the reference: LayerNorm -> nn.SelfAttention -> + shortcut, followed by the FFN's
// LayerNorm; flax.nn.SelfAttention = q/k/v DenseGeneral, q / sqrt(depth), softmax(q k^T) v, out DenseGeneral).
// Replaces three launches (QKV GEMM -> attention -> out-projection GEMM) whose fp32 q/k/v round trip through HBM
// (1536 B written + 1536 B read per token) was the whole cost: here q, k, v never leave the SM.
//
// One CTA pair (cta_group::2) owns 256 tokens = 8 samples of 32 positions; each CTA holds 4 samples, one per TMEM lane
// quadrant, so a warp's 32 TMEM lanes are exactly the 32 positions of one sample.
//   GEMM qkv : D[0:384) (TMEM) = A[256 x 128] . Wqkv[128 x 384]          (three N = 128 MMAs per k-step, K = 128)
//   (kTrain additionally stores q | k | v, the probabilities and the attention output for the backward pass)
//   attention: 16 epilogue warps (4 per quadrant); a warp handles H/4 heads of its sample, one at a time: q/k/v head
//              slices TMEM -> registers (+ bias, q / sqrt(dh)) -> tf32 rows in a per-warp shared-memory tile -> Q K^T and
//              P V on mma.sync m16n8k8 tf32 (a 32x32x16 problem per head is far below a tcgen05 tile), fp32 softmax
//              with the reference's max-subtracted exp / sum on the accumulator fragments; o head slice -> bf16 ->
//              written straight into the canonical K-major SWIZZLE_128B operand layout of the next GEMM
//   GEMM out : D[384:512) = O[256 x 128] . Wo[128 x 128]
//   epilogue : + bo + residual -> h_mid (fp32) ; single-pass LayerNorm -> a2 (bf16)
// The weights (Wq|Wk|Wv|Wo, 64 KB per CTA) are fetched once per CTA and stay in shared memory for all its tiles.
//
// Warp roles (640 threads): 0 TMA producer, 1 MMA issuer (leader CTA), 2 TMEM allocator, 4..19 epilogue.
#pragma once
#include "gemm_tcgen05.cuh"
#include "kernels.cuh"
#include "row_epilogue.cuh"

namespace smd {

struct AttnBlockArgs {
  const float* b_qkv;            // [384] = bq | bk | bv
  const float* b_o;              // [128]
  const float* residual;         // fp32 [M][128] (may alias out_f32)
  float* out_f32;                // fp32 [M][128]
  const float* ln_gamma;         // [128] LayerNorm of the new residual stream -> out_bf16
  const float* ln_beta;
  __nv_bfloat16* out_bf16;       // bf16 [M][128]
  // training (kTrain): what the backward pass needs (csrc/backward.cu), in the layouts of the three-launch path
  float* qkv_out;                // fp32 [M][384] = (q | k | v) + bias, q unscaled
  float* probs_out;              // fp32 [M / 32][H][32][32] softmax probabilities
  __nv_bfloat16* o_out;          // bf16 [M][128] attention output (operand of the out-projection's weight gradient)
  int M, H;                      // tokens; heads (dh = 128 / H in {8, 16})
};

struct AttnSmem {
  static constexpr int kA = 32768;          // [2 k-blocks][128 rows][128 B]
  static constexpr int kW = 16384;          // per-CTA half of a 128-column weight block: [2 k-blocks][64 k][64 n]
  static constexpr int kO = 32768;          // attention output as the out-projection's A operand
  static constexpr int offA = 0;
  static constexpr int offW = offA + kA;                 // q, k, v, o blocks
  static constexpr int offO = offW + 4 * kW;
  static constexpr int offBar = offO + kO;
  static constexpr int kBarBytes = 256;
  static constexpr int offScr = offBar + kBarBytes;
  static constexpr int kEpiWarps = 16;
  static constexpr int kKvFloats = 2 * 32 * 20;          // k and v tiles of one head: [32][dh + 4], dh <= 16
  static constexpr int kScrPerWarp = kKvFloats;          // later the 32 x 32 transpose tile + 2 x 32 LayerNorm partials of the final epilogue
  static_assert(kScrPerWarp >= 1024 + 64, "scratch too small for the final epilogue");
  static constexpr int kScrBytes = kEpiWarps * kScrPerWarp * 4;
  static constexpr int kTotal = offScr + kScrBytes + 1024;
  static constexpr int kThreads = 128 + 32 * kEpiWarps;
};

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

template <int DH>
__device__ __forceinline__ void tmem_ld_head(uint32_t taddr, float (&out)[DH]) {
  if constexpr (DH == 16) {
    uint32_t r[16];
    tmem_ld_32x16(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = __uint_as_float(r[i]);
  } else {
    uint32_t r[8];
    tmem_ld_32x8(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = __uint_as_float(r[i]);
  }
}

template <int DH, bool kTrain>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(AttnSmem::kThreads, 1)
attn_block_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWqkv,
                  const __grid_constant__ CUtensorMap tmWo, const AttnBlockArgs p) {
  using S = AttnSmem;
  constexpr int PITCH = DH + 4;
  extern __shared__ uint8_t attn_smem_raw[];
  uint8_t* smem = attn_smem_raw + ((1024u - (smem_u32(attn_smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::offBar);
  uint64_t* w_full = bars + 0;
  uint64_t* a_full = bars + 1;
  uint64_t* a_empty = bars + 2;
  uint64_t* qkv_full = bars + 3;
  uint64_t* o_full = bars + 4;
  uint64_t* d2_full = bars + 5;
  uint64_t* d2_empty = bars + 6;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_tiles = (p.M + 255) / 256;
  const int group = blockIdx.x / 2, num_groups = gridDim.x / 2;

  pdl_trigger();
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmWqkv);
    tma_prefetch_desc(&tmWo);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(w_full, 1);
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    mbar_init(qkv_full, 1);
    mbar_init(o_full, 2 * S::kEpiWarps);
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, 2 * S::kEpiWarps);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<2>(tmem_ptr_smem, 512);
    tmem_relinquish<2>();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      // weights once: block b (q, k, v) = columns [128 b, 128 b + 128) of Wqkv [K = 128][N = 384]; this CTA's 64 columns
      if (leader) mbar_arrive_expect_tx(w_full, 2u * 4u * S::kW);
      const uint32_t w_bar = mapa_shared(smem_u32(w_full), 0);
      for (int b = 0; b < 3; ++b)
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d_2sm(&tmWqkv, w_bar, smem + S::offW + b * S::kW + kb * 8192, b * 128 + static_cast<int>(rank) * 64, 64 * kb);
      for (int kb = 0; kb < 2; ++kb)
        tma_load_2d_2sm(&tmWo, w_bar, smem + S::offW + 3 * S::kW + kb * 8192, static_cast<int>(rank) * 64, 64 * kb);
      const uint32_t a_full_l = mapa_shared(smem_u32(a_full), 0);
      uint32_t nt = 0;
      for (int tile = group; tile < num_tiles; tile += num_groups, ++nt) {
        const int m_row0 = tile * 256 + static_cast<int>(rank) * 128;
        mbar_wait(a_empty, (nt & 1u) ^ 1u);
        if (leader) mbar_arrive_expect_tx(a_full, 2u * S::kA);
        for (int kb = 0; kb < 2; ++kb) tma_load_2d_2sm(&tmA, a_full_l, smem + S::offA + kb * 16384, 64 * kb, m_row0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one lane) =====================
    if (leader && elect_one()) {
      const uint32_t idesc = make_idesc_bf16(256, 128, 0, 1);   // A / O K-major, weights MN-major
      const uint32_t sA = smem_u32(smem + S::offA), sO = smem_u32(smem + S::offO);
      mbar_wait(w_full, 0);
      uint32_t nt = 0;
      for (int tile = group; tile < num_tiles; tile += num_groups, ++nt) {
        mbar_wait(a_full, nt & 1u);
        // (the q/k/v accumulators of the previous tile were drained before its o_full arrived, which this thread
        // waited for below)
        tcgen05_fence_after();
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const uint32_t sW = smem_u32(smem + S::offW + b * S::kW);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = make_smem_desc_sw128(sA + kb * 16384 + k * 32, 0u, 1024u);
              const uint64_t bd = make_smem_desc_sw128(sW + kb * 8192 + k * 2048, 8192u, 1024u);
              umma_bf16<2>(tmem_base + b * 128u, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
            }
        }
        umma_commit<2>(a_empty);      // the A tile may be refilled
        umma_commit<2>(qkv_full);
        mbar_wait_cluster(o_full, nt & 1u);
        mbar_wait_cluster(d2_empty, (nt & 1u) ^ 1u);
        tcgen05_fence_after();
        {
          const uint32_t sW = smem_u32(smem + S::offW + 3 * S::kW);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = make_smem_desc_sw128(sO + kb * 16384 + k * 32, 0u, 1024u);
              const uint64_t bd = make_smem_desc_sw128(sW + kb * 8192 + k * 2048, 8192u, 1024u);
              umma_bf16<2>(tmem_base + 384u, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
            }
        }
        umma_commit<2>(d2_full);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps =====================
    const uint32_t q = warp & 3u;                      // TMEM lane quadrant = sample within this CTA
    const int eg = static_cast<int>(warp - 4u) >> 2;   // 0..3: head group / 32-column quarter of the output
    float* scr_all = reinterpret_cast<float*>(smem + S::offScr);
    float* scr = scr_all + (warp - 4u) * S::kScrPerWarp;
    uint32_t* sA = reinterpret_cast<uint32_t*>(scr);                 // [32][PITCH] tf32: q rows, later v rows
    uint32_t* sB = reinterpret_cast<uint32_t*>(scr) + 32 * PITCH;    // [32][PITCH] tf32: k rows
    const int g = static_cast<int>(lane >> 2), t = static_cast<int>(lane & 3u);   // mma.sync fragment coordinates
    const uint32_t o_full_l = mapa_shared(smem_u32(o_full), 0);
    const uint32_t d2_empty_l = mapa_shared(smem_u32(d2_empty), 0);
    const uint32_t r_in_tile = q * 32u + lane;         // row of this thread inside the CTA's 128-row tile
    const int heads_per_warp = p.H / 4;
    const float qscale = rsqrtf(static_cast<float>(DH));
    const uint32_t lane_base = tmem_base + ((q * 32u) << 16);
    uint32_t nt = 0;
    for (int tile = group; tile < num_tiles; tile += num_groups, ++nt) {
      const int row = tile * 256 + static_cast<int>(rank) * 128 + static_cast<int>(r_in_tile);
      [[maybe_unused]] const bool row_ok = row < p.M;
      mbar_wait(qkv_full, nt & 1u);
      tcgen05_fence_after();
      for (int hh = 0; hh < heads_per_warp; ++hh) {
        const int h = eg * heads_per_warp + hh;
        const int hc = h * DH;                         // first column of this head inside a 128-wide block
        __syncwarp();
        float qv[DH], kv[DH], vv[DH];
        tmem_ld_head<DH>(lane_base + static_cast<uint32_t>(hc), qv);
        tmem_ld_head<DH>(lane_base + 128u + static_cast<uint32_t>(hc), kv);
        tmem_ld_head<DH>(lane_base + 256u + static_cast<uint32_t>(hc), vv);
        // + bias, q / sqrt(depth) (flax), rounded to tf32 once; q and k rows -> this warp's two [32][DH + 4] tiles
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          const float4 bq = __ldg(reinterpret_cast<const float4*>(p.b_qkv + hc + d));
          const float4 bk = __ldg(reinterpret_cast<const float4*>(p.b_qkv + 128 + hc + d));
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.b_qkv + 256 + hc + d));
          *reinterpret_cast<uint4*>(sA + lane * PITCH + d) =
              make_uint4(to_tf32((qv[d] + bq.x) * qscale), to_tf32((qv[d + 1] + bq.y) * qscale),
                         to_tf32((qv[d + 2] + bq.z) * qscale), to_tf32((qv[d + 3] + bq.w) * qscale));
          *reinterpret_cast<uint4*>(sB + lane * PITCH + d) =
              make_uint4(to_tf32(kv[d] + bk.x), to_tf32(kv[d + 1] + bk.y), to_tf32(kv[d + 2] + bk.z), to_tf32(kv[d + 3] + bk.w));
          vv[d] += bv.x; vv[d + 1] += bv.y; vv[d + 2] += bv.z; vv[d + 3] += bv.w;
          if constexpr (kTrain) {
            if (row_ok) {
              float* qd = p.qkv_out + static_cast<size_t>(row) * 384 + hc + d;
              *reinterpret_cast<float4*>(qd) = make_float4(qv[d] + bq.x, qv[d + 1] + bq.y, qv[d + 2] + bq.z, qv[d + 3] + bq.w);
              *reinterpret_cast<float4*>(qd + 128) = make_float4(kv[d] + bk.x, kv[d + 1] + bk.y, kv[d + 2] + bk.z, kv[d + 3] + bk.w);
              *reinterpret_cast<float4*>(qd + 256) = make_float4(vv[d], vv[d + 1], vv[d + 2], vv[d + 3]);
            }
          }
        }
        __syncwarp();
        // ---- S = (Q / sqrt(dh)) K^T on mma.sync m16n8k8 tf32: 2 m-tiles x 4 n-tiles, DH / 8 k-steps
        float sc[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt2 = 0; nt2 < 4; ++nt2)
#pragma unroll
            for (int i = 0; i < 4; ++i) sc[mt][nt2][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < DH / 8; ++ks) {
          uint32_t a[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const uint32_t* q0 = sA + (16 * mt + g) * PITCH + 8 * ks + t;
            a[mt][0] = q0[0]; a[mt][1] = q0[8 * PITCH]; a[mt][2] = q0[4]; a[mt][3] = q0[8 * PITCH + 4];
          }
#pragma unroll
          for (int nt2 = 0; nt2 < 4; ++nt2) {
            const uint32_t* k0 = sB + (8 * nt2 + g) * PITCH + 8 * ks + t;
            const uint32_t b0 = k0[0], b1 = k0[4];
            mma_tf32_16x8x8(sc[0][nt2], a[0], b0, b1);
            mma_tf32_16x8x8(sc[1][nt2], a[1], b0, b1);
          }
        }
        __syncwarp();
        // v rows take over the q tile (q is no longer needed)
#pragma unroll
        for (int d = 0; d < DH; d += 4)
          *reinterpret_cast<uint4*>(sA + lane * PITCH + d) = make_uint4(to_tf32(vv[d]), to_tf32(vv[d + 1]), to_tf32(vv[d + 2]), to_tf32(vv[d + 3]));
        // ---- row softmax (max-subtracted exp / sum, fp32): a row lives in the 4 lanes of a quad, 8 values per lane
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
          for (int hr = 0; hr < 2; ++hr) {
            float mx = -INFINITY;
#pragma unroll
            for (int nt2 = 0; nt2 < 4; ++nt2) mx = fmaxf(mx, fmaxf(sc[mt][nt2][2 * hr], sc[mt][nt2][2 * hr + 1]));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float sum = 0.f;
#pragma unroll
            for (int nt2 = 0; nt2 < 4; ++nt2) {
              const float e0 = __expf(sc[mt][nt2][2 * hr] - mx), e1 = __expf(sc[mt][nt2][2 * hr + 1] - mx);
              sc[mt][nt2][2 * hr] = e0; sc[mt][nt2][2 * hr + 1] = e1;
              sum += e0 + e1;
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const float inv = 1.0f / sum;
#pragma unroll
            for (int nt2 = 0; nt2 < 4; ++nt2) { sc[mt][nt2][2 * hr] *= inv; sc[mt][nt2][2 * hr + 1] *= inv; }
          }
        }
        if constexpr (kTrain) {
          // accumulator layout = the layout attention_bwd_mma_kernel reads back: rows 16 mt + g (+ 8), keys 8 nt + 2t (+ 1)
          const int smp = (tile * 256 + static_cast<int>(rank) * 128 + static_cast<int>(q) * 32) >> 5;
          if (smp * 32 < p.M) {
            float* pr = p.probs_out + (static_cast<size_t>(smp) * p.H + hc / DH) * 1024;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
              for (int nt2 = 0; nt2 < 4; ++nt2) {
                *reinterpret_cast<float2*>(pr + (16 * mt + g) * 32 + 8 * nt2 + 2 * t) = make_float2(sc[mt][nt2][0], sc[mt][nt2][1]);
                *reinterpret_cast<float2*>(pr + (16 * mt + g + 8) * 32 + 8 * nt2 + 2 * t) = make_float2(sc[mt][nt2][2], sc[mt][nt2][3]);
              }
          }
        }
        __syncwarp();
        // ---- O = P V: the probabilities feed the second product straight from the accumulator registers (within a
        // block of 8 keys, k-slot t holds key 2t and slot t + 4 key 2t + 1; the V fragment is read with the same
        // permutation -- a sum over keys does not care about their order)
        float acc[2][DH / 8][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int n2 = 0; n2 < DH / 8; ++n2)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][n2][i] = 0.f;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          uint32_t a[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            a[mt][0] = to_tf32(sc[mt][kb][0]); a[mt][1] = to_tf32(sc[mt][kb][2]);
            a[mt][2] = to_tf32(sc[mt][kb][1]); a[mt][3] = to_tf32(sc[mt][kb][3]);
          }
#pragma unroll
          for (int n2 = 0; n2 < DH / 8; ++n2) {
            const uint32_t* v0 = sA + (8 * kb + 2 * t) * PITCH + 8 * n2 + g;
            const uint32_t b0 = v0[0], b1 = v0[PITCH];
            mma_tf32_16x8x8(acc[0][n2], a[0], b0, b1);
            mma_tf32_16x8x8(acc[1][n2], a[1], b0, b1);
          }
        }
        // o head slice -> bf16 -> canonical K-major SWIZZLE_128B operand of the out-projection: 16-byte chunk c of row r
        // lives at r * 128 + ((c ^ (r & 7)) << 4) inside k-block (column / 64); a lane owns column pairs of 4 rows
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int n2 = 0; n2 < DH / 8; ++n2)
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
              const uint32_t rr = q * 32u + static_cast<uint32_t>(16 * mt + g + 8 * hr);
              const int col = hc + 8 * n2 + 2 * t;
              uint8_t* dst = smem + S::offO + (col >> 6) * 16384 + rr * 128u +
                             (((static_cast<uint32_t>(col & 63) >> 3) ^ (rr & 7u)) << 4) + static_cast<uint32_t>(col & 7) * 2u;
              const uint32_t ob = pack_bf16x2(acc[mt][n2][2 * hr], acc[mt][n2][2 * hr + 1]);
              *reinterpret_cast<uint32_t*>(dst) = ob;
              if constexpr (kTrain) {
                const int grow = tile * 256 + static_cast<int>(rank) * 128 + static_cast<int>(rr);
                if (grow < p.M) *reinterpret_cast<uint32_t*>(p.o_out + static_cast<size_t>(grow) * 128 + col) = ob;
              }
            }
      }
      // q/k/v accumulators drained and this warp's part of O written
      tcgen05_fence_before();
      fence_proxy_async_smem();    // O stores become visible to the tensor core's (async proxy) reads
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(o_full_l);
      // ---------------- final epilogue: D2 + bo + residual -> LayerNorm (csrc/row_epilogue.cuh) ----------------
      {
        const int c0 = eg * 32;
        const RowEpiArgs ea{p.b_o, p.residual, p.out_f32, p.ln_gamma, p.ln_beta, p.out_bf16, p.M};
        const int row0 = tile * 256 + static_cast<int>(rank) * 128 + static_cast<int>(q) * 32;
        float4 pre[8];
        row_epi_prefetch(ea, row0, c0, lane, pre);
        mbar_wait(d2_full, nt & 1u);
        tcgen05_fence_after();
        __syncwarp();
        row_epi_stage(pre, scr, lane);         // the warp's q/k/v tiles are dead: the scratch becomes its transpose tile
        uint32_t r[32];
        tmem_ld_32x32(lane_base + 384u + static_cast<uint32_t>(c0), r);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(d2_empty_l);
        row_epi_finish(ea, r, scr, scr + 1024, scr_all + 1024, S::kScrPerWarp, q, row0, c0, lane);
      }
    }
  }

  // ===================== teardown =====================
  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc<2>(tmem_base, 512);
}

}  // namespace smd
