// Fused transformer FFN for sm_100a:  out = LayerNorm_next( gelu(a W1 + b1) W2 + b2 + residual )
// (models/ncsn.py:160-166 in the reference: Dense(2048) -> gelu -> Dense(128) -> + residual, followed by the next
// block's LayerNorm).  The 2048-wide hidden activation never leaves the SM (inference); training additionally
// streams it out (pre- and post-GELU, bf16) because the backward pass needs it.
//
// One CTA pair (cta_group::2) owns 256 tokens.  The hidden dimension is processed in 16 chunks of 128:
//   GEMM1_j : D1[b] (TMEM, 128 cols)  = A[256 x 128] . W1[:, chunk j]            (K = 128)
//   epi1_j  : D1[b] -> +b1 -> gelu -> bf16 -> shared memory, written directly in the canonical K-major SWIZZLE_128B
//             operand layout (each thread owns one row: 4 x 16-byte chunks per 32 columns, chunk index XOR row&7)
//   GEMM2_j : D2 (TMEM, 128 cols)    += H_j[256 x 128] . W2[chunk j, :]          (K = 128)
// The single MMA-issuing thread runs the two chains (G1 of the next chunks, G2 of the converted ones) in whatever order
// their inputs become ready (mbarrier.test_wait polling), so the tensor pipe works ahead while the epilogue warps
// convert: bias + GELU on packed fp32 pairs (FADD2 / FMUL2 / FFMA2).  Final epilogue (csrc/row_epilogue.cuh):
// D2 + b2 + residual, single-pass full-row LayerNorm, fp32 residual stream and bf16 operand for the next GEMM, every
// global access through a warp-private 32 x 32 transpose tile borrowed from the then idle H buffers.
//
// Warp roles (640 threads): 0 / 3 TMA producers (A tile once per tile and W1 chunks, 4-deep ring / W2 chunks, 3-deep;
// each CTA stages its own 128 A rows and its half of every weight chunk), 1 MMA issuer (leader CTA), 2 TMEM allocator,
// 4..19 epilogue: two groups of 8 (two per TMEM lane quadrant, 64 columns each); group g converts chunks j = g (mod 2),
// so the groups run half a period apart and one group's TMEM-load / barrier latency hides under the other's math.
// Every cross-CTA hand-off is an mbarrier: tcgen05.commit multicasts "slot free / accumulator ready" to both CTAs,
// epilogue warps of both CTAs arrive remotely on the leader's "D1 drained / H written" barriers.
#pragma once
#include "gemm_tcgen05.cuh"
#include "row_epilogue.cuh"

namespace smd {

struct FfnFusedArgs {
  const float* b1;               // [Md]
  const float* b2;               // [128]
  const float* residual;         // fp32 [M][128] (may alias out_f32)
  float* out_f32;                // fp32 [M][128]
  const float* ln_gamma;         // [128] LayerNorm applied to the new residual stream -> out_bf16
  const float* ln_beta;
  __nv_bfloat16* out_bf16;       // bf16 [M][128]
  __nv_bfloat16* hidden_pre;     // bf16 [M][Md] pre-GELU (training) or null
  __nv_bfloat16* hidden;         // bf16 [M][Md] post-GELU (training) or null
  int M, Md;
};

struct FfnSmem {
  static constexpr int kA = 32768;          // [2 k-blocks][128 rows][128 B]
  static constexpr int kW = 16384;          // per-CTA half of a weight chunk: [2 k-blocks][64 k][64 n]
  static constexpr int kW1Stages = 4, kW2Stages = 3, kHStages = 2;
  static constexpr int kH = 32768;          // [2 k-blocks][128 rows][128 B]
  static constexpr int offA = 0;
  static constexpr int offW1 = offA + kA;
  static constexpr int offW2 = offW1 + kW1Stages * kW;
  static constexpr int offH = offW2 + kW2Stages * kW;
  static constexpr int offBar = offH + kHStages * kH;
  static constexpr int kBarBytes = 256;
  static constexpr int offScr = offBar + kBarBytes;
  static constexpr int kEpiWarps = 16;       // four per TMEM lane quadrant, 32 of a chunk's 128 columns each
  static constexpr int kScrBytes = kEpiWarps * 32 * 2 * 4;
  static constexpr int kTotal = offScr + kScrBytes + 1024;
  static constexpr int kThreads = 128 + 32 * kEpiWarps;
};

// kTrain: also stream the pre- and post-GELU hidden activations out (bf16) for the backward pass
template <bool kTrain>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FfnSmem::kThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const FfnFusedArgs p) {
  using S = FfnSmem;
  extern __shared__ uint8_t ffn_smem_raw[];
  uint8_t* smem = ffn_smem_raw + ((1024u - (smem_u32(ffn_smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::offBar);
  uint64_t* a_full = bars + 0;
  uint64_t* a_empty = bars + 1;
  uint64_t* d2_full = bars + 2;
  uint64_t* d2_empty = bars + 3;
  uint64_t* d1_full = bars + 4;     // [2]
  uint64_t* d1_empty = bars + 6;    // [2]
  uint64_t* h_full = bars + 8;      // [2]
  uint64_t* h_empty = bars + 10;    // [2]
  uint64_t* w1_full = bars + 12;    // [kW1Stages]
  uint64_t* w1_empty = w1_full + S::kW1Stages;
  uint64_t* w2_full = w1_empty + S::kW1Stages;
  uint64_t* w2_empty = w2_full + S::kW2Stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w2_empty + S::kW2Stages);
  static_assert((12 + 2 * S::kW1Stages + 2 * S::kW2Stages) * 8 + 8 <= S::kBarBytes, "barrier block too small");

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_tiles = (p.M + 255) / 256;
  const int group = blockIdx.x / 2, num_groups = gridDim.x / 2;
  const int nchunks = p.Md / 128;

  pdl_trigger();
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < S::kW1Stages; ++i) { mbar_init(&w1_full[i], 1); mbar_init(&w1_empty[i], 1); }
    for (int i = 0; i < S::kW2Stages; ++i) { mbar_init(&w2_full[i], 1); mbar_init(&w2_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], S::kEpiWarps);   // one 8-warp group per CTA serves a buffer
      mbar_init(&h_full[i], S::kEpiWarps); mbar_init(&h_empty[i], 1);
    }
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
    // ===================== TMA producer: A tiles and W1 chunks =====================
    if (elect_one()) {
      const uint32_t a_full_l = mapa_shared(smem_u32(a_full), 0);
      uint32_t n1 = 0, nt = 0;
      for (int tile = group; tile < num_tiles; tile += num_groups, ++nt) {
        const int m_row0 = tile * 256 + static_cast<int>(rank) * 128;
        mbar_wait(a_empty, (nt & 1u) ^ 1u);
        if (leader) mbar_arrive_expect_tx(a_full, 2u * S::kA);
        for (int kb = 0; kb < 2; ++kb) tma_load_2d_2sm(&tmA, a_full_l, smem + S::offA + kb * 16384, 64 * kb, m_row0);
        for (int j = 0; j < nchunks; ++j, ++n1) {
          const uint32_t s = n1 % S::kW1Stages;
          mbar_wait(&w1_empty[s], ((n1 / S::kW1Stages) & 1u) ^ 1u);
          if (leader) mbar_arrive_expect_tx(&w1_full[s], 2u * S::kW);
          const uint32_t bar = mapa_shared(smem_u32(&w1_full[s]), 0);
          uint8_t* dst = smem + S::offW1 + s * S::kW;
          for (int kb = 0; kb < 2; ++kb)   // W1 is [K = 128][N = Md]: box = 64 n x 64 k
            tma_load_2d_2sm(&tmW1, bar, dst + kb * 8192, j * 128 + static_cast<int>(rank) * 64, 64 * kb);
        }
      }
    }
  } else if (warp == 3) {
    // ===================== TMA producer: W2 chunks (its own thread, so a full W2 ring never holds W1 back) ==========
    if (elect_one()) {
      uint32_t n2 = 0;
      for (int tile = group; tile < num_tiles; tile += num_groups) {
        for (int j = 0; j < nchunks; ++j, ++n2) {
          const uint32_t s = n2 % S::kW2Stages;
          mbar_wait(&w2_empty[s], ((n2 / S::kW2Stages) & 1u) ^ 1u);
          if (leader) mbar_arrive_expect_tx(&w2_full[s], 2u * S::kW);
          const uint32_t bar = mapa_shared(smem_u32(&w2_full[s]), 0);
          uint8_t* dst = smem + S::offW2 + s * S::kW;
          for (int kb = 0; kb < 2; ++kb)   // W2 is [K = Md][N = 128]
            tma_load_2d_2sm(&tmW2, bar, dst + kb * 8192, static_cast<int>(rank) * 64, j * 128 + 64 * kb);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one lane) =====================
    if (leader && elect_one()) {
      const uint32_t idesc = make_idesc_bf16(256, 128, 0, 1);   // A / H K-major, weights MN-major
      const uint32_t sA = smem_u32(smem + S::offA);
      uint32_t n1 = 0, n2 = 0, nd1 = 0, nh = 0, nt = 0;
      for (int tile = group; tile < num_tiles; tile += num_groups, ++nt) {
        mbar_wait(a_full, nt & 1u);
        // The two chains -- first GEMM of chunk g1 (needs its W1 slot and the accumulator buffer chunk g1 - 2 used) and
        // second GEMM of chunk g2 (needs the epilogue's H_g2 and its W2 slot) -- are issued in whatever order their
        // inputs become ready: the two epilogue groups then drift half a period apart instead of the later one waiting
        // behind the other group's second GEMM every chunk.
        int g1 = 0, g2 = 0;
        unsigned long long t0 = global_timer_ns();
        uint32_t spins = 0;
        while (g2 < nchunks) {
          bool moved = false;
          if (g1 < nchunks) {
            const uint32_t s = n1 % S::kW1Stages, b = nd1 & 1u;
            if (mbar_test(&w1_full[s], (n1 / S::kW1Stages) & 1u) && mbar_test_cluster(&d1_empty[b], ((nd1 >> 1) & 1u) ^ 1u)) {
              tcgen05_fence_after();
              const uint32_t sW = smem_u32(smem + S::offW1 + s * S::kW);
#pragma unroll
              for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t ad = make_smem_desc_sw128(sA + kb * 16384 + k * 32, 0u, 1024u);
                  const uint64_t bd = make_smem_desc_sw128(sW + kb * 8192 + k * 2048, 8192u, 1024u);
                  umma_bf16<2>(tmem_base + b * 128u, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                }
              umma_commit<2>(&w1_empty[s]);
              umma_commit<2>(&d1_full[b]);
              if (g1 + 1 == nchunks) umma_commit<2>(a_empty);   // the A tile is no longer needed: the producer may fetch the next one
              ++n1; ++nd1; ++g1;
              moved = true;
            }
          }
          if (g2 < g1) {
            const uint32_t hb = nh & 1u, s2 = n2 % S::kW2Stages;
            if (mbar_test_cluster(&h_full[hb], (nh >> 1) & 1u) && mbar_test(&w2_full[s2], (n2 / S::kW2Stages) & 1u)) {
              if (g2 == 0) mbar_wait_cluster(d2_empty, (nt & 1u) ^ 1u);
              tcgen05_fence_after();
              const uint32_t sH = smem_u32(smem + S::offH + hb * S::kH);
              const uint32_t sW = smem_u32(smem + S::offW2 + s2 * S::kW);
#pragma unroll
              for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t ad = make_smem_desc_sw128(sH + kb * 16384 + k * 32, 0u, 1024u);
                  const uint64_t bd = make_smem_desc_sw128(sW + kb * 8192 + k * 2048, 8192u, 1024u);
                  umma_bf16<2>(tmem_base + 256u, ad, bd, idesc, (g2 | kb | k) != 0 ? 1u : 0u);
                }
              umma_commit<2>(&w2_empty[s2]);
              umma_commit<2>(&h_empty[hb]);
              if (g2 + 1 == nchunks) umma_commit<2>(d2_full);
              ++n2; ++nh; ++g2;
              moved = true;
            }
          }
          if (moved) { spins = 0; t0 = global_timer_ns(); }
          else if ((++spins & 0x3FFu) == 0 && global_timer_ns() - t0 > SMD_WAIT_LIMIT_NS) __trap();
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps =====================
    const uint32_t q = warp & 3u;                      // TMEM lane quadrant
    const int eg = static_cast<int>(warp - 4u) >> 2;   // final epilogue: which 32-column quarter of the 128-wide output
    const int grp = eg >> 1, half = eg & 1;            // chunk loop: group (chunk parity) and 64-column half of a chunk
    float* scr_all = reinterpret_cast<float*>(smem + S::offScr);
    float* scr = scr_all + (warp - 4u) * 64;
    const uint32_t d1_empty_l0 = mapa_shared(smem_u32(&d1_empty[0]), 0), d1_empty_l1 = mapa_shared(smem_u32(&d1_empty[1]), 0);
    const uint32_t h_full_l0 = mapa_shared(smem_u32(&h_full[0]), 0), h_full_l1 = mapa_shared(smem_u32(&h_full[1]), 0);
    const uint32_t d2_empty_l = mapa_shared(smem_u32(d2_empty), 0);
    const uint32_t r_in_tile = q * 32u + lane;         // row of this thread inside the CTA's 128-row tile
    const uint32_t swz = r_in_tile & 7u;
    const int c0 = eg * 32;                            // this warp's columns inside a chunk / inside the 128-wide output
    uint32_t ng = 0, nt = 0;
    for (int tile = group; tile < num_tiles; tile += num_groups, ++nt) {
      const int row = tile * 256 + static_cast<int>(rank) * 128 + static_cast<int>(r_in_tile);
      [[maybe_unused]] const bool row_ok = row < p.M;
      // chunk j belongs to warp group j & 1 (D1 buffer and H buffer j & 1 as well): while one group converts its chunk
      // the other is half a period away, so TMEM-load / barrier latencies of one hide under the math of the other
      for (int jj = 0; jj < nchunks / 2; ++jj, ++ng) {
        const int j = 2 * jj + grp;
        const uint32_t ph = ng & 1u;
        uint8_t* hrow = smem + S::offH + grp * S::kH + half * 16384 + r_in_tile * 128u;
        mbar_wait(&d1_full[grp], ph);
        tcgen05_fence_after();
        __syncwarp();
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          const int cc = half * 64 + ps * 32;          // column inside the chunk
          const int gcol = j * 128 + cc;               // hidden-unit index
          float4 bq[8];
          {
            const float4* b4 = reinterpret_cast<const float4*>(p.b1 + gcol);
#pragma unroll
            for (int i = 0; i < 8; ++i) bq[i] = __ldg(b4 + i);
          }
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + grp * 128u + static_cast<uint32_t>(cc), r);
          tmem_ld_wait();
          if (ps == 1) {
            // D1 is in registers: hand the accumulator back before the math
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(grp ? d1_empty_l1 : d1_empty_l0);
          }
          // bias add and GELU on packed fp32 pairs (FADD2 / FMUL2 / FFMA2): half the issue slots of the scalar form
          uint64_t v2[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v2[2 * i] = f32x2_add(f32x2_pack(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), f32x2_pack(bq[i].x, bq[i].y));
            v2[2 * i + 1] = f32x2_add(f32x2_pack(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), f32x2_pack(bq[i].z, bq[i].w));
          }
          if (kTrain && p.hidden_pre != nullptr && row_ok) {
            uint4* dst = reinterpret_cast<uint4*>(p.hidden_pre + static_cast<size_t>(row) * p.Md + gcol);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t w[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float a, b;
                f32x2_unpack(v2[4 * i + k], a, b);
                w[k] = pack_bf16x2(a, b);
              }
              dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
          uint4 pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float a, b;
              f32x2_unpack(gelu_tanh_x2(v2[4 * i + k]), a, b);
              w[k] = pack_bf16x2(a, b);
            }
            pk[i] = make_uint4(w[0], w[1], w[2], w[3]);
          }
          // the previous chunk of this buffer (two chunks back) must have been consumed by its second GEMM
          if (ps == 0) mbar_wait(&h_empty[grp], ph ^ 1u);
          // canonical K-major SWIZZLE_128B: 16-byte chunk c of row r lives at r * 128 + ((c ^ (r & 7)) << 4);
          // columns [cc, cc + 32) of the chunk are k-block cc / 64 (= half), 16-byte chunks (cc % 64) / 8 .. + 3
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t ch = static_cast<uint32_t>(ps * 4 + i);
            *reinterpret_cast<uint4*>(hrow + ((ch ^ swz) << 4)) = pk[i];
          }
          if (kTrain && p.hidden != nullptr && row_ok) {
            uint4* dst = reinterpret_cast<uint4*>(p.hidden + static_cast<size_t>(row) * p.Md + gcol);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = pk[i];
          }
        }
        fence_proxy_async_smem();    // this thread's H stores become visible to the tensor core's (async proxy) reads
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(grp ? h_full_l1 : h_full_l0);
      }
      // ---------------- final epilogue: D2 + b2 + residual -> LayerNorm (csrc/row_epilogue.cuh) ----------------
      // the H region is idle between this tile's last second GEMM (d2_full) and the next tile's first chunk: every
      // warp borrows 4 KB of it as its transpose tile
      {
        const RowEpiArgs ea{p.b2, p.residual, p.out_f32, p.ln_gamma, p.ln_beta, p.out_bf16, p.M};
        const int row0 = tile * 256 + static_cast<int>(rank) * 128 + static_cast<int>(q) * 32;
        float* tsc = reinterpret_cast<float*>(smem + S::offH) + (warp - 4u) * 1024u;
        float4 pre[8];
        row_epi_prefetch(ea, row0, c0, lane, pre);
        mbar_wait(d2_full, nt & 1u);
        tcgen05_fence_after();
        __syncwarp();
        row_epi_stage(pre, tsc, lane);
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + 256u + static_cast<uint32_t>(c0), r);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(d2_empty_l);
        row_epi_finish(ea, r, tsc, scr, scr_all, 64, q, row0, c0, lane);
        // the next tile's H stores of any warp may land in any warp's transpose tile
        asm volatile("bar.sync 5, %0;" ::"n"(32 * S::kEpiWarps) : "memory");
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
