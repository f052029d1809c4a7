// Host side of the tcgen05 GEMM: tensor-map construction (cuTensorMapEncodeTiled through the runtime's
// driver-entry-point lookup, so libsmd.so has no link-time dependency on libcuda) and the launcher.
#pragma once
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <atomic>
#include <string>
#include "gemm_tcgen05.cuh"
#include "ffn_fused.cuh"
#include "attn_block.cuh"
#include "pdl_launch.cuh"

namespace smd {

extern std::atomic<long long> g_launches;
void set_error(const std::string& msg);

inline PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  return fn;
}

// bf16 row-major matrix [rows][cols]; box = box_rows x 64 columns (128 bytes, SWIZZLE_128B).
inline bool make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  auto fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return false; }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: code " + std::to_string(static_cast<int>(r)) + " rows=" +
              std::to_string(rows) + " cols=" + std::to_string(cols) + " box_rows=" + std::to_string(box_rows));
    return false;
  }
  return true;
}

struct GemmOp {
  CUtensorMap tmA, tmB;
  CUtensorMap tmA_lo, tmB_lo;   // strict-precision mode: the lo halves of both operands (has_lo)
  bool has_lo = false;
  int N = 0, K = 0, BN = 0, cg = 1, a_mn = 0, b_mn = 0, k_splits = 1;
};

inline int choose_bn(int N, int cg) {
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  const int q = 16 * cg;
  if (N < 256) return ((N + q - 1) / q) * q;
  return 256;
}

// A: K-major [a_rows][K] (a_mn=0) or MN-major [K][a_rows] (a_mn=1); same for B with N rows.
inline bool make_gemm_op(GemmOp* op, const void* A, uint64_t a_rows, const void* B, uint64_t b_rows_total, int N,
                         int K, int BN, int cg, int a_mn, int b_mn, uint64_t k_rows_a = 0, uint64_t k_rows_b = 0,
                         size_t lo_bytes = 0) {
  if (lo_bytes) {
    GemmOp lo;
    if (!make_gemm_op(&lo, static_cast<const uint8_t*>(A) + lo_bytes, a_rows, static_cast<const uint8_t*>(B) + lo_bytes,
                      b_rows_total, N, K, BN, cg, a_mn, b_mn, k_rows_a, k_rows_b, 0)) return false;
    op->tmA_lo = lo.tmA; op->tmB_lo = lo.tmB; op->has_lo = true;
  }
  op->N = N; op->K = K; op->BN = BN; op->cg = cg; op->a_mn = a_mn; op->b_mn = b_mn; op->k_splits = 1;
  if (K % 64 != 0) { set_error("GEMM K must be a multiple of 64"); return false; }
  if (BN % (16 * cg) != 0 || BN > 256 || BN < 16 * cg) { set_error("bad BN " + std::to_string(BN)); return false; }
  if (b_mn && (BN / cg) % 64 != 0) { set_error("MN-major B needs BN/cta_group % 64 == 0"); return false; }
  bool ok;
  if (!a_mn) ok = make_tmap_bf16(&op->tmA, A, a_rows, static_cast<uint64_t>(K), 128);
  else ok = make_tmap_bf16(&op->tmA, A, k_rows_a ? k_rows_a : static_cast<uint64_t>(K), a_rows, 64);
  if (!ok) return false;
  if (!b_mn) ok = make_tmap_bf16(&op->tmB, B, b_rows_total, static_cast<uint64_t>(K), static_cast<uint32_t>(BN / cg));
  else ok = make_tmap_bf16(&op->tmB, B, k_rows_b ? k_rows_b : static_cast<uint64_t>(K), b_rows_total, 64);
  return ok;
}

inline int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// feature bits a launch needs, and whether it qualifies for a specialised (fast-path-only) instantiation
inline uint32_t epi_needs(const GemmEpilogue& e) {
  uint32_t f = 0;
  if (e.bias) f |= F_BIAS;
  if (e.residual) f |= F_RES;
  if (e.out_f32) f |= F_F32;
  if (e.out_bf16) f |= F_BF16;
  if (e.out_bf16_pre) f |= F_PRE;
  if (e.row_stats) f |= F_STATS;
  if (e.ln_gamma) f |= F_LN;
  if (e.gelu_grad_of) f |= F_GG;
  if (e.atomic_out) f |= F_ATOMIC;
  if (e.act != ACT_NONE) f |= F_ACT;
  if (e.out_scale != 0.0f && e.out_scale != 1.0f) f |= F_SCALE;
  return f;
}
inline bool epi_clean(const GemmOp& op, const GemmEpilogue& e) {
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return (op.N % 32 == 0) && (op.BN % 32 == 0) && (!e.residual || ((e.ld_res & 3) == 0 && a16(e.residual))) &&
         (!e.out_f32 || ((e.ld_f32 & 3) == 0 && a16(e.out_f32))) &&
         ((!e.out_bf16 && !e.out_bf16_pre) || (e.ld_bf16 & 7) == 0) && (!e.out_bf16 || a16(e.out_bf16)) &&
         (!e.out_bf16_pre || a16(e.out_bf16_pre)) && (!e.gelu_grad_of || ((e.ld_gg & 7) == 0 && a16(e.gelu_grad_of))) &&
         (!e.bias || a16(e.bias)) && (!e.ln_gamma || (a16(e.ln_gamma) && a16(e.ln_beta)));
}

// Number of epilogue warps of the instantiation launch_gemm will pick (mirrors launch_gemm_cg): the per-tile statistics
// partials (GemmEpilogue::stats_part) have (epi warps / 4) slots per n-tile.
inline int epi_warps_for(const GemmOp& op, const GemmEpilogue& ep) {
  if (ep.lnf_part != nullptr || ep.lo_delta != 0 || !epi_clean(op, ep)) return 8;
  const uint32_t need = epi_needs(ep);
  const bool short_k = op.K <= 256;
  if ((need & ~kEpiF32) == 0) return short_k ? 12 : 8;
  if ((need & ~kEpiAtomic) == 0 || (need & ~kEpiF32Res) == 0) return 8;
  if ((need & ~kEpiAct) == 0) return short_k ? 12 : 8;
  return 8;
}
inline int stats_slots_for(const GemmOp& op, const GemmEpilogue& ep) {
  return ((op.N + op.BN - 1) / op.BN) * (epi_warps_for(op, ep) / 4);
}

template <int kCG, uint32_t kF, int kEW = 8>
inline cudaError_t launch_gemm_inst(const GemmOp& op, int M, const GemmEpilogue& ep, cudaStream_t st) {
  using SM = GemmSmem<kCG, kEW, lnf_kind(kF), scr_floats(kF)>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<kCG, kF, kEW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SM::kTotal);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  GemmShape sh;
  sh.M = M; sh.N = op.N; sh.K = op.K; sh.BN = op.BN; sh.a_mn = op.a_mn; sh.b_mn = op.b_mn;
  // normalise the split count so that every split owns at least one 64-wide k block
  const int num_kb = op.K / kBK;
  int splits = op.k_splits < 1 ? 1 : op.k_splits;
  if (splits > num_kb) splits = num_kb;
  const int per = (num_kb + splits - 1) / splits;
  splits = (num_kb + per - 1) / per;
  sh.k_splits = splits;
  const int rows_per_tile = kBM * kCG;
  const int tiles = ((M + rows_per_tile - 1) / rows_per_tile) * ((op.N + op.BN - 1) / op.BN) * splits;
  int groups = device_sm_count() / kCG;
  if (tiles < groups) groups = tiles;
  if constexpr ((kF & F_LNF) != 0 && (kF & F_RAGGED) == 0) {
    // the n-tiles of one row block exchange LayerNorm partials: keep them in the same scheduling round
    const int num_n = (op.N + op.BN - 1) / op.BN;
    if (groups > num_n) groups -= groups % num_n;
  }
  if (groups < 1) groups = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(groups * kCG));
  cfg.blockDim = dim3(SM::kThreads);
  cfg.dynamicSmemBytes = SM::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = kCG;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<kCG, kF, kEW>, op.tmA, op.tmB, sh, ep);
}

template <int kCG>
inline cudaError_t launch_gemm_cg(const GemmOp& op, int M, const GemmEpilogue& ep, cudaStream_t st) {
  if (ep.lnf_part != nullptr) {
    // LN-fused two-pass epilogue (the caller arms it only for full, aligned tiles; see smd_api.cu::arm_lnf)
    if (!epi_clean(op, ep) || op.N % op.BN != 0 || op.BN % 64 != 0 || op.k_splits > 1) return cudaErrorInvalidValue;
    if constexpr (kCG == 2) {   // (the 64 KB parking buffer only fits next to the 32 KB pipeline slots of CTA pairs)
      // epilogue warps per kind: SMD_LNF_WARPS_A / _B = 8 (default) | 12 (three warps per TMEM quadrant, <= 128 registers;
      // measured slower: 280 / 371 us against 217 / 298 us per launch at 32000 tokens)
      static const int wa = [] { const char* v = getenv("SMD_LNF_WARPS_A"); return (v && atoi(v) == 12) ? 12 : 8; }();
      static const int wb = [] { const char* v = getenv("SMD_LNF_WARPS_B"); return (v && atoi(v) == 12) ? 12 : 8; }();
      if (ep.residual != nullptr || ep.out_f32 != nullptr)
        return wb == 12 ? launch_gemm_inst<kCG, kEpiLnfB, 12>(op, M, ep, st) : launch_gemm_inst<kCG, kEpiLnfB, 8>(op, M, ep, st);
      return wa == 12 ? launch_gemm_inst<kCG, kEpiLnfA, 12>(op, M, ep, st) : launch_gemm_inst<kCG, kEpiLnfA, 8>(op, M, ep, st);
    } else {
      return cudaErrorInvalidValue;
    }
  }
  if (ep.lo_delta != 0) return launch_gemm_inst<kCG, kEpiStrict>(op, M, ep, st);   // strict-precision mode (bf16x3)
  const uint32_t need = epi_needs(ep);
  if (epi_clean(op, ep)) {
    auto fits = [&](uint32_t kind) { return (need & ~kind) == 0; };
    // short-K GEMMs are epilogue-bound: give them a third epilogue warp per TMEM quadrant
    const bool short_k = op.K <= 256;
    if (fits(kEpiF32)) return short_k ? launch_gemm_inst<kCG, kEpiF32, 12>(op, M, ep, st) : launch_gemm_inst<kCG, kEpiF32>(op, M, ep, st);
    if (fits(kEpiAtomic)) return launch_gemm_inst<kCG, kEpiAtomic>(op, M, ep, st);
    if (fits(kEpiF32Res)) return launch_gemm_inst<kCG, kEpiF32Res>(op, M, ep, st);
    if (fits(kEpiAct)) return short_k ? launch_gemm_inst<kCG, kEpiAct, 12>(op, M, ep, st) : launch_gemm_inst<kCG, kEpiAct>(op, M, ep, st);
    if (fits(kEpiGG)) return launch_gemm_inst<kCG, kEpiGG>(op, M, ep, st);
    if ((need & F_LN) && fits(kEpiLn) && op.N == op.BN && op.BN <= 128 && op.BN % 64 == 0 && op.k_splits <= 1)
      return launch_gemm_inst<kCG, kEpiLn>(op, M, ep, st);
  }
  return launch_gemm_inst<kCG, kEpiGeneric>(op, M, ep, st);
}

inline cudaError_t launch_gemm(const GemmOp& op, int M, const GemmEpilogue& ep, cudaStream_t st) {
  return op.cg == 2 ? launch_gemm_cg<2>(op, M, ep, st) : launch_gemm_cg<1>(op, M, ep, st);
}


// ---------------------------------------------------------------------------------------------------
// fused FFN (ffn_fused.cuh)
// ---------------------------------------------------------------------------------------------------
struct FfnOp {
  CUtensorMap tmA, tmW1, tmW2;
  bool ok = false;
};
// A: bf16 [rows][128] K-major; W1: bf16 (128, Md) row-major; W2: bf16 (Md, 128) row-major (both MN-major B operands)
inline bool make_ffn_op(FfnOp* op, const void* A, uint64_t rows, const void* W1, const void* W2, int Md) {
  op->ok = make_tmap_bf16(&op->tmA, A, rows, 128, 128) && make_tmap_bf16(&op->tmW1, W1, 128, static_cast<uint64_t>(Md), 64) &&
           make_tmap_bf16(&op->tmW2, W2, static_cast<uint64_t>(Md), 128, 64);
  return op->ok;
}
inline bool ffn_fused_enabled() {
  static const bool on = [] { const char* v = getenv("SMD_FFN_FUSED"); return !(v && v[0] == '0'); }();
  return on;
}
// SMD_FFN_FUSED=2: use the fused kernel at every size and in training too (tests)
inline bool ffn_fused_forced() {
  static const bool on = [] { const char* v = getenv("SMD_FFN_FUSED"); return v && v[0] == '2'; }();
  return on;
}
inline cudaError_t launch_ffn_fused(const FfnOp& op, const FfnFusedArgs& a, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(ffn_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnSmem::kTotal);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ffn_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnSmem::kTotal);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = (a.M + 255) / 256;
  int pairs = device_sm_count() / 2;
  if (tiles < pairs) pairs = tiles;
  if (pairs < 1) pairs = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(2 * pairs));
  cfg.blockDim = dim3(FfnSmem::kThreads);
  cfg.dynamicSmemBytes = FfnSmem::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (a.hidden_pre != nullptr || a.hidden != nullptr) return cudaLaunchKernelEx(&cfg, ffn_fused_kernel<true>, op.tmA, op.tmW1, op.tmW2, a);
  return cudaLaunchKernelEx(&cfg, ffn_fused_kernel<false>, op.tmA, op.tmW1, op.tmW2, a);
}


// ---------------------------------------------------------------------------------------------------
// fused attention block (attn_block.cuh)
// ---------------------------------------------------------------------------------------------------
struct AttnOp {
  CUtensorMap tmA, tmWqkv, tmWo;
  bool ok = false;
};
// A: bf16 [rows][128] K-major; Wqkv: bf16 (128, 384) row-major; Wo: bf16 (128, 128) row-major (MN-major B operands)
inline bool make_attn_op(AttnOp* op, const void* A, uint64_t rows, const void* Wqkv, const void* Wo) {
  op->ok = make_tmap_bf16(&op->tmA, A, rows, 128, 128) && make_tmap_bf16(&op->tmWqkv, Wqkv, 128, 384, 64) &&
           make_tmap_bf16(&op->tmWo, Wo, 128, 128, 64);
  return op->ok;
}
// SMD_ATTN_BLOCK=0 keeps the three-launch path (QKV GEMM, attention kernel, out-projection GEMM)
inline bool attn_block_enabled() {
  static const bool on = [] { const char* v = getenv("SMD_ATTN_BLOCK"); return !(v && v[0] == '0'); }();
  return on;
}
// SMD_ATTN_BLOCK_TRAIN=1: the training forward uses the block kernel too (kTrain: q | k | v, probabilities and the
// attention output are written out for the backward pass).  Off by default: at batch 128 only 16 CTA pairs are busy and
// the scattered saves make the launch 34 us against 25 us for the three-launch path (profiles/r02_bench_train_attn*.json).
inline bool attn_block_train_enabled() {
  static const bool on = [] { const char* v = getenv("SMD_ATTN_BLOCK_TRAIN"); return v && v[0] == '1'; }();
  return on;
}
inline cudaError_t launch_attn_block(const AttnOp& op, const AttnBlockArgs& a, cudaStream_t st) {
  const int dh = 128 / a.H;
  if (dh != 8 && dh != 16) return cudaErrorInvalidValue;
  const bool train = a.qkv_out != nullptr;
  if (train && (a.probs_out == nullptr || a.o_out == nullptr)) return cudaErrorInvalidValue;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_block_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kTotal);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(attn_block_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kTotal);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(attn_block_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kTotal);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(attn_block_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kTotal);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = (a.M + 255) / 256;
  int pairs = device_sm_count() / 2;
  if (tiles < pairs) pairs = tiles;
  if (pairs < 1) pairs = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(2 * pairs));
  cfg.blockDim = dim3(AttnSmem::kThreads);
  cfg.dynamicSmemBytes = AttnSmem::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (train) {
    if (dh == 16) return cudaLaunchKernelEx(&cfg, attn_block_kernel<16, true>, op.tmA, op.tmWqkv, op.tmWo, a);
    return cudaLaunchKernelEx(&cfg, attn_block_kernel<8, true>, op.tmA, op.tmWqkv, op.tmWo, a);
  }
  if (dh == 16) return cudaLaunchKernelEx(&cfg, attn_block_kernel<16, false>, op.tmA, op.tmWqkv, op.tmWo, a);
  return cudaLaunchKernelEx(&cfg, attn_block_kernel<8, false>, op.tmA, op.tmWqkv, op.tmWo, a);
}

}  // namespace smd
