// Saved-activation bookkeeping for the training step (forward keeps what backward needs).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include <functional>
#include <string>
#include <vector>
#include "../../include/smd.h"
#include "gemm_host.cuh"

namespace smd {

static constexpr int kEt = 128;

struct TrainState {
  bool enabled = false;
  int L = 0, K = 0, H = 0, Md = 0, C = 0, S = 0, B = 0;
  size_t Mp = 0;
  // byte offsets into the workspace
  std::vector<size_t> off_h;           // fp32 [Mp][128] x (2L+1): residual stream after embed / attn / ffn
  std::vector<size_t> off_a1, off_a2;  // bf16 [Mp][128] x L: LN1 / LN2 outputs (GEMM operands)
  std::vector<size_t> off_qkv;         // fp32 [Mp][384] x L
  std::vector<size_t> off_probs;       // fp32 [B][H][32][32] x L
  std::vector<size_t> off_o;           // bf16 [Mp][128] x L
  std::vector<size_t> off_hidden_pre;  // bf16 [Mp][Md] x L (pre-GELU)
  std::vector<size_t> off_hidden;      // bf16 [Mp][Md] x L (post-GELU)
  size_t off_a_post = 0;               // bf16 [Mp][128]
  std::vector<size_t> off_u;           // fp32 [Mp][Md] x (K+1)
  std::vector<size_t> off_r1;          // fp32 [Mp][Md] x K
  std::vector<size_t> off_act_a, off_act_b;  // bf16 [Mp][Md] x K
  size_t off_act_out = 0;              // bf16 [Mp][Md]
  // FiLM generator saves, per block: enc is shared
  std::vector<size_t> off_e1pre, off_e1, off_e2;  // fp32 [B][512] x K
  // backward scratch
  size_t off_g32a = 0, off_g32b = 0;   // fp32 [Mp][Md] gradient ping-pong (tail) -- also used [Mp][128]-wide in the trunk
  // bf16 gradient operands of the tail, one buffer per use (their dW GEMMs run on the weight-gradient stream):
  std::vector<size_t> off_du16;        // bf16 [Mp][Md] x (K+1): gradient wrt the residual stream u_j entering block j
  std::vector<size_t> off_dr16t;       // bf16 [Mp][Md] x K: gradient wrt r1 (after LayerNorm-b backward)
  size_t off_dh = 0, off_dh2 = 0;      // fp32 [Mp][128]
  // per-layer bf16 gradient operands: the dW GEMMs that read them run on their own stream, so no buffer is
  // rewritten within one backward pass
  std::vector<size_t> off_dh16a;       // bf16 [Mp][128] x L: gradient entering layer l (from LN1 of l+1 / post-LN)
  std::vector<size_t> off_dh16b;       // bf16 [Mp][128] x L: gradient at the attention output (from LN2)
  std::vector<size_t> off_dr16;        // bf16 [Mp][Md]  x L: gradient at the FFN pre-activation
  std::vector<size_t> off_dqkv16;      // bf16 [Mp][384] x L
  size_t off_dh16_in = 0;              // bf16 [Mp][128]: LN1 output of layer 0 (unused operand)
  size_t off_dqkv32 = 0;               // fp32 [Mp][384]
  size_t off_dpred16 = 0;              // bf16 [Mp][Cp64]
  size_t off_dpred32 = 0;              // fp32 [Mp][C]
  size_t off_dss = 0;                  // fp32 [B][2Md]
  size_t off_de = 0, off_de2 = 0;      // fp32 [B][512] x2
  size_t off_loss = 0;                 // fp32 [B]
  size_t off_loss_ctr = 0;             // u32: blocks of the loss kernel that have finished
  size_t off_wplain = 0;               // bf16 plain (in,out) copies of every GEMM weight (dX operands)
  std::vector<size_t> off_w_qkv, off_w_o, off_w_ffn1, off_w_ffn2, off_w_a, off_w_b;
  size_t off_w_post = 0, off_w_out = 0, off_w_in = 0;
  std::vector<size_t> off_w_ss;        // bf16 [512][2Md] per block
  size_t off_e2_16 = 0, off_dss16 = 0; // bf16 [Bp][512], [Bp][2Md]
  // backward GEMM descriptors (built by train_bind)
  std::vector<GemmOp> dWb, dXb, dWa, dXa, dWss, dXss;
  std::vector<GemmOp> dW2, dX2, dW1, dX1, dWo, dXo, dWqkv, dXqkv;
  GemmOp dWout, dXout, dWpost, dXpost, dWin;

  float* h(uint8_t* ws, int i) const { return reinterpret_cast<float*>(ws + off_h[i]); }
  __nv_bfloat16* a1(uint8_t* ws, int l) const { return reinterpret_cast<__nv_bfloat16*>(ws + off_a1[l]); }
  __nv_bfloat16* a2(uint8_t* ws, int l) const { return reinterpret_cast<__nv_bfloat16*>(ws + off_a2[l]); }
  float* qkv(uint8_t* ws, int l) const { return reinterpret_cast<float*>(ws + off_qkv[l]); }
  float* probs(uint8_t* ws, int l) const { return reinterpret_cast<float*>(ws + off_probs[l]); }
  __nv_bfloat16* o(uint8_t* ws, int l) const { return reinterpret_cast<__nv_bfloat16*>(ws + off_o[l]); }
  __nv_bfloat16* hidden_pre(uint8_t* ws, int l) const { return reinterpret_cast<__nv_bfloat16*>(ws + off_hidden_pre[l]); }
  __nv_bfloat16* hidden(uint8_t* ws, int l) const { return reinterpret_cast<__nv_bfloat16*>(ws + off_hidden[l]); }
  __nv_bfloat16* a_post(uint8_t* ws) const { return reinterpret_cast<__nv_bfloat16*>(ws + off_a_post); }
  float* u(uint8_t* ws, int k) const { return reinterpret_cast<float*>(ws + off_u[k]); }
  float* r1(uint8_t* ws, int k) const { return reinterpret_cast<float*>(ws + off_r1[k]); }
  __nv_bfloat16* act_a(uint8_t* ws, int k) const { return reinterpret_cast<__nv_bfloat16*>(ws + off_act_a[k]); }
  __nv_bfloat16* act_b(uint8_t* ws, int k) const { return reinterpret_cast<__nv_bfloat16*>(ws + off_act_b[k]); }
  __nv_bfloat16* act_out(uint8_t* ws) const { return reinterpret_cast<__nv_bfloat16*>(ws + off_act_out); }
  template <typename T>
  T* at(uint8_t* ws, size_t off) const { return reinterpret_cast<T*>(ws + off); }
};

// `add(name, bytes)` reserves a 1024-aligned workspace region and returns its byte offset.
void train_workspace(TrainState& ts, const smd_config& c, int Mp, int K,
                     const std::function<size_t(const std::string&, size_t)>& add);

// Re-point a K-major A operand at another activation buffer of the same geometry.
inline bool retarget_a(GemmOp* op, const void* A, uint64_t rows) {
  return make_tmap_bf16(&op->tmA, A, rows, static_cast<uint64_t>(op->K), 128);
}

}  // namespace smd
