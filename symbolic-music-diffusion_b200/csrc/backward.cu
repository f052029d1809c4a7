// Backward pass of the DDPM objective (placeholder until the hand-written backward lands in this file).
#include "train.cuh"

using namespace smd;

extern "C" int smd_ddpm_grads(smd_plan* plan, const float* params, const float* x0, const float* used_alpha,
                              const float* eps, int batch, int global_batch, float* grads, float* loss_sum,
                              smd_stream_t stream) {
  (void)plan; (void)params; (void)x0; (void)used_alpha; (void)eps; (void)batch; (void)global_batch; (void)grads;
  (void)loss_sum; (void)stream;
  set_error("smd_ddpm_grads: backward pass not built yet");
  return SMD_ERR_STATE;
}
