/* A plain C11 client of include/smd.h: proves the boundary is a C ABI (no C++ / torch types in any signature).
 * Built and run by tests/test_abi.py::test_plain_c_client with gcc; needs no GPU (only plan / layout calls). */
#include <stdio.h>
#include <string.h>
#include "smd.h"

int main(void) {
  if (smd_version() < 100) { fprintf(stderr, "version\n"); return 1; }
  smd_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.arch = SMD_ARCH_TRANSFORMER_DDPM;
  cfg.num_layers = 6; cfg.num_heads = 8; cfg.num_mlp_layers = 2; cfg.mlp_dims = 2048;
  cfg.seq_len = 32; cfg.channels = 42; cfg.max_batch = 128; cfg.cta_group = 2; cfg.training = 1;
  smd_plan* plan = NULL;
  if (smd_plan_create(&cfg, &plan) != SMD_OK) { fprintf(stderr, "create: %s\n", smd_last_error()); return 2; }
  const int n = smd_num_tensors(plan);
  long long total = 0, first = 0, count = 0;
  for (int i = 0; i < n; ++i) {
    char name[96];
    long long off;
    int shape[4], ndim;
    if (smd_tensor_info(plan, i, name, (int)sizeof name, &off, shape, &ndim) != SMD_OK) return 3;
    long long sz = 1;
    for (int d = 0; d < ndim; ++d) sz *= shape[d];
    total += sz;
    if (i == 0 && (strcmp(name, "in.kernel") != 0 || off != 0)) return 4;
  }
  if (smd_grads_tail_range(plan, &first, &count) != SMD_OK) return 5;
  /* the reference's report_model count for the base configuration (utils/train_utils.py:121-131) */
  printf("tensors=%d params=%lld arena=%lld workspace=%zu tail=[%lld,+%lld)\n", n, total, smd_arena_floats(plan),
         smd_workspace_bytes(plan), first, count);
  /* an invalid configuration must fail with SMD_ERR_INVALID and a message, not crash */
  cfg.seq_len = 16;
  smd_plan* bad = NULL;
  if (smd_plan_create(&cfg, &bad) != SMD_ERR_INVALID || strlen(smd_last_error()) == 0) return 6;
  smd_plan_destroy(plan);
  return total == 25579946LL ? 0 : 7;
}
