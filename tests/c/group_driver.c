/* C99 driver of the multi-GPU entry points of include/gbm_b200.h -- what a non-Python host (the Julia shim's
 * ccall sequence) does: one process, a local group of N GPUs, the whole gwaslmm in one collective call.
 * Usage: group_driver N_GPUS n p seed kind grm_type model  < y (n doubles, binary, stdin)
 * Output (stdout, text; every line starts with "gbm", NCCL may print its banner there too):
 * "gbm l ploidy packed", then l lines "gbm idx z". */
#include <stdio.h>
#include <stdlib.h>

#include "gbm_b200.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc__ = (call);                                                       \
    if (rc__ != GBM_OK) {                                                    \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, gbm_last_error()); \
      return 10 + rc__;                                                      \
    }                                                                        \
  } while (0)

int main(int argc, char** argv) {
  if (argc != 8) return 2;
  const int n_gpus = atoi(argv[1]);
  const int64_t n = atoll(argv[2]), p = atoll(argv[3]);
  const uint64_t seed = (uint64_t)atoll(argv[4]);
  const int kind = atoi(argv[5]), grm_type = atoi(argv[6]), model = atoi(argv[7]);
  double* y = (double*)malloc(sizeof(double) * (size_t)n);
  double* z = (double*)malloc(sizeof(double) * (size_t)p);
  int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)p);
  if (!y || !z || !idx) return 3;
  if (fread(y, sizeof(double), (size_t)n, stdin) != (size_t)n) return 4;
  gbm_group* grp = 0;
  gbm_sharded* m = 0;
  int world = 0, n_local = 0, first = -1, packed = -1;
  int64_t l = 0, j;
  gbm_gwas_timing tm;
  CHECK(gbm_group_create_local(n_gpus, 0, &grp));
  CHECK(gbm_group_info(grp, &world, &n_local, &first));
  if (world != n_gpus || n_local != n_gpus || first != 0) return 5;
  CHECK(gbm_sharded_generate(grp, seed, n, p, kind, 1, &m, &packed));
  CHECK(gbm_sharded_gwas(m, y, model, grm_type, 0, z, 0, 0, 0, 0, 0, 0, idx, &l, 0, &tm));
  printf("gbm %lld %d %d\n", (long long)l, (int)tm.ploidy, packed);
  for (j = 0; j < l; ++j) printf("gbm %lld %.17g\n", (long long)idx[j], z[idx[j] - 1]);
  CHECK(gbm_sharded_free(m));
  CHECK(gbm_group_free(grp));
  free(y);
  free(z);
  free(idx);
  return 0;
}
