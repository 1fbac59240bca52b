// One object per (kernel kind, mode): nvcc -DFFVD_SPLIT_BUILD -DFFVD_INST_KIND=k -DFFVD_INST_MODE=m -c fused_inst.cu
// (the fused kernel templates dominate the build time; ten of these compile in parallel, see the Makefile).
#include "fused.cuh"
#include "fused_table.cuh"

#define FFVD_CAT2(a, b, c) a##b##_##c
#define FFVD_CAT(a, b, c) FFVD_CAT2(a, b, c)

ffvd_fused_fn FFVD_CAT(ffvd_fused_lookup_, FFVD_INST_KIND, FFVD_INST_MODE)(int rb, int ngw, int nw, int minb) {
  return ffvd_fused_pick<FFVD_INST_KIND, FFVD_INST_MODE>(rb, ngw, nw, minb);
}
