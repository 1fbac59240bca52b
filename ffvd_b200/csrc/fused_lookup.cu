// Split build: dispatch (kind, mode) to the per-object lookup functions defined by fused_inst.cu.
#define FFVD_SPLIT_BUILD 1
#include "fused_table.cuh"

#define FFVD_DECL(k, m) ffvd_fused_fn ffvd_fused_lookup_##k##_##m(int, int, int, int);
FFVD_DECL(0, 0) FFVD_DECL(0, 1) FFVD_DECL(0, 2) FFVD_DECL(0, 3) FFVD_DECL(0, 4)
FFVD_DECL(1, 0) FFVD_DECL(1, 1) FFVD_DECL(1, 2) FFVD_DECL(1, 3) FFVD_DECL(1, 4)
#undef FFVD_DECL

ffvd_fused_fn ffvd_fused_lookup(int kind, int mode, int rb, int ngw, int nw, int minb) {
#define FFVD_CASE(k, m) if (kind == k && mode == m) return ffvd_fused_lookup_##k##_##m(rb, ngw, nw, minb);
  FFVD_CASE(0, 0) FFVD_CASE(0, 1) FFVD_CASE(0, 2) FFVD_CASE(0, 3) FFVD_CASE(0, 4)
  FFVD_CASE(1, 0) FFVD_CASE(1, 1) FFVD_CASE(1, 2) FFVD_CASE(1, 3) FFVD_CASE(1, 4)
#undef FFVD_CASE
  return nullptr;
}
