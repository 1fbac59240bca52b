// Which (tile shape, warp count, CTAs/SM) instantiations of the fused kernel exist, and how they are looked up.
// Monolithic builds (dev / diagnostic) instantiate them inside capi.cu; the production build compiles one
// fused_inst.cu object per (kernel kind, mode) in parallel (FFVD_SPLIT_BUILD) and looks them up by value.
#pragma once
#include "ffvd_common.cuh"

typedef void (*ffvd_fused_fn)(const DevProblem*, int, long long, double*);

#ifdef FFVD_SPLIT_BUILD
ffvd_fused_fn ffvd_fused_lookup(int kind, int mode, int rb, int ngw, int nw, int minb);
#endif

#if !defined(FFVD_SPLIT_BUILD) || defined(FFVD_INST_KIND)
template <int KIND, int MODE>
static ffvd_fused_fn ffvd_fused_pick(int rb, int ngw, int nw, int minb) {
  using namespace ffvd;
  ffvd_fused_fn kern = nullptr;
#define FFVD_PICK(RB_, NGW_, NW_, MINB_) \
  if (ngw == NGW_ && rb == RB_ && nw == NW_ && minb == MINB_) kern = fused_kernel<KIND, RB_, NGW_, MODE, NW_, MINB_>;
#ifdef FFVD_DEV_MINIMAL
  // kernel-development build (seconds instead of minutes): SE uncollapsed only, a few tile shapes
  if constexpr (KIND == 0 && MODE == MODE_UNCOLLAPSED) {
    FFVD_PICK(8, 1, 16, 1) FFVD_PICK(8, 2, 8, 1) FFVD_PICK(4, 4, 8, 1) FFVD_PICK(8, 1, 8, 1)
    FFVD_PICK(4, 1, 16, 1)
    FFVD_PICK(4, 4, 4, 2) FFVD_PICK(8, 2, 4, 2)       // half-width CTAs, two per SM: Mp = 256 (BT = 32), Mp = 128 (BT = 64)
    FFVD_PICK(8, 1, 4, 2)                             // Mp = 64
  }
#else
  FFVD_PICK(8, 1, 16, 1) FFVD_PICK(8, 2, 8, 1) FFVD_PICK(4, 3, 8, 1) FFVD_PICK(4, 4, 8, 1)
  FFVD_PICK(2, 6, 8, 1) FFVD_PICK(2, 8, 8, 1) FFVD_PICK(1, 12, 8, 1) FFVD_PICK(1, 16, 8, 1)
  FFVD_PICK(4, 1, 16, 1)                     // half-height tiles at Mp = 128 for launches with fewer work items than SMs
  FFVD_PICK(4, 4, 4, 2) FFVD_PICK(8, 2, 4, 2)   // half-width CTAs (4 warps, 255 registers), two per SM: Mp = 256 (BT = 32), Mp = 128
  FFVD_PICK(8, 1, 4, 2)                         // Mp = 64 (M <= 64): one 16-column group per warp
#endif
#undef FFVD_PICK
  return kern;
}
#endif
