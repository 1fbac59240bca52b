// C ABI of libffvd_b200.so (see include/ffvd_b200.h).  Host-side orchestration only: tensor
// validation, staging of host tensors, workspace management and kernel launches.  There is no
// CPU compute path in this library.
#include "../../include/ffvd_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "elementwise.cuh"
#ifdef FFVD_SPLIT_BUILD
#include "ffvd_common.cuh"
#include "fused_modes.cuh"
#else
#include "fused.cuh"
#endif
#include "fused_table.cuh"
#include "prep_post.cuh"
#include "blocked_chol.cuh"
#include "comm.cuh"
#include "dense_small.cuh"

using namespace ffvd;

static thread_local std::string g_last_error;
static int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fail(FFVD_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));             \
  } while (0)

struct ffvd_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  std::vector<void*> pinned;      // descriptor copies the captured H2D nodes read at every replay
  int64_t kernels = 0;            // kernel nodes (launch accounting)
};

struct ffvd_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 0;
  int max_smem = 0;
  int64_t launches = 0;
  // workspace arena (grow-only, re-zeroed when the layout changes)
  char* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<long long> arena_key;
  DevProblem* d_probs = nullptr;
  OutPtrs* d_outs = nullptr;
  double* kscr = nullptr;      // points into the arena (set by ensure_arena)
  // pools of the last collapsed evaluation (read by ffvd_collapse_u_mean right after it)
  // Kzz factors of the last full preparation (FFVD_FLAG_REUSE_KZZ)
  bool kzz_valid = false;
  std::vector<long long> kzz_key;
  size_t last_off_cvec = 0, last_off_HxT = 0;
  int last_Mp = 0, last_nb = 0;
  char* det_buf = nullptr;         // FFVD_FLAG_DETERMINISTIC: private accumulator copies, x-bar planes, kzz_bwd partials (grow-only)
  size_t det_cap = 0;
  bool force_blocked = false;      // keep the Cholesky factor L itself (blocked path): conditional(return_Lm=True)
  bool p1_pending = false;         // FFVD_FLAG_COLLAPSED_P1_ONLY left its statistics in the arena; FFVD_FLAG_COLLAPSED_RESUME consumes them
  std::vector<long long> p1_key;
  size_t p1_off_S = 0, p1_bytes = 0; int p1_nb = 0, p1_Mp = 0;
  int cur_rb = 0;                  // tile height (row blocks of 8) chosen for the call in flight, 0 = fused_cfg's default
  int probs_cap = 0;
  int* h_status = nullptr;     // pinned
  size_t h_status_cap = 0;
  // ring of event pairs bracketing the fused-kernel launches (bench.py reads the average)
  static const int kRing = 256;
  cudaEvent_t ev0[kRing], ev1[kRing];
  long long ev_count = 0;
  // CUDA-graph capture (ffvd_graph_capture_begin / _end): work is captured on cap_stream, replayed into the caller's stream
  bool capturing = false;
  cudaStream_t user_stream = nullptr, cap_stream = nullptr;
  struct ffvd_graph* cur_graph = nullptr;
  // NCCL communicator (ffvd_comm_init) and the packed all-reduce buffer
  NcclComm comm = nullptr;
  int comm_rank = 0, comm_nranks = 1;
  double* pack_buf = nullptr;
  size_t pack_cap = 0;
};

extern "C" int ffvd_version(void) { return 100; }
extern "C" const char* ffvd_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* ffvd_status_string(int s) {
  switch (s) {
    case FFVD_OK: return "ok";
    case FFVD_E_BADARG: return "bad argument";
    case FFVD_E_DTYPE: return "tensor is not float64";
    case FFVD_E_SHAPE: return "shape mismatch or tensor not C-contiguous";
    case FFVD_E_DEVICE: return "tensor on a different device than the context";
    case FFVD_E_CUDA: return "CUDA runtime error";
    case FFVD_E_UNSUPPORTED: return "option not supported by this build";
    case FFVD_E_LIMIT: return "size beyond this build's limits";
    case FFVD_E_STALE: return "FFVD_FLAG_REUSE_KZZ: Z / kernel hyper-parameters changed since the cached factors were built";
    default: return s > 0 ? "matrix not positive definite (status = 1-based failing pivot)" : "unknown status";
  }
}

extern "C" int ffvd_ctx_create(int device, void* stream, ffvd_ctx** out) {
  if (!out) return fail(FFVD_E_BADARG, "out is null");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(FFVD_E_DEVICE, "no such CUDA device");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(FFVD_E_DEVICE, "libffvd_b200 requires an sm_100a (B200) device");
  ffvd_ctx* c = new ffvd_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->max_smem = (int)prop.sharedMemPerBlockOptin;
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return fail(FFVD_E_CUDA, cudaGetErrorString(e)); }
    c->own_stream = true;
  }
  for (int i = 0; i < ffvd_ctx::kRing; ++i) { cudaEventCreate(&c->ev0[i]); cudaEventCreate(&c->ev1[i]); }
  *out = c;
  return FFVD_OK;
}

extern "C" int ffvd_ctx_destroy(ffvd_ctx* c) {
  if (!c) return FFVD_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->arena) cudaFree(c->arena);
  if (c->d_probs) cudaFree(c->d_probs);
  if (c->d_outs) cudaFree(c->d_outs);
  if (c->h_status) cudaFreeHost(c->h_status);
  if (c->det_buf) cudaFree(c->det_buf);
  if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
  if (c->comm) { if (const NcclApi* a = nccl_api(nullptr)) a->CommDestroy(c->comm); c->comm = nullptr; }
  if (c->pack_buf) cudaFree(c->pack_buf);
  for (int i = 0; i < ffvd_ctx::kRing; ++i) { cudaEventDestroy(c->ev0[i]); cudaEventDestroy(c->ev1[i]); }
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return FFVD_OK;
}

extern "C" int ffvd_ctx_synchronize(ffvd_ctx* c) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return FFVD_OK;
}
extern "C" int64_t ffvd_ctx_launch_count(ffvd_ctx* c) { return c ? c->launches : 0; }
extern "C" int ffvd_ctx_fused_time(ffvd_ctx* c, int reset, double* total_ms, int64_t* count) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  const long long n = c->ev_count < ffvd_ctx::kRing ? c->ev_count : ffvd_ctx::kRing;
  double tot = 0.0;
  for (long long i = 0; i < n; ++i) {
    float ms = 0.0f;
    CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0[i], c->ev1[i]));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (count) *count = n;
  if (reset) c->ev_count = 0;
  return FFVD_OK;
}

extern "C" int ffvd_debug_phase_clocks(ffvd_ctx* c, int reset, uint64_t* out16) {
#ifdef FFVD_PHASE_TIMING
  if (!c || !out16) return fail(FFVD_E_BADARG, "null argument");
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaMemcpyFromSymbol(out16, g_phase_clocks, 16 * sizeof(uint64_t)));
  if (reset) {
    uint64_t z[16] = {0};
    CUDA_TRY(cudaMemcpyToSymbol(g_phase_clocks, z, sizeof z));
  }
  return FFVD_OK;
#else
  (void)c; (void)reset; (void)out16;
  return fail(FFVD_E_UNSUPPORTED, "library was not built with -DFFVD_PHASE_TIMING");
#endif
}

// ---------------------------------------------------------------------------------------------
// CUDA graphs: the reference drives 22 session.run calls per outer iteration (base_model.py:915-933, :944-950); on small
// problems (the bundled data: T <= 512, M = 100) an evaluation is ~10 launches of a few microseconds each, so launch gaps
// and the per-call host work dominate.  A captured sequence of ffvd_nll_grads_* / ffvd_sghmc_update / ffvd_adam_update
// calls replays as ONE launch.  Capture happens on a stream owned by the context (torch's legacy default stream cannot be
// captured); the graph is launched into the context's own stream.
extern "C" int ffvd_graph_capture_begin(ffvd_ctx* c) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  if (c->capturing) return fail(FFVD_E_BADARG, "a capture is already in progress");
  CUDA_TRY(cudaSetDevice(c->device));
  if (!c->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamSynchronize(c->stream));            // everything issued so far is done before the capture starts
  c->cur_graph = new ffvd_graph();
  cudaError_t e = cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeRelaxed);
  if (e != cudaSuccess) { delete c->cur_graph; c->cur_graph = nullptr; return fail(FFVD_E_CUDA, cudaGetErrorString(e)); }
  c->user_stream = c->stream;
  c->stream = c->cap_stream;
  c->capturing = true;
  return FFVD_OK;
}

static void graph_free(ffvd_graph* g) {
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  for (void* p : g->pinned) cudaFreeHost(p);
  delete g;
}

extern "C" int ffvd_graph_capture_end(ffvd_ctx* c, ffvd_graph** out) {
  if (!c || !out) return fail(FFVD_E_BADARG, "null argument");
  if (!c->capturing) return fail(FFVD_E_BADARG, "no capture in progress");
  ffvd_graph* g = c->cur_graph;
  cudaError_t e = cudaStreamEndCapture(c->cap_stream, &g->graph);
  c->stream = c->user_stream;
  c->capturing = false;
  c->cur_graph = nullptr;
  if (e != cudaSuccess || !g->graph) { graph_free(g); cudaGetLastError(); return fail(FFVD_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e)); }
  e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) { graph_free(g); return fail(FFVD_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
  size_t n = 0;
  cudaGraphGetNodes(g->graph, nullptr, &n);
  std::vector<cudaGraphNode_t> nodes(n);
  if (n) cudaGraphGetNodes(g->graph, nodes.data(), &n);
  for (auto& nd : nodes) {
    cudaGraphNodeType ty;
    if (cudaGraphNodeGetType(nd, &ty) == cudaSuccess && ty == cudaGraphNodeTypeKernel) g->kernels++;
  }
  *out = g;
  return FFVD_OK;
}

extern "C" int ffvd_graph_launch(ffvd_ctx* c, ffvd_graph* g) {
  if (!c || !g || !g->exec) return fail(FFVD_E_BADARG, "null argument");
  if (c->capturing) return fail(FFVD_E_BADARG, "cannot launch a graph while capturing");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaGraphLaunch(g->exec, c->stream));
  c->launches += g->kernels;
  return FFVD_OK;
}

extern "C" int64_t ffvd_graph_kernel_count(ffvd_graph* g) { return g ? g->kernels : 0; }

extern "C" int ffvd_graph_destroy(ffvd_ctx* c, ffvd_graph* g) {
  if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
  graph_free(g);
  return FFVD_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU: NCCL communicator per context
#define NCCL_TRY(api, expr)                                                                                        \
  do {                                                                                                             \
    int _r = (expr);                                                                                               \
    if (_r != 0) return fail(FFVD_E_CUDA, std::string(#expr) + ": NCCL error " + ((api)->GetErrorString ? (api)->GetErrorString(_r) : "?")); \
  } while (0)

extern "C" int ffvd_comm_unique_id(void* id128) {
  if (!id128) return fail(FFVD_E_BADARG, "id128 is null");
  std::string err;
  const NcclApi* a = nccl_api(&err);
  if (!a) return fail(FFVD_E_UNSUPPORTED, err);
  NcclUniqueId id;
  NCCL_TRY(a, a->GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return FFVD_OK;
}

extern "C" int ffvd_comm_init(ffvd_ctx* c, const void* id128, int rank, int nranks) {
  if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(FFVD_E_BADARG, "ffvd_comm_init: bad argument");
  std::string err;
  const NcclApi* a = nccl_api(&err);
  if (!a) return fail(FFVD_E_UNSUPPORTED, err);
  CUDA_TRY(cudaSetDevice(c->device));
  if (c->comm) { NCCL_TRY(a, a->CommDestroy(c->comm)); c->comm = nullptr; }
  NcclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NCCL_TRY(a, a->CommInitRank(&c->comm, nranks, id, rank));
  c->comm_rank = rank; c->comm_nranks = nranks;
  return FFVD_OK;
}

extern "C" int ffvd_comm_destroy(ffvd_ctx* c) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  if (c->comm) {
    const NcclApi* a = nccl_api(nullptr);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (a) NCCL_TRY(a, a->CommDestroy(c->comm));
    c->comm = nullptr;
  }
  c->comm_rank = 0; c->comm_nranks = 1;
  return FFVD_OK;
}

extern "C" int ffvd_comm_info(ffvd_ctx* c, int* rank, int* nranks, int* nccl_version) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  if (rank) *rank = c->comm_rank;
  if (nranks) *nranks = c->comm ? c->comm_nranks : 1;
  if (nccl_version) {
    *nccl_version = 0;
    const NcclApi* a = nccl_api(nullptr);
    if (a && a->GetVersion) a->GetVersion(nccl_version);
  }
  return FFVD_OK;
}

// ---------------------------------------------------------------------------------------------
// tensor import / staging
struct Tens {
  double* d = nullptr;     // device pointer the kernels use (always float64)
  void* host = nullptr;    // host pointer if the tensor lives on the host
  float* dev32 = nullptr;  // float32 tensors: the float32 device copy (the caller's own memory for device tensors)
  bool own32 = false;      // dev32 was allocated by the call (host float32 tensor)
  bool f32 = false;
  size_t numel = 0;
  int ndim = 0;
  int64_t shape[4] = {1, 1, 1, 1};
  bool staged = false, is_out = false, present = false;
};

// float32 mode (BASELINE north star: <= 1e-4 in float32): tensors may cross the ABI as float32; they are widened on the device
// into float64 workspace copies, ALL arithmetic stays float64 (the reference has no float32 path), and
// outputs are rounded back once.  What float32 buys is half the HBM / PCIe footprint of the caller's X, x-bar and SG-HMC
// state; the tensor-pipe work is unchanged.
__global__ void cvt_f32_to_f64_kernel(const float* __restrict__ src, double* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = (double)src[i];
}
__global__ void cvt_f64_to_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = (float)src[i];
}
static int cvt_grid(size_t n) {
  size_t b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

struct Call {
  ffvd_ctx* c;
  std::vector<Tens> staged;    // tensors that need work after the kernels: host copies back, float32 narrowing, frees
  bool touched_host = false;
  explicit Call(ffvd_ctx* ctx) : c(ctx) {}
  int import(DLManagedTensor* mt, bool is_out, Tens& t, const char* name, bool optional = false) {
    t = Tens();
    if (!mt) {
      if (optional) return FFVD_OK;
      return fail(FFVD_E_BADARG, std::string(name) + " is null");
    }
    const DLTensor& dl = mt->dl_tensor;
    const bool is64 = dl.dtype.code == 2 && dl.dtype.bits == 64 && dl.dtype.lanes == 1;
    const bool is32 = dl.dtype.code == 2 && dl.dtype.bits == 32 && dl.dtype.lanes == 1;
    if (!is64 && !is32) return fail(FFVD_E_DTYPE, std::string(name) + " must be float64 (or float32: widened on the device)");
    if (dl.ndim > 4) return fail(FFVD_E_SHAPE, std::string(name) + ": rank > 4");
    t.ndim = dl.ndim;
    t.numel = 1;
    for (int i = 0; i < dl.ndim; ++i) { t.shape[i] = dl.shape[i]; t.numel *= (size_t)dl.shape[i]; }
    if (dl.strides) {
      int64_t expect = 1;
      for (int i = dl.ndim - 1; i >= 0; --i) {
        if (dl.shape[i] != 1 && dl.strides[i] != expect)
          return fail(FFVD_E_SHAPE, std::string(name) + " must be C-contiguous");
        expect *= dl.shape[i];
      }
    }
    t.present = true;
    t.is_out = is_out;
    t.f32 = is32;
    char* base = (char*)dl.data + dl.byte_offset;
    const size_t esz = is32 ? 4 : 8;
    const bool on_dev = dl.device.device_type == kDLCUDA || dl.device.device_type == kDLCUDAManaged;
    if (on_dev) {
      if (dl.device.device_type == kDLCUDA && dl.device.device_id != c->device)
        return fail(FFVD_E_DEVICE, std::string(name) + " lives on another GPU");
      if (!is32) { t.d = (double*)base; return FFVD_OK; }
      t.dev32 = (float*)base;
    } else {
      if (dl.device.device_type != kDLCPU && dl.device.device_type != kDLCUDAHost)
        return fail(FFVD_E_DEVICE, std::string(name) + ": unsupported DLPack device type");
      // host tensor: explicit staging copy (a transfer, not a compute fallback)
      if (c->capturing) return fail(FFVD_E_DEVICE, std::string(name) + ": host tensors cannot be used while a CUDA graph is being captured");
      t.host = base;
      touched_host = true;
    }
    if (c->capturing) return fail(FFVD_E_DTYPE, std::string(name) + ": float32 tensors cannot be used while a CUDA graph is being captured");
    t.staged = true;
    if (t.numel == 0) return FFVD_OK;
    void* raw = nullptr;                                   // device copy in the tensor's own dtype
    if (t.host) {
      CUDA_TRY(cudaMallocAsync(&raw, t.numel * esz, c->stream));
      if (!is_out) CUDA_TRY(cudaMemcpyAsync(raw, t.host, t.numel * esz, cudaMemcpyHostToDevice, c->stream));
      if (is32) { t.dev32 = (float*)raw; t.own32 = true; }
    }
    if (is32) {
      CUDA_TRY(cudaMallocAsync((void**)&t.d, t.numel * sizeof(double), c->stream));
      if (!is_out) { cvt_f32_to_f64_kernel<<<cvt_grid(t.numel), 256, 0, c->stream>>>(t.dev32, t.d, t.numel); c->launches++; }
    } else {
      t.d = (double*)raw;
    }
    staged.push_back(t);
    return FFVD_OK;
  }
  int finish(bool force_sync = false) {
    for (auto& t : staged) {
      if (!t.is_out || !t.numel) continue;
      if (t.f32) { cvt_f64_to_f32_kernel<<<cvt_grid(t.numel), 256, 0, c->stream>>>(t.d, t.dev32, t.numel); c->launches++; }
      if (t.host) CUDA_TRY(cudaMemcpyAsync(t.host, t.f32 ? (void*)t.dev32 : (void*)t.d, t.numel * (t.f32 ? 4 : 8), cudaMemcpyDeviceToHost, c->stream));
    }
    for (auto& t : staged) {
      if (t.d && (t.host || t.f32)) CUDA_TRY(cudaFreeAsync(t.d, c->stream));
      if (t.own32 && t.dev32) CUDA_TRY(cudaFreeAsync(t.dev32, c->stream));
    }
    staged.clear();
    if (touched_host || force_sync) CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaGetLastError());
    return FFVD_OK;
  }
  ~Call() {
    for (auto& t : staged) {
      if (t.d && (t.host || t.f32)) cudaFreeAsync(t.d, c->stream);
      if (t.own32 && t.dev32) cudaFreeAsync(t.dev32, c->stream);
    }
  }
  // in-place tensors (SG-HMC / Adam state) are imported as inputs; mark them for the copy / narrowing back
  void mark_inout(const Tens& t) {
    for (auto& s : staged)
      if (s.d == t.d) s.is_out = true;
  }
};
#define TRY(expr)            \
  do {                       \
    int _s = (expr);         \
    if (_s != FFVD_OK) return _s; \
  } while (0)

static int grid1d(size_t n, int threads = 256, int cap = 148 * 16) {
  size_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > (size_t)cap) b = cap;
  return (int)b;
}

// ---------------------------------------------------------------------------------------------
// workspace
struct Layout {
  int nprob, nb, D, M, Mp, Din, Dy, nk;   // nk = kernels per problem (D, or 1 when shared)
  long long sumS;
  bool collapsed;
  size_t off_ZT, off_ZTs, off_hyp, off_hq, off_UT, off_Linv, off_LinvT, off_Sacc, off_Wk, off_Nmat, off_Hx, off_HxT, off_ubar, off_cvec, off_wvec, off_rs,
      off_small, off_terms, off_status, off_utmp, off_kscr, off_Lfac, off_Dinv, off_status2, off_guard, off_collb, off_gZd, total;
  int nfac;                        // matrices in the blocked-factorisation pools: nprob * max(nk, nb)
  size_t zero_begin, zero_end;     // region re-zeroed before every evaluation
  size_t small_per;                // doubles of small accumulators per problem
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static const int kMaxCtasPerSm = 2;      // co-resident fused CTAs per SM the K-tile scratch is sized for (FFVD_MINB <= 2)
static Layout make_layout(const ffvd_ctx* c, int nprob, int nb, int nk, int D, int M, int Mp, int Din, int Dy, long long sumS,
                          bool collapsed, bool need_acc) {
  Layout L;
  L.nprob = nprob; L.nb = nb; L.nk = nk; L.D = D; L.M = M; L.Mp = Mp; L.Din = Din; L.Dy = Dy; L.sumS = sumS; L.collapsed = collapsed;
  size_t o = 0;
  const size_t mm = (size_t)Mp * Mp * sizeof(double);
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  L.off_ZT = take((size_t)nprob * 64 * Mp * 8);          // Z~^T and its fragment-ordered copy
  L.off_ZTs = take((size_t)nprob * nk * FFVD_ZTS_ROWS * Mp * 8);
  L.off_hyp = take((size_t)nprob * nk * 72 * 8);
  L.off_hq = take((size_t)nprob * D * 4 * 8);
  L.off_UT = take((size_t)nprob * D * Mp * 8);
  L.off_kscr = take((size_t)c->num_sms * kMaxCtasPerSm * 64 * Mp * 8);   // per-CTA K-tile scratch of the fused kernel
  L.off_guard = take((size_t)nprob * 4 * 8);             // REUSE_KZZ content guard (outside the per-call zero region)
  L.off_Linv = take((size_t)nprob * nk * mm);
  L.off_LinvT = take((size_t)nprob * nk * mm);
  L.off_utmp = take((size_t)nprob * M * D * 8);
  L.nfac = nprob * (nb > nk ? nb : nk);
  L.off_Lfac = take((size_t)nprob * nk * mm);
  L.off_Dinv = take((size_t)L.nfac * (Mp / 64) * 64 * 64 * 8);
  L.off_Wk = take(need_acc ? (size_t)nprob * nb * mm : 0);
  L.off_Nmat = take(collapsed ? (size_t)nprob * nb * mm : 0);
  L.off_Hx = take(collapsed ? (size_t)nprob * nb * mm : 0);
  L.off_HxT = take(collapsed ? (size_t)nprob * nb * mm : 0);
  L.off_cvec = take(collapsed ? (size_t)nprob * nb * Mp * 8 : 0);
  L.off_wvec = take((size_t)nprob * nb * Mp * 8);        // uncollapsed: L^{-T} u ; collapsed: w'
  L.off_rs = take(need_acc ? (size_t)nprob * nb * Mp * 8 : 0);
  L.zero_begin = o;
  L.off_gZd = take(need_acc ? (size_t)nprob * D * Mp * 32 * 8 : 0);     // raw Z-bar products per output dim; with S the "per-CTA copies" region
  L.off_Sacc = take(need_acc ? (size_t)nprob * nb * mm : 0);           //   [off_gZd, off_ubar) of the deterministic mode
  L.off_ubar = take(need_acc ? (size_t)nprob * nb * Mp * 8 : 0);
  L.small_per = (size_t)M * Din + (size_t)D * Din + D + D + (size_t)D * Dy + Dy + Dy;
  L.off_small = take((size_t)nprob * L.small_per * 8);
  L.off_terms = take((size_t)sumS * FFVD_NTERMS_RAW * 8);
  L.off_status = take((size_t)nprob * (nb > nk ? nb : nk) * sizeof(int));
  L.off_collb = take((size_t)nprob * nb * 4 * 8);       // collapsed: per (s,d) scalar terms (plain stores; finalize sums them in order)
  L.off_status2 = take((size_t)L.nfac * sizeof(int));
  L.zero_end = o;
  L.total = o;
  return L;
}

static int ensure_arena(ffvd_ctx* c, const Layout& L) {
  std::vector<long long> key = {L.nprob, L.nb, L.nk, L.D, L.M, L.Mp, L.Din, L.Dy, L.sumS, L.collapsed ? 1 : 0, (long long)L.total};
  if (c->capturing && (L.total > c->arena_bytes || key != c->arena_key || c->probs_cap < L.nprob))
    return fail(FFVD_E_BADARG, "CUDA-graph capture: make the same call once outside the capture first (the workspace must already be laid out)");
  if (L.total > c->arena_bytes) {
    if (c->arena) { CUDA_TRY(cudaStreamSynchronize(c->stream)); CUDA_TRY(cudaFree(c->arena)); c->arena = nullptr; c->arena_bytes = 0; }
    CUDA_TRY(cudaMalloc((void**)&c->arena, L.total));
    c->arena_bytes = L.total;
    c->arena_key.clear();
  }
  if (key != c->arena_key) {
    CUDA_TRY(cudaMemsetAsync(c->arena, 0, L.total, c->stream));   // padding of Linv etc. must be zero
    c->arena_key = key;
    c->kzz_valid = false;
  }
  c->kscr = (double*)(c->arena + L.off_kscr);
  if (c->probs_cap < L.nprob) {
    if (c->d_probs) { CUDA_TRY(cudaStreamSynchronize(c->stream)); CUDA_TRY(cudaFree(c->d_probs)); CUDA_TRY(cudaFree(c->d_outs)); }
    CUDA_TRY(cudaMalloc((void**)&c->d_probs, sizeof(DevProblem) * L.nprob));
    CUDA_TRY(cudaMalloc((void**)&c->d_outs, sizeof(OutPtrs) * L.nprob));
    c->probs_cap = L.nprob;
  }
  const size_t ns = (size_t)L.nprob * (L.nb > L.nk ? L.nb : L.nk) + (size_t)L.nfac + 2 + (size_t)L.nprob * 8;   // + guard words
  if (c->h_status_cap < ns) {
    if (c->h_status) CUDA_TRY(cudaFreeHost(c->h_status));
    CUDA_TRY(cudaMallocHost((void**)&c->h_status, ns * sizeof(int)));
    c->h_status_cap = ns;
  }
  return FFVD_OK;
}

static void bind_problem(ffvd_ctx* c, const Layout& L, int p, long long s_begin, DevProblem& P) {
  char* a = c->arena;
  const size_t mm = (size_t)L.Mp * L.Mp;
  P.ZT = (double*)(a + L.off_ZT) + (size_t)p * 64 * L.Mp;
  P.Zf = P.ZT + (size_t)32 * L.Mp;
  P.ZTs = (double*)(a + L.off_ZTs) + (size_t)p * L.nk * FFVD_ZTS_ROWS * L.Mp;
  P.hyp = (double*)(a + L.off_hyp) + (size_t)p * L.nk * 72;
  P.hq = (double*)(a + L.off_hq) + (size_t)p * L.D * 4;
  P.UT = (double*)(a + L.off_UT) + (size_t)p * L.D * L.Mp;
  P.Linv = (double*)(a + L.off_Linv) + (size_t)p * L.nk * mm;
  P.LinvT = (double*)(a + L.off_LinvT) + (size_t)p * L.nk * mm;
  P.Sacc = (double*)(a + L.off_Sacc) + (size_t)p * L.nb * mm;
  P.Wk = (double*)(a + L.off_Wk) + (size_t)p * L.nb * mm;
  P.Nmat = (double*)(a + L.off_Nmat) + (size_t)p * L.nb * mm;
  P.Hx = (double*)(a + L.off_Hx) + (size_t)p * L.nb * mm;
  P.HxT = (double*)(a + L.off_HxT) + (size_t)p * L.nb * mm;
  P.ubar = (double*)(a + L.off_ubar) + (size_t)p * L.nb * L.Mp;
  P.cvec = (double*)(a + L.off_cvec) + (size_t)p * L.nb * L.Mp;
  P.wvec = (double*)(a + L.off_wvec) + (size_t)p * L.nb * L.Mp;
  P.rs = (double*)(a + L.off_rs) + (size_t)p * L.nb * L.Mp;
  double* sm = (double*)(a + L.off_small) + (size_t)p * L.small_per;
  P.gZ = sm; sm += (size_t)L.M * L.Din;
  P.gl = sm; sm += (size_t)L.D * L.Din;
  P.gv = sm; sm += L.D;
  P.gQ = sm; sm += L.D;
  P.gC = sm; sm += (size_t)L.D * L.Dy;
  P.gd = sm; sm += L.Dy;
  P.gR = sm;
  P.terms_raw = (double*)(a + L.off_terms) + (size_t)s_begin * FFVD_NTERMS_RAW;
  P.status = (int*)(a + L.off_status) + (size_t)p * (L.nb > L.nk ? L.nb : L.nk);
  P.guard = (unsigned long long*)(a + L.off_guard) + (size_t)p * 4;
  P.collb = (double*)(a + L.off_collb) + (size_t)p * L.nb * 4;
  P.gZd = (double*)(a + L.off_gZd) + (size_t)p * L.D * L.Mp * 32;
}

// ---------------------------------------------------------------------------------------------
// kernel dispatch helpers
// Tile configuration per padded M: RB row blocks (BT = 8*RB time steps per tile) and NW warps per CTA.
// Tile configuration of the fused kernel for a padded inducing-point count.  Overridable for experiments with
// FFVD_RB / FFVD_NW / FFVD_MINB (only the combinations instantiated below exist).
struct FusedCfg { int rb, nw, minb; };
#ifndef FFVD_DUAL_DEFAULT
#define FFVD_DUAL_DEFAULT 0
#endif
// supported padded sizes: 64 and 128 * {1,2,3,4,6,8,12,16}
static int pad_M(int M) {
  static const int sizes[] = {64, 128, 256, 384, 512, 768, 1024, 1536, 2048};
  for (int s : sizes)
    if (M <= s) return s;
  return -1;
}
static FusedCfg fused_cfg(int Mp) {
  FusedCfg cfg;
  const int ngw = Mp / 128;
  cfg.rb = (ngw <= 2) ? 8 : (ngw <= 4 ? 4 : (ngw <= 8 ? 2 : 1));
  cfg.nw = (ngw == 1) ? 16 : 8;
  cfg.minb = 1;
  if (Mp == 64) {            // M <= 64: four column warps (one 16-column group each), two CTAs per SM; padding to 128 would
    cfg.rb = 8; cfg.nw = 4; cfg.minb = 2;      // quadruple the tensor work of both contractions and of the SYRK
    if (const char* e = getenv("FFVD_RB")) cfg.rb = atoi(e);
    return cfg;
  }
  // Half-width CTAs (4 warps x 255 registers, BT x Mp tile of half the rows), TWO per SM: the non-tensor phases of one CTA
  // (K tile, staging, statistics, flushes) run under the contraction phases of the other.  Only where the B-operand
  // stream from L2 can afford half the reuse (Mp <= 256).  FFVD_DUAL=0/1 overrides.
  bool dual = FFVD_DUAL_DEFAULT != 0;
  if (const char* e = getenv("FFVD_DUAL")) dual = atoi(e) != 0;
  if (dual && ngw == 2) { cfg.rb = 4; cfg.nw = 4; cfg.minb = 2; }
  if (dual && ngw == 1) { cfg.rb = 8; cfg.nw = 4; cfg.minb = 2; }
  if (const char* e = getenv("FFVD_RB")) cfg.rb = atoi(e);
  if (const char* e = getenv("FFVD_NW")) cfg.nw = atoi(e);
  if (const char* e = getenv("FFVD_MINB")) cfg.minb = atoi(e);
  return cfg;
}

template <int KIND, int MODE>
static int launch_fused(ffvd_ctx* c, int Mp, int Din, const DevProblem* d_probs, int nprob, long long total_items) {
  FusedCfg cfg = fused_cfg(Mp);
  if (c->cur_rb) cfg.rb = c->cur_rb;
  const int ngw = Mp / (16 * (cfg.nw < 8 ? cfg.nw : 8));       // 16-column groups per column warp
#ifdef FFVD_SPLIT_BUILD
  ffvd_fused_fn kern = ffvd_fused_lookup(KIND, MODE, cfg.rb, ngw, cfg.nw, cfg.minb);    // instantiated in fused_inst.cu objects
#else
  ffvd_fused_fn kern = ffvd_fused_pick<KIND, MODE>(cfg.rb, ngw, cfg.nw, cfg.minb);
#endif
  if (!kern) return fail(FFVD_E_BADARG, "no fused kernel instantiated for this (Mp, FFVD_RB, FFVD_NW, FFVD_MINB)");
  const int RB = cfg.rb;
  const size_t smem = fused_smem_bytes(RB, Mp, cfg.nw, (Din + 1 + 7) / 8);
  if ((int)smem > c->max_smem) return fail(FFVD_E_LIMIT, "fused kernel shared memory exceeds the device limit");
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * cfg.nw, smem));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > cfg.minb) per_sm = cfg.minb;
  if (per_sm > kMaxCtasPerSm) per_sm = kMaxCtasPerSm;           // the K-tile scratch holds num_sms * kMaxCtasPerSm tiles
  const long long cap = (long long)c->num_sms * per_sm;      // persistent grid: every CTA resident
  long long grid = total_items < cap ? total_items : cap;
  if (grid < 1) return FFVD_OK;
  const int slot = (int)(c->ev_count % ffvd_ctx::kRing);
  if (!c->capturing) cudaEventRecord(c->ev0[slot], c->stream);
  kern<<<(int)grid, 32 * cfg.nw, smem, c->stream>>>(d_probs, nprob, total_items, c->kscr);
  if (!c->capturing) { cudaEventRecord(c->ev1[slot], c->stream); c->ev_count++; }
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return FFVD_OK;
}
static int rb_of(int Mp) { return fused_cfg(Mp).rb; }
// Tile height for a call with `rows` x-rows per (sample, output dim) pair and `pairs` such pairs: the default, except that
// at Mp = 128 half-height tiles are used while they still fit one wave of the persistent grid -- a launch with fewer
// work items than SMs (the single-chain case: T = 512, D = 4 -> 64 items) is one item's latency long, and half the rows
// is nearly half the latency (C1: fused kernel 38 -> 24 us).
static int rb_for_call(const ffvd_ctx* c, int Mp, long long pairs, long long rows) {
  const int rb = rb_of(Mp);
  if (Mp != 128 || rb != 8 || getenv("FFVD_RB") || fused_cfg(Mp).nw == 4) return rb;      // (Mp = 64 keeps full-height tiles)
  const long long items4 = pairs * ((rows + 31) / 32);
  return items4 <= c->num_sms ? 4 : rb;
}
// output dims per block of the work-item order: keep the operand matrices of one block (16 Mp^2 bytes per dim) within
// ~16 MB of L2.  FFVD_DBLK overrides (experiments).
static int dblk_of(int Mp, int D) {
  int b = (int)((size_t)16 * 1024 * 1024 / ((size_t)16 * Mp * Mp));
  if (const char* e = getenv("FFVD_DBLK")) b = atoi(e);
  if (b < 1) b = 1;
  if (b > D) b = D;
  return b;
}

// Work-item form (see fused_kernel): one item per (sample, tile, block of dblk dims) -- the CTA stages the x tile once and
// loops over the dims -- when that still leaves >= 32 items per CTA of the persistent grid (tail <= 3%); one item per
// (sample, tile, d) otherwise.  FFVD_DL=0/1 overrides (A/B timing).
static void set_items(const ffvd_ctx* c, DevProblem& P, long long pairs_st /* S * ntiles */, bool force_loop = false) {
  const int nblk = (P.D + P.dblk - 1) / P.dblk;
  bool loop = P.dblk > 1 && pairs_st * nblk >= (long long)32 * c->num_sms;
  if (const char* e = getenv("FFVD_DL")) loop = atoi(e) != 0 && P.dblk > 1;
  if (force_loop) loop = P.dblk > 1;        // deterministic mode: ONE item per (sample, tile) so that x-bar rows have one writer CTA
  P.dl = loop ? P.dblk : 1;
  P.nitems = loop ? pairs_st * nblk : pairs_st * P.D;
}

static int blocked_factor_invert(ffvd_ctx* c, double* A, double* Dinv, double* X, double* XT, int* status, int nbatch,
                                 int M, int Mp);

static bool use_blocked(const ffvd_ctx* c, int M, int Mp) {
  if (c->force_blocked) return true;
  if (const char* e = getenv("FFVD_BLOCKED_CHOL")) return atoi(e) != 0;
  if (chol_fast_fits(M, Mp, (size_t)c->max_smem)) return false;     // register-resident single-CTA path (M <= 119)
  // Above it the multi-kernel blocked path (whose 64 x 64 diagonal blocks go through the same register-resident
  // routines) wins at every M: 0.21 ms at Mp = 128, 0.48 ms at Mp = 256 for four matrices against 0.41 ... 0.83 ms of the
  // generic single-CTA shared-memory routines (tools/prep_paths.py), which stay reachable with FFVD_BLOCKED_CHOL=0.
  return true;
}

template <int KIND>
static int launch_prep(ffvd_ctx* c, const Layout& L, double jitter, bool reuse = false, long long ident = 0, bool want_ltu = false,
                       bool* ltu_done = nullptr) {
  long long jbits;
  memcpy(&jbits, &jitter, sizeof jbits);
  std::vector<long long> key = c->arena_key;
  key.push_back(KIND); key.push_back(jbits); key.push_back(ident);
  const bool reused = reuse && c->kzz_valid && key == c->kzz_key;     // factors of the previous call are still in the arena
  if (ltu_done) *ltu_done = false;
  // hyp / hq / U^T every call (U changes under REUSE_KZZ); the scaled Z~^T (SE) only when the factors are rebuilt
  // content guard: record the hash of Z / logv / logl when the factors are (re)built, verify it when they are reused
  hyper_kernel<<<dim3(L.nk > L.D ? L.nk : L.D, L.nprob), 128, 0, c->stream>>>(c->d_probs, KIND, L.nk, (KIND == 0 && !reused) ? 1 : 0,
                                                                              reused ? 2 : 1);
  c->launches++;
  if (reused) return FFVD_OK;
  c->kzz_valid = true;          // check_status clears it again if the factorisation fails; ASYNC callers are covered by the
  c->kzz_key = key;             // device-side guard (finalize_kernel drops the record of a failed factorisation)
  if (use_blocked(c, L.M, L.Mp)) {
    double* Lfac = (double*)(c->arena + L.off_Lfac);
    const size_t nel = (size_t)L.Mp * L.Mp > (size_t)32 * L.Mp ? (size_t)L.Mp * L.Mp : (size_t)32 * L.Mp;
    kzz_fill_kernel<KIND><<<dim3((unsigned)((nel + 255) / 256), L.nk, L.nprob), 256, 0, c->stream>>>(c->d_probs, Lfac, jitter);
    c->launches++;
    return blocked_factor_invert(c, Lfac, (double*)(c->arena + L.off_Dinv), (double*)(c->arena + L.off_Linv),
                                 (double*)(c->arena + L.off_LinvT), (int*)(c->arena + L.off_status2), L.nprob * L.nk, L.M, L.Mp);
  }
  size_t smem = (size_t)2 * L.Mp * 8 + (size_t)L.M * (L.M + 1) * 8;
  int mode = 1;
  if (chol_fast_fits(L.M, L.Mp, (size_t)c->max_smem) && !getenv("FFVD_NO_FAST_CHOL")) {
    smem = chol_fast_smem_doubles(L.M, L.Mp) * 8;      // L and L^{-1} both in shared memory (M <= ~116)
    mode = 2;
  }
  CUDA_TRY(cudaFuncSetAttribute(kzz_prep_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int do_ltu = (mode == 2 && want_ltu) ? 1 : 0;     // w = L^{-T} u from the shared-memory inverse (else ltu_kernel)
  kzz_prep_kernel<KIND><<<dim3(L.nk, L.nprob), 512, smem, c->stream>>>(c->d_probs, jitter, mode, do_ltu);
  if (ltu_done) *ltu_done = do_ltu != 0;
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return FFVD_OK;
}

// tri: BGEMM_A_UPPER / BGEMM_B_LOWER when an operand is L^{-T} / L^{-1} (exact zeros in the other triangle): see bgemm_nn_kernel
enum { BGEMM_A_UPPER = 1, BGEMM_B_LOWER = 2 };
static int launch_bgemm(ffvd_ctx* c, double* C, const double* A, const double* B, int n, double alpha, int batch,
                        BatchMap mC, BatchMap mA, BatchMap mB, int tri = 0) {
  // half-height tiles while the launch would not fill the GPU with 64 x 64 ones
  if ((long long)(n / 64) * (n / 64) * batch < c->num_sms)
    bgemm_nn_kernel<32><<<dim3(n / 64, n / 32, batch), 256, 0, c->stream>>>(C, A, B, n, alpha, mC, mA, mB, tri);
  else
    bgemm_nn_kernel<64><<<dim3(n / 64, n / 64, batch), 256, 0, c->stream>>>(C, A, B, n, alpha, mC, mA, mB, tri);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return FFVD_OK;
}

static int check_status(ffvd_ctx* c, const Layout& L) {
  const size_t n1 = (size_t)L.nprob * (L.nb > L.nk ? L.nb : L.nk), n2 = (size_t)L.nfac;
  CUDA_TRY(cudaMemcpyAsync(c->h_status, c->arena + L.off_status, n1 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->h_status + n1, c->arena + L.off_status2, n2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  unsigned long long* hg = (unsigned long long*)(c->h_status + n1 + n2 + ((n1 + n2) & 1));
  CUDA_TRY(cudaMemcpyAsync(hg, c->arena + L.off_guard, (size_t)L.nprob * 4 * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (int p = 0; p < L.nprob; ++p)
    if (hg[(size_t)p * 4 + 3] != 0) {
      c->kzz_valid = false;
      return fail(FFVD_E_STALE, "FFVD_FLAG_REUSE_KZZ: the contents of Z / logv / logl differ from those the cached Cholesky factors "
                                "were built from (results of this call are NaN); call again without the flag");
    }
  for (size_t i = 0; i < n1 + n2; ++i)
    if (c->h_status[i] != 0) {
      c->kzz_valid = false;      // never reuse the factors of a failed factorisation
      char buf[160];
      snprintf(buf, sizeof buf, "Cholesky failed: matrix %zu not positive definite at pivot %d", i < n1 ? i : i - n1, c->h_status[i]);
      g_last_error = buf;
      return c->h_status[i];
    }
  return FFVD_OK;
}

// Blocked Cholesky + inverse of nbatch matrices: A (in: SPD, lower filled; out: L), X <- L^{-1}, XT <- L^{-T}.
static int blocked_factor_invert(ffvd_ctx* c, double* A, double* Dinv, double* X, double* XT, int* status, int nbatch,
                                 int M, int Mp) {
  const int nblk = Mp / 64;
  const size_t sm_potrf = (size_t)(2 * 64 * 65 + 64) * 8, sm_trtri = (size_t)64 * 68 * 8;
  CUDA_TRY(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_potrf));
  CUDA_TRY(cudaFuncSetAttribute(trtri_column_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_trtri));
  for (int kb = 0; kb < nblk; ++kb) {
    potrf_diag_kernel<<<nbatch, 256, sm_potrf, c->stream>>>(A, Dinv, status, kb, M, Mp); c->launches++;
    const int rem = nblk - kb - 1;
    if (rem > 0) {
      trsm_panel_kernel<<<dim3(rem, nbatch), 256, 0, c->stream>>>(A, Dinv, kb, Mp); c->launches++;
      syrk_trailing_kernel<<<dim3(rem * (rem + 1) / 2, nbatch), 256, 0, c->stream>>>(A, kb, Mp); c->launches++;
    }
  }
  trtri_column_kernel<<<dim3(nblk, nbatch), 256, sm_trtri, c->stream>>>(A, Dinv, X, Mp); c->launches++;
  transpose_pad_kernel<<<dim3(Mp / 32, Mp / 32, nbatch), dim3(32, 8), 0, c->stream>>>(X, XT, M, Mp); c->launches++;
  CUDA_TRY(cudaGetLastError());
  return FFVD_OK;
}

// ---------------------------------------------------------------------------------------------
struct ProbT {
  Tens X, Z, U, logv, logl, logQ, C, d, logR, Y, ctrl;
  Tens nll, terms, gX, gZ, gU, glogv, glogl, glogQ, gC, gd, glogR;
  int S, T, D, Din, nc, Dy, M;
  double* gx_scratch = nullptr;
};

static int import_problem(Call& call, int kind, const ffvd_problem* p, const ffvd_outputs* o, ProbT& t) {
  TRY(call.import(p->X, false, t.X, "X"));
  TRY(call.import(p->Z, false, t.Z, "Z"));
  TRY(call.import(p->U, false, t.U, "U"));
  TRY(call.import(p->logv, false, t.logv, "logv"));
  TRY(call.import(p->logl, false, t.logl, "logl", kind != FFVD_KERNEL_SE));
  TRY(call.import(p->logQ, false, t.logQ, "logQ"));
  TRY(call.import(p->C, false, t.C, "C"));
  TRY(call.import(p->d, false, t.d, "d"));
  TRY(call.import(p->logR, false, t.logR, "logR"));
  TRY(call.import(p->Y, false, t.Y, "Y"));
  TRY(call.import(p->ctrl, false, t.ctrl, "ctrl", true));
  if (t.X.ndim == 2) { t.S = 1; t.T = (int)t.X.shape[0] - 1; t.D = (int)t.X.shape[1]; }
  else if (t.X.ndim == 3) { t.S = (int)t.X.shape[0]; t.T = (int)t.X.shape[1] - 1; t.D = (int)t.X.shape[2]; }
  else return fail(FFVD_E_SHAPE, "X must be (T+1,D) or (S,T+1,D)");
  if (t.T < 1 || t.D < 1 || t.S < 1) return fail(FFVD_E_SHAPE, "X is empty");
  if (t.Z.ndim != 2) return fail(FFVD_E_SHAPE, "Z must be (M,Din)");
  t.M = (int)t.Z.shape[0]; t.Din = (int)t.Z.shape[1];
  t.nc = t.Din - t.D;
  if (t.nc < 0) return fail(FFVD_E_SHAPE, "Z has fewer columns than X");
  if (t.Din > FFVD_MAX_DIN) return fail(FFVD_E_LIMIT, "Din > 31");
  if (t.nc > 0) {
    if (!t.ctrl.present || t.ctrl.ndim != 2 || t.ctrl.shape[0] < t.T || t.ctrl.shape[1] != t.nc)
      return fail(FFVD_E_SHAPE, "ctrl must be (>=T, Din-D)");
    if (t.ctrl.shape[0] != t.T) return fail(FFVD_E_SHAPE, "ctrl must have exactly T rows");
  }
  if (t.U.ndim != 2 || t.U.shape[0] != t.M || t.U.shape[1] != t.D) return fail(FFVD_E_SHAPE, "U must be (M,D)");
  if (t.logv.numel != (size_t)t.D) return fail(FFVD_E_SHAPE, "logv must be (D)");
  if (kind == FFVD_KERNEL_SE && t.logl.numel != (size_t)t.D * t.Din) return fail(FFVD_E_SHAPE, "logl must be (D,Din)");
  if (t.logQ.numel != (size_t)t.D) return fail(FFVD_E_SHAPE, "logQ must be (D)");
  if (t.Y.ndim != 2 || t.Y.shape[0] != t.T) return fail(FFVD_E_SHAPE, "Y must be (T,Dy)");
  t.Dy = (int)t.Y.shape[1];
  if (t.C.numel != (size_t)t.D * t.Dy) return fail(FFVD_E_SHAPE, "C must be (D,Dy)");
  if (t.d.numel != (size_t)t.Dy) return fail(FFVD_E_SHAPE, "d must be (Dy)");
  if (t.logR.numel != (size_t)t.Dy * t.Dy) return fail(FFVD_E_SHAPE, "logR must be (Dy,Dy)");
  if (o) {
    TRY(call.import(o->nll, true, t.nll, "nll", true));
    TRY(call.import(o->terms, true, t.terms, "terms", true));
    TRY(call.import(o->g_X, true, t.gX, "g_X", true));
    TRY(call.import(o->g_Z, true, t.gZ, "g_Z", true));
    TRY(call.import(o->g_U, true, t.gU, "g_U", true));
    TRY(call.import(o->g_logv, true, t.glogv, "g_logv", true));
    TRY(call.import(o->g_logl, true, t.glogl, "g_logl", true));
    TRY(call.import(o->g_logQ, true, t.glogQ, "g_logQ", true));
    TRY(call.import(o->g_C, true, t.gC, "g_C", true));
    TRY(call.import(o->g_d, true, t.gd, "g_d", true));
    TRY(call.import(o->g_logR, true, t.glogR, "g_logR", true));
    auto chk = [&](const Tens& g, size_t n, const char* nm) -> int {
      if (g.present && g.numel != n) return fail(FFVD_E_SHAPE, std::string(nm) + " has the wrong number of elements");
      return FFVD_OK;
    };
    TRY(chk(t.nll, t.S, "nll")); TRY(chk(t.terms, (size_t)t.S * 6, "terms")); TRY(chk(t.gX, t.X.numel, "g_X"));
    TRY(chk(t.gZ, t.Z.numel, "g_Z")); TRY(chk(t.gU, t.U.numel, "g_U")); TRY(chk(t.glogv, t.D, "g_logv"));
    TRY(chk(t.glogl, (size_t)t.D * t.Din, "g_logl")); TRY(chk(t.glogQ, t.D, "g_logQ")); TRY(chk(t.gC, t.C.numel, "g_C"));
    TRY(chk(t.gd, t.Dy, "g_d")); TRY(chk(t.glogR, t.logR.numel, "g_logR"));
  }
  return FFVD_OK;
}

template <int KIND>
static int run_nll(ffvd_ctx* c, int collapsed, int nprob, const ffvd_problem* probs, int flags, double jitter,
                   const ffvd_outputs* outs) {
  if (!c || !probs || !outs || nprob < 1) return fail(FFVD_E_BADARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  Call call(c);
  std::vector<ProbT> pt(nprob);
  long long sumS = 0;
  for (int p = 0; p < nprob; ++p) {
    TRY(import_problem(call, KIND, &probs[p], &outs[p], pt[p]));
    if (p > 0 && (pt[p].M != pt[0].M || pt[p].D != pt[0].D || pt[p].Din != pt[0].Din || pt[p].Dy != pt[0].Dy))
      return fail(FFVD_E_SHAPE, "batched problems must share M, D, Din, Dy");
    if (collapsed && pt[p].S != pt[0].S) return fail(FFVD_E_UNSUPPORTED, "collapsed batched problems must share S");
    sumS += pt[p].S;
  }
  const int M = pt[0].M, D = pt[0].D, Din = pt[0].Din, Dy = pt[0].Dy;
  const int Mp = pad_M(M);
  if (Mp < 0) return fail(FFVD_E_LIMIT, "M > 2048 is not supported by this build");
  const bool no_grads = (flags & FFVD_FLAG_NO_GRADS) != 0;
  const bool det = (flags & FFVD_FLAG_DETERMINISTIC) != 0;
  if (det && c->capturing) return fail(FFVD_E_UNSUPPORTED, "FFVD_FLAG_DETERMINISTIC is not captured in CUDA graphs");
  if (det && (flags & (FFVD_FLAG_COLLAPSED_P1_ONLY | FFVD_FLAG_COLLAPSED_RESUME))) return fail(FFVD_E_UNSUPPORTED, "FFVD_FLAG_DETERMINISTIC with a split collapsed evaluation is not built");
  const int nb = collapsed ? pt[0].S * D : D;
  Layout L = make_layout(c, nprob, nb, D, D, M, Mp, Din, Dy, sumS, collapsed != 0, true);
  TRY(ensure_arena(c, L));
  long long pairs = 0, maxT = 0;
  for (auto& t : pt) { pairs += (long long)D * t.S; maxT = t.T > maxT ? t.T : maxT; }
  const int RB = rb_for_call(c, Mp, pairs, maxT), BT = 8 * RB;
  c->cur_rb = RB;

  std::vector<DevProblem> hp(nprob);
  std::vector<OutPtrs> ho(nprob);
  long long item = 0, s_begin = 0;
  for (int p = 0; p < nprob; ++p) {
    ProbT& t = pt[p];
    DevProblem& P = hp[p];
    memset(&P, 0, sizeof P);
    P.X = t.X.d; P.Z = t.Z.d; P.U = t.U.d; P.logv = t.logv.d; P.logl = t.logl.d; P.logQ = t.logQ.d; P.C = t.C.d;
    P.dvec = t.d.d; P.logR = t.logR.d; P.Y = t.Y.d; P.ctrl = t.ctrl.d;
    P.S = t.S; P.T = t.T; P.D = D; P.Din = Din; P.nc = t.nc; P.Dy = Dy; P.M = M; P.Mp = Mp; P.Dx = D; P.xrows = t.T + 1; P.hs = 1;
    P.ntiles = (t.T + BT - 1) / BT;
    P.dblk = det ? D : dblk_of(Mp, D);
    P.item_begin = item;
    set_items(c, P, (long long)t.S * P.ntiles, det);
    item += P.nitems;
    bind_problem(c, L, p, s_begin, P);
    s_begin += t.S;
    if (t.gX.present) P.gX = t.gX.d;
    else if (!no_grads || collapsed) {   // collapsed pass 1 always accumulates the emission gradient
      if (c->capturing) return fail(FFVD_E_BADARG, "CUDA-graph capture: pass g_X (no scratch allocation inside a graph)");
      CUDA_TRY(cudaMallocAsync((void**)&t.gx_scratch, t.X.numel * 8, c->stream));
      P.gX = t.gx_scratch;
    }
    if (P.gX && nprob <= 4 && !(flags & FFVD_FLAG_COLLAPSED_RESUME)) CUDA_TRY(cudaMemsetAsync(P.gX, 0, t.X.numel * 8, c->stream));
    OutPtrs& O = ho[p];
    O.nll = t.nll.d; O.terms = t.terms.d; O.g_Z = t.gZ.d; O.g_U = t.gU.d; O.g_logv = t.glogv.d; O.g_logl = t.glogl.d;
    O.g_logQ = t.glogQ.d; O.g_C = t.gC.d; O.g_d = t.gd.d; O.g_logR = t.glogR.d;
  }
  const long long total_items = item;
  // ---- FFVD_FLAG_DETERMINISTIC: private accumulator copies (one per CTA for the S region, one per (CTA, warp) for u-bar / small
  //      gradients / raw terms), three extra x-bar planes and the kzz_bwd partials; see det_ptr1 / det_ptr2 (ffvd_common.cuh)
  const int det_copies = c->num_sms * kMaxCtasPerSm;
  const size_t det_r1 = L.off_ubar - L.off_gZd, det_r2 = L.off_status - L.off_ubar;
  const int kzz_nblk = (M + 7) / 8;
  size_t det_off2 = 0, det_offp = 0, det_offk = 0, det_total = 0, xtot = 0;
  if (det) {
    for (auto& t : pt) xtot += t.X.numel;
    det_off2 = align_up((size_t)det_copies * det_r1, 256);
    det_offp = align_up(det_off2 + (size_t)det_copies * FFVD_DET_MAX_WARPS * det_r2, 256);
    det_offk = align_up(det_offp + 3 * xtot * 8, 256);
    det_total = det_offk + (size_t)nprob * nb * ((size_t)M * Din + (size_t)kzz_nblk * (Din + 1)) * 8;
    if (det_total > c->det_cap) {
      if (c->det_buf) { CUDA_TRY(cudaStreamSynchronize(c->stream)); CUDA_TRY(cudaFree(c->det_buf)); c->det_buf = nullptr; c->det_cap = 0; }
      CUDA_TRY(cudaMalloc((void**)&c->det_buf, det_total));
      c->det_cap = det_total;
    }
    size_t xo = 0;
    for (int p = 0; p < nprob; ++p) {
      DevProblem& P = hp[p];
      P.det1 = c->det_buf; P.det_base1 = c->arena + L.off_gZd; P.det_stride1 = (long long)det_r1;
      P.det2 = c->det_buf + det_off2; P.det_base2 = c->arena + L.off_ubar; P.det_stride2 = (long long)det_r2;
      if (P.gX) { P.gXp = (double*)(c->det_buf + det_offp) + xo; P.gXp_stride = (long long)xtot; }
      P.kzzpart = (double*)(c->det_buf + det_offk) + (size_t)p * nb * ((size_t)M * Din + (size_t)kzz_nblk * (Din + 1));
      xo += pt[p].X.numel;
    }
    CUDA_TRY(cudaMemsetAsync(c->det_buf + det_offp, 0, 3 * xtot * 8, c->stream));
  }
  auto det_zero = [&]() -> int {           // before every fused launch
    if (!det) return FFVD_OK;
    CUDA_TRY(cudaMemsetAsync(c->det_buf, 0, det_offp, c->stream));
    return FFVD_OK;
  };
  auto det_reduce = [&]() -> int {         // after every fused launch: shared accumulators += the private copies, in index order
    if (!det) return FFVD_OK;
    det_reduce_kernel<<<grid1d(det_r1 / 8), 256, 0, c->stream>>>((double*)(c->arena + L.off_gZd), c->det_buf, det_r1, det_r1 / 8, det_copies);
    det_reduce_kernel<<<grid1d(det_r2 / 8), 256, 0, c->stream>>>((double*)(c->arena + L.off_ubar), c->det_buf + det_off2, det_r2, det_r2 / 8,
                                                                 det_copies * FFVD_DET_MAX_WARPS);
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return FFVD_OK;
  };
  c->last_off_cvec = L.off_cvec; c->last_off_HxT = L.off_HxT; c->last_Mp = Mp; c->last_nb = nb;
  const bool p1_only = (flags & FFVD_FLAG_COLLAPSED_P1_ONLY) != 0, resume = (flags & FFVD_FLAG_COLLAPSED_RESUME) != 0;
  if ((p1_only || resume) && (!collapsed || no_grads || nprob != 1)) return fail(FFVD_E_BADARG, "COLLAPSED_P1_ONLY / COLLAPSED_RESUME need one collapsed problem with gradients");
  if ((p1_only || resume) && !pt[0].gX.present) return fail(FFVD_E_BADARG, "COLLAPSED_P1_ONLY / COLLAPSED_RESUME need g_X (pass 1 leaves the emission gradient in it)");
  if (p1_only && resume) return fail(FFVD_E_BADARG, "COLLAPSED_P1_ONLY and COLLAPSED_RESUME are two calls");
  std::vector<long long> callkey = c->arena_key;
  for (auto& t : pt) for (const double* q : {t.X.d, t.Z.d, t.gX.d}) callkey.push_back((long long)(uintptr_t)q);
  if (resume && !(c->p1_pending && callkey == c->p1_key))
    return fail(FFVD_E_BADARG, "COLLAPSED_RESUME without a matching COLLAPSED_P1_ONLY call on this context");
  c->p1_pending = false;
  if (c->capturing && !(flags & FFVD_FLAG_ASYNC)) return fail(FFVD_E_BADARG, "CUDA-graph capture needs FFVD_FLAG_ASYNC (no status read-back inside a graph)");
  {
    const void *srcP = hp.data(), *srcO = ho.data();
    if (c->capturing) {       // the captured copy nodes read their source at every replay: give them memory that outlives this call
      void *pp = nullptr, *po = nullptr;
      CUDA_TRY(cudaMallocHost(&pp, sizeof(DevProblem) * nprob)); c->cur_graph->pinned.push_back(pp);
      CUDA_TRY(cudaMallocHost(&po, sizeof(OutPtrs) * nprob)); c->cur_graph->pinned.push_back(po);
      memcpy(pp, hp.data(), sizeof(DevProblem) * nprob); memcpy(po, ho.data(), sizeof(OutPtrs) * nprob);
      srcP = pp; srcO = po;
    }
    CUDA_TRY(cudaMemcpyAsync(c->d_probs, srcP, sizeof(DevProblem) * nprob, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_outs, srcO, sizeof(OutPtrs) * nprob, cudaMemcpyHostToDevice, c->stream));
  }
  if (!resume) {
    CUDA_TRY(cudaMemsetAsync(c->arena + L.zero_begin, 0, L.zero_end - L.zero_begin, c->stream));
    if (nprob > 4) {           // many small problems: one launch instead of nprob memsets
      size_t maxn = 0;
      for (auto& t : pt) maxn = t.X.numel > maxn ? t.X.numel : maxn;
      zero_gx_kernel<<<dim3(grid1d(maxn, 256, 64), nprob), 256, 0, c->stream>>>(c->d_probs); c->launches++;
    }
  }

  long long ident = 1469598103934665603LL;      // FNV-1a over the addresses of Z / logv / logl of every problem: the factors in
  for (auto& t : pt)                             // the arena belong to exactly these tensors
    for (const double* q : {t.Z.d, t.logv.d, t.logl.d}) ident = (ident ^ (long long)(uintptr_t)q) * 1099511628211LL;
  bool ltu_done = false;
  if (!resume) TRY(launch_prep<KIND>(c, L, jitter, (flags & FFVD_FLAG_REUSE_KZZ) != 0, ident, !collapsed && !no_grads, &ltu_done));
  const int nz = nprob * nb;
  const BatchMap idm = {1, 1, 1};                // z -> z
  const BatchMap lmap = {nb, D, D};              // z -> (z / nb) * D + z % D
  // the pools are contiguous over problems, so problem 0's pointers are the pool bases
  double* Sacc = hp[0].Sacc; double* Wk = hp[0].Wk; double* Linv = hp[0].Linv; double* LinvT = hp[0].LinvT;
  double* Nmat = hp[0].Nmat; double* Hx = hp[0].Hx; double* HxT = hp[0].HxT;
  const dim3 gsym((unsigned)((size_t)M * M + 255) / 256, nb, nprob);

  if (no_grads && !collapsed) {
    TRY(det_zero());
    TRY((launch_fused<KIND, MODE_FORWARD>(c, Mp, Din, c->d_probs, nprob, total_items)));
    TRY(det_reduce());
  } else if (!collapsed) {
    if (!ltu_done) {
      const int chunks = std::max(1, std::min((M + 31) / 32, (4 * c->num_sms) / std::max(1, D * nprob)));     // fill the GPU, one 32-row chunk at least
      ltu_kernel<<<dim3(D, nprob, chunks), 256, 0, c->stream>>>(c->d_probs); c->launches++;
    }
    TRY(det_zero());
    TRY((launch_fused<KIND, MODE_UNCOLLAPSED>(c, Mp, Din, c->d_probs, nprob, total_items)));
    TRY(det_reduce());
    zbar_post_kernel<KIND><<<dim3(D + 1, nprob), 256, 0, c->stream>>>(c->d_probs); c->launches++;
    symmetrize_lower_kernel<<<gsym, 256, 0, c->stream>>>(c->d_probs, 0); c->launches++;
  } else {
    if (!resume) {
      TRY(det_zero());
      TRY((launch_fused<KIND, MODE_COLLAPSED_P1>(c, Mp, Din, c->d_probs, nprob, total_items)));
      TRY(det_reduce());
      symmetrize_lower_kernel<<<gsym, 256, 0, c->stream>>>(c->d_probs, 1); c->launches++;
    }
    if (p1_only) {
      // time-sharded collapsed bound (conditionals_multi_output.py:246-254 sums over ALL transitions): the caller now
      // all-reduces the statistics S = F^T F and b = F^T delta (ffvd_collapsed_stats_allreduce / _get / _set) and calls
      // again with FFVD_FLAG_COLLAPSED_RESUME
      c->p1_pending = true; c->p1_key = callkey;
      c->p1_off_S = L.off_Sacc; c->p1_bytes = (L.off_ubar - L.off_Sacc) + (size_t)nprob * nb * Mp * 8; c->p1_nb = nb; c->p1_Mp = Mp;
      CUDA_TRY(cudaGetLastError());
      int st1 = FFVD_OK;
      if (!(flags & 8)) st1 = check_status(c, L);
      TRY(call.finish());
      return st1;
    }
    if (use_blocked(c, M, Mp)) {
      collapsed_fill_kernel<<<dim3((unsigned)(((size_t)Mp * Mp + 255) / 256), nb, nprob), 256, 0, c->stream>>>(c->d_probs); c->launches++;
      TRY(blocked_factor_invert(c, Wk, (double*)(c->arena + L.off_Dinv), Hx, HxT, (int*)(c->arena + L.off_status2), nz, M, Mp));
      collapsed_logdet_kernel<<<dim3(nb, nprob), 256, 0, c->stream>>>(c->d_probs); c->launches++;
    } else {
      const bool fast = chol_fast_fits(M, Mp, (size_t)c->max_smem) && !getenv("FFVD_NO_FAST_CHOL");
      const size_t sm_chol = fast ? chol_fast_smem_doubles(M, Mp) * 8 : (size_t)2 * Mp * 8;
      CUDA_TRY(cudaFuncSetAttribute(collapsed_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_chol));
      collapsed_chol_kernel<<<dim3(nb, nprob), 512, sm_chol, c->stream>>>(c->d_probs, fast ? 1 : 0); c->launches++;
    }
    if (!no_grads) {
      TRY(launch_bgemm(c, Wk, HxT, Hx, Mp, 1.0, nz, idm, idm, idm, BGEMM_A_UPPER | BGEMM_B_LOWER));          // H^{-1} = L_H^{-T} L_H^{-1}
      collapsed_vec_kernel<<<dim3(nb, nprob), 1024, 0, c->stream>>>(c->d_probs); c->launches++;   // Wk <- Mat'
      TRY(launch_bgemm(c, Hx, Wk, Linv, Mp, 1.0, nz, idm, idm, lmap, BGEMM_B_LOWER));        // Mat' L^{-1}
      TRY(launch_bgemm(c, Nmat, LinvT, Hx, Mp, 1.0, nz, idm, lmap, idm, BGEMM_A_UPPER));     // N = L^{-T} Mat' L^{-1}
      TRY(det_zero());
      TRY((launch_fused<KIND, MODE_COLLAPSED_P2>(c, Mp, Din, c->d_probs, nprob, total_items)));
      TRY(det_reduce());
      zbar_post_kernel<KIND><<<dim3(D + 1, nprob), 256, 0, c->stream>>>(c->d_probs); c->launches++;
      TRY(launch_bgemm(c, HxT, Wk, Sacc, Mp, 1.0, nz, idm, idm, idm));        // Mat' S
      symmetrize_lower_kernel<<<gsym, 256, 0, c->stream>>>(c->d_probs, 2); c->launches++;   // Sacc <- Gs
    } else {
      // forward only still needs c for the quadratic term
      TRY(launch_bgemm(c, Wk, HxT, Hx, Mp, 1.0, nz, idm, idm, idm, BGEMM_A_UPPER | BGEMM_B_LOWER));
      collapsed_vec_kernel<<<dim3(nb, nprob), 1024, 0, c->stream>>>(c->d_probs); c->launches++;
    }
  }
  if (!no_grads && !(collapsed && (flags & FFVD_FLAG_NO_REPLICATED))) {
    // (time-sharded collapsed bound: G = Mat' S + c b^T is built from the all-reduced statistics, identical on every rank,
    //  so only the rank without FFVD_FLAG_NO_REPLICATED pushes it through the Cholesky backward)
    TRY(launch_bgemm(c, Wk, Sacc, Linv, Mp, 1.0, nz, idm, idm, lmap, BGEMM_B_LOWER));        // Gs L^{-1}
    TRY(launch_bgemm(c, Sacc, LinvT, Wk, Mp, -0.5, nz, idm, lmap, idm, BGEMM_A_UPPER));      // Kbar_zz = -1/2 L^{-T} Gs L^{-1}
    const dim3 grow((M + 7) / 8, nb, nprob);
    if (getenv("FFVD_SPLIT_KZZ_BWD")) {                     // the two-kernel form (kept for A/B timing)
      wz_kernel<KIND><<<grow, 256, 0, c->stream>>>(c->d_probs); c->launches++;
      kzz_bwd_kernel<KIND><<<grow, 256, 0, c->stream>>>(c->d_probs); c->launches++;
    } else {
      kzz_bwd_fused_kernel<KIND><<<grow, 256, 0, c->stream>>>(c->d_probs); c->launches++;
      if (det) { kzz_bwd_reduce_kernel<<<nprob, 256, 0, c->stream>>>(c->d_probs, nb, (int)grow.x); c->launches++; }
    }
  }
  {
    int gx_blocks = 0;
    if (!no_grads) {
      size_t maxn = 0;
      for (auto& t : pt) maxn = t.X.numel > maxn ? t.X.numel : maxn;
      gx_blocks = grid1d(maxn);
    }
    finalize_kernel<KIND><<<dim3(1 + gx_blocks, nprob), 256, 0, c->stream>>>(c->d_probs, c->d_outs, collapsed, flags, gx_blocks,
                                                                             nb > D ? nb : D, (const int*)(c->arena + L.off_status2), L.nfac);
    c->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  for (auto& t : pt)
    if (t.gx_scratch) CUDA_TRY(cudaFreeAsync(t.gx_scratch, c->stream));
  int st = FFVD_OK;
  if (!(flags & 8)) st = check_status(c, L);
  TRY(call.finish());
  return st;
}

// One all-reduce (sum, float64) of a list of device tensors packed into one buffer.
static int allreduce_list(ffvd_ctx* c, Call& call, DLManagedTensor* const* list, const char* const* names, int n) {
  if (!c->comm || c->comm_nranks == 1) return FFVD_OK;          // single rank: the sum is the value
  if (c->capturing) return fail(FFVD_E_UNSUPPORTED, "collectives are not captured in CUDA graphs");
  const NcclApi* a = nccl_api(nullptr);
  if (!a) return fail(FFVD_E_UNSUPPORTED, "NCCL is not loaded");
  PackList pl;
  pl.n = 0; pl.off[0] = 0;
  for (int i = 0; i < n; ++i) {
    if (!list[i]) continue;
    Tens t;
    TRY(call.import(list[i], false, t, names[i]));
    if (t.staged) return fail(FFVD_E_DEVICE, std::string(names[i]) + ": the collective works on device tensors only");
    if (t.numel == 0) continue;
    if (pl.n == 12) return fail(FFVD_E_LIMIT, "at most 12 tensors per packed all-reduce");
    pl.ptr[pl.n] = t.d;
    pl.off[pl.n + 1] = pl.off[pl.n] + (long long)t.numel;
    pl.n++;
  }
  const size_t total = (size_t)pl.off[pl.n];
  if (!total) return FFVD_OK;
  if (c->pack_cap < total) {
    if (c->pack_buf) { CUDA_TRY(cudaStreamSynchronize(c->stream)); CUDA_TRY(cudaFree(c->pack_buf)); }
    CUDA_TRY(cudaMalloc((void**)&c->pack_buf, total * sizeof(double)));
    c->pack_cap = total;
  }
  pack_kernel<<<grid1d(total), 256, 0, c->stream>>>(pl, c->pack_buf, 0); c->launches++;
  NCCL_TRY(a, a->AllReduce(c->pack_buf, c->pack_buf, total, kNcclFloat64, kNcclSum, c->comm, c->stream));
  pack_kernel<<<grid1d(total), 256, 0, c->stream>>>(pl, c->pack_buf, 1); c->launches++;
  CUDA_TRY(cudaGetLastError());
  return FFVD_OK;
}

extern "C" int ffvd_allreduce_shared(ffvd_ctx* c, const ffvd_outputs* o, int with_scalars) {
  if (!c || !o) return fail(FFVD_E_BADARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  Call call(c);
  DLManagedTensor* list[10] = {o->g_Z, o->g_U, o->g_logv, o->g_logl, o->g_logQ, o->g_C, o->g_d, o->g_logR,
                               with_scalars ? o->nll : nullptr, with_scalars ? o->terms : nullptr};
  const char* names[10] = {"g_Z", "g_U", "g_logv", "g_logl", "g_logQ", "g_C", "g_d", "g_logR", "nll", "terms"};
  TRY(allreduce_list(c, call, list, names, 10));
  return call.finish();
}

extern "C" int ffvd_allreduce(ffvd_ctx* c, DLManagedTensor* t) {
  if (!c || !t) return fail(FFVD_E_BADARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  Call call(c);
  const char* nm = "tensor";
  TRY(allreduce_list(c, call, &t, &nm, 1));
  return call.finish();
}

extern "C" int ffvd_nll_grads_uncollapsed(ffvd_ctx* c, int kind, const ffvd_problem* p, int flags, double jitter,
                                          const ffvd_outputs* o) {
  if (kind == FFVD_KERNEL_SE) return run_nll<0>(c, 0, 1, p, flags, jitter, o);
  if (kind == FFVD_KERNEL_LINEAR) return run_nll<1>(c, 0, 1, p, flags, jitter, o);
  return fail(FFVD_E_BADARG, "unknown kernel kind");
}
extern "C" int ffvd_nll_grads_collapsed(ffvd_ctx* c, int kind, const ffvd_problem* p, int flags, double jitter,
                                        const ffvd_outputs* o) {
  if (kind == FFVD_KERNEL_SE) return run_nll<0>(c, 1, 1, p, flags, jitter, o);
  if (kind == FFVD_KERNEL_LINEAR) return run_nll<1>(c, 1, 1, p, flags, jitter, o);
  return fail(FFVD_E_BADARG, "unknown kernel kind");
}
// ---- statistics of a pending collapsed pass 1 (time sharding): S = F^T F (nb,Mp,Mp) and b = F^T delta (nb,Mp)
extern "C" int ffvd_collapsed_stats_shape(ffvd_ctx* c, int* nb, int* Mp) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  if (!c->p1_pending) return fail(FFVD_E_BADARG, "no pending FFVD_FLAG_COLLAPSED_P1_ONLY evaluation on this context");
  if (nb) *nb = c->p1_nb;
  if (Mp) *Mp = c->p1_Mp;
  return FFVD_OK;
}

extern "C" int ffvd_collapsed_stats_allreduce(ffvd_ctx* c) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  if (!c->p1_pending) return fail(FFVD_E_BADARG, "no pending FFVD_FLAG_COLLAPSED_P1_ONLY evaluation on this context");
  if (!c->comm || c->comm_nranks == 1) return FFVD_OK;
  const NcclApi* a = nccl_api(nullptr);
  if (!a) return fail(FFVD_E_UNSUPPORTED, "NCCL is not loaded");
  CUDA_TRY(cudaSetDevice(c->device));
  // S and b are adjacent in the arena (b follows S after alignment padding, which is zero on every rank)
  double* base = (double*)(c->arena + c->p1_off_S);
  NCCL_TRY(a, a->AllReduce(base, base, c->p1_bytes / sizeof(double), kNcclFloat64, kNcclSum, c->comm, c->stream));
  return FFVD_OK;
}

static int collapsed_stats_copy(ffvd_ctx* c, DLManagedTensor* S, DLManagedTensor* b, bool set) {
  if (!c || !S || !b) return fail(FFVD_E_BADARG, "null argument");
  if (!c->p1_pending) return fail(FFVD_E_BADARG, "no pending FFVD_FLAG_COLLAPSED_P1_ONLY evaluation on this context");
  CUDA_TRY(cudaSetDevice(c->device));
  Call call(c);
  Tens tS, tb;
  TRY(call.import(S, !set, tS, "S"));
  TRY(call.import(b, !set, tb, "b"));
  const size_t nS = (size_t)c->p1_nb * c->p1_Mp * c->p1_Mp, nb_ = (size_t)c->p1_nb * c->p1_Mp;
  if (tS.numel != nS || tb.numel != nb_) return fail(FFVD_E_SHAPE, "S must be (nb,Mp,Mp) and b (nb,Mp): see ffvd_collapsed_stats_shape");
  double* dS = (double*)(c->arena + c->p1_off_S);
  double* db = (double*)(c->arena + c->p1_off_S + c->p1_bytes) - nb_;
  if (set) {
    CUDA_TRY(cudaMemcpyAsync(dS, tS.d, nS * 8, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(db, tb.d, nb_ * 8, cudaMemcpyDeviceToDevice, c->stream));
  } else {
    CUDA_TRY(cudaMemcpyAsync(tS.d, dS, nS * 8, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(tb.d, db, nb_ * 8, cudaMemcpyDeviceToDevice, c->stream));
  }
  return call.finish();
}
extern "C" int ffvd_collapsed_stats_get(ffvd_ctx* c, DLManagedTensor* S_out, DLManagedTensor* b_out) { return collapsed_stats_copy(c, S_out, b_out, false); }
extern "C" int ffvd_collapsed_stats_set(ffvd_ctx* c, DLManagedTensor* S_in, DLManagedTensor* b_in) { return collapsed_stats_copy(c, S_in, b_in, true); }

// (S,D,Mp) -> U_mean (S,M,D) and (S,D,Mp,Mp) -> LHinvT (S,D,M,M)
__global__ void extract_collapsed_kernel(const double* __restrict__ cvec, const double* __restrict__ HxT, double* __restrict__ Umean,
                                         double* __restrict__ LHinvT, int S, int D, int M, int Mp) {
  const size_t n1 = (size_t)S * M * D, n2 = LHinvT ? (size_t)S * D * M * M : 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += (size_t)gridDim.x * blockDim.x) {
    if (i < n1) {
      const int d = (int)(i % D), m = (int)((i / D) % M), s = (int)(i / ((size_t)D * M));
      Umean[i] = cvec[((size_t)s * D + d) * Mp + m];
    } else {
      const size_t j = i - n1;
      const int c0 = (int)(j % M), r0 = (int)((j / M) % M);
      const size_t b = j / ((size_t)M * M);
      LHinvT[j] = HxT[b * Mp * Mp + (size_t)r0 * Mp + c0];
    }
  }
}

extern "C" int ffvd_collapse_u_mean(ffvd_ctx* c, int kind, const ffvd_problem* p, double jitter, DLManagedTensor* U_mean_out,
                                    DLManagedTensor* LHinvT_out) {
  if (!c || !p || !U_mean_out) return fail(FFVD_E_BADARG, "null argument");
  ffvd_outputs o;
  memset(&o, 0, sizeof o);
  // collapsed forward pass: S = F^T F, b = F^T delta, H = S/Q + I, chol(H), c = H^{-1} b / Q  (left in the context's pools)
  int st;
  if (kind == FFVD_KERNEL_SE) st = run_nll<0>(c, 1, 1, p, FFVD_FLAG_NO_GRADS, jitter, &o);
  else if (kind == FFVD_KERNEL_LINEAR) st = run_nll<1>(c, 1, 1, p, FFVD_FLAG_NO_GRADS, jitter, &o);
  else return fail(FFVD_E_BADARG, "unknown kernel kind");
  if (st != FFVD_OK) return st;
  Call call(c);
  Tens tX, tZ, tu, tl;
  TRY(call.import(p->X, false, tX, "X"));
  TRY(call.import(p->Z, false, tZ, "Z"));
  TRY(call.import(U_mean_out, true, tu, "U_mean_out"));
  TRY(call.import(LHinvT_out, true, tl, "LHinvT_out", true));
  const int S = tX.ndim == 3 ? (int)tX.shape[0] : 1, D = (int)tX.shape[tX.ndim - 1], M = (int)tZ.shape[0];
  if (tu.numel != (size_t)S * M * D) return fail(FFVD_E_SHAPE, "U_mean_out must be (S,M,D) [(M,D) for 2-d X]");
  if (tl.present && tl.numel != (size_t)S * D * M * M) return fail(FFVD_E_SHAPE, "LHinvT_out must be (S,D,M,M)");
  // NOTE: staging X / Z again for host tensors is wasteful but harmless; the pools were written by run_nll above
  extract_collapsed_kernel<<<grid1d((size_t)S * M * D + (tl.present ? (size_t)S * D * M * M : 0)), 256, 0, c->stream>>>(
      (const double*)(c->arena + c->last_off_cvec), (const double*)(c->arena + c->last_off_HxT), tu.d, tl.present ? tl.d : nullptr, S, D, M,
      c->last_Mp);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  return call.finish();
}

extern "C" int ffvd_nll_grads_batched(ffvd_ctx* c, int kind, int collapsed, int nprob, const ffvd_problem* p, int flags,
                                      double jitter, const ffvd_outputs* o) {
  if (kind == FFVD_KERNEL_SE) return run_nll<0>(c, collapsed, nprob, p, flags, jitter, o);
  if (kind == FFVD_KERNEL_LINEAR) return run_nll<1>(c, collapsed, nprob, p, flags, jitter, o);
  return fail(FFVD_E_BADARG, "unknown kernel kind");
}

// ---------------------------------------------------------------------------------------------
// operator-level entry points
extern "C" int ffvd_kernel_K(ffvd_ctx* c, int kind, DLManagedTensor* X, DLManagedTensor* X2, DLManagedTensor* logv,
                             DLManagedTensor* logl, DLManagedTensor* out) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  CUDA_TRY(cudaSetDevice(c->device));
  Call call(c);
  Tens tX, tX2, tv, tl, to;
  TRY(call.import(X, false, tX, "X"));
  TRY(call.import(X2, false, tX2, "X2", true));
  TRY(call.import(logv, false, tv, "logv"));
  TRY(call.import(logl, false, tl, "logl", kind != FFVD_KERNEL_SE));
  TRY(call.import(out, true, to, "out"));
  if (tX.ndim != 2) return fail(FFVD_E_SHAPE, "X must be (N,Din)");
  const int N = (int)tX.shape[0], Din = (int)tX.shape[1];
  const Tens& t2 = tX2.present ? tX2 : tX;
  if (t2.ndim != 2 || t2.shape[1] != Din) return fail(FFVD_E_SHAPE, "X2 must be (N2,Din)");
  const int N2 = (int)t2.shape[0];
  if (tv.numel != 1) return fail(FFVD_E_SHAPE, "logv must be a scalar");
  if (kind == FFVD_KERNEL_SE && tl.numel != (size_t)Din) return fail(FFVD_E_SHAPE, "logl must be (Din)");
  if (to.numel != (size_t)N * N2) return fail(FFVD_E_SHAPE, "out must be (N,N2)");
  if (to.numel) {
    const int grid = grid1d(to.numel);
    if (kind == FFVD_KERNEL_SE) kernel_K_kernel<0><<<grid, 256, 0, c->stream>>>(tX.d, t2.d, N, N2, Din, tv.d, tl.d, to.d);
    else if (kind == FFVD_KERNEL_LINEAR) kernel_K_kernel<1><<<grid, 256, 0, c->stream>>>(tX.d, t2.d, N, N2, Din, tv.d, nullptr, to.d);
    else return fail(FFVD_E_BADARG, "unknown kernel kind");
    c->launches++;
  }
  return call.finish();
}

extern "C" int ffvd_kernel_Kdiag(ffvd_ctx* c, int kind, DLManagedTensor* X, DLManagedTensor* logv, DLManagedTensor* logl,
                                 DLManagedTensor* out) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  CUDA_TRY(cudaSetDevice(c->device));
  Call call(c);
  Tens tX, tv, tl, to;
  TRY(call.import(X, false, tX, "X"));
  TRY(call.import(logv, false, tv, "logv"));
  TRY(call.import(logl, false, tl, "logl", true));
  TRY(call.import(out, true, to, "out"));
  if (tX.ndim != 2) return fail(FFVD_E_SHAPE, "X must be (N,Din)");
  const int N = (int)tX.shape[0], Din = (int)tX.shape[1];
  if (tv.numel != 1) return fail(FFVD_E_SHAPE, "logv must be a scalar");
  if (to.numel != (size_t)N) return fail(FFVD_E_SHAPE, "out must be (N)");
  if (N) {
    if (kind == FFVD_KERNEL_SE) kernel_Kdiag_kernel<0><<<grid1d(N), 256, 0, c->stream>>>(tX.d, N, Din, tv.d, to.d);
    else if (kind == FFVD_KERNEL_LINEAR) kernel_Kdiag_kernel<1><<<grid1d(N), 256, 0, c->stream>>>(tX.d, N, Din, tv.d, to.d);
    else return fail(FFVD_E_BADARG, "unknown kernel kind");
    c->launches++;
  }
  return call.finish();
}

// shared prep for kernel_pre_cal / conditional: one DevProblem with only the Z-side fields.
static int setup_zside(ffvd_ctx* c, int kind, const Tens& tZ, const Tens& tv, const Tens& tl, int nk, int R, double jitter,
                       Layout& L, DevProblem& P, bool need_scratch = false, bool reuse = false) {
  if (c->capturing) return fail(FFVD_E_UNSUPPORTED, "only ffvd_nll_grads_*, ffvd_sghmc_update and ffvd_adam_update can be captured in a CUDA graph");
  const int M = (int)tZ.shape[0], Din = (int)tZ.shape[1];
  if (Din > FFVD_MAX_DIN) return fail(FFVD_E_LIMIT, "Din > 31");
  const int Mp = pad_M(M);
  if (Mp < 0) return fail(FFVD_E_LIMIT, "M > 2048 is not supported by this build");
  L = make_layout(c, 1, need_scratch ? R : nk, nk, R, M, Mp, Din, 1, 1, false, need_scratch);
  TRY(ensure_arena(c, L));
  memset(&P, 0, sizeof P);
  P.Z = tZ.d; P.logv = tv.d; P.logl = tl.d;
  P.M = M; P.Mp = Mp; P.Din = Din; P.D = R; P.hs = (nk == 1 && R != 1) ? 0 : 1; P.S = 1;
  if (nk == 1) P.hs = 0;
  bind_problem(c, L, 0, 0, P);
  CUDA_TRY(cudaMemcpyAsync(c->d_probs, &P, sizeof P, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemsetAsync(c->arena + L.zero_begin, 0, L.zero_end - L.zero_begin, c->stream));
  long long ident = 1469598103934665603LL;
  for (const double* q : {tZ.d, tv.d, tl.d}) ident = (ident ^ (long long)(uintptr_t)q) * 1099511628211LL;
  if (tZ.staged || tv.staged || tl.staged) reuse = false;      // staged host tensors get fresh device copies every call
  if (kind == FFVD_KERNEL_SE) TRY(launch_prep<0>(c, L, jitter, reuse, ident));
  else if (kind == FFVD_KERNEL_LINEAR) TRY(launch_prep<1>(c, L, jitter, reuse, ident));
  else return fail(FFVD_E_BADARG, "unknown kernel kind");
  return FFVD_OK;
}

extern "C" int ffvd_kernel_pre_cal(ffvd_ctx* c, int kind, DLManagedTensor* Z, DLManagedTensor* logv, DLManagedTensor* logl,
                                   double jitter, DLManagedTensor* LinvT_out) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  CUDA_TRY(cudaSetDevice(c->device));
  Call call(c);
  Tens tZ, tv, tl, to;
  TRY(call.import(Z, false, tZ, "Z"));
  TRY(call.import(logv, false, tv, "logv"));
  TRY(call.import(logl, false, tl, "logl", kind != FFVD_KERNEL_SE));
  TRY(call.import(LinvT_out, true, to, "LinvT_out"));
  if (tZ.ndim != 2) return fail(FFVD_E_SHAPE, "Z must be (M,Din)");
  const int M = (int)tZ.shape[0], Din = (int)tZ.shape[1];
  const int D = (int)tv.numel;
  if (D < 1) return fail(FFVD_E_SHAPE, "logv must be (D)");
  if (kind == FFVD_KERNEL_SE && tl.numel != (size_t)D * Din) return fail(FFVD_E_SHAPE, "logl must be (D,Din)");
  if (to.numel != (size_t)D * M * M) return fail(FFVD_E_SHAPE, "LinvT_out must be (D,M,M)");
  Layout L; DevProblem P;
  TRY(setup_zside(c, kind, tZ, tv, tl, D, D, jitter, L, P));
  P.hs = 1;
  for (int d = 0; d < D; ++d)
    CUDA_TRY(cudaMemcpy2DAsync(to.d + (size_t)d * M * M, (size_t)M * 8, P.LinvT + (size_t)d * P.Mp * P.Mp, (size_t)P.Mp * 8,
                               (size_t)M * 8, M, cudaMemcpyDeviceToDevice, c->stream));
  int st = check_status(c, L);
  TRY(call.finish());
  return st;
}

// u'[:,d] = L_d^{-1} f[:,d]  (non-white conditional).  grid (R); block 256
__global__ void unwhiten_kernel(const DevProblem* __restrict__ probs, const double* __restrict__ f, double* __restrict__ out) {
  const DevProblem& P = probs[0];
  const int d = blockIdx.x, M = P.M, Mp = P.Mp, D = P.D;
  const double* Li = P.Linv + (size_t)d * P.hs * Mp * Mp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = warp; m < M; m += 8) {
    double t = 0.0;
    for (int n = lane; n <= m; n += 32) t = fma(Li[(size_t)m * Mp + n], f[(size_t)n * D + d], t);
    t = warp_sum(t);
    if (lane == 0) out[(size_t)m * D + d] = t;
  }
}

extern "C" int ffvd_conditional_ex(ffvd_ctx* c, int kind, int shared_kernel, DLManagedTensor* Xnew, DLManagedTensor* Z,
                                   DLManagedTensor* logv, DLManagedTensor* logl, DLManagedTensor* f, DLManagedTensor* q_sqrt,
                                   int white, int full_cov, double jitter, int flags, DLManagedTensor* mean_out,
                                   DLManagedTensor* var_out);

extern "C" int ffvd_conditional(ffvd_ctx* c, int kind, int shared_kernel, DLManagedTensor* Xnew, DLManagedTensor* Z,
                                DLManagedTensor* logv, DLManagedTensor* logl, DLManagedTensor* f, DLManagedTensor* q_sqrt,
                                int white, int full_cov, double jitter, DLManagedTensor* mean_out, DLManagedTensor* var_out) {
  return ffvd_conditional_ex(c, kind, shared_kernel, Xnew, Z, logv, logl, f, q_sqrt, white, full_cov, jitter, 0, mean_out, var_out);
}

extern "C" int ffvd_conditional_ex(ffvd_ctx* c, int kind, int shared_kernel, DLManagedTensor* Xnew, DLManagedTensor* Z,
                                   DLManagedTensor* logv, DLManagedTensor* logl, DLManagedTensor* f, DLManagedTensor* q_sqrt,
                                   int white, int full_cov, double jitter, int flags, DLManagedTensor* mean_out,
                                   DLManagedTensor* var_out) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  if (full_cov) return fail(FFVD_E_UNSUPPORTED, "conditional: full_cov=True (N x N covariances) is not built; the reference driver "
                                                "runs with full_cov=False (FFVD_Main.py:267)");
  if (q_sqrt && !white) return fail(FFVD_E_UNSUPPORTED, "conditional: q_sqrt with white=False is not built");
  CUDA_TRY(cudaSetDevice(c->device));
  Call call(c);
  Tens tX, tZ, tv, tl, tf, tm, tvar, tq;
  TRY(call.import(q_sqrt, false, tq, "q_sqrt", true));
  TRY(call.import(Xnew, false, tX, "Xnew"));
  TRY(call.import(Z, false, tZ, "Z"));
  TRY(call.import(logv, false, tv, "logv"));
  TRY(call.import(logl, false, tl, "logl", kind != FFVD_KERNEL_SE));
  TRY(call.import(f, false, tf, "f"));
  TRY(call.import(mean_out, true, tm, "mean_out"));
  TRY(call.import(var_out, true, tvar, "var_out"));
  if (tX.ndim != 2 || tZ.ndim != 2 || tf.ndim != 2) return fail(FFVD_E_SHAPE, "Xnew, Z, f must be 2-d");
  const int N = (int)tX.shape[0], Din = (int)tZ.shape[1], M = (int)tZ.shape[0], R = (int)tf.shape[1];
  if (tX.shape[1] != Din || tf.shape[0] != M) return fail(FFVD_E_SHAPE, "Xnew (N,Din), Z (M,Din), f (M,R) expected");
  const int nk = shared_kernel ? 1 : R;
  if (tv.numel != (size_t)nk) return fail(FFVD_E_SHAPE, "logv must have one entry per kernel");
  if (kind == FFVD_KERNEL_SE && tl.numel != (size_t)nk * Din) return fail(FFVD_E_SHAPE, "logl must be (kernels,Din)");
  if (tm.numel != (size_t)N * R || tvar.numel != (size_t)N * R) return fail(FFVD_E_SHAPE, "mean/var must be (N,R)");
  if (N == 0) return call.finish();
  Layout L; DevProblem P;
  TRY(setup_zside(c, kind, tZ, tv, tl, nk, R, jitter, L, P, tq.present && tq.ndim == 3, (flags & FFVD_FLAG_REUSE_KZZ) != 0));
  const int RB = rb_for_call(c, P.Mp, R, N), BT = 8 * RB;
  c->cur_rb = RB;
  P.hs = shared_kernel ? 0 : 1;
  P.X = tX.d; P.S = 1; P.T = N; P.xrows = N; P.Dx = Din; P.nc = 0; P.Dy = 1;
  P.ntiles = (N + BT - 1) / BT; P.dblk = dblk_of(P.Mp, R); P.item_begin = 0;
  set_items(c, P, (long long)P.ntiles);
  P.cond_mean = tm.d; P.cond_var = tvar.d;
  if (tq.present) {
    if (tq.ndim == 2) {                       // (M,R) per-point scales, conditionals_multi_output.py:51-52
      if (tq.shape[0] != M || tq.shape[1] != R) return fail(FFVD_E_SHAPE, "2-d q_sqrt must be (M,R)");
      P.qmode = 2; P.nq = R; P.qmat = tq.d;
    } else if (tq.ndim == 3) {                // (R,M,M) or (1,M,M) factors, :53-62
      if (tq.shape[1] != M || tq.shape[2] != M || (tq.shape[0] != R && tq.shape[0] != 1))
        return fail(FFVD_E_SHAPE, "3-d q_sqrt must be (R,M,M) or (1,M,M)");
      P.qmode = 3; P.nq = (int)tq.shape[0];
      double* qpad = (double*)(c->arena + L.off_Wk);      // zero padded (nq,Mp,Mp) copies
      CUDA_TRY(cudaMemsetAsync(qpad, 0, (size_t)P.nq * P.Mp * P.Mp * 8, c->stream));
      for (int i = 0; i < P.nq; ++i)
        CUDA_TRY(cudaMemcpy2DAsync(qpad + (size_t)i * P.Mp * P.Mp, (size_t)P.Mp * 8, tq.d + (size_t)i * M * M, (size_t)M * 8,
                                   (size_t)M * 8, M, cudaMemcpyDeviceToDevice, c->stream));
      P.qmat = qpad;
    } else {
      return fail(FFVD_E_SHAPE, "Bad dimension for q_sqrt");       // conditionals_multi_output.py:57-59
    }
  }
  double* utmp = (double*)(c->arena + L.off_utmp);
  P.U = tf.d;
  CUDA_TRY(cudaMemcpyAsync(c->d_probs, &P, sizeof P, cudaMemcpyHostToDevice, c->stream));
  if (!white) {
    unwhiten_kernel<<<R, 256, 0, c->stream>>>(c->d_probs, tf.d, utmp); c->launches++;
    P.U = utmp;
    CUDA_TRY(cudaMemcpyAsync(c->d_probs, &P, sizeof P, cudaMemcpyHostToDevice, c->stream));
  }
  // U (possibly un-whitened) was bound after setup_zside ran the prep kernels: refresh U^T
  hyper_kernel<<<dim3(nk > R ? nk : R, 1), 128, 0, c->stream>>>(c->d_probs, kind, nk, 0, 0); c->launches++;
  if (kind == FFVD_KERNEL_SE) TRY((launch_fused<0, MODE_COND>(c, P.Mp, P.Din, c->d_probs, 1, P.nitems)));
  else TRY((launch_fused<1, MODE_COND>(c, P.Mp, P.Din, c->d_probs, 1, P.nitems)));
  int st = FFVD_OK;
  if (!(flags & FFVD_FLAG_ASYNC)) st = check_status(c, L);
  TRY(call.finish());
  return st;
}

// conditionals.py:6-107 / conditionals_multi_output.py:6-120 on explicit matrices: every option of base_conditional
// (full_cov, q_sqrt 2-d / 3-d, white or not, return_Lm).  For a handful of prediction points; the hot-path branch
// (white, diagonal variances) is ffvd_conditional_ex.
static void launch_gemm_small(ffvd_ctx* c, double* C, int ldc, const double* A, int lda, int ta, const double* B, int ldb, int tb,
                              int m, int n, int k, double alpha, double beta) {
  dgemm_small_kernel<<<dim3((n + 31) / 32, (m + 31) / 32), dim3(32, 8), 0, c->stream>>>(C, ldc, A, lda, ta, B, ldb, tb, m, n, k, alpha, beta);
  c->launches++;
}

extern "C" int ffvd_conditional_dense(ffvd_ctx* c, int kind, int shared_kernel, DLManagedTensor* Xnew, DLManagedTensor* Z,
                                      DLManagedTensor* logv, DLManagedTensor* logl, DLManagedTensor* f, DLManagedTensor* q_sqrt,
                                      int white, int full_cov, double jitter, DLManagedTensor* mean_out, DLManagedTensor* var_out,
                                      DLManagedTensor* Lm_out) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  if (kind != FFVD_KERNEL_SE && kind != FFVD_KERNEL_LINEAR) return fail(FFVD_E_BADARG, "unknown kernel kind");
  CUDA_TRY(cudaSetDevice(c->device));
  Call call(c);
  Tens tX, tZ, tv, tl, tf, tm, tvar, tq, tL;
  TRY(call.import(q_sqrt, false, tq, "q_sqrt", true));
  TRY(call.import(Xnew, false, tX, "Xnew"));
  TRY(call.import(Z, false, tZ, "Z"));
  TRY(call.import(logv, false, tv, "logv"));
  TRY(call.import(logl, false, tl, "logl", kind != FFVD_KERNEL_SE));
  TRY(call.import(f, false, tf, "f"));
  TRY(call.import(mean_out, true, tm, "mean_out"));
  TRY(call.import(var_out, true, tvar, "var_out"));
  TRY(call.import(Lm_out, true, tL, "Lm_out", true));
  if (tX.ndim != 2 || tZ.ndim != 2 || tf.ndim != 2) return fail(FFVD_E_SHAPE, "Xnew, Z, f must be 2-d");
  const int N = (int)tX.shape[0], Din = (int)tZ.shape[1], M = (int)tZ.shape[0], R = (int)tf.shape[1];
  if (tX.shape[1] != Din || tf.shape[0] != M) return fail(FFVD_E_SHAPE, "Xnew (N,Din), Z (M,Din), f (M,R) expected");
  const int nk = shared_kernel ? 1 : R;
  if (tv.numel != (size_t)nk) return fail(FFVD_E_SHAPE, "logv must have one entry per kernel");
  if (kind == FFVD_KERNEL_SE && tl.numel != (size_t)nk * Din) return fail(FFVD_E_SHAPE, "logl must be (kernels,Din)");
  if (tm.numel != (size_t)N * R) return fail(FFVD_E_SHAPE, "mean must be (N,R)");
  if (tvar.numel != (full_cov ? (size_t)R * N * N : (size_t)N * R)) return fail(FFVD_E_SHAPE, "var must be (N,R), or (R,N,N) with full_cov");
  if (tL.present && tL.numel != (size_t)nk * M * M) return fail(FFVD_E_SHAPE, "Lm_out must be (kernels,M,M)");
  int qmode = 0, nq = 0;
  if (tq.present) {
    if (tq.ndim == 2) { if (tq.shape[0] != M || tq.shape[1] != R) return fail(FFVD_E_SHAPE, "2-d q_sqrt must be (M,R)"); qmode = 2; nq = R; }
    else if (tq.ndim == 3) {
      if (tq.shape[1] != M || tq.shape[2] != M || (tq.shape[0] != R && tq.shape[0] != 1)) return fail(FFVD_E_SHAPE, "3-d q_sqrt must be (R,M,M) or (1,M,M)");
      qmode = 3; nq = (int)tq.shape[0];
    } else return fail(FFVD_E_SHAPE, "Bad dimension for q_sqrt");
  }
  if (N == 0) return call.finish();
  Layout L; DevProblem P;
  c->force_blocked = tL.present;
  int st0 = setup_zside(c, kind, tZ, tv, tl, nk, R, jitter, L, P, false, false);
  c->force_blocked = false;
  c->kzz_valid = false;                     // (a forced blocked factorisation must not be mistaken for the default path's)
  if (st0 != FFVD_OK) return st0;
  const int Mp = P.Mp;
  double *Kmn = nullptr, *A = nullptr, *A2 = nullptr, *LTA = nullptr, *Knn = nullptr;
  const size_t mn = (size_t)M * N;
  CUDA_TRY(cudaMallocAsync((void**)&Kmn, mn * 8, c->stream));
  CUDA_TRY(cudaMallocAsync((void**)&A, mn * 8, c->stream));
  CUDA_TRY(cudaMallocAsync((void**)&A2, mn * 8, c->stream));
  CUDA_TRY(cudaMallocAsync((void**)&LTA, mn * 8, c->stream));
  CUDA_TRY(cudaMallocAsync((void**)&Knn, (full_cov ? (size_t)N * N : (size_t)N) * 8, c->stream));
  const double* Ause = nullptr;
  for (int r = 0; r < R; ++r) {
    const int k = shared_kernel ? 0 : r;
    double* var_r = full_cov ? tvar.d + (size_t)r * N * N : tvar.d + r;
    const int ldv = full_cov ? N : R;
    if (r == 0 || !shared_kernel) {
      const double* lv = tv.d + k;
      const double* ll = kind == FFVD_KERNEL_SE ? tl.d + (size_t)k * Din : nullptr;
      const double* Li = P.Linv + (size_t)k * Mp * Mp;
      const double* LiT = P.LinvT + (size_t)k * Mp * Mp;
      if (kind == FFVD_KERNEL_SE) kernel_K_kernel<0><<<grid1d(mn), 256, 0, c->stream>>>(tZ.d, tX.d, M, N, Din, lv, ll, Kmn);
      else kernel_K_kernel<1><<<grid1d(mn), 256, 0, c->stream>>>(tZ.d, tX.d, M, N, Din, lv, nullptr, Kmn);
      c->launches++;
      if (full_cov) {
        if (kind == FFVD_KERNEL_SE) kernel_K_kernel<0><<<grid1d((size_t)N * N), 256, 0, c->stream>>>(tX.d, tX.d, N, N, Din, lv, ll, Knn);
        else kernel_K_kernel<1><<<grid1d((size_t)N * N), 256, 0, c->stream>>>(tX.d, tX.d, N, N, Din, lv, nullptr, Knn);
      } else {
        if (kind == FFVD_KERNEL_SE) kernel_Kdiag_kernel<0><<<grid1d(N), 256, 0, c->stream>>>(tX.d, N, Din, lv, Knn);
        else kernel_Kdiag_kernel<1><<<grid1d(N), 256, 0, c->stream>>>(tX.d, N, Din, lv, Knn);
      }
      c->launches++;
      launch_gemm_small(c, A, N, Li, Mp, 0, Kmn, N, 0, M, N, M, 1.0, 0.0);                       // A = L^{-1} Kmn        (:37 / cmo:37)
      Ause = A;
      if (!white) { launch_gemm_small(c, A2, N, LiT, Mp, 0, A, N, 0, M, N, M, 1.0, 0.0); Ause = A2; }   // A = L^{-T} A (:45-46)
    }
    // fvar = Knn - A^T A with the FIRST A (before the non-white substitution), conditionals.py:39-44
    if (full_cov) {
      CUDA_TRY(cudaMemcpyAsync(var_r, Knn, (size_t)N * N * 8, cudaMemcpyDeviceToDevice, c->stream));
      launch_gemm_small(c, var_r, N, A, N, 1, A, N, 0, N, N, M, -1.0, 1.0);
    } else {
      colsumsq_kernel<<<grid1d(N), 256, 0, c->stream>>>(var_r, ldv, Knn, A, M, N, nullptr, 0, -1.0, 0); c->launches++;
    }
    launch_gemm_small(c, tm.d + r, R, Ause, N, 1, tf.d + r, R, 0, N, 1, M, 1.0, 0.0);              // fmean = A^T f          (:48)
    if (qmode == 3) {
      const double* qr = tq.d + (size_t)(nq == 1 ? 0 : r) * M * M;
      launch_gemm_small(c, LTA, N, qr, M, 1, Ause, N, 0, M, N, M, 1.0, 0.0);                       // LTA = q^T A            (:53-55)
      if (full_cov) launch_gemm_small(c, var_r, N, LTA, N, 1, LTA, N, 0, N, N, M, 1.0, 1.0);
      else { colsumsq_kernel<<<grid1d(N), 256, 0, c->stream>>>(var_r, ldv, nullptr, LTA, M, N, nullptr, 0, 1.0, 1); c->launches++; }
    } else if (qmode == 2) {
      if (full_cov) {
        scale_rows_kernel<<<grid1d(mn), 256, 0, c->stream>>>(LTA, Ause, M, N, tq.d + r, R); c->launches++;   // LTA = A * q[:, r] (:51-52)
        launch_gemm_small(c, var_r, N, LTA, N, 1, LTA, N, 0, N, N, M, 1.0, 1.0);
      } else {
        colsumsq_kernel<<<grid1d(N), 256, 0, c->stream>>>(var_r, ldv, nullptr, Ause, M, N, tq.d + r, R, 1.0, 1); c->launches++;
      }
    }
  }
  if (tL.present) {
    const double* Lfac = (const double*)(c->arena + L.off_Lfac);
    for (int k = 0; k < nk; ++k) {
      tril_copy_kernel<<<grid1d((size_t)M * M), 256, 0, c->stream>>>(tL.d + (size_t)k * M * M, Lfac + (size_t)k * Mp * Mp, M, Mp); c->launches++;
    }
  }
  for (double* p : {Kmn, A, A2, LTA, Knn}) CUDA_TRY(cudaFreeAsync(p, c->stream));
  CUDA_TRY(cudaGetLastError());
  int st = check_status(c, L);
  TRY(call.finish());
  return st;
}

extern "C" int ffvd_logdensity_norm_diag(ffvd_ctx* c, DLManagedTensor* y, DLManagedTensor* ymean, DLManagedTensor* Rchols,
                                         int vec, DLManagedTensor* out) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  CUDA_TRY(cudaSetDevice(c->device));
  Call call(c);
  Tens ty, tm, tr, to;
  TRY(call.import(y, false, ty, "y"));
  TRY(call.import(ymean, false, tm, "ymean"));
  TRY(call.import(Rchols, false, tr, "Rchols"));
  TRY(call.import(out, true, to, "out"));
  if (ty.ndim != 2 || tm.numel != ty.numel) return fail(FFVD_E_SHAPE, "y, ymean must be (N,Dy)");
  const int N = (int)ty.shape[0], Dy = (int)ty.shape[1];
  if (tr.numel != (size_t)Dy) return fail(FFVD_E_SHAPE, "Rchols must be (Dy)");
  if (to.numel != (vec ? (size_t)N : (size_t)N * Dy)) return fail(FFVD_E_SHAPE, "out has the wrong size");
  if (N) { logdensity_diag_kernel<<<grid1d(N), 256, 0, c->stream>>>(ty.d, tm.d, tr.d, N, Dy, vec, to.d); c->launches++; }
  return call.finish();
}

// Diagnostic: the fused kernels' branch-free exp routine applied elementwise (accuracy tests against libm).
__global__ void debug_exp_kernel(const double* __restrict__ x, double* __restrict__ out, size_t n) {
  __shared__ double tab[64];
  if (threadIdx.x < 64) tab[threadIdx.x] = g_exp2_tab[threadIdx.x];
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double a[1] = {x[i]};
    exp_nonpos_n<1>(a, tab);
    out[i] = a[0];
  }
}
extern "C" int ffvd_debug_exp(ffvd_ctx* c, DLManagedTensor* x, DLManagedTensor* out) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  CUDA_TRY(cudaSetDevice(c->device));
  Call call(c);
  Tens tx, to;
  TRY(call.import(x, false, tx, "x"));
  TRY(call.import(out, true, to, "out"));
  if (tx.numel != to.numel) return fail(FFVD_E_SHAPE, "x and out must have the same size");
  if (tx.numel) { debug_exp_kernel<<<grid1d(tx.numel), 256, 0, c->stream>>>(tx.d, to.d, tx.numel); c->launches++; }
  return call.finish();
}

extern "C" int ffvd_logdensity_norm(ffvd_ctx* c, DLManagedTensor* y, DLManagedTensor* ymean, DLManagedTensor* Rchols,
                                    DLManagedTensor* out) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  CUDA_TRY(cudaSetDevice(c->device));
  Call call(c);
  Tens ty, tm, tr, to;
  TRY(call.import(y, false, ty, "y"));
  TRY(call.import(ymean, false, tm, "ymean"));
  TRY(call.import(Rchols, false, tr, "Rchols"));
  TRY(call.import(out, true, to, "out"));
  if (tm.ndim != 2) return fail(FFVD_E_SHAPE, "ymean must be (N,Dy)");
  const int N = (int)tm.shape[0], Dy = (int)tm.shape[1];
  if (Dy < 1 || Dy > FFVD_MAX_DY) return fail(FFVD_E_LIMIT, "logdensity_norm: Dy must be in 1..64");
  if (ty.numel != (size_t)N * Dy && ty.numel != (size_t)Dy) return fail(FFVD_E_SHAPE, "y must be (N,Dy), or one row (Dy) broadcast over ymean");
  if (tr.numel != (size_t)Dy * Dy) return fail(FFVD_E_SHAPE, "Rchols must be (Dy,Dy)");
  if (to.numel != (size_t)N) return fail(FFVD_E_SHAPE, "out must be (N)");
  if (N) {
    logdensity_full_kernel<<<grid1d(N, 128), 128, 0, c->stream>>>(ty.d, ty.numel == (size_t)Dy && N != 1 ? 1 : N, tm.d, tr.d, N, Dy, to.d);
    c->launches++;
  }
  return call.finish();
}

extern "C" int ffvd_sghmc_update(ffvd_ctx* c, DLManagedTensor* theta, DLManagedTensor* grad, DLManagedTensor* noise,
                                 DLManagedTensor* xi, DLManagedTensor* g, DLManagedTensor* g2, DLManagedTensor* p,
                                 double epsilon, double mdecay, double X_N, int burn_in) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  CUDA_TRY(cudaSetDevice(c->device));
  Call call(c);
  Tens tt, tg, tn, txi, tgg, tg2, tp;
  // in-place tensors are imported as inputs (copied in) and flagged as outputs (copied back)
  TRY(call.import(theta, false, tt, "theta"));
  TRY(call.import(grad, false, tg, "grad"));
  TRY(call.import(noise, false, tn, "noise"));
  TRY(call.import(xi, false, txi, "xi"));
  TRY(call.import(g, false, tgg, "g"));
  TRY(call.import(g2, false, tg2, "g2"));
  TRY(call.import(p, false, tp, "p"));
  const size_t n = tt.numel;
  if (tg.numel != n || tn.numel != n || txi.numel != n || tgg.numel != n || tg2.numel != n || tp.numel != n)
    return fail(FFVD_E_SHAPE, "all SG-HMC tensors must have the same number of elements");
  call.mark_inout(tt); call.mark_inout(tp);
  if (burn_in) { call.mark_inout(txi); call.mark_inout(tgg); call.mark_inout(tg2); }
  if (n) {
    const double eps_scaled = epsilon / sqrt(X_N);
    const int grid = grid1d(n, 256, c->num_sms * 8);
    const uintptr_t al = (uintptr_t)tt.d | (uintptr_t)tg.d | (uintptr_t)tn.d | (uintptr_t)txi.d | (uintptr_t)tgg.d | (uintptr_t)tg2.d | (uintptr_t)tp.d;
    const bool vec = (al & 15) == 0;
#define FFVD_SGHMC(B, V) sghmc_kernel<B, V><<<grid, 256, 0, c->stream>>>(tt.d, tg.d, tn.d, txi.d, tgg.d, tg2.d, tp.d, n, epsilon, mdecay, eps_scaled)
    if (burn_in) { if (vec) FFVD_SGHMC(1, 1); else FFVD_SGHMC(1, 0); }
    else { if (vec) FFVD_SGHMC(0, 1); else FFVD_SGHMC(0, 0); }
#undef FFVD_SGHMC
    c->launches++;
  }
  return call.finish();
}

extern "C" int ffvd_adam_update(ffvd_ctx* c, DLManagedTensor* theta, DLManagedTensor* grad, DLManagedTensor* m,
                                DLManagedTensor* v, double lr, double beta1, double beta2, double eps, int64_t step) {
  if (!c) return fail(FFVD_E_BADARG, "ctx is null");
  if (step < 1) return fail(FFVD_E_BADARG, "step is 1-based");
  CUDA_TRY(cudaSetDevice(c->device));
  Call call(c);
  Tens tt, tg, tm, tv;
  TRY(call.import(theta, false, tt, "theta"));
  TRY(call.import(grad, false, tg, "grad"));
  TRY(call.import(m, false, tm, "m"));
  TRY(call.import(v, false, tv, "v"));
  const size_t n = tt.numel;
  if (tg.numel != n || tm.numel != n || tv.numel != n) return fail(FFVD_E_SHAPE, "all Adam tensors must have the same size");
  call.mark_inout(tt); call.mark_inout(tm); call.mark_inout(tv);
  if (n) {
    const double lr_t = lr * sqrt(1.0 - pow(beta2, (double)step)) / (1.0 - pow(beta1, (double)step));
    const bool vec = ((((uintptr_t)tt.d | (uintptr_t)tg.d | (uintptr_t)tm.d | (uintptr_t)tv.d)) & 15) == 0;
    if (vec) adam_kernel<1><<<grid1d(n, 256, c->num_sms * 8), 256, 0, c->stream>>>(tt.d, tg.d, tm.d, tv.d, n, lr_t, beta1, beta2, eps);
    else adam_kernel<0><<<grid1d(n, 256, c->num_sms * 8), 256, 0, c->stream>>>(tt.d, tg.d, tm.d, tv.d, n, lr_t, beta1, beta2, eps);
    c->launches++;
  }
  return call.finish();
}
