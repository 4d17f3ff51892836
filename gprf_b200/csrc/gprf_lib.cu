// Host side of libgprf_b200.so: device context, per-evaluation launch plan and
// the C-ABI declared in include/gprf_b200.h.  No torch types cross this
// boundary; PyTorch (when used at all) only lends device pointers and a stream.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/gprf_b200.h"
#include "gprf_kernels.cuh"
#include "partition.cuh"
#include "resident.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <queue>
#include <mutex>

using namespace gprf;

#define CUDA_OK(call)                                                                     \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      char buf_[512];                                                                     \
      snprintf(buf_, sizeof buf_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      h->err = buf_;                                                                      \
      return GPRF_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

struct gprf_ctx {
  int device = 0;
  long long n = 0;
  int dx = 0, dy = 0, yr = 0, nya = 0, dfn = 0, wfn = 0, nls = 0;
  std::string err;

  cudaStream_t stream = nullptr;   // used by the host-buffer entry
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_ms = 0.f;
  int last_launches = 0;

  double* dY = nullptr;
  double* dX = nullptr;            // staging for host-buffer entry
  double* dOut = nullptr;          // [ll, gth(5), gradX(n*dx)]
  double* hX = nullptr;            // pinned
  double* hOut = nullptr;          // pinned

  // structure
  bool have_structure = false;
  int B = 0, E = 0, U = 0;
  long long plen = 0;
  std::vector<UnitDesc> units;
  std::vector<int> all_list;       // active, non-empty units, largest first
  int ntmax = 0;
  UnitDesc* dUnits = nullptr;
  int* dList = nullptr;      // scratch list (jitter retries)
  int* dListAll = nullptr;   // all active, non-empty units
  long long* dPerm = nullptr;
  long long* dBlockPtr = nullptr;
  int* dPosBlock = nullptr;
  int *dAdjPtr = nullptr, *dAdjEdge = nullptr, *dAdjSide = nullptr;
  size_t capUnits = 0, capPerm = 0, capPos = 0, capB = 0, capAdj = 0, capAdj2 = 0, capAdjPtr = 0;
  UnitDesc* hUnits = nullptr;      // pinned staging
  int* hList = nullptr;
  cudaEvent_t evStage = nullptr;
  std::vector<int> edges, deg;
  std::vector<long long> block_ptr_h;
  std::vector<unsigned char> explicit_mask, lpt, seen;
  bool use_explicit_mask = false, adj_dirty = true, blocks_from_device = false;
  bool raw_weights = false;        // masked evaluations: every active unit counts with weight 1
  // optimiser glue (gprf_set_x_prior / gprf_neg_objective)
  double* dPriorMean = nullptr;    // n x dx
  double* dPriorPartial = nullptr;
  unsigned* dPriorCounter = nullptr;
  double prior_ivar[MAX_DX] = {0, 0, 0}, prior_gscale[MAX_DX] = {1, 1, 1};
  bool have_prior = false, prior_apply = false;
  int shard_rank = 0, shard_world = 1;
  bool keep_kinv = false;          // store K^-1 tiles (gprf_set_keep_kinv)
  bool share_on = true;            // edges reuse block i's factor tiles (gprf_set_factor_reuse)
  bool any_share = false;
  int fused_share_min = -1;        // fused pairs reuse too once there are this many of them (-1: 8 per SM)
  int n_fused_parents = 0;         // all_list = [tiled parents | other tiled | fused parents | other fused]
  int n_tiled_parents = 0;
  int n_share_units = 0;
  long long n_share_tiles = 0;     // potrf/trtri/forward-solve tile tasks not executed thanks to the reuse
  int n_sm = 148;
  int panel_order = 0;             // 0 task-major (default: measured faster), 1 unit-major, -1 by launch size (GPRF_PANEL_ORDER)
  // Units of up to fused_nt 64-point tiles take k_unit_fused (one CTA per unit).  Default 0: every
  // unit runs the tile pipeline, which is the faster schedule on B200 for every workload measured
  // (README config 0.853 vs 0.904 ms, 8-rank shard of it 0.469 vs 0.611, cfg4 16.9 vs 18.6 ms).
  int fused_nt = 0;
  int fused_mixed_nt = 0;          // ... but only up to this many when larger units run the tile pipeline anyway
  int fused_eff = 0;               // threshold in force for the current structure (rebuild_units)
  unsigned long long* dTrace = nullptr;   // debug trace of the fused kernel (gprf_debug_trace)
  size_t capTrace = 0;

  // device partitioner (K8)
  int part_kind = 0, part_B = 0, part_mode = 0, part_root = 0, part_launches = 0;
  double part_wrap_add = 22.0, part_wrap_mod = 360.0;
  double *dPartA = nullptr, *dPartB = nullptr, *dPartC = nullptr;
  int* dPartChild = nullptr;
  int *dOwner = nullptr, *dIota = nullptr, *dIdxSorted = nullptr;
  size_t capOwner = 0, capIota = 0, capIdxS = 0, capCub = 0, capHB = 0;
  void* dCub = nullptr;
  long long* hBlockPtr = nullptr;  // pinned

  double* arena = nullptr;
  size_t arena_cap = 0;            // doubles
  // per-evaluation scratch, one allocation so that one memset clears it:
  // [ll per unit | jitter per unit | info per unit | nfail]
  void* dScratch = nullptr;
  size_t scratch_bytes = 0;
  double *dLLu = nullptr, *dGthU = nullptr, *dJitter = nullptr;
  int *dInfo = nullptr, *dNfail = nullptr;
  int* hNfail = nullptr;           // pinned
  bool units_built = false;        // descriptors on the device match block_ptr_h / edges / switches
  std::vector<double> jitter;
  std::vector<int> tries;

  // resident (shared-memory) unit path, resident.cuh
  bool res_enable = true;
  int bucket_split = -1;                               // placement CTAs of the split bucketing launch (GPRF_BUCKET_SPLIT; 0: one CTA, -1: one per 4096 points, at most 8)
  bool res_defer = true, res_sort_blocks = true;       // GPRF_RES_DEFER / GPRF_RES_SORTBLK
  bool res_early = true;                               // pairs start once the parent's W is exported (GPRF_RES_EARLY=0: wait for all factor exports)
  long long res_spin_limit = res::SPIN_LIMIT_CYCLES;   // watchdog of the resident kernel's spin waits (cycles)
  bool dev_blocks_valid = false;   // dPerm / dPosBlock / dBlockPtr describe the current blocks
  bool host_blocks_stale = false;  // ... and are newer than block_ptr_h / the unit descriptors
  cudaStream_t blocks_stream = nullptr;   // stream the device-held blocks were last written on
  bool res_static_dirty = true;    // edges / degrees / unit mask have to be uploaded again
  bool last_resident = false;      // the last evaluation ran on the resident path
  bool plan_fused = false;         // the launch plan of the next resident evaluation is already on the device
  int pending_part_launches = 0;   // re-blocking launches of gprf_reblock_device, counted into the next evaluation
  bool bucket_small = true;        // single-CTA bucketing for small problems (GPRF_BUCKET_SMALL=0: cub radix sort)
  long long res_evals = 0, res_fallbacks = 0;
  int res_last_status = 0;
  double* dResExports = nullptr;
  double* dResScratch = nullptr;
  double *dResLL = nullptr, *dResGth = nullptr, *dResGx = nullptr;
  int *dResInfo = nullptr, *dResOrderB = nullptr, *dResOrderP = nullptr, *dResCounts = nullptr, *dResReady = nullptr;
  int res_epoch = 0;
  int *dResEdges = nullptr, *dResDeg = nullptr;
  unsigned char* dResActive = nullptr;
  bool res_have_mask = false;
  size_t capResExp = 0, capResScr = 0, capResU = 0, capResB = 0, capResE = 0, capResAct = 0;
  int* hResStatus = nullptr;       // pinned: [status]
  double* dResDbg = nullptr;
  int* dResListPtr = nullptr;      // n_sm + 1 bounds of the CTAs' unit lists
  int res_dbg_unit = -1, res_dbg_phase = -1;
  std::vector<unsigned char> res_mask;

  // optional per-kernel-family timing (CUDA events around every launch)
  bool profile = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Rec { int fam; size_t a, b; };
  std::vector<Rec> recs;
  float fam_ms[GPRF_N_FAMILIES] = {0};
  int fam_launches[GPRF_N_FAMILIES] = {0};

  size_t prof_begin(int fam, cudaStream_t st) {
    if (!profile) return 0;
    while (ev_pool.size() < ev_used + 2) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev_pool.push_back(e);
    }
    size_t a = ev_used;
    ev_used += 2;
    cudaEventRecord(ev_pool[a], st);
    recs.push_back({fam, a, a + 1});
    return a;
  }
  void prof_end(size_t a, cudaStream_t st) {
    if (profile) cudaEventRecord(ev_pool[a + 1], st);
  }
  void prof_resolve() {
    for (int f = 0; f < GPRF_N_FAMILIES; ++f) { fam_ms[f] = 0.f; fam_launches[f] = 0; }
    for (auto& r : recs) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev_pool[r.a], ev_pool[r.b]) == cudaSuccess) fam_ms[r.fam] += ms;
      fam_launches[r.fam]++;
    }
    recs.clear();
    ev_used = 0;
  }
};

#define LAUNCH(fam, ...)                          \
  do {                                            \
    size_t pa_ = h->prof_begin(fam, st);          \
    __VA_ARGS__;                                  \
    h->prof_end(pa_, st);                         \
    ++launches;                                   \
  } while (0)

static size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

template <class Tp>
static cudaError_t ensure(Tp** p, size_t* cap, size_t need) {
  if (need <= *cap && *p) return cudaSuccess;
  if (*p) cudaFree(*p);
  *p = nullptr;
  size_t c = std::max<size_t>(need + need / 4, 16);
  cudaError_t e = cudaMalloc((void**)p, c * sizeof(Tp));
  *cap = (e == cudaSuccess) ? c : 0;
  return e;
}

static int make_cov(const gprf_ctx* h, const double* theta, int ncov, CovParams* cp) {
  if (ncov != 2 + h->nls) return GPRF_ERR_ARG;
  cp->nv = theta[0];
  cp->s2 = theta[1];
  cp->dx = h->dx;
  cp->nls = h->nls;
  cp->dfn = h->dfn;
  for (int t = 0; t < MAX_NLS; ++t) {
    if (t < h->nls) {
      double l = theta[2 + t];
      cp->il2[t] = 1.0 / (l * l);
      cp->il3[t] = 1.0 / (l * l * l);
    } else {
      cp->il2[t] = 0.0;
      cp->il3[t] = 0.0;
    }
  }
  return GPRF_OK;
}

static const size_t PIPE_BYTES = PIPE_ALLOC_DOUBLES * sizeof(double);

// k_unit_fused is instantiated in its own translation units (gprf_fused.cu, one per covariance
// family, compiled in parallel by build.sh); these are their host-side launchers.
namespace gprf {
template <int DFN, int WFN> void fused_set_attr();
template <int DFN, int WFN>
void fused_launch(const EvalParams& P, double* ll_u, double* gth_u, int want_grad, int nunits, cudaStream_t st);
}

namespace gprf {
namespace res {
template <int DFN, int WFN> int resident_set_attr();
template <int DFN, int WFN> void resident_launch(const ResParams& P, int grid, cudaStream_t st);
}
}

template <int DFN, int WFN>
static void set_attrs_t() {
  res::resident_set_attr<DFN, WFN>();
  cudaFuncSetAttribute(k_potrf_diag<DFN, WFN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_BYTES);
  cudaFuncSetAttribute(k_potrf_panel<DFN, WFN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_BYTES);
  cudaFuncSetAttribute(k_grad<DFN, WFN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_BYTES);
  fused_set_attr<DFN, WFN>();
}
static void set_attrs() {
  set_attrs_t<0, 0>();
  set_attrs_t<0, 1>();
  set_attrs_t<1, 0>();
  set_attrs_t<1, 1>();
  cudaFuncSetAttribute(k_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_BYTES);
  cudaFuncSetAttribute(k_trtri, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_BYTES);
  cudaFuncSetAttribute(k_alpha, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_BYTES);
}

#define DISPATCH_COV(h, CALL)                                  \
  do {                                                         \
    if ((h)->dfn == 0 && (h)->wfn == 0) { CALL(0, 0); }        \
    else if ((h)->dfn == 0 && (h)->wfn == 1) { CALL(0, 1); }   \
    else if ((h)->dfn == 1 && (h)->wfn == 0) { CALL(1, 0); }   \
    else { CALL(1, 1); }                                       \
  } while (0)

// ---------------------------------------------------------------------------
extern "C" int gprf_abi_version(void) { return 1; }

extern "C" const char* gprf_strerror(int code) {
  switch (code) {
    case GPRF_OK: return "ok";
    case GPRF_ERR_NOT_PD: return "not positive definite, even with jitter.";
    case GPRF_ERR_NONPOS_DIAG: return "not pd: non-positive diagonal elements";
    case GPRF_ERR_ARG: return "invalid argument";
    case GPRF_ERR_CUDA: return "CUDA error";
    case GPRF_ERR_NO_STRUCTURE: return "gprf_set_structure has not been called";
  }
  return "unknown";
}

extern "C" const char* gprf_last_error(gprf_handle h) { return h ? h->err.c_str() : ""; }

extern "C" int gprf_create(gprf_handle* out, int device, long long n, int dx, int dy, const double* Y,
                           int dfn_id, int wfn_id) {
  if (!out || n <= 0 || dx < 1 || dx > MAX_DX || dy < 1 || !Y) return GPRF_ERR_ARG;
  if (dfn_id < 0 || dfn_id > 1 || wfn_id < 0 || wfn_id > 1) return GPRF_ERR_ARG;
  if (dfn_id == GPRF_DFN_LLD && dx != 3) return GPRF_ERR_ARG;
  gprf_ctx* h = new gprf_ctx();
  *out = h;
  h->device = device;
  h->n = n;
  h->dx = dx;
  h->dy = dy;
  h->yr = ((dy + T - 1) / T) * T;
  h->nya = h->yr / T;
  h->dfn = dfn_id;
  h->wfn = wfn_id;
  h->nls = (dfn_id == GPRF_DFN_LLD) ? 2 : dx;
  CUDA_OK(cudaSetDevice(device));
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreate(&h->ev0));
  CUDA_OK(cudaEventCreate(&h->ev1));
  CUDA_OK(cudaEventCreateWithFlags(&h->evStage, cudaEventDisableTiming));
  const size_t outlen = 1 + MAX_NCOV + (size_t)n * dx;
  CUDA_OK(cudaMalloc((void**)&h->dY, (size_t)n * dy * sizeof(double)));
  CUDA_OK(cudaMalloc((void**)&h->dX, (size_t)n * dx * sizeof(double)));
  CUDA_OK(cudaMalloc((void**)&h->dOut, outlen * sizeof(double)));
  CUDA_OK(cudaMallocHost((void**)&h->hX, (size_t)n * dx * sizeof(double)));
  CUDA_OK(cudaMallocHost((void**)&h->hOut, outlen * sizeof(double)));
  CUDA_OK(cudaMallocHost((void**)&h->hNfail, sizeof(int)));
  CUDA_OK(cudaMemcpy(h->dY, Y, (size_t)n * dy * sizeof(double), cudaMemcpyHostToDevice));
  set_attrs();
  if (const char* e = getenv("GPRF_FUSED_NT")) {
    h->fused_nt = atoi(e);
    h->fused_mixed_nt = std::min(h->fused_nt, 4);
  }
  if (const char* e = getenv("GPRF_PANEL_ORDER")) h->panel_order = atoi(e);
  if (const char* e = getenv("GPRF_RESIDENT")) h->res_enable = atoi(e) != 0;
  if (const char* e = getenv("GPRF_RES_EARLY")) h->res_early = atoi(e) != 0;
  if (const char* e = getenv("GPRF_BUCKET_SPLIT")) h->bucket_split = std::max(-1, std::min(8, atoi(e)));
  if (const char* e = getenv("GPRF_RES_DEFER")) h->res_defer = atoi(e) != 0;
  if (const char* e = getenv("GPRF_RES_SORTBLK")) h->res_sort_blocks = atoi(e) != 0;
  if (const char* e = getenv("GPRF_RES_WATCHDOG_S")) {
    const double sec = atof(e);
    if (sec > 0) h->res_spin_limit = (long long)(sec * 2.0e9);
  }
  if (const char* e = getenv("GPRF_BUCKET_SMALL")) h->bucket_small = atoi(e) != 0;
  CUDA_OK(cudaMallocHost((void**)&h->hResStatus, 4 * sizeof(int)));
  if (const char* e = getenv("GPRF_FUSED_SHARE_MIN")) h->fused_share_min = atoi(e);
  cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, device);
  if (const char* e = getenv("GPRF_FUSED_MIXED_NT")) h->fused_mixed_nt = atoi(e);
  h->fused_mixed_nt = std::min(h->fused_mixed_nt, h->fused_nt);
  CUDA_OK(cudaGetLastError());
  return GPRF_OK;
}

// Debug: enable (n_ctas > 0) / read back the per-phase timestamps of the fused kernel.
extern "C" int gprf_debug_trace(gprf_handle h, int n_ctas, unsigned long long* out) {
  if (!h) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  if (out && h->dTrace) {
    CUDA_OK(cudaMemcpy(out, h->dTrace,
                       std::min((size_t)n_ctas * 2 * TRACE_SLOTS, h->capTrace) * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost));
    return GPRF_OK;
  }
  if (h->dTrace) cudaFree(h->dTrace);
  h->dTrace = nullptr;
  h->capTrace = 0;
  if (n_ctas > 0) {
    const size_t len = (size_t)n_ctas * 2 * TRACE_SLOTS;
    CUDA_OK(cudaMalloc((void**)&h->dTrace, len * sizeof(unsigned long long)));
    CUDA_OK(cudaMemset(h->dTrace, 0, len * sizeof(unsigned long long)));
    h->capTrace = len;
  }
  return GPRF_OK;
}

static int rebuild_units(gprf_ctx* h, cudaStream_t st);
// The factor-reuse plan depends on these switches: rebuild the unit descriptors.
static int replan(gprf_ctx* h) {
  if (!h->have_structure) return GPRF_OK;
  CUDA_OK(cudaSetDevice(h->device));
  int rc = rebuild_units(h, h->stream);
  if (rc != GPRF_OK) return rc;
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return GPRF_OK;
}

extern "C" int gprf_set_keep_kinv(gprf_handle h, int on) {
  if (!h) return GPRF_ERR_ARG;
  h->keep_kinv = on != 0;
  return replan(h);
}

extern "C" int gprf_set_fused_nt(gprf_handle h, int nt) {
  if (!h || nt < 0) return GPRF_ERR_ARG;
  h->fused_nt = nt;
  h->fused_mixed_nt = std::min(nt, 4);
  return replan(h);
}

extern "C" int gprf_set_factor_reuse(gprf_handle h, int on) {
  if (!h) return GPRF_ERR_ARG;
  h->share_on = on != 0;
  if (on == 2) h->fused_share_min = 1;        // fused pairs too, however few
  else if (on == 1 && h->fused_share_min == 1) h->fused_share_min = -1;
  return replan(h);
}

extern "C" int gprf_factor_reuse_stats(gprf_handle h, int* n_units, long long* n_tiles) {
  if (!h) return GPRF_ERR_ARG;
  if (n_units) *n_units = h->n_share_units;
  if (n_tiles) *n_tiles = h->n_share_tiles;
  return GPRF_OK;
}

extern "C" int gprf_destroy(gprf_handle h) {
  if (!h) return GPRF_OK;
  cudaSetDevice(h->device);
  cudaFree(h->dY); cudaFree(h->dX); cudaFree(h->dOut);
  cudaFreeHost(h->hX); cudaFreeHost(h->hOut);
  cudaFree(h->dUnits); cudaFree(h->dList); cudaFree(h->dListAll); cudaFree(h->dPerm); cudaFree(h->dBlockPtr);
  cudaFree(h->dPosBlock); cudaFree(h->dAdjPtr); cudaFree(h->dAdjEdge); cudaFree(h->dAdjSide);
  cudaFree(h->arena); cudaFree(h->dScratch); cudaFree(h->dGthU);
  if (h->hNfail) cudaFreeHost(h->hNfail);
  cudaFree(h->dPartA); cudaFree(h->dPartB); cudaFree(h->dPartC); cudaFree(h->dPartChild);
  cudaFree(h->dOwner); cudaFree(h->dIota); cudaFree(h->dIdxSorted); cudaFree(h->dCub);
  cudaFree(h->dResExports); cudaFree(h->dResScratch); cudaFree(h->dResLL); cudaFree(h->dResGth); cudaFree(h->dResGx);
  cudaFree(h->dResInfo); cudaFree(h->dResOrderB); cudaFree(h->dResOrderP); cudaFree(h->dResCounts); cudaFree(h->dResReady);
  cudaFree(h->dResEdges); cudaFree(h->dResDeg); cudaFree(h->dResActive); cudaFree(h->dResDbg); cudaFree(h->dResListPtr);
  cudaFree(h->dPriorMean); cudaFree(h->dPriorPartial); cudaFree(h->dPriorCounter);
  if (h->hResStatus) cudaFreeHost(h->hResStatus);
  if (h->hUnits) cudaFreeHost(h->hUnits);
  if (h->hList) cudaFreeHost(h->hList);
  if (h->hBlockPtr) cudaFreeHost(h->hBlockPtr);
  if (h->evStage) cudaEventDestroy(h->evStage);
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return GPRF_OK;
}

// ---------------------------------------------------------------------------
// structure = edges (static) + blocks (per update_X) -> units
// ---------------------------------------------------------------------------
static int validate_edges(gprf_ctx* h, int B, int E, const int* edges) {
  for (int e = 0; e < E; ++e) {
    int i = edges[2 * e], j = edges[2 * e + 1];
    if (i < 0 || i >= B || j < 0 || j >= B || i == j) {
      h->err = "edge endpoints must be distinct block ids";
      return GPRF_ERR_ARG;
    }
  }
  return GPRF_OK;
}

// Adjacency (CSR of incident edges per block, in edge order = fixed summation order) and degrees.
static int rebuild_adjacency(gprf_ctx* h) {
  const int B = h->B, E = h->E;
  const int* edges = h->edges.data();
  int rc = validate_edges(h, B, E, edges);
  if (rc != GPRF_OK) return rc;
  h->deg.assign(B, 0);
  for (int e = 0; e < E; ++e) {
    h->deg[edges[2 * e]]++;
    h->deg[edges[2 * e + 1]]++;
  }
  std::vector<int> adj_ptr(B + 1, 0), adj_edge(2 * (size_t)E + 1), adj_side(2 * (size_t)E + 1);
  for (int b = 0; b < B; ++b) adj_ptr[b + 1] = adj_ptr[b] + h->deg[b];
  {
    std::vector<int> fill(adj_ptr.begin(), adj_ptr.end() - 1);
    for (int e = 0; e < E; ++e) {
      int i = edges[2 * e], j = edges[2 * e + 1];
      adj_edge[fill[i]] = e; adj_side[fill[i]++] = 0;
      adj_edge[fill[j]] = e; adj_side[fill[j]++] = 1;
    }
  }
  CUDA_OK(ensure(&h->dAdjPtr, &h->capAdjPtr, (size_t)B + 1));
  CUDA_OK(ensure(&h->dAdjEdge, &h->capAdj, (size_t)2 * E + 1));
  CUDA_OK(ensure(&h->dAdjSide, &h->capAdj2, (size_t)2 * E + 1));
  CUDA_OK(cudaMemcpy(h->dAdjPtr, adj_ptr.data(), (size_t)(B + 1) * sizeof(int), cudaMemcpyHostToDevice));
  if (E > 0) {
    CUDA_OK(cudaMemcpy(h->dAdjEdge, adj_edge.data(), (size_t)2 * E * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(h->dAdjSide, adj_side.data(), (size_t)2 * E * sizeof(int), cudaMemcpyHostToDevice));
  }
  h->adj_dirty = false;
  h->res_static_dirty = true;
  return GPRF_OK;
}

// Multi-GPU sharding of the units over `world` ranks - the same rule as
// gprf_b200/dist.py::shard_units.  The unit of assignment is a GROUP: block i together with
// every edge (i, j) whose rows start with block i (gprf.py:310-330), so that a pair always
// finds its parent's factor on its own rank (UnitDesc::share).  Groups are placed by a
// longest-processing-time greedy on the work model W(s) = s^3 + 4 s^2 * 50.
static void lpt_mask(const std::vector<double>& sizes, int B, const int* edges, int rank, int world,
                     std::vector<unsigned char>& mask) {
  const size_t U = sizes.size();
  mask.assign(U, 1);
  if (world <= 1) return;
  auto W = [](double s) { return s * s * s + 4.0 * 50.0 * s * s; };
  std::vector<double> cost(B);
  for (int b = 0; b < B; ++b) cost[b] = W(sizes[b]);
  for (size_t u = B; u < U; ++u) cost[edges[2 * (u - B)]] += W(sizes[u]);
  std::vector<int> order(B);
  for (int b = 0; b < B; ++b) order[b] = b;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  typedef std::pair<double, int> LR;
  std::priority_queue<LR, std::vector<LR>, std::greater<LR>> heap;
  for (int r = 0; r < world; ++r) heap.push(LR(0.0, r));
  std::vector<int> owner(B);
  for (int b : order) {
    LR t = heap.top();
    heap.pop();
    owner[b] = t.second;
    heap.push(LR(t.first + cost[b], t.second));
  }
  for (int b = 0; b < B; ++b) mask[b] = owner[b] == rank ? 1 : 0;
  for (size_t u = B; u < U; ++u) mask[u] = owner[edges[2 * (u - B)]] == rank ? 1 : 0;
}

// The resident path (resident.cuh) is tried when the structure is of the small-unit kind: blocks of
// ~100 points (at most 128 each, checked on the device per evaluation).  The test is a function of
// static quantities only, so that every rank of a sharded job takes the same decision.
static bool res_eligible(const gprf_ctx* h) {
  if (!h->res_enable || h->keep_kinv || h->dy > 8 * res::RNYB || h->B < 1) return false;
  if ((size_t)(h->B + h->E) * sizeof(int) > 160 * 1024) return false;     // plan kernel's shared arrays
  return (double)h->n / (double)h->B <= 112.0;
}

// Unit sizes the multi-GPU split works on.  With the resident path the split must not depend on the
// per-evaluation block sizes (they never reach the host): nominal sizes, every block alike.
static void shard_sizes(const gprf_ctx* h, std::vector<double>& sizes) {
  const int B = h->B, E = h->E;
  sizes.assign((size_t)B + E, 0.0);
  if (res_eligible(h)) {
    for (int b = 0; b < B; ++b) sizes[b] = 100.0;
    for (int e = 0; e < E; ++e) sizes[B + e] = 200.0;
    return;
  }
  const long long* block_ptr = h->block_ptr_h.data();
  for (int b = 0; b < B; ++b) sizes[b] = (double)(block_ptr[b + 1] - block_ptr[b]);
  for (int e = 0; e < E; ++e) sizes[B + e] = sizes[h->edges[2 * e]] + sizes[h->edges[2 * e + 1]];
}

// Units from h->block_ptr_h + edges (+ mask / shard).  Uploads descriptors on `st` from pinned staging.
static int rebuild_units(gprf_ctx* h, cudaStream_t st) {
  const int B = h->B;
  if ((int)h->block_ptr_h.size() != B + 1) return GPRF_ERR_NO_STRUCTURE;
  if (h->adj_dirty) {
    int rc = rebuild_adjacency(h);
    if (rc != GPRF_OK) return rc;
  }
  const int E = h->E, U = B + E;
  h->U = U;
  const long long* block_ptr = h->block_ptr_h.data();
  const int* edges = h->edges.data();
  const unsigned char* mask = nullptr;
  if (h->use_explicit_mask) {
    if ((int)h->explicit_mask.size() != U) return GPRF_ERR_ARG;
    mask = h->explicit_mask.data();
  } else if (h->shard_world > 1) {
    std::vector<double> sizes;
    shard_sizes(h, sizes);
    lpt_mask(sizes, B, edges, h->shard_rank, h->shard_world, h->lpt);
    mask = h->lpt.data();
  }
  h->units.assign(U, UnitDesc());
  size_t off = 0;
  h->ntmax = 0;
  h->all_list.clear();
  for (int uix = 0; uix < U; ++uix) {
    UnitDesc& u = h->units[uix];
    int bi, bj = -1;
    if (uix < B) {
      bi = uix;
      u.weight = h->raw_weights ? 1.0 : 1.0 - (double)h->deg[bi];
    } else {
      bi = edges[2 * (uix - B)];
      bj = edges[2 * (uix - B) + 1];
      u.weight = 1.0;
    }
    u.ni = (int)(block_ptr[bi + 1] - block_ptr[bi]);
    u.a_start = (int)block_ptr[bi];
    int nj = 0;
    u.b_start = 0;
    if (bj >= 0) {
      nj = (int)(block_ptr[bj + 1] - block_ptr[bj]);
      u.b_start = (int)block_ptr[bj];
    }
    u.s = u.ni + nj;
    u.nt = (u.s + T - 1) / T;
    u.sp = u.nt * T;
    u.active = (!mask || mask[uix]) ? 1 : 0;
    u.share = 0;
    u.pstore = u.pshare = 0;
    u.p_sp = u.p_nt = 0;
    u.p_m_off = u.p_d_off = u.p_ld_off = u.p_k_off = u.p_kp_off = u.p_ap_off = 0;
    u.kp_off = u.ap_off = 0;
    if (u.active && u.s > 0) {
      const size_t sp = u.sp, nt = u.nt;
      u.m_off = off;   off += align16((sp + h->yr) * sp);
      u.d_off = off;   off += align16(2 * nt * T * T);
      u.al_off = off;  off += align16(sp * h->yr);
      u.xs_off = off;  off += align16(sp * XD);
      u.part_off = off; off += align16((nt * (nt + 1) / 2) * PART_STRIDE);
      u.ld_off = off;  off += align16(nt * (1 + h->nya));   // logdet partials, then ||Z||^2 partials
      u.gx_off = off;  off += align16(sp * 3);
      u.k_off = off;   off += align16(sp * sp);
      if (uix < B && h->share_on && !h->keep_kinv && h->deg[bi] > 0 && u.ni >= T) {
        // room for the partial U-products this block's pairs may start from (UnitDesc::pstore)
        u.kp_off = off;  off += align16(sp * sp);
        u.ap_off = off;  off += align16(sp * h->yr);
      }
      h->all_list.push_back(uix);
      h->ntmax = std::max(h->ntmax, u.nt);
    } else {
      u.m_off = u.d_off = u.al_off = u.xs_off = u.part_off = u.ld_off = u.gx_off = u.k_off = 0;
    }
  }
  std::stable_sort(h->all_list.begin(), h->all_list.end(),
                   [&](int a, int b) { return h->units[a].s > h->units[b].s; });
  // Factor reuse (UnitDesc::share): a pair that takes the tiled schedule reads the tiles that lie
  // entirely inside block i from block i's own unit when that unit is evaluated on this device.
  // A fused parent is complete before the tiled launches start (launch_units); a tiled parent
  // advances level by level in the same launches, one level ahead of every read.
  // A fused unit occupies one CTA for its whole life: a handful of 5-8 tile units next to a tile
  // pipeline that is running anyway is a ~2 ms latency-bound tail (n=200k: the 500-point blocks),
  // while as extra tiles of the pipeline's launches they are nearly free.
  h->fused_eff = (h->ntmax > h->fused_nt) ? std::min(h->fused_nt, h->fused_mixed_nt) : h->fused_nt;
  h->any_share = false;
  h->n_share_units = 0;
  h->n_share_tiles = 0;
  h->n_fused_parents = 0;
  h->n_tiled_parents = 0;
  if (h->share_on && !h->keep_kinv) {
    // Fused pairs can reuse as well, but only behind a launch of their own for the parent blocks;
    // that extra dependent launch pays off once the pairs fill the GPU several times over.
    const int fmin = h->fused_share_min >= 0 ? h->fused_share_min : 8 * h->n_sm;
    auto candidate = [&](int e) {
      const UnitDesc& u = h->units[B + e];
      const UnitDesc& p = h->units[edges[2 * e]];
      return u.active && u.s > 0 && p.active && p.s > 0 && u.ni / T >= 1;
    };
    int n_fused_cand = 0;
    for (int e = 0; e < E; ++e)
      if (candidate(e) && h->units[B + e].nt <= h->fused_eff) ++n_fused_cand;
    const bool fused_share = n_fused_cand >= std::max(fmin, 1);
    std::vector<unsigned char> is_parent(B, 0);
    for (int e = 0; e < E; ++e) {
      UnitDesc& u = h->units[B + e];
      UnitDesc& p = h->units[edges[2 * e]];
      if (!candidate(e)) continue;
      if (u.nt <= h->fused_eff && !fused_share) continue;
      const int m = u.ni / T;
      u.share = m;
      u.p_sp = p.sp;
      u.p_nt = p.nt;
      u.p_m_off = p.m_off;
      u.p_d_off = p.d_off;
      u.p_ld_off = p.ld_off;
      u.p_k_off = p.k_off;
      // The partial U-products need the parents' alpha / K^-1 launches in front of everybody
      // else's: worth it from two shared tiles on (README config, one shared tile of four: the two
      // extra dependent launches cost 35 us of a 0.85 ms step and save 5 % of one family).
      if (m >= 2 && p.kp_off != 0) {
        u.pshare = m;
        u.p_kp_off = p.kp_off;
        u.p_ap_off = p.ap_off;
        p.pstore = m;                             // same m for every pair of this parent
        is_parent[edges[2 * e]] = 2;
      } else if (is_parent[edges[2 * e]] == 0) {
        is_parent[edges[2 * e]] = 1;
      }
      h->any_share = true;
      h->n_share_units++;
      // diag + below-diagonal panel tiles + trtri tiles of the leading square, forward-solve tiles
      h->n_share_tiles += (long long)m * (m + 1) / 2 + (long long)m * (m - 1) / 2 + (long long)m * h->nya;
    }
    // fused parents first among the fused units (the tiled prefix stays in front)
    auto first_fused = std::find_if(h->all_list.begin(), h->all_list.end(),
                                    [&](int uix) { return h->units[uix].nt <= h->fused_eff; });
    auto mid = std::stable_partition(first_fused, h->all_list.end(),
                                     [&](int uix) { return uix < B && is_parent[uix]; });
    h->n_fused_parents = (int)(mid - first_fused);
    // tiled parents: the alpha / K^-1 launches run them first (their pairs start from the partial
    // sums they store); kept at the FRONT of the tiled prefix of the list
    auto tmid = std::stable_partition(h->all_list.begin(), first_fused,
                                      [&](int uix) { return uix < B && is_parent[uix] == 2; });
    h->n_tiled_parents = (int)(tmid - h->all_list.begin());
  }
  if (off > h->arena_cap || !h->arena) {
    CUDA_OK(cudaStreamSynchronize(st));
    size_t acap = h->arena_cap;
    CUDA_OK(ensure(&h->arena, &acap, off));
    h->arena_cap = acap;
  }
  if ((size_t)U > h->capUnits || !h->dUnits) {
    CUDA_OK(cudaStreamSynchronize(st));
    cudaFree(h->dUnits); cudaFree(h->dList); cudaFree(h->dListAll); cudaFree(h->dScratch); cudaFree(h->dGthU);
    if (h->hUnits) cudaFreeHost(h->hUnits);
    if (h->hList) cudaFreeHost(h->hList);
    h->dUnits = nullptr; h->dList = nullptr; h->dListAll = nullptr; h->dScratch = nullptr; h->dGthU = nullptr;
    h->hUnits = nullptr; h->hList = nullptr;
    size_t cu = (size_t)U + U / 4 + 16;
    CUDA_OK(cudaMalloc((void**)&h->dUnits, cu * sizeof(UnitDesc)));
    CUDA_OK(cudaMalloc((void**)&h->dList, cu * sizeof(int)));
    CUDA_OK(cudaMalloc((void**)&h->dListAll, cu * sizeof(int)));
    h->scratch_bytes = cu * (2 * sizeof(double) + sizeof(int)) + 16;
    CUDA_OK(cudaMalloc(&h->dScratch, h->scratch_bytes));
    h->dLLu = (double*)h->dScratch;
    h->dJitter = h->dLLu + cu;
    h->dInfo = (int*)(h->dJitter + cu);
    h->dNfail = h->dInfo + cu;
    CUDA_OK(cudaMalloc((void**)&h->dGthU, cu * MAX_NCOV * sizeof(double)));
    CUDA_OK(cudaMallocHost((void**)&h->hUnits, cu * sizeof(UnitDesc)));
    CUDA_OK(cudaMallocHost((void**)&h->hList, cu * sizeof(int)));
    h->capUnits = cu;
  }
  // pinned staging must not be overwritten while a previous upload is in flight
  CUDA_OK(cudaEventSynchronize(h->evStage));
  memcpy(h->hUnits, h->units.data(), (size_t)U * sizeof(UnitDesc));
  if (!h->all_list.empty()) memcpy(h->hList, h->all_list.data(), h->all_list.size() * sizeof(int));
  CUDA_OK(cudaMemcpyAsync(h->dUnits, h->hUnits, (size_t)U * sizeof(UnitDesc), cudaMemcpyHostToDevice, st));
  if (!h->all_list.empty())
    CUDA_OK(cudaMemcpyAsync(h->dListAll, h->hList, h->all_list.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaEventRecord(h->evStage, st));
  h->jitter.assign(U, 0.0);
  h->tries.assign(U, 0);
  h->have_structure = true;
  h->units_built = true;
  return GPRF_OK;
}

static int store_blocks_host(gprf_ctx* h, int B, const long long* block_ptr, const long long* perm) {
  const long long plen = block_ptr[B];
  if (block_ptr[0] != 0 || plen < 0 || plen > h->n || (plen > 0 && !perm)) {
    h->err = "block_ptr must start at 0 and cover at most n points";
    return GPRF_ERR_ARG;
  }
  for (int b = 0; b < B; ++b)
    if (block_ptr[b + 1] < block_ptr[b]) return GPRF_ERR_ARG;
  h->seen.assign((size_t)h->n, 0);
  for (long long p = 0; p < plen; ++p) {
    if (perm[p] < 0 || perm[p] >= h->n || h->seen[(size_t)perm[p]]) {
      h->err = "block index lists must be disjoint indices in [0, n)";
      return GPRF_ERR_ARG;
    }
    h->seen[(size_t)perm[p]] = 1;
  }
  if (B != h->B) h->adj_dirty = true;
  h->B = B;
  h->plen = plen;
  h->block_ptr_h.assign(block_ptr, block_ptr + B + 1);
  std::vector<int> pos_block((size_t)plen);
  for (int b = 0; b < B; ++b)
    for (long long p = block_ptr[b]; p < block_ptr[b + 1]; ++p) pos_block[(size_t)p] = b;
  CUDA_OK(ensure(&h->dPerm, &h->capPerm, (size_t)std::max<long long>(h->n, 1)));
  CUDA_OK(ensure(&h->dPosBlock, &h->capPos, (size_t)std::max<long long>(h->n, 1)));
  CUDA_OK(ensure(&h->dBlockPtr, &h->capB, (size_t)B + 1));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  if (plen > 0) {
    CUDA_OK(cudaMemcpy(h->dPerm, perm, (size_t)plen * sizeof(long long), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(h->dPosBlock, pos_block.data(), (size_t)plen * sizeof(int), cudaMemcpyHostToDevice));
  }
  CUDA_OK(cudaMemcpy(h->dBlockPtr, block_ptr, (size_t)(B + 1) * sizeof(long long), cudaMemcpyHostToDevice));
  h->blocks_from_device = false;
  h->units_built = false;
  h->dev_blocks_valid = true;
  h->host_blocks_stale = false;
  return GPRF_OK;
}

extern "C" int gprf_set_edges(gprf_handle h, int E, const int* edges, int shard_rank, int shard_world) {
  if (!h || E < 0 || (E > 0 && !edges) || shard_world < 1 || shard_rank < 0 || shard_rank >= shard_world)
    return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  h->edges.assign(edges, edges + 2 * (size_t)E);
  h->E = E;
  h->shard_rank = shard_rank;
  h->shard_world = shard_world;
  h->use_explicit_mask = false;
  h->adj_dirty = true;
  h->have_structure = false;
  h->units_built = false;
  if ((int)h->block_ptr_h.size() == h->B + 1 && h->B > 0) {
    int rc = rebuild_units(h, h->stream);
    if (rc != GPRF_OK) return rc;
    CUDA_OK(cudaStreamSynchronize(h->stream));
  }
  return GPRF_OK;
}

// Evaluate only the units with mask != 0 (NULL: all units / the shard of gprf_set_edges again).
// raw_weights != 0: every active unit counts once (no (1 - deg_i) factors), so that a mask holding a
// single unit returns that unit's own (ll, gradients) - llgrad_unary / llgrad_joint of gprf.py:299-330
// on the live structure.
extern "C" int gprf_set_unit_mask(gprf_handle h, const unsigned char* mask, int n_units, int raw_weights) {
  if (!h) return GPRF_ERR_ARG;
  if (mask && n_units != h->B + h->E) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  h->use_explicit_mask = (mask != nullptr);
  if (mask) h->explicit_mask.assign(mask, mask + n_units);
  h->raw_weights = (mask != nullptr) && raw_weights != 0;
  h->res_static_dirty = true;
  h->have_structure = false;
  h->units_built = false;
  if (!h->host_blocks_stale && (int)h->block_ptr_h.size() == h->B + 1 && h->B > 0) {
    int rc = rebuild_units(h, h->stream);
    if (rc != GPRF_OK) return rc;
    CUDA_OK(cudaStreamSynchronize(h->stream));
  }
  return GPRF_OK;
}

extern "C" int gprf_set_blocks(gprf_handle h, int B, const long long* block_ptr, const long long* perm) {
  if (!h || B < 1 || !block_ptr) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  h->have_structure = false;
  int rc = store_blocks_host(h, B, block_ptr, perm);
  if (rc != GPRF_OK) return rc;
  rc = rebuild_units(h, h->stream);
  if (rc != GPRF_OK) return rc;
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return GPRF_OK;
}

extern "C" int gprf_set_structure(gprf_handle h, int B, const long long* block_ptr, const long long* perm,
                                  int E, const int* edges, const unsigned char* unit_mask) {
  if (!h || B < 1 || !block_ptr || E < 0 || (E > 0 && !edges)) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  h->have_structure = false;
  int rc = validate_edges(h, B, E, edges);
  if (rc != GPRF_OK) return rc;
  h->edges.assign(edges, edges + 2 * (size_t)E);
  h->E = E;
  h->adj_dirty = true;
  h->use_explicit_mask = (unit_mask != nullptr);
  if (unit_mask) h->explicit_mask.assign(unit_mask, unit_mask + B + E);
  else { h->shard_rank = 0; h->shard_world = 1; }
  rc = store_blocks_host(h, B, block_ptr, perm);
  if (rc != GPRF_OK) return rc;
  rc = rebuild_units(h, h->stream);
  if (rc != GPRF_OK) return rc;
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return GPRF_OK;
}

// ---------------------------------------------------------------------------
// device partitioners (K8)
// ---------------------------------------------------------------------------
extern "C" int gprf_set_grid_partitioner(gprf_handle h, int B, const double* centers, const double* csq,
                                         int dot_mode) {
  if (!h || B < 1 || !centers || !csq || dot_mode < 0 || dot_mode > 2) return GPRF_ERR_ARG;
  if ((size_t)B * (h->dx + 1) * sizeof(double) > 200 * 1024) {
    h->err = "too many block centres for the shared-memory assignment kernel";
    return GPRF_ERR_ARG;
  }
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  cudaFree(h->dPartA); cudaFree(h->dPartB);
  h->dPartA = h->dPartB = nullptr;
  CUDA_OK(cudaMalloc((void**)&h->dPartA, (size_t)B * h->dx * sizeof(double)));
  CUDA_OK(cudaMalloc((void**)&h->dPartB, (size_t)B * sizeof(double)));
  CUDA_OK(cudaMemcpy(h->dPartA, centers, (size_t)B * h->dx * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(h->dPartB, csq, (size_t)B * sizeof(double), cudaMemcpyHostToDevice));
  h->part_kind = 1;
  h->part_B = B;
  h->part_mode = dot_mode;
  return GPRF_OK;
}

extern "C" int gprf_set_tree_partitioner(gprf_handle h, int n_nodes, const double* center, const double* direction,
                                         const double* cut, const int* child, int root, int n_leaves,
                                         double wrap_add, double wrap_mod, int dot_mode) {
  if (!h || n_nodes < 0 || n_leaves < 1 || dot_mode < 0 || dot_mode > 2 || h->dx < 2) return GPRF_ERR_ARG;
  if (n_nodes > 0 && (!center || !direction || !cut || !child)) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  cudaFree(h->dPartA); cudaFree(h->dPartB); cudaFree(h->dPartC); cudaFree(h->dPartChild);
  h->dPartA = h->dPartB = h->dPartC = nullptr;
  h->dPartChild = nullptr;
  const size_t nn = (size_t)std::max(n_nodes, 1);
  CUDA_OK(cudaMalloc((void**)&h->dPartA, nn * 2 * sizeof(double)));
  CUDA_OK(cudaMalloc((void**)&h->dPartB, nn * 2 * sizeof(double)));
  CUDA_OK(cudaMalloc((void**)&h->dPartC, nn * sizeof(double)));
  CUDA_OK(cudaMalloc((void**)&h->dPartChild, nn * 2 * sizeof(int)));
  if (n_nodes > 0) {
    CUDA_OK(cudaMemcpy(h->dPartA, center, nn * 2 * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(h->dPartB, direction, nn * 2 * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(h->dPartC, cut, nn * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(h->dPartChild, child, nn * 2 * sizeof(int), cudaMemcpyHostToDevice));
  }
  h->part_kind = 2;
  h->part_B = n_leaves;
  h->part_mode = dot_mode;
  h->part_root = root;
  h->part_wrap_add = wrap_add;
  h->part_wrap_mod = wrap_mod;
  return GPRF_OK;
}

static int res_sync_static(gprf_ctx* h);
static int res_alloc(gprf_ctx* h, int grid);

// f = -(ll + x_prior), g = -(gradX + prior gradient) * grad_scale, in place on out_dev (k_x_prior)
static int apply_prior(gprf_ctx* h, const double* X_dev, double* out_dev, int grad_X, cudaStream_t st) {
  if (!h->prior_apply) return GPRF_OK;
  PriorParams Q;
  Q.X = X_dev;
  Q.mean = h->dPriorMean;
  for (int d = 0; d < MAX_DX; ++d) {
    Q.ivar[d] = h->prior_ivar[d];
    Q.gscale[d] = h->prior_gscale[d];
  }
  Q.n = h->n;
  Q.dx = h->dx;
  Q.want_gx = grad_X ? 1 : 0;
  Q.partial = h->dPriorPartial;
  Q.counter = h->dPriorCounter;
  const long long tot = (long long)h->n * h->dx;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((tot + 255) / 256, 2LL * h->n_sm));
  k_x_prior<<<grid, 256, 0, st>>>(Q, out_dev);
  CUDA_OK(cudaGetLastError());
  return GPRF_OK;
}
static void res_plan_params(gprf_ctx* h, res::PlanParams* Q);

// Block membership of X_dev on the device: assignment kernel, stable radix sort, bounds.
// Leaves perm / pos_block / block_ptr on the device.  Nothing is read back: the resident path never
// needs the block sizes on the host; reblock_finish() fetches them for the tile pipeline.
static int reblock_launch(gprf_ctx* h, const double* X_dev, cudaStream_t st) {
  if (h->part_kind == 0) {
    h->err = "no device partitioner set";
    return GPRF_ERR_ARG;
  }
  const long long n = h->n;
  const int B = h->part_B;
  CUDA_OK(ensure(&h->dPerm, &h->capPerm, (size_t)n));
  CUDA_OK(ensure(&h->dPosBlock, &h->capPos, (size_t)n));
  CUDA_OK(ensure(&h->dBlockPtr, &h->capB, (size_t)B + 1));
  CUDA_OK(ensure(&h->dOwner, &h->capOwner, (size_t)n));
  CUDA_OK(ensure(&h->dIdxSorted, &h->capIdxS, (size_t)n));
  if (!h->dIota || h->capIota < (size_t)n) {
    CUDA_OK(ensure(&h->dIota, &h->capIota, (size_t)n));
    k_iota<<<(unsigned)((h->capIota + 255) / 256), 256, 0, st>>>(h->dIota, (long long)h->capIota);
  }
  int bits = 1;
  while ((1 << bits) < B) ++bits;
  size_t need = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, need, h->dOwner, h->dPosBlock, h->dIota, h->dIdxSorted, (int)n, 0, bits, st);
  if (need > h->capCub || !h->dCub) {
    CUDA_OK(cudaStreamSynchronize(st));
    cudaFree(h->dCub);
    h->dCub = nullptr;
    CUDA_OK(cudaMalloc(&h->dCub, need + 256));
    h->capCub = need + 256;
  }
  // small CTAs: at n = 10^4 a 256-thread CTA per 256 points would light up only 40 of the 148 SMs
  const int tb = n >= 256LL * 4 * h->n_sm ? 256 : 64;
  const unsigned gb = (unsigned)((n + tb - 1) / tb);
  if (h->part_kind == 1) {
    const size_t sm = (size_t)B * (h->dx + 1) * sizeof(double);
#define CALL_ASSIGN(MODE)                                                                                         \
  do {                                                                                                            \
    if (sm > 48 * 1024) cudaFuncSetAttribute(k_assign_grid<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
    k_assign_grid<MODE><<<(unsigned)((n * GT + 127) / 128), 128, sm, st>>>(X_dev, n, h->dx, h->dPartA, h->dPartB, B, h->dOwner); \
  } while (0)
    if (h->part_mode == 0) CALL_ASSIGN(0);
    else if (h->part_mode == 1) CALL_ASSIGN(1);
    else CALL_ASSIGN(2);
  } else {
    TreeParams Tp;
    Tp.center = h->dPartA;
    Tp.direction = h->dPartB;
    Tp.cut = h->dPartC;
    Tp.child = h->dPartChild;
    Tp.root = h->part_root;
    Tp.wrap_add = h->part_wrap_add;
    Tp.wrap_mod = h->part_wrap_mod;
    if (h->part_mode == 0) k_assign_tree<0><<<gb, tb, 0, st>>>(X_dev, n, h->dx, Tp, h->dOwner);
    else if (h->part_mode == 1) k_assign_tree<1><<<gb, tb, 0, st>>>(X_dev, n, h->dx, Tp, h->dOwner);
    else k_assign_tree<2><<<gb, tb, 0, st>>>(X_dev, n, h->dx, Tp, h->dOwner);
  }
  h->plan_fused = false;
  if (n <= BK_MAXN && B <= BK_MAXB && h->bucket_small) {
    // one CTA: stable bucketing + bounds, and the resident path's launch plan when that path follows
    BucketPlan bp;
    memset(&bp, 0, sizeof(bp));
    size_t smp = 0;                      // the plan's share
    if (B == h->B && res_eligible(h) && res_sync_static(h) == GPRF_OK &&
        res_alloc(h, std::max(1, std::min(h->B + h->E, h->n_sm))) == GPRF_OK) {
      bp.enabled = 1;
      res_plan_params(h, &bp.Q);
      smp = res::res_plan_smem(h->B, h->E, BK_WARPS * 32);
      h->plan_fused = true;
    }
    // split launch (CTA 0: totals + plan, CTAs 1 .. npl: placement over BK_WARPS * npl sub-ranges) when there
    // are enough points for the finer ranges and their histograms fit; else one CTA does everything
    int npl = (n >= 4096) ? (h->bucket_split >= 0 ? h->bucket_split : (int)std::min<long long>(8, n / 4096)) : 0;
    while (npl > 0 && ((size_t)BK_WARPS * npl * B + B + 1) * sizeof(int) > 160 * 1024) npl >>= 1;
    size_t smb = npl > 0 ? std::max(((size_t)BK_WARPS * npl * B + B + 1) * sizeof(int), ((size_t)B + 1) * sizeof(int) + smp)
                         : ((size_t)BK_WARPS * B + B + 1) * sizeof(int) + smp;
    if (smb > 200 * 1024) {
      h->err = "bucket kernel: structure too large for one CTA";
      return GPRF_ERR_ARG;
    }
    if (smb > 48 * 1024) cudaFuncSetAttribute(k_bucket_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
    k_bucket_small<<<1 + npl, BK_WARPS * 32, smb, st>>>(h->dOwner, n, B, h->dBlockPtr, h->dPerm, h->dPosBlock, bp);
    CUDA_OK(cudaGetLastError());
    h->part_launches = 2;
  } else {
    size_t tmp = h->capCub;
    cub::DeviceRadixSort::SortPairs(h->dCub, tmp, h->dOwner, h->dPosBlock, h->dIota, h->dIdxSorted, (int)n, 0, bits, st);
    const long long nthr = std::max<long long>(n, B + 1);
    k_block_bounds<<<(unsigned)((nthr + tb - 1) / tb), tb, 0, st>>>(h->dPosBlock, h->dIdxSorted, n, B, h->dBlockPtr,
                                                                      h->dPerm);
    CUDA_OK(cudaGetLastError());
    h->part_launches = 5;
  }
  if (B != h->B) {
    h->adj_dirty = true;
    h->units_built = false;
  }
  h->B = B;
  h->plen = n;
  h->dev_blocks_valid = true;
  h->host_blocks_stale = true;
  h->have_structure = false;
  h->blocks_stream = st;
  return GPRF_OK;
}

// Bring the host view (block_ptr_h, unit descriptors of the tile pipeline) up to date with the
// device-held blocks: one D2H copy + synchronisation, descriptors rebuilt only when a size changed.
static int reblock_finish(gprf_ctx* h, cudaStream_t st) {
  const int B = h->B;
  if (!h->hBlockPtr || h->capHB < (size_t)B + 1) {
    if (h->hBlockPtr) cudaFreeHost(h->hBlockPtr);
    h->hBlockPtr = nullptr;
    CUDA_OK(cudaMallocHost((void**)&h->hBlockPtr, ((size_t)B + 1) * sizeof(long long)));
    h->capHB = (size_t)B + 1;
  }
  CUDA_OK(cudaMemcpyAsync(h->hBlockPtr, h->dBlockPtr, ((size_t)B + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  h->host_blocks_stale = false;
  // Unit descriptors depend on the block SIZES only (which points sit in a block is in dPerm,
  // already rewritten on the device): late in an optimisation few points change block, and then
  // nothing has to be rebuilt or uploaded.
  if (h->units_built && h->blocks_from_device && (int)h->block_ptr_h.size() == B + 1 &&
      memcmp(h->block_ptr_h.data(), h->hBlockPtr, ((size_t)B + 1) * sizeof(long long)) == 0) {
    h->have_structure = true;
    return GPRF_OK;
  }
  h->block_ptr_h.assign(h->hBlockPtr, h->hBlockPtr + B + 1);
  h->blocks_from_device = true;
  return rebuild_units(h, st);
}

static int reblock_device(gprf_ctx* h, const double* X_dev, cudaStream_t st) {
  int rc = reblock_launch(h, X_dev, st);
  if (rc != GPRF_OK) return rc;
  return reblock_finish(h, st);
}

// Host view needed (tile pipeline, gprf_get_blocks) while the blocks live on the device only.
static int ensure_host_blocks(gprf_ctx* h, cudaStream_t st) {
  if (h->host_blocks_stale) return reblock_finish(h, st);
  if (!h->units_built && (int)h->block_ptr_h.size() == h->B + 1) return rebuild_units(h, st);
  return GPRF_OK;
}

extern "C" int gprf_reblock_device(gprf_handle h, const double* X_dev, void* stream) {
  if (!h || !X_dev) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  h->have_structure = false;
  const int rc = res_eligible(h) ? reblock_launch(h, X_dev, (cudaStream_t)stream)   // host view on demand
                                 : reblock_device(h, X_dev, (cudaStream_t)stream);
  if (rc == GPRF_OK) h->pending_part_launches = h->part_launches;
  return rc;
}

extern "C" int gprf_reblock(gprf_handle h, const double* X) {
  if (!h || !X) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  h->have_structure = false;
  const size_t xb = (size_t)h->n * h->dx * sizeof(double);
  memcpy(h->hX, X, xb);
  CUDA_OK(cudaMemcpyAsync(h->dX, h->hX, xb, cudaMemcpyHostToDevice, h->stream));
  int rc = res_eligible(h) ? reblock_launch(h, h->dX, h->stream) : reblock_device(h, h->dX, h->stream);
  if (rc != GPRF_OK) return rc;
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return GPRF_OK;
}

extern "C" int gprf_block_count(gprf_handle h, int* n_blocks, long long* plen) {
  if (!h) return GPRF_ERR_ARG;
  if (h->host_blocks_stale) {
    CUDA_OK(cudaSetDevice(h->device));
    int rc = ensure_host_blocks(h, h->blocks_stream);
    if (rc != GPRF_OK) return rc;
  }
  if ((int)h->block_ptr_h.size() != h->B + 1) return GPRF_ERR_NO_STRUCTURE;
  if (n_blocks) *n_blocks = h->B;
  if (plen) *plen = h->plen;
  return GPRF_OK;
}

extern "C" int gprf_get_blocks(gprf_handle h, long long* block_ptr, long long* perm) {
  if (!h) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  if (h->host_blocks_stale) {
    int rc = ensure_host_blocks(h, h->blocks_stream);
    if (rc != GPRF_OK) return rc;
  }
  if ((int)h->block_ptr_h.size() != h->B + 1) return GPRF_ERR_NO_STRUCTURE;
  if (h->blocks_stream) CUDA_OK(cudaStreamSynchronize(h->blocks_stream));
  if (block_ptr) memcpy(block_ptr, h->block_ptr_h.data(), (size_t)(h->B + 1) * sizeof(long long));
  if (perm && h->plen > 0) {
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaMemcpy(perm, h->dPerm, (size_t)h->plen * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  return GPRF_OK;
}

// Enqueue the per-unit pipeline for the units in `list` (already on device as
// dList[0..nlist)).  Returns the number of kernel launches.
// Units of up to h->fused_nt tiles go through k_unit_fused (one CTA per unit, one launch);
// larger ones through the multi-launch tile pipeline.  `nt_of(i)` = tile count of list entry i
// (lists are sorted largest first, so the large units form a prefix).
static int launch_units_tiled(gprf_ctx* h, const EvalParams& P, const int* list_host, int nlist, int ntmax,
                              bool want_grad, cudaStream_t st, int n_parents);

static int launch_fused(gprf_ctx* h, const EvalParams& P, int base, int cnt, bool want_grad, cudaStream_t st) {
  int launches = 0;
  if (cnt <= 0) return 0;
  EvalParams Pc = P;
  Pc.ulist = P.ulist + base;
  Pc.trace = h->dTrace;
  Pc.trace_ctas = (int)(h->capTrace / (2 * TRACE_SLOTS));
#define CALL_FUSED(D, W) fused_launch<D, W>(Pc, h->dLLu, h->dGthU, want_grad ? 1 : 0, cnt, st)
  LAUNCH(8, DISPATCH_COV(h, CALL_FUSED));
  return launches;
}

// `n_parents`: the fused units [nlarge, nlarge + n_parents) of the list are parents of pairs that
// reuse their factor (main pass only): they are launched first, everything else after them.
static int launch_units(gprf_ctx* h, const EvalParams& P, const int* list_host, int nlist, int ntmax,
                        bool want_grad, cudaStream_t st, int n_parents = 0, int n_tiled_parents = 0) {
  int launches = 0;
  if (nlist == 0) return 0;
  int nlarge = 0, ntl = 0;
  while (nlarge < nlist && h->units[list_host[nlarge]].nt > h->fused_eff) {
    ntl = std::max(ntl, h->units[list_host[nlarge]].nt);
    ++nlarge;
  }
  launches += launch_fused(h, P, nlarge, n_parents, want_grad, st);
  if (nlarge > 0)
    launches += launch_units_tiled(h, P, list_host, nlarge, ntl, want_grad, st, n_tiled_parents);
  launches += launch_fused(h, P, nlarge + n_parents, nlist - nlarge - n_parents, want_grad, st);
  (void)ntmax;
  return launches;
}

// `n_parents`: the first n_parents units of the list store partial U-products that other units of
// the list start from (UnitDesc::pstore): their alpha / K^-1 launches go first.
static int launch_units_tiled(gprf_ctx* h, const EvalParams& P, const int* list_host, int nlist, int ntmax,
                              bool want_grad, cudaStream_t st, int n_parents) {
  int launches = 0;
  if (nlist == 0) return 0;
  const int CH = 32768;   // gridDim.y limit is 65535
  for (int base = 0; base < nlist; base += CH) {
    const int cnt = std::min(CH, nlist - base);
    EvalParams Pc = P;
    Pc.ulist = P.ulist + base;
    LAUNCH(0, (k_prep<<<dim3(ntmax, cnt), NTHREADS, T * (T + 1) * sizeof(double), st>>>(Pc)));
    for (int k = 0; k < ntmax; ++k) {
#define CALL_DIAG(D, W) k_potrf_diag<D, W><<<dim3(1, cnt), NTHREADS, PIPE_BYTES, st>>>(Pc, k)
      if (k == 0) LAUNCH(1, DISPATCH_COV(h, CALL_DIAG));   // diag(k+1) rides in panel(k)
      const int gx = ntmax - k - 1 + h->nya;
      // unit-major once the launch is several waves long (see k_potrf_panel)
      const int um = h->panel_order >= 0 ? h->panel_order : ((long long)cnt * gx >= 4LL * 2 * h->n_sm ? 1 : 0);
#define CALL_PANEL(D, W) \
  k_potrf_panel<D, W><<<um ? dim3(gx, cnt) : dim3(cnt, gx), NTHREADS, PIPE_BYTES, st>>>(Pc, k, um)
      LAUNCH(2, DISPATCH_COV(h, CALL_PANEL));
    }
    if (want_grad) {
      for (int d = 1; d < ntmax; ++d) {
        LAUNCH(3, (k_trtri<<<dim3(ntmax - d, cnt), NTHREADS, PIPE_BYTES, st>>>(Pc, d)));
      }
      const int npc = std::max(0, std::min(n_parents - base, cnt));
      for (int part = 0; part < 2; ++part) {
        const int c0 = part == 0 ? 0 : npc, c1 = part == 0 ? npc : cnt;
        if (c1 <= c0) continue;
        EvalParams Pp = Pc;
        Pp.ulist = Pc.ulist + c0;
        int ntp = 0;                        // parents are blocks: far fewer tiles than ntmax
        for (int q = c0; q < c1; ++q) ntp = std::max(ntp, h->units[list_host[base + q]].nt);
        LAUNCH(4, (k_alpha<<<dim3(ntp * h->nya, c1 - c0), NTHREADS, PIPE_BYTES, st>>>(Pp)));
#define CALL_GRAD(D, W) k_grad<D, W><<<dim3(ntp * (ntp + 1) / 2, c1 - c0), NTHREADS, PIPE_BYTES, st>>>(Pp)
        LAUNCH(5, DISPATCH_COV(h, CALL_GRAD));
      }
    }
    LAUNCH(6, (k_unit_finalize<<<cnt, NTHREADS, 0, st>>>(Pc, h->dLLu, h->dGthU, want_grad ? 1 : 0)));
  }
  return launches;
}

// ---------------------------------------------------------------------------
// Resident path (resident.cuh): plan -> block units -> pair units -> combine, 4 launches, nothing
// read back before the end of the evaluation.  The caller copies out / the status word and
// synchronises once; a non-zero status sends the evaluation through the tile pipeline.
// ---------------------------------------------------------------------------
static int res_sync_static(gprf_ctx* h) {
  if (h->adj_dirty) {
    int rc = rebuild_adjacency(h);
    if (rc != GPRF_OK) return rc;
  }
  if (!h->res_static_dirty) return GPRF_OK;
  h->plan_fused = false;
  const int B = h->B, E = h->E, U = B + E;
  CUDA_OK(ensure(&h->dResEdges, &h->capResE, (size_t)2 * E + 2));
  CUDA_OK(ensure(&h->dResDeg, &h->capResB, (size_t)B + 1));
  if (E > 0)
    CUDA_OK(cudaMemcpy(h->dResEdges, h->edges.data(), (size_t)2 * E * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(h->dResDeg, h->deg.data(), (size_t)B * sizeof(int), cudaMemcpyHostToDevice));
  h->res_have_mask = false;
  if (h->use_explicit_mask) {
    if ((int)h->explicit_mask.size() != U) return GPRF_ERR_ARG;
    h->res_mask = h->explicit_mask;
    h->res_have_mask = true;
  } else if (h->shard_world > 1) {
    std::vector<double> sizes;
    shard_sizes(h, sizes);
    lpt_mask(sizes, B, h->edges.data(), h->shard_rank, h->shard_world, h->res_mask);
    h->res_have_mask = true;
  }
  if (h->res_have_mask) {
    CUDA_OK(ensure(&h->dResActive, &h->capResAct, (size_t)U + 1));
    CUDA_OK(cudaMemcpy(h->dResActive, h->res_mask.data(), (size_t)U, cudaMemcpyHostToDevice));
  }
  h->res_static_dirty = false;
  return GPRF_OK;
}

static void res_plan_params(gprf_ctx* h, res::PlanParams* Q) {
  Q->block_ptr = h->dBlockPtr;
  Q->edges = h->dResEdges;
  Q->active = h->res_have_mask ? h->dResActive : nullptr;
  Q->B = h->B;
  Q->E = h->E;
  Q->G = std::max(1, std::min(h->B + h->E, h->n_sm));
  Q->sort_blocks = h->res_sort_blocks ? 1 : 0;
  Q->list_ptr = h->dResListPtr;
  Q->order = h->dResOrderB;        // sized for all units
  Q->counts = h->dResCounts;
  Q->status = h->dResCounts + 4;
  Q->info = h->dResInfo;
}

static int res_alloc(gprf_ctx* h, int grid) {
  const size_t B = (size_t)h->B, U = (size_t)h->B + h->E;
  CUDA_OK(ensure(&h->dResExports, &h->capResExp, B * (size_t)res::EXP_STRIDE));
  CUDA_OK(ensure(&h->dResScratch, &h->capResScr, (size_t)grid * (size_t)res::SCR_STRIDE));
  if (U > h->capResU || !h->dResLL) {
    cudaFree(h->dResLL); cudaFree(h->dResGth); cudaFree(h->dResGx); cudaFree(h->dResInfo);
    cudaFree(h->dResOrderB); cudaFree(h->dResOrderP); cudaFree(h->dResCounts); cudaFree(h->dResReady);
    h->dResLL = h->dResGth = h->dResGx = nullptr;
    h->dResInfo = h->dResOrderB = h->dResOrderP = h->dResCounts = h->dResReady = nullptr;
    const size_t cu = U + U / 4 + 16;
    CUDA_OK(cudaMalloc((void**)&h->dResLL, cu * sizeof(double)));
    CUDA_OK(cudaMalloc((void**)&h->dResGth, cu * MAX_NCOV * sizeof(double)));
    CUDA_OK(cudaMalloc((void**)&h->dResGx, cu * res::GX_STRIDE * sizeof(double)));
    CUDA_OK(cudaMalloc((void**)&h->dResInfo, cu * sizeof(int)));
    CUDA_OK(cudaMalloc((void**)&h->dResOrderB, cu * sizeof(int)));
    CUDA_OK(cudaMalloc((void**)&h->dResOrderP, cu * sizeof(int)));
    CUDA_OK(cudaMalloc((void**)&h->dResCounts, 8 * sizeof(int)));
    CUDA_OK(cudaMalloc((void**)&h->dResReady, 3 * cu * sizeof(int)));
    CUDA_OK(cudaMemset(h->dResReady, 0, 3 * cu * sizeof(int)));
    h->res_epoch = 0;
    CUDA_OK(cudaMemset(h->dResLL, 0, cu * sizeof(double)));
    CUDA_OK(cudaMemset(h->dResGth, 0, cu * MAX_NCOV * sizeof(double)));
    CUDA_OK(cudaMemset(h->dResGx, 0, cu * res::GX_STRIDE * sizeof(double)));
    h->capResU = cu;
  }
  if (!h->dResDbg) CUDA_OK(cudaMalloc((void**)&h->dResDbg, 2 * 160 * 160 * sizeof(double)));
  if (!h->dResListPtr) CUDA_OK(cudaMalloc((void**)&h->dResListPtr, ((size_t)h->n_sm + 2) * sizeof(int)));
  return GPRF_OK;
}

// Enqueue one evaluation on the resident path.  out_dev = [ll, grad theta (5), gradX (n dx)].
static int run_resident(gprf_ctx* h, const double* X_dev, const CovParams& cp, int grad_X, int grad_cov,
                        double* out_dev, cudaStream_t st, int* launches_out, double* status_dev = nullptr) {
  int rc = res_sync_static(h);
  if (rc != GPRF_OK) return rc;
  const int B = h->B, E = h->E;
  rc = res_alloc(h, std::max(1, std::min(B + E, h->n_sm)));
  if (rc != GPRF_OK) return rc;
  int launches = 0;
  const size_t outlen = 1 + MAX_NCOV + (grad_X ? (size_t)h->n * h->dx : 0);
  // k_res_combine writes every entry of out when the blocks cover all points (always after re-blocking)
  if (h->plen != (long long)h->n) CUDA_OK(cudaMemsetAsync(out_dev, 0, outlen * sizeof(double), st));
  if (h->plan_fused) {
    h->plan_fused = false;           // the bucketing kernel of this evaluation's re-blocking built the plan
  } else {
    res::PlanParams Q;
    res_plan_params(h, &Q);
    const size_t plan_sm = res::res_plan_smem(B, E, 512);
    if (plan_sm > 48 * 1024)
      cudaFuncSetAttribute(res::k_res_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan_sm);
    LAUNCH(9, (res::k_res_plan<<<1, 512, plan_sm, st>>>(Q)));
  }
  res::ResParams P;
  P.X = X_dev;
  P.Y = h->dY;
  P.perm = h->dPerm;
  P.block_ptr = h->dBlockPtr;
  P.edges = h->dResEdges;
  P.B = B;
  P.dx = h->dx;
  P.dy = h->dy;
  P.nyb = (h->dy + 7) / 8;
  P.want_grad = (grad_X || grad_cov) ? 1 : 0;
  P.cp = cp;
  P.exports = h->dResExports;
  P.scratch = h->dResScratch;
  P.ll_u = h->dResLL;
  P.gth_u = h->dResGth;
  P.gx_u = h->dResGx;
  P.info = h->dResInfo;
  P.status = h->dResCounts + 4;
  P.spin_limit = h->res_spin_limit;
  P.defer_ok = h->res_defer ? 1 : 0;
  P.dbg_unit = h->res_dbg_unit;
  P.dbg_phase = h->res_dbg_phase;
  P.dbg_out = h->dResDbg;
  // debug timeline (gprf_debug_trace with >= n_sm CTAs)
  P.trace = (h->dTrace && h->capTrace >= (size_t)h->n_sm * 2 * res::RTRACE_SLOTS) ? h->dTrace : nullptr;
  P.ready = h->dResReady;
  P.ready2 = h->dResReady + h->capResU;
  P.ready0 = h->res_early ? h->dResReady + 2 * h->capResU : nullptr;
  P.epoch = ++h->res_epoch;
  P.order = h->dResOrderB;
  P.n_order = h->dResCounts + 0;
  P.list_ptr = h->dResListPtr;
  // ONE launch: block units first, the pair units behind them wait for their parent's exports
  const int grid_u = std::max(1, std::min(B + E, h->n_sm));
  if (P.trace) P.trace = h->dTrace;
#define CALL_RESU(D, W) res::resident_launch<D, W>(P, grid_u, st)
  LAUNCH(11, DISPATCH_COV(h, CALL_RESU));
  res::ResCombine C;
  C.perm = h->dPerm;
  C.pos_block = h->dPosBlock;
  C.block_ptr = h->dBlockPtr;
  C.adj_ptr = h->dAdjPtr;
  C.adj_edge = h->dAdjEdge;
  C.adj_side = h->dAdjSide;
  C.edges = h->dResEdges;
  C.deg = h->dResDeg;
  C.raw = h->raw_weights ? 1 : 0;
  C.active = h->res_have_mask ? h->dResActive : nullptr;
  C.ll_u = h->dResLL;
  C.gth_u = h->dResGth;
  C.gx_u = h->dResGx;
  C.status = h->dResCounts + 4;
  C.B = B;
  C.E = E;
  C.dx = h->dx;
  C.plen = h->plen;
  const unsigned gc = 1 + (grad_X ? (unsigned)((h->plen + 255) / 256) : 0);
  LAUNCH(12, (res::k_res_combine<<<gc, 256, 0, st>>>(C, out_dev, grad_X ? 1 : 0, grad_cov ? 1 : 0, status_dev)));
  CUDA_OK(cudaGetLastError());
  rc = apply_prior(h, X_dev, out_dev, grad_X, st);
  if (rc != GPRF_OK) return rc;
  if (h->prior_apply) ++launches;
  if (launches_out) *launches_out = launches;
  return GPRF_OK;
}

// Try the resident path for one evaluation.  Returns GPRF_OK with *done = 1 when the results are in
// out_dev / host_out; *done = 0 means "run the tile pipeline" (not eligible, a unit did not fit, or
// a pivot failed and the jitter rule has to be applied).
static int try_resident(gprf_ctx* h, const double* X_dev, const double* theta, int ncov, int grad_X, int grad_cov,
                        double* out_dev, cudaStream_t st, double* host_out, int* done) {
  *done = 0;
  h->last_resident = false;
  if (!res_eligible(h) || !h->dev_blocks_valid || h->plen <= 0) return GPRF_OK;
  CovParams cp;
  int rc = make_cov(h, theta, ncov, &cp);
  if (rc != GPRF_OK) {
    h->err = "theta must have 2 + (number of lengthscales) entries";
    return rc;
  }
  int launches = 0;
  // One resident launch per process at a time: its CTAs wait for each other (a pair spins on its parent
  // block's flag), which is only safe while the whole grid can become resident.  Two handles driven from
  // two host threads are serialised here (the evaluation is synchronous anyway); the kernel's watchdog
  // (ST_TIMEOUT) covers what a mutex cannot (other processes on the same GPU).
  static std::mutex res_mutex;
  std::lock_guard<std::mutex> res_lock(res_mutex);
  CUDA_OK(cudaEventRecord(h->ev0, st));
  rc = run_resident(h, X_dev, cp, grad_X, grad_cov, out_dev, st, &launches);
  if (rc != GPRF_OK) return rc;
  CUDA_OK(cudaEventRecord(h->ev1, st));
  const size_t outlen = 1 + MAX_NCOV + (grad_X ? (size_t)h->n * h->dx : 0);
  if (host_out) CUDA_OK(cudaMemcpyAsync(host_out, out_dev, outlen * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(h->hResStatus, h->dResCounts + 4, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  h->res_evals++;
  h->res_last_status = h->hResStatus[0];
  if (h->hResStatus[0] != 0) {
    h->res_fallbacks++;
    return GPRF_OK;
  }
  h->last_launches = launches;
  h->last_resident = true;
  *done = 1;
  return GPRF_OK;
}

// On return the stream has drained: out_dev (and host_out, a pinned buffer, when given) hold the results.
static int run_eval(gprf_ctx* h, const double* X_dev, const double* theta, int ncov, int grad_X, int grad_cov,
                    double* out_dev, cudaStream_t st, int* failed_unit, double* host_out = nullptr) {
  if (failed_unit) *failed_unit = -1;
  if (!h->have_structure) return GPRF_ERR_NO_STRUCTURE;
  CovParams cp;
  int rc = make_cov(h, theta, ncov, &cp);
  if (rc != GPRF_OK) {
    h->err = "theta must have 2 + (number of lengthscales) entries";
    return rc;
  }
  const bool want_grad = grad_X || grad_cov;
  const int U = h->U;
  EvalParams P;
  P.units = h->dUnits;
  P.ulist = h->dList;
  P.arena = h->arena;
  P.X = X_dev;
  P.Y = h->dY;
  P.perm = h->dPerm;
  P.jitter = h->dJitter;
  P.info = h->dInfo;
  P.nfail = h->dNfail;
  P.dx = h->dx;
  P.dy = h->dy;
  P.yr = h->yr;
  P.nya = h->nya;
  P.keep_kinv = h->keep_kinv ? 1 : 0;
  P.no_share = 0;
  P.cp = cp;
  P.trace = nullptr;        // set for the fused launches only (launch_units)
  P.trace_ctas = 0;

  std::fill(h->jitter.begin(), h->jitter.end(), 0.0);
  std::fill(h->tries.begin(), h->tries.end(), 0);
  CUDA_OK(cudaEventRecord(h->ev0, st));
  CUDA_OK(cudaMemsetAsync(h->dScratch, 0, h->scratch_bytes, st));
  const size_t outlen = 1 + MAX_NCOV + (grad_X ? (size_t)h->n * h->dx : 0);
  CUDA_OK(cudaMemsetAsync(out_dev, 0, outlen * sizeof(double), st));
  const int nlist = (int)h->all_list.size();
  P.ulist = h->dListAll;
  int launches = launch_units(h, P, h->all_list.data(), nlist, h->ntmax, want_grad, st,
                              h->any_share ? h->n_fused_parents : 0, h->any_share ? h->n_tiled_parents : 0);
  P.ulist = h->dList;
  CUDA_OK(cudaGetLastError());

  // weighted sums over the units (gprf.py:245-291) and, for the host-buffer entries, the D2H copy
  auto combine = [&]() -> int {
    LAUNCH(7, (k_combine_scalars<<<1, 256, 0, st>>>(h->dUnits, U, h->dLLu, h->dGthU, grad_cov ? 1 : 0, out_dev)));
    if (grad_X && h->plen > 0) {
      CombineParams C;
      C.units = h->dUnits;
      C.arena = h->arena;
      C.perm = h->dPerm;
      C.pos_block = h->dPosBlock;
      C.block_ptr = h->dBlockPtr;
      C.adj_ptr = h->dAdjPtr;
      C.adj_edge = h->dAdjEdge;
      C.adj_side = h->dAdjSide;
      C.B = h->B;
      C.dx = h->dx;
      C.plen = h->plen;
      const int tb = 64;
      LAUNCH(7, (k_combine_gradx<<<(unsigned)((h->plen + tb - 1) / tb), tb, 0, st>>>(C, out_dev + 1 + MAX_NCOV)));
    }
    CUDA_OK(cudaGetLastError());
    {
      int prc = apply_prior(h, X_dev, out_dev, grad_X, st);
      if (prc != GPRF_OK) return prc;
    }
    CUDA_OK(cudaEventRecord(h->ev1, st));
    if (host_out) CUDA_OK(cudaMemcpyAsync(host_out, out_dev, outlen * sizeof(double), cudaMemcpyDeviceToHost, st));
    return GPRF_OK;
  };
  // The combination is enqueued before the Cholesky status is known (one host round trip per
  // evaluation instead of two); in the rare case of a failed unit it is simply redone below.
  rc = combine();
  if (rc != GPRF_OK) return rc;

  // jitter rule (gpy_linalg.py:77-97): host-driven retry of the failed units only
  CUDA_OK(cudaMemcpyAsync(h->hNfail, h->dNfail, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  int nfail = *h->hNfail;
  const bool any_retry = nfail > 0;
  std::vector<int> info;
  while (nfail > 0) {
    if (!(cp.s2 + cp.nv > 0.0)) return GPRF_ERR_NONPOS_DIAG;
    info.resize(U);
    CUDA_OK(cudaMemcpy(info.data(), h->dInfo, (size_t)U * sizeof(int), cudaMemcpyDeviceToHost));
    if (!P.no_share && h->any_share) {
      // A pair that read its leading tiles from block i never ran their pivots: a failure of the
      // parent inside the shared tiles is the pair's failure too (its own dpotrf would have
      // stopped at the same pivot, gpy_linalg.py:81-83).
      for (int e = 0; e < h->E; ++e) {
        const UnitDesc& u = h->units[h->B + e];
        const int pi = info[h->edges[2 * e]];
        if (u.share > 0 && info[h->B + e] == 0 && pi != 0 && pi - 1 < u.share * T) info[h->B + e] = pi;
      }
    }
    P.no_share = 1;            // retried units factor all of their own tiles
    std::vector<int> failed;
    int ntm = 0;
    for (int uix = 0; uix < U; ++uix) {
      if (info[uix] == 0) continue;
      if (h->tries[uix] >= 5) {
        if (failed_unit) *failed_unit = uix;
        return GPRF_ERR_NOT_PD;
      }
      h->jitter[uix] = (cp.s2 + cp.nv) * 1e-6 * std::pow(10.0, h->tries[uix]);
      h->tries[uix]++;
      if (!std::isfinite(h->jitter[uix])) {
        if (failed_unit) *failed_unit = uix;
        return GPRF_ERR_NOT_PD;
      }
      failed.push_back(uix);
      ntm = std::max(ntm, h->units[uix].nt);
    }
    CUDA_OK(cudaMemcpyAsync(h->dJitter, h->jitter.data(), (size_t)U * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemsetAsync(h->dInfo, 0, (size_t)U * sizeof(int), st));
    CUDA_OK(cudaMemsetAsync(h->dNfail, 0, sizeof(int), st));
    std::stable_sort(failed.begin(), failed.end(), [&](int a, int b) { return h->units[a].s > h->units[b].s; });
    CUDA_OK(cudaMemcpyAsync(h->dList, failed.data(), failed.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    launches += launch_units(h, P, failed.data(), (int)failed.size(), ntm, want_grad, st);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(h->hNfail, h->dNfail, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    nfail = *h->hNfail;
  }
  if (any_retry) {
    // (entries of out no block covers are not rewritten by the combination: cleared so that the prior
    // epilogue, which works in place, is not applied to them twice)
    if (h->prior_apply) CUDA_OK(cudaMemsetAsync(out_dev, 0, outlen * sizeof(double), st));
    rc = combine();
    if (rc != GPRF_OK) return rc;
    CUDA_OK(cudaStreamSynchronize(st));
  }
  h->last_launches = launches;
  return GPRF_OK;
}

extern "C" int gprf_llgrad_device(gprf_handle h, const double* X_dev, const double* theta, int ncov,
                                  int grad_X, int grad_cov, double* out_dev, void* stream, int* failed_unit) {
  if (!h || !X_dev || !theta || !out_dev) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (failed_unit) *failed_unit = -1;
  int done = 0;
  int rc = try_resident(h, X_dev, theta, ncov, grad_X, grad_cov, out_dev, st, nullptr, &done);
  if (rc != GPRF_OK) return rc;
  if (!done) {
    rc = ensure_host_blocks(h, st);
    if (rc != GPRF_OK) return rc;
    rc = run_eval(h, X_dev, theta, ncov, grad_X, grad_cov, out_dev, st, failed_unit);
    if (rc != GPRF_OK) return rc;
  }
  h->last_launches += h->pending_part_launches;
  h->pending_part_launches = 0;
  CUDA_OK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  if (h->profile) h->prof_resolve();
  return GPRF_OK;
}

// Device-resident evaluation WITHOUT a host round trip: on the resident path the launches are enqueued on
// `stream` and the call returns at once; the evaluation's status word is written (as a double, 0 = ok) to
// *status_dev by the combination kernel, next to the results in out_dev.  Meant for multi-GPU use
// (gprf.py:218-233 runs the units in a process pool): the status travels inside the all-reduce of the packed
// results, the host synchronises once after the collective, and only when the reduced status is non-zero
// (a unit needs the jitter rule, a block does not fit) every rank repeats the evaluation with
// gprf_llgrad_device.  *enqueued = 1: asynchronous; 0: the structure is not of the resident kind, the
// evaluation ran synchronously (tile pipeline, jitter rule and errors as in gprf_llgrad_device) and
// *status_dev is 0.
extern "C" int gprf_llgrad_device_nosync(gprf_handle h, const double* X_dev, const double* theta, int ncov,
                                         int grad_X, int grad_cov, double* out_dev, double* status_dev, void* stream,
                                         int* enqueued, int* failed_unit) {
  if (!h || !X_dev || !theta || !out_dev || !status_dev || !enqueued) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  *enqueued = 0;
  if (failed_unit) *failed_unit = -1;
  if (res_eligible(h) && h->dev_blocks_valid && h->plen > 0) {
    CovParams cp;
    int rc = make_cov(h, theta, ncov, &cp);
    if (rc != GPRF_OK) {
      h->err = "theta must have 2 + (number of lengthscales) entries";
      return rc;
    }
    int launches = 0;
    CUDA_OK(cudaEventRecord(h->ev0, st));
    rc = run_resident(h, X_dev, cp, grad_X, grad_cov, out_dev, st, &launches, status_dev);
    if (rc != GPRF_OK) return rc;
    CUDA_OK(cudaEventRecord(h->ev1, st));
    h->res_evals++;
    h->res_last_status = 0;               // unknown to the host; the caller reports failures by redoing
    h->last_launches = launches + h->pending_part_launches;
    h->pending_part_launches = 0;
    h->last_resident = true;
    *enqueued = 1;
    return GPRF_OK;
  }
  int rc = gprf_llgrad_device(h, X_dev, theta, ncov, grad_X, grad_cov, out_dev, stream, failed_unit);
  if (rc != GPRF_OK) return rc;
  CUDA_OK(cudaMemsetAsync(status_dev, 0, sizeof(double), st));
  return GPRF_OK;
}

extern "C" int gprf_llgrad(gprf_handle h, const double* X, const double* theta, int ncov, int grad_X,
                           int grad_cov, double* ll, double* gradX, double* gradTheta, int* failed_unit) {
  if (!h || !X || !theta || !ll) return GPRF_ERR_ARG;
  if ((grad_X && !gradX) || (grad_cov && !gradTheta)) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  const size_t xb = (size_t)h->n * h->dx * sizeof(double);
  memcpy(h->hX, X, xb);
  CUDA_OK(cudaMemcpyAsync(h->dX, h->hX, xb, cudaMemcpyHostToDevice, h->stream));
  if (failed_unit) *failed_unit = -1;
  int done = 0;
  int rc = try_resident(h, h->dX, theta, ncov, grad_X, grad_cov, h->dOut, h->stream, h->hOut, &done);
  if (rc != GPRF_OK) return rc;
  if (!done) {
    rc = ensure_host_blocks(h, h->stream);
    if (rc != GPRF_OK) return rc;
    rc = run_eval(h, h->dX, theta, ncov, grad_X, grad_cov, h->dOut, h->stream, failed_unit, h->hOut);
    if (rc != GPRF_OK) return rc;
  }
  CUDA_OK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  if (h->profile) h->prof_resolve();
  *ll = h->hOut[0];
  if (grad_cov)
    for (int t = 0; t < ncov; ++t) gradTheta[t] = h->hOut[1 + t];
  if (grad_X) memcpy(gradX, h->hOut + 1 + MAX_NCOV, xb);
  return GPRF_OK;
}

extern "C" int gprf_llgrad_reblock(gprf_handle h, const double* X, const double* theta, int ncov, int grad_X,
                                   int grad_cov, double* ll, double* gradX, double* gradTheta, int* failed_unit) {
  if (!h || !X || !theta || !ll) return GPRF_ERR_ARG;
  if ((grad_X && !gradX) || (grad_cov && !gradTheta)) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  const size_t xb = (size_t)h->n * h->dx * sizeof(double);
  memcpy(h->hX, X, xb);
  CUDA_OK(cudaMemcpyAsync(h->dX, h->hX, xb, cudaMemcpyHostToDevice, h->stream));
  h->have_structure = false;
  if (failed_unit) *failed_unit = -1;
  int rc = reblock_launch(h, h->dX, h->stream);
  if (rc != GPRF_OK) return rc;
  int done = 0;
  rc = try_resident(h, h->dX, theta, ncov, grad_X, grad_cov, h->dOut, h->stream, h->hOut, &done);
  if (rc != GPRF_OK) return rc;
  if (!done) {
    rc = reblock_finish(h, h->stream);
    if (rc != GPRF_OK) return rc;
    rc = run_eval(h, h->dX, theta, ncov, grad_X, grad_cov, h->dOut, h->stream, failed_unit, h->hOut);
    if (rc != GPRF_OK) return rc;
  }
  CUDA_OK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  if (h->profile) h->prof_resolve();
  h->last_launches += h->part_launches;
  *ll = h->hOut[0];
  if (grad_cov)
    for (int t = 0; t < ncov; ++t) gradTheta[t] = h->hOut[1 + t];
  if (grad_X) memcpy(gradX, h->hOut + 1 + MAX_NCOV, xb);
  return GPRF_OK;
}

// ---- optimiser glue ---------------------------------------------------------------------------------------------
extern "C" int gprf_set_x_prior(gprf_handle h, const double* mean, const double* inv_var, const double* grad_scale) {
  if (!h) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  if (!mean) {
    h->have_prior = false;
    return GPRF_OK;
  }
  if (!inv_var) return GPRF_ERR_ARG;
  const size_t nb = (size_t)h->n * h->dx * sizeof(double);
  if (!h->dPriorMean) CUDA_OK(cudaMalloc((void**)&h->dPriorMean, std::max<size_t>(nb, 8)));
  if (!h->dPriorPartial) CUDA_OK(cudaMalloc((void**)&h->dPriorPartial, (size_t)2 * h->n_sm * sizeof(double)));
  if (!h->dPriorCounter) {
    CUDA_OK(cudaMalloc((void**)&h->dPriorCounter, sizeof(unsigned)));
    CUDA_OK(cudaMemset(h->dPriorCounter, 0, sizeof(unsigned)));
  }
  CUDA_OK(cudaMemcpy(h->dPriorMean, mean, nb, cudaMemcpyHostToDevice));
  for (int d = 0; d < h->dx; ++d) {
    h->prior_ivar[d] = inv_var[d];
    h->prior_gscale[d] = grad_scale ? grad_scale[d] : 1.0;
  }
  h->have_prior = true;
  return GPRF_OK;
}

extern "C" int gprf_neg_objective(gprf_handle h, const double* X, const double* theta, int ncov, int grad_cov,
                                  int reblock, double* f, double* g, double* gradTheta, int* failed_unit) {
  if (!h || !X || !theta || !f || !g) return GPRF_ERR_ARG;
  if (!h->have_prior) {
    h->err = "gprf_neg_objective: no prior set (gprf_set_x_prior)";
    return GPRF_ERR_ARG;
  }
  h->prior_apply = true;
  const int rc = reblock ? gprf_llgrad_reblock(h, X, theta, ncov, 1, grad_cov, f, g, gradTheta, failed_unit)
                         : gprf_llgrad(h, X, theta, ncov, 1, grad_cov, f, g, gradTheta, failed_unit);
  h->prior_apply = false;
  return rc;
}

extern "C" int gprf_unit_results(gprf_handle h, double* ll_units, double* jitter_units) {
  if (!h || (!h->have_structure && !h->last_resident)) return GPRF_ERR_NO_STRUCTURE;
  CUDA_OK(cudaSetDevice(h->device));
  if (h->last_resident) {
    const size_t U = (size_t)h->B + h->E;
    if (ll_units) CUDA_OK(cudaMemcpy(ll_units, h->dResLL, U * sizeof(double), cudaMemcpyDeviceToHost));
    if (jitter_units) memset(jitter_units, 0, U * sizeof(double));
    return GPRF_OK;
  }
  if (ll_units) CUDA_OK(cudaMemcpy(ll_units, h->dLLu, (size_t)h->U * sizeof(double), cudaMemcpyDeviceToHost));
  if (jitter_units) memcpy(jitter_units, h->jitter.data(), (size_t)h->U * sizeof(double));
  return GPRF_OK;
}

// ---- resident path: switches, statistics, debug ------------------------------------------------
extern "C" int gprf_set_resident(gprf_handle h, int on) {
  if (!h) return GPRF_ERR_ARG;
  h->res_enable = on != 0;
  h->res_static_dirty = true;
  h->units_built = false;          // the multi-GPU split depends on it (shard_sizes)
  if ((int)h->block_ptr_h.size() == h->B + 1 && h->B > 0 && !h->host_blocks_stale) return replan(h);
  return GPRF_OK;
}

extern "C" int gprf_resident_stats(gprf_handle h, long long* evals, long long* fallbacks, int* last_status) {
  if (!h) return GPRF_ERR_ARG;
  if (evals) *evals = h->res_evals;
  if (fallbacks) *fallbacks = h->res_fallbacks;
  if (last_status) *last_status = h->res_last_status;
  return GPRF_OK;
}

extern "C" int gprf_resident_layout(long long* out, int n) {
  const long long v[] = {res::EMAXB, res::RNYB, res::RBLK, res::EXP_W, res::EXP_KINV, res::EXP_ZY, res::EXP_AROW,
                         res::EXP_SCAL, res::EXP_STRIDE, res::GX_STRIDE, res::R_CAP_DOUBLES, res::EXP_KSAVE};
  const int m = (int)(sizeof(v) / sizeof(v[0]));
  if (!out || n < m) return m;
  for (int i = 0; i < m; ++i) out[i] = v[i];
  return m;
}

extern "C" int gprf_set_resident_debug(gprf_handle h, int unit, int phase) {
  if (!h) return GPRF_ERR_ARG;
  h->res_dbg_unit = unit;
  h->res_dbg_phase = phase;
  return GPRF_OK;
}

// out_dump: 2 x 128 x 128 doubles (R1, R2 of the unit / phase selected with gprf_set_resident_debug);
// out_export: EXP_STRIDE doubles, the raw export record of `block`; out_unit: ll, grad theta (5),
// gradX rows (GX_STRIDE) of unit `unit`.  Any pointer may be NULL.
extern "C" int gprf_get_resident_debug(gprf_handle h, double* out_dump, int block, double* out_export, int unit,
                                       double* out_unit) {
  if (!h || !h->dResLL) return GPRF_ERR_NO_STRUCTURE;
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaDeviceSynchronize());
  if (out_dump) CUDA_OK(cudaMemcpy(out_dump, h->dResDbg, 2 * 160 * 160 * sizeof(double), cudaMemcpyDeviceToHost));
  if (out_export) {
    if (block < 0 || block >= h->B) return GPRF_ERR_ARG;
    CUDA_OK(cudaMemcpy(out_export, h->dResExports + (size_t)block * res::EXP_STRIDE, res::EXP_STRIDE * sizeof(double),
                       cudaMemcpyDeviceToHost));
  }
  if (out_unit) {
    if (unit < 0 || unit >= h->B + h->E) return GPRF_ERR_ARG;
    CUDA_OK(cudaMemcpy(out_unit, h->dResLL + unit, sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(out_unit + 1, h->dResGth + (size_t)unit * MAX_NCOV, MAX_NCOV * sizeof(double),
                       cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(out_unit + 1 + MAX_NCOV, h->dResGx + (size_t)unit * res::GX_STRIDE,
                       res::GX_STRIDE * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return GPRF_OK;
}

extern "C" int gprf_last_timing(gprf_handle h, float* ms, int* launches) {
  if (!h) return GPRF_ERR_ARG;
  if (ms) *ms = h->last_ms;
  if (launches) *launches = h->last_launches;
  return GPRF_OK;
}

extern "C" int gprf_set_profiling(gprf_handle h, int on) {
  if (!h) return GPRF_ERR_ARG;
  h->profile = (on != 0);
  return GPRF_OK;
}

extern "C" const char* gprf_family_name(int fam) {
  static const char* names[GPRF_N_FAMILIES] = {"prep", "potrf_diag", "potrf_panel", "trtri",
                                               "alpha", "kinv_grad", "unit_finalize", "combine", "unit_fused",
                                               "res_plan", "res_blocks", "res_pairs", "res_combine"};
  return (fam >= 0 && fam < GPRF_N_FAMILIES) ? names[fam] : "";
}

extern "C" int gprf_family_timing(gprf_handle h, float* ms, int* launches) {
  if (!h) return GPRF_ERR_ARG;
  for (int f = 0; f < GPRF_N_FAMILIES; ++f) {
    if (ms) ms[f] = h->fam_ms[f];
    if (launches) launches[f] = h->fam_launches[f];
  }
  return GPRF_OK;
}

extern "C" int gprf_debug_unit(gprf_handle h, int unit, int* s, int* sp, int* yr, double* M, double* alpha,
                               double* gx_unit) {
  if (!h || !h->have_structure) return GPRF_ERR_NO_STRUCTURE;
  if (unit < 0 || unit >= h->U) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  const UnitDesc& u = h->units[unit];
  if (s) *s = u.s;
  if (sp) *sp = u.sp;
  if (yr) *yr = h->yr;
  if (!u.active || u.s == 0) return GPRF_OK;
  if (M) CUDA_OK(cudaMemcpy(M, h->arena + u.m_off, (size_t)(u.sp + h->yr) * u.sp * sizeof(double), cudaMemcpyDeviceToHost));
  if (alpha) CUDA_OK(cudaMemcpy(alpha, h->arena + u.al_off, (size_t)u.sp * h->yr * sizeof(double), cudaMemcpyDeviceToHost));
  if (gx_unit) CUDA_OK(cudaMemcpy(gx_unit, h->arena + u.gx_off, (size_t)u.sp * 3 * sizeof(double), cudaMemcpyDeviceToHost));
  return GPRF_OK;
}

// out_host (s x dy) = L_unit Y_unit of the last evaluation's factorisation (tile pipeline; the unit must
// be active and dy <= 64).  See k_unit_lmul.
extern "C" int gprf_unit_lmul(gprf_handle h, int unit, double* out_host) {
  if (!h || !out_host) return GPRF_ERR_ARG;
  if (!h->have_structure || h->last_resident) return GPRF_ERR_NO_STRUCTURE;
  if (unit < 0 || unit >= h->U || h->dy > 64) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  const UnitDesc& u = h->units[unit];
  if (!u.active || u.s == 0) return GPRF_OK;
  if (unit >= h->B) {
    h->err = "gprf_unit_lmul is defined for block units";
    return GPRF_ERR_ARG;
  }
  double* d_out = nullptr;
  CUDA_OK(cudaMalloc((void**)&d_out, (size_t)u.s * h->dy * sizeof(double)));
  k_unit_lmul<<<(u.s + 63) / 64, 256, 0, h->stream>>>(h->arena + u.m_off, u.sp, u.s, h->dY, h->dy, h->dPerm, u.a_start, d_out);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d_out, (size_t)u.s * h->dy * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_out);
  CUDA_OK(e);
  return GPRF_OK;
}

extern "C" int gprf_kernel_matrix(gprf_handle h, const double* X1, long long n1, const double* X2, long long n2,
                                  const double* theta, int ncov, double* K) {
  if (!h || !X1 || !theta || !K || n1 < 0) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  CovParams cp;
  if (make_cov(h, theta, ncov, &cp) != GPRF_OK) return GPRF_ERR_ARG;
  const bool self = (X2 == nullptr);
  if (self) n2 = n1;
  if (n1 == 0 || n2 == 0) return GPRF_OK;
  double *d1 = nullptr, *d2 = nullptr, *dK = nullptr;
  CUDA_OK(cudaMalloc((void**)&d1, (size_t)n1 * h->dx * sizeof(double)));
  CUDA_OK(cudaMemcpy(d1, X1, (size_t)n1 * h->dx * sizeof(double), cudaMemcpyHostToDevice));
  if (!self) {
    CUDA_OK(cudaMalloc((void**)&d2, (size_t)n2 * h->dx * sizeof(double)));
    CUDA_OK(cudaMemcpy(d2, X2, (size_t)n2 * h->dx * sizeof(double), cudaMemcpyHostToDevice));
  }
  CUDA_OK(cudaMalloc((void**)&dK, (size_t)n1 * n2 * sizeof(double)));
  const int grid = (int)std::min<long long>((n1 * n2 + 255) / 256, 148 * 16);
#define CALL_KM(D, W) k_kernel_matrix<D, W><<<grid, 256>>>(d1, n1, self ? d1 : d2, n2, h->dx, cp, self ? 1 : 0, dK)
  DISPATCH_COV(h, CALL_KM);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpy(K, dK, (size_t)n1 * n2 * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(d1); cudaFree(d2); cudaFree(dK);
  return GPRF_OK;
}

extern "C" int gprf_kernel_deriv(gprf_handle h, const double* X, long long n, const double* theta, int ncov,
                                 int mode, int p, int which, double* out) {
  if (!h || !X || !theta || !out || n < 0 || mode < 0 || mode > 1 || which < 0) return GPRF_ERR_ARG;
  if (mode == 0 && (p < 0 || p >= n || which >= h->dx)) return GPRF_ERR_ARG;
  if (mode == 1 && which >= h->nls) return GPRF_ERR_ARG;
  CUDA_OK(cudaSetDevice(h->device));
  CovParams cp;
  if (make_cov(h, theta, ncov, &cp) != GPRF_OK) return GPRF_ERR_ARG;
  if (n == 0) return GPRF_OK;
  const long long tot = mode == 0 ? n : n * n;
  double *dXq = nullptr, *dO = nullptr;
  CUDA_OK(cudaMalloc((void**)&dXq, (size_t)n * h->dx * sizeof(double)));
  CUDA_OK(cudaMalloc((void**)&dO, (size_t)tot * sizeof(double)));
  CUDA_OK(cudaMemcpy(dXq, X, (size_t)n * h->dx * sizeof(double), cudaMemcpyHostToDevice));
  const int grid = (int)std::min<long long>((tot + 255) / 256, 148 * 16);
#define CALL_KD(D, W) k_kernel_deriv<D, W><<<grid, 256>>>(dXq, n, h->dx, cp, mode, p, which, dO)
  DISPATCH_COV(h, CALL_KD);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpy(out, dO, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(dXq);
  cudaFree(dO);
  return GPRF_OK;
}

extern "C" int gprf_block_max_kernel(gprf_handle h, const double* X, const double* theta, int ncov, double* maxk) {
  if (!h || !X || !theta || !maxk) return GPRF_ERR_ARG;
  if (!h->have_structure) return GPRF_ERR_NO_STRUCTURE;
  CUDA_OK(cudaSetDevice(h->device));
  CovParams cp;
  if (make_cov(h, theta, ncov, &cp) != GPRF_OK) return GPRF_ERR_ARG;
  const size_t xb = (size_t)h->n * h->dx * sizeof(double);
  CUDA_OK(cudaMemcpy(h->dX, X, xb, cudaMemcpyHostToDevice));
  double* dM = nullptr;
  const size_t BB = (size_t)h->B * h->B;
  CUDA_OK(cudaMalloc((void**)&dM, BB * sizeof(double)));
  CUDA_OK(cudaMemset(dM, 0, BB * sizeof(double)));
#define CALL_MK(D, W) k_block_maxk<D, W><<<(unsigned)BB, 256>>>(h->dX, h->dx, h->dPerm, h->dBlockPtr, h->B, cp, dM)
  DISPATCH_COV(h, CALL_MK);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpy(maxk, dM, BB * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(dM);
  return GPRF_OK;
}
