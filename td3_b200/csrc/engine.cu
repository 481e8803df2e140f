// engine.cu -- host side of libtd3b200.so: builds the launch plan of one TD3 update
// (TD3_featured.py:123-171 / TD3_particles.py:167-224) as a list of "stages" (stage.cuh),
// captures it into CUDA graphs and exposes the C ABI declared in include/td3_b200.h.
//
// No CPU fallback exists anywhere in this file: every entry point launches sm_100a kernels
// or fails with an error code.
#include "../../include/td3_b200.h"

#include <cuda.h>   // CUtensorMap types only: the driver entry point is looked up at run time, nothing links libcuda

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "misc.cuh"
#include "stage.cuh"
#include "tc.cuh"
#include "front.cuh"
#include "apply.cuh"
#include "persist.cuh"
#ifndef TD3_NO_PIPE_KERNEL
#include "tcpipe.cuh"
#endif
#include "infer.cuh"
#include "enc.cuh"
#include "encbwd.cuh"
#include "chain.cuh"

using namespace td3;

constexpr int kMaxProgStages = 128;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                           \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess) return fail(TD3_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_));   \
  } while (0)

inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Per-device scratch allocations of the stand-alone entry points (rb_sample_indices, td3_gemm, set_encoder_*): keyed by
// (current device, slot), grown on demand, never shared between devices.  Calls on one stream are ordered; concurrent
// streams of one device must not use the same entry point at the same time (stated in td3_b200.h).
void* device_scratch(int slot, size_t bytes, bool zero = false) {
  static std::map<std::pair<int, int>, std::pair<void*, size_t>> pool;
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  auto& e = pool[{dev, slot}];
  if (e.second < bytes) {
    if (e.first) { cudaDeviceSynchronize(); cudaFree(e.first); }
    e.first = nullptr; e.second = 0;
    if (cudaMalloc(&e.first, bytes) != cudaSuccess) return nullptr;
    e.second = bytes;
    if (zero) cudaMemset(e.first, 0, bytes);
  }
  return e.first;
}

// ------------------------------------------------------------------------------------
// launch list
// ------------------------------------------------------------------------------------
struct Launch {
  enum Kind { STAGE, GATHER, LOSS, EW, TICK, HEAD, WN, FRONT, DPSYNC, ENC, CHAIN, ENCBWD_W2, ENCBWD_X } kind = STAGE;
  ChainParams chain{};
  DpSyncParams dpsync{};
  EncParams enc{};
  EncBwdParams encb{};
  HeadParams head{};
  WnParams wn{};
  FrontParams front{};
  int smem_bytes = 0;
  StageParams stage{};
  GatherParams gather{};
  LossParams loss{};
  EwParams ew{};
  DwParams dw{};        // EW launches: first-layer gradient tiles that step their own parameters (apply.cuh)
  AdamTick tick{};
  int grid_x = 1, grid_y = 1;
};

constexpr int kChainSmemMax = 232448 - 1024;     // dynamic shared memory of a chain CTA (static barriers take the rest)
constexpr int kChainSmemFixed = 1024 + kChX0Bytes + 2 * kChMaxW * 4 + 5 * kChRows * 4;   // alignment, X0, bias, head weights, reductions

extern int g_sm_count;

int ensure_kernel_attrs() {
  static bool done = false;
  if (done) return TD3_OK;
  CUDA_TRY(cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CUDA_TRY(cudaFuncSetAttribute(head_kernel_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CUDA_TRY(cudaFuncSetAttribute(front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFrontSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(front_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFrontWideSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(enc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEncSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(enc_bwd_w2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEbW2SmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(enc_bwd_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEbXSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmemMax));
  CUDA_TRY(cudaFuncSetAttribute(stage_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(stage_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(stage_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmemBytesSmall));
  CUDA_TRY(cudaFuncSetAttribute(stage_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmemBytes));
#ifndef TD3_NO_PIPE_KERNEL
  CUDA_TRY(cudaFuncSetAttribute(stage_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPipeSmemBytes));
#endif
  CUDA_TRY(cudaFuncSetAttribute(persistent_update_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(persistent_update_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmemBytes));
  done = true;
  return TD3_OK;
}

// every launch allows programmatic dependent launch of its successor (see misc.cuh: pdl_*)
template <typename Kernel, typename Params>
cudaError_t launch_pdl(Kernel k, dim3 grid, dim3 block, size_t smem, cudaStream_t s, const Params& p, int cluster = 1) {
  static const bool use_pdl = getenv("TD3_PDL") != nullptr;   // measured: no gain on B200 graphs (124.1 vs 121.9 us/update), off by default
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    cfg.gridDim.x = (grid.x + cluster - 1) / cluster * cluster;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, k, p);
}

// A fork of the graph being captured on `s`: work launched on g_fork.side after fork_begin() depends only on what `s` held
// at that point; fork_join() makes `s` wait for it again (run_launch before any launch that is not a plain stage,
// run_seq / the capture sites at their end: a capture must not end with unjoined work).
struct ForkState {
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool pending = false;
  int dev = -1;                 // device the stream and events belong to
};
thread_local ForkState g_fork;

// stream + events of the fork for the CURRENT device, created outside any capture (capture(), the prefix timer and every
// eager head launch call this); a thread that moves to another device gets new ones on its next eager launch
static int fork_ensure() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return TD3_ERR_CUDA;
  if (g_fork.side && g_fork.dev == dev) return TD3_OK;
  if (g_fork.pending) return TD3_ERR_STATE;
  if (g_fork.side) {            // (destroying another device's handles is legal from any device)
    cudaStreamDestroy(g_fork.side);
    cudaEventDestroy(g_fork.ev_fork);
    cudaEventDestroy(g_fork.ev_join);
    g_fork = ForkState{};
  }
  cudaStream_t st = nullptr;
  if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return TD3_ERR_CUDA;
  if (cudaEventCreateWithFlags(&g_fork.ev_fork, cudaEventDisableTiming) != cudaSuccess) return TD3_ERR_CUDA;
  if (cudaEventCreateWithFlags(&g_fork.ev_join, cudaEventDisableTiming) != cudaSuccess) return TD3_ERR_CUDA;
  g_fork.side = st;
  g_fork.dev = dev;
  return TD3_OK;
}

static int fork_begin(cudaStream_t s) {
  int dev = -1;
  cudaGetDevice(&dev);
  if (!g_fork.side || g_fork.pending || g_fork.dev != dev) return TD3_ERR_STATE;   // no fork: the caller launches behind the head instead
  if (cudaEventRecord(g_fork.ev_fork, s) != cudaSuccess) return TD3_ERR_CUDA;
  if (cudaStreamWaitEvent(g_fork.side, g_fork.ev_fork, 0) != cudaSuccess) return TD3_ERR_CUDA;
  g_fork.pending = true;
  return TD3_OK;
}

static cudaError_t fork_join(cudaStream_t s) {
  if (!g_fork.pending) return cudaSuccess;
  g_fork.pending = false;
  return cudaStreamWaitEvent(s, g_fork.ev_join, 0);
}

static bool stage_all_tc(const StageParams& S) {
  for (int q = 0; q < S.n_problems; ++q)
    if (!(S.p[q].kind == PK_GEMM && S.p[q].use_tc)) return false;
  return S.n_problems > 0;
}

int run_launch(const Launch& L, cudaStream_t s) {
  cudaError_t e = cudaSuccess;
  if (L.kind != Launch::STAGE) {
    e = fork_join(s);
    if (e != cudaSuccess) return fail(TD3_ERR_CUDA, "fork_join: %s", cudaGetErrorString(e));
  }
  switch (L.kind) {
    case Launch::STAGE: {
      if (L.stage.total_tiles <= 0) return TD3_OK;
      int rc = ensure_kernel_attrs();
      if (rc != TD3_OK) return rc;
      // fp32-only stages need the FFMA tile's buffers only: a small footprint lets the successor's CTAs co-reside
#ifndef TD3_NO_PIPE_KERNEL
      if (L.stage.any_tc && L.stage.pipe_tiles > 0)
        e = launch_pdl(stage_pipe_kernel, dim3(std::min(g_sm_count, L.stage.pipe_tiles)), dim3(kStageThreads), kPipeSmemBytes, s, L.stage);
      else
#endif
      if (L.stage.any_tc && L.stage.small_ring)    // many tiles per SM: the instance with the coalesced epilogue (small ring excludes clusters)
        e = launch_pdl(stage_kernel<true, true>, dim3(L.stage.total_tiles), dim3(kStageThreads), kDynSmemBytesSmall, s, L.stage);
      else if (L.stage.any_tc && stage_all_tc(L.stage) && !getenv("TD3_NO_LEAN"))
        e = launch_pdl(stage_kernel<true, false, true>, dim3(L.stage.total_tiles), dim3(kStageThreads), kDynSmemBytes, s, L.stage, L.stage.cluster);
      else if (L.stage.any_tc)
        e = launch_pdl(stage_kernel<true>, dim3(L.stage.total_tiles), dim3(kStageThreads), kDynSmemBytes, s, L.stage, L.stage.cluster);
      else e = launch_pdl(stage_kernel<false>, dim3(L.stage.total_tiles), dim3(kStageThreads), kSmemBytes + 1024, s, L.stage);
      break;
    }
    case Launch::GATHER: {
      // rows with a long 16-byte aligned segment (particle sets): bulk-copy staging through shared memory
      int bulk = 0;
      for (int i = 0; i < L.gather.n_seg; ++i) bulk |= L.gather.seg_len[i] >= kGatherBulkMinFloats;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(L.grid_x, L.grid_y); cfg.blockDim = dim3(256); cfg.stream = s;
      cfg.dynamicSmemBytes = bulk ? 8 * kGatherBulkBytes : 0;
      e = cudaLaunchKernelEx(&cfg, gather_kernel, L.gather, bulk);
      break;
    }
    case Launch::LOSS:
      e = launch_pdl(loss_kernel, dim3(L.grid_x), dim3(256), 0, s, L.loss);
      break;
    case Launch::EW:
      if (L.dw.n_tiles > 0) {
        static const bool use_pdl = getenv("TD3_PDL") != nullptr;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(L.dw.n_tiles + L.grid_x); cfg.blockDim = dim3(kDwThreads); cfg.dynamicSmemBytes = (size_t)dw_smem_bytes(L.dw.rows_per_pass); cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
        e = cudaLaunchKernelEx(&cfg, apply_kernel, L.ew, L.dw);
      } else {
        e = launch_pdl(adam_polyak_kernel, dim3(L.grid_x), dim3(kEwThreads), 0, s, L.ew);
      }
      break;
    case Launch::TICK:
      e = launch_pdl(adam_tick_kernel, dim3(1), dim3(32), 0, s, L.tick);
      break;
    case Launch::HEAD: {
      int rc = ensure_kernel_attrs();
      if (rc != TD3_OK) return rc;
      if (L.grid_x > g_sm_count && L.smem_bytes <= 100 * 1024 && !getenv("TD3_NO_WIDE_HEAD"))
        e = launch_pdl(head_kernel_wide, dim3(L.grid_x), dim3(kHeadThreads), (size_t)L.smem_bytes, s, L.head);
      else
        e = launch_pdl(head_kernel, dim3(L.grid_x), dim3(kHeadThreads), (size_t)L.smem_bytes, s, L.head);
      if (e == cudaSuccess && L.head.defer_finish) {
        // loss / host mirror / optimiser tick: beside the next stage when the stream is being captured (a fork of the
        // graph, joined by fork_join() before the next launch that is not a plain stage), behind the head otherwise
        const int n_agents = L.grid_x / std::max(1, L.head.n_cta);
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(s, &cs);
        cudaStream_t fs = s;
        if (cs != cudaStreamCaptureStatusActive) fork_ensure();
        else if (fork_begin(s) == TD3_OK) fs = g_fork.side;
        g_launches.fetch_add(1, std::memory_order_relaxed);
        head_finish_kernel<<<n_agents, 32, 0, fs>>>(L.head);
        e = cudaGetLastError();
        if (fs != s && e == cudaSuccess) e = cudaEventRecord(g_fork.ev_join, fs);
      }
      break;
    }
    case Launch::WN:
      e = launch_pdl(wn_kernel, dim3(L.grid_x), dim3(256), 0, s, L.wn);
      break;
    case Launch::FRONT: {
      int rc = ensure_kernel_attrs();
      if (rc != TD3_OK) return rc;
      if (L.front.job_groups > 0)
        e = launch_pdl(front_wide_kernel, dim3(L.grid_x), dim3(kFrontWideThreads), (size_t)front_wide_smem_bytes(L.front.rows_per_tile, L.front.head != 0, L.front.w_window), s, L.front);
      else
        e = launch_pdl(front_kernel, dim3(L.grid_x), dim3(256), (size_t)front_smem_bytes(L.front.rows_per_tile, L.front.head != 0), s, L.front);
      break;
    }
    case Launch::DPSYNC:
      dp_signal_wait_kernel<<<1, 32, 0, s>>>(L.dpsync);
      e = cudaGetLastError();
      break;
    case Launch::ENC: {
      int rc = ensure_kernel_attrs();
      if (rc != TD3_OK) return rc;
      e = launch_pdl(enc_fwd_kernel, dim3(L.grid_x), dim3(kEncThreads), (size_t)kEncSmemBytes, s, L.enc);
      break;
    }
    case Launch::ENCBWD_W2: {
      int rc = ensure_kernel_attrs();
      if (rc != TD3_OK) return rc;
      e = launch_pdl(enc_bwd_w2_kernel, dim3(L.grid_x), dim3(kEbThreads), (size_t)kEbW2SmemBytes, s, L.encb);
      break;
    }
    case Launch::ENCBWD_X: {
      int rc = ensure_kernel_attrs();
      if (rc != TD3_OK) return rc;
      e = launch_pdl(enc_bwd_x_kernel, dim3(L.grid_x), dim3(kEbXThreads), (size_t)kEbXSmemBytes, s, L.encb);
      break;
    }
    case Launch::CHAIN: {
      int rc = ensure_kernel_attrs();
      if (rc != TD3_OK) return rc;
      e = launch_pdl(chain_kernel, dim3(L.chain.n_ctas), dim3(kChThreads), (size_t)L.smem_bytes, s, L.chain);
      break;
    }
  }
  if (e != cudaSuccess) return fail(TD3_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CUDA_TRY(cudaGetLastError());
  return TD3_OK;
}

// A stage under construction: problems that may run concurrently.
using ProblemList = std::vector<Problem>;

Problem blank_problem(int kind) {
  Problem p;
  memset(&p, 0, sizeof(p));
  p.kind = kind;
  p.groups_inner = 1;
  p.c_dups = 1;
  p.ksplit = 1;
  p.map_a = p.map_b = -1;
  return p;
}

struct GroupShape { int n_outer, n_inner; };

Problem make_gemm(int M, int N, int K, const float* A, int lda, bool a_rc, const float* B, int ldb, bool b_rc, float* C,
                  int ldc, int epi) {
  Problem p = blank_problem(PK_GEMM);
  p.epi = epi;
  p.M = M; p.N = N; p.K = K;
  p.A = A; p.lda = lda; p.a_rc = a_rc;
  p.B = B; p.ldb = ldb; p.b_rc = b_rc;
  p.C = C; p.ldc = ldc;
  return p;
}

// ------------------------------------------------------------------------------------
// TMA tensor maps for the tcgen05 tile's operands (tc.cuh).  One map per (operand, group).
// ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Appends `groups` maps for operand A (is_a) or B of a tensor-core GEMM problem; false when the operand cannot be
// described to the TMA unit (rows not 16-byte aligned): the tile then stages it with cp.async instead.
bool encode_operand_maps(const Problem& p, bool is_a, int n_outer, int n_inner, std::vector<CUtensorMap>& out) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const float* base = is_a ? p.A : p.B;
  const long long ld = is_a ? p.lda : p.ldb, go = is_a ? p.a_go : p.b_go, gi = is_a ? p.a_gi : p.b_gi;
  const int rc = is_a ? p.a_rc : p.b_rc;
  const long long O = is_a ? p.M : p.N;
  if (!aligned16(base) || (ld & 3) || (go & 3) || (gi & 3) || ld <= 0) return false;
  const size_t first = out.size();
  for (int o = 0; o < n_outer; ++o)
    for (int i = 0; i < n_inner; ++i) {
      CUtensorMap m;
      cuuint64_t gdim[2], gstride[1] = {(cuuint64_t)ld * 4};
      cuuint32_t box[2], estr[2] = {1, 1};
      CUtensorMapSwizzle sw;
      if (rc) {        // memory [O rows][K cols]: K-major atoms, box = 32 reduction steps x (128 | NT) rows
        gdim[0] = (cuuint64_t)p.K; gdim[1] = (cuuint64_t)O;
        box[0] = 32; box[1] = is_a ? (cuuint32_t)(128 / std::max(1, p.tc_cluster)) : (cuuint32_t)p.tc_nt;
        sw = CU_TENSOR_MAP_SWIZZLE_128B;
      } else {         // memory [K rows][O cols]: MN-major atoms with 32-byte swizzle base, box = 32 MN x 32 reduction rows
        gdim[0] = (cuuint64_t)O; gdim[1] = (cuuint64_t)p.K;
        box[0] = 32; box[1] = 32;
        sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
      }
      void* addr = const_cast<float*>(base + o * go + i * gi);
      CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, addr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        out.resize(first);
        return false;
      }
      out.push_back(m);
    }
  return true;
}

constexpr int kMaxTensorMaps = 8192;    // (128 bytes each; a 64-agent population needs ~2500)

// Give every tensor-core problem of `seqs` its maps; `dev_maps` is the device array they will be copied to.
bool attach_tensor_maps(std::initializer_list<std::vector<Launch>*> seqs, CUtensorMap* dev_maps, std::vector<CUtensorMap>& host) {
  bool ok_all = true;
  for (auto* seq : seqs)
    for (Launch& L : *seq) {
      if (L.kind != Launch::STAGE) continue;
      int n_param = 0;        // maps of this launch that also travel in its kernel parameters
      for (int q = 0; q < L.stage.n_problems; ++q) {
        Problem& p = L.stage.p[q];
        p.tmapA = p.tmapB = nullptr;
        p.map_a = p.map_b = -1;
        if (p.kind != PK_GEMM || !p.use_tc) continue;
        const int groups = p.tile_count / std::max(1, p.tiles_per_group);
        const int n_inner = std::max(1, p.groups_inner), n_outer = std::max(1, groups / n_inner);
        const size_t ia = host.size();
        if (!encode_operand_maps(p, true, n_outer, n_inner, host)) { ok_all = false; continue; }
        const size_t ib = host.size();
        if (!encode_operand_maps(p, false, n_outer, n_inner, host) || (int)host.size() > kMaxTensorMaps) {
          host.resize(ia);
          ok_all = false;
          continue;
        }
        p.tmapA = dev_maps + ia;
        p.tmapB = dev_maps + ib;
        const int na = (int)(ib - ia), nb = (int)(host.size() - ib);
        if (n_param + na + nb <= kStageMaps && !getenv("TD3_NO_PARAM_MAPS")) {
          static_assert(sizeof(CUtensorMap) == sizeof(TensorMapBlob), "tensor map blob size");
          memcpy(&L.stage.maps[n_param], &host[ia], (size_t)na * sizeof(CUtensorMap));
          p.map_a = n_param;
          n_param += na;
          memcpy(&L.stage.maps[n_param], &host[ib], (size_t)nb * sizeof(CUtensorMap));
          p.map_b = n_param;
          n_param += nb;
        }
      }
    }
  return ok_all;
}


int g_sm_count = 148;
// tensor-core policy of the plan being built: 0 = fp32 FFMA tiles only, 1 = TF32 tcgen05 tiles where eligible
thread_local int g_tc_mode = 0;

void finalize_problem(Problem& p, GroupShape gs) {
  p.groups_inner = gs.n_inner;
  const int groups = gs.n_outer * gs.n_inner;
  switch (p.kind) {
    case PK_GEMM: {
      auto ok4 = [](long long v) { return (v & 3) == 0; };
      p.a_vec = aligned16(p.A) && ok4(p.lda) && ok4(p.a_go) && ok4(p.a_gi) && ok4(p.a_rc ? p.K : p.M);
      p.b_vec = aligned16(p.B) && ok4(p.ldb) && ok4(p.b_go) && ok4(p.b_gi) && ok4(p.b_rc ? p.K : p.N);
      p.c_vec = aligned16(p.C) && ok4(p.ldc) && ok4(p.c_go) && ok4(p.c_gi) && ok4(p.c_split) && ok4(p.c_dup_stride);
      p.aux_vec = p.aux0 && aligned16(p.aux0) && ok4(p.ldaux) && ok4(p.aux0_go) && ok4(p.aux0_gi);
      // TF32 mode: contractions whose operands the TMA unit can address (16-byte aligned rows) and whose reduction
      // is long enough to fill a 128-row tensor-core tile; no fused row-sum there (PK_COLSUM does the bias gradient)
      auto tma_ok = [&](const float* base, long long ld, long long go_, long long gi_) {
        return aligned16(base) && ok4(ld) && ok4(go_) && ok4(gi_) && ld > 0;
      };
      p.use_tc = g_tc_mode && p.K >= 64 && !(p.aux1 && p.epi == EPI_STORE) && tma_ok(p.A, p.lda, p.a_go, p.a_gi) &&
                 tma_ok(p.B, p.ldb, p.b_go, p.b_gi) && encode_tiled_fn() != nullptr && !getenv("TD3_NO_TMA");
      if (!p.use_tc && p.B_master) { p.B = p.B_master; p.B_master = nullptr; }   // FFMA tiles read the fp32 master weights
      if (p.use_tc) {
        p.tiles_m = (p.M + 127) / 128;
        int nt = p.N <= 16 ? 16 : 32;
        for (int cand : {64, 128}) {   // wider tiles only when there are plenty of them (A is re-read once per N tile)
          const long long tiles = (long long)p.tiles_m * ((p.N + cand - 1) / cand) * p.ksplit * groups;
          if (p.N > nt && tiles >= 2LL * g_sm_count) nt = cand;   // (from 1 or 0.6 tiles per SM: measured no gain on 8 / 32-agent populations)
        }
        p.tc_nt = nt;
        p.tiles_n = (p.N + nt - 1) / nt;
      } else {
        p.tiles_m = (p.M + kBM - 1) / kBM;
        p.tiles_n = (p.N + kBN - 1) / kBN;
      }
      p.tiles_per_group = p.tiles_m * p.tiles_n * p.ksplit;
      break;
    }
    case PK_COLSUM:
      p.tiles_n = (p.N + 31) / 32;
      p.tiles_per_group = p.tiles_n * p.ksplit;
      break;
    case PK_LN_FWD:
    case PK_LN_BWD_ROWS:
      p.tiles_per_group = (p.M + 7) / 8;
      break;
    case PK_LN_BWD_COLS:
      p.tiles_per_group = (p.N + 31) / 32;
      break;
    case PK_POOL_FWD:
      p.tiles_n = 1;
      p.tiles_per_group = p.M;          // one CTA per sample, all channels (<= 256)
      break;
    case PK_SMALLK_FWD:
      p.tiles_per_group = (p.M + 127) / 128;
      break;
    case PK_SMALLK_DW:
      p.tiles_per_group = p.ksplit;
      break;
    case PK_POOL_BWD:
      p.tiles_per_group = (int)(((long long)p.M * p.K + kPoolBwdRows - 1) / kPoolBwdRows);
      break;
    case PK_REDUCE_SPLITS:
      p.tiles_per_group = (p.M + kReduceTile - 1) / kReduceTile;
      break;
    case PK_NEG_MEAN:
      p.tiles_per_group = 1;
      break;
    case PK_ENC_FUSED:
    case PK_ENC_BWD_W2:
    case PK_ENC_BWD_X:
      p.tiles_per_group = p.M / kEncTile;
      break;
  }
  p.tile_count = p.tiles_per_group * groups;
}

// thread-local planning switches: tensor-core policy is declared further down next to finalize_problem
thread_local int g_cluster_mode = 0;     // 1: stage launches may use thread-block clusters (A-panel multicast)


// Final tile layout of a stage launch: tensor-core problems first, one cluster size for the launch, every TC
// problem's N-tile count padded to a multiple of it (padding tiles fetch their share of the A panel and store nothing).
void layout_stage(Launch& L) {
  StageParams& S = L.stage;
  const int n = S.n_problems;
  std::stable_partition(S.p, S.p + n, [](const Problem& p) { return p.kind == PK_GEMM && p.use_tc; });
  // Waves: a tensor-core launch runs one CTA per SM, so tile number sm_count + 1 waits for a whole tile to finish, and
  // the K loop of a tile costs the same for every N width.  When a launch is over one wave, widen
  // the N tiles of its most numerous tensor-core problems as long as that removes whole waves.
  if (!getenv("TD3_NO_WAVE_FIT")) {
    auto count = [&](const int* nt) {
      long long total = 0;
      for (int q = 0; q < n; ++q) {
        const Problem& p = S.p[q];
        if (p.kind == PK_GEMM && p.use_tc) {
          const int g = p.tile_count / std::max(1, p.tiles_per_group);
          total += (long long)g * p.tiles_m * ((p.N + nt[q] - 1) / nt[q]) * p.ksplit;
        } else {
          total += p.tile_count;
        }
      }
      return total;
    };
    int nt[kMaxProblemsPerStage];
    bool any = false;
    for (int q = 0; q < n; ++q) { nt[q] = S.p[q].tc_nt; any |= S.p[q].kind == PK_GEMM && S.p[q].use_tc; }
    if (any && count(nt) > g_sm_count) {
      // greedy walk towards wider tiles; keep the first configuration with the fewest waves
      auto waves = [&](const int* v) { return (count(v) + g_sm_count - 1) / g_sm_count; };
      int best_nt[kMaxProblemsPerStage];
      std::copy(nt, nt + n, best_nt);
      long long best_waves = waves(nt);
      while (best_waves > 1) {
        int best = -1;
        long long best_tiles = 0;
        for (int q = 0; q < n; ++q) {
          const Problem& p = S.p[q];
          if (!(p.kind == PK_GEMM && p.use_tc) || nt[q] >= 128 || p.N <= nt[q]) continue;
          const long long t = (long long)(p.tile_count / std::max(1, p.tiles_per_group)) * p.tiles_m * ((p.N + nt[q] - 1) / nt[q]) * p.ksplit;
          if (t > best_tiles) { best_tiles = t; best = q; }
        }
        if (best < 0) break;
        nt[best] *= 2;
        if (waves(nt) < best_waves) {
          best_waves = waves(nt);
          std::copy(nt, nt + n, best_nt);
        }
      }
      std::copy(best_nt, best_nt + n, nt);
      for (int q = 0; q < n; ++q) {
          Problem& p = S.p[q];
          if (!(p.kind == PK_GEMM && p.use_tc) || nt[q] == p.tc_nt) continue;
          const int g = p.tile_count / std::max(1, p.tiles_per_group);
          p.tc_nt = nt[q];
          p.tiles_n = (p.N + p.tc_nt - 1) / p.tc_nt;
          p.tiles_per_group = p.tiles_m * p.tiles_n * p.ksplit;
          p.tile_count = p.tiles_per_group * g;
        }
    }
  }
  int groups[kMaxProblemsPerStage];
  long long other_tiles = 0;
  bool any_tc = false;
  for (int q = 0; q < n; ++q) {
    Problem& p = S.p[q];
    groups[q] = p.tile_count / std::max(1, p.tiles_per_group);
    if (p.kind == PK_GEMM && p.use_tc) any_tc = true;
    else other_tiles += p.tile_count;
  }
  int c = 1;
  if (g_cluster_mode && any_tc) {
    auto total_for = [&](int cand) {
      long long total = other_tiles;
      for (int q = 0; q < n; ++q) {
        const Problem& p = S.p[q];
        if (!(p.kind == PK_GEMM && p.use_tc)) continue;
        const long long tn = ((p.N + p.tc_nt - 1) / p.tc_nt + cand - 1) / cand * cand;
        total += (long long)groups[q] * p.tiles_m * tn * p.ksplit;
      }
      return total;
    };
    const long long base = total_for(1);
    // single wave: padding tiles are free (idle SMs) and help fetch the A panel; several waves: only if padding is <= 10 %
    for (int cand : {8, 4, 2}) {
      const long long t = total_for(cand);
      if (t <= g_sm_count || (base > g_sm_count && t * 10 <= base * 11)) { c = cand; break; }
    }
  }
  int tiles = 0;
  for (int q = 0; q < n; ++q) {
    Problem& p = S.p[q];
    if (p.kind == PK_GEMM && p.use_tc) {
      p.tc_cluster = c;
      p.tiles_n = ((p.N + p.tc_nt - 1) / p.tc_nt + c - 1) / c * c;
      p.tiles_per_group = p.tiles_m * p.tiles_n * p.ksplit;
      p.tile_count = p.tiles_per_group * groups[q];
    }
    p.tile_begin = tiles;
    tiles += p.tile_count;
  }
  S.total_tiles = tiles;
  S.any_tc = any_tc ? 1 : 0;
  S.cluster = c;
  // many tiles per SM (particle-encoder backward, batch-8192 data-parallel update): half the ring, two CTAs per SM
  {
    const char* thr = getenv("TD3_SMALL_RING_TILES");       // tiles per SM from which the small ring is used (tuning knob)
    const double per_sm = thr ? atof(thr) : 1.5;
    S.small_ring = (any_tc && c == 1 && tiles >= per_sm * g_sm_count && !getenv("TD3_NO_SMALL_RING")) ? 1 : 0;
    // ... or, when every tensor-core problem keeps its whole reduction in one tile, one persistent CTA per SM whose
    // producer / MMA / epilogue warps are pipelined across tiles (tcpipe.cuh)
    S.pipe_tiles = 0;
    // Measured (profiles/r02g_pipe_kernel_prefix_times.txt): NOT faster than two small-ring CTAs per SM -- the forward stages
    // of a population are bound by L2 -> shared-memory operand traffic (fp32 operands: 32 flop per byte at 128 x 128), and a
    // backward stage's column-sum / FFMA tiles run after the pipelined tiles instead of beside them -- so it is opt-in.
    const char* pthr = getenv("TD3_PIPE_TILES");            // tiles per SM from which the pipelined kernel is used (tuning knob)
    const double pipe_per_sm = pthr ? atof(pthr) : 1.5;
    if (any_tc && c == 1 && getenv("TD3_PIPE")) {
      int tc_tiles = 0;
      bool ok = true;
      for (int q = 0; q < n; ++q) {
        const Problem& p = S.p[q];
        if (p.kind == PK_GEMM && p.use_tc) {
          ok = ok && p.ksplit == 1 && p.tile_begin == tc_tiles && p.K >= 32;
          tc_tiles += p.tile_count;
        }
      }
      if (ok && tc_tiles >= pipe_per_sm * g_sm_count) { S.pipe_tiles = tc_tiles; S.small_ring = 0; }
    }
  }
}

// the fused set-encoder forward of one pass (enc.cuh) as a launch: the problem record carries
//   A = particles, B / bias = conv1 weight / bias (fp32 masters), aux0 = conv2 weight (the TF32 copy when there is one),
//   aux1 = conv2 bias, aux2 / aux3 = optional h1 / h2 outputs, C = partial means, M = rows, K = D, N = particles per sample
bool make_enc_launch(const Problem& p, Launch& L) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  L = Launch{};
  L.kind = Launch::ENC;
  EncParams& E = L.enc;
  memset(&E, 0, sizeof(E));
  const int groups = p.tile_count / std::max(1, p.tiles_per_group);
  const int n_inner = std::max(1, p.groups_inner), n_outer = std::max(1, groups / n_inner);
  if (groups > kEncMaxGroups) return false;
  E.P = p.A; E.p_go = p.a_go;
  E.W1 = p.B; E.b1 = p.bias; E.b2 = p.aux1; E.w_go = p.b_go; E.w_gi = p.b_gi;
  E.h1 = p.aux2; E.h1_go = p.aux2_go; E.h1_gi = p.aux2_gi;
  E.h2 = p.aux3; E.h2_go = p.aux3_go; E.h2_gi = p.aux3_gi;
  E.part = p.C; E.part_go = p.c_go; E.part_gi = p.c_gi;
  E.bits = reinterpret_cast<unsigned int*>(const_cast<void*>(p.tmapA)); E.bits_go = p.aux2_go; E.bits_gi = p.aux2_gi;   // (see build_forward)
  if (E.bits) { E.h1 = nullptr; E.h2 = nullptr; }
  E.rows = p.M; E.D = p.K; E.n_inner = n_inner; E.n_groups = groups; E.tiles_per_group = p.tiles_per_group;
  for (int o = 0; o < n_outer; ++o)
    for (int i = 0; i < n_inner; ++i) {
      CUtensorMap m;
      cuuint64_t gdim[2] = {(cuuint64_t)kEncH, (cuuint64_t)kEncO}, gstride[1] = {(cuuint64_t)kEncH * 4};
      cuuint32_t box[2] = {32, (cuuint32_t)kEncO}, estr[2] = {1, 1};
      void* addr = const_cast<float*>(p.aux0 + o * p.aux0_go + i * p.aux0_gi);
      if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, addr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
      static_assert(sizeof(CUtensorMap) == sizeof(TensorMapBlob), "tensor map blob size");
      memcpy(&E.w2_map[o * n_inner + i], &m, sizeof(m));
    }
  L.grid_x = std::min(g_sm_count, groups * p.tiles_per_group);
  return true;
}

// the fused set-encoder backward (encbwd.cuh) as a launch: the problem record carries
//   A = particles, B / bias = conv1 weight / bias, aux0 = conv2 weight (TF32 shadow), aux1 / lda = d(pooled), aux2 / ldb = pooled,
//   tmapA = bitmaps (strides in aux3_go / aux3_gi), C = split-K partials, M = rows, K = D, N = particles per sample,
//   ksplit = CTAs per group (per channel half for the X kernel), c_split = offset of the kernel's region in the partial buffer
Launch make_enc_bwd_launch(const Problem& p) {
  Launch L;
  L.kind = p.kind == PK_ENC_BWD_W2 ? Launch::ENCBWD_W2 : Launch::ENCBWD_X;
  EncBwdParams& E = L.encb;
  memset(&E, 0, sizeof(E));
  const int groups = p.tile_count / std::max(1, p.tiles_per_group);
  E.P = p.A; E.p_go = p.a_go;
  E.W1 = p.B; E.b1 = p.bias; E.W2 = p.aux0; E.w_go = p.b_go; E.w_gi = p.b_gi;
  E.dpool = p.aux1; E.ld_dpool = p.lda; E.dpool_go = p.aux1_go; E.dpool_gi = p.aux1_gi;
  E.pooled = p.aux2; E.ld_pooled = p.ldb; E.pooled_go = p.aux2_go; E.pooled_gi = p.aux2_gi;
  E.bits = reinterpret_cast<const unsigned int*>(p.tmapA); E.bits_go = p.aux3_go; E.bits_gi = p.aux3_gi;
  E.part = p.C + p.c_split; E.part_go = p.c_go; E.part_gi = p.c_gi;
  E.rows = p.M; E.D = p.K; E.n_particles = p.N; E.n_inner = std::max(1, p.groups_inner); E.n_groups = groups;
  E.tiles_per_group = p.tiles_per_group; E.ks = p.ksplit;
  E.inv_n = 1.f / (float)p.N;
  L.grid_x = groups * p.ksplit * (p.kind == PK_ENC_BWD_X ? 2 : 1);
  return L;
}

int emit_stage(std::vector<Launch>& seq, const ProblemList& probs_in) {
  ProblemList probs;
  for (const Problem& p : probs_in) {
    if (p.kind == PK_ENC_FUSED) {
      Launch L;
      if (!make_enc_launch(p, L)) return fail(TD3_ERR_CUDA, "fused set-encoder: tensor map for conv2 could not be encoded");
      seq.push_back(L);
    } else if (p.kind == PK_ENC_BWD_W2 || p.kind == PK_ENC_BWD_X) {
      seq.push_back(make_enc_bwd_launch(p));
    } else {
      probs.push_back(p);
    }
  }
  // split into chunks of kMaxProblemsPerStage (they stay independent, so extra launches are still correct)
  size_t i = 0;
  while (i < probs.size()) {
    Launch L;
    L.kind = Launch::STAGE;
    int n = 0, tiles = 0;
    while (i < probs.size() && n < kMaxProblemsPerStage) {
      Problem p = probs[i++];
      if (p.tile_count <= 0) continue;
      p.tile_begin = tiles;
      tiles += p.tile_count;
      if (p.kind == PK_GEMM && p.use_tc) L.stage.any_tc = 1;
      L.stage.p[n++] = p;
    }
    L.stage.n_problems = n;
    L.stage.total_tiles = tiles;
    if (n > 0) {
      layout_stage(L);
      seq.push_back(L);
    }
  }
  return TD3_OK;
}

// fold the problems of stage launch `src` into `dst` (both STAGE launches of mutually independent work)
bool merge_stage(Launch& dst, const Launch& src) {
  if (dst.kind != Launch::STAGE || src.kind != Launch::STAGE) return false;
  if (dst.stage.n_problems + src.stage.n_problems > kMaxProblemsPerStage) return false;
  for (int q = 0; q < src.stage.n_problems; ++q) {
    Problem p = src.stage.p[q];
    p.tile_begin = dst.stage.total_tiles;
    dst.stage.total_tiles += p.tile_count;
    dst.stage.p[dst.stage.n_problems++] = p;
  }
  dst.stage.any_tc |= src.stage.any_tc;
  layout_stage(dst);
  return true;
}

// merge several per-pass stage sequences so that stage s of every pass shares one launch
std::vector<ProblemList> zip_stages(const std::vector<std::vector<ProblemList>>& passes) {
  size_t depth = 0;
  for (auto& p : passes) depth = std::max(depth, p.size());
  std::vector<ProblemList> out(depth);
  for (auto& p : passes)
    for (size_t s = 0; s < p.size(); ++s) out[s].insert(out[s].end(), p[s].begin(), p[s].end());
  return out;
}

// ------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------
struct Bump {
  float* base = nullptr;
  long long used = 0;
  std::map<std::string, std::pair<long long, long long>> regions;
  float* take(long long n, const char* name = nullptr) {
    n = std::max<long long>(n, 1);
    const long long off = used;
    used += round_up(n, 64);
    if (name) regions[name] = {off, n};
    return base ? base + off : nullptr;
  }
};

// Activations of one forward pass over `groups` networks of identical shape.
struct PassBuf {
  float* x0 = nullptr; int ld0 = 0; long long x0_go = 0, x0_gi = 0;    // trunk input [B, ld0]
  float* x0n = nullptr; float* mean0 = nullptr; float* rstd0 = nullptr; // input LayerNorm (lnorm1)
  long long x0n_go = 0, x0n_gi = 0;
  float* r[TD3_MAX_LINEAR] = {};      // relu(z_l)            [B, w_l]
  float* n[TD3_MAX_LINEAR] = {};      // LN(relu(z_l))        (norm == layer)
  float* mean[TD3_MAX_LINEAR] = {};
  float* rstd[TD3_MAX_LINEAR] = {};
  long long h_go[TD3_MAX_LINEAR] = {}, h_gi[TD3_MAX_LINEAR] = {};
  long long s_go = 0, s_gi = 0;       // strides of the [B] statistics
  const float* P = nullptr; long long P_go = 0;       // particles [B*N, D]
  float* h1 = nullptr; float* h2 = nullptr;           // encoder activations
  long long h1_go = 0, h1_gi = 0, h2_go = 0, h2_gi = 0;
  float* part = nullptr; long long part_go = 0, part_gi = 0;   // fused encoder: partial means [rows / 128, enc_out]
  unsigned int* enc_bits = nullptr; long long bits_go = 0, bits_gi = 0;   // ReLU bitmaps instead of h1 / h2 (fused backward, encbwd.cuh)
  bool keep_enc_acts = true;                          // false: nothing reads h1 / h2 after the forward pass (target networks, Q1)
};

struct OutSpec {
  float* out = nullptr; int ld = 0; long long go = 0, gi = 0;
  int epi = EPI_BIAS;
  float* aux0 = nullptr; int ldaux = 0; long long aux0_go = 0, aux0_gi = 0;
  float f0 = 1.f, f1 = 0.f;
  int dups = 1; long long dup_stride = 0;
};

struct ParamRef {   // packed parameters of `n_inner` networks per agent
  const float* base = nullptr;
  long long go = 0, gi = 0;
  const float* tc = nullptr;   // TF32 mode: round-to-nearest shadow of `base` (same layout) for the tensor-core tiles
};

// weight matrix at `off` as the B operand of a contraction: the TF32 shadow when there is one (finalize_problem
// switches back to the master copy if the problem ends up on the FFMA tile)
void weight_operand(Problem& p, const ParamRef& W, long long off) {
  p.B = W.base + off;
  p.B_master = nullptr;
  if (W.tc) { p.B = W.tc + off; p.B_master = W.base + off; }
}

struct GradRef {
  float* base = nullptr;
  long long go = 0, gi = 0;
};

}  // namespace

struct td3_agent {
  td3_agent_config cfg{};
  td3_param_set actor{}, critic{};
  bool params_bound = false;
  unsigned long long* state_u64 = nullptr;   // [0] sample step [1] critic Adam t [2] actor Adam t [3] rb size
  float* state_f32 = nullptr;                // critic_loss[n_agents], actor_loss[n_agents]
  unsigned long long* host_status = nullptr; // mapped pinned host words {loss bits, sequence}[n_agents], or nullptr
  bool host_status_live = false;             // the planned update writes them (fused critic head)
  unsigned int* head_seq = nullptr;          // per-agent count of critic updates completed under the current plan
  long long batch = 0, global_batch = 0, batch_offset = 0;
  Bump ws;
  long long ws_floats = 0;

  // --- workspace pointers (valid after plan) ---
  int ld_a = 0, ld_q = 0, in_a = 0, in_q = 0, qw = 1;
  float *xa = nullptr, *xa2 = nullptr, *xq = nullptr, *xq2 = nullptr, *xpi = nullptr, *xapi = nullptr;
  long long xa_go = 0, xq_go = 0, xq_gi = 0, xpi_go = 0;
  float *P = nullptr, *P2 = nullptr;
  float *r = nullptr, *nd = nullptr, *eps = nullptr, *noise_in = nullptr;
  long long *idx = nullptr, *idx_in = nullptr;
  float *q = nullptr, *tq = nullptr, *y = nullptr, *dq = nullptr, *q_pi = nullptr, *dq_pi = nullptr, *tanh_y = nullptr;
  PassBuf pb_at, pb_ct, pb_c, pb_a, pb_q1;
  // weight normalisation: effective parameters (W = g v / ||v|| in every weight_v slot) the contractions read
  float *eff_a = nullptr, *eff_at = nullptr, *eff_c = nullptr, *eff_ct = nullptr;
  // TF32 mode without weight normalisation: round-to-nearest shadows of the four packed parameter buffers, read by the
  // tensor-core tiles (tcgen05 truncates its operands; a rounded operand is read exactly).  Kept current by the
  // optimiser kernels (Adam / Polyak / apply write master and shadow together); rebuilt from the masters when the
  // caller says the parameters changed under us (td3_agent_params_changed) and after every plan.
  float *sh_a = nullptr, *sh_at = nullptr, *sh_c = nullptr, *sh_ct = nullptr;
  bool shadow_dirty = true;
  // data-parallel update over peer-mapped gradient buffers (td3_dp_bind_peers): the Adam launches sum the ranks' gradients
  int dp_world = 0, dp_rank = 0;
  bool dp_fused = false;
  const float* dp_critic_grads[kMaxPeers] = {};
  const float* dp_actor_grads[kMaxPeers] = {};
  unsigned int* dp_flags[kMaxPeers] = {};
  // row-local front kernels (front.cuh): template of the sampling launch (gather + first layers), filled by plan_sample
  bool front_on = false;
  FrontParams front_sample{};

  std::vector<Launch> seq_sample, seq_target, seq_critic_fb, seq_critic_apply, seq_actor_fb, seq_actor_apply;
  // fused middle of a policy update (critic backward with the actor forward riding along, critic Adam, rest of the
  // actor step): what the CUDA graph and the persistent program run instead of critic_fb + critic_apply + actor_fb
  std::vector<Launch> seq_policy_mid;
  // what train_n runs after seq_sample (CUDA graph and plain-launch modes): either the sequences above back to back, or
  // -- plain MLPs -- the tail-fused form: first-layer gradient + Adam (+ Polyak) in ONE launch (apply.cuh) and the
  // actor's forward layers riding along with the target pass
  std::vector<Launch> seq_run_critic, seq_run_policy;
  // layer-fused chains (chain.cuh): what train_n runs INSTEAD of seq_sample + seq_run_* when the shapes allow it
  // (the first launch of each samples the batch itself: plan_sample fills its replay view)
  std::vector<Launch> seq_chain_critic, seq_chain_policy;
  bool chain_on = false;
  bool tail_fused = false;
  int n_actor_fwd = 0;                       // leading launches of seq_actor_fb that are the actor's forward pass
  const float* plan_rows = nullptr;
  long long plan_row_stride = 0, plan_rb_agent_stride = 0;
  int plan_rng_mode = -1;

  struct Graphs {
    cudaGraphExec_t critic_only = nullptr, with_actor = nullptr;
    long long nodes_critic_only = 0, nodes_with_actor = 0;     // kernels per replay (launch accounting)
    const float* rows = nullptr;
    long long row_stride = 0;
    int rng_mode = -1;
  } graphs;
  long long last_rb_size = -1;
  cudaStream_t cap_stream = nullptr;
  // persistent-kernel programs (device copies live in the workspace region "program")
  StageRec* prog_dev = nullptr;
  long long* prof_dev = nullptr;
  CUtensorMap* tmaps_dev = nullptr;
  int n_prog_critic = 0, n_prog_policy = 0, n_bar_critic = 0, n_bar_policy = 0;
  bool prog_dirty = true;
  int persist_grid = 0;
  int cluster_mode = 1;                      // stage launches use clusters (graph / launches modes); the persistent kernel does not
  unsigned int bar_count = 0;               // value of the device barrier counter once all queued launches finish
  bool bar_reset = true;                    // zero the counter before the next launch (fresh state block)         // capture happens here: the caller's stream may be the legacy stream
};

namespace {

void drop_graphs(td3_agent* a) {
  if (a->graphs.critic_only) cudaGraphExecDestroy(a->graphs.critic_only);
  if (a->graphs.with_actor) cudaGraphExecDestroy(a->graphs.with_actor);
  a->graphs = td3_agent::Graphs{};
}

// ------------------------------------------------------------------------------------
// forward pass builder
// ------------------------------------------------------------------------------------
void set_groups(Problem& p, long long a_go, long long a_gi, long long b_go, long long b_gi, long long c_go, long long c_gi) {
  p.a_go = a_go; p.a_gi = a_gi; p.b_go = b_go; p.b_gi = b_gi; p.c_go = c_go; p.c_gi = c_gi;
}

// K6 (enc.cuh / encbwd.cuh) applies: TF32 mode, the reference's 256 / 128 encoder widths, whole 128-particle tiles per sample
bool enc_fusable(const td3_agent_config& cfg, const td3_net_layout& net, GroupShape gs, int B) {
  return cfg.variant == TD3_VARIANT_PARTICLES && g_tc_mode && net.enc_hidden == kEncH && net.enc_out == kEncO &&
         cfg.particle_dim <= 7 && ((long long)B * cfg.n_particles) % kEncTile == 0 && cfg.n_particles % kEncTile == 0 &&
         gs.n_outer * gs.n_inner <= kEncMaxGroups && encode_tiled_fn() != nullptr && !getenv("TD3_NO_ENC_FUSION");
}

std::vector<ProblemList> build_forward(const td3_agent_config& cfg, const td3_net_layout& net, ParamRef W, GroupShape gs,
                                       int B, const PassBuf& pb, const OutSpec& out, int pool_dups = 1,
                                       long long pool_dup_stride = 0, bool skip_last = false, bool skip_first = false) {
  // skip_last: the output layer is computed by the fused head kernel (misc.cuh: head_body) or a front kernel
  // skip_first: the first layer's relu(x W^T + b) is computed by a front kernel (front.cuh)
  std::vector<ProblemList> st;
  const bool ln = cfg.norm == TD3_NORM_LAYER;
  const bool enc = cfg.variant == TD3_VARIANT_PARTICLES;
  const int L = net.n_linear;
  // TF32 mode: whatever a later tensor-core contraction reads as an operand is stored rounded to nearest TF32 by its
  // producer (stage.cuh: Problem::rn_out), and weights come from the rounded shadows (ParamRef::tc)
  const int tf = g_tc_mode ? 1 : 0;
  // K6 (enc.cuh): layer 1 -> layer 2 on tcgen05 -> partial pooling in ONE persistent launch, activations on chip
  const bool fuse_enc = enc_fusable(cfg, net, gs, B) && pb.part;
  if (fuse_enc) {
    const int rows = B * cfg.n_particles;
    Problem ef = blank_problem(PK_ENC_FUSED);
    ef.M = rows; ef.K = cfg.particle_dim; ef.N = cfg.n_particles;
    ef.A = pb.P; ef.a_go = pb.P_go;
    ef.B = W.base + net.c1w_off; ef.b_go = W.go; ef.b_gi = W.gi;
    ef.bias = W.base + net.c1b_off; ef.bias_go = W.go; ef.bias_gi = W.gi;
    ef.aux0 = const_cast<float*>((W.tc ? W.tc : W.base) + net.c2w_off); ef.aux0_go = W.go; ef.aux0_gi = W.gi;
    ef.aux1 = const_cast<float*>(W.base + net.c2b_off); ef.aux1_go = W.go; ef.aux1_gi = W.gi;
    if (pb.keep_enc_acts && pb.enc_bits) {            // the backward pass recomputes from bitmaps: no activation stores
      ef.tmapA = pb.enc_bits; ef.aux2_go = pb.bits_go; ef.aux2_gi = pb.bits_gi;
    } else if (pb.keep_enc_acts) {
      ef.aux2 = pb.h1; ef.aux2_go = pb.h1_go; ef.aux2_gi = pb.h1_gi;
      ef.aux3 = pb.h2; ef.aux3_go = pb.h2_go; ef.aux3_gi = pb.h2_gi;
    }
    ef.C = pb.part; ef.c_go = pb.part_go; ef.c_gi = pb.part_gi;
    finalize_problem(ef, gs);
    st.push_back({ef});
    // relu(mean over the sample's N / 128 partial means), into the first enc_out columns of the trunk input (:57-59)
    Problem pl = blank_problem(PK_POOL_FWD);
    pl.M = B; pl.N = net.enc_out; pl.K = cfg.n_particles / kEncTile;
    pl.A = pb.part; pl.lda = net.enc_out; pl.a_go = pb.part_go; pl.a_gi = pb.part_gi;
    pl.C = pb.x0; pl.ldc = pb.ld0; pl.c_go = pb.x0_go; pl.c_gi = pb.x0_gi;
    pl.c_dups = pool_dups; pl.c_dup_stride = pool_dup_stride;
    pl.rn_out = tf;
    finalize_problem(pl, gs);
    st.push_back({pl});
  } else if (enc) {
    const int rows = B * cfg.n_particles;
    // conv1 == per-particle linear D -> enc_hidden (TD3_particles.py:29,54)
    Problem e1 = make_gemm(rows, net.enc_hidden, cfg.particle_dim, pb.P, cfg.particle_dim, true, W.base + net.c1w_off,
                           cfg.particle_dim, true, pb.h1, net.enc_hidden, EPI_BIAS_RELU);
    set_groups(e1, pb.P_go, 0, W.go, W.gi, pb.h1_go, pb.h1_gi);
    e1.bias = W.base + net.c1b_off; e1.bias_go = W.go; e1.bias_gi = W.gi;
    if (cfg.particle_dim <= 8) e1.kind = PK_SMALLK_FWD;   // D-long reduction: dedicated HBM-write-bound tile
    e1.rn_out = tf;
    finalize_problem(e1, gs);
    st.push_back({e1});
    // conv2 (1x1) == linear enc_hidden -> enc_out (:30,56)
    Problem e2 = make_gemm(rows, net.enc_out, net.enc_hidden, pb.h1, net.enc_hidden, true, W.base + net.c2w_off,
                           net.enc_hidden, true, pb.h2, net.enc_out, EPI_BIAS_RELU);
    set_groups(e2, pb.h1_go, pb.h1_gi, W.go, W.gi, pb.h2_go, pb.h2_gi);
    e2.bias = W.base + net.c2b_off; e2.bias_go = W.go; e2.bias_gi = W.gi;
    weight_operand(e2, W, net.c2w_off);
    finalize_problem(e2, gs);
    st.push_back({e2});
    // avg_pool over particles + relu, written into the first enc_out columns of the trunk input (:57-59)
    Problem pl = blank_problem(PK_POOL_FWD);
    pl.M = B; pl.N = net.enc_out; pl.K = cfg.n_particles;
    pl.A = pb.h2; pl.lda = net.enc_out; pl.a_go = pb.h2_go; pl.a_gi = pb.h2_gi;
    pl.C = pb.x0; pl.ldc = pb.ld0; pl.c_go = pb.x0_go; pl.c_gi = pb.x0_gi;
    pl.c_dups = pool_dups; pl.c_dup_stride = pool_dup_stride;
    pl.rn_out = tf;
    finalize_problem(pl, gs);
    st.push_back({pl});
  }
  const float* in = pb.x0;
  int ld_in = pb.ld0;
  long long in_go = pb.x0_go, in_gi = pb.x0_gi;
  if (enc && ln) {   // lnorm1 on the concatenated input (:60-61)
    Problem p = blank_problem(PK_LN_FWD);
    p.M = B; p.N = net.dims[0];
    p.A = pb.x0; p.lda = pb.ld0; p.a_go = pb.x0_go; p.a_gi = pb.x0_gi;
    p.B = W.base + net.ln_in_g_off; p.b_go = W.go; p.b_gi = W.gi;
    p.bias = W.base + net.ln_in_b_off; p.bias_go = W.go; p.bias_gi = W.gi;
    p.C = pb.x0n; p.ldc = pb.ld0; p.c_go = pb.x0n_go; p.c_gi = pb.x0n_gi;
    p.aux2 = pb.mean0; p.aux3 = pb.rstd0; p.aux2_go = p.aux3_go = pb.s_go; p.aux2_gi = p.aux3_gi = pb.s_gi;
    p.f0 = 1e-5f;
    p.rn_out = tf;
    finalize_problem(p, gs);
    st.push_back({p});
    in = pb.x0n; in_go = pb.x0n_go; in_gi = pb.x0n_gi;
  }
  for (int l = 0; l < L; ++l) {
    const int K = net.dims[l], N = net.dims[l + 1];
    const bool last = l == L - 1;
    if (last && skip_last) break;
    Problem g = make_gemm(B, N, K, in, ld_in, true, W.base + net.w_off[l], K, true, last ? out.out : pb.r[l],
                          last ? out.ld : N, last ? out.epi : EPI_BIAS_RELU);
    set_groups(g, in_go, in_gi, W.go, W.gi, last ? out.go : pb.h_go[l], last ? out.gi : pb.h_gi[l]);
    g.bias = W.base + net.b_off[l]; g.bias_go = W.go; g.bias_gi = W.gi;
    if (last) {
      g.aux0 = out.aux0; g.ldaux = out.ldaux; g.aux0_go = out.aux0_go; g.aux0_gi = out.aux0_gi;
      g.f0 = out.f0; g.f1 = out.f1;
      g.c_dups = out.dups; g.c_dup_stride = out.dup_stride;
    }
    weight_operand(g, W, net.w_off[l]);
    // hidden activations feed the next layer's contraction (through LayerNorm when there is one: then its output is
    // the operand); an action written into a network input feeds that network's first layer; Q values are results
    g.rn_out = tf && (last ? out.epi != EPI_BIAS : !ln);
    finalize_problem(g, gs);
    if (!(l == 0 && skip_first)) st.push_back({g});
    if (last) break;
    in = pb.r[l]; ld_in = N; in_go = pb.h_go[l]; in_gi = pb.h_gi[l];
    if (ln) {   // post-ReLU LayerNorm (TD3_featured.py:44-46)
      Problem p = blank_problem(PK_LN_FWD);
      p.M = B; p.N = N;
      p.A = pb.r[l]; p.lda = N; p.a_go = pb.h_go[l]; p.a_gi = pb.h_gi[l];
      p.B = W.base + net.ln_g_off[l]; p.b_go = W.go; p.b_gi = W.gi;
      p.bias = W.base + net.ln_b_off[l]; p.bias_go = W.go; p.bias_gi = W.gi;
      p.C = pb.n[l]; p.ldc = N; p.c_go = pb.h_go[l]; p.c_gi = pb.h_gi[l];
      p.aux2 = pb.mean[l]; p.aux3 = pb.rstd[l]; p.aux2_go = p.aux3_go = pb.s_go; p.aux2_gi = p.aux3_gi = pb.s_gi;
      p.f0 = 1e-5f;
      p.rn_out = tf;
      finalize_problem(p, gs);
      st.push_back({p});
      in = pb.n[l];
    }
  }
  return st;
}

// ------------------------------------------------------------------------------------
// backward pass builder.  dout = d(loss)/d(network output) [B, w_L] (ld = ld_dout).
//   want_dw : write parameter gradients into G (same layout as the parameters)
//   dx0_mode: 0 none, 1 d(input columns [col0, col0+ncols)) -> dx0 (epilogue dx0_epi)
// Scratch: dz[2] ping-pong [groups][B*maxw], dn [groups][B*maxw].
// ------------------------------------------------------------------------------------
struct BwdScratch {
  float* dz[2] = {nullptr, nullptr};
  float* dn = nullptr;
  long long go = 0, gi = 0;          // strides of dz/dn
  float* dx0_full = nullptr; long long dx0_full_go = 0, dx0_full_gi = 0;   // [B, ld0] (encoder / lnorm1 path)
  float* dh2 = nullptr; float* dh1 = nullptr; long long dh2_go = 0, dh2_gi = 0, dh1_go = 0, dh1_gi = 0;
  float* part = nullptr; long long part_go = 0, part_gi = 0;               // split-K partials
  long long part_cap = 0;
};

struct Dz0Info { const float* dz = nullptr; int ld = 0; long long go = 0, gi = 0; };

struct Dx0Spec {
  int mode = 0;
  int col0 = 0, ncols = 0;
  float* out = nullptr; int ld = 0; long long go = 0, gi = 0;
  int epi = EPI_STORE;
  float* aux0 = nullptr; int ldaux = 0; long long aux0_go = 0, aux0_gi = 0;
  float f0 = 1.f;
};

int choose_ksplit(int tiles, int groups, int K) {
  if (K <= 2048) return 1;
  int want = std::max(1, 592 / std::max(1, tiles * groups));
  int cap = std::max(1, K / 512);
  return std::max(1, std::min(std::min(want, cap), 128));
}

// ------------------------------------------------------------------------------------
// Backward pass of the particle-set encoder (TD3_particles.py:53-58 under autograd): from the gradient w.r.t. the pooled
// features to the gradients of conv1 / conv2.  Used by build_backward and by the set_encoder_bwd export.
// ------------------------------------------------------------------------------------
constexpr int kEncPartials = 148;      // split-K partials the encoder-backward buffers hold per group (one per CTA at most)

struct EncBwdArgs {
  int B = 0, n_particles = 0, D = 0, H = 0, O = 0;
  const float* dpool = nullptr; int ld_dpool = 0; long long dpool_go = 0, dpool_gi = 0;     // d(loss)/d(pooled) [B, >= O]
  const float* pooled = nullptr; int ld_pooled = 0; long long pooled_go = 0, pooled_gi = 0; // relu(mean) (the outer ReLU's gate)
  const float* h1 = nullptr; const float* h2 = nullptr; long long h1_go = 0, h1_gi = 0, h2_go = 0, h2_gi = 0;
  const float* P = nullptr; long long P_go = 0;
  float* dh2 = nullptr; float* dh1 = nullptr; long long dh2_go = 0, dh2_gi = 0, dh1_go = 0, dh1_gi = 0;
  float* part = nullptr; long long part_go = 0, part_gi = 0;                                 // split-K partials
  const float* W2 = nullptr; const float* W2_tc = nullptr; long long w_go = 0, w_gi = 0;
  float* gW1 = nullptr; float* gb1 = nullptr; float* gW2 = nullptr; float* gb2 = nullptr; long long g_go = 0, g_gi = 0;
  const unsigned int* bits = nullptr; long long bits_go = 0, bits_gi = 0;   // ReLU bitmaps of the forward pass: fused backward
  const float* W1 = nullptr; const float* b1 = nullptr;
};

std::vector<ProblemList> enc_backward_stages(const EncBwdArgs& e, GroupShape gs) {
  std::vector<ProblemList> st;
  const int groups = gs.n_outer * gs.n_inner;
  const int tf = g_tc_mode ? 1 : 0;
  const int B = e.B;
  if (e.bits) {
    // fused backward (encbwd.cuh): two persistent tcgen05 launches leave per-CTA partials, one stage reduces them
    const int rows = B * e.n_particles, H = e.H, O = e.O, D = e.D;
    const int tiles = rows / kEncTile;
    const int ks2 = std::max(1, std::min(tiles, std::min(kEncPartials, g_sm_count / std::max(1, groups))));
    const int ks1 = std::max(1, std::min(tiles, std::min(kEncPartials, g_sm_count / std::max(1, 2 * groups))));
    auto fused = [&](int kind, int ks, long long part_ofs) {
      Problem p = blank_problem(kind);
      p.M = rows; p.K = D; p.N = e.n_particles;
      p.A = e.P; p.a_go = e.P_go;
      p.B = e.W1; p.bias = e.b1; p.b_go = e.w_go; p.b_gi = e.w_gi;
      p.aux0 = const_cast<float*>(e.W2_tc ? e.W2_tc : e.W2);
      p.aux1 = const_cast<float*>(e.dpool); p.lda = e.ld_dpool; p.aux1_go = e.dpool_go; p.aux1_gi = e.dpool_gi;
      p.aux2 = const_cast<float*>(e.pooled); p.ldb = e.ld_pooled; p.aux2_go = e.pooled_go; p.aux2_gi = e.pooled_gi;
      p.tmapA = e.bits; p.aux3_go = e.bits_go; p.aux3_gi = e.bits_gi;
      p.C = e.part; p.c_go = e.part_go; p.c_gi = e.part_gi; p.c_split = part_ofs;
      p.ksplit = ks;
      finalize_problem(p, gs);
      return p;
    };
    st.push_back({fused(PK_ENC_BWD_W2, ks2, 0)});
    float* part1 = e.part + (long long)ks2 * (O * H + O);
    st.push_back({fused(PK_ENC_BWD_X, ks1, (long long)ks2 * (O * H + O))});
    auto reduce = [&](const float* src, int n, int ks, float* dst) {
      Problem r = blank_problem(PK_REDUCE_SPLITS);
      r.M = n; r.K = ks; r.c_split = n;
      r.A = src; r.a_go = e.part_go; r.a_gi = e.part_gi;
      r.C = dst; r.c_go = e.g_go; r.c_gi = e.g_gi;
      finalize_problem(r, gs);
      return r;
    };
    st.push_back({reduce(e.part, O * H, ks2, e.gW2), reduce(e.part + (long long)ks2 * O * H, O, ks2, e.gb2),
                  reduce(part1, H * D, ks1, e.gW1), reduce(part1 + (long long)ks1 * H * D, H, ks1, e.gb1)});
    return st;
  }
  {
      const int rows = B * e.n_particles;
      const int H = e.H, O = e.O, D = e.D;
      // dH2 = dpool/N gated by both ReLUs (:56-57)
      Problem pbk = blank_problem(PK_POOL_BWD);
      pbk.M = B; pbk.N = O; pbk.K = e.n_particles;
      pbk.A = e.dpool; pbk.lda = e.ld_dpool; pbk.a_go = e.dpool_go; pbk.a_gi = e.dpool_gi;
      pbk.aux0 = const_cast<float*>(e.pooled); pbk.ldaux = e.ld_pooled; pbk.aux0_go = e.pooled_go; pbk.aux0_gi = e.pooled_gi;
      pbk.aux1 = const_cast<float*>(e.h2); pbk.ldb = O; pbk.aux1_go = e.h2_go; pbk.aux1_gi = e.h2_gi;
      pbk.C = e.dh2; pbk.ldc = O; pbk.c_go = e.dh2_go; pbk.c_gi = e.dh2_gi;
      pbk.rn_out = tf;
      finalize_problem(pbk, gs);
      st.push_back({pbk});
      // conv2: dW2 = dH2^T H1 (split-K over B*N), dH1 = (dH2 W2) * (H1 > 0)
      const int tiles2 = ((O + 31) / 32) * ((H + 31) / 32);
      // tensor-core dW2: many short slices (64 chunks each) so that the long reduction balances against the dX tiles of
      // the same stage under the static tile -> CTA assignment
      const int ks2 = g_tc_mode ? std::max(1, std::min(128, rows / 2048)) : choose_ksplit(tiles2, groups, rows);
      Problem dw2 = make_gemm(O, H, rows, e.dh2, O, false, e.h1, H, false, ks2 > 1 ? e.part : e.gW2, H,
                              EPI_STORE);
      dw2.ksplit = ks2; dw2.c_split = (long long)O * H;
      set_groups(dw2, e.dh2_go, e.dh2_gi, e.h1_go, e.h1_gi, ks2 > 1 ? e.part_go : e.g_go, ks2 > 1 ? e.part_gi : e.g_gi);
      finalize_problem(dw2, gs);
      Problem cs2 = blank_problem(PK_COLSUM);
      const bool dw2_tc = dw2.use_tc;
      if (dw2_tc) {
        cs2.N = O; cs2.K = rows; cs2.ksplit = ks2; cs2.c_split = O;
        cs2.A = e.dh2; cs2.lda = O; cs2.a_go = e.dh2_go; cs2.a_gi = e.dh2_gi;
        cs2.C = ks2 > 1 ? e.part + (long long)ks2 * O * H : e.gb2;
        cs2.c_go = ks2 > 1 ? e.part_go : e.g_go; cs2.c_gi = ks2 > 1 ? e.part_gi : e.g_gi;
        finalize_problem(cs2, gs);
      } else {
        dw2.aux1 = ks2 > 1 ? e.part + (long long)ks2 * O * H : e.gb2;
        dw2.aux1_go = ks2 > 1 ? e.part_go : e.g_go; dw2.aux1_gi = ks2 > 1 ? e.part_gi : e.g_gi;
        finalize_problem(dw2, gs);
      }
      Problem dx2 = make_gemm(rows, H, O, e.dh2, O, true, e.W2, H, false, e.dh1, H, EPI_RELU_MASK);
      set_groups(dx2, e.dh2_go, e.dh2_gi, e.w_go, e.w_gi, e.dh1_go, e.dh1_gi);
      dx2.aux0 = const_cast<float*>(e.h1); dx2.ldaux = H; dx2.aux0_go = e.h1_go; dx2.aux0_gi = e.h1_gi;
      if (e.W2_tc) { dx2.B = e.W2_tc; dx2.B_master = e.W2; }
      dx2.rn_out = tf;
      finalize_problem(dx2, gs);
      if (dw2_tc) st.push_back({dw2, dx2, cs2});
      else st.push_back({dw2, dx2});
      ProblemList s3;
      if (ks2 > 1) {
        Problem r1 = blank_problem(PK_REDUCE_SPLITS);
        r1.M = O * H; r1.K = ks2; r1.c_split = (long long)O * H;
        r1.A = e.part; r1.a_go = e.part_go; r1.a_gi = e.part_gi;
        r1.C = e.gW2; r1.c_go = e.g_go; r1.c_gi = e.g_gi;
        finalize_problem(r1, gs);
        s3.push_back(r1);
        Problem r2 = blank_problem(PK_REDUCE_SPLITS);
        r2.M = O; r2.K = ks2; r2.c_split = O;
        r2.A = e.part + (long long)ks2 * O * H; r2.a_go = e.part_go; r2.a_gi = e.part_gi;
        r2.C = e.gb2; r2.c_go = e.g_go; r2.c_gi = e.g_gi;
        finalize_problem(r2, gs);
        s3.push_back(r2);
      }
      // conv1: dW1 = dH1^T P (split-K), second half of the partial buffer
      const int tiles1 = ((H + 31) / 32) * ((D + 31) / 32);
      // the small-K tile reads dH1 once, coalesced: give every SM a couple of row slices
      const int ks1 = (D <= 8 && H <= kStageThreads) ? std::max(1, std::min(128, rows / 512)) : choose_ksplit(tiles1, groups, rows);
      float* part1 = e.part + (long long)ks2 * (O * H + O);
      Problem dw1 = make_gemm(H, D, rows, e.dh1, H, false, e.P, D, false, ks1 > 1 ? part1 : e.gW1, D,
                              EPI_STORE);
      dw1.ksplit = ks1; dw1.c_split = (long long)H * D;
      set_groups(dw1, e.dh1_go, e.dh1_gi, e.P_go, 0, ks1 > 1 ? e.part_go : e.g_go, ks1 > 1 ? e.part_gi : e.g_gi);
      dw1.aux1 = ks1 > 1 ? part1 + (long long)ks1 * H * D : e.gb1;
      dw1.aux1_go = ks1 > 1 ? e.part_go : e.g_go; dw1.aux1_gi = ks1 > 1 ? e.part_gi : e.g_gi;
      if (D <= 8 && H <= kStageThreads) dw1.kind = PK_SMALLK_DW;
      finalize_problem(dw1, gs);
      s3.push_back(dw1);
      st.push_back(s3);
      if (ks1 > 1) {
        ProblemList s4;
        Problem r1 = blank_problem(PK_REDUCE_SPLITS);
        r1.M = H * D; r1.K = ks1; r1.c_split = (long long)H * D;
        r1.A = part1; r1.a_go = e.part_go; r1.a_gi = e.part_gi;
        r1.C = e.gW1; r1.c_go = e.g_go; r1.c_gi = e.g_gi;
        finalize_problem(r1, gs);
        s4.push_back(r1);
        Problem r2 = blank_problem(PK_REDUCE_SPLITS);
        r2.M = H; r2.K = ks1; r2.c_split = H;
        r2.A = part1 + (long long)ks1 * H * D; r2.a_go = e.part_go; r2.a_gi = e.part_gi;
        r2.C = e.gb1; r2.c_go = e.g_go; r2.c_gi = e.g_gi;
        finalize_problem(r2, gs);
        s4.push_back(r2);
        st.push_back(s4);
      }
  }
  return st;
}

std::vector<ProblemList> build_backward(const td3_agent_config& cfg, const td3_net_layout& net, ParamRef W, GradRef G,
                                        GroupShape gs, int B, const PassBuf& pb, const float* dout, int ld_dout,
                                        long long dout_go, long long dout_gi, bool want_dw, const Dx0Spec& dx0,
                                        const BwdScratch& sc, bool head_done = false, bool din_top_done = false,
                                        Dz0Info* dz0 = nullptr) {
  // dz0 != nullptr: the first layer's dW/db are NOT emitted; *dz0 says where the gradient w.r.t. its pre-activation
  // ends up, for the fused gradient + optimiser tiles of the apply launch (apply.cuh)
  // din_top_done: the gradient w.r.t. the output layer's input is already in sc.dz[0] (front.cuh); the output layer's
  // dW/db still have to be computed and ride along with the next layer's stage
  // head_done: the output layer's dW/db and the gradient w.r.t. its input were produced by the fused head kernel
  // (into sc.dz[0], or sc.dn with LayerNorm): the walk starts below it
  std::vector<ProblemList> st;
  const bool ln = cfg.norm == TD3_NORM_LAYER;
  const bool enc = cfg.variant == TD3_VARIANT_PARTICLES;
  const int L = net.n_linear;
  const int groups = gs.n_outer * gs.n_inner;
  const int tf = g_tc_mode ? 1 : 0;      // gradients w.r.t. activations are operands of the next dW / dX contractions
  const float* dz = dout;
  int ld_dz = ld_dout;
  long long dz_go = dout_go, dz_gi = dout_gi;
  int pp = 0;
  ProblemList carry;
  for (int l = L - 1; l >= 0; --l) {
    const int K = net.dims[l], N = net.dims[l + 1];
    ProblemList stage = carry;
    carry.clear();
    // input activation of layer l
    const float* in; int ld_in; long long in_go, in_gi;
    if (l == 0) {
      const bool lnin = enc && ln;
      in = lnin ? pb.x0n : pb.x0; ld_in = pb.ld0;
      in_go = lnin ? pb.x0n_go : pb.x0_go; in_gi = lnin ? pb.x0n_gi : pb.x0_gi;
    } else {
      in = ln ? pb.n[l - 1] : pb.r[l - 1]; ld_in = K; in_go = pb.h_go[l - 1]; in_gi = pb.h_gi[l - 1];
    }
    const bool skip = head_done && l == L - 1 && l > 0;
    if (l == 0 && dz0) { dz0->dz = dz; dz0->ld = ld_dz; dz0->go = dz_go; dz0->gi = dz_gi; }
    if (want_dw && !skip && !(l == 0 && dz0)) {   // dW_l[n,k] = sum_b dz[b,n] in[b,k] ; db_l[n] = sum_b dz[b,n]
      Problem p = make_gemm(N, K, B, dz, ld_dz, false, in, ld_in, false, G.base + net.w_off[l], K, EPI_STORE);
      set_groups(p, dz_go, dz_gi, in_go, in_gi, G.go, G.gi);
      finalize_problem(p, gs);
      if (p.use_tc) {   // tensor-core tile: the bias gradient is a separate column-sum problem of the same stage
        Problem cs = blank_problem(PK_COLSUM);
        cs.N = N; cs.K = B; cs.A = dz; cs.lda = ld_dz; cs.a_go = dz_go; cs.a_gi = dz_gi;
        cs.C = G.base + net.b_off[l]; cs.c_go = G.go; cs.c_gi = G.gi;
        finalize_problem(cs, gs);
        stage.push_back(p);
        stage.push_back(cs);
      } else {          // FFMA tile: row sums of the A operand fall out of the fragments already in registers
        p.aux1 = G.base + net.b_off[l]; p.aux1_go = G.go; p.aux1_gi = G.gi;
        finalize_problem(p, gs);
        stage.push_back(p);
      }
    }
    if (l > 0) {     // d(in_l) = dz . W_l
      const bool mask_now = !ln;
      float* dst = mask_now ? sc.dz[pp] : sc.dn;
      if (din_top_done && l == L - 1 && mask_now) {
        carry = stage;
        dz = sc.dz[pp]; ld_dz = K; dz_go = sc.go; dz_gi = sc.gi;
        pp ^= 1;
        continue;
      }
      if (!skip) {
        Problem p = make_gemm(B, K, N, dz, ld_dz, true, W.base + net.w_off[l], K, false, dst, K,
                              mask_now ? EPI_RELU_MASK : EPI_STORE);
        set_groups(p, dz_go, dz_gi, W.go, W.gi, sc.go, sc.gi);
        if (mask_now) { p.aux0 = pb.r[l - 1]; p.ldaux = K; p.aux0_go = pb.h_go[l - 1]; p.aux0_gi = pb.h_gi[l - 1]; }
        weight_operand(p, W, net.w_off[l]);
        p.rn_out = tf && mask_now;        // with LayerNorm the fp32 LN-backward tile consumes this and rounds its own output
        finalize_problem(p, gs);
        stage.push_back(p);
      }
      if (!stage.empty()) st.push_back(stage);
      if (ln) {      // through LayerNorm then ReLU
        ProblemList s2;
        Problem rr = blank_problem(PK_LN_BWD_ROWS);
        rr.epi = EPI_RELU_MASK;
        rr.M = B; rr.N = K;
        rr.A = sc.dn; rr.lda = K; rr.a_go = sc.go; rr.a_gi = sc.gi;
        rr.B = W.base + net.ln_g_off[l - 1]; rr.b_go = W.go; rr.b_gi = W.gi;
        rr.aux0 = pb.r[l - 1]; rr.ldaux = K; rr.aux0_go = pb.h_go[l - 1]; rr.aux0_gi = pb.h_gi[l - 1];
        rr.aux2 = pb.mean[l - 1]; rr.aux3 = pb.rstd[l - 1];
        rr.aux2_go = rr.aux3_go = pb.s_go; rr.aux2_gi = rr.aux3_gi = pb.s_gi;
        rr.C = sc.dz[pp]; rr.ldc = K; rr.c_go = sc.go; rr.c_gi = sc.gi;
        rr.rn_out = tf;
        finalize_problem(rr, gs);
        s2.push_back(rr);
        if (want_dw) {
          Problem cc = blank_problem(PK_LN_BWD_COLS);
          cc.M = B; cc.N = K;
          cc.A = sc.dn; cc.lda = K; cc.a_go = sc.go; cc.a_gi = sc.gi;
          cc.aux0 = pb.r[l - 1]; cc.ldaux = K; cc.aux0_go = pb.h_go[l - 1]; cc.aux0_gi = pb.h_gi[l - 1];
          cc.aux2 = pb.mean[l - 1]; cc.aux3 = pb.rstd[l - 1];
          cc.aux2_go = cc.aux3_go = pb.s_go; cc.aux2_gi = cc.aux3_gi = pb.s_gi;
          cc.C = G.base + net.ln_g_off[l - 1]; cc.c_go = G.go; cc.c_gi = G.gi;
          cc.aux1 = G.base + net.ln_b_off[l - 1]; cc.aux1_go = G.go; cc.aux1_gi = G.gi;
          finalize_problem(cc, gs);
          s2.push_back(cc);
        }
        st.push_back(s2);
      }
      dz = sc.dz[pp]; ld_dz = K; dz_go = sc.go; dz_gi = sc.gi;
      pp ^= 1;
      continue;
    }
    // ---- l == 0 : gradient w.r.t. the trunk input ----
    const bool lnin = enc && ln;
    const bool need_enc_bwd = enc && want_dw;
    const bool need_full = lnin || need_enc_bwd;
    if (!need_full && dx0.mode == 1) {
      // only a column slice of d(x0) is needed (the action columns): slice W_0
      Problem p = make_gemm(B, dx0.ncols, N, dz, ld_dz, true, W.base + net.w_off[0] + dx0.col0, K, false, dx0.out, dx0.ld,
                            dx0.epi);
      set_groups(p, dz_go, dz_gi, W.go, W.gi, dx0.go, dx0.gi);
      p.aux0 = dx0.aux0; p.ldaux = dx0.ldaux; p.aux0_go = dx0.aux0_go; p.aux0_gi = dx0.aux0_gi; p.f0 = dx0.f0;
      weight_operand(p, W, net.w_off[0] + dx0.col0);
      p.rn_out = tf;
      finalize_problem(p, gs);
      stage.push_back(p);
      st.push_back(stage);
      break;
    }
    if (need_full || dx0.mode == 1) {
      float* dst = lnin ? sc.dn : sc.dx0_full;
      const int ldd = lnin ? K : pb.ld0;
      Problem p = make_gemm(B, K, N, dz, ld_dz, true, W.base + net.w_off[0], K, false, dst, ldd, EPI_STORE);
      set_groups(p, dz_go, dz_gi, W.go, W.gi, lnin ? sc.go : sc.dx0_full_go, lnin ? sc.gi : sc.dx0_full_gi);
      weight_operand(p, W, net.w_off[0]);
      finalize_problem(p, gs);
      stage.push_back(p);
    }
    if (!stage.empty()) st.push_back(stage);
    if (lnin) {
      ProblemList s2;
      Problem rr = blank_problem(PK_LN_BWD_ROWS);
      rr.epi = EPI_STORE;
      rr.M = B; rr.N = K;
      rr.A = sc.dn; rr.lda = K; rr.a_go = sc.go; rr.a_gi = sc.gi;
      rr.B = W.base + net.ln_in_g_off; rr.b_go = W.go; rr.b_gi = W.gi;
      rr.aux0 = pb.x0; rr.ldaux = pb.ld0; rr.aux0_go = pb.x0_go; rr.aux0_gi = pb.x0_gi;
      rr.aux2 = pb.mean0; rr.aux3 = pb.rstd0; rr.aux2_go = rr.aux3_go = pb.s_go; rr.aux2_gi = rr.aux3_gi = pb.s_gi;
      rr.C = sc.dx0_full; rr.ldc = pb.ld0; rr.c_go = sc.dx0_full_go; rr.c_gi = sc.dx0_full_gi;
      finalize_problem(rr, gs);
      s2.push_back(rr);
      if (want_dw) {
        Problem cc = blank_problem(PK_LN_BWD_COLS);
        cc.M = B; cc.N = K;
        cc.A = sc.dn; cc.lda = K; cc.a_go = sc.go; cc.a_gi = sc.gi;
        cc.aux0 = pb.x0; cc.ldaux = pb.ld0; cc.aux0_go = pb.x0_go; cc.aux0_gi = pb.x0_gi;
        cc.aux2 = pb.mean0; cc.aux3 = pb.rstd0; cc.aux2_go = cc.aux3_go = pb.s_go; cc.aux2_gi = cc.aux3_gi = pb.s_gi;
        cc.C = G.base + net.ln_in_g_off; cc.c_go = G.go; cc.c_gi = G.gi;
        cc.aux1 = G.base + net.ln_in_b_off; cc.aux1_go = G.go; cc.aux1_gi = G.gi;
        finalize_problem(cc, gs);
        s2.push_back(cc);
      }
      st.push_back(s2);
    }
    if (need_enc_bwd) {
      EncBwdArgs e;
      e.B = B; e.n_particles = cfg.n_particles; e.D = cfg.particle_dim; e.H = net.enc_hidden; e.O = net.enc_out;
      e.dpool = sc.dx0_full; e.ld_dpool = pb.ld0; e.dpool_go = sc.dx0_full_go; e.dpool_gi = sc.dx0_full_gi;
      e.pooled = pb.x0; e.ld_pooled = pb.ld0; e.pooled_go = pb.x0_go; e.pooled_gi = pb.x0_gi;
      e.h1 = pb.h1; e.h1_go = pb.h1_go; e.h1_gi = pb.h1_gi; e.h2 = pb.h2; e.h2_go = pb.h2_go; e.h2_gi = pb.h2_gi;
      e.P = pb.P; e.P_go = pb.P_go;
      e.dh2 = sc.dh2; e.dh2_go = sc.dh2_go; e.dh2_gi = sc.dh2_gi; e.dh1 = sc.dh1; e.dh1_go = sc.dh1_go; e.dh1_gi = sc.dh1_gi;
      e.part = sc.part; e.part_go = sc.part_go; e.part_gi = sc.part_gi;
      e.W2 = W.base + net.c2w_off; e.W2_tc = W.tc ? W.tc + net.c2w_off : nullptr; e.w_go = W.go; e.w_gi = W.gi;
      e.gW1 = G.base + net.c1w_off; e.gb1 = G.base + net.c1b_off; e.gW2 = G.base + net.c2w_off; e.gb2 = G.base + net.c2b_off;
      e.g_go = G.go; e.g_gi = G.gi;
      e.bits = pb.enc_bits; e.bits_go = pb.bits_go; e.bits_gi = pb.bits_gi;
      e.W1 = W.base + net.c1w_off; e.b1 = W.base + net.c1b_off;
      for (auto& s2 : enc_backward_stages(e, gs)) st.push_back(s2);
    }
  }
  return st;
}

long long partial_floats(const td3_agent_config& cfg, const td3_net_layout& net, int B) {
  if (cfg.variant != TD3_VARIANT_PARTICLES) return 1;
  const long long H = net.enc_hidden, O = net.enc_out, D = cfg.particle_dim;
  return kEncPartials * (O * H + O) + kEncPartials * (H * D + H);
}

// allocate the activations of a pass
void alloc_pass(Bump& ws, const td3_agent_config& cfg, const td3_net_layout& net, GroupShape gs, int B, PassBuf& pb,
                bool own_x0, bool need_enc_acts) {
  const int groups = gs.n_outer * gs.n_inner;
  const bool ln = cfg.norm == TD3_NORM_LAYER;
  const bool enc = cfg.variant == TD3_VARIANT_PARTICLES;
  if (own_x0) {
    pb.ld0 = (int)round_up(net.dims[0], 4);
    pb.x0 = ws.take((long long)groups * B * pb.ld0);
    pb.x0_gi = (long long)B * pb.ld0;
    pb.x0_go = pb.x0_gi * gs.n_inner;
  }
  pb.s_gi = B; pb.s_go = (long long)B * gs.n_inner;
  if (enc && ln) {
    pb.x0n = ws.take((long long)groups * B * pb.ld0);
    pb.x0n_gi = (long long)B * pb.ld0; pb.x0n_go = pb.x0n_gi * gs.n_inner;
    pb.mean0 = ws.take((long long)groups * B);
    pb.rstd0 = ws.take((long long)groups * B);
  }
  for (int l = 0; l + 1 < net.n_linear; ++l) {
    const long long w = net.dims[l + 1];
    pb.h_gi[l] = B * w; pb.h_go[l] = B * w * gs.n_inner;
    pb.r[l] = ws.take(groups * B * w);
    if (ln) {
      pb.n[l] = ws.take(groups * B * w);
      pb.mean[l] = ws.take((long long)groups * B);
      pb.rstd[l] = ws.take((long long)groups * B);
    }
  }
  if (enc && need_enc_acts && enc_fusable(cfg, net, gs, B) && !getenv("TD3_NO_ENC_BWD_FUSION")) {
    // fused backward: 64 bytes of ReLU bitmaps per particle instead of the 1.5 KB of h1 / h2
    const long long rows = (long long)B * cfg.n_particles;
    pb.bits_gi = rows * 16; pb.bits_go = pb.bits_gi * gs.n_inner;
    pb.enc_bits = reinterpret_cast<unsigned int*>(ws.take((long long)groups * rows * 16));
  } else if (enc && need_enc_acts) {
    const long long rows = (long long)B * cfg.n_particles;
    pb.h1_gi = rows * net.enc_hidden; pb.h1_go = pb.h1_gi * gs.n_inner;
    pb.h2_gi = rows * net.enc_out; pb.h2_go = pb.h2_gi * gs.n_inner;
    pb.h1 = ws.take(groups * rows * net.enc_hidden);
    pb.h2 = ws.take(groups * rows * net.enc_out);
  }
  if (enc) {
    const long long tiles = ((long long)B * cfg.n_particles + kEncTile - 1) / kEncTile;
    pb.part_gi = tiles * net.enc_out; pb.part_go = pb.part_gi * gs.n_inner;
    pb.part = ws.take(groups * tiles * net.enc_out);
  }
}

void alloc_bwd(Bump& ws, const td3_agent_config& cfg, const td3_net_layout& net, GroupShape gs, int B, BwdScratch& sc,
               bool enc_bwd) {
  const int groups = gs.n_outer * gs.n_inner;
  long long maxw = 4;
  for (int l = 0; l <= net.n_linear; ++l) maxw = std::max<long long>(maxw, round_up(net.dims[l], 4));
  sc.gi = B * maxw; sc.go = sc.gi * gs.n_inner;
  sc.dz[0] = ws.take(groups * B * maxw);
  sc.dz[1] = ws.take(groups * B * maxw);
  sc.dn = ws.take(groups * B * maxw);
  const long long ld0 = round_up(net.dims[0], 4);
  sc.dx0_full_gi = B * ld0; sc.dx0_full_go = sc.dx0_full_gi * gs.n_inner;
  sc.dx0_full = ws.take(groups * B * ld0);
  if (cfg.variant == TD3_VARIANT_PARTICLES && enc_bwd && enc_fusable(cfg, net, gs, B) && !getenv("TD3_NO_ENC_BWD_FUSION")) {
    sc.part_cap = partial_floats(cfg, net, B);      // fused backward (encbwd.cuh): no dH2 / dH1 in memory, only the partials
    sc.part_gi = sc.part_cap; sc.part_go = sc.part_cap * gs.n_inner;
    sc.part = ws.take(groups * sc.part_cap);
  } else if (cfg.variant == TD3_VARIANT_PARTICLES && enc_bwd) {
    const long long rows = (long long)B * cfg.n_particles;
    sc.dh2_gi = rows * net.enc_out; sc.dh2_go = sc.dh2_gi * gs.n_inner;
    sc.dh1_gi = rows * net.enc_hidden; sc.dh1_go = sc.dh1_gi * gs.n_inner;
    sc.dh2 = ws.take(groups * rows * net.enc_out);
    sc.dh1 = ws.take(groups * rows * net.enc_hidden);
    sc.part_cap = partial_floats(cfg, net, B);
    sc.part_gi = sc.part_cap; sc.part_go = sc.part_cap * gs.n_inner;
    sc.part = ws.take(groups * sc.part_cap);
  }
}

// out[b, j] = epi(in[b, j]) for a [B, ncols] column slice, expressed as a product with a tiny identity matrix so
// that it reuses the GEMM tile path and its epilogues (only the particles + LayerNorm actor step needs it).
Problem make_slice(int B, int ncols, const float* in, int ld_in, const float* eye, float* out, int ld_out, int epi) {
  Problem p = make_gemm(B, ncols, ncols, in, ld_in, true, eye, ncols, false, out, ld_out, epi);
  return p;
}

// ---- row-local front kernels (front.cuh) ----
// first layer of `net` as a front job: out = relu(x[:, :K] W_0^T + b_0) into the pass's r[0]
FrontNet front_first_layer(const td3_net_layout& net, ParamRef W, int n_inner, const PassBuf& pb, int rn_out = 0) {
  FrontNet n;
  memset(&n, 0, sizeof(n));
  n.rn_out = rn_out;
  n.W = W.base + net.w_off[0]; n.bias = W.base + net.b_off[0]; n.w_go = W.go; n.w_gi = W.gi;
  n.ws_c = net.dims[0]; n.ws_k = 1;
  n.out = pb.r[0]; n.out_go = pb.h_go[0]; n.out_gi = pb.h_gi[0]; n.ldo = net.dims[1];
  n.K = net.dims[0]; n.N = net.dims[1]; n.n_inner = n_inner; n.act_col = -1;
  return n;
}

// job table + grid of a front launch
void front_finish(Launch& L, int B, int nA) {
  FrontParams& F = L.front;
  // thousands of rows (a population, or a large data-parallel batch): the row-block-major kernel (front_wide_body), one
  // tile per (agent, row block, group of jobs); one agent at batch 256 keeps (8 rows, one job) tiles -- more CTAs in
  // flight on a path that is pure latency
  const bool many = (long long)nA * B >= 1024 && !getenv("TD3_FRONT_ROWS8");
  const bool wide = many && !getenv("TD3_NO_FRONT_WIDE");
  F.rows_per_tile = many ? kFrontRowsWide : kFrontRows;
  if (wide && (long long)nA * ((B + 15) / 16) <= 2 * 148) F.rows_per_tile = 16;
  F.batch = B; F.n_agents = nA; F.row_blocks = (B + F.rows_per_tile - 1) / F.rows_per_tile;
  int jobs = 0;
  for (int i = 0; i < F.n_nets; ++i) {
    F.net[i].job_begin = jobs;
    F.net[i].col_blocks = (F.net[i].N + kFrontCols - 1) / kFrontCols;
    jobs += F.net[i].n_inner * F.net[i].col_blocks;
  }
  F.jobs = jobs;
  L.kind = Launch::FRONT;
  if (wide) {
    // a tile = (agent, row block, group of (network, twin) units); the groups' weight matrices sit in shared memory all
    // at once when they fit (a tile then waits for memory once)
    std::vector<int> unit_floats;
    for (int i = 0; i < F.n_nets; ++i)
      for (int t = 0; t < F.net[i].n_inner; ++t) unit_floats.push_back(front_unit_floats(F.net[i].N, F.net[i].K));
    const int units = (int)unit_floats.size();
    int g = 1;
    while (g < units && (long long)nA * F.row_blocks * g < 128) ++g;
    F.job_groups = g; F.units = units;
    int window = 0;
    for (int jg = 0; jg < g; ++jg) {
      int need = 0;
      for (int u = units * jg / g; u < units * (jg + 1) / g; ++u) need += unit_floats[u];
      window = std::max(window, need);
    }
    const int cap = kFrontWideSmemBytes / 4 - front_wide_smem_floats_fixed(F.rows_per_tile, F.head != 0);
    F.w_window = std::max(std::min(window, cap), *std::max_element(unit_floats.begin(), unit_floats.end()));
    L.grid_x = nA * F.row_blocks * g;
  } else {
    F.job_groups = 0;
    L.grid_x = nA * F.row_blocks * jobs;
  }
}

// ---- weight normalisation (misc.cuh: wn_body) ----
// Row blocks (and, for the effective-parameter pass, pass-through chunks) of one network of a family.
bool make_wn_layout(const td3_net_layout& net, bool with_copies, WnLayout& Y) {
  memset(&Y, 0, sizeof(Y));
  const int L = net.n_linear;
  if (L > kWnMaxLayers) return false;
  Y.n_layers = L;
  int units = 0;
  std::vector<std::pair<long long, long long>> vs;
  for (int l = 0; l < L; ++l) {
    Y.w_off[l] = net.w_off[l]; Y.g_off[l] = net.wg_off[l];
    Y.out[l] = net.dims[l + 1]; Y.in[l] = net.dims[l];
    Y.unit_begin[l] = units;
    units += (Y.out[l] + 7) / 8;
    vs.push_back({net.w_off[l], net.w_off[l] + (long long)Y.out[l] * Y.in[l]});
  }
  if (with_copies) {     // everything that is not a weight_v slot is copied through unchanged
    std::sort(vs.begin(), vs.end());
    long long at = 0;
    vs.push_back({net.n_floats, net.n_floats});
    for (auto& iv : vs) {
      if (iv.first > at) {
        if (Y.n_copies >= kWnMaxCopies) return false;
        Y.c_off[Y.n_copies] = at; Y.c_len[Y.n_copies] = (int)(iv.first - at); Y.c_unit_begin[Y.n_copies] = units;
        units += (Y.c_len[Y.n_copies] + kWnCopyChunk - 1) / kWnCopyChunk;
        ++Y.n_copies;
      }
      at = std::max(at, iv.second);
    }
  }
  Y.units = units;
  return true;
}

struct WnJobSpec { const float* src; float* dst; long long net_stride; int n_nets; int family; };   // family: 0 actor, 1 critic

bool make_wn_launch(const td3_agent_config& c, int mode, std::initializer_list<WnJobSpec> jobs, Launch& L) {
  L = Launch{};
  L.kind = Launch::WN;
  WnParams& P = L.wn;
  memset(&P, 0, sizeof(P));
  P.mode = mode; P.rows_per_tile = 8;
  P.rn_out = (mode == 0 && g_tc_mode) ? 1 : 0;
  if (!make_wn_layout(c.actor, mode == 0, P.lay[0]) || !make_wn_layout(c.q, mode == 0, P.lay[1])) return false;
  int tiles = 0;
  for (const WnJobSpec& j : jobs) {
    if (P.n_jobs >= kWnMaxJobs) return false;
    WnJob& J = P.job[P.n_jobs++];
    J.src = j.src; J.dst = j.dst; J.net_stride = j.net_stride; J.n_nets = j.n_nets; J.layout = j.family;
    J.tile_begin = tiles;
    tiles += j.n_nets * P.lay[j.family].units;
  }
  P.total_tiles = tiles;
  L.grid_x = tiles;
  return tiles > 0;
}

Launch dp_sync_launch(const td3_agent* a, int signal_slot, int wait_slot) {
  Launch L;
  L.kind = Launch::DPSYNC;
  DpSyncParams& P = L.dpsync;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < a->dp_world; ++i) P.peer_flags[i] = a->dp_flags[i];
  P.counts = reinterpret_cast<unsigned int*>(a->state_u64 + 12);     // four 32-bit counters in the reserved state words
  P.world = a->dp_world; P.rank = a->dp_rank; P.signal_slot = signal_slot; P.wait_slot = wait_slot;
  return L;
}

int plan_agent(td3_agent* a, long long batch) {
  const td3_agent_config& c = a->cfg;
  const int B = (int)batch, nA = c.n_agents, nq = c.n_q;
  const bool enc = c.variant == TD3_VARIANT_PARTICLES;
  const bool ln = c.norm == TD3_NORM_LAYER;
  const bool wn = c.norm == TD3_NORM_WEIGHT;
  const GroupShape g_actor{nA, 1}, g_crit{nA, nq}, g_q1{nA, 1};
  Bump& ws = a->ws;
  ws.used = 0;
  ws.regions.clear();
  g_tc_mode = c.precision == TD3_PRECISION_TF32 ? 1 : 0;
  // A-panel multicast over thread-block clusters is implemented and tested but measured no faster on B200 (the K loop
  // is bound by the MMA's shared-memory A read, not by the operand stream: DESIGN.md section 5): opt-in.
  // Single agents only: with a population's groups the cluster launches fault (found in round 2, not debugged: the knob
  // buys nothing there either, since small-ring / many-tile launches exclude it).
  g_cluster_mode = (a->cluster_mode && nA == 1 && getenv("TD3_CLUSTER")) ? 1 : 0;
  cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, 0);
  if (g_sm_count <= 0) g_sm_count = 148;
  const int A = c.action_dim, S = c.state_dim, E = enc ? c.actor.enc_out : 0;
  a->in_a = c.actor.dims[0];
  a->in_q = c.q.dims[0];
  a->qw = c.q.dims[c.q.n_linear];
  a->ld_a = (int)round_up(a->in_a, 4);
  a->ld_q = (int)round_up(a->in_q, 4);
  const int ld_q = a->ld_q, ld_a = a->ld_a, qw = a->qw;
  if (a->in_a != E + S || a->in_q != E + S + A)
    return fail(TD3_ERR_INVALID, "layout dims[0] (%d,%d) inconsistent with enc_out+state_dim(+action_dim) (%d,%d)",
                a->in_a, a->in_q, E + S, E + S + A);

  // ---- batch staging ----
  a->idx = reinterpret_cast<long long*>(ws.take(2LL * nA * B, "indices"));
  a->idx_in = reinterpret_cast<long long*>(ws.take(2LL * nA * B, "indices_in"));
  a->noise_in = ws.take((long long)nA * B * A, "noise_in");
  a->eps = ws.take((long long)nA * B * A, "eps");
  a->r = ws.take((long long)nA * B, "reward");
  a->nd = ws.take((long long)nA * B, "not_done");
  // trunk inputs.  Q inputs are per twin only when an encoder makes the pooled columns differ.
  const int xq_inner = enc ? nq : 1;
  a->xq_gi = enc ? (long long)B * ld_q : 0;
  a->xq_go = (long long)B * ld_q * xq_inner;
  a->xq = ws.take((long long)nA * xq_inner * B * ld_q, "x_q");
  a->xq2 = ws.take((long long)nA * xq_inner * B * ld_q, "x_q_next");
  a->xpi_go = (long long)B * ld_q;
  a->xpi = ws.take((long long)nA * B * ld_q, "x_q_pi");
  if (enc) {
    a->xa_go = (long long)B * ld_a;
    a->xa = ws.take((long long)nA * B * ld_a, "x_actor");
    a->xa2 = ws.take((long long)nA * B * ld_a, "x_actor_next");
    const long long pn = (long long)B * c.n_particles * c.particle_dim;
    a->P = ws.take(nA * pn, "particles");
    a->P2 = ws.take(nA * pn, "particles_next");
  }
  a->q = ws.take((long long)nA * nq * B * qw, "q");
  a->tq = ws.take((long long)nA * nq * B * qw, "tq");
  a->y = ws.take((long long)nA * B * qw, "target_q");
  a->dq = ws.take((long long)nA * nq * B * qw, "dq");
  a->q_pi = ws.take((long long)nA * B * qw, "q_pi");
  a->dq_pi = ws.take((long long)nA * B * qw, "dq_pi");
  a->tanh_y = ws.take((long long)nA * B * A, "tanh_y");
  float* da = ws.take((long long)nA * B * A, "d_action");
  float* eye = ws.take((long long)A * A, "eye");
  // fused critic heads (misc.cuh: head_body): small output width, hidden width within the kernel's register tile
  const int Lq = c.q.n_linear, wq_last = Lq >= 2 ? c.q.dims[Lq - 1] : 0;
  const bool fuse_heads = Lq >= 2 && qw <= kHeadMaxQw && wq_last <= kHeadMaxW && !getenv("TD3_NO_HEAD_FUSION");
  const int head_ctas = (B + kHeadRows - 1) / kHeadRows;
  // the head's own dW/db: a GEMM problem of the first backward stage (from dq and the hidden activations) instead of a
  // cross-CTA reduction inside the head kernel, wherever the backward walk can start below the head (no LayerNorm)
  const bool head_dw_in_stage = fuse_heads && !ln && !getenv("TD3_HEAD_DW");
  const long long head_per_g = (long long)qw * wq_last + qw;
  const long long head_part_go = (long long)head_ctas * nq * head_per_g + head_ctas;
  float* head_part = ws.take(fuse_heads ? nA * head_part_go * 2 : 1, "head_partials");
  unsigned int* head_counter = reinterpret_cast<unsigned int*>(ws.take(2LL * nA, "head_counters"));
  unsigned int* head_seq = reinterpret_cast<unsigned int*>(ws.take(nA, "head_seq"));
  a->host_status_live = fuse_heads && a->host_status != nullptr && !getenv("TD3_NO_HOST_STATUS");
  a->head_seq = head_seq;
  a->eff_a = a->eff_at = a->eff_c = a->eff_ct = nullptr;
  if (wn) {
    a->eff_a = ws.take((long long)nA * c.actor.n_floats, "effective_actor");
    a->eff_at = ws.take((long long)nA * c.actor.n_floats, "effective_actor_target");
    a->eff_c = ws.take((long long)nA * nq * c.q.n_floats, "effective_critic");
    a->eff_ct = ws.take((long long)nA * nq * c.q.n_floats, "effective_critic_target");
  }
  const bool tf = g_tc_mode != 0;
  a->sh_a = a->sh_at = a->sh_c = a->sh_ct = nullptr;
  if (tf && !wn) {
    a->sh_a = ws.take((long long)nA * c.actor.n_floats, "tf32_actor");
    a->sh_at = ws.take((long long)nA * c.actor.n_floats, "tf32_actor_target");
    a->sh_c = ws.take((long long)nA * nq * c.q.n_floats, "tf32_critic");
    a->sh_ct = ws.take((long long)nA * nq * c.q.n_floats, "tf32_critic_target");
  }
  a->prof_dev = reinterpret_cast<long long*>(ws.take(2LL * 2 * 128 * 3, "prof"));
  a->tmaps_dev = reinterpret_cast<CUtensorMap*>(ws.take((long long)kMaxTensorMaps * (long long)(sizeof(CUtensorMap) / 4), "tensor_maps"));
  a->prog_dev = reinterpret_cast<StageRec*>(ws.take(2LL * kMaxProgStages * (long long)(sizeof(StageRec) / 4), "program"));

  // ---- passes ----
  auto share_x0 = [&](PassBuf& pb, float* x, int ld, long long go, long long gi) {
    pb.x0 = x; pb.ld0 = ld; pb.x0_go = go; pb.x0_gi = gi;
  };
  PassBuf &at = a->pb_at, &ct = a->pb_ct, &cc = a->pb_c, &pa = a->pb_a, &q1 = a->pb_q1;
  at = ct = cc = pa = q1 = PassBuf{};
  at.keep_enc_acts = ct.keep_enc_acts = q1.keep_enc_acts = false;   // only the passes with an encoder backward (critics, actor) keep h1 / h2
  if (enc) {
    share_x0(at, a->xa2, ld_a, a->xa_go, 0);
    share_x0(pa, a->xa, ld_a, a->xa_go, 0);
    at.P = a->P2; pa.P = a->P; ct.P = a->P2; cc.P = a->P; q1.P = a->P;
    at.P_go = pa.P_go = ct.P_go = cc.P_go = q1.P_go = (long long)B * c.n_particles * c.particle_dim;
  } else {
    share_x0(at, a->xq2, ld_q, a->xq_go, 0);   // actor input = first S columns of the Q input
    share_x0(pa, a->xpi, ld_q, a->xpi_go, 0);
  }
  share_x0(ct, a->xq2, ld_q, a->xq_go, a->xq_gi);
  share_x0(cc, a->xq, ld_q, a->xq_go, a->xq_gi);
  share_x0(q1, a->xpi, ld_q, a->xpi_go, 0);
  alloc_pass(ws, c, c.actor, g_actor, B, at, false, true);
  alloc_pass(ws, c, c.q, g_crit, B, ct, false, true);
  alloc_pass(ws, c, c.q, g_crit, B, cc, false, true);
  alloc_pass(ws, c, c.actor, g_actor, B, pa, false, true);
  alloc_pass(ws, c, c.q, g_q1, B, q1, false, true);
  BwdScratch sc_c, sc_q1, sc_a;
  alloc_bwd(ws, c, c.q, g_crit, B, sc_c, true);
  alloc_bwd(ws, c, c.q, g_q1, B, sc_q1, false);
  alloc_bwd(ws, c, c.actor, g_actor, B, sc_a, true);
  // layer-fused chains (chain.cuh): dZ of every layer at once (the weight-gradient stage runs after the whole chain),
  // per-tile flags / loss terms, launch epoch and completion counter
  const int ch_tiles = (B + kChRows - 1) / kChRows;
  float* ch_cz[TD3_MAX_LINEAR] = {};
  float* ch_az[TD3_MAX_LINEAR] = {};
  if (!enc) {
    for (int l = 0; l + 1 < c.q.n_linear; ++l) ch_cz[l] = ws.take((long long)nA * nq * B * c.q.dims[l + 1]);
    for (int l = 0; l < c.actor.n_linear; ++l) ch_az[l] = ws.take((long long)nA * B * c.actor.dims[l + 1]);
  }
  unsigned int* ch_sync = reinterpret_cast<unsigned int*>(ws.take(64 + (long long)nA * ch_tiles * nq, "chain_sync"));
  float* ch_loss_part = ws.take(2LL * nA * ch_tiles * nq, "chain_loss_terms");
  a->ws_floats = ws.used;
  if (!ws.base) return TD3_OK;   // sizing pass only

  // ---- parameter references ----
  const long long qn = c.q.n_floats, an = c.actor.n_floats;
  // what the contractions read: the packed parameters, or their weight-normalised copies
  const float* pa_w = wn ? a->eff_a : a->actor.params;
  const float* pat_w = wn ? a->eff_at : a->actor.target;
  const float* pc_w = wn ? a->eff_c : a->critic.params;
  const float* pct_w = wn ? a->eff_ct : a->critic.target;
  ParamRef Wa{pa_w, an, 0, a->sh_a}, Wat{pat_w, an, 0, a->sh_at};
  ParamRef Wc{pc_w, qn * nq, qn, a->sh_c}, Wct{pct_w, qn * nq, qn, a->sh_ct};
  ParamRef Wq1{pc_w, qn * nq, 0, a->sh_c};
  GradRef Ga{a->actor.grad, an, 0}, Gc{a->critic.grad, qn * nq, qn};

  a->seq_sample.clear(); a->seq_target.clear(); a->seq_critic_fb.clear(); a->seq_critic_apply.clear();
  a->seq_actor_fb.clear(); a->seq_actor_apply.clear();

  // Row-local links of the chain (sampling + first layers, actor output layer + the critic's first layer, and their
  // mirror in the actor's backward pass) run as front kernels when the shapes fit (front.cuh).
  const int La = c.actor.n_linear;
  const bool front = !enc && A <= kFrontMaxA && S + A <= kFrontMaxK && La >= 2 && Lq >= 2 &&
                     c.actor.dims[La - 1] <= kFrontMaxKh && c.q.dims[1] <= kFrontMaxKh && !getenv("TD3_NO_FRONT");
  a->front_on = front;
  // tail fusion (apply.cuh): plain MLPs whose first layer sits at the head of each network's packed block
  auto l0_fusable = [](const td3_net_layout& n) {
    return n.n_linear >= 2 && n.dims[0] <= kDwMaxK && (n.dims[1] & 3) == 0 && n.w_off[0] == 0 && (n.w_off[1] & 3) == 0 &&
           n.b_off[0] >= (long long)n.dims[0] * n.dims[1] && n.w_off[1] >= n.b_off[0] + n.dims[1] && (n.n_floats & 3) == 0;
  };
  const bool dp_fused = a->dp_fused && a->dp_world >= 1;
  const bool fuse_tail = front && !dp_fused && !ln && !wn && l0_fusable(c.q) && l0_fusable(c.actor) && (ld_q & 3) == 0 &&
                         Lq >= 2 && qw <= kHeadMaxQw && c.q.dims[Lq - 1] <= kHeadMaxW && !getenv("TD3_NO_HEAD_FUSION") &&
                         !getenv("TD3_NO_TAIL_FUSION");
  a->tail_fused = fuse_tail;
  // first-layer outputs of the front kernels feed the second layers' tensor-core tiles (through LayerNorm when on,
  // whose own output is then the rounded operand)
  const int rn_front = tf && !ln ? 1 : 0;
  std::vector<Launch> v_cb, v_abf, v_tp;     // fused-mode pieces: critic head + backward, actor backward, target pass + actor forward
  std::vector<ProblemList> s_abf;
  Dz0Info dz0_c, dz0_a;
  size_t n_actor_pre_bwd = 0;
  if (front) {
    // sampling launch: gather + target-actor L1 on s' + online-critic L1 on [s, a]; plan_sample adds the replay view
    Launch L;
    FrontParams& F = L.front;
    memset(&F, 0, sizeof(F));
    F.gather = 1; F.A = A;
    F.n_nets = 2;
    F.net[0] = front_first_layer(c.q, Wc, nq, cc, rn_front);      // x_off 0 first: job 0 stages (and scatters) the whole row
    F.net[0].x_off = 0;
    F.net[1] = front_first_layer(c.actor, Wat, 1, at, rn_front);
    F.net[1].x_off = S + A;
    front_finish(L, B, nA);
    a->front_sample = F;
  }

  if (wn) {   // refresh all four effective buffers (the caller may have loaded or edited parameters since the last update)
    Launch L;
    if (!make_wn_launch(c, 0, {{a->actor.params, a->eff_a, an, nA, 0}, {a->actor.target, a->eff_at, an, nA, 0},
                               {a->critic.params, a->eff_c, qn, nA * nq, 1}, {a->critic.target, a->eff_ct, qn, nA * nq, 1}}, L))
      return fail(TD3_ERR_INVALID, "weight normalisation: unsupported layout");
    a->seq_target.push_back(L);
  }

  // ---- target step (TD3_featured.py:129-142) ----
  {
    OutSpec o;   // next_action into the action columns of the target-critic input(s)
    o.out = a->xq2 + E + S; o.ld = ld_q; o.go = a->xq_go; o.gi = 0;
    o.epi = EPI_BIAS_TANH_NOISE; o.aux0 = a->eps; o.ldaux = A; o.aux0_go = (long long)B * A;
    o.f0 = enc ? 1.f : c.max_action;
    o.f1 = c.clamp_target_action ? c.max_action : 0.f;
    o.dups = xq_inner; o.dup_stride = a->xq_gi;
    auto s_at = build_forward(c, c.actor, Wat, g_actor, B, at, o, 1, 0, front, front);
    OutSpec oq;
    oq.out = a->tq; oq.ld = qw; oq.gi = (long long)B * qw; oq.go = oq.gi * nq; oq.epi = EPI_BIAS;
    auto s_ct = build_forward(c, c.q, Wct, g_crit, B, ct, oq, 1, 0, fuse_heads, front);
    // the online critic forward is independent of the target path: run it alongside the target actor
    OutSpec oc;
    oc.out = a->q; oc.ld = qw; oc.gi = (long long)B * qw; oc.go = oc.gi * nq; oc.epi = EPI_BIAS;
    auto s_cc = build_forward(c, c.q, Wc, g_crit, B, cc, oc, 1, 0, fuse_heads, front);
    for (auto& st : zip_stages({s_at, s_cc})) emit_stage(a->seq_target, st);
    if (front) {
      // next_action = clamp(max_action * tanh(h W_L^T + b_L) + eps) (TD3_featured.py:131-138), then both target
      // critics' first layer on [s', next_action]
      Launch L;
      FrontParams& F = L.front;
      memset(&F, 0, sizeof(F));
      F.head = 1; F.A = A;
      F.h = ln ? at.n[La - 2] : at.r[La - 2]; F.h_go = at.h_go[La - 2]; F.ldh = c.actor.dims[La - 1]; F.Kh = c.actor.dims[La - 1];
      F.Wh = pat_w + c.actor.w_off[La - 1]; F.bh = pat_w + c.actor.b_off[La - 1]; F.wh_go = an;
      F.hs_j = c.actor.dims[La - 1]; F.hs_k = 1;
      F.head_epi = EPI_BIAS_TANH_NOISE; F.aux_in = a->eps; F.aux_go = (long long)B * A;
      F.a_out = a->xq2 + S; F.a_go = a->xq_go; F.a_ld = ld_q;
      F.f0 = o.f0; F.f1 = o.f1;
      F.n_nets = 1;
      F.net[0] = front_first_layer(c.q, Wct, nq, ct, rn_front);
      F.net[0].x = a->xq2; F.net[0].x_go = a->xq_go; F.net[0].ldx = ld_q; F.net[0].x_off = 0; F.net[0].act_col = S;
      front_finish(L, B, nA);
      a->seq_target.push_back(L);
    }
    for (auto& st : s_ct) emit_stage(a->seq_target, st);
  }
  // ---- critic loss + backward (:145-152) ----
  if (dp_fused) a->seq_critic_fb.push_back(dp_sync_launch(a, -1, 1));   // every rank is done reading my previous critic gradient
  {
    Launch L;
    L.kind = Launch::LOSS;
    L.grid_x = nA;
    LossParams& lp = L.loss;
    lp.q = a->q; lp.tq = a->tq; lp.r = a->r; lp.nd = a->nd; lp.y = a->y; lp.dq = a->dq; lp.loss = a->state_f32;
    lp.batch = B; lp.width = qw; lp.ldq = qw; lp.n_q = nq;
    lp.q_gi = (long long)B * qw; lp.q_go = lp.q_gi * nq; lp.y_go = (long long)B * qw; lp.r_go = B;
    lp.discount = c.discount;
    lp.inv_norm = 1.f / (float)((a->global_batch > 0 ? a->global_batch : batch) * qw);
    lp.tick = AdamTick{a->state_u64, 0, 0, c.lr_critic, c.beta1, c.beta2};
    lp.rn_out = tf ? 1 : 0;
    if (fuse_heads) {       // both heads + loss + head backward in one launch
      const float inv_norm = lp.inv_norm;
      const AdamTick tick = lp.tick;
      L = Launch{};
      L.kind = Launch::HEAD;
      HeadParams& H = L.head;
      memset(&H, 0, sizeof(H));
      const int l2 = Lq - 2;
      H.h = ln ? cc.n[l2] : cc.r[l2]; H.h_go = cc.h_go[l2]; H.h_gi = cc.h_gi[l2];
      H.ht = ln ? ct.n[l2] : ct.r[l2]; H.ht_go = ct.h_go[l2]; H.ht_gi = ct.h_gi[l2];
      H.W = pc_w + c.q.w_off[Lq - 1]; H.b = pc_w + c.q.b_off[Lq - 1]; H.w_go = qn * nq; H.w_gi = qn;
      H.Wt = pct_w + c.q.w_off[Lq - 1]; H.bt = pct_w + c.q.b_off[Lq - 1]; H.wt_go = qn * nq; H.wt_gi = qn;
      H.r = a->r; H.nd = a->nd; H.r_go = B;
      H.q = a->q; H.tq = a->tq; H.dq = a->dq; H.q_gi = (long long)B * qw; H.q_go = H.q_gi * nq;
      H.y = a->y; H.y_go = (long long)B * qw;
      H.dz = ln ? sc_c.dn : sc_c.dz[0]; H.dz_go = sc_c.go; H.dz_gi = sc_c.gi;
      H.gW = a->critic.grad + c.q.w_off[Lq - 1]; H.gb = a->critic.grad + c.q.b_off[Lq - 1]; H.g_go = qn * nq; H.g_gi = qn;
      H.part = head_part; H.part_go = head_part_go; H.counter = head_counter; H.loss = a->state_f32;
      H.batch = B; H.w = wq_last; H.qw = qw; H.n_q = nq; H.ldh = wq_last; H.lddz = wq_last; H.n_cta = head_ctas;
      H.mode = 0; H.relu_mask = ln ? 0 : 1; H.skip_dw = head_dw_in_stage ? 1 : 0;
      H.defer_finish = (head_dw_in_stage && !getenv("TD3_NO_DEFER_FINISH")) ? 1 : 0;
      H.rn_out = tf ? 1 : 0;
      H.discount = c.discount; H.inv_norm = inv_norm; H.tick = tick;
      H.host_status = a->host_status_live ? a->host_status : nullptr; H.seq = head_seq;
      L.grid_x = nA * head_ctas;
      L.smem_bytes = (int)((2LL * nq * qw * wq_last + 4 * kHeadMaxQw + kHeadRows * 2 * (kHeadMaxQw + 1) +
                            (head_dw_in_stage ? 0 : (long long)kHeadRows * nq * head_per_g) + 16) * sizeof(float));
    }
    a->seq_critic_fb.push_back(L);
    Dx0Spec none;
    auto s_b = build_backward(c, c.q, Wc, Gc, g_crit, B, cc, a->dq, qw, (long long)B * qw * nq, (long long)B * qw, true,
                              none, sc_c, fuse_heads && !head_dw_in_stage, head_dw_in_stage);
    for (auto& st : s_b) emit_stage(a->seq_critic_fb, st);
    if (fuse_tail) {
      v_cb.push_back(L);
      auto s_bf = build_backward(c, c.q, Wc, Gc, g_crit, B, cc, a->dq, qw, (long long)B * qw * nq, (long long)B * qw, true,
                                 none, sc_c, fuse_heads && !head_dw_in_stage, head_dw_in_stage, &dz0_c);
      for (auto& st : s_bf) emit_stage(v_cb, st);
    }
    if (wn) {     // dL/dW -> (dL/dg, dL/dv) in place
      Launch Lw;
      if (!make_wn_launch(c, 1, {{a->critic.params, a->critic.grad, qn, nA * nq, 1}}, Lw))
        return fail(TD3_ERR_INVALID, "weight normalisation: unsupported layout");
      a->seq_critic_fb.push_back(Lw);
    }
  }
  // ---- critic Adam (:153) ----
  {
    Launch L;
    L.kind = Launch::EW;
    EwParams& e = L.ew;
    e.n_ranges = 1;
    e.beta1 = c.beta1; e.beta2 = c.beta2; e.eps = c.adam_eps; e.tau = c.tau;
    EwRange& r = e.r[0];
    r.p = a->critic.params; r.g = a->critic.grad; r.m = a->critic.exp_avg; r.v = a->critic.exp_avg_sq; r.tgt = nullptr;
    r.n = qn * nq * nA; r.blk_begin = 0; r.sc_ptr = reinterpret_cast<const float*>(a->state_u64 + 10); r.do_adam = 1; r.do_polyak = 0;
    r.p_sh = a->sh_c;
    L.grid_x = (int)((r.n + kEwPerBlock - 1) / kEwPerBlock);
    if (dp_fused) {
      for (int i = 0; i < a->dp_world; ++i) r.g_peer[i] = a->dp_critic_grads[i];
      r.n_peer = a->dp_world;
      a->seq_critic_apply.push_back(dp_sync_launch(a, 0, 0));     // my gradient is complete; wait for everybody's
      a->seq_critic_apply.push_back(L);
      a->seq_critic_apply.push_back(dp_sync_launch(a, 1, -1));    // done reading the peers' gradients
    } else {
      a->seq_critic_apply.push_back(L);
    }
  }
  // ---- actor step (:159-163): actor fwd, Q1 fwd with the stepped critic, -mean, backward ----
  {
    OutSpec o;
    o.out = a->xpi + E + S; o.ld = ld_q; o.go = a->xpi_go; o.gi = 0;
    o.epi = EPI_BIAS_TANH; o.aux0 = a->tanh_y; o.ldaux = A; o.aux0_go = (long long)B * A;
    o.f0 = enc ? 1.f : c.max_action;
    auto s_a = build_forward(c, c.actor, Wa, g_actor, B, pa, o, 1, 0, front);
    for (auto& st : s_a) emit_stage(a->seq_actor_fb, st);
    a->n_actor_fwd = (int)a->seq_actor_fb.size();
    if (dp_fused) a->seq_actor_fb.push_back(dp_sync_launch(a, -1, 3));   // every rank is done reading my previous actor gradient
    if (wn) {     // Q1 below reads the critic the Adam step just changed
      Launch Lw;
      if (!make_wn_launch(c, 0, {{a->critic.params, a->eff_c, qn, nA * nq, 1}}, Lw))
        return fail(TD3_ERR_INVALID, "weight normalisation: unsupported layout");
      a->seq_actor_fb.push_back(Lw);
    }
    OutSpec oq;
    oq.out = a->q_pi; oq.ld = qw; oq.go = (long long)B * qw; oq.epi = EPI_BIAS;
    if (front) {
      // action = max_action * tanh(h W_L^T + b_L) (:159), then Q1's first layer on [s, action] with the stepped critic
      Launch L;
      FrontParams& F = L.front;
      memset(&F, 0, sizeof(F));
      F.head = 1; F.A = A;
      F.h = ln ? pa.n[La - 2] : pa.r[La - 2]; F.h_go = pa.h_go[La - 2]; F.ldh = c.actor.dims[La - 1]; F.Kh = c.actor.dims[La - 1];
      F.Wh = pa_w + c.actor.w_off[La - 1]; F.bh = pa_w + c.actor.b_off[La - 1]; F.wh_go = an;
      F.hs_j = c.actor.dims[La - 1]; F.hs_k = 1;
      F.head_epi = EPI_BIAS_TANH; F.aux_out = a->tanh_y; F.aux_go = (long long)B * A;
      F.a_out = a->xpi + S; F.a_go = a->xpi_go; F.a_ld = ld_q;
      F.f0 = o.f0;
      F.n_nets = 1;
      F.net[0] = front_first_layer(c.q, Wq1, 1, q1, rn_front);
      F.net[0].x = a->xpi; F.net[0].x_go = a->xpi_go; F.net[0].ldx = ld_q; F.net[0].x_off = 0; F.net[0].act_col = S;
      front_finish(L, B, nA);
      a->seq_actor_fb.push_back(L);
    }
    auto s_q = build_forward(c, c.q, Wq1, g_q1, B, q1, oq, 1, 0, fuse_heads, front);
    for (auto& st : s_q) emit_stage(a->seq_actor_fb, st);
    if (fuse_heads) {       // Q1 head + (-mean) + head backward + actor optimiser tick
      Launch L;
      L.kind = Launch::HEAD;
      HeadParams& H = L.head;
      memset(&H, 0, sizeof(H));
      const int l2 = Lq - 2;
      H.h = ln ? q1.n[l2] : q1.r[l2]; H.h_go = q1.h_go[l2]; H.h_gi = q1.h_gi[l2];
      H.W = pc_w + c.q.w_off[Lq - 1]; H.b = pc_w + c.q.b_off[Lq - 1]; H.w_go = qn * nq; H.w_gi = 0;
      H.q = a->q_pi; H.q_go = (long long)B * qw; H.q_gi = 0;
      H.dz = ln ? sc_q1.dn : sc_q1.dz[0]; H.dz_go = sc_q1.go; H.dz_gi = sc_q1.gi;
      H.part = head_part + nA * head_part_go; H.part_go = head_part_go; H.counter = head_counter + nA; H.loss = a->state_f32 + nA;
      H.batch = B; H.w = wq_last; H.qw = qw; H.n_q = 1; H.ldh = wq_last; H.lddz = wq_last; H.n_cta = head_ctas;
      H.mode = 1; H.relu_mask = ln ? 0 : 1;
      H.defer_finish = getenv("TD3_NO_DEFER_FINISH") ? 0 : 1;
      H.rn_out = tf ? 1 : 0;
      H.inv_norm = 1.f / (float)((a->global_batch > 0 ? a->global_batch : batch) * qw);
      H.tick = AdamTick{a->state_u64, 1, 0, c.lr_actor, c.beta1, c.beta2};
      L.grid_x = nA * head_ctas;
      L.smem_bytes = (int)((1LL * qw * wq_last + 4 * kHeadMaxQw + kHeadRows * 2 * (kHeadMaxQw + 1) + 16) * sizeof(float));
      a->seq_actor_fb.push_back(L);
    }
    // actor_loss = -mean(Q1) (read-back only) + bump the actor Adam step counter
    {
      Problem nm = blank_problem(PK_NEG_MEAN);
      nm.M = B; nm.N = qw; nm.A = a->q_pi; nm.lda = qw; nm.a_go = (long long)B * qw;
      nm.C = a->state_f32 + nA; nm.c_go = 1; nm.f0 = -1.f;
      finalize_problem(nm, g_q1);
      // first backward stage of Q1 shares the launch with the read-back
      // without LayerNorm the action gradient and the actor's top-layer input gradient are one front kernel
      const bool front_bwd = front && !ln;
      Dx0Spec dx;
      dx.mode = front_bwd ? 0 : 1; dx.col0 = E + S; dx.ncols = A;
      const bool lnin = enc && ln;
      dx.out = da; dx.ld = A; dx.go = (long long)B * A;
      dx.epi = lnin ? EPI_STORE : EPI_TANH_GRAD;
      dx.aux0 = a->tanh_y; dx.ldaux = A; dx.aux0_go = (long long)B * A; dx.f0 = enc ? 1.f : c.max_action;
      auto s_qb = build_backward(c, c.q, Wq1, GradRef{}, g_q1, B, q1, a->dq_pi, qw, (long long)B * qw, 0, false, dx,
                                 sc_q1, fuse_heads);
      if (!fuse_heads && !s_qb.empty()) s_qb[0].push_back(nm);
      else if (!fuse_heads) emit_stage(a->seq_actor_fb, {nm});
      for (auto& st : s_qb) emit_stage(a->seq_actor_fb, st);
      if (front_bwd) {
        // d(action) = (dz1 . W_0[:, S:S+A]) * max_action * (1 - tanh^2)  (through Q1's first layer and the actor's tanh),
        // then the actor's top layer: dz = (d(action) . W_L) * (hidden > 0)
        Launch L;
        FrontParams& F = L.front;
        memset(&F, 0, sizeof(F));
        F.head = 1; F.A = A;
        F.h = sc_q1.dz[(Lq - 2) & 1]; F.h_go = sc_q1.go; F.ldh = c.q.dims[1]; F.Kh = c.q.dims[1];
        F.Wh = pc_w + c.q.w_off[0] + S; F.bh = nullptr; F.wh_go = qn * nq; F.hs_j = 1; F.hs_k = c.q.dims[0];
        F.head_epi = EPI_TANH_GRAD; F.aux_in = a->tanh_y; F.aux_go = (long long)B * A;
        F.a_out = da; F.a_go = (long long)B * A; F.a_ld = A;
        F.f0 = dx.f0;
        F.n_nets = 1;
        FrontNet& n = F.net[0];
        n.W = pa_w + c.actor.w_off[La - 1]; n.bias = nullptr; n.w_go = an; n.w_gi = 0;
        n.ws_c = 1; n.ws_k = c.actor.dims[La - 1];
        n.out = sc_a.dz[0]; n.out_go = sc_a.go; n.out_gi = sc_a.gi; n.ldo = c.actor.dims[La - 1];
        n.mask = pa.r[La - 2]; n.mask_go = pa.h_go[La - 2]; n.mask_gi = pa.h_gi[La - 2];
        n.K = A; n.N = c.actor.dims[La - 1]; n.n_inner = 1; n.act_col = 0;
        n.rn_out = tf ? 1 : 0;
        front_finish(L, B, nA);
        a->seq_actor_fb.push_back(L);
      }
      if (lnin) {
        // d(action) = slice of the full input gradient, then through tanh: identity-GEMM slice
        Problem sl = make_slice(B, A, sc_q1.dx0_full + E + S, ld_q, eye, da, A, EPI_TANH_GRAD);
        set_groups(sl, sc_q1.dx0_full_go, 0, 0, 0, (long long)B * A, 0);
        sl.aux0 = a->tanh_y; sl.ldaux = A; sl.aux0_go = (long long)B * A; sl.f0 = 1.f;
        sl.rn_out = tf ? 1 : 0;
        finalize_problem(sl, g_q1);
        emit_stage(a->seq_actor_fb, {sl});
      }
    }
    if (!fuse_heads) {
      Launch L;
      L.kind = Launch::TICK;
      L.tick = AdamTick{a->state_u64, 1, 0, c.lr_actor, c.beta1, c.beta2};
      a->seq_actor_fb.push_back(L);
    }
    Dx0Spec none;
    auto s_ab = build_backward(c, c.actor, Wa, Ga, g_actor, B, pa, da, A, (long long)B * A, 0, true, none, sc_a, false,
                               front && !ln);
    n_actor_pre_bwd = a->seq_actor_fb.size();
    for (auto& st : s_ab) emit_stage(a->seq_actor_fb, st);
    if (fuse_tail) {
      s_abf = build_backward(c, c.actor, Wa, Ga, g_actor, B, pa, da, A, (long long)B * A, 0, true, none, sc_a, false,
                             front && !ln, &dz0_a);
      for (auto& st : s_abf) emit_stage(v_abf, st);
    }
    if (wn) {
      Launch Lw;
      if (!make_wn_launch(c, 1, {{a->actor.params, a->actor.grad, an, nA, 0}}, Lw))
        return fail(TD3_ERR_INVALID, "weight normalisation: unsupported layout");
      a->seq_actor_fb.push_back(Lw);
    }
  }
  // ---- actor Adam + Polyak of both targets (:164-171) ----
  {
    Launch L;
    L.kind = Launch::EW;
    EwParams& e = L.ew;
    e.n_ranges = 2;
    e.beta1 = c.beta1; e.beta2 = c.beta2; e.eps = c.adam_eps; e.tau = c.tau;
    EwRange& r0 = e.r[0];
    r0.p = a->critic.params; r0.tgt = a->critic.target; r0.n = qn * nq * nA; r0.blk_begin = 0;
    r0.do_adam = 0; r0.do_polyak = 1;
    r0.tgt_sh = a->sh_ct;
    const long long b0 = (r0.n + kEwPerBlock - 1) / kEwPerBlock;
    EwRange& r1 = e.r[1];
    r1.p = a->actor.params; r1.g = a->actor.grad; r1.m = a->actor.exp_avg; r1.v = a->actor.exp_avg_sq;
    r1.tgt = a->actor.target; r1.n = an * nA; r1.blk_begin = b0; r1.sc_ptr = reinterpret_cast<const float*>(a->state_u64 + 11);
    r1.do_adam = 1; r1.do_polyak = 1;
    r1.p_sh = a->sh_a; r1.tgt_sh = a->sh_at;
    L.grid_x = (int)(b0 + (r1.n + kEwPerBlock - 1) / kEwPerBlock);
    if (dp_fused) {
      for (int i = 0; i < a->dp_world; ++i) r1.g_peer[i] = a->dp_actor_grads[i];
      r1.n_peer = a->dp_world;
      a->seq_actor_apply.push_back(dp_sync_launch(a, 2, 2));
      a->seq_actor_apply.push_back(L);
      a->seq_actor_apply.push_back(dp_sync_launch(a, 3, -1));
    } else {
      a->seq_actor_apply.push_back(L);
    }
  }
  // The actor's forward pass (TD3_featured.py:159, actor(state)) reads nothing the critic update writes, so in the
  // fused policy update its layers ride along with the critic's backward stages instead of owning barriers.
  {
    a->seq_policy_mid.clear();
    size_t ai = 0;
    for (const Launch& L : a->seq_critic_fb) {
      Launch m = L;
      if (L.kind == Launch::STAGE && (int)ai < a->n_actor_fwd && merge_stage(m, a->seq_actor_fb[ai])) ++ai;
      a->seq_policy_mid.push_back(m);
    }
    for (; (int)ai < a->n_actor_fwd; ++ai) a->seq_policy_mid.push_back(a->seq_actor_fb[ai]);
    for (const Launch& L : a->seq_critic_apply) a->seq_policy_mid.push_back(L);
    for (size_t i = a->n_actor_fwd; i < a->seq_actor_fb.size(); ++i) a->seq_policy_mid.push_back(a->seq_actor_fb[i]);
  }
  if (fuse_tail) {
    // the actor's forward layers (they read the sampled states and the actor only) ride along with the target pass
    a->seq_policy_mid.clear();
    size_t ai = 0;
    // the actor's first layer (K = S <= 32) is a job of the front kernel between the two stages; its second layer
    // then rides with the stage after it
    bool l1_in_front = false;
    for (const Launch& L : a->seq_target) {
      Launch m = L;
      if (L.kind == Launch::FRONT && ai == 0 && a->n_actor_fwd >= 1 && L.front.n_nets < kFrontMaxNets && !L.front.gather) {
        FrontNet& n = m.front.net[m.front.n_nets++];
        n = front_first_layer(c.actor, Wa, 1, pa, rn_front);
        n.x = a->xpi; n.x_go = a->xpi_go; n.ldx = ld_q; n.x_off = 0; n.act_col = -1;
        front_finish(m, B, nA);
        l1_in_front = true;
        ++ai;
      } else if (L.kind == Launch::STAGE && l1_in_front && (int)ai < a->n_actor_fwd && merge_stage(m, a->seq_actor_fb[ai])) {
        ++ai;
      }
      v_tp.push_back(m);
    }
    for (; (int)ai < a->n_actor_fwd; ++ai) v_tp.push_back(a->seq_actor_fb[ai]);
  }

  // ---- layer-fused chains (chain.cuh), opt-in: TD3_CHAIN=1 ----
  // Plain MLPs of the featured variant in TF32 mode: the row-local part of the update is ONE launch (per policy_freq
  // phase), the weight gradients of layers >= 1 are one tensor-core stage, the first layer's gradient + the optimiser
  // are the apply launch: 3 launches for a critic-only update, 6 for a policy update.  Correct (the whole GPU suite
  // passes with it on) but MEASURED SLOWER than the stage-per-layer graph on B200 (cfg2: 120 / 260 us per critic-only /
  // policy update against 53 / 104 us; 8- and 16-agent populations: on par): one CTA per 64-row tile streams every
  // layer's full weight matrix through a 3-slot ring, and at ~2 k cycles of L2 round trip per slot that is ~25 B/clk per
  // SM -- a 400 x 300 layer takes 23 k cycles where the stage kernel spreads it over 20-130 SMs (DESIGN.md section 9,
  // profiles/r02b_chain_*).  Kept as the measured record of that design and as a second implementation the parity
  // tests cross-check; not on the default path.
  std::vector<Launch> v_ch_cdw, v_ch_adw;
  Launch ch_A, ch_Apol, ch_P;
  bool chain = false;
  {
    auto widths_ok = [&](const td3_net_layout& n) {
      if (n.n_linear < 2 || n.n_linear - 1 > kChMaxMask) return false;
      for (int l = 1; l < n.n_linear; ++l)
        if ((n.dims[l] & 3) || n.dims[l] > kChMaxW || n.dims[l] < 8) return false;
      if ((n.n_floats & 3) != 0) return false;
      for (int l = 1; l < n.n_linear; ++l)
        if (n.w_off[l] & 3) return false;
      return true;
    };
    chain = fuse_tail && tf && !enc && !ln && !wn && qw == 1 && nq >= 1 && nq <= 2 && S + A <= 32 && A <= 8 &&
            widths_ok(c.q) && widths_ok(c.actor) && a->global_batch == 0 && a->batch_offset == 0 && a->sh_c != nullptr &&
            encode_tiled_fn() != nullptr && getenv("TD3_CHAIN") && !getenv("TD3_NO_CHAIN");
  }
  if (chain) {
    EncodeTiledFn encf = encode_tiled_fn();
    bool ok = true;
    // 3-D tensor map over the weight matrix at `w` ([rows][cols] fp32, row-major) of `n_nets` networks `stride` floats apart
    auto add_map = [&](Launch& L, int& n_maps, const float* w, long long stride, int n_nets, int rows, int cols, bool mn, int nb) {
      if (n_maps >= kChMaxMaps || !aligned16(w) || (cols & 3) || (stride & 3)) { ok = false; return 0; }
      CUtensorMap m;
      cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)n_nets};
      cuuint64_t gstride[2] = {(cuuint64_t)cols * 4, (cuuint64_t)std::max<long long>(stride, 4) * 4};
      cuuint32_t box[3] = {32, mn ? 32u : (cuuint32_t)nb, 1}, estr[3] = {1, 1, 1};
      CUresult r = encf(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(w), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { ok = false; return 0; }
      static_assert(sizeof(CUtensorMap) == sizeof(TensorMapBlob), "tensor map blob size");
      memcpy(&L.chain.maps[n_maps], &m, sizeof(m));
      return n_maps++;
    };
    struct Fam { const float* P; const float* Psh; long long go, gi; int z_o, z_i, n_nets; long long stride; };
    const Fam f_at{a->actor.target, a->sh_at, an, 0, 1, 0, nA, an};
    const Fam f_a{a->actor.params, a->sh_a, an, 0, 1, 0, nA, an};
    const Fam f_ct{a->critic.target, a->sh_ct, qn * nq, qn, nq, 1, nA * nq, qn};
    const Fam f_c{a->critic.params, a->sh_c, qn * nq, qn, nq, 1, nA * nq, qn};
    const Fam f_q1{a->critic.params, a->sh_c, qn * nq, 0, nq, 0, nA * nq, qn};
    int slot_need = 4096, maxw = 32;
    auto blocks = [&](ChainStep& st) {
      st.n_blk = (st.N + 255) / 256;
      st.nb = (int)round_up((st.N + st.n_blk - 1) / st.n_blk, 8);
      const bool mn = st.b_mode == CB_TMA_MN || st.b_mode == CB_MANUAL_MN;
      if (st.b_mode == CB_MANUAL_MN) { st.n_blk = 1; st.nb = (int)round_up(st.N, 8); }
      st.blk_bytes = mn ? (st.b_mode == CB_MANUAL_MN ? 4096 : ((st.nb + 31) / 32) * 4096) : st.nb * 128;
      slot_need = std::max(slot_need, (int)round_up(st.blk_bytes, 1024));
      maxw = std::max(maxw, std::max(st.a_x0 ? 32 : st.K, st.N));
    };
    auto base_step = [&](const Fam& f) {
      ChainStep st;
      memset(&st, 0, sizeof(st));
      st.P = f.P; st.Psh = f.Psh; st.p_go = f.go; st.p_gi = f.gi; st.z_o = f.z_o; st.z_i = f.z_i;
      st.bias_off = -1; st.mask_w = st.mask_r = -1; st.map = -1;
      return st;
    };
    // forward through linear layer l of `n` (hidden layers: ReLU; the actor's output layer: tanh)
    auto fwd_step = [&](Launch& L, int& n_maps, const td3_net_layout& n, const Fam& f, int l, bool out_layer) {
      ChainStep st = base_step(f);
      st.K = n.dims[l]; st.N = n.dims[l + 1];
      st.w_off = n.w_off[l]; st.bias_off = n.b_off[l];
      st.epi = out_layer ? CE_ACTOR_OUT : CE_HIDDEN;
      if (l == 0) { st.b_mode = CB_MANUAL_K; st.a_x0 = 1; st.ld_w = n.dims[0]; }
      else st.b_mode = CB_TMA_K;
      blocks(st);
      if (l > 0) st.map = add_map(L, n_maps, f.Psh + n.w_off[l], f.stride, f.n_nets, n.dims[l + 1], n.dims[l], false, st.nb);
      return st;
    };
    // backward through linear layer l >= 1: d(in) = dZ_l . W_l, gated by the ReLU of hidden layer l - 1
    auto bwd_step = [&](Launch& L, int& n_maps, const td3_net_layout& n, const Fam& f, int l) {
      ChainStep st = base_step(f);
      st.K = n.dims[l + 1]; st.N = n.dims[l];
      st.w_off = n.w_off[l];
      st.epi = CE_BWD_MASK; st.b_mode = CB_TMA_MN;
      blocks(st);
      st.map = add_map(L, n_maps, f.Psh + n.w_off[l], f.stride, f.n_nets, n.dims[l + 1], n.dims[l], true, st.nb);
      return st;
    };
    auto push = [&](Launch& L, int& n_steps, const ChainStep& st) {
      if (n_steps >= kChMaxSteps) { ok = false; return; }
      L.chain.steps[n_steps++] = st;
    };
    auto common = [&](Launch& L) {
      L = Launch{};
      L.kind = Launch::CHAIN;
      ChainParams& P = L.chain;
      memset(&P, 0, sizeof(P));
      P.batch = B; P.n_agents = nA; P.tiles = ch_tiles; P.S = S; P.A = A; P.n_q = nq; P.ld_q = ld_q; P.ld_tanh = A;
      P.xq = a->xq; P.xq_go = a->xq_go; P.xq2 = a->xq2; P.xq2_go = a->xq_go; P.xpi = a->xpi; P.xpi_go = a->xpi_go;
      P.r = a->r; P.nd = a->nd; P.r_go = B;
      P.tanh_y = a->tanh_y; P.tanh_go = (long long)B * A;
      P.q = a->q; P.tq = a->tq; P.dq = a->dq; P.q_gi = B; P.q_go = (long long)B * nq;
      P.y = a->y; P.y_go = B; P.q_pi = a->q_pi; P.qpi_go = B;
      P.lp_go = (long long)ch_tiles * nq;
      P.loss = a->state_f32;
      P.flags = ch_sync + 64; P.flag_go = (long long)ch_tiles * nq;
      P.epoch = ch_sync; P.done = ch_sync + 1;
      P.discount = c.discount; P.inv_norm = 1.f / (float)(batch * qw);
      P.max_action = c.max_action; P.clamp_action = c.clamp_target_action ? c.max_action : 0.f;
      P.prof = getenv("TD3_CHAIN_PROF") ? a->prof_dev : nullptr;
    };
    const int Lq_ = c.q.n_linear;
    // critic-side launch; with_actor: the online actor's forward pass rides along (policy steps)
    auto build_A = [&](Launch& L, bool with_actor) {
      common(L);
      ChainParams& P = L.chain;
      P.loss_part = ch_loss_part;
      P.fin_mode = 0;
      P.tick = AdamTick{a->state_u64, 0, 0, c.lr_critic, c.beta1, c.beta2};
      P.host_status = a->host_status_live ? a->host_status : nullptr; P.seq = head_seq;
      int n_steps = 0, n_maps = 0, n_roles = 0, first = 0;
      auto begin_role = [&](int kind, int inner) {
        P.role_kind[n_roles] = kind; P.role_inner[n_roles] = inner; P.role_first[n_roles] = first; P.role_step0[n_roles] = n_steps;
      };
      auto end_role = [&](int inner) {
        P.role_nsteps[n_roles] = n_steps - P.role_step0[n_roles];
        first += nA * inner * ch_tiles;
        ++n_roles;
        P.role_first[n_roles] = first;
      };
      // TARGET: target actor, then target critic twin t on [s', a']
      begin_role(CR_TARGET, nq);
      for (int l = 0; l + 1 < La; ++l) push(L, n_steps, fwd_step(L, n_maps, c.actor, f_at, l, false));
      { ChainStep st = fwd_step(L, n_maps, c.actor, f_at, La - 1, true); st.flavour = 0; push(L, n_steps, st); }
      for (int l = 0; l + 1 < Lq_; ++l) {
        ChainStep st = fwd_step(L, n_maps, c.q, f_ct, l, false);
        if (l == Lq_ - 2) { st.post = CP_TQ; st.head_w_off = c.q.w_off[Lq_ - 1]; st.head_b_off = c.q.b_off[Lq_ - 1]; }
        push(L, n_steps, st);
      }
      end_role(nq);
      if (with_actor) {
        begin_role(CR_ACTOR, 1);
        for (int l = 0; l + 1 < La; ++l) {
          ChainStep st = fwd_step(L, n_maps, c.actor, f_a, l, false);
          st.out = pa.r[l]; st.out_go = pa.h_go[l]; st.out_gi = pa.h_gi[l]; st.ld_out = c.actor.dims[l + 1];
          push(L, n_steps, st);
        }
        { ChainStep st = fwd_step(L, n_maps, c.actor, f_a, La - 1, true); st.flavour = 1; push(L, n_steps, st); }
        end_role(1);
      }
      // CRITIC: online critic twin t forward, loss gradient, dX chain
      begin_role(CR_CRITIC, nq);
      for (int l = 0; l + 1 < Lq_; ++l) {
        ChainStep st = fwd_step(L, n_maps, c.q, f_c, l, false);
        st.out = cc.r[l]; st.out_go = cc.h_go[l]; st.out_gi = cc.h_gi[l]; st.ld_out = c.q.dims[l + 1];
        st.mask_w = l;
        if (l == Lq_ - 2) {
          st.post = CP_CRITIC; st.head_w_off = c.q.w_off[Lq_ - 1]; st.head_b_off = c.q.b_off[Lq_ - 1];
          st.out2 = ch_cz[l]; st.out2_gi = (long long)B * c.q.dims[l + 1]; st.out2_go = st.out2_gi * nq;
        }
        push(L, n_steps, st);
      }
      for (int l = Lq_ - 2; l >= 1; --l) {
        ChainStep st = bwd_step(L, n_maps, c.q, f_c, l);
        st.mask_r = l - 1;
        st.out = ch_cz[l - 1]; st.out_gi = (long long)B * c.q.dims[l]; st.out_go = st.out_gi * nq; st.ld_out = c.q.dims[l];
        push(L, n_steps, st);
      }
      end_role(nq);
      P.n_roles = n_roles; P.n_ctas = first;
    };
    // policy launch: Q1(s, pi(s)) with the stepped critic, -mean, backward to the action and through the actor
    auto build_P = [&](Launch& L) {
      common(L);
      ChainParams& P = L.chain;
      P.loss_part = ch_loss_part + (long long)nA * ch_tiles * nq;
      P.fin_mode = 1;
      if (P.prof) P.prof += 4 * 48;
      P.tick = AdamTick{a->state_u64, 1, 0, c.lr_actor, c.beta1, c.beta2};
      int n_steps = 0, n_maps = 0;
      P.role_kind[0] = CR_POLICY; P.role_inner[0] = 1; P.role_first[0] = 0; P.role_step0[0] = 0;
      for (int l = 0; l + 1 < Lq_; ++l) {
        ChainStep st = fwd_step(L, n_maps, c.q, f_q1, l, false);
        st.mask_w = l;
        if (l == Lq_ - 2) { st.post = CP_Q1; st.head_w_off = c.q.w_off[Lq_ - 1]; st.head_b_off = c.q.b_off[Lq_ - 1]; st.ld_out = c.q.dims[l + 1]; }
        push(L, n_steps, st);
      }
      for (int l = Lq_ - 2; l >= 1; --l) {
        ChainStep st = bwd_step(L, n_maps, c.q, f_q1, l);
        st.mask_r = l - 1;
        push(L, n_steps, st);
      }
      {   // d([s, a]) = dZ_0 . W_0, the action columns through max_action * tanh
        ChainStep st = base_step(f_q1);
        st.K = c.q.dims[1]; st.N = c.q.dims[0]; st.w_off = c.q.w_off[0]; st.ld_w = c.q.dims[0];
        st.b_mode = CB_MANUAL_MN; st.epi = CE_DX_TANH;
        st.out = ch_az[La - 1]; st.out_go = (long long)B * A; st.ld_out = A;
        blocks(st);
        push(L, n_steps, st);
      }
      {   // through the actor's output layer: K = A
        ChainStep st = bwd_step(L, n_maps, c.actor, f_a, La - 1);
        st.gmask = pa.r[La - 2]; st.gmask_go = pa.h_go[La - 2]; st.gmask_gi = 0;
        st.out = ch_az[La - 2]; st.out_go = (long long)B * c.actor.dims[La - 1]; st.ld_out = c.actor.dims[La - 1];
        push(L, n_steps, st);
      }
      for (int l = La - 2; l >= 1; --l) {
        ChainStep st = bwd_step(L, n_maps, c.actor, f_a, l);
        st.gmask = pa.r[l - 1]; st.gmask_go = pa.h_go[l - 1]; st.gmask_gi = 0;
        st.out = ch_az[l - 1]; st.out_go = (long long)B * c.actor.dims[l]; st.ld_out = c.actor.dims[l];
        push(L, n_steps, st);
      }
      P.role_nsteps[0] = n_steps;
      P.role_first[1] = nA * ch_tiles;
      P.n_roles = 1; P.n_ctas = nA * ch_tiles;
    };
    build_A(ch_A, false);
    build_A(ch_Apol, true);
    build_P(ch_P);
    // shared-memory geometry (the same for the three launches: the worst case of all steps)
    const int act_bytes = kChRows * 128 * ((maxw + 31) / 32);
    const int n_mask = std::max(Lq_ - 1, 1);
    const int fixed = kChainSmemFixed + act_bytes + n_mask * kChMaskBytes;
    const int n_slots = std::min(kChMaxSlots, (kChainSmemMax - fixed) / slot_need);
    if (n_slots < 2) ok = false;
    for (Launch* L : {&ch_A, &ch_Apol, &ch_P}) {
      L->chain.act_bytes = act_bytes; L->chain.slot_bytes = slot_need; L->chain.n_slots = n_slots; L->chain.n_mask = n_mask;
      L->smem_bytes = fixed + n_slots * slot_need;
      for (int i = 0; i < kChMaxSteps; ++i) {        // small column blocks: several 32-step chunks per ring-slot use
        ChainStep& st = L->chain.steps[i];
        if (st.K <= 0) continue;
        const int nch = (st.K + 31) / 32;
        st.cps = (st.n_blk > 1 || st.b_mode == CB_MANUAL_K) ? 1 : std::max(1, std::min(nch, slot_need / st.blk_bytes));
      }
    }
    // weight gradients of layers >= 1 (dW_l = dZ_l^T . H_{l-1}, db_l = column sums of dZ_l): one stage per family
    auto dw_stage = [&](std::vector<Launch>& seq, const td3_net_layout& n, GradRef G, GroupShape gs, const PassBuf& pb,
                        float* const* dzs, const float* dz_top, int ld_top, long long top_go, long long top_gi, int n_inner) {
      ProblemList stage;
      for (int l = n.n_linear - 1; l >= 1; --l) {
        const int K = n.dims[l], N = n.dims[l + 1];
        const bool top = l == n.n_linear - 1;
        const float* dz = top ? dz_top : dzs[l];
        const int ld_dz = top ? ld_top : N;
        const long long dz_gi = top ? top_gi : (long long)B * N, dz_go = top ? top_go : (long long)B * N * n_inner;
        Problem p = make_gemm(N, K, B, dz, ld_dz, false, pb.r[l - 1], K, false, G.base + n.w_off[l], K, EPI_STORE);
        set_groups(p, dz_go, dz_gi, pb.h_go[l - 1], pb.h_gi[l - 1], G.go, G.gi);
        finalize_problem(p, gs);
        if (p.use_tc) {
          Problem cs = blank_problem(PK_COLSUM);
          cs.N = N; cs.K = B; cs.A = dz; cs.lda = ld_dz; cs.a_go = dz_go; cs.a_gi = dz_gi;
          cs.C = G.base + n.b_off[l]; cs.c_go = G.go; cs.c_gi = G.gi;
          finalize_problem(cs, gs);
          stage.push_back(p);
          stage.push_back(cs);
        } else {
          p.aux1 = G.base + n.b_off[l]; p.aux1_go = G.go; p.aux1_gi = G.gi;
          finalize_problem(p, gs);
          stage.push_back(p);
        }
      }
      emit_stage(seq, stage);
    };
    dw_stage(v_ch_cdw, c.q, Gc, g_crit, cc, ch_cz, a->dq, qw, (long long)B * qw * nq, (long long)B * qw, nq);
    dw_stage(v_ch_adw, c.actor, Ga, g_actor, pa, ch_az, ch_az[La - 1], A, (long long)B * A, 0, 1);
    if (!ok) chain = false;
    cudaMemset(ch_sync, 0, (size_t)(64 + (long long)nA * ch_tiles * nq) * sizeof(unsigned int));
  }
  a->chain_on = chain;
  a->seq_chain_critic.clear(); a->seq_chain_policy.clear();
  // TMA descriptors of every tensor-core operand (pointers are fixed from here on: torch owns the buffers).  Merged
  // sequences get their own: merging stages can change a launch's cluster size and with it the A-panel box.
  if (g_tc_mode) {
    std::vector<CUtensorMap> host;
    if (!attach_tensor_maps({&a->seq_target, &a->seq_critic_fb, &a->seq_actor_fb, &a->seq_policy_mid, &v_cb, &v_abf, &v_tp, &v_ch_cdw, &v_ch_adw},
                            a->tmaps_dev, host))
      return fail(TD3_ERR_CUDA, "cuTensorMapEncodeTiled failed for a tensor-core operand (or more than %d maps needed)", kMaxTensorMaps);
    if (!host.empty())
      cudaMemcpy(a->tmaps_dev, host.data(), host.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  }
  // ---- what train_n runs after the sampling launch ----
  a->seq_run_critic.clear(); a->seq_run_policy.clear();
  auto append = [](std::vector<Launch>& dst, const std::vector<Launch>& src, size_t from = 0, size_t to = (size_t)-1) {
    for (size_t i = from; i < std::min(to, src.size()); ++i) dst.push_back(src[i]);
  };
  if (fuse_tail) {
    auto fill_dw = [&](DwParams& D, const td3_net_layout& n, const Dz0Info& z, const PassBuf& pb, const td3_param_set& ps,
                       int n_inner, long long net_floats, const float* sc_ptr, bool polyak, float* sh, float* sh_t) {
      memset(&D, 0, sizeof(D));
      D.batch = B; D.n_agents = nA; D.n_inner = n_inner; D.N = n.dims[1]; D.K = n.dims[0];
      D.rows_per_pass = nA > 1 ? kDwRowsSmall : kDwRows;
      D.col_blocks = (D.N + kDwCols - 1) / kDwCols;
      D.n_tiles = nA * n_inner * D.col_blocks;
      D.dz = z.dz; D.ld_dz = z.ld; D.dz_go = z.go; D.dz_gi = z.gi;
      D.x = pb.x0; D.ldx = pb.ld0; D.x_go = pb.x0_go; D.x_gi = pb.x0_gi;
      D.p = ps.params; D.g = ps.grad; D.m = ps.exp_avg; D.v = ps.exp_avg_sq; D.tgt = polyak ? ps.target : nullptr;
      D.do_polyak = polyak ? 1 : 0;
      D.p_sh = sh; D.tgt_sh = polyak ? sh_t : nullptr;
      D.p_go = net_floats * n_inner; D.p_gi = net_floats; D.w_off = n.w_off[0]; D.b_off = n.b_off[0];
      D.sc_ptr = sc_ptr;
    };
    Launch apply_c = a->seq_critic_apply[0];
    apply_c.ew.r[0].skip_period = qn; apply_c.ew.r[0].skip_len = c.q.w_off[1];
    fill_dw(apply_c.dw, c.q, dz0_c, cc, a->critic, nq, qn, reinterpret_cast<const float*>(a->state_u64 + 10), false, a->sh_c, a->sh_ct);
    Launch apply_a = a->seq_actor_apply[0];
    apply_a.ew.r[1].skip_period = an; apply_a.ew.r[1].skip_len = c.actor.w_off[1];
    fill_dw(apply_a.dw, c.actor, dz0_a, pa, a->actor, 1, an, reinterpret_cast<const float*>(a->state_u64 + 11), true, a->sh_a, a->sh_at);
    if (!dz0_c.dz || !dz0_a.dz || (dz0_c.ld & 3) || (dz0_a.ld & 3))
      return fail(TD3_ERR_STATE, "tail fusion: first-layer gradient not located");
    append(a->seq_run_critic, a->seq_target);
    append(a->seq_run_critic, v_cb);
    a->seq_run_critic.push_back(apply_c);
    append(a->seq_run_policy, v_tp);
    append(a->seq_run_policy, v_cb);
    a->seq_run_policy.push_back(apply_c);
    append(a->seq_run_policy, a->seq_actor_fb, (size_t)a->n_actor_fwd, n_actor_pre_bwd);
    append(a->seq_run_policy, v_abf);
    a->seq_run_policy.push_back(apply_a);
    if (chain) {
      // the apply launches of the chains read the first-layer dZ the chains wrote
      Launch apc = apply_c, apa = apply_a;
      apc.dw.dz = ch_cz[0]; apc.dw.ld_dz = c.q.dims[1]; apc.dw.dz_gi = (long long)B * c.q.dims[1]; apc.dw.dz_go = apc.dw.dz_gi * nq;
      apa.dw.dz = ch_az[0]; apa.dw.ld_dz = c.actor.dims[1]; apa.dw.dz_gi = 0; apa.dw.dz_go = (long long)B * c.actor.dims[1];
      a->seq_chain_critic.push_back(ch_A);
      append(a->seq_chain_critic, v_ch_cdw);
      a->seq_chain_critic.push_back(apc);
      a->seq_chain_policy.push_back(ch_Apol);
      append(a->seq_chain_policy, v_ch_cdw);
      a->seq_chain_policy.push_back(apc);
      a->seq_chain_policy.push_back(ch_P);
      append(a->seq_chain_policy, v_ch_adw);
      a->seq_chain_policy.push_back(apa);
    }
  } else {
    append(a->seq_run_critic, a->seq_target);
    append(a->seq_run_critic, a->seq_critic_fb);
    append(a->seq_run_critic, a->seq_critic_apply);
    append(a->seq_run_policy, a->seq_target);
    append(a->seq_run_policy, a->seq_policy_mid);
    append(a->seq_run_policy, a->seq_actor_apply);
  }
  // constant buffers: dq_pi = -1/(B*qw) (d(-mean)/dQ1), identity for the slice problem
  {
    std::vector<float> h((size_t)nA * B * qw, -1.f / (float)((a->global_batch > 0 ? a->global_batch : batch) * qw));
    cudaMemcpy(a->dq_pi, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
    std::vector<float> I((size_t)A * A, 0.f);
    for (int i = 0; i < A; ++i) I[(size_t)i * A + i] = 1.f;
    cudaMemcpy(eye, I.data(), I.size() * sizeof(float), cudaMemcpyHostToDevice);
  }
  a->batch = batch;
  a->plan_rows = nullptr;
  a->plan_rng_mode = -1;
  a->prog_dirty = true;
  a->shadow_dirty = true;
  drop_graphs(a);
  return TD3_OK;
}

// (re)build the sampling launch for a replay view / rng mode
int plan_sample(td3_agent* a, const td3_replay_view* rb, int rng_mode) {
  if (a->plan_rows == rb->rows && a->plan_row_stride == rb->row_stride && a->plan_rng_mode == rng_mode &&
      a->plan_rb_agent_stride == rb->agent_stride && !a->seq_sample.empty())
    return TD3_OK;
  const td3_agent_config& c = a->cfg;
  const int B = (int)a->batch, nA = c.n_agents, nq = c.n_q;
  const bool enc = c.variant == TD3_VARIANT_PARTICLES;
  const int A = c.action_dim, S = c.state_dim, E = enc ? c.actor.enc_out : 0;
  const int ld_q = a->ld_q, ld_a = a->ld_a;
  Launch L;
  L.kind = Launch::GATHER;
  GatherParams& G = L.gather;
  memset(&G, 0, sizeof(G));
  G.rows = rb->rows; G.row_stride = rb->row_stride; G.rb_agent_stride = rb->agent_stride;
  G.size = 0; G.size_ptr = a->state_u64 + 3;
  G.batch = B; G.n_agents = nA; G.rng_mode = rng_mode;
  G.idx_in = a->idx_in; G.idx_out = a->idx; G.step_ptr = a->state_u64; G.seed = c.seed;
  G.eps_out = a->eps; G.noise_in = a->noise_in; G.action_dim = A;
  G.policy_noise = c.policy_noise; G.noise_clip = c.noise_clip;
  G.elem_offset = (int)a->batch_offset;
  int n = 0;
  // TF32 mode, plain gather launch: the network inputs (not the particle sets, rewards or flags) are operands of
  // tensor-core first layers -> stored rounded to nearest TF32.  The front kernel's first layers are fp32 FFMA.
  const bool rn_inputs = c.precision == TD3_PRECISION_TF32 && !a->front_on;
  auto seg = [&](int off, int len, float* dst, int ld, long long astride, bool net_input = true) {
    G.seg_off[n] = off; G.seg_len[n] = len; G.dst[n] = dst; G.dst_ld[n] = ld; G.dst_agent_stride[n] = astride;
    if (net_input && rn_inputs && n < 32) G.seg_rn |= 1u << n;
    ++n;
  };
  const int xq_inner = enc ? nq : 1;
  if (!enc) {
    // row = [s | a | s2 | r | nd]
    seg(0, S, a->xq, ld_q, a->xq_go);
    seg(0, S, a->xpi, ld_q, a->xpi_go);
    seg(S, A, a->xq + S, ld_q, a->xq_go);
    seg(S + A, S, a->xq2, ld_q, a->xq_go);
    seg(2 * S + A, 1, a->r, 1, B, false);
    seg(2 * S + A + 1, 1, a->nd, 1, B, false);
  } else {
    // row = [particles | particles2 | f | a | f2 | r | nd], the particle sets padded to 16-byte multiples (bulk copies)
    const int PN = c.n_particles * c.particle_dim, PNp = (PN + 3) / 4 * 4;
    const int o_p = 0, o_p2 = PNp, o_f = 2 * PNp, o_a = o_f + S, o_f2 = o_a + A, o_r = o_f2 + S;
    seg(o_f, S, a->xa + E, ld_a, a->xa_go);
    seg(o_f, S, a->xpi + E, ld_q, a->xpi_go);
    for (int g = 0; g < xq_inner; ++g) seg(o_f, S, a->xq + g * a->xq_gi + E, ld_q, a->xq_go);
    for (int g = 0; g < xq_inner; ++g) seg(o_a, A, a->xq + g * a->xq_gi + E + S, ld_q, a->xq_go);
    seg(o_f2, S, a->xa2 + E, ld_a, a->xa_go);
    for (int g = 0; g < xq_inner; ++g) seg(o_f2, S, a->xq2 + g * a->xq_gi + E, ld_q, a->xq_go);
    seg(o_p, PN, a->P, PN, (long long)B * PN, false);
    seg(o_p2, PN, a->P2, PN, (long long)B * PN, false);
    seg(o_r, 1, a->r, 1, B, false);
    seg(o_r + 1, 1, a->nd, 1, B, false);
  }
  if (n > kMaxSeg) return fail(TD3_ERR_INVALID, "too many gather segments (%d)", n);
  G.n_seg = n;
  // the chain launches sample their rows themselves (chain.cuh): same view, same index / noise source
  for (auto* seq : {&a->seq_chain_critic, &a->seq_chain_policy})
    if (!seq->empty()) (*seq)[0].chain.g = G;
  const long long jobs = (long long)nA * B;
  L.grid_x = (int)((jobs + 7) / 8);
  L.grid_y = rb->row_floats > 2048 ? (int)std::min<long long>(32, (rb->row_floats + 2047) / 2048) : 1;
  G.slices = L.grid_y;
  if (a->front_on) {       // sampling + the first layers that read nothing but the sampled rows, in one launch
    G.slices = 1;
    G.row_floats = 2 * S + A + 2;
    const GatherParams gp = G;
    L = Launch{};
    L.front = a->front_sample;
    L.front.g = gp;
    front_finish(L, B, nA);
  }
  a->seq_sample.clear();
  a->seq_sample.push_back(L);
  a->plan_rows = rb->rows; a->plan_row_stride = rb->row_stride; a->plan_rng_mode = rng_mode;
  a->plan_rb_agent_stride = rb->agent_stride;
  a->prog_dirty = true;
  return TD3_OK;
}

// ------------------------------------------------------------------------------------
// persistent-kernel programs: the launch sequences re-expressed as stage records
// ------------------------------------------------------------------------------------
int append_records(std::vector<StageRec>& prog, const std::vector<Launch>& seq, AdamTick* pending_tick) {
  for (const Launch& L : seq) {
    if (L.kind == Launch::TICK) {             // folded into the next record: CTA 0 performs it on entry
      *pending_tick = L.tick;
      continue;
    }
    StageRec r;
    memset(&r, 0, sizeof(r));
    r.barrier_after = 1;
    r.tick = *pending_tick;
    *pending_tick = AdamTick{};
    switch (L.kind) {
      case Launch::STAGE:
        r.kind = SK_STAGE; r.u.st = L.stage; r.main_tiles = L.stage.total_tiles;
        break;
      case Launch::GATHER:
        r.kind = SK_GATHER; r.u.g = L.gather; r.main_tiles = L.grid_x * L.grid_y; r.gather_grid_x = L.grid_x;
        break;
      case Launch::LOSS:
        r.kind = SK_LOSS; r.u.l = L.loss; r.main_tiles = L.grid_x;
        break;
      case Launch::EW:
        r.kind = SK_EW_ONLY; r.ew = L.ew; r.ew_tiles = L.grid_x;
        if (L.dw.n_tiles > 0) { r.kind = SK_APPLY; r.u.d = L.dw; r.main_tiles = L.dw.n_tiles; }
        break;
      case Launch::HEAD:
        r.kind = SK_HEAD; r.u.h = L.head; r.main_tiles = L.grid_x;
        r.u.h.defer_finish = 0;                 // inside the persistent kernel the last CTA finishes
        break;
      case Launch::WN:
        r.kind = SK_WN; r.u.w = L.wn; r.main_tiles = L.grid_x;
        break;
      case Launch::FRONT:
        r.kind = SK_FRONT; r.u.f = L.front; r.main_tiles = L.grid_x;
        break;
      default:     // kernels with their own persistent CTA structure (fused set-encoder, chain) or host-side protocols (DP flags)
        return fail(TD3_ERR_UNSUPPORTED, "exec_mode persistent cannot run this plan (fused set-encoder / chain / data-parallel "
                                         "launches are kernels of their own): use exec_mode graph or launches");
    }
    prog.push_back(r);
  }
  return TD3_OK;
}

int build_programs(td3_agent* a, cudaStream_t s) {
  if (!a->prog_dirty) return TD3_OK;
  std::vector<StageRec> pc, pp;
  AdamTick pend{};
  // the same sequences the CUDA graphs replay (tail-fused where the plan allows it)
  for (auto* seq : {&a->seq_sample, &a->seq_run_critic})
    if (int rc = append_records(pc, *seq, &pend)) return rc;
  for (auto* seq : {&a->seq_sample, &a->seq_run_policy})
    if (int rc = append_records(pp, *seq, &pend)) return rc;
  if (pc.empty() || pp.empty() || (int)pc.size() > kMaxProgStages || (int)pp.size() > kMaxProgStages)
    return fail(TD3_ERR_STATE, "persistent program has %zu / %zu stages (max %d)", pc.size(), pp.size(), kMaxProgStages);
  // When the first stage of an update is a plain gather it touches nothing the optimiser stage before it writes (the
  // barrier after the gather covers the parameters): no barrier between updates.  A front sampling stage also runs
  // first layers, i.e. it READS parameters the previous update's optimiser stage is still writing on other CTAs.
  if (pc.front().kind == SK_GATHER) {
    pc.back().barrier_after = 0;
    pp.back().barrier_after = 0;
  }
  a->n_bar_critic = a->n_bar_policy = 0;
  for (auto& r : pc) a->n_bar_critic += r.barrier_after;
  for (auto& r : pp) a->n_bar_policy += r.barrier_after;
  CUDA_TRY(cudaMemcpyAsync(a->prog_dev, pc.data(), pc.size() * sizeof(StageRec), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(a->prog_dev + kMaxProgStages, pp.data(), pp.size() * sizeof(StageRec), cudaMemcpyHostToDevice, s));
  a->n_prog_critic = (int)pc.size();
  a->n_prog_policy = (int)pp.size();
  int max_tiles = 1;
  for (auto* p : {&pc, &pp})
    for (auto& r : *p) max_tiles = std::max(max_tiles, r.main_tiles + r.ew_tiles);
  int rc = ensure_kernel_attrs();
  if (rc != TD3_OK) return rc;
  int dev = 0, sms = 0, per_sm = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const bool tc = a->cfg.precision == TD3_PRECISION_TF32;
  if (tc) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, persistent_update_kernel<true>, kStageThreads, kDynSmemBytes));
  else CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, persistent_update_kernel<false>, kStageThreads, kDynSmemBytes));
  if (per_sm < 1) return fail(TD3_ERR_CUDA, "persistent kernel does not fit on an SM");
  if (const char* e = getenv("TD3_PERSIST_CTAS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e)));
  a->persist_grid = std::min(sms * per_sm, max_tiles);
  a->prog_dirty = false;
  return TD3_OK;
}

int launch_persistent(td3_agent* a, long long total_it, int iterations, cudaStream_t s) {
  PersistArgs args;
  args.prog_critic = a->prog_dev;
  args.prog_policy = a->prog_dev + kMaxProgStages;
  args.n_critic = a->n_prog_critic;
  args.n_policy = a->n_prog_policy;
  args.total_it = total_it;
  args.iterations = iterations;
  args.policy_freq = a->cfg.policy_freq;
  args.barrier = reinterpret_cast<unsigned int*>(a->state_u64 + 8);
  if (a->bar_reset) {
    CUDA_TRY(cudaMemsetAsync(args.barrier, 0, sizeof(unsigned int), s));
    a->bar_count = 0;
    a->bar_reset = false;
  }
  args.barrier_base = a->bar_count;
  unsigned int bar_add = 0;
  {   // barriers this launch will execute: one per stage except the last of every update
    const long long pf = a->cfg.policy_freq;
    const long long n_pol = (total_it + iterations) / pf - total_it / pf;
    const long long n_bar = n_pol * a->n_bar_policy + ((long long)iterations - n_pol) * a->n_bar_critic;
    bar_add = (unsigned int)(n_bar * a->persist_grid);
  }
  static const bool want_prof = getenv("TD3_PERSIST_PROF") != nullptr;
  args.prof = want_prof ? a->prof_dev : nullptr;
  void* kargs[] = {&args};
  const void* fn = a->cfg.precision == TD3_PRECISION_TF32 ? (const void*)persistent_update_kernel<true>
                                                           : (const void*)persistent_update_kernel<false>;
  cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(a->persist_grid), dim3(kStageThreads), kargs, (size_t)kDynSmemBytes, s);
  if (e != cudaSuccess) {
    a->bar_reset = true;
    return fail(TD3_ERR_CUDA, "cudaLaunchCooperativeKernel: %s", cudaGetErrorString(e));
  }
  a->bar_count += bar_add;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return TD3_OK;
}

int run_seq(const std::vector<Launch>& seq, cudaStream_t s) {
  for (const Launch& L : seq) {
    int rc = run_launch(L, s);
    if (rc != TD3_OK) return rc;
  }
  if (fork_join(s) != cudaSuccess) return fail(TD3_ERR_CUDA, "fork_join at the end of a launch sequence failed");
  return TD3_OK;
}

__global__ void set_u64_kernel(unsigned long long* p, unsigned long long v) { *p = v; }

// TF32 shadows of the parameters: rebuilt from the masters whenever they may have changed outside the optimiser kernels
int ensure_shadows(td3_agent* a, cudaStream_t s) {
  if (!a->shadow_dirty) return TD3_OK;
  if (a->sh_a && a->ws.base) {
    const td3_agent_config& c = a->cfg;
    const long long an = (long long)c.n_agents * c.actor.n_floats, cn = (long long)c.n_agents * c.n_q * c.q.n_floats;
    RoundCopyParams R;
    memset(&R, 0, sizeof(R));
    const float* src[4] = {a->actor.params, a->actor.target, a->critic.params, a->critic.target};
    float* dst[4] = {a->sh_a, a->sh_at, a->sh_c, a->sh_ct};
    const long long n[4] = {an, an, cn, cn};
    long long blk = 0;
    for (int i = 0; i < 4; ++i) {
      R.src[i] = src[i]; R.dst[i] = dst[i]; R.n[i] = n[i]; R.blk_begin[i] = blk;
      blk += (n[i] + kEwPerBlock - 1) / kEwPerBlock;
    }
    R.n_ranges = 4;
    round_copy_kernel<<<(unsigned)blk, kEwThreads, 0, s>>>(R);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
  }
  a->shadow_dirty = false;
  return TD3_OK;
}

int check_ready(td3_agent* a, bool need_plan = true) {
  if (!a) return fail(TD3_ERR_INVALID, "null agent");
  if (!a->params_bound || !a->state_u64) return fail(TD3_ERR_STATE, "td3_agent_bind_params/bind_state not called");
  if (need_plan && (a->batch <= 0 || !a->ws.base)) return fail(TD3_ERR_STATE, "td3_agent_plan not called");
  return TD3_OK;
}

int check_rb(td3_agent* a, const td3_replay_view* rb) {
  if (!rb || !rb->rows) return fail(TD3_ERR_INVALID, "null replay view");
  if (rb->size <= 0) return fail(TD3_ERR_INVALID, "cannot sample from an empty replay buffer (size=%lld)", (long long)rb->size);
  const td3_agent_config& c = a->cfg;
  long long want = c.variant == TD3_VARIANT_PARTICLES
                       ? 2LL * (c.state_dim + ((long long)c.n_particles * c.particle_dim + 3) / 4 * 4) + c.action_dim + 2
                       : 2LL * c.state_dim + c.action_dim + 2;
  if (rb->row_floats != want) return fail(TD3_ERR_INVALID, "replay row has %lld floats, agent expects %lld", (long long)rb->row_floats, want);
  return TD3_OK;
}

int sync_rb_size(td3_agent* a, const td3_replay_view* rb, cudaStream_t s) {
  if (a->last_rb_size != rb->size) {
    set_u64_kernel<<<1, 1, 0, s>>>(a->state_u64 + 3, (unsigned long long)rb->size);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    a->last_rb_size = rb->size;
  }
  return TD3_OK;
}

int capture(td3_agent* a, bool with_actor, cudaGraphExec_t* out, long long* n_nodes) {
  cudaGraph_t graph = nullptr;
  const long long launches_before = g_launches.load();
  if (!a->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&a->cap_stream, cudaStreamNonBlocking));
  cudaStream_t s = a->cap_stream;
  fork_ensure();
  CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int rc = a->chain_on ? TD3_OK : run_seq(a->seq_sample, s);
  if (rc == TD3_OK) rc = run_seq(a->chain_on ? (with_actor ? a->seq_chain_policy : a->seq_chain_critic)
                                             : (with_actor ? a->seq_run_policy : a->seq_run_critic), s);
  cudaError_t e = cudaStreamEndCapture(s, &graph);
  *n_nodes = g_launches.load() - launches_before;            // captured, not executed: move to per-replay accounting
  g_launches.fetch_sub(*n_nodes, std::memory_order_relaxed);
  if (rc != TD3_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return fail(TD3_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(out, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail(TD3_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  return TD3_OK;
}

}  // namespace

// ====================================================================================
// C ABI
// ====================================================================================
extern "C" {

int td3_abi_version(void) { return TD3_ABI_VERSION; }
void td3_struct_sizes(int64_t* out) {   // lets a foreign-language binding check its struct mirrors at load time
  out[0] = (int64_t)sizeof(td3_net_layout);
  out[1] = (int64_t)sizeof(td3_param_set);
  out[2] = (int64_t)sizeof(td3_agent_config);
  out[3] = (int64_t)sizeof(td3_replay_view);
}
const char* td3_last_error(void) { return g_err.c_str(); }
int64_t td3_launch_count(void) { return g_launches.load(); }
#ifdef TD3_TC_DEBUG
int td3_debug_set(int idx, int val) {
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpyToSymbol(td3::g_tc_dbg, &val, sizeof(int), idx * sizeof(int)));
  return TD3_OK;
}
#endif
#ifdef TD3_TILE_PROF
int td3_debug_tile_prof(long long* out) {
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpyFromSymbol(out, td3::g_tp, sizeof(long long) * 128 * 16));
  return TD3_OK;
}
#endif

int td3_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp p;
  CUDA_TRY(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (name && name_len > 0) {
    strncpy(name, p.name, name_len - 1);
    name[name_len - 1] = 0;
  }
  if (p.major != 10) return fail(TD3_ERR_UNSUPPORTED, "libtd3b200 is built for sm_100a only; device is sm_%d%d", p.major, p.minor);
  return TD3_OK;
}

int rb_add_rows(float* rows, int64_t row_stride, int64_t row_floats, int64_t max_size, int64_t ptr, const float* host_rows,
                int64_t n_rows, void* stream) {
  if (!rows || !host_rows || n_rows < 0 || ptr < 0 || ptr >= max_size || n_rows > max_size || row_floats > row_stride)
    return fail(TD3_ERR_INVALID, "rb_add_rows: bad arguments (ptr=%lld n=%lld max=%lld)", (long long)ptr, (long long)n_rows, (long long)max_size);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t first = std::min<int64_t>(n_rows, max_size - ptr);
  if (n_rows == 1) {   // ReplayBuffer_*.add: one contiguous row
    CUDA_TRY(cudaMemcpyAsync(rows + ptr * row_stride, host_rows, row_floats * sizeof(float), cudaMemcpyHostToDevice, s));
    return TD3_OK;
  }
  // host rows are packed row_floats apart; device rows row_stride apart
  if (first > 0)
    CUDA_TRY(cudaMemcpy2DAsync(rows + ptr * row_stride, row_stride * sizeof(float), host_rows, row_floats * sizeof(float),
                               row_floats * sizeof(float), first, cudaMemcpyHostToDevice, s));
  if (n_rows > first)
    CUDA_TRY(cudaMemcpy2DAsync(rows, row_stride * sizeof(float), host_rows + first * row_floats, row_floats * sizeof(float),
                               row_floats * sizeof(float), n_rows - first, cudaMemcpyHostToDevice, s));
  return TD3_OK;
}

int rb_sample_indices(const td3_replay_view* rb, const int64_t* idx_dev, int64_t batch, int32_t n_seg, const int64_t* seg_off,
                      const int64_t* seg_len, float* const* dst, const int64_t* dst_ld, void* stream) {
  if (!rb || !rb->rows || !idx_dev || batch <= 0 || n_seg <= 0 || n_seg > kMaxSeg)
    return fail(TD3_ERR_INVALID, "rb_sample_indices: bad arguments");
  unsigned long long* zero_step = static_cast<unsigned long long*>(device_scratch(0, sizeof(unsigned long long), true));
  long long* idx_scratch = static_cast<long long*>(device_scratch(1, (size_t)batch * sizeof(long long)));
  if (!zero_step || !idx_scratch) return fail(TD3_ERR_CUDA, "rb_sample_indices: scratch allocation failed");
  Launch L;
  L.kind = Launch::GATHER;
  GatherParams& G = L.gather;
  memset(&G, 0, sizeof(G));
  G.rows = rb->rows; G.row_stride = rb->row_stride; G.size = rb->size;
  G.batch = (int)batch; G.n_agents = 1; G.rng_mode = 1; G.n_seg = n_seg;
  for (int i = 0; i < n_seg; ++i) {
    G.seg_off[i] = (int)seg_off[i]; G.seg_len[i] = (int)seg_len[i]; G.dst[i] = dst[i]; G.dst_ld[i] = (int)dst_ld[i];
  }
  G.idx_in = reinterpret_cast<const long long*>(idx_dev);
  G.idx_out = idx_scratch;
  G.step_ptr = zero_step;
  G.action_dim = 0;
  L.grid_x = (int)((batch + 7) / 8);
  L.grid_y = rb->row_floats > 2048 ? (int)std::min<long long>(32, (rb->row_floats + 2047) / 2048) : 1;
  G.slices = L.grid_y;
  return run_launch(L, (cudaStream_t)stream);
}

int rb_philox_indices(int64_t* idx_dev, int64_t batch, int64_t size, uint64_t seed, uint64_t stream_id, uint64_t step,
                      void* stream) {
  if (!idx_dev || batch <= 0) return fail(TD3_ERR_INVALID, "rb_philox_indices: bad arguments");
  if (size <= 0) return fail(TD3_ERR_INVALID, "cannot sample from an empty replay buffer (size=%lld)", (long long)size);
  philox_indices_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<long long*>(idx_dev), batch, size, seed, (unsigned)stream_id, step);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CUDA_TRY(cudaGetLastError());
  return TD3_OK;
}

int adam_polyak_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, float* target, int64_t n, int64_t t,
                     double lr, double beta1, double beta2, double eps, double tau, void* stream) {
  if (!params || n <= 0) return fail(TD3_ERR_INVALID, "adam_polyak_step: bad arguments");
  if (!grad && !target) return fail(TD3_ERR_INVALID, "adam_polyak_step: nothing to do (grad and target both NULL)");
  if (grad && (!exp_avg || !exp_avg_sq || t < 1)) return fail(TD3_ERR_INVALID, "adam_polyak_step: Adam needs moments and t >= 1");
  Launch L;
  L.kind = Launch::EW;
  EwParams& e = L.ew;
  e.n_ranges = 1;
  e.beta1 = beta1; e.beta2 = beta2; e.eps = eps; e.tau = tau;
  EwRange& r = e.r[0];
  r.p = params; r.g = grad; r.m = exp_avg; r.v = exp_avg_sq; r.tgt = target; r.n = n; r.blk_begin = 0;
  r.sc_ptr = nullptr;
  r.step_size = (float)(lr / (1.0 - pow(beta1, (double)t)));      // torch/optim/adam.py: step_size, bias_correction2_sqrt
  r.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)t));
  r.do_adam = grad != nullptr; r.do_polyak = target != nullptr;
  L.grid_x = (int)((n + kEwPerBlock - 1) / kEwPerBlock);
  return run_launch(L, (cudaStream_t)stream);
}

int td3_gemm(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, int32_t a_rc, const float* B, int64_t ldb, int32_t b_rc,
             float* C, int64_t ldc, const float* bias, int32_t relu, int32_t use_tc, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || !A || !B || !C) return fail(TD3_ERR_INVALID, "td3_gemm: bad arguments");
  Problem p = make_gemm((int)M, (int)N, (int)K, A, (int)lda, a_rc != 0, B, (int)ldb, b_rc != 0, C, (int)ldc,
                        bias ? (relu ? EPI_BIAS_RELU : EPI_BIAS) : EPI_STORE);
  p.bias = bias;
  g_tc_mode = use_tc ? 1 : 0;
  g_cluster_mode = getenv("TD3_CLUSTER") ? 1 : 0;
  cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, 0);
  finalize_problem(p, GroupShape{1, 1});
  std::vector<Launch> seq;
  emit_stage(seq, {p});
  if (use_tc) {
    CUtensorMap* scratch = static_cast<CUtensorMap*>(device_scratch(2, 8 * sizeof(CUtensorMap)));
    if (!scratch) return fail(TD3_ERR_CUDA, "td3_gemm: scratch allocation failed");
    std::vector<CUtensorMap> host;
    if (!p.use_tc)
      return fail(TD3_ERR_UNSUPPORTED, "td3_gemm: operands are not eligible for the tcgen05 tile (TMA needs 16-byte aligned rows; K >= 64)");
    if (!attach_tensor_maps({&seq}, scratch, host)) return fail(TD3_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    if (!host.empty())
      CUDA_TRY(cudaMemcpyAsync(scratch, host.data(), host.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  }
  return run_seq(seq, (cudaStream_t)stream);
}

// ---- the particle-set encoder on caller buffers (set_encoder_fwd / set_encoder_bwd) ----------------------
namespace {
int run_problem_stages(std::vector<ProblemList>& stages, bool any_tc, cudaStream_t s) {
  std::vector<Launch> seq;
  for (auto& st : stages) {
    int rc = emit_stage(seq, st);
    if (rc != TD3_OK) return rc;
  }
  if (any_tc) {
    CUtensorMap* scratch = static_cast<CUtensorMap*>(device_scratch(3, 32 * sizeof(CUtensorMap)));
    if (!scratch) return fail(TD3_ERR_CUDA, "set_encoder: scratch allocation failed");
    std::vector<CUtensorMap> host;
    if (!attach_tensor_maps({&seq}, scratch, host) || host.size() > 32) return fail(TD3_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    if (!host.empty())
      CUDA_TRY(cudaMemcpyAsync(scratch, host.data(), host.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, s));
  }
  return run_seq(seq, s);
}
}  // namespace

int64_t set_encoder_workspace_floats(int64_t batch, int64_t n_particles, int64_t particle_dim, int64_t enc_hidden, int64_t enc_out) {
  const long long rows = batch * n_particles;
  return rows * (enc_hidden + enc_out) + kEncPartials * (enc_out * enc_hidden + enc_out) +
         kEncPartials * (enc_hidden * particle_dim + enc_hidden) + (rows / kEncTile + 1) * enc_out + 1024;
}

int set_encoder_fwd(const float* particles, int64_t batch, int64_t n_particles, int64_t particle_dim, int64_t enc_hidden,
                    int64_t enc_out, const float* conv1_w, const float* conv1_b, const float* conv2_w, const float* conv2_b,
                    float* pooled, int64_t ld_pooled, float* h1, float* h2, float* workspace, int64_t workspace_floats,
                    int32_t use_tc, void* stream) {
  if (!particles || !conv1_w || !conv1_b || !conv2_w || !conv2_b || !pooled || !workspace || batch <= 0 || n_particles <= 0 ||
      particle_dim <= 0 || enc_hidden <= 0 || enc_out <= 0 || enc_out > 256 || ld_pooled < enc_out)
    return fail(TD3_ERR_INVALID, "set_encoder_fwd: bad arguments");
  if (workspace_floats < set_encoder_workspace_floats(batch, n_particles, particle_dim, enc_hidden, enc_out))
    return fail(TD3_ERR_INVALID, "set_encoder_fwd: workspace too small (see set_encoder_workspace_floats)");
  const int B = (int)batch, N = (int)n_particles, D = (int)particle_dim, H = (int)enc_hidden, O = (int)enc_out;
  const long long rows = (long long)B * N;
  const GroupShape gs{1, 1};
  g_tc_mode = use_tc ? 1 : 0;
  g_cluster_mode = 0;
  cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, 0);
  float* h1w = h1 ? h1 : workspace;
  float* h2w = h2 ? h2 : workspace + rows * H;
  float* part = workspace + rows * (H + O);
  std::vector<ProblemList> st;
  if (use_tc) {
    if (H != kEncH || O != kEncO || D > 7 || rows % kEncTile != 0 || N % kEncTile != 0 || !encode_tiled_fn() || !aligned16(conv2_w))
      return fail(TD3_ERR_UNSUPPORTED, "set_encoder_fwd(use_tc): the fused tcgen05 encoder needs 256 / 128 channels, particle_dim <= 7 and "
                                       "n_particles a multiple of 128");
    Problem ef = blank_problem(PK_ENC_FUSED);
    ef.M = (int)rows; ef.K = D; ef.N = N;
    ef.A = particles;
    ef.B = conv1_w; ef.bias = conv1_b;
    ef.aux0 = const_cast<float*>(conv2_w); ef.aux1 = const_cast<float*>(conv2_b);
    ef.aux2 = h1; ef.aux3 = h2;
    ef.C = part;
    finalize_problem(ef, gs);
    st.push_back({ef});
    Problem pl = blank_problem(PK_POOL_FWD);
    pl.M = B; pl.N = O; pl.K = N / kEncTile;
    pl.A = part; pl.lda = O;
    pl.C = pooled; pl.ldc = (int)ld_pooled;
    finalize_problem(pl, gs);
    st.push_back({pl});
  } else {
    Problem e1 = make_gemm((int)rows, H, D, particles, D, true, conv1_w, D, true, h1w, H, EPI_BIAS_RELU);
    e1.bias = conv1_b;
    if (D <= 8) e1.kind = PK_SMALLK_FWD;
    finalize_problem(e1, gs);
    st.push_back({e1});
    Problem e2 = make_gemm((int)rows, O, H, h1w, H, true, conv2_w, H, true, h2w, O, EPI_BIAS_RELU);
    e2.bias = conv2_b;
    finalize_problem(e2, gs);
    st.push_back({e2});
    Problem pl = blank_problem(PK_POOL_FWD);
    pl.M = B; pl.N = O; pl.K = N;
    pl.A = h2w; pl.lda = O;
    pl.C = pooled; pl.ldc = (int)ld_pooled;
    finalize_problem(pl, gs);
    st.push_back({pl});
  }
  return run_problem_stages(st, false, (cudaStream_t)stream);
}

int set_encoder_bwd(const float* particles, int64_t batch, int64_t n_particles, int64_t particle_dim, int64_t enc_hidden,
                    int64_t enc_out, const float* conv2_w, const float* h1, const float* h2, const float* pooled, int64_t ld_pooled,
                    const float* d_pooled, int64_t ld_d_pooled, float* g_conv1_w, float* g_conv1_b, float* g_conv2_w,
                    float* g_conv2_b, float* workspace, int64_t workspace_floats, int32_t use_tc, void* stream) {
  if (!particles || !conv2_w || !h1 || !h2 || !pooled || !d_pooled || !g_conv1_w || !g_conv1_b || !g_conv2_w || !g_conv2_b ||
      !workspace || batch <= 0 || n_particles <= 0 || particle_dim <= 0 || enc_hidden <= 0 || enc_out <= 0)
    return fail(TD3_ERR_INVALID, "set_encoder_bwd: bad arguments");
  if (workspace_floats < set_encoder_workspace_floats(batch, n_particles, particle_dim, enc_hidden, enc_out))
    return fail(TD3_ERR_INVALID, "set_encoder_bwd: workspace too small (see set_encoder_workspace_floats)");
  const long long rows = batch * n_particles;
  g_tc_mode = use_tc ? 1 : 0;
  g_cluster_mode = 0;
  cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, 0);
  EncBwdArgs e;
  e.B = (int)batch; e.n_particles = (int)n_particles; e.D = (int)particle_dim; e.H = (int)enc_hidden; e.O = (int)enc_out;
  e.dpool = d_pooled; e.ld_dpool = (int)ld_d_pooled;
  e.pooled = pooled; e.ld_pooled = (int)ld_pooled;
  e.h1 = h1; e.h2 = h2; e.P = particles;
  e.dh1 = workspace; e.dh2 = workspace + rows * enc_hidden;
  e.part = workspace + rows * (enc_hidden + enc_out);
  e.W2 = conv2_w;
  e.gW1 = g_conv1_w; e.gb1 = g_conv1_b; e.gW2 = g_conv2_w; e.gb2 = g_conv2_b;
  std::vector<ProblemList> st = enc_backward_stages(e, GroupShape{1, 1});
  return run_problem_stages(st, use_tc != 0, (cudaStream_t)stream);
}

// The fused pair (enc.cuh / encbwd.cuh): the forward keeps 64 bytes of ReLU bitmaps per particle instead of h1 / h2, the
// backward recomputes from the particles and those bitmaps.
static bool enc_fused_shape_ok(int64_t batch, int64_t n_particles, int64_t particle_dim, const float* conv2_w) {
  return particle_dim >= 1 && particle_dim <= 7 && n_particles % kEncTile == 0 && batch >= 1 && encode_tiled_fn() != nullptr &&
         aligned16(conv2_w) && batch * n_particles < (1ll << 31);
}

int set_encoder_fwd_bits(const float* particles, int64_t batch, int64_t n_particles, int64_t particle_dim, const float* conv1_w,
                         const float* conv1_b, const float* conv2_w, const float* conv2_b, float* pooled, int64_t ld_pooled,
                         uint32_t* relu_bits, float* workspace, int64_t workspace_floats, void* stream) {
  if (!particles || !conv1_w || !conv1_b || !conv2_w || !conv2_b || !pooled || !relu_bits || !workspace || ld_pooled < kEncO)
    return fail(TD3_ERR_INVALID, "set_encoder_fwd_bits: bad arguments");
  if (!enc_fused_shape_ok(batch, n_particles, particle_dim, conv2_w))
    return fail(TD3_ERR_UNSUPPORTED, "set_encoder_fwd_bits: the fused tcgen05 encoder needs particle_dim <= 7 and n_particles a multiple of 128");
  if (workspace_floats < set_encoder_workspace_floats(batch, n_particles, particle_dim, kEncH, kEncO))
    return fail(TD3_ERR_INVALID, "set_encoder_fwd_bits: workspace too small (see set_encoder_workspace_floats)");
  const int B = (int)batch, N = (int)n_particles, D = (int)particle_dim;
  const long long rows = (long long)B * N;
  const GroupShape gs{1, 1};
  g_tc_mode = 1;
  g_cluster_mode = 0;
  cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, 0);
  float* part = workspace + rows * (kEncH + kEncO);
  std::vector<ProblemList> st;
  Problem ef = blank_problem(PK_ENC_FUSED);
  ef.M = (int)rows; ef.K = D; ef.N = N;
  ef.A = particles;
  ef.B = conv1_w; ef.bias = conv1_b;
  ef.aux0 = const_cast<float*>(conv2_w); ef.aux1 = const_cast<float*>(conv2_b);
  ef.tmapA = relu_bits;
  ef.C = part;
  finalize_problem(ef, gs);
  st.push_back({ef});
  Problem pl = blank_problem(PK_POOL_FWD);
  pl.M = B; pl.N = kEncO; pl.K = N / kEncTile;
  pl.A = part; pl.lda = kEncO;
  pl.C = pooled; pl.ldc = (int)ld_pooled;
  finalize_problem(pl, gs);
  st.push_back({pl});
  return run_problem_stages(st, false, (cudaStream_t)stream);
}

int set_encoder_bwd_fused(const float* particles, int64_t batch, int64_t n_particles, int64_t particle_dim, const float* conv1_w,
                          const float* conv1_b, const float* conv2_w, const uint32_t* relu_bits, const float* pooled,
                          int64_t ld_pooled, const float* d_pooled, int64_t ld_d_pooled, float* g_conv1_w, float* g_conv1_b,
                          float* g_conv2_w, float* g_conv2_b, float* workspace, int64_t workspace_floats, void* stream) {
  if (!particles || !conv1_w || !conv1_b || !conv2_w || !relu_bits || !pooled || !d_pooled || !g_conv1_w || !g_conv1_b || !g_conv2_w ||
      !g_conv2_b || !workspace)
    return fail(TD3_ERR_INVALID, "set_encoder_bwd_fused: bad arguments");
  if (!enc_fused_shape_ok(batch, n_particles, particle_dim, conv2_w))
    return fail(TD3_ERR_UNSUPPORTED, "set_encoder_bwd_fused: needs particle_dim <= 7 and n_particles a multiple of 128");
  if (workspace_floats < set_encoder_workspace_floats(batch, n_particles, particle_dim, kEncH, kEncO))
    return fail(TD3_ERR_INVALID, "set_encoder_bwd_fused: workspace too small (see set_encoder_workspace_floats)");
  const long long rows = batch * n_particles;
  g_tc_mode = 1;
  g_cluster_mode = 0;
  cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, 0);
  EncBwdArgs e;
  e.B = (int)batch; e.n_particles = (int)n_particles; e.D = (int)particle_dim; e.H = kEncH; e.O = kEncO;
  e.dpool = d_pooled; e.ld_dpool = (int)ld_d_pooled;
  e.pooled = pooled; e.ld_pooled = (int)ld_pooled;
  e.P = particles;
  e.part = workspace + rows * (kEncH + kEncO);
  e.W2 = conv2_w; e.W1 = conv1_w; e.b1 = conv1_b;
  e.bits = relu_bits;
  e.gW1 = g_conv1_w; e.gb1 = g_conv1_b; e.gW2 = g_conv2_w; e.gb2 = g_conv2_b;
  std::vector<ProblemList> st = enc_backward_stages(e, GroupShape{1, 1});
  return run_problem_stages(st, false, (cudaStream_t)stream);
}

int td3_agent_create(const td3_agent_config* cfg, td3_agent** out) {
  if (!cfg || !out) return fail(TD3_ERR_INVALID, "td3_agent_create: null argument");
  if (cfg->n_q < 1 || cfg->n_q > 2) return fail(TD3_ERR_INVALID, "n_q must be 1 or 2");
  if (cfg->n_agents < 1) return fail(TD3_ERR_INVALID, "n_agents must be >= 1");
  if (cfg->actor.n_linear < 1 || cfg->actor.n_linear > TD3_MAX_LINEAR || cfg->q.n_linear < 1 || cfg->q.n_linear > TD3_MAX_LINEAR)
    return fail(TD3_ERR_INVALID, "n_linear out of range");
  if (cfg->policy_freq < 1) return fail(TD3_ERR_INVALID, "policy_freq must be >= 1");
  if (cfg->variant == TD3_VARIANT_PARTICLES && (cfg->n_particles < 1 || cfg->particle_dim < 1))
    return fail(TD3_ERR_INVALID, "particles variant needs n_particles, particle_dim >= 1");
  if (cfg->actor.dims[cfg->actor.n_linear] != cfg->action_dim) return fail(TD3_ERR_INVALID, "actor output width != action_dim");
  td3_agent* a = new td3_agent();
  a->cfg = *cfg;
  *out = a;
  return TD3_OK;
}

int td3_agent_destroy(td3_agent* agent) {
  if (!agent) return TD3_OK;
  drop_graphs(agent);
  if (agent->cap_stream) cudaStreamDestroy(agent->cap_stream);
  delete agent;
  return TD3_OK;
}

int td3_agent_bind_params(td3_agent* a, const td3_param_set* actor, const td3_param_set* critic) {
  if (!a || !actor || !critic) return fail(TD3_ERR_INVALID, "td3_agent_bind_params: null argument");
  const td3_param_set* sets[2] = {actor, critic};
  for (auto* s : sets)
    if (!s->params || !s->target || !s->grad || !s->exp_avg || !s->exp_avg_sq || !aligned16(s->params) || !aligned16(s->target) ||
        !aligned16(s->grad) || !aligned16(s->exp_avg) || !aligned16(s->exp_avg_sq))
      return fail(TD3_ERR_INVALID, "td3_agent_bind_params: null or misaligned (16 B) buffer");
  a->actor = *actor;
  a->critic = *critic;
  a->params_bound = true;
  a->batch = 0;
  drop_graphs(a);
  return TD3_OK;
}

int td3_agent_bind_state(td3_agent* a, void* state_dev, int64_t n_bytes) {
  if (!a || !state_dev) return fail(TD3_ERR_INVALID, "td3_agent_bind_state: null argument");
  const int64_t need = 16 * 8 + 2 * (int64_t)a->cfg.n_agents * 4;
  if (n_bytes < need) return fail(TD3_ERR_INVALID, "state buffer too small: %lld < %lld bytes", (long long)n_bytes, (long long)need);
  a->state_u64 = reinterpret_cast<unsigned long long*>(state_dev);
  a->state_f32 = reinterpret_cast<float*>(a->state_u64 + 16);
  a->bar_reset = true;
  a->batch = 0;
  a->last_rb_size = -1;
  drop_graphs(a);
  return TD3_OK;
}

int td3_agent_bind_host_status(td3_agent* a, void* host_words) {
  if (!a) return fail(TD3_ERR_INVALID, "td3_agent_bind_host_status: null agent");
  a->host_status = reinterpret_cast<unsigned long long*>(host_words);
  a->batch = 0;                 // the head launch carries the pointer: plan again
  drop_graphs(a);
  return TD3_OK;
}

int td3_agent_host_status_live(const td3_agent* a) { return a && a->host_status_live ? 1 : 0; }
int td3_agent_chain_active(const td3_agent* a) { return a && a->chain_on ? 1 : 0; }

int td3_dp_bind_peers(td3_agent* a, int32_t world, int32_t rank, float* const* critic_grad_peers, float* const* actor_grad_peers,
                      uint32_t* const* flag_peers) {
  if (!a || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || !critic_grad_peers || !actor_grad_peers || !flag_peers)
    return fail(TD3_ERR_INVALID, "td3_dp_bind_peers: bad arguments (world %d, at most %d ranks of one NVLink domain)", world, kMaxPeers);
  for (int i = 0; i < world; ++i) {
    if (!critic_grad_peers[i] || !actor_grad_peers[i] || !flag_peers[i] || !aligned16(critic_grad_peers[i]) || !aligned16(actor_grad_peers[i]))
      return fail(TD3_ERR_INVALID, "td3_dp_bind_peers: null or misaligned peer buffer of rank %d", i);
    a->dp_critic_grads[i] = critic_grad_peers[i];
    a->dp_actor_grads[i] = actor_grad_peers[i];
    a->dp_flags[i] = flag_peers[i];
  }
  a->dp_world = world; a->dp_rank = rank;
  a->batch = 0;                  // the Adam launches carry the peer pointers: plan again
  drop_graphs(a);
  return TD3_OK;
}

int td3_dp_set_fused_reduce(td3_agent* a, int32_t on) {
  if (!a) return fail(TD3_ERR_INVALID, "td3_dp_set_fused_reduce: null agent");
  const bool want = on != 0;
  if (want && a->dp_world < 1) return fail(TD3_ERR_STATE, "td3_dp_set_fused_reduce: td3_dp_bind_peers not called");
  if (want == a->dp_fused) return TD3_OK;
  a->dp_fused = want;
  drop_graphs(a);
  if (a->batch > 0 && a->ws.base) {
    CUDA_TRY(cudaDeviceSynchronize());
    return plan_agent(a, a->batch);
  }
  return TD3_OK;
}

int dp_allreduce_grads(float* out, const float* const* peer_grads, int32_t world, int64_t n, void* stream) {
  if (!out || !peer_grads || world < 1 || world > kMaxPeers || n <= 0) return fail(TD3_ERR_INVALID, "dp_allreduce_grads: bad arguments");
  PeerSumParams P;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < world; ++i) {
    if (!peer_grads[i] || !aligned16(peer_grads[i])) return fail(TD3_ERR_INVALID, "dp_allreduce_grads: null or misaligned buffer of rank %d", i);
    P.peer[i] = peer_grads[i];
  }
  P.out = out; P.n = n; P.world = world;
  peer_sum_kernel<<<(unsigned)((n + kEwPerBlock - 1) / kEwPerBlock), kEwThreads, 0, (cudaStream_t)stream>>>(P);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CUDA_TRY(cudaGetLastError());
  return TD3_OK;
}

int td3_agent_prepare(td3_agent* a, const td3_replay_view* rb, void* stream) {
  int rc = check_ready(a);
  if (rc == TD3_OK) rc = check_rb(a, rb);
  if (rc == TD3_OK) rc = sync_rb_size(a, rb, (cudaStream_t)stream);
  if (rc == TD3_OK) rc = ensure_shadows(a, (cudaStream_t)stream);
  return rc;
}

int td3_agent_params_changed(td3_agent* a) {
  if (!a) return fail(TD3_ERR_INVALID, "td3_agent_params_changed: null agent");
  a->shadow_dirty = true;
  return TD3_OK;
}

int64_t td3_agent_workspace_floats(const td3_agent* agent, int64_t batch) {
  if (!agent || batch <= 0) return -1;
  td3_agent tmp;
  tmp.cfg = agent->cfg;
  tmp.ws.base = nullptr;
  if (plan_agent(&tmp, batch) != TD3_OK) return -1;
  return tmp.ws_floats;
}

int td3_agent_plan(td3_agent* a, int64_t batch, float* workspace, int64_t workspace_floats, void* stream) {
  int rc = check_ready(a, false);
  if (rc != TD3_OK) return rc;
  if (batch <= 0 || !workspace || !aligned16(workspace)) return fail(TD3_ERR_INVALID, "td3_agent_plan: bad batch/workspace");
  const int64_t need = td3_agent_workspace_floats(a, batch);
  if (need < 0) return TD3_ERR_INVALID;
  if (workspace_floats < need) return fail(TD3_ERR_INVALID, "workspace too small: %lld < %lld floats", (long long)workspace_floats, (long long)need);
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  a->ws.base = workspace;
  rc = plan_agent(a, batch);
  if (rc != TD3_OK) a->batch = 0;
  else if (a->head_seq) CUDA_TRY(cudaMemset(a->head_seq, 0, sizeof(unsigned int) * a->cfg.n_agents));
  return rc;
}

int td3_agent_region(const td3_agent* a, const char* name, int64_t* offset, int64_t* n_floats) {
  if (!a || !name) return fail(TD3_ERR_INVALID, "td3_agent_region: null argument");
  auto it = a->ws.regions.find(name);
  if (it == a->ws.regions.end()) return fail(TD3_ERR_INVALID, "unknown workspace region '%s'", name);
  if (offset) *offset = it->second.first;
  if (n_floats) *n_floats = it->second.second;
  return TD3_OK;
}

int td3_agent_set_global_batch(td3_agent* a, int64_t global_batch, int64_t batch_offset) {
  if (!a || global_batch < 0 || batch_offset < 0) return fail(TD3_ERR_INVALID, "td3_agent_set_global_batch: bad argument");
  a->global_batch = global_batch;
  a->batch_offset = batch_offset;
  a->plan_rows = nullptr;        // the sampling launch carries the offset: rebuild it
  if (a->batch > 0 && a->ws.base) return plan_agent(a, a->batch);
  return TD3_OK;
}

int td3_sample_batch(td3_agent* a, const td3_replay_view* rb, int32_t rng_mode, void* stream) {
  int rc = check_ready(a);
  if (rc == TD3_OK) rc = check_rb(a, rb);
  if (rc == TD3_OK) rc = plan_sample(a, rb, rng_mode);
  if (rc == TD3_OK) rc = sync_rb_size(a, rb, (cudaStream_t)stream);
  if (rc == TD3_OK) rc = ensure_shadows(a, (cudaStream_t)stream);
  if (rc == TD3_OK) rc = run_seq(a->seq_sample, (cudaStream_t)stream);
  return rc;
}

int td3_target_step(td3_agent* a, void* stream) {
  int rc = check_ready(a);
  if (rc == TD3_OK) rc = ensure_shadows(a, (cudaStream_t)stream);
  return rc == TD3_OK ? run_seq(a->seq_target, (cudaStream_t)stream) : rc;
}

int td3_critic_step(td3_agent* a, int32_t apply, void* stream) {
  int rc = check_ready(a);
  if (rc == TD3_OK) rc = ensure_shadows(a, (cudaStream_t)stream);
  if (rc == TD3_OK) rc = run_seq(a->seq_critic_fb, (cudaStream_t)stream);
  if (rc == TD3_OK && apply) rc = run_seq(a->seq_critic_apply, (cudaStream_t)stream);
  return rc;
}

int td3_critic_apply(td3_agent* a, void* stream) {
  int rc = check_ready(a);
  return rc == TD3_OK ? run_seq(a->seq_critic_apply, (cudaStream_t)stream) : rc;
}

int td3_actor_step(td3_agent* a, int32_t apply, void* stream) {
  int rc = check_ready(a);
  if (rc == TD3_OK) rc = ensure_shadows(a, (cudaStream_t)stream);
  if (rc == TD3_OK) rc = run_seq(a->seq_actor_fb, (cudaStream_t)stream);
  if (rc == TD3_OK && apply) rc = run_seq(a->seq_actor_apply, (cudaStream_t)stream);
  return rc;
}

int td3_actor_apply(td3_agent* a, void* stream) {
  int rc = check_ready(a);
  return rc == TD3_OK ? run_seq(a->seq_actor_apply, (cudaStream_t)stream) : rc;
}

int td3_train_n(td3_agent* a, const td3_replay_view* rb, int64_t total_it, int32_t iterations, int32_t rng_mode, int32_t use_graph,
                void* stream) {
  int rc = check_ready(a);
  if (rc == TD3_OK) rc = check_rb(a, rb);
  if (rc != TD3_OK) return rc;
  if (iterations < 1) return fail(TD3_ERR_INVALID, "iterations must be >= 1");
  if (rng_mode == TD3_RNG_INJECTED && iterations != 1)
    return fail(TD3_ERR_INVALID, "injected indices/noise cover exactly one update: iterations must be 1");
  cudaStream_t s = (cudaStream_t)stream;
  {   // the persistent kernel walks un-clustered tile layouts, the stage launches clustered ones: re-plan on a switch
    const int want = use_graph == 2 ? 0 : 1;
    if (want != a->cluster_mode) {
      a->cluster_mode = want;
      CUDA_TRY(cudaStreamSynchronize(s));
      rc = plan_agent(a, a->batch);
      if (rc != TD3_OK) return rc;
    }
  }
  rc = plan_sample(a, rb, rng_mode);
  if (rc == TD3_OK) rc = sync_rb_size(a, rb, s);
  if (rc == TD3_OK) rc = ensure_shadows(a, s);
  if (rc != TD3_OK) return rc;
  if (use_graph == 2) {            // persistent kernel: all `iterations` updates in cooperative launches
    rc = build_programs(a, s);
    for (int done = 0; rc == TD3_OK && done < iterations;) {
      const int n = std::min(iterations - done, 8192);
      rc = launch_persistent(a, total_it + done, n, s);
      done += n;
    }
    return rc;
  }
  if (use_graph) {
    auto& g = a->graphs;
    if (g.rows != rb->rows || g.row_stride != rb->row_stride || g.rng_mode != rng_mode || !g.critic_only) {
      drop_graphs(a);
      rc = capture(a, false, &g.critic_only, &g.nodes_critic_only);
      if (rc == TD3_OK) rc = capture(a, true, &g.with_actor, &g.nodes_with_actor);
      if (rc != TD3_OK) { drop_graphs(a); return rc; }
      g.rows = rb->rows; g.row_stride = rb->row_stride; g.rng_mode = rng_mode;
    }
  }
  for (int it = 0; it < iterations; ++it) {
    const int64_t step_no = total_it + it + 1;                      // self.total_it += 1 (TD3_featured.py:124)
    const bool policy_step = (step_no % a->cfg.policy_freq) == 0;   // :156
    if (use_graph) {
      CUDA_TRY(cudaGraphLaunch(policy_step ? a->graphs.with_actor : a->graphs.critic_only, s));
      g_launches.fetch_add(policy_step ? a->graphs.nodes_with_actor : a->graphs.nodes_critic_only, std::memory_order_relaxed);
    } else {
      if (a->chain_on) {
        rc = run_seq(policy_step ? a->seq_chain_policy : a->seq_chain_critic, s);
      } else {
        rc = run_seq(a->seq_sample, s);
        if (rc == TD3_OK) rc = run_seq(policy_step ? a->seq_run_policy : a->seq_run_critic, s);
      }
      if (rc != TD3_OK) return rc;
    }
  }
  return TD3_OK;
}

// ---- diagnostics: marginal cost of every launch of an update, measured in situ ----------
// Captures the sampling launch plus the first k launches of the critic-only (with_actor = 0) or policy update into a
// graph for k = 1..n and times `reps` back-to-back replays of each with CUDA events: us_out[k-1] = microseconds per
// replay of the k-launch prefix, kinds_out[k-1] = Launch::Kind of launch k (0 stage, 1 gather, 3 Adam/apply, 5 head,
// 7 front).  Differences of consecutive entries are what each launch adds to the chain with everything before it
// warm/cold exactly as in the real update.  The prefixes mutate the agent (they are real work): tools only.
int td3_debug_prefix_times(td3_agent* a, const td3_replay_view* rb, int32_t with_actor, int32_t reps, float* us_out,
                           int32_t* kinds_out, int32_t cap, int32_t* n_out) {
  int rc = check_ready(a);
  if (rc == TD3_OK) rc = check_rb(a, rb);
  if (rc == TD3_OK) rc = plan_sample(a, rb, TD3_RNG_PHILOX);
  if (rc != TD3_OK) return rc;
  if (!a->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&a->cap_stream, cudaStreamNonBlocking));
  cudaStream_t s = a->cap_stream;
  rc = sync_rb_size(a, rb, s);
  if (rc == TD3_OK) rc = ensure_shadows(a, s);
  if (rc != TD3_OK) return rc;
  std::vector<Launch> all;
  if (a->chain_on) {
    all = with_actor ? a->seq_chain_policy : a->seq_chain_critic;
  } else {
    all = a->seq_sample;
    for (const Launch& L : (with_actor ? a->seq_run_policy : a->seq_run_critic)) all.push_back(L);
  }
  const int n = (int)std::min<size_t>(all.size(), (size_t)cap);
  const long long launches_before = g_launches.load();
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  fork_ensure();
  for (int k = 1; k <= n; ++k) {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < k && rc == TD3_OK; ++i) rc = run_launch(all[i], s);
    fork_join(s);
    cudaError_t e = cudaStreamEndCapture(s, &graph);
    if (rc != TD3_OK || e != cudaSuccess) return rc != TD3_OK ? rc : fail(TD3_ERR_CUDA, "prefix capture: %s", cudaGetErrorString(e));
    CUDA_TRY(cudaGraphInstantiate(&exec, graph, 0));
    for (int i = 0; i < 20; ++i) CUDA_TRY(cudaGraphLaunch(exec, s));
    CUDA_TRY(cudaEventRecord(e0, s));
    for (int i = 0; i < reps; ++i) CUDA_TRY(cudaGraphLaunch(exec, s));
    CUDA_TRY(cudaEventRecord(e1, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    us_out[k - 1] = ms * 1000.f / (float)reps;
    kinds_out[k - 1] = (int)all[k - 1].kind;
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  g_launches.store(launches_before);      // captured launches are not executed launches
  *n_out = n;
  return TD3_OK;
}

// ---- small-batch inference on caller buffers ------------------------------------------
static int copy_cols(float* dst, int ld, const float* src, int w, int64_t rows, cudaStream_t s) {
  CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)ld * sizeof(float), src, (size_t)w * sizeof(float), (size_t)w * sizeof(float),
                             (size_t)rows, cudaMemcpyDeviceToDevice, s));
  return TD3_OK;
}

int td3_actor_forward(td3_agent* a, int32_t which, int32_t agent_index, const float* state, const float* particles,
                      int64_t batch, float* action_out, void* stream) {
  int rc = check_ready(a);
  if (rc != TD3_OK) return rc;
  const td3_agent_config& c = a->cfg;
  if (!state || !action_out || batch < 1 || batch > a->batch || agent_index < 0 || agent_index >= c.n_agents)
    return fail(TD3_ERR_INVALID, "td3_actor_forward: bad arguments (batch %lld, planned %lld)", (long long)batch, (long long)a->batch);
  const bool enc = c.variant == TD3_VARIANT_PARTICLES;
  if (enc && !particles) return fail(TD3_ERR_INVALID, "td3_actor_forward: particles required");
  cudaStream_t s = (cudaStream_t)stream;
  const int E = enc ? c.actor.enc_out : 0, S = c.state_dim, A = c.action_dim;
  g_tc_mode = 0;                           // B = 1 latency path: exact fp32 tiles
  PassBuf pb = a->pb_a;                    // agent 0's slot of the online-actor activations
  rc = copy_cols(pb.x0 + E, pb.ld0, state, S, batch, s);
  if (rc != TD3_OK) return rc;
  if (enc)
    CUDA_TRY(cudaMemcpyAsync(const_cast<float*>(pb.P), particles, sizeof(float) * batch * c.n_particles * c.particle_dim,
                             cudaMemcpyDeviceToDevice, s));
  const long long w_at = (long long)agent_index * c.actor.n_floats;
  const float* w_raw = (which ? a->actor.target : a->actor.params) + w_at;
  std::vector<Launch> seq;
  if (c.norm == TD3_NORM_WEIGHT) {
    float* eff = (which ? a->eff_at : a->eff_a) + w_at;
    Launch Lw;
    if (!make_wn_launch(c, 0, {{w_raw, eff, c.actor.n_floats, 1, 0}}, Lw)) return fail(TD3_ERR_INVALID, "weight normalisation: unsupported layout");
    seq.push_back(Lw);
    w_raw = eff;
  }
  ParamRef W{w_raw, 0, 0};
  OutSpec o;
  o.out = action_out; o.ld = A; o.epi = EPI_BIAS_TANH; o.aux0 = a->tanh_y; o.ldaux = A;
  o.f0 = enc ? 1.f : c.max_action;
  for (auto& st : build_forward(c, c.actor, W, GroupShape{1, 1}, (int)batch, pb, o)) emit_stage(seq, st);
  return run_seq(seq, s);
}

int td3_infer_wait(const uint32_t* host_flags, int32_t n, uint32_t seq, int64_t timeout_us);

int td3_infer_b1(td3_agent* a, int32_t net, int32_t which, int32_t agent_index, const float* host_in, float* host_out, uint32_t seq,
                 int64_t wait_us, void* stream) {
  if (!a || !a->params_bound) return fail(TD3_ERR_STATE, "td3_infer_b1: td3_agent_bind_params not called");
  const td3_agent_config& c = a->cfg;
  if (!host_in || !host_out || agent_index < 0 || agent_index >= c.n_agents || net < 0 || net > 1)
    return fail(TD3_ERR_INVALID, "td3_infer_b1: bad arguments");
  if (c.variant != TD3_VARIANT_FEATURED || c.norm == TD3_NORM_WEIGHT)
    return fail(TD3_ERR_UNSUPPORTED, "td3_infer_b1: plain-MLP networks only (the particle encoder goes through td3_actor_forward / td3_critic_forward)");
  const td3_net_layout& L = net == 0 ? c.actor : c.q;
  for (int l = 0; l <= L.n_linear; ++l)
    if (L.dims[l] > kInferMaxWidth) return fail(TD3_ERR_UNSUPPORTED, "td3_infer_b1: layer width %d > %d", L.dims[l], kInferMaxWidth);
  InferParams P;
  memset(&P, 0, sizeof(P));
  const td3_param_set& ps = net == 0 ? a->actor : a->critic;
  const int n_nets = net == 0 ? 1 : c.n_q;
  P.x_host = host_in; P.out_host = host_out;
  P.W = (which ? ps.target : ps.params) + (long long)agent_index * L.n_floats * n_nets;
  P.net_stride = L.n_floats; P.n_nets = n_nets; P.n_linear = L.n_linear;
  P.ln = c.norm == TD3_NORM_LAYER ? 1 : 0;
  P.final_tanh = net == 0 ? 1 : 0;
  P.out_scale = net == 0 ? c.max_action : 1.f;
  P.seq = seq;
  for (int l = 0; l <= L.n_linear; ++l) P.dims[l] = L.dims[l];
  if (L.dims[0] <= kInferInline) {       // the row rides in the kernel parameters: the device never reads host memory
    P.x_inline = 1;
    memcpy(P.x, host_in, sizeof(float) * L.dims[0]);
  }
  for (int l = 0; l < L.n_linear; ++l) {
    P.w_off[l] = L.w_off[l]; P.b_off[l] = L.b_off[l]; P.lng_off[l] = L.ln_g_off[l]; P.lnb_off[l] = L.ln_b_off[l];
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_nets * kInferCluster); cfg.blockDim = dim3(kInferThreads); cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kInferCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, infer_b1_kernel, P));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (wait_us > 0) return td3_infer_wait(reinterpret_cast<const uint32_t*>(host_out + (size_t)n_nets * L.dims[L.n_linear]), n_nets, seq, wait_us);
  return TD3_OK;
}

int td3_infer_wait(const uint32_t* host_flags, int32_t n, uint32_t seq, int64_t timeout_us) {
  if (!host_flags || n < 1) return fail(TD3_ERR_INVALID, "td3_infer_wait: bad arguments");
  const volatile uint32_t* f = host_flags;
  long long spins = 0;
  timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (;;) {
    bool done = true;
    for (int i = 0; i < n; ++i) done &= f[i] == seq;
    if (done) return TD3_OK;
    if ((++spins & 1023) == 0) {
      clock_gettime(CLOCK_MONOTONIC, &t1);
      const long long us = (t1.tv_sec - t0.tv_sec) * 1000000LL + (t1.tv_nsec - t0.tv_nsec) / 1000;
      if (us > timeout_us) {
        cudaError_t err = cudaGetLastError();
        return fail(TD3_ERR_CUDA, "td3_infer_wait: result did not arrive within %lld us (%s)", (long long)timeout_us, cudaGetErrorString(err));
      }
    }
  }
}

int td3_critic_forward(td3_agent* a, int32_t which, int32_t agent_index, const float* state, const float* particles,
                       const float* action, int64_t batch, float* q_out, void* stream) {
  int rc = check_ready(a);
  if (rc != TD3_OK) return rc;
  const td3_agent_config& c = a->cfg;
  if (!state || !action || !q_out || batch < 1 || batch > a->batch || agent_index < 0 || agent_index >= c.n_agents)
    return fail(TD3_ERR_INVALID, "td3_critic_forward: bad arguments (batch %lld, planned %lld)", (long long)batch, (long long)a->batch);
  const bool enc = c.variant == TD3_VARIANT_PARTICLES;
  if (enc && !particles) return fail(TD3_ERR_INVALID, "td3_critic_forward: particles required");
  cudaStream_t s = (cudaStream_t)stream;
  const int E = enc ? c.q.enc_out : 0, S = c.state_dim, A = c.action_dim, nq = c.n_q, qw = a->qw;
  g_tc_mode = 0;
  PassBuf pb = a->pb_c;
  const int copies = enc ? nq : 1;
  for (int g = 0; g < copies; ++g) {
    rc = copy_cols(pb.x0 + g * pb.x0_gi + E, pb.ld0, state, S, batch, s);
    if (rc == TD3_OK) rc = copy_cols(pb.x0 + g * pb.x0_gi + E + S, pb.ld0, action, A, batch, s);
    if (rc != TD3_OK) return rc;
  }
  if (enc)
    CUDA_TRY(cudaMemcpyAsync(const_cast<float*>(pb.P), particles, sizeof(float) * batch * c.n_particles * c.particle_dim,
                             cudaMemcpyDeviceToDevice, s));
  const long long w_at = (long long)agent_index * c.q.n_floats * nq;
  const float* w_raw = (which ? a->critic.target : a->critic.params) + w_at;
  std::vector<Launch> seq;
  if (c.norm == TD3_NORM_WEIGHT) {
    float* eff = (which ? a->eff_ct : a->eff_c) + w_at;
    Launch Lw;
    if (!make_wn_launch(c, 0, {{w_raw, eff, c.q.n_floats, nq, 1}}, Lw)) return fail(TD3_ERR_INVALID, "weight normalisation: unsupported layout");
    seq.push_back(Lw);
    w_raw = eff;
  }
  ParamRef W{w_raw, 0, c.q.n_floats};
  OutSpec o;
  o.out = q_out; o.ld = qw; o.gi = batch * qw; o.epi = EPI_BIAS;
  for (auto& st : build_forward(c, c.q, W, GroupShape{1, nq}, (int)batch, pb, o)) emit_stage(seq, st);
  return run_seq(seq, s);
}

}  // extern "C"
