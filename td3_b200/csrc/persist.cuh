// persist.cuh -- the persistent update kernel: `iterations` complete TD3 updates in ONE cooperative
// launch.  The update is a program of stage records in global memory (the same stages the
// stage-per-launch path runs); every CTA walks the program, takes the tiles blockIdx.x, blockIdx.x +
// gridDim.x, ... of each stage and meets the others at a device-wide barrier where the next stage
// consumes what this one produced.  Compared with one kernel (or graph node) per stage this removes the
// launch gap between dependent stages (~19 per update) and lets the next stage's descriptor be fetched
// while the barrier is still filling.
#pragma once
#include "misc.cuh"
#include "stage.cuh"
#include "tc.cuh"
#include "front.cuh"
#include "apply.cuh"

namespace td3 {

// dynamic shared memory of both kernels: 1 KB alignment slack + the tensor-core operand ring.  The FFMA tile's
// cp.async ring, epilogue tile and bias strip (kSmemBytes) alias the start of the same region.
constexpr int kDynSmemBytes = 1024 + kTcRingBytes + 512;   // + bias strip of the TC epilogue
constexpr int kDynSmemBytesSmall = 1024 + kTcSlots * kTcSub + 512;   // small ring (StageParams::small_ring): two CTAs per SM
static_assert(kSmemBytes <= kTcSlots * kTcSub, "FFMA tile buffers must fit inside the small TC ring");
static_assert(kSmemBytes <= kTcRingBytes, "FFMA tile buffers must fit inside the TC ring");
static_assert(kFrontWideSmemBytes <= kTcRingBytes && kFrontSmemBytes <= kTcRingBytes, "front tiles must fit inside the TC ring");

__device__ __forceinline__ unsigned char* aligned_smem(unsigned char* raw) {
  const unsigned int a = smem_u32(raw);
  return raw + ((1024u - (a & 1023u)) & 1023u);
}

// one tile of a stage: look up the problem the tile index falls into and run it.  kTc selects which GEMM tile is
// compiled in: the tensor-core kernels carry no FFMA GEMM code and vice versa (instruction footprint).
template <bool kTc, bool kCo = false, bool kLean = false>
__device__ __forceinline__ void run_stage_tile(const StageParams& S, int tile_global, unsigned char* ring, TcState* tc,
                                               bool param_maps = false) {
  int pi = 0;
#pragma unroll
  for (int q = 1; q < kMaxProblemsPerStage; ++q)
    if (q < S.n_problems && tile_global >= S.p[q].tile_begin) pi = q;
  const Problem& P = S.p[pi];
  const int tile = tile_global - P.tile_begin;
  if constexpr (kLean) {                      // every problem of the launch is a tensor-core GEMM (host: StageParams::all_tc)
    gemm_tile_tc<kCo>(P, tile, ring, tc, param_maps ? S.maps : nullptr);
    return;
  }
  float* smem = reinterpret_cast<float*>(ring);
  switch (P.kind) {
    case PK_GEMM:
      if constexpr (kTc) {
        if (P.use_tc) { gemm_tile_tc<kCo>(P, tile, ring, tc, param_maps ? S.maps : nullptr); break; }
      }
      gemm_tile(P, tile, smem);
      break;
    case PK_LN_FWD: ln_fwd_tile(P, tile); break;
    case PK_LN_BWD_ROWS: ln_bwd_rows_tile(P, tile); break;
    case PK_LN_BWD_COLS: ln_bwd_cols_tile(P, tile, smem); break;
    case PK_POOL_FWD: pool_fwd_tile(P, tile, smem); break;
    case PK_POOL_BWD: pool_bwd_tile(P, tile); break;
    case PK_REDUCE_SPLITS: reduce_splits_tile(P, tile); break;
    case PK_NEG_MEAN: neg_mean_tile(P, tile, smem); break;
    case PK_COLSUM: colsum_tile(P, tile, smem); break;
    case PK_SMALLK_FWD: smallk_fwd_tile(P, tile, smem); break;
    case PK_SMALLK_DW: smallk_dw_tile(P, tile, smem); break;
    default: break;
  }
}

// stage-per-launch form (phase-by-phase API, CUDA-graph mode, B=small inference)
// kCo: the many-tile (small-ring) launches' instance, with the coalesced epilogue of tc.cuh
// kLean: the launch holds tensor-core GEMM problems only -- an instance without the other tile kinds' code
template <bool kTc, bool kCo = false, bool kLean = false>
__global__ void __launch_bounds__(kStageThreads, 2) stage_kernel(const __grid_constant__ StageParams S) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ TcState tc;
  unsigned char* ring = aligned_smem(smem_raw);
  pdl_launch_dependents();
  const int cluster = S.cluster;
  if constexpr (kTc) {
    if (S.any_tc) tc_setup(&tc, cluster > 1 ? cluster : 1, S.small_ring ? 1 : kTcGroup);   // TMEM allocation + mbarrier init overlap the previous stage's tail
    if (cluster > 1) cluster_sync_all();     // every peer's barriers exist before anybody multicasts into its shared memory
  }
  pdl_wait();                                // the previous stage's outputs are complete and visible from here on
#ifdef TD3_TILE_PROF
  if (threadIdx.x == 0) tc.prof_stage = 0;   // stage-per-launch form: block 0's stamps land in slot 0
  __syncthreads();
#endif
  if ((int)blockIdx.x < S.total_tiles) run_stage_tile<kTc, kCo, kLean>(S, blockIdx.x, ring, &tc, true);   // S is __grid_constant__
  if constexpr (kTc) {
    if (cluster > 1) cluster_sync_all();     // nobody leaves while a peer may still write its shared memory or barriers
    if (S.any_tc) tc_teardown(&tc);
  }
}

enum StageKind : int { SK_STAGE = 0, SK_GATHER = 1, SK_LOSS = 2, SK_EW_ONLY = 3, SK_HEAD = 4, SK_WN = 5, SK_FRONT = 6,
                       SK_APPLY = 7 };   // SK_APPLY: first-layer dW + optimiser tiles (apply.cuh) followed by element-wise blocks

struct alignas(16) StageRec {
  int kind;
  int main_tiles;                   // tiles of the main part (problem table / gather blocks / loss blocks)
  int ew_tiles;                     // Adam/Polyak blocks appended after the main tiles (independent of them)
  int barrier_after;                // 0: the next stage does not depend on this one
  int gather_grid_x, pad0, pad1, pad2;
  AdamTick tick;                    // tick.state != nullptr: CTA 0 performs the actor optimiser tick on entering the stage
  union Main {
    StageParams st;
    GatherParams g;
    LossParams l;
    HeadParams h;
    WnParams w;
    FrontParams f;
    DwParams d;
  } u;
  EwParams ew;
};

struct PersistArgs {
  const StageRec* prog_critic;      // update without the delayed actor step
  const StageRec* prog_policy;      // update with actor step + Polyak
  int n_critic, n_policy;
  long long total_it;               // the reference's counter before the first update of this launch
  int iterations, policy_freq;
  unsigned int* barrier;            // arrival counter
  unsigned int barrier_base;        // its value when this launch starts (host-tracked)
  long long* prof;                  // optional [2 CTAs][stages][3] clock64 stamps of the last iteration, or nullptr
};

// Device-wide barrier for a co-resident (cooperatively launched) grid: one monotonically increasing
// arrival counter; barrier number n of the launch completes when it reaches base + n * gridDim.x (the
// host passes `base`, the value the counter has when the launch starts).  One release-reduction and
// one acquire-load per CTA: the critical path is a single L2 round trip after the last arrival.
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(ctr) : "memory");
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(ctr) : "memory");
    } while ((int)(v - target) < 0);
  }
  __syncthreads();
}

__device__ __forceinline__ void fetch_rec(StageRec* dst, const StageRec* src) {
  constexpr int kWords = (int)(sizeof(StageRec) / 16);
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (int i = threadIdx.x; i < kWords; i += blockDim.x) d4[i] = __ldcg(s4 + i);
}

template <bool kTc>
__global__ void __launch_bounds__(kStageThreads, 1) persistent_update_kernel(const __grid_constant__ PersistArgs a) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ StageRec rec[2];
  __shared__ TcState tc;
  __shared__ float red[8];
  unsigned char* ring = aligned_smem(smem_raw);
  if constexpr (kTc) tc_setup(&tc);
  int slot = 0;
  unsigned int bar_target = a.barrier_base;
  {
    const bool pol0 = ((a.total_it + 1) % a.policy_freq) == 0;
    fetch_rec(&rec[0], pol0 ? a.prog_policy : a.prog_critic);
  }
  __syncthreads();
  for (int it = 0; it < a.iterations; ++it) {
    const bool policy = ((a.total_it + it + 1) % a.policy_freq) == 0;     // TD3_featured.py:124,156
    const StageRec* prog = policy ? a.prog_policy : a.prog_critic;
    const int n = policy ? a.n_policy : a.n_critic;
    for (int s = 0; s < n; ++s) {
      const StageRec& R = rec[slot];
      const bool prof = a.prof && it == a.iterations - 1 && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
      long long* pr = prof ? a.prof + ((blockIdx.x == 0 ? 0 : 128) + s) * 3 : nullptr;
      if (prof) pr[0] = clock64();
#ifdef TD3_TILE_PROF
      if (threadIdx.x == 0) tc.prof_stage = (a.prof && it == a.iterations - 1) ? s : -1;
      __syncthreads();
#endif
      if (blockIdx.x == 0 && threadIdx.x == 0 && R.tick.state) adam_tick(R.tick);
      const int total = R.main_tiles + R.ew_tiles;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        if (tile >= R.main_tiles) {
          adam_polyak_body(R.ew, tile - R.main_tiles);
        } else if (R.kind == SK_STAGE) {
          run_stage_tile<kTc>(R.u.st, tile, ring, &tc);
        } else if (R.kind == SK_GATHER) {
          gather_body(R.u.g, tile % R.gather_grid_x, tile / R.gather_grid_x);
        } else if (R.kind == SK_LOSS) {
          __syncthreads();
          loss_body(R.u.l, tile, red);
        } else if (R.kind == SK_HEAD) {
          __syncthreads();
          head_body(R.u.h, tile, reinterpret_cast<float*>(ring));
        } else if (R.kind == SK_WN) {
          wn_body(R.u.w, tile);
        } else if (R.kind == SK_FRONT) {
          if (R.u.f.job_groups > 0) front_wide_body(R.u.f, tile, reinterpret_cast<float*>(ring));
          else front_body(R.u.f, tile, reinterpret_cast<float*>(ring));
        } else if (R.kind == SK_APPLY) {
          __syncthreads();
          dw_adam_body(R.u.d, R.ew, tile, reinterpret_cast<float*>(ring));
        }
      }
      // descriptor of the next stage (static data) while the others are still working
      const StageRec* next = nullptr;
      if (s + 1 < n) next = prog + s + 1;
      else if (it + 1 < a.iterations)
        next = (((a.total_it + it + 2) % a.policy_freq) == 0) ? a.prog_policy : a.prog_critic;
      const int barrier_after = R.barrier_after;
      if (next) fetch_rec(&rec[slot ^ 1], next);
      slot ^= 1;
      if (prof) pr[1] = clock64();
      if (barrier_after) {
        bar_target += gridDim.x;
        grid_barrier(a.barrier, bar_target);
      } else {
        __syncthreads();
      }
      if (prof) pr[2] = clock64();
    }
  }
  if constexpr (kTc) tc_teardown(&tc);
}

}  // namespace td3
