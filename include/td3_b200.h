/*
 * td3_b200.h -- C ABI of libtd3b200.so: the B200-native (sm_100a) TD3 update hot path.
 *
 * The reference (yannikkellerde/TD3) has no FFI: its boundary is a duck-typed Python
 * object surface.  This header is the boundary a maintainer of the
 * reference binds with ctypes (INTEGRATION.md shows the stub); every entry point names
 * the reference code it replaces.  Conventions:
 *   - plain C types only; device memory is BORROWED (the caller -- PyTorch in the
 *     shipped host layer -- owns every allocation, so state_dict() views stay valid);
 *   - every call takes the CUDA stream to launch on (cudaStream_t passed as void*),
 *     never synchronises the host, and returns 0 on success or a negative td3_status;
 *     the message is available from td3_last_error() (thread-local);
 *   - all floating-point data is IEEE fp32, indices are int64.
 */
#ifndef TD3_B200_H_
#define TD3_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TD3_ABI_VERSION 1
#define TD3_MAX_LINEAR 8          /* linear layers per network, incl. the output layer */
#define TD3_MAX_SEGMENTS 16       /* destination segments of one replay row */

typedef enum td3_status {
  TD3_OK = 0,
  TD3_ERR_INVALID = -1,           /* bad argument (ValueError on the Python side) */
  TD3_ERR_CUDA = -2,              /* a CUDA runtime call failed (RuntimeError) */
  TD3_ERR_STATE = -3,             /* call order / plan missing */
  TD3_ERR_UNSUPPORTED = -4
} td3_status;

typedef enum td3_norm {
  TD3_NORM_NONE = 0,
  TD3_NORM_LAYER = 1,             /* post-ReLU nn.LayerNorm on every hidden layer (TD3_featured.py:44-46, TD3_particles.py:60-67) */
  TD3_NORM_WEIGHT = 2             /* torch.nn.utils.weight_norm on every `linears` module (TD3_particles.py:48-50): parameters are
                                     weight_g [out,1] and weight_v [out,in], the layer uses W = g * v / ||v||_row */
} td3_norm;
typedef enum td3_variant { TD3_VARIANT_FEATURED = 0, TD3_VARIANT_PARTICLES = 1 } td3_variant;
typedef enum td3_precision {
  TD3_PRECISION_FP32 = 0,         /* every contraction on strict-fp32 FFMA tiles (matches the CPU reference to ~1e-6) */
  TD3_PRECISION_TF32 = 1          /* contractions with 16-byte-copyable operands and K >= 64 on tcgen05 kind::tf32 tiles,
                                     fp32 accumulation in tensor memory; everything else (and all storage) stays fp32 */
} td3_precision;
typedef enum td3_rng_mode {
  TD3_RNG_PHILOX = 0,             /* on-device Philox4x32-10 indices + noise */
  TD3_RNG_INJECTED = 1            /* indices / N(0,1) draws supplied by the caller (parity mode) */
} td3_rng_mode;

/* ----------------------------------------------------------------------------------- */
/* Packed network layout.  One network = one contiguous fp32 buffer whose tensors sit   */
/* at the offsets below, in the reference's nn.Module.parameters() order, each          */
/* reference-shaped and contiguous (so the host exposes them as state_dict views).      */
/* Replaces: Actor/Q (TD3_featured.py:15-81), Actor/Q_network (TD3_particles.py:19-119) */
/* ----------------------------------------------------------------------------------- */
typedef struct td3_net_layout {
  int32_t n_linear;                        /* len(self.linears) */
  int32_t dims[TD3_MAX_LINEAR + 1];        /* dims[0] = trunk input width, dims[n_linear] = output width */
  int64_t w_off[TD3_MAX_LINEAR];           /* linears.i.weight  [dims[i+1], dims[i]] row-major (weight norm: linears.i.weight_v) */
  int64_t wg_off[TD3_MAX_LINEAR];          /* linears.i.weight_g [dims[i+1]]   (norm == weight only) */
  int64_t b_off[TD3_MAX_LINEAR];           /* linears.i.bias    [dims[i+1]] */
  int64_t ln_g_off[TD3_MAX_LINEAR];        /* lnorms.i.weight   [dims[i+1]]   (norm == layer) */
  int64_t ln_b_off[TD3_MAX_LINEAR];        /* lnorms.i.bias */
  /* particle-set encoder (TD3_particles.py:29-32,53-58); unused for the featured variant */
  int32_t enc_hidden;                      /* 256 = conv1 out channels */
  int32_t enc_out;                         /* 128 = conv2 out channels */
  int64_t c1w_off, c1b_off;                /* conv1.weight [enc_hidden,1,1,D], conv1.bias */
  int64_t c2w_off, c2b_off;                /* conv2.weight [enc_out,enc_hidden,1], conv2.bias */
  int64_t ln_in_g_off, ln_in_b_off;        /* lnorm1.weight/bias [dims[0]] (norm == layer) */
  int64_t n_floats;                        /* padded size of this network in the packed buffer */
} td3_net_layout;

/* The five packed buffers of one network family (online, target, grad, Adam moments). */
typedef struct td3_param_set {
  float* params;      /* online network   (actor or critic = [q1 | q2]) */
  float* target;      /* target network, same layout */
  float* grad;        /* gradient of the last backward, same layout */
  float* exp_avg;     /* Adam first moment  (torch.optim.Adam state "exp_avg") */
  float* exp_avg_sq;  /* Adam second moment (state "exp_avg_sq") */
} td3_param_set;

/* Agent configuration.  Replaces TD3_base.__init__ (TD3_base.py:7-24) and the TD3
 * constructors (TD3_featured.py:100-110, TD3_particles.py:139-151). */
typedef struct td3_agent_config {
  int32_t variant;                 /* td3_variant */
  int32_t norm;                    /* td3_norm */
  int32_t n_q;                     /* 2 = clipped double Q (CDQ), 1 = single critic (TD3_particles.py:131) */
  int32_t state_dim;               /* featured: S.  particles: F (feature vector width) */
  int32_t action_dim;              /* A */
  int32_t n_particles;             /* particles: N (0 for featured) */
  int32_t particle_dim;            /* particles: D */
  int32_t clamp_target_action;     /* featured 1 (TD3_featured.py:135-137), particles 0 (TD3_particles.py:179-181) */
  int32_t n_agents;                /* independent agents stepped in lock-step in one launch (>= 1) */
  int32_t precision;               /* td3_precision */
  float max_action;                /* actor output scale; particles actor ignores it (TD3_particles.py:68-69) -> pass 1 */
  float discount, policy_noise, noise_clip;   /* used as fp32 scalars against fp32 tensors, as torch casts them */
  /* Python-float (double) hyper-parameters: torch derives 1-beta1, 1-beta2, 1-tau and the bias corrections in
   * double before casting to fp32, so they must arrive un-rounded for <= 2 ulp Adam parity. */
  double tau, lr_actor, lr_critic, beta1, beta2, adam_eps;
  int32_t policy_freq;
  int32_t reserved1;
  uint64_t seed;                   /* Philox key */
  td3_net_layout actor;            /* layout of ONE actor */
  td3_net_layout q;                /* layout of ONE Q network; the critic buffer is n_q of them back to back */
} td3_agent_config;

/* Device-resident replay buffer view: array-of-rows, one transition per row.
 * Row = [state | action | next_state | reward | not_done] (featured) or
 *       [particles | next_particles | feat | action | next_feat | reward | not_done] (particles; each particle set
 *        padded to a multiple of 4 floats, so both start 16-byte aligned: the gather stages them with cp.async.bulk),
 * fp32, row_stride floats apart.  Replaces the five/seven float64 NumPy arrays of
 * my_replay_buffer.py:16-22,81-85 (fp32 round-to-nearest at add == FloatTensor(float64) at sample). */
typedef struct td3_replay_view {
  const float* rows;
  int64_t row_stride;              /* floats between consecutive rows (multiple of 4) */
  int64_t row_floats;              /* payload floats per row */
  int64_t max_size;
  int64_t size;                    /* rows currently valid: sampling is uniform over [0, size) */
  int64_t agent_stride;            /* floats between the buffers of consecutive agents (n_agents > 1) */
} td3_replay_view;

typedef struct td3_agent td3_agent;      /* opaque */

/* ---- library ---------------------------------------------------------------------- */
int td3_abi_version(void);
/* sizeof(td3_net_layout, td3_param_set, td3_agent_config, td3_replay_view): binding self-check. */
void td3_struct_sizes(int64_t* out4);
const char* td3_last_error(void);
/* SM count / name of the current device; fails (TD3_ERR_CUDA) when no sm_100 device is present. */
int td3_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len);

/* The stand-alone entry points below (rb_sample_indices, td3_gemm, set_encoder_*) keep small scratch allocations per
 * device (the device current at the call).  Calls on one stream are ordered; two streams of the same device must not
 * run the same stand-alone entry point concurrently.  Agents (td3_agent_*) own all their memory and have no such limit. */

/* ---- replay buffer (my_replay_buffer.py) ------------------------------------------ */
/* add(): copy n packed fp32 rows from (pinned) host memory into rows[ptr .. ptr+n) with
 * ring wrap-around.  Replaces ReplayBuffer_*.add (my_replay_buffer.py:46-56,109-117); the
 * caller advances ptr/size exactly as :55-56. */
int rb_add_rows(float* rows, int64_t row_stride, int64_t row_floats, int64_t max_size, int64_t ptr,
                const float* host_rows, int64_t n_rows, void* stream);
/* sample() given indices: out segment k receives columns [seg_off[k], seg_off[k]+seg_len[k]) of
 * row idx[b] at dst[k] + b*dst_ld[k].  Replaces the fancy-index gather + FloatTensor cast +
 * H2D of my_replay_buffer.py:61-69,122-128; bit-exact for identical indices. */
int rb_sample_indices(const td3_replay_view* rb, const int64_t* idx_dev, int64_t batch, int32_t n_seg,
                      const int64_t* seg_off, const int64_t* seg_len, float* const* dst, const int64_t* dst_ld,
                      void* stream);
/* np.random.randint(0, size, batch) replaced by Philox4x32-10 (my_replay_buffer.py:59,120):
 * idx[b] = mulhi64(philox(seed; stream_id, step, b), size). */
int rb_philox_indices(int64_t* idx_dev, int64_t batch, int64_t size, uint64_t seed, uint64_t stream_id,
                      uint64_t step, void* stream);

/* ---- elementwise optimiser kernels over packed buffers ----------------------------- */
/* torch.optim.Adam.step (TD3_featured.py:153,164) on n floats, t = step number (1-based), and
 * optionally the Polyak update target = tau*p + (1-tau)*target (TD3_featured.py:167-171) fused in.
 * target == NULL -> Adam only.  grad == NULL -> Polyak only. */
int adam_polyak_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, float* target,
                     int64_t n, int64_t t, double lr, double beta1, double beta2, double eps, double tau, void* stream);

/* ---- dense contraction on its own (torch.nn.Linear / addmm / mm call sites: TD3_featured.py:39-48,73-81) ---- */
/* C[M,N] = op(A)[M,K] . op(B)[K,N] (+ bias[N], optional ReLU).  a_rc / b_rc = 1: the operand is stored with the
 * reduction index contiguous (A as [M,K] row-major, B as [N,K] row-major, i.e. an nn.Linear weight); 0: stored with
 * the reduction index as the row (A as [K,M], B as [K,N]).  use_tc = 1 runs the tcgen05 TF32 tile (fails with
 * TD3_ERR_UNSUPPORTED when the operands are not eligible), 0 the fp32 FFMA tile.  Used by the kernel tests and the
 * GEMM roofline measurements. */
int td3_gemm(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, int32_t a_rc, const float* B, int64_t ldb,
             int32_t b_rc, float* C, int64_t ldc, const float* bias, int32_t relu, int32_t use_tc, void* stream);

/* ---- particle-set encoder on caller buffers (TD3_particles.py:29-32, 53-58 / 104-109) ------------------------------
 * The encoder both particle networks share: conv1 = Conv2d(1, enc_hidden, (1, D)) == a per-particle linear layer
 * D -> enc_hidden, ReLU, conv2 = Conv1d(enc_hidden, enc_out, 1) == per-particle linear enc_hidden -> enc_out, ReLU,
 * AvgPool2d((1, N)) over the particles, ReLU.  particles is [batch * n_particles, particle_dim] (row = one particle),
 * the weights are the reference modules' tensors viewed as [enc_hidden, D] and [enc_out, enc_hidden].
 * set_encoder_fwd writes pooled[b, 0:enc_out] (row stride ld_pooled) and, when the pointers are given, the activations
 * h1 [rows, enc_hidden] and h2 [rows, enc_out] a backward pass needs.  use_tc = 1 runs the fused tcgen05 kernel
 * (csrc/enc.cuh: both layers and the pooling partial sums in one persistent launch, TF32 operands, fp32 accumulation; needs
 * enc_hidden = 256, enc_out = 128, D <= 7, n_particles % 128 == 0, else TD3_ERR_UNSUPPORTED); use_tc = 0 the strict-fp32
 * tiles.  set_encoder_bwd turns d(loss)/d(pooled) into the gradients of conv1 / conv2 (what autograd produces for the
 * reference's backward pass through _encode); it reads the h1 / h2 / pooled a forward call stored.
 * workspace: at least set_encoder_workspace_floats(...) floats of device memory, contents undefined afterwards. */
int64_t set_encoder_workspace_floats(int64_t batch, int64_t n_particles, int64_t particle_dim, int64_t enc_hidden, int64_t enc_out);
int set_encoder_fwd(const float* particles, int64_t batch, int64_t n_particles, int64_t particle_dim, int64_t enc_hidden,
                    int64_t enc_out, const float* conv1_w, const float* conv1_b, const float* conv2_w, const float* conv2_b,
                    float* pooled, int64_t ld_pooled, float* h1, float* h2, float* workspace, int64_t workspace_floats,
                    int32_t use_tc, void* stream);
int set_encoder_bwd(const float* particles, int64_t batch, int64_t n_particles, int64_t particle_dim, int64_t enc_hidden,
                    int64_t enc_out, const float* conv2_w, const float* h1, const float* h2, const float* pooled, int64_t ld_pooled,
                    const float* d_pooled, int64_t ld_d_pooled, float* g_conv1_w, float* g_conv1_b, float* g_conv2_w,
                    float* g_conv2_b, float* workspace, int64_t workspace_floats, int32_t use_tc, void* stream);

/* The fused forward / backward pair of the same encoder (csrc/enc.cuh + csrc/encbwd.cuh; enc_hidden = 256, enc_out = 128,
 * D <= 7, n_particles % 128 == 0, TF32 operands, fp32 accumulation).  Instead of h1 / h2 (1.5 KB per particle) the forward
 * keeps relu_bits: 16 * batch * n_particles uint32 words of ReLU sign bits (4 words per particle for h2, the same bits
 * transposed per 128-particle tile, and 8 words per particle for h1, transposed).  set_encoder_bwd_fused recomputes h1
 * from the particles and produces the same four gradients as set_encoder_bwd in two persistent tcgen05 launches plus a
 * fixed-order reduction of per-CTA partials (deterministic).  This is what TD3_particles.train runs in TF32 mode
 * (TD3_particles.py:167-224 under autograd through :53-58). */
int set_encoder_fwd_bits(const float* particles, int64_t batch, int64_t n_particles, int64_t particle_dim, const float* conv1_w,
                         const float* conv1_b, const float* conv2_w, const float* conv2_b, float* pooled, int64_t ld_pooled,
                         uint32_t* relu_bits, float* workspace, int64_t workspace_floats, void* stream);
int set_encoder_bwd_fused(const float* particles, int64_t batch, int64_t n_particles, int64_t particle_dim, const float* conv1_w,
                          const float* conv1_b, const float* conv2_w, const uint32_t* relu_bits, const float* pooled,
                          int64_t ld_pooled, const float* d_pooled, int64_t ld_d_pooled, float* g_conv1_w, float* g_conv1_b,
                          float* g_conv2_w, float* g_conv2_b, float* workspace, int64_t workspace_floats, void* stream);

/* ---- agent ------------------------------------------------------------------------- */
int td3_agent_create(const td3_agent_config* cfg, td3_agent** out);
int td3_agent_destroy(td3_agent* agent);
/* Bind the packed parameter buffers (borrowed). */
int td3_agent_bind_params(td3_agent* agent, const td3_param_set* actor, const td3_param_set* critic);
/* Bind the small persistent device state block (borrowed, zero-initialised by the caller):
 *   u64[0] sampling step (Philox counter)   u64[1] critic Adam step   u64[2] actor Adam step
 *   u64[3] live replay size                 u64[4..15] reserved
 *   then fp32: critic_loss[n_agents], actor_loss[n_agents]  (last update's losses; the reference never
 *   reads them back -- TD3_featured.py:148,159 -- they exist for tests and metrics).
 * n_bytes >= 128 + 8*n_agents.  The Adam steps are what torch.optim.Adam keeps in state["step"]. */
int td3_agent_bind_state(td3_agent* agent, void* state_dev, int64_t n_bytes);
/* Optional host mirror of the critic loss (the reference computes critic_loss at TD3_featured.py:148 and never reads
 * it; a training loop that logs it would call .item(), i.e. drain the stream).  host_words = n_agents 8-byte words in
 * page-locked host memory that the device can address (cudaHostAlloc / torch pin_memory under UVA), zero-initialised.
 * The fused critic-head kernel of every update stores {low 32 bits: loss as fp32, high 32 bits: number of critic
 * updates completed by this plan} to word[agent] as soon as the loss exists; a host thread that polls the sequence
 * half gets the step's result without waiting for the optimiser kernels behind it.  Call before td3_agent_plan (it
 * invalidates the plan).  td3_agent_host_status_live() says whether the planned update writes the words (it does
 * whenever the critic head is fused: output width <= 4, last hidden width <= 512). */
int td3_agent_bind_host_status(td3_agent* agent, void* host_words);
int td3_agent_host_status_live(const td3_agent* agent);
/* 1 when the planned update runs as the layer-fused chain launches (csrc/chain.cuh: TD3_featured.py:129-163 as one
 * launch per phase; opt-in through the environment variable TD3_CHAIN=1, plain MLPs in TF32 mode), 0 when it runs as
 * the default stage-per-layer sequence. */
int td3_agent_chain_active(const td3_agent* agent);
/* precision = TD3_PRECISION_TF32 keeps round-to-nearest TF32 copies of the four packed parameter buffers for the tensor
 * cores (tcgen05 kind::tf32 truncates its operands; a pre-rounded operand is read exactly, which removes the bias of
 * truncation).  The optimiser kernels keep them current.  Whoever writes the parameter buffers from outside -- the
 * host mirror's load_state_dict / load (TD3_base.py:37-50), or any in-place edit of a state_dict view -- calls this
 * afterwards; the next entry point that runs an update rebuilds the copies first (one launch). */
int td3_agent_params_changed(td3_agent* agent);
/* Things an update does lazily on entry -- publishing rb->size to the device word the sampling kernel reads, rebuilding
 * the TF32 copies after td3_agent_params_changed -- done now, on `stream`.  A caller that replays a captured CUDA graph
 * of the phase calls below (td3_b200/data_parallel.py) calls this before every replay. */
int td3_agent_prepare(td3_agent* agent, const td3_replay_view* rb, void* stream);

/* ---- data-parallel update over one NVLink domain (BASELINE config 5b; the reference is single-device) ------------
 * The batch of ONE agent is split over `world` ranks (td3_agent_set_global_batch); what has to be exchanged is the sum
 * of the packed gradients before each Adam step (the step being sharded: TD3_featured.py:151-153 and :162-164).
 *
 * dp_allreduce_grads: out[e] = sum over r of peer_grads[r][e], added in rank order, read through peer mappings (or plain
 * device pointers on one GPU).  The caller orders it after the producers (stream order / its own flags).
 *
 * td3_dp_bind_peers + td3_dp_set_fused_reduce(1): the same sum folded INTO the optimiser kernels -- critic_grad_peers /
 * actor_grad_peers are the `world` ranks' gradient buffers (this rank's own td3_param_set::grad among them, all in
 * symmetric, peer-mapped memory), flag_peers their 64-word flag arrays (zero-initialised).  td3_critic_apply /
 * td3_actor_apply then run [signal "my gradient is complete" + wait for every rank's] -> Adam reading the peers'
 * gradients over NVLink -> [signal "done reading"], and td3_critic_step / td3_actor_step wait for every rank's "done
 * reading" before they overwrite the gradient.  No all-reduce pass, no reduced copy, bit-identical sums on all ranks.
 * Needs the unfused phase calls (td3_sample_batch .. td3_actor_apply); counters live in the bound state block. */
int dp_allreduce_grads(float* out, const float* const* peer_grads, int32_t world, int64_t n, void* stream);
int td3_dp_bind_peers(td3_agent* agent, int32_t world, int32_t rank, float* const* critic_grad_peers,
                      float* const* actor_grad_peers, uint32_t* const* flag_peers);
int td3_dp_set_fused_reduce(td3_agent* agent, int32_t on);
/* Workspace (activations, batch staging) for a given batch size, in floats. */
int64_t td3_agent_workspace_floats(const td3_agent* agent, int64_t batch);
/* Bind the workspace and build the launch plan for `batch`.  Drops captured graphs. */
int td3_agent_plan(td3_agent* agent, int64_t batch, float* workspace, int64_t workspace_floats, void* stream);
/* Offset (floats) and length of a named workspace region ("q", "target_q", "critic_loss",
 * "next_action", "indices", "noise_in", "indices_in", ...): test / metric read-back and injection. */
int td3_agent_region(const td3_agent* agent, const char* name, int64_t* offset, int64_t* n_floats);

/* `iterations` full TD3 updates (TD3_featured.py:123-171 / TD3_particles.py:167-224), each:
 * sample -> target step -> twin-critic fwd/bwd + Adam -> every policy_freq-th call actor step +
 * Polyak.  total_it is the reference's counter BEFORE the first of these updates.  In
 * TD3_RNG_INJECTED mode the caller has filled the regions "indices_in" (int64 [n_agents][batch]) and
 * "noise_in" (N(0,1) draws, fp32 [n_agents][batch][A]) and iterations must be 1.
 * exec_mode: 2 = persistent cooperative kernel (all iterations in one launch, device-wide barriers between the
 * dependent stages), 1 = CUDA-graph replay of one kernel per stage, 0 = plain stage-by-stage launches. */
int td3_train_n(td3_agent* agent, const td3_replay_view* rb, int64_t total_it, int32_t iterations,
                int32_t rng_mode, int32_t exec_mode, void* stream);
/* The phases of one update, launched individually (no graph): parity tests and the DP critic
 * (gradient all-reduce between critic_backward and critic_apply). */
int td3_sample_batch(td3_agent* agent, const td3_replay_view* rb, int32_t rng_mode, void* stream);
int td3_target_step(td3_agent* agent, void* stream);      /* TD3_featured.py:129-142 */
int td3_critic_step(td3_agent* agent, int32_t apply, void* stream);  /* :145-153 (apply=0: stop after backward) */
int td3_critic_apply(td3_agent* agent, void* stream);     /* Adam on the critic (:153) */
int td3_actor_step(td3_agent* agent, int32_t apply, void* stream);   /* :159-164 */
int td3_actor_apply(td3_agent* agent, void* stream);      /* actor Adam + Polyak of both targets (:164-171) */
/* Data-parallel shard of a larger batch: losses and gradients are normalised by 1/global_batch instead of
 * 1/batch, and local row b draws element (b + batch_offset) of the global batch's Philox index / noise sequence, so the
 * global batch does not depend on the number of ranks.  global_batch = 0 restores single-device behaviour. */
int td3_agent_set_global_batch(td3_agent* agent, int64_t global_batch, int64_t batch_offset);

/* B=small inference on caller buffers (device pointers):
 * actor(x)  -> select_action (TD3_featured.py:113-115, TD3_particles.py:153-157)
 * critic(x,u) -> eval_q      (TD3_featured.py:117-121, TD3_particles.py:159-164)
 * which: 0 = online net, 1 = target net; agent_index selects one of n_agents.  particles == NULL for
 * the featured variant.  batch must not exceed the planned batch.  q_out is [n_q][batch][q_width].
 * These reuse the training workspace (nothing in it persists between updates). */
int td3_actor_forward(td3_agent* agent, int32_t which, int32_t agent_index, const float* state, const float* particles,
                      int64_t batch, float* action_out, void* stream);
int td3_critic_forward(td3_agent* agent, int32_t which, int32_t agent_index, const float* state, const float* particles,
                       const float* action, int64_t batch, float* q_out, void* stream);
/* select_action / eval_q at batch 1 (TD3_featured.py:113-121; main.py:44-45,250), plain-MLP networks: ONE kernel per
 * call.  host_in = the input row ([state] for net 0 = actor, [state | action] for net 1 = the twin critics), host_out =
 * [n_nets * out_dim results][n_nets sequence words], both in page-locked host memory the device can address: the kernel
 * reads the row over PCIe, walks the layers with activations in shared memory (strict fp32, LayerNorm as :44-46) and
 * stores results and then `seq` into the sequence words.  td3_infer_wait spins on those words (host side, no CUDA
 * call) until they equal seq.  Ordered after everything enqueued on `stream` before it, like the reference's .cpu(). */
int td3_infer_b1(td3_agent* agent, int32_t net, int32_t which, int32_t agent_index, const float* host_in, float* host_out,
                 uint32_t seq, int64_t wait_us, void* stream);   /* wait_us > 0: also td3_infer_wait(.., wait_us) before returning */
int td3_infer_wait(const uint32_t* host_flags, int32_t n, uint32_t seq, int64_t timeout_us);

/* Number of kernel launches issued by this library since load (bench.py "gpu_launches"). */
int64_t td3_launch_count(void);

/* Diagnostics (tools/prefix_times.py): microseconds per replay of the first k launches of an update, k = 1..n,
 * each prefix captured into its own CUDA graph and timed with CUDA events.  Mutates the agent. */
int td3_debug_prefix_times(td3_agent* a, const td3_replay_view* rb, int32_t with_actor, int32_t reps, float* us_out,
                           int32_t* kinds_out, int32_t cap, int32_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* TD3_B200_H_ */
