/*
 * b200ppo.h — C ABI of the B200-native PPO hot path (sm_100a).
 *
 * Drop-in boundary for the hot path of emiwar/nnx-ppo (reference v0.2.1).  The reference is pure
 * Python/JAX and has no FFI seam of its own; each entry point below replaces the *computation* of
 * the reference function cited next to it, and is what an XLA-FFI / ctypes binding for that
 * function binds (see INTEGRATION.md for the reference-side stubs).
 *
 * Conventions (all entry points):
 *   - plain C types only; every pointer marked "dev" is a device pointer owned by the caller;
 *   - `stream` is a cudaStream_t passed as void*; launchers never allocate, free or synchronise,
 *     so they can be captured into a CUDA graph;
 *   - return 0 on success, a negative B200PPO_E* code on invalid arguments, or a positive
 *     cudaError_t value if a launch failed (b200ppo_error_string decodes all three);
 *   - re-entrant: no global mutable state except one-time cudaFuncSetAttribute opt-ins.
 *
 * Layouts: everything is float32 row-major unless noted; rollout tensors are time-major
 * [T][B][...] with B = number of envs on this rank; masks are uint8 (0/1); indices are int32.
 */
#ifndef B200PPO_H_
#define B200PPO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PPO_MAX_LAYERS 8

#define B200PPO_ACT_NONE 0
#define B200PPO_ACT_RELU 1
#define B200PPO_ACT_SWISH 2
#define B200PPO_ACT_TANH 3

#define B200PPO_EINVAL (-1)   /* bad shape / null pointer / unsupported plan            */
#define B200PPO_ELIMIT (-2)   /* size exceeds a compiled-in limit       */
#define B200PPO_EALIGN (-3)   /* pointer or offset not aligned as documented            */

/* One stack of Dense layers (reference: networks/feedforward.py:13-51, factories.py:14-45):
 * layer l computes z = h W_l + b_l with W_l [dims[l]][dims[l+1]] row-major at w_off[l] and b_l at
 * b_off[l] (float offsets into the parameter arena, multiples of 4); hidden layers apply `act`,
 * the last layer is linear. */
typedef struct b200ppo_chain {
  int32_t n_layers;
  int32_t act;
  int32_t dims[B200PPO_MAX_LAYERS + 1];
  int32_t pad_;
  int64_t w_off[B200PPO_MAX_LAYERS];
  int64_t b_off[B200PPO_MAX_LAYERS];
} b200ppo_chain;

/* The network "plan": what make_mlp_actor_critic builds (reference: networks/factories.py:72-146):
 * Normalizer? -> PPOAdapter(action = actor MLP + NormalTanhSampler, value = critic MLP). */
typedef struct b200ppo_plan {
  int32_t obs_dim;
  int32_t act_dim;
  int32_t normalize;        /* 1 if a Normalizer precedes the adapter                      */
  int32_t pad_;
  float entropy_weight;     /* NormalTanhSampler args, sampling_layers.py:69-80            */
  float min_std;
  float std_scale;
  float pad2_;
  int64_t n_params;         /* arena length in floats (including alignment padding)       */
  b200ppo_chain actor;      /* out = 2 * act_dim                                           */
  b200ppo_chain critic;     /* out = 1                                                     */
} b200ppo_plan;

/* Hyper-parameters of one minibatch update (reference: PPOConfig, algorithms/config.py:11-30). */
typedef struct b200ppo_hparams {
  float gamma;              /* discounting_factor */
  float lambda_;            /* gae_lambda         */
  float clip_range;
  float critic_loss_weight;
  float learning_rate;
  float adam_b1, adam_b2, adam_eps;
  float weight_decay;       /* < 0: plain adam; >= 0: adamw with this decay               */
  float grad_clip;          /* <= 0: no clip_by_global_norm                               */
  int32_t normalize_advantages;
  int32_t world_size;       /* data-parallel ranks sharing this update (means are global) */
  int32_t rank;             /* this process' rank (only read when bufs->comm is set)      */
} b200ppo_hparams;

/* Device copy of the numeric hyper-parameters (b200ppo_update_bufs.hparams_dev): float[B200PPO_HP_FLOATS].
 * The reference keeps gae_lambda / discounting_factor as TRACED scalars of its jitted step
 * (ppo.py:105) and the learning rate inside the optax state, i.e. they are run-time data of the
 * device program; a captured CUDA graph follows a schedule the same way when the kernels read them
 * from device memory instead of from their (captured) launch arguments. */
#define B200PPO_HP_GAMMA 0
#define B200PPO_HP_LAMBDA 1
#define B200PPO_HP_CLIP_RANGE 2
#define B200PPO_HP_CRITIC_WEIGHT 3
#define B200PPO_HP_LEARNING_RATE 4
#define B200PPO_HP_ADAM_B1 5
#define B200PPO_HP_ADAM_B2 6
#define B200PPO_HP_ADAM_EPS 7
#define B200PPO_HP_WEIGHT_DECAY 8
#define B200PPO_HP_GRAD_CLIP 9
#define B200PPO_HP_FLOATS 16

#define B200PPO_METRICS_STRIDE 12

/* Device buffers of one minibatch update.  ws is a scratch arena of
 * b200ppo_update_workspace_bytes() bytes, 256-byte aligned. */
typedef struct b200ppo_update_bufs {
  /* rollout buffer, time-major over the B envs of this rank */
  const float* obs;            /* dev [T][B][O]  raw observations (Normalizer rollout_extras)    */
  const float* raw_action;     /* dev [T][B][A]  sampler rollout_extras                           */
  const float* loglik_old;     /* dev [T][B]                                                       */
  const float* reward;         /* dev [T][B]                                                       */
  const uint8_t* done;         /* dev [T][B]                                                       */
  const uint8_t* truncated;    /* dev [T][B]                                                       */
  const float* next_obs_last;  /* dev [B][O]     next_obs[-1] (ppo.py:433)                         */
  const int32_t* inds;         /* dev [mb]       env indices of this minibatch (ppo.py:297)        */
  /* normalizer (read-only during updates) */
  const float* norm_mean;      /* dev [O]                                                          */
  const float* norm_std;       /* dev [O]  from b200ppo_norm_prepare                               */
  /* parameters and optimizer state (updated in place) */
  float* params;               /* dev [P]                                                          */
  float* adam_mu;              /* dev [P]                                                          */
  float* adam_nu;              /* dev [P]                                                          */
  /* counters: [0..1] sampler stream key, [2] sampler count base, [3] adam count base */
  const uint32_t* rng_state;   /* dev uint32[4]                                                    */
  float* metrics_out;          /* dev [B200PPO_METRICS_STRIDE]: [0] actor loss, [1] critic loss,
                                * [2] regularisation loss, [3] grad norm (with gradient clipping),
                                * [4] clipping fraction, [5] E[target], [6] E[target^2], [7] E[adv],
                                * [8] E[adv^2] of the advantages the surrogate uses (normalised when
                                * normalize_advantages: ppo.py:477-480 reassigns `advantages` before it is
                                * logged at :523) - all divided by the global sample count; [9], [10] the
                                * normalisation constants mean(a), std(a) + 1e-8 (0, 1 when off) divided by
                                * world_size: a SUM over ranks returns every entry as its global value */
  void* ws;                    /* dev scratch                                                      */
  /* optional peer-memory exchange (NULL: the caller all-reduces adv_sums and the gradient between
   * the stages).  dev uint64[world_size]: base address of every rank's comm buffer
   * (b200ppo_comm_alloc / b200ppo_comm_ipc_open), entry [rank] being this rank's own. */
  const uint64_t* comm;
  /* optional dev uint8[n_params]: 0 marks a structural zero (the off-diagonal blocks when per-key
   * encoders — containers.py Concat — are laid out as one block-diagonal Dense layer); such entries
   * are excluded from the gradient norm and never updated; 2 marks the second copy of a tied
   * parameter (param_tie below): updated, not counted in the norm.  NULL: every entry is a parameter. */
  const uint8_t* param_mask;
  /* optional dev float[B200PPO_HP_FLOATS]: when non-NULL the kernels read the numeric hyper-parameters
   * (gamma, lambda, clip range, critic weight, learning rate, Adam b1 / b2 / eps, weight-decay and
   * clip-threshold VALUES) from here instead of from `hp`; `hp` still decides the launch structure
   * (weight decay on / off, clipping on / off, normalize_advantages, world size). */
  const float* hparams_dev;
  /* optional dev uint32[1]: monotonic base of the peer-exchange epoch (epoch = base + update_index +
   * 1), advanced by b200ppo_iter_finalize.  NULL: the Adam count rng_state[3] is the base (it goes
   * backwards when a training state is rolled back, which would make stale flags look current). */
  const uint32_t* comm_epoch;
  /* optional dev int32[n_params]: tied parameters (a trunk shared by the actor and the critic chain is stored
   * once per chain).  param_tie[i] = index of the other copy of parameter i, or -1.  The RED stage adds the
   * two copies' gradients (d loss / d w = actor path + critic path, containers.py Sequential feeding one
   * feature tensor to both PPOAdapter ports), so both copies take the same optimizer step and stay
   * bit-identical; mark the second copy 2 in param_mask to keep it out of the gradient norm.  NULL: no ties. */
  const int32_t* param_tie;
} b200ppo_update_bufs;

/* -------- library -------------------------------------------------------------------------- */
int b200ppo_version(void);
const char* b200ppo_error_string(int code);
int b200ppo_num_sms(void);

/* -------- K7: threefry device functions, exposed for parity tests -------------------------- *
 * jax.random.bits / normal / randint over a flat shape (n,) from a key held by value.          */
int b200ppo_random_bits(void* stream, uint32_t k0, uint32_t k1, int64_t n, uint32_t* out /*dev*/);
int b200ppo_random_normal(void* stream, uint32_t k0, uint32_t k1, int64_t n, float* out /*dev*/);

/* -------- K6: minibatch permutation indices (ppo.py:287-294) ------------------------------- *
 * out[e][:] = jax.random.permutation(fold_in(new_key, e), n) for e < n_epochs, bit-exact.      *
 * new_key: dev uint32[2].  scratch: dev, b200ppo_permutation_scratch_bytes(n, n_epochs) bytes. */
int64_t b200ppo_permutation_scratch_bytes(int32_t n, int32_t n_epochs);
int b200ppo_permutation(void* stream, const uint32_t* new_key /*dev*/, int32_t n, int32_t n_epochs,
                        int32_t* out /*dev [n_epochs][n]*/, void* scratch /*dev*/);

/* -------- K2: generalised advantage estimation (ppo.py:351-394) ---------------------------- */
int b200ppo_gae(void* stream, const float* rewards /*dev [T][B]*/,
                const float* values_excl_last /*dev [T][B]*/, const float* last_value /*dev [B]*/,
                const uint8_t* done /*dev [T][B]*/, const uint8_t* truncation /*dev [T][B]*/,
                int32_t T, int32_t B, float lambda_, float gamma, float* advantages /*dev [T][B]*/);

/* -------- K5: Normalizer (normalizer.py:63-136) -------------------------------------------- *
 * norm_prepare : std[o] = counter > 0 ? sqrt(max(M2/counter, 1e-6)) : 10                        *
 * norm_batch_stats: batch mean / M2 of x[n_rows][O] -> batch_stats[2*O] (two-pass, fp32)        *
 * norm_merge   : Chan-merge `world` batch statistics (rank order) into (mean, M2, counter)      */
int b200ppo_norm_prepare(void* stream, const float* M2 /*dev [O]*/, const float* counter /*dev [1]*/,
                         int32_t O, float* std_out /*dev [O]*/);
int64_t b200ppo_norm_scratch_bytes(int32_t O);
int b200ppo_norm_batch_stats(void* stream, const float* x /*dev [n_rows][O]*/, int64_t n_rows,
                             int32_t O, float* batch_stats /*dev [2*O]*/, void* scratch /*dev*/);
int b200ppo_norm_merge(void* stream, const float* batch_stats /*dev [world][2*O]*/,
                       int32_t world, float n_per_rank, int32_t O, float* mean /*dev [O]*/,
                       float* M2 /*dev [O]*/, float* counter /*dev [1]*/);

/* -------- K1: policy step (rollout.py:18; sampling_layers.py:82-113; adapter.py:75-117) ----- *
 * One network call on B rows: normalize -> actor -> NormalTanhSampler -> critic.                *
 * mode 0: sample (rollout_extras=None); 1: replay raw_action_in; |2: deterministic (mean).      *
 * count_offset is added to rng_state[2]; the call consumes 2 counts (1 if deterministic).       *
 * ws (nullable, b200ppo_policy_workspace_bytes() bytes): receives [B][2A] = [mu | sigma], the     *
 * sampler's `metrics` dict (sampling_layers.py:111); everything else lives in shared memory.      */
int64_t b200ppo_policy_workspace_bytes(const b200ppo_plan* plan, int32_t B);
int b200ppo_policy_step(void* stream, const b200ppo_plan* plan, const float* params /*dev*/,
                        const float* norm_mean /*dev*/, const float* norm_std /*dev*/,
                        const float* obs /*dev [B][O]*/, int32_t B, int32_t mode,
                        const uint32_t* rng_state /*dev*/, uint32_t count_offset,
                        const float* raw_action_in /*dev [B][A] or NULL*/,
                        float* raw_action /*dev [B][A]*/, float* action /*dev [B][A]*/,
                        float* loglik /*dev [B]*/, float* value /*dev [B]*/,
                        float* reg_loss /*dev [B] or NULL*/, void* ws /*dev [B][2A] or NULL*/);

/* -------- K1 (persistent): fused T-step rollout on the synthetic env (rollout.py:11-73) ---- *
 * env state (in/out): obs [B][O], step_counter int32 [B], term_state uint32 [B].                *
 * iter_keys: dev uint32[4] = reset_key, new_key (ppo.py:271).  Advances nothing on its own;     *
 * consumes sampler counts rng_state[2] + 2t, +2t+1 for t < T.                                   */
typedef struct b200ppo_synth_env {
  int32_t obs_dim, act_dim, max_len, term_thresh16;
  const float* Wo;             /* dev [O][O] */
  const float* Wa;             /* dev [A][O] */
} b200ppo_synth_env;

int b200ppo_synth_reset(void* stream, const b200ppo_synth_env* env, const uint32_t* keys /*dev [B][2]*/,
                        int32_t B, float* obs, int32_t* step_counter, uint32_t* term_state);
int b200ppo_synth_init_keys(void* stream, uint32_t k0, uint32_t k1, int32_t B, uint32_t* keys_out);
/* jax.random.split(key, n)[first : first + count] with the key read from DEVICE memory (dev uint32[2]), so *
 * that a captured graph follows the per-iteration keys (rollout.py:57-59: split(reset_key, (T, B))).      */
int b200ppo_split_keys_dev(void* stream, const uint32_t* key /*dev*/, int64_t first, int32_t count,
                           uint32_t* keys_out /*dev [count][2]*/);
/* keys_out[i] = element `index` of jax.random.split(keys[i]) for each of `rows` row keys (dev uint32[rows][2]):
 * the key flow of a vmapped `env.reset(key)` that splits its key first (wrappers, rollout.py:39). */
int b200ppo_split_rows(void* stream, const uint32_t* keys, int32_t rows, uint32_t index, uint32_t* keys_out);
int b200ppo_rollout_synth(void* stream, const b200ppo_plan* plan, const b200ppo_synth_env* env,
                          const float* params /*dev*/, const float* norm_mean, const float* norm_std,
                          const uint32_t* rng_state /*dev*/, const uint32_t* iter_keys /*dev*/,
                          int32_t T, int32_t B,
                          float* env_obs /*dev [B][O] in/out*/, int32_t* env_counter /*dev [B]*/,
                          uint32_t* env_term /*dev [B]*/,
                          float* obs /*dev [T][B][O]*/, float* raw_action /*dev [T][B][A]*/,
                          float* action /*dev [T][B][A]*/, float* loglik /*dev [T][B]*/,
                          float* reward /*dev [T][B]*/, uint8_t* done /*dev [T][B]*/,
                          uint8_t* truncated /*dev [T][B]*/, float* next_obs_last /*dev [B][O]*/);

/* -------- K1 (persistent, evaluation): eval_rollout (rollout.py:97-148) on the synthetic env --- *
 * L policy + env steps per env from the given (freshly reset) env state, which is only read: no  *
 * transition record, no reset; done is sticky, episode_reward sums the rewards of the steps      *
 * taken while the env was not yet done, lifespan counts the steps that did not end in done.      *
 * mode: 0 = sample (sampler counts rng_state[2] + 2t), 2 = deterministic (networks.eval(),       *
 * ppo.py:122; no sample draw).  The caller advances the sampler count by L (mode 2) or 2L.        */
int b200ppo_eval_synth(void* stream, const b200ppo_plan* plan, const b200ppo_synth_env* env,
                       const float* params /*dev*/, const float* norm_mean, const float* norm_std,
                       const uint32_t* rng_state /*dev*/, int32_t mode, int32_t L, int32_t B,
                       const float* env_obs /*dev [B][O]*/, const int32_t* env_counter /*dev [B]*/,
                       const uint32_t* env_term /*dev [B]*/, float* episode_reward /*dev [B]*/,
                       float* lifespan /*dev [B]*/);

/* -------- K3 + K4: one minibatch update (ppo.py:296-317 update_step, 397-531 ppo_loss) ------ *
 * Stages (bit mask, launched in this order):                                                    *
 *   FWD   critic+actor forward over the gathered minibatch (+ bootstrap rows)                   *
 *   GAE   per-env reverse scan from the fresh values, advantage moment sums -> adv_sums          *
 *   LOSS  clipped surrogate / value / entropy terms and their gradients w.r.t. network outputs  *
 *   BWD   backward through both MLPs (dX chain, then dW split over row ranges)                  *
 *   RED   reduce dW partials to the flat gradient (b200ppo_update_grad_ptr) + its squared norm  *
 *   ADAM  optax adam / adamw (+ clip_by_global_norm) from the flat gradient                     *
 * A data-parallel caller all-reduces adv_sums after GAE and the flat gradient after RED.        *
 * update_index selects the sampler count offset (2*(T+1) per update after the 2*T of the        *
 * rollout) and the adam step (rng_state[3] + update_index + 1).                                 */
#define B200PPO_STAGE_FWD 1
#define B200PPO_STAGE_GAE 2
#define B200PPO_STAGE_LOSS 4
#define B200PPO_STAGE_BWD 8
#define B200PPO_STAGE_RED 16
#define B200PPO_STAGE_ADAM 32
#define B200PPO_STAGE_ALL 63
/* finer selection inside BWD (either bit alone runs only that kernel; BWD = both) */
#define B200PPO_STAGE_BWD_DX 64
#define B200PPO_STAGE_BWD_DW 128
/* with FWD: skip the weight pre-split launch of the tensor-core path.  Only valid when the previous
 * call on this workspace was an update whose ADAM stage ran (it refreshes the split operand planes
 * together with the parameters) and nothing else has touched `params` since. */
#define B200PPO_STAGE_NO_PREP 256
/* modifier of STAGE_LOSS: the distillation head (distillation.py:160-232) instead of the PPO surrogate - mean
 * negative log-likelihood of bufs->raw_action (the teacher's mean, raw space) under the current policy + the
 * entropy regulariser; no GAE stage needed, d loss / d value = 0; metrics [0] = NLL, [1] = 0, [2] = regulariser */
#define B200PPO_STAGE_NLL 512

int64_t b200ppo_update_workspace_bytes(const b200ppo_plan* plan, int32_t T, int32_t mb);
int b200ppo_update(void* stream, const b200ppo_plan* plan, const b200ppo_hparams* hp,
                   const b200ppo_update_bufs* bufs, int32_t T, int32_t B, int32_t mb,
                   uint32_t rng_count_offset, int32_t update_index, int32_t stages);
/* -------- peer-memory exchange of the data-parallel update (replaces the two per-update NCCL
 * all-reduces of DESIGN.md section 5: advantage moment sums and the flat gradient) ---------------
 * Every rank owns one comm buffer (cudaMalloc'ed, exported with CUDA IPC) with, double-buffered on the
 * parity of the exchange epoch, one slot per source rank for the advantage sums and for the gradient.
 * Everything is PUSHED: a sender stores each 32-bit word together with the epoch as ONE 8-byte word
 * {payload, epoch} (single-copy atomic; the idea of NCCL's LL protocol) straight into every rank's buffer
 * over NVLink; a receiver polls the words of its OWN buffer until the epoch half matches.  No flags, no
 * system fences, no remote loads, one one-way NVLink latency per exchange:
 *   GAE kernel  -> its last block pushes (sum a, sum a^2) into every rank's slot;
 *   loss kernel -> collects all ranks' sums in rank order;
 *   Adam kernel -> every thread pushes its (locally reduced) gradient element and sums the W slots of its
 *                  own buffer in rank order (bit-identical parameters on every rank); with
 *                  clip_by_global_norm the same launch also produces the squared norm of the summed
 *                  gradient and the clipped optimizer step is a second launch.
 * epoch = (bufs->comm_epoch ? *comm_epoch : adam count) + update_index + 1 (monotonic, >= 1).            */
#define B200PPO_MAX_RANKS 16
int64_t b200ppo_comm_bytes(const b200ppo_plan* plan, int32_t world_size);
int b200ppo_comm_alloc(int64_t bytes, void** out);                 /* zero-initialised device memory */
int b200ppo_comm_free(void* p);
int b200ppo_comm_ipc_get(void* p, uint8_t* handle64);              /* 64-byte cudaIpcMemHandle_t     */
int b200ppo_comm_ipc_open(const uint8_t* handle64, void** out);
int b200ppo_comm_ipc_close(void* p);

/* -------- recurrent actor: one time step of obs -> Normalizer -> Dense(act) -> LSTM -> Dense ------
 * Replaces, per time step, `networks(state, obs)` for an actor built as
 * Sequential([Dense, LSTM, Dense]) (networks/recurrent.py:89-161 over flax's OptimizedLSTMCell;
 * the architecture of recurrent_test.py:245-258) in the rollout (rollout.py:18) and in the replay
 * scan of ppo_loss (ppo.py:409-431), and its reverse-mode step.  Gate order (i, f, g, o); the LSTM
 * kernels are stored as one matrix Wcat = [Wi; Wh] ([pre_dim + hidden, 4*hidden], row-major) with
 * the hidden kernels' bias [4*hidden].  After a forward step the carry of rows whose `done` flag is
 * set is zeroed (reset_state); the backward step applies the same mask to the incoming carry
 * gradient, so nothing flows across an episode boundary.  The host loops over time. */
typedef struct b200ppo_lstm_plan {
  int32_t obs_dim, pre_dim, hidden, out_dim;   /* out_dim = 2 * action_size                       */
  int32_t act;                                  /* activation of the pre Dense (B200PPO_ACT_*)     */
  int32_t normalize;
  int64_t w1_off, b1_off;                       /* pre Dense  [obs_dim, pre_dim], [pre_dim]        */
  int64_t wcat_off, bl_off;                     /* LSTM       [pre_dim + hidden, 4 hidden], [4 hidden] */
  int64_t w2_off, b2_off;                       /* post Dense [hidden, out_dim], [out_dim]         */
  int64_t n_params;
  /* trainable_initial_state (recurrent.py:85-87, 135-141, 154-157): arena offsets of the learned carry a
   * reset hands out, [hidden] each - init_c_off for carry slot 0 (the reference's `initial_h` Param, which
   * flax's cell reads as c), init_h_off for slot 1; both 0 (a zero-initialised struct): reset to zeros.
   * Sequence API only (the b200ppo_lstm_step_* kernels reject a plan that has them). */
  int64_t init_c_off, init_h_off;
} b200ppo_lstm_plan;
/* floats of the per-step activation cache the backward step needs, for `rows` rows */
int64_t b200ppo_lstm_cache_floats(const b200ppo_lstm_plan* plan, int32_t rows);
/* obs: dev [.][obs_dim]; row j reads obs[inds ? inds[j] : j]; done (nullable): dev uint8, indexed
 * like obs; c, h: dev [rows][hidden] carry, updated in place; y: dev [rows][out_dim];
 * cache (nullable): dev, b200ppo_lstm_cache_floats() floats. */
int b200ppo_lstm_step_fwd(void* stream, const b200ppo_lstm_plan* plan, const float* params,
                          const float* norm_mean, const float* norm_std, const float* obs,
                          const int32_t* inds, const uint8_t* done, int32_t rows, float* c, float* h,
                          float* y, float* cache);
/* d_y: dev [rows][out_dim]; dc, dh: dev [rows][hidden], on entry the gradient w.r.t. the carry this
 * step handed on (zeros for the last step), on exit w.r.t. the carry it received.  Weight gradients,
 * two ways: (a) all four *_out NULL: accumulated into grad (dev [n_params], zero it first) with
 * atomics; (b) deterministic: the step writes its GEMM operands - cat_out [rows][pre_dim + hidden]
 * = [u | h_in], hn_out [rows][hidden] = h', da_out [rows][4 hidden], dz_out [rows][pre_dim] - and
 * b200ppo_lstm_weight_grads() reduces the operands of ALL steps (buffers laid out step-major) with
 * three batched GEMMs and fixed-order partial sums (grad may be NULL then). */
int b200ppo_lstm_step_bwd(void* stream, const b200ppo_lstm_plan* plan, const float* params,
                          const float* d_y, const float* cache, const int32_t* inds,
                          const uint8_t* done, int32_t rows, float* dc, float* dh, float* grad,
                          float* cat_out, float* hn_out, float* da_out, float* dz_out);
int64_t b200ppo_lstm_wgrad_scratch_floats(const b200ppo_lstm_plan* plan, int32_t rows_total);
/* cache / cat / hn / da / dz / d_y: the step-major [steps * rows][width] operand matrices; writes (not
 * accumulates) the recurrent actor's weight and bias gradients into grad (dev [n_params]). */
int b200ppo_lstm_weight_grads(void* stream, const b200ppo_lstm_plan* plan, const float* cache,
                              const float* cat, const float* hn, const float* da, const float* dz,
                              const float* d_y, int32_t rows_total, float* grad, float* scratch);

/* -------- recurrent actor, whole sequences on the tensor cores (tcgen05 3xTF32) -----------------------
 * The replay scan of ppo_loss over T steps (ppo.py:411-431) and its reverse-mode pass, for the same
 * Dense(act) -> LSTM -> Dense actor.  Only h_{t-1} Wh is sequential: the input projection, the post
 * Dense and every weight gradient are batched GEMMs over all T * rows samples; the T step launches
 * compute 128-row x 16-unit tiles with the gate math, the reset-on-done select and the activation cache
 * in the GEMM epilogue; when all tiles fit on the device at once the forward recurrence is ONE persistent
 * launch (weights resident in shared memory, the carry handed between CTAs through arrival counters).  Needs hidden % 16 == 0 and pre_dim % 4 == 0 (b200ppo_lstm_seq_supported;
 * otherwise use the step kernels above).  ws: b200ppo_lstm_seq_workspace_floats() floats, 256-byte aligned, shared by a
 * forward call (keep_cache = 1) and the backward call that follows it.
 *   x     dev [T*rows][obs_dim] inputs in step-major order; normalised on load when norm_mean / norm_std
 *         are given (rollout: raw observations), else taken as is (replay: the gathered + normalised
 *         observations the critic pass produced)
 *   done  dev uint8 [T][B] (nullable): row j of step t resets its carry after the step when
 *         done[t*B + (inds ? inds[j] : j)] is set
 *   c, h  dev [rows][hidden]: in = carry entering step 0, out = carry after step T-1 (reset applied)
 *   y     dev [T*rows][out_dim] actor outputs
 *   keep_cache  bit 0: keep the activation cache in ws for the backward call; bit 1: the split weight planes in
 *         ws are current (same params as the previous call on this ws: the rollout's T calls) - skip their launch
 * backward: d_y dev [T*rows][out_dim]; grad dev [n_params]: the recurrent actor's weight and bias
 * gradients are WRITTEN (fixed-order sums: bit-reproducible), other entries untouched.               */
int b200ppo_lstm_seq_supported(const b200ppo_lstm_plan* plan);
/* 1 (also B200PPO_LSTM_PERSIST=1): the forward recurrence of a replay is one persistent launch when its tiles
 * fit on the device at once; 0 (default: measured no faster inside the captured iteration): one launch per
 * step.  Returns the previous setting (-1: unset). */
int b200ppo_lstm_set_persistent(int on);
int64_t b200ppo_lstm_seq_workspace_floats(const b200ppo_lstm_plan* plan, int32_t T, int32_t rows);
int b200ppo_lstm_seq_num_launches(const b200ppo_lstm_plan* plan, int32_t T, int32_t rows, int32_t backward);
int b200ppo_lstm_seq_forward(void* stream, const b200ppo_lstm_plan* plan, const float* params,
                             const float* norm_mean, const float* norm_std, const float* x,
                             const uint8_t* done, const int32_t* inds, int32_t B, float* c, float* h,
                             int32_t T, int32_t rows, float* ws, float* y, int32_t keep_cache);
int b200ppo_lstm_seq_backward(void* stream, const b200ppo_lstm_plan* plan, const float* params,
                              const float* x, const float* d_y, const uint8_t* done, const int32_t* inds,
                              int32_t B, int32_t T, int32_t rows, float* ws, float* grad);
/* parity hooks of the generic tile kernels: C[M][N] = A[M][K] B (b_nt = 0: B [K][N]; 1: B [N][K]), with
 * k_slices > 1 the partial products land in C[slice][M][N]; W[Kdim][N] = A^T D, b[N] = column sums of D
 * (scratch: S * (Kdim + 1) * N floats). */
int b200ppo_rg_gemm_test(void* stream, const float* A, const float* B, float* C, int32_t M, int32_t N,
                         int32_t K, int32_t b_nt, int32_t k_slices);
int b200ppo_rg_tn_test(void* stream, const float* A, const float* D, int32_t rows, int32_t Kdim, int32_t N,
                       int32_t S, float* scratch, float* W, float* b);

/* NormalTanhSampler (sampling_layers.py:88-147) on actor outputs y [B][2A] = [mu | rho] computed
 * elsewhere (the recurrent step): mode bit 0 = replay stored raw actions, bit 1 = deterministic;
 * keys as in b200ppo_policy_step (sample: count, entropy: the next count; deterministic skips the
 * sample draw); element (row, d) draws index row*A + d.  reg_loss (nullable): -entropy_weight * H. */
int b200ppo_sampler_step(void* stream, const float* y, int32_t B, int32_t A, int32_t mode,
                         float min_std, float std_scale, float entropy_weight,
                         const uint32_t* rng_state, uint32_t count_offset,
                         const float* raw_action_in, float* raw_action, float* action,
                         float* loglik, float* reg_loss);

/* GEMM engine of the update: 0 = fp32 FFMA on CUDA cores, 1 = tcgen05 3xTF32 (default; error-   *
 * compensated, fp32-level accuracy), 2 = tcgen05 plain TF32 (not fp32 parity).  Also selectable  *
 * with the environment variable B200PPO_GEMM=ffma|tf32x3|tf32.  Returns the previous mode.       */
int b200ppo_set_gemm_mode(int mode);
/* Kernel-structure switches of the update (A/B measurements, cross-checks; environment: B200PPO_FUSE_GAE_LOSS,  *
 * B200PPO_DW_MN): fuse_gae_loss_on = 1: GAE and loss of a call that runs both stages are ONE launch (default),     *
 * 0: two; dw_mn_on = 1: the weight-gradient kernel takes its operands as MN-major swizzled TMA boxes wherever the  *
 * layer's widths are multiples of 4 (default), 0: it transposes them in shared memory.  -1 keeps a setting.      *
 * Returns the previous settings (bit 0 | bit 1).  The row-split table of a plan depends on dw_mn_on.              */
int b200ppo_set_update_paths(int fuse_gae_loss_on, int dw_mn_on);
/* Dense engine of the rollout (b200ppo_rollout_synth / _ws): 1 = tensor cores (default: the fused kernel on warp-level *
 * mma.sync m16n8k8 tiles, error-compensated 3xTF32, while the weights fit an SM's shared memory; the batched per-step  *
 * tcgen05 GEMM path of b200ppo_rollout_synth_ws beyond that), 0 = fused kernel on fp32 FFMA tiles, 2 = batched path     *
 * whenever a workspace is passed; also B200PPO_ROLLOUT=ffma|mma|wide.  Mode 1 picks between the two tensor-core        *
 * kernels by the number of 16-env tiles (more tiles than SMs: one CTA per SM with two env groups sharing the weights;  *
 * else one tile per CTA); 3 / 4 = mode 1 with the one-tile / two-group kernel forced (tests, A/B; B200PPO_ROLLOUT_MMA=1|2). *
 * Returns the previous mode (0..2).                                                                                    */
int b200ppo_set_rollout_mode(int mode);
/* b200ppo_rollout_synth with a caller-owned scratch buffer (b200ppo_rollout_synth_workspace_bytes(plan, B) bytes; 0 =  *
 * this plan keeps its weights resident in the fused kernel and needs none).  With a workspace, networks / envs whose    *
 * weights do not fit shared memory (e.g. 768-wide dict-observation encoders) run every Dense layer of a step as one      *
 * tcgen05 tile GEMM over all B envs instead of re-streaming the weights per 16-env tile; same results contract.          */
/* The synthetic env's side of ONE rollout step for a policy evaluated elsewhere (the recurrent actor's            *
 * b200ppo_lstm_seq_forward): `_begin` once per rollout (splits the env matrix into operand planes, copies the env   *
 * observation into the step's input tile and into obs[0]); `_step` for t = 0 .. T-1: NormalTanhSampler on the actor *
 * outputs y [B][ldy] (sample count rng_state[2] + 2 t, rollout.py:18 / sampling_layers.py:88-113), env step, reward, *
 * done / truncation, reset select (rollout.py:19-45), record row t and obs[t + 1]; env state advanced in place.       */
int64_t b200ppo_synth_env_step_workspace_bytes(int32_t O, int32_t A, int32_t B);
int b200ppo_synth_env_begin(void* stream, const b200ppo_synth_env* env, int32_t B, const float* env_obs /*dev [B][O]*/,
                            float* obs0 /*dev [B][O]*/, void* ws, int64_t ws_bytes);
int b200ppo_synth_env_step(void* stream, const b200ppo_synth_env* env, const float* y, int32_t ldy, float min_std,
                           float std_scale, const uint32_t* rng_state, const uint32_t* iter_keys, int32_t t, int32_t T,
                           int32_t B, float* env_obs, int32_t* env_counter, uint32_t* env_term, float* obs,
                           float* raw_action, float* action, float* loglik, float* reward, uint8_t* done,
                           uint8_t* truncated, float* next_obs_last, void* ws, int64_t ws_bytes);
int64_t b200ppo_rollout_synth_workspace_bytes(const b200ppo_plan* plan, int32_t B);
int b200ppo_rollout_synth_num_launches(const b200ppo_plan* plan, int32_t T, int32_t B, int32_t with_ws);
int b200ppo_rollout_synth_ws(void* stream, const b200ppo_plan* plan, const b200ppo_synth_env* env,
                             const float* params, const float* norm_mean, const float* norm_std,
                             const uint32_t* rng_state, const uint32_t* iter_keys, int32_t T, int32_t B,
                             float* env_obs, int32_t* env_counter, uint32_t* env_term, float* obs,
                             float* raw_action, float* action, float* loglik, float* reward, uint8_t* done,
                             uint8_t* truncated, float* next_obs_last, void* ws, int64_t ws_bytes);
/* Programmatic dependent launch between the kernels of b200ppo_update (a kernel's prologue overlaps its
 * predecessor's tail; griddepcontrol.wait before the first dependent access).  0 = plain stream order
 * (default), 1 = every launch, 2 = only the small latency-bound kernels (GAE, loss, Adam), 3 = those and the
 * kernel that follows one; also B200PPO_PDL=0..3 (measurements: profiles/r2_notes.md).  Returns the previous
 * setting; any other value only queries.  Takes effect for launches (and graph captures) made afterwards. */
int b200ppo_set_pdl(int on);
/* Profiling aid (synchronous): clock64 phase stamps of CTA 0 of the last tensor-core update       *
 * kernel -> out_host; returns -(1000 + count).                                                   */
int b200ppo_debug_timestamps(long long* out_host, int32_t max_n);
/* Profiling aid (synchronous): globaltimer (ns) of every CTA of the last tensor-core update kernel, *
 * three per CTA (entry, TMEM / barrier set-up done, exit) -> out_host[3 * max_ctas]; returns the   *
 * number of CTA slots copied (<= 1024).                                                           */
int b200ppo_debug_cta_times(unsigned long long* out_host, int32_t max_ctas);
int b200ppo_debug_select(int flags);     /* bit 0: the dW kernel keeps dX's stamps; bits 8..: blockIdx.x of the stamped CTA */
/* Host-only: the weight-gradient kernel's work table for (plan, T, mb): M-tile item i (actor layers first, one    *
 * item per 128 input columns of a layer) runs as out_splits[i] CTAs of out_rows_per_split[i] rows each (the sum of   *
 * the splits never exceeds the SM count: one CTA per SM).  Returns the number of items written (<= max_items).     */
int b200ppo_update_dw_splits(const b200ppo_plan* plan, int32_t T, int32_t mb, int32_t* out_splits,
                             int32_t* out_rows_per_split, int32_t max_items);
/* Number of kernels b200ppo_update launches for the given stage mask (for launch accounting).   */
int b200ppo_update_num_launches(const b200ppo_plan* plan, const b200ppo_hparams* hp, int32_t T,
                                int32_t mb, int32_t stages);
/* Pointers into the workspace that a data-parallel caller reduces across ranks. */
double* b200ppo_update_adv_sums_ptr(const b200ppo_plan* plan, int32_t T, int32_t mb, void* ws);
float* b200ppo_update_grad_ptr(const b200ppo_plan* plan, int32_t T, int32_t mb, void* ws);
/* Debug / parity views into the workspace (after the corresponding stage has run). */
float* b200ppo_update_debug_ptr(const b200ppo_plan* plan, int32_t T, int32_t mb, void* ws,
                                int32_t which /*0 adv, 1 values, 2 actor out, 3 d_y, 4 d_v, 5 gathered + normalised obs [(T+1)*mb][O]*/);

/* End-of-iteration bookkeeping: rng_state[2] += rng_advance; rng_state[3] += adam_advance;
 * comm_epoch (nullable, dev uint32[1]) += adam_advance. */
int b200ppo_iter_finalize(void* stream, uint32_t* rng_state /*dev*/, uint32_t rng_advance,
                          uint32_t adam_advance, uint32_t* comm_epoch /*dev or NULL*/);

/* -------- tensor-core bring-up / parity hook: C[M][N] = A[M][K] * B[K][N] with tcgen05.mma ---- *
 * kind::tf32, fp32 accumulation in TMEM; split = 0: plain TF32 operands, 1: error-compensated   *
 * 3xTF32 (hi/lo operand split).  N multiple of 16 in [16, 256].                                 */
int b200ppo_tc_gemm_test(void* stream, const float* A /*dev*/, const float* B /*dev*/, float* C /*dev*/,
                         int32_t M, int32_t N, int32_t K, int32_t split);

/* Bring-up / parity harness of the MN-major tensor-core operand path (the dW kernel's): C[K][N] = H^T D  *
 * for row-major H [rows][K], D [rows][N] (device pointers), operands fetched by swizzled TMA boxes, no    *
 * software transposition.  K <= 128, N <= 256, K % 4 == N % 4 == 0; split != 0: error-compensated 3xTF32. */
int b200ppo_tc_mn_test(void* stream, const float* H /*dev*/, const float* D /*dev*/, float* C /*dev*/, int32_t rows,
                       int32_t K, int32_t N, int32_t split, float* dbg /*dev, nullable: (4 + ceil(N/32)) * 512 floats, the
                       operand tiles of the last 16-row stage as they landed in shared memory*/);
/* Microbenchmark (profiling aid): cycles for `iters` dependent tcgen05.mma (M=128, N, K=8), for  *
 * `iters` serialized bulk copies of `bytes`, and for the MMAs over `nacc` accumulators.           */
int b200ppo_tc_microbench(void* stream, const float* src /*dev*/, long long* out /*dev [4]*/, int32_t N,
                          int32_t iters, int32_t bytes, int32_t nacc, int32_t blocks);

/* -------- measurement helper: register-resident FFMA loop (fp32 CUDA-core peak) ------------ */
int b200ppo_ffma_peak(void* stream, int32_t iters, float* sink /*dev [blocks*threads]*/,
                      int32_t blocks, int32_t threads);

#ifdef __cplusplus
}
#endif
#endif /* B200PPO_H_ */
