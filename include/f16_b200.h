/* f16_b200.h -- C ABI of the B200-native batched F-16 longitudinal simulator.
 *
 * Drop-in boundary for the per-step NumPy path of lalapopa/F16-model
 * (F16model/model + F16model/env).  The reference has no FFI of its own (it is
 * pure Python); each entry point below names the reference function it
 * replaces (paths relative to the reference root).  All pointers are DEVICE
 * pointers unless the name ends in _host; the caller (PyTorch, or any other
 * allocator) owns every buffer, the library only enqueues work on the given
 * CUDA stream and keeps no reference past the call.  No torch types cross the
 * boundary.  Functions return 0 on success or a negative F16_E* code and never
 * throw; f16_last_error() returns a thread-local message.
 *
 * Batch layout: structure-of-arrays.  A "state" is real[9][n] in the field
 * order of States.to_array() (model/_interface.py:18-31):
 *     0 Ox [m], 1 Oy [m], 2 wz [rad/s], 3 theta [rad], 4 V [m/s],
 *     5 alpha [rad], 6 stab [rad], 7 dstab [rad/s], 8 Pa [%]
 * `real` is float (F16_F32) or double (F16_F64) as selected by `dtype`.
 */
#ifndef F16_B200_H
#define F16_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F16_ABI_VERSION 2

enum { F16_F32 = 0, F16_F64 = 1 };

enum {
    F16_OK = 0,
    F16_EINVAL = -1,  /* bad argument (null pointer, bad dtype, n <= 0, ...) */
    F16_ECUDA = -2,   /* CUDA runtime error, see f16_last_error() */
    F16_ENODEV = -3   /* no CUDA device / wrong architecture */
};

/* How an episode's initial state is produced by f16_env_reset and by the
 * in-kernel autoreset (reference: F16._get_init_state, env/env.py:172-195). */
enum {
    F16_INIT_RANDOM = 0,  /* get_random_state(), env/env.py:221-243, counter RNG */
    F16_INIT_FIXED = 1,   /* config["init_state"]/["init_control"]: cfg.x0 for every env */
    F16_INIT_PER_ENV = 2  /* explicit per-env states in f16_env_buffers.init_state */
};

enum { F16_OBS_AOS = 0 /* obs[n][5], the reference's layout */, F16_OBS_SOA = 1 /* obs[5][n] */ };

/* f16_config.flags: which kernel / route a call takes.  All variants compute the same numbers bit for bit
 * (tested); the flags exist so that a caller -- and the parity tests -- can pin one per call. */
enum {
    F16_FLAG_ONE_ENV_PER_THREAD = 1, /* fp32 autoreset step: run the one-env-per-thread kernel on every tile instead of
                                        the two-envs-per-thread kernel (the default on full 256/512-env tiles) */
    F16_FLAG_NO_PDL = 2,             /* launch without programmatic dependent launch */
    F16_FLAG_HOST_STAGED = 4         /* f16_env_step_host: never take the zero-copy route */
};

/* Scalar configuration of a batch of envs (reference: the config dict of
 * F16.__init__, env/env.py:18-35, plus what SyncVectorEnv adds). */
typedef struct f16_config {
    int32_t dtype;        /* F16_F32 | F16_F64 */
    int32_t fast_math;    /* F32 only: 1 = MUFU intrinsics for sin/cos/log2/exp2/rcp (within the 1e-4 bound) */
    int64_t n_envs;
    double dt;            /* config["dt"] */
    int32_t horizon;      /* episode_length value at which `done` fires: tn/dt - 1 (env/env.py:115) */
    int32_t ref_len;      /* samples per reference signal: len(arange(0, tn, dt)) (env/task.py:12) */
    int32_t n_ref;        /* number of signals in the bank */
    int32_t autoreset;    /* 1 = gymnasium SyncVectorEnv semantics: reset in the same step, return the reset obs */
    int32_t init_mode;    /* F16_INIT_* */
    int32_t obs_layout;   /* F16_OBS_* */
    int32_t noise;        /* config["noise"] truthy: wz_obs += radians(N(0, 0.5)) (env/env.py:139-152) */
    int32_t flags;        /* F16_FLAG_* bits, 0 = defaults: per-call kernel / route selection (was `reserved`) */
    uint64_t seed;        /* key of the counter-based RNG (reset + noise) */
    int64_t env_id_base;  /* global id of env 0 of this shard: RNG streams are keyed by the GLOBAL env id, so
                             results do not depend on how the batch is sharded over GPUs */
    double x0[9];         /* F16_INIT_FIXED: the initial state */
    int32_t seg_len;      /* 0: ref_bank holds whole signals, real[n_ref][ref_len].  > 0: SEGMENT mode -- every signal is a
                             concatenation of up to four segments of seg_len samples, then zeros (the random `combo` scenario,
                             env/task.py:65-103); ref_bank is the segment bank real[seg_len][seg_stride] (sample-major: row r of
                             segment sample t at [t * seg_stride + r]), an env's ref_idx packs four 8-bit rows of it (segment s in
                             bits [8s, 8s+8)), and a reset copies ref_codes[uniform draw < n_ref].  Needs ref_len <= 4 * seg_len. */
    int32_t seg_stride;   /* segment mode: rows per sample of the segment bank (>= number of rows, <= 256) */
} f16_config;

/* Persistent per-env buffers (device).  What the reference keeps in the F16
 * object: model.state_prev, episode_length, total_return, done, ref_signal. */
typedef struct f16_env_buffers {
    void *state;             /* real[9][n]   r/w */
    int32_t *ep_len;         /* [n]          r/w  F16.episode_length */
    void *total_return;      /* real[n]      r/w  F16.total_return */
    int32_t *ref_idx;        /* [n]          r/w  row of ref_bank followed by each env */
    uint8_t *done;           /* [n]          r/w  F16.done (sticky until reset; always 0 after an autoreset) */
    uint32_t *episode_no;    /* [n]          r/w  episodes started so far (RNG counter) */
    const void *ref_bank;    /* real[n_ref][ref_len]  wz_ref = theta_ref/2 (env/task.py:25-26); segment mode: real[seg_len][seg_stride] */
    const void *init_state;  /* real[9][n]   F16_INIT_PER_ENV only, else NULL */
    void *pos_comp;          /* real[2][n]   r/w  Kahan carry of the Ox, Oy position integrals, or NULL: keeps the
                                fp32 mode's positions within ~1e-7 of the fp64 reference over an episode (plain fp32
                                accumulation drifts up to 1 ulp per step). Leave NULL in fp64 parity mode. */
    const uint64_t *seed_dev; /* [1] optional DEVICE copy of the RNG key, or NULL.  When set, resets and observation noise read
                                the key from here instead of f16_config.seed (kernel arguments are frozen into a captured CUDA
                                graph; a device scalar is not), so re-seeding (F16.reset(seed), env/env.py:155-158) also reaches
                                autoresets replayed from a graph. */
    const uint32_t *ref_codes; /* [n_ref] segment mode only (else NULL): the packed segment rows of every signal of the scenario */
} f16_env_buffers;

/* Per-call tensors of f16_env_step. */
typedef struct f16_step_io {
    const void *actions;  /* real[K][n]  normalised stick in [-1,1] (clipped like env/env.py:204) */
    void *obs;            /* real[n][5] (or [5][n]): observation after the LAST of the K steps; with
                             obs_every_step: real[K][n][5] (or [K][5][n]) */
    void *reward;         /* real[K][n] */
    uint8_t *done;        /* [K][n]  1 on the step that terminated an episode (and, without autoreset, after) */
    void *final_obs;      /* real[n][5]/[5][n]: terminal observation of the most recent episode that ended
                             inside this call (gymnasium "final_observation"); only written where done fired;
                             may be NULL */
    int32_t *final_ep_len;   /* [n] episode_length of that episode ("final_info"); may be NULL */
    void *final_return;      /* real[n] total_return of that episode; may be NULL */
    int32_t obs_every_step;  /* 0: only the last observation is written */
    int32_t reserved;
    double *stats;           /* double[4] accumulators (+=, never cleared by the library): episodes finished, sum of their
                                total_return, sum of their episode_length, spare.  One atomic set per group of lanes that
                                terminate together (active-mask election + shuffle reduction).  May be NULL. */
} f16_step_io;

/* ---- library ---------------------------------------------------------- */
int f16_abi_version(void);
const char *f16_last_error(void);
/* Upload the table images for `device` (idempotent; done lazily otherwise). */
int f16_init(int device);
/* SMs of `device`, CTA size and the resident-CTA count the step kernel is launched with. */
int f16_launch_geometry(int device, int dtype, int *n_sm, int *cta_threads, int *ctas_per_sm);

/* Pipe-rate probe: one launch of a persistent grid that does nothing but `iters` x 32 independent operations per thread
 * of one kind -- the MEASURED denominator of the rooflines MEASURED_PEAKS.json has no entry for (bench.py times it with
 * CUDA events beside the kernel it reports on).  *ops_per_launch receives the operations one launch performs (an FMA
 * counts once).  scratch: >= 4 bytes of device memory. */
enum { F16_PEAK_FP32_FMA = 0, F16_PEAK_FP64_FMA = 1, F16_PEAK_MUFU_TANH = 2 };
int f16_peak_probe(int kind, int iters, double *ops_per_launch, void *scratch, void *stream);

/* ---- env level -------------------------------------------------------- */
/* F16.reset (env/env.py:154-170) for every env whose mask byte is non-zero
 * (mask == NULL: all).  Draws/copies the initial state, zeroes the counters,
 * picks a reference signal (uniformly from the bank unless n_ref == 1) and
 * writes the reset observation.  flags: F16_RESET_KEEP_REF keeps each env's
 * reference row; F16_RESET_KEEP_STATE keeps the state the caller has just
 * written (explicit initial states) and only clears the counters and computes
 * the observation. */
enum { F16_RESET_KEEP_REF = 1, F16_RESET_KEEP_STATE = 2 };
int f16_env_reset(const f16_config *cfg, const f16_env_buffers *env, const uint8_t *mask, int flags,
                  void *obs /* real[n][5] or [5][n], may be NULL */, void *stream);

/* K consecutive F16.step calls (env/env.py:46-70) for all n envs in ONE kernel
 * launch; state stays in registers across the K steps. */
int f16_env_step(const f16_config *cfg, const f16_env_buffers *env, const f16_step_io *io, int K, void *stream);

/* Same step through HOST buffers.  The env's persistent buffers stay on the device; the call blocks until
 * the results are in the host buffers.  Two routes, both on the GPU:
 *   F16_HOST_ROUTE_ZERO_COPY  every buffer is pinned AND mapped (cudaHostAlloc / cudaHostRegister), fp32 gym
 *       step (K = 1, autoreset): the step kernel itself reads the actions from and bulk-stores obs / reward /
 *       done to host memory -- one launch, transfers overlapped with the arithmetic tile by tile;
 *   F16_HOST_ROUTE_STAGED     otherwise: H2D of actions, kernel, D2H of obs/reward/done, chunked and
 *       double-buffered over internal streams.
 * Both routes run on the library's own streams: work the caller has queued on other streams against the env's
 * buffers (a reset, a device-side step) must have completed before the call.
 * f16_host_route() names the route the calling thread's last f16_env_step_host took. */
enum { F16_HOST_ROUTE_STAGED = 0, F16_HOST_ROUTE_ZERO_COPY = 1 };
int f16_env_step_host(const f16_config *cfg, const f16_env_buffers *env, const void *actions_host, void *obs_host,
                      void *reward_host, uint8_t *done_host, int K, int device);
int f16_host_route(void);

/* ---- model level ------------------------------------------------------ */
/* ODE_3DoF.solve (model/ODE_3DoF.py:8-60): dx = f(x, u); x real[9][n], u real[2][n] (stab cmd [rad],
 * throttle), dx real[9][n] (all nine derivatives, including the V_dot the integrator discards). */
int f16_rhs(int dtype, int64_t n, const void *x, const void *u, void *dx, void *stream);
/* K x F16model.step (model/_interface.py:116-122): explicit Euler with the SAME control held for the K
 * steps, wz clipped to +-60 deg/s, V frozen.  x real[9][n] is updated in place. */
int f16_model_step(int dtype, int64_t n, void *x, const void *u, double dt, int K, void *stream);

/* ---- data level (parity probes of the table/atmosphere code) ----------- */
/* coeff.get_Cx/get_Cy/get_Mz (data/coeff.py:37,8,225) at beta=0, lef=0, sb=0: out real[3][n] */
int f16_coeffs(int dtype, int64_t n, const void *alpha, const void *fi, const void *wz, const void *V, void *out,
               void *stream);
/* thrust.get_thrust(H, M, Pa) (data/thrust.py:6-14): out real[n] */
int f16_thrust(int dtype, int64_t n, const void *H, const void *M, const void *Pa, void *out, void *stream);
/* atmosphere.get_density / get_speed_of_sound (data/atmosphere.py:5-11): rho, a real[n] */
int f16_atmosphere(int dtype, int64_t n, const void *h, void *rho, void *a, void *stream);

/* ---- trim search ------------------------------------------------------- */
/* find_trim_position.optimize_loop / Cost.calculate (F16model/utils/find_trim_position.py:27-66) for a grid of
 * flight conditions: for every (V[c], H[c]) cell and every start u0 = starts[s] = (stab [rad], throttle,
 * alpha [rad]) one Nelder-Mead search with scipy.optimize.fmin's parameters (xtol = ftol = 1e-10, maxiter 1000,
 * maxfun 5000) minimises the reference's weighted squared-derivative cost.  All fp64.
 * out_x double[n_cells][n_starts][3], out_cost double[n_cells][n_starts], out_iters int32[n_cells][n_starts] or NULL.
 * The caller takes the arg-min over starts (the reference stops at the first start with cost <= 1e-5). */
int f16_trim(int n_cells, const void *V, const void *H, int n_starts, const void *starts, void *out_x, void *out_cost,
             void *out_iters, void *stream);

/* ---- agent level -------------------------------------------------------- */
/* Agent.get_action_and_value / get_value (control/ppo/ppo_model_gsde.py:81-99,158-160): the two tanh MLPs
 * 5->64->64->1 (actor_mean, critic) on the tensor cores (tcgen05, TF32 operands, fp32 accumulate), fused with the
 * Gaussian policy head.  obs float[n][5]; blob = weights packed by f16_policy_blob layout (see
 * f16-model_b200/policy.py); logstd float[1]; outputs float[n], each may be NULL.
 * action = mean + exp(logstd) * eps with eps ~ N(0,1) from the counter RNG keyed by (seed, env id, episode_no[i],
 * ep_len[i]) -- the env's own counters, so CUDA-graph replays draw fresh noise -- or by (seed, env id, step) when
 * ep_len is NULL; eps = 0 when sample == 0; logprob is the Gaussian log-density of that action. */
int f16_policy_blob_bytes(void);
int f16_policy_forward(int64_t n, const void *obs, const void *blob, const void *logstd, void *mean, void *value,
                       void *action, void *logprob, int sample, uint64_t seed, uint32_t step, const int32_t *ep_len,
                       const uint32_t *episode_no, int64_t env_id_base, void *stream);

/* The same forward with fp16 tensor-core operands and the rollout kernel's schedule (four independent 128-row groups per CTA,
 * biases inside the MMAs, split-phase layer 2): weights = the fp16 blob of f16_ppo_grad16 / f16_rollout
 * (f16_ppo_train_blob16_bytes, kept current by f16_ppo_adam), so acting, the closed-loop rollout and the update's forward pass
 * round alike.  Same arguments, RNG keys and outputs as f16_policy_forward; 1.5x its throughput at 2^20 rows. */
int f16_policy_forward16(int64_t n, const void *obs, const void *train_blob16, const void *logstd, void *mean, void *value,
                         void *action, void *logprob, int sample, uint64_t seed, uint32_t step, const int32_t *ep_len,
                         const uint32_t *episode_no, int64_t env_id_base, void *stream);

/* The closed-loop rollout of Agent.train (control/ppo/ppo_model_gsde.py:216-230) in ONE persistent kernel launch:
 *     for k in range(K):  action, logprob, value = policy(obs);  obs, reward, done = F16.step(action)   (autoreset)
 * for all n envs, with the state and the observation held on the SM between the steps: the two MLPs run on the tensor
 * cores (tcgen05, fp16 operands, fp32 accumulate in TMEM), the Gaussian head draws its noise from the counter RNG keyed
 * by (policy_seed, global env id, episode_no, ep_len) exactly like f16_policy_forward does with the env's counters, and
 * the env step is the arithmetic of f16_env_step (fp32, bit-identical given the same actions).  Only the rollout rows
 * are written.  Requirements: dtype F16_F32, autoreset 1, obs_layout F16_OBS_AOS, noise 0, pos_comp NULL.
 * train_blob16 = the fp16 weight blob of f16_ppo_grad16 (f16_ppo_train_blob16_bytes), kept current by f16_ppo_adam. */
typedef struct f16_rollout_io {
    const void *train_blob16; /* weights, 16-byte aligned */
    const void *logstd;       /* float[1] */
    const void *obs0;         /* float[n][5]: the observation the first step acts on (of the last reset / step) */
    void *obs_next;           /* float[K][n][5]: observation returned by step k */
    void *actions;            /* float[K][n] */
    void *logprobs;           /* float[K][n] */
    void *values;             /* float[K][n] */
    void *rewards;            /* float[K][n] */
    uint8_t *dones;           /* [K][n] */
    void *next_value;         /* float[n]: critic value of obs_next[K-1] (GAE bootstrap); may be NULL */
    void *final_obs;          /* as in f16_step_io; may be NULL */
    int32_t *final_ep_len;    /* may be NULL */
    void *final_return;       /* may be NULL */
    double *stats;            /* double[4] episode statistics accumulators; may be NULL */
    int32_t sample;           /* 0: eps = 0 (deterministic actions) */
    int32_t reserved;
    uint64_t policy_seed;
} f16_rollout_io;
int f16_rollout(const f16_config *cfg, const f16_env_buffers *env, const f16_rollout_io *io, int K, void *stream);

/* One clipped-PPO minibatch gradient in ONE kernel (the body of Agent.train's minibatch loop,
 * control/ppo/ppo_model_gsde.py:303-364, without the optimiser step): forward of actor_mean and critic on the rows
 * mb_idx[0..n) of the flattened rollout buffers (obs float[N][5]; actions, logprobs, adv, ret, val float[N]), the loss
 * (ratio clipping with clip_coef, optionally clipped value loss, vf_coef, ent_coef, optional advantage normalisation
 * with adv_stats = {mean, std} of the minibatch advantages) and the backward pass of both MLPs.  train_blob = the
 * forward blob followed by W2^T in the same packing (f16_ppo_train_blob_bytes); grad receives, ACCUMULATED with atomics
 * (zero it first), f16_ppo_grad_floats floats: the parameter gradients in torch order per net
 * (w1[2][64][5], b1[2][64], w2[2][64][64], b2[2][64], w3[2][64], b3[2], logstd) and five sums over the minibatch
 * (-logratio, (ratio-1)-logratio, |ratio-1| > clip, value loss, policy loss). */
/* mean and unbiased std of the advantages of n_minibatches consecutive minibatches of n rows each (the advantage normalisation
 * of ppo_model_gsde.py:318-321): minibatch y = adv[mb_idx[y*n .. (y+1)*n)] -> out float[y][2]; scratch = n_minibatches x 4
 * doubles, zeroed once by the caller (the kernel re-arms it).  One launch can serve a whole epoch. */
int f16_ppo_adv_stats(int64_t n, int n_minibatches, const void *adv, const int64_t *mb_idx, void *out, void *scratch, void *stream);
int f16_ppo_train_blob_bytes(void);
int f16_ppo_train_blob16_bytes(void);
int f16_ppo_grad_floats(void);
int f16_ppo_grad(int64_t n, const void *train_blob, const void *obs, const void *actions, const void *logprobs,
                 const void *adv, const void *ret, const void *val, const int64_t *mb_idx, const void *logstd,
                 const void *adv_stats, void *grad, double clip_coef, double vf_coef, double ent_coef, int norm_adv,
                 int clip_vloss, void *stream);
/* Same computation with fp16 tensor-core operands (the 11 significant bits of TF32) and the weight gradients on the
 * tensor cores as well; train_blob16 (f16_ppo_train_blob16_bytes) = W1, W2, W2^T in fp16, then the fp32 biases / heads.
 * mb_idx and the gathered row arrays are requested BEFORE the launch's grid dependency resolves (the launch is chained to the
 * optimiser kernel with programmatic dependent launch): they must come from plain, fully ordered launches (any torch kernel,
 * f16_ppo_randperm, f16_ppo_pack_rows), not from a kernel that releases its programmatic dependents early.
 * scratch: optional f16_ppo_grad16_scratch_floats(device) floats.  With it every CTA stores a private gradient and a second
 * small kernel sums them into grad, which is then OVERWRITTEN (no zeroing needed); without it grad is accumulated with atomics. */
/* The minibatch permutation of an epoch (np.random.permutation(batch) in Agent.train, control/ppo/ppo_model_gsde.py:296-300) without
 * a sort: out[i] = pi(i), pi a keyed bijection of [0, n) -- four rounds of (odd multiply, xor-shift, add) on ceil(log2 n) bits,
 * cycle-walked into range.  One pass over 8 bytes per element (torch.randperm's radix sort costs 0.9 ms for 2^23 elements, a
 * fifth of the fused update); every (n, seed) gives the same permutation on every device. */
int f16_ppo_randperm(int64_t n, uint64_t seed, int64_t *out, void *stream);

/* Packed minibatch rows (f16_ppo_pack_rows): a minibatch row is gathered through mb_idx from six arrays, i.e. from 4-5 different
 * 32-byte DRAM sectors; packed once per update into rows float[N][8] = (obs[5], action, logprob, adv) -- one aligned sector -- and
 * rv float[N][2] = (ret, val), the gather of f16_ppo_grad16 is one sector per thread.  f16_ppo_grad16 reads the packed form when
 * `actions` is NULL: `obs` then points to rows, `ret` to rv, and logprobs / adv / val are ignored. */
int f16_ppo_pack_rows(int64_t N, const void *obs, const void *actions, const void *logprobs, const void *adv, const void *ret,
                      const void *val, void *rows, void *rv, void *stream);
int f16_ppo_grad16_scratch_floats(int device);
int f16_ppo_grad16(int64_t n, const void *train_blob16, const void *obs, const void *actions, const void *logprobs,
                   const void *adv, const void *ret, const void *val, const int64_t *mb_idx, const void *logstd,
                   const void *adv_stats, void *grad, double clip_coef, double vf_coef, double ent_coef, int norm_adv,
                   int clip_vloss, void *scratch, void *stream);

/* f16_ppo_grad16 fused with the trainer's only collective, the gradient all-reduce over the ranks of one node
 * (torch DistributedDataParallel's role around ppo_model_gsde.py:360-364; NCCL needs 20-29 us per call for this 37 KB message):
 * the reduction kernel that sums the CTAs' private gradients also exchanges them through NVLink / NVSwitch PEER MEMORY -- every
 * rank pushes its 32-element slices as (value, call number) 8-byte words into all ranks' exchange buffers (the payload carries its
 * own flag: no fence, no round trip), polls its own buffer until every rank's words show the current call and adds the rows in
 * rank order (bit-identical results on every rank) -- and leaves the MEAN over the ranks in grad.  No NCCL call, no extra launch;
 * replayable from a CUDA graph (the call counters live in peer->seq).
 * Setup, once: every rank f16_peer_alloc()s an exchange buffer of f16_peer_buffer_bytes(world) bytes (zeroed), the 64-byte IPC
 * handles are exchanged by any means (torch.distributed.all_gather), every rank f16_peer_open()s the others' handles, uploads the
 * `world` pointers (own buffer at [rank]) as a device array, zeroes f16_peer_seq_words() uint32 counters, and all ranks meet at a
 * barrier before the first call.  Every rank must make the same sequence of calls.  A rank that never arrives traps the waiting
 * kernel after ~10 s instead of hanging the device. */
typedef struct f16_peer_group {
    int32_t rank, world;   /* 1 <= world <= 32, ranks of ONE node (peer access between their devices) */
    void *const *bufs;     /* DEVICE array void*[world]: rank r's exchange buffer as mapped into this process */
    uint32_t *seq;         /* DEVICE uint32[f16_peer_seq_words()], zeroed once */
} f16_peer_group;
int64_t f16_peer_buffer_bytes(int world);
int f16_peer_seq_words(void);
int f16_peer_alloc(int64_t bytes, void **ptr, void *ipc_handle /* 64 bytes out */);
int f16_peer_open(const void *ipc_handle /* 64 bytes */, void **ptr);
int f16_peer_close(void *ptr);
int f16_peer_free(void *ptr);
int f16_ppo_grad16_peer(int64_t n, const void *train_blob16, const void *obs, const void *actions, const void *logprobs,
                        const void *adv, const void *ret, const void *val, const int64_t *mb_idx, const void *logstd,
                        const void *adv_stats, void *grad, double clip_coef, double vf_coef, double ent_coef, int norm_adv,
                        int clip_vloss, void *scratch, const f16_peer_group *peer, void *stream);

/* One whole minibatch step of Agent.train's update loop (control/ppo/ppo_model_gsde.py:303-364) in TWO launches: the gradient
 * kernel of f16_ppo_grad16, then ONE multi-CTA kernel that sums the CTAs' private gradients (with `peer`: also all-reduces them over
 * NVLink peer memory, as f16_ppo_grad16_peer does), takes the global gradient norm at a grid-wide arrival counter, applies
 * clip_grad_norm_ + Adam(amsgrad) to its 32 parameters and scatters each new parameter to its places in the three weight blobs --
 * f16_ppo_grad16 + f16_ppo_adam without the single-CTA optimiser launch (13 us of dependent latency per minibatch).
 * scatter: int32[n_params][4], 16-byte aligned = the inverse of f16_ppo_adam's gather permutations: for parameter j the word
 * index it occupies in fwd_blob, in tail_blob (the fp32 tail of the fp16 training blob) and up to two fp16 words of half_blob
 * (W2 and W2^T), -1 = none.  sync: f16_ppo_step16_sync_bytes() bytes of device memory, zeroed once (arrival counter, per-CTA norm
 * shares and call counters).  grad receives the (mean) gradient and the five statistics as f16_ppo_grad16 leaves them; diag as in
 * f16_ppo_adam (diag[2] is ADDED to).  step and lr are device scalars: the call replays from a CUDA graph. */
typedef struct f16_ppo_step16_args {
    int64_t n;                 /* minibatch rows */
    const void *train_blob16;
    const void *obs, *actions, *logprobs, *adv, *ret, *val;   /* packed rows when actions == NULL, as in f16_ppo_grad16 */
    const int64_t *mb_idx;
    const void *logstd, *adv_stats;
    void *grad;
    double clip_coef, vf_coef, ent_coef;
    int32_t norm_adv, clip_vloss;
    void *scratch;             /* f16_ppo_grad16_scratch_floats(device) floats, required */
    const f16_peer_group *peer; /* or NULL */
    int32_t n_params, reserved;
    void *param, *exp_avg, *exp_avg_sq, *max_exp_avg_sq, *step;
    const void *lr;
    double beta1, beta2, eps, max_grad_norm;
    const int32_t *scatter;
    void *fwd_blob, *tail_blob, *half_blob;
    void *diag;                /* float[6] or NULL */
    double inv_rows;
    void *sync;
} f16_ppo_step16_args;
int f16_ppo_step16_sync_bytes(void);
int f16_ppo_step16(const f16_ppo_step16_args *a, void *stream);

/* A whole EPOCH of Agent.train's update loop (control/ppo/ppo_model_gsde.py:295-364: the inner `for start in range(0, batch_size,
 * minibatch_size)` loop) in ONE persistent cooperative launch on one rank: n_minibatches steps, each the gradient of f16_ppo_grad16
 * followed by the optimiser step of f16_ppo_step16, with the optimiser step taken INSIDE the kernel -- the CTAs meet at a grid-wide
 * counter, each sums the partial gradients of the parameters it owns, the squared-norm shares are exchanged, the owners apply
 * clip_grad_norm_ + Adam(amsgrad) and scatter the new parameters into the three blobs, and every CTA re-reads the training blob.
 * Arguments as f16_ppo_step16 except: mb_idx holds n_minibatches * n indices (minibatch k reads [k n, (k + 1) n)), adv_stats holds
 * n_minibatches (mean, std) pairs, and sync points to f16_ppo_epoch16_sync_bytes() bytes of its own (not the buffer f16_ppo_step16
 * uses; no initialisation needed).  With `peer` the reduction phase also takes the ranks' mean over NVLink peer memory (same buffers,
 * same call numbering as f16_ppo_step16 / f16_ppo_grad16_peer: the calls can be mixed; every rank must launch the same grid).  diag, step: as after n_minibatches calls of f16_ppo_step16; grad holds
 * the last minibatch's gradient and statistics.  The call replays from a CUDA graph. */
int f16_ppo_epoch16_sync_bytes(void);
int f16_ppo_epoch16(const f16_ppo_step16_args *a, int n_minibatches, void *stream);

/* The optimiser half of the minibatch step (ppo_model_gsde.py:360-364: clip_grad_norm_(max_grad_norm) then
 * Adam(amsgrad=True).step()) on the flat parameter vector (same field order as the gradient of f16_ppo_grad), followed
 * by re-packing the rollout blob and the training blob from the updated parameters: blob word k = param[perm[k] - 1],
 * 0 where perm[k] == 0; a third destination (half_perm / half_blob) is written as fp16.  Optional epilogue: diag[6]
 * receives (old_approx_kl, approx_kl, clip fraction ADDED, value loss, policy loss, entropy) from the five sums
 * f16_ppo_grad left behind the parameter gradients (grad[n_params .. n_params+5)) times inv_rows, and zero_grad clears
 * the gradient buffer (n_params + 5 floats) for the next minibatch.  step and lr are device scalars, so the call can be
 * replayed from a CUDA graph. */
int f16_ppo_adam(int n_params, void *param, const void *grad, void *exp_avg, void *exp_avg_sq, void *max_exp_avg_sq,
                 void *step, const void *lr, double beta1, double beta2, double eps, double max_grad_norm,
                 const int64_t *fwd_perm, void *fwd_blob, int n_fwd, const int64_t *train_perm, void *train_blob,
                 int n_train, const int64_t *half_perm, void *half_blob, int n_half, void *diag, const void *logstd,
                 double inv_rows, int zero_grad, void *stream);

/* The GAE reverse scan of Agent.train (control/ppo/ppo_model_gsde.py:244-268): float [T][n] rewards, values;
 * done_after u8[T][n] = the done flag returned by step t (the reference's dones[t+1], and next_done for the last
 * step) exactly as f16_env_step writes it; bootstrap next_value float[n] -> advantages, returns float[T][n]. */
int f16_gae(int T, int64_t n, const void *rewards, const void *values, const uint8_t *done_after, const void *next_value,
            double gamma, double gae_lambda, void *advantages, void *returns, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* F16_B200_H */
