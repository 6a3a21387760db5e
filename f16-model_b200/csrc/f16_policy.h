// f16_policy.h -- packed weights and launch arguments of the tensor-core policy forward.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace f16 {

// Weights of actor_mean and critic (control/ppo/ppo_model_gsde.py:37-52: Linear(5,64) Tanh
// Linear(64,64) Tanh Linear(64,1) each), pre-arranged in the tensor core's canonical K-major
// no-swizzle layout: [16-byte K-chunk][output row][4 floats].  Packed by policy.py.
struct alignas(16) PolicyBlob {
    float w1[2][128][4];      // rows 0-63 actor, 64-127 critic; K = 5 inputs padded to 8
    float w2[2][16][64][4];   // [actor|critic][K-chunk][row][4]
    float b1[128];            // actor | critic
    float b2[128];
    float w3[128];            // output layer weights, actor | critic
    float b3[4];              // {actor bias, critic bias, 0, 0}
};

struct PolicyArgs {
    const PolicyBlob *blob;   // device
    const float *obs;         // [n][5]
    const float *logstd;      // [1] state-independent log std of the Gaussian policy, or nullptr
    float *mean;              // [n] or nullptr
    float *value;             // [n] or nullptr
    float *action;            // [n] or nullptr: mean + std * eps
    float *logprob;           // [n] or nullptr
    int64_t n;
    int sample;               // 0: eps = 0 (deterministic action)
    uint32_t step;            // RNG counter (rollout step), used when ep_len is null
    const int32_t *ep_len;    // [n] env step counters / episode counters as RNG counters, or nullptr
    const uint32_t *episode_no;
    uint64_t seed;
    int64_t id_base;          // global id of env 0 (sharding-invariant noise)
};

// ---- fused PPO minibatch gradient (forward + clipped-PPO loss + backward of both MLPs in one kernel) ----
// weights for the training kernel: the forward blob plus W2^T (the B operand of dH1 = dZ2 * W2), same K-major packing
struct alignas(16) TrainBlob {
    PolicyBlob fwd;
    float w2t[2][16][64][4];   // [actor|critic][K-chunk over OUTPUT features j][row = input feature i][4]
};

// gradient of the minibatch loss, laid out like the torch parameters (actor | critic), plus the loss statistics
struct PolicyGrad {
    float w1[2][64][5];
    float b1[2][64];
    float w2[2][64][64];
    float b2[2][64];
    float w3[2][64];
    float b3[2];
    float logstd;
    float stats[5];            // sums over the minibatch: -logratio, (ratio-1)-logratio, |ratio-1| > clip, value loss, policy loss
};

struct PpoGradArgs {
    const TrainBlob *blob;
    const float *obs;          // [N][5] rollout buffer (flattened T*n)
    const float *actions, *logprobs, *adv, *ret, *val;   // [N]
    const int64_t *mb_idx;     // [n] rows of this minibatch
    const float *logstd;       // [1]
    const float *adv_stats;    // [2] mean, std of the minibatch advantages (used when norm_adv)
    PolicyGrad *grad;          // accumulated with atomics: zero it first
    int64_t n;
    float clip_coef, vf_coef, ent_coef;
    int norm_adv, clip_vloss;
    int packed;                // ppo_grad16 only: obs = rows float[N][8] (obs[5], action, logprob, adv), ret = float[N][2] (ret, val)
};

cudaError_t launch_ppo_grad(const PpoGradArgs &P, int n_sm, cudaStream_t st);

// fp16-operand variant of the training kernel: every tensor-core operand (X, W1, H1, W2, W2^T, dZ2, dZ1) is fp16 -- the
// same 11 significant bits as TF32 -- because MN-major no-swizzle operands exist for kind::f16 but not for kind::tf32:
// the [8-feature chunk][row][8 halfs] buffers are then both the (row x feature) operands of the forward / backward-data
// products and the (feature x row) operands of the weight gradients, which move onto the tensor cores with no copy.
struct alignas(16) TrainBlob16 {
    __half w1[2][128][8];      // layer 1, K = 16: k < 5 the observation weights, the rest zero
    __half w2[2][8][64][8];    // [net][K-chunk over inputs i][row = output j][8]
    __half w2t[2][8][64][8];   // [net][K-chunk over outputs j][row = input i][8]
    float b1[128], b2[128], w3[128], b3[4];
};
// partials: optional [n_sm][sizeof(PolicyGrad)/4] floats.  With it every CTA stores its own gradient (no atomics: 148 CTAs adding
// into the same 9 k addresses cost a fifth of the kernel) and a second small kernel sums the CTAs' rows into P.grad (overwritten).
// peer: optional cross-rank exchange (see ppo_grad_reduce_peer_kernel): the reduction kernel then also all-reduces (means) the gradient
// over the ranks of the node through peer memory; needs `partials`.
struct PeerArgs {
    uint2 *const *bufs;   // DEVICE array of `world` pointers: every rank's exchange buffer as mapped into this process (own one at [rank])
    uint32_t *seq;        // DEVICE call counters of this rank, one per reduction CTA, zeroed once
    int rank, world;
};
// gradient reduction (+ optional peer exchange) fused with clip_grad_norm_ + Adam(amsgrad) + blob re-packing: ppo_reduce_adam_kernel
struct ReduceAdamArgs {
    const float *partials;   // filled in by the launcher
    int n_rows, n_out;
    float *grad;
    uint2 *const *bufs;      // peer exchange (nullptr: single rank), see PeerArgs
    uint32_t *seq;           // peer exchange: its call counters (PeerArgs::seq)
    int rank, world, n_pad;
    int n;                   // parameters (n_out - 5)
    float *param, *exp_avg, *exp_avg_sq, *max_exp_avg_sq, *step;
    const float *lr;
    float beta1, beta2, eps, max_grad_norm;
    const int32_t *scatter;  // [n][4]: destination word of parameter j in fwd_blob, tail_blob, half_blob (twice), -1 = none; 16-byte aligned
    float *fwd_blob, *tail_blob;
    __half *half_blob;
    float *diag;             // [6] or nullptr
    float inv_rows;
    unsigned char *sync;     // ppo_step16_sync_bytes() bytes, zeroed once: arrival counter, per-CTA squared-norm shares, per-CTA call counters
};
size_t ppo_step16_sync_bytes();
cudaError_t launch_ppo_step16(const PpoGradArgs &P, const TrainBlob16 *blob16, float *partials, ReduceAdamArgs RA, int n_sm, cudaStream_t st);
// the n_mb minibatch steps of an epoch in ONE persistent cooperative launch (RA.bufs set: with the cross-rank exchange in its reduction phase): P.mb_idx holds n_mb * P.n indices, P.adv_stats n_mb
// pairs; RA.sync: ppo_epoch16_sync_bytes() bytes of its own (the launcher zeroes the meeting counter)
size_t ppo_epoch16_sync_bytes();
cudaError_t launch_ppo_epoch16(const PpoGradArgs &P, const TrainBlob16 *blob16, float *partials, ReduceAdamArgs RA, int n_mb, int n_sm, cudaStream_t st);
int ppo_grad_floats_padded();
int ppo_grad_reduce_ctas();
cudaError_t launch_ppo_grad16(const PpoGradArgs &P, const TrainBlob16 *blob16, float *partials, int n_sm, cudaStream_t st, const PeerArgs *peer = nullptr);

// clip_grad_norm_ + Adam(amsgrad) + re-packing of the two weight blobs, one CTA (the whole model is 9.3 k parameters)
struct PpoAdamArgs {
    int n;                      // parameters
    float *param;               // [n] flat parameters (PolicyGrad field order), updated in place
    const float *grad;          // [n]
    float *exp_avg, *exp_avg_sq, *max_exp_avg_sq;   // [n] Adam state
    float *step;                // [1] step counter (float, as torch keeps it when capturable)
    const float *lr;            // [1]
    float beta1, beta2, eps, max_grad_norm;
    const int64_t *fwd_perm;    // [n_fwd]: blob word k = param[perm[k] - 1], or 0 where perm[k] == 0
    float *fwd_blob;
    int n_fwd;
    const int64_t *train_perm;
    float *train_blob;
    int n_train;
    const int64_t *half_perm;   // optional third destination, written as fp16 (the tensor-core operands of TrainBlob16)
    __half *half_blob;
    int n_half;
    // optional epilogue of the minibatch step: diagnostics from the gradient kernel's five sums (grad[n .. n+5)), then the
    // gradient buffer is cleared for the next minibatch
    float *diag;                // [6] old_approx_kl, approx_kl, clip fraction (summed), value loss, policy loss, entropy; or nullptr
    float *grad_rw;             // the same buffer as grad, writable: zeroed (n + 5 floats) when zero_grad != 0
    const float *logstd;        // [1]
    float inv_rows;             // 1 / minibatch rows
    int zero_grad;
};
cudaError_t launch_ppo_adam(const PpoAdamArgs &A, cudaStream_t st);
cudaError_t launch_ppo_randperm(int64_t n, uint64_t seed, int64_t *out, int n_sm, cudaStream_t st);
cudaError_t launch_ppo_pack_rows(int64_t n, const float *obs, const float *actions, const float *logprobs, const float *adv, const float *ret,
                                 const float *val, float *rows, float *rv, int n_sm, cudaStream_t st);
// mean and unbiased std of adv[mb_idx[y*n .. (y+1)*n)] -> out[y][2] for y < n_minibatches; scratch = n_minibatches x 4 doubles,
// zero before the first call (the kernel re-zeroes it)
cudaError_t launch_adv_stats(const float *adv, const int64_t *mb_idx, int64_t n, int n_minibatches, float *out, double *scratch, int n_sm,
                             cudaStream_t st);

cudaError_t launch_gae(int T, int64_t n, const float *rewards, const float *values, const uint8_t *done_after, const float *next_value,
                       float gamma, float lam, float *advantages, float *returns, int n_sm, cudaStream_t st);
// ---- closed-loop rollout: K x (policy forward -> Gaussian head -> F16.step) in one persistent kernel (f16_rollout.cu) ----
template <typename R>
struct EnvArgs;
struct RolloutArgs {
    const TrainBlob16 *blob;      // fp16 tensor-core operands W1, W2 and the fp32 biases / heads (the W2^T block is skipped)
    const float *logstd;          // [1]
    const float *obs0;            // [n][5]   the observation the first step acts on
    float *obs_next;              // [K][n][5] observation returned by step k (the reset observation where an episode ended)
    float *actions, *logprobs, *values, *rewards;   // [K][n]
    uint8_t *dones;               // [K][n]
    float *next_value;            // [n] critic value of the observation after the last step (GAE bootstrap), or nullptr
    float *final_obs;             // gymnasium final_observation / final_info of the most recent episode that ended, or nullptr
    int32_t *final_ep_len;
    float *final_return;
    double *stats;                // [4] episode statistics accumulators, or nullptr
    int K;
    int sample;                   // 0: deterministic actions (eps = 0)
    uint64_t policy_seed;         // key of the exploration-noise stream (the counter is the env's own: id, episode, step)
    int pace;                     // groups of a CTA admitted to their tanh section at a time (0 = no pacing)
};
cudaError_t launch_rollout(const EnvArgs<float> &A, const RolloutArgs &R, bool fast, int n_sm, int pdl, cudaStream_t st);

size_t policy_smem_bytes();
cudaError_t launch_policy_forward(const PolicyArgs &P, int n_sm, int pdl, cudaStream_t st);
// fp16-operand forward with the rollout kernel's schedule (f16_rollout.cu): P.blob is ignored, the weights come from blob16
cudaError_t launch_policy_forward16(const PolicyArgs &P, const TrainBlob16 *blob16, int n_sm, int pdl, cudaStream_t st);

}  // namespace f16
