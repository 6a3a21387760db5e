// f16_policy.h -- packed weights and launch arguments of the tensor-core policy forward.
#pragma once
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

cudaError_t launch_gae(int T, int64_t n, const float *rewards, const float *values, const uint8_t *done_after, const float *next_value,
                       float gamma, float lam, float *advantages, float *returns, int n_sm, cudaStream_t st);
size_t policy_smem_bytes();
cudaError_t launch_policy_forward(const PolicyArgs &P, int n_sm, int pdl, cudaStream_t st);

}  // namespace f16
