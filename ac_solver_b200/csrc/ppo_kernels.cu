// ppo_kernels.cu -- the non-GEMM part of the PPO update path, fused:
//   reference: ac_solver/agents/training.py:230-250 (GAE), :262-330 (losses); paths relative to /root/reference.
// The actor / critic matrices stay in torch (cuBLAS); what the reference spends ~40 small elementwise
// and reduction launches per minibatch on -- log-softmax, entropy, ratio, the clipped / KL-penalised
// policy loss, the clipped value loss, advantage normalisation, and their backward -- is ONE pass
// here that produces the loss terms AND the gradients with respect to the logits and the values,
// so the whole minibatch step (forward GEMMs, this kernel, backward GEMMs, clip, Adam) is a short
// fixed launch sequence that is captured in a CUDA graph (agents/training.py).  GAE is a
// backward-in-time scan, one thread per environment, coalesced over environments.
#include <algorithm>
#include <cstdint>

#include <cuda_runtime.h>

#include "../../include/acsolver_b200.h"
#include "acs_internal.h"

namespace acs {

// training.py:236-249, same fp32 operation order (no FMA contraction: results equal torch's elementwise ops)
__global__ void __launch_bounds__(128) gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                  const float* __restrict__ dones, const float* __restrict__ next_value,
                                                  const float* __restrict__ next_done, float* __restrict__ adv,
                                                  float* __restrict__ ret, int T, int64_t N, float gamma, float gamma_lambda) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float nextnonterminal = __fsub_rn(1.0f, next_done[i]);
    float nextvalues = next_value[i];
    float last = 0.0f;
    for (int t = T - 1; t >= 0; --t) {
        const int64_t o = (int64_t)t * N + i;
        const float v = values[o];
        // delta = rewards[t] + gamma * nextvalues * nextnonterminal - values[t]
        const float delta = __fsub_rn(__fadd_rn(rewards[o], __fmul_rn(__fmul_rn(gamma, nextvalues), nextnonterminal)), v);
        // advantages[t] = lastgaelam = delta + gamma * gae_lambda * nextnonterminal * lastgaelam
        last = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lambda, nextnonterminal), last));
        adv[o] = last;
        ret[o] = __fadd_rn(last, v);
        nextnonterminal = __fsub_rn(1.0f, dones[o]);
        nextvalues = v;
    }
}

constexpr int kLossThreads = 256;
constexpr int kLossTerms = 8;  // pg, v, entropy, kl, clipfrac, (unused x3)

__device__ __forceinline__ void block_reduce_store(double (&acc)[kLossTerms], double* partial) {
    __shared__ double sh[kLossThreads / 32][kLossTerms];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kLossTerms; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (lane == 0) sh[wid][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLossTerms) {
        double v = 0;
        for (int w = 0; w < kLossThreads / 32; ++w) v += sh[w][threadIdx.x];
        partial[(size_t)blockIdx.x * kLossTerms + threadIdx.x] = v;
    }
}

// sum and sum of squares of the minibatch advantages (double accumulators; fixed reduction order)
__global__ void __launch_bounds__(kLossThreads) adv_stats_kernel(const float* __restrict__ adv, int64_t B, double* partial) {
    double acc[kLossTerms] = {};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        const double a = adv[i];
        acc[0] += a;
        acc[1] += a * a;
    }
    block_reduce_store(acc, partial);
}
// stats = {mean, 1 / (std_unbiased + 1e-8)}  (training.py:282-285; torch.std is the unbiased estimator)
__global__ void adv_stats_final_kernel(const double* partial, int nblocks, int64_t B, float* stats) {
    if (threadIdx.x != 0) return;
    double s = 0, q = 0;
    for (int b = 0; b < nblocks; ++b) {
        s += partial[(size_t)b * kLossTerms];
        q += partial[(size_t)b * kLossTerms + 1];
    }
    const double mean = s / (double)B;
    double var = B > 1 ? (q - s * mean) / (double)(B - 1) : 0.0;
    if (var < 0) var = 0;
    stats[0] = (float)mean;
    stats[1] = (float)(1.0 / (sqrt(var) + 1e-8));
}

struct LossParams {
    const float* logits;      // [B, A]
    const float* newvalue;    // [B]
    const int64_t* action;    // [B]
    const float* old_logprob; // [B]
    const float* adv;         // [B] raw advantages
    const float* ret;         // [B]
    const float* old_value;   // [B]
    const float* adv_stats;   // {mean, inv_std} or null (no normalisation)
    const float* beta;        // device scalar (KL-penalty coefficient) when !loss_clip
    float* dlogits;           // [B, A]  d loss / d logits
    float* dvalue;            // [B]     d loss / d newvalue
    double* partial;          // [blocks][kLossTerms]
    int64_t B;
    int A;
    int loss_clip, clip_vloss;
    float clip_coef, ent_coef, vf_coef;
};

// one thread per sample: forward terms and the gradient of
//   loss = mean(pg) - ent_coef * mean(entropy) + vf_coef * v_loss            (training.py:287-318)
template <int MAXA>
__global__ void __launch_bounds__(kLossThreads) ppo_loss_kernel(const LossParams P) {
    double acc[kLossTerms] = {};
    const float invB = 1.0f / (float)P.B;
    const float beta = P.loss_clip ? 0.0f : *P.beta;
    const float mean = P.adv_stats ? P.adv_stats[0] : 0.0f, inv_std = P.adv_stats ? P.adv_stats[1] : 1.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.B; i += (int64_t)gridDim.x * blockDim.x) {
        float z[MAXA];
        float zmax = -INFINITY;
#pragma unroll
        for (int k = 0; k < MAXA; ++k) {
            z[k] = k < P.A ? P.logits[i * P.A + k] : -INFINITY;
            zmax = fmaxf(zmax, z[k]);
        }
        float se = 0.f;
#pragma unroll
        for (int k = 0; k < MAXA; ++k) se += k < P.A ? __expf(z[k] - zmax) : 0.f;
        const float lse = zmax + __logf(se);
        const int a = (int)P.action[i];
        // an action outside [0, A) has no log-probability (torch's Categorical.log_prob raises): poison the row
        float H = 0.f, logp_a = (a >= 0 && a < P.A) ? 0.f : __int_as_float(0x7fc00000);
        float p[MAXA];
#pragma unroll
        for (int k = 0; k < MAXA; ++k) {
            const float lp = z[k] - lse;
            p[k] = k < P.A ? __expf(lp) : 0.f;
            if (k < P.A) H -= p[k] * lp;
            if (k == a) logp_a = lp;
        }
        const float logratio = logp_a - P.old_logprob[i];
        const float ratio = __expf(logratio);
        const float kl_var = (ratio - 1.0f) - logratio;
        const float A = (P.adv[i] - mean) * inv_std;
        const float pg1 = -A * ratio;
        float pg, dpg_dlogratio;
        if (P.loss_clip) {
            const float rc = fminf(fmaxf(ratio, 1.0f - P.clip_coef), 1.0f + P.clip_coef);
            const float pg2 = -A * rc;
            pg = fmaxf(pg1, pg2);
            // inside the clip range both branches coincide (gradient -A*ratio through either); outside, the
            // clamped branch is constant
            dpg_dlogratio = pg1 >= pg2 ? -A * ratio : 0.0f;
        } else {
            pg = pg1 + beta * kl_var;
            dpg_dlogratio = -A * ratio + beta * (ratio - 1.0f);
        }
        const float v = P.newvalue[i], r = P.ret[i];
        float vterm, dv;
        if (P.clip_vloss) {
            const float ov = P.old_value[i];
            const float dlt = v - ov;
            const float vc = ov + fminf(fmaxf(dlt, -P.clip_coef), P.clip_coef);
            const float lu = (v - r) * (v - r), lc = (vc - r) * (vc - r);
            vterm = fmaxf(lu, lc);
            const float du = 2.0f * (v - r);
            const float dc = (dlt >= -P.clip_coef && dlt <= P.clip_coef) ? 2.0f * (vc - r) : 0.0f;
            dv = lu > lc ? du : (lc > lu ? dc : 0.5f * (du + dc));
        } else {
            vterm = (v - r) * (v - r);
            dv = 2.0f * (v - r);
        }
        P.dvalue[i] = 0.5f * P.vf_coef * dv * invB;
#pragma unroll
        for (int k = 0; k < MAXA; ++k) {
            if (k < P.A) {
                const float lp = z[k] - lse;
                const float dH = -p[k] * (lp + H);  // d entropy / d logit k
                P.dlogits[i * P.A + k] = (dpg_dlogratio * ((k == a ? 1.0f : 0.0f) - p[k]) - P.ent_coef * dH) * invB;
            }
        }
        acc[0] += pg;
        acc[1] += vterm;
        acc[2] += H;
        acc[3] += kl_var;
        acc[4] += fabsf(ratio - 1.0f) > P.clip_coef ? 1.0 : 0.0;
    }
    block_reduce_store(acc, P.partial);
}

// out = {loss, pg_loss, v_loss, entropy, approx_kl, clipfrac}
__global__ void ppo_loss_final_kernel(const double* partial, int nblocks, int64_t B, float ent_coef, float vf_coef, float* out) {
    if (threadIdx.x != 0) return;
    double s[5] = {};
    for (int b = 0; b < nblocks; ++b)
        for (int k = 0; k < 5; ++k) s[k] += partial[(size_t)b * kLossTerms + k];
    const double pg = s[0] / B, vl = 0.5 * s[1] / B, H = s[2] / B;
    out[0] = (float)(pg - ent_coef * H + vf_coef * vl);
    out[1] = (float)pg;
    out[2] = (float)vl;
    out[3] = (float)H;
    out[4] = (float)(s[3] / B);
    out[5] = (float)(s[4] / B);
}

}  // namespace acs

using namespace acs;

extern "C" {

int acs_gae(const float* d_rewards, const float* d_values, const float* d_dones, const float* d_next_value,
            const float* d_next_done, float* d_advantages, float* d_returns, int T, int64_t N, double gamma, double gae_lambda,
            void* stream) {
    if (T < 0 || N < 0) {
        acs::set_last_error("gae: bad shape");
        return ACS_ERR_INVALID;
    }
    if (T == 0 || N == 0) return ACS_OK;
    if (!d_rewards || !d_values || !d_dones || !d_next_value || !d_next_done || !d_advantages || !d_returns) {
        acs::set_last_error("gae: null buffer");
        return ACS_ERR_INVALID;
    }
    // the reference multiplies fp32 tensors by the Python floats gamma and gamma * gae_lambda
    gae_kernel<<<(unsigned)((N + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        d_rewards, d_values, d_dones, d_next_value, d_next_done, d_advantages, d_returns, T, N, (float)gamma,
        (float)(gamma * gae_lambda));
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        acs::set_last_error(cudaGetErrorString(e));
        return ACS_ERR_CUDA;
    }
    return ACS_OK;
}

int acs_ppo_loss_workspace_bytes(void) { return 1024 * kLossTerms * (int)sizeof(double) + 64; }

int acs_ppo_loss(const float* d_logits, const float* d_newvalue, const int64_t* d_action, const float* d_old_logprob,
                 const float* d_adv, const float* d_returns, const float* d_old_value, const float* d_beta, float* d_dlogits,
                 float* d_dvalue, float* d_out8, void* d_workspace, int64_t B, int n_actions, int norm_adv, int loss_clip,
                 int clip_vloss, double clip_coef, double ent_coef, double vf_coef, void* stream) {
    if (B < 1 || n_actions < 1 || n_actions > 16) {
        acs::set_last_error("ppo_loss: need B >= 1 and 1 <= n_actions <= 16");
        return ACS_ERR_INVALID;
    }
    if (!d_logits || !d_newvalue || !d_action || !d_old_logprob || !d_adv || !d_returns || !d_old_value || !d_dlogits ||
        !d_dvalue || !d_out8 || !d_workspace || (!loss_clip && !d_beta)) {
        acs::set_last_error("ppo_loss: null buffer");
        return ACS_ERR_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* partial = static_cast<double*>(d_workspace);
    float* stats = reinterpret_cast<float*>(static_cast<char*>(d_workspace) + 1024 * kLossTerms * sizeof(double));
    const int blocks = (int)std::min<int64_t>(1024, (B + kLossThreads - 1) / kLossThreads);
    if (norm_adv) {
        adv_stats_kernel<<<blocks, kLossThreads, 0, s>>>(d_adv, B, partial);
        adv_stats_final_kernel<<<1, 32, 0, s>>>(partial, blocks, B, stats);
    }
    LossParams P{};
    P.logits = d_logits;
    P.newvalue = d_newvalue;
    P.action = d_action;
    P.old_logprob = d_old_logprob;
    P.adv = d_adv;
    P.ret = d_returns;
    P.old_value = d_old_value;
    P.adv_stats = norm_adv ? stats : nullptr;
    P.beta = d_beta;
    P.dlogits = d_dlogits;
    P.dvalue = d_dvalue;
    P.partial = partial;
    P.B = B;
    P.A = n_actions;
    P.loss_clip = loss_clip ? 1 : 0;
    P.clip_vloss = clip_vloss ? 1 : 0;
    P.clip_coef = (float)clip_coef;
    P.ent_coef = (float)ent_coef;
    P.vf_coef = (float)vf_coef;
    ppo_loss_kernel<16><<<blocks, kLossThreads, 0, s>>>(P);
    ppo_loss_final_kernel<<<1, 32, 0, s>>>(partial, blocks, B, (float)ent_coef, (float)vf_coef, d_out8);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        acs::set_last_error(cudaGetErrorString(e));
        return ACS_ERR_CUDA;
    }
    return ACS_OK;
}

}  // extern "C"

// ---- rollout bookkeeping of the PPO loop, fused (agents/training.py:139-228 of the reference) ------------
// The reference's rollout step is a dozen small tensor writes around the policy and the environment
// (obs[step] = next_obs, dones[step] = next_done, values / actions / logprobs[step], rewards[step],
// episodic return / length bookkeeping, the deque of finished episodes).  On the device each of them is a
// kernel launch of a few microseconds inside the rollout graph; here they are two kernels:
//   sample_record (before the env step): Categorical sampling by the Gumbel-max trick from a counter-based
//       generator, log-probability of the drawn action, and the writes into the [T, N] buffers at time t;
//   finish (after the env step and the reward wrappers): rewards[t], next_done, episodic statistics.
namespace acs {

__device__ __forceinline__ unsigned long long rl_mix(unsigned long long x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

struct RolloutRecordParams {
    const int8_t* state;       // [N, width] current observations
    const float* next_done;    // [N]
    const float* logits;       // [N, A]
    const float* value;        // [N]
    const long long* ctr;      // device scalars {t, draw counter}
    int8_t* obs_buf;           // [T, N, width]
    float* dones_buf;          // [T, N]
    float* values_buf;         // [T, N]
    float* logprobs_buf;       // [T, N]
    long long* actions_buf;    // [T, N]
    uint8_t* action_u8;        // [N]  the drawn actions for the environment kernel
    long long N, T;
    int width, A;
    unsigned long long seed;
};

__global__ void __launch_bounds__(256) rollout_sample_record_kernel(const RolloutRecordParams P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N) return;
    const long long t = P.ctr[0], draw = P.ctr[1];
    if (t < 0 || t >= P.T) return;  // the caller rolls out at most T steps between resets of the time index
    const long long o = t * P.N + i;
    // Categorical(logits).sample() as argmax(logits + Gumbel noise); log_prob of the drawn action
    float zmax = -INFINITY, best = -INFINITY;
    int a = 0;
    float z[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        z[k] = k < P.A ? P.logits[i * P.A + k] : -INFINITY;
        zmax = fmaxf(zmax, z[k]);
    }
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (k < P.A) {
            se += __expf(z[k] - zmax);
            const unsigned long long h = rl_mix(P.seed ^ rl_mix((unsigned long long)draw * 0x9E3779B97F4A7C15ull + (unsigned long long)i * 16ull + k));
            const float u = ((float)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
            const float g = z[k] - __logf(-__logf(u));
            if (g > best) {
                best = g;
                a = k;
            }
        }
    }
    float za = z[0];
#pragma unroll
    for (int k = 1; k < 16; ++k) za = (k == a) ? z[k] : za;
    P.actions_buf[o] = a;
    P.action_u8[i] = (uint8_t)a;
    P.logprobs_buf[o] = za - (zmax + __logf(se));
    P.values_buf[o] = P.value[i];
    P.dones_buf[o] = P.next_done[i];
    const int8_t* src = P.state + i * P.width;
    int8_t* dst = P.obs_buf + o * P.width;
    if ((P.width & 7) == 0) {
        for (int k = 0; k < P.width / 8; ++k) reinterpret_cast<uint2*>(dst)[k] = reinterpret_cast<const uint2*>(src)[k];
    } else {
        for (int k = 0; k < P.width; ++k) dst[k] = src[k];
    }
}

struct RolloutFinishParams {
    const float* reward;       // [N] reward of the step as the loop receives it (after the wrappers)
    const uint8_t* done;       // [N]
    const uint8_t* truncated;  // [N]
    const long long* ctr;      // {t, draw counter}
    float* rewards_buf;        // [T, N]
    float* next_done;          // [N]
    float* ep_return;          // [N]
    float* ep_length;          // [N]
    float* ring_ret;           // [ring]  returns / lengths of the last `ring` finished episodes
    float* ring_len;
    unsigned long long* counters;  // {episodes finished}
    long long N, T;
    int ring;
};

__global__ void __launch_bounds__(256) rollout_finish_kernel(const RolloutFinishParams P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N) return;
    const long long t = P.ctr[0];
    if (t < 0 || t >= P.T) return;
    const float r = P.reward[i];
    P.rewards_buf[t * P.N + i] = r;
    const bool d = P.done[i] != 0;
    P.next_done[i] = d ? 1.0f : 0.0f;  // torch.Tensor(done): terminated only (training.py:226-228)
    const float er = P.ep_return[i] + r, el = P.ep_length[i] + 1.0f;
    if (d || P.truncated[i]) {  // training.py:191-196: the episode's return and length enter the deque(maxlen=100)
        const unsigned long long k = atomicAdd(&P.counters[0], 1ull);
        P.ring_ret[k % (unsigned long long)P.ring] = er;
        P.ring_len[k % (unsigned long long)P.ring] = el;
        P.ep_return[i] = 0.f;
        P.ep_length[i] = 0.f;
    } else {
        P.ep_return[i] = er;
        P.ep_length[i] = el;
    }
}

}  // namespace acs

extern "C" {

int acs_rollout_sample_record(const int8_t* d_state, const float* d_next_done, const float* d_logits, const float* d_value,
                              const int64_t* d_ctr2, int8_t* d_obs_buf, float* d_dones_buf, float* d_values_buf,
                              float* d_logprobs_buf, int64_t* d_actions_buf, uint8_t* d_action_u8, int64_t N, int64_t T, int width,
                              int n_actions, uint64_t seed, void* stream) {
    if (N < 0 || T < 1 || width < 1 || n_actions < 1 || n_actions > 16) {
        acs::set_last_error("rollout_sample_record: bad shape (1 <= n_actions <= 16)");
        return ACS_ERR_INVALID;
    }
    if (N == 0) return ACS_OK;
    if (!d_state || !d_next_done || !d_logits || !d_value || !d_ctr2 || !d_obs_buf || !d_dones_buf || !d_values_buf ||
        !d_logprobs_buf || !d_actions_buf || !d_action_u8) {
        acs::set_last_error("rollout_sample_record: null buffer");
        return ACS_ERR_INVALID;
    }
    acs::RolloutRecordParams P{};
    P.state = d_state;
    P.next_done = d_next_done;
    P.logits = d_logits;
    P.value = d_value;
    P.ctr = reinterpret_cast<const long long*>(d_ctr2);
    P.obs_buf = d_obs_buf;
    P.dones_buf = d_dones_buf;
    P.values_buf = d_values_buf;
    P.logprobs_buf = d_logprobs_buf;
    P.actions_buf = reinterpret_cast<long long*>(d_actions_buf);
    P.action_u8 = d_action_u8;
    P.N = N;
    P.T = T;
    P.width = width;
    P.A = n_actions;
    P.seed = seed;
    acs::rollout_sample_record_kernel<<<(unsigned)((N + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(P);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        acs::set_last_error(cudaGetErrorString(e));
        return ACS_ERR_CUDA;
    }
    return ACS_OK;
}

int acs_rollout_finish(const float* d_reward, const uint8_t* d_done, const uint8_t* d_truncated, const int64_t* d_ctr2,
                       float* d_rewards_buf, float* d_next_done, float* d_ep_return, float* d_ep_length, float* d_ring_ret,
                       float* d_ring_len, uint64_t* d_counters, int64_t N, int64_t T, int ring, void* stream) {
    if (N < 0 || T < 1 || ring < 1) {
        acs::set_last_error("rollout_finish: bad shape");
        return ACS_ERR_INVALID;
    }
    if (N == 0) return ACS_OK;
    if (!d_reward || !d_done || !d_truncated || !d_ctr2 || !d_rewards_buf || !d_next_done || !d_ep_return || !d_ep_length ||
        !d_ring_ret || !d_ring_len || !d_counters) {
        acs::set_last_error("rollout_finish: null buffer");
        return ACS_ERR_INVALID;
    }
    acs::RolloutFinishParams P{};
    P.reward = d_reward;
    P.done = d_done;
    P.truncated = d_truncated;
    P.ctr = reinterpret_cast<const long long*>(d_ctr2);
    P.rewards_buf = d_rewards_buf;
    P.next_done = d_next_done;
    P.ep_return = d_ep_return;
    P.ep_length = d_ep_length;
    P.ring_ret = d_ring_ret;
    P.ring_len = d_ring_len;
    P.counters = reinterpret_cast<unsigned long long*>(d_counters);
    P.N = N;
    P.T = T;
    P.ring = ring;
    acs::rollout_finish_kernel<<<(unsigned)((N + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(P);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        acs::set_last_error(cudaGetErrorString(e));
        return ACS_ERR_CUDA;
    }
    return ACS_OK;
}

}  // extern "C"
