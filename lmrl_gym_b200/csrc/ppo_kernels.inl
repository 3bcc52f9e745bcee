// ppo_kernels.inl — PPO data path after the forward passes (C ABI part 3).
//
// Reference: PPOInference.get_ppo_data_from_token_trajectory_chain (LLM_RL/algorithms/ppo/base_interface.py:572-669) — numpy / python
// loops over chains on the host: KL-shaped rewards (:572-584), get_action_state_next_state_idxs (:230-243), the GAE scan
// get_advantages_and_returns (:253-293), batch-wide whitening (:245-251, :608-615) — and ppo_loss_fn (:71-142).
// Here: one thread per chain walks its tokens backwards (the recurrence is sequential per chain, chains are independent), every
// float32 operation in the reference's order with no FMA contraction so the un-whitened scan is BIT-EXACT; the whitening statistics
// are reduced in double in a fixed order (deterministic; within 1 ulp of numpy's pairwise float32 mean / var).
// HBM-bound streaming work: 5 floats + 1 byte read, 3 floats written per output token.
#include <cuda_runtime.h>
#include <stdint.h>

namespace lmrl {

struct PpoArgs {
    int n_chains;
    const int64_t* chain_off;        // [n_chains + 1] token offsets; values of chain c start at chain_off[c] + c (one bootstrap slot each)
    const float* policy_lp; const float* initial_lp; const float* values; const float* rewards; const uint8_t* sta;
    float gamma, gamma_lam, kl_weight;
    float* adv; float* ret; float* kls;
    double* part;                    // [n_chains][2] per-chain (sum adv, sum adv^2)
    int* part_n;                     // [n_chains] action tokens of the chain
};

// the library is compiled with --use_fast_math (expf -> ex2.approx): the KL estimate exp(lr) - 1 - lr cancels catastrophically and feeds
// the adaptive KL controller, so exp goes through double precision here (unaffected by the flag; the result rounds to the correctly
// rounded float32 exp in all but ~1e-9 of cases) — per-token work, not a hot loop
__device__ __forceinline__ float exp_precise(float x) { return (float)exp((double)x); }

__global__ void ppo_gae_kernel(PpoArgs a) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_chains) return;
    const int64_t t0 = a.chain_off[c], t1 = a.chain_off[c + 1];
    const float* v = a.values + t0 + c;
    float next_v = v[t1 - t0];                       // the bootstrap slot: last value * (1 - done), filled by the host wrapper
    float last = 0.f;
    bool first = true;
    double s = 0.0, s2 = 0.0;
    int n = 0;
    for (int64_t t = t1 - 1; t >= t0; --t) {
        if (!a.sta[t]) { a.adv[t] = 0.f; a.ret[t] = 0.f; a.kls[t] = 0.f; continue; }
        const float lr = __fsub_rn(a.policy_lp[t], a.initial_lp[t]);                     // log_ratio (:572-575)
        a.kls[t] = __fsub_rn(__fsub_rn(exp_precise(lr), 1.f), lr);                              // exp(lr) - 1 - lr (:579)
        const float r = __fsub_rn(a.rewards[t], __fmul_rn(a.kl_weight, lr));             // :581-584
        const float sv = v[t - t0];
        const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(a.gamma, next_v)), sv);     // :284
        last = first ? delta : __fadd_rn(delta, __fmul_rn(a.gamma_lam, last));           // :285 (lastgaelam starts as the int 0)
        first = false;
        a.adv[t] = last;
        a.ret[t] = __fadd_rn(last, sv);                                                  // :288
        next_v = sv;                                                                     // the next state of the previous action token
        s += (double)last; s2 += (double)last * (double)last; ++n;
    }
    a.part[2 * c] = s; a.part[2 * c + 1] = s2; a.part_n[c] = n;
}

// fixed-order reduction of the per-chain partials -> stats[0] = mean, stats[1] = 1 / sqrt(var + 1e-8)   (single block)
__global__ void ppo_whiten_stats_kernel(int n_chains, const double* __restrict__ part, const int* __restrict__ part_n, float* __restrict__ stats) {
    __shared__ double sh[3][256];
    double s = 0.0, s2 = 0.0, n = 0.0;
    for (int c = threadIdx.x; c < n_chains; c += 256) { s += part[2 * c]; s2 += part[2 * c + 1]; n += (double)part_n[c]; }
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = s2; sh[2][threadIdx.x] = n;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; sh[2][threadIdx.x] += sh[2][threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double cnt = sh[2][0] > 0.0 ? sh[2][0] : 1.0;
        const double mean = sh[0][0] / cnt;
        const double var = fmax(sh[1][0] / cnt - mean * mean, 0.0);
        const float meanf = (float)mean, varf = (float)var;
        stats[0] = meanf;
        stats[1] = __frcp_rn(__fsqrt_rn(__fadd_rn(varf, 1e-8f)));                        // jnp.reciprocal(jnp.sqrt(var + 1e-8)) (:249)
    }
}

__global__ void ppo_whiten_kernel(int64_t n, const uint8_t* __restrict__ sta, const float* __restrict__ stats, float* __restrict__ adv) {
    const float mean = stats[0], inv = stats[1];
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        if (sta[t]) adv[t] = __fmul_rn(__fsub_rn(adv[t], mean), inv);
}

// ---- ppo_loss_fn (:71-142): nine masked sums over [batch * (time - 1)] elements, per-block partials in double -------
constexpr int PPO_LOSS_TERMS = 12;
struct PpoLossArgs {
    int64_t n;
    const float* mask;               // should_take_action * attention_mask
    const float* logprobs; const float* values; const float* old_logprobs; const float* old_values; const float* old_adv; const float* old_ret;
    float cliprange_value, ratio_lo, ratio_hi;
    double* part;                    // [gridDim.x][PPO_LOSS_TERMS]
};

__global__ void ppo_loss_partial_kernel(PpoLossArgs a) {
    double acc[PPO_LOSS_TERMS];
#pragma unroll
    for (int k = 0; k < PPO_LOSS_TERMS; ++k) acc[k] = 0.0;
    const float lo_r = a.ratio_lo, hi_r = a.ratio_hi;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const float m = a.mask[i], v = a.values[i], ov = a.old_values[i], ret = a.old_ret[i], adv = a.old_adv[i];
        const float vc = fminf(fmaxf(v, __fsub_rn(ov, a.cliprange_value)), __fadd_rn(ov, a.cliprange_value));
        const float d1 = __fsub_rn(v, ret), d2 = __fsub_rn(vc, ret);
        const float vf1 = __fmul_rn(d1, d1), vf2 = __fmul_rn(d2, d2);
        const float lr = __fmul_rn(__fsub_rn(a.logprobs[i], a.old_logprobs[i]), m);
        const float ratio = exp_precise(lr);
        const float pg1 = __fmul_rn(-adv, ratio), pg2 = __fmul_rn(-adv, fminf(fmaxf(ratio, lo_r), hi_r));
        const float dm = __fmul_rn(d1, m);
        acc[0] += (double)m;
        acc[1] += (double)__fmul_rn(fmaxf(vf1, vf2), m);
        acc[2] += (double)((vf2 > vf1 ? 1.f : 0.f) * m);
        acc[3] += (double)__fsub_rn(__fsub_rn(ratio, 1.f), lr);
        acc[4] += (double)__fmul_rn(fmaxf(pg1, pg2), m);
        acc[5] += (double)((pg2 > pg1 ? 1.f : 0.f) * m);
        acc[6] += (double)__fmul_rn(dm, dm);
        acc[7] += (double)__fmul_rn(ratio, m);
        acc[8] += (double)__fmul_rn(v, m);
        acc[9] += (double)__fmul_rn(ov, m);
        acc[10] += (double)__fmul_rn(ret, m);
        acc[11] += 1.0;
    }
    __shared__ double sh[256];
    for (int k = 0; k < PPO_LOSS_TERMS; ++k) {
        sh[threadIdx.x] = acc[k];
        __syncthreads();
        for (int o = 128; o; o >>= 1) {
            if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) a.part[(size_t)blockIdx.x * PPO_LOSS_TERMS + k] = sh[0];
        __syncthreads();
    }
}

__global__ void ppo_loss_final_kernel(int n_blocks, const double* __restrict__ part, double* __restrict__ out) {
    const int k = threadIdx.x;
    if (k >= PPO_LOSS_TERMS) return;
    double s = 0.0;
    for (int b = 0; b < n_blocks; ++b) s += part[(size_t)b * PPO_LOSS_TERMS + k];
    out[k] = s;
}

}  // namespace lmrl

using namespace lmrl;

extern "C" {

int lmrl_ppo_advantages_host(lmrl_engine* e, int32_t n_chains, const int64_t* chain_off, const float* policy_logprobs, const float* initial_logprobs,
                             const float* values, const float* rewards, const uint8_t* should_take_action, double gamma, double lam, double kl_weight,
                             int32_t whiten, float* advantages, float* returns, float* kls) {
    LMRL_CHECK(e && chain_off && policy_logprobs && initial_logprobs && values && rewards && should_take_action && advantages && returns && kls,
               "lmrl_ppo_advantages: null argument");
    LMRL_CHECK(n_chains > 0 && chain_off[0] == 0, "lmrl_ppo_advantages: need at least one chain and chain_off[0] == 0");
    for (int c = 0; c < n_chains; ++c) LMRL_CHECK(chain_off[c + 1] >= chain_off[c], "lmrl_ppo_advantages: chain_off must be non-decreasing");
    LMRL_CUDA(cudaSetDevice(e->device));
    const int64_t N = chain_off[n_chains];
    cudaStream_t st = e->stream;
    // scratch lives only for this call (the PPO data pass runs once per collection round)
    char* buf = nullptr;
    const size_t fN = sizeof(float) * (size_t)std::max<int64_t>(N, 1), fV = sizeof(float) * (size_t)(N + n_chains);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_off = take(sizeof(int64_t) * (n_chains + 1)), o_plp = take(fN), o_ilp = take(fN), o_val = take(fV), o_rew = take(fN), o_sta = take((size_t)std::max<int64_t>(N, 1)),
                 o_adv = take(fN), o_ret = take(fN), o_kls = take(fN), o_part = take(sizeof(double) * 2 * n_chains), o_pn = take(sizeof(int) * n_chains),
                 o_stats = take(2 * sizeof(float));
    LMRL_CUDA(cudaMalloc(&buf, off));
    auto fail_free = [&](cudaError_t err) { cudaFree(buf); return lmrl::fail("lmrl_ppo_advantages: %s", cudaGetErrorString(err)); };
    cudaError_t err = cudaSuccess;
    auto h2d = [&](size_t o, const void* src, size_t bytes) { if (err == cudaSuccess && bytes) err = cudaMemcpyAsync(buf + o, src, bytes, cudaMemcpyHostToDevice, st); };
    h2d(o_off, chain_off, sizeof(int64_t) * (n_chains + 1));
    h2d(o_plp, policy_logprobs, sizeof(float) * N); h2d(o_ilp, initial_logprobs, sizeof(float) * N); h2d(o_val, values, fV);
    h2d(o_rew, rewards, sizeof(float) * N); h2d(o_sta, should_take_action, (size_t)N);
    if (err != cudaSuccess) return fail_free(err);
    PpoArgs a{};
    a.n_chains = n_chains; a.chain_off = reinterpret_cast<const int64_t*>(buf + o_off);
    a.policy_lp = reinterpret_cast<const float*>(buf + o_plp); a.initial_lp = reinterpret_cast<const float*>(buf + o_ilp);
    a.values = reinterpret_cast<const float*>(buf + o_val); a.rewards = reinterpret_cast<const float*>(buf + o_rew);
    a.sta = reinterpret_cast<const uint8_t*>(buf + o_sta);
    a.gamma = (float)gamma; a.gamma_lam = (float)(gamma * lam); a.kl_weight = (float)kl_weight;   // python scalars are 'weak': cast once, then float32 arithmetic
    a.adv = reinterpret_cast<float*>(buf + o_adv); a.ret = reinterpret_cast<float*>(buf + o_ret); a.kls = reinterpret_cast<float*>(buf + o_kls);
    a.part = reinterpret_cast<double*>(buf + o_part); a.part_n = reinterpret_cast<int*>(buf + o_pn);
    ppo_gae_kernel<<<(n_chains + 127) / 128, 128, 0, st>>>(a);
    e->launches++;
    if (whiten && N > 0) {
        float* stats = reinterpret_cast<float*>(buf + o_stats);
        ppo_whiten_stats_kernel<<<1, 256, 0, st>>>(n_chains, a.part, a.part_n, stats);
        ppo_whiten_kernel<<<(int)std::min<int64_t>((N + 255) / 256, 148 * 8), 256, 0, st>>>(N, a.sta, stats, a.adv);
        e->launches += 2;
    }
    auto d2h = [&](void* dst, size_t o, size_t bytes) { if (err == cudaSuccess && bytes) err = cudaMemcpyAsync(dst, buf + o, bytes, cudaMemcpyDeviceToHost, st); };
    d2h(advantages, o_adv, sizeof(float) * N); d2h(returns, o_ret, sizeof(float) * N); d2h(kls, o_kls, sizeof(float) * N);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) return fail_free(err);
    cudaFree(buf);
    return 0;
}

int lmrl_ppo_loss_host(lmrl_engine* e, int64_t n, const float* mask, const float* logprobs, const float* values, const float* old_logprobs,
                       const float* old_values, const float* old_advantages, const float* old_returns, double cliprange_value, double cliprange,
                       double value_loss_coef, double* out /*[16]*/) {
    LMRL_CHECK(e && mask && logprobs && values && old_logprobs && old_values && old_advantages && old_returns && out, "lmrl_ppo_loss: null argument");
    LMRL_CHECK(n > 0, "lmrl_ppo_loss: empty batch");
    LMRL_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 4);
    const size_t fN = (sizeof(float) * (size_t)n + 255) / 256 * 256;
    char* buf = nullptr;
    LMRL_CUDA(cudaMalloc(&buf, 7 * fN + sizeof(double) * PPO_LOSS_TERMS * (blocks + 1)));
    const float* src[7] = {mask, logprobs, values, old_logprobs, old_values, old_advantages, old_returns};
    cudaError_t err = cudaSuccess;
    for (int k = 0; k < 7 && err == cudaSuccess; ++k) err = cudaMemcpyAsync(buf + k * fN, src[k], sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, st);
    PpoLossArgs a{};
    a.n = n;
    const float** dst[7] = {&a.mask, &a.logprobs, &a.values, &a.old_logprobs, &a.old_values, &a.old_adv, &a.old_ret};
    for (int k = 0; k < 7; ++k) *dst[k] = reinterpret_cast<const float*>(buf + k * fN);
    a.cliprange_value = (float)cliprange_value; a.ratio_lo = (float)(1.0 - cliprange); a.ratio_hi = (float)(1.0 + cliprange);
    a.part = reinterpret_cast<double*>(buf + 7 * fN);
    double* d_out = a.part + (size_t)blocks * PPO_LOSS_TERMS;
    double sums[PPO_LOSS_TERMS];
    if (err == cudaSuccess) {
        ppo_loss_partial_kernel<<<blocks, 256, 0, st>>>(a);
        ppo_loss_final_kernel<<<1, 32, 0, st>>>(blocks, a.part, d_out);
        e->launches += 2;
        err = cudaMemcpyAsync(sums, d_out, sizeof(sums), cudaMemcpyDeviceToHost, st);
    }
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    if (err == cudaSuccess) err = cudaGetLastError();
    cudaFree(buf);
    LMRL_CHECK(err == cudaSuccess, "lmrl_ppo_loss: %s", cudaGetErrorString(err));
    const double cnt = sums[0];
    const double vf_loss = 0.5 * sums[1] / cnt, pg_loss = sums[4] / cnt;
    out[0] = pg_loss + value_loss_coef * vf_loss;      // total_loss
    out[1] = pg_loss;                                          // policy_loss
    out[2] = vf_loss;                                          // value_loss
    out[3] = sums[6] / cnt;                                    // values_error
    out[4] = sums[2] / cnt;                                    // value clipfrac
    out[5] = sums[3] / cnt;                                    // approx_kl
    out[6] = sums[5] / cnt;                                    // policy clipfrac
    out[7] = sums[7] / cnt;                                    // ratio
    out[8] = cnt / sums[11];                                   // padding_percentage
    out[9] = sums[8] / cnt;                                    // values mean
    out[10] = sums[9] / cnt;                                   // old_values mean
    out[11] = sums[10] / cnt;                                  // returns mean
    out[12] = cnt;
    out[13] = out[14] = out[15] = 0.0;
    return 0;
}

}  // extern "C"

// ---- AdamW (the optimizer of the PPO / BC / ILQL drivers: optax.adamw(lr, b1 0.9, b2 0.95, eps 1e-8, weight_decay, mask), ------
// train_ppo_gpt2.py:144-165) as ONE streaming kernel over a flat fp32 parameter segment: reads p, g, m, v (16 B), writes p, m, v (12 B)
// per parameter — HBM-bound, 28 B / parameter.  optax order of operations: m, v moments; bias-corrected m_hat / (sqrt(v_hat) + eps);
// + weight_decay * p; * (-lr).
namespace lmrl {
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                             float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, float grad_scale) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * grad_scale;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float upd = __fdiv_rn(__fdiv_rn(mi, bc1), __fadd_rn(__fsqrt_rn(__fdiv_rn(vi, bc2)), eps)) + wd * p[i];      // IEEE division / sqrt whatever --use_fast_math says
        p[i] -= lr * upd;
    }
}
}  // namespace lmrl

extern "C" {

// All pointers are DEVICE pointers (fp32, n elements); enqueued on the engine stream, no synchronisation.  step is 1-based.
// grad_scale multiplies the gradient first (1 / world_size after a summing all-reduce, or a loss-scale inverse).
int lmrl_adamw_dev(lmrl_engine* e, float* params, const float* grads, float* m, float* v, int64_t n, double lr, double b1, double b2, double eps,
                   double weight_decay, int64_t step, double grad_scale) {
    LMRL_CHECK(e && params && grads && m && v, "lmrl_adamw: null argument");
    LMRL_CHECK(n >= 0 && step >= 1, "lmrl_adamw: n must be >= 0 and step >= 1");
    if (n == 0) return 0;
    LMRL_CUDA(cudaSetDevice(e->device));
    const float bc1 = (float)(1.0 - pow(b1, (double)step)), bc2 = (float)(1.0 - pow(b2, (double)step));
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    adamw_kernel<<<blocks, 256, 0, e->stream>>>(params, grads, m, v, n, (float)lr, (float)b1, (float)b2, (float)eps, (float)weight_decay, bc1, bc2, (float)grad_scale);
    e->launches++;
    LMRL_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
