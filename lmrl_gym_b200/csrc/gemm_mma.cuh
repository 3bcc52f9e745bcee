// gemm_mma.cuh — C[M,N] = A[M,K] · W[N,K]^T with fused epilogues, bf16 in / fp32 accumulate.
//
// First-generation tensor-core path of the decoder (mma.sync m16n8k16, cp.async 3-stage pipeline, XOR-swizzled
// shared tiles, ldmatrix fragments).  Persistent CTAs stride over 128x128 output tiles; the row count M is read
// from device memory so a whole rollout turn can be enqueued without the host knowing how many rows are live.
// The tcgen05/TMEM/TMA kernel in gemm_tcgen05.cuh replaces this one on the large shapes; this one stays as the
// cross-check in the GPU tests and for shapes the tcgen05 path does not cover.
//
// Epilogues:
//   EPI_QKV     + bias; Q -> q[M,d] bf16, K/V -> slab KV cache at (env, pos)               (c_attn + cache append)
//   EPI_RESID   + bias + residual (fp32, in place)                                          (attn c_proj, mlp c_proj)
//   EPI_GELU    + bias, gelu_new, -> bf16                                                   (mlp c_fc)
//   EPI_SAMPLE  logits never leave the SM: scale by 1/T, add Gumbel noise from a counter hash, per-row arg-max
//               merged across N tiles with a packed 64-bit atomicMax                        (tied LM head + sampler)
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <utility>
#include "pdl.cuh"

namespace lmrl {

enum { EPI_QKV = 0, EPI_RESID = 1, EPI_GELU = 2, EPI_SAMPLE = 3, EPI_RELU = 4, EPI_LOGPROB = 5 };   // RELU / LOGPROB: tcgen05 kernel only

struct GemmArgs {
    const __nv_bfloat16* A;      // [Mcap, K]
    const __nv_bfloat16* W;      // [N, K]
    const float* bias;           // [N] or nullptr
    int N, K;
    const int* M_dev;            // live rows
    // EPI_RESID
    float* resid;                // [M, N]
    // EPI_GELU / EPI_QKV(q)
    __nv_bfloat16* out;          // [M, N] (GELU) or [M, d] (QKV)
    // EPI_QKV
    __nv_bfloat16* kcache;       // this layer's page pool: [n_pages, H, 32, 64] (see KvPages in gpt2_kernels.cuh)
    __nv_bfloat16* vcache;
    const int* row_env;
    const int* row_pos;
    const int* kv_table;         // [B][kv_ppe] page of (env, position / 32)
    int d_model, n_head, kv_ppe;
    // EPI_SAMPLE
    unsigned long long* best;    // [rows] packed (ordered score << 32 | ~col)
    const uint64_t* row_key;     // [rows] per-row sampling key
    float inv_temp;
    int do_sample;
    int n_vocab;
    float* logits_out;           // optional [rows, n_vocab] fp32 dump (parity hook)
    // ---- tcgen05 kernel only ----
    // K segments: the mainloop runs n_seg segments per output tile; segment s multiplies kb_count 64-wide k-blocks of
    // A starting at k-block a_kb0 with W matrix w_idx (one of up to three [N, K_w] matrices, k-blocks from 0) into
    // accumulator acc.  An ordinary GEMM is one segment {0, 0, 0, K/64}.  The value-perturbed sampler
    // (value_rl_base/gpt2/generation.py:112-119) uses A = [h_pi | beta*relu1 | beta*relu2] and two accumulators:
    //   acc0 = h_pi·wte^T + beta*q1,  acc1 = h_pi·wte^T + beta*q2,  logit = min(acc0 + bias, acc1 + bias2).
    struct Seg { int acc, a_kb0, w_idx, kb_count; } seg[4];
    int n_seg, n_acc;
    const float* bias2;          // bias of accumulator 1 (EPI_SAMPLE with n_acc == 2)
    // EPI_LOGPROB (teacher-forced scoring): per row, per 128-column half tile, the running (max, sum exp(logit/T - max)) of the
    // masked logits, and the scaled logit of the row's target token; merged over tiles by logprob_finalize_kernel
    float2* lp_partial;          // [rows][2 * N / 256]
    const int* lp_target;        // [rows] target token id
    float* lp_tlogit;            // [rows] logit/T of the target
    int out_col0;                // EPI_RELU / EPI_GELU: column offset of the output box inside the tmOut tensor
    float out_scale;             // EPI_RELU: out = out_scale * relu(acc + bias)
    // EPI_RESID + fused LayerNorm (tcgen05 one-CTA kernel): the CTA whose tile completes an m-tile of the residual stream (every
    // N tile of those 128 rows has been reduce-added: per-m-tile arrival counter) normalises the 128 finished rows and writes the
    // bf16 operand of the next GEMM, so the decoder stack launches no LayerNorm kernels.  ln_out == nullptr: plain EPI_RESID.
    const float* ln_gamma;       // [N]
    const float* ln_beta;        // [N]
    __nv_bfloat16* ln_out;       // [M, N] bf16
    int* ln_cnt;                 // [ceil(Mcap / 128)] arrival counters, zero between launches
    float ln_eps;
    // LayerNorm folded into the CONSUMER GEMM (tcgen05 one-CTA kernel; Gpt2Dev::ln_fold).  LN(x)·W^T + b = rstd * (x · W''^T) + b' with
    // W''[n][k] = gamma[k] W[n][k] - mean_k(gamma[k] W[n][k]) (row-centred, so the mean of x drops out of the product) and
    // b'[n] = b[n] + sum_k beta[k] W[n][k]: the consumer reads the RAW residual stream as a bf16 copy and only needs 1/sigma per row.
    //   producer (EPI_RESID with xb_out): reads x, writes x + acc + bias (fp32) and its bf16 copy, and the per-row partial
    //            (sum, sum of squares) of its columns into stat_out[row][stat_slot0 + 2 * n_tile + half] — fixed slots, no atomics: deterministic
    //   consumer (EPI_QKV / EPI_GELU with stat_in): rstd from the stat_n partial pairs of the row, epilogue value = rstd * acc + bias'
    __nv_bfloat16* xb_out;       // [M, N] bf16 copy of the updated residual stream
    float2* stat_out;            // [M][8] partial (sum, sumsq)
    const float2* stat_in;       // [M][8]
    int stat_n;                  // partial pairs per row to add up (consumer)
};

constexpr int GEMM_BM = 128, GEMM_BN = 128, GEMM_BK = 64, GEMM_STAGES = 3, GEMM_THREADS = 256;
constexpr int GEMM_STAGE_BYTES = (GEMM_BM + GEMM_BN) * GEMM_BK * 2;
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

__device__ __forceinline__ void mma_bf16(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float gelu_new(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    return 0.5f * x * (1.0f + tanhf(k0 * (x + k1 * x * x * x)));
}

// ---- counter-based sampling noise (also restated in oracle/gpt2.py::gumbel_noise) -------------------------------
__device__ __host__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
// two multiply-fold rounds (Philox-style 32x32 -> 64 products): 2 IMAD.WIDE + 3 LOP3 per draw on the sampling epilogue
__device__ __host__ __forceinline__ uint32_t noise_bits(uint64_t key, uint32_t col) {
    uint64_t p = (uint64_t)(col ^ (uint32_t)key) * 0xD2511F53ull;
    const uint32_t x = (uint32_t)(p >> 32) ^ (uint32_t)p ^ (uint32_t)(key >> 32);
    p = (uint64_t)x * 0xCD9E8D57ull;
    return (uint32_t)(p >> 32) ^ (uint32_t)p;
}
// Gumbel noise g = -log(-log(u)), u = (k + 0.5) / 2^24 from the top 24 hash bits.  Round 1 formed u in float32 and took two fast logs:
// for k = 2^24 - 1 the float rounds to exactly 1.0 (noise = +inf: that column wins whatever its logit, probability 2^-24 per draw) and
// near 1 the MUFU.LG2 error (~1e-7 absolute) is as large as -log(u) itself, so the upper tail of the noise (> ~14) — exactly the draws
// that let an unlikely token win — was wrong: ~0.5 % of sampled tokens disagreed with the float64 oracle where 0.05 % of low-probability
// winners are expected.  Now the inner logarithm switches to the series of -log1p(-v), v = 1 - u = (2^24 - k - 0.5) / 2^24 (exact in
// float32) for v < 2^-6; relative error of -log(u) <= 1e-6 everywhere, noise error <= 2e-6.
__device__ __forceinline__ float gumbel_from_bits(uint32_t h) {
    const uint32_t k = h >> 8;
    const float u = fmaf((float)k, 1.0f / 16777216.0f, 0.5f / 16777216.0f);
    const float v = fmaf((float)(16777215u - k), 1.0f / 16777216.0f, 0.5f / 16777216.0f);
    const float series = v * fmaf(v, fmaf(v, 0.33333334f, 0.5f), 1.0f);     // v + v^2 / 2 + v^3 / 3
    const float e = v < 0.015625f ? series : -__logf(u);
    return -__logf(e);
}
// order-preserving float -> uint32
__device__ __host__ __forceinline__ uint32_t float_ordered(float f) {
    uint32_t b;
#ifdef __CUDA_ARCH__
    b = __float_as_uint(f);
#else
    memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ unsigned long long pack_best(float score, int col) {
    return ((unsigned long long)float_ordered(score) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)col);
}

template <int EPI>
__device__ __forceinline__ void epi_store(const GemmArgs& g, int row, int col, float v0, float v1) {
    if (g.bias) { v0 += __ldg(g.bias + col); v1 += __ldg(g.bias + col + 1); }
    if constexpr (EPI == EPI_RESID) {
        float2* p = reinterpret_cast<float2*>(g.resid + (size_t)row * g.N + col);
        float2 r = *p;
        r.x += v0; r.y += v1;
        *p = r;
    } else if constexpr (EPI == EPI_GELU) {
        *reinterpret_cast<__nv_bfloat162*>(g.out + (size_t)row * g.N + col) = __floats2bfloat162_rn(gelu_new(v0), gelu_new(v1));
    } else if constexpr (EPI == EPI_QKV) {
        const int d = g.d_model;
        const int part = col / d, c = col - part * d;
        const __nv_bfloat162 v = __floats2bfloat162_rn(v0, v1);
        if (part == 0) {
            *reinterpret_cast<__nv_bfloat162*>(g.out + (size_t)row * d + c) = v;
        } else {
            const int e = g.row_env[row], p = g.row_pos[row];
            const int head = c >> 6, dd = c & 63;
            __nv_bfloat16* cache = part == 1 ? g.kcache : g.vcache;
            const int page = g.kv_table[(size_t)e * g.kv_ppe + (p >> 5)];
            *reinterpret_cast<__nv_bfloat162*>(cache + (((size_t)page * g.n_head + head) * 32 + (p & 31)) * 64 + dd) = v;
        }
    }
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_bf16_mma_kernel(const GemmArgs g) {
    extern __shared__ __align__(128) unsigned char gemm_smem[];
    pdl_enter();
    const int M = *g.M_dev;
    if (M <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_m = warp >> 2, warp_n = warp & 3;
    const int n_tiles = g.N / GEMM_BN, m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
    const int total = n_tiles * m_tiles, kt = g.K / GEMM_BK;
    const uint32_t smem_base = smem_u32(gemm_smem);

    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * GEMM_BM, n0 = (tile % n_tiles) * GEMM_BN;
        const __nv_bfloat16* Ag = g.A + (size_t)m0 * g.K;
        const __nv_bfloat16* Wg = g.W + (size_t)n0 * g.K;

        auto load_stage = [&](int stage, int kb) {
            const uint32_t sA = smem_base + stage * GEMM_STAGE_BYTES, sB = sA + GEMM_BM * GEMM_BK * 2;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = tid + i * GEMM_THREADS, row = c >> 3, ch = c & 7;
                const uint32_t off = row * 128 + ((ch ^ (row & 7)) << 4);
                cp_async16(sA + off, Ag + (size_t)row * g.K + kb * GEMM_BK + ch * 8);
                cp_async16(sB + off, Wg + (size_t)row * g.K + kb * GEMM_BK + ch * 8);
            }
        };

        float acc[4][4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll
        for (int s = 0; s < GEMM_STAGES - 1; ++s) {
            if (s < kt) load_stage(s, s);
            cp_async_commit();
        }
        for (int kb = 0; kb < kt; ++kb) {
            cp_async_wait<GEMM_STAGES - 2>();
            __syncthreads();
            if (kb + GEMM_STAGES - 1 < kt) load_stage((kb + GEMM_STAGES - 1) % GEMM_STAGES, kb + GEMM_STAGES - 1);
            cp_async_commit();
            const uint32_t sA = smem_base + (kb % GEMM_STAGES) * GEMM_STAGE_BYTES, sB = sA + GEMM_BM * GEMM_BK * 2;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t a[4][4], b[4][2];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    const int row = warp_m * 64 + mi * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                    const int ch = ks * 2 + (lane >> 4);
                    ldmatrix_x4(sA + row * 128 + ((ch ^ (row & 7)) << 4), a[mi][0], a[mi][1], a[mi][2], a[mi][3]);
                }
#pragma unroll
                for (int nj = 0; nj < 2; ++nj) {
                    const int row = warp_n * 32 + nj * 16 + (lane & 7) + (lane >> 4) * 8;
                    const int ch = ks * 2 + ((lane >> 3) & 1);
                    ldmatrix_x4(sB + row * 128 + ((ch ^ (row & 7)) << 4), b[nj * 2][0], b[nj * 2][1], b[nj * 2 + 1][0], b[nj * 2 + 1][1]);
                }
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) mma_bf16(acc[mi][ni], a[mi], b[ni][0], b[ni][1]);
            }
        }
        cp_async_wait<0>();
        __syncthreads();   // every warp is done with shared memory before the next tile's prologue

        const int gq = lane >> 2, tq = lane & 3;
        if constexpr (EPI != EPI_SAMPLE) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const int row = m0 + warp_m * 64 + mi * 16 + gq, col = n0 + warp_n * 32 + ni * 8 + tq * 2;
                    if (row < M) epi_store<EPI>(g, row, col, acc[mi][ni][0], acc[mi][ni][1]);
                    if (row + 8 < M) epi_store<EPI>(g, row + 8, col, acc[mi][ni][2], acc[mi][ni][3]);
                }
        } else {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row = m0 + warp_m * 64 + mi * 16 + gq + h * 8;
                    const bool live = row < M;
                    const uint64_t key = (live && g.do_sample) ? g.row_key[row] : 0ull;
                    float best = -INFINITY;
                    int best_col = 0x7FFFFFFF;
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int col = n0 + warp_n * 32 + ni * 8 + tq * 2 + q;
                            const float logit = acc[mi][ni][h * 2 + q];
                            if (live && col < g.n_vocab) {
                                if (g.logits_out) g.logits_out[(size_t)row * g.n_vocab + col] = logit;
                                float s = logit * g.inv_temp;
                                if (g.do_sample) s += gumbel_from_bits(noise_bits(key, (uint32_t)col));
                                if (s > best) { best = s; best_col = col; }
                            }
                        }
#pragma unroll
                    for (int off = 1; off <= 2; off <<= 1) {
                        const float ob = __shfl_xor_sync(0xFFFFFFFFu, best, off);
                        const int oc = __shfl_xor_sync(0xFFFFFFFFu, best_col, off);
                        if (ob > best || (ob == best && oc < best_col)) { best = ob; best_col = oc; }
                    }
                    if (live && tq == 0 && best_col != 0x7FFFFFFF) atomicMax(g.best + row, pack_best(best, best_col));
                }
        }
    }
}

}  // namespace lmrl
