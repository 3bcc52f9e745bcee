// gemv.cuh — weight-streaming GEMV path of the decoder for SMALL batches (decode passes of at most 8 rows).
//
// The reference's own drivers evaluate with 1-32 concurrent episodes (wordle/bc/eval_bc_gpt2.py:38-39, maze/bc/eval_bc.py:164-173;
// BASELINE.json configs[0] is batch 1).  At that size a decode pass is pure weight streaming: Y[R, N] = X[R, K] . W[N, K]^T reads
// 2*N*K bytes of W for 2*R*N*K flops, and 128-row tensor-core tiles would be >= 94 % padding.  Here
//   * the R <= 8 activation rows are staged ONCE per CTA in shared memory as fp32 — and when the operand is a LayerNorm of the
//     fp32 residual stream (ln_1 -> QKV, ln_2 -> FC, ln_f -> LM head) the CTA normalises the rows itself while staging them
//     (values rounded to bf16 exactly like the tensor-core path's operand), so the small-batch decoder launches no LayerNorm;
//   * a CTA (4 warps) takes one group of 4 consecutive W rows (output features) at a time and splits K across its warps, so that ALL of
//     the group's 16-byte coalesced loads are in flight at once (one memory round trip per group instead of K / 2048 dependent ones:
//     at this size the kernel is a latency chain, not a bandwidth problem); 4 x R fp32 accumulators per lane, each shared-memory
//     read of X feeding 4 columns; the four warps' partial sums meet in shared memory;
//   * the 4R partial sums are reduced across the warp with a halving butterfly (4R - 1 + log2(32 / 4R) shuffles instead of
//     5 * 4R) that leaves ONE finished (column, row) element in each lane, which then runs the epilogue for its element:
//     bias + KV-page append / q (QKV), fp32 residual add, GeLU -> bf16, or the fused temperature + Gumbel arg-max of the LM head.
// HBM-bound by construction: algorithmic bytes per launch = 2*N*K (W) + O(R*(K + N)); bench.py reports the achieved GB/s.
#pragma once
#include "gpt2_kernels.cuh"

namespace lmrl {

constexpr int GV_WARPS = 4;                     // warps per CTA (small CTAs: a 768 x 768 matrix still spreads over 48 SMs)
constexpr int GV_COLS = 4;                      // output features per warp pass
constexpr int GV_MAX_ROWS = 8;

struct GemvArgs {
    // operand: either LayerNorm(x_f32[gather ? gather[r] : r]) or A[r] (bf16)
    const float* x_f32; const float* ln_gamma; const float* ln_beta; float ln_eps; const int* gather;
    const __nv_bfloat16* A; int a_pitch;
    const __nv_bfloat16* W;      // [N][K]
    const float* bias;           // [N] or nullptr
    int N, K;
    const int* M_dev;            // live rows (<= R)
    // EPI_RESID
    float* resid;                // [M][N]
    // EPI_GELU / EPI_QKV (q)
    __nv_bfloat16* out;          // [M][N] (GELU) or [M][d] (QKV)
    // EPI_QKV
    __nv_bfloat16* kcache; __nv_bfloat16* vcache; const int* row_env; const int* row_pos; const int* kv_table; int d_model, n_head, kv_ppe;
    // EPI_SAMPLE
    unsigned long long* best; const uint64_t* row_key; float inv_temp; int do_sample; int n_vocab; float* logits_out;
};

// halving butterfly: on entry every lane holds NV partial sums v[0..NV); on exit lane L holds in v[0] the warp total of the value
// with index sum_s ((L >> (4 - s)) & 1) << s  (s = 0 .. log2(NV) - 1), replicated over the remaining low lane bits
template <int NV>
__device__ __forceinline__ void warp_reduce_many(float (&v)[NV], int lane) {
    int off = 16;
#pragma unroll
    for (int n = NV; n > 1; n >>= 1, off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float keep = hi ? v[2 * i + 1] : v[2 * i], send = hi ? v[2 * i] : v[2 * i + 1];
            v[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, off);
        }
    }
    for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xFFFFFFFFu, v[0], off);
}
template <int NV>
__device__ __forceinline__ int warp_reduce_many_index(int lane) {
    int idx = 0, bit = 4;
#pragma unroll
    for (int s = 0; (1 << s) < NV; ++s, --bit) idx |= ((lane >> bit) & 1) << s;
    return idx;
}

template <int EPI, int R, bool LN>
__global__ void __launch_bounds__(GV_WARPS * 32) gemv_kernel(const GemvArgs a) {
    // [R][2][K / 8] float4: fp32 copies of the bf16-valued activations; plane h holds elements 8c + 4h .. 8c + 4h + 3 of every 8-element
    // chunk c, so that the 32 lanes of a warp (consecutive chunks) read consecutive float4 from either plane: no bank conflicts
    extern __shared__ __align__(16) float gv_x[];
    pdl_trigger();                                                // (programmatic dependent launch: the next kernel may begin its own prologue)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = a.K;
    // the weights do not depend on the previous kernel's output: pull this CTA's first group of W rows towards L2 while the
    // activation rows are being staged
    if ((int)blockIdx.x < a.N / GV_COLS) {
        const char* w0 = reinterpret_cast<const char*>(a.W + (size_t)blockIdx.x * GV_COLS * K);
        for (int off = (int)threadIdx.x * 128; off < GV_COLS * K * 2; off += GV_WARPS * 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(w0 + off));
    }
    pdl_wait();                                                   // everything below reads what the previous kernel wrote
    const int M = min(*a.M_dev, R);
    if (M <= 0) return;
    // ---- stage the activation rows (zero rows beyond M) ----
    if constexpr (LN) {
        for (int r = warp; r < R; r += GV_WARPS) {
            float4* dst = reinterpret_cast<float4*>(gv_x + (size_t)r * K);
            const int nv = K >> 7;                                 // float4 per lane (K = d_model, multiple of 128, <= 1024)
            if (r >= M) {
                for (int i = 0; i < nv; ++i) dst[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
                continue;                                          // (all of the row's 2 x K / 8 float4, whatever the plane order)
            }
            const float4* src = reinterpret_cast<const float4*>(a.x_f32 + (size_t)(a.gather ? a.gather[r] : r) * K);
            float4 v[8];
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < nv) { v[i] = src[lane + 32 * i]; s += v[i].x + v[i].y + v[i].z + v[i].w; }
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
            const float mean = s / K;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < nv) {
                    const float b0 = v[i].x - mean, b1 = v[i].y - mean, b2 = v[i].z - mean, b3 = v[i].w - mean;
                    q += b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3;
                }
#pragma unroll
            for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
            const float rstd = rsqrtf(q / K + a.ln_eps);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < nv) {
                    const int c = (lane + 32 * i) * 4;
                    const float4 g4 = *reinterpret_cast<const float4*>(a.ln_gamma + c), b4 = *reinterpret_cast<const float4*>(a.ln_beta + c);
                    float4 o4;                                     // rounded to bf16 like the tensor-core path's LayerNorm output
                    o4.x = __bfloat162float(__float2bfloat16((v[i].x - mean) * rstd * g4.x + b4.x));
                    o4.y = __bfloat162float(__float2bfloat16((v[i].y - mean) * rstd * g4.y + b4.y));
                    o4.z = __bfloat162float(__float2bfloat16((v[i].z - mean) * rstd * g4.z + b4.z));
                    o4.w = __bfloat162float(__float2bfloat16((v[i].w - mean) * rstd * g4.w + b4.w));
                    const int j = lane + 32 * i;                    // float4 j = elements 4j .. 4j + 3 = half (j & 1) of chunk j / 2
                    dst[(j & 1) * (K >> 3) + (j >> 1)] = o4;
                }
        }
    } else {
        const int chunks = K >> 3;                                 // 8 bf16 per 16-byte chunk
        for (int i = threadIdx.x; i < R * chunks; i += GV_WARPS * 32) {
            const int r = i / chunks, c = i - r * chunks;
            float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (r < M) bf16x8_to_float(*reinterpret_cast<const uint4*>(a.A + (size_t)r * a.a_pitch + 8 * c), f);
            float4* dst = reinterpret_cast<float4*>(gv_x + (size_t)r * K);
            dst[c] = make_float4(f[0], f[1], f[2], f[3]);
            dst[chunks + c] = make_float4(f[4], f[5], f[6], f[7]);
        }
    }
    __syncthreads();
    // ---- stream W: one group of GV_COLS consecutive output features per CTA pass, K split over the warps ----
    constexpr int NV = GV_COLS * R;
    constexpr int KSTEPS = 4;                                      // K <= 4096: at most 512 chunks = 4 per lane of the CTA
    __shared__ float gv_red[2][GV_WARPS][32];
    const int my_idx = warp_reduce_many_index<NV>(lane);           // the (column, row) element this lane finishes
    const int my_c = my_idx / R, my_r = my_idx - my_c * R;
    const bool writer = warp == 0 && (lane & ((32 / NV) - 1)) == 0 && my_r < M;   // one lane per element (NV < 32: low lane bits replicate)
    const int n_groups = a.N / GV_COLS, kchunks = K >> 3;
    float best_s = -INFINITY; int best_c = 0x7FFFFFFF;             // EPI_SAMPLE: running arg-max of this lane's row over its columns (warp 0)
    int it = 0;
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x, ++it) {
        const int n0 = grp * GV_COLS;
        const uint4* w4 = reinterpret_cast<const uint4*>(a.W + (size_t)n0 * K);
        uint4 wv[KSTEPS][GV_COLS];
#pragma unroll
        for (int st = 0; st < KSTEPS; ++st) {                       // every load of the group first
            const int ch = (st * GV_WARPS + warp) * 32 + lane;
#pragma unroll
            for (int c = 0; c < GV_COLS; ++c) wv[st][c] = ch < kchunks ? __ldg(w4 + (size_t)c * kchunks + ch) : make_uint4(0, 0, 0, 0);
        }
        float acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.f;
#pragma unroll
        for (int st = 0; st < KSTEPS; ++st) {
            const int ch = (st * GV_WARPS + warp) * 32 + lane;
            if (ch < kchunks) {
                float xf[R][8];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float4* xs = reinterpret_cast<const float4*>(gv_x + (size_t)r * K);
                    const float4 x0 = xs[ch], x1 = xs[kchunks + ch];
                    xf[r][0] = x0.x; xf[r][1] = x0.y; xf[r][2] = x0.z; xf[r][3] = x0.w; xf[r][4] = x1.x; xf[r][5] = x1.y; xf[r][6] = x1.z; xf[r][7] = x1.w;
                }
#pragma unroll
                for (int c = 0; c < GV_COLS; ++c) {
                    float wf[8];
                    bf16x8_to_float(wv[st][c], wf);
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int t = 0; t < 8; ++t) acc[c * R + r] = fmaf(wf[t], xf[r][t], acc[c * R + r]);
                }
            }
        }
        const int n = n0 + my_c;
        const float bias_n = (warp == 0 && a.bias) ? __ldg(a.bias + n) : 0.f;     // issued before the shuffle tree, consumed after it
        warp_reduce_many<NV>(acc, lane);
        gv_red[it & 1][warp][lane] = acc[0];                         // the warps' K-partials of element (my_c, my_r) meet here
        __syncthreads();                                            // (double buffered: one barrier per group)
        if (warp != 0) continue;
        float val = bias_n;
#pragma unroll
        for (int w = 0; w < GV_WARPS; ++w) val += gv_red[it & 1][w][lane];
        if constexpr (EPI == EPI_SAMPLE) {
            if (my_r < M && n < a.n_vocab) {
                if (a.logits_out && writer) a.logits_out[(size_t)my_r * a.n_vocab + n] = val;
                float s = val * a.inv_temp;
                if (a.do_sample) s += gumbel_from_bits(noise_bits(a.row_key[my_r], (uint32_t)n));
                if (s > best_s) { best_s = s; best_c = n; }       // columns visited in increasing order: the lowest column wins ties
            }
        } else if (writer) {
            if constexpr (EPI == EPI_RESID) {
                a.resid[(size_t)my_r * a.N + n] += val;
            } else if constexpr (EPI == EPI_GELU) {
                a.out[(size_t)my_r * a.N + n] = __float2bfloat16(gelu_new(val));
            } else if constexpr (EPI == EPI_QKV) {
                const int d = a.d_model, part = n / d, c = n - part * d;
                if (part == 0) {
                    a.out[(size_t)my_r * d + c] = __float2bfloat16(val);
                } else {
                    const int e = a.row_env[my_r], p = a.row_pos[my_r];
                    const int page = a.kv_table[(size_t)e * a.kv_ppe + (p >> KV_PAGE_SHIFT)];
                    (part == 1 ? a.kcache : a.vcache)[kv_row_off(page, a.n_head, c >> 6, p) + (c & 63)] = __float2bfloat16(val);
                }
            }
        }
    }
    if constexpr (EPI == EPI_SAMPLE) {
        // lanes of one row differ in the column bits of the index (the lane bits that select my_c) and in the replicated low bits: fold them
        // all — lanes with another row keep out by comparing the row id
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float os = __shfl_xor_sync(0xFFFFFFFFu, best_s, o);
            const int oc = __shfl_xor_sync(0xFFFFFFFFu, best_c, o), orow = __shfl_xor_sync(0xFFFFFFFFu, my_r, o);
            if (orow == my_r && (os > best_s || (os == best_s && oc < best_c))) { best_s = os; best_c = oc; }
        }
        if (warp == 0 && my_r < M && best_c != 0x7FFFFFFF) atomicMax(a.best + my_r, pack_best(best_s, best_c));
    }
}

}  // namespace lmrl
