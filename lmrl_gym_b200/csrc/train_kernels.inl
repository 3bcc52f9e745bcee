// train_kernels.inl — hand-written backward pass of the GPT-2 policy + value head for the PPO gradient step (C ABI part 5).
//
// Reference: GPT2PPOTrain._step (LLM_RL/algorithms/ppo/gpt2/interface.py:72-211): jax.value_and_grad of
//   loss(policy_params, value_head_params) = ppo_loss_fn(logprobs(policy(input_ids)), values(LinearHead(hidden[-1])), ...)
// Here the chain rule is split at the per-token quantities: the caller supplies dL/dlogprob[t] and dL/dvalue[t] (the derivative of
// ppo_loss_fn, a closed-form elementwise expression — lmrl_gym_b200/ppo_train.py::ppo_loss_grads) and this file propagates them
// through the LM head, the value head and the decoder stack into fp32 parameter gradients (HF GPT-2 layout, accumulated):
//
//   forward  : the engine's own prefill pass (gpt2_engine.inl) with the residual stream saved before every layer;
//   backward : per layer, the layer is recomputed from its saved input (activation checkpointing: 37 KB per row for 12 layers instead
//              of ~290 KB), then   dgrad  dX = dY · W      -> gemm_bf16_tcgen05 (A = dY, W operand = the HF-layout copy of the weight)
//                                  wgrad  dW = X^T · dY    -> gemm_bf16_tcgen05 over TRANSPOSED operands, fp32 accumulation straight into
//                                                             the gradient by the TMA reduce-add epilogue
//              plus small HBM-bound kernels: LayerNorm / GeLU / causal-attention backward, column sums for the biases, transposes.
// Operands are bf16, accumulation fp32 (the usual mixed-precision recipe; the reference trains in fp32 or bf16 activations).
#include <cuda_bf16.h>

namespace lmrl {

// ---- small kernels --------------------------------------------------------------------------------------------------------
// src [M][C] (bf16, or fp32 converted on the fly) -> dst [C][Mp] bf16, columns >= M zero-filled (Mp = row count padded to 64)
template <typename T>
__global__ void transpose_to_bf16_rows_kernel(const T* __restrict__ src, const int* __restrict__ M_dev, int C, __nv_bfloat16* __restrict__ dst, int Mp,
                                              __nv_bfloat16* __restrict__ copy /* optional row-major bf16 copy [M][C] */) {
    __shared__ float tile[32][33];
    const int M = *M_dev;
    const int m0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int m = m0 + i, c = c0 + threadIdx.x;
        float v = 0.f;
        if (m < M && c < C) {
            if constexpr (sizeof(T) == 4) v = src[(size_t)m * C + c]; else v = __bfloat162float(src[(size_t)m * C + c]);
            if (copy) copy[(size_t)m * C + c] = __float2bfloat16(v);
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, m = m0 + threadIdx.x;
        if (c < C && m < Mp) dst[(size_t)c * Mp + m] = __float2bfloat16(tile[threadIdx.x][i]);
    }
}

// out[c] += sum over rows of src[m][c].  Launched with enough row groups (gridDim.y) that a thread walks only a few rows, four loads in flight.
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ src, const int* __restrict__ M_dev, int C, float* __restrict__ out) {
    const int M = *M_dev;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    auto ld = [&](int m) -> float {
        if constexpr (sizeof(T) == 4) return src[(size_t)m * C + c]; else return __bfloat162float(src[(size_t)m * C + c]);
    };
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const int step = gridDim.y;
    int m = blockIdx.y;
    for (; m + 3 * step < M; m += 4 * step) { a0 += ld(m); a1 += ld(m + step); a2 += ld(m + 2 * step); a3 += ld(m + 3 * step); }
    for (; m < M; m += step) a0 += ld(m);
    const float acc = (a0 + a1) + (a2 + a3);
    if (acc != 0.f) atomicAdd(out + c, acc);
}

__device__ __forceinline__ float gelu_new_grad(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float t = tanhf(k0 * (x + k1 * x * x * x));
    return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k0 * (1.f + 3.f * k1 * x * x);
}
// h = gelu_new(hpre)   (forward recompute: the FC GEMM stores the pre-activation, this makes the FC2 operand); 16-byte accesses, C % 8 == 0
__global__ void gelu_fwd_kernel(const __nv_bfloat16* __restrict__ hpre, const int* __restrict__ M_dev, int C, __nv_bfloat16* __restrict__ h) {
    const size_t n8 = ((size_t)*M_dev * C) >> 3;
    const uint4* src = reinterpret_cast<const uint4*>(hpre);
    uint4* dst = reinterpret_cast<uint4*>(h);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        uint4 u = src[i];
        __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(p[k]); p[k] = __floats2bfloat162_rn(gelu_new(f.x), gelu_new(f.y)); }
        dst[i] = u;
    }
}
// dhpre = dh * gelu'(hpre), in place over dh
__global__ void gelu_bwd_kernel(const __nv_bfloat16* __restrict__ hpre, const int* __restrict__ M_dev, int C, __nv_bfloat16* __restrict__ dh) {
    const size_t n8 = ((size_t)*M_dev * C) >> 3;
    const uint4* src = reinterpret_cast<const uint4*>(hpre);
    uint4* dst = reinterpret_cast<uint4*>(dh);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 x = src[i];
        uint4 g = dst[i];
        const __nv_bfloat162* px = reinterpret_cast<const __nv_bfloat162*>(&x);
        __nv_bfloat162* pg = reinterpret_cast<__nv_bfloat162*>(&g);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 fx = __bfloat1622float2(px[k]), fg = __bfloat1622float2(pg[k]);
            pg[k] = __floats2bfloat162_rn(fg.x * gelu_new_grad(fx.x), fg.y * gelu_new_grad(fx.y));
        }
        dst[i] = g;
    }
}

// LayerNorm backward, one warp per row:  y = (x - mean) * rstd * g + b
//   dx (+)= rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat));   dg += dy * xhat;   db += dy
// A lane owns the columns lane, lane + 32, ...: its dg / db partial sums stay in registers over all the rows its warp visits and are
// folded into the CTA's shared-memory row, then into global memory, once at the end.  d <= 1024.
template <typename TDY>
__global__ void ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const TDY* __restrict__ dy, int d, float eps,
                              const int* __restrict__ M_dev, float* __restrict__ dx, int accumulate, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    extern __shared__ float ln_sm[];            // [2][d]
    const int M = *M_dev;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) ln_sm[c] = 0.f;
    __syncthreads();
    float dgr[32], dbr[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) { dgr[k] = 0.f; dbr[k] = 0.f; }
    for (int r = blockIdx.x * nwarp + warp; r < M; r += gridDim.x * nwarp) {
        const float* xr = x + (size_t)r * d;
        const TDY* dr = dy + (size_t)r * d;
        float xv[32], dv[32];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const int c = lane + 32 * k;
            xv[k] = 0.f; dv[k] = 0.f;
            if (c < d) {
                xv[k] = xr[c];
                if constexpr (sizeof(TDY) == 4) dv[k] = dr[c]; else dv[k] = __bfloat162float(dr[c]);
                s += xv[k];
            }
        }
        const float mean = warp_sum(s) / d;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) if (lane + 32 * k < d) { const float t = xv[k] - mean; q += t * t; }
        const float rstd = rsqrtf(warp_sum(q) / d + eps);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const int c = lane + 32 * k;
            if (c < d) {
                const float xh = (xv[k] - mean) * rstd, dg = dv[k] * gamma[c];
                xv[k] = xh;
                s1 += dg; s2 += dg * xh;
                dgr[k] += dv[k] * xh; dbr[k] += dv[k];
            }
        }
        s1 = warp_sum(s1) / d; s2 = warp_sum(s2) / d;
        float* o = dx + (size_t)r * d;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const int c = lane + 32 * k;
            if (c < d) {
                const float v = rstd * (dv[k] * gamma[c] - s1 - xv[k] * s2);
                o[c] = accumulate ? o[c] + v : v;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const int c = lane + 32 * k;
        if (c < d) { atomicAdd(ln_sm + c, dgr[k]); atomicAdd(ln_sm + d + c, dbr[k]); }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        if (ln_sm[c] != 0.f) atomicAdd(dgamma + c, ln_sm[c]);
        if (ln_sm[d + c] != 0.f) atomicAdd(dbeta + c, ln_sm[d + c]);
    }
}

// causal self-attention backward for full sequences (every slot starts at position 0), one CTA per (slot, head).
//   P = softmax(scale * Q K^T) (causal);  O = P V;   given dO:  dV = P^T dO,  dP = dO V^T,  dS = P * (dP - rowsum(P * dP)) * scale,
//   dQ = dS K,  dK = dS^T Q.
// The Q, K, V, dO tiles of the (sequence, head) are staged in shared memory as bf16 with a 144-byte row pitch, read with 16-byte loads (conflict-free for lanes
// that walk rows).  Phase 1: one warp per query row, lanes over the keys j <= i: scores, soft-max statistics, dP, D_i and
// dQ_i (64 fp32 accumulators per lane, reduced with shuffles); the row statistics stay in shared memory.  Phase 2: one warp per key
// row, lanes over the queries i >= j: dK_j and dV_j (128 accumulators per lane).  No atomics, no tensor cores: attention is ~3 % of the
// training FLOPs at these sequence lengths (T <= 384).
constexpr int ATTB_THREADS = 256;
constexpr int ATTB_PITCH = 72;                     // bf16 elements per staged row: 64 + 8 padding -> 144 B, 16-byte aligned, and the 8 lanes of a
                                                   // quarter-warp reading 16 B each from consecutive rows cover all 32 banks (36 j + 4 c mod 32)
constexpr int ATTB_MAX_T = 384;
struct Bf8 { uint4 v; };                           // eight bf16
__device__ __forceinline__ void bf8_unpack(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(p[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
// dot of a register-resident row (8 x uint4) with a shared-memory row
__device__ __forceinline__ float attb_dot_rs(const uint4 (&a)[8], const __nv_bfloat16* b) {
    const uint4* pb = reinterpret_cast<const uint4*>(b);
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};                 // eight independent chains: the warp count per SM is low, ILP hides the FMA latency
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float x[8], y[8];
        bf8_unpack(a[c], x); bf8_unpack(pb[c], y);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = fmaf(x[i], y[i], s[i]);
    }
    return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}
__global__ void __launch_bounds__(ATTB_THREADS, 1) attn_bwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kcache,
                                                                   const __nv_bfloat16* __restrict__ vcache, const __nv_bfloat16* __restrict__ dout, RowsDev rows,
                                                                   int n_head, int d_model, KvPages kv, float scale, __nv_bfloat16* __restrict__ dqkv) {
    extern __shared__ __align__(16) unsigned char ab_raw[];
    const int slot = blockIdx.x / n_head, head = blockIdx.x % n_head;
    if (slot >= *rows.n_slots) return;
    const int env = rows.slot_env[slot], r0 = rows.slot_row0[slot], T = rows.slot_nrow[slot];
    __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(ab_raw);
    __nv_bfloat16* Ks = Qs + (size_t)T * ATTB_PITCH;
    __nv_bfloat16* Vs = Ks + (size_t)T * ATTB_PITCH;
    __nv_bfloat16* Os = Vs + (size_t)T * ATTB_PITCH;
    float* st_m = reinterpret_cast<float*>(Os + (size_t)T * ATTB_PITCH);          // T * 144 B per tile: 16-byte aligned
    float* st_l = st_m + T;
    float* st_D = st_l + T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = ATTB_THREADS >> 5;
    const int* ptab = kv.table + (size_t)env * kv.ppe;                       // position r of the sequence: row r % 32 of page ptab[r / 32]
    auto kv_off = [&](int r) { return kv_row_off(ptab[r >> KV_PAGE_SHIFT], n_head, head, r); };
    for (int i = tid; i < T * 8; i += ATTB_THREADS) {                        // one 16-byte chunk per thread per step
        const int r = i >> 3, w = i & 7;
        const size_t grow = (size_t)(r0 + r) * d_model + head * 64 + 8 * w;
        reinterpret_cast<uint4*>(Qs + (size_t)r * ATTB_PITCH)[w] = *reinterpret_cast<const uint4*>(q + grow);
        reinterpret_cast<uint4*>(Os + (size_t)r * ATTB_PITCH)[w] = *reinterpret_cast<const uint4*>(dout + grow);
        reinterpret_cast<uint4*>(Ks + (size_t)r * ATTB_PITCH)[w] = *reinterpret_cast<const uint4*>(kcache + kv_off(r) + 8 * w);
        reinterpret_cast<uint4*>(Vs + (size_t)r * ATTB_PITCH)[w] = *reinterpret_cast<const uint4*>(vcache + kv_off(r) + 8 * w);
    }
    __syncthreads();
    // ---- phase 1: dQ and the row statistics (warp per query row, lanes over keys)
    for (int i = warp; i < T; i += nwarp) {
        uint4 qreg[8], oreg[8];                                              // the row and its dO, packed, in registers
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            qreg[c] = reinterpret_cast<const uint4*>(Qs + (size_t)i * ATTB_PITCH)[c];
            oreg[c] = reinterpret_cast<const uint4*>(Os + (size_t)i * ATTB_PITCH)[c];
        }
        float sc[ATTB_MAX_T / 32];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < ATTB_MAX_T / 32; ++k) {
            const int j = lane + 32 * k;
            sc[k] = -INFINITY;
            if (32 * k <= i) {                                               // warp-uniform: chunks past the diagonal are skipped
                if (j <= i) { sc[k] = attb_dot_rs(qreg, Ks + (size_t)j * ATTB_PITCH) * scale; mx = fmaxf(mx, sc[k]); }
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < ATTB_MAX_T / 32; ++k) { sc[k] = (lane + 32 * k <= i) ? __expf(sc[k] - mx) : 0.f; sum += sc[k]; }
        const float linv = 1.f / warp_sum(sum);
        float dp[ATTB_MAX_T / 32];
        float Dp = 0.f;
#pragma unroll
        for (int k = 0; k < ATTB_MAX_T / 32; ++k) {
            const int j = lane + 32 * k;
            dp[k] = 0.f;
            if (32 * k <= i) {
                if (j <= i) { sc[k] *= linv; dp[k] = attb_dot_rs(oreg, Vs + (size_t)j * ATTB_PITCH); Dp += sc[k] * dp[k]; }
            }
        }
        const float D = warp_sum(Dp);
        float dq[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) dq[c] = 0.f;
#pragma unroll
        for (int k = 0; k < ATTB_MAX_T / 32; ++k) {
            const int j = lane + 32 * k;
            if (32 * k <= i) {
                if (j <= i) {
                    const float ds = sc[k] * (dp[k] - D) * scale;
                    const uint4* kj = reinterpret_cast<const uint4*>(Ks + (size_t)j * ATTB_PITCH);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float f[8];
                        bf8_unpack(kj[c], f);
#pragma unroll
                        for (int u = 0; u < 8; ++u) dq[8 * c + u] = fmaf(ds, f[u], dq[8 * c + u]);
                    }
                }
            }
        }
        if (lane == 0) { st_m[i] = mx; st_l[i] = linv; st_D[i] = D; }
        float mine0 = 0.f, mine1 = 0.f;                                      // lane l keeps dims 2l, 2l + 1 of the reduced row
#pragma unroll
        for (int c = 0; c < 64; ++c) {
            const float v = warp_sum(dq[c]);
            if ((c >> 1) == lane) { if (c & 1) mine1 = v; else mine0 = v; }
        }
        *reinterpret_cast<__nv_bfloat162*>(dqkv + (size_t)(r0 + i) * 3 * d_model + head * 64 + 2 * lane) = __floats2bfloat162_rn(mine0, mine1);
    }
    __syncthreads();
    // ---- phase 2: dK and dV (warp per key row, lanes over the queries i >= j)
    for (int j = warp; j < T; j += nwarp) {
        uint4 kreg[8], vreg[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            kreg[c] = reinterpret_cast<const uint4*>(Ks + (size_t)j * ATTB_PITCH)[c];
            vreg[c] = reinterpret_cast<const uint4*>(Vs + (size_t)j * ATTB_PITCH)[c];
        }
        float dk[64], dv[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
        for (int i = j + lane; i < T; i += 32) {
            const __nv_bfloat16* qi = Qs + (size_t)i * ATTB_PITCH;
            const __nv_bfloat16* oi = Os + (size_t)i * ATTB_PITCH;
            const float p = __expf(attb_dot_rs(kreg, qi) * scale - st_m[i]) * st_l[i];
            const float ds = p * (attb_dot_rs(vreg, oi) - st_D[i]) * scale;
            const uint4* q4 = reinterpret_cast<const uint4*>(qi);
            const uint4* o4 = reinterpret_cast<const uint4*>(oi);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float fq[8], fo[8];
                bf8_unpack(q4[c], fq); bf8_unpack(o4[c], fo);
#pragma unroll
                for (int u = 0; u < 8; ++u) { dk[8 * c + u] = fmaf(ds, fq[u], dk[8 * c + u]); dv[8 * c + u] = fmaf(p, fo[u], dv[8 * c + u]); }
            }
        }
        float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) {
            const float a = warp_sum(dk[c]), b = warp_sum(dv[c]);
            if ((c >> 1) == lane) { if (c & 1) { k1 = a; v1 = b; } else { k0 = a; v0 = b; } }
        }
        __nv_bfloat16* o = dqkv + (size_t)(r0 + j) * 3 * d_model + head * 64 + 2 * lane;
        *reinterpret_cast<__nv_bfloat162*>(o + d_model) = __floats2bfloat162_rn(k0, k1);
        *reinterpret_cast<__nv_bfloat162*>(o + 2 * d_model) = __floats2bfloat162_rn(v0, v1);
    }
}

// ---- tensor-core version of the same computation (mma.sync m16n8k16, bf16 in / fp32 accumulate) --------------------------------
// Same staging and the same two phases, in 16-row blocks per warp:
//   phase 1 (query block): pass A  S = Q K^T -> running row max / sum;   pass B  S again, dP = dO V^T, dS = P (dP - D) scale,
//                          dQ += dS K  (dS re-used straight from the accumulator fragments as the A operand, K through ldmatrix.trans);
//   phase 2 (key block)  : S^T = K Q^T and dP^T = V dO^T directly (no transposes), P^T / dS^T from the per-query statistics,
//                          dK += dS^T Q,  dV += P^T dO.
// D_i = rowsum(P * dP) = dO_i . O_i is taken from the forward output (the FlashAttention-2 identity).  Rows >= T of the last block are
// zero-filled in shared memory and masked.
__device__ __forceinline__ void attb_load_a(const __nv_bfloat16* tile, int row0, int g, int t, uint32_t (&a)[4][4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const uint32_t* r0p = reinterpret_cast<const uint32_t*>(tile + (size_t)(row0 + g) * ATTB_PITCH);
        const uint32_t* r1p = reinterpret_cast<const uint32_t*>(tile + (size_t)(row0 + g + 8) * ATTB_PITCH);
        a[kk][0] = r0p[8 * kk + t]; a[kk][1] = r1p[8 * kk + t]; a[kk][2] = r0p[8 * kk + 4 + t]; a[kk][3] = r1p[8 * kk + 4 + t];
    }
}
// acc[4] = A_blk (16 x 64, fragments a) . B_rows[n0 .. n0+7]^T   (B rows are k-contiguous in shared memory)
__device__ __forceinline__ void attb_mma_rowsT(const uint32_t (&a)[4][4], const __nv_bfloat16* tile, int n0, int g, int t, float (&acc)[4]) {
    acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
    const uint32_t* rp = reinterpret_cast<const uint32_t*>(tile + (size_t)(n0 + g) * ATTB_PITCH);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) mma_bf16(acc, a[kk], rp[8 * kk + t], rp[8 * kk + 4 + t]);
}
// out[8][4] += A (16 x 16, fragment a) . tile[k0 .. k0+15][0 .. 63]   (tile rows are the k index: B through ldmatrix.trans)
__device__ __forceinline__ void attb_mma_rows(const uint32_t (&a)[4], uint32_t tile_addr, int k0, int lane, float (&out)[8][4]) {
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
        const int row = k0 + (lane & 7) + ((lane >> 3) & 1) * 8, col = 16 * dp + (lane >> 4) * 8;
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(tile_addr + (uint32_t)(row * ATTB_PITCH + col) * 2u, b0, b1, b2, b3);
        mma_bf16(out[2 * dp], a, b0, b1);
        mma_bf16(out[2 * dp + 1], a, b2, b3);
    }
}
__global__ void __launch_bounds__(ATTB_THREADS, 1) attn_bwd_mma_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kcache,
                                                                       const __nv_bfloat16* __restrict__ vcache, const __nv_bfloat16* __restrict__ dout,
                                                                       const __nv_bfloat16* __restrict__ att, RowsDev rows, int n_head, int d_model, KvPages kv,
                                                                       float scale, __nv_bfloat16* __restrict__ dqkv) {
    extern __shared__ __align__(16) unsigned char ab_raw[];
    const int slot = blockIdx.x / n_head, head = blockIdx.x % n_head;
    if (slot >= *rows.n_slots) return;
    const int env = rows.slot_env[slot], r0 = rows.slot_row0[slot], T = rows.slot_nrow[slot];
    const int Tp = (T + 15) & ~15, nblk = Tp >> 4;
    __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(ab_raw);
    __nv_bfloat16* Ks = Qs + (size_t)Tp * ATTB_PITCH;
    __nv_bfloat16* Vs = Ks + (size_t)Tp * ATTB_PITCH;
    __nv_bfloat16* Os = Vs + (size_t)Tp * ATTB_PITCH;
    float* st_m = reinterpret_cast<float*>(Os + (size_t)Tp * ATTB_PITCH);
    float* st_l = st_m + Tp;
    float* st_D = st_l + Tp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = ATTB_THREADS >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int* ptab = kv.table + (size_t)env * kv.ppe;                       // position r of the sequence: row r % 32 of page ptab[r / 32]
    auto kv_off = [&](int r) { return kv_row_off(ptab[r >> KV_PAGE_SHIFT], n_head, head, r); };
    for (int i = tid; i < Tp * 8; i += ATTB_THREADS) {
        const int r = i >> 3, w = i & 7;
        uint4 vq = make_uint4(0, 0, 0, 0), vo = vq, vk = vq, vv = vq;
        if (r < T) {
            const size_t grow = (size_t)(r0 + r) * d_model + head * 64 + 8 * w;
            vq = *reinterpret_cast<const uint4*>(q + grow);
            vo = *reinterpret_cast<const uint4*>(dout + grow);
            vk = *reinterpret_cast<const uint4*>(kcache + kv_off(r) + 8 * w);
            vv = *reinterpret_cast<const uint4*>(vcache + kv_off(r) + 8 * w);
        }
        reinterpret_cast<uint4*>(Qs + (size_t)r * ATTB_PITCH)[w] = vq;
        reinterpret_cast<uint4*>(Os + (size_t)r * ATTB_PITCH)[w] = vo;
        reinterpret_cast<uint4*>(Ks + (size_t)r * ATTB_PITCH)[w] = vk;
        reinterpret_cast<uint4*>(Vs + (size_t)r * ATTB_PITCH)[w] = vv;
    }
    for (int i = tid; i < Tp; i += ATTB_THREADS) {                            // D_i = dO_i . O_i
        float D = 0.f;
        if (i < T) {
            const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(dout + (size_t)(r0 + i) * d_model + head * 64);
            const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(att + (size_t)(r0 + i) * d_model + head * 64);
#pragma unroll 8
            for (int c = 0; c < 32; ++c) { const float2 x = __bfloat1622float2(a[c]), y = __bfloat1622float2(b[c]); D = fmaf(x.x, y.x, D); D = fmaf(x.y, y.y, D); }
        }
        st_D[i] = D;
    }
    __syncthreads();
    const uint32_t Ks_addr = smem_u32(Ks), Qs_addr = smem_u32(Qs), Os_addr = smem_u32(Os);
    // ---- phase 1: query blocks (heaviest first so the warps finish together)
    for (int qb = nblk - 1 - warp; qb >= 0; qb -= nwarp) {
        const int i0 = 16 * qb;
        uint32_t qa[4][4], oa[4][4];
        attb_load_a(Qs, i0, g, t, qa);
        attb_load_a(Os, i0, g, t, oa);
        const int rowa = i0 + g, rowb = i0 + g + 8;
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
        for (int jb = 0; jb < 2 * (qb + 1); ++jb) {                           // pass A: row statistics
            float s[4];
            attb_mma_rowsT(qa, Ks, 8 * jb, g, t, s);
            const int key = 8 * jb + 2 * t;
            s[0] = key <= rowa ? s[0] * scale : -INFINITY; s[1] = key + 1 <= rowa ? s[1] * scale : -INFINITY;
            s[2] = key <= rowb ? s[2] * scale : -INFINITY; s[3] = key + 1 <= rowb ? s[3] * scale : -INFINITY;
            float b0 = fmaxf(s[0], s[1]), b1 = fmaxf(s[2], s[3]);
            b0 = fmaxf(b0, __shfl_xor_sync(0xFFFFFFFFu, b0, 1)); b0 = fmaxf(b0, __shfl_xor_sync(0xFFFFFFFFu, b0, 2));
            b1 = fmaxf(b1, __shfl_xor_sync(0xFFFFFFFFu, b1, 1)); b1 = fmaxf(b1, __shfl_xor_sync(0xFFFFFFFFu, b1, 2));
            const float n0 = fmaxf(m0, b0), n1 = fmaxf(m1, b1);
            if (n0 > -INFINITY) { l0 = l0 * __expf(m0 - n0) + __expf(s[0] - n0) + __expf(s[1] - n0); m0 = n0; }
            if (n1 > -INFINITY) { l1 = l1 * __expf(m1 - n1) + __expf(s[2] - n1) + __expf(s[3] - n1); m1 = n1; }
        }
        l0 += __shfl_xor_sync(0xFFFFFFFFu, l0, 1); l0 += __shfl_xor_sync(0xFFFFFFFFu, l0, 2);
        l1 += __shfl_xor_sync(0xFFFFFFFFu, l1, 1); l1 += __shfl_xor_sync(0xFFFFFFFFu, l1, 2);
        const float li0 = l0 > 0.f ? 1.f / l0 : 0.f, li1 = l1 > 0.f ? 1.f / l1 : 0.f;
        if (t == 0) { st_m[rowa] = m0; st_l[rowa] = li0; st_m[rowb] = m1; st_l[rowb] = li1; }
        const float D0 = st_D[rowa], D1 = st_D[rowb];
        float dq[8][4];
#pragma unroll
        for (int d8 = 0; d8 < 8; ++d8) { dq[d8][0] = dq[d8][1] = dq[d8][2] = dq[d8][3] = 0.f; }
        for (int kb = 0; kb <= qb; ++kb) {                                    // pass B: dS and dQ += dS K, sixteen keys at a time
            uint32_t a[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float s[4], dp[4];
                attb_mma_rowsT(qa, Ks, 16 * kb + 8 * h, g, t, s);
                attb_mma_rowsT(oa, Vs, 16 * kb + 8 * h, g, t, dp);
                const int key = 16 * kb + 8 * h + 2 * t;
                const float p0 = key <= rowa ? __expf(s[0] * scale - m0) * li0 : 0.f, p1 = key + 1 <= rowa ? __expf(s[1] * scale - m0) * li0 : 0.f;
                const float p2 = key <= rowb ? __expf(s[2] * scale - m1) * li1 : 0.f, p3 = key + 1 <= rowb ? __expf(s[3] * scale - m1) * li1 : 0.f;
                a[2 * h] = pack_bf16x2(p0 * (dp[0] - D0) * scale, p1 * (dp[1] - D0) * scale);
                a[2 * h + 1] = pack_bf16x2(p2 * (dp[2] - D1) * scale, p3 * (dp[3] - D1) * scale);
            }
            attb_mma_rows(a, Ks_addr, 16 * kb, lane, dq);
        }
#pragma unroll
        for (int d8 = 0; d8 < 8; ++d8) {
            if (rowa < T) *reinterpret_cast<__nv_bfloat162*>(dqkv + (size_t)(r0 + rowa) * 3 * d_model + head * 64 + 8 * d8 + 2 * t) = __floats2bfloat162_rn(dq[d8][0], dq[d8][1]);
            if (rowb < T) *reinterpret_cast<__nv_bfloat162*>(dqkv + (size_t)(r0 + rowb) * 3 * d_model + head * 64 + 8 * d8 + 2 * t) = __floats2bfloat162_rn(dq[d8][2], dq[d8][3]);
        }
    }
    __syncthreads();
    // ---- phase 2: key blocks (heaviest = first block, first)
    for (int kb = warp; kb < nblk; kb += nwarp) {
        const int j0 = 16 * kb;
        uint32_t ka[4][4], va[4][4];
        attb_load_a(Ks, j0, g, t, ka);
        attb_load_a(Vs, j0, g, t, va);
        const int keya = j0 + g, keyb = j0 + g + 8;
        float dk[8][4], dv[8][4];
#pragma unroll
        for (int d8 = 0; d8 < 8; ++d8) { dk[d8][0] = dk[d8][1] = dk[d8][2] = dk[d8][3] = 0.f; dv[d8][0] = dv[d8][1] = dv[d8][2] = dv[d8][3] = 0.f; }
        for (int qb = kb; qb < nblk; ++qb) {                                  // sixteen queries at a time
            uint32_t pa[4], da[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float s[4], dp[4];
                attb_mma_rowsT(ka, Qs, 16 * qb + 8 * h, g, t, s);              // S^T block: rows = keys, columns = queries
                attb_mma_rowsT(va, Os, 16 * qb + 8 * h, g, t, dp);             // dP^T block
                const int qq = 16 * qb + 8 * h + 2 * t;
                const float mq0 = st_m[qq], lq0 = st_l[qq], Dq0 = st_D[qq], mq1 = st_m[qq + 1], lq1 = st_l[qq + 1], Dq1 = st_D[qq + 1];
                const bool v0 = qq < T, v1 = qq + 1 < T;
                const float p0 = (v0 && qq >= keya) ? __expf(s[0] * scale - mq0) * lq0 : 0.f, p1 = (v1 && qq + 1 >= keya) ? __expf(s[1] * scale - mq1) * lq1 : 0.f;
                const float p2 = (v0 && qq >= keyb) ? __expf(s[2] * scale - mq0) * lq0 : 0.f, p3 = (v1 && qq + 1 >= keyb) ? __expf(s[3] * scale - mq1) * lq1 : 0.f;
                pa[2 * h] = pack_bf16x2(p0, p1); pa[2 * h + 1] = pack_bf16x2(p2, p3);
                da[2 * h] = pack_bf16x2(p0 * (dp[0] - Dq0) * scale, p1 * (dp[1] - Dq1) * scale);
                da[2 * h + 1] = pack_bf16x2(p2 * (dp[2] - Dq0) * scale, p3 * (dp[3] - Dq1) * scale);
            }
            attb_mma_rows(da, Qs_addr, 16 * qb, lane, dk);
            attb_mma_rows(pa, Os_addr, 16 * qb, lane, dv);
        }
#pragma unroll
        for (int d8 = 0; d8 < 8; ++d8) {
            if (keya < T) {
                __nv_bfloat16* o = dqkv + (size_t)(r0 + keya) * 3 * d_model + head * 64 + 8 * d8 + 2 * t;
                *reinterpret_cast<__nv_bfloat162*>(o + d_model) = __floats2bfloat162_rn(dk[d8][0], dk[d8][1]);
                *reinterpret_cast<__nv_bfloat162*>(o + 2 * d_model) = __floats2bfloat162_rn(dv[d8][0], dv[d8][1]);
            }
            if (keyb < T) {
                __nv_bfloat16* o = dqkv + (size_t)(r0 + keyb) * 3 * d_model + head * 64 + 8 * d8 + 2 * t;
                *reinterpret_cast<__nv_bfloat162*>(o + d_model) = __floats2bfloat162_rn(dk[d8][2], dk[d8][3]);
                *reinterpret_cast<__nv_bfloat162*>(o + 2 * d_model) = __floats2bfloat162_rn(dv[d8][2], dv[d8][3]);
            }
        }
    }
}

// dwte[tok] += dx_row, dwpe[pos] += dx_row   (one warp per row)
__global__ void embed_bwd_kernel(const float* __restrict__ dx, RowsDev rows, int d, float* __restrict__ dwte, float* __restrict__ dwpe) {
    const int M = *rows.M;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < M; r += warps) {
        float* a = dwte + (size_t)rows.row_tok[r] * d;
        float* b = dwpe + (size_t)rows.row_pos[r] * d;
        for (int c = lane; c < d; c += 32) { const float v = dx[(size_t)r * d + c]; atomicAdd(a + c, v); atomicAdd(b + c, v); }
    }
}

// LM head: logits row (fp32, dumped by the head GEMM) -> dlogits = g * (onehot(target) - softmax) as bf16 [v_pad] (padding columns 0).
// One CTA per scored row.
__global__ void dlogits_kernel(const float* __restrict__ logits, int n_vocab, int v_pad, const int* __restrict__ targets, const float* __restrict__ g,
                               const int* __restrict__ n_dev, float inv_temp, __nv_bfloat16* __restrict__ dl) {
    __shared__ float red[256];
    const int r = blockIdx.x;
    __nv_bfloat16* o = dl + (size_t)r * v_pad;
    if (r >= *n_dev) { for (int c = threadIdx.x; c < v_pad; c += blockDim.x) o[c] = __float2bfloat16(0.f); return; }
    const float* l = logits + (size_t)r * n_vocab;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < n_vocab; c += blockDim.x) mx = fmaxf(mx, l[c] * inv_temp);
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int s = 128; s; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
    mx = red[0];
    __syncthreads();
    float sum = 0.f;
    for (int c = threadIdx.x; c < n_vocab; c += blockDim.x) sum += __expf(l[c] * inv_temp - mx);
    red[threadIdx.x] = sum;
    __syncthreads();
    for (int s = 128; s; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
    const float inv = 1.f / red[0], gr = g[r] * inv_temp;
    const int tgt = targets[r];
    for (int c = threadIdx.x; c < v_pad; c += blockDim.x) {
        float v = 0.f;
        if (c < n_vocab) v = gr * ((c == tgt ? 1.f : 0.f) - __expf(l[c] * inv_temp - mx) * inv);
        o[c] = __float2bfloat16(v);
    }
}

// Fused PPO variant of dlogits_kernel: the row's log-prob is formed from the dumped logits and the policy part of ppo_loss_fn
// (ppo/base_interface.py:113-127) is differentiated in place:  ratio = exp(lp - old_lp),  loss term = max(-A ratio, -A clip(ratio)),
// g = d term / d lp / n.  acc[0..3] (double) += policy term, ratio - 1 - log ratio, clip indicator, ratio.
__global__ void dlogits_ppo_kernel(const float* __restrict__ logits, int n_vocab, int v_pad, const int* __restrict__ targets, const float* __restrict__ old_lp,
                                   const float* __restrict__ adv, const int* __restrict__ n_dev, float ratio_lo, float ratio_hi, __nv_bfloat16* __restrict__ dl,
                                   double* __restrict__ acc) {
    __shared__ float red[256];
    const int r = blockIdx.x;
    __nv_bfloat16* o = dl + (size_t)r * v_pad;
    const int n = *n_dev;
    if (r >= n) { for (int c = threadIdx.x; c < v_pad; c += blockDim.x) o[c] = __float2bfloat16(0.f); return; }
    const float* l = logits + (size_t)r * n_vocab;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < n_vocab; c += blockDim.x) mx = fmaxf(mx, l[c]);
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int s = 128; s; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
    mx = red[0];
    __syncthreads();
    float sum = 0.f;
    for (int c = threadIdx.x; c < n_vocab; c += blockDim.x) sum += __expf(l[c] - mx);
    red[threadIdx.x] = sum;
    __syncthreads();
    for (int s = 128; s; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
    const float total = red[0], inv = 1.f / total;
    const int tgt = targets[r];
    const float lp = l[tgt] - mx - logf(total);
    const float lr = lp - old_lp[r], ratio = expf(lr), A = adv[r];
    const float pg1 = -A * ratio, pg2 = -A * fminf(fmaxf(ratio, ratio_lo), ratio_hi);
    const bool clipped = pg2 > pg1;
    const float dterm = (clipped && !(ratio >= ratio_lo && ratio <= ratio_hi)) ? 0.f : -A * ratio;
    const float gr = dterm / (float)n;
    if (threadIdx.x == 0) {
        atomicAdd(acc + 0, (double)fmaxf(pg1, pg2)); atomicAdd(acc + 1, (double)((ratio - 1.f) - lr)); atomicAdd(acc + 2, clipped ? 1.0 : 0.0);
        atomicAdd(acc + 3, (double)ratio);
    }
    for (int c = threadIdx.x; c < v_pad; c += blockDim.x) {
        float v = 0.f;
        if (c < n_vocab) v = gr * ((c == tgt ? 1.f : 0.f) - __expf(l[c] - mx) * inv);
        o[c] = __float2bfloat16(v);
    }
}
// value part of ppo_loss_fn (:101-111) on the scored rows: dv[row] = coef * 0.5 * d max((v - R)^2, (clip(v) - R)^2) / dv / n;
// acc[4] += max(vf1, vf2), acc[5] += clip indicator, acc[6] += (v - R)^2, acc[7] += v
__global__ void ppo_value_grad_kernel(const float* __restrict__ values, const int* __restrict__ rows, const float* __restrict__ old_v, const float* __restrict__ ret,
                                      const int* __restrict__ n_dev, float cliprange_value, float coef, float* __restrict__ dv, double* __restrict__ acc) {
    const int n = *n_dev;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int row = rows[r];
    const float v = values[row], ov = old_v[r], R = ret[r];
    const float lo = ov - cliprange_value, hi = ov + cliprange_value;
    const float vc = fminf(fmaxf(v, lo), hi);
    const float vf1 = (v - R) * (v - R), vf2 = (vc - R) * (vc - R);
    const bool clipped = vf2 > vf1;
    const float d = clipped ? ((v >= lo && v <= hi) ? 2.f * (vc - R) : 0.f) : 2.f * (v - R);
    dv[row] = coef * 0.5f * d / (float)n;
    atomicAdd(acc + 4, (double)fmaxf(vf1, vf2)); atomicAdd(acc + 5, clipped ? 1.0 : 0.0); atomicAdd(acc + 6, (double)vf1); atomicAdd(acc + 7, (double)v);
}

// dst[idx[r]] += src[r]   (rows of d floats)
__global__ void scatter_add_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, const int* __restrict__ n_dev, int d, float* __restrict__ dst) {
    const int n = *n_dev;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps)
        for (int c = lane; c < d; c += 32) dst[(size_t)idx[r] * d + c] += src[(size_t)r * d + c];
}

// LinearHead value v_t = LN_f(x_t) . w + b  (fp32, as value_heads_kernel computes it):  dfx[t] += dv_t w;  dw += dv_t LN_f(x_t);  db += dv_t
__global__ void value_head_bwd_kernel(const float* __restrict__ x, const float* __restrict__ lnf_g, const float* __restrict__ lnf_b, const float* __restrict__ w,
                                      const float* __restrict__ dv, int d, float eps, const int* __restrict__ M_dev, float* __restrict__ dfx,
                                      float* __restrict__ dw, float* __restrict__ db) {
    extern __shared__ float vh_sm[];             // [d + 1]
    const int M = *M_dev;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int c = threadIdx.x; c <= d; c += blockDim.x) vh_sm[c] = 0.f;
    __syncthreads();
    for (int r = blockIdx.x * nwarp + warp; r < M; r += gridDim.x * nwarp) {
        const float g = dv[r];
        if (g == 0.f) continue;
        const float* xr = x + (size_t)r * d;
        float s = 0.f;
        for (int c = lane; c < d; c += 32) s += xr[c];
        const float mean = warp_sum(s) / d;
        float q = 0.f;
        for (int c = lane; c < d; c += 32) { const float t = xr[c] - mean; q += t * t; }
        const float rstd = rsqrtf(warp_sum(q) / d + eps);
        for (int c = lane; c < d; c += 32) {
            atomicAdd(vh_sm + c, g * ((xr[c] - mean) * rstd * lnf_g[c] + lnf_b[c]));
            dfx[(size_t)r * d + c] += g * w[c];
        }
        if (lane == 0) atomicAdd(vh_sm + d, g);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) if (vh_sm[c] != 0.f) atomicAdd(dw + c, vh_sm[c]);
    if (threadIdx.x == 0 && vh_sm[d] != 0.f) atomicAdd(db, vh_sm[d]);
}

// out[r][c] = in[r][c] as bf16 (rows < M)
__global__ void f32_to_bf16_rows_kernel(const float* __restrict__ src, const int* __restrict__ M_dev, int C, __nv_bfloat16* __restrict__ dst) {
    const size_t n = (size_t)*M_dev * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = __float2bfloat16(src[i]);
}

// ---- training state -----------------------------------------------------------------------------------------------------------
struct TrainLayerT { __nv_bfloat16 *w_qkv, *w_proj, *w_fc, *w_fc2; };      // HF-layout ([in][out]) bf16 copies: the W operand of the dgrad GEMMs
struct TrainDev {
    bool ready = false;
    int m_cap = 0, lp_cap = 0;                 // row capacity / scored-row capacity
    std::vector<TrainLayerT> wt;
    __nv_bfloat16* wte_t = nullptr;            // [d][v_pad] transposed embedding (dgrad of the tied LM head)
    float* xsave = nullptr;                    // [L + 1][m_cap][d] residual stream before every layer and after the last
    float *dx = nullptr, *dfx = nullptr;       // [m_cap][d] gradient of the residual stream / of the final-LN output
    __nv_bfloat16 *xn1 = nullptr, *hpre = nullptr, *dxb = nullptr, *dxn = nullptr, *datt = nullptr, *dqkv = nullptr, *dh = nullptr;
    __nv_bfloat16 *tA = nullptr, *tB = nullptr;  // transposed operands of the wgrad GEMMs: [<= max(4d, v_pad)][Mp] each
    float* zero_bias = nullptr;                // zeros, the bias of the plain-store GEMM epilogue
    int* consts = nullptr;                     // device ints: {d, 3d, 4d, n_vocab_rows, n_lp}
    // LM head
    float* lp_logits = nullptr;                // [lp_cap][n_vocab] fp32
    __nv_bfloat16* dl = nullptr;               // [lp_cap][v_pad]
    float* dxs = nullptr;                      // [lp_cap][d] fp32
    int *lp_rows = nullptr, *lp_targets = nullptr;
    float *lp_g = nullptr, *dv = nullptr;      // [lp_cap], [m_cap]
    float *lp_aux = nullptr;                   // [4][lp_cap] fused PPO step: old log-probs, advantages, old values, returns of the scored rows
    float* values = nullptr;                   // [m_cap] LinearHead value of every row (fused PPO step)
    double* loss_acc = nullptr;                // [8] loss sums of the fused PPO step
    unsigned long long* best = nullptr;
    uint64_t* row_key = nullptr;
};

}  // namespace lmrl

using namespace lmrl;

void lmrl_train_free(Gpt2Dev* g) {
    delete g->train;
    g->train = nullptr;
}

// bf16 HF-layout copies of tower 0's matrices (the engine stores W as [out][in]; dgrad needs [in][out] as its K-major W operand)
static int train_refresh_weight_copies(lmrl_engine* e) {
    Gpt2Dev& G = *e->gpt2;
    if (!G.train || !G.train->ready) return 0;
    TrainDev& Tr = *G.train;
    const lmrl_gpt2_dims& D = G.dims;
    const int d = D.d_model, ff = D.d_ff;
    cudaStream_t st = e->stream;
    auto tr = [&](const __nv_bfloat16* src, int rows, int cols, __nv_bfloat16* dst, const int* rows_dev) {     // [rows][cols] -> [cols][rows]
        transpose_to_bf16_rows_kernel<__nv_bfloat16><<<dim3((rows + 31) / 32, (cols + 31) / 32), dim3(32, 8), 0, st>>>(src, rows_dev, cols, dst, rows, nullptr);
    };
    for (int l = 0; l < D.n_layer; ++l) {
        const LayerDev& L = G.tw[0].layers[l];
        tr(L.w_qkv, 3 * d, d, Tr.wt[l].w_qkv, Tr.consts + 1);
        tr(L.w_proj, d, d, Tr.wt[l].w_proj, Tr.consts + 0);
        tr(L.w_fc, ff, d, Tr.wt[l].w_fc, Tr.consts + 2);
        tr(L.w_fc2, d, ff, Tr.wt[l].w_fc2, Tr.consts + 0);
    }
    tr(G.tw[0].wte, G.v_pad, d, Tr.wte_t, Tr.consts + 5);
    LMRL_CUDA(cudaGetLastError());
    return 0;
}

extern "C" {

int lmrl_train_enable(lmrl_engine* e, int32_t max_rows, int32_t max_scored_rows) {
    LMRL_CHECK(e && e->gpt2 && e->gpt2->tw[0].loaded, "lmrl_train_enable: load the policy (tower 0) first");
    LMRL_CUDA(cudaSetDevice(e->device));
    Gpt2Dev& G = *e->gpt2;
    if (G.train && G.train->ready) return 0;
    const lmrl_gpt2_dims& D = G.dims;
    const int d = D.d_model, ff = D.d_ff, L = D.n_layer;
    LMRL_CHECK(e->cfg.max_input_length <= 384, "lmrl_train_enable: the attention backward kernel keeps dK / dV of one sequence in shared memory: max_input_length %d > 384",
               e->cfg.max_input_length);
    LMRL_CHECK(d % 256 == 0 && ff % 256 == 0, "lmrl_train_enable: d_model and d_ff must be multiples of 256");
    if (!G.train) G.train = new TrainDev();
    TrainDev& Tr = *G.train;
    Tr.m_cap = (int)std::min<size_t>((size_t)G.m_cap, ((size_t)std::max(max_rows, 128) + 127) / 128 * 128);
    Tr.lp_cap = (std::max(max_scored_rows, 128) + 127) / 128 * 128;
    const size_t M = (size_t)Tr.m_cap, P = (size_t)Tr.lp_cap;
    Tr.wt.resize(L);
    int rc = 0;
    for (int l = 0; l < L && !rc; ++l)
        rc |= e->alloc(&Tr.wt[l].w_qkv, (size_t)3 * d * d) | e->alloc(&Tr.wt[l].w_proj, (size_t)d * d) | e->alloc(&Tr.wt[l].w_fc, (size_t)d * ff) | e->alloc(&Tr.wt[l].w_fc2, (size_t)d * ff);
    const size_t t_elems = std::max((size_t)G.v_pad * P, (size_t)std::max(ff, 3 * d) * M);
    rc |= e->alloc(&Tr.wte_t, (size_t)d * G.v_pad) | e->alloc(&Tr.xsave, (size_t)(L + 1) * M * d) | e->alloc(&Tr.dx, M * d) | e->alloc(&Tr.dfx, M * d) |
          e->alloc(&Tr.xn1, M * d) | e->alloc(&Tr.hpre, M * ff) | e->alloc(&Tr.dxb, M * d) | e->alloc(&Tr.dxn, M * d) | e->alloc(&Tr.datt, M * d) |
          e->alloc(&Tr.dqkv, M * 3 * d) | e->alloc(&Tr.dh, M * ff) | e->alloc(&Tr.tA, t_elems) | e->alloc(&Tr.tB, (size_t)std::max(ff, 3 * d) * M) |
          e->alloc(&Tr.zero_bias, (size_t)std::max(ff, G.v_pad)) | e->alloc(&Tr.consts, 8) | e->alloc(&Tr.lp_logits, P * D.n_vocab) | e->alloc(&Tr.dl, P * G.v_pad) |
          e->alloc(&Tr.dxs, P * d) | e->alloc(&Tr.lp_rows, P) | e->alloc(&Tr.lp_targets, P) | e->alloc(&Tr.lp_g, P) | e->alloc(&Tr.dv, M) | e->alloc(&Tr.lp_aux, 4 * P) | e->alloc(&Tr.values, M) | e->alloc(&Tr.loss_acc, 8) | e->alloc(&Tr.best, P) |
          e->alloc(&Tr.row_key, P);
    if (rc) return -1;
    const int consts[8] = {d, 3 * d, ff, (int)G.v_pad, 0, (int)G.v_pad, 0, 0};      // [4] scored rows, [6] wte rows of the gradient: per call
    LMRL_CUDA(cudaMemcpyAsync(Tr.consts, consts, sizeof consts, cudaMemcpyHostToDevice, e->stream));
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    Tr.ready = true;
    return train_refresh_weight_copies(e);
}

}  // extern "C"

// one decoder layer forward for training: like tower_layers' body, but LN1's output and the FC pre-activation are kept
// (Tr.xn1, Tr.hpre) and the residual stream G.x ends as x_mid (with_fc2 = false) or as the layer output (true)
static int train_layer_forward(lmrl_engine* e, int l, int n_seq, bool with_fc2) {
    Gpt2Dev& G = *e->gpt2;
    TrainDev& Tr = *G.train;
    const lmrl_gpt2_dims& D = G.dims;
    const TowerDev& T = G.tw[0];
    const LayerDev& L = T.layers[l];
    cudaStream_t st = e->stream;
    const int ew = G.sm_count * 8;
    LMRL_K(layernorm_kernel, ew, 256, 0, st, G.x, L.ln1_g, L.ln1_b, D.d_model, D.ln_eps, G.rows.M, (const int*)nullptr, Tr.xn1, D.d_model);
    GemmArgs a{};
    a.A = Tr.xn1; a.W = L.w_qkv; a.bias = L.b_qkv; a.N = 3 * D.d_model; a.K = D.d_model; a.M_dev = G.rows.M;
    a.out = G.q; a.kcache = T.kv + (size_t)l * G.kv_layer_stride; a.vcache = a.kcache + G.kv_layer_stride / 2;
    a.row_env = G.rows.row_env; a.row_pos = G.rows.row_pos; a.d_model = D.d_model; a.n_head = D.n_head; a.kv_table = G.kvp.table; a.kv_ppe = G.kvp.ppe;
    if (launch_gemm<EPI_QKV>(e, a, a.N, 0, (size_t)Tr.m_cap, a.N, 0)) return -1;
    LMRL_K(attention_prefill_kernel, attp_grid(n_seq, D.n_head, G.sm_count), ATTP_WARPS * 32, ATTP_WARPS * ATTP_WARP_SMEM, st, G.q,
           a.kcache, a.vcache, G.rows, D.n_head, D.d_model, G.kvp, 0.125f * 1.4426950408889634f, G.attn_wide, G.att);
    GemmArgs p{};
    p.A = G.att; p.W = L.w_proj; p.bias = L.b_proj; p.N = D.d_model; p.K = D.d_model; p.M_dev = G.rows.M; p.resid = G.x;
    if (launch_gemm<EPI_RESID>(e, p, p.N, 0, G.m_cap, p.N, 1)) return -1;
    LMRL_K(layernorm_kernel, ew, 256, 0, st, G.x, L.ln2_g, L.ln2_b, D.d_model, D.ln_eps, G.rows.M, (const int*)nullptr, G.xn, D.d_model);
    GemmArgs f{};                                   // plain store of the pre-activation: the QKV epilogue with every tile a "Q" tile
    f.A = G.xn; f.W = L.w_fc; f.bias = L.b_fc; f.N = D.d_ff; f.K = D.d_model; f.M_dev = G.rows.M; f.out = Tr.hpre; f.d_model = D.d_ff;
    if (launch_gemm<EPI_QKV>(e, f, f.N, 0, (size_t)Tr.m_cap, f.N, 2)) return -1;
    gelu_fwd_kernel<<<ew, 256, 0, st>>>(Tr.hpre, G.rows.M, D.d_ff, G.h);
    if (with_fc2) {
        GemmArgs f2{};
        f2.A = G.h; f2.W = L.w_fc2; f2.bias = L.b_fc2; f2.N = D.d_model; f2.K = D.d_ff; f2.M_dev = G.rows.M; f2.resid = G.x;
        if (launch_gemm<EPI_RESID>(e, f2, f2.N, 0, G.m_cap, f2.N, 3)) return -1;
    }
    e->launches += 5;
    return 0;
}

struct LayerGrads { float *ln1_g, *ln1_b, *attn_w, *attn_b, *proj_w, *proj_b, *ln2_g, *ln2_b, *fc_w, *fc_b, *fc2_w, *fc2_b; };

// backward of layer l.  On entry Tr.dx = dL/d(layer output); on exit Tr.dx = dL/d(layer input).  The layer has just been recomputed
// (train_layer_forward(l, false)): G.x = x_mid, Tr.xn1, G.q / KV slab, G.att, G.xn = xn2, Tr.hpre, G.h are live.
static int train_layer_backward(lmrl_engine* e, int l, int n_seq, int Mp, const LayerGrads& gr) {
    Gpt2Dev& G = *e->gpt2;
    TrainDev& Tr = *G.train;
    const lmrl_gpt2_dims& D = G.dims;
    const LayerDev& L = G.tw[0].layers[l];
    const TrainLayerT& W = Tr.wt[l];
    cudaStream_t st = e->stream;
    const int d = D.d_model, ff = D.d_ff, ew = G.sm_count * 8;
    const int* M_dev = G.rows.M;
    const int* c_d = Tr.consts + 0;
    const int* c_ff = Tr.consts + 2;
    auto tgrid = [&](int C) { return dim3((Mp + 31) / 32, (C + 31) / 32); };
    // dgrad: out[M][n_out] (bf16) = A[M][k_in] . W_hf[n_out][k_in]^T
    auto dgrad = [&](const __nv_bfloat16* A, int k_in, const __nv_bfloat16* W_hf, int n_out, __nv_bfloat16* out) -> int {
        GemmArgs g{};
        g.A = A; g.W = W_hf; g.bias = Tr.zero_bias; g.N = n_out; g.K = k_in; g.M_dev = M_dev; g.out = out; g.d_model = n_out;
        return launch_gemm<EPI_QKV>(e, g, n_out, 0, (size_t)Tr.m_cap, (size_t)n_out, 0);
    };
    // wgrad: grad[k_in][n_out] (fp32, HF layout) += Xt[k_in][Mp] . dYt[n_out][Mp]^T
    auto wgrad = [&](const __nv_bfloat16* Xt, int k_in, const int* k_in_dev, const __nv_bfloat16* dYt, int n_out, float* grad) -> int {
        GemmArgs g{};
        g.A = Xt; g.W = dYt; g.bias = nullptr; g.N = n_out; g.K = Mp; g.M_dev = k_in_dev; g.resid = grad;
        return launch_gemm<EPI_RESID>(e, g, n_out, 0, (size_t)k_in, (size_t)n_out, 1);
    };
    // ---- mlp.c_proj (fc2): x_out = x_mid + h . Wfc2 + b
    transpose_to_bf16_rows_kernel<float><<<tgrid(d), dim3(32, 8), 0, st>>>(Tr.dx, M_dev, d, Tr.tB, Mp, Tr.dxb);          // dxb = bf16(dx), tB = dx^T
    transpose_to_bf16_rows_kernel<__nv_bfloat16><<<tgrid(ff), dim3(32, 8), 0, st>>>(G.h, M_dev, ff, Tr.tA, Mp, nullptr);  // tA = h^T
    if (wgrad(Tr.tA, ff, c_ff, Tr.tB, d, gr.fc2_w)) return -1;
    colsum_kernel<float><<<dim3((d + 127) / 128, 512), 128, 0, st>>>(Tr.dx, M_dev, d, gr.fc2_b);
    if (dgrad(Tr.dxb, d, W.w_fc2, ff, Tr.dh)) return -1;                                                                  // dh = dx . Wfc2^T
    gelu_bwd_kernel<<<ew, 256, 0, st>>>(Tr.hpre, M_dev, ff, Tr.dh);                                                       // dh := dhpre
    // ---- mlp.c_fc: hpre = xn2 . Wfc + b
    transpose_to_bf16_rows_kernel<__nv_bfloat16><<<tgrid(d), dim3(32, 8), 0, st>>>(G.xn, M_dev, d, Tr.tA, Mp, nullptr);   // tA = xn2^T
    transpose_to_bf16_rows_kernel<__nv_bfloat16><<<tgrid(ff), dim3(32, 8), 0, st>>>(Tr.dh, M_dev, ff, Tr.tB, Mp, nullptr); // tB = dhpre^T
    if (wgrad(Tr.tA, d, c_d, Tr.tB, ff, gr.fc_w)) return -1;
    colsum_kernel<__nv_bfloat16><<<dim3((ff + 127) / 128, 256), 128, 0, st>>>(Tr.dh, M_dev, ff, gr.fc_b);
    if (dgrad(Tr.dh, ff, W.w_fc, d, Tr.dxn)) return -1;                                                                   // dxn2
    // ---- ln_2:  dx += LN'(x_mid) dxn2
    ln_bwd_kernel<__nv_bfloat16><<<G.sm_count * 2, 256, 2 * d * sizeof(float), st>>>(G.x, L.ln2_g, Tr.dxn, d, D.ln_eps, M_dev, Tr.dx, 1, gr.ln2_g, gr.ln2_b);
    // ---- attn.c_proj: x_mid = x_in + att . Wproj + b
    transpose_to_bf16_rows_kernel<float><<<tgrid(d), dim3(32, 8), 0, st>>>(Tr.dx, M_dev, d, Tr.tB, Mp, Tr.dxb);
    transpose_to_bf16_rows_kernel<__nv_bfloat16><<<tgrid(d), dim3(32, 8), 0, st>>>(G.att, M_dev, d, Tr.tA, Mp, nullptr);
    if (wgrad(Tr.tA, d, c_d, Tr.tB, d, gr.proj_w)) return -1;
    colsum_kernel<float><<<dim3((d + 127) / 128, 512), 128, 0, st>>>(Tr.dx, M_dev, d, gr.proj_b);
    if (dgrad(Tr.dxb, d, W.w_proj, d, Tr.datt)) return -1;
    // ---- attention
    {
        const __nv_bfloat16* kc = G.tw[0].kv + (size_t)l * G.kv_layer_stride;
        auto smem_for = [](size_t T) { const size_t Tp = (T + 15) / 16 * 16; return Tp * ATTB_PITCH * 2 * 4 + Tp * 3 * sizeof(float) + 16; };
        const size_t smem = smem_for((size_t)e->cfg.max_input_length);
        static bool configured = false;
        if (!configured) {
            LMRL_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(ATTB_MAX_T)));
            LMRL_CUDA(cudaFuncSetAttribute(attn_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(ATTB_MAX_T)));
            configured = true;
        }
        const char* sel = getenv("LMRL_ATTN_BWD");                        // "cuda" selects the CUDA-core kernel (cross-check in the tests)
        if (sel && sel[0] == 'c')
            attn_bwd_kernel<<<n_seq * D.n_head, ATTB_THREADS, smem, st>>>(G.q, kc, kc + G.kv_layer_stride / 2, Tr.datt, G.rows, D.n_head, d, G.kvp, 0.125f, Tr.dqkv);
        else
            attn_bwd_mma_kernel<<<n_seq * D.n_head, ATTB_THREADS, smem, st>>>(G.q, kc, kc + G.kv_layer_stride / 2, Tr.datt, G.att, G.rows, D.n_head, d, G.kvp, 0.125f,
                                                                              Tr.dqkv);
    }
    // ---- attn.c_attn: qkv = xn1 . Wqkv + b
    transpose_to_bf16_rows_kernel<__nv_bfloat16><<<tgrid(d), dim3(32, 8), 0, st>>>(Tr.xn1, M_dev, d, Tr.tA, Mp, nullptr);
    transpose_to_bf16_rows_kernel<__nv_bfloat16><<<tgrid(3 * d), dim3(32, 8), 0, st>>>(Tr.dqkv, M_dev, 3 * d, Tr.tB, Mp, nullptr);
    if (wgrad(Tr.tA, d, c_d, Tr.tB, 3 * d, gr.attn_w)) return -1;
    colsum_kernel<__nv_bfloat16><<<dim3((3 * d + 127) / 128, 256), 128, 0, st>>>(Tr.dqkv, M_dev, 3 * d, gr.attn_b);
    if (dgrad(Tr.dqkv, 3 * d, W.w_qkv, d, Tr.dxn)) return -1;                                                             // dxn1
    // ---- ln_1:  dx += LN'(x_in) dxn1
    ln_bwd_kernel<__nv_bfloat16><<<G.sm_count * 2, 256, 2 * d * sizeof(float), st>>>(Tr.xsave + (size_t)l * Tr.m_cap * d, L.ln1_g, Tr.dxn, d, D.ln_eps, M_dev, Tr.dx, 1,
                                                                                      gr.ln1_g, gr.ln1_b);
    e->launches += 16;
    LMRL_CUDA(cudaGetLastError());
    return 0;
}

static int policy_grad_impl(lmrl_engine* e, const int32_t* tokens, const int32_t* n_tok, int32_t pitch, int32_t n, const float* dlogprob, const float* dvalue,
                            float temperature, const lmrl_gpt2_grads* grads, const lmrl_ppo_batch* ppo) {
    LMRL_CHECK(e && e->has_gpt2 && e->gpt2->train && e->gpt2->train->ready, "lmrl_policy_grad: call lmrl_train_enable first");
    LMRL_CHECK(tokens && n_tok && grads && (ppo || (dlogprob && dvalue)), "lmrl_policy_grad: null argument");
    LMRL_CHECK(!ppo || (ppo->should_take_action && ppo->old_logprobs && ppo->old_values && ppo->old_advantages && ppo->old_returns), "lmrl_ppo_grad: null batch array");
    LMRL_CUDA(cudaSetDevice(e->device));
    Gpt2Dev& G = *e->gpt2;
    TrainDev& Tr = *G.train;
    const lmrl_gpt2_dims& D = G.dims;
    const TowerDev& T = G.tw[0];
    const int d = D.d_model, L = D.n_layer;
    const int max_in = e->cfg.max_input_length;
    cudaStream_t st = e->stream;
    // rows in env order; scored rows = positions with a non-zero upstream log-prob gradient
    std::vector<int32_t> lp_rows, lp_targets;
    std::vector<float> lp_g, dv, aux[4];
    int M = 0;
    for (int i = 0; i < n; ++i) {
        LMRL_CHECK(n_tok[i] >= 0 && n_tok[i] <= max_in && n_tok[i] <= pitch, "lmrl_policy_grad: sequence %d has %d tokens (max_input_length %d)", i, n_tok[i], max_in);
        for (int t = 0; t < n_tok[i]; ++t) {
            const size_t o = (size_t)i * pitch + t;
            LMRL_CHECK(tokens[o] >= 0 && tokens[o] < G.v_pad, "lmrl_policy_grad: token id %d out of range", tokens[o]);
            if (!ppo) dv.push_back(dvalue[o]);
            if (t + 1 < n_tok[i] && (ppo ? ppo->should_take_action[o] != 0 : dlogprob[o] != 0.f)) {
                LMRL_CHECK(tokens[o + 1] < D.n_vocab, "lmrl_policy_grad: scored target id %d is a padding id", tokens[o + 1]);
                lp_rows.push_back(M + t); lp_targets.push_back(tokens[o + 1]);
                if (ppo) { aux[0].push_back(ppo->old_logprobs[o]); aux[1].push_back(ppo->old_advantages[o]); aux[2].push_back(ppo->old_values[o]); aux[3].push_back(ppo->old_returns[o]); }
                else lp_g.push_back(dlogprob[o]);
            }
        }
        M += n_tok[i];
    }
    const int n_lp = (int)lp_rows.size();
    LMRL_CHECK(M > 0 && M <= Tr.m_cap, "lmrl_policy_grad: %d rows exceed the training capacity %d (lmrl_train_enable)", M, Tr.m_cap);
    LMRL_CHECK(n_lp <= Tr.lp_cap, "lmrl_policy_grad: %d scored tokens exceed the capacity %d (lmrl_train_enable)", n_lp, Tr.lp_cap);
    LMRL_CHECK(n <= e->cfg.max_envs, "lmrl_policy_grad: %d sequences exceed engine capacity %d", n, e->cfg.max_envs);
    const int Mp = (M + 63) / 64 * 64;
    if (load_prompts(e, tokens, n_tok, pitch, n)) return -1;
    if (ppo) {
        LMRL_CUDA(cudaMemsetAsync(Tr.dv, 0, sizeof(float) * M, st));
        LMRL_CUDA(cudaMemsetAsync(Tr.loss_acc, 0, sizeof(double) * 8, st));
    } else {
        LMRL_CUDA(cudaMemcpyAsync(Tr.dv, dv.data(), sizeof(float) * M, cudaMemcpyHostToDevice, st));
    }
    if (n_lp) {
        LMRL_CUDA(cudaMemcpyAsync(Tr.lp_rows, lp_rows.data(), sizeof(int32_t) * n_lp, cudaMemcpyHostToDevice, st));
        LMRL_CUDA(cudaMemcpyAsync(Tr.lp_targets, lp_targets.data(), sizeof(int32_t) * n_lp, cudaMemcpyHostToDevice, st));
        if (ppo) {
            for (int k = 0; k < 4; ++k)
                LMRL_CUDA(cudaMemcpyAsync(Tr.lp_aux + (size_t)k * Tr.lp_cap, aux[k].data(), sizeof(float) * n_lp, cudaMemcpyHostToDevice, st));
        } else {
            LMRL_CUDA(cudaMemcpyAsync(Tr.lp_g, lp_g.data(), sizeof(float) * n_lp, cudaMemcpyHostToDevice, st));
        }
    }
    const int call_consts[3] = {n_lp, (int)G.v_pad, grads->n_vocab_rows};
    LMRL_CHECK(grads->n_vocab_rows > 0 && grads->n_vocab_rows <= G.v_pad, "lmrl_policy_grad: grads->n_vocab_rows %d out of range", grads->n_vocab_rows);
    LMRL_CHECK(!grads->vh_w || G.vh[0].has_lin, "lmrl_policy_grad: value-head gradients requested but no LinearHead is loaded in head set 0");
    LMRL_CUDA(cudaMemcpyAsync(Tr.consts + 4, call_consts, sizeof call_consts, cudaMemcpyHostToDevice, st));
    LMRL_CUDA(cudaStreamSynchronize(st));                      // host vectors die at return; the rest is enqueued asynchronously
    const int saved_profile = e->profile;
    e->profile = 0;
    struct Restore { lmrl_engine* e; int p; ~Restore() { e->profile = p; } } restore{e, saved_profile};
    G.cur = &G.lane[0];
    e->cur_pass_slot = -1;
    e->cur_mode = 0;
    const int ew = G.sm_count * 8;
    // LMRL_TRAIN_TIMING=1: CUDA-event timing of the sections, printed to stderr (diagnostics only)
    const bool timing = getenv("LMRL_TRAIN_TIMING") != nullptr;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    auto mark = [&](const char* name) {
        if (!timing) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, st);
        marks.push_back({name, ev});
    };
    mark("start");
    // ---- forward, saving the residual stream before every layer
    LMRL_K(clear_gen_kernel, (n + 255) / 256, 256, 0, st, n, G.gen);
    LMRL_K(build_rows_kernel, 1, 1024, 0, st, 0, 0, n, (const uint8_t*)nullptr, G.stream.tokens, G.stream.len, G.stream.kv_len, G.stream.cap, G.gen, G.rows, G.kvp, -1);
    LMRL_K(fill_prefill_rows_kernel, (n * 32 + 255) / 256, 256, 0, st, G.stream.tokens, G.stream.len, G.stream.kv_len, G.stream.cap, G.gen, G.rows, G.kvp, -1);
    LMRL_K(embed_kernel, ew, 256, 0, st, T.wte, T.wpe, d, G.rows, G.x);
    const size_t xbytes = sizeof(float) * (size_t)M * d, xstride = (size_t)Tr.m_cap * d;
    for (int l = 0; l < L; ++l) {
        LMRL_CUDA(cudaMemcpyAsync(Tr.xsave + (size_t)l * xstride, G.x, xbytes, cudaMemcpyDeviceToDevice, st));
        if (train_layer_forward(e, l, n, true)) return -1;
    }
    LMRL_CUDA(cudaMemcpyAsync(Tr.xsave + (size_t)L * xstride, G.x, xbytes, cudaMemcpyDeviceToDevice, st));
    LMRL_CUDA(cudaMemsetAsync(Tr.dfx, 0, xbytes, st));
    mark("forward");
    // ---- tied LM head: dlogits = g (onehot - softmax) on the scored rows
    if (n_lp) {
        const int* n_lp_dev = Tr.consts + 4;
        const int Kp = (n_lp + 63) / 64 * 64, n_lp_pad = (n_lp + 127) / 128 * 128;
        LMRL_K(layernorm_kernel, ew, 256, 0, st, G.x, T.lnf_g, T.lnf_b, d, D.ln_eps, n_lp_dev, (const int*)Tr.lp_rows, Tr.dxn, d);       // dxn := LN_f of the scored rows
        LMRL_CUDA(cudaMemsetAsync(Tr.best, 0, sizeof(unsigned long long) * n_lp, st));
        GemmArgs h{};
        h.A = Tr.dxn; h.W = T.wte; h.bias = nullptr; h.N = G.v_pad; h.K = d; h.M_dev = n_lp_dev; h.best = Tr.best; h.row_key = Tr.row_key; h.inv_temp = 1.0f;
        h.do_sample = 0; h.n_vocab = D.n_vocab; h.logits_out = Tr.lp_logits;
        if (launch_gemm<EPI_SAMPLE>(e, h, D.n_vocab, 1, (size_t)Tr.m_cap, G.v_pad, 4)) return -1;
        const float inv_temp = 1.0f / (temperature > 0.f ? temperature : 1.0f);
        if (ppo)
            dlogits_ppo_kernel<<<n_lp_pad, 256, 0, st>>>(Tr.lp_logits, D.n_vocab, G.v_pad, Tr.lp_targets, Tr.lp_aux, Tr.lp_aux + Tr.lp_cap, n_lp_dev,
                                                         (float)(1.0 - ppo->cliprange), (float)(1.0 + ppo->cliprange), Tr.dl, Tr.loss_acc);
        else
            dlogits_kernel<<<n_lp_pad, 256, 0, st>>>(Tr.lp_logits, D.n_vocab, G.v_pad, Tr.lp_targets, Tr.lp_g, n_lp_dev, inv_temp, Tr.dl);
        // dgrad: d(LN_f out)[scored rows] = dl . wte      (fp32, then scattered into dfx)
        LMRL_CUDA(cudaMemsetAsync(Tr.dxs, 0, sizeof(float) * (size_t)n_lp_pad * d, st));
        GemmArgs g{};
        g.A = Tr.dl; g.W = Tr.wte_t; g.bias = nullptr; g.N = d; g.K = G.v_pad; g.M_dev = n_lp_dev; g.resid = Tr.dxs;
        if (launch_gemm<EPI_RESID>(e, g, d, 0, (size_t)Tr.lp_cap, (size_t)d, 1)) return -1;
        scatter_add_rows_kernel<<<ew, 256, 0, st>>>(Tr.dxs, Tr.lp_rows, n_lp_dev, d, Tr.dfx);
        // wgrad: dwte += dl^T . LN_f(x)
        transpose_to_bf16_rows_kernel<__nv_bfloat16><<<dim3((Kp + 31) / 32, (G.v_pad + 31) / 32), dim3(32, 8), 0, st>>>(Tr.dl, n_lp_dev, G.v_pad, Tr.tA, Kp, nullptr);
        transpose_to_bf16_rows_kernel<__nv_bfloat16><<<dim3((Kp + 31) / 32, (d + 31) / 32), dim3(32, 8), 0, st>>>(Tr.dxn, n_lp_dev, d, Tr.tB, Kp, nullptr);
        GemmArgs wg{};
        wg.A = Tr.tA; wg.W = Tr.tB; wg.bias = nullptr; wg.N = d; wg.K = Kp; wg.M_dev = Tr.consts + 6; wg.resid = grads->wte;
        if (launch_gemm<EPI_RESID>(e, wg, d, 0, (size_t)grads->n_vocab_rows, (size_t)d, 1)) return -1;
    }
    mark("lm_head_backward");
    // ---- value head (LinearHead on the fp32 final LayerNorm), then ln_f
    if (ppo && n_lp) {                               // values of every row, then the value part of the loss on the scored rows
        ValueHeadArgs va{};
        va.slot_q1 = va.slot_q2 = va.slot_v = -1; va.targets = G.rows.row_tok;
        va.x = G.x; va.lnf_g = T.lnf_g; va.lnf_b = T.lnf_b; va.lin_w = G.vh[0].lin_w; va.lin_b = G.vh[0].lin_b; va.d = d; va.eps = D.ln_eps; va.has_lin = 1;
        va.v_out = Tr.values;
        LMRL_K(value_heads_kernel, std::max(1, std::min((M + 7) / 8, G.sm_count * 8)), 256, 0, st, G.rows.M, va);
        ppo_value_grad_kernel<<<(n_lp + 255) / 256, 256, 0, st>>>(Tr.values, Tr.lp_rows, Tr.lp_aux + 2 * (size_t)Tr.lp_cap, Tr.lp_aux + 3 * (size_t)Tr.lp_cap, Tr.consts + 4,
                                                                  (float)ppo->cliprange_value, (float)ppo->value_loss_coef, Tr.dv, Tr.loss_acc);
    }
    if (grads->vh_w)
        value_head_bwd_kernel<<<G.sm_count * 2, 256, (d + 1) * sizeof(float), st>>>(G.x, T.lnf_g, T.lnf_b, G.vh[0].lin_w, Tr.dv, d, D.ln_eps, G.rows.M, Tr.dfx,
                                                                                     grads->vh_w, grads->vh_b);
    ln_bwd_kernel<float><<<G.sm_count * 2, 256, 2 * d * sizeof(float), st>>>(G.x, T.lnf_g, Tr.dfx, d, D.ln_eps, G.rows.M, Tr.dx, 0, grads->lnf_g, grads->lnf_b);
    // ---- decoder layers, last to first: recompute the layer from its saved input, then its backward
    for (int l = L - 1; l >= 0; --l) {
        LMRL_CUDA(cudaMemcpyAsync(G.x, Tr.xsave + (size_t)l * xstride, xbytes, cudaMemcpyDeviceToDevice, st));
        if (train_layer_forward(e, l, n, false)) return -1;
        if (l == L - 1) mark("recompute_last_layer");
        const LayerGrads gr{grads->ln1_g[l], grads->ln1_b[l], grads->attn_w[l], grads->attn_b[l], grads->proj_w[l], grads->proj_b[l],
                            grads->ln2_g[l], grads->ln2_b[l], grads->fc_w[l], grads->fc_b[l], grads->fc2_w[l], grads->fc2_b[l]};
        if (train_layer_backward(e, l, n, Mp, gr)) return -1;
        if (l == L - 1) mark("backward_last_layer");
    }
    mark("layers_backward");
    embed_bwd_kernel<<<ew, 256, 0, st>>>(Tr.dx, G.rows, d, grads->wte, grads->wpe);
    mark("embedding_backward");
    LMRL_CUDA(cudaGetLastError());
    LMRL_CUDA(cudaStreamSynchronize(st));
    if (timing) {
        fprintf(stderr, "[lmrl_policy_grad] rows=%d scored=%d:", M, n_lp);
        for (size_t i = 1; i < marks.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
            fprintf(stderr, " %s %.2f ms", marks[i].first, ms);
        }
        fprintf(stderr, "\n");
        for (auto& m : marks) cudaEventDestroy(m.second);
    }
    if (ppo && ppo->loss_out) {
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        LMRL_CUDA(cudaMemcpy(acc, Tr.loss_acc, sizeof acc, cudaMemcpyDeviceToHost));
        const double cnt = n_lp > 0 ? (double)n_lp : 1.0;
        const double pg = acc[0] / cnt, vf = 0.5 * acc[4] / cnt;
        double* o = ppo->loss_out;
        o[0] = pg + ppo->value_loss_coef * vf; o[1] = pg; o[2] = vf; o[3] = acc[6] / cnt; o[4] = acc[5] / cnt; o[5] = acc[1] / cnt; o[6] = acc[2] / cnt;
        o[7] = acc[3] / cnt; o[8] = cnt / std::max(1.0, (double)M); o[9] = acc[7] / cnt; o[10] = 0.0; o[11] = 0.0; o[12] = (double)n_lp;
        o[13] = o[14] = o[15] = 0.0;
    }
    return 0;
}

extern "C" {

int lmrl_policy_grad_host(lmrl_engine* e, const int32_t* tokens, const int32_t* n_tok, int32_t pitch, int32_t n, const float* dlogprob, const float* dvalue,
                          float temperature, const lmrl_gpt2_grads* grads) {
    return policy_grad_impl(e, tokens, n_tok, pitch, n, dlogprob, dvalue, temperature, grads, nullptr);
}

int lmrl_ppo_grad_host(lmrl_engine* e, const int32_t* tokens, const int32_t* n_tok, int32_t pitch, int32_t n, const lmrl_ppo_batch* batch, const lmrl_gpt2_grads* grads) {
    LMRL_CHECK(batch, "lmrl_ppo_grad: null batch");
    LMRL_CHECK(e && e->gpt2 && e->gpt2->vh[0].has_lin, "lmrl_ppo_grad: the LinearHead value head (head set 0) is not loaded");
    return policy_grad_impl(e, tokens, n_tok, pitch, n, nullptr, nullptr, 1.0f, grads, batch);
}

}  // extern "C"
