// gpt2_kernels.cuh — the non-GEMM kernels of the GPT-2 KV-cache decoder over a RAGGED batch of rows.
//
// A "row" is one (environment, position) pair that needs a forward pass this launch: during prefill an env
// contributes its not-yet-cached stream tokens, during decode one row per generating env.  Row counts live in
// device memory (RowsDev::M) so a whole rollout is enqueued without host round trips.
//
// Replaces, per SURVEY.md §2.2: K1 (wte[ids]+wpe[pos], ln_1), K3 (causal attention with left-pad position ids,
// value_rl_base/gpt2/generation.py:152-171 — here positions are simply the index in the env's own stream),
// the LayerNorms of K5/K7, and the HF sampler bookkeeping (EOS / max_new_tokens).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include "gemm_mma.cuh"

namespace lmrl {

struct RowsDev {
    int* M;            // [1] live rows this pass
    int* n_slots;      // [1] envs contributing rows this pass
    int* row_env;      // [Mcap]
    int* row_pos;      // [Mcap]
    int* row_tok;      // [Mcap]
    int* slot_env;     // [B]   compact list of contributing envs
    int* slot_row0;    // [B]   first row of the slot
    int* slot_nrow;    // [B]
    int* slot_last;    // [B]   row index of the slot's last row (the one that is sampled from)
    int* ctx_total;    // [1]   sum over slots of (last position + 1): cached key positions attention must read this pass
};

struct GenDev {
    int* tok;          // [B][max_new] sampled ids this turn
    int* len;          // [B]
    uint8_t* active;   // [B] still generating this turn
    int* cur;          // [B] tokens whose KV is (about to be) valid == position of the next fed token
    int max_new;
    int eos_id;
};

// ---- paged KV cache ---------------------------------------------------------------------------------------------------
// The K / V of one layer live in a pool of pages of KV_PAGE positions, [n_pages][n_head][KV_PAGE][64] bf16 each for K and for V
// (4 KB contiguous per (page, head)): position p of env e is row (p % KV_PAGE) of page table[e][p / KV_PAGE].  ONE table and one
// allocator serve every layer and every tower (they all cache the same (env, position) pairs).  Pages are handed out on the
// device when the row builder assigns a position whose page is still unset (bump counter); an env keeps its pages when its
// stream rolls back (kv_len rollback after an action, LEFT truncation) and the whole table is cleared when a new batch starts, so
// memory follows the positions actually reached instead of max_input_length + max_new_tokens per env (reference cache shape:
// value_rl_base/gpt2/generation.py:123-150, a dense [B, max_length] slab per layer).
constexpr int KV_PAGE = 32, KV_PAGE_SHIFT = 5;
struct KvPages {
    int* table;        // [B][ppe]  page id or -1
    int* n_alloc;      // [1] pages handed out since the last reset
    int* overflow;     // [1] set when the pool ran out (the host turns it into an error)
    int ppe;           // pages per env = ceil(t_cap / KV_PAGE)  (<= 32: a warp holds one env's table, one entry per lane)
    int n_pages;       // pool capacity
};
__device__ __forceinline__ size_t kv_row_off(int page, int n_head, int head, int pos) {
    return (((size_t)page * n_head + head) * KV_PAGE + (pos & (KV_PAGE - 1))) * 64;
}
__device__ __forceinline__ void kv_ensure_page(const KvPages& kv, int e, int pg) {
    int* slot = kv.table + (size_t)e * kv.ppe + pg;
    if (*slot < 0) {
        int id = atomicAdd(kv.n_alloc, 1);
        if (id >= kv.n_pages) { *kv.overflow = 1; id = 0; }
        *slot = id;
    }
}
__global__ void kv_reset_kernel(KvPages kv, int n_env) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_env * kv.ppe) kv.table[i] = -1;
    if (i == 0) { *kv.n_alloc = 0; *kv.overflow = 0; }
}

// Joined-history tokenisation (ppo/gpt2/interface.py:519-522 -> value_rl_base/base_interface.py:111-112: the reference encodes the
// JOINED history string every turn).  The token stream is kept per Text item (append-only, the reference's train-time convention,
// environment.py:313-322); with GPT-2's pre-tokeniser and merges the two differ in exactly one place for env text: a prompt that ENDS in
// a run of k >= 2 newlines (an action line followed by the empty feedback line of an invalid guess; k = 3 when the action itself was
// empty) is ONE pre-token there, which the merge 'Ċ'+'Ċ' -> 'ĊĊ' (id 628 in GPT-2) turns into floor(k / 2) '\n\n' tokens followed by
// k % 2 plain newlines, while the same newlines in the middle of a prompt stay single tokens (`\s+(?!\S)` backs off by one before
// a non-space).  So the fold is a property of the prompt VIEW: positions before the trailing run are the stream's, the run is fed
// folded, and because the run's K / V (and its first newline's, possibly cached as a plain token by an earlier turn) belong to this
// turn's view only, the cache is valid up to the run's start afterwards.
struct PromptView { int kv; int len; int run_start; int n_pairs; };   // rows cover positions [kv, len); tokens at >= run_start are folded
__device__ __forceinline__ PromptView prompt_view(const int32_t* __restrict__ tok, int len, int kv_len, int nl_id, int nlnl_id) {
    PromptView v{kv_len, len, len, 0};
    if (nlnl_id < 0 || len < 2 || tok[len - 1] != nl_id || tok[len - 2] != nl_id) return v;
    int rs = len - 2;
    while (rs > 0 && tok[rs - 1] == nl_id) --rs;
    const int k = len - rs;
    v.run_start = rs; v.n_pairs = k >> 1;
    v.len = rs + (k >> 1) + (k & 1);
    v.kv = min(kv_len, rs);
    return v;
}
__device__ __forceinline__ int prompt_view_token(const PromptView& v, const int32_t* __restrict__ tok, int pos, int nl_id, int nlnl_id) {
    return pos < v.run_start ? tok[pos] : (pos - v.run_start < v.n_pairs ? nlnl_id : nl_id);
}

// ---- row builder: exclusive scan over envs (single CTA; B <= 16384) ---------------------------------------------
// mode 0 (prefill): env e contributes stream tokens [kv_len[e], len[e]) if it is not done
// mode 1 (decode) : env e contributes its last sampled token at position cur[e] if it is still generating
// e0 / B: the env range of the calling lane (the whole batch: 0 / n)
__global__ void build_rows_kernel(int mode, int e0, int B, const uint8_t* __restrict__ done, const int32_t* __restrict__ stream_tok,
                                  const int32_t* __restrict__ stream_len, const int32_t* __restrict__ kv_len, int stream_cap,
                                  GenDev gen, RowsDev rows, KvPages kv, int nlnl_id) {
    pdl_enter();
    __shared__ int warp_sums[2][32];
    __shared__ int carry[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    if (tid == 0) { carry[0] = 0; carry[1] = 0; *rows.ctx_total = 0; }
    __syncthreads();
    int ctx = 0;
    for (int base = 0; base < B; base += blockDim.x) {          // envs [e0, e0 + B): the lane's share of the batch
        const int e = e0 + base + tid;
        int n = 0;
        if (base + tid < B) {
            if (mode == 0) {
                const PromptView v = prompt_view(stream_tok + (size_t)e * stream_cap, stream_len[e], kv_len[e], gen.eos_id, nlnl_id);
                n = (done && done[e]) ? 0 : max(v.len - v.kv, 0);
                if (n > 0) ctx += v.len;
            }
            else { n = gen.active[e] ? 1 : 0; if (n > 0) ctx += gen.cur[e] + 1; }
        }
        const int has = n > 0;
        int xs = n, xh = has;      // inclusive warp scans of rows and of slots
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xFFFFFFFFu, xs, o), b = __shfl_up_sync(0xFFFFFFFFu, xh, o);
            if (lane >= o) { xs += a; xh += b; }
        }
        if (lane == 31) { warp_sums[0][warp] = xs; warp_sums[1][warp] = xh; }
        __syncthreads();
        if (warp == 0) {
            int a = lane < nwarp ? warp_sums[0][lane] : 0, b = lane < nwarp ? warp_sums[1][lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ua = __shfl_up_sync(0xFFFFFFFFu, a, o), ub = __shfl_up_sync(0xFFFFFFFFu, b, o);
                if (lane >= o) { a += ua; b += ub; }
            }
            warp_sums[0][lane] = a; warp_sums[1][lane] = b;     // inclusive over warps
        }
        __syncthreads();
        const int row0 = carry[0] + (warp ? warp_sums[0][warp - 1] : 0) + xs - n;
        const int slot = carry[1] + (warp ? warp_sums[1][warp - 1] : 0) + xh - has;
        if (has) {
            rows.slot_env[slot] = e; rows.slot_row0[slot] = row0; rows.slot_nrow[slot] = n; rows.slot_last[slot] = row0 + n - 1;
            if (mode == 1) {
                const int p = gen.cur[e];
                rows.row_env[row0] = e; rows.row_pos[row0] = p; rows.row_tok[row0] = gen.tok[e * gen.max_new + gen.len[e] - 1];
                gen.cur[e] = p + 1;
                kv_ensure_page(kv, e, p >> KV_PAGE_SHIFT);                  // the page this row's K / V go to
            }
        }
        __syncthreads();
        if (tid == 0) { carry[0] += warp_sums[0][nwarp - 1]; carry[1] += warp_sums[1][nwarp - 1]; }
        __syncthreads();
    }
    if (tid == 0) { *rows.M = carry[0]; *rows.n_slots = carry[1]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) ctx += __shfl_xor_sync(0xFFFFFFFFu, ctx, o);
    if (lane == 0 && ctx) atomicAdd(rows.ctx_total, ctx);
}

// prefill: expand each slot into its rows (one warp per slot); also resets the per-turn generation state
__global__ void fill_prefill_rows_kernel(const int32_t* __restrict__ stream_tok, const int32_t* __restrict__ stream_len,
                                         int32_t* __restrict__ kv_len, int stream_cap, GenDev gen, RowsDev rows, KvPages kv, int nlnl_id) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slot >= *rows.n_slots) return;
    const int e = rows.slot_env[slot], r0 = rows.slot_row0[slot], n = rows.slot_nrow[slot];
    const int32_t* stok = stream_tok + (size_t)e * stream_cap;
    const PromptView v = prompt_view(stok, stream_len[e], kv_len[e], gen.eos_id, nlnl_id);      // the same view build_rows_kernel sized n from
    const int p0 = v.kv;
    {   // pages of positions [p0, p0 + n): one lane per page (ppe <= 32)
        const int pg = (p0 >> KV_PAGE_SHIFT) + lane;
        if (pg <= ((p0 + n - 1) >> KV_PAGE_SHIFT)) kv_ensure_page(kv, e, pg);
    }
    for (int i = lane; i < n; i += 32) {
        rows.row_env[r0 + i] = e; rows.row_pos[r0 + i] = p0 + i;
        rows.row_tok[r0 + i] = prompt_view_token(v, stok, p0 + i, gen.eos_id, nlnl_id);
    }
    // a folded run's K / V belong to this turn's prompt view only: the next prefill starts at the run again
    __syncwarp();                                              // every lane has read kv_len[e] (prompt_view) before it changes
    if (lane == 0) { gen.cur[e] = p0 + n; gen.len[e] = 0; gen.active[e] = 1; kv_len[e] = min(p0 + n, v.run_start); }
}

__global__ void clear_gen_kernel(int B, GenDev gen) {
    pdl_enter();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < B) { gen.active[e] = 0; gen.len[e] = 0; }
}

// ---- K1: x = wte[tok] + wpe[pos] (fp32 residual stream) ------------------------------------------------------------
__global__ void embed_kernel(const __nv_bfloat16* __restrict__ wte, const __nv_bfloat16* __restrict__ wpe, int d, RowsDev rows,
                             float* __restrict__ x) {
    pdl_enter();
    const int M = *rows.M;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < M; r += warps) {
        const __nv_bfloat16* a = wte + (size_t)rows.row_tok[r] * d;
        const __nv_bfloat16* b = wpe + (size_t)rows.row_pos[r] * d;
        for (int c = lane * 8; c < d; c += 256) {
            const uint4 ua = *reinterpret_cast<const uint4*>(a + c), ub = *reinterpret_cast<const uint4*>(b + c);
            const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&ua);
            const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&ub);
            float o[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 fa = __bfloat1622float2(pa[q]), fb = __bfloat1622float2(pb[q]);
                o[2 * q] = fa.x + fb.x; o[2 * q + 1] = fa.y + fb.y;
            }
            float4* dst = reinterpret_cast<float4*>(x + (size_t)r * d + c);
            dst[0] = make_float4(o[0], o[1], o[2], o[3]);
            dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
}

// ---- LayerNorm: fp32 row -> bf16 row; one warp per row, d <= 1024, d % 128 == 0 ------------------------------------
// gather != nullptr: output row j takes input row gather[j] (final LN of the sampled rows only), count from n_dev
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 int d, float eps, const int* __restrict__ n_dev, const int* __restrict__ gather,
                                 __nv_bfloat16* __restrict__ out, int out_pitch) {
    pdl_enter();
    const int n = *n_dev;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int nv = d >> 7;                 // float4 per lane
    for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += warps) {
        const int r = gather ? gather[j] : j;
        const float4* src = reinterpret_cast<const float4*>(x + (size_t)r * d);
        float4 v[8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < nv) { v[i] = src[lane + 32 * i]; s += v[i].x + v[i].y + v[i].z + v[i].w; }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        const float mean = s / d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < nv) {
                const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
                q += a * a + b * b + c * c + e * e;
            }
#pragma unroll
        for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
        const float rstd = rsqrtf(q / d + eps);
        __nv_bfloat16* dst = out + (size_t)j * out_pitch;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < nv) {
                const int c = (lane + 32 * i) * 4;
                const float4 g4 = *reinterpret_cast<const float4*>(gamma + c), b4 = *reinterpret_cast<const float4*>(beta + c);
                const __nv_bfloat162 lo = __floats2bfloat162_rn((v[i].x - mean) * rstd * g4.x + b4.x, (v[i].y - mean) * rstd * g4.y + b4.y);
                const __nv_bfloat162 hi = __floats2bfloat162_rn((v[i].z - mean) * rstd * g4.z + b4.z, (v[i].w - mean) * rstd * g4.w + b4.w);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(dst + c) = pk;
            }
    }
}

// ---- value heads of the per-position forward (lmrl_forward_host): one warp per row ---------------------------------
// Q heads: q_k(h_t)[a] with a = the NEXT token of the sequence — dense2 is never applied to the whole vocabulary, the warp
// reads the one dense2 column it needs (row `a` of the transposed [v_pad][hidden] bf16 matrix, 1.5 KB) and folds it with the
// row's relu(dense1) (bf16, written by the ReLU-epilogue GEMM).  V head: MLPHead with output_dim 1 (same fold with its single
// dense2 column), or LinearHead on the final-LN hidden, recomputed here in fp32 from the residual stream (ppo value head).
// HBM-bound: per row (n_heads + 1) * hidden * 2 B of operands -> 4..12 B of results.
struct ValueHeadArgs {
    const __nv_bfloat16* fr; int fr_pitch; int hidden;
    const __nv_bfloat16* fx; int q1_lin, q2_lin;      // LinearHead Q heads: q_k = fx_row . w2q_k[target] + b (fx = final-LN hidden, [M][d])
    int slot_q1, slot_q2, slot_v;
    const __nv_bfloat16* w2q1; const float* b2q1;
    const __nv_bfloat16* w2q2; const float* b2q2;
    const float* w2v; float b2v;
    const int* targets;
    const float* x; const float* lnf_g; const float* lnf_b; const float* lin_w; float lin_b; int d; float eps; int has_lin;
    float* q1_out; float* q2_out; float* v_out;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

__device__ __forceinline__ float dot_bf16_bf16(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, int n, int lane) {
    float acc = 0.f;
    for (int c = lane * 8; c < n; c += 256) {
        const uint4 ua = *reinterpret_cast<const uint4*>(a + c), ub = *reinterpret_cast<const uint4*>(b + c);
        const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&ua);
        const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&ub);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 fa = __bfloat1622float2(pa[q]), fb = __bfloat1622float2(pb[q]);
            acc = fmaf(fa.x, fb.x, acc); acc = fmaf(fa.y, fb.y, acc);
        }
    }
    return warp_sum(acc);
}

__global__ void value_heads_kernel(const int* __restrict__ n_dev, ValueHeadArgs a) {
    pdl_enter();
    const int n = *n_dev;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
        const int tgt = a.targets[r];
        const __nv_bfloat16* rr = a.fr + (size_t)r * a.fr_pitch;
        if (a.slot_q1 >= 0 || a.q1_lin) {
            const float q = dot_bf16_bf16(a.q1_lin ? a.fx + (size_t)r * a.d : rr + a.slot_q1 * a.hidden, a.w2q1 + (size_t)tgt * a.hidden, a.hidden, lane);
            if (lane == 0) a.q1_out[r] = q + a.b2q1[tgt];
        }
        if (a.slot_q2 >= 0 || a.q2_lin) {
            const float q = dot_bf16_bf16(a.q2_lin ? a.fx + (size_t)r * a.d : rr + a.slot_q2 * a.hidden, a.w2q2 + (size_t)tgt * a.hidden, a.hidden, lane);
            if (lane == 0) a.q2_out[r] = q + a.b2q2[tgt];
        }
        if (a.slot_v >= 0) {
            const __nv_bfloat16* rv = rr + a.slot_v * a.hidden;
            float acc = 0.f;
            for (int c = lane; c < a.hidden; c += 32) acc = fmaf(__bfloat162float(rv[c]), a.w2v[c], acc);
            acc = warp_sum(acc);
            if (lane == 0) a.v_out[r] = acc + a.b2v;
        } else if (a.has_lin) {
            const float* xr = a.x + (size_t)r * a.d;
            float s = 0.f;
            for (int c = lane; c < a.d; c += 32) s += xr[c];
            const float mean = warp_sum(s) / a.d;
            float q = 0.f;
            for (int c = lane; c < a.d; c += 32) { const float t = xr[c] - mean; q = fmaf(t, t, q); }
            const float rstd = rsqrtf(warp_sum(q) / a.d + a.eps);
            float acc = 0.f;
            for (int c = lane; c < a.d; c += 32) acc = fmaf((xr[c] - mean) * rstd * a.lnf_g[c] + a.lnf_b[c], a.lin_w[c], acc);
            acc = warp_sum(acc);
            if (lane == 0) a.v_out[r] = acc + a.lin_b;
        }
    }
}

// ---- K3: causal attention over the slab KV cache, one WARP per (slot, head) ----------------------------------------
// HBM-bound batched GEMV (~93 % of the decode bytes).  A warp streams the (env, head) K and V slabs
// straight from global memory with 16-byte loads: lane = (key group g = lane/8, dim chunk s = lane%8), i.e. every
// load instruction of the warp covers 4 consecutive keys x 128 B = 512 contiguous bytes.  ONE pass with an online
// soft-max: per step of 16 keys a lane has 4 K and 4 V loads in flight together, the 8 lanes of a key group fold the
// dot product with shuffles, every key group keeps its own running (max, sum, P·V accumulator) and the four groups
// are merged once at the end (flash-decoding style).  No shared memory.  During prefill the warp loops over the
// slot's rows (K/V re-reads hit L1/L2); the tensor-core kernel below is the one used for prefill passes.
constexpr int ATT_WARPS = 8;

__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float f[8]) {
    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xFFFF0000u);
    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xFFFF0000u);
    f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xFFFF0000u);
    f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xFFFF0000u);
}

__global__ void __launch_bounds__(ATT_WARPS * 32, 3) attention_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kcache,
                                                                      const __nv_bfloat16* __restrict__ vcache, RowsDev rows, int n_head,
                                                                      int d_model, KvPages kv, float scale, __nv_bfloat16* __restrict__ out) {
    pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, s = lane & 7;
    const int total = *rows.n_slots * n_head;
    for (int w = blockIdx.x * ATT_WARPS + warp; w < total; w += gridDim.x * ATT_WARPS) {
        const int slot = w / n_head, head = w - slot * n_head;
        const int e = rows.slot_env[slot], r0 = rows.slot_row0[slot], n = rows.slot_nrow[slot];
        // this env's page table, one entry per lane; a step of 16 keys lies inside one page (KV_PAGE = 32)
        const int my_page = lane < kv.ppe ? kv.table[(size_t)e * kv.ppe + lane] : 0;
        const uint4* gK = reinterpret_cast<const uint4*>(kcache) + s;       // + kv_row_off(page, head, j) / 8 per key
        const uint4* gV = reinterpret_cast<const uint4*>(vcache) + s;
        for (int i = 0; i < n; ++i) {
            const int r = r0 + i, nkeys = rows.row_pos[r] + 1;
            float qf[8];
            {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(q + (size_t)r * d_model + head * 64) + s);
                bf16x8_to_float(u, qf);
#pragma unroll
                for (int t = 0; t < 8; ++t) qf[t] *= scale;
            }
            float m = -INFINITY, l = 0.f;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int j0 = 0; j0 < nkeys; j0 += 16) {
                uint4 kk[4], vv[4];
                const int page = __shfl_sync(0xFFFFFFFFu, my_page, j0 >> KV_PAGE_SHIFT);
                const size_t pbase = (((size_t)page * n_head + head) * KV_PAGE + (j0 & (KV_PAGE - 1))) * 8;      // in uint4 units (8 per key)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 4 * u + g;
                    const bool ok = j < nkeys;
                    kk[u] = ok ? __ldg(gK + pbase + (size_t)(4 * u + g) * 8) : make_uint4(0, 0, 0, 0);
                    vv[u] = ok ? __ldg(gV + pbase + (size_t)(4 * u + g) * 8) : make_uint4(0, 0, 0, 0);
                }
                float sc[4];
                float cm = -INFINITY;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float kf[8];
                    bf16x8_to_float(kk[u], kf);
                    float dot = 0.f;
#pragma unroll
                    for (int t = 0; t < 8; ++t) dot = fmaf(qf[t], kf[t], dot);
                    dot += __shfl_xor_sync(0xFFFFFFFFu, dot, 1);
                    dot += __shfl_xor_sync(0xFFFFFFFFu, dot, 2);
                    dot += __shfl_xor_sync(0xFFFFFFFFu, dot, 4);
                    sc[u] = (j0 + 4 * u + g < nkeys) ? dot : -INFINITY;
                    cm = fmaxf(cm, sc[u]);
                }
                if (cm > -INFINITY) {                                   // uniform inside a key group
                    const float m_new = fmaxf(m, cm);
                    const float corr = __expf(m - m_new);               // m = -inf on the first step: exp(-inf) = 0
                    l *= corr;
#pragma unroll
                    for (int t = 0; t < 8; ++t) acc[t] *= corr;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float p = __expf(sc[u] - m_new);          // masked keys: exp(-inf) = 0
                        float vf[8];
                        bf16x8_to_float(vv[u], vf);
                        l += p;
#pragma unroll
                        for (int t = 0; t < 8; ++t) acc[t] = fmaf(p, vf[t], acc[t]);
                    }
                    m = m_new;
                }
            }
            // merge the four key groups (lanes differing in bits 3 and 4 hold the same dims)
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
                const float m_o = __shfl_xor_sync(0xFFFFFFFFu, m, o), l_o = __shfl_xor_sync(0xFFFFFFFFu, l, o);
                const float m_n = fmaxf(m, m_o);                        // finite: key 0 always exists
                const float c1 = m > -INFINITY ? __expf(m - m_n) : 0.f, c2 = m_o > -INFINITY ? __expf(m_o - m_n) : 0.f;
                l = l * c1 + l_o * c2;
#pragma unroll
                for (int t = 0; t < 8; ++t) acc[t] = acc[t] * c1 + __shfl_xor_sync(0xFFFFFFFFu, acc[t], o) * c2;
                m = m_n;
            }
            if (g == 0) {
                const float inv = 1.0f / l;
                __nv_bfloat162 o2[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) o2[t] = __floats2bfloat162_rn(acc[2 * t] * inv, acc[2 * t + 1] * inv);
                *reinterpret_cast<uint4*>(out + (size_t)r * d_model + head * 64 + s * 8) = *reinterpret_cast<const uint4*>(o2);
            }
        }
    }
}

// ---- K3, software-pipelined variant ---------------------------------------------------------------------------------------
// Same arithmetic as attention_kernel (bit-identical results), but the K / V loads of the NEXT 16-key step are issued before the current
// step is consumed (two register buffers), so an item's dependent memory round trips drop from ceil(ctx / 16) to about half: at the
// contexts of a Wordle rollout (30-110 keys) the streaming kernel is bound by those round trips per (slot, head), not by bandwidth.
__device__ __forceinline__ void att_load_step(const uint4* __restrict__ gK, const uint4* __restrict__ gV, int my_page, int n_head, int head, int j0,
                                              int g, int nkeys, uint4 (&kk)[4], uint4 (&vv)[4]) {
    const int page = __shfl_sync(0xFFFFFFFFu, my_page, j0 >> KV_PAGE_SHIFT);
    const size_t pbase = (((size_t)page * n_head + head) * KV_PAGE + (j0 & (KV_PAGE - 1))) * 8;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const bool ok = j0 + 4 * u + g < nkeys;
        kk[u] = ok ? __ldg(gK + pbase + (size_t)(4 * u + g) * 8) : make_uint4(0, 0, 0, 0);
        vv[u] = ok ? __ldg(gV + pbase + (size_t)(4 * u + g) * 8) : make_uint4(0, 0, 0, 0);
    }
}
__device__ __forceinline__ void att_consume_step(const uint4 (&kk)[4], const uint4 (&vv)[4], const float (&qf)[8], int j0, int g, int nkeys,
                                                 float& m, float& l, float (&acc)[8]) {
    float sc[4];
    float cm = -INFINITY;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        float kf[8];
        bf16x8_to_float(kk[u], kf);
        float dot = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t) dot = fmaf(qf[t], kf[t], dot);
        dot += __shfl_xor_sync(0xFFFFFFFFu, dot, 1);
        dot += __shfl_xor_sync(0xFFFFFFFFu, dot, 2);
        dot += __shfl_xor_sync(0xFFFFFFFFu, dot, 4);
        sc[u] = (j0 + 4 * u + g < nkeys) ? dot : -INFINITY;
        cm = fmaxf(cm, sc[u]);
    }
    if (cm > -INFINITY) {
        const float m_new = fmaxf(m, cm);
        const float corr = __expf(m - m_new);
        l *= corr;
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] *= corr;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float p = __expf(sc[u] - m_new);
            float vf[8];
            bf16x8_to_float(vv[u], vf);
            l += p;
#pragma unroll
            for (int t = 0; t < 8; ++t) acc[t] = fmaf(p, vf[t], acc[t]);
        }
        m = m_new;
    }
}

__global__ void __launch_bounds__(ATT_WARPS * 32, 2) attention_pipe_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kcache,
                                                                           const __nv_bfloat16* __restrict__ vcache, RowsDev rows, int n_head,
                                                                           int d_model, KvPages kv, float scale, __nv_bfloat16* __restrict__ out) {
    pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, s = lane & 7;
    const int total = *rows.n_slots * n_head;
    const uint4* gK = reinterpret_cast<const uint4*>(kcache) + s;
    const uint4* gV = reinterpret_cast<const uint4*>(vcache) + s;
    for (int w = blockIdx.x * ATT_WARPS + warp; w < total; w += gridDim.x * ATT_WARPS) {
        const int slot = w / n_head, head = w - slot * n_head;
        const int e = rows.slot_env[slot], r0 = rows.slot_row0[slot], n = rows.slot_nrow[slot];
        const int my_page = lane < kv.ppe ? kv.table[(size_t)e * kv.ppe + lane] : 0;
        for (int i = 0; i < n; ++i) {
            const int r = r0 + i, nkeys = rows.row_pos[r] + 1;
            uint4 ka[4], va[4], kb[4], vb[4];
            att_load_step(gK, gV, my_page, n_head, head, 0, g, nkeys, ka, va);            // in flight while q arrives
            float qf[8];
            {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(q + (size_t)r * d_model + head * 64) + s);
                bf16x8_to_float(u, qf);
#pragma unroll
                for (int t = 0; t < 8; ++t) qf[t] *= scale;
            }
            float m = -INFINITY, l = 0.f;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int j0 = 0; j0 < nkeys; j0 += 32) {
                const bool has_b = j0 + 16 < nkeys;
                if (has_b) att_load_step(gK, gV, my_page, n_head, head, j0 + 16, g, nkeys, kb, vb);
                att_consume_step(ka, va, qf, j0, g, nkeys, m, l, acc);
                if (has_b) {
                    if (j0 + 32 < nkeys) att_load_step(gK, gV, my_page, n_head, head, j0 + 32, g, nkeys, ka, va);
                    att_consume_step(kb, vb, qf, j0 + 16, g, nkeys, m, l, acc);
                }
            }
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
                const float m_o = __shfl_xor_sync(0xFFFFFFFFu, m, o), l_o = __shfl_xor_sync(0xFFFFFFFFu, l, o);
                const float m_n = fmaxf(m, m_o);
                const float c1 = m > -INFINITY ? __expf(m - m_n) : 0.f, c2 = m_o > -INFINITY ? __expf(m_o - m_n) : 0.f;
                l = l * c1 + l_o * c2;
#pragma unroll
                for (int t = 0; t < 8; ++t) acc[t] = acc[t] * c1 + __shfl_xor_sync(0xFFFFFFFFu, acc[t], o) * c2;
                m = m_n;
            }
            if (g == 0) {
                const float inv = 1.0f / l;
                __nv_bfloat162 o2[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) o2[t] = __floats2bfloat162_rn(acc[2 * t] * inv, acc[2 * t + 1] * inv);
                *reinterpret_cast<uint4*>(out + (size_t)r * d_model + head * 64 + s * 8) = *reinterpret_cast<const uint4*>(o2);
            }
        }
    }
}

// ---- K3, several heads per warp ---------------------------------------------------------------------------------------------
// Same arithmetic as attention_kernel (bit-identical results).  A warp takes HPW consecutive heads of ONE slot: the slot's metadata chain
// (slot -> env / first row -> page table / position: two dependent round trips before the first K / V byte can be asked for) is paid once
// per HPW heads instead of once per head, and the first K / V loads of a head are issued before its q is converted.
template <int HPW>
__global__ void __launch_bounds__(ATT_WARPS * 32, 3) attention_hpw_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kcache,
                                                                          const __nv_bfloat16* __restrict__ vcache, RowsDev rows, int n_head,
                                                                          int d_model, KvPages kv, float scale, __nv_bfloat16* __restrict__ out) {
    pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, s = lane & 7;
    const int groups = n_head / HPW;                                       // head groups per slot (n_head % HPW == 0: checked by the host)
    const int total = *rows.n_slots * groups;
    const uint4* gK = reinterpret_cast<const uint4*>(kcache) + s;
    const uint4* gV = reinterpret_cast<const uint4*>(vcache) + s;
    for (int w = blockIdx.x * ATT_WARPS + warp; w < total; w += gridDim.x * ATT_WARPS) {
        const int slot = w / groups, head0 = (w - slot * groups) * HPW;
        const int e = rows.slot_env[slot], r0 = rows.slot_row0[slot], n = rows.slot_nrow[slot];
        const int my_page = lane < kv.ppe ? kv.table[(size_t)e * kv.ppe + lane] : 0;
        for (int i = 0; i < n; ++i) {
            const int r = r0 + i, nkeys = rows.row_pos[r] + 1;
#pragma unroll 1
            for (int hh = 0; hh < HPW; ++hh) {
                const int head = head0 + hh;
                uint4 kk[4], vv[4];
                att_load_step(gK, gV, my_page, n_head, head, 0, g, nkeys, kk, vv);        // in flight while q arrives
                float qf[8];
                {
                    const uint4 u = __ldg(reinterpret_cast<const uint4*>(q + (size_t)r * d_model + head * 64) + s);
                    bf16x8_to_float(u, qf);
#pragma unroll
                    for (int t = 0; t < 8; ++t) qf[t] *= scale;
                }
                float m = -INFINITY, l = 0.f;
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                for (int j0 = 0; j0 < nkeys; j0 += 16) {
                    if (j0 > 0) att_load_step(gK, gV, my_page, n_head, head, j0, g, nkeys, kk, vv);
                    att_consume_step(kk, vv, qf, j0, g, nkeys, m, l, acc);
                }
#pragma unroll
                for (int o = 8; o <= 16; o <<= 1) {
                    const float m_o = __shfl_xor_sync(0xFFFFFFFFu, m, o), l_o = __shfl_xor_sync(0xFFFFFFFFu, l, o);
                    const float m_n = fmaxf(m, m_o);
                    const float c1 = m > -INFINITY ? __expf(m - m_n) : 0.f, c2 = m_o > -INFINITY ? __expf(m_o - m_n) : 0.f;
                    l = l * c1 + l_o * c2;
#pragma unroll
                    for (int t = 0; t < 8; ++t) acc[t] = acc[t] * c1 + __shfl_xor_sync(0xFFFFFFFFu, acc[t], o) * c2;
                    m = m_n;
                }
                if (g == 0) {
                    const float inv = 1.0f / l;
                    __nv_bfloat162 o2[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) o2[t] = __floats2bfloat162_rn(acc[2 * t] * inv, acc[2 * t + 1] * inv);
                    *reinterpret_cast<uint4*>(out + (size_t)r * d_model + head * 64 + s * 8) = *reinterpret_cast<const uint4*>(o2);
                }
            }
        }
    }
}

// ---- K3 (prefill): tensor-core attention for slots that bring many rows ------------------------------------------
// One warp per (slot, head).  MT m16 tiles of query rows (16 or 32 rows of the slot) are held as mma A fragments; the (env, head)
// K / V slabs are staged 32 keys at a time into per-warp shared memory (cp.async, 128-byte rows, XOR-swizzled 16-byte
// chunks), double buffered so the next key block is in flight while the current one is consumed with ldmatrix (K) /
// ldmatrix.trans (V) by mma.sync m16n8k16 with a flash-style online soft-max (fp32 running max / sum, P re-packed from
// the S accumulators as the next A operand).  With 32-row passes (MT = 2, the default) every staged K / V block and every
// ldmatrix fragment feeds both m-tiles, so a Wordle turn's ~20 new rows per episode stream the episode's K / V exactly once
// (round 1: 16-row passes, the second one re-reading K / V through L2 — DRAM traffic 1.25x the algorithmic bytes, 0.55 of HBM).
// Slots with more than 32 new rows take further passes; the passes that will be followed by another one ask L2 to keep the lines.
constexpr int ATTP_WARPS = 12;                                  // 12 x 16 KB of staging, one CTA per SM: 3 warps per SM sub-partition (16 K registers each) -> 168 registers per thread
constexpr int ATTP_KEYS = 32;                                   // keys per staged block
constexpr int ATTP_WARP_SMEM = 2 * 2 * ATTP_KEYS * 128;        // 2 buffers x (K block + V block)
constexpr int ATTP_CTAS_PER_SM = 1;

// cp.async with an L2 eviction-priority hint (see above)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async16_hint(uint32_t dst, const void* src, uint64_t policy) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "l"(policy));
}
// src_bytes = 16: plain copy; 0: nothing is read from global memory and the 16 shared-memory bytes are zero-filled (keys past the
// last visible one: masked anyway, but stale shared memory could hold NaN patterns and 0 * NaN would poison P.V)
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async16_zfill_hint(uint32_t dst, const void* src, uint32_t src_bytes, uint64_t policy) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;\n" ::"r"(dst), "l"(src), "r"(src_bytes), "l"(policy));
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// one pass over the keys for query rows [rb, rb + 16 * MT) of a slot (rows beyond n are computed on zero queries and not stored)
template <int MT>
__device__ __forceinline__ void attp_rows(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ gK, const __nv_bfloat16* __restrict__ gV,
                                          int r0, int n, int rb, int p0, int head, int d_model, int n_head, int my_page, float scale_log2e, bool keep,
                                          uint64_t pol_last, uint32_t sbase, int lane, __nv_bfloat16* __restrict__ out) {
    const int g = lane >> 2, t = lane & 3;
    const int nr = min(16 * MT, n - rb);                   // live query rows in this pass
    const int n_keys = p0 + rb + nr;                       // keys visible to the last live row
    auto stage = [&](int buf, int k0) {                    // 32 keys of K and V = one KV page: 2 x 256 chunks of 16 B, 16 per lane
        const uint32_t sK = sbase + buf * (2 * ATTP_KEYS * 128), sV = sK + ATTP_KEYS * 128;
        const int page = __shfl_sync(0xFFFFFFFFu, my_page, k0 >> KV_PAGE_SHIFT);     // keys past the last live one stay inside the page (masked)
        const size_t pbase = ((size_t)page * n_head + head) * KV_PAGE * 64;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int idx = c * 32 + lane, key = idx >> 3, ch = idx & 7;
            const uint32_t off = key * 128 + ((ch ^ (key & 7)) << 4);
            const uint32_t nb = k0 + key < n_keys ? 16u : 0u;         // the last block of a context: only the visible keys are read (round 1
                                                                      // fetched the whole 32-key block: +16 keys per (env, head) on average)
            if (keep) {
                cp_async16_zfill_hint(sK + off, gK + pbase + (size_t)key * 64 + ch * 8, nb, pol_last);
                cp_async16_zfill_hint(sV + off, gV + pbase + (size_t)key * 64 + ch * 8, nb, pol_last);
            } else {
                cp_async16_zfill(sK + off, gK + pbase + (size_t)key * 64 + ch * 8, nb);
                cp_async16_zfill(sV + off, gV + pbase + (size_t)key * 64 + ch * 8, nb);
            }
        }
        cp_async_commit();
    };
    __syncwarp();
    stage(0, 0);
    // Q fragments: qa[mt][kk][0..3]
    uint32_t qa[MT][4][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int i = rb + mt * 16 + g + hh * 8;
            const uint32_t* src = reinterpret_cast<const uint32_t*>(q + (size_t)(r0 + i) * d_model + head * 64);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                qa[mt][kk][hh] = i < n ? __ldg(src + kk * 8 + t) : 0u;            // dims 16kk + 2t, +1
                qa[mt][kk][hh + 2] = i < n ? __ldg(src + kk * 8 + 4 + t) : 0u;    // dims 16kk + 8 + 2t, +1
            }
        }
    float o[MT][8][4];
    float mrow[MT][2], lrow[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        mrow[mt][0] = mrow[mt][1] = -INFINITY; lrow[mt][0] = lrow[mt][1] = 0.f;
#pragma unroll
        for (int d8 = 0; d8 < 8; ++d8)
#pragma unroll
            for (int c = 0; c < 4; ++c) o[mt][d8][c] = 0.f;
    }
    int buf = 0;
    for (int k0 = 0; k0 < n_keys; k0 += ATTP_KEYS, buf ^= 1) {
        if (k0 + ATTP_KEYS < n_keys) { stage(buf ^ 1, k0 + ATTP_KEYS); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncwarp();
        const uint32_t sK = sbase + buf * (2 * ATTP_KEYS * 128), sV = sK + ATTP_KEYS * 128;
        // S = Q K^T  (4 key n8-blocks); every K fragment feeds all m-tiles
        float sacc[MT][4][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) sacc[mt][j][c] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {               // two k16 steps per ldmatrix.x4
                const int key = j * 8 + (lane & 7), ch = kp * 4 + (lane >> 3);
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(sK + key * 128 + ((ch ^ (key & 7)) << 4), b0, b1, b2, b3);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    mma_bf16(sacc[mt][j], qa[mt][2 * kp], b0, b1);
                    mma_bf16(sacc[mt][j], qa[mt][2 * kp + 1], b2, b3);
                }
            }
        }
        // scale, causal mask, online soft-max
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int pos = p0 + rb + mt * 16 + g + hh * 8;
                float bm = -INFINITY;
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int key = k0 + j * 8 + 2 * t + c;
                        float v = sacc[mt][j][hh * 2 + c] * scale_log2e;
                        v = key <= pos ? v : -INFINITY;
                        sacc[mt][j][hh * 2 + c] = v;
                        bm = fmaxf(bm, v);
                    }
                bm = fmaxf(bm, __shfl_xor_sync(0xFFFFFFFFu, bm, 1));
                bm = fmaxf(bm, __shfl_xor_sync(0xFFFFFFFFu, bm, 2));
                const float m_new = fmaxf(mrow[mt][hh], bm);          // finite: key 0 is visible to every row
                const float alpha = exp2f(mrow[mt][hh] - m_new);
                mrow[mt][hh] = m_new;
                float ls = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float pz = exp2f(sacc[mt][j][hh * 2 + c] - m_new);
                        sacc[mt][j][hh * 2 + c] = pz;
                        ls += pz;
                    }
                lrow[mt][hh] = lrow[mt][hh] * alpha + ls;
#pragma unroll
                for (int d8 = 0; d8 < 8; ++d8) { o[mt][d8][hh * 2] *= alpha; o[mt][d8][hh * 2 + 1] *= alpha; }
            }
        // O += P V   (two k16 key steps x 8 dim n8-blocks); every V fragment feeds all m-tiles
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t pa[MT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                pa[mt][0] = pack_bf16x2(sacc[mt][2 * ks][0], sacc[mt][2 * ks][1]);
                pa[mt][1] = pack_bf16x2(sacc[mt][2 * ks][2], sacc[mt][2 * ks][3]);
                pa[mt][2] = pack_bf16x2(sacc[mt][2 * ks + 1][0], sacc[mt][2 * ks + 1][1]);
                pa[mt][3] = pack_bf16x2(sacc[mt][2 * ks + 1][2], sacc[mt][2 * ks + 1][3]);
            }
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {               // two dim n8-blocks per ldmatrix.x4.trans
                const int key = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ch = dp * 2 + (lane >> 4);
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(sV + key * 128 + ((ch ^ (key & 7)) << 4), b0, b1, b2, b3);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    mma_bf16(o[mt][2 * dp], pa[mt], b0, b1);
                    mma_bf16(o[mt][2 * dp + 1], pa[mt], b2, b3);
                }
            }
        }
        __syncwarp();                                      // everyone is done with this buffer before it is refilled
    }
    // normalise and store
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            float l = lrow[mt][hh];
            l += __shfl_xor_sync(0xFFFFFFFFu, l, 1);
            l += __shfl_xor_sync(0xFFFFFFFFu, l, 2);
            const int i = rb + mt * 16 + g + hh * 8;
            if (i < n) {
                const float inv = 1.0f / l;
                __nv_bfloat16* dst = out + (size_t)(r0 + i) * d_model + head * 64;
#pragma unroll
                for (int d8 = 0; d8 < 8; ++d8)
                    *reinterpret_cast<uint32_t*>(dst + d8 * 8 + 2 * t) = pack_bf16x2(o[mt][d8][hh * 2] * inv, o[mt][d8][hh * 2 + 1] * inv);
            }
        }
}

__global__ void __launch_bounds__(ATTP_WARPS * 32, ATTP_CTAS_PER_SM) attention_prefill_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kcache,
                                                                               const __nv_bfloat16* __restrict__ vcache, RowsDev rows, int n_head,
                                                                               int d_model, KvPages kv, float scale_log2e, int wide,
                                                                               __nv_bfloat16* __restrict__ out) {
    pdl_enter();
    extern __shared__ __align__(128) unsigned char attp_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t sbase = smem_u32(attp_smem) + warp * ATTP_WARP_SMEM;
    const int total = *rows.n_slots * n_head;
    const uint64_t pol_last = l2_policy_evict_last();
    for (int w = blockIdx.x * ATTP_WARPS + warp; w < total; w += gridDim.x * ATTP_WARPS) {
        const int slot = w / n_head, head = w - slot * n_head;
        const int e = rows.slot_env[slot], r0 = rows.slot_row0[slot], n = rows.slot_nrow[slot];
        const int p0 = rows.row_pos[r0];
        const int my_page = lane < kv.ppe ? kv.table[(size_t)e * kv.ppe + lane] : 0;     // this env's page table, one entry per lane
        for (int rb = 0; rb < n;) {
            if (wide && n - rb > 16) {
                attp_rows<2>(q, kcache, vcache, r0, n, rb, p0, head, d_model, n_head, my_page, scale_log2e, rb + 32 < n, pol_last, sbase, lane, out);
                rb += 32;
            } else {
                attp_rows<1>(q, kcache, vcache, r0, n, rb, p0, head, d_model, n_head, my_page, scale_log2e, rb + 16 < n, pol_last, sbase, lane, out);
                rb += 16;
            }
        }
    }
}
static inline int attp_grid(int n_env, int n_head, int sm_count) {
    return std::max(1, std::min((n_env * n_head + ATTP_WARPS - 1) / ATTP_WARPS, sm_count * ATTP_CTAS_PER_SM));
}

// ---- sampler bookkeeping ------------------------------------------------------------------------------------------------
// per-row sampling key: fmix of (base seed, episode key, turn, token index) — independent of batch layout / GPU count
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void prepare_sample_kernel(RowsDev rows, GenDev gen, const uint64_t* __restrict__ episode_key, const int32_t* __restrict__ turn_of_env,
                                      int turn_const, uint64_t base_seed, uint64_t* __restrict__ row_key, unsigned long long* __restrict__ best) {
    pdl_enter();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= *rows.n_slots) return;
    const int e = rows.slot_env[j];
    const uint64_t ek = episode_key ? episode_key[e] : (uint64_t)e;
    const uint64_t turn = turn_of_env ? (uint64_t)turn_of_env[e] : (uint64_t)turn_const;
    row_key[j] = splitmix64(splitmix64(base_seed ^ splitmix64(ek)) + (turn << 32) + (uint64_t)gen.len[e]);
    best[j] = 0ull;
}

// ---- top-k / top-p filtered sampling (HF logits warpers; None in every reference Wordle / Maze driver, so this is the slow path) ----
// One CTA per sampled row.  The row's fp32 logits (already dumped by the LM-head GEMM) are scaled by 1/T and parked in shared
// memory (n_vocab floats), then
//   top-k (FlaxTopKLogitsWarper): the k-th largest value is found by bisection on the order-preserving integer image of the
//          floats (32 counting passes over shared memory); ties at the threshold keep the lowest indices, as lax.top_k does;
//   top-p (FlaxTopPLogitsWarper): in descending order a token is kept iff the probability mass of the tokens ranked before it
//          is < top_p (the token that crosses top_p is kept); the cut value is found by the same bisection on the mass of
//          strictly larger scores, ties again resolved by index;
//   the survivor set is sampled with the engine's Gumbel-max noise (same (row key, column) counter hash as the fused path).
// Kept set and arg-max are written as the packed `best` word finalize_sample_kernel expects.
constexpr int FS_THREADS = 256;

__device__ __forceinline__ float block_sum_f(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < FS_THREADS / 32; ++i) t += red[i];
    return t;
}
__device__ __forceinline__ float block_max_f(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = -INFINITY;
    for (int i = 0; i < FS_THREADS / 32; ++i) t = fmaxf(t, red[i]);
    return t;
}
// exclusive prefix of `v` over threads in thread order (each thread owns a contiguous index range, so this is index order)
__device__ __forceinline__ int block_excl_scan_i(int v, int* red, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
    __syncthreads();
    if (lane == 31) red[warp] = x;
    __syncthreads();
    int base = 0, tot = 0;
    for (int i = 0; i < FS_THREADS / 32; ++i) { if (i < warp) base += red[i]; tot += red[i]; }
    *total = tot;
    return base + x - v;
}

__global__ void __launch_bounds__(FS_THREADS) filtered_sample_kernel(const float* __restrict__ logits, int n_vocab, RowsDev rows, const uint64_t* __restrict__ row_key,
                                                                     float inv_temp, int do_sample, int top_k, float top_p,
                                                                     unsigned long long* __restrict__ best) {
    pdl_enter();
    extern __shared__ float fs_row[];                  // [n_vocab] scaled logits; -inf once filtered out
    __shared__ float red_f[FS_THREADS / 32];
    __shared__ int red_i[FS_THREADS / 32];
    const int j = blockIdx.x;
    if (j >= *rows.n_slots) return;
    const int tid = threadIdx.x;
    const int per = (n_vocab + FS_THREADS - 1) / FS_THREADS, lo = min(tid * per, n_vocab), hi = min(lo + per, n_vocab);   // my index range
    const float* src = logits + (size_t)j * n_vocab;
    float mx = -INFINITY;
    for (int i = tid; i < n_vocab; i += FS_THREADS) { const float v = src[i] * inv_temp; fs_row[i] = v; mx = fmaxf(mx, v); }
    mx = block_max_f(mx, red_f);                       // (contains the __syncthreads that publish fs_row)
    if (top_k > 0 && top_k < n_vocab) {
        // largest key T with count(x >= T) >= k  == the k-th largest value
        uint32_t klo = 0u, khi = 0xFFFFFFFFu;
        while (klo < khi) {
            const uint32_t mid = klo + (uint32_t)(((uint64_t)khi - klo + 1) >> 1);
            int c = 0;
            for (int i = tid; i < n_vocab; i += FS_THREADS) c += float_ordered(fs_row[i]) >= mid;
            int tot;
            (void)block_excl_scan_i(c, red_i, &tot);
            if (tot >= top_k) klo = mid; else khi = mid - 1;
        }
        int c_gt = 0, c_eq = 0;
        for (int i = lo; i < hi; ++i) { const uint32_t u = float_ordered(fs_row[i]); c_gt += u > klo; c_eq += u == klo; }
        int tot_gt, tot_eq;
        (void)block_excl_scan_i(c_gt, red_i, &tot_gt);
        int rank = block_excl_scan_i(c_eq, red_i, &tot_eq);          // how many threshold-valued entries precede my range
        const int keep_eq = top_k - tot_gt;                          // lowest-index ties that still fit
        for (int i = lo; i < hi; ++i) {
            const uint32_t u = float_ordered(fs_row[i]);
            if (u < klo) fs_row[i] = -INFINITY;
            else if (u == klo) { if (rank >= keep_eq) fs_row[i] = -INFINITY; ++rank; }
        }
        __syncthreads();
    }
    if (top_p > 0.f && top_p < 1.f) {
        float z = 0.f;
        for (int i = tid; i < n_vocab; i += FS_THREADS) z += __expf(fs_row[i] - mx);
        z = block_sum_f(z, red_f);
        const float target = top_p * z;                              // compare unnormalised masses
        // smallest key T with mass(x > T) < top_p: every value >= T has less than top_p ranked before its first occurrence
        uint32_t klo = 0u, khi = 0xFFFFFFFFu;
        while (klo < khi) {
            const uint32_t mid = klo + (uint32_t)(((uint64_t)khi - klo) >> 1);
            float m = 0.f;
            for (int i = tid; i < n_vocab; i += FS_THREADS) { const float v = fs_row[i]; m += float_ordered(v) > mid ? __expf(v - mx) : 0.f; }
            m = block_sum_f(m, red_f);
            if (m < target) khi = mid; else klo = mid + 1;
        }
        // ties at the cut value: the r-th of them (index order) is kept iff mass_gt + r * p(value) < top_p
        float m_gt = 0.f; int c_eq = 0; float p_eq = 0.f;
        for (int i = lo; i < hi; ++i) { const float v = fs_row[i]; const uint32_t u = float_ordered(v); if (u > klo) m_gt += __expf(v - mx); else if (u == klo) { ++c_eq; p_eq = __expf(v - mx); } }
        m_gt = block_sum_f(m_gt, red_f);
        p_eq = block_max_f(p_eq, red_f);
        int tot_eq;
        int rank = block_excl_scan_i(c_eq, red_i, &tot_eq);
        for (int i = lo; i < hi; ++i) {
            const uint32_t u = float_ordered(fs_row[i]);
            if (u < klo) fs_row[i] = -INFINITY;
            else if (u == klo) { if (!(m_gt + (float)rank * p_eq < target) && rank > 0) fs_row[i] = -INFINITY; ++rank; }
        }
        __syncthreads();
    }
    // Gumbel-max over the survivors (lowest column wins ties, like the fused sampler)
    const uint64_t key = row_key[j];
    float bs = -INFINITY; int bc = 0x7FFFFFFF;
    for (int i = tid; i < n_vocab; i += FS_THREADS) {
        float v = fs_row[i];
        if (v > -INFINITY) {
            if (do_sample) v += gumbel_from_bits(noise_bits(key, (uint32_t)i));
            if (v > bs) { bs = v; bc = i; }
        }
    }
    if (bc != 0x7FFFFFFF) atomicMax(best + j, pack_best(bs, bc));
}

__global__ void finalize_sample_kernel(RowsDev rows, GenDev gen, const unsigned long long* __restrict__ best) {
    pdl_enter();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= *rows.n_slots) return;
    const int e = rows.slot_env[j];
    const int tok = (int)(0xFFFFFFFFu - (uint32_t)(best[j] & 0xFFFFFFFFull));
    const int k = gen.len[e];
    gen.tok[e * gen.max_new + k] = tok;
    gen.len[e] = k + 1;
    if (tok == gen.eos_id || k + 1 >= gen.max_new) gen.active[e] = 0;
}

}  // namespace lmrl
