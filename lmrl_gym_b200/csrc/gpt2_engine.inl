// gpt2_engine.inl — GPT-2 KV-cache decoder, fused LM-head sampler and the fused rollout loop (C ABI part 2).
//
// Reference path being replaced: GPT2PPOPolicy.act -> generate_from_str -> pjit _generate ->
// HF FlaxGenerationMixin._sample (prefill + lax.while_loop of single-token forwards) around the env's Python step.
// Here one host thread enqueues, per turn: [row build -> embed -> L x (LN, QKV GEMM + KV append, attention,
// proj GEMM + residual, LN, FC GEMM + GeLU, FC2 GEMM + residual) -> final LN of the sampled rows -> LM-head GEMM
// with fused Gumbel arg-max] x max_new_tokens, then the env step kernel.  Row counts and done masks stay on the
// device; nothing is copied to the host until the rollout is over.
#include "gpt2_kernels.cuh"
#include "gemm_tcgen05.cuh"
#include "gemm_tcgen05_2cta.cuh"
#include "gemv.cuh"
#include <map>
#include <tuple>

namespace lmrl {

struct LayerDev {
    float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    __nv_bfloat16 *w_qkv, *w_proj, *w_fc, *w_fc2;      // [N][K] bf16
    float *b_qkv, *b_proj, *b_fc, *b_fc2;
    // LayerNorm-folded copies for the consumer-side fold (GemmArgs::stat_in): W''[n][k] = gamma[k] W[n][k] - mean_k(.), b' = b + beta . W
    __nv_bfloat16 *w_qkv_f = nullptr, *w_fc_f = nullptr;
    float *b_qkv_f = nullptr, *b_fc_f = nullptr;
};

// one GPT-2 parameter set + its KV slab.  Tower 0 = the sampled policy (PPO / BC policy, or pi_beta of ILQL / MC);
// tower 1 = the value base of ILQL / MC (value_rl_base/gpt2/generation.py:77-88), same architecture, own KV cache.
struct TowerDev {
    bool loaded = false;
    bool fold_valid = false;          // the LayerNorm-folded weight copies (LayerDev::w_qkv_f ...) match the current weights
    __nv_bfloat16 *wte = nullptr, *wpe = nullptr;
    float *lnf_g = nullptr, *lnf_b = nullptr;
    std::vector<LayerDev> layers;
    __nv_bfloat16* kv = nullptr;      // [L][2][n_pages][H][KV_PAGE][64] page pools (one page table for all layers and towers: Gpt2Dev::kvp)
};

// Q heads of the value-perturbed policy: MLPHead = dense2(relu(dense1(h))) (heads/mlp_head.py:139-148), one or two of them
struct QHeadsDev {
    int n_q = 0, hidden = 0;
    int linear = 0;                   // 1: LinearHead Q heads (heads/linear_head.py:112-119, the MC-returns drivers): q = dense(h), no first layer;
                                      //    w2 then holds beta * kernel and the value tower's final LN writes h straight into xl[:, d:2d]
    float beta = 0.f;
    __nv_bfloat16* w1cat = nullptr;   // [n_q * hidden][d]   dense1 kernels, transposed, stacked along N
    float* b1cat = nullptr;           // [n_q * hidden]
    __nv_bfloat16* w2[2] = {nullptr, nullptr};   // [v_pad][hidden]  dense2 kernel, transposed, zero-padded rows
    float* b2s[2] = {nullptr, nullptr};          // [v_pad]  beta * dense2 bias
};

// value heads used by the teacher-forced per-position forward (lmrl_forward_host): up to two Q heads and one V head, each an
// MLPHead (heads/mlp_head.py:139-148), or a LinearHead V (heads/linear_head.py:112-119: the PPO value head).  Unlike QHeadsDev
// nothing is scaled by beta: these return the raw head outputs (ilql/gpt2/score_fn.py:45-58, ppo/base_interface.py:529).
struct ValueHeadsDev {
    int n_mlp = 0, hidden = 0;        // MLP heads stacked in w1cat, in the order q1, q2, v (absent ones skipped)
    int slot_q1 = -1, slot_q2 = -1, slot_v = -1;     // index of each head inside w1cat / the ReLU buffer
    __nv_bfloat16* w1cat = nullptr;   // [3 * hidden][d]
    float* b1cat = nullptr;           // [3 * hidden]
    __nv_bfloat16* w2q[2] = {nullptr, nullptr};      // [v_pad][hidden] dense2 kernels of the Q heads, transposed
    float* b2q[2] = {nullptr, nullptr};              // [v_pad]
    float* w2v = nullptr;             // [hidden] dense2 kernel of the V MLP head (output_dim 1)
    float* lin_w = nullptr;           // [d] LinearHead kernel
    float b2v = 0.f, lin_b = 0.f;
    bool has_lin = false, allocated = false;
    bool q_lin[2] = {false, false};   // Q head k is a LinearHead: w2q[k] = its kernel (transposed), applied to the final-LN hidden itself
};

struct TrainDev;                          // train_kernels.inl: buffers of the hand-written backward pass
struct Gpt2Dev;
struct VecCopy;
}  // namespace lmrl
void lmrl_train_free(lmrl::Gpt2Dev* g);
static int train_refresh_weight_copies(lmrl_engine* e);
namespace lmrl {

// One lane = the activation buffers, row map and sampling scratch of one forward pass in flight.  Lane 0 is the whole batch (every
// host-facing call, profiled runs); lanes 1 and 2 are the two halves of lane 0's buffers, each with its own stream, used by the fused
// rollout to run two half-batches concurrently: while one half is in its HBM-bound kernels (KV attention, LayerNorm) the other
// half's tensor-bound GEMMs keep the tensor pipe busy, and the tail of every one-wave decode GEMM is filled by the other lane.
struct Lane {
    float* x = nullptr;
    __nv_bfloat16 *xn = nullptr, *q = nullptr, *att = nullptr, *h = nullptr, *xl = nullptr, *xl1 = nullptr;
    RowsDev rows{};
    uint64_t* row_key = nullptr;
    unsigned long long *best = nullptr, *best_scratch = nullptr;
    cudaStream_t st = nullptr;
    int* ln_cnt = nullptr;    // per-m-tile arrival counters of the LayerNorm fused into the residual GEMMs
    float2* stat[2] = {nullptr, nullptr};   // [m_cap][8] per-row partial (sum, sumsq) of the residual stream: [0] feeds ln_1 consumers, [1] ln_2 consumers
    int e0 = 0;               // first env of the lane
    size_t m_cap = 0;         // row capacity of the lane's activation buffers
    size_t xl_rows = 0;       // rows behind xl / xl1
};

struct Gpt2Dev {
    lmrl_gpt2_dims dims{};
    int v_pad = 0;            // LM-head rows padded to a multiple of 128
    int t_cap = 0;            // KV positions per env
    int m_cap = 0;            // row capacity of the activation buffers
    int stream_cap = 0;
    TowerDev tw[3];                   // 0 policy / pi_beta, 1 value base (or PPO initial policy), 2 target value base (forward only; shares tower 1's KV slab)
    QHeadsDev qh;
    ValueHeadsDev vh[2];              // head sets of lmrl_forward_host: 0 online, 1 target
    int policy_kind = 0;              // 0 = LM policy (logits of tower 0); 1 = value-perturbed (pi_beta + beta * min q); 2 = beta * min q alone (pi_beta None)
    size_t kv_layer_stride = 0;       // elements between layers (covers K and V)
    KvPages kvp{};                    // page table + allocator of the paged KV cache (gpt2_kernels.cuh)
    // activations
    float* x = nullptr;
    // xl: [B_pad][3 d] = [final-LN hidden of tower 0 | beta*relu(dense1) of Q head 1 | of Q head 2] (A operand of the sampler);
    // xl1: [B_pad][d] final-LN hidden of the value tower
    __nv_bfloat16 *xn = nullptr, *q = nullptr, *att = nullptr, *h = nullptr, *xl = nullptr, *xl1 = nullptr;
    int xl_cols = 0;
    // teacher-forced scoring (lmrl_score_host): head operands for up to s_cap scored rows, partial log-sum-exp pairs, results
    __nv_bfloat16 *xs = nullptr, *xs1 = nullptr;
    int s_cap = 0;
    float2* lp_partial = nullptr; float* lp_tlogit = nullptr; float* lp_out = nullptr;
    int *score_rows = nullptr, *score_targets = nullptr, *score_n = nullptr;
    // per-position forward (lmrl_forward_host): final-LN hidden of every row, ReLU(dense1) of the value heads, per-row outputs
    int f_cap = 0;
    __nv_bfloat16 *fx = nullptr, *fr = nullptr;
    int* f_targets = nullptr;
    float2* f_lp_partial = nullptr;
    float *f_tlogit = nullptr, *f_lp = nullptr, *f_q1 = nullptr, *f_q2 = nullptr, *f_v = nullptr;
    RowsDev rows{};
    TrainDev* train = nullptr;
    Lane lane[3];
    Lane* cur = nullptr;              // lane the next tower_layers / forward_and_sample call runs on
    // fused rollouts may split the batch over lanes 1 and 2 (see lanes_usable).  Off by default: measured on B200 it is SLOWER
    // (4096-episode Wordle rollout, graphs on: 215.1 ms one lane, 220.1 ms two lanes; letters variant 80.4 vs 82.6).  With the GEMM
    // capped at 128 registers (__maxnreg__, no spills) a GEMM CTA and an attention CTA do fit one SM together, and the gap narrows
    // (213.7 vs 217.2 ms; 81.4 vs 81.6) but does not close: the half-batch GEMMs of a decode pass (16 m-tiles) are less efficient
    // than the full-batch ones and that outweighs the overlap.  Results are bit-identical either way.
    int two_lanes = 0;
    cudaStream_t stream_b = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int* lane_counts = nullptr;       // [6] M / n_slots / ctx_total of lanes 1 and 2
    GenDev gen{};
    StreamDev stream{};
    uint64_t* row_key = nullptr;
    unsigned long long* best = nullptr;
    unsigned long long* best_scratch = nullptr;   // sink for the fused arg-max when top-k / top-p re-samples from dumped logits
    uint64_t* episode_key = nullptr;  // [B]
    float* logits = nullptr;          // parity hook buffer, allocated on first use
    size_t logits_rows = 0;
    // rollout logs (device)
    float* ep_return = nullptr; int32_t* ep_length = nullptr; uint8_t* ep_done = nullptr;
    float* step_reward = nullptr; uint8_t* step_valid = nullptr; uint8_t* step_fb = nullptr; int32_t* step_aux = nullptr;
    int32_t* act_tokens = nullptr; int32_t* act_len = nullptr;
    int* live = nullptr;              // [1] scratch of the early-exit test
    // CUDA graph of one whole fused rollout (every launch is static: row counts live on the device).  A configuration is run
    // eagerly once (first-use attribute calls, tensor-map creation), captured on its second use and replayed afterwards.
    struct RolloutKey {
        int env_kind, n, do_sample, policy_kind, gemm_impl, has_opt, two_lanes;
        float temperature; uint64_t seed; int top_k; float top_p;
        bool operator==(const RolloutKey& o) const {
            return top_k == o.top_k && top_p == o.top_p && env_kind == o.env_kind && n == o.n && do_sample == o.do_sample && policy_kind == o.policy_kind && gemm_impl == o.gemm_impl &&
                   has_opt == o.has_opt && two_lanes == o.two_lanes && temperature == o.temperature && seed == o.seed;
        }
    };
    // the text path's generate call replays ONE captured decode pass max_new_tokens - 1 times (the turn index is read from device memory,
    // so the same graph serves every turn of an episode)
    cudaGraphExec_t decode_graph = nullptr;
    RolloutKey decode_key{}, decode_warm{};
    bool decode_warm_valid = false;
    int64_t decode_graph_launches = 0;
    int32_t* gen_turn = nullptr;      // [B] turn index of the current generate call
    cudaGraphExec_t graph_exec = nullptr;
    RolloutKey graph_key{}, warm_key{};
    bool warm_valid = false;
    int64_t graph_launches = 0;
    int use_graphs = 1;
    int narrow_decode_tiles = 1;      // 128 x 192 tiles for the N = 768 residual GEMMs of decode passes
    // LayerNorm fused into the residual GEMMs (GemmArgs::ln_out): the decoder stack launches one LayerNorm per tower pass (layer 0's
    // ln_1) instead of two per layer.  Bit-identical to the separate launches, but measured SLOWER on B200 (4096-episode Wordle
    // rollout: 247.7 vs 211.3 ms): the CTA that completes an m-tile has to wait for the TMA reductions of all its peers to land
    // and then normalises 128 rows at one SM's L2 latency (~20 us per m-tile), which lands on the critical path of every residual
    // GEMM (proj prefill 8.5 -> 35.8 ms per step), while the stand-alone layernorm_kernel runs at 6.1 TB/s (74 us per 98 k rows,
    // 6.6 us at 4096 rows).  Default 0 = separate layernorm_kernel launches; 1 is kept selectable (lmrl_set_option "fuse_ln").
    int fuse_ln = 0;
    int attn_wide = 1;                // prefill attention: 32 query rows per warp pass (two m16 tiles share every staged K / V block)
    int gemv_max_rows = 8;            // batches of at most this many envs decode through the weight-streaming GEMV kernels
    // LayerNorm folded into the CONSUMER GEMMs (see GemmArgs::stat_in): the residual GEMMs write x, its bf16 copy and per-row partial
    // statistics; QKV / FC run on the raw bf16 copy with gamma-folded, row-centred weights and scale their accumulators by 1 / sigma.
    // No LayerNorm launches in the decoder stack (one statistics pass per tower pass for layer 0).
    int ln_fold = 0;
    int att_grid_mult = 8;            // decode attention: grid cap in CTAs per SM (grid-stride beyond it)
    // decode-pass attention: 4 (default) = SIMT streaming kernel, a warp takes 4 consecutive heads of a slot (the slot's metadata chain is paid once
    // per 4 heads; falls back to one head per warp when n_head % 4 != 0), 3 = 2 heads per warp, 0 = one head per warp, 1 = the tensor-core
    // (cp.async staged) prefill kernel with one row per slot, 2 = one head per warp with software-pipelined K / V loads
    int att_decode_impl = 4;
    int nlnl_id = -1;                 // id of the tokenizer's '\n\n' token (joined-history prompts fold a trailing newline pair into it); -1 = none
    int fold_nlnl = -1;               // nlnl_id while a fused Wordle rollout is being enqueued, -1 on the text paths (their callers tokenise themselves)
    // programmatic dependent launch between the kernels of a pass (pdl.cuh).  Off by default: measured on B200 it is SLOWER
    // for this path (4096-episode Wordle rollout: 208.2 ms with graphs and no PDL, 223.8 ms with PDL; letters variant 81.0 vs 84.2)
    int use_pdl = 0;
    int sm_count = 148;
    // 2 (default) = tcgen05 with CTA pairs (cta_group::2, 256 x 256 tiles) on the plain GEMMs of prefill passes and one CTA per 128 x 256
    // tile everywhere else; 1 = one CTA per tile everywhere; 0 = mma.sync cross-check.  The pair kernel became the default once it
    // got the shared-memory bias staging of the one-CTA epilogue: 2-5 % per prefill GEMM, -1.2 ms per 4096-episode rollout (same box)
    int gemm_impl = 2;
    VecCopy* vec_tab_dev = nullptr;   // copy table of lmrl_load_gpt2_tower's device-to-device path
    std::map<std::tuple<const void*, size_t, size_t, int>, CUtensorMap> tmaps;   // (base pointer, row bound, columns, box rows | f32 flag) -> descriptor
    std::vector<void*> owned;         // allocations made outside engine::alloc (freed in lmrl_gpt2_free)
};

__global__ void transpose_to_bf16_kernel(const float* __restrict__ src, int K, int N, __nv_bfloat16* __restrict__ dst, int n_take = -1) {
    // src [K][N] fp32 (HF Conv1D / flax Dense layout) -> dst [N][K] bf16; only the first n_take columns of src when n_take >= 0
    __shared__ float tile[32][33];
    const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    const int n_lim = n_take >= 0 ? n_take : N;
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        if (k0 + i < K && n0 + threadIdx.x < n_lim) tile[i][threadIdx.x] = src[(size_t)(k0 + i) * N + n0 + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        if (n0 + i < n_lim && k0 + threadIdx.x < K) dst[(size_t)(n0 + i) * K + k0 + threadIdx.x] = __float2bfloat16(tile[threadIdx.x][i]);
}

// LayerNorm fold of one [K][N] fp32 matrix (HF Conv1D layout) into dst [N][K] bf16: W''[n][k] = gamma[k] W[k][n] - mean_k(gamma[k] W[k][n]);
// bfold[n] = bias[n] + sum_k beta[k] W[k][n].  One block per output feature n (load-time work: the strided reads do not matter).
__global__ void fold_ln_weights_kernel(const float* __restrict__ src, int K, int N, const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const float* __restrict__ bias, __nv_bfloat16* __restrict__ dst, float* __restrict__ bfold) {
    __shared__ float red[2][32];
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float sg = 0.f, sb = 0.f;
    for (int k = tid; k < K; k += blockDim.x) { const float w = src[(size_t)k * N + n]; sg = fmaf(gamma[k], w, sg); sb = fmaf(beta[k], w, sb); }
#pragma unroll
    for (int o = 16; o; o >>= 1) { sg += __shfl_xor_sync(0xFFFFFFFFu, sg, o); sb += __shfl_xor_sync(0xFFFFFFFFu, sb, o); }
    if (lane == 0) { red[0][warp] = sg; red[1][warp] = sb; }
    __syncthreads();
    float tg = 0.f, tb = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { tg += red[0][i]; tb += red[1][i]; }
    const float mean = tg / K;
    for (int k = tid; k < K; k += blockDim.x) dst[(size_t)n * K + k] = __float2bfloat16(gamma[k] * src[(size_t)k * N + n] - mean);
    if (tid == 0) bfold[n] = bias[n] + tb;
}

// first statistics of a tower pass (the embedded rows): bf16 copy of x and the full-row (sum, sumsq) into partial slot 0; one warp per row
__global__ void xb_stats_kernel(const float* __restrict__ x, int d, const int* __restrict__ n_dev, __nv_bfloat16* __restrict__ xb, float2* __restrict__ stat) {
    pdl_enter();
    const int n = *n_dev;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
        float s1 = 0.f, s2 = 0.f;
        for (int c = lane * 4; c < d; c += 128) {
            const float4 v = *reinterpret_cast<const float4*>(x + (size_t)r * d + c);
            s1 += (v.x + v.y) + (v.z + v.w);
            s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
            uint2 pk;
            pk.x = pack_bf16x2_tc(v.x, v.y); pk.y = pack_bf16x2_tc(v.z, v.w);
            *reinterpret_cast<uint2*>(xb + (size_t)r * d + c) = pk;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o); s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o); }
        if (lane == 0) stat[(size_t)r * 8] = make_float2(s1, s2);
    }
}

// the ~100 small fp32 vectors of a parameter set (LayerNorm scales / offsets, biases) when the source is already on the device (a trainer's
// weight hand-over after every optimizer step): ONE launch over a table instead of one cudaMemcpyAsync per vector
struct VecCopy { const float* src; float* dst; int n; int pad; };
__global__ void batched_vec_copy_kernel(const VecCopy* __restrict__ tab) {
    const VecCopy c = tab[blockIdx.x];
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) c.dst[i] = c.src[i];
}

__global__ void convert_to_bf16_kernel(const float* __restrict__ src, size_t n, __nv_bfloat16* __restrict__ dst) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = __float2bfloat16(src[i]);
}

__global__ void init_stream_kernel(int n, const int32_t* __restrict__ header, int n_header, StreamDev st,
                                   const uint64_t* __restrict__ seeds, uint64_t* __restrict__ episode_key, float* ep_return,
                                   int32_t* ep_length, uint8_t* ep_done) {
    pdl_enter();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    for (int i = 0; i < n_header; ++i) {
        st.tokens[(size_t)e * st.cap + i] = header[i];
        if (st.is_action) { st.is_action[(size_t)e * st.cap + i] = 0; st.tok_reward[(size_t)e * st.cap + i] = 0.f; }
    }
    if (st.truncated) st.truncated[e] = 0;
    st.len[e] = n_header;
    st.kv_len[e] = 0;
    episode_key[e] = seeds[e];
    ep_return[e] = 0.f; ep_length[e] = 0; ep_done[e] = 0;
}

// after the env step: append this turn's results to the per-episode logs
__global__ void log_turn_kernel(int n, int max_turns, GenDev gen, StepOut step,
                                float* ep_return, int32_t* ep_length, uint8_t* ep_done, float* step_reward, uint8_t* step_valid,
                                uint8_t* step_fb, int32_t* step_aux, int32_t* act_tokens, int32_t* act_len) {
    pdl_enter();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int glen = gen.len[e];
    if (glen <= 0) return;                      // env did not act this turn (already done)
    const int t = ep_length[e];                 // transitions logged so far == index of this turn
    if (t < 0 || t >= max_turns) return;
    const size_t o = (size_t)e * max_turns + t;
    step_reward[o] = step.reward[e];
    step_valid[o] = step.valid[e];
    step_aux[o] = step.aux[e];
    for (int i = 0; i < 5; ++i) step_fb[o * 5 + i] = step.fb[e * 5 + i];
    for (int k = 0; k < gen.max_new; ++k) act_tokens[o * gen.max_new + k] = k < glen ? gen.tok[e * gen.max_new + k] : -1;
    act_len[o] = glen;
    ep_return[e] += step.reward[e];
    ep_length[e] = t + 1;
    ep_done[e] = step.done[e];
}

// ---- kernel microbenchmark helpers (lmrl_bench_kernels): synthetic row map + random activations ----------------------
__global__ void bench_rows_kernel(int M, int rpe, int ctx, int n_slots, RowsDev rows, KvPages kv) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) { *rows.M = M; *rows.n_slots = n_slots; }
    if (r < n_slots * kv.ppe) kv.table[r] = r % kv.n_pages;          // env e owns pages [e * ppe, (e + 1) * ppe)
    if (r >= M) return;
    const int e = r / rpe, i = r - e * rpe;
    rows.row_env[r] = e; rows.row_pos[r] = ctx - rpe + i; rows.row_tok[r] = (r * 7919) % 50000;
    if (i == 0) { rows.slot_env[e] = e; rows.slot_row0[e] = r; rows.slot_nrow[e] = rpe; rows.slot_last[e] = r + rpe - 1; }
}
__global__ void bench_fill_bf16_kernel(__nv_bfloat16* dst, size_t n, uint32_t seed, float scale) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16(scale * ((float)(fmix32((uint32_t)i ^ seed) >> 8) * (2.0f / 16777216.0f) - 1.0f));
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D row-major tensor map with a 128-byte-swizzled box of (128 bytes) x box_rows: bf16 -> 64 columns, fp32 -> 32 columns
static int make_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, bool f32) {
    static PFN_encodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        LMRL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        LMRL_CHECK(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available from this driver");
        encode = reinterpret_cast<PFN_encodeTiled>(fn);
    }
    const cuuint64_t esz = f32 ? 4 : 2;
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {cols * esz};
    const cuuint32_t box[2] = {(cuuint32_t)(128 / esz), box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                              gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LMRL_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu)", (int)r, (unsigned long long)rows,
               (unsigned long long)cols);
    return 0;
}
static int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t K, uint32_t box_rows) {
    return make_tmap(out, base, rows, K, box_rows, false);
}

template <int EPI, int BN = TC_BN, bool XB = false>
static int launch_gemm_tc(cudaStream_t st, int pdl, int sm_count, const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmW1,
                          const CUtensorMap& tmW2, const CUtensorMap& tmOut, const GemmArgs& g) {
    static bool configured = false;
    if (!configured) {
        LMRL_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<EPI, BN, XB>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcTile<BN>::SMEM_BYTES));
        configured = true;
    }
    LMRL_CUDA(launch_k(pdl, 1, gemm_bf16_tcgen05_kernel<EPI, BN, XB>, dim3(sm_count), dim3(TcCfg<EPI>::THREADS), (size_t)TcTile<BN>::SMEM_BYTES, st, tmA, tmW,
                       tmW1, tmW2, tmOut, g));
    return 0;
}

template <int EPI>
static int launch_gemm_tc2(cudaStream_t st, int pdl, int sm_count, const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmOut, const GemmArgs& g) {
    static bool configured = false;
    if (!configured) {
        LMRL_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_2cta_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES));
        configured = true;
    }
    LMRL_CUDA(launch_k(pdl, 2, gemm_bf16_tcgen05_2cta_kernel<EPI>, dim3((unsigned)(sm_count & ~1)), dim3(TC2_THREADS), (size_t)TC2_SMEM_BYTES, st, tmA, tmW, tmOut, g));
    return 0;
}

template <int EPI>
static int launch_gemm_mma(cudaStream_t st, int pdl, int sm_count, const GemmArgs& g) {
    static bool configured = false;
    if (!configured) {
        LMRL_CUDA(cudaFuncSetAttribute(gemm_bf16_mma_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
        configured = true;
    }
    LMRL_CUDA(launch_k(pdl, 1, gemm_bf16_mma_kernel<EPI>, dim3(sm_count * 2), dim3(GEMM_THREADS), (size_t)GEMM_SMEM_BYTES, st, g));
    return 0;
}

// extra operands of a multi-segment GEMM (the value-perturbed sampler): A row pitch and the two further W matrices
struct GemmExtra {
    size_t a_cols = 0;                       // columns of the A tensor (0 = g.K)
    const __nv_bfloat16* W1 = nullptr; size_t w1_cols = 0;
    const __nv_bfloat16* W2 = nullptr; size_t w2_cols = 0;
    size_t out_cols = 0;                     // columns of the output tensor (0 = g.N)
};

// a_rows / w_rows: allocated rows behind g.A / g.W (TMA bounds)
template <int EPI>
static int launch_gemm(lmrl_engine* e, GemmArgs g, int n_true, int count_kind, size_t a_rows, size_t w_rows, int kind,
                       const GemmExtra& x = GemmExtra()) {
    Gpt2Dev& G = *e->gpt2;
    if (g.n_seg == 0) { g.n_seg = 1; g.n_acc = 1; g.seg[0] = {0, 0, 0, g.K / TC_BK}; }
    const bool tc = G.gemm_impl >= 1 && g.N % TC_BN == 0 && g.K % TC_BK == 0;
    LMRL_CHECK(tc || (g.n_seg == 1 && EPI != EPI_RELU && EPI != EPI_LOGPROB), "this GEMM needs the tcgen05 kernel (multi-segment / ReLU epilogue): N=%d K=%d", g.N, g.K);
    const CUtensorMap *ta = nullptr, *tw = nullptr, *tw1 = nullptr, *tw2 = nullptr, *to = nullptr;
    bool use2 = false, narrow = false;
    if (tc) {
        auto get = [&](const void* base, size_t rows, size_t cols, uint32_t box_rows, bool f32) -> const CUtensorMap* {
            auto key = std::make_tuple(base, rows, cols, (int)box_rows | (f32 ? 1 << 16 : 0));
            auto it = G.tmaps.find(key);
            if (it == G.tmaps.end()) {
                CUtensorMap m;
                if (make_tmap(&m, base, rows, cols, box_rows, f32)) return nullptr;
                it = G.tmaps.emplace(key, m).first;
            }
            return &it->second;
        };
        // CTA-pair kernel: plain GEMMs of the prefill passes (many rows); each CTA stages 128 of the 256 W rows
        use2 = G.gemm_impl == 2 && e->cur_mode == 0 && g.n_seg == 1 && EPI != EPI_SAMPLE && EPI != EPI_LOGPROB && !x.W1 && !g.ln_out && !g.xb_out && !g.stat_in;
        // decode pass (one row per env), N = 768 residual GEMMs: 192-wide tiles fill the SMs better than 256-wide ones
        narrow = EPI == EPI_RESID && G.narrow_decode_tiles && e->cur_mode == 1 && g.N == 768 && g.n_seg == 1 && !use2;
        ta = get(g.A, a_rows, x.a_cols ? x.a_cols : (size_t)g.K, TC_BM, false);
        tw = get(g.W, w_rows, (size_t)g.K, use2 ? TC2_B_ROWS : (narrow ? 192 : TC_BN), false);
        tw1 = x.W1 ? get(x.W1, w_rows, x.w1_cols, TC_BN, false) : tw;
        tw2 = x.W2 ? get(x.W2, w_rows, x.w2_cols, TC_BN, false) : tw;
        // output map (TMA store / reduce-add source boxes are 32 rows x 128 bytes); the sampling epilogue stores nothing
        if (EPI == EPI_RESID) to = get(g.resid, a_rows, (size_t)g.N, 32, true);
        else if (EPI == EPI_GELU || EPI == EPI_RELU) to = get(g.out, a_rows, x.out_cols ? x.out_cols : (size_t)g.N, 32, false);
        else if (EPI == EPI_QKV) to = get(g.out, a_rows, (size_t)g.d_model, 32, false);
        else to = ta;
        if (!ta || !tw || !tw1 || !tw2 || !to) return -1;
    }
    LMRL_CHECK(!(g.ln_out || g.xb_out || g.stat_in) || (tc && !use2), "fused / folded LayerNorm needs the tcgen05 one-CTA GEMM (N=%d K=%d impl=%d)", g.N, g.K, G.gemm_impl);
    double k_total = 0.0;
    for (int i = 0; i < g.n_seg; ++i) k_total += (double)g.seg[i].kb_count * TC_BK;
    if (e->prof_begin(0, 2.0 * (double)n_true * k_total, count_kind, 2 * kind + e->cur_mode)) return -1;
    int rc;
    if constexpr (EPI != EPI_SAMPLE && EPI != EPI_LOGPROB) {
        if (use2) {
            rc = launch_gemm_tc2<EPI>(G.cur->st, G.use_pdl, G.sm_count, *ta, *tw, *to, g);
            e->prof_end();
            e->launches++;
            return rc;
        }
    }
    if constexpr (EPI == EPI_RESID) {
        if (narrow || g.xb_out) {
            if (g.xb_out) rc = narrow ? launch_gemm_tc<EPI_RESID, 192, true>(G.cur->st, G.use_pdl, G.sm_count, *ta, *tw, *tw, *tw, *to, g)
                                      : launch_gemm_tc<EPI_RESID, TC_BN, true>(G.cur->st, G.use_pdl, G.sm_count, *ta, *tw, *tw, *tw, *to, g);
            else rc = launch_gemm_tc<EPI_RESID, 192>(G.cur->st, G.use_pdl, G.sm_count, *ta, *tw, *tw, *tw, *to, g);
            e->prof_end();
            e->launches++;
            return rc;
        }
    }
    if constexpr (EPI == EPI_RELU || EPI == EPI_LOGPROB) rc = launch_gemm_tc<EPI>(G.cur->st, G.use_pdl, G.sm_count, *ta, *tw, *tw1, *tw2, *to, g);
    else rc = tc ? launch_gemm_tc<EPI>(G.cur->st, G.use_pdl, G.sm_count, *ta, *tw, *tw1, *tw2, *to, g) : launch_gemm_mma<EPI>(G.cur->st, G.use_pdl, G.sm_count, g);
    e->prof_end();
    e->launches++;
    return rc;
}

// weight-streaming GEMV (gemv.cuh) of a small-batch decode pass: R = rows staged per CTA (1, 2, 4 or 8 >= the batch's envs)
template <int EPI, bool LN, int R>
static int launch_gemv_r(lmrl_engine* e, const GemvArgs& a) {
    Gpt2Dev& G = *e->gpt2;
    const size_t smem = (size_t)R * a.K * sizeof(float);
    static bool configured = false;
    if (!configured) {
        LMRL_CUDA(cudaFuncSetAttribute(gemv_kernel<EPI, R, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, GV_MAX_ROWS * 4096 * (int)sizeof(float)));
        configured = true;
    }
    const int n_groups = a.N / GV_COLS;
    const int grid = std::max(1, std::min(n_groups, G.sm_count * 8));          // one group of 4 output features per CTA pass
    LMRL_CUDA(launch_k(G.use_pdl, 1, gemv_kernel<EPI, R, LN>, dim3(grid), dim3(GV_WARPS * 32), smem, G.cur->st, a));
    return 0;
}
template <int EPI, bool LN>
static int launch_gemv(lmrl_engine* e, const GemvArgs& a, int R, int n_true, int count_kind, int kind) {
    LMRL_CHECK(a.N % GV_COLS == 0 && a.K % 128 == 0 && a.K <= 4096, "GEMV path: unsupported shape N=%d K=%d", a.N, a.K);
    if (e->prof_begin(0, 2.0 * (double)n_true * a.K, count_kind, 2 * kind + e->cur_mode)) return -1;
    int rc;
    switch (R) {
        case 1: rc = launch_gemv_r<EPI, LN, 1>(e, a); break;
        case 2: rc = launch_gemv_r<EPI, LN, 2>(e, a); break;
        case 4: rc = launch_gemv_r<EPI, LN, 4>(e, a); break;
        default: rc = launch_gemv_r<EPI, LN, 8>(e, a); break;
    }
    e->prof_end();
    e->launches++;
    return rc;
}
// rows staged by the GEMV path for a decode pass over n_env envs; 0 = the pass takes the tensor-core tiles
static int gemv_rows(const Gpt2Dev& G, int mode, int n_env) {
    if (mode != 1 || G.gemv_max_rows <= 0 || n_env > std::min(G.gemv_max_rows, GV_MAX_ROWS) || G.gemm_impl == 0) return 0;
    return n_env <= 1 ? 1 : n_env <= 2 ? 2 : n_env <= 4 ? 4 : 8;
}

// every kernel of the rollout path is launched through launch_k so that programmatic dependent launch can be switched on
#define LMRL_K(kernel, grid, block, smem, st, ...) \
    ::lmrl::launch_k((e && e->gpt2) ? e->gpt2->use_pdl : 0, 1, kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)

#define LMRL_LAUNCH(cat, ...)                          \
    do {                                               \
        if ((cat) >= 0 && e->prof_begin(cat)) return -1; \
        __VA_ARGS__;                                   \
        {                                              \
            const cudaError_t lerr__ = cudaPeekAtLastError(); \
            if (lerr__ != cudaSuccess) return ::lmrl::fail("kernel launch failed: %s (%s:%d)", cudaGetErrorString(lerr__), __FILE__, __LINE__); \
        }                                              \
        e->prof_end();                                 \
        e->launches++;                                 \
    } while (0)

// teacher-forced scoring request: which rows of the pass feed the LM head and which token each is scored against
struct ScoreReq {
    const int* rows;       // [n_score] row indices (device)
    const int* targets;    // [n_score] target token of every scored row (device)
    const int* n_dev;      // [1] n_score (device)
    float* out;            // [n_score] log p(target | prefix) (device)
    int n_score;           // host copy (grid sizing only)
};

// merge the per-(tile, half) partial (max, sum) pairs of one row: out = logit_target - logsumexp
__global__ void logprob_finalize_kernel(const float2* __restrict__ partial, int n_part, const float* __restrict__ tlogit, const int* __restrict__ n_dev,
                                        float* __restrict__ out) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int n = *n_dev;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
        float m = -INFINITY, s = 0.f;
        for (int i = lane; i < n_part; i += 32) {
            const float2 p = partial[(size_t)r * n_part + i];
            if (p.x > -INFINITY) {
                const float mn = fmaxf(m, p.x);
                s = s * __expf(m - mn) + p.y * __expf(p.x - mn);
                m = mn;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float mo = __shfl_xor_sync(0xFFFFFFFFu, m, o), so = __shfl_xor_sync(0xFFFFFFFFu, s, o);
            const float mn = fmaxf(m, mo);
            const float a = m > -INFINITY ? s * __expf(m - mn) : 0.f, b = mo > -INFINITY ? so * __expf(mo - mn) : 0.f;
            s = a + b; m = mn;
        }
        if (lane == 0) out[r] = tlogit[r] - (m + __logf(s));
    }
}

// profiling: copy the live row / slot / context counts of the pass just built to the pinned host array (FLOP and byte accounting)
static void note_pass_counts(lmrl_engine* e) {
    Gpt2Dev& G = *e->gpt2;
    if (!(e->profile && e->n_passes < e->max_passes)) return;
    e->cur_pass_slot = e->n_passes++;
    const RowsDev& R = G.cur->rows;
    cudaMemcpyAsync(e->h_counts + 3 * e->cur_pass_slot, R.M, sizeof(int), cudaMemcpyDeviceToHost, G.cur->st);
    cudaMemcpyAsync(e->h_counts + 3 * e->cur_pass_slot + 1, R.n_slots, sizeof(int), cudaMemcpyDeviceToHost, G.cur->st);
    cudaMemcpyAsync(e->h_counts + 3 * e->cur_pass_slot + 2, R.ctx_total, sizeof(int), cudaMemcpyDeviceToHost, G.cur->st);
}

// the decoder layers of one tower over the current row map: x (fp32 residual stream, already embedded) -> x
static int tower_layers(lmrl_engine* e, const TowerDev& T, int mode, int n_env) {
    Gpt2Dev& G = *e->gpt2;
    Lane& Ln = *G.cur;
    const lmrl_gpt2_dims& D = G.dims;
    cudaStream_t st = Ln.st;
    const int ew_blocks = G.sm_count * 8;
    const int att_blocks = std::min((n_env * D.n_head + ATT_WARPS - 1) / ATT_WARPS, G.sm_count * G.att_grid_mult);
    // ln_2 of this layer rides on the attention-projection GEMM, ln_1 of the next layer on the MLP-projection GEMM (tcgen05
    // one-CTA kernel only); layer 0's ln_1 is the one LayerNorm launch of the pass
    const bool fuse = G.fuse_ln && G.gemm_impl >= 1 && D.d_model % TC_BN == 0 && Ln.ln_cnt != nullptr;
    const int gv = gemv_rows(G, mode, n_env);
    // consumer-side LayerNorm fold: no LayerNorm launches; stat_n = partial pairs a residual GEMM writes per row (2 per N tile of its config)
    const bool fold = !gv && !fuse && (G.ln_fold == 1 || (G.ln_fold == 2 && mode == 1)) && G.gemm_impl >= 1 && D.d_model % TC_BN == 0 && Ln.stat[0] != nullptr && T.layers[0].w_qkv_f != nullptr;
    LMRL_CHECK(!fold || T.fold_valid, "ln_fold is on but this tower's LayerNorm-folded weight copies are stale (the weights were re-loaded while ln_fold was off; "
               "the copies are refreshed on every load while it is on): load the weights again (set_params / the trainer's hand-over) with ln_fold set");
    const int resid_parts = 2 * (D.d_model / ((G.narrow_decode_tiles && mode == 1 && D.d_model == 768) ? 192 : TC_BN));
    int n_a = 1;                                                            // partial pairs behind ln_1's input (layer 0: one full-row pair)
    if (fold) LMRL_LAUNCH(2, LMRL_K(xb_stats_kernel, ew_blocks, 256, 0, st, Ln.x, D.d_model, Ln.rows.M, Ln.xn, Ln.stat[0]));
    for (int l = 0; l < D.n_layer; ++l) {
        const LayerDev& L = T.layers[l];
        if (fold) {
            GemmArgs a{};
            a.A = Ln.xn; a.W = L.w_qkv_f; a.bias = L.b_qkv_f; a.N = 3 * D.d_model; a.K = D.d_model; a.M_dev = Ln.rows.M;
            a.out = Ln.q; a.kcache = T.kv + (size_t)l * G.kv_layer_stride; a.vcache = a.kcache + G.kv_layer_stride / 2;
            a.row_env = Ln.rows.row_env; a.row_pos = Ln.rows.row_pos; a.d_model = D.d_model; a.n_head = D.n_head; a.kv_table = G.kvp.table; a.kv_ppe = G.kvp.ppe;
            a.stat_in = Ln.stat[0]; a.stat_n = n_a; a.ln_eps = D.ln_eps;
            if (launch_gemm<EPI_QKV>(e, a, a.N, 0, Ln.m_cap, a.N, 0)) return -1;
            if (e->prof_begin(1, 4.0 * D.d_model, 2, 10 + mode, 4.0 * D.d_model, 0)) return -1;
            if (mode == 0 || G.att_decode_impl == 1)
                LMRL_LAUNCH(-1, LMRL_K(attention_prefill_kernel, attp_grid(n_env, D.n_head, G.sm_count), ATTP_WARPS * 32, ATTP_WARPS * ATTP_WARP_SMEM, st, Ln.q, a.kcache, a.vcache, Ln.rows,
                                       D.n_head, D.d_model, G.kvp, 0.125f * 1.4426950408889634f, G.attn_wide, Ln.att));
            else if (G.att_decode_impl == 2)
                LMRL_LAUNCH(-1, LMRL_K(attention_pipe_kernel, att_blocks, ATT_WARPS * 32, 0, st, Ln.q, a.kcache, a.vcache, Ln.rows, D.n_head, D.d_model, G.kvp, 0.125f, Ln.att));
            else if (G.att_decode_impl == 3 && D.n_head % 2 == 0)
                LMRL_LAUNCH(-1, LMRL_K(attention_hpw_kernel<2>, std::min(att_blocks, (n_env * D.n_head / 2 + ATT_WARPS - 1) / ATT_WARPS), ATT_WARPS * 32, 0, st, Ln.q, a.kcache, a.vcache, Ln.rows, D.n_head, D.d_model, G.kvp, 0.125f, Ln.att));
            else if (G.att_decode_impl == 4 && D.n_head % 4 == 0)
                LMRL_LAUNCH(-1, LMRL_K(attention_hpw_kernel<4>, std::min(att_blocks, (n_env * D.n_head / 4 + ATT_WARPS - 1) / ATT_WARPS), ATT_WARPS * 32, 0, st, Ln.q, a.kcache, a.vcache, Ln.rows, D.n_head, D.d_model, G.kvp, 0.125f, Ln.att));
            else
                LMRL_LAUNCH(-1, LMRL_K(attention_kernel, att_blocks, ATT_WARPS * 32, 0, st, Ln.q, a.kcache, a.vcache, Ln.rows, D.n_head, D.d_model, G.kvp, 0.125f, Ln.att));
            GemmArgs p{};
            p.A = Ln.att; p.W = L.w_proj; p.bias = L.b_proj; p.N = D.d_model; p.K = D.d_model; p.M_dev = Ln.rows.M; p.resid = Ln.x;
            p.xb_out = Ln.xn; p.stat_out = Ln.stat[1];
            if (launch_gemm<EPI_RESID>(e, p, p.N, 0, Ln.m_cap, p.N, 1)) return -1;
            GemmArgs f{};
            f.A = Ln.xn; f.W = L.w_fc_f; f.bias = L.b_fc_f; f.N = D.d_ff; f.K = D.d_model; f.M_dev = Ln.rows.M; f.out = Ln.h;
            f.stat_in = Ln.stat[1]; f.stat_n = resid_parts; f.ln_eps = D.ln_eps;
            if (launch_gemm<EPI_GELU>(e, f, f.N, 0, Ln.m_cap, f.N, 2)) return -1;
            GemmArgs f2{};
            f2.A = Ln.h; f2.W = L.w_fc2; f2.bias = L.b_fc2; f2.N = D.d_model; f2.K = D.d_ff; f2.M_dev = Ln.rows.M; f2.resid = Ln.x;
            if (l + 1 < D.n_layer) { f2.xb_out = Ln.xn; f2.stat_out = Ln.stat[0]; }
            if (launch_gemm<EPI_RESID>(e, f2, f2.N, 0, Ln.m_cap, f2.N, 3)) return -1;
            n_a = resid_parts;
            continue;
        }
        if (gv) {
            // small-batch decode pass: five launches per layer, the two LayerNorms folded into the QKV / FC GEMVs' operand staging
            __nv_bfloat16* kc = T.kv + (size_t)l * G.kv_layer_stride;
            GemvArgs a{};
            a.x_f32 = Ln.x; a.ln_gamma = L.ln1_g; a.ln_beta = L.ln1_b; a.ln_eps = D.ln_eps;
            a.W = L.w_qkv; a.bias = L.b_qkv; a.N = 3 * D.d_model; a.K = D.d_model; a.M_dev = Ln.rows.M; a.out = Ln.q;
            a.kcache = kc; a.vcache = kc + G.kv_layer_stride / 2; a.row_env = Ln.rows.row_env; a.row_pos = Ln.rows.row_pos;
            a.kv_table = G.kvp.table; a.kv_ppe = G.kvp.ppe; a.d_model = D.d_model; a.n_head = D.n_head;
            if (launch_gemv<EPI_QKV, true>(e, a, gv, a.N, 0, 0)) return -1;
            if (e->prof_begin(1, 4.0 * D.d_model, 2, 10 + mode, 4.0 * D.d_model, 0)) return -1;
            LMRL_LAUNCH(-1, LMRL_K(attention_kernel, att_blocks, ATT_WARPS * 32, 0, st, Ln.q, a.kcache, a.vcache, Ln.rows, D.n_head, D.d_model, G.kvp, 0.125f, Ln.att));
            GemvArgs p{};
            p.A = Ln.att; p.a_pitch = D.d_model; p.W = L.w_proj; p.bias = L.b_proj; p.N = D.d_model; p.K = D.d_model; p.M_dev = Ln.rows.M; p.resid = Ln.x;
            if (launch_gemv<EPI_RESID, false>(e, p, gv, p.N, 0, 1)) return -1;
            GemvArgs f{};
            f.x_f32 = Ln.x; f.ln_gamma = L.ln2_g; f.ln_beta = L.ln2_b; f.ln_eps = D.ln_eps;
            f.W = L.w_fc; f.bias = L.b_fc; f.N = D.d_ff; f.K = D.d_model; f.M_dev = Ln.rows.M; f.out = Ln.h;
            if (launch_gemv<EPI_GELU, true>(e, f, gv, f.N, 0, 2)) return -1;
            GemvArgs f2{};
            f2.A = Ln.h; f2.a_pitch = D.d_ff; f2.W = L.w_fc2; f2.bias = L.b_fc2; f2.N = D.d_model; f2.K = D.d_ff; f2.M_dev = Ln.rows.M; f2.resid = Ln.x;
            if (launch_gemv<EPI_RESID, false>(e, f2, gv, f2.N, 0, 3)) return -1;
            continue;
        }
        if (!fuse || l == 0)
            LMRL_LAUNCH(2, LMRL_K(layernorm_kernel, ew_blocks, 256, 0, st, Ln.x, L.ln1_g, L.ln1_b, D.d_model, D.ln_eps, Ln.rows.M, nullptr, Ln.xn, D.d_model));
        GemmArgs a{};
        a.A = Ln.xn; a.W = L.w_qkv; a.bias = L.b_qkv; a.N = 3 * D.d_model; a.K = D.d_model; a.M_dev = Ln.rows.M;
        a.out = Ln.q; a.kcache = T.kv + (size_t)l * G.kv_layer_stride; a.vcache = a.kcache + G.kv_layer_stride / 2;
        a.row_env = Ln.rows.row_env; a.row_pos = Ln.rows.row_pos; a.d_model = D.d_model; a.n_head = D.n_head; a.kv_table = G.kvp.table; a.kv_ppe = G.kvp.ppe;
        if (launch_gemm<EPI_QKV>(e, a, a.N, 0, Ln.m_cap, a.N, 0)) return -1;
        // algorithmic bytes: every cached K and V position of the contributing envs once (bf16) + q in / out rows
        if (e->prof_begin(1, 4.0 * D.d_model, 2, 10 + mode, 4.0 * D.d_model, 0)) return -1;
        if (mode == 0 || G.att_decode_impl == 1)
            LMRL_LAUNCH(-1, LMRL_K(attention_prefill_kernel, attp_grid(n_env, D.n_head, G.sm_count), ATTP_WARPS * 32, ATTP_WARPS * ATTP_WARP_SMEM, st, Ln.q, a.kcache, a.vcache, Ln.rows, D.n_head, D.d_model, G.kvp,
                                                                                          0.125f * 1.4426950408889634f, G.attn_wide, Ln.att));
        else if (G.att_decode_impl == 2)
            LMRL_LAUNCH(-1, LMRL_K(attention_pipe_kernel, att_blocks, ATT_WARPS * 32, 0, st, Ln.q, a.kcache, a.vcache, Ln.rows, D.n_head, D.d_model, G.kvp,
                                                                                                  0.125f, Ln.att));
        else if (G.att_decode_impl == 3 && D.n_head % 2 == 0)
            LMRL_LAUNCH(-1, LMRL_K(attention_hpw_kernel<2>, std::min(att_blocks, (n_env * D.n_head / 2 + ATT_WARPS - 1) / ATT_WARPS), ATT_WARPS * 32, 0, st, Ln.q, a.kcache, a.vcache, Ln.rows, D.n_head, D.d_model, G.kvp, 0.125f, Ln.att));
        else if (G.att_decode_impl == 4 && D.n_head % 4 == 0)
            LMRL_LAUNCH(-1, LMRL_K(attention_hpw_kernel<4>, std::min(att_blocks, (n_env * D.n_head / 4 + ATT_WARPS - 1) / ATT_WARPS), ATT_WARPS * 32, 0, st, Ln.q, a.kcache, a.vcache, Ln.rows, D.n_head, D.d_model, G.kvp, 0.125f, Ln.att));
        else
            LMRL_LAUNCH(-1, LMRL_K(attention_kernel, att_blocks, ATT_WARPS * 32, 0, st, Ln.q, a.kcache, a.vcache, Ln.rows, D.n_head, D.d_model, G.kvp,
                                                                                             0.125f, Ln.att));
        GemmArgs p{};
        p.A = Ln.att; p.W = L.w_proj; p.bias = L.b_proj; p.N = D.d_model; p.K = D.d_model; p.M_dev = Ln.rows.M; p.resid = Ln.x;
        if (fuse) { p.ln_gamma = L.ln2_g; p.ln_beta = L.ln2_b; p.ln_out = Ln.xn; p.ln_cnt = Ln.ln_cnt; p.ln_eps = D.ln_eps; }
        if (launch_gemm<EPI_RESID>(e, p, p.N, 0, Ln.m_cap, p.N, 1)) return -1;
        if (!fuse)
            LMRL_LAUNCH(2, LMRL_K(layernorm_kernel, ew_blocks, 256, 0, st, Ln.x, L.ln2_g, L.ln2_b, D.d_model, D.ln_eps, Ln.rows.M, nullptr, Ln.xn, D.d_model));
        GemmArgs f{};
        f.A = Ln.xn; f.W = L.w_fc; f.bias = L.b_fc; f.N = D.d_ff; f.K = D.d_model; f.M_dev = Ln.rows.M; f.out = Ln.h;
        if (launch_gemm<EPI_GELU>(e, f, f.N, 0, Ln.m_cap, f.N, 2)) return -1;
        GemmArgs f2{};
        f2.A = Ln.h; f2.W = L.w_fc2; f2.bias = L.b_fc2; f2.N = D.d_model; f2.K = D.d_ff; f2.M_dev = Ln.rows.M; f2.resid = Ln.x;
        if (fuse && l + 1 < D.n_layer) {
            const LayerDev& Nx = T.layers[l + 1];
            f2.ln_gamma = Nx.ln1_g; f2.ln_beta = Nx.ln1_b; f2.ln_out = Ln.xn; f2.ln_cnt = Ln.ln_cnt; f2.ln_eps = D.ln_eps;
        }
        if (launch_gemm<EPI_RESID>(e, f2, f2.N, 0, Ln.m_cap, f2.N, 3)) return -1;
    }
    return 0;
}

// One forward pass over the current ragged batch + sampling of one token per contributing env.
// mode 0 = prefill of un-cached stream tokens, mode 1 = single-token decode.
// score != nullptr: no sampling; the LM head runs over score->rows and writes log-probabilities of score->targets instead.
static int forward_and_sample(lmrl_engine* e, int mode, int n_env, const uint8_t* done, const int32_t* turn_of_env, int turn_const,
                              const lmrl_sampling& s, float* logits_out, const ScoreReq* score = nullptr) {
    Gpt2Dev& G = *e->gpt2;
    Lane& Ln = *G.cur;
    const lmrl_gpt2_dims& D = G.dims;
    cudaStream_t st = Ln.st;
    const int ew_blocks = G.sm_count * 8;
    e->cur_pass_slot = -1;
    e->cur_mode = mode;
    LMRL_LAUNCH(2, LMRL_K(build_rows_kernel, 1, 1024, 0, st, mode, Ln.e0, n_env, done, G.stream.tokens, G.stream.len, G.stream.kv_len, G.stream.cap, G.gen, Ln.rows, G.kvp, G.fold_nlnl));
    note_pass_counts(e);
    if (mode == 0)
        LMRL_LAUNCH(2, LMRL_K(fill_prefill_rows_kernel, (n_env * 32 + 255) / 256, 256, 0, st, G.stream.tokens, G.stream.len, G.stream.kv_len, G.stream.cap, G.gen, Ln.rows, G.kvp, G.fold_nlnl));
    // towers of this policy: 0 = LM policy; 0 and 1 = pi_beta + value base (kind 1); 1 alone = value base without pi_beta (kind 2)
    const int t_first = G.policy_kind == 2 ? 1 : 0, t_last = G.policy_kind >= 1 ? 1 : 0;
    LMRL_LAUNCH(2, LMRL_K(embed_kernel, ew_blocks, 256, 0, st, G.tw[t_first].wte, G.tw[t_first].wpe, D.d_model, Ln.rows, Ln.x));
    // head operand buffers: the per-env sampling buffers, or the (larger) scoring buffers
    __nv_bfloat16* xl = score ? G.xs : Ln.xl;
    __nv_bfloat16* xl1 = score ? G.xs1 : Ln.xl1;
    const size_t xl_rows = score ? (size_t)G.s_cap : Ln.xl_rows;
    const int* head_n = score ? score->n_dev : Ln.rows.n_slots;
    const int* head_gather = score ? score->rows : Ln.rows.slot_last;
    // small-batch decode pass of the plain LM policy: the final LayerNorm and the LM head + sampler are one GEMV launch
    const int gv_head = (!score && G.policy_kind == 0) ? gemv_rows(G, mode, n_env) : 0;
    for (int t = t_first; t <= t_last; ++t) {
        const TowerDev& T = G.tw[t];
        if (t > t_first) LMRL_LAUNCH(2, LMRL_K(embed_kernel, ew_blocks, 256, 0, st, T.wte, T.wpe, D.d_model, Ln.rows, Ln.x));
        if (tower_layers(e, T, mode, n_env)) return -1;
        if (gv_head) break;
        // final LayerNorm of the sampled rows only: tower 0 -> xl[:, 0:d], value tower -> xl1 (MLP Q heads: operand of their first
        // layer) or straight into xl[:, d:2d] (LinearHead Q heads: operand of the head GEMM's second K segment)
        const bool to_xl = t == 0 || G.qh.linear;
        LMRL_LAUNCH(2, LMRL_K(layernorm_kernel, ew_blocks, 256, 0, st, Ln.x, T.lnf_g, T.lnf_b, D.d_model, D.ln_eps, head_n, head_gather,
                                                                  t == 0 ? xl : (G.qh.linear ? xl + D.d_model : xl1), to_xl ? G.xl_cols : D.d_model));
    }
    // top-k / top-p: the LM head dumps fp32 logits and a per-row filter kernel samples from the survivors (slow path)
    const bool filtered = !score && (s.top_k > 0 || (s.top_p > 0.f && s.top_p < 1.f));
    if (filtered && !logits_out) {
        const size_t need = (size_t)e->cfg.max_envs * D.n_vocab;
        if (G.logits_rows < need) {
            LMRL_CUDA(cudaMalloc(&G.logits, need * sizeof(float)));
            G.owned.push_back(G.logits);
            G.logits_rows = need;
        }
        logits_out = G.logits;
    }
    if (!score)
        LMRL_LAUNCH(2, LMRL_K(prepare_sample_kernel, (n_env + 255) / 256, 256, 0, st, Ln.rows, G.gen, G.episode_key, turn_of_env, turn_const, s.seed, Ln.row_key, Ln.best));
    if (gv_head) {
        GemvArgs hv{};
        hv.x_f32 = Ln.x; hv.ln_gamma = G.tw[0].lnf_g; hv.ln_beta = G.tw[0].lnf_b; hv.ln_eps = D.ln_eps; hv.gather = head_gather;
        hv.W = G.tw[0].wte; hv.N = G.v_pad; hv.K = D.d_model; hv.M_dev = head_n;
        hv.best = filtered ? Ln.best_scratch : Ln.best; hv.row_key = Ln.row_key; hv.inv_temp = 1.0f / (s.temperature > 0.f ? s.temperature : 1.0f);
        hv.do_sample = s.do_sample; hv.n_vocab = D.n_vocab; hv.logits_out = logits_out;
        if (launch_gemv<EPI_SAMPLE, true>(e, hv, gv_head, D.n_vocab, 1, 4)) return -1;
        if (filtered) {
            static bool configured_gv = false;
            const size_t fs_smem = (size_t)D.n_vocab * sizeof(float);
            LMRL_CHECK(fs_smem <= 220 * 1024, "top-k / top-p sampling keeps one row of logits in shared memory: n_vocab %d is too large", D.n_vocab);
            if (!configured_gv) {
                LMRL_CUDA(cudaFuncSetAttribute(filtered_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fs_smem));
                configured_gv = true;
            }
            LMRL_LAUNCH(2, LMRL_K(filtered_sample_kernel, n_env, FS_THREADS, fs_smem, st, logits_out, D.n_vocab, Ln.rows, Ln.row_key, hv.inv_temp, s.do_sample,
                                                                                    s.top_k, s.top_p, Ln.best));
        }
        LMRL_LAUNCH(2, LMRL_K(finalize_sample_kernel, (n_env + 255) / 256, 256, 0, st, Ln.rows, G.gen, Ln.best));
        e->cur_pass_slot = -1;
        LMRL_CUDA(cudaGetLastError());
        return 0;
    }
    GemmArgs h{};
    h.A = xl; h.W = G.tw[0].wte; h.bias = nullptr; h.N = G.v_pad; h.K = D.d_model; h.M_dev = head_n;
    h.best = filtered ? Ln.best_scratch : Ln.best; h.row_key = Ln.row_key; h.inv_temp = 1.0f / (s.temperature > 0.f ? s.temperature : 1.0f);
    h.do_sample = s.do_sample;
    h.n_vocab = D.n_vocab; h.logits_out = logits_out;
    GemmExtra hx;
    hx.a_cols = (size_t)G.xl_cols;
    const int kb_d = D.d_model / TC_BK;
    if (G.policy_kind >= 1) {
        // Q heads, first layer: xl[:, d + k*hidden ...] = beta * relu(xl1 · W1_k^T + b1_k)   (mlp_head.py:146-147; beta folded in here)
        const QHeadsDev& Q = G.qh;
        if (!Q.linear) {
            GemmArgs r{};
            r.A = xl1; r.W = Q.w1cat; r.bias = Q.b1cat; r.N = Q.n_q * Q.hidden; r.K = D.d_model; r.M_dev = head_n; r.out = xl;
            r.out_col0 = D.d_model; r.out_scale = Q.beta;
            GemmExtra rx;
            rx.out_cols = (size_t)G.xl_cols;
            if (launch_gemm<EPI_RELU>(e, r, r.N, 1, xl_rows, (size_t)r.N, 4, rx)) return -1;
        }
        // logits = min_k(h_pi · wte^T + beta * (relu_k · W2_k^T + b2_k))   (generation.py:112-119); LinearHead Q heads: both heads read
        // the value tower's hidden h in xl[:, d:2d] and their kernels carry beta
        const int kb_h = Q.hidden / TC_BK;
        const int a2 = Q.linear ? kb_d : kb_d + kb_h;          // A-operand column block of the second head
        h.n_acc = Q.n_q; h.n_seg = 2 * Q.n_q;
        h.seg[0] = {0, 0, 0, kb_d}; h.seg[1] = {0, kb_d, 1, kb_h};
        h.seg[2] = {1, 0, 0, kb_d}; h.seg[3] = {1, a2, 2, kb_h};
        if (G.policy_kind == 2) {                     // pi_beta is None: logits = beta * min_k q_k (generation.py:118-119) — no LM-head segment
            h.n_seg = Q.n_q;
            h.seg[0] = {0, kb_d, 1, kb_h}; h.seg[1] = {1, a2, 2, kb_h};
        }
        h.bias = Q.b2s[0]; h.bias2 = Q.b2s[1];
        hx.W1 = Q.w2[0]; hx.w1_cols = (size_t)Q.hidden;
        if (Q.n_q == 2) { hx.W2 = Q.w2[1]; hx.w2_cols = (size_t)Q.hidden; }
    }
    if (score) {
        h.lp_partial = G.lp_partial; h.lp_target = score->targets; h.lp_tlogit = G.lp_tlogit;
        if (launch_gemm<EPI_LOGPROB>(e, h, D.n_vocab, 1, xl_rows, G.v_pad, 4, hx)) return -1;
        LMRL_LAUNCH(2, LMRL_K(logprob_finalize_kernel, std::max(1, std::min((score->n_score + 7) / 8, G.sm_count * 8)), 256, 0, st, 
                           G.lp_partial, 2 * (G.v_pad / TC_BN), G.lp_tlogit, score->n_dev, score->out));
    } else {
        if (launch_gemm<EPI_SAMPLE>(e, h, D.n_vocab, 1, xl_rows, G.v_pad, 4, hx)) return -1;
        if (filtered) {
            static bool configured = false;
            const size_t fs_smem = (size_t)D.n_vocab * sizeof(float);
            LMRL_CHECK(fs_smem <= 220 * 1024, "top-k / top-p sampling keeps one row of logits in shared memory: n_vocab %d is too large", D.n_vocab);
            if (!configured) {
                LMRL_CUDA(cudaFuncSetAttribute(filtered_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fs_smem));
                configured = true;
            }
            LMRL_LAUNCH(2, LMRL_K(filtered_sample_kernel, n_env, FS_THREADS, fs_smem, st, logits_out, D.n_vocab, Ln.rows, Ln.row_key, h.inv_temp, s.do_sample,
                                                                                    s.top_k, s.top_p, Ln.best));
        }
        LMRL_LAUNCH(2, LMRL_K(finalize_sample_kernel, (n_env + 255) / 256, 256, 0, st, Ln.rows, G.gen, Ln.best));
    }
    e->cur_pass_slot = -1;
    LMRL_CUDA(cudaGetLastError());
    return 0;
}

static int gpt2_alloc_state(lmrl_engine* e) {
    Gpt2Dev& G = *e->gpt2;
    const lmrl_gpt2_dims& D = G.dims;
    const int B = e->cfg.max_envs;
    const int max_in = e->cfg.max_input_length, max_new = e->cfg.max_new_tokens, T = e->cfg.max_turns;
    G.t_cap = (max_in + max_new + 7) / 8 * 8;
    G.stream_cap = max_in + MAX_LINE_TOKENS;
    G.m_cap = ((size_t)B * max_in + 127) / 128 * 128;
    // paged KV: pages of KV_PAGE positions; the pool defaults to every env at full length (it can never run out), lmrl_config.kv_pages
    // sizes it for the lengths a workload really reaches (the row builder raises an error flag when the pool is exhausted)
    G.kvp.ppe = (G.t_cap + KV_PAGE - 1) / KV_PAGE;
    LMRL_CHECK(G.kvp.ppe <= 32, "max_input_length + max_new_tokens = %d exceeds %d KV positions per env", G.t_cap, 32 * KV_PAGE);
    G.kvp.n_pages = e->cfg.kv_pages > 0 ? e->cfg.kv_pages : B * G.kvp.ppe;
    LMRL_CHECK(G.kvp.n_pages >= G.kvp.ppe, "lmrl_config.kv_pages = %d is smaller than one env's %d pages", G.kvp.n_pages, G.kvp.ppe);
    if (e->alloc(&G.kvp.table, (size_t)B * G.kvp.ppe) || e->alloc(&G.kvp.n_alloc, 1) || e->alloc(&G.kvp.overflow, 1)) return -1;
    LMRL_CUDA(cudaMemsetAsync(G.kvp.table, 0xFF, sizeof(int) * (size_t)B * G.kvp.ppe, e->stream));
    G.kv_layer_stride = (size_t)2 * G.kvp.n_pages * D.d_model * KV_PAGE;
    if (e->alloc(&G.x, (size_t)G.m_cap * D.d_model) || e->alloc(&G.xn, (size_t)G.m_cap * D.d_model) ||
        e->alloc(&G.q, (size_t)G.m_cap * D.d_model) || e->alloc(&G.att, (size_t)G.m_cap * D.d_model) ||
        e->alloc(&G.h, (size_t)G.m_cap * D.d_ff) || e->alloc(&G.xl, (size_t)((B + 127) / 128 * 128) * 3 * D.d_model) ||
        e->alloc(&G.xl1, (size_t)((B + 127) / 128 * 128) * D.d_model))
        return -1;
    G.xl_cols = 3 * D.d_model;
    if (e->alloc(&G.rows.M, 1) || e->alloc(&G.rows.n_slots, 1) || e->alloc(&G.rows.ctx_total, 1) || e->alloc(&G.rows.row_env, G.m_cap) || e->alloc(&G.rows.row_pos, G.m_cap) ||
        e->alloc(&G.rows.row_tok, G.m_cap) || e->alloc(&G.rows.slot_env, B) || e->alloc(&G.rows.slot_row0, B) ||
        e->alloc(&G.rows.slot_nrow, B) || e->alloc(&G.rows.slot_last, B))
        return -1;
    if (e->alloc(&G.gen.tok, (size_t)B * max_new) || e->alloc(&G.gen.len, B) || e->alloc(&G.gen.active, B) || e->alloc(&G.gen.cur, B)) return -1;
    G.gen.max_new = max_new;
    G.gen.eos_id = e->tok.t.newline_id;
    if (e->alloc(&G.stream.tokens, (size_t)B * G.stream_cap) || e->alloc(&G.stream.len, B) || e->alloc(&G.stream.kv_len, B) ||
        e->alloc(&G.stream.is_action, (size_t)B * G.stream_cap) || e->alloc(&G.stream.tok_reward, (size_t)B * G.stream_cap) || e->alloc(&G.stream.truncated, B))
        return -1;
    G.stream.cap = G.stream_cap; G.stream.max_input = max_in;
    if (e->alloc(&G.row_key, B) || e->alloc(&G.best, B) || e->alloc(&G.best_scratch, B) || e->alloc(&G.episode_key, B) || e->alloc(&G.live, 1) ||
        e->alloc(&G.gen_turn, B))
        return -1;
    if (e->alloc(&G.ep_return, B) || e->alloc(&G.ep_length, B) || e->alloc(&G.ep_done, B) || e->alloc(&G.step_reward, (size_t)B * T) ||
        e->alloc(&G.step_valid, (size_t)B * T) || e->alloc(&G.step_fb, (size_t)B * T * 5) || e->alloc(&G.step_aux, (size_t)B * T) ||
        e->alloc(&G.act_tokens, (size_t)B * T * max_new) || e->alloc(&G.act_len, (size_t)B * T))
        return -1;
    LMRL_CUDA(cudaFuncSetAttribute(attention_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTP_WARPS * ATTP_WARP_SMEM));
    // lanes: 0 = the whole batch; 1 / 2 = the two halves of the same buffers (usable when both halves stay tile-aligned)
    Lane& L0 = G.lane[0];
    L0.x = G.x; L0.xn = G.xn; L0.q = G.q; L0.att = G.att; L0.h = G.h; L0.xl = G.xl; L0.xl1 = G.xl1; L0.rows = G.rows; L0.row_key = G.row_key;
    L0.best = G.best; L0.best_scratch = G.best_scratch; L0.st = e->stream; L0.e0 = 0; L0.m_cap = (size_t)G.m_cap; L0.xl_rows = (size_t)(B + 127) / 128 * 128;
    G.cur = &G.lane[0];
    const size_t m_tiles = ((size_t)G.m_cap + TC_BM - 1) / TC_BM + 1;
    if (e->alloc(&L0.ln_cnt, m_tiles) || e->alloc(&G.lane[1].ln_cnt, m_tiles) || e->alloc(&G.lane[2].ln_cnt, m_tiles)) return -1;
    if (e->alloc(&L0.stat[0], (size_t)G.m_cap * 8) || e->alloc(&L0.stat[1], (size_t)G.m_cap * 8)) return -1;
    if (e->alloc(&G.lane_counts, 6)) return -1;
    LMRL_CUDA(cudaStreamCreateWithFlags(&G.stream_b, cudaStreamNonBlocking));
    LMRL_CUDA(cudaEventCreateWithFlags(&G.ev_fork, cudaEventDisableTiming));
    LMRL_CUDA(cudaEventCreateWithFlags(&G.ev_join, cudaEventDisableTiming));
    const size_t half_rows = (size_t)(B / 2) * max_in, hb = (size_t)(B / 2);
    for (int k = 1; k <= 2; ++k) {
        Lane& L = G.lane[k];
        const size_t ro = k == 2 ? half_rows : 0, so = k == 2 ? hb : 0;      // row / slot offset of the half
        L.x = G.x + ro * D.d_model; L.xn = G.xn + ro * D.d_model; L.q = G.q + ro * D.d_model; L.att = G.att + ro * D.d_model; L.h = G.h + ro * D.d_ff;
        L.xl = G.xl + so * G.xl_cols; L.xl1 = G.xl1 + so * D.d_model;
        L.rows = G.rows;
        L.rows.M = G.lane_counts + 3 * (k - 1); L.rows.n_slots = L.rows.M + 1; L.rows.ctx_total = L.rows.M + 2;
        L.rows.row_env += ro; L.rows.row_pos += ro; L.rows.row_tok += ro;
        L.rows.slot_env += so; L.rows.slot_row0 += so; L.rows.slot_nrow += so; L.rows.slot_last += so;
        L.row_key = G.row_key + so; L.best = G.best + so; L.best_scratch = G.best_scratch + so;
        L.st = k == 2 ? G.stream_b : e->stream;
        L.stat[0] = G.lane[0].stat[0] + ro * 8; L.stat[1] = G.lane[0].stat[1] + ro * 8;
        L.m_cap = half_rows; L.xl_rows = hb;
    }
    return 0;
}

// two half-batch lanes need tile-aligned halves: B/2 slots a multiple of 128 (xl rows) and (B/2) * max_input rows a multiple of 128
static bool lanes_usable(const lmrl_engine* e) {
    const int B = e->cfg.max_envs;
    return B % 256 == 0 && ((size_t)(B / 2) * e->cfg.max_input_length) % 128 == 0;
}

}  // namespace lmrl

using namespace lmrl;

void lmrl_gpt2_free(lmrl_engine* e) {
    if (!e || !e->gpt2) return;
    if (e->gpt2->graph_exec) cudaGraphExecDestroy(e->gpt2->graph_exec);
    if (e->gpt2->decode_graph) cudaGraphExecDestroy(e->gpt2->decode_graph);
    if (e->gpt2->train) lmrl_train_free(e->gpt2);
    if (e->gpt2->stream_b) cudaStreamDestroy(e->gpt2->stream_b);
    if (e->gpt2->ev_fork) cudaEventDestroy(e->gpt2->ev_fork);
    if (e->gpt2->ev_join) cudaEventDestroy(e->gpt2->ev_join);
    for (void* p : e->gpt2->owned) cudaFree(p);
    delete e->gpt2;
    e->gpt2 = nullptr;
}

extern "C" {

int lmrl_load_gpt2(lmrl_engine* e, const lmrl_gpt2_dims* dims, const lmrl_gpt2_weights* w) { return lmrl_load_gpt2_tower(e, 0, dims, w); }

static int* option_slot(lmrl::Gpt2Dev& G, const char* name);

int lmrl_load_gpt2_tower(lmrl_engine* e, int32_t tower, const lmrl_gpt2_dims* dims, const lmrl_gpt2_weights* w) {
    LMRL_CHECK(e && dims && w, "lmrl_load_gpt2: null argument");
    LMRL_CHECK(tower >= 0 && tower <= 2, "lmrl_load_gpt2_tower: tower must be 0 (policy / pi_beta), 1 (value base / initial policy) or 2 (target value base)");
    LMRL_CHECK(tower == 0 || (e->gpt2 && e->gpt2->tw[0].loaded), "lmrl_load_gpt2_tower: load tower 0 first");
    LMRL_CHECK(e->has_tok, "lmrl_load_gpt2: load the tokenizer tables first (EOS id, header)");
    LMRL_CHECK(dims->d_model % 128 == 0 && dims->d_model <= 1024 && dims->d_model == dims->n_head * 64 && dims->d_ff % 128 == 0,
               "lmrl_load_gpt2: unsupported shape d_model=%d n_head=%d d_ff=%d (need head_dim 64, d_model<=1024, multiples of 128)",
               dims->d_model, dims->n_head, dims->d_ff);
    LMRL_CHECK(w->n_vocab_rows >= dims->n_vocab, "lmrl_load_gpt2: wte has fewer rows than n_vocab");
    LMRL_CHECK(e->cfg.max_input_length + e->cfg.max_new_tokens <= dims->n_pos, "lmrl_load_gpt2: max_input_length + max_new_tokens exceeds n_pos");
    LMRL_CUDA(cudaSetDevice(e->device));
    const bool first = e->gpt2 == nullptr;
    if (first) {
        e->gpt2 = new Gpt2Dev();
        e->gpt2->dims = *dims;
        cudaDeviceProp prop;
        LMRL_CUDA(cudaGetDeviceProperties(&prop, e->device));
        e->gpt2->sm_count = prop.multiProcessorCount;
        // A/B aid: LMRL_OPTIONS="name=value,name=value" presets lmrl_set_option switches for every engine of the process
        if (const char* env = getenv("LMRL_OPTIONS")) {
            std::string all(env);
            for (size_t p0 = 0; p0 < all.size();) {
                size_t p1 = all.find(',', p0);
                if (p1 == std::string::npos) p1 = all.size();
                const std::string kv = all.substr(p0, p1 - p0);
                const size_t eq = kv.find('=');
                if (eq != std::string::npos) {
                    int* slot = option_slot(*e->gpt2, kv.substr(0, eq).c_str());
                    if (!slot) {                                          // nothing is loaded yet: leave the engine as it was
                        delete e->gpt2;
                        e->gpt2 = nullptr;
                        return lmrl::fail("LMRL_OPTIONS: unknown option '%s'", kv.substr(0, eq).c_str());
                    }
                    *slot = atoi(kv.c_str() + eq + 1);
                }
                p0 = p1 + 1;
            }
        }
    } else {
        LMRL_CHECK(memcmp(&e->gpt2->dims, dims, sizeof *dims) == 0, "lmrl_load_gpt2: set_params with a different architecture");
    }
    Gpt2Dev& GG = *e->gpt2;
    TowerDev& G = GG.tw[tower];
    const bool fresh = !G.loaded;
    const int d = dims->d_model, ff = dims->d_ff, L = dims->n_layer;
    GG.v_pad = (dims->n_vocab + TC_BN - 1) / TC_BN * TC_BN;   // LM-head rows padded to the widest N tile
    // staging buffer for fp32 uploads
    const size_t stage_elems = std::max((size_t)w->n_vocab_rows * d, (size_t)d * std::max(ff, 3 * d));
    float* stage = nullptr;
    // device-resident sources (a trainer's fp32 master weights) are converted in place: no staging copy, no per-matrix synchronisation
    auto on_device = [&](const void* p) -> bool {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeDevice;
    };
    const bool dev_src = on_device(w->wte);
    if (!dev_src) LMRL_CUDA(cudaMalloc(&stage, stage_elems * sizeof(float)));
    const float* cur_src = nullptr;
    auto up = [&](const float* host, size_t n) -> int {
        if (dev_src) { cur_src = host; return 0; }
        LMRL_CUDA(cudaMemcpyAsync(stage, host, n * sizeof(float), cudaMemcpyDefault, e->stream));
        cur_src = stage;
        return 0;
    };
    const bool do_fold = fresh || GG.ln_fold != 0;
    std::vector<VecCopy> vec_tab;                                       // device-resident sources: gathered, copied by one kernel at the end
    auto vec = [&](float** dst, const float* host, size_t n) -> int {
        if (fresh && e->alloc(dst, n)) return -1;
        if (dev_src && !fresh && !do_fold) { vec_tab.push_back(VecCopy{host, *dst, (int)n, 0}); return 0; }   // (the fold kernels read these vectors: no deferral then)
        LMRL_CUDA(cudaMemcpyAsync(*dst, host, n * sizeof(float), cudaMemcpyDefault, e->stream));
        return 0;
    };
    auto mat_t = [&](__nv_bfloat16** dst, const float* host, int K, int N) -> int {   // [K][N] fp32 -> [N][K] bf16
        if (fresh && e->alloc(dst, (size_t)K * N)) return -1;
        if (up(host, (size_t)K * N)) return -1;
        transpose_to_bf16_kernel<<<dim3((N + 31) / 32, (K + 31) / 32), dim3(32, 8), 0, e->stream>>>(cur_src, K, N, *dst);
        if (!dev_src) LMRL_CUDA(cudaStreamSynchronize(e->stream));      // the staging buffer is reused by the next matrix
        return 0;
    };
    // [K][N] fp32 -> the plain [N][K] bf16 copy AND the LayerNorm-folded copy + folded bias (gamma / beta / bias already on the device).
    // A weight hand-over from a trainer (every optimizer step) refreshes the folded copies only while ln_fold is on; otherwise they are
    // marked stale and the fold stays off for this tower until a load with ln_fold set
    auto mat_t_fold = [&](__nv_bfloat16** dst, __nv_bfloat16** dst_f, float** bias_f, const float* host, int K, int N, const float* gamma,
                          const float* beta, const float* bias) -> int {
        if (fresh && (e->alloc(dst, (size_t)K * N) || e->alloc(dst_f, (size_t)K * N) || e->alloc(bias_f, (size_t)N))) return -1;
        if (up(host, (size_t)K * N)) return -1;
        transpose_to_bf16_kernel<<<dim3((N + 31) / 32, (K + 31) / 32), dim3(32, 8), 0, e->stream>>>(cur_src, K, N, *dst);
        if (do_fold) fold_ln_weights_kernel<<<N, 256, 0, e->stream>>>(cur_src, K, N, gamma, beta, bias, *dst_f, *bias_f);
        if (!dev_src) LMRL_CUDA(cudaStreamSynchronize(e->stream));      // the staging buffer is reused by the next matrix
        return 0;
    };
    int rc = 0;
    if (fresh) { rc |= e->alloc(&G.wte, (size_t)GG.v_pad * d); rc |= e->alloc(&G.wpe, (size_t)dims->n_pos * d); G.layers.resize(L); }
    if (!rc) {
        const int rows = std::min(w->n_vocab_rows, GG.v_pad);
        rc |= up(w->wte, (size_t)rows * d);
        convert_to_bf16_kernel<<<1024, 256, 0, e->stream>>>(cur_src, (size_t)rows * d, G.wte);
        if (!dev_src) cudaStreamSynchronize(e->stream);
        rc |= up(w->wpe, (size_t)dims->n_pos * d);
        convert_to_bf16_kernel<<<1024, 256, 0, e->stream>>>(cur_src, (size_t)dims->n_pos * d, G.wpe);
        if (!dev_src) cudaStreamSynchronize(e->stream);
        rc |= vec(&G.lnf_g, w->lnf_g, d); rc |= vec(&G.lnf_b, w->lnf_b, d);
    }
    for (int l = 0; l < L && !rc; ++l) {
        LayerDev& Ld = G.layers[l];
        rc |= vec(&Ld.ln1_g, w->ln1_g[l], d); rc |= vec(&Ld.ln1_b, w->ln1_b[l], d);
        rc |= vec(&Ld.ln2_g, w->ln2_g[l], d); rc |= vec(&Ld.ln2_b, w->ln2_b[l], d);
        rc |= vec(&Ld.b_qkv, w->attn_b[l], 3 * d);
        if (!rc) rc |= mat_t_fold(&Ld.w_qkv, &Ld.w_qkv_f, &Ld.b_qkv_f, w->attn_w[l], d, 3 * d, Ld.ln1_g, Ld.ln1_b, Ld.b_qkv);
        rc |= mat_t(&Ld.w_proj, w->proj_w[l], d, d); rc |= vec(&Ld.b_proj, w->proj_b[l], d);
        rc |= vec(&Ld.b_fc, w->fc_b[l], ff);
        if (!rc) rc |= mat_t_fold(&Ld.w_fc, &Ld.w_fc_f, &Ld.b_fc_f, w->fc_w[l], d, ff, Ld.ln2_g, Ld.ln2_b, Ld.b_fc);
        rc |= mat_t(&Ld.w_fc2, w->fc2_w[l], ff, d); rc |= vec(&Ld.b_fc2, w->fc2_b[l], d);
    }
    if (!rc && !vec_tab.empty()) {
        LMRL_CHECK(vec_tab.size() <= 512, "lmrl_load_gpt2: %zu small vectors exceed the copy table", vec_tab.size());
        if (!GG.vec_tab_dev && e->alloc(&GG.vec_tab_dev, 512)) return -1;
        LMRL_CUDA(cudaMemcpyAsync(GG.vec_tab_dev, vec_tab.data(), vec_tab.size() * sizeof(VecCopy), cudaMemcpyHostToDevice, e->stream));
        batched_vec_copy_kernel<<<(unsigned)vec_tab.size(), 256, 0, e->stream>>>(GG.vec_tab_dev);
    }
    cudaStreamSynchronize(e->stream);
    if (stage) cudaFree(stage);
    if (rc) return -1;
    if (first && gpt2_alloc_state(e)) return -1;
    // tower 2 is only ever run by the self-contained per-position forward (kv_len = 0 on entry): it borrows tower 1's slab
    if (fresh && tower == 2 && GG.tw[1].kv) G.kv = GG.tw[1].kv;
    else if (fresh && e->alloc(&G.kv, GG.kv_layer_stride * L)) return -1;
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    LMRL_CUDA(cudaGetLastError());
    G.loaded = true;
    G.fold_valid = do_fold;
    e->has_gpt2 = true;
    if (tower == 0 && GG.train) return train_refresh_weight_copies(e);      // the backward pass keeps HF-layout copies of tower 0
    return 0;
}

int lmrl_load_q_heads(lmrl_engine* e, const lmrl_mlp_head* q1, const lmrl_mlp_head* q2, float beta) {
    LMRL_CHECK(e && e->gpt2 && e->gpt2->tw[0].loaded && e->gpt2->tw[1].loaded, "lmrl_load_q_heads: load pi_beta (tower 0) and the value base (tower 1) first");
    LMRL_CHECK(q1 && q1->w2, "lmrl_load_q_heads: q1 is required");
    const bool linear = q1->w1 == nullptr;          // LinearHead: only the one kernel (w2) and its optional bias (b2)
    LMRL_CHECK(linear || (q1->b1 && q1->b2), "lmrl_load_q_heads: an MLPHead needs w1, b1, w2, b2");
    LMRL_CHECK(!q2 || (q2->w2 && (q2->w1 == nullptr) == linear && (linear || (q2->b1 && q2->b2))), "lmrl_load_q_heads: q1 and q2 must be heads of the same kind");
    LMRL_CUDA(cudaSetDevice(e->device));
    Gpt2Dev& G = *e->gpt2;
    QHeadsDev& Q = G.qh;
    const int d = G.dims.d_model, hid = q1->hidden, n_q = q2 ? 2 : 1;
    LMRL_CHECK(hid == d, "lmrl_load_q_heads: hidden_dim %d must equal d_model %d (MLPHeadConfig of train_ilql_gpt2.py:224-231; a LinearHead passes d_model)", hid, d);
    LMRL_CHECK(!q2 || q2->hidden == hid, "lmrl_load_q_heads: q1 and q2 differ in hidden_dim");
    LMRL_CHECK(q1->n_out >= G.dims.n_vocab && (!q2 || q2->n_out >= G.dims.n_vocab), "lmrl_load_q_heads: head output_dim is smaller than n_vocab");
    LMRL_CHECK(Q.n_q == 0 || (Q.n_q == n_q && Q.linear == (linear ? 1 : 0)), "lmrl_load_q_heads: set_params with a different number or kind of Q heads");
    const bool fresh = Q.n_q == 0;
    if (fresh) {
        if (!linear && (e->alloc(&Q.w1cat, (size_t)n_q * hid * d) || e->alloc(&Q.b1cat, (size_t)n_q * hid))) return -1;
        for (int k = 0; k < n_q; ++k)
            if (e->alloc(&Q.w2[k], (size_t)G.v_pad * hid) || e->alloc(&Q.b2s[k], (size_t)G.v_pad)) return -1;
    }
    float* stage = nullptr;
    LMRL_CUDA(cudaMalloc(&stage, (size_t)hid * std::max(q1->n_out, q2 ? q2->n_out : 0) * sizeof(float)));
    const lmrl_mlp_head* qs[2] = {q1, q2};
    for (int k = 0; k < n_q; ++k) {
        const lmrl_mlp_head* q = qs[k];
        if (!linear) {
            // dense1 kernel [d][hid] -> rows [k*hid, (k+1)*hid) of w1cat [.][d]
            LMRL_CUDA(cudaMemcpyAsync(stage, q->w1, (size_t)d * hid * sizeof(float), cudaMemcpyHostToDevice, e->stream));
            transpose_to_bf16_kernel<<<dim3((hid + 31) / 32, (d + 31) / 32), dim3(32, 8), 0, e->stream>>>(stage, d, hid, Q.w1cat + (size_t)k * hid * d);
            LMRL_CUDA(cudaStreamSynchronize(e->stream));
            LMRL_CUDA(cudaMemcpyAsync(Q.b1cat + (size_t)k * hid, q->b1, (size_t)hid * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        }
        // dense2 (or LinearHead) kernel [hid][n_out] -> w2 [v_pad][hid] (rows beyond n_out stay zero; ids beyond n_vocab are masked by the
        // sampler).  MLPHead: beta is folded into the first layer's output; LinearHead: into this kernel (one bf16 rounding of beta * w).
        const int n_rows = std::min(q->n_out, G.v_pad);
        if (linear) {
            std::vector<float> scaled((size_t)hid * q->n_out);
            for (size_t i = 0; i < scaled.size(); ++i) scaled[i] = beta * q->w2[i];
            LMRL_CUDA(cudaMemcpyAsync(stage, scaled.data(), scaled.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream));
            LMRL_CUDA(cudaStreamSynchronize(e->stream));
        } else {
            LMRL_CUDA(cudaMemcpyAsync(stage, q->w2, (size_t)hid * q->n_out * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        }
        transpose_to_bf16_kernel<<<dim3((n_rows + 31) / 32, (hid + 31) / 32), dim3(32, 8), 0, e->stream>>>(stage, hid, q->n_out, Q.w2[k], n_rows);
        LMRL_CUDA(cudaStreamSynchronize(e->stream));
        std::vector<float> b2((size_t)G.v_pad, 0.f);
        if (q->b2)
            for (int i = 0; i < n_rows; ++i) b2[i] = beta * q->b2[i];
        LMRL_CUDA(cudaMemcpyAsync(Q.b2s[k], b2.data(), b2.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        LMRL_CUDA(cudaStreamSynchronize(e->stream));
    }
    cudaFree(stage);
    LMRL_CUDA(cudaGetLastError());
    Q.n_q = n_q; Q.hidden = hid; Q.beta = beta; Q.linear = linear ? 1 : 0;
    G.policy_kind = 1;
    return 0;
}

int lmrl_set_policy_kind(lmrl_engine* e, int32_t kind) {
    LMRL_CHECK(e && e->gpt2 && e->gpt2->tw[0].loaded, "lmrl_set_policy_kind: no GPT-2 parameters loaded");
    LMRL_CHECK(kind == 0 || ((kind == 1 || kind == 2) && e->gpt2->tw[1].loaded && e->gpt2->qh.n_q > 0),
               "lmrl_set_policy_kind: kinds 1 / 2 (value-perturbed with / without pi_beta) need the value tower and Q heads; kind must be 0, 1 or 2");
    e->gpt2->policy_kind = kind;
    return 0;
}

// after a stream synchronisation: turn the allocator's overflow flag into an error
static int kv_check_overflow(lmrl_engine* e) {
    Gpt2Dev& G = *e->gpt2;
    int h[2] = {0, 0};
    LMRL_CUDA(cudaMemcpy(&h[0], G.kvp.overflow, sizeof(int), cudaMemcpyDeviceToHost));
    if (!h[0]) return 0;
    LMRL_CUDA(cudaMemcpy(&h[1], G.kvp.n_alloc, sizeof(int), cudaMemcpyDeviceToHost));
    return lmrl::fail("paged KV cache exhausted: the batch asked for %d pages of %d positions, the pool has %d (lmrl_config.kv_pages); results of this call are invalid",
                      h[1], KV_PAGE, G.kvp.n_pages);
}

static int load_prompts(lmrl_engine* e, const int32_t* tokens, const int32_t* n_tok, int32_t pitch, int32_t n) {
    Gpt2Dev& G = *e->gpt2;
    LMRL_CHECK(n > 0 && n <= e->cfg.max_envs, "batch of %d rows exceeds engine capacity %d", n, e->cfg.max_envs);
    const int max_in = e->cfg.max_input_length;
    // LEFT truncation on the host side of the copy: keep the last max_input tokens of every row
    std::vector<int32_t> lens(n), packed((size_t)n * max_in, 0);
    for (int i = 0; i < n; ++i) {
        const int len = std::min(n_tok[i], max_in), skip = n_tok[i] - len;
        LMRL_CHECK(n_tok[i] <= pitch && n_tok[i] >= 0, "row %d: n_tok %d exceeds pitch %d", i, n_tok[i], pitch);
        lens[i] = len;
        memcpy(packed.data() + (size_t)i * max_in, tokens + (size_t)i * pitch + skip, sizeof(int32_t) * len);
    }
    LMRL_CUDA(cudaMemcpy2DAsync(G.stream.tokens, sizeof(int32_t) * G.stream.cap, packed.data(), sizeof(int32_t) * max_in,
                                sizeof(int32_t) * max_in, n, cudaMemcpyHostToDevice, e->stream));
    LMRL_CUDA(cudaMemcpyAsync(G.stream.len, lens.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream));
    LMRL_CUDA(cudaMemsetAsync(G.stream.kv_len, 0, sizeof(int32_t) * n, e->stream));
    kv_reset_kernel<<<(e->cfg.max_envs * G.kvp.ppe + 255) / 256, 256, 0, e->stream>>>(G.kvp, e->cfg.max_envs);      // a new batch: every KV page is free again
    LMRL_CUDA(cudaStreamSynchronize(e->stream));   // host staging vectors die at return
    return 0;
}

int lmrl_generate_host(lmrl_engine* e, const int32_t* tokens, const int32_t* n_tok, int32_t pitch, int32_t n,
                       const uint64_t* episode_keys, int32_t turn, const lmrl_sampling* s, int32_t* out_tokens, int32_t* out_len) {
    LMRL_CHECK(e && e->has_gpt2, "lmrl_generate: no GPT-2 parameters loaded");
    LMRL_CHECK(tokens && n_tok && s && out_tokens && out_len, "lmrl_generate: null argument");
    LMRL_CUDA(cudaSetDevice(e->device));
    Gpt2Dev& G = *e->gpt2;
    if (load_prompts(e, tokens, n_tok, pitch, n)) return -1;
    if (episode_keys) LMRL_CUDA(cudaMemcpyAsync(G.episode_key, episode_keys, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, e->stream));
    else {
        std::vector<uint64_t> idx(n);
        for (int i = 0; i < n; ++i) idx[i] = (uint64_t)i;
        LMRL_CUDA(cudaMemcpyAsync(G.episode_key, idx.data(), sizeof(uint64_t) * n, cudaMemcpyHostToDevice, e->stream));
        LMRL_CUDA(cudaStreamSynchronize(e->stream));
    }
    LMRL_LAUNCH(2, LMRL_K(clear_gen_kernel, (n + 255) / 256, 256, 0, e->stream, n, G.gen));
    {
        std::vector<int32_t> turns(n, turn);
        LMRL_CUDA(cudaMemcpyAsync(G.gen_turn, turns.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream));
        LMRL_CUDA(cudaStreamSynchronize(e->stream));
    }
    if (forward_and_sample(e, 0, n, nullptr, G.gen_turn, 0, *s, nullptr)) return -1;                 // prefill pass + first token
    const int n_decode = e->cfg.max_new_tokens - 1;
    const bool graphable = G.use_graphs && !e->profile && n_decode >= 2;
    const Gpt2Dev::RolloutKey key{-1, n, s->do_sample, G.policy_kind, G.gemm_impl, G.gemv_max_rows, G.fuse_ln, s->temperature, s->seed, s->top_k, s->top_p};
    if (graphable && G.decode_graph && G.decode_key == key) {
        for (int k = 0; k < n_decode; ++k) LMRL_CUDA(cudaGraphLaunch(G.decode_graph, e->stream));
        e->launches += G.decode_graph_launches * n_decode;
    } else if (graphable && G.decode_warm_valid && G.decode_warm == key) {
        // second use of this configuration: capture one decode pass, then replay it for every remaining token
        if (G.decode_graph) { cudaGraphExecDestroy(G.decode_graph); G.decode_graph = nullptr; }
        const int64_t before = e->launches;
        LMRL_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = forward_and_sample(e, 1, n, nullptr, G.gen_turn, 0, *s, nullptr);
        cudaGraph_t graph = nullptr;
        const cudaError_t err = cudaStreamEndCapture(e->stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return -1; }
        LMRL_CHECK(err == cudaSuccess && graph, "lmrl_generate: CUDA graph capture of the decode pass failed: %s", cudaGetErrorString(err));
        G.decode_graph_launches = e->launches - before;
        e->launches = before;
        const cudaError_t ierr = cudaGraphInstantiate(&G.decode_graph, graph, 0);
        cudaGraphDestroy(graph);
        LMRL_CHECK(ierr == cudaSuccess, "lmrl_generate: cudaGraphInstantiate failed: %s", cudaGetErrorString(ierr));
        G.decode_key = key;
        for (int k = 0; k < n_decode; ++k) LMRL_CUDA(cudaGraphLaunch(G.decode_graph, e->stream));
        e->launches += G.decode_graph_launches * n_decode;
    } else {
        if (graphable) { G.decode_warm = key; G.decode_warm_valid = true; }
        for (int k = 0; k < n_decode; ++k)
            if (forward_and_sample(e, 1, n, nullptr, G.gen_turn, 0, *s, nullptr)) return -1;
    }
    LMRL_CUDA(cudaMemcpyAsync(out_tokens, G.gen.tok, sizeof(int32_t) * (size_t)n * e->cfg.max_new_tokens, cudaMemcpyDeviceToHost, e->stream));
    LMRL_CUDA(cudaMemcpyAsync(out_len, G.gen.len, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, e->stream));
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    return kv_check_overflow(e);
}

int lmrl_logits_host(lmrl_engine* e, const int32_t* tokens, const int32_t* n_tok, int32_t pitch, int32_t n, float* logits) {
    LMRL_CHECK(e && e->has_gpt2, "lmrl_logits: no GPT-2 parameters loaded");
    LMRL_CUDA(cudaSetDevice(e->device));
    Gpt2Dev& G = *e->gpt2;
    if (load_prompts(e, tokens, n_tok, pitch, n)) return -1;
    const size_t need = (size_t)e->cfg.max_envs * G.dims.n_vocab;
    if (G.logits_rows < need) {
        LMRL_CUDA(cudaMalloc(&G.logits, need * sizeof(float)));
        G.owned.push_back(G.logits);
        G.logits_rows = need;
    }
    LMRL_K(clear_gen_kernel, (n + 255) / 256, 256, 0, e->stream, n, G.gen);
    lmrl_sampling s{1.0f, 0, 0, 0, 0.f};
    if (forward_and_sample(e, 0, n, nullptr, nullptr, 0, s, G.logits)) return -1;
    // rows come back in slot order == env order for envs with tokens; scatter on the host
    std::vector<int32_t> slot_env(n);
    int n_slots = 0;
    LMRL_CUDA(cudaMemcpyAsync(&n_slots, G.rows.n_slots, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    LMRL_CUDA(cudaMemcpyAsync(slot_env.data(), G.rows.slot_env, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, e->stream));
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    for (int j = 0; j < n_slots; ++j)
        LMRL_CUDA(cudaMemcpyAsync(logits + (size_t)slot_env[j] * G.dims.n_vocab, G.logits + (size_t)j * G.dims.n_vocab,
                                  sizeof(float) * G.dims.n_vocab, cudaMemcpyDeviceToHost, e->stream));
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    return kv_check_overflow(e);
}

}  // extern "C" (reopened below)

static int score_alloc(lmrl_engine* e) {
    Gpt2Dev& G = *e->gpt2;
    if (G.s_cap) return 0;
    const int B = e->cfg.max_envs;
    const int cap = std::max(4096, (8 * B + 127) / 128 * 128);
    const size_t n_part = (size_t)2 * (G.v_pad / TC_BN);
    if (e->alloc(&G.xs, (size_t)cap * G.xl_cols) || e->alloc(&G.xs1, (size_t)cap * G.dims.d_model) || e->alloc(&G.lp_partial, (size_t)cap * n_part) ||
        e->alloc(&G.lp_tlogit, cap) || e->alloc(&G.lp_out, cap) || e->alloc(&G.score_rows, cap) || e->alloc(&G.score_targets, cap) ||
        e->alloc(&G.score_n, 1))
        return -1;
    G.s_cap = cap;
    return 0;
}

extern "C" {

int lmrl_score_host(lmrl_engine* e, const int32_t* tokens, const int32_t* n_tok, int32_t pitch, int32_t n, const int32_t* prefix_len,
                    float temperature, float* logprob_sum, float* token_logprobs) {
    LMRL_CHECK(e && e->has_gpt2, "lmrl_score: no GPT-2 parameters loaded");
    LMRL_CHECK(tokens && n_tok && prefix_len && logprob_sum, "lmrl_score: null argument");
    LMRL_CUDA(cudaSetDevice(e->device));
    Gpt2Dev& G = *e->gpt2;
    if (score_alloc(e)) return -1;
    const int max_in = e->cfg.max_input_length;
    // rows are assigned in env order, n_tok[i] rows for sequence i (positions 0..n_tok-1): scored rows are the positions
    // prefix_len-1 .. n_tok-2, each against the next token (ppo/score_fn.py:112-118)
    std::vector<int32_t> rows, targets, owner;
    int row0 = 0;
    for (int i = 0; i < n; ++i) {
        LMRL_CHECK(n_tok[i] >= 0 && n_tok[i] <= max_in && n_tok[i] <= pitch, "lmrl_score: sequence %d has %d tokens (max_input_length %d)", i, n_tok[i], max_in);
        LMRL_CHECK(prefix_len[i] >= 1 && prefix_len[i] <= n_tok[i], "lmrl_score: sequence %d: prefix_len %d must be in [1, %d]", i, prefix_len[i], n_tok[i]);
        for (int t = prefix_len[i] - 1; t + 1 < n_tok[i]; ++t) {
            rows.push_back(row0 + t);
            targets.push_back(tokens[(size_t)i * pitch + t + 1]);
            owner.push_back(i);
        }
        row0 += n_tok[i];
    }
    const int n_score = (int)rows.size();
    LMRL_CHECK(n_score <= G.s_cap, "lmrl_score: %d scored tokens exceed the capacity %d of one call", n_score, G.s_cap);
    for (int i = 0; i < n; ++i) logprob_sum[i] = 0.f;
    if (n_score == 0) return 0;
    if (load_prompts(e, tokens, n_tok, pitch, n)) return -1;
    LMRL_CUDA(cudaMemcpyAsync(G.score_rows, rows.data(), sizeof(int32_t) * n_score, cudaMemcpyHostToDevice, e->stream));
    LMRL_CUDA(cudaMemcpyAsync(G.score_targets, targets.data(), sizeof(int32_t) * n_score, cudaMemcpyHostToDevice, e->stream));
    LMRL_CUDA(cudaMemcpyAsync(G.score_n, &n_score, sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    LMRL_K(clear_gen_kernel, (n + 255) / 256, 256, 0, e->stream, n, G.gen);
    lmrl_sampling s{temperature, 0, 0, 0, 0.f};
    ScoreReq req{G.score_rows, G.score_targets, G.score_n, G.lp_out, n_score};
    if (forward_and_sample(e, 0, n, nullptr, nullptr, 0, s, nullptr, &req)) return -1;
    std::vector<float> lp(n_score);
    LMRL_CUDA(cudaMemcpyAsync(lp.data(), G.lp_out, sizeof(float) * n_score, cudaMemcpyDeviceToHost, e->stream));
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    if (kv_check_overflow(e)) return -1;
    for (int j = 0; j < n_score; ++j) logprob_sum[owner[j]] += lp[j];          // fp32 sum in token order, as .sum(axis=0) of the reference
    if (token_logprobs) memcpy(token_logprobs, lp.data(), sizeof(float) * n_score);
    return 0;
}

}  // extern "C" (reopened below)

static int forward_alloc(lmrl_engine* e) {
    Gpt2Dev& G = *e->gpt2;
    if (G.f_cap) return 0;
    const int B = e->cfg.max_envs, d = G.dims.d_model;
    const int cap = (int)std::min<size_t>((size_t)G.m_cap, (size_t)(std::max(8192, 16 * B) + 127) / 128 * 128);
    const size_t n_part = (size_t)2 * (G.v_pad / TC_BN);
    if (e->alloc(&G.fx, (size_t)cap * d) || e->alloc(&G.fr, (size_t)cap * 3 * d) || e->alloc(&G.f_targets, cap) ||
        e->alloc(&G.f_lp_partial, (size_t)cap * n_part) || e->alloc(&G.f_tlogit, cap) || e->alloc(&G.f_lp, cap) || e->alloc(&G.f_q1, cap) ||
        e->alloc(&G.f_q2, cap) || e->alloc(&G.f_v, cap))
        return -1;
    G.f_cap = cap;
    return 0;
}

extern "C" {

int lmrl_load_value_heads(lmrl_engine* e, int32_t head_set, const lmrl_value_heads* h) {
    LMRL_CHECK(e && e->gpt2 && e->gpt2->tw[0].loaded, "lmrl_load_value_heads: load GPT-2 parameters first");
    LMRL_CHECK(h && (head_set == 0 || head_set == 1), "lmrl_load_value_heads: head_set must be 0 (online) or 1 (target)");
    LMRL_CHECK(!(h->v && h->v_linear_w), "lmrl_load_value_heads: give the V head either as an MLPHead or as a LinearHead, not both");
    LMRL_CUDA(cudaSetDevice(e->device));
    Gpt2Dev& G = *e->gpt2;
    ValueHeadsDev& H = G.vh[head_set];
    const int d = G.dims.d_model;
    const lmrl_mlp_head* mlp[3] = {h->q1, h->q2, h->v};
    for (int k = 0; k < 3; ++k) {
        if (!mlp[k]) continue;
        const bool lin_q = k < 2 && mlp[k]->w1 == nullptr;          // LinearHead Q head (MC returns): kernel in w2, optional bias in b2
        LMRL_CHECK(mlp[k]->w2 && (lin_q || (mlp[k]->w1 && mlp[k]->b1 && mlp[k]->b2)), "lmrl_load_value_heads: head %d has a null array", k);
        LMRL_CHECK(mlp[k]->hidden == d, "lmrl_load_value_heads: hidden_dim %d must equal d_model %d", mlp[k]->hidden, d);
        LMRL_CHECK(k == 2 ? mlp[k]->n_out == 1 : mlp[k]->n_out >= G.dims.n_vocab, "lmrl_load_value_heads: head %d has output_dim %d", k, mlp[k]->n_out);
    }
    LMRL_CHECK((3 * d) % TC_BN == 0 && d % TC_BN == 0, "lmrl_load_value_heads: d_model %d must be a multiple of %d", d, TC_BN);
    if (!H.allocated) {
        if (e->alloc(&H.w1cat, (size_t)3 * d * d) || e->alloc(&H.b1cat, (size_t)3 * d) || e->alloc(&H.w2v, d) || e->alloc(&H.lin_w, d)) return -1;
        for (int k = 0; k < 2; ++k)
            if (e->alloc(&H.w2q[k], (size_t)G.v_pad * d) || e->alloc(&H.b2q[k], (size_t)G.v_pad)) return -1;
        H.allocated = true;
    }
    int n_out_max = 1;
    for (int k = 0; k < 2; ++k) if (mlp[k]) n_out_max = std::max(n_out_max, mlp[k]->n_out);
    // staging buffer of the fp32 kernels — only when there is an MLP / Q head to transpose: a PPO LinearHead hand-over (every optimizer
    // step) must not pay a cudaMalloc + cudaFree, which synchronise the device and, with peer access enabled by NCCL, cost 1-13 ms
    float* stage = nullptr;
    if (mlp[0] || mlp[1] || mlp[2]) LMRL_CUDA(cudaMalloc(&stage, (size_t)d * std::max(n_out_max, d) * sizeof(float)));
    H.n_mlp = 0; H.hidden = d; H.slot_q1 = H.slot_q2 = H.slot_v = -1; H.has_lin = false; H.q_lin[0] = H.q_lin[1] = false;
    for (int k = 0; k < 3; ++k) {
        const lmrl_mlp_head* q = mlp[k];
        if (!q) continue;
        if (k < 2 && q->w1 == nullptr) {
            H.q_lin[k] = true;                                      // no first layer, no slot in the ReLU buffer
        } else {
            const int slot = H.n_mlp++;
            (k == 0 ? H.slot_q1 : k == 1 ? H.slot_q2 : H.slot_v) = slot;
            LMRL_CUDA(cudaMemcpyAsync(stage, q->w1, (size_t)d * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
            transpose_to_bf16_kernel<<<dim3((d + 31) / 32, (d + 31) / 32), dim3(32, 8), 0, e->stream>>>(stage, d, d, H.w1cat + (size_t)slot * d * d);
            LMRL_CUDA(cudaStreamSynchronize(e->stream));
            LMRL_CUDA(cudaMemcpyAsync(H.b1cat + (size_t)slot * d, q->b1, (size_t)d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        }
        if (k < 2) {
            const int n_rows = std::min(q->n_out, G.v_pad);
            LMRL_CUDA(cudaMemcpyAsync(stage, q->w2, (size_t)d * q->n_out * sizeof(float), cudaMemcpyHostToDevice, e->stream));
            transpose_to_bf16_kernel<<<dim3((n_rows + 31) / 32, (d + 31) / 32), dim3(32, 8), 0, e->stream>>>(stage, d, q->n_out, H.w2q[k], n_rows);
            LMRL_CUDA(cudaStreamSynchronize(e->stream));
            LMRL_CUDA(cudaMemsetAsync(H.b2q[k], 0, (size_t)G.v_pad * sizeof(float), e->stream));
            if (q->b2) LMRL_CUDA(cudaMemcpyAsync(H.b2q[k], q->b2, (size_t)n_rows * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        } else {
            LMRL_CUDA(cudaMemcpyAsync(H.w2v, q->w2, (size_t)d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
            H.b2v = q->b2[0];
        }
        LMRL_CUDA(cudaStreamSynchronize(e->stream));
    }
    // unused ReLU columns must not hold stale weights of an earlier, larger head set: the GEMM only covers n_mlp * d columns
    if (h->v_linear_w) {
        LMRL_CUDA(cudaMemcpyAsync(H.lin_w, h->v_linear_w, (size_t)d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        H.lin_b = h->v_linear_b ? h->v_linear_b[0] : 0.f;
        H.has_lin = true;
        LMRL_CUDA(cudaStreamSynchronize(e->stream));
    }
    if (stage) cudaFree(stage);
    LMRL_CUDA(cudaGetLastError());
    return 0;
}

int lmrl_forward_host(lmrl_engine* e, int32_t tower, int32_t head_set, const int32_t* tokens, const int32_t* n_tok, int32_t pitch, int32_t n,
                      float temperature, lmrl_forward_out* out) {
    LMRL_CHECK(e && e->has_gpt2, "lmrl_forward: no GPT-2 parameters loaded");
    LMRL_CHECK(tokens && n_tok && out && n >= 0 && pitch > 0, "lmrl_forward: null argument");
    LMRL_CHECK(tower >= 0 && tower <= 2 && e->gpt2->tw[tower].loaded, "lmrl_forward: tower %d is not loaded", tower);
    LMRL_CHECK(head_set == 0 || head_set == 1, "lmrl_forward: head_set must be 0 or 1");
    LMRL_CUDA(cudaSetDevice(e->device));
    Gpt2Dev& G = *e->gpt2;
    const lmrl_gpt2_dims& D = G.dims;
    const ValueHeadsDev& H = G.vh[head_set];
    const bool want_lp = out->token_logprobs != nullptr, want_v = out->v != nullptr, want_q1 = out->q1 != nullptr, want_q2 = out->q2 != nullptr;
    LMRL_CHECK(!want_q1 || H.slot_q1 >= 0 || H.q_lin[0], "lmrl_forward: q1 requested but head set %d has no Q1 head", head_set);
    LMRL_CHECK(!want_q2 || H.slot_q2 >= 0 || H.q_lin[1], "lmrl_forward: q2 requested but head set %d has no Q2 head", head_set);
    LMRL_CHECK(!want_v || H.slot_v >= 0 || H.has_lin, "lmrl_forward: v requested but head set %d has no V head", head_set);
    if (forward_alloc(e)) return -1;
    const int max_in = e->cfg.max_input_length;
    for (int i = 0; i < n; ++i) {
        LMRL_CHECK(n_tok[i] >= 0 && n_tok[i] <= max_in && n_tok[i] <= pitch, "lmrl_forward: sequence %d has %d tokens (max_input_length %d, pitch %d)", i, n_tok[i], max_in, pitch);
        for (int t = 0; t < n_tok[i]; ++t)
            LMRL_CHECK(tokens[(size_t)i * pitch + t] >= 0 && tokens[(size_t)i * pitch + t] < G.v_pad, "lmrl_forward: sequence %d position %d: token id %d out of range", i, t, tokens[(size_t)i * pitch + t]);
    }
    float* outs[4] = {out->token_logprobs, out->v, out->q1, out->q2};
    for (float* o : outs) if (o) memset(o, 0, sizeof(float) * (size_t)n * pitch);
    const TowerDev& T = G.tw[tower];
    cudaStream_t st = e->stream;
    const int ew_blocks = G.sm_count * 8;
    const bool need_mlp = ((want_q1 && H.slot_q1 >= 0) || (want_q2 && H.slot_q2 >= 0) || (want_v && H.slot_v >= 0)) && H.n_mlp > 0;
    const bool need_lin_q = (want_q1 && H.q_lin[0]) || (want_q2 && H.q_lin[1]);
    std::vector<int32_t> targets;
    std::vector<float> host[4];
    for (int i0 = 0; i0 < n;) {
        int i1 = i0, rows = 0;
        while (i1 < n && i1 - i0 < e->cfg.max_envs && rows + n_tok[i1] <= G.f_cap) { rows += n_tok[i1]; ++i1; }
        LMRL_CHECK(i1 > i0, "lmrl_forward: sequence %d does not fit the row capacity %d", i0, G.f_cap);
        const int nc = i1 - i0;
        if (rows > 0) {
            targets.assign(rows, 0);
            for (int i = i0, r = 0; i < i1; ++i)
                for (int t = 0; t < n_tok[i]; ++t, ++r) targets[r] = t + 1 < n_tok[i] ? std::min(tokens[(size_t)i * pitch + t + 1], G.v_pad - 1) : 0;
            if (load_prompts(e, tokens + (size_t)i0 * pitch, n_tok + i0, pitch, nc)) return -1;
            LMRL_CUDA(cudaMemcpyAsync(G.f_targets, targets.data(), sizeof(int32_t) * rows, cudaMemcpyHostToDevice, st));
            e->cur_pass_slot = -1;
            e->cur_mode = 0;
            LMRL_LAUNCH(2, LMRL_K(clear_gen_kernel, (nc + 255) / 256, 256, 0, st, nc, G.gen));
            LMRL_LAUNCH(2, LMRL_K(build_rows_kernel, 1, 1024, 0, st, 0, 0, nc, (const uint8_t*)nullptr, G.stream.tokens, G.stream.len, G.stream.kv_len, G.stream.cap, G.gen, G.rows, G.kvp, -1));
            note_pass_counts(e);
            LMRL_LAUNCH(2, LMRL_K(fill_prefill_rows_kernel, (nc * 32 + 255) / 256, 256, 0, st, G.stream.tokens, G.stream.len, G.stream.kv_len, G.stream.cap, G.gen, G.rows, G.kvp, -1));
            LMRL_LAUNCH(2, LMRL_K(embed_kernel, ew_blocks, 256, 0, st, T.wte, T.wpe, D.d_model, G.rows, G.x));
            if (tower_layers(e, T, 0, nc)) return -1;
            if (want_lp || need_mlp || need_lin_q)
                LMRL_LAUNCH(2, LMRL_K(layernorm_kernel, ew_blocks, 256, 0, st, G.x, T.lnf_g, T.lnf_b, D.d_model, D.ln_eps, G.rows.M, (const int*)nullptr, G.fx, D.d_model));
            if (want_lp) {
                GemmArgs h{};
                h.A = G.fx; h.W = T.wte; h.bias = nullptr; h.N = G.v_pad; h.K = D.d_model; h.M_dev = G.rows.M;
                h.inv_temp = 1.0f / (temperature > 0.f ? temperature : 1.0f); h.n_vocab = D.n_vocab;
                h.lp_partial = G.f_lp_partial; h.lp_target = G.f_targets; h.lp_tlogit = G.f_tlogit;
                if (launch_gemm<EPI_LOGPROB>(e, h, D.n_vocab, 0, (size_t)G.f_cap, G.v_pad, 4)) return -1;
                LMRL_LAUNCH(2, LMRL_K(logprob_finalize_kernel, std::max(1, std::min((rows + 7) / 8, G.sm_count * 8)), 256, 0, st, G.f_lp_partial, 2 * (G.v_pad / TC_BN),
                                      G.f_tlogit, G.rows.M, G.f_lp));
            }
            if (need_mlp) {
                GemmArgs r{};
                r.A = G.fx; r.W = H.w1cat; r.bias = H.b1cat; r.N = H.n_mlp * H.hidden; r.K = D.d_model; r.M_dev = G.rows.M; r.out = G.fr;
                r.out_col0 = 0; r.out_scale = 1.0f;
                GemmExtra rx;
                rx.out_cols = (size_t)3 * D.d_model;
                if (launch_gemm<EPI_RELU>(e, r, r.N, 0, (size_t)G.f_cap, (size_t)3 * D.d_model, 4, rx)) return -1;
            }
            if (want_v || want_q1 || want_q2) {
                ValueHeadArgs a{};
                a.fr = G.fr; a.fr_pitch = 3 * D.d_model; a.hidden = H.hidden;
                a.slot_q1 = want_q1 ? H.slot_q1 : -1; a.slot_q2 = want_q2 ? H.slot_q2 : -1; a.slot_v = want_v ? H.slot_v : -1;
                a.fx = G.fx; a.q1_lin = (want_q1 && H.q_lin[0]) ? 1 : 0; a.q2_lin = (want_q2 && H.q_lin[1]) ? 1 : 0;
                a.w2q1 = H.w2q[0]; a.b2q1 = H.b2q[0]; a.w2q2 = H.w2q[1]; a.b2q2 = H.b2q[1]; a.w2v = H.w2v; a.b2v = H.b2v;
                a.targets = G.f_targets;
                a.x = G.x; a.lnf_g = T.lnf_g; a.lnf_b = T.lnf_b; a.lin_w = H.lin_w; a.lin_b = H.lin_b; a.d = D.d_model; a.eps = D.ln_eps;
                a.has_lin = (want_v && H.slot_v < 0 && H.has_lin) ? 1 : 0;
                a.q1_out = G.f_q1; a.q2_out = G.f_q2; a.v_out = G.f_v;
                LMRL_LAUNCH(2, LMRL_K(value_heads_kernel, std::max(1, std::min((rows + 7) / 8, G.sm_count * 8)), 256, 0, st, G.rows.M, a));
            }
            const float* dev[4] = {G.f_lp, G.f_v, G.f_q1, G.f_q2};
            for (int k = 0; k < 4; ++k)
                if (outs[k]) { host[k].resize(rows); LMRL_CUDA(cudaMemcpyAsync(host[k].data(), dev[k], sizeof(float) * rows, cudaMemcpyDeviceToHost, st)); }
            LMRL_CUDA(cudaStreamSynchronize(st));
            LMRL_CUDA(cudaGetLastError());
            if (kv_check_overflow(e)) return -1;
            for (int i = i0, r = 0; i < i1; ++i)
                for (int t = 0; t < n_tok[i]; ++t, ++r) {
                    const bool has_next = t + 1 < n_tok[i];
                    const size_t o = (size_t)i * pitch + t;
                    if (outs[0] && has_next) outs[0][o] = host[0][r];
                    if (outs[1]) outs[1][o] = host[1][r];
                    if (outs[2] && has_next) outs[2][o] = host[2][r];
                    if (outs[3] && has_next) outs[3][o] = host[3][r];
                }
        }
        i0 = i1;
    }
    e->prof_close();
    return 0;
}

static int rollout_enqueue(lmrl_engine* e, int32_t env_kind, const uint64_t* seeds_dev, const int32_t* opt_pos_dev, int32_t n, const lmrl_sampling& s) {
    Gpt2Dev& G = *e->gpt2;
    LMRL_CHECK(env_kind == LMRL_ENV_WORDLE || env_kind == LMRL_ENV_MAZE, "fused rollout: unknown env_kind %d", env_kind);
    LMRL_CHECK(n > 0 && n <= e->cfg.max_envs, "fused rollout: %d envs exceed engine capacity %d", n, e->cfg.max_envs);
    const bool wordle = env_kind == LMRL_ENV_WORDLE;
    if (wordle) {
        LMRL_CHECK(e->has_wordle, "fused rollout: no Wordle vocabulary loaded");
        LMRL_CHECK(e->cfg.max_new_tokens * e->max_token_bytes <= MAX_ACTION_CP, "fused rollout: max_new_tokens * longest token exceeds %d bytes", MAX_ACTION_CP);
    } else {
        LMRL_CHECK(e->has_maze && e->maze.obs_tokens, "fused rollout: no maze (with its observation token table) loaded");
        LMRL_CHECK(e->maze.obs_pitch + MAZE_ACT_PITCH + e->cfg.max_input_length <= G.stream_cap, "fused rollout: observation does not fit the token stream");
    }
    const int T = e->cfg.max_turns;
    cudaStream_t st = e->stream;
    const int blocks = (n + ENV_WARPS_PER_BLOCK - 1) / ENV_WARPS_PER_BLOCK;
    const uint8_t* done = wordle ? e->wordle.done : e->maze.done;
    if (wordle) {
        LMRL_LAUNCH(3, LMRL_K(wordle_reset_kernel, blocks, ENV_WARPS_PER_BLOCK * 32, ENV_WARPS_PER_BLOCK * MT_N * sizeof(uint32_t), st, e->wordle, seeds_dev, n));
        LMRL_LAUNCH(3, LMRL_K(init_stream_kernel, (n + 255) / 256, 256, 0, st, n, e->d_header, (int)e->header_tokens.size(), G.stream, seeds_dev,
                                                                           G.episode_key, G.ep_return, G.ep_length, G.ep_done));
    } else {
        LMRL_LAUNCH(3, LMRL_K(maze_reset_kernel, blocks, ENV_WARPS_PER_BLOCK * 32, ENV_WARPS_PER_BLOCK * MT_N * sizeof(uint32_t), st, e->maze, seeds_dev, opt_pos_dev, nullptr, n));
        LMRL_LAUNCH(3, LMRL_K(init_stream_kernel, (n + 255) / 256, 256, 0, st, n, e->d_header, 0, G.stream, seeds_dev,
                                                                           G.episode_key, G.ep_return, G.ep_length, G.ep_done));
        LMRL_LAUNCH(3, LMRL_K(maze_init_stream_kernel, (n + 127) / 128, 128, 0, st, e->maze, G.stream, n));
    }
    LMRL_LAUNCH(2, LMRL_K(kv_reset_kernel, (e->cfg.max_envs * G.kvp.ppe + 255) / 256, 256, 0, st, G.kvp, e->cfg.max_envs));     // a new batch: every KV page is free again
    ActionSrc src{};
    src.tokens = G.gen.tok; src.ntokens = G.gen.len; src.tpitch = G.gen.max_new;
    // two half-batch lanes on two streams (see struct Lane); not under the per-launch profiler (its events live on one stream) and not
    // with top-k / top-p (one shared logits dump)
    const bool filtered = s.top_k > 0 || (s.top_p > 0.f && s.top_p < 1.f);
    const bool split = G.two_lanes && lanes_usable(e) && !e->profile && !filtered && n >= 256;
    const int n1 = std::min(e->cfg.max_envs / 2, (n + 1) / 2);
    G.lane[1].e0 = 0; G.lane[2].e0 = n1;
    struct RestoreLane { Gpt2Dev& g; ~RestoreLane() { g.cur = &g.lane[0]; g.fold_nlnl = -1; } } restore_lane{G};
    G.fold_nlnl = wordle ? G.nlnl_id : -1;
    for (int turn = 0; turn < T; ++turn) {
        LMRL_LAUNCH(2, LMRL_K(clear_gen_kernel, (n + 255) / 256, 256, 0, st, n, G.gen));
        if (!split) {
            for (int k = 0; k < e->cfg.max_new_tokens; ++k)
                if (forward_and_sample(e, k == 0 ? 0 : 1, n, done, G.ep_length, 0, s, nullptr)) return -1;
        } else {
            // fork: lane 2 (second half of the envs) runs on its own stream next to lane 1; join before the env step
            LMRL_CUDA(cudaEventRecord(G.ev_fork, st));
            LMRL_CUDA(cudaStreamWaitEvent(G.stream_b, G.ev_fork, 0));
            for (int k = 0; k < e->cfg.max_new_tokens; ++k) {
                G.cur = &G.lane[1];
                if (forward_and_sample(e, k == 0 ? 0 : 1, n1, done, G.ep_length, 0, s, nullptr)) return -1;
                G.cur = &G.lane[2];
                if (forward_and_sample(e, k == 0 ? 0 : 1, n - n1, done, G.ep_length, 0, s, nullptr)) return -1;
            }
            G.cur = &G.lane[0];
            LMRL_CUDA(cudaEventRecord(G.ev_join, G.stream_b));
            LMRL_CUDA(cudaStreamWaitEvent(st, G.ev_join, 0));
        }
        if (wordle)
            LMRL_LAUNCH(3, LMRL_K(wordle_step_kernel, blocks, ENV_WARPS_PER_BLOCK * 32, ENV_WARPS_PER_BLOCK * sizeof(WordleScratch), st, 
                               e->wordle, e->tok, src, nullptr, e->step, G.stream, n));
        else
            LMRL_LAUNCH(3, LMRL_K(maze_fused_step_kernel, (n + 127) / 128, 128, 0, st, e->maze, e->tok, src, e->step, G.stream, n));
        LMRL_LAUNCH(3, LMRL_K(log_turn_kernel, (n + 255) / 256, 256, 0, st, n, T, G.gen, e->step, G.ep_return, G.ep_length, G.ep_done,
                                                                        G.step_reward, G.step_valid, G.step_fb, G.step_aux, G.act_tokens, G.act_len));
        // long episodes (Maze: up to max_steps + 1 turns): every 8th turn ask the device whether anything is still running
        if (T > 8 && (turn & 7) == 7 && turn + 1 < T) {
            LMRL_CUDA(cudaMemsetAsync(G.live, 0, sizeof(int), st));
            LMRL_LAUNCH(2, LMRL_K(count_live_kernel, 32, 256, 0, st, done, n, G.live));
            int live = 0;
            LMRL_CUDA(cudaMemcpyAsync(&live, G.live, sizeof(int), cudaMemcpyDeviceToHost, st));
            LMRL_CUDA(cudaStreamSynchronize(st));
            if (live == 0) break;
        }
    }
    e->prof_close();
    LMRL_CUDA(cudaGetLastError());
    return 0;
}

static int rollout_export(lmrl_engine* e, int32_t n, lmrl_rollout_out* out, cudaMemcpyKind kind) {
    Gpt2Dev& G = *e->gpt2;
    const size_t T = e->cfg.max_turns, K = e->cfg.max_new_tokens;
    if (!out) return 0;
    if (out->ep_return) LMRL_CUDA(cudaMemcpyAsync(out->ep_return, G.ep_return, sizeof(float) * n, kind, e->stream));
    if (out->ep_length) LMRL_CUDA(cudaMemcpyAsync(out->ep_length, G.ep_length, sizeof(int32_t) * n, kind, e->stream));
    if (out->ep_done) LMRL_CUDA(cudaMemcpyAsync(out->ep_done, G.ep_done, n, kind, e->stream));
    if (out->step_reward) LMRL_CUDA(cudaMemcpyAsync(out->step_reward, G.step_reward, sizeof(float) * n * T, kind, e->stream));
    if (out->step_valid) LMRL_CUDA(cudaMemcpyAsync(out->step_valid, G.step_valid, n * T, kind, e->stream));
    if (out->step_feedback) LMRL_CUDA(cudaMemcpyAsync(out->step_feedback, G.step_fb, n * T * 5, kind, e->stream));
    if (out->step_aux) LMRL_CUDA(cudaMemcpyAsync(out->step_aux, G.step_aux, sizeof(int32_t) * n * T, kind, e->stream));
    if (out->act_tokens) LMRL_CUDA(cudaMemcpyAsync(out->act_tokens, G.act_tokens, sizeof(int32_t) * n * T * K, kind, e->stream));
    if (out->act_len) LMRL_CUDA(cudaMemcpyAsync(out->act_len, G.act_len, sizeof(int32_t) * n * T, kind, e->stream));
    return 0;
}

}  // extern "C" (reopened below)

// seeds / options already sit in e->d_seeds / e->d_opt_pos: run the rollout eagerly, or replay its CUDA graph
static int rollout_run(lmrl_engine* e, int32_t env_kind, bool has_opt, int32_t n, const lmrl_sampling& s) {
    Gpt2Dev& G = *e->gpt2;
    const int32_t* opt = has_opt ? e->d_opt_pos : nullptr;
    const bool graphable = G.use_graphs && !e->profile && e->cfg.max_turns <= 8;     // longer rollouts poll the device for early exit
    if (!graphable) return rollout_enqueue(e, env_kind, e->d_seeds, opt, n, s);
    const Gpt2Dev::RolloutKey key{env_kind, n, s.do_sample, G.policy_kind, G.gemm_impl, has_opt ? 1 : 0, G.two_lanes, s.temperature, s.seed, s.top_k, s.top_p};
    if (G.graph_exec && G.graph_key == key) {
        LMRL_CUDA(cudaGraphLaunch(G.graph_exec, e->stream));
        e->launches += G.graph_launches;
        return 0;
    }
    if (!(G.warm_valid && G.warm_key == key)) {            // first use of this configuration: eager
        G.warm_key = key; G.warm_valid = true;
        return rollout_enqueue(e, env_kind, e->d_seeds, opt, n, s);
    }
    if (G.graph_exec) { cudaGraphExecDestroy(G.graph_exec); G.graph_exec = nullptr; }
    const int64_t before = e->launches;
    LMRL_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = rollout_enqueue(e, env_kind, e->d_seeds, opt, n, s);
    cudaGraph_t graph = nullptr;
    const cudaError_t err = cudaStreamEndCapture(e->stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return -1; }
    LMRL_CHECK(err == cudaSuccess && graph, "fused rollout: CUDA graph capture failed: %s", cudaGetErrorString(err));
    G.graph_launches = e->launches - before;
    e->launches = before;
    const cudaError_t ierr = cudaGraphInstantiate(&G.graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    LMRL_CHECK(ierr == cudaSuccess, "fused rollout: cudaGraphInstantiate failed: %s", cudaGetErrorString(ierr));
    G.graph_key = key;
    LMRL_CUDA(cudaGraphLaunch(G.graph_exec, e->stream));
    e->launches += G.graph_launches;
    return 0;
}

extern "C" {

int lmrl_set_use_pdl(lmrl_engine* e, int32_t enable) {
    LMRL_CHECK(e && e->gpt2, "lmrl_set_use_pdl: no GPT-2 parameters loaded");
    e->gpt2->use_pdl = enable ? 1 : 0;
    if (e->gpt2->graph_exec) { cudaGraphExecDestroy(e->gpt2->graph_exec); e->gpt2->graph_exec = nullptr; }
    if (e->gpt2->decode_graph) { cudaGraphExecDestroy(e->gpt2->decode_graph); e->gpt2->decode_graph = nullptr; }
    e->gpt2->warm_valid = false; e->gpt2->decode_warm_valid = false;
    return 0;
}

int lmrl_kv_stats(lmrl_engine* e, int32_t* pages_in_use, int32_t* pages_total, int32_t* page_positions) {
    LMRL_CHECK(e && e->gpt2, "lmrl_kv_stats: no GPT-2 parameters loaded");
    LMRL_CUDA(cudaSetDevice(e->device));
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    int used = 0;
    LMRL_CUDA(cudaMemcpy(&used, e->gpt2->kvp.n_alloc, sizeof(int), cudaMemcpyDeviceToHost));
    if (pages_in_use) *pages_in_use = std::min(used, e->gpt2->kvp.n_pages);
    if (pages_total) *pages_total = e->gpt2->kvp.n_pages;
    if (page_positions) *page_positions = KV_PAGE;
    return kv_check_overflow(e);
}

static int* option_slot(lmrl::Gpt2Dev& G, const char* name) {
    if (!strcmp(name, "fuse_ln")) return &G.fuse_ln;
    if (!strcmp(name, "attn_wide")) return &G.attn_wide;
    if (!strcmp(name, "gemv_max_rows")) return &G.gemv_max_rows;
    if (!strcmp(name, "tok_nlnl_id")) return &G.nlnl_id;
    if (!strcmp(name, "att_grid_mult")) return &G.att_grid_mult;
    if (!strcmp(name, "ln_fold")) return &G.ln_fold;
    if (!strcmp(name, "att_decode_impl")) return &G.att_decode_impl;
    return nullptr;
}

int lmrl_set_option(lmrl_engine* e, const char* name, int32_t value) {
    LMRL_CHECK(e && e->gpt2 && name, "lmrl_set_option: no GPT-2 parameters loaded");
    int* slot = option_slot(*e->gpt2, name);
    LMRL_CHECK(slot, "lmrl_set_option: unknown option '%s'", name);
    *slot = value;
    if (e->gpt2->graph_exec) { cudaGraphExecDestroy(e->gpt2->graph_exec); e->gpt2->graph_exec = nullptr; }
    if (e->gpt2->decode_graph) { cudaGraphExecDestroy(e->gpt2->decode_graph); e->gpt2->decode_graph = nullptr; }
    e->gpt2->warm_valid = false; e->gpt2->decode_warm_valid = false;
    return 0;
}

int lmrl_get_option(lmrl_engine* e, const char* name, int32_t* value) {
    LMRL_CHECK(e && e->gpt2 && name && value, "lmrl_get_option: no GPT-2 parameters loaded");
    int* slot = option_slot(*e->gpt2, name);
    LMRL_CHECK(slot, "lmrl_get_option: unknown option '%s'", name);
    *value = *slot;
    return 0;
}

int lmrl_set_two_lanes(lmrl_engine* e, int32_t enable) {
    LMRL_CHECK(e && e->gpt2, "lmrl_set_two_lanes: no GPT-2 parameters loaded");
    e->gpt2->two_lanes = enable ? 1 : 0;
    return 0;
}

int lmrl_set_use_graphs(lmrl_engine* e, int32_t enable) {
    LMRL_CHECK(e && e->gpt2, "lmrl_set_use_graphs: no GPT-2 parameters loaded");
    e->gpt2->use_graphs = enable ? 1 : 0;
    return 0;
}

int lmrl_rollout_host(lmrl_engine* e, int32_t env_kind, const uint64_t* seeds, const int32_t* opt_pos, int32_t n,
                      const lmrl_sampling* s, lmrl_rollout_out* out) {
    LMRL_CHECK(e && e->has_gpt2, "lmrl_rollout: no GPT-2 parameters loaded");
    LMRL_CHECK(seeds && s, "lmrl_rollout: null argument");
    LMRL_CUDA(cudaSetDevice(e->device));
    LMRL_CHECK(n > 0 && n <= e->cfg.max_envs, "lmrl_rollout: %d envs exceed engine capacity %d", n, e->cfg.max_envs);
    LMRL_CUDA(cudaMemcpyAsync(e->d_seeds, seeds, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, e->stream));
    if (opt_pos) LMRL_CUDA(cudaMemcpyAsync(e->d_opt_pos, opt_pos, sizeof(int32_t) * 2 * n, cudaMemcpyHostToDevice, e->stream));
    if (rollout_run(e, env_kind, opt_pos != nullptr, n, *s)) return -1;
    if (rollout_export(e, n, out, cudaMemcpyDeviceToHost)) return -1;
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    return kv_check_overflow(e);
}

int lmrl_rollout_dev(lmrl_engine* e, int32_t env_kind, const uint64_t* seeds_dev, const int32_t* opt_pos_dev, int32_t n,
                     const lmrl_sampling* s, lmrl_rollout_out* out_dev) {
    LMRL_CHECK(e && e->has_gpt2, "lmrl_rollout: no GPT-2 parameters loaded");
    LMRL_CHECK(seeds_dev && s, "lmrl_rollout: null argument");
    LMRL_CUDA(cudaSetDevice(e->device));
    LMRL_CHECK(n > 0 && n <= e->cfg.max_envs, "lmrl_rollout: %d envs exceed engine capacity %d", n, e->cfg.max_envs);
    if (seeds_dev != e->d_seeds) LMRL_CUDA(cudaMemcpyAsync(e->d_seeds, seeds_dev, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, e->stream));
    if (opt_pos_dev && opt_pos_dev != e->d_opt_pos)
        LMRL_CUDA(cudaMemcpyAsync(e->d_opt_pos, opt_pos_dev, sizeof(int32_t) * 2 * n, cudaMemcpyDeviceToDevice, e->stream));
    if (rollout_run(e, env_kind, opt_pos_dev != nullptr, n, *s)) return -1;
    return rollout_export(e, n, out_dev, cudaMemcpyDeviceToDevice);
}

int lmrl_rollout_chains_host(lmrl_engine* e, int32_t n, int32_t pitch, int32_t* tokens, int32_t* n_tok, uint8_t* is_action, float* reward, uint8_t* done,
                             uint8_t* truncated) {
    LMRL_CHECK(e && e->has_gpt2 && e->has_wordle, "lmrl_rollout_chains: needs a Wordle engine with a policy (run a fused Wordle rollout first)");
    LMRL_CHECK(tokens && n_tok && is_action && reward && done && truncated, "lmrl_rollout_chains: null argument");
    LMRL_CHECK(n > 0 && n <= e->cfg.max_envs && pitch > 0, "lmrl_rollout_chains: bad batch %d / pitch %d", n, pitch);
    LMRL_CUDA(cudaSetDevice(e->device));
    Gpt2Dev& G = *e->gpt2;
    cudaStream_t st = e->stream;
    const int w = std::min(pitch, G.stream.cap);
    LMRL_CUDA(cudaMemcpy2DAsync(tokens, sizeof(int32_t) * pitch, G.stream.tokens, sizeof(int32_t) * G.stream.cap, sizeof(int32_t) * w, n, cudaMemcpyDeviceToHost, st));
    LMRL_CUDA(cudaMemcpy2DAsync(is_action, pitch, G.stream.is_action, G.stream.cap, w, n, cudaMemcpyDeviceToHost, st));
    LMRL_CUDA(cudaMemcpy2DAsync(reward, sizeof(float) * pitch, G.stream.tok_reward, sizeof(float) * G.stream.cap, sizeof(float) * w, n, cudaMemcpyDeviceToHost, st));
    LMRL_CUDA(cudaMemcpyAsync(n_tok, G.stream.len, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    LMRL_CUDA(cudaMemcpyAsync(done, G.ep_done, n, cudaMemcpyDeviceToHost, st));
    LMRL_CUDA(cudaMemcpyAsync(truncated, G.stream.truncated, n, cudaMemcpyDeviceToHost, st));
    LMRL_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < n; ++i) LMRL_CHECK(n_tok[i] <= pitch, "lmrl_rollout_chains: episode %d has %d tokens, pitch is %d", i, n_tok[i], pitch);
    return 0;
}

int lmrl_wordle_scripted_rollout_dev(lmrl_engine* e, const uint64_t* seeds_dev, int32_t n, float prob_smart, uint64_t policy_seed,
                                     lmrl_rollout_out* out) {
    LMRL_CHECK(e && e->has_wordle, "lmrl_wordle_scripted_rollout: no Wordle vocabulary loaded");
    LMRL_CHECK(seeds_dev && out && out->ep_return && out->ep_length && out->ep_done, "lmrl_wordle_scripted_rollout: seeds and the per-episode outputs are required");
    LMRL_CHECK(n > 0 && n <= e->cfg.max_envs, "lmrl_wordle_scripted_rollout: %d envs exceed engine capacity %d", n, e->cfg.max_envs);
    LMRL_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    const int T = e->cfg.max_turns;
    const int blocks = (n + ENV_WARPS_PER_BLOCK - 1) / ENV_WARPS_PER_BLOCK;
    LMRL_LAUNCH(3, LMRL_K(wordle_reset_kernel, blocks, ENV_WARPS_PER_BLOCK * 32, ENV_WARPS_PER_BLOCK * MT_N * sizeof(uint32_t), st, e->wordle, seeds_dev, n));
    LMRL_LAUNCH(3, LMRL_K(clear_episode_log_kernel, (n + 255) / 256, 256, 0, st, n, out->ep_return, out->ep_length, out->ep_done));
    ActionSrc src{};
    src.scripted = 1; src.prob_smart = prob_smart; src.policy_seed = policy_seed; src.episode_seed = seeds_dev;
    StreamDev none{};
    for (int turn = 0; turn < std::min(T, (int)WORDLE_TRIES); ++turn) {
        LMRL_LAUNCH(3, LMRL_K(wordle_step_kernel, blocks, ENV_WARPS_PER_BLOCK * 32, ENV_WARPS_PER_BLOCK * sizeof(WordleScratch), st, 
                           e->wordle, e->tok, src, nullptr, e->step, none, n));
        LMRL_LAUNCH(3, LMRL_K(log_scripted_turn_kernel, (n + 255) / 256, 256, 0, st, n, T, e->cfg.max_new_tokens, e->wordle, e->step, out->ep_return,
                                                                                 out->ep_length, out->ep_done, out->step_reward, out->step_valid,
                                                                                 out->step_feedback, out->step_aux, out->act_tokens, out->act_len));
    }
    LMRL_CUDA(cudaGetLastError());
    return 0;
}

int lmrl_set_gemm_impl(lmrl_engine* e, int32_t impl) {
    LMRL_CHECK(e && e->gpt2, "lmrl_set_gemm_impl: no GPT-2 parameters loaded");
    LMRL_CHECK(impl >= 0 && impl <= 2, "lmrl_set_gemm_impl: impl must be 0 (mma.sync), 1 (tcgen05, one CTA per tile) or 2 (tcgen05, CTA pairs on prefill passes)");
    e->gpt2->gemm_impl = impl;
    // a captured rollout / decode graph holds the old kernels
    if (e->gpt2->graph_exec) { cudaGraphExecDestroy(e->gpt2->graph_exec); e->gpt2->graph_exec = nullptr; }
    if (e->gpt2->decode_graph) { cudaGraphExecDestroy(e->gpt2->decode_graph); e->gpt2->decode_graph = nullptr; }
    e->gpt2->warm_valid = false; e->gpt2->decode_warm_valid = false;
    return 0;
}

int lmrl_test_gemm_host(lmrl_engine* e, int32_t impl, int32_t M, int32_t N, int32_t K, const uint16_t* A_bf16, const uint16_t* W_bf16,
                        const float* bias, float* resid_inout) {
    LMRL_CHECK(e && A_bf16 && W_bf16 && resid_inout, "lmrl_test_gemm: null argument");
    LMRL_CHECK(M > 0 && N % 64 == 0 && K % 64 == 0 && (impl == 3 || N % 128 == 0) && (impl == 0 || impl == 3 || N % TC_BN == 0),
               "lmrl_test_gemm: unsupported shape %d x %d x %d", M, N, K);
    LMRL_CHECK(impl >= 0 && impl <= 3, "lmrl_test_gemm: impl must be 0, 1, 2 or 3");
    LMRL_CHECK(impl != 3 || N % 192 == 0, "lmrl_test_gemm: impl 3 (192-wide tiles) needs N %% 192 == 0");
    LMRL_CUDA(cudaSetDevice(e->device));
    const size_t m_pad = ((size_t)M + 255) / 256 * 256;
    __nv_bfloat16 *dA = nullptr, *dW = nullptr;
    float *dB = nullptr, *dR = nullptr;
    int* dM = nullptr;
    LMRL_CUDA(cudaMalloc(&dA, m_pad * K * 2)); LMRL_CUDA(cudaMalloc(&dW, (size_t)N * K * 2));
    LMRL_CUDA(cudaMalloc(&dB, (size_t)N * 4)); LMRL_CUDA(cudaMalloc(&dR, m_pad * N * 4)); LMRL_CUDA(cudaMalloc(&dM, 4));
    LMRL_CUDA(cudaMemsetAsync(dA, 0, m_pad * K * 2, e->stream));
    LMRL_CUDA(cudaMemcpyAsync(dA, A_bf16, (size_t)M * K * 2, cudaMemcpyHostToDevice, e->stream));
    LMRL_CUDA(cudaMemcpyAsync(dW, W_bf16, (size_t)N * K * 2, cudaMemcpyHostToDevice, e->stream));
    if (bias) LMRL_CUDA(cudaMemcpyAsync(dB, bias, (size_t)N * 4, cudaMemcpyHostToDevice, e->stream));
    LMRL_CUDA(cudaMemcpyAsync(dR, resid_inout, (size_t)M * N * 4, cudaMemcpyHostToDevice, e->stream));
    LMRL_CUDA(cudaMemcpyAsync(dM, &M, 4, cudaMemcpyHostToDevice, e->stream));
    GemmArgs g{};
    g.A = dA; g.W = dW; g.bias = bias ? dB : nullptr; g.N = N; g.K = K; g.M_dev = dM; g.resid = dR;
    g.n_seg = 1; g.n_acc = 1; g.seg[0] = {0, 0, 0, K / TC_BK};
    cudaDeviceProp prop;
    LMRL_CUDA(cudaGetDeviceProperties(&prop, e->device));
    int rc = 0;
    if (impl == 3) {
        CUtensorMap ta, tw, to;
        rc = make_tmap_bf16(&ta, dA, m_pad, K, TC_BM) || make_tmap_bf16(&tw, dW, N, K, 192) || make_tmap(&to, dR, m_pad, N, 32, true);
        if (!rc) rc = launch_gemm_tc<EPI_RESID, 192>(e->stream, 0, prop.multiProcessorCount, ta, tw, tw, tw, to, g);
    } else if (impl == 2) {
        CUtensorMap ta, tw, to;
        rc = make_tmap_bf16(&ta, dA, m_pad, K, TC_BM) || make_tmap_bf16(&tw, dW, N, K, TC2_B_ROWS) || make_tmap(&to, dR, m_pad, N, 32, true);
        if (!rc) rc = launch_gemm_tc2<EPI_RESID>(e->stream, 0, prop.multiProcessorCount, ta, tw, to, g);
    } else if (impl == 1) {
        CUtensorMap ta, tw, to;
        rc = make_tmap_bf16(&ta, dA, m_pad, K, TC_BM) || make_tmap_bf16(&tw, dW, N, K, TC_BN) || make_tmap(&to, dR, m_pad, N, 32, true);
        if (!rc) rc = launch_gemm_tc<EPI_RESID>(e->stream, 0, prop.multiProcessorCount, ta, tw, tw, tw, to, g);
    } else {
        rc = launch_gemm_mma<EPI_RESID>(e->stream, 0, prop.multiProcessorCount, g);
    }
    e->launches++;
    cudaError_t err = cudaStreamSynchronize(e->stream);
    if (!rc && err == cudaSuccess) err = cudaMemcpy(resid_inout, dR, (size_t)M * N * 4, cudaMemcpyDeviceToHost);
    cudaFree(dA); cudaFree(dW); cudaFree(dB); cudaFree(dR); cudaFree(dM);
    if (rc) return -1;
    LMRL_CHECK(err == cudaSuccess, "lmrl_test_gemm: %s", cudaGetErrorString(err));
    return 0;
}

int lmrl_bench_kernels(lmrl_engine* e, int32_t M, int32_t ctx, int32_t iters, float* ms_out) {
    LMRL_CHECK(e && e->has_gpt2 && ms_out, "lmrl_bench_kernels: engine without GPT-2 parameters or null output");
    LMRL_CUDA(cudaSetDevice(e->device));
    Gpt2Dev& G = *e->gpt2;
    const lmrl_gpt2_dims& D = G.dims;
    const int B = e->cfg.max_envs;
    LMRL_CHECK(M > 0 && M <= G.m_cap && (M <= B || M % B == 0), "lmrl_bench_kernels: M=%d must be <= max_envs or a multiple of it (capacity %d)", M, G.m_cap);
    const int rpe = M <= B ? 1 : M / B, n_slots = M <= B ? M : B;
    LMRL_CHECK(ctx >= rpe && ctx <= G.t_cap && iters > 0, "lmrl_bench_kernels: ctx=%d must be in [%d, %d]", ctx, rpe, G.t_cap);
    cudaStream_t st = e->stream;
    bench_rows_kernel<<<(std::max(M, n_slots * G.kvp.ppe) + 255) / 256, 256, 0, st>>>(M, rpe, ctx, n_slots, G.rows, G.kvp);
    bench_fill_bf16_kernel<<<1024, 256, 0, st>>>(G.xn, (size_t)M * D.d_model, 0x1234u, 1.0f);
    bench_fill_bf16_kernel<<<1024, 256, 0, st>>>(G.att, (size_t)M * D.d_model, 0x2345u, 1.0f);
    bench_fill_bf16_kernel<<<1024, 256, 0, st>>>(G.q, (size_t)M * D.d_model, 0x3456u, 1.0f);
    bench_fill_bf16_kernel<<<1024, 256, 0, st>>>(G.h, (size_t)M * D.d_ff, 0x4567u, 1.0f);
    bench_fill_bf16_kernel<<<1024, 256, 0, st>>>(G.xl, (size_t)n_slots * G.xl_cols, 0x5678u, 1.0f);
    LMRL_CUDA(cudaMemsetAsync(G.x, 0, (size_t)M * D.d_model * sizeof(float), st));
    LMRL_CUDA(cudaMemsetAsync(G.best, 0, sizeof(unsigned long long) * n_slots, st));
    const LayerDev& L = G.tw[0].layers[0];
    GemmArgs a{};
    a.A = G.xn; a.W = L.w_qkv; a.bias = L.b_qkv; a.N = 3 * D.d_model; a.K = D.d_model; a.M_dev = G.rows.M;
    a.out = G.q; a.kcache = G.tw[0].kv; a.vcache = a.kcache + G.kv_layer_stride / 2;
    a.row_env = G.rows.row_env; a.row_pos = G.rows.row_pos; a.d_model = D.d_model; a.n_head = D.n_head; a.kv_table = G.kvp.table; a.kv_ppe = G.kvp.ppe;
    GemmArgs p{};
    p.A = G.att; p.W = L.w_proj; p.bias = L.b_proj; p.N = D.d_model; p.K = D.d_model; p.M_dev = G.rows.M; p.resid = G.x;
    GemmArgs f{};
    f.A = G.xn; f.W = L.w_fc; f.bias = L.b_fc; f.N = D.d_ff; f.K = D.d_model; f.M_dev = G.rows.M; f.out = G.h;
    GemmArgs f2{};
    f2.A = G.h; f2.W = L.w_fc2; f2.bias = L.b_fc2; f2.N = D.d_model; f2.K = D.d_ff; f2.M_dev = G.rows.M; f2.resid = G.x;
    if (G.ln_fold && G.gemm_impl >= 1 && L.w_qkv_f && G.tw[0].fold_valid && G.lane[0].stat[0] && D.d_model % TC_BN == 0) {
        // the LayerNorm-folded launch sequence: residual GEMMs with the bf16 copy + statistics epilogue, consumers scaling by 1 / sigma
        const int parts = 2 * (D.d_model / ((G.narrow_decode_tiles && rpe == 1 && D.d_model == 768) ? 192 : TC_BN));
        a.W = L.w_qkv_f; a.bias = L.b_qkv_f; a.stat_in = G.lane[0].stat[0]; a.stat_n = parts; a.ln_eps = D.ln_eps;
        f.W = L.w_fc_f; f.bias = L.b_fc_f; f.stat_in = G.lane[0].stat[1]; f.stat_n = parts; f.ln_eps = D.ln_eps;
        p.xb_out = G.xn; p.stat_out = G.lane[0].stat[1];
        f2.xb_out = G.xn; f2.stat_out = G.lane[0].stat[0];
        LMRL_CUDA(cudaMemsetAsync(G.lane[0].stat[0], 0, (size_t)M * 8 * sizeof(float2), st));
        LMRL_CUDA(cudaMemsetAsync(G.lane[0].stat[1], 0, (size_t)M * 8 * sizeof(float2), st));
    }
    GemmArgs h{};
    h.A = G.xl; h.W = G.tw[0].wte; h.bias = nullptr; h.N = G.v_pad; h.K = D.d_model; h.M_dev = G.rows.n_slots;
    h.best = G.best; h.row_key = G.row_key; h.inv_temp = 1.0f; h.do_sample = 1; h.n_vocab = D.n_vocab; h.logits_out = nullptr;
    const size_t xl_rows = (size_t)(B + 127) / 128 * 128;
    GemmExtra hx;
    hx.a_cols = (size_t)G.xl_cols;
    const int ew_blocks = G.sm_count * 8;
    const int att_blocks = std::min((n_slots * D.n_head + ATT_WARPS - 1) / ATT_WARPS, G.sm_count * 8);
    cudaEvent_t ev[2];
    LMRL_CUDA(cudaEventCreate(&ev[0])); LMRL_CUDA(cudaEventCreate(&ev[1]));
    const int saved_profile = e->profile;
    e->profile = 0;
    e->cur_mode = rpe > 1 ? 0 : 1;                     // prefill-shaped batches take the prefill kernel selection
    int rc = 0;
    for (int k = 0; k < 8 && !rc; ++k) {
        for (int it = -2; it < iters && !rc; ++it) {      // two untimed warm-up launches
            if (it == 0) cudaEventRecord(ev[0], st);
            switch (k) {
                case 0: rc = launch_gemm<EPI_QKV>(e, a, a.N, 0, G.m_cap, a.N, 0); break;
                case 1: rc = launch_gemm<EPI_RESID>(e, p, p.N, 0, G.m_cap, p.N, 1); break;
                case 2: rc = launch_gemm<EPI_GELU>(e, f, f.N, 0, G.m_cap, f.N, 2); break;
                case 3: rc = launch_gemm<EPI_RESID>(e, f2, f2.N, 0, G.m_cap, f2.N, 3); break;
                case 4: rc = launch_gemm<EPI_SAMPLE>(e, h, D.n_vocab, 1, xl_rows, G.v_pad, 4, hx); break;
                case 5:
                    if (rpe > 1)
                        LMRL_K(attention_prefill_kernel, attp_grid(n_slots, D.n_head, G.sm_count), ATTP_WARPS * 32, ATTP_WARPS * ATTP_WARP_SMEM, st, G.q, a.kcache, a.vcache, G.rows, D.n_head, D.d_model, G.kvp,
                                                                                      0.125f * 1.4426950408889634f, G.attn_wide, G.att);
                    else if (G.att_decode_impl == 2)
                        LMRL_K(attention_pipe_kernel, att_blocks, ATT_WARPS * 32, 0, st, G.q, a.kcache, a.vcache, G.rows, D.n_head, D.d_model, G.kvp, 0.125f, G.att);
                    else if (G.att_decode_impl == 3 && D.n_head % 2 == 0)
                        LMRL_K(attention_hpw_kernel<2>, std::min(att_blocks, (n_slots * D.n_head / 2 + ATT_WARPS - 1) / ATT_WARPS), ATT_WARPS * 32, 0, st, G.q, a.kcache, a.vcache, G.rows, D.n_head, D.d_model, G.kvp, 0.125f, G.att);
                    else if (G.att_decode_impl == 4 && D.n_head % 4 == 0)
                        LMRL_K(attention_hpw_kernel<4>, std::min(att_blocks, (n_slots * D.n_head / 4 + ATT_WARPS - 1) / ATT_WARPS), ATT_WARPS * 32, 0, st, G.q, a.kcache, a.vcache, G.rows, D.n_head, D.d_model, G.kvp, 0.125f, G.att);
                    else
                        LMRL_K(attention_kernel, att_blocks, ATT_WARPS * 32, 0, st, G.q, a.kcache, a.vcache, G.rows, D.n_head, D.d_model, G.kvp, 0.125f, G.att);
                    break;
                case 6: LMRL_K(layernorm_kernel, ew_blocks, 256, 0, st, G.x, L.ln1_g, L.ln1_b, D.d_model, D.ln_eps, G.rows.M, nullptr, G.xn, D.d_model); break;
                case 7: LMRL_K(embed_kernel, ew_blocks, 256, 0, st, G.tw[0].wte, G.tw[0].wpe, D.d_model, G.rows, G.x); break;
            }
        }
        cudaEventRecord(ev[1], st);
        cudaError_t err = cudaStreamSynchronize(st);
        if (err != cudaSuccess) { e->profile = saved_profile; return lmrl::fail("lmrl_bench_kernels: kernel %d failed: %s", k, cudaGetErrorString(err)); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[0], ev[1]);
        ms_out[k] = ms / iters;
    }
    e->profile = saved_profile;
    cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
    return rc;
}

int lmrl_profile_enable(lmrl_engine* e, int32_t enable) {
    LMRL_CHECK(e, "null engine");
    LMRL_CUDA(cudaSetDevice(e->device));
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    e->profile = enable ? 1 : 0;
    e->prof.clear();
    e->prof_closed_at = 0;
    e->n_passes = 0;
    if (enable && !e->h_counts) {
        e->max_passes = 1 << 16;
        LMRL_CUDA(cudaMallocHost(&e->h_counts, sizeof(int32_t) * 3 * e->max_passes));
    }
    return 0;
}

int lmrl_profile_read(lmrl_engine* e, lmrl_profile_out* out) {
    LMRL_CHECK(e && out, "null argument");
    LMRL_CUDA(cudaSetDevice(e->device));
    if (e->prof_closed_at != e->prof.size()) e->prof_close();
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    memset(out, 0, sizeof *out);
    for (const auto& r : e->prof) {
        float ms = 0.f;
        LMRL_CUDA(cudaEventElapsedTime(&ms, e->ev_pool[r.ev], e->ev_pool[r.ev + 1]));
        out->ms[r.cat] += ms;
        out->launches[r.cat] += 1;
        double work = r.count_slot >= 0 ? r.work_per_row * (double)e->h_counts[r.count_slot] : 0.0;
        if (r.count_slot2 >= 0 && r.work2_per_row != 0.0) work += r.work2_per_row * (double)e->h_counts[r.count_slot2];
        out->work[r.cat] += work;
        if (r.tag >= 0 && r.tag < 12) { out->tag_ms[r.tag] += ms; out->tag_work[r.tag] += work; }
    }
    for (int p = 0; p < e->n_passes; ++p) { out->rows += e->h_counts[3 * p]; out->sampled += e->h_counts[3 * p + 1]; }
    return 0;
}

}  // extern "C"
