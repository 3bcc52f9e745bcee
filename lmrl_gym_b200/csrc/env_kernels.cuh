// env_kernels.cuh — device-resident Wordle / Maze environments: one warp per environment.
//
// Kernels E1..E6 of SURVEY.md §2.2 (reset, action decode + normalise + validate, Wordle transition, observation
// render, Maze transition, rollout bookkeeping).  All state is int32/uint32 in HBM, structure-of-arrays over the
// environment index; the per-env logic is env_core.h (shared with the host test build), the vocabulary scans are
// warp-parallel here (ballot + popc + fns, order preserving: the target is the k-th survivor in file order,
// reference llm_rl_scripts/wordle/env/game.py:154, 179).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "env_core.h"
#include "pdl.cuh"

namespace lmrl {

constexpr int ENV_WARPS_PER_BLOCK = 8;
constexpr int MAX_ACTION_CP = 256;      // code points kept per action (longer actions are invalid anyway)
constexpr int MAX_LINE_TOKENS = 4 * MAX_ACTION_CP + 8;

struct TokenizerDev {
    const uint8_t* token_bytes;   // flat bytes of all tokens
    const int32_t* offsets;       // [n_tokens + 1]
    int32_t n_tokens;
    TokTables t;
};

struct WordleDev {
    // tables
    const uint32_t* words; int32_t n_words;
    int32_t require_in_vocab; float bad_word_reward;
    // state, [B] or [B][k]
    uint32_t* know;        // [B][10]  here[5], ban[5]
    uint32_t* mt;          // [B][624]
    int32_t* mt_idx;       // [B]
    int32_t* n_actions;    // [B]
    uint32_t* hist;        // [B][6] packed well-formed actions, WORD_NONE otherwise
    int32_t* n_filtered;   // [B]
    uint8_t* done;         // [B]
    uint32_t* target;      // [B] last sampled target (diagnostic / tests)
};

struct MazeDev {
    MazeCfg cfg;
    const uint8_t* grid;          // [H*W]
    const int32_t* free_cells;    // [n_free] y*W+x, row-major order (np.argwhere order, env.py:195)
    int32_t n_free;
    int32_t goal_y, goal_x;       // the single valid goal (maze_utils.py:11-18)
    // state
    int32_t* pos;                 // [B][2]
    int32_t* goal;                // [B][2]
    int32_t* num_steps;           // [B]
    uint8_t* done;                // [B]
    // fused rollout only: canonical token ids of the closed text set + the sliding last_k window over history items
    const int32_t* obs_tokens;    // [H*W][obs_pitch] observation sentence of every cell (default goal)
    const int32_t* obs_len;       // [H*W]
    int32_t obs_pitch;
    const int32_t* act_tokens;    // [4][MAZE_ACT_PITCH] canonical ids of 'move left\n' ... (what re-tokenising the history gives)
    const int32_t* act_len;       // [4]
    int32_t* item_len;            // [B][MAZE_RING] token length of every history item inside the window, oldest first
    int32_t* n_items;             // [B]
    int32_t* cut;                 // [B] leading tokens of the window that LEFT truncation removed from the stream
};
constexpr int MAZE_RING = 72;     // last_k <= 64 items (+ the action / observation pair being appended)
constexpr int MAZE_ACT_PITCH = 8;
constexpr int MAZE_ACT_BYTES = 64;

// step outputs (one row per env)
struct StepOut {
    float* reward;        // [B]
    uint8_t* done;        // [B]
    uint8_t* valid;       // [B]   Wordle: action was a legal guess (reward is a Python int);  Maze: known move
    uint8_t* fb;          // [B][5] Wordle feedback codes (0 g, 1 y, 2 b, 255 none)
    int32_t* aux;         // [B]   Wordle: filtered-vocabulary size after the step;  Maze: outcome (0/1/2)
};

// the per-env token stream the policy reads (fused path)
struct StreamDev {
    int32_t* tokens;      // [B][cap]
    int32_t* len;         // [B]  tokens in the stream
    int32_t* kv_len;      // [B]  leading tokens whose KV entries are valid
    int32_t cap;          // row pitch
    int32_t max_input;    // LEFT truncation length (BlockingStrategy max_length); stream is shifted when exceeded
    // token-trajectory annotations of the stream (PPO data path: TokenTrajectory.from_text_trajectory, environment.py:346-383), Wordle only:
    uint8_t* is_action;   // [B][cap] the token belongs to an action Text, or nullptr
    float* tok_reward;    // [B][cap] the Text item's reward on its LAST token, 0 elsewhere
    uint8_t* truncated;   // [B] the stream was LEFT-truncated at least once (it no longer starts at the episode's first token)
};

struct ActionSrc {
    // text path: UTF-8 bytes per env
    const uint8_t* bytes; const int32_t* nbytes; int32_t pitch;
    // fused path: sampled token ids per env (bytes == nullptr)
    const int32_t* tokens; const int32_t* ntokens; int32_t tpitch;
    // scripted path (env microbenchmark): the device equivalent of RandomMixturePolicy (wordle/env/scripted_policies.py:114-126):
    // with probability prob_smart a uniformly random word of the FILTERED vocabulary, else of the whole vocabulary;
    // draws come from a counter hash of (policy_seed, episode seed, turn) instead of the reference's global `random`
    int32_t scripted; float prob_smart; uint64_t policy_seed; const uint64_t* episode_seed;
};

__device__ __forceinline__ uint64_t env_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- E1: reset -------------------------------------------------------------------------------------------
__global__ void wordle_reset_kernel(WordleDev w, const uint64_t* __restrict__ seeds, int n) {
    pdl_enter();
    extern __shared__ uint32_t smem[];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int e = blockIdx.x * ENV_WARPS_PER_BLOCK + warp;
    if (e >= n) return;
    uint32_t* mt = smem + warp * MT_N;
    __shared__ int idx_s[ENV_WARPS_PER_BLOCK];
    if (lane == 0) mt_seed(mt, &idx_s[warp], seeds[e]);
    __syncwarp();
    for (int i = lane; i < MT_N; i += 32) w.mt[(size_t)e * MT_N + i] = mt[i];
    if (lane < 10) w.know[e * 10 + lane] = 0;
    if (lane < WORDLE_TRIES) w.hist[e * WORDLE_TRIES + lane] = WORD_NONE;
    if (lane == 0) {
        w.mt_idx[e] = idx_s[warp];
        w.n_actions[e] = 0;
        w.n_filtered[e] = w.n_words;       // the initial table accepts every word (game.py:54-55)
        w.done[e] = 0;
        w.target[e] = WORD_NONE;
    }
}

// warp-parallel: k-th (0-based) word accepted by m, in table order; WORD_NONE if fewer than k+1
__device__ __forceinline__ uint32_t warp_select_kth(const WordleMasks& m, const uint32_t* __restrict__ words, int n_words, int k) {
    const int lane = lane_id();
    int seen = 0;
    for (int base = 0; base < n_words; base += 32) {
        const int i = base + lane;
        const uint32_t w = i < n_words ? __ldg(words + i) : 0u;
        const unsigned ok = __ballot_sync(0xFFFFFFFFu, i < n_words && word_ok(m, w));
        const int c = __popc(ok);
        if (k < seen + c) {
            const int src = __fns(ok, 0, k - seen + 1);
            return __shfl_sync(0xFFFFFFFFu, w, src);
        }
        seen += c;
    }
    return WORD_NONE;
}

// warp-parallel: number of accepted words and the first one
__device__ __forceinline__ int warp_count_first(const WordleMasks& m, const uint32_t* __restrict__ words, int n_words, uint32_t* first) {
    const int lane = lane_id();
    int seen = 0;
    uint32_t f = WORD_NONE;
    for (int base = 0; base < n_words; base += 32) {
        const int i = base + lane;
        const uint32_t w = i < n_words ? __ldg(words + i) : 0u;
        const unsigned ok = __ballot_sync(0xFFFFFFFFu, i < n_words && word_ok(m, w));
        if (seen == 0 && ok) f = __shfl_sync(0xFFFFFFFFu, w, __ffs(ok) - 1);
        seen += __popc(ok);
    }
    *first = f;
    return seen;
}

__device__ __forceinline__ bool warp_in_vocab(const uint32_t* __restrict__ words, int n_words, uint32_t a) {
    const int lane = lane_id();
    bool hit = false;
    for (int base = 0; base < n_words; base += 32) {
        const int i = base + lane;
        hit = hit || (i < n_words && __ldg(words + i) == a);
    }
    return __any_sync(0xFFFFFFFFu, hit);
}

// ---- E2 + E3 + E4: one Wordle step for every active env ----------------------------------------------------
// Per warp scratch in shared memory: code points [MAX_ACTION_CP] + MT state [624].
struct WordleScratch {
    uint32_t cp[MAX_ACTION_CP];
    uint32_t mt[MT_N];
    uint8_t bytes[MAX_ACTION_CP];
    int idx;
};

__global__ void wordle_step_kernel(WordleDev w, TokenizerDev tk, ActionSrc src, const uint8_t* __restrict__ skip,
                                   StepOut out, StreamDev st, int n) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int e = blockIdx.x * ENV_WARPS_PER_BLOCK + warp;
    if (e >= n) return;
    if (w.done[e] || (skip && skip[e])) return;      // finished episodes idle (environment.py:82, 187)
    WordleScratch& s = reinterpret_cast<WordleScratch*>(smem_raw)[warp];

    // E2: action text -> code points -> strip / de-space -> packed guess
    int ncp = 0;
    if (src.scripted) {
        WordleKnow k0;
#pragma unroll
        for (int i = 0; i < 5; ++i) { k0.here[i] = w.know[e * 10 + i]; k0.ban[i] = w.know[e * 10 + 5 + i]; }
        const uint64_t h = env_mix64(env_mix64(src.policy_seed ^ env_mix64(src.episode_seed[e])) + (uint64_t)w.n_actions[e]);
        const bool smart = (float)((h >> 40) & 0xFFFFFFull) * (1.0f / 16777216.0f) < src.prob_smart;
        const int nf = w.n_filtered[e];
        uint32_t word;
        if (smart && nf > 0) word = warp_select_kth(wordle_masks(k0), w.words, w.n_words, (int)((uint32_t)h % (uint32_t)nf));
        else word = __ldg(w.words + (uint32_t)h % (uint32_t)w.n_words);
        if (lane == 0)
            for (int i = 0; i < 5; ++i) s.cp[i] = 'a' + ((word >> (5 * i)) & 31u);
        ncp = 5;
    } else if (lane == 0) {
        const uint8_t* bytes; int nb;
        if (src.bytes) { bytes = src.bytes + (size_t)e * src.pitch; nb = min(src.nbytes[e], MAX_ACTION_CP); }
        else {
            nb = 0;
            const int nt = src.ntokens[e];
            for (int k = 0; k < nt; ++k) {
                const int t = src.tokens[(size_t)e * src.tpitch + k];
                if (t < 0 || t >= tk.n_tokens) continue;
                for (int o = tk.offsets[t]; o < tk.offsets[t + 1] && nb < MAX_ACTION_CP; ++o) s.bytes[nb++] = tk.token_bytes[o];
            }
            bytes = s.bytes;
        }
        ncp = utf8_decode_replace(bytes, nb, s.cp, MAX_ACTION_CP);
        ncp = normalise_action(s.cp, ncp);
    }
    ncp = __shfl_sync(0xFFFFFFFFu, ncp, 0);
    __syncwarp();
    const uint32_t guess = pack_action(s.cp, ncp);
    const bool wellformed = guess != WORD_NONE;
    const bool in_vocab = wellformed && warp_in_vocab(w.words, w.n_words, guess);
    const bool valid = wellformed && (!w.require_in_vocab || in_vocab);       // game.py:214

    WordleKnow k;
#pragma unroll
    for (int i = 0; i < 5; ++i) { k.here[i] = w.know[e * 10 + i]; k.ban[i] = w.know[e * 10 + 5 + i]; }
    int n_act = w.n_actions[e];
    int n_filt = w.n_filtered[e];
    uint32_t first = WORD_NONE;

    if (valid) {
        // E3: target = rng.choice(filtered) (game.py:219) -> transition (:220) -> re-filter (:221)
        for (int i = lane; i < MT_N; i += 32) s.mt[i] = w.mt[(size_t)e * MT_N + i];
        __syncwarp();
        int kth = 0;
        if (lane == 0) { s.idx = w.mt_idx[e]; kth = (int)mt_randbelow(s.mt, &s.idx, (uint32_t)n_filt); }
        kth = __shfl_sync(0xFFFFFFFFu, kth, 0);
        __syncwarp();
        for (int i = lane; i < MT_N; i += 32) w.mt[(size_t)e * MT_N + i] = s.mt[i];
        const uint32_t target = warp_select_kth(wordle_masks(k), w.words, w.n_words, kth);
        wordle_apply(k, guess, target);
        n_filt = warp_count_first(wordle_masks(k), w.words, w.n_words, &first);
        if (lane < 5) { w.know[e * 10 + lane] = k.here[lane]; w.know[e * 10 + 5 + lane] = k.ban[lane]; }
        if (lane == 0) { w.mt_idx[e] = s.idx; w.n_filtered[e] = n_filt; w.target[e] = target; }
    }

    // history, reward (game.py:290-293), terminal (:295-296), feedback (:273-288)
    bool in_hist = false;                                              // filtered[0] in action_history (incl. this action)
    if (n_filt == 1 && first != WORD_NONE) {
        in_hist = (wellformed && guess == first);
        for (int i = 0; i < min(n_act, WORDLE_TRIES); ++i) in_hist = in_hist || (w.hist[e * WORDLE_TRIES + i] == first);
    }
    __syncwarp();
    if (lane == 0 && n_act < WORDLE_TRIES) w.hist[e * WORDLE_TRIES + n_act] = wellformed ? guess : WORD_NONE;
    n_act += 1;
    // reward() tests `len != 5 or not in vocab` regardless of require_words_in_vocab (game.py:291)
    const float reward = !(wellformed && in_vocab) ? w.bad_word_reward : ((n_filt == 1 && in_hist) ? 0.0f : -1.0f);
    const bool done = (n_act == WORDLE_TRIES) || (reward == 0.0f);
    uint8_t fb[5] = {255, 255, 255, 255, 255};
    if (valid) wordle_feedback(k, guess, fb);

    if (lane == 0) {
        w.n_actions[e] = n_act;
        w.done[e] = done ? 1 : 0;
        out.reward[e] = reward;
        out.done[e] = done ? 1 : 0;
        out.valid[e] = valid ? 1 : 0;
        out.aux[e] = n_filt;
        for (int i = 0; i < 5; ++i) out.fb[e * 5 + i] = fb[i];
        // E4: append the canonical rendering of (action line, feedback line) to the policy's token stream
        if (st.tokens) {
            int32_t* row = st.tokens + (size_t)e * st.cap;
            int len = st.len[e];
            const int keep = len;                       // KV of everything before the action stays valid
            len += render_spaced_line(s.cp, ncp, tk.t, row + len, st.cap - len);
            const int act_end = len;
            len += render_feedback_line(fb, valid, tk.t, row + len, st.cap - len);
            if (st.is_action) {                         // the action Text carries the step's reward on its last token, the feedback none
                uint8_t* ia = st.is_action + (size_t)e * st.cap;
                float* tr = st.tok_reward + (size_t)e * st.cap;
                for (int i = keep; i < len; ++i) { ia[i] = i < act_end ? 1 : 0; tr[i] = 0.f; }
                if (act_end > keep) tr[act_end - 1] = reward;
            }
            int kv = min(keep, st.kv_len[e]);           // (a folded '\n\n' prompt tail leaves kv_len one short of the stream: gpt2_kernels.cuh)
            if (len > st.max_input) {                   // LEFT truncation (BlockingStrategy, eval_bc_gpt2.py:100-104)
                const int drop = len - st.max_input;
                for (int i = 0; i < st.max_input; ++i) row[i] = row[i + drop];
                if (st.is_action) {
                    uint8_t* ia = st.is_action + (size_t)e * st.cap;
                    float* tr = st.tok_reward + (size_t)e * st.cap;
                    for (int i = 0; i < st.max_input; ++i) { ia[i] = ia[i + drop]; tr[i] = tr[i + drop]; }
                    st.truncated[e] = 1;
                }
                len = st.max_input;
                kv = 0;                                 // positions shifted: the whole window is re-prefilled
            }
            st.len[e] = len;
            st.kv_len[e] = kv;
        }
    }
}

// ---- Maze --------------------------------------------------------------------------------------------------
// reset: random.seed(seed); goal = choice(valid_goals) (one goal: 1-bit rejection draws); start = choice(free - goal)
// unless options carry them (env.py:186-202).  opt_pos/opt_goal: [B][2], -1 = not given.
__global__ void maze_reset_kernel(MazeDev m, const uint64_t* __restrict__ seeds, const int32_t* __restrict__ opt_pos,
                                  const int32_t* __restrict__ opt_goal, int n) {
    pdl_enter();
    extern __shared__ uint32_t smem[];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int e = blockIdx.x * ENV_WARPS_PER_BLOCK + warp;
    if (e >= n) return;
    uint32_t* mt = smem + warp * MT_N;
    if (lane == 0) {
        int idx;
        mt_seed(mt, &idx, seeds[e]);
        int gy = m.goal_y, gx = m.goal_x;
        if (opt_goal && opt_goal[e * 2] >= 0) { gy = opt_goal[e * 2]; gx = opt_goal[e * 2 + 1]; }
        else (void)mt_randbelow(mt, &idx, 1u);                       // choice over the single valid goal
        int py, px;
        if (opt_pos && opt_pos[e * 2] >= 0) { py = opt_pos[e * 2]; px = opt_pos[e * 2 + 1]; }
        else {
            // free cells in row-major order with the goal removed (list.remove: first occurrence)
            int n_free = 0, goal_at = -1;
            for (int i = 0; i < m.n_free; ++i) if (goal_at < 0 && m.free_cells[i] == gy * m.cfg.W + gx) goal_at = i;
            n_free = m.n_free - (goal_at >= 0 ? 1 : 0);
            int k = (int)mt_randbelow(mt, &idx, (uint32_t)n_free);
            if (goal_at >= 0 && k >= goal_at) ++k;
            py = m.free_cells[k] / m.cfg.W; px = m.free_cells[k] % m.cfg.W;
        }
        m.pos[e * 2] = py; m.pos[e * 2 + 1] = px;
        m.goal[e * 2] = gy; m.goal[e * 2 + 1] = gx;
        m.num_steps[e] = 0;
        m.done[e] = 0;
    }
}

__global__ void maze_step_kernel(MazeDev m, ActionSrc src, const uint8_t* __restrict__ skip, StepOut out,
                                 int32_t* __restrict__ action_id, int n) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (m.done[e] || (skip && skip[e])) return;
    const uint8_t* bytes = src.bytes + (size_t)e * src.pitch;
    const int a = maze_match_action(bytes, src.nbytes[e]);
    const int steps = m.num_steps[e];
    const MazeStep r = maze_step(m.cfg, m.grid, m.pos[e * 2], m.pos[e * 2 + 1], m.goal[e * 2], m.goal[e * 2 + 1], steps, a);
    m.pos[e * 2] = r.y; m.pos[e * 2 + 1] = r.x;
    if (r.outcome == 0) m.num_steps[e] = steps + 1;                 // env.py:177 (not on Success / Failure)
    m.done[e] = (uint8_t)r.done;
    out.reward[e] = r.reward;
    out.done[e] = (uint8_t)r.done;
    out.valid[e] = a >= 0;
    out.aux[e] = r.outcome;
    if (action_id) action_id[e] = a;
}

// ---- fused Maze loop (policy tokens in, policy tokens out; maze/env/env.py:161-184 + the policy-side string handling) ----
// window bookkeeping: remove `n_drop` oldest items from the stream of env e (tokens already cut by truncation are accounted first)
__device__ __forceinline__ int maze_drop_items(MazeDev& m, int e, int n_drop, int32_t* row, int len) {
    int32_t* il = m.item_len + (size_t)e * MAZE_RING;
    int n = m.n_items[e], cut = m.cut[e], removed = 0;
    for (int i = 0; i < n_drop && i < n; ++i) {
        const int L = il[i];
        const int gone = min(cut, L);                  // part of this item that truncation had already removed
        cut -= gone;
        removed += L - gone;
    }
    for (int i = n_drop; i < n; ++i) il[i - n_drop] = il[i];
    m.n_items[e] = max(n - n_drop, 0);
    m.cut[e] = cut;
    if (removed > 0) for (int i = removed; i < len; ++i) row[i - removed] = row[i];
    return len - removed;
}

// after maze_reset_kernel: the policy's first prompt is the observation of the start cell (env.py:204-214)
__global__ void maze_init_stream_kernel(MazeDev m, StreamDev st, int n) {
    pdl_enter();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int cell = m.pos[e * 2] * m.cfg.W + m.pos[e * 2 + 1];
    const int L = m.obs_len[cell];
    int32_t* row = st.tokens + (size_t)e * st.cap;
    for (int i = 0; i < L; ++i) row[i] = m.obs_tokens[(size_t)cell * m.obs_pitch + i];
    int len = L;
    m.item_len[(size_t)e * MAZE_RING] = L; m.n_items[e] = 1; m.cut[e] = 0;
    if (len > st.max_input) {
        const int drop = len - st.max_input;
        for (int i = 0; i < st.max_input; ++i) row[i] = row[i + drop];
        len = st.max_input; m.cut[e] = drop;
    }
    st.len[e] = len; st.kv_len[e] = 0;
}

// one env per thread: sampled ids -> action text (decode, `removesuffix('\n') + '\n'` as every reference driver's
// out_str_process) -> MazeEnv.step -> next prompt window (last_k items, unknown action collapses to the observation)
__global__ void maze_fused_step_kernel(MazeDev m, TokenizerDev tk, ActionSrc src, StepOut out, StreamDev st, int n) {
    pdl_enter();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (m.done[e]) return;
    uint8_t bytes[MAZE_ACT_BYTES + 2];
    int nb = 0;
    bool overflow = false;
    const int nt = src.ntokens[e];
    for (int k = 0; k < nt; ++k) {
        const int t = src.tokens[(size_t)e * src.tpitch + k];
        if (t < 0 || t >= tk.n_tokens) continue;
        for (int o = tk.offsets[t]; o < tk.offsets[t + 1]; ++o) {
            if (nb < MAZE_ACT_BYTES) bytes[nb++] = tk.token_bytes[o];
            else overflow = true;
        }
    }
    if (nb == 0 || bytes[nb - 1] != (uint8_t)'\n') bytes[nb++] = (uint8_t)'\n';
    const int a = overflow ? -1 : maze_match_action(bytes, nb);
    const int steps = m.num_steps[e];
    const int cell0 = m.pos[e * 2] * m.cfg.W + m.pos[e * 2 + 1];
    const MazeStep r = maze_step(m.cfg, m.grid, m.pos[e * 2], m.pos[e * 2 + 1], m.goal[e * 2], m.goal[e * 2 + 1], steps, a);
    m.pos[e * 2] = r.y; m.pos[e * 2 + 1] = r.x;
    if (r.outcome == 0) m.num_steps[e] = steps + 1;
    m.done[e] = (uint8_t)r.done;
    const int cell = r.y * m.cfg.W + r.x;
    out.reward[e] = r.reward;
    out.done[e] = (uint8_t)r.done;
    out.valid[e] = a >= 0;
    out.aux[e] = r.outcome | (cell << 8) | (cell0 << 20);
    for (int i = 0; i < 5; ++i) out.fb[e * 5 + i] = 255;
    if (r.done) return;
    int32_t* row = st.tokens + (size_t)e * st.cap;
    int32_t* il = m.item_len + (size_t)e * MAZE_RING;
    const int oL = m.obs_len[cell];
    const int32_t* obs = m.obs_tokens + (size_t)cell * m.obs_pitch;
    int len = st.len[e], kv;
    if (a < 0) {                                        // env.py:179-180: the history collapses to the new observation
        for (int i = 0; i < oL; ++i) row[i] = obs[i];
        len = oL; il[0] = oL; m.n_items[e] = 1; m.cut[e] = 0; kv = 0;
    } else {
        kv = len;                                       // everything before the action keeps its KV entries
        const int aL = m.act_len[a];
        for (int i = 0; i < aL; ++i) row[len + i] = m.act_tokens[a * MAZE_ACT_PITCH + i];
        len += aL;
        for (int i = 0; i < oL; ++i) row[len + i] = obs[i];
        len += oL;
        int ni = m.n_items[e];
        il[ni] = aL; il[ni + 1] = oL; ni += 2;
        m.n_items[e] = ni;
        if (ni > m.cfg.last_k) {                        // env.py:183 new_history[-last_k:]
            len = maze_drop_items(m, e, ni - m.cfg.last_k, row, len);
            kv = 0;                                     // positions restart at the new window start
        }
    }
    if (len > st.max_input) {                           // LEFT truncation of the policy input (BlockingStrategy)
        const int drop = len - st.max_input;
        for (int i = 0; i < st.max_input; ++i) row[i] = row[i + drop];
        len = st.max_input; m.cut[e] += drop; kv = 0;
    }
    st.len[e] = len; st.kv_len[e] = kv;
}

// scripted rollouts: per-turn log (an env acted this turn iff its action count ran ahead of the logged length)
__global__ void log_scripted_turn_kernel(int n, int max_turns, int act_pitch, WordleDev w, StepOut step, float* ep_return, int32_t* ep_length,
                                         uint8_t* ep_done, float* step_reward, uint8_t* step_valid, uint8_t* step_fb, int32_t* step_aux,
                                         int32_t* act_tokens, int32_t* act_len) {
    pdl_enter();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int t = ep_length[e];
    if (w.n_actions[e] != t + 1 || t >= max_turns) return;
    const size_t o = (size_t)e * max_turns + t;
    if (step_reward) step_reward[o] = step.reward[e];
    if (step_valid) step_valid[o] = step.valid[e];
    if (step_aux) step_aux[o] = step.aux[e];
    if (step_fb) for (int i = 0; i < 5; ++i) step_fb[o * 5 + i] = step.fb[e * 5 + i];
    if (act_tokens) act_tokens[o * act_pitch] = (int32_t)w.hist[e * WORDLE_TRIES + t];      // the packed 5-letter guess
    if (act_len) act_len[o] = 1;
    ep_return[e] += step.reward[e];
    ep_length[e] = t + 1;
    ep_done[e] = step.done[e];
}
__global__ void clear_episode_log_kernel(int n, float* ep_return, int32_t* ep_length, uint8_t* ep_done) {
    pdl_enter();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) { ep_return[e] = 0.f; ep_length[e] = 0; ep_done[e] = 0; }
}

// live-episode count for the periodic early-exit test of long rollouts
__global__ void count_live_kernel(const uint8_t* __restrict__ done, int n, int* __restrict__ live) {
    pdl_enter();
    int c = 0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) c += done[e] ? 0 : 1;
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(live, c);
}

}  // namespace lmrl
