// envcore_host.cpp — host build of env_core.h for the CPU test-suite (tests/test_envcore_host.py).
// Same per-environment logic the kernels run, with serial vocabulary scans instead of warp ballots.  TEST ONLY:
// the product never calls into this library.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "env_core.h"
using namespace lmrl;

extern "C" {

void hc_mt_u32_stream(uint64_t seed, int n, uint32_t* out) {
    std::vector<uint32_t> mt(MT_N);
    int idx;
    mt_seed(mt.data(), &idx, seed);
    for (int i = 0; i < n; ++i) out[i] = mt_u32(mt.data(), &idx);
}

void hc_mt_randbelow_stream(uint64_t seed, const uint32_t* ns, int n, uint32_t* out) {
    std::vector<uint32_t> mt(MT_N);
    int idx;
    mt_seed(mt.data(), &idx, seed);
    for (int i = 0; i < n; ++i) out[i] = mt_randbelow(mt.data(), &idx, ns[i]);
}

int hc_utf8_decode(const uint8_t* s, int n, uint32_t* out, int cap) { return utf8_decode_replace(s, n, out, cap); }

int hc_normalise(const uint8_t* s, int n, uint32_t* out, int cap) {
    int m = utf8_decode_replace(s, n, out, cap);
    return normalise_action(out, m);
}

int hc_render_action_line(const uint8_t* s, int n, const int32_t* tok_first, const int32_t* tok_space, const int32_t* byte_tok,
                          int space_id, int newline_id, int32_t* out, int cap) {
    uint32_t cp[1024];
    int m = utf8_decode_replace(s, n, cp, 1024);
    m = normalise_action(cp, m);
    TokTables t{tok_first, tok_space, byte_tok, space_id, newline_id};
    return render_spaced_line(cp, m, t, out, cap);
}

// One Wordle episode: actions are UTF-8 strings concatenated in `bytes` with lengths `lens`.
// Outputs per step: reward, done, valid, fb[5], n_filtered, target.  Returns steps taken.
int hc_wordle_episode(const uint32_t* words, int n_words, int require, float bad_reward, uint64_t seed,
                      const uint8_t* bytes, const int32_t* lens, int n_actions, float* reward, uint8_t* done,
                      uint8_t* valid, uint8_t* fb, int32_t* n_filtered, uint32_t* targets) {
    std::vector<uint32_t> mt(MT_N);
    int idx;
    mt_seed(mt.data(), &idx, seed);
    WordleKnow k;
    wordle_clear(k);
    uint32_t hist[WORDLE_TRIES];
    int n_act = 0, n_filt = n_words;
    int off = 0;
    for (int a = 0; a < n_actions; ++a) {
        uint32_t cp[1024];
        int ncp = utf8_decode_replace(bytes + off, lens[a], cp, 1024);
        off += lens[a];
        ncp = normalise_action(cp, ncp);
        const uint32_t guess = pack_action(cp, ncp);
        const bool wellformed = guess != WORD_NONE;
        bool in_vocab = false;
        if (wellformed) for (int i = 0; i < n_words; ++i) in_vocab = in_vocab || words[i] == guess;
        const bool ok = wellformed && (!require || in_vocab);
        uint32_t first = WORD_NONE, target = WORD_NONE;
        if (ok) {
            int kth = (int)mt_randbelow(mt.data(), &idx, (uint32_t)n_filt);
            WordleMasks m = wordle_masks(k);
            for (int i = 0, seen = 0; i < n_words; ++i)
                if (word_ok(m, words[i])) { if (seen == kth) { target = words[i]; break; } ++seen; }
            wordle_apply(k, guess, target);
            m = wordle_masks(k);
            n_filt = 0;
            for (int i = 0; i < n_words; ++i)
                if (word_ok(m, words[i])) { if (n_filt == 0) first = words[i]; ++n_filt; }
        }
        bool in_hist = false;
        if (n_filt == 1 && first != WORD_NONE) {
            in_hist = wellformed && guess == first;
            for (int i = 0; i < n_act && i < WORDLE_TRIES; ++i) in_hist = in_hist || hist[i] == first;
        }
        if (n_act < WORDLE_TRIES) hist[n_act] = wellformed ? guess : WORD_NONE;
        ++n_act;
        const float r = !(wellformed && in_vocab) ? bad_reward : ((n_filt == 1 && in_hist) ? 0.0f : -1.0f);
        const bool d = n_act == WORDLE_TRIES || r == 0.0f;
        uint8_t f[5] = {255, 255, 255, 255, 255};
        if (ok) wordle_feedback(k, guess, f);
        reward[a] = r; done[a] = d; valid[a] = ok; n_filtered[a] = n_filt; targets[a] = target;
        memcpy(fb + 5 * a, f, 5);
        if (d) return a + 1;
    }
    return n_actions;
}

// One Maze episode on raw action strings.  Returns steps taken.
int hc_maze_episode(const uint8_t* grid, int H, int W, int goal_y, int goal_x, int max_steps, const float* rewards3,
                    uint64_t seed, int opt_y, int opt_x, const uint8_t* bytes, const int32_t* lens, int n_actions,
                    int32_t* start, float* reward, uint8_t* done, int32_t* outcome, int32_t* pos) {
    std::vector<uint32_t> mt(MT_N);
    int idx;
    mt_seed(mt.data(), &idx, seed);
    (void)mt_randbelow(mt.data(), &idx, 1u);
    std::vector<int> cells;
    for (int i = 0; i < H * W; ++i) if (grid[i] == 0 && i != goal_y * W + goal_x) cells.push_back(i);
    int y, x;
    if (opt_y >= 0) { y = opt_y; x = opt_x; }
    else { int k = (int)mt_randbelow(mt.data(), &idx, (uint32_t)cells.size()); y = cells[k] / W; x = cells[k] % W; }
    start[0] = y; start[1] = x;
    MazeCfg c{H, W, max_steps, 1, rewards3[0], rewards3[1], rewards3[2]};
    int steps = 0, off = 0;
    for (int a = 0; a < n_actions; ++a) {
        const int act = maze_match_action(bytes + off, lens[a]);
        off += lens[a];
        MazeStep r = maze_step(c, grid, y, x, goal_y, goal_x, steps, act);
        y = r.y; x = r.x;
        if (r.outcome == 0) ++steps;
        reward[a] = r.reward; done[a] = (uint8_t)r.done; outcome[a] = r.outcome; pos[2 * a] = y; pos[2 * a + 1] = x;
        if (r.done) return a + 1;
    }
    return n_actions;
}

}  // extern "C"

// ---- Text-Nav (textnav_core.h): one episode on the host, for tests/test_textnav_host.py ---------------------------------
#include "textnav_core.h"
extern "C" {

int hc_textnav_start_room(uint64_t seed) {
    std::vector<uint32_t> mt(MT_N);
    int idx;
    mt_seed(mt.data(), &idx, seed);
    return tn_choice_room((int)mt_randbelow(mt.data(), &idx, 8u));
}

// commands concatenated in `bytes` with lengths `lens`; per step: applied code, state (12 ints), admissible codes (32) + count.
// states[0] / adm[0] describe the state after reset.  Returns steps taken.
int hc_textnav_episode(int start_room, const uint8_t* bytes, const int32_t* lens, int n_cmds, int32_t* applied, int32_t* states, int32_t* adm, int32_t* n_adm) {
    int32_t s[TN_STATE];
    tn_reset(s, start_room);
    memcpy(states, s, sizeof s);
    n_adm[0] = tn_admissible(s, adm);
    int off = 0;
    for (int k = 0; k < n_cmds; ++k) {
        applied[k] = tn_step(s, bytes + off, lens[k]);
        off += lens[k];
        memcpy(states + (size_t)(k + 1) * TN_STATE, s, sizeof s);
        n_adm[k + 1] = tn_admissible(s, adm + (size_t)(k + 1) * TN_MAX_ADM);
    }
    return n_cmds;
}

int hc_textnav_cmd_string(int code, uint8_t* out) { return tn_cmd_string(code >> 8, (code >> 4) & 15, code & 15, out); }

}
