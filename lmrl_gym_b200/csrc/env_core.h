// env_core.h — per-environment integer logic of the rollout path, written once for device and host.
//
// Everything here is serial per environment and free of CUDA intrinsics so that the same source is compiled
// (a) by nvcc into the sm_100a kernels of lmrl_b200.cu and (b) by g++ into csrc/libenvcore_host.so, which the
// CPU test-suite drives against the golden vectors (tests/test_envcore_host.py).  The warp-parallel parts
// (vocabulary scans) live in env_kernels.cuh and call the predicates defined here.
//
// Reference behaviour restated (file:line relative to the reference repo):
//   MT19937 / random.Random(seed) / choice      CPython stdlib, used at llm_rl_scripts/wordle/env/env.py:53,
//                                               wordle/env/game.py:179, maze/env/env.py:193-202
//   Wordle knowledge table + filter             wordle/env/game.py:17-92, 134-156
//   Wordle feedback / reward / terminal         wordle/env/game.py:213-222, 273-296
//   action normalisation / re-rendering         wordle/env/env.py:7-26
//   Maze step / reward / observation            maze/env/env.py:8-131, 161-214
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LMRL_HD __host__ __device__ __forceinline__
#else
#define LMRL_HD static inline
#endif

namespace lmrl {

// ------------------------------------------------------------------------------------------------------
// MT19937, CPython flavour.  State = 624 words + index, caller-owned (global or shared memory).
// ------------------------------------------------------------------------------------------------------
constexpr int MT_N = 624;
constexpr int MT_M = 397;

LMRL_HD void mt_seed(uint32_t* mt, int* idx, uint64_t seed) {
    // init_genrand(19650218) then init_by_array(key) with key = 32-bit little-endian words of the seed
    uint32_t key[2] = {(uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32)};
    const int klen = key[1] ? 2 : 1;
    mt[0] = 19650218u;
    for (int i = 1; i < MT_N; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    int i = 1, j = 0;
    for (int k = MT_N; k > 0; --k) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        ++i; ++j;
        if (i >= MT_N) { mt[0] = mt[MT_N - 1]; i = 1; }
        if (j >= klen) j = 0;
    }
    for (int k = MT_N - 1; k > 0; --k) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        ++i;
        if (i >= MT_N) { mt[0] = mt[MT_N - 1]; i = 1; }
    }
    mt[0] = 0x80000000u;
    *idx = MT_N;
}

LMRL_HD void mt_twist(uint32_t* mt) {
    for (int k = 0; k < MT_N; ++k) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % MT_N] & 0x7FFFFFFFu);
        mt[k] = mt[(k + MT_M) % MT_N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
    }
}

LMRL_HD uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    return y;
}

LMRL_HD uint32_t mt_u32(uint32_t* mt, int* idx) {
    if (*idx >= MT_N) { mt_twist(mt); *idx = 0; }
    return mt_temper(mt[(*idx)++]);
}

// random.Random._randbelow_with_getrandbits for 0 < n < 2^32
LMRL_HD uint32_t mt_randbelow(uint32_t* mt, int* idx, uint32_t n) {
    int k = 0;
    for (uint32_t t = n; t; t >>= 1) ++k;  // n.bit_length()
    uint32_t r = mt_u32(mt, idx) >> (32 - k);
    while (r >= n) r = mt_u32(mt, idx) >> (32 - k);
    return r;
}

LMRL_HD uint64_t mt_getrandbits64(uint32_t* mt, int* idx) {
    uint64_t lo = mt_u32(mt, idx);
    uint64_t hi = mt_u32(mt, idx);
    return lo | (hi << 32);
}

// ------------------------------------------------------------------------------------------------------
// UTF-8 -> code points with CPython's errors='replace' segmentation (Objects/stringlib/codecs.h utf8_decode
// + unicode_decode_utf8): one U+FFFD per maximal invalid prefix; a truncated-but-valid tail is ONE U+FFFD.
// This is what tokenizer.batch_decode does to arbitrary sampled byte-level tokens.
// ------------------------------------------------------------------------------------------------------
LMRL_HD bool is_cont(uint8_t b) { return (b & 0xC0) == 0x80; }

LMRL_HD int utf8_decode_replace(const uint8_t* s, int n, uint32_t* out, int cap) {
    int i = 0, m = 0;
    while (i < n && m < cap) {
        const uint8_t c = s[i];
        if (c < 0x80) { out[m++] = c; ++i; continue; }
        int need = 0;            // total length of the sequence
        if (c >= 0xC2 && c < 0xE0) need = 2;
        else if (c >= 0xE0 && c < 0xF0) need = 3;
        else if (c >= 0xF0 && c < 0xF5) need = 4;
        if (need == 0) { out[m++] = 0xFFFD; ++i; continue; }            // invalid start byte
        const int avail = n - i;
        int bad = 0;             // >0: error spanning `bad` bytes
        bool truncated = false;
        if (avail < 2) truncated = true;
        else {
            const uint8_t c2 = s[i + 1];
            bool ok2 = is_cont(c2);
            if (need == 3 && ok2) ok2 = !((c2 < 0xA0) ? (c == 0xE0) : (c == 0xED));
            if (need == 4 && ok2) ok2 = !((c2 < 0x90) ? (c == 0xF0) : (c == 0xF4));
            if (!ok2) bad = 1;
            else if (need >= 3) {
                if (avail < 3) truncated = true;
                else if (!is_cont(s[i + 2])) bad = 2;
                else if (need == 4) {
                    if (avail < 4) truncated = true;
                    else if (!is_cont(s[i + 3])) bad = 3;
                }
            }
        }
        if (bad) { out[m++] = 0xFFFD; i += bad; continue; }
        if (truncated) { out[m++] = 0xFFFD; i = n; break; }
        uint32_t cp;
        if (need == 2) cp = ((uint32_t)(c & 0x1F) << 6) | (s[i + 1] & 0x3F);
        else if (need == 3) cp = ((uint32_t)(c & 0x0F) << 12) | ((uint32_t)(s[i + 1] & 0x3F) << 6) | (s[i + 2] & 0x3F);
        else cp = ((uint32_t)(c & 0x07) << 18) | ((uint32_t)(s[i + 1] & 0x3F) << 12) | ((uint32_t)(s[i + 2] & 0x3F) << 6) | (s[i + 3] & 0x3F);
        out[m++] = cp;
        i += need;
    }
    return m;
}

LMRL_HD int utf8_encode(uint32_t cp, uint8_t* out) {
    if (cp < 0x80) { out[0] = (uint8_t)cp; return 1; }
    if (cp < 0x800) { out[0] = 0xC0 | (cp >> 6); out[1] = 0x80 | (cp & 0x3F); return 2; }
    if (cp < 0x10000) { out[0] = 0xE0 | (cp >> 12); out[1] = 0x80 | ((cp >> 6) & 0x3F); out[2] = 0x80 | (cp & 0x3F); return 3; }
    out[0] = 0xF0 | (cp >> 18); out[1] = 0x80 | ((cp >> 12) & 0x3F); out[2] = 0x80 | ((cp >> 6) & 0x3F); out[3] = 0x80 | (cp & 0x3F);
    return 4;
}

// str.isspace() per code point (what str.strip() removes, wordle/env/env.py:23)
LMRL_HD bool py_isspace(uint32_t c) {
    if (c <= 0x20) return (c >= 0x09 && c <= 0x0D) || (c >= 0x1C && c <= 0x20);
    if (c < 0x85) return false;
    return c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 ||
           c == 0x202F || c == 0x205F || c == 0x3000;
}

// text.strip().replace(' ', '') in place; returns new length
LMRL_HD int normalise_action(uint32_t* cp, int n) {
    int a = 0, b = n;
    while (a < b && py_isspace(cp[a])) ++a;
    while (b > a && py_isspace(cp[b - 1])) --b;
    int m = 0;
    for (int i = a; i < b; ++i)
        if (cp[i] != 0x20) cp[m++] = cp[i];
    return m;
}

// ------------------------------------------------------------------------------------------------------
// Wordle.  Knowledge table transposed to position-major bit sets over the alphabet:
//   here[i] bit c  <=>  letter c is HERE at position i          (CharKnowledge.HERE)
//   ban[i]  bit c  <=>  letter c is NOT_HERE at position i      (CharKnowledge.NOT_HERE)
//   neither        <=>  POSSIBLE
// A word is one uint32, letter i in bits [5i,5i+5).
// ------------------------------------------------------------------------------------------------------
constexpr uint32_t WORD_NONE = 0xFFFFFFFFu;
constexpr int WORDLE_TRIES = 6;

struct WordleKnow {
    uint32_t here[5];
    uint32_t ban[5];
};

struct WordleMasks {          // derived from WordleKnow once per filter pass
    uint32_t here[5], ban[5];
    uint32_t must;            // letters that are neither all-POSSIBLE nor all-NOT_HERE  => must occur in the word
    uint32_t absent;          // letters NOT_HERE at all five positions
};

LMRL_HD void wordle_clear(WordleKnow& k) {
    for (int i = 0; i < 5; ++i) { k.here[i] = 0; k.ban[i] = 0; }
}

LMRL_HD WordleMasks wordle_masks(const WordleKnow& k) {
    WordleMasks m;
    uint32_t any = 0, all = 0xFFFFFFFFu;
    for (int i = 0; i < 5; ++i) {
        m.here[i] = k.here[i]; m.ban[i] = k.ban[i];
        any |= k.here[i] | k.ban[i];
        all &= k.ban[i];
    }
    m.absent = all;
    m.must = any & ~all;
    return m;
}

LMRL_HD uint32_t word_presence(uint32_t w) {
    return (1u << (w & 31)) | (1u << ((w >> 5) & 31)) | (1u << ((w >> 10) & 31)) | (1u << ((w >> 15) & 31)) | (1u << ((w >> 20) & 31));
}

// game.py:53-80 for all 26 letters at once
LMRL_HD bool word_ok(const WordleMasks& m, uint32_t w) {
    uint32_t viol = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const uint32_t bit = 1u << ((w >> (5 * i)) & 31);
        viol |= (m.ban[i] & bit) | (m.here[i] & ~bit);
    }
    viol |= m.must & ~word_presence(w);
    return viol == 0;
}

// game.py:82-92 — sequential per guessed letter, duplicates included
LMRL_HD void wordle_apply(WordleKnow& k, uint32_t guess, uint32_t target) {
    const uint32_t tp = word_presence(target);
    for (int i = 0; i < 5; ++i) {
        const uint32_t c = (guess >> (5 * i)) & 31, t = (target >> (5 * i)) & 31, bit = 1u << c;
        if (c == t) { k.here[i] |= bit; k.ban[i] &= ~bit; }
        else if (tp & bit) { k.ban[i] |= bit; k.here[i] &= ~bit; }
        else for (int j = 0; j < 5; ++j) { k.ban[j] |= bit; k.here[j] &= ~bit; }
    }
}

// game.py:279-287: 0 = g, 1 = y, 2 = b, 255 = nothing emitted for this letter
LMRL_HD void wordle_feedback(const WordleKnow& k, uint32_t guess, uint8_t fb[5]) {
    uint32_t all = 0xFFFFFFFFu;
    for (int i = 0; i < 5; ++i) all &= k.ban[i];
    for (int i = 0; i < 5; ++i) {
        const uint32_t bit = 1u << ((guess >> (5 * i)) & 31);
        fb[i] = (k.here[i] & bit) ? 0 : (all & bit) ? 2 : (k.ban[i] & bit) ? 1 : 255;
    }
}

// len==5 and all a-z (game.py:214) -> packed word, else WORD_NONE
LMRL_HD uint32_t pack_action(const uint32_t* cp, int n) {
    if (n != 5) return WORD_NONE;
    uint32_t w = 0;
    for (int i = 0; i < 5; ++i) {
        if (cp[i] < 'a' || cp[i] > 'z') return WORD_NONE;
        w |= (cp[i] - 'a') << (5 * i);
    }
    return w;
}

// ------------------------------------------------------------------------------------------------------
// Token rendering of env text (canonical ids from the tokenizer tables, see tokenizer.py::device_tables)
// ------------------------------------------------------------------------------------------------------
struct TokTables {
    const int32_t* tok_first;   // [128] id of 'c' at line start, -1 if not one token
    const int32_t* tok_space;   // [128] id of ' c', -1 if not one token
    const int32_t* byte_tok;    // [256] id of the single byte
    int32_t space_id, newline_id;
};

// ' '.join(chars) + '\n'  (wordle/env/env.py:11) as token ids; returns count (never more than cap)
LMRL_HD int render_spaced_line(const uint32_t* cp, int n, const TokTables& t, int32_t* out, int cap) {
    int m = 0;
    for (int k = 0; k < n; ++k) {
        const uint32_t c = cp[k];
        int32_t one = -1;
        if (c < 128) one = k == 0 ? t.tok_first[c] : t.tok_space[c];
        if (one >= 0) { if (m < cap) out[m++] = one; continue; }
        if (k > 0 && m < cap) out[m++] = t.space_id;
        uint8_t b[4];
        const int nb = utf8_encode(c, b);
        for (int q = 0; q < nb; ++q) if (m < cap) out[m++] = t.byte_tok[b[q]];
    }
    if (m < cap) out[m++] = t.newline_id;
    return m;
}

// 'g y b b b\n' or '\n' (wordle/env/env.py:13-16)
LMRL_HD int render_feedback_line(const uint8_t fb[5], bool valid, const TokTables& t, int32_t* out, int cap) {
    int m = 0;
    if (valid) {
        const char sym[3] = {'g', 'y', 'b'};
        int k = 0;
        for (int i = 0; i < 5; ++i) {
            if (fb[i] > 2) continue;
            const int c = sym[fb[i]];
            if (m < cap) out[m++] = k == 0 ? t.tok_first[c] : t.tok_space[c];
            ++k;
        }
    }
    if (m < cap) out[m++] = t.newline_id;
    return m;
}

// ------------------------------------------------------------------------------------------------------
// Maze (maze/env/env.py).  Grid is row-major uint8 (1 = wall), at most 32x32.
// ------------------------------------------------------------------------------------------------------
struct MazeCfg {
    int H, W;
    int max_steps;        // < 0: no limit
    int last_k;
    float r_goal, r_unknown, r_other;   // env.py:109-131
};

// action ids: 0 left, 1 right, 2 up, 3 down (env.py:94-99), -1 unknown string
LMRL_HD int maze_match_action(const uint8_t* s, int n) {
    const char* names[4] = {"move left\n", "move right\n", "move up\n", "move down\n"};
    for (int a = 0; a < 4; ++a) {
        int L = 0;
        while (names[a][L]) ++L;
        if (L != n) continue;
        bool eq = true;
        for (int i = 0; i < L; ++i) eq = eq && (s[i] == (uint8_t)names[a][i]);
        if (eq) return a;
    }
    return -1;
}

struct MazeStep { int y, x; float reward; int done; int outcome; };  // outcome 0 running, 1 success, 2 failure
                                                                      // (num_steps is advanced by the caller)

// env.py:161-184 without the text; num_steps is the value *before* the call
LMRL_HD MazeStep maze_step(const MazeCfg& c, const uint8_t* grid, int y, int x, int gy, int gx, int num_steps, int action) {
    MazeStep r;
    r.y = y; r.x = x;
    if (c.max_steps >= 0 && num_steps >= c.max_steps) { r.reward = -1.0f; r.done = 1; r.outcome = 2; return r; }
    if (action >= 0) {
        const int dy = action == 2 ? -1 : action == 3 ? 1 : 0;
        const int dx = action == 0 ? -1 : action == 1 ? 1 : 0;
        if (grid[(y + dy) * c.W + (x + dx)] == 0) { r.y = y + dy; r.x = x + dx; }
    }
    const bool goal = (r.y == gy && r.x == gx);
    r.reward = goal ? c.r_goal : (action < 0 ? c.r_unknown : c.r_other);
    r.done = goal ? 1 : 0;
    r.outcome = goal ? 1 : 0;
    return r;
}

}  // namespace lmrl
