// engine.h — the opaque engine behind include/lmrl_b200.h: device allocations, tables, env state, stream.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/lmrl_b200.h"
#include "env_kernels.cuh"

namespace lmrl {

extern thread_local std::string g_last_error;

inline int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return -1;
}

#define LMRL_CUDA(x)                                                                                   \
    do {                                                                                               \
        cudaError_t err__ = (x);                                                                       \
        if (err__ != cudaSuccess)                                                                      \
            return ::lmrl::fail("%s failed: %s (%s:%d)", #x, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

#define LMRL_CHECK(cond, ...)                      \
    do {                                           \
        if (!(cond)) return ::lmrl::fail(__VA_ARGS__); \
    } while (0)

struct Gpt2Dev;   // gpt2_engine.cuh

}  // namespace lmrl

struct lmrl_engine {
    lmrl_config cfg{};
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<void*> allocs;
    int64_t bytes = 0;
    int64_t launches = 0;

    // tables
    bool has_tok = false, has_wordle = false, has_maze = false, has_gpt2 = false;
    lmrl::TokenizerDev tok{};
    int32_t* d_byte_tok = nullptr; int32_t* d_tok_first = nullptr; int32_t* d_tok_space = nullptr;
    std::vector<int32_t> header_tokens;
    int32_t* d_header = nullptr;
    int max_token_bytes = 0;
    lmrl::WordleDev wordle{};
    lmrl::MazeDev maze{};
    int32_t* d_maze_obs_tokens = nullptr; int32_t* d_maze_obs_len = nullptr; int maze_obs_pitch = 0;

    // step I/O staging (device)
    uint8_t* d_act_bytes = nullptr; int32_t* d_act_len = nullptr; int act_pitch = 0;
    uint8_t* d_skip = nullptr;
    uint64_t* d_seeds = nullptr;
    int32_t* d_opt_pos = nullptr; int32_t* d_opt_goal = nullptr;
    lmrl::StepOut step{};
    int32_t* d_action_id = nullptr;

    lmrl::Gpt2Dev* gpt2 = nullptr;

    // Text-Nav (textnav_kernels.inl): [B][12] state, last applied command code, admissible command codes [B][32] + counts, start-room options
    struct TextNavDev { int32_t *state = nullptr, *applied = nullptr, *adm = nullptr, *n_adm = nullptr, *start = nullptr; bool ready = false, reset_done = false; } textnav;

    // optional per-launch timing (bench.py): CUDA events around every launch, grouped by category
    // work = work_per_row * h_counts[count_slot] (+ work2_per_row * h_counts[count_slot2]): GEMM FLOPs per live row,
    // attention bytes per cached key position (+ per row)
    struct ProfRec { int cat; int ev; double work_per_row; int count_slot; int tag; double work2_per_row; int count_slot2; };
    int profile = 0;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<ProfRec> prof;
    int32_t* h_counts = nullptr;      // pinned: [3 * max_passes] (rows, slots, sum of context lengths) per forward pass
    int n_passes = 0, max_passes = 0;
    int cur_pass_slot = -1;
    int cur_mode = 0;                 // 0 prefill pass, 1 decode pass (profiling tag)

    // One event per launch boundary: launch i runs from event i to event i+1 (the next launch's begin, or the closing event that
    // profile_read records), so a launch's time includes the gap to its successor — two events per launch cost ~6 % of a step.
    int prof_begin(int cat, double work_per_row = 0.0, int count_kind = 0, int tag = -1, double work2_per_row = 0.0, int count_kind2 = 0) {
        if (!profile) return 0;
        const int i = (int)prof.size();
        while ((int)ev_pool.size() < i + 2) {
            cudaEvent_t ev;
            if (cudaEventCreate(&ev) != cudaSuccess) return lmrl::fail("cudaEventCreate failed");
            ev_pool.push_back(ev);
        }
        prof.push_back({cat, i, work_per_row, cur_pass_slot < 0 ? -1 : 3 * cur_pass_slot + count_kind, tag, work2_per_row,
                        cur_pass_slot < 0 ? -1 : 3 * cur_pass_slot + count_kind2});
        cudaEventRecord(ev_pool[i], stream);
        return 0;
    }
    void prof_end() {}
    size_t prof_closed_at = 0;
    // closing event of the most recent launch; the next prof_begin re-records the same event, moving the boundary forward
    void prof_close() {
        if (!profile || prof.empty()) return;
        cudaEventRecord(ev_pool[prof.size()], stream);
        prof_closed_at = prof.size();
    }

    template <typename T>
    int alloc(T** p, size_t count) {
        void* q = nullptr;
        const size_t nbytes = (count ? count : 1) * sizeof(T);
        cudaError_t err = cudaMalloc(&q, nbytes);
        if (err != cudaSuccess) return lmrl::fail("cudaMalloc(%zu bytes) failed: %s", nbytes, cudaGetErrorString(err));
        cudaMemsetAsync(q, 0, nbytes, stream);
        allocs.push_back(q);
        bytes += (int64_t)nbytes;
        *p = reinterpret_cast<T*>(q);
        return 0;
    }
    template <typename T>
    int upload(T** p, const T* host, size_t count) {
        if (alloc(p, count)) return -1;
        if (count) {
            cudaError_t err = cudaMemcpyAsync(*p, host, count * sizeof(T), cudaMemcpyHostToDevice, stream);
            if (err != cudaSuccess) return lmrl::fail("H2D copy failed: %s", cudaGetErrorString(err));
            cudaStreamSynchronize(stream);   // host buffer may be transient
        }
        return 0;
    }
};
