// textnav_kernels.inl — Text-Nav environments as device-resident int32 state (C ABI part 4).
// One thread per environment: 12 int32 of state, byte-string command in, state + applied-command code + admissible-command codes out.
// Integer / latency-bound, O(B) work; the host renders text from the codes (lmrl_gym_b200/textnav.py).
#include "textnav_core.h"

namespace lmrl {

__global__ void textnav_reset_kernel(int32_t* __restrict__ state, const uint64_t* __restrict__ seeds, const int32_t* __restrict__ start_room, int n,
                                     int32_t* __restrict__ adm, int32_t* __restrict__ n_adm) {
    extern __shared__ uint32_t tn_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * ENV_WARPS_PER_BLOCK + warp;
    if (e >= n || lane != 0) return;
    int room;
    if (start_room && start_room[e] >= 0) room = start_room[e];
    else {
        uint32_t* mt = tn_smem + warp * MT_N;                         // random.Random(seed).choice(rooms): one _randbelow(8)
        int idx;
        mt_seed(mt, &idx, seeds[e]);
        room = tn_choice_room((int)mt_randbelow(mt, &idx, 8u));
    }
    int32_t* s = state + (size_t)e * TN_STATE;
    tn_reset(s, room);
    n_adm[e] = tn_admissible(s, adm + (size_t)e * TN_MAX_ADM);
}

__global__ void textnav_step_kernel(int32_t* __restrict__ state, const uint8_t* __restrict__ bytes, const int32_t* __restrict__ nbytes, int pitch,
                                    const uint8_t* __restrict__ skip, int n, int32_t* __restrict__ applied, int32_t* __restrict__ adm,
                                    int32_t* __restrict__ n_adm, float* __restrict__ reward, uint8_t* __restrict__ done) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    int32_t s[TN_STATE];
    for (int i = 0; i < TN_STATE; ++i) s[i] = state[(size_t)e * TN_STATE + i];
    int code = -3;
    if (!(skip && skip[e])) {
        code = tn_step(s, bytes + (size_t)e * pitch, min(nbytes[e], pitch));
        for (int i = 0; i < TN_STATE; ++i) state[(size_t)e * TN_STATE + i] = s[i];
    }
    applied[e] = code;
    n_adm[e] = tn_admissible(s, adm + (size_t)e * TN_MAX_ADM);
    reward[e] = (float)s[9];                                          // env.py:69: the reward of a step is the game's current score
    done[e] = (uint8_t)s[10];
}

}  // namespace lmrl

using namespace lmrl;

typedef lmrl_engine::TextNavDev TextNavDev;
static int textnav_alloc(lmrl_engine* e, TextNavDev** out) {
    TextNavDev* t = &e->textnav;
    if (!t->ready) {
        const size_t B = (size_t)e->cfg.max_envs;
        if (e->alloc(&t->state, B * TN_STATE) || e->alloc(&t->applied, B) || e->alloc(&t->adm, B * TN_MAX_ADM) || e->alloc(&t->n_adm, B) || e->alloc(&t->start, B)) return -1;
        t->ready = true;
    }
    *out = t;
    return 0;
}
static int textnav_export(lmrl_engine* e, TextNavDev* t, int n, int32_t* state, int32_t* applied, int32_t* adm, int32_t* n_adm) {
    cudaStream_t st = e->stream;
    if (state) LMRL_CUDA(cudaMemcpyAsync(state, t->state, sizeof(int32_t) * TN_STATE * n, cudaMemcpyDeviceToHost, st));
    if (applied) LMRL_CUDA(cudaMemcpyAsync(applied, t->applied, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    if (adm) LMRL_CUDA(cudaMemcpyAsync(adm, t->adm, sizeof(int32_t) * TN_MAX_ADM * n, cudaMemcpyDeviceToHost, st));
    if (n_adm) LMRL_CUDA(cudaMemcpyAsync(n_adm, t->n_adm, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    return 0;
}

extern "C" {

int lmrl_textnav_reset_host(lmrl_engine* e, const uint64_t* seeds, const int32_t* start_room, int32_t n, int32_t* state, int32_t* adm, int32_t* n_adm) {
    LMRL_CHECK(e && seeds, "lmrl_textnav_reset: null argument");
    LMRL_CHECK(n > 0 && n <= e->cfg.max_envs, "lmrl_textnav_reset: %d envs exceed engine capacity %d", n, e->cfg.max_envs);
    LMRL_CUDA(cudaSetDevice(e->device));
    TextNavDev* t;
    if (textnav_alloc(e, &t)) return -1;
    cudaStream_t st = e->stream;
    LMRL_CUDA(cudaMemcpyAsync(e->d_seeds, seeds, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, st));
    if (start_room) {
        for (int i = 0; i < n; ++i) LMRL_CHECK(start_room[i] < TN_ROOMS, "lmrl_textnav_reset: start_room[%d] = %d", i, start_room[i]);
        LMRL_CUDA(cudaMemcpyAsync(t->start, start_room, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    }
    const int blocks = (n + ENV_WARPS_PER_BLOCK - 1) / ENV_WARPS_PER_BLOCK;
    textnav_reset_kernel<<<blocks, ENV_WARPS_PER_BLOCK * 32, ENV_WARPS_PER_BLOCK * MT_N * sizeof(uint32_t), st>>>(t->state, e->d_seeds, start_room ? t->start : nullptr, n,
                                                                                                                  t->adm, t->n_adm);
    e->launches++;
    LMRL_CUDA(cudaMemsetAsync(t->applied, 0xFF, sizeof(int32_t) * n, st));
    t->reset_done = true;
    if (textnav_export(e, t, n, state, nullptr, adm, n_adm)) return -1;
    LMRL_CUDA(cudaStreamSynchronize(st));
    LMRL_CUDA(cudaGetLastError());
    return 0;
}

int lmrl_textnav_step_host(lmrl_engine* e, const uint8_t* action_bytes, const int32_t* action_len, int32_t pitch, const uint8_t* skip, int32_t n,
                           float* reward, uint8_t* done, int32_t* applied, int32_t* state, int32_t* adm, int32_t* n_adm) {
    LMRL_CHECK(e && action_bytes && action_len && reward && done, "lmrl_textnav_step: null argument");
    LMRL_CHECK(n > 0 && n <= e->cfg.max_envs, "lmrl_textnav_step: %d envs exceed engine capacity %d", n, e->cfg.max_envs);
    LMRL_CHECK(pitch > 0 && pitch <= e->act_pitch, "lmrl_textnav_step: pitch %d exceeds the action staging width %d", pitch, e->act_pitch);
    LMRL_CUDA(cudaSetDevice(e->device));
    TextNavDev* t;
    if (textnav_alloc(e, &t)) return -1;
    LMRL_CHECK(t->reset_done, "lmrl_textnav_step: reset must be called before step");
    cudaStream_t st = e->stream;
    LMRL_CUDA(cudaMemcpyAsync(e->d_act_bytes, action_bytes, (size_t)n * pitch, cudaMemcpyHostToDevice, st));
    LMRL_CUDA(cudaMemcpyAsync(e->d_act_len, action_len, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    if (skip) LMRL_CUDA(cudaMemcpyAsync(e->d_skip, skip, n, cudaMemcpyHostToDevice, st));
    textnav_step_kernel<<<(n + 127) / 128, 128, 0, st>>>(t->state, e->d_act_bytes, e->d_act_len, pitch, skip ? e->d_skip : nullptr, n, t->applied, t->adm, t->n_adm,
                                                          e->step.reward, e->step.done);
    e->launches++;
    LMRL_CUDA(cudaMemcpyAsync(reward, e->step.reward, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    LMRL_CUDA(cudaMemcpyAsync(done, e->step.done, n, cudaMemcpyDeviceToHost, st));
    if (textnav_export(e, t, n, state, applied, adm, n_adm)) return -1;
    LMRL_CUDA(cudaStreamSynchronize(st));
    LMRL_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
