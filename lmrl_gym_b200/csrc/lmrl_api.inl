// lmrl_api.cu — C-ABI entry points (include/lmrl_b200.h): lifecycle, tables, text-path environments.
// The policy / fused-rollout entry points are in gpt2_engine.cu.



namespace lmrl {
thread_local std::string g_last_error;
}
using namespace lmrl;

extern "C" {

const char* lmrl_last_error(void) { return g_last_error.c_str(); }
int lmrl_abi_version(void) { return 2; }

int lmrl_create(const lmrl_config* cfg, lmrl_engine** out) {
    LMRL_CHECK(cfg && out, "lmrl_create: null argument");
    LMRL_CHECK(cfg->max_envs > 0, "lmrl_create: max_envs must be positive");
    int ndev = 0;
    LMRL_CUDA(cudaGetDeviceCount(&ndev));
    LMRL_CHECK(cfg->device >= 0 && cfg->device < ndev, "lmrl_create: device %d not present (%d devices)", cfg->device, ndev);
    LMRL_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    LMRL_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    LMRL_CHECK(prop.major == 10, "lmrl_create: this library is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
    lmrl_engine* e = new lmrl_engine();
    e->cfg = *cfg;
    e->device = cfg->device;
    LMRL_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    if (const char* t = getenv("LMRL_MBAR_TIMEOUT_S")) {            // debugging aid: mbarrier wait bound of the tcgen05 kernels (default 2 s)
        const unsigned long long ns = (unsigned long long)(atof(t) * 1e9);
        if (ns > 0) LMRL_CUDA(cudaMemcpyToSymbol(lmrl::g_mbar_timeout_ns, &ns, sizeof ns));
    }
    const int B = cfg->max_envs;
    e->act_pitch = MAX_ACTION_CP;
    if (e->alloc(&e->d_act_bytes, (size_t)B * e->act_pitch) || e->alloc(&e->d_act_len, B) || e->alloc(&e->d_skip, B) ||
        e->alloc(&e->d_seeds, B) || e->alloc(&e->d_opt_pos, (size_t)B * 2) || e->alloc(&e->d_opt_goal, (size_t)B * 2) ||
        e->alloc(&e->step.reward, B) || e->alloc(&e->step.done, B) || e->alloc(&e->step.valid, B) ||
        e->alloc(&e->step.fb, (size_t)B * 5) || e->alloc(&e->step.aux, B) || e->alloc(&e->d_action_id, B)) {
        lmrl_destroy(e);
        return -1;
    }
    *out = e;
    return 0;
}

int lmrl_destroy(lmrl_engine* e) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);

    lmrl_gpt2_free(e);
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    if (e->h_counts) cudaFreeHost(e->h_counts);
    for (void* p : e->allocs) cudaFree(p);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return 0;
}

int lmrl_sync(lmrl_engine* e) {
    LMRL_CHECK(e, "null engine");
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

int64_t lmrl_device_bytes(lmrl_engine* e) { return e ? e->bytes : 0; }
void* lmrl_stream(lmrl_engine* e) { return e ? (void*)e->stream : nullptr; }
int64_t lmrl_launch_count(lmrl_engine* e, int32_t reset) {
    if (!e) return 0;
    const int64_t v = e->launches;
    if (reset) e->launches = 0;
    return v;
}

int lmrl_load_tokenizer(lmrl_engine* e, const uint8_t* token_bytes, const int32_t* offsets, int32_t n_tokens,
                        const int32_t* byte_tok, const int32_t* tok_first, const int32_t* tok_space, int32_t space_id,
                        int32_t newline_id, const int32_t* header_tokens, int32_t n_header) {
    LMRL_CHECK(e && token_bytes && offsets && byte_tok && tok_first && tok_space, "lmrl_load_tokenizer: null argument");
    LMRL_CUDA(cudaSetDevice(e->device));
    uint8_t* d_bytes; int32_t* d_off;
    if (e->upload(&d_bytes, token_bytes, (size_t)offsets[n_tokens]) || e->upload(&d_off, offsets, (size_t)n_tokens + 1) ||
        e->upload(&e->d_byte_tok, byte_tok, 256) || e->upload(&e->d_tok_first, tok_first, 128) ||
        e->upload(&e->d_tok_space, tok_space, 128))
        return -1;
    e->max_token_bytes = 0;
    for (int i = 0; i < n_tokens; ++i) e->max_token_bytes = std::max(e->max_token_bytes, offsets[i + 1] - offsets[i]);
    e->tok.token_bytes = d_bytes; e->tok.offsets = d_off; e->tok.n_tokens = n_tokens;
    e->tok.t.byte_tok = e->d_byte_tok; e->tok.t.tok_first = e->d_tok_first; e->tok.t.tok_space = e->d_tok_space;
    e->tok.t.space_id = space_id; e->tok.t.newline_id = newline_id;
    e->header_tokens.assign(header_tokens, header_tokens + n_header);
    if (e->upload(&e->d_header, header_tokens, (size_t)n_header)) return -1;
    e->has_tok = true;
    return 0;
}

int lmrl_load_wordle(lmrl_engine* e, const uint32_t* packed_words, int32_t n_words, int32_t require, float bad_word_reward) {
    LMRL_CHECK(e && packed_words && n_words > 0, "lmrl_load_wordle: bad argument");
    LMRL_CUDA(cudaSetDevice(e->device));
    const int B = e->cfg.max_envs;
    uint32_t* d_words;
    if (e->upload(&d_words, packed_words, (size_t)n_words)) return -1;
    WordleDev& w = e->wordle;
    w.words = d_words; w.n_words = n_words; w.require_in_vocab = require; w.bad_word_reward = bad_word_reward;
    if (!e->has_wordle) {
        if (e->alloc(&w.know, (size_t)B * 10) || e->alloc(&w.mt, (size_t)B * MT_N) || e->alloc(&w.mt_idx, B) ||
            e->alloc(&w.n_actions, B) || e->alloc(&w.hist, (size_t)B * WORDLE_TRIES) || e->alloc(&w.n_filtered, B) ||
            e->alloc(&w.done, B) || e->alloc(&w.target, B))
            return -1;
    }
    e->has_wordle = true;
    return 0;
}

int lmrl_load_maze(lmrl_engine* e, const uint8_t* grid, int32_t H, int32_t W, int32_t goal_y, int32_t goal_x,
                   int32_t max_steps, int32_t last_k, const float* rewards3, const int32_t* obs_tokens,
                   const int32_t* obs_len, int32_t obs_pitch, const int32_t* act_tokens, const int32_t* act_len) {
    LMRL_CHECK(e && grid && rewards3 && H > 2 && W > 2 && H <= 64 && W <= 64, "lmrl_load_maze: bad argument");
    LMRL_CUDA(cudaSetDevice(e->device));
    const int B = e->cfg.max_envs;
    MazeDev& m = e->maze;
    uint8_t* d_grid;
    if (e->upload(&d_grid, grid, (size_t)H * W)) return -1;
    std::vector<int32_t> free_cells;
    for (int i = 0; i < H * W; ++i) if (grid[i] == 0) free_cells.push_back(i);
    int32_t* d_free;
    if (e->upload(&d_free, free_cells.data(), free_cells.size())) return -1;
    m.cfg.H = H; m.cfg.W = W; m.cfg.max_steps = max_steps; m.cfg.last_k = last_k;
    m.cfg.r_goal = rewards3[0]; m.cfg.r_unknown = rewards3[1]; m.cfg.r_other = rewards3[2];
    m.grid = d_grid; m.free_cells = d_free; m.n_free = (int)free_cells.size(); m.goal_y = goal_y; m.goal_x = goal_x;
    if (!e->has_maze) {
        if (e->alloc(&m.pos, (size_t)B * 2) || e->alloc(&m.goal, (size_t)B * 2) || e->alloc(&m.num_steps, B) || e->alloc(&m.done, B))
            return -1;
    }
    if (!e->has_maze) {
        if (e->alloc(&m.item_len, (size_t)B * MAZE_RING) || e->alloc(&m.n_items, B) || e->alloc(&m.cut, B)) return -1;
    }
    m.obs_tokens = nullptr; m.obs_len = nullptr; m.act_tokens = nullptr; m.act_len = nullptr;
    if (obs_tokens && obs_len && act_tokens && act_len) {
        LMRL_CHECK(last_k >= 1 && last_k <= 64, "lmrl_load_maze: the fused loop supports 1 <= last_k <= 64 (got %d)", last_k);
        for (int a = 0; a < 4; ++a) LMRL_CHECK(act_len[a] > 0 && act_len[a] <= MAZE_ACT_PITCH, "lmrl_load_maze: action %d has %d tokens (max %d)", a, act_len[a], MAZE_ACT_PITCH);
        int32_t *d_at, *d_al;
        if (e->upload(&e->d_maze_obs_tokens, obs_tokens, (size_t)H * W * obs_pitch) || e->upload(&e->d_maze_obs_len, obs_len, (size_t)H * W) ||
            e->upload(&d_at, act_tokens, (size_t)4 * MAZE_ACT_PITCH) || e->upload(&d_al, act_len, 4))
            return -1;
        e->maze_obs_pitch = obs_pitch;
        m.obs_tokens = e->d_maze_obs_tokens; m.obs_len = e->d_maze_obs_len; m.obs_pitch = obs_pitch;
        m.act_tokens = d_at; m.act_len = d_al;
    }
    e->has_maze = true;
    return 0;
}

static int stage_actions(lmrl_engine* e, const uint8_t* action_bytes, const int32_t* action_len, int32_t pitch,
                         const uint8_t* skip, int32_t n) {
    LMRL_CHECK(n > 0 && n <= e->cfg.max_envs, "batch of %d envs exceeds engine capacity %d", n, e->cfg.max_envs);
    LMRL_CHECK(pitch > 0 && pitch <= e->act_pitch, "action pitch %d exceeds %d bytes", pitch, e->act_pitch);
    LMRL_CUDA(cudaMemcpy2DAsync(e->d_act_bytes, e->act_pitch, action_bytes, pitch, pitch, n, cudaMemcpyHostToDevice, e->stream));
    LMRL_CUDA(cudaMemcpyAsync(e->d_act_len, action_len, sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream));
    if (skip) LMRL_CUDA(cudaMemcpyAsync(e->d_skip, skip, n, cudaMemcpyHostToDevice, e->stream));
    else LMRL_CUDA(cudaMemsetAsync(e->d_skip, 0, n, e->stream));
    return 0;
}

int lmrl_wordle_reset_host(lmrl_engine* e, const uint64_t* seeds, int32_t n) {
    LMRL_CHECK(e && e->has_wordle, "lmrl_wordle_reset: no vocabulary loaded");
    LMRL_CHECK(seeds && n > 0 && n <= e->cfg.max_envs, "lmrl_wordle_reset: bad batch size %d", n);
    LMRL_CUDA(cudaSetDevice(e->device));
    LMRL_CUDA(cudaMemcpyAsync(e->d_seeds, seeds, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, e->stream));
    const int blocks = (n + ENV_WARPS_PER_BLOCK - 1) / ENV_WARPS_PER_BLOCK;
    wordle_reset_kernel<<<blocks, ENV_WARPS_PER_BLOCK * 32, ENV_WARPS_PER_BLOCK * MT_N * sizeof(uint32_t), e->stream>>>(e->wordle, e->d_seeds, n);
    e->launches++;
    LMRL_CUDA(cudaGetLastError());
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

int lmrl_wordle_step_host(lmrl_engine* e, const uint8_t* action_bytes, const int32_t* action_len, int32_t pitch,
                          const uint8_t* skip, int32_t n, float* reward, uint8_t* done, uint8_t* valid, uint8_t* feedback5,
                          int32_t* n_filtered) {
    LMRL_CHECK(e && e->has_wordle, "lmrl_wordle_step: no vocabulary loaded");
    LMRL_CUDA(cudaSetDevice(e->device));
    if (stage_actions(e, action_bytes, action_len, pitch, skip, n)) return -1;
    ActionSrc src{};
    src.bytes = e->d_act_bytes; src.nbytes = e->d_act_len; src.pitch = e->act_pitch;
    StreamDev st{};   // text path: no token stream
    const int blocks = (n + ENV_WARPS_PER_BLOCK - 1) / ENV_WARPS_PER_BLOCK;
    wordle_step_kernel<<<blocks, ENV_WARPS_PER_BLOCK * 32, ENV_WARPS_PER_BLOCK * sizeof(WordleScratch), e->stream>>>(
        e->wordle, e->tok, src, e->d_skip, e->step, st, n);
    e->launches++;
    LMRL_CUDA(cudaGetLastError());
    if (reward) LMRL_CUDA(cudaMemcpyAsync(reward, e->step.reward, sizeof(float) * n, cudaMemcpyDeviceToHost, e->stream));
    if (done) LMRL_CUDA(cudaMemcpyAsync(done, e->wordle.done, n, cudaMemcpyDeviceToHost, e->stream));
    if (valid) LMRL_CUDA(cudaMemcpyAsync(valid, e->step.valid, n, cudaMemcpyDeviceToHost, e->stream));
    if (feedback5) LMRL_CUDA(cudaMemcpyAsync(feedback5, e->step.fb, (size_t)n * 5, cudaMemcpyDeviceToHost, e->stream));
    if (n_filtered) LMRL_CUDA(cudaMemcpyAsync(n_filtered, e->wordle.n_filtered, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, e->stream));
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

int lmrl_wordle_peek_host(lmrl_engine* e, int32_t n, uint32_t* know10, uint32_t* target, int32_t* n_actions) {
    LMRL_CHECK(e && e->has_wordle && n > 0 && n <= e->cfg.max_envs, "lmrl_wordle_peek: bad argument");
    LMRL_CUDA(cudaSetDevice(e->device));
    if (know10) LMRL_CUDA(cudaMemcpyAsync(know10, e->wordle.know, sizeof(uint32_t) * 10 * n, cudaMemcpyDeviceToHost, e->stream));
    if (target) LMRL_CUDA(cudaMemcpyAsync(target, e->wordle.target, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, e->stream));
    if (n_actions) LMRL_CUDA(cudaMemcpyAsync(n_actions, e->wordle.n_actions, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, e->stream));
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

int lmrl_maze_reset_host(lmrl_engine* e, const uint64_t* seeds, const int32_t* opt_pos, const int32_t* opt_goal, int32_t n,
                         int32_t* pos_out, int32_t* goal_out) {
    LMRL_CHECK(e && e->has_maze, "lmrl_maze_reset: no maze loaded");
    LMRL_CHECK(seeds && n > 0 && n <= e->cfg.max_envs, "lmrl_maze_reset: bad batch size %d", n);
    LMRL_CUDA(cudaSetDevice(e->device));
    LMRL_CUDA(cudaMemcpyAsync(e->d_seeds, seeds, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, e->stream));
    if (opt_pos) LMRL_CUDA(cudaMemcpyAsync(e->d_opt_pos, opt_pos, sizeof(int32_t) * 2 * n, cudaMemcpyHostToDevice, e->stream));
    if (opt_goal) LMRL_CUDA(cudaMemcpyAsync(e->d_opt_goal, opt_goal, sizeof(int32_t) * 2 * n, cudaMemcpyHostToDevice, e->stream));
    const int blocks = (n + ENV_WARPS_PER_BLOCK - 1) / ENV_WARPS_PER_BLOCK;
    maze_reset_kernel<<<blocks, ENV_WARPS_PER_BLOCK * 32, ENV_WARPS_PER_BLOCK * MT_N * sizeof(uint32_t), e->stream>>>(
        e->maze, e->d_seeds, opt_pos ? e->d_opt_pos : nullptr, opt_goal ? e->d_opt_goal : nullptr, n);
    e->launches++;
    LMRL_CUDA(cudaGetLastError());
    if (pos_out) LMRL_CUDA(cudaMemcpyAsync(pos_out, e->maze.pos, sizeof(int32_t) * 2 * n, cudaMemcpyDeviceToHost, e->stream));
    if (goal_out) LMRL_CUDA(cudaMemcpyAsync(goal_out, e->maze.goal, sizeof(int32_t) * 2 * n, cudaMemcpyDeviceToHost, e->stream));
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

int lmrl_maze_step_host(lmrl_engine* e, const uint8_t* action_bytes, const int32_t* action_len, int32_t pitch,
                        const uint8_t* skip, int32_t n, float* reward, uint8_t* done, uint8_t* known, int32_t* outcome,
                        int32_t* pos_out) {
    LMRL_CHECK(e && e->has_maze, "lmrl_maze_step: no maze loaded");
    LMRL_CUDA(cudaSetDevice(e->device));
    if (stage_actions(e, action_bytes, action_len, pitch, skip, n)) return -1;
    ActionSrc src{};
    src.bytes = e->d_act_bytes; src.nbytes = e->d_act_len; src.pitch = e->act_pitch;
    maze_step_kernel<<<(n + 127) / 128, 128, 0, e->stream>>>(e->maze, src, e->d_skip, e->step, e->d_action_id, n);
    e->launches++;
    LMRL_CUDA(cudaGetLastError());
    if (reward) LMRL_CUDA(cudaMemcpyAsync(reward, e->step.reward, sizeof(float) * n, cudaMemcpyDeviceToHost, e->stream));
    if (done) LMRL_CUDA(cudaMemcpyAsync(done, e->maze.done, n, cudaMemcpyDeviceToHost, e->stream));
    if (known) LMRL_CUDA(cudaMemcpyAsync(known, e->step.valid, n, cudaMemcpyDeviceToHost, e->stream));
    if (outcome) LMRL_CUDA(cudaMemcpyAsync(outcome, e->step.aux, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, e->stream));
    if (pos_out) LMRL_CUDA(cudaMemcpyAsync(pos_out, e->maze.pos, sizeof(int32_t) * 2 * n, cudaMemcpyDeviceToHost, e->stream));
    LMRL_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

}  // extern "C"
