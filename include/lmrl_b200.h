/* lmrl_b200.h — C ABI of the B200 batched-rollout engine for LMRL-Gym.
 *
 * The reference has no FFI on this path: the boundary is a pair of Python ABCs in LLM_RL/environment.py
 * (BatchedTextEnv.reset/step :56-69, BatchedTextPolicy.act :120-123) driven by interact_environment (:154-207) and
 * text_env_eval (:211-267).  Each entry point below names the reference interface it replaces; INTEGRATION.md shows
 * the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C: opaque handle, pointers + sizes, int status (0 = ok, negative = error; text via lmrl_last_error()).
 *   - functions ending in _host take HOST pointers, do their own H2D/D2H copies on the engine stream and return
 *     after the results are in the caller's buffers.  Functions ending in _dev take DEVICE pointers, enqueue on the
 *     engine stream and return without synchronising.
 *   - one engine = one GPU (one process per GPU; rollouts shard over engines with no collective).
 *   - no exceptions / longjmp cross the ABI; the library never falls back to a CPU path.
 */
#ifndef LMRL_B200_H
#define LMRL_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lmrl_engine lmrl_engine;

#define LMRL_ENV_NONE   0
#define LMRL_ENV_WORDLE 1
#define LMRL_ENV_MAZE   2

/* maze observation renderers (llm_rl_scripts/maze/env/env.py:15-81) */
#define LMRL_MAZE_DESCRIBE_GOAL_WALLS      0  /* describe_observation               */
#define LMRL_MAZE_DESCRIBE_GOAL_POS_WALLS  1  /* describe_observation_give_position */
#define LMRL_MAZE_DESCRIBE_WALLS           2  /* describe_observation_only_walls    */

typedef struct lmrl_config {
    int32_t device;            /* CUDA ordinal */
    int32_t max_envs;          /* B: concurrent environments on this GPU */
    int32_t max_input_length;  /* BlockingStrategy max_length (LEFT pad / LEFT truncate), eval_bc_gpt2.py:100-104 */
    int32_t max_new_tokens;    /* GenerationConfig.max_new_tokens, eval_bc_gpt2.py:90-95 */
    int32_t max_turns;         /* upper bound on env steps per episode (6 Wordle, max_steps+1 Maze) */
    int32_t kv_pages;          /* pages (32 positions each, per layer and tower) of the paged KV cache; 0 = max_envs * ceil((max_input_length +
                                  max_new_tokens) / 32): every env at full length.  A smaller pool fits more envs per GPU when episodes stay
                                  short; a batch that runs out of pages fails with an error (lmrl_kv_stats reports the use) */
    int32_t reserved[2];
} lmrl_config;

typedef struct lmrl_gpt2_dims {
    int32_t n_layer, n_head, d_model, d_ff, n_vocab /* rows of wte actually sampled from (unpadded) */, n_pos;
    float ln_eps;
    int32_t reserved;
} lmrl_gpt2_dims;

/* fp32 weights in HF GPT-2 layout (Conv1D: [in, out]); converted to bf16 device layout at load.  The arrays may live in host OR
 * device memory (copied with cudaMemcpyDefault): a trainer that keeps fp32 master weights on the GPU hands its pointers over
 * directly (the caller orders its own stream before the call). */
typedef struct lmrl_gpt2_weights {
    const float* wte;            /* [n_vocab_rows, d]   (n_vocab_rows >= n_vocab; extra rows = padding ids) */
    int32_t n_vocab_rows;
    int32_t reserved;
    const float* wpe;            /* [n_pos, d] */
    const float* const* ln1_g;   /* [L] -> [d] */
    const float* const* ln1_b;
    const float* const* attn_w;  /* [L] -> [d, 3d]  c_attn.weight */
    const float* const* attn_b;  /* [L] -> [3d] */
    const float* const* proj_w;  /* [L] -> [d, d]   attn c_proj.weight */
    const float* const* proj_b;
    const float* const* ln2_g;
    const float* const* ln2_b;
    const float* const* fc_w;    /* [L] -> [d, d_ff] */
    const float* const* fc_b;
    const float* const* fc2_w;   /* [L] -> [d_ff, d] */
    const float* const* fc2_b;
    const float* lnf_g;
    const float* lnf_b;
} lmrl_gpt2_weights;

typedef struct lmrl_sampling {
    float temperature;     /* GenerationConfig.temperature (None -> 1.0) */
    int32_t do_sample;     /* 0 = greedy argmax */
    uint64_t seed;         /* base of the counter-based sampling RNG, keyed by (seed, episode seed, turn, token) */
    int32_t top_k;         /* GenerationConfig.top_k (None / 0 = off): keep the k highest scores (FlaxTopKLogitsWarper) */
    float top_p;           /* GenerationConfig.top_p (None / 0 / >= 1 = off): nucleus filtering (FlaxTopPLogitsWarper), after top_k */
} lmrl_sampling;

const char* lmrl_last_error(void);
int lmrl_abi_version(void);

/* ---- lifecycle ------------------------------------------------------------------------------------------------ */
int lmrl_create(const lmrl_config* cfg, lmrl_engine** out);
int lmrl_destroy(lmrl_engine* e);
int lmrl_sync(lmrl_engine* e);
int64_t lmrl_device_bytes(lmrl_engine* e);          /* HBM held by this engine */
void* lmrl_stream(lmrl_engine* e);                  /* cudaStream_t the engine launches on */

/* ---- tables --------------------------------------------------------------------------------------------------- */
/* replaces tokenizer.batch_decode / tokenizer.encode of value_rl_base/base_interface.py:112,126 for the closed set of
 * env strings: id -> bytes for every id, plus single-token ids of 'c' / ' c' for ASCII c and of each single byte */
int lmrl_load_tokenizer(lmrl_engine* e, const uint8_t* token_bytes, const int32_t* offsets, int32_t n_tokens,
                        const int32_t* byte_tok /*[256]*/, const int32_t* tok_first /*[128]*/,
                        const int32_t* tok_space /*[128]*/, int32_t space_id, int32_t newline_id,
                        const int32_t* header_tokens, int32_t n_header);

/* replaces Vocabulary.from_file + WordleEnvironment.__init__ (wordle/env/game.py:164-170, env.py:40-44) */
int lmrl_load_wordle(lmrl_engine* e, const uint32_t* packed_words, int32_t n_words, int32_t require_words_in_vocab,
                     float bad_word_reward);

/* replaces setup_maze_env / MazeEnv.__init__ (maze/env/maze_utils.py:9-52, env.py:134-159).
 * rewards = {at goal, unknown action, otherwise} (env.py:109-131).  Fused rollout only (all four may be NULL for the
 * text path): obs_tokens/obs_len = token ids of the observation sentence of every cell y*W+x for the default goal;
 * act_tokens [4][8] / act_len [4] = canonical ids of 'move left\n', 'move right\n', 'move up\n', 'move down\n'. */
int lmrl_load_maze(lmrl_engine* e, const uint8_t* grid, int32_t H, int32_t W, int32_t goal_y, int32_t goal_x,
                   int32_t max_steps, int32_t last_k, const float* rewards3,
                   const int32_t* obs_tokens, const int32_t* obs_len, int32_t obs_pitch,
                   const int32_t* act_tokens, const int32_t* act_len);

/* replaces load_params / GPT2PPOPolicy.set_params (ppo/gpt2/interface.py:548-549) */
int lmrl_load_gpt2(lmrl_engine* e, const lmrl_gpt2_dims* dims, const lmrl_gpt2_weights* w);

/* ---- ILQL / MC value-perturbed policy (GPT2ValuePolicy, value_rl_base/gpt2/interface.py:281-320) ----------------------
 * params tuple of the reference = (pi_beta, base, q1_head, q2_head-or-None) (value_rl_base/gpt2/interface.py:101):
 *   tower 0 = pi_beta, tower 1 = the value base (same architecture, own KV cache), then the Q heads.
 * Every sampled token uses  logits = pi_beta_logits + beta * min(q1(h), q2(h)),  h = last hidden state (after ln_f) of the
 * value base (generation.py:87-119);  q_k(h) = dense2_k(relu(dense1_k(h)))  (heads/mlp_head.py:139-148). */
int lmrl_load_gpt2_tower(lmrl_engine* e, int32_t tower, const lmrl_gpt2_dims* dims, const lmrl_gpt2_weights* w);
typedef struct lmrl_mlp_head {
    const float* w1;   /* [d_model, hidden]  dense1.kernel */
    const float* b1;   /* [hidden] */
    const float* w2;   /* [hidden, n_out]    dense2.kernel (n_out >= n_vocab; extra columns are ignored) */
    const float* b2;   /* [n_out] */
    int32_t hidden;
    int32_t n_out;
} lmrl_mlp_head;
/* q2 may be NULL (MC returns: one head, wordle/mc/eval_mc_gpt2.py:103-116).  Switches the engine to the value-perturbed policy.
 * A head with w1 == NULL is a LinearHead (heads/linear_head.py:112-119, the Q head of the MC-returns drivers,
 * wordle/mc/train_mc_returns_gpt2.py:198-206): q(h) = h . w2 + b2 with w2 [d_model, n_out], hidden = d_model, b1 ignored, b2 may be
 * NULL (use_bias = False).  q1 and q2 must be of the same kind. */
int lmrl_load_q_heads(lmrl_engine* e, const lmrl_mlp_head* q1, const lmrl_mlp_head* q2, float beta);
/* 0 = plain LM policy (logits of tower 0), 1 = value-perturbed policy (pi_beta + beta * min q),
 * 2 = the pi_beta = None variant (logits = beta * min q, generation.py:118-119: only the value tower runs) */
int lmrl_set_policy_kind(lmrl_engine* e, int32_t kind);

/* ---- environments, text path (BatchedTextEnv.reset / .step, environment.py:56-69) ------------------------------- */
/* seeds: one uint64 per env (abs(seed) of random.Random(seed), wordle/env/env.py:53) */
int lmrl_wordle_reset_host(lmrl_engine* e, const uint64_t* seeds, int32_t n);
/* action text as UTF-8 bytes, row i at action_bytes + i*pitch; skip[i] != 0 leaves env i untouched (done rows).
 * outputs per env: reward, done, valid (legal guess => the reference reward is a Python int), feedback codes
 * (0 g, 1 y, 2 b, 255 none), size of the filtered vocabulary after the step */
int lmrl_wordle_step_host(lmrl_engine* e, const uint8_t* action_bytes, const int32_t* action_len, int32_t pitch,
                          const uint8_t* skip, int32_t n, float* reward, uint8_t* done, uint8_t* valid,
                          uint8_t* feedback5, int32_t* n_filtered);
/* opt_pos / opt_goal: [n][2] (y,x) or -1,-1 when the option is absent (maze/env/env.py:190-200); may be NULL */
int lmrl_maze_reset_host(lmrl_engine* e, const uint64_t* seeds, const int32_t* opt_pos, const int32_t* opt_goal,
                         int32_t n, int32_t* pos_out /*[n][2]*/, int32_t* goal_out /*[n][2]*/);
int lmrl_maze_step_host(lmrl_engine* e, const uint8_t* action_bytes, const int32_t* action_len, int32_t pitch,
                        const uint8_t* skip, int32_t n, float* reward, uint8_t* done, uint8_t* known,
                        int32_t* outcome /*0 running, 1 Success, 2 Failure*/, int32_t* pos_out /*[n][2]*/);
/* test hooks: raw env state */
int lmrl_wordle_peek_host(lmrl_engine* e, int32_t n, uint32_t* know10, uint32_t* target, int32_t* n_actions);

/* ---- policy, token path (BatchedTextPolicy.act -> generate_from_str, ppo/gpt2/interface.py:507-546) ------------ */
/* prompts: LEFT-truncated to max_input_length by the callee; row i = tokens + i*pitch, n_tok[i] ids (0 = inactive row).
 * out_tokens [n][max_new_tokens], out_len[n] (generation stops after the newline/EOS id or at max_new_tokens).
 * episode_keys: one uint64 per row feeding the counter-based sampler (NULL = row index). */
int lmrl_generate_host(lmrl_engine* e, const int32_t* tokens, const int32_t* n_tok, int32_t pitch, int32_t n,
                       const uint64_t* episode_keys, int32_t turn, const lmrl_sampling* s,
                       int32_t* out_tokens, int32_t* out_len);
/* parity hook: teacher-forced forward of one ragged batch, fp32 logits of the LAST position of every row */
int lmrl_logits_host(lmrl_engine* e, const int32_t* tokens, const int32_t* n_tok, int32_t pitch, int32_t n,
                     float* logits /*[n][n_vocab]*/);

/* teacher-forced scoring (build_bc_score_fn / build_ppo_score_fn, ppo/score_fn.py:69-126): for every sequence the sum over
 * t >= prefix_len of log softmax(logits_{t-1} / temperature)[token_t] under the loaded policy (plain LM or value-perturbed);
 * ids >= n_vocab are masked to -inf (ppo/gpt2/interface.py:330).  Sequences longer than max_input_length are rejected.
 * token_logprobs (optional): the individual terms, sequence by sequence, in token order. */
int lmrl_score_host(lmrl_engine* e, const int32_t* tokens, const int32_t* n_tok, int32_t pitch, int32_t n,
                    const int32_t* prefix_len, float temperature, float* logprob_sum /*[n]*/, float* token_logprobs);

/* ---- teacher-forced per-position forward (SURVEY.md §8 f1 advantage scoring, f2 PPO data path) ---------------------
 * Replaces, for one tower and one head set, ValueRLInference.forward (value_rl_base/base_interface.py:129-168: q1 / q2 / v of
 * every position), ILQLInference.forward (ilql/base_interface.py:313-344: value_inference + target_value_inference),
 * GPT2Inference.get_logits_from_tokens + log_softmax + take_along_axis (ilql/gpt2/score_fn.py:60-64) and PPOInference.forward
 * + token_logprobs_from_logits (ppo/base_interface.py:404-411, 519-529) — without materialising [rows, vocab] logits or Q values.
 *
 * Towers: 0 = policy / pi_beta, 1 = value base (ILQL / MC) or the PPO initial policy, 2 = target value base (forward only).
 * Head sets: 0 = online heads (q1, q2, v), 1 = target heads (q1_target, q2_target).  A head set is independent of the
 * beta-scaled Q heads lmrl_load_q_heads installs for sampling. */
typedef struct lmrl_value_heads {
    const lmrl_mlp_head* q1;      /* MLPHead d -> hidden -> n_out (>= n_vocab), or a LinearHead (w1 == NULL: kernel in w2, bias in b2 or NULL), or NULL */
    const lmrl_mlp_head* q2;      /* or NULL */
    const lmrl_mlp_head* v;       /* MLPHead d -> hidden -> 1 (ILQL V head), or NULL */
    const float* v_linear_w;      /* LinearHead kernel [d_model] (output_dim 1; the PPO value head), or NULL */
    const float* v_linear_b;      /* [1], or NULL when use_bias = False */
} lmrl_value_heads;
int lmrl_load_value_heads(lmrl_engine* e, int32_t head_set, const lmrl_value_heads* h);

typedef struct lmrl_forward_out {  /* HOST pointers, each [n][pitch] fp32, any may be NULL; unwritten cells are set to 0 */
    float* token_logprobs;  /* [i][t] = log softmax(lm_logits_t / temperature)[tokens[i][t+1]]   for t in [0, n_tok[i] - 1) */
    float* v;               /* [i][t] = V head of position t                                      for t in [0, n_tok[i]) */
    float* q1;              /* [i][t] = Q1(h_t)[tokens[i][t+1]]                                   for t in [0, n_tok[i] - 1) */
    float* q2;              /* [i][t] = Q2(h_t)[tokens[i][t+1]] */
} lmrl_forward_out;
/* Sequences start at position 0 (RIGHT padding of the reference: pad cells simply do not exist here); n_tok[i] <= max_input_length.
 * Any n: the call walks the batch in chunks of at most max_envs sequences / the engine's row capacity. */
int lmrl_forward_host(lmrl_engine* e, int32_t tower, int32_t head_set, const int32_t* tokens, const int32_t* n_tok, int32_t pitch,
                      int32_t n, float temperature, lmrl_forward_out* out);

/* ---- Text-Nav (llm_rl_scripts/text_nav/env/env.py:13-72 over the TextWorld game of env/make_game.py:30-107) --------------
 * Device-resident int32 state, 12 words per env: {room, loc[7 objects], open-container mask, score, done, moves}; encodings in
 * csrc/textnav_core.h.  reset: start room = random.Random(seed).choice(rooms) (make_game.py:82-83) unless start_room[i] >= 0.
 * step: `command.strip()` (env.py:38), exact match against the command strings TextWorld lists; reward = current score, done =
 * game over (env.py:69-70).  applied[i] = command code, -1 unknown text, -2 not possible here, -3 skipped / game already over.
 * adm [n][32] / n_adm [n]: codes of the admissible commands in TextWorld's (alphabetical) order, from which the host renders
 * "Admissible commands: ..." (env.py:44-45).  All pointers HOST; state / applied / adm / n_adm may be NULL. */
int lmrl_textnav_reset_host(lmrl_engine* e, const uint64_t* seeds, const int32_t* start_room, int32_t n, int32_t* state /*[n][12]*/,
                            int32_t* adm, int32_t* n_adm);
int lmrl_textnav_step_host(lmrl_engine* e, const uint8_t* action_bytes, const int32_t* action_len, int32_t pitch, const uint8_t* skip,
                           int32_t n, float* reward, uint8_t* done, int32_t* applied, int32_t* state, int32_t* adm, int32_t* n_adm);

/* ---- chess rules for the Chess-endgames environment (llm_rl_scripts/chess/env/env.py:29-186; python-chess is not a dependency) ---
 * HOST-side: the env's opponent is a UCI engine process, the GPU runs the policy.  One lmrl_chess = one board with its move stack.
 * lmrl_chess_fen            board.fen(): en-passant square only when a legal en-passant capture exists, cleaned castling rights
 * lmrl_chess_push_san       board.push_san(): 0 pushed, 1 null move (reported, not pushed; env.py:116-118), -1 illegal / ambiguous /
 *                           malformed (board unchanged, text in lmrl_last_error)
 * lmrl_chess_push_uci       board.san(move) into san_out, then board.push(move) (the opponent's reply, env.py:163-170)
 * lmrl_chess_status         bit 0 checkmate, 1 stalemate, 2 insufficient material, 3 seventy-five moves, 4 fivefold repetition
 *                           (any of them = board.is_game_over()), bit 5 = side to move in check
 * lmrl_chess_perft          leaf count of the legal-move tree (known-answer test hook) */
typedef struct lmrl_chess lmrl_chess;
lmrl_chess* lmrl_chess_new(const char* fen /* NULL = initial position */);
void lmrl_chess_free(lmrl_chess* c);
int lmrl_chess_fen(lmrl_chess* c, char* out, int32_t cap);
int lmrl_chess_legal_moves(lmrl_chess* c, char* out_uci /* space separated */, int32_t cap);
int lmrl_chess_push_san(lmrl_chess* c, const char* san, char* uci_out, int32_t cap);
int lmrl_chess_push_uci(lmrl_chess* c, const char* uci, char* san_out, int32_t cap);
int lmrl_chess_pop(lmrl_chess* c);
int lmrl_chess_status(lmrl_chess* c);
int lmrl_chess_turn(lmrl_chess* c);          /* 0 white, 1 black */
int lmrl_chess_is_valid(lmrl_chess* c);      /* board.is_valid() as far as generated endgames need it (env.py:227-243) */
int64_t lmrl_chess_perft(lmrl_chess* c, int32_t depth);

/* ---- PPO data path after the forward passes ------------------------------------------------------
 * Replaces the host loops of PPOInference.get_ppo_data_from_token_trajectory_chain (ppo/base_interface.py:572-669): KL-shaped
 * rewards  r - kl_weight * (logp - logp0) * is_action  (:572-584), get_action_state_next_state_idxs (:230-243), the GAE scan
 * get_advantages_and_returns (:253-293) over ACTION tokens only, and whitening over the whole batch (:245-251, :608-615).
 * Chains are flattened in CombinedTokenTrajectoryChain order (:303-334): chain c owns output tokens [chain_off[c], chain_off[c+1])
 * of the [N] arrays and values[chain_off[c] + c .. chain_off[c+1] + c] (its N_c state values followed by ONE bootstrap slot =
 * last value * (1 - done), :555-570).  Outputs are [N]; non-action tokens get 0.  kls = exp(lr) - 1 - lr at action tokens (:577-579).
 * gamma / lam / kl_weight are python scalars in the reference ("weak" against float32 arrays): cast to float once, float32 maths. */
int lmrl_ppo_advantages_host(lmrl_engine* e, int32_t n_chains, const int64_t* chain_off, const float* policy_logprobs,
                             const float* initial_logprobs, const float* values /*[N + n_chains]*/, const float* rewards,
                             const uint8_t* should_take_action, double gamma, double lam, double kl_weight, int32_t whiten,
                             float* advantages, float* returns, float* kls);
/* ppo_loss_fn (ppo/base_interface.py:71-142) over n = batch * (time - 1) elements; mask = should_take_action * attention_mask.
 * out[0..12] = total_loss, policy_loss, value_loss, values_error, value clipfrac, approx_kl, policy clipfrac, ratio,
 * padding_percentage, mean values, mean old_values, mean returns, n (mask sum). */
int lmrl_ppo_loss_host(lmrl_engine* e, int64_t n, const float* mask, const float* logprobs, const float* values,
                       const float* old_logprobs, const float* old_values, const float* old_advantages, const float* old_returns,
                       double cliprange_value, double cliprange, double value_loss_coef, double* out /*[16]*/);

/* trajectory -> token chain on the device: after a fused WORDLE rollout the engine still holds every episode's canonical token stream;
 * this exports it annotated as a TokenTrajectory (environment.py:346-383; the PPO driver's one-trajectory-per-episode conversion,
 * wordle/ppo/train_ppo_gpt2.py:307-335): tokens [n][pitch], n_tok [n], is_action [n][pitch], reward [n][pitch] (an action's reward on
 * its last token), done [n], truncated [n] (the stream was LEFT-truncated during the rollout and no longer starts at the first
 * token: the caller falls back to the text path for that episode).  No tokeniser runs on the host.  HOST pointers. */
int lmrl_rollout_chains_host(lmrl_engine* e, int32_t n, int32_t pitch, int32_t* tokens, int32_t* n_tok, uint8_t* is_action, float* reward,
                             uint8_t* done, uint8_t* truncated);

/* ---- gradient step of the policy + value head (GPT2PPOTrain._step, ppo/gpt2/interface.py:72-211) ----------------------------------
 * The chain rule is split at the per-token quantities: the caller supplies dL/dlogprob and dL/dvalue (the closed-form derivative of
 * ppo_loss_fn) and the engine back-propagates them through the tied LM head, the LinearHead value head (head set 0) and the decoder
 * stack of tower 0, ACCUMULATING fp32 gradients into the caller's DEVICE arrays (HF GPT-2 layout, same shapes as lmrl_gpt2_weights).
 * dlogprob[i][t] = dL/d log p(tokens[i][t+1] | tokens[i][..t]) for t < n_tok[i] - 1;  dvalue[i][t] = dL/d value(position t). */
typedef struct lmrl_gpt2_grads {
    float* wte;                  /* [n_vocab_rows, d] */
    int32_t n_vocab_rows;
    int32_t reserved;
    float* wpe;                  /* [n_pos, d] */
    float* const* ln1_g; float* const* ln1_b; float* const* attn_w; float* const* attn_b; float* const* proj_w; float* const* proj_b;
    float* const* ln2_g; float* const* ln2_b; float* const* fc_w; float* const* fc_b; float* const* fc2_w; float* const* fc2_b;   /* [L] -> as lmrl_gpt2_weights */
    float* lnf_g; float* lnf_b;
    float* vh_w;                 /* [d]  LinearHead kernel gradient, or NULL */
    float* vh_b;                 /* [1] */
} lmrl_gpt2_grads;
/* allocate the training buffers (row / scored-row capacities of one lmrl_policy_grad_host call) and the HF-layout weight copies */
int lmrl_train_enable(lmrl_engine* e, int32_t max_rows, int32_t max_scored_rows);
int lmrl_policy_grad_host(lmrl_engine* e, const int32_t* tokens, const int32_t* n_tok, int32_t pitch, int32_t n, const float* dlogprob,
                          const float* dvalue, float temperature, const lmrl_gpt2_grads* grads);

/* the same with ppo_loss_fn (ppo/base_interface.py:71-142) differentiated ON THE DEVICE: one forward with saved activations yields the
 * log-probs of the action tokens and the LinearHead values, the loss gradient with respect to them is formed in the LM-head / value
 * kernels, and the backward pass continues as in lmrl_policy_grad_host.  HOST arrays [n][pitch], index t = quantities of position t
 * (the token predicted there is tokens[t+1]); should_take_action[t] != 0 marks the action tokens (the loss mask).
 * loss_out[0..12] as lmrl_ppo_loss_host (total, policy, value, values_error, value clipfrac, approx_kl, policy clipfrac, ratio, ...). */
typedef struct lmrl_ppo_batch {
    const uint8_t* should_take_action;
    const float* old_logprobs; const float* old_values; const float* old_advantages; const float* old_returns;
    double cliprange_value, cliprange, value_loss_coef;
    double* loss_out;            /* [16] or NULL */
} lmrl_ppo_batch;
int lmrl_ppo_grad_host(lmrl_engine* e, const int32_t* tokens, const int32_t* n_tok, int32_t pitch, int32_t n, const lmrl_ppo_batch* batch,
                       const lmrl_gpt2_grads* grads);

/* optax.adamw(learning_rate, b1, b2, eps, weight_decay[, mask]) (train_ppo_gpt2.py:144-165) over one flat fp32 segment, in place.
 * DEVICE pointers, enqueued on the engine stream (lmrl_stream), no synchronisation; step is 1-based; grad_scale multiplies the
 * gradient first (1 / world_size after a summing all-reduce).  Segments under the weight-decay mask pass weight_decay = 0. */
int lmrl_adamw_dev(lmrl_engine* e, float* params, const float* grads, float* m, float* v, int64_t n, double lr, double b1, double b2,
                   double eps, double weight_decay, int64_t step, double grad_scale);

/* ---- fused rollout (interact_environment's while-loop, environment.py:176-206, with no host round trip) -------- */
typedef struct lmrl_rollout_out {      /* all HOST pointers for _host, DEVICE pointers for _dev; any may be NULL */
    float* ep_return;       /* [n]  sum of rewards (fp32 accumulation in step order)            :238 */
    int32_t* ep_length;     /* [n]  number of transitions                                        :240 */
    uint8_t* ep_done;       /* [n]  done flag of the last transition                             :239 */
    float* step_reward;     /* [n][max_turns] */
    uint8_t* step_valid;    /* [n][max_turns] */
    uint8_t* step_feedback; /* [n][max_turns][5]  (Wordle) */
    int32_t* step_aux;      /* [n][max_turns]  Wordle: filtered size; Maze: outcome | cell_after << 8 | cell_before << 20 */
    int32_t* act_tokens;    /* [n][max_turns][max_new_tokens]  sampled ids per turn */
    int32_t* act_len;       /* [n][max_turns] */
} lmrl_rollout_out;

int lmrl_rollout_host(lmrl_engine* e, int32_t env_kind, const uint64_t* seeds, const int32_t* opt_pos, int32_t n,
                      const lmrl_sampling* s, lmrl_rollout_out* out);
int lmrl_rollout_dev(lmrl_engine* e, int32_t env_kind, const uint64_t* seeds_dev, const int32_t* opt_pos_dev, int32_t n,
                     const lmrl_sampling* s, lmrl_rollout_out* out_dev);
/* scripted-action variant (env kernels only; actions come from the device RandomMixturePolicy equivalent,
 * wordle/env/scripted_policies.py:114-126): used by the env microbenchmark */
int lmrl_wordle_scripted_rollout_dev(lmrl_engine* e, const uint64_t* seeds_dev, int32_t n, float prob_smart,
                                     uint64_t policy_seed, lmrl_rollout_out* out_dev);

/* fused rollouts of up to 8 turns are captured into a CUDA graph on the second use of a configuration and replayed afterwards
 * (default on; per-launch profiling runs eagerly) */
int lmrl_set_use_graphs(lmrl_engine* e, int32_t enable);
/* fused rollouts run the two halves of the batch as two concurrent lanes on two streams (HBM-bound attention of one half overlaps the
 * tensor-bound GEMMs of the other; results are identical — every env's arithmetic is independent of the batch split).  Default OFF:
 * measured slower on B200 (the GEMM and attention CTAs do not fit one SM together, so the lanes serialise). */
int lmrl_set_two_lanes(lmrl_engine* e, int32_t enable);
/* programmatic dependent launch between the kernels of a pass (default OFF — measured slower on B200 for this path): a kernel's
 * CTAs are scheduled and run their prologue behind the tail of the previous kernel, then wait for its completion before touching
 * global memory */
int lmrl_set_use_pdl(lmrl_engine* e, int32_t enable);
/* named integer switches of the policy engine (A/B measurement and cross-checks; every value gives identical results up to the
 * documented float tolerance):
 *   "fuse_ln"        1 LayerNorm fused into the residual GEMMs (the CTA completing an m-tile normalises its rows; bit-identical,
 *                    measured slower on B200), 0 (default) separate layernorm_kernel launches
 *   "attn_wide"      1 (default) prefill attention takes 32 query rows per warp pass (K / V streamed once for slots of <= 32 new rows),
 *                    0 16-row passes
 *   "gemv_max_rows"  decode passes of batches with at most this many envs run the weight-streaming GEMV kernels instead of
 *                    128-row tensor-core tiles (default 8; 0 disables)
 *   "att_decode_impl" decode-pass attention: 4 (default) the streaming SIMT kernel with 4 consecutive heads of a slot per warp (the slot's
 *                    metadata chain slot -> env -> page table is paid once per 4 heads: +14 % at context 24, +4 % at 40, -5 % at 80 against
 *                    one head per warp in isolation, -1 % per Wordle rollout), 3 two heads per warp, 0 one head per warp, 2 one head per warp
 *                    with software-pipelined K / V loads (0, 2, 3, 4 are bit-identical), 1 the cp.async-staged tensor-core prefill kernel
 *                    with one row per slot (1 and 2 measured 1-2 % slower per rollout); "att_grid_mult" caps the SIMT kernels' grid at
 *                    that many CTAs per SM (default 8)
 *   "ln_fold"        1 LayerNorm folded into the CONSUMER GEMMs: the residual GEMMs write x, its bf16 copy and per-row partial (sum, sum of
 *                    squares); QKV / FC run on the raw copy against gamma-folded row-centred weights and scale their accumulators by
 *                    1 / sigma.  No LayerNorm launches in the decoder stack; results within the float tolerance of the default path
 *                    (different rounding points); measured 0-1 % faster per rollout, 0 (default)
 *   "tok_nlnl_id"    id of the tokenizer's '\n\n' token, -1 (default) = none.  The reference encodes the JOINED history every turn
 *                    (ppo/gpt2/interface.py:519-522), so with GPT-2's merges a prompt ending in an action line + an empty feedback
 *                    line ends in that single token; the fused rollout reproduces this in its prompt view (the stream itself stays
 *                    tokenised per Text item).  The Python policy classes set it from the tokenizer.
 * Unknown names fail.  Changing a switch drops the captured rollout graph. */
int lmrl_set_option(lmrl_engine* e, const char* name, int32_t value);
int lmrl_get_option(lmrl_engine* e, const char* name, int32_t* value);

/* paged KV cache: pages handed out since the current batch started, pool capacity, positions per page; fails if the pool overflowed */
int lmrl_kv_stats(lmrl_engine* e, int32_t* pages_in_use, int32_t* pages_total, int32_t* page_positions);

/* counters for bench.py: kernels launched by this engine since the last reset */
int64_t lmrl_launch_count(lmrl_engine* e, int32_t reset);

/* GEMM implementation switch (default 2 = tcgen05/TMEM/TMA with CTA pairs (cta_group::2, 256x256 tiles) on the plain GEMMs of prefill
 * passes and one CTA per 128x256 tile elsewhere; 1 = one CTA per tile everywhere; 0 = the mma.sync kernel kept as an independent
 * cross-check) and a unit-test hook:
 * resid[M,N] += A[M,K] · W[N,K]^T + bias through the chosen implementation (bf16 bit patterns in); the hook also accepts
 * impl 3 = the 128x192-tile tcgen05 kernel the engine uses for the N = 768 GEMMs of decode passes */
int lmrl_set_gemm_impl(lmrl_engine* e, int32_t impl);
int lmrl_test_gemm_host(lmrl_engine* e, int32_t impl, int32_t M, int32_t N, int32_t K, const uint16_t* A_bf16,
                        const uint16_t* W_bf16, const float* bias, float* resid_inout);

/* kernel microbenchmark on the engine's own buffers and layer-0 weights: M synthetic rows (M <= max_envs: one row per
 * env at position ctx-1, i.e. a decode pass; otherwise M/max_envs rows per env ending at position ctx-1, a prefill pass).
 * ms_out[8] = mean device time of {QKV, attn-proj, FC, FC2, LM-head+sampler GEMMs, attention, LayerNorm, embed}. */
int lmrl_bench_kernels(lmrl_engine* e, int32_t M, int32_t ctx, int32_t iters, float* ms_out);

/* per-launch device timing for the roofline report: CUDA events are recorded on the engine stream around every kernel
 * launch while enabled.  Categories: 0 GEMM (tensor), 1 attention, 2 element-wise / bookkeeping, 3 env kernels. */
#define LMRL_PROF_CATS 4
typedef struct lmrl_profile_out {
    double ms[LMRL_PROF_CATS];        /* summed launch durations */
    double work[LMRL_PROF_CATS];      /* algorithmic FLOPs (cat 0: 2*M*N*K with the live row count M of each launch) */
    int64_t launches[LMRL_PROF_CATS];
    /* finer split, index = 2*kind + (0 prefill pass | 1 decode pass); kind: 0 QKV GEMM, 1 attn-proj GEMM, 2 FC GEMM,
     * 3 FC2 GEMM, 4 LM-head GEMM, 5 attention */
    double tag_ms[12];
    double tag_work[12];
    int64_t rows;                     /* rows that went through the decoder (prefill + decode) */
    int64_t sampled;                  /* tokens sampled */
} lmrl_profile_out;
int lmrl_profile_enable(lmrl_engine* e, int32_t enable);      /* also clears the records */
int lmrl_profile_read(lmrl_engine* e, lmrl_profile_out* out); /* synchronises the stream */

#ifdef __cplusplus
}
#endif
#endif /* LMRL_B200_H */
