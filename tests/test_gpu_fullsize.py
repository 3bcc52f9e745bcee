"""Full-size properties (BASELINE.json configs[1]: 4096 concurrent Wordle episodes, GPT-2-small shapes) that do not need the
oracle to run 4096 episodes:

  * determinism: the same seeds give bit-identical logs, eagerly launched and replayed from the CUDA graph, as one lane or as two
    concurrent half-batch lanes on two streams, at the full batch and at a ragged one;
  * batch-layout independence: sampling noise is keyed by (seed, turn, token) and no kernel mixes rows, so permuting the
    seeds permutes the per-episode logs exactly — the property that makes multi-GPU sharding result-invariant;
  * a random sample of the 4096 episodes replays bit-exactly through the CPU oracle env (texts, rewards, done, return);
  * summary statistics equal the float32 reductions of LLM_RL/environment.py:244-265 over the logged returns.

The policy is the random-init GPT-2-small with the LM head steered to letters (``steer_to_letters``) so that guesses are
well-formed and, with require_words_in_vocab=False, exercise target draws / transitions / wins.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _logs_equal(a, b, perm=None):
    for name in ("ep_return", "ep_length", "ep_done", "step_reward", "step_valid", "step_feedback", "step_aux", "act_tokens", "act_len"):
        x, y = getattr(a, name), getattr(b, name)
        if perm is not None:
            y = y[perm]
        if not np.array_equal(x, y):
            return name
    return None


def _logs_equal_prefix(a, b, n):
    for name in ("ep_return", "ep_length", "ep_done", "step_reward", "step_valid", "step_feedback", "step_aux", "act_tokens", "act_len"):
        if not np.array_equal(getattr(a, name)[:n], getattr(b, name)[:n]):
            return name
    return None


def test_full_size_rollout_properties():
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params, steer_to_letters
    from lmrl_gym_b200.rollout import WordleEpisode, episode_returns, run_fused_batch
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200.vocab import load_vocab_words
    from lmrl_gym_b200.wordle import B200WordleEnv
    from oracle.mt19937 import seed_stream
    from oracle.wordle import ReformatWordleOracle
    B = 4096
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config.small()
    params = steer_to_letters(init_gpt2_params(cfg, seed=0), tok)
    eng = Engine(max_envs=B, max_input_length=256, max_new_tokens=6, max_turns=6, tokenizer=tok)
    env = B200WordleEnv(eng, "wordle_official.txt", require_words_in_vocab=False)
    pol = B200GPT2Policy(eng, params, cfg, seed=0)
    gen = seed_stream(0)
    seeds = [next(gen) for _ in range(B)]
    eng.set_use_graphs(False)
    eng.set_two_lanes(False)
    one_lane = run_fused_batch(eng, env, pol, seeds)       # the whole batch as one lane on one stream
    eng.set_two_lanes(True)
    two = run_fused_batch(eng, env, pol, seeds)            # two concurrent half-batch lanes on two streams
    assert _logs_equal(one_lane, two) is None, _logs_equal(one_lane, two)
    ragged = run_fused_batch(eng, env, pol, seeds[:1001])   # odd batch: lanes of 501 and 500 envs
    assert _logs_equal_prefix(one_lane, ragged, 1001) is None
    eng.set_two_lanes(False)
    eager = one_lane
    eng.set_use_graphs(True)
    first = run_fused_batch(eng, env, pol, seeds)          # eager again (first use of the configuration with graphs on)
    captured = run_fused_batch(eng, env, pol, seeds)       # captured + launched
    replayed = run_fused_batch(eng, env, pol, seeds)       # replayed
    for other in (first, captured, replayed):
        assert _logs_equal(eager, other) is None, _logs_equal(eager, other)
    perm = np.random.default_rng(1).permutation(B)
    shuffled = run_fused_batch(eng, env, pol, [seeds[i] for i in perm])
    inv = np.empty(B, np.int64)
    inv[perm] = np.arange(B)
    assert _logs_equal(eager, shuffled, inv) is None, _logs_equal(eager, shuffled, inv)
    # every episode terminated, lengths in 1..6, both env branches were taken, somebody won
    assert eager.ep_done.all() and eager.ep_length.min() >= 1 and eager.ep_length.max() == 6
    legal = eager.step_valid.astype(bool) & (np.arange(6)[None, :] < eager.ep_length[:, None])
    assert legal.sum() > 1000 and (~legal).sum() > 1000
    # replay a sample through the oracle env
    words = load_vocab_words("wordle_official.txt")
    rets = episode_returns(eager, env)
    for i in np.random.default_rng(2).choice(B, size=48, replace=False):
        ep = WordleEpisode(eager, int(i), tok, env, None)
        o = ReformatWordleOracle(words, require_words_in_vocab=False)
        h = o.reset(seeds[i])
        total = 0.0
        for tr in ep:
            assert [[x.text, x.is_action] for x in tr.pre_action_history] == [[a, b] for a, b in h]
            h, r, d = o.step(h + ((tr.post_action_history[-1].text, True),))
            assert [[x.text, x.is_action] for x in tr.post_transition_history] == [[a, b] for a, b in h]
            assert tr.reward == r and type(tr.reward) is type(r) and tr.done == d
            total += r
        assert abs(total - rets[i]) <= 1e-5 and abs(float(eager.ep_return[i]) - total) <= 1e-5
    eng.close()
