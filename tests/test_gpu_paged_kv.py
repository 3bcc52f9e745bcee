"""GPU tests of the paged KV cache (north_star "paged-KV attention"; reference cache shape: a dense [B, max_length] slab per layer,
value_rl_base/gpt2/generation.py:123-150).

  * a pool sized for the positions a batch really reaches gives bit-identical logits / tokens / rollout logs to the full-size pool
    (page ids differ, the arithmetic does not);
  * pages in use follow the live positions (ceil(len / 32) per env), are kept across turns and released when a new batch starts;
  * a pool that is too small fails loudly (no silent overwrite of another episode's cache);
  * 8192 concurrent Wordle episodes of GPT-2-small fit one GPU with a pool sized for the ~140 tokens an episode reaches
    (the slab would have taken 2 x 39.9 GB for the cache alone).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from lmrl_gym_b200.gpt2 import GPT2Config, init_gpt2_params  # noqa: E402
from lmrl_gym_b200.tokenizer import ByteTokenizer  # noqa: E402


def _cfg():
    return GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)


def test_small_pool_is_bit_identical_and_counts_pages():
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    tok = ByteTokenizer.synthetic()
    cfg = _cfg()
    params = init_gpt2_params(cfg, seed=2)
    rng = np.random.default_rng(0)
    lens = rng.integers(1, 200, size=96)
    lens[:4] = (1, 32, 33, 200)
    prompts = [rng.integers(0, cfg.n_vocab, size=int(n)).tolist() for n in lens]
    keys = [int(k) for k in rng.integers(0, 2**63, size=96)]
    need_prefill = int(sum((n + 31) // 32 for n in lens))
    outs = []
    for kv_pages in (0, need_prefill + 96):            # full pool (128 envs x 7 pages) vs just enough (+1 page per row for the decoded tokens)
        eng = Engine(max_envs=128, max_input_length=200, max_new_tokens=5, tokenizer=tok, kv_pages=kv_pages)
        pol = B200GPT2Policy(eng, params, cfg, seed=3)
        logits = pol.logits(prompts)
        used, total, page = eng.kv_stats()
        assert page == 32 and total == (kv_pages or 128 * 7) and used == need_prefill, (used, need_prefill, total)
        gens = pol.generate_tokens(prompts, episode_keys=keys, turn=2)
        used2, _, _ = eng.kv_stats()
        grown = int(sum((n + len(g) - 1 + 31) // 32 for n, g in zip(lens, gens)))      # the last sampled token is never fed back
        assert used2 == grown, (used2, grown)
        outs.append((logits, gens))
        eng.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]


def test_exhausted_pool_fails_loudly():
    from lmrl_gym_b200._lib import LmrlError
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    tok = ByteTokenizer.synthetic()
    cfg = _cfg()
    eng = Engine(max_envs=64, max_input_length=128, max_new_tokens=4, tokenizer=tok, kv_pages=40)
    pol = B200GPT2Policy(eng, init_gpt2_params(cfg, seed=2), cfg, seed=3)
    rng = np.random.default_rng(1)
    ok = [rng.integers(0, cfg.n_vocab, size=30).tolist() for _ in range(40)]            # 40 pages: fits exactly
    assert np.isfinite(pol.logits(ok)).all()
    too_many = [rng.integers(0, cfg.n_vocab, size=40).tolist() for _ in range(40)]      # 80 pages
    with pytest.raises(LmrlError, match="paged KV cache exhausted"):
        pol.generate_tokens(too_many)
    assert np.isfinite(pol.logits(ok)).all()                                            # the next batch starts from an empty pool
    eng.close()


def test_fused_rollout_with_a_small_pool_matches_the_full_pool():
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, steer_to_letters
    from lmrl_gym_b200.rollout import run_fused_batch
    from lmrl_gym_b200.wordle import B200WordleEnv
    from oracle.mt19937 import seed_stream
    tok = ByteTokenizer.synthetic()
    cfg = _cfg()
    params = steer_to_letters(init_gpt2_params(cfg, seed=4), tok)
    gen = seed_stream(3)
    seeds = [next(gen) for _ in range(256)]
    logs = []
    for kv_pages in (0, 256 * 6):                       # an episode reaches ~8 + 6 * 24 = 152 tokens: 5-6 pages of the 9 a full-length env owns
        eng = Engine(max_envs=256, max_input_length=256, max_new_tokens=6, max_turns=6, tokenizer=tok, kv_pages=kv_pages)
        env = B200WordleEnv(eng, "wordle_official.txt", require_words_in_vocab=False)
        pol = B200GPT2Policy(eng, params, cfg, seed=1)
        for graphs in (False, True, True):
            eng.set_use_graphs(graphs)
            lg = run_fused_batch(eng, env, pol, seeds)
            logs.append(lg)
        used, total, _ = eng.kv_stats()
        assert 256 <= used <= 256 * 6 and total == (kv_pages or 256 * 9)
        eng.close()
    for lg in logs[1:]:
        for name in ("ep_return", "ep_length", "ep_done", "step_reward", "step_valid", "step_feedback", "step_aux", "act_tokens", "act_len"):
            assert np.array_equal(getattr(lg, name), getattr(logs[0], name)), name


def test_8192_wordle_episodes_fit_one_gpu():
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    from lmrl_gym_b200.rollout import run_fused_batch
    from lmrl_gym_b200.wordle import B200WordleEnv
    from oracle.mt19937 import seed_stream
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config.small()
    B = 8192
    eng = Engine(max_envs=B, max_input_length=256, max_new_tokens=6, max_turns=6, tokenizer=tok, kv_pages=B * 6)
    env = B200WordleEnv(eng, "wordle_official.txt")
    pol = B200GPT2Policy(eng, init_gpt2_params(cfg, seed=0), cfg, seed=0)
    gen = seed_stream(0)
    lg = run_fused_batch(eng, env, pol, [next(gen) for _ in range(B)])
    used, total, _ = eng.kv_stats()
    print(f"B=8192: device bytes {eng.device_bytes / 1e9:.1f} GB, KV pages in use {used} of {total}")
    assert (lg.ep_length == 6).all() and (lg.ep_return == -6.0).all() and used <= total
    assert eng.device_bytes < 150e9
    eng.close()
