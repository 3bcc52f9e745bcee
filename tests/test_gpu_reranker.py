"""GPU parity of teacher-forced scoring (lmrl_score_host) and the re-ranker policies against the torch-fp32 oracle.

Tolerance: a token log-probability is (logit - logsumexp); with logits within 2e-2 of the bf16-emulating oracle the
term is within 4e-2, the sum over an action of <= 6 tokens within 1.5e-1 (observed ~1e-2)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from lmrl_gym_b200.gpt2 import GPT2Config, init_gpt2_params  # noqa: E402
from lmrl_gym_b200.tokenizer import ByteTokenizer  # noqa: E402

TOL_TERM = 4e-2


def test_score_matches_oracle_and_reranker_picks_argmax():
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.environment import Text
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    from lmrl_gym_b200.maze import describe, maze_grid
    from lmrl_gym_b200.reranker import (ReRankerPolicy, ReRankerSamplePolicy, build_bc_score_fn, maze_proposal_function,
                                        score_token_sequences, token_history)
    from oracle import gpt2 as og
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)
    params = init_gpt2_params(cfg, seed=9)
    eng = Engine(max_envs=64, max_input_length=96, max_new_tokens=4, tokenizer=tok)
    pol = B200GPT2Policy(eng, params, cfg)
    rng = np.random.default_rng(0)
    # raw token sequences, ragged, several prefix lengths (incl. scoring everything after the first token, and nothing)
    seqs = [rng.integers(0, cfg.n_vocab, size=int(n)).tolist() for n in rng.integers(2, 96, size=40)]
    prefix = [int(rng.integers(1, len(s) + 1)) for s in seqs]
    prefix[0], prefix[1] = 1, len(seqs[1])
    sums, terms = score_token_sequences(pol, seqs, prefix, return_token_logprobs=True)
    worst = 0.0
    for i in range(0, 40, 3):
        want = og.sequence_logprob(params, seqs[i], prefix[i], cfg.n_head, cfg.n_vocab, emulate_bf16=True)
        assert len(want) == len(terms[i]) == len(seqs[i]) - prefix[i]
        if len(want):
            worst = max(worst, np.abs(terms[i] - want).max())
            assert abs(float(sums[i]) - want.sum()) <= TOL_TERM * max(len(want), 1)
    assert sums[1] == 0.0
    print(f"token log-prob: max |diff| vs bf16-emulating oracle {worst:.4g}")
    assert worst <= TOL_TERM
    # temperature
    s2 = score_token_sequences(pol, seqs[:4], prefix[:4], temperature=0.7)
    want = og.sequence_logprob(params, seqs[2], prefix[2], cfg.n_head, cfg.n_vocab, temperature=0.7, emulate_bf16=True)
    assert abs(float(s2[2]) - want.sum()) <= TOL_TERM * max(len(want), 1) / 0.7
    # re-ranker over the four maze moves: picks the oracle's arg-max (unless the top two are within tolerance)
    grid = maze_grid("double_t_maze")
    score_fn = build_bc_score_fn(pol)
    policy = ReRankerPolicy(maze_proposal_function, score_fn)
    for pos in [(1, 1), (5, 6), (8, 6), (1, 11)]:
        h = (Text(describe(1, grid, pos, (8, 6)), False),)
        out = policy.act(h)
        props = maze_proposal_function(h)
        want = []
        for p in props:
            full, pre = token_history(p, tok), len(token_history(p[:-1], tok))
            want.append(og.sequence_logprob(params, full, pre, cfg.n_head, cfg.n_vocab, emulate_bf16=True).sum())
        got = np.asarray(score_fn(props))
        assert np.abs(got - np.asarray(want)).max() <= 4 * TOL_TERM
        best = int(np.argmax(want))
        chosen = [p[-1].text for p in props].index(out[-1].text)
        assert chosen == best or want[best] - want[chosen] <= 8 * TOL_TERM
    sp = ReRankerSamplePolicy(maze_proposal_function, score_fn, rng=np.random.default_rng(1))
    assert sp.act(h)[-1].text in [p[-1].text for p in props]
    eng.close()


def test_score_value_policy_uses_perturbed_logits():
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2ValuePolicy, init_mlp_head_params
    from lmrl_gym_b200.reranker import score_token_sequences
    from oracle import gpt2 as og
    import torch
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)
    params = (init_gpt2_params(cfg, seed=1), init_gpt2_params(cfg, seed=2), init_mlp_head_params(cfg, seed=3), init_mlp_head_params(cfg, seed=4))
    eng = Engine(max_envs=16, max_input_length=48, max_new_tokens=2, tokenizer=tok)
    pol = B200GPT2ValuePolicy(eng, params, cfg, beta=4.0)
    rng = np.random.default_rng(3)
    seqs = [rng.integers(0, cfg.n_vocab, size=int(n)).tolist() for n in (12, 30, 7)]
    prefix = [10, 27, 6]
    sums, terms = score_token_sequences(pol, seqs, prefix, return_token_logprobs=True)
    for i, s in enumerate(seqs):
        want = []
        for t in range(prefix[i], len(s)):
            logits = og.value_policy_last_logits(params, s[:t], cfg.n_head, cfg.n_vocab, 4.0, emulate_bf16=True)
            want.append(float(torch.log_softmax(torch.from_numpy(logits).double(), -1)[s[t]]))
        assert np.abs(terms[i] - np.asarray(want)).max() <= 0.25, (i, terms[i], want)
    eng.close()
