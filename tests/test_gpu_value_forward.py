"""GPU parity of the per-position teacher-forced forward (lmrl_forward_host) against the torch-fp32 oracle:
token log-probabilities of every position (PPOInference.token_logprobs_from_logits, ppo/base_interface.py:404-411), the Q heads
gathered at the next token and the V head of every position (ValueRLInference._forward, value_rl_base/gpt2/interface.py:150-230;
ilql/gpt2/score_fn.py:45-58), the PPO LinearHead value (heads/linear_head.py:112-119), and the ILQL / MC advantage score functions
built on them.  The policy side of the reference is unpinned (JAX stack absent), so these are tolerance tests against the oracle
restatement; the score functions' host logic is pinned separately (tests/test_oracle_score_fns.py).

Tolerances (2-layer d=256 GPT-2, lecun-normal heads): Q / V are O(1..5) sums over 256 bf16 relu activations -> 6e-2 vs the oracle that
rounds where the engine rounds, 1.5e-1 vs plain fp32; log-probs 3e-2 / 8e-2; the linear value head is computed in fp32 from the
residual stream: 2e-2 / 6e-2.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from lmrl_gym_b200.gpt2 import GPT2Config, init_gpt2_params, init_mlp_head_params  # noqa: E402
from lmrl_gym_b200.tokenizer import ByteTokenizer  # noqa: E402


@pytest.fixture(scope="module")
def tok():
    return ByteTokenizer.synthetic()


def _cfg():
    return GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)


def _heads(cfg, seed):
    rng = np.random.default_rng(seed + 50)
    hs = []
    for k in range(3):
        h = init_mlp_head_params(cfg, seed=seed + k, n_out=(1 if k == 2 else None), layer2_bias_init=(-4.4 if k < 2 else 0.5))
        h["dense1"]["bias"] = (0.1 * rng.standard_normal(cfg.d_model)).astype(np.float32)
        hs.append(h)
    return hs


def _seqs(cfg, n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, cfg.n_vocab, size=int(m)).tolist() for m in rng.integers(lo, hi, size=n)]


def test_forward_value_heads_and_logprobs_match_oracle(tok):
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    from lmrl_gym_b200 import value_inference as VI
    from oracle import gpt2 as og
    cfg = _cfg()
    pi, base, tbase = (init_gpt2_params(cfg, seed=s) for s in (0, 1, 2))
    q1, q2, v = _heads(cfg, 10)
    tq1, tq2, _ = _heads(cfg, 20)
    eng = Engine(max_envs=48, max_input_length=64, max_new_tokens=4, tokenizer=tok)
    B200GPT2Policy(eng, pi, cfg)
    online = VI.B200ValueRLInference(eng, base, q1, q2, v, cfg, tower=1, head_set=0)
    target = VI.B200ValueRLInference(eng, tbase, tq1, tq2, None, cfg, tower=2, head_set=1)
    seqs = _seqs(cfg, 130, 1, 65, 3)            # 130 sequences > max_envs: the call walks them in chunks
    r_on, r_tg = online.forward_ragged(seqs), target.forward_ragged(seqs)
    r_lp = VI.forward_tokens(eng, 0, 0, seqs, logprobs=True)
    worst = dict(q_emu=0.0, q_f32=0.0, v_emu=0.0, v_f32=0.0, lp_emu=0.0, lp_f32=0.0)
    for i in range(0, 130, 7):
        s, n = seqs[i], len(seqs[i])
        for res, (b, a1, a2, av) in ((r_on, (base, q1, q2, v)), (r_tg, (tbase, tq1, tq2, None))):
            for emu in (True, False):
                o = og.value_forward_gathered(b, a1, a2, av, s, cfg.n_head, emulate_bf16=emu)
                key = "emu" if emu else "f32"
                for name in ("q1", "q2"):
                    if n > 1:
                        worst["q_" + key] = max(worst["q_" + key], float(np.abs(res[name][i, :n - 1] - o[name]).max()))
                if av is not None:
                    worst["v_" + key] = max(worst["v_" + key], float(np.abs(res["v"][i, :n] - o["v"]).max()))
            assert (res["q1"][i, max(n - 1, 0):] == 0).all()            # cells past the sequence stay 0
        if n > 1:
            for emu in (True, False):
                lp = og.token_logprobs_all(pi, s, cfg.n_head, cfg.n_vocab, emulate_bf16=emu)
                key = "lp_emu" if emu else "lp_f32"
                worst[key] = max(worst[key], float(np.abs(r_lp["logprobs"][i, :n - 1] - lp).max()))
    print("per-position forward, max |diff| vs oracle:", {k: round(x, 5) for k, x in worst.items()})
    assert worst["q_emu"] <= 6e-2 and worst["q_f32"] <= 1.5e-1
    assert worst["v_emu"] <= 6e-2 and worst["v_f32"] <= 1.5e-1
    assert worst["lp_emu"] <= 3e-2 and worst["lp_f32"] <= 8e-2
    # towers really differ (the target call did not silently reuse the online networks)
    assert np.abs(r_on["q1"] - r_tg["q1"]).mean() > 1e-2
    # forward() on a RIGHT-padded matrix == the ragged call
    pad = cfg.n_vocab
    mat = np.full((20, 64), pad, np.int32)
    for j in range(20):
        mat[j, :len(seqs[j])] = seqs[j]
    out = online.forward(mat)
    assert out.q1sa.shape == (20, 63) and out.v.shape == (20, 64)
    w = min(63, r_on["q1"].shape[1])
    np.testing.assert_array_equal(out.q1sa[:, :w], r_on["q1"][:20, :w])
    np.testing.assert_array_equal(out.v[:, :w], r_on["v"][:20, :w])
    eng.close()


def test_forward_linear_value_head_matches_oracle(tok):
    """PPO: policy log-probs + LinearHead values from tower 0, initial-policy log-probs from tower 1."""
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, _upload_params
    from lmrl_gym_b200 import value_inference as VI
    from oracle import gpt2 as og
    cfg = _cfg()
    pol, init = init_gpt2_params(cfg, seed=5), init_gpt2_params(cfg, seed=6)
    lin = dict(dense=dict(kernel=(np.random.default_rng(9).standard_normal((cfg.d_model, 1)) / 16).astype(np.float32), bias=np.asarray([0.125], np.float32)))
    eng = Engine(max_envs=32, max_input_length=48, max_new_tokens=4, tokenizer=tok)
    B200GPT2Policy(eng, pol, cfg)
    _upload_params(eng, cfg, init, tower=1)
    VI.upload_value_heads(eng, 0, v_linear=lin)
    seqs = _seqs(cfg, 40, 2, 49, 8)
    r0 = VI.forward_tokens(eng, 0, 0, seqs, logprobs=True, v=True)
    r1 = VI.forward_tokens(eng, 1, 0, seqs, logprobs=True)
    wv = wv32 = wl = 0.0
    for i in range(0, 40, 3):
        s, n = seqs[i], len(seqs[i])
        wv = max(wv, float(np.abs(r0["v"][i, :n] - og.linear_value_forward(pol, lin, s, cfg.n_head, emulate_bf16=True)).max()))
        wv32 = max(wv32, float(np.abs(r0["v"][i, :n] - og.linear_value_forward(pol, lin, s, cfg.n_head)).max()))
        wl = max(wl, float(np.abs(r0["logprobs"][i, :n - 1] - og.token_logprobs_all(pol, s, cfg.n_head, cfg.n_vocab, emulate_bf16=True)).max()))
        wl = max(wl, float(np.abs(r1["logprobs"][i, :n - 1] - og.token_logprobs_all(init, s, cfg.n_head, cfg.n_vocab, emulate_bf16=True)).max()))
    print(f"linear value head: max|diff| {wv:.4g} (emulating oracle) {wv32:.4g} (fp32); log-probs of both towers {wl:.4g}")
    assert wv <= 2e-2 and wv32 <= 6e-2 and wl <= 3e-2
    eng.close()


@pytest.mark.parametrize("kind", ["ilql", "mc"])
def test_advantage_score_fns_and_reranker(tok, kind):
    """build_ilql_score_fn / build_mc_score_fn on the device == the oracle score function fed with the oracle networks; a
    ReRankerPolicy over the four Maze moves picks the oracle's arg-max when its margin exceeds the tolerance."""
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    from lmrl_gym_b200 import value_inference as VI
    from lmrl_gym_b200 import reranker as RR
    from lmrl_gym_b200.environment import Text
    from oracle import gpt2 as og
    cfg = _cfg()
    pi, base, tbase = (init_gpt2_params(cfg, seed=s) for s in (30, 31, 32))
    q1, q2, v = _heads(cfg, 40)
    tq1, tq2, _ = _heads(cfg, 45)
    eng = Engine(max_envs=32, max_input_length=96, max_new_tokens=4, tokenizer=tok)
    pi_pol = B200GPT2Policy(eng, pi, cfg)
    if kind == "ilql":
        inf = VI.B200ILQLInference(VI.B200ValueRLInference(eng, base, q1, q2, v, cfg, tower=1, head_set=0),
                                   VI.B200ValueRLInference(eng, tbase, tq1, tq2, None, cfg, tower=2, head_set=1))
        fn = RR.build_ilql_score_fn(inf, pi_pol, tok, max_length=96, value_weight=1.0, logit_weight=0.5, bsize=8)
    else:
        inf = VI.B200ValueRLInference(eng, base, q1, None, None, cfg, tower=1, head_set=0)
        fn = RR.build_mc_score_fn(inf, pi_pol, tok, max_length=96, value_weight=1.0, logit_weight=0.5, bsize=8)
    obs = ["The goal is at position 8, 6. Your current position is at position 5, 3. There are walls to your left, below you.\n",
           "There are walls above you.\n", "Your current position is at position 1, 1 0.\n"]
    n_ok = n_clear = 0
    for k, o in enumerate(obs):
        hist = (Text(o, False),) if k != 1 else (Text(obs[0], False), Text("move up\n", True), Text(o, False))
        props = RR.maze_proposal_function(hist)
        got = np.asarray(fn(props), np.float64)
        want = []
        for p in props:
            ids = RR.token_history(p, tok)
            pre = len(RR.token_history(p[:-1], tok))
            if kind == "ilql":
                tg = og.value_forward_gathered(tbase, tq1, tq2, None, ids, cfg.n_head, emulate_bf16=True)
                on = og.value_forward_gathered(base, None, None, v, ids, cfg.n_head, emulate_bf16=True)
                per = np.minimum(tg["q1"], tg["q2"]) - on["v"][:-1]
            else:
                per = og.value_forward_gathered(base, q1, None, None, ids, cfg.n_head, emulate_bf16=True)["q1"]
            lp = og.token_logprobs_all(pi, ids, cfg.n_head, cfg.n_vocab, emulate_bf16=True)
            want.append(per[pre - 1:].sum() + 0.5 * lp[pre - 1:].sum())
        want = np.asarray(want)
        n_act = len(RR.token_history(props[0], tok)) - len(RR.token_history(props[0][:-1], tok))
        tol = n_act * (6e-2 * 2 + 0.5 * 3e-2)
        assert np.abs(got - want).max() <= tol, (got, want)
        choice = RR.ReRankerPolicy(RR.maze_proposal_function, fn).act(hist)
        srt = np.sort(want)
        if srt[-1] - srt[-2] > 2 * tol:
            n_clear += 1
            n_ok += choice[-1].text == RR.MAZE_ACTIONS[int(np.argmax(want))]
    assert n_ok == n_clear
    eng.close()


def test_forward_linear_q_head_matches_oracle(tok):
    """MC returns: the Q head is a LinearHead over the base model's last hidden state (train_mc_returns_gpt2.py:198-206,
    mc_returns/score_fn.py): q(h_t)[tok_{t+1}] = h_t . W[:, tok_{t+1}] + b — one kernel column gathered per position; an MLP V head next
    to it keeps working (mixed head kinds in one head set)."""
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    from lmrl_gym_b200 import value_inference as VI
    from oracle import gpt2 as og
    cfg = _cfg()
    pi, base = init_gpt2_params(cfg, seed=0), init_gpt2_params(cfg, seed=1)
    rng = np.random.default_rng(8)
    lin = lambda bias: dict(dense=dict(kernel=(rng.standard_normal((cfg.d_model, cfg.n_vocab_rows)) / np.sqrt(cfg.d_model)).astype(np.float32),  # noqa: E731
                                       bias=(0.3 * rng.standard_normal(cfg.n_vocab_rows)).astype(np.float32) if bias else None))
    q1, q2 = lin(True), lin(False)
    _, _, v = _heads(cfg, 10)
    eng = Engine(max_envs=48, max_input_length=64, max_new_tokens=4, tokenizer=tok)
    B200GPT2Policy(eng, pi, cfg)
    inf = VI.B200ValueRLInference(eng, base, q1, q2, v, cfg, tower=1, head_set=0)
    seqs = _seqs(cfg, 40, 1, 65, 4)
    res = inf.forward_ragged(seqs)
    worst = dict(emu=0.0, f32=0.0, v=0.0)
    for i in range(0, 40, 3):
        s, n = seqs[i], len(seqs[i])
        for emu in (True, False):
            o = og.value_forward_gathered(base, q1, q2, v, s, cfg.n_head, emulate_bf16=emu)
            for name in ("q1", "q2"):
                if n > 1:
                    worst["emu" if emu else "f32"] = max(worst["emu" if emu else "f32"], float(np.abs(res[name][i, :n - 1] - o[name]).max()))
            if emu:
                worst["v"] = max(worst["v"], float(np.abs(res["v"][i, :n] - o["v"]).max()))
    print("linear Q heads, per-position forward, max |diff| vs oracle:", {k: round(x, 5) for k, x in worst.items()})
    assert worst["emu"] <= 3e-2 and worst["f32"] <= 8e-2 and worst["v"] <= 6e-2
    assert np.abs(res["q1"] - res["q2"]).mean() > 1e-2
    eng.close()
