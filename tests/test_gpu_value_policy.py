"""GPU parity of the ILQL / MC value-perturbed policy against the torch-fp32 oracle restatement of
value_rl_base/gpt2/generation.py:36-121 and heads/mlp_head.py:139-148.

Tolerances: the perturbation beta*q is an O(1)-magnitude sum over `hidden` bf16 activations (relative precision 2^-9),
where the LM logits of a random-init GPT-2 are O(0.5); one-ulp differences of the value tower's hidden state flip the
bf16 rounding of ~20% of the relu activations (ulp 0.03-0.06 at beta*relu ~ 8), i.e. ~0.02 rms / ~0.1 worst case over
50 257 x 25 logits of magnitude ~17 (dense2 bias -4.4 * beta).  Bounds for beta = 4 with lecun-normal heads:
  * vs the oracle that rounds where the engine rounds:  max |diff| <= 1.2e-1
  * vs the plain fp32 oracle:                            max |diff| <= 2.5e-1, mean <= 3e-2
Sampled ids must be the arg-max of the oracle's (logits/T + Gumbel noise) up to twice the emulated tolerance.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from lmrl_gym_b200.gpt2 import GPT2Config, init_gpt2_params, init_mlp_head_params  # noqa: E402
from lmrl_gym_b200.tokenizer import ByteTokenizer  # noqa: E402
from lmrl_gym_b200.vocab import load_vocab_words  # noqa: E402

TOL_EMU, TOL_F32, TOL_F32_MEAN = 1.2e-1, 2.5e-1, 3e-2
BETA = 4.0


@pytest.fixture(scope="module")
def tok():
    return ByteTokenizer.synthetic()


def _cfg():
    return GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)


def _params(cfg, two_heads=True, seed=0):
    rng = np.random.default_rng(seed + 100)
    q1 = init_mlp_head_params(cfg, seed=seed + 2)
    q1["dense1"]["bias"] = (0.1 * rng.standard_normal(cfg.d_model)).astype(np.float32)
    q2 = init_mlp_head_params(cfg, seed=seed + 3) if two_heads else None
    return (init_gpt2_params(cfg, seed=seed), init_gpt2_params(cfg, seed=seed + 1), q1, q2)


@pytest.mark.parametrize("two_heads,with_pi_beta", [(True, True), (False, True), (True, False), (False, False)])
def test_value_policy_logits_match_oracle(tok, two_heads, with_pi_beta):
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2ValuePolicy
    from oracle import gpt2 as og
    cfg = _cfg()
    params = _params(cfg, two_heads)
    if not with_pi_beta:
        params = (None,) + params[1:]          # the pi_beta = None variant: logits = beta * min q (generation.py:118-119)
    eng = Engine(max_envs=160, max_input_length=64, max_new_tokens=4, tokenizer=tok)
    pol = B200GPT2ValuePolicy(eng, params, cfg, beta=BETA)
    rng = np.random.default_rng(0)
    prompts = [rng.integers(0, cfg.n_vocab, size=int(n)).tolist() for n in rng.integers(1, 64, size=150)]
    got = pol.logits(prompts)
    worst_emu = worst_f32 = 0.0
    for i in range(0, 150, 6):
        emu = og.value_policy_last_logits(params, prompts[i], cfg.n_head, cfg.n_vocab, BETA, emulate_bf16=True)
        f32 = og.value_policy_last_logits(params, prompts[i], cfg.n_head, cfg.n_vocab, BETA)
        worst_emu = max(worst_emu, np.abs(got[i] - emu).max())
        worst_f32 = max(worst_f32, np.abs(got[i] - f32).max())
        assert np.abs(got[i] - f32).mean() <= TOL_F32_MEAN
    print(f"value policy ({'ILQL 2 heads' if two_heads else 'MC 1 head'}): max|logit diff| vs bf16-emulating oracle {worst_emu:.4g}, vs fp32 oracle {worst_f32:.4g}")
    assert worst_emu <= TOL_EMU and worst_f32 <= TOL_F32
    # the perturbation is really there: the engine's logits differ from pi_beta's own by about beta * q
    if with_pi_beta:
        pi_only = og.forward_last_logits(params[0], prompts[0], cfg.n_head, cfg.n_vocab, emulate_bf16=True)
        assert np.abs(got[0] - pi_only).mean() > 1.0
    eng.close()


def test_value_policy_generate_and_set_params(tok):
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2ValuePolicy
    from oracle import gpt2 as og
    cfg = _cfg()
    params = _params(cfg, True, seed=10)
    eng = Engine(max_envs=64, max_input_length=32, max_new_tokens=5, tokenizer=tok)
    pol = B200GPT2ValuePolicy(eng, params, cfg, beta=BETA, seed=5)
    rng = np.random.default_rng(2)
    prompts = [rng.integers(0, cfg.n_vocab, size=int(n)).tolist() for n in rng.integers(1, 32, size=40)]
    keys = [int(k) for k in rng.integers(0, 2**63, size=40)]

    def check_gens(p):
        gens = pol.generate_tokens(prompts, episode_keys=keys, turn=1)
        for i in range(0, 40, 5):
            for k, t in enumerate(gens[i]):
                logits = og.value_policy_last_logits(p, prompts[i] + gens[i][:k], cfg.n_head, cfg.n_vocab, BETA, emulate_bf16=True)
                s = og.sample_scores(logits, og.row_key(5, keys[i], 1, k), 1.0, True)
                assert s[t] >= s.max() - 2 * TOL_EMU, (i, k)
        return gens

    g0 = check_gens(params)
    params2 = _params(cfg, True, seed=20)
    pol.set_params(params2)
    g1 = check_gens(params2)
    assert g0 != g1
    eng.close()


def test_value_policy_fused_rollout_replays_through_oracle_env(tok):
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2ValuePolicy
    from lmrl_gym_b200.wordle import B200WordleEnv
    from lmrl_gym_b200.environment import text_env_eval
    from oracle import gpt2 as og
    from oracle.mt19937 import seed_stream
    from oracle.wordle import ReformatWordleOracle
    cfg = _cfg()
    params = _params(cfg, True, seed=30)
    B = 128
    eng = Engine(max_envs=B, max_input_length=128, max_new_tokens=4, max_turns=6, tokenizer=tok)
    env = B200WordleEnv(eng, "wordle_official.txt")
    pol = B200GPT2ValuePolicy(eng, params, cfg, beta=BETA, seed=3)
    inter, summary = text_env_eval(env, pol, n_rollouts=B, seed_generator=seed_stream(0), bsize=B, verbose=False)
    words = load_vocab_words("wordle_official.txt")
    seeds = seed_stream(0)
    seed_list = []
    for ep in inter:
        seed = next(seeds)
        seed_list.append(seed)
        o = ReformatWordleOracle(words)
        h = o.reset(seed)
        for tr in ep:
            h, r, d = o.step(h + ((tr.post_action_history[-1].text, True),))
            assert [[x.text, x.is_action] for x in tr.post_transition_history] == [[a, b] for a, b in h]
            assert tr.reward == r and tr.done == d
    for i in (0, 77):
        ep = inter[i]
        logs, j = ep._logs, ep._i
        for t, tr in enumerate(ep):
            prompt = tok.encode("".join(x.text for x in tr.pre_action_history))[-128:]
            gen = logs.act_tokens[j, t, :logs.act_len[j, t]].tolist()
            for k, tk in enumerate(gen):
                logits = og.value_policy_last_logits(params, prompt + gen[:k], cfg.n_head, cfg.n_vocab, BETA, emulate_bf16=True)
                s = og.sample_scores(logits, og.row_key(3, seed_list[i], t, k), 1.0, True)
                assert s[tk] >= s.max() - 2 * TOL_EMU, (i, t, k)
    eng.close()


def _linear_head(cfg, seed, bias=True):
    rng = np.random.default_rng(seed)
    k = (rng.standard_normal((cfg.d_model, cfg.n_vocab_rows)) / np.sqrt(cfg.d_model)).astype(np.float32)
    return dict(dense=dict(kernel=k, bias=(0.3 * rng.standard_normal(cfg.n_vocab_rows)).astype(np.float32) if bias else None))


@pytest.mark.parametrize("n_heads,with_pi_beta,bias", [(1, True, True), (2, True, True), (1, False, False)],
                         ids=["mc-one-linear-head", "two-linear-heads", "no-pi-beta-no-bias"])
def test_value_policy_with_linear_q_heads(tok, n_heads, with_pi_beta, bias):
    """The MC-returns policy: the Q head is a LinearHead over the value base's last hidden state (wordle/mc/eval_mc_gpt2.py:93-107,
    heads/linear_head.py:112-119): logits = pi_beta + beta * (h W + b)."""
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2ValuePolicy
    from oracle import gpt2 as og
    cfg = _cfg()
    q1 = _linear_head(cfg, 21, bias)
    q2 = _linear_head(cfg, 22, bias) if n_heads == 2 else None
    params = (init_gpt2_params(cfg, seed=30) if with_pi_beta else None, init_gpt2_params(cfg, seed=31), q1, q2)
    eng = Engine(max_envs=64, max_input_length=48, max_new_tokens=4, tokenizer=tok)
    pol = B200GPT2ValuePolicy(eng, params, cfg, beta=BETA, seed=3)
    rng = np.random.default_rng(1)
    prompts = [rng.integers(0, cfg.n_vocab, size=int(n)).tolist() for n in rng.integers(1, 48, size=40)]
    got = pol.logits(prompts)
    worst_emu = worst_f32 = 0.0
    for i in range(0, 40, 4):
        emu = og.value_policy_last_logits(params, prompts[i], cfg.n_head, cfg.n_vocab, BETA, emulate_bf16=True)
        f32 = og.value_policy_last_logits(params, prompts[i], cfg.n_head, cfg.n_vocab, BETA)
        worst_emu = max(worst_emu, np.abs(got[i] - emu).max())
        worst_f32 = max(worst_f32, np.abs(got[i] - f32).max())
    print(f"linear Q heads ({n_heads}): max|logit diff| vs bf16-emulating oracle {worst_emu:.4g}, vs fp32 oracle {worst_f32:.4g}")
    assert worst_emu <= TOL_EMU and worst_f32 <= TOL_F32
    if n_heads == 2:                                    # min over the heads, not one of them
        one = og.value_policy_last_logits(params[:3] + (None,), prompts[0], cfg.n_head, cfg.n_vocab, BETA, emulate_bf16=True)
        assert np.abs(got[0] - one).max() > 0.5
    # decode path (KV caches of both towers) + sampler: every sampled id is the oracle's arg-max up to tolerance
    keys = [int(k) for k in rng.integers(0, 2**63, size=40)]
    gens = pol.generate_tokens(prompts, episode_keys=keys, turn=2)
    for i in range(0, 40, 8):
        for k, t in enumerate(gens[i]):
            logits = og.value_policy_last_logits(params, prompts[i] + gens[i][:k], cfg.n_head, cfg.n_vocab, BETA, emulate_bf16=True)
            sc = og.sample_scores(logits, og.row_key(3, keys[i], 2, k))
            assert sc[t] >= sc.max() - 2 * TOL_EMU, (i, k)
    # set_params with heads of the other kind is refused (buffers are laid out per kind)
    with pytest.raises(Exception):
        pol.set_params(_params(cfg, n_heads == 2))
    eng.close()


@pytest.mark.parametrize("beta", [16.0, 32.0])
@pytest.mark.parametrize("heads", ["ilql-two-mlp-heads", "mc-linear-head"])
def test_value_policy_at_reference_operating_point(tok, beta, heads):
    """GPT-2-small shape at the reference's own beta (wordle/ilql/eval_ilql_gpt2.py:47 policy_beta = 16; the sweeps go to 32), two MLP Q
    heads (ILQL) and one LinearHead (MC returns).  beta is a power of two, so folding it into the bf16 operand is exact; the error of
    beta * q is 2^-9 relative to |beta * q| ~ beta * 4.4, hence bounds that scale with beta / 4 from the beta = 4 ones above.  What the
    sampler sees is the ORDER of the logits: the arg-max of the engine's logits must be the oracle's arg-max up to that bound on every
    row, and the greedy-token agreement rate with the plain fp32 oracle is reported and bounded below."""
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2ValuePolicy
    from oracle import gpt2 as og
    cfg = GPT2Config.small()
    rng = np.random.default_rng(5)
    if heads == "ilql-two-mlp-heads":
        q1, q2 = init_mlp_head_params(cfg, seed=12), init_mlp_head_params(cfg, seed=13)
        q1["dense1"]["bias"] = (0.1 * rng.standard_normal(cfg.d_model)).astype(np.float32)
    else:
        q1, q2 = _linear_head(cfg, 21), None
    params = (init_gpt2_params(cfg, seed=40), init_gpt2_params(cfg, seed=41), q1, q2)
    eng = Engine(max_envs=32, max_input_length=48, max_new_tokens=3, tokenizer=tok)
    pol = B200GPT2ValuePolicy(eng, params, cfg, beta=beta, seed=3)
    prompts = [rng.integers(0, cfg.n_vocab, size=int(n)).tolist() for n in rng.integers(3, 48, size=24)]
    prompts[0] = tok.encode("Wordle:\nc r a n e\nb y b b g\n")
    got = pol.logits(prompts)
    tol_emu, tol_f32 = TOL_EMU * beta / 4.0, TOL_F32 * beta / 4.0
    worst_emu = worst_f32 = 0.0
    agree = []
    for i in range(0, 24, 2):
        emu = og.value_policy_last_logits(params, prompts[i], cfg.n_head, cfg.n_vocab, beta, emulate_bf16=True)
        f32 = og.value_policy_last_logits(params, prompts[i], cfg.n_head, cfg.n_vocab, beta)
        worst_emu = max(worst_emu, np.abs(got[i] - emu).max())
        worst_f32 = max(worst_f32, np.abs(got[i] - f32).max())
        agree.append(int(np.argmax(got[i]) == np.argmax(f32)))
        assert f32[np.argmax(got[i])] >= f32.max() - 2 * tol_f32, i          # the engine's greedy token is within tolerance of the oracle's best
        assert emu[np.argmax(got[i])] >= emu.max() - 2 * tol_emu, i
    scale = float(np.abs(f32).mean())
    print(f"{heads} beta={beta:g} GPT-2-small: max|logit diff| vs bf16-emulating oracle {worst_emu:.4g} (bound {tol_emu:.3g}), vs fp32 oracle "
          f"{worst_f32:.4g} (bound {tol_f32:.3g}); mean |logit| {scale:.1f}; greedy-token agreement with the fp32 oracle {np.mean(agree):.2f}")
    assert worst_emu <= tol_emu and worst_f32 <= tol_f32
    assert np.mean(agree) >= 0.5
    # decode path with both KV caches at this shape: sampled ids are the oracle's arg-max within the bound
    keys = [int(k) for k in rng.integers(0, 2**63, size=24)]
    gens = pol.generate_tokens(prompts, episode_keys=keys, turn=1)
    for i in (0, 7, 19):
        for k, t in enumerate(gens[i]):
            logits = og.value_policy_last_logits(params, prompts[i] + gens[i][:k], cfg.n_head, cfg.n_vocab, beta, emulate_bf16=True)
            sc = og.sample_scores(logits, og.row_key(3, keys[i], 1, k))
            assert sc[t] >= sc.max() - 2 * tol_emu, (i, k)
    eng.close()
