"""GPU parity of the PPO data path: the device kernels of csrc/ppo_kernels.inl against the vectors the REFERENCE's own
code produced (tests/golden/ppo_data.json: get_ppo_data_from_token_trajectory_chain, ppo_loss_fn), at full size against the oracle, and
the whole B200PPOInference (device forward of both policies + LinearHead values + device GAE) against the torch-fp32 oracle.

Bars: the GAE scan (returns, un-whitened advantages) is BIT-EXACT (float32, reference operation order); whitened advantages and the loss
reductions within 2e-6 relative (different but deterministic summation order); kls within 1e-6 absolute (fast exp).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import ppo as O  # noqa: E402
from tests.test_oracle_ppo import G, PAD, table_forward  # noqa: E402
from tests.test_ppo_host import numpy_advantages, ref_chain  # noqa: E402

from lmrl_gym_b200 import ppo as P  # noqa: E402
from lmrl_gym_b200.gpt2 import GPT2Config, init_gpt2_params  # noqa: E402
from lmrl_gym_b200.tokenizer import ByteTokenizer  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    from lmrl_gym_b200.engine import Engine
    e = Engine(max_envs=32, max_input_length=64, max_new_tokens=4, tokenizer=ByteTokenizer.synthetic())
    yield e
    e.close()


class TableInference(P.B200PPOInference):
    """The fixture's table model as the forward, the REAL device kernels for everything after it."""

    def __init__(self, engine):
        self._lmrl_engine = engine
        self.pad_token_id = PAD
        self.has_initial_policy = True


@pytest.mark.parametrize("case", G["cases"], ids=lambda c: c["name"])
def test_device_ppo_data_matches_reference_vectors(eng, case):
    inf = TableInference(eng)
    datas, kls = inf.get_ppo_data_from_token_trajectory_chain([ref_chain(c) for c in case["chains"]], bsize=case["bsize"], max_length=case["max_length"],
                                                              gamma=case["gamma"], lam=case["lam"], kl_weight=case["kl_weight"],
                                                              use_advantage_whitening=case["whiten"], _forward=table_forward)
    np.testing.assert_allclose(kls, case["all_kls"], rtol=0, atol=1e-6)
    assert len(datas) == len(case["ppo_data"])
    for d, g in zip(datas, case["ppo_data"]):
        assert d.input_ids.tolist() == g["input_ids"]
        assert np.asarray(d.should_take_action).astype(int).tolist() == g["should_take_action"]
        np.testing.assert_array_equal(d.old_logprobs, np.asarray(g["old_logprobs"], np.float32))
        np.testing.assert_array_equal(d.old_values, np.asarray(g["old_values"], np.float32))
        np.testing.assert_array_equal(d.old_returns, np.asarray(g["old_returns"], np.float32))          # bit-exact scan
        if case["whiten"]:
            np.testing.assert_allclose(d.old_advantages, g["old_advantages"], rtol=2e-6, atol=2e-6)
        else:
            np.testing.assert_array_equal(d.old_advantages, np.asarray(g["old_advantages"], np.float32))


@pytest.mark.parametrize("g", G["losses"], ids=lambda g: "x".join(map(str, g["shape"])))
def test_device_loss_matches_reference_vectors(eng, g):
    sh = g["shape"]
    a = lambda k: np.asarray(g[k], np.float32).reshape(sh)  # noqa: E731
    loss, logs = P.ppo_loss(eng, a("attention_mask"), a("logprobs"), a("values"), np.asarray(g["should_take_action"]).reshape(sh).astype(bool),
                            a("old_logprobs"), a("old_values"), a("old_advantages"), a("old_returns"), cliprange_value=g["cliprange_value"],
                            cliprange=g["cliprange"], value_loss_coef=g["value_loss_coef"])
    for mine, ref in (("total_loss", "loss"), ("policy_loss", "policy_loss"), ("value_loss", "value_loss"), ("approx_kl", "approx_kl"),
                      ("policy_clipfrac", "pg_clipfrac"), ("value_clipfrac", "vf_clipfrac"), ("values_error", "values_error"), ("ratio", "ratio"),
                      ("padding_percentage", "padding_percentage")):
        assert abs(logs[mine] - g[ref]) <= 2e-6 * max(1.0, abs(g[ref])), (mine, logs[mine], g[ref])


def test_device_advantages_full_size_against_oracle(eng):
    """4096 chains x up to 6 turns of (observation, action) tokens — the shape of one Wordle collection round."""
    rng = np.random.default_rng(0)
    n_chains = 4096
    lens = rng.integers(8, 160, size=n_chains)
    chain_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    N = int(chain_off[-1])
    sta = (rng.random(N) < 0.45)
    sta[chain_off[:-1][:7]] = False
    sta[chain_off[3]:chain_off[4]] = False              # a chain with no action tokens at all
    plp = -np.abs(rng.standard_normal(N)).astype(np.float32)
    ilp = (plp + 0.1 * rng.standard_normal(N)).astype(np.float32)
    val = rng.standard_normal(N + n_chains).astype(np.float32)
    rew = (rng.standard_normal(N) * sta).astype(np.float32)
    for whiten in (False, True):
        adv, ret, kls = P.ppo_advantages(eng, chain_off, plp, ilp, val, rew, sta, gamma=0.99, lam=0.95, kl_weight=0.001, whiten=whiten)
        a0, r0, k0 = numpy_advantages(chain_off, plp, ilp, val, rew, sta.astype(np.uint8), gamma=0.99, lam=0.95, kl_weight=0.001, whiten=whiten)
        np.testing.assert_array_equal(ret, r0)
        np.testing.assert_allclose(kls, k0, rtol=0, atol=1e-6)
        if whiten:
            np.testing.assert_allclose(adv, a0, rtol=5e-6, atol=5e-6)
            m = sta.astype(bool)
            assert abs(float(adv[m].mean())) < 1e-4 and abs(float(adv[m].std()) - 1.0) < 1e-3      # the property whitening exists for
        else:
            np.testing.assert_array_equal(adv, a0)
        assert (adv[~sta] == 0).all() and (ret[~sta] == 0).all()


def test_ppo_inference_end_to_end_against_oracle(eng):
    """Text trajectory chains -> B200PPOInference (device forward of policy / initial policy / value head + device GAE) vs the oracle's
    ppo_data_from_chains fed by the torch-fp32 GPT-2 restatement.  Policy side unpinned -> tolerances (bf16 engine vs fp32 oracle)."""
    from lmrl_gym_b200.environment import Text, TextTrajectory, TextTrajectoryChain
    from oracle import gpt2 as og
    cfg = GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)
    tok = eng.tokenizer
    pol, ini = init_gpt2_params(cfg, seed=50), init_gpt2_params(cfg, seed=50)
    rng = np.random.default_rng(3)
    for L in pol["layers"]:                                  # the trained policy has drifted a little from the initial one
        L["fc_w"] = (L["fc_w"] + 0.01 * rng.standard_normal(L["fc_w"].shape)).astype(np.float32)
    lin = dict(dense=dict(kernel=(rng.standard_normal((cfg.d_model, 1)) / 16).astype(np.float32), bias=np.asarray([-0.5], np.float32)))
    inf = P.B200PPOInference(eng, pol, lin, initial_policy_params=ini, config=cfg, pad_token_id=tok.pad_token_id)
    words = ["crane", "slate", "qqqqq", "b o a r d", "x"]
    chains = []
    for c in range(10):
        hist, trajs = (Text("Wordle:\n", False),), []
        n_turn = 1 + c % 3
        for t in range(n_turn):
            hist = hist + (Text(" ".join(words[(c + t) % 5]) + "\n", True),)
            rew = tuple(0.0 for _ in hist[:-1]) + (-1.0 if t + 1 < n_turn or c % 2 else 0.0,)
            hist_obs = hist + (Text("b y g b b\n", False),)
            trajs.append(TextTrajectory(hist_obs, rew + (0.0,), t + 1 == n_turn and c % 4 != 3))
            hist = hist_obs
        nxt = None
        for tr in reversed(trajs):
            nxt = TextTrajectoryChain(tr, nxt)
        chains.append(nxt)
    kw = dict(gamma=1.0, lam=0.95, kl_weight=0.05)
    datas, kls = inf.get_ppo_data_from_text_trajectory_chain(chains, bsize=8, max_length=64, use_advantage_whitening=False, **kw)

    def oracle_forward(tokens):
        a, b, v = (np.zeros((tokens.shape[0], tokens.shape[1] - 1), np.float32), np.zeros((tokens.shape[0], tokens.shape[1] - 1), np.float32),
                   np.zeros(tokens.shape, np.float32))
        for i, row in enumerate(tokens):
            n = int((row != tok.pad_token_id).sum())
            s = row[:n].tolist()
            a[i, :n - 1] = og.token_logprobs_all(ini, s, cfg.n_head, cfg.n_vocab, emulate_bf16=True)
            b[i, :n - 1] = og.token_logprobs_all(pol, s, cfg.n_head, cfg.n_vocab, emulate_bf16=True)
            v[i, :n] = og.linear_value_forward(pol, lin, s, cfg.n_head, emulate_bf16=True)
        return a, b, v

    o_chains = []
    for ch in chains:
        l = P.TokenTrajectoryChain.from_text_trajectory_chain(ch, tok).to_list()
        o_chains.append([(t.tokens, t.is_action, t.reward, bool(t.done)) for t in l])
    o_datas, o_kls = O.ppo_data_from_chains(o_chains, oracle_forward, tok.pad_token_id, 8, 64, use_advantage_whitening=False, **kw)
    assert len(datas) == len(o_datas) == sum(len(c) for c in o_chains)
    worst = dict(lp=0.0, v=0.0, adv=0.0, ret=0.0)
    for d, o in zip(datas, o_datas):
        assert d.input_ids.tolist() == o.input_ids.tolist() and d.should_take_action.tolist() == o.should_take_action.tolist()
        for k, name in (("lp", "old_logprobs"), ("v", "old_values"), ("adv", "old_advantages"), ("ret", "old_returns")):
            worst[k] = max(worst[k], float(np.abs(getattr(d, name) - getattr(o, name)).max()))
    print("B200PPOInference vs oracle, max |diff|:", {k: round(v, 5) for k, v in worst.items()}, "kl max diff", float(np.abs(kls - o_kls).max()))
    assert worst["lp"] <= 3e-2 and worst["v"] <= 2e-2 and worst["ret"] <= 8e-2 and worst["adv"] <= 1e-1
    assert np.abs(kls - o_kls).max() <= 5e-3
    # the loss of the data under the policy that produced it: ratio == 1, approx_kl == 0, nothing clipped
    blk = lambda xs, pad, T, dt: P._block_right(xs, pad, T, dt)  # noqa: E731
    ids = blk([d.input_ids for d in datas], tok.pad_token_id, 64, np.int32)
    loss, logs = inf.eval_loss(ids, blk([d.should_take_action for d in datas], False, 63, np.bool_), blk([d.old_logprobs for d in datas], 0.0, 63, np.float32),
                               blk([d.old_values for d in datas], 0.0, 63, np.float32), blk([d.old_advantages for d in datas], 0.0, 63, np.float32),
                               blk([d.old_returns for d in datas], 0.0, 63, np.float32))
    assert abs(logs["ratio"] - 1.0) < 1e-6 and abs(logs["approx_kl"]) < 1e-6 and logs["policy_clipfrac"] == 0.0 and logs["value_clipfrac"] == 0.0
    adv_mean = float(np.concatenate([d.old_advantages[d.should_take_action.astype(bool)] for d in datas]).mean())
    assert abs(logs["policy_loss"] + adv_mean) < 1e-5


def test_ppo_inference_forward_from_str_and_bc_eval_loss(eng):
    """forward_from_str (ppo/base_interface.py:437-462) == forward of the separately encoded, RIGHT-blocked strings, checked against the
    oracle's log-probs; eval_loss with a BC batch adds bc_loss * bc_loss_weight and nests the logs (:451-464)."""
    from lmrl_gym_b200.ppo_train import bc_batch_from_str_segments
    from oracle import gpt2 as og
    cfg = GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)
    tok = eng.tokenizer
    pol = init_gpt2_params(cfg, seed=60)
    lin = dict(dense=dict(kernel=(np.random.default_rng(4).standard_normal((cfg.d_model, 1)) / 16).astype(np.float32), bias=np.asarray([0.2], np.float32)))
    inf = P.B200PPOInference(eng, pol, lin, initial_policy_params=pol, config=cfg, pad_token_id=tok.pad_token_id)
    strs = ["Wordle:\nc r a n e\nb y b b g\n", "Wordle:\n", "Wordle:\ns l a t e\n"]
    out = inf.forward_from_str(strs)
    T = max(len(tok.encode(x)) for x in strs)
    assert out.policy_logprobs.shape == (3, T - 1) and out.values.shape == (3, T)
    for i, x in enumerate(strs):
        ids = list(tok.encode(x))
        lp = og.token_logprobs_all(pol, ids, cfg.n_head, cfg.n_vocab, emulate_bf16=True)
        assert np.abs(out.policy_logprobs[i, :len(ids) - 1] - lp).max() <= 3e-2
        assert (out.policy_logprobs[i, len(ids) - 1:] == 0).all() and (out.values[i, len(ids):] == 0).all()
    np.testing.assert_array_equal(out.initial_policy_logprobs, out.policy_logprobs)         # same weights in both towers
    cut = inf.forward_from_str(strs, P.BlockingStrategy(P.Padding.RIGHT, P.Truncation.LEFT, 6), token_process=lambda t: t[:-1])
    assert cut.values.shape == (3, 6)
    with pytest.raises(AssertionError):
        inf.forward_from_str(strs, P.BlockingStrategy(P.Padding.LEFT, P.Truncation.LEFT, 8))
    # eval_loss with the BC term
    ids = P.block_sequences([tok.encode(x) for x in strs], tok.pad_token_id, np.int32, P.BlockingStrategy(P.Padding.RIGHT, P.Truncation.RIGHT, T))
    sta = np.zeros((3, T - 1), bool)
    sta[0, 3:9] = sta[2, 3:] = True
    zeros = np.zeros((3, T - 1), np.float32)
    args = (ids, sta, out.policy_logprobs, out.values[:, :-1], zeros + 0.5, zeros)
    base_loss, base_logs = inf.eval_loss(*args)
    bc = bc_batch_from_str_segments([[("Wordle:\n", 0.0), ("c r a n e\n", 1.0)], [("Wordle:\n", 0.0), ("a b\n", 1.0)]], tok, 16, tok.pad_token_id)
    total, logs = inf.eval_loss(*args, bc_data_input_ids=bc["input_ids"], bc_data_input_training_mask=bc["input_training_mask"], bc_loss_weight=2.0)
    want = 0.0
    for row, m in zip(bc["input_ids"], bc["input_training_mask"]):
        n = int((row != tok.pad_token_id).sum())
        want -= float((og.token_logprobs_all(pol, row[:n].tolist(), cfg.n_head, cfg.n_vocab, emulate_bf16=True) * m[1:n]).sum())
    want /= float(bc["input_training_mask"][:, 1:].sum())
    assert abs(logs["bc"]["loss"] - want) <= 3e-2 and abs(total - (base_loss + 2.0 * logs["bc"]["loss"])) <= 1e-5
    assert logs["total_loss"] == total and abs(logs["ppo"]["policy_loss"] - base_logs["policy_loss"]) <= 1e-6
    assert inf.get_ppo_data_from_text_trajectory_chain_iterble.__func__ is inf.get_ppo_data_from_text_trajectory_chain.__func__


def test_ppo_data_throughput_vs_cpu_port():
    """Measurement, not parity: the same PPO collection pass (GPT-2-small, Wordle-shaped chains) through B200PPOInference and through
    the CPU port (torch-CPU GPT-2 oracle + oracle/ppo.py on all host threads, bounded sample).  Printed for DESIGN.md; the only
    assertion is the sanity bound GPU >= 50x CPU port."""
    import os
    import time
    import torch
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    from oracle import gpt2 as og
    sys_path_scripts = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ppo_data_bench", os.path.join(sys_path_scripts, "ppo_data_bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config.small()
    rng = np.random.default_rng(0)
    e = Engine(max_envs=512, max_input_length=128, max_new_tokens=6, tokenizer=tok)
    pol, ini = init_gpt2_params(cfg, seed=0), init_gpt2_params(cfg, seed=1)
    lin = dict(dense=dict(kernel=(rng.standard_normal((cfg.d_model, 1)) / 28).astype(np.float32), bias=np.zeros(1, np.float32)))
    B200GPT2Policy(e, pol, cfg)
    inf = P.B200PPOInference(e, None, lin, initial_policy_params=ini, config=cfg, pad_token_id=tok.pad_token_id)
    chains = mod.wordle_chains(tok, 512, rng)
    kw = dict(max_length=128, gamma=1.0, lam=0.95, kl_weight=0.001, use_advantage_whitening=True)
    inf.get_ppo_data_from_token_trajectory_chain(chains[:64], bsize=1 << 20, verbose=False, **kw)
    n_tok = sum(t.tokens.shape[0] for c in chains for t in c.to_list())
    t0 = time.perf_counter()
    datas, _ = inf.get_ppo_data_from_token_trajectory_chain(chains, bsize=1 << 20, verbose=False, **kw)
    gpu = n_tok / (time.perf_counter() - t0)
    e.close()
    torch.set_num_threads(os.cpu_count() or 1)
    sub = chains[:6]

    def fwd(tokens):
        a = np.zeros((tokens.shape[0], tokens.shape[1] - 1), np.float32); b = np.zeros_like(a); vv = np.zeros(tokens.shape, np.float32)
        for i, row in enumerate(tokens):
            n = int((row != tok.pad_token_id).sum()); s = row[:n].tolist()
            a[i, :n - 1] = og.token_logprobs_all(ini, s, cfg.n_head, cfg.n_vocab)
            b[i, :n - 1] = og.token_logprobs_all(pol, s, cfg.n_head, cfg.n_vocab)
            vv[i, :n] = og.linear_value_forward(pol, lin, s, cfg.n_head)
        return a, b, vv

    oc = [[(t.tokens, t.is_action, t.reward, bool(t.done)) for t in c.to_list()] for c in sub]
    n_sub = sum(t[0].shape[0] for c in oc for t in c)
    t0 = time.perf_counter()
    o_datas, _ = O.ppo_data_from_chains(oc, fwd, tok.pad_token_id, 64, 128, gamma=1.0, lam=0.95, kl_weight=0.001)
    cpu = n_sub / (time.perf_counter() - t0)
    print(f"PPO data pass: B200 {gpu:.0f} tokens/s (512 chains, host in / host out) vs CPU port {cpu:.0f} tokens/s ({torch.get_num_threads()} threads, {len(sub)} chains)")
    for d, o in zip(datas[:len(o_datas)], o_datas):
        assert np.abs(d.old_returns - o.old_returns).max() < 0.3
    assert gpu > 50 * cpu
