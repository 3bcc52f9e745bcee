"""PPO data path oracle (oracle/ppo.py) against vectors produced by the reference's own functions
(tests/golden/ppo_data.json, written by oracle/gen_golden_algos.py from LLM_RL/algorithms/ppo/base_interface.py)."""
import json
import os

import numpy as np
import pytest

from oracle import ppo as O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ppo_data.json")))
V, PAD = G["vocab"], G["pad"]
TAB = {k: np.asarray(v, np.float32).reshape((V, V) if len(v) == V * V else (V,)) for k, v in G["tables"].items()}


def logits_of(tab, b):
    t = np.arange(b.shape[1], dtype=np.float32)[None, :, None]
    return (tab[b] * (1.0 + 0.01 * t)).astype(np.float32)


def token_logprobs(logits, tokens):
    x = logits[:, :-1].astype(np.float32)
    m = x.max(-1, keepdims=True)
    lse = np.log(np.exp(x - m).sum(-1, dtype=np.float32)) + m[..., 0]
    return (np.take_along_axis(x, tokens[:, 1:][..., None], axis=-1)[..., 0] - lse).astype(np.float32)


def table_forward(tokens):
    vals = (TAB["val"][tokens] + 0.02 * np.arange(tokens.shape[1], dtype=np.float32)[None]).astype(np.float32)
    return token_logprobs(logits_of(TAB["init"], tokens), tokens), token_logprobs(logits_of(TAB["pol"], tokens), tokens), vals


def to_chain(c):
    return [(np.asarray(t["tokens"], np.int32), np.asarray(t["is_action"], bool), np.asarray(t["reward"], np.float32), bool(t["done"])) for t in c]


@pytest.mark.parametrize("case", G["cases"], ids=lambda c: c["name"])
def test_ppo_data_matches_reference(case):
    chains = [to_chain(c) for c in case["chains"]]
    datas, kls = O.ppo_data_from_chains(chains, table_forward, PAD, case["bsize"], case["max_length"], gamma=case["gamma"], lam=case["lam"],
                                        kl_weight=case["kl_weight"], use_advantage_whitening=case["whiten"])
    np.testing.assert_allclose(kls, case["all_kls"], rtol=1e-5, atol=1e-7)
    assert len(datas) == len(case["ppo_data"])
    for d, g in zip(datas, case["ppo_data"]):
        assert d.input_ids.tolist() == g["input_ids"]
        assert d.should_take_action.astype(int).tolist() == g["should_take_action"]
        for k in ("old_logprobs", "old_values", "old_advantages", "old_returns"):
            np.testing.assert_allclose(getattr(d, k), g[k], rtol=2e-5, atol=2e-6, err_msg=k)


@pytest.mark.parametrize("g", G["gae"], ids=lambda g: "x".join(map(str, g["shape"])))
def test_gae_matches_reference(g):
    sh = g["shape"]
    sv, nv, r = (np.asarray(g[k], np.float32).reshape(sh) for k in ("state_values", "next_state_values", "action_rewards"))
    adv, ret = O.advantages_and_returns(sv, nv, r, g["gamma"], g["lam"], use_whitening=g["whiten"])
    if not g["whiten"]:
        assert adv.reshape(-1).tolist() == g["advantages"]          # the scan itself is bit-exact
    np.testing.assert_allclose(adv.reshape(-1), g["advantages"], rtol=1e-6, atol=1e-6)
    assert ret.reshape(-1).tolist() == g["returns"]


def test_idxs_match_reference():
    for g in G["idxs"]:
        a, s, n = O.action_state_next_state_idxs(np.asarray(g["should_take_action"], bool))
        assert (a.tolist(), s.tolist(), n.tolist()) == (g["action"], g["state"], g["next_state"])


@pytest.mark.parametrize("g", G["losses"], ids=lambda g: "x".join(map(str, g["shape"])))
def test_loss_matches_reference(g):
    sh = g["shape"]
    a = lambda k: np.asarray(g[k], np.float32).reshape(sh)  # noqa: E731
    loss, logs = O.ppo_loss(a("attention_mask"), a("logprobs"), a("values"), np.asarray(g["should_take_action"]).reshape(sh).astype(bool), a("old_logprobs"),
                            a("old_values"), a("old_advantages"), a("old_returns"), cliprange_value=g["cliprange_value"], cliprange=g["cliprange"],
                            value_loss_coef=g["value_loss_coef"])
    assert abs(float(loss) - g["loss"]) <= 1e-6 * max(1.0, abs(g["loss"]))
    for k_mine, k_ref in (("policy_loss", "policy_loss"), ("value_loss", "value_loss"), ("approx_kl", "approx_kl"), ("pg_clipfrac", "pg_clipfrac"),
                          ("vf_clipfrac", "vf_clipfrac"), ("values_error", "values_error"), ("ratio", "ratio"), ("padding_percentage", "padding_percentage")):
        assert abs(logs[k_mine] - g[k_ref]) <= 1e-6 * max(1.0, abs(g[k_ref])), k_mine


def test_kl_controller_matches_reference():
    ctl = O.AdaptiveKLController(0.001, 0.1, 10000)
    for g in G["kl_controller"]:
        ctl.update(g["current"], g["n_steps"])
        assert ctl.value == pytest.approx(g["value"], rel=1e-12)
