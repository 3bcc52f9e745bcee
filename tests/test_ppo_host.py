"""Host logic of lmrl_gym_b200/ppo.py (chain combination, flattening, PPOData slicing) against the reference-generated vectors of
tests/golden/ppo_data.json, with the two device calls replaced: the forward by the fixture's table model, the advantage kernels by the
oracle's numpy functions.  The device kernels themselves are checked on the GPU in tests/test_gpu_ppo.py."""
import json
import os

import numpy as np
import pytest

from oracle import ppo as O
from tests.test_oracle_ppo import G, PAD, table_forward

from lmrl_gym_b200 import ppo as P
from lmrl_gym_b200.environment import Text, TextTrajectory, TextTrajectoryChain


def ref_chain(c):
    nxt = None
    for t in reversed(c):
        nxt = P.TokenTrajectoryChain(P.TokenTrajectory(np.asarray(t["tokens"], np.int32), np.asarray(t["is_action"], np.bool_),
                                                       np.asarray(t["reward"], np.float32), np.asarray(t["done"], np.bool_)), nxt)
    return nxt


def numpy_advantages(chain_off, plp, ilp, val, rew, sta, *, gamma, lam, kl_weight, whiten):
    """What lmrl_ppo_advantages_host computes, via the oracle's functions (flattened layout of include/lmrl_b200.h)."""
    N = int(chain_off[-1])
    adv, ret, kls = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros(N, np.float32)
    for c in range(len(chain_off) - 1):
        o0, o1 = int(chain_off[c]), int(chain_off[c + 1])
        s = sta[o0:o1].astype(bool)
        v = val[o0 + c:o1 + c + 1]
        lr = (plp[o0:o1] - ilp[o0:o1]) * s.astype(np.float32)
        r = rew[o0:o1] - kl_weight * lr
        a, st, nx = O.action_state_next_state_idxs(s)
        A, R = O.advantages_and_returns(v[st][None], v[nx][None], r[a][None], gamma, lam, use_whitening=False)
        adv[o0:o1][a] = A[0]; ret[o0:o1][a] = R[0]
        kls[o0:o1][a] = (np.exp(lr) - 1 - lr)[a]
    if whiten:
        m = sta.astype(bool)
        adv[m] = O.whiten(adv[m])
    return adv, ret, kls


class FakeInference(P.B200PPOInference):
    def __init__(self):
        self.pad_token_id = PAD
        self.has_initial_policy = True


@pytest.mark.parametrize("case", G["cases"], ids=lambda c: c["name"])
def test_ppo_data_host_logic(case):
    inf = FakeInference()
    datas, kls = inf.get_ppo_data_from_token_trajectory_chain([ref_chain(c) for c in case["chains"]], bsize=case["bsize"], max_length=case["max_length"],
                                                              gamma=case["gamma"], lam=case["lam"], kl_weight=case["kl_weight"],
                                                              use_advantage_whitening=case["whiten"], _forward=table_forward, _advantages=numpy_advantages)
    np.testing.assert_allclose(kls, case["all_kls"], rtol=1e-5, atol=1e-7)
    assert len(datas) == len(case["ppo_data"])
    for d, g in zip(datas, case["ppo_data"]):
        assert d.input_ids.tolist() == g["input_ids"]
        assert np.asarray(d.should_take_action).astype(int).tolist() == g["should_take_action"]
        for k in ("old_logprobs", "old_values", "old_advantages", "old_returns"):
            np.testing.assert_allclose(getattr(d, k), g[k], rtol=2e-5, atol=2e-6, err_msg=k)


def test_token_types_from_text():
    class Tok:
        @staticmethod
        def encode(s):
            return [ord(c) for c in s]
    chain = TextTrajectoryChain(TextTrajectory((Text("ab", False), Text("cde", True)), (0.0, -1.0), False),
                                TextTrajectoryChain(TextTrajectory((Text("ab", False), Text("cde", True), Text("f", False), Text("gh", True)), (0.0, -1.0, 0.0, 2.0), True), None))
    tc = P.TokenTrajectoryChain.from_text_trajectory_chain(chain, Tok)
    l = tc.to_list()
    assert len(l) == 2 and l[0].tokens.tolist() == [97, 98, 99, 100, 101]
    assert l[0].is_action.tolist() == [False, False, True, True, True] and l[0].reward.tolist() == [0, 0, 0, 0, -1.0]
    assert l[1].reward.tolist() == [0, 0, 0, 0, -1.0, 0, 0, 2.0] and bool(l[1].done)
    th = P.TokenHistory.from_text_history(chain.text_trajectory.text_history, Tok)
    assert th.tokens.tolist() == [97, 98, 99, 100, 101] and th.is_action.sum() == 3
    cc = P.CombinedTokenTrajectoryChain.from_token_trajectory_chain(tc)
    assert cc.chunk_lens == [4, 7] and cc.should_take_action.tolist() == [False, True, True, True, False, True, True, True, False, True, True]
    assert cc.rewards.tolist() == [0, 0, 0, -1.0, 0, 0, 0, -1.0, 0, 0, 2.0]
    with pytest.raises(AssertionError):
        P.CombinedTokenTrajectoryChain.from_token_trajectory_chain(tc, max_length=3)     # would cut an action


def test_kl_controller():
    ctl = P.AdaptiveKLController(0.001, 0.1, 10000)
    for g in G["kl_controller"]:
        ctl.update(g["current"], g["n_steps"])
        assert ctl.value == pytest.approx(g["value"], rel=1e-12)


def _random_chains(rng, n_chains):
    """Rollout-shaped chains: every later trajectory of a chain extends the previous one (the history grows), rewards on action ends."""
    chains = []
    for _ in range(n_chains):
        toks, act, rew = [], [], []
        chain = []
        n_traj = 1 + int(rng.integers(0, 4))
        for k in range(n_traj):
            ns, na = 1 + int(rng.integers(0, 7)), 1 + int(rng.integers(0, 5))
            toks += rng.integers(0, 30, ns).tolist(); act += [False] * ns; rew += [0.0] * ns
            toks += rng.integers(0, 30, na).tolist(); act += [True] * na; rew += [0.0] * (na - 1) + [float(np.float32(rng.normal()))]
            done = bool(k == n_traj - 1 and rng.random() < 0.5)
            chain.append(dict(tokens=list(toks), is_action=list(act), reward=list(rew), done=done))
        chains.append(chain)
    return chains


@pytest.mark.parametrize("seed", range(8))
def test_ppo_data_host_logic_equals_oracle_on_random_chains(seed):
    """Beyond the four reference-generated cases: random growing-history chains, with and without whitening / padding — the product's
    chain combination + flattening + slicing (device calls stubbed) must equal the reference-pinned oracle on the same table model."""
    rng = np.random.default_rng(seed)
    chains = _random_chains(rng, 1 + int(rng.integers(0, 9)))
    longest = max(len(t["tokens"]) for c in chains for t in c)
    max_length = None if seed % 2 else longest + int(rng.integers(1, 5))
    kw = dict(gamma=float(rng.choice([1.0, 0.99, 0.9])), lam=float(rng.choice([1.0, 0.95, 0.5])), kl_weight=float(rng.choice([0.0, 0.001, 0.1])))
    whiten = bool(seed % 3)
    inf = FakeInference()
    datas, kls = inf.get_ppo_data_from_token_trajectory_chain([ref_chain(c) for c in chains], bsize=3, max_length=max_length, use_advantage_whitening=whiten,
                                                              _forward=table_forward, _advantages=numpy_advantages, **kw)
    o_chains = [[(np.asarray(t["tokens"], np.int32), np.asarray(t["is_action"], bool), np.asarray(t["reward"], np.float32), bool(t["done"])) for t in c] for c in chains]
    o_datas, o_kls = O.ppo_data_from_chains(o_chains, table_forward, PAD, 3, max_length, use_advantage_whitening=whiten, **kw)
    np.testing.assert_allclose(kls, o_kls, rtol=1e-6, atol=1e-7)
    assert len(datas) == len(o_datas)
    for d, o in zip(datas, o_datas):
        assert d.input_ids.tolist() == o.input_ids.tolist() and np.asarray(d.should_take_action).tolist() == o.should_take_action.tolist()
        for k in ("old_logprobs", "old_values", "old_advantages", "old_returns"):
            np.testing.assert_allclose(getattr(d, k), getattr(o, k), rtol=2e-6, atol=2e-6, err_msg=k)


def test_ppo_data_blocking_and_datasets_match_reference(golden_dir):
    """``PPOData.block`` / ``PPODataset`` / ``PPOIterableDataset`` (ppo/data.py:9-158) against the reference's own code run on the same
    items (tests/golden/ppo_blocking.json, oracle/gen_golden_algos.py::gen_blocking): values, shapes and dtypes for the four
    padding x truncation strategies."""
    import json
    import os
    from lmrl_gym_b200 import ppo as P
    g = json.load(open(os.path.join(golden_dir, "ppo_blocking.json")))

    class Tok:
        pad_token_id = g["pad"]
    dt = dict(input_ids=np.int32, should_take_action=np.bool_, old_logprobs=np.float32, old_values=np.float32, old_advantages=np.float32,
              old_returns=np.float32)
    items = [P.PPOData(**{k: np.asarray(v, dtype=dt[k]) for k, v in it.items()}) for it in g["items"]]
    assert len(g["cases"]) == 8
    for case in g["cases"]:
        bs = P.BlockingStrategy(P.Padding[case["padding"]], P.Truncation[case["truncation"]], case["max_length"])
        blocked = P.PPOData.block(items, bs, Tok())
        for k, want in case["blocked"].items():
            want = np.asarray(want)
            assert blocked[k].shape == want.shape and str(blocked[k].dtype) == case["dtypes"][k], (k, blocked[k].dtype)
            np.testing.assert_allclose(blocked[k].astype(np.float64), want.astype(np.float64), rtol=0, atol=1e-9)
        ds = P.PPODataset.from_ppo_data_list(items, Tok(), bs)
        assert len(ds) == case["dataset_len"]
        rows = ds[np.asarray([1, 3])]
        for k, want in case["dataset_rows_1_3"].items():
            np.testing.assert_allclose(rows[k].astype(np.float64), np.asarray(want, dtype=np.float64), rtol=0, atol=1e-9)
            assert rows[k].dtype == dt[k]
        stream = list(P.PPOIterableDataset.from_ppo_data_iterable(items, Tok(), bs))
        assert len(stream) == case["iterable_len"]
        for k, want in case["iterable_item_4"].items():
            np.testing.assert_allclose(stream[4][k].astype(np.float64), np.asarray(want, dtype=np.float64), rtol=0, atol=1e-9)
    # the trainer's own blocking (RIGHT / RIGHT) is the same function
    from lmrl_gym_b200.ppo_train import block_ppo_data
    mine = block_ppo_data(items, g["pad"], 8)
    ref = P.PPOData.block(items, P.BlockingStrategy(P.Padding.RIGHT, P.Truncation.RIGHT, 8), Tok())
    for k in ref:
        np.testing.assert_array_equal(mine[k], ref[k])
    # max_length None: the longest sequence
    free = P.PPOData.block(items, P.BlockingStrategy(P.Padding.LEFT, P.Truncation.LEFT, None), Tok())
    assert free["input_ids"].shape == (6, 12) and free["old_values"].shape == (6, 11)
