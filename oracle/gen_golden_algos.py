"""Generate tests/golden/{score_fns,ppo_data}.json by RUNNING THE REFERENCE's own algorithm code (build container only).

    python -m oracle.gen_golden_algos            # needs /root/reference (read-only)

What runs is the reference, unmodified, imported from /root/reference through ``oracle/ref_shims.py`` (jax.numpy -> numpy):
  * ``build_ilql_score_fn`` (LLM_RL/algorithms/ilql/gpt2/score_fn.py:11-68) and ``build_mc_score_fn``
    (mc_returns/score_fn.py:11-60) over toy "inference" objects whose Q / V / logits come from small seeded tables;
  * ``PPOInference.get_ppo_data_from_token_trajectory_chain`` (ppo/base_interface.py:464-669) called as an unbound function on
    a stand-in ``self`` whose ``forward`` returns table-driven logits and values; plus ``get_advantages_and_returns``,
    ``whiten``, ``get_action_state_next_state_idxs``, ``ppo_loss_fn`` and ``AdaptiveKLController`` on seeded inputs.
The tables and inputs are stored next to the outputs, so the tests need nothing but the JSON.  Test infrastructure only.
"""
from __future__ import annotations

import json
import os
import types

import numpy as np

from oracle import ref_shims

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
V = 32
PAD = V - 1


class ToyTokenizer:
    """One id per character (ord % 31); id 31 is the pad."""
    pad_token_id = PAD

    def encode(self, text):
        return [ord(c) % (V - 1) for c in text]


def fl(a):
    return np.asarray(a, dtype=np.float64).round(9).tolist() if np.asarray(a).dtype.kind == "f" else np.asarray(a).tolist()


def f32list(a):
    return [float(np.float32(x)) for x in np.asarray(a).reshape(-1)]


# ---------------------------------------------------------------------------------------------------------------------
def gen_score_fns():
    from LLM_RL.environment import Text
    from LLM_RL.algorithms.ilql.gpt2.score_fn import build_ilql_score_fn
    from LLM_RL.algorithms.mc_returns.score_fn import build_mc_score_fn
    rng = np.random.default_rng(7)
    tabs = {k: rng.standard_normal((V, V)).astype(np.float32) for k in ("q1", "q2", "tq1", "tq2", "pi")}
    tabs["v"] = rng.standard_normal(V).astype(np.float32)

    def table_q(tab, batch):      # [b, t, V]: row of the current token, drifting with the position
        b = np.asarray(batch)
        t = np.arange(b.shape[1], dtype=np.float32)[None, :, None]
        return (tab[b] * (1.0 + 0.01 * t)).astype(np.float32)

    def table_v(batch):
        b = np.asarray(batch)
        return (tabs["v"][b] + 0.02 * np.arange(b.shape[1], dtype=np.float32)[None]).astype(np.float32)

    class VO(types.SimpleNamespace):
        pass

    class ILQLToy:                # ILQLForwardOutput: .output (online q1, q2, v) and .target_output (target q1, q2)
        def forward(self, batch):
            return VO(output=VO(q1=table_q(tabs["q1"], batch), q2=table_q(tabs["q2"], batch), v=table_v(batch)),
                      target_output=VO(q1=table_q(tabs["tq1"], batch), q2=table_q(tabs["tq2"], batch), v=None))

    class BareToy:                # a ValueRLForwardOutput directly: the reference's AttributeError fallback
        def forward(self, batch):
            return VO(q1=table_q(tabs["q1"], batch), q2=table_q(tabs["q2"], batch), v=table_v(batch))

    class PiToy:
        def get_logits_from_tokens(self, batch):
            return table_q(tabs["pi"], batch)

    tok = ToyTokenizer()
    words = ["move left\n", "move right\n", "up\n", "a", "stay here now\n", "zz top\n"]
    obs = ["You see a wall.\n", "ok\n", "The goal is near, keep going!\n", "x"]
    histories = []
    for i in range(14):
        h = []
        for j in range(1 + i % 4):
            h.append((obs[(i + j) % len(obs)], False))
            if j < i % 4:
                h.append((words[(i * 3 + j) % len(words)], True))
        h.append((words[i % len(words)], True))
        histories.append(h)
    cases = []
    for kind, inf in (("ilql_target", ILQLToy()), ("ilql_bare", BareToy()), ("mc", BareToy())):
        for max_length, vw, lw, bsize in ((128, 1.0, None, 4), (128, 0.7, 0.3, 5), (24, 1.0, 2.0, 3), (9, 1.5, None, 14)):
            build = build_mc_score_fn if kind == "mc" else build_ilql_score_fn
            fn = build(inf, PiToy() if lw is not None else None, tok, max_length, vw, lw, bsize)
            ths = [tuple(Text(t, a) for t, a in h) for h in histories]
            out = fn(ths)
            cases.append(dict(kind=kind, max_length=max_length, value_weight=vw, logit_weight=lw, bsize=bsize, scores=f32list(out)))
    return dict(vocab=V, pad=PAD, tables={k: f32list(v) for k, v in tabs.items()}, histories=histories, cases=cases,
                note="q[b,t,:] = tab[tok[b,t]] * (1 + 0.01 t); v[b,t] = vtab[tok[b,t]] + 0.02 t; pi logits like q with the 'pi' table")


# ---------------------------------------------------------------------------------------------------------------------
def make_chains(rng, n_chains, max_traj, long=False):
    """Random token-trajectory chains shaped like rollouts: [state tokens, action tokens]* per trajectory, reward on the last token of
    every action, later trajectories of a chain start with a state."""
    chains = []
    for c in range(n_chains):
        chain = []
        n_traj = 1 + int(rng.integers(0, max_traj))
        for k in range(n_traj):
            toks, act, rew = [], [], []
            for seg in range(1 + int(rng.integers(0, 3))):
                ns = 1 + int(rng.integers(0, 6 if not long else 12))
                na = 1 + int(rng.integers(0, 4))
                toks += rng.integers(0, V - 1, ns).tolist(); act += [False] * ns; rew += [0.0] * ns
                toks += rng.integers(0, V - 1, na).tolist(); act += [True] * na
                rew += [0.0] * (na - 1) + [float(np.float32(rng.normal()))]
            if rng.random() < 0.3:      # trailing state text (e.g. the final observation)
                ns = 1 + int(rng.integers(0, 3))
                toks += rng.integers(0, V - 1, ns).tolist(); act += [False] * ns; rew += [0.0] * ns
            done = bool(k == n_traj - 1 and rng.random() < 0.6)
            chain.append(dict(tokens=toks, is_action=act, reward=rew, done=done))
        chains.append(chain)
    return chains


def gen_ppo():
    import LLM_RL.algorithms.ppo.base_interface as B
    from LLM_RL.environment import TokenTrajectory, TokenTrajectoryChain
    rng = np.random.default_rng(11)
    tabs = dict(init=rng.standard_normal((V, V)).astype(np.float32), pol=None, val=rng.standard_normal(V).astype(np.float32))
    tabs["pol"] = (tabs["init"] + 0.3 * rng.standard_normal((V, V))).astype(np.float32)

    def logits_of(tab, batch):
        b = np.asarray(batch)
        t = np.arange(b.shape[1], dtype=np.float32)[None, :, None]
        return (tab[b] * (1.0 + 0.01 * t)).astype(np.float32)

    class FakeSelf:
        initial_policy_model = types.SimpleNamespace(config=types.SimpleNamespace(mesh=None))
        policy_model = types.SimpleNamespace(config=types.SimpleNamespace(mesh=None))
        initial_policy_params = object()
        tokenizer = ToyTokenizer()
        token_logprobs_from_logits = staticmethod(B.PPOInference.token_logprobs_from_logits)

        def forward(self, tokens_batch, train=False, prng_key=None):
            b = np.asarray(tokens_batch)
            vals = (tabs["val"][b] + 0.02 * np.arange(b.shape[1], dtype=np.float32)[None]).astype(np.float32)
            return types.SimpleNamespace(initial_policy_raw_output=types.SimpleNamespace(logits=logits_of(tabs["init"], b)),
                                         policy_raw_output=types.SimpleNamespace(logits=logits_of(tabs["pol"], b)), values=vals)

    def to_ref_chain(chain):
        nxt = None
        for t in reversed(chain):
            nxt = TokenTrajectoryChain(TokenTrajectory(np.asarray(t["tokens"], np.int32), np.asarray(t["is_action"], np.bool_),
                                                       np.asarray(t["reward"], np.float32), np.asarray(t["done"], np.bool_)), nxt)
        return nxt

    cases = []
    for name, n_chains, max_traj, max_length, bsize, gamma, lam, klw, wh in (
            ("basic", 6, 3, None, 4, 1.0, 0.95, 0.001, True), ("nowhiten", 5, 2, None, 3, 0.99, 0.9, 0.05, False),
            ("single", 1, 1, None, 8, 1.0, 0.95, 0.0, True), ("padded", 8, 4, 64, 5, 0.9, 1.0, 0.01, True)):
        chains = make_chains(rng, n_chains, max_traj)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            datas, kls = B.PPOInference.get_ppo_data_from_token_trajectory_chain(
                FakeSelf(), [to_ref_chain(c) for c in chains], bsize=bsize, max_length=max_length, verbose=False, gamma=gamma, lam=lam,
                kl_weight=klw, use_advantage_whitening=wh)
        cases.append(dict(name=name, chains=chains, max_length=max_length, bsize=bsize, gamma=gamma, lam=lam, kl_weight=klw, whiten=wh,
                          all_kls=f32list(kls),
                          ppo_data=[dict(input_ids=np.asarray(d.input_ids).tolist(), should_take_action=np.asarray(d.should_take_action).astype(int).tolist(),
                                         old_logprobs=f32list(d.old_logprobs), old_values=f32list(d.old_values),
                                         old_advantages=f32list(d.old_advantages), old_returns=f32list(d.old_returns)) for d in datas]))
    # the pure functions on seeded inputs
    gae = []
    for b, n, gamma, lam, wh in ((1, 7, 1.0, 0.95, False), (3, 12, 0.99, 0.9, True), (2, 1, 0.5, 0.5, False), (4, 33, 1.0, 1.0, True)):
        sv, nv, r = (rng.standard_normal((b, n)).astype(np.float32) for _ in range(3))
        adv, ret = B.get_advantages_and_returns(sv, nv, r, gamma=gamma, lam=lam, use_whitening=wh)
        gae.append(dict(gamma=gamma, lam=lam, whiten=wh, shape=[b, n], state_values=f32list(sv), next_state_values=f32list(nv), action_rewards=f32list(r),
                        advantages=f32list(adv), returns=f32list(ret)))
    idxs = []
    for pat in ([0, 0, 1, 1, 0, 1], [1, 1, 1], [0, 0, 0], [0, 1, 0, 0, 1, 1, 0], [1]):
        a, s, nx = B.get_action_state_next_state_idxs(np.asarray(pat, dtype=np.bool_))
        idxs.append(dict(should_take_action=pat, action=a.tolist(), state=s.tolist(), next_state=nx.tolist()))
    losses = []
    for b, t, cv, cr, vc in ((4, 9, 0.2, 0.2, 1.0), (2, 17, 0.1, 0.3, 0.5), (1, 3, 0.2, 0.2, 1.0)):
        am = (rng.random((b, t)) < 0.85).astype(np.float32)
        sta = rng.random((b, t)) < 0.5
        sta[0, 0] = True; am[0, 0] = 1.0
        lp, olp = -np.abs(rng.standard_normal((b, t))).astype(np.float32), None
        olp = (lp + 0.2 * rng.standard_normal((b, t))).astype(np.float32)
        vals, ovals, oadv, oret = (rng.standard_normal((b, t)).astype(np.float32) for _ in range(4))
        loss, logs = B.ppo_loss_fn(am, lp, vals, sta, olp, ovals, oadv, oret, cliprange_value=cv, cliprange=cr, value_loss_coef=vc)
        losses.append(dict(shape=[b, t], cliprange_value=cv, cliprange=cr, value_loss_coef=vc, attention_mask=f32list(am), should_take_action=sta.astype(int).reshape(-1).tolist(),
                           logprobs=f32list(lp), old_logprobs=f32list(olp), values=f32list(vals), old_values=f32list(ovals), old_advantages=f32list(oadv),
                           old_returns=f32list(oret), loss=float(loss), policy_loss=float(logs["losses"]["policy_loss"]), value_loss=float(logs["losses"]["value_loss"]),
                           approx_kl=float(logs["policy"]["approx_kl"]), pg_clipfrac=float(logs["policy"]["clipfrac"]), vf_clipfrac=float(logs["values"]["clipfrac"]),
                           values_error=float(logs["values"]["values_error"]), ratio=float(logs["ratio"]), padding_percentage=float(logs["padding_percentage"])))
    kl = []
    ctl = B.AdaptiveKLController(0.001, 0.1, 10000)
    for cur, n in ((0.05, 128), (0.5, 128), (0.1, 64), (0.0, 256)):
        ctl.update(cur, n)
        kl.append(dict(current=cur, n_steps=n, value=ctl.value))
    return dict(vocab=V, pad=PAD, tables={k: f32list(v) for k, v in tabs.items()}, cases=cases, gae=gae, idxs=idxs, losses=losses, kl_controller=kl,
                note="logits[b,t,:] = tab[tok[b,t]] * (1 + 0.01 t); values[b,t] = vtab[tok[b,t]] + 0.02 t")


def gen_blocking():
    """``PPOData.block`` / ``PPODataset`` / ``PPOIterableDataset`` of the reference (ppo/data.py:9-158) on ragged seeded items under the
    four padding x truncation strategies.  JaxSeq's ``block_sequences`` underneath is the restatement in ref_shims (JaxSeq is not vendored);
    what this pins is the reference's composition: pad values, the max_length - 1 rule, dtypes, dataset indexing."""
    from LLM_RL.algorithms.ppo.data import PPOData, PPODataset, PPOIterableDataset
    from JaxSeq.utils import BlockingStrategy, Padding, Truncation
    rng = np.random.default_rng(11)
    items = []
    for n in (5, 9, 2, 12, 7, 8):
        f = lambda: rng.standard_normal(n - 1).astype(np.float32)  # noqa: E731
        items.append(PPOData(rng.integers(0, PAD, n).astype(np.int32), rng.random(n - 1) < 0.5, f(), f(), f(), f()))
    tok = ToyTokenizer()
    cases = []
    for pad in (Padding.RIGHT, Padding.LEFT):
        for trunc in (Truncation.RIGHT, Truncation.LEFT):
            for max_length in (8, 16):
                bs = BlockingStrategy(pad, trunc, max_length)
                blocked = PPOData.block(items, bs, tok)
                ds = PPODataset.from_ppo_data_list(items, tok, bs)
                it, src = [], iter(PPOIterableDataset.from_ppo_data_iterable(items, tok, bs))     # the reference's iterator has __next__ only
                while True:
                    try:
                        it.append(next(src))
                    except StopIteration:
                        break
                cases.append(dict(padding=pad.name, truncation=trunc.name, max_length=max_length,
                                  blocked={k: fl(np.asarray(v)) for k, v in blocked.items()}, dtypes={k: str(np.asarray(v).dtype) for k, v in blocked.items()},
                                  dataset_len=len(ds), dataset_rows_1_3={k: fl(np.asarray(v)) for k, v in ds[np.asarray([1, 3])].items()},
                                  iterable_item_4={k: fl(np.asarray(v)) for k, v in it[4].items()}, iterable_len=len(it)))
    return dict(pad=PAD, items=[{k: fl(getattr(d, k)) for k in d._fields} for d in items], cases=cases)


def main():
    import sys
    ref_shims.install()
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    for name, fn in (("score_fns", gen_score_fns), ("ppo_data", gen_ppo), ("ppo_blocking", gen_blocking)):
        if only and name not in only:
            continue
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(fn(), f, separators=(",", ":"))
        print("wrote", name, os.path.getsize(os.path.join(OUT, name + ".json")), "bytes")


if __name__ == "__main__":
    main()
