"""ILQL / MC advantage score functions: the oracle restatement AND the product's host logic (lmrl_gym_b200/reranker.py, with a
table-driven stand-in for the device forward) against vectors produced by the reference's own build_ilql_score_fn /
build_mc_score_fn (tests/golden/score_fns.json, written by oracle/gen_golden_algos.py)."""
import json
import os

import numpy as np
import pytest

from oracle import score_fns as O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "score_fns.json")))
V, PAD = G["vocab"], G["pad"]
TAB = {k: np.asarray(v, np.float32).reshape((V, V) if len(v) == V * V else (V,)) for k, v in G["tables"].items()}


def enc(text):
    return [ord(c) % (V - 1) for c in text]


class Tok:
    pad_token_id = PAD
    encode = staticmethod(enc)


def table_q(tab, batch):
    t = np.arange(batch.shape[1], dtype=np.float32)[None, :, None]
    return (tab[batch] * (1.0 + 0.01 * t)).astype(np.float32)


def table_v(batch):
    return (TAB["v"][batch] + 0.02 * np.arange(batch.shape[1], dtype=np.float32)[None]).astype(np.float32)


def token_histories():
    full, pre = [], []
    for h in G["histories"]:
        full.append([i for t, _ in h for i in enc(t)])
        pre.append(len([i for t, _ in h[:-1] for i in enc(t)]))
    return full, pre


@pytest.mark.parametrize("case", G["cases"], ids=lambda c: f"{c['kind']}-{c['max_length']}-{c['logit_weight']}")
def test_oracle_matches_reference(case):
    full, pre = token_histories()
    kind = "mc" if case["kind"] == "mc" else "ilql"
    qk = ("tq1", "tq2") if case["kind"] == "ilql_target" else ("q1", "q2")
    got = O.score_histories(kind, full, pre, PAD, case["max_length"], case["value_weight"], case["logit_weight"], case["bsize"],
                            q_fn=lambda b: (table_q(TAB[qk[0]], b), table_q(TAB[qk[1]], b)), v_fn=table_v, pi_fn=lambda b: table_q(TAB["pi"], b))
    np.testing.assert_allclose(got, case["scores"], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("case", G["cases"], ids=lambda c: f"{c['kind']}-{c['max_length']}-{c['logit_weight']}")
def test_product_host_logic_matches_reference(case, monkeypatch):
    """reranker.build_ilql_score_fn / build_mc_score_fn with the device forward replaced by the same toy tables."""
    from lmrl_gym_b200 import reranker, value_inference as VI
    from lmrl_gym_b200.environment import Text

    def gathered(q1k, q2k, batch):
        sa = lambda tab: np.take_along_axis(table_q(TAB[tab], batch)[:, :-1], batch[:, 1:][..., None], axis=2)[..., 0]  # noqa: E731
        return VI.ValueRLGatheredOutput(q1sa=sa(q1k), q2sa=sa(q2k), v=table_v(batch))

    class Bare:
        def forward(self, batch):
            return gathered("q1", "q2", batch)

    class WithTarget:
        def forward(self, batch):
            return VI.ILQLGatheredOutput(output=gathered("q1", "q2", batch), target_output=gathered("tq1", "tq2", batch))

    def fake_logprobs(engine, tower, batch, pad, temperature=1.0):
        x = table_q(TAB["pi"], batch)
        m = x.max(-1, keepdims=True)
        lp = x - m - np.log(np.exp(x - m).sum(-1, keepdims=True))
        return np.take_along_axis(lp[:, :-1], batch[:, 1:][..., None], axis=2)[..., 0]

    monkeypatch.setattr(VI, "token_logprobs", fake_logprobs)

    class Pi:
        _lmrl_engine = None

    inf = WithTarget() if case["kind"] == "ilql_target" else Bare()
    build = reranker.build_mc_score_fn if case["kind"] == "mc" else reranker.build_ilql_score_fn
    fn = build(inf, Pi() if case["logit_weight"] is not None else None, Tok, case["max_length"], case["value_weight"], case["logit_weight"], case["bsize"])
    got = fn([tuple(Text(t, a) for t, a in h) for h in G["histories"]])
    np.testing.assert_allclose(got, case["scores"], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("seed", range(6))
def test_product_host_logic_equals_oracle_on_random_histories(seed, monkeypatch):
    """Random histories / weights / truncation lengths beyond the reference-generated cases: reranker's host logic (device forward replaced
    by the toy tables) == the reference-pinned oracle score function."""
    import random
    from lmrl_gym_b200 import reranker, value_inference as VI
    from lmrl_gym_b200.environment import Text
    rng = random.Random(seed)
    alphabet = "abcdefghij klmnop\n"
    hists = []
    for _ in range(rng.randint(1, 9)):
        h = []
        for k in range(rng.randint(0, 3)):
            h.append(("".join(rng.choice(alphabet) for _ in range(rng.randint(1, 12))), False))
            h.append(("".join(rng.choice(alphabet) for _ in range(rng.randint(1, 6))), True))
        h.append(("".join(rng.choice(alphabet) for _ in range(rng.randint(1, 12))), False))
        h.append(("".join(rng.choice(alphabet) for _ in range(rng.randint(1, 6))), True))
        hists.append(h)
    kind = ("ilql", "mc")[seed % 2]
    max_length, vw = rng.choice([8, 20, 64]), rng.choice([1.0, 0.3])
    lw = rng.choice([None, 0.5])
    bsize = rng.randint(1, 5)

    def gathered(batch):
        sa = lambda tab: np.take_along_axis(table_q(TAB[tab], batch)[:, :-1], batch[:, 1:][..., None], axis=2)[..., 0]  # noqa: E731
        return VI.ValueRLGatheredOutput(q1sa=sa("q1"), q2sa=sa("q2"), v=table_v(batch))

    class Inf:
        def forward(self, batch):
            return gathered(batch)

    def fake_logprobs(engine, tower, batch, pad, temperature=1.0):
        x = table_q(TAB["pi"], batch)
        m = x.max(-1, keepdims=True)
        lp = x - m - np.log(np.exp(x - m).sum(-1, keepdims=True))
        return np.take_along_axis(lp[:, :-1], batch[:, 1:][..., None], axis=2)[..., 0]

    monkeypatch.setattr(VI, "token_logprobs", fake_logprobs)

    class Pi:
        _lmrl_engine = None

    build = reranker.build_mc_score_fn if kind == "mc" else reranker.build_ilql_score_fn
    fn = build(Inf(), Pi() if lw is not None else None, Tok, max_length, vw, lw, bsize)
    got = fn([tuple(Text(t, a) for t, a in h) for h in hists])
    full = [[i for t, _ in h for i in enc(t)] for h in hists]
    pre = [len([i for t, _ in h[:-1] for i in enc(t)]) for h in hists]
    want = O.score_histories(kind, full, pre, PAD, max_length, vw, lw, bsize, q_fn=lambda b: (table_q(TAB["q1"], b), table_q(TAB["q2"], b)), v_fn=table_v,
                             pi_fn=lambda b: table_q(TAB["pi"], b))
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)


def test_bc_score_fn_follows_the_reference_truncation_indexing(monkeypatch):
    """ppo/score_fn.py:86,118: rows are LEFT-truncated to max_length but sliced with the UNtruncated prefix length, and an empty prefix
    selects `[-1:]` of the padded row.  A numpy restatement of those two lines over a table of per-token log-probs is the oracle; the
    device call is stubbed with the same table (the BC / PPO score fn used to shift the prefix by the cut)."""
    import numpy as np
    from lmrl_gym_b200 import reranker as RR
    from lmrl_gym_b200.environment import Text

    class Tok:
        pad_token_id = 0

        def encode(self, s):
            return [ord(c) for c in s]

    class Pol:
        class _lmrl_engine:
            tokenizer = Tok()

    def lp(prev, nxt):                                   # a fixed "model": log p(next | prev) depends on the pair only
        return -0.01 * ((prev * 7 + nxt * 13) % 97) - 0.5

    def stub(policy, seqs, prefix_lens, **kw):
        return np.asarray([sum(lp(s[t - 1], s[t]) for t in range(p, len(s))) for s, p in zip(seqs, prefix_lens)], np.float32)

    monkeypatch.setattr(RR, "score_token_sequences", stub)

    def reference(h, max_length):
        full = [t for it in h for t in Tok().encode(it.text)]
        pre = len([t for it in h[:-1] for t in Tok().encode(it.text)])
        row = full[-max_length:] + [0] * (max_length - min(len(full), max_length))
        mask = [float(t != 0) for t in row]
        action_logits = [lp(row[j], row[j + 1]) * mask[j + 1] for j in range(max_length - 1)]
        return float(np.sum(action_logits[(pre - 1):], dtype=np.float64))

    cases = [
        (Text("abcdefgh", False), Text("xyz", True)),                     # fits: the three action tokens
        (Text("abcdefghijklmnop", False), Text("xyz", True)),             # truncated: prefix_len - 1 lands past the row's end -> 0.0
        (Text("abcdefghijk", False), Text("xy", True)),                   # truncated by 1: the slice starts one token late
        (Text("", False), Text("hello", True)),                           # empty prefix, short row: [-1:] is padding -> 0.0
        (Text("", False), Text("abcdefghijklmnopq", True)),               # empty prefix, full row: the last column only
        (Text("ab", False), Text("c", True), Text("de", False), Text("fgh", True)),
    ]
    fn = RR.build_bc_score_fn(Pol(), max_length=12)
    got = fn(cases)
    want = [reference(h, 12) for h in cases]
    assert np.allclose(got, want, atol=1e-6), (got, want)
    assert got[1] == 0.0 and got[3] == 0.0 and got[2] != 0.0 and got[4] != 0.0
