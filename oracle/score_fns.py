"""ILQL / MC advantage score functions — numpy restatement of the reference (oracle; test infrastructure only).

Follows LLM_RL/algorithms/ilql/gpt2/score_fn.py:11-68 and LLM_RL/algorithms/mc_returns/score_fn.py:11-60.
PINNED: ``tests/golden/score_fns.json`` holds the outputs of the reference's own ``build_ilql_score_fn`` /
``build_mc_score_fn`` (run by ``oracle/gen_golden_algos.py`` through ``oracle/ref_shims.py``) on table-driven toy models.

The model enters through three callables over a RIGHT-padded id matrix ``batch [b, T]``:
    q_fn(batch)  -> (q1 [b, T, V], q2 [b, T, V] or None)     Q heads used for the score (target heads when the inference has them, :45-48)
    v_fn(batch)  -> v [b, T] or None                          online V head (ILQL only)
    pi_fn(batch) -> logits [b, T, V] or None                  pi_beta (only with logit_weight)
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np


def _log_softmax(x):
    x = np.asarray(x, dtype=np.float32)
    m = x.max(axis=-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=-1, keepdims=True, dtype=np.float32))


def score_histories(kind: str, token_histories: Sequence[Sequence[int]], prefix_lens: Sequence[int], pad: int, max_length: int,
                    value_weight: float, logit_weight: Optional[float], bsize: int, q_fn: Callable, v_fn: Optional[Callable],
                    pi_fn: Optional[Callable]) -> List[float]:
    """token_histories: per candidate the ids of the whole history (last item = the action); prefix_lens: ids before the action.
    Rows are LEFT-truncated to the last max_length ids and RIGHT-padded (:32); the action span is taken from ``prefix_len - 1`` of
    the UNtruncated history (:54) — after truncation that slice starts too late or is empty, exactly as in the reference."""
    tokens = np.full((len(token_histories), max_length), pad, dtype=np.int32)
    for i, t in enumerate(token_histories):
        t = list(t)[-max_length:]
        tokens[i, :len(t)] = t
    out: List[float] = []
    for i in range(0, len(tokens), bsize):
        batch = tokens[i:i + bsize]
        mask = (batch != pad).astype(np.float32)
        q1, q2 = q_fn(batch)
        qs = q1 if (kind == "mc" or q2 is None) else np.minimum(q1, q2)
        qsa = np.take_along_axis(qs[:, :-1], batch[:, 1:][..., None], axis=2)[..., 0]
        per_tok = qsa * mask[:, 1:] if kind == "mc" else (qsa - v_fn(batch)[:, :-1]) * mask[:, 1:]
        adv = np.zeros(batch.shape[0], np.float32)
        for x in range(batch.shape[0]):
            adv[x] = value_weight * per_tok[x][(prefix_lens[i + x] - 1):].sum(axis=0)
        if logit_weight is not None:
            lp = np.take_along_axis(_log_softmax(pi_fn(batch))[:, :-1], batch[:, 1:][..., None], axis=2)[..., 0] * mask[:, 1:]
            for x in range(batch.shape[0]):
                adv[x] += logit_weight * lp[x][(prefix_lens[i + x] - 1):].sum(axis=0)
        out.extend(float(a) for a in adv)
    return out
