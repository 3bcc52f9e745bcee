"""Re-ranker (scoring) policies on the device engine — SURVEY.md §8(f) rank 1.

Mirrors ``LLM_RL/algorithms/ppo/reranker_policy.py`` (``ReRankerPolicy`` :22-32, ``ReRankerSamplePolicy`` :5-20) and the
log-probability score functions ``build_bc_score_fn`` / ``build_ppo_score_fn`` (``ppo/score_fn.py:10-126``): a proposal
function expands a history into candidate histories (e.g. ``maze_proposal_function``, maze/env/env.py:101-102: the four
moves), every candidate is scored by the sum of the log-probabilities of its action tokens under the policy
(teacher-forced, one prefill-only pass: ``lmrl_score_host``), and the best (or a softmax sample) is returned.  Used by
the Maze drivers' "move accuracy" evaluation (maze/bc/eval_bc.py:182-212).

The advantage-based score functions of ILQL / MC (``ilql/gpt2/score_fn.py``, ``mc_returns/score_fn.py``: Q - V of the
value heads) are not built; a ``B200GPT2ValuePolicy`` is scored with the log-softmax of its perturbed logits.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence

import numpy as np

from ._lib import check, ptr
from .environment import Text, TextHistory, TextPolicy

MAZE_ACTIONS = ("move left\n", "move right\n", "move up\n", "move down\n")    # manhatten_actions key order, maze/env/env.py:94-99


def maze_proposal_function(text_history: TextHistory) -> List[TextHistory]:
    """maze/env/env.py:101-102."""
    return [tuple(text_history) + (Text(a, True),) for a in MAZE_ACTIONS]


def token_history(text_history: TextHistory, tokenizer) -> List[int]:
    """``TokenHistory.from_text_history`` (LLM_RL/environment.py:303-327): every Text item is tokenised on its own."""
    out: List[int] = []
    for item in text_history:
        out.extend(tokenizer.encode(item.text))
    return out


def score_token_sequences(policy, seqs: Sequence[Sequence[int]], prefix_lens: Sequence[int], temperature: float = 1.0,
                          return_token_logprobs: bool = False):
    """sum_{t >= prefix} log p(token_t | tokens_<t) for every sequence, on the policy's engine."""
    e = policy._lmrl_engine
    n = len(seqs)
    out = np.zeros(n, np.float32)
    per_tok: List[np.ndarray] = [None] * n
    cap = e.max_envs
    for lo in range(0, n, cap):
        chunk = seqs[lo:lo + cap]
        pl = np.asarray(prefix_lens[lo:lo + cap], dtype=np.int32)
        pitch = max(8, max(len(s) for s in chunk))
        toks = np.zeros((len(chunk), pitch), np.int32)
        lens = np.zeros(len(chunk), np.int32)
        for i, s in enumerate(chunk):
            toks[i, :len(s)] = s
            lens[i] = len(s)
        sums = np.zeros(len(chunk), np.float32)
        n_scored = int(np.maximum(lens - pl, 0).sum())
        terms = np.zeros(max(n_scored, 1), np.float32)
        check(e.lib.lmrl_score_host(e.handle, ptr(toks), ptr(lens), pitch, len(chunk), ptr(pl), C.c_float(temperature), ptr(sums), ptr(terms)))
        out[lo:lo + len(chunk)] = sums
        o = 0
        for i in range(len(chunk)):
            k = max(int(lens[i] - pl[i]), 0)
            per_tok[lo + i] = terms[o:o + k].copy()
            o += k
    return (out, per_tok) if return_token_logprobs else out


def build_bc_score_fn(policy, max_length: Optional[int] = None, bsize: Optional[int] = None) -> Callable[[List[TextHistory]], List[float]]:
    """``build_bc_score_fn`` / ``build_ppo_score_fn`` (ppo/score_fn.py:69-126, 10-67) bound to a device policy: the score of a
    history whose last item is an action is the summed log-probability of that action's tokens.  ``max_length`` LEFT-truncates the
    token history like the reference's ``tokens[-max_length:]`` (the prefix length shrinks with it); ``bsize`` is accepted for
    signature compatibility (the engine batches by its own capacity)."""
    tok = policy._lmrl_engine.tokenizer

    def score_fn(text_histories: List[TextHistory]) -> List[float]:
        assert all(h[-1].is_action for h in text_histories)
        seqs, prefix = [], []
        for h in text_histories:
            full = token_history(h, tok)
            pre = len(token_history(h[:-1], tok))
            if max_length is not None and len(full) > max_length:
                cut = len(full) - max_length
                full, pre = full[cut:], max(pre - cut, 1)
            seqs.append(full)
            prefix.append(max(pre, 1))
        return [float(x) for x in score_token_sequences(policy, seqs, prefix)]

    return score_fn


build_ppo_score_fn = build_bc_score_fn


class ReRankerPolicy(TextPolicy):
    """ppo/reranker_policy.py:22-32."""

    def __init__(self, proposal_fn: Callable[[TextHistory], List[TextHistory]], score_fn: Callable[[List[TextHistory]], List[float]]):
        self.proposal_fn = proposal_fn
        self.score_fn = score_fn

    def act(self, text_history: TextHistory) -> TextHistory:
        proposals = self.proposal_fn(text_history)
        scores = self.score_fn(proposals)
        return proposals[np.argmax(np.asarray(scores, dtype=np.float32)).item()]


class ReRankerSamplePolicy(TextPolicy):
    """ppo/reranker_policy.py:5-20 (softmax sample over the scores; ``rng`` replaces the reference's global ``np.random``)."""

    def __init__(self, proposal_fn, score_fn, rng: Optional[np.random.Generator] = None):
        self.proposal_fn = proposal_fn
        self.score_fn = score_fn
        self.rng = rng

    def act(self, text_history: TextHistory) -> TextHistory:
        proposals = self.proposal_fn(text_history)
        scores = np.asarray(self.score_fn(proposals), dtype=np.float32)
        p = np.exp(scores) / np.exp(scores).sum()
        selected = (self.rng.choice(len(p), p=p) if self.rng is not None else np.random.choice(len(p), p=p))
        return proposals[int(selected)]
