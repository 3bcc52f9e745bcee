"""Re-ranker (scoring) policies on the device engine — SURVEY.md §8(f) rank 1.

Mirrors ``LLM_RL/algorithms/ppo/reranker_policy.py`` (``ReRankerPolicy`` :22-32, ``ReRankerSamplePolicy`` :5-20) and the
log-probability score functions ``build_bc_score_fn`` / ``build_ppo_score_fn`` (``ppo/score_fn.py:10-126``): a proposal
function expands a history into candidate histories (e.g. ``maze_proposal_function``, maze/env/env.py:101-102: the four
moves), every candidate is scored by the sum of the log-probabilities of its action tokens under the policy
(teacher-forced, one prefill-only pass: ``lmrl_score_host``), and the best (or a softmax sample) is returned.  Used by
the Maze drivers' "move accuracy" evaluation (maze/bc/eval_bc.py:182-212).

``build_ilql_score_fn`` / ``build_mc_score_fn`` are the advantage-based score functions of ILQL / MC
(``ilql/gpt2/score_fn.py:11-68``, ``mc_returns/score_fn.py:11-60``): value_weight * sum over the action tokens of
(min(Q1, Q2)(s, a) - V(s)) [ILQL] or Q1(s, a) [MC], plus logit_weight * the summed pi_beta log-probabilities — on the
per-position device forward (``lmrl_forward_host``, value_inference.py).
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
    history whose last item is an action is the summed log-probability of that action's tokens.

    Truncation follows the reference to the letter (as ``_advantage_score_fn`` below already did): the token history is LEFT-truncated
    to ``max_length`` (``tokens[-max_length:]``, :86) but the slice that selects the action's log-probs uses the UNtruncated prefix
    length, ``masked_action_logits[x][(prefix_len[x]-1):]`` (:118) — so under truncation fewer (possibly no) action tokens are summed,
    and an empty prefix (``prefix_len = 0`` -> ``[-1:]``) selects the last column of the padded row (0.0 unless the row is full).
    ``max_length=None`` scores untruncated rows; ``bsize`` is accepted for signature compatibility (the engine batches by its own
    capacity)."""
    tok = policy._lmrl_engine.tokenizer

    def score_fn(text_histories: List[TextHistory]) -> List[float]:
        assert all(h[-1].is_action for h in text_histories)
        seqs, prefix, where = [], [], []
        out = [0.0] * len(text_histories)
        for i, h in enumerate(text_histories):
            full = token_history(h, tok)
            pre = len(token_history(h[:-1], tok))
            width = len(full) if max_length is None else max_length          # columns of the reference's padded row
            row = full[-width:]
            start = pre - 1 if pre >= 1 else width - 2                       # first scored column of action_logits (python slice semantics)
            if start > len(row) - 2 or len(row) < 2:
                continue                                                     # the slice is empty or covers only padding: the reference sums to 0.0
            seqs.append(row)
            prefix.append(start + 1)
            where.append(i)
        if seqs:
            for i, x in zip(where, score_token_sequences(policy, seqs, prefix)):
                out[i] = float(x)
        return out

    return score_fn


build_ppo_score_fn = build_bc_score_fn


def _block_right(token_histories: Sequence[Sequence[int]], max_length: int, pad_token_id: int) -> np.ndarray:
    """``tokens[-max_length:]`` then RIGHT padding to max_length (ilql/gpt2/score_fn.py:32)."""
    out = np.full((len(token_histories), max_length), pad_token_id, dtype=np.int32)
    for i, t in enumerate(token_histories):
        t = list(t)[-max_length:]
        out[i, :len(t)] = t
    return out


def _advantage_score_fn(kind: str, inference, pi_beta_policy, tokenizer, max_length: int, value_weight: float,
                        logit_weight: Optional[float], bsize: int):
    assert (pi_beta_policy is None and logit_weight is None) or (pi_beta_policy is not None and logit_weight is not None)
    from .value_inference import token_logprobs
    pad = tokenizer.pad_token_id

    def score_fn(text_histories: List[TextHistory], done: Optional[List] = None) -> List[float]:
        assert all(h[-1].is_action for h in text_histories)
        prefix_lens = [len(token_history(h[:-1], tokenizer)) for h in text_histories]
        tokens = _block_right([token_history(h, tokenizer) for h in text_histories], max_length, pad)
        advantages: List[float] = []
        for i in range(0, len(text_histories), bsize):
            batch = tokens[i:i + bsize]
            mask = (batch != pad).astype(np.float32)
            values = inference.forward(batch)
            if kind == "ilql":
                # score_fn.py:45-48: the target networks' Q when the inference has them, else the online ones; V is always online
                tgt = getattr(values, "target_output", None)
                out = getattr(values, "output", values)
                qsrc = tgt if tgt is not None else out
                qsa = np.minimum(qsrc.q1sa, qsrc.q2sa)
                per_tok = (qsa - out.v[:, :-1]) * mask[:, 1:]
            else:
                per_tok = values.q1sa * mask[:, 1:]                      # mc_returns/score_fn.py:44-49
            advs = np.empty(batch.shape[0], np.float32)
            for x in range(batch.shape[0]):
                # the reference indexes the (possibly LEFT-truncated) row with the UNtruncated prefix length (score_fn.py:54)
                advs[x] = np.float32(value_weight) * per_tok[x][(prefix_lens[i + x] - 1):].sum(axis=0, dtype=np.float32)
            if logit_weight is not None:
                lp = token_logprobs(pi_beta_policy._lmrl_engine, 0, batch, pad) * mask[:, 1:]
                for x in range(batch.shape[0]):
                    advs[x] += np.float32(logit_weight) * lp[x][(prefix_lens[i + x] - 1):].sum(axis=0, dtype=np.float32)
            advantages.extend(advs.tolist())
        return advantages

    return score_fn


def build_ilql_score_fn(inference, pi_beta_inference, tokenizer, max_length: int, value_weight: float, logit_weight: Optional[float],
                        bsize: int):
    """``build_ilql_score_fn`` (ilql/gpt2/score_fn.py:11-68).  ``inference``: a ``B200ILQLInference`` (Q from the target networks,
    V from the online one) or a bare ``B200ValueRLInference`` (the reference's AttributeError fallback, :47-48, :57-58);
    ``pi_beta_inference``: a ``B200GPT2Policy`` on the same engine (tower 0) or None."""
    return _advantage_score_fn("ilql", inference, pi_beta_inference, tokenizer, max_length, value_weight, logit_weight, bsize)


def build_mc_score_fn(inference, pi_beta_inference, tokenizer, max_length: int, value_weight: float, logit_weight: Optional[float],
                      bsize: int):
    """``build_mc_score_fn`` (mc_returns/score_fn.py:11-60): value_weight * sum of Q1(s, a) over the action tokens."""
    return _advantage_score_fn("mc", inference, pi_beta_inference, tokenizer, max_length, value_weight, logit_weight, bsize)


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
