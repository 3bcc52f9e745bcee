"""Teacher-forced per-position forward on the device engine: value heads and token log-probabilities.

Host mirrors of the reference's inference objects for the paths either side of the rollout:

    ValueRLInference.forward      LLM_RL/algorithms/value_rl_base/base_interface.py:129-168 (+ gpt2/interface.py:150-215)
    ILQLInference.forward         LLM_RL/algorithms/ilql/base_interface.py:313-344  (value_inference + target_value_inference)
    GPT2Inference.get_logits_from_tokens -> log_softmax -> take_along_axis   (ilql/gpt2/score_fn.py:60-64)

The reference materialises ``q1, q2`` as ``[batch, time, vocab]`` arrays and then gathers the taken token
(``take_along_axis(qs[:, :-1], batch[:, 1:])``); every consumer in the repository (score functions, ILQL / MC losses)
only ever uses that gather.  The engine therefore returns the gathered values directly (``q1sa[b, t] = Q1(h_t)[tok_{t+1}]``)
— one dense2 column per position is read instead of a vocabulary-wide GEMM (``csrc/gpt2_kernels.cuh::value_heads_kernel``).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, NamedTuple, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check, ptr
from .engine import Engine
from .gpt2 import GPT2Config, _upload_params, head_struct


def upload_value_heads(engine: Engine, head_set: int, q1: Optional[Dict] = None, q2: Optional[Dict] = None, v: Optional[Dict] = None,
                       v_linear: Optional[Dict] = None, role: str = "ILQL / MC value heads") -> None:
    """q1 / q2 / v: flax ``MLPHead`` params ``{'dense1': {'kernel','bias'}, 'dense2': {...}}`` (heads/mlp_head.py:99-148); q1 / q2 may
    instead be ``LinearHead`` params ``{'dense': {'kernel' [d, n_out], 'bias'}}`` (the MC-returns Q head);
    v_linear: flax ``LinearHead`` params ``{'dense': {'kernel' [d, 1], 'bias' [1] (optional)}}`` (heads/linear_head.py:91-119)."""
    engine.claim_slot("value-head set", int(head_set), role)
    keep: list = []
    hs = _lib.lmrl_value_heads()
    for name, q in (("q1", q1), ("q2", q2), ("v", v)):
        s = head_struct(q, keep)
        if s is not None:
            setattr(hs, name, C.pointer(s))
    if v_linear is not None:
        w = np.ascontiguousarray(np.asarray(v_linear["dense"]["kernel"], dtype=np.float32).reshape(-1))
        keep.append(w)
        hs.v_linear_w = w.ctypes.data_as(_lib._PF)
        if v_linear["dense"].get("bias") is not None:
            b = np.ascontiguousarray(np.asarray(v_linear["dense"]["bias"], dtype=np.float32).reshape(-1))
            keep.append(b)
            hs.v_linear_b = b.ctypes.data_as(_lib._PF)
    check(engine.lib.lmrl_load_value_heads(engine.handle, int(head_set), C.byref(hs)))


def forward_tokens(engine: Engine, tower: int, head_set: int, seqs: Sequence[Sequence[int]], *, logprobs: bool = False, v: bool = False,
                   q1: bool = False, q2: bool = False, temperature: float = 1.0, pitch: Optional[int] = None) -> Dict[str, np.ndarray]:
    """One teacher-forced pass of ``tower`` over every sequence (each starts at position 0).  Returns ``[n, pitch]`` fp32 arrays:
    ``logprobs[i, t] = log p(seq[t+1] | seq[:t+1])``, ``q1 / q2[i, t] = Q(h_t)[seq[t+1]]`` (t < len-1), ``v[i, t]`` (t < len);
    cells past a sequence's end are 0."""
    n = len(seqs)
    lens = np.asarray([len(s) for s in seqs], dtype=np.int32)
    pitch = int(pitch if pitch is not None else max(8, int(lens.max()) if n else 8))
    toks = np.zeros((n, pitch), np.int32)
    for i, s in enumerate(seqs):
        toks[i, :len(s)] = np.asarray(s, dtype=np.int32)
    res: Dict[str, np.ndarray] = {}
    out = _lib.lmrl_forward_out()
    for name, want, field in (("logprobs", logprobs, "token_logprobs"), ("v", v, "v"), ("q1", q1, "q1"), ("q2", q2, "q2")):
        if want:
            res[name] = np.zeros((n, pitch), np.float32)
            setattr(out, field, ptr(res[name]))
    check(engine.lib.lmrl_forward_host(engine.handle, int(tower), int(head_set), ptr(toks), ptr(lens), pitch, n, C.c_float(temperature), C.byref(out)))
    res["n_tok"] = lens
    return res


def _seqs_from_padded(tokens: np.ndarray, pad_token_id: int) -> List[np.ndarray]:
    """RIGHT-padded ``[b, t]`` id matrix -> ragged sequences (the reference derives the attention mask from ``tokens != pad``,
    JaxSeq initialize_attn_mask_pos_ids; with RIGHT padding the real tokens are the leading ones)."""
    tokens = np.asarray(tokens)
    out = []
    for row in tokens:
        real = row != pad_token_id
        n = int(real.sum())
        assert bool(real[:n].all()), "only RIGHT padding is supported by the device forward (pad id inside a sequence)"
        out.append(row[:n].astype(np.int32))
    return out


class ValueRLGatheredOutput(NamedTuple):
    """What every consumer takes from ``ValueRLForwardOutput`` (value_rl_base/base_interface.py:20-24): ``q1sa / q2sa [b, t-1]`` are
    ``take_along_axis(q[:, :-1], tokens[:, 1:])`` (0 on padding), ``v [b, t]`` the V head of every real position (0 on padding)."""
    q1sa: np.ndarray
    q2sa: Optional[np.ndarray]
    v: Optional[np.ndarray]


class B200ValueRLInference:
    """``ValueRLInference`` (base tower + q1 / q2 / v heads) on one engine tower + head set.

    tower 1 / head set 0 for the online networks, tower 2 / head set 1 for the targets (``ILQLInference.target_value_inference``);
    ``pad_token_id`` defaults to the tokenizer's (50257 for GPT-2 + <|pad|>)."""

    def __init__(self, engine: Engine, base_params: Dict, q1_head_params: Dict, q2_head_params: Optional[Dict] = None,
                 v_head_params: Optional[Dict] = None, config: Optional[GPT2Config] = None, tower: int = 1, head_set: int = 0,
                 pad_token_id: Optional[int] = None):
        self._lmrl_engine = engine
        self.config = config or GPT2Config.small()
        self.tower, self.head_set = int(tower), int(head_set)
        self.has_q2, self.has_v = q2_head_params is not None, v_head_params is not None
        self.pad_token_id = int(pad_token_id if pad_token_id is not None else self.config.n_vocab)
        _upload_params(engine, self.config, base_params, tower=self.tower)
        upload_value_heads(engine, self.head_set, q1=q1_head_params, q2=q2_head_params, v=v_head_params)

    def forward_ragged(self, seqs: Sequence[Sequence[int]], pitch: Optional[int] = None) -> Dict[str, np.ndarray]:
        return forward_tokens(self._lmrl_engine, self.tower, self.head_set, seqs, q1=True, q2=self.has_q2, v=self.has_v, pitch=pitch)

    def forward(self, input_ids: np.ndarray) -> ValueRLGatheredOutput:
        input_ids = np.asarray(input_ids)
        r = self.forward_ragged(_seqs_from_padded(input_ids, self.pad_token_id), pitch=max(8, input_ids.shape[1]))
        t = input_ids.shape[1]
        return ValueRLGatheredOutput(q1sa=r["q1"][:, :t - 1], q2sa=r["q2"][:, :t - 1] if self.has_q2 else None, v=r["v"][:, :t] if self.has_v else None)


class ILQLGatheredOutput(NamedTuple):
    output: ValueRLGatheredOutput
    target_output: ValueRLGatheredOutput


class B200ILQLInference:
    """``ILQLInference.forward`` (ilql/base_interface.py:313-344): the online and the target value networks over the same batch."""

    def __init__(self, value_inference: B200ValueRLInference, target_value_inference: B200ValueRLInference):
        self.value_inference = value_inference
        self.target_value_inference = target_value_inference

    def forward(self, input_ids: np.ndarray) -> ILQLGatheredOutput:
        return ILQLGatheredOutput(output=self.value_inference.forward(input_ids), target_output=self.target_value_inference.forward(input_ids))


def token_logprobs(engine: Engine, tower: int, input_ids: np.ndarray, pad_token_id: int, temperature: float = 1.0) -> np.ndarray:
    """``take_along_axis(log_softmax(get_logits_from_tokens(batch))[:, :-1], batch[:, 1:])`` -> ``[b, t-1]`` (0 on padding)."""
    input_ids = np.asarray(input_ids)
    r = forward_tokens(engine, tower, 0, _seqs_from_padded(input_ids, pad_token_id), logprobs=True, temperature=temperature,
                       pitch=max(8, input_ids.shape[1]))
    return r["logprobs"][:, :input_ids.shape[1] - 1]
