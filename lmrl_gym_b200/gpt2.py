"""GPT-2 sampling policy on the device engine, behind the reference's BatchedTextPolicy interface.

Drop-in for ``GPT2PPOPolicy`` (LLM_RL/algorithms/ppo/gpt2/interface.py:480-549): ``act(text_history, done)`` joins
each history into a prompt (environment.py:18), LEFT-truncates it to ``max_input_length`` tokens, samples up to
``max_new_tokens`` ids (stop at the newline/EOS id, as ``GenerationConfig(eos_token_id=tokenizer.encode('\\n')[0])``
of wordle/ppo/train_ppo_gpt2.py:223), decodes the completion and appends it as an action ``Text``;
``set_params(params)`` swaps the weights (interface.py:548-549).  When the env passed to ``text_env_eval`` is a device
env on the same ``Engine`` the whole loop is fused on the GPU instead (lmrl_gym_b200/rollout.py).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Tuple, Callable, Dict, List, Optional

import numpy as np

from . import _lib
from ._lib import check, ptr
from .engine import Engine
from .environment import BatchedTextPolicy, Text, TextHistory, text_history_to_str


@dataclass
class GPT2Config:
    n_layer: int = 12
    n_head: int = 12
    d_model: int = 768
    d_ff: int = 3072
    n_vocab: int = 50257          # ids the sampler may emit (the reference masks padded ids to -inf, interface.py:330)
    n_vocab_rows: int = 50258     # rows of wte (50257 + <|pad|>)
    n_pos: int = 1024
    ln_eps: float = 1e-5

    @classmethod
    def small(cls):
        return cls()

    @classmethod
    def medium(cls):
        return cls(n_layer=24, n_head=16, d_model=1024, d_ff=4096)

    @property
    def matmul_params(self) -> int:
        d, f = self.d_model, self.d_ff
        return self.n_layer * (3 * d * d + d * d + 2 * d * f) + d * self.n_vocab


def init_gpt2_params(cfg: GPT2Config, seed: int = 0) -> Dict:
    """Seeded random init with HF GPT-2's scheme (normal(0, 0.02); c_proj scaled by 1/sqrt(2L); LN = 1/0)."""
    rng = np.random.default_rng(seed)
    d, f, L = cfg.d_model, cfg.d_ff, cfg.n_layer

    def nrm(shape, std=0.02):
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std))

    proj_std = 0.02 / np.sqrt(2.0 * L)
    layers = []
    for _ in range(L):
        layers.append(dict(
            ln1_g=np.ones(d, np.float32), ln1_b=np.zeros(d, np.float32),
            attn_w=nrm((d, 3 * d)), attn_b=np.zeros(3 * d, np.float32),
            proj_w=nrm((d, d), proj_std), proj_b=np.zeros(d, np.float32),
            ln2_g=np.ones(d, np.float32), ln2_b=np.zeros(d, np.float32),
            fc_w=nrm((d, f)), fc_b=np.zeros(f, np.float32),
            fc2_w=nrm((f, d), proj_std), fc2_b=np.zeros(d, np.float32)))
    return dict(wte=nrm((cfg.n_vocab_rows, d)), wpe=nrm((cfg.n_pos, d)), lnf_g=np.ones(d, np.float32),
                lnf_b=np.zeros(d, np.float32), layers=layers)


def steer_to_letters(params: Dict, tokenizer, logit: float = 15.0, scale: float = 50.0, seed: int = 123) -> Dict:
    """Synthetic-weights utility (no pretrained or trained Wordle policy exists offline): make the tied LM head context-free
    and concentrated on single-letter tokens + newline, so a random-init GPT-2 emits actions with the TOKEN SHAPE of a
    trained Wordle policy (about five letters then '\n').  The final LN emits the constant vector scale*u (gamma = 0), the
    embedding rows of the 52 letter tokens ('c' and ' c') are set to (logit/scale)*u and the newline row a little higher
    so that P(newline) = 1/6 per step; every other id keeps its N(0, 0.02) row (logit ~ N(0, 1))."""
    letters = [tokenizer.encode(c)[0] for c in "abcdefghijklmnopqrstuvwxyz"] + [tokenizer.encode("x " + c)[1] for c in "abcdefghijklmnopqrstuvwxyz"]
    d = params["wte"].shape[1]
    u = np.random.default_rng(seed).standard_normal(d)
    u /= np.linalg.norm(u)
    out = dict(params)
    wte = params["wte"].copy()
    wte[letters] = (logit / scale) * u
    wte[tokenizer.newline_id()] = ((logit + np.log(len(letters) / 5.0)) / scale) * u
    out["wte"] = wte.astype(np.float32)
    out["lnf_g"] = np.zeros_like(params["lnf_g"])
    out["lnf_b"] = (scale * u).astype(np.float32)
    return out


def params_from_hf_state_dict(sd, cfg: GPT2Config) -> Dict:
    """``transformers.GPT2LMHeadModel().state_dict()`` (torch) -> the dict layout above."""
    g = lambda k: np.ascontiguousarray(sd[k].detach().cpu().float().numpy())  # noqa: E731
    layers = []
    for l in range(cfg.n_layer):
        p = f"transformer.h.{l}."
        layers.append(dict(
            ln1_g=g(p + "ln_1.weight"), ln1_b=g(p + "ln_1.bias"), attn_w=g(p + "attn.c_attn.weight"), attn_b=g(p + "attn.c_attn.bias"),
            proj_w=g(p + "attn.c_proj.weight"), proj_b=g(p + "attn.c_proj.bias"), ln2_g=g(p + "ln_2.weight"), ln2_b=g(p + "ln_2.bias"),
            fc_w=g(p + "mlp.c_fc.weight"), fc_b=g(p + "mlp.c_fc.bias"), fc2_w=g(p + "mlp.c_proj.weight"), fc2_b=g(p + "mlp.c_proj.bias")))
    return dict(wte=g("transformer.wte.weight"), wpe=g("transformer.wpe.weight"), lnf_g=g("transformer.ln_f.weight"),
                lnf_b=g("transformer.ln_f.bias"), layers=layers)


_LAYER_KEYS = ["ln1_g", "ln1_b", "attn_w", "attn_b", "proj_w", "proj_b", "ln2_g", "ln2_b", "fc_w", "fc_b", "fc2_w", "fc2_b"]


def init_mlp_head_params(cfg: GPT2Config, seed: int = 0, hidden: Optional[int] = None, n_out: Optional[int] = None,
                         layer2_bias_init: float = -4.4, init_range: Optional[float] = None) -> Dict:
    """Seeded init of an ``MLPHead`` (heads/mlp_head.py:99-137) with the ILQL script's config
    (``MLPHeadConfig(input_dim=d, hidden_dim=d, output_dim=vocab, layer2_bias_init=-4.4)``, train_ilql_gpt2.py:224-231):
    lecun-normal kernels (std = 1/sqrt(fan_in)) unless ``init_range`` is given, zero dense1 bias, constant dense2 bias."""
    rng = np.random.default_rng(seed)
    d = cfg.d_model
    hidden = hidden or d
    n_out = n_out or cfg.n_vocab_rows
    s1 = init_range if init_range is not None else 1.0 / np.sqrt(d)
    s2 = init_range if init_range is not None else 1.0 / np.sqrt(hidden)
    return dict(dense1=dict(kernel=rng.standard_normal((d, hidden), dtype=np.float32) * np.float32(s1), bias=np.zeros(hidden, np.float32)),
                dense2=dict(kernel=rng.standard_normal((hidden, n_out), dtype=np.float32) * np.float32(s2),
                            bias=np.full(n_out, layer2_bias_init, np.float32)))


def _upload_params(engine: Engine, cfg: GPT2Config, params: Dict, tower: int = 0, role: Optional[str] = None):
    engine.claim_slot("tower", int(tower), role or ("policy" if tower == 0 else "value base"))
    L = cfg.n_layer
    assert len(params["layers"]) == L
    keep = []  # keep contiguous fp32 copies alive during the call

    def arr(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        keep.append(a)
        return a.ctypes.data_as(_lib._PF)

    w = _lib.lmrl_gpt2_weights()
    w.wte = arr(params["wte"])
    w.n_vocab_rows = int(params["wte"].shape[0])
    w.wpe = arr(params["wpe"])
    w.lnf_g = arr(params["lnf_g"])
    w.lnf_b = arr(params["lnf_b"])
    for key in _LAYER_KEYS:
        ptrs = (_lib._PF * L)(*[arr(layer[key]) for layer in params["layers"]])
        keep.append(ptrs)
        setattr(w, key, C.cast(ptrs, _lib._PPF))
    dims = _lib.lmrl_gpt2_dims(n_layer=cfg.n_layer, n_head=cfg.n_head, d_model=cfg.d_model, d_ff=cfg.d_ff, n_vocab=cfg.n_vocab,
                               n_pos=cfg.n_pos, ln_eps=cfg.ln_eps)
    check(engine.lib.lmrl_load_gpt2_tower(engine.handle, int(tower), C.byref(dims), C.byref(w)))


def head_struct(q: Optional[Dict], keep: list):
    """flax head params -> ``lmrl_mlp_head`` (None stays None): an MLPHead ``{'dense1': {'kernel','bias'}, 'dense2': {...}}``
    (heads/mlp_head.py:99-148) or a LinearHead ``{'dense': {'kernel','bias'-or-None}}`` (heads/linear_head.py:91-119: w1 / b1 stay NULL, the
    kernel travels in w2).  ``keep`` receives every object whose memory the struct points into."""
    if q is None:
        return None
    h = _lib.lmrl_mlp_head()
    if "dense" in q:
        fields = (("w2", q["dense"]["kernel"]),) + ((("b2", q["dense"]["bias"]),) if q["dense"].get("bias") is not None else ())
        h.hidden, h.n_out = int(np.shape(q["dense"]["kernel"])[0]), int(np.shape(q["dense"]["kernel"])[1])
    else:
        fields = (("w1", q["dense1"]["kernel"]), ("b1", q["dense1"]["bias"]), ("w2", q["dense2"]["kernel"]), ("b2", q["dense2"]["bias"]))
        h.hidden, h.n_out = int(np.shape(q["dense1"]["kernel"])[1]), int(np.shape(q["dense2"]["kernel"])[1])
    for name, a in fields:
        a = np.ascontiguousarray(a, dtype=np.float32)
        keep.append(a)
        setattr(h, name, a.ctypes.data_as(_lib._PF))
    keep.append(h)
    return h


def _upload_q_heads(engine: Engine, q1: Dict, q2: Optional[Dict], beta: float):
    engine.claim_slot("sampler heads", 0, "Q heads of the value-perturbed policy")
    keep: list = []
    h1, h2 = head_struct(q1, keep), head_struct(q2, keep)
    check(engine.lib.lmrl_load_q_heads(engine.handle, C.byref(h1), None if h2 is None else C.byref(h2), float(beta)))


def default_out_str_process(s: str) -> str:
    # the lambda every reference Wordle/Maze driver passes (e.g. wordle/bc/eval_bc_gpt2.py:105)
    return s.removesuffix("\n") + "\n"


def nlnl_token_id(tok) -> int:
    """Id of the tokenizer's "\\n\\n" token when prompts ending in a run of k newlines encode that run as floor(k / 2) of it followed by
    k % 2 plain newlines (GPT-2's pre-tokeniser + the 'ĊĊ' merge; what the device's prompt view reproduces), else -1 (no fold: the run
    stays k newline tokens, which is exact for tokenizers without such a token)."""
    nl = tok.encode("\n")
    e2 = tok.encode("e\n\n")
    if len(nl) != 1 or len(e2) < 1 or tok.token_bytes[e2[-1]] != b"\n\n":
        return -1
    nlnl, head = int(e2[-1]), tok.encode("e")
    ok = (tok.encode("e\n\n\n") == head + [nlnl, nl[0]] and tok.encode("e\n\n\n\n") == head + [nlnl, nlnl] and
          tok.encode("e\n\nx") == head + [nl[0], nl[0]] + tok.encode("x"))
    return nlnl if ok else -1


class B200GPT2Policy(BatchedTextPolicy):
    def __init__(self, engine: Engine, params: Dict, config: Optional[GPT2Config] = None, do_sample: bool = True,
                 temperature: Optional[float] = None, seed: int = 0, in_str_process: Optional[Callable[[str], str]] = None,
                 out_str_process: Optional[Callable[[str], str]] = None, top_k: Optional[int] = None, top_p: Optional[float] = None):
        if engine.tokenizer is None:
            raise _lib.LmrlError("Engine needs a tokenizer before a policy can be bound to it")
        self._lmrl_engine = engine
        self.config = config or GPT2Config.small()
        self.do_sample = bool(do_sample)
        self.temperature = 1.0 if temperature is None else float(temperature)
        self.seed = int(seed)
        # GenerationConfig.top_k / top_p (eval_bc_gpt2.py:42-44, 90-95; None in every shipped driver): HF warper order is
        # temperature -> top_k -> top_p
        self.top_k = 0 if top_k is None else int(top_k)
        self.top_p = 0.0 if top_p is None else float(top_p)
        self.in_str_process = in_str_process
        self.out_str_process = out_str_process
        self._calls = 0
        _upload_params(engine, self.config, params)
        # joined-history tokenisation (ppo/gpt2/interface.py:519-522): a tokenizer that encodes a trailing "\n\n" as ONE token (GPT-2's
        # merges do: id 628) makes the fused rollout fold the newline pair at the end of a prompt into that token
        engine.set_option("tok_nlnl_id", nlnl_token_id(engine.tokenizer))
        if os.environ.get("LMRL_NO_GRAPHS") == "1":        # debugging aid (compute-sanitizer racecheck, ncu): eager launches instead of graph replay
            engine.set_use_graphs(False)
        if os.environ.get("LMRL_GEMM_IMPL"):                # debugging aid: 0 = mma.sync GEMMs (tools that cannot instrument tcgen05 / TMA kernels)
            engine.set_gemm_impl(int(os.environ["LMRL_GEMM_IMPL"]))

    @property
    def fusable(self) -> bool:
        return self.in_str_process is None and self.out_str_process is None

    def sampling(self) -> "_lib.lmrl_sampling":
        return _lib.lmrl_sampling(temperature=self.temperature, do_sample=int(self.do_sample), seed=self.seed & ((1 << 64) - 1),
                                  top_k=self.top_k, top_p=self.top_p)

    def set_params(self, params: Dict) -> None:
        _upload_params(self._lmrl_engine, self.config, params)

    def generate_tokens(self, prompts: List[List[int]], episode_keys=None, turn: int = 0):
        e = self._lmrl_engine
        n = len(prompts)
        pitch = max(8, max((len(p) for p in prompts), default=8))
        toks = np.zeros((n, pitch), np.int32)
        lens = np.zeros(n, np.int32)
        for i, p in enumerate(prompts):
            toks[i, :len(p)] = p
            lens[i] = len(p)
        out = np.zeros((n, e.max_new_tokens), np.int32)
        out_len = np.zeros(n, np.int32)
        keys = None if episode_keys is None else np.asarray(episode_keys, dtype=np.uint64)
        s = self.sampling()
        check(e.lib.lmrl_generate_host(e.handle, ptr(toks), ptr(lens), pitch, n, ptr(keys), int(turn), C.byref(s), ptr(out), ptr(out_len)))
        return [out[i, :out_len[i]].tolist() for i in range(n)]

    def logits(self, prompts: List[List[int]]) -> np.ndarray:
        """Parity hook: fp32 logits of the last position of every prompt."""
        e = self._lmrl_engine
        n = len(prompts)
        pitch = max(8, max(len(p) for p in prompts))
        toks = np.zeros((n, pitch), np.int32)
        lens = np.zeros(n, np.int32)
        for i, p in enumerate(prompts):
            toks[i, :len(p)] = p
            lens[i] = len(p)
        out = np.zeros((n, self.config.n_vocab), np.float32)
        check(e.lib.lmrl_logits_host(e.handle, ptr(toks), ptr(lens), pitch, n, ptr(out)))
        return out

    def act(self, text_history: List[Optional[TextHistory]], done: Optional[List[bool]] = None) -> List[Optional[TextHistory]]:
        if done is None:
            done = [False] * len(text_history)
        tok = self._lmrl_engine.tokenizer
        in_f = self.in_str_process or (lambda x: x)
        out_f = self.out_str_process or default_out_str_process
        live = [i for i, d in enumerate(done) if not d]
        out: List[Optional[TextHistory]] = [None] * len(text_history)
        cap = self._lmrl_engine.max_envs
        for lo in range(0, len(live), cap):
            idx = live[lo:lo + cap]
            prompts = [tok.encode(in_f(text_history_to_str(text_history[i]))) for i in idx]
            gens = self.generate_tokens(prompts, episode_keys=[self._calls * (1 << 20) + i for i in idx], turn=0)
            for i, g in zip(idx, gens):
                out[i] = tuple(text_history[i]) + (Text(out_f(tok.decode(g)), True),)
        self._calls += 1
        return out


class B200GPT2ValuePolicy(B200GPT2Policy):
    """``GPT2ValuePolicy`` (LLM_RL/algorithms/value_rl_base/gpt2/interface.py:240-329) on the engine: every sampled token
    uses ``logits = pi_beta(x) + beta * min(Q1(h), Q2(h))`` with ``h`` the last hidden state of the value base
    (value_rl_base/gpt2/generation.py:36-121).  ``params`` is the reference's tuple
    ``(pi_beta_params, base_params, q1_head_params, q2_head_params-or-None)`` (interface.py:101); GPT-2 parameter sets use
    the dict layout of ``init_gpt2_params``, heads the flax layout ``{'dense1': {'kernel','bias'}, 'dense2': {...}}`` (MLPHead) or
    ``{'dense': {'kernel','bias'}}`` (LinearHead, ``q = dense(h)``).
    ILQL passes two MLP Q heads (beta 16 in wordle/ilql/eval_ilql_gpt2.py:47), MC returns one LinearHead (wordle/mc/eval_mc_gpt2.py:93-107,
    train_mc_returns_gpt2.py:198-206)."""

    def __init__(self, engine: Engine, params: Tuple, config: Optional[GPT2Config] = None, beta: float = 16.0, **kw):
        pi_beta, base, q1, q2 = params
        # pi_beta = None (generation.py:116-119: logits = beta * min q): the engine still wants a tower 0 (embedding tables of the row
        # builder, buffer shapes), so the value base stands in for it and policy kind 2 skips its layer stack and its LM-head segment
        super().__init__(engine, pi_beta if pi_beta is not None else base, config, **kw)
        self.beta = float(beta)
        _upload_params(engine, self.config, base, tower=1)
        _upload_q_heads(engine, q1, q2, self.beta)
        if pi_beta is None:
            check(engine.lib.lmrl_set_policy_kind(engine.handle, 2))

    def set_params(self, policy_params: Tuple) -> None:
        pi_beta, base, q1, q2 = policy_params
        _upload_params(self._lmrl_engine, self.config, pi_beta if pi_beta is not None else base, tower=0)
        _upload_params(self._lmrl_engine, self.config, base, tower=1)
        _upload_q_heads(self._lmrl_engine, q1, q2, self.beta)
        if pi_beta is None:
            check(self._lmrl_engine.lib.lmrl_set_policy_kind(self._lmrl_engine.handle, 2))
