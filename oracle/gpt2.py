"""GPT-2 decoder + sampler — fp32 torch-CPU restatement (oracle; test infrastructure only).  PARITY UNPINNED.

The reference's policy arithmetic lives in third-party code that is neither under /root/reference nor installable
here: Sea-Snell/JaxSEQ (git HEAD, unpinned: setup_jaxseq.sh:1-3) driving transformers==4.26.1 FlaxGPT2LMHeadModel /
FlaxGenerationMixin on jax==0.4.7 (requirements.txt:1-20).  Call sites in the reference:
LLM_RL/algorithms/ppo/gpt2/interface.py:527-535 and value_rl_base/gpt2/interface.py:98-105.  What is restated here
is the published GPT-2 architecture as HF implements it (GPT2Config defaults: pre-LN blocks, fused c_attn, 64-wide
heads, gelu_new, LN eps 1e-5, learned positions, tied LM head) and the semantics of HF sampling the reference uses
(do_sample, temperature, EOS = '\\n', max_new_tokens; top-k/top-p None: wordle/bc/eval_bc_gpt2.py:42-44, 90-95).
``forward_last_logits`` is itself checked against ``transformers.GPT2LMHeadModel`` (torch) with identical weights in
tests/test_oracle_gpt2.py.

Sampling: ``jax.random.categorical`` cannot be matched bit-for-bit, so the engine defines its own counter-based
Gumbel-max sampler; ``row_key`` / ``gumbel_noise`` restate that definition (csrc/gemm_mma.cuh::noise_bits,
csrc/gpt2_kernels.cuh::prepare_sample_kernel) so that tests can check the chosen token is the arg-max of
(oracle logits / T + noise) up to the logit tolerance.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

M64 = (1 << 64) - 1


def bf16_round(a: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def _t(a, emulate):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return t.to(torch.bfloat16).to(torch.float32) if emulate else t


def _r(t: torch.Tensor, emulate: bool) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32) if emulate else t


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def layer_norm(x, g, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def forward_hidden(params: Dict, tokens: Sequence[int], n_head: int, eps: float = 1e-5, weights_bf16: bool = True,
                   emulate_bf16: bool = False, _residual_only: bool = False) -> torch.Tensor:
    """Full causal forward of one sequence; returns the final-LN hidden states [T, d].

    weights_bf16: matrices/embeddings rounded to bf16 first (what the engine stores), biases/LN params stay fp32.
    emulate_bf16: additionally round activations where the engine does (LN outputs, q/k/v, attention output,
                  GeLU output)."""
    W = lambda a: _t(a, weights_bf16)  # noqa: E731
    F = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))  # noqa: E731
    T = len(tokens)
    ids = torch.as_tensor(list(tokens), dtype=torch.long)
    x = W(params["wte"])[ids] + W(params["wpe"])[:T]
    d = x.shape[1]
    hd = d // n_head
    mask = torch.tril(torch.ones(T, T, dtype=torch.bool))
    for L in params["layers"]:
        xn = _r(layer_norm(x, F(L["ln1_g"]), F(L["ln1_b"]), eps), emulate_bf16)
        qkv = _r(xn @ W(L["attn_w"]) + F(L["attn_b"]), emulate_bf16)
        q, k, v = qkv.split(d, dim=1)
        q = q.view(T, n_head, hd).transpose(0, 1)
        k = k.view(T, n_head, hd).transpose(0, 1)
        v = v.view(T, n_head, hd).transpose(0, 1)
        s = (q @ k.transpose(1, 2)) / math.sqrt(hd)
        s = s.masked_fill(~mask, float("-inf"))
        a = _r((torch.softmax(s, dim=-1) @ v).transpose(0, 1).reshape(T, d), emulate_bf16)
        x = x + a @ W(L["proj_w"]) + F(L["proj_b"])
        xn = _r(layer_norm(x, F(L["ln2_g"]), F(L["ln2_b"]), eps), emulate_bf16)
        h = _r(gelu_new(xn @ W(L["fc_w"]) + F(L["fc_b"])), emulate_bf16)
        x = x + h @ W(L["fc2_w"]) + F(L["fc2_b"])
    if _residual_only:
        return x
    return _r(layer_norm(x, F(params["lnf_g"]), F(params["lnf_b"]), eps), emulate_bf16)


def _forward_residual(params: Dict, tokens: Sequence[int], n_head: int, **kw) -> torch.Tensor:
    """The fp32 residual stream after the last block (before ln_f)."""
    return forward_hidden(params, tokens, n_head, _residual_only=True, **kw)


def forward_last_logits(params: Dict, tokens: Sequence[int], n_head: int, n_vocab: int, **kw) -> np.ndarray:
    h = forward_hidden(params, tokens, n_head, **kw)
    wte = _t(params["wte"], kw.get("weights_bf16", True))[:n_vocab]
    return (h[-1] @ wte.T).numpy()


def sequence_logprob(params: Dict, tokens: Sequence[int], prefix_len: int, n_head: int, n_vocab: int, temperature: float = 1.0, **kw) -> np.ndarray:
    """build_bc_score_fn (LLM_RL/algorithms/ppo/score_fn.py:112-118): log softmax(logits_{t-1})[token_t] for t >= prefix_len."""
    h = forward_hidden(params, tokens, n_head, **kw)
    wte = _t(params["wte"], kw.get("weights_bf16", True))[:n_vocab]
    logp = torch.log_softmax((h @ wte.T).double() / temperature, dim=-1).numpy()
    return np.asarray([logp[t - 1, tokens[t]] for t in range(max(prefix_len, 1), len(tokens))], dtype=np.float64)


# ---- ILQL / MC value-perturbed logits -------------------------------------------------------------------------------

def mlp_head(h: "torch.Tensor", q: Dict, weights_bf16: bool = True) -> "torch.Tensor":
    """MLPHead.__call__ (LLM_RL/heads/mlp_head.py:139-148): dense2(relu(dense1(x))); a ``{"dense": ...}`` dict is a LinearHead
    (heads/linear_head.py:112-119): dense(x)."""
    if "dense" in q:
        b = q["dense"].get("bias")
        out = h @ _t(q["dense"]["kernel"], weights_bf16)
        return out if b is None else out + torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32))
    w1, w2 = _t(q["dense1"]["kernel"], weights_bf16), _t(q["dense2"]["kernel"], weights_bf16)
    b1 = torch.from_numpy(np.ascontiguousarray(q["dense1"]["bias"], dtype=np.float32))
    b2 = torch.from_numpy(np.ascontiguousarray(q["dense2"]["bias"], dtype=np.float32))
    return torch.relu(h @ w1 + b1) @ w2 + b2


def value_policy_last_logits(params, tokens: Sequence[int], n_head: int, n_vocab: int, beta: float, emulate_bf16: bool = False,
                             weights_bf16: bool = True, **kw) -> np.ndarray:
    """GPT2ValueRLGeneration.__call__ (LLM_RL/algorithms/value_rl_base/gpt2/generation.py:36-121) for the last position:
    logits = pi_beta_logits + beta * min(q1(h), q2(h)) (or beta * q1(h) when q2 is None, :112-119), h = hidden_states[-1]
    of the value base (:87; HF appends the post-ln_f state last).  ``emulate_bf16`` rounds where the engine rounds: both
    final hidden states and beta*relu(dense1) are stored as bf16 before the last matmul."""
    pi_beta, base, q1, q2 = params
    h_v = forward_hidden(base, tokens, n_head, emulate_bf16=emulate_bf16, weights_bf16=weights_bf16, **kw)[-1]
    if pi_beta is None:                      # generation.py:118-119: logits = beta * q
        pi_logits = torch.zeros(n_vocab)
    else:
        h_pi = forward_hidden(pi_beta, tokens, n_head, emulate_bf16=emulate_bf16, weights_bf16=weights_bf16, **kw)[-1]
        pi_logits = h_pi @ _t(pi_beta["wte"], weights_bf16)[:n_vocab].T
    if not emulate_bf16:
        qs = [mlp_head(h_v, q, weights_bf16)[:n_vocab] for q in (q1, q2) if q is not None]
        qmin = qs[0] if len(qs) == 1 else torch.minimum(qs[0], qs[1])
        return (pi_logits + beta * qmin).numpy()
    outs = []
    for q in (q1, q2):
        if q is None:
            continue
        if "dense" in q:                     # LinearHead: the engine folds beta into the bf16 kernel, h_v is already bf16-rounded
            b = q["dense"].get("bias")
            bias = torch.zeros(n_vocab) if b is None else torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32))[:n_vocab]
            outs.append(pi_logits + h_v @ _r(beta * _t(q["dense"]["kernel"], False), True)[:, :n_vocab] + beta * bias)
            continue
        b1 = torch.from_numpy(np.ascontiguousarray(q["dense1"]["bias"], dtype=np.float32))
        b2 = torch.from_numpy(np.ascontiguousarray(q["dense2"]["bias"], dtype=np.float32))
        r = _r(beta * torch.relu(h_v @ _t(q["dense1"]["kernel"], weights_bf16) + b1), True)
        outs.append(pi_logits + r @ _t(q["dense2"]["kernel"], weights_bf16)[:, :n_vocab] + beta * b2[:n_vocab])
    return (outs[0] if len(outs) == 1 else torch.minimum(outs[0], outs[1])).numpy()


# ---- per-position teacher-forced forward: value heads and token log-probs (lmrl_forward_host) ------------------------

def value_forward_gathered(base: Dict, q1: Optional[Dict], q2: Optional[Dict], v: Optional[Dict], tokens: Sequence[int], n_head: int,
                           emulate_bf16: bool = False, weights_bf16: bool = True) -> Dict[str, np.ndarray]:
    """ValueRLInference._forward (LLM_RL/algorithms/value_rl_base/gpt2/interface.py:150-230) for one sequence, reduced to what
    every consumer gathers: q_k(h_t)[tok_{t+1}] for t < T-1 (ilql/gpt2/score_fn.py:49) and v(h_t) for every t.
    ``emulate_bf16`` rounds relu(dense1) to bf16 like the engine's ReLU-epilogue GEMM."""
    h = forward_hidden(base, tokens, n_head, emulate_bf16=emulate_bf16, weights_bf16=weights_bf16)
    nxt = torch.as_tensor(list(tokens[1:]), dtype=torch.long)
    out: Dict[str, np.ndarray] = {}
    for name, q in (("q1", q1), ("q2", q2)):
        if q is None:
            continue
        if "dense" in q:                     # LinearHead Q head (heads/linear_head.py:112-119): q = dense(h)
            w = _t(q["dense"]["kernel"], weights_bf16)[:, nxt]
            b = q["dense"].get("bias")
            out[name] = ((h[:-1] * w.T).sum(-1) + (0.0 if b is None else torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32))[nxt])).numpy()
            continue
        b1 = torch.from_numpy(np.ascontiguousarray(q["dense1"]["bias"], dtype=np.float32))
        b2 = torch.from_numpy(np.ascontiguousarray(q["dense2"]["bias"], dtype=np.float32))
        r = _r(torch.relu(h[:-1] @ _t(q["dense1"]["kernel"], weights_bf16) + b1), emulate_bf16)
        w2 = _t(q["dense2"]["kernel"], weights_bf16)[:, nxt]                 # [hidden, T-1]: the one dense2 column each position needs
        out[name] = ((r * w2.T).sum(-1) + b2[nxt]).numpy()
    if v is not None:
        b1 = torch.from_numpy(np.ascontiguousarray(v["dense1"]["bias"], dtype=np.float32))
        r = _r(torch.relu(h @ _t(v["dense1"]["kernel"], weights_bf16) + b1), emulate_bf16)
        w2 = torch.from_numpy(np.ascontiguousarray(v["dense2"]["kernel"], dtype=np.float32))[:, 0]
        out["v"] = (r @ w2 + float(np.asarray(v["dense2"]["bias"]).reshape(-1)[0])).numpy()
    return out


def linear_value_forward(params: Dict, v_linear: Dict, tokens: Sequence[int], n_head: int, emulate_bf16: bool = False, weights_bf16: bool = True) -> np.ndarray:
    """PPO value head: LinearHead (LLM_RL/heads/linear_head.py:112-119) on the policy's last hidden state, every position
    (ppo/gpt2/interface.py forward: ``value_head_model.apply(..., policy_output.hidden_states[-1])``).  The engine recomputes the final
    LayerNorm in fp32 for this head, so no rounding of its output is emulated."""
    F = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))  # noqa: E731
    x = _forward_residual(params, tokens, n_head, weights_bf16=weights_bf16, emulate_bf16=emulate_bf16)
    hn = layer_norm(x, F(params["lnf_g"]), F(params["lnf_b"]), 1e-5)
    w = F(np.asarray(v_linear["dense"]["kernel"]).reshape(-1))
    b = float(np.asarray(v_linear["dense"].get("bias", [0.0])).reshape(-1)[0]) if v_linear["dense"].get("bias") is not None else 0.0
    return (hn @ w + b).numpy()


def token_logprobs_all(params: Dict, tokens: Sequence[int], n_head: int, n_vocab: int, temperature: float = 1.0, **kw) -> np.ndarray:
    """PPOInference.token_logprobs_from_logits (ppo/base_interface.py:404-411): log softmax(logits_t)[tok_{t+1}], t < T-1; ids beyond
    n_vocab masked to -inf (ppo/gpt2/interface.py:330)."""
    return sequence_logprob(params, tokens, 1, n_head, n_vocab, temperature, **kw)


# ---- the engine's sampler, restated -------------------------------------------------------------------------------

def splitmix64(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def row_key(base_seed: int, episode_key: int, turn: int, k: int) -> int:
    return splitmix64((splitmix64(base_seed ^ splitmix64(episode_key)) + (turn << 32) + k) & M64)


def _fmix32(h: np.ndarray) -> np.ndarray:
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = (h.astype(np.uint64) * np.uint64(0x85EBCA6B)).astype(np.uint32)
    h ^= h >> np.uint32(13)
    h = (h.astype(np.uint64) * np.uint64(0xC2B2AE35)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    return h


def noise_bits(key: int, n_vocab: int) -> np.ndarray:
    """csrc/gemm_mma.cuh::noise_bits — two multiply-fold rounds keyed by the 64-bit row key."""
    m32 = np.uint64(0xFFFFFFFF)
    col = np.arange(n_vocab, dtype=np.uint64)
    p = ((col ^ np.uint64(key & 0xFFFFFFFF)) & m32) * np.uint64(0xD2511F53)
    x = ((p >> np.uint64(32)) ^ (p & m32) ^ np.uint64(key >> 32)) & m32
    p = x * np.uint64(0xCD9E8D57)
    return (((p >> np.uint64(32)) ^ (p & m32)) & m32).astype(np.uint32)


def gumbel_noise(key: int, n_vocab: int) -> np.ndarray:
    h = noise_bits(key, n_vocab)
    u = ((h >> np.uint32(8)).astype(np.float64) + 0.5) / 16777216.0
    return -np.log(-np.log(u))


def warp_top_k(scores: np.ndarray, top_k: int) -> np.ndarray:
    """FlaxTopKLogitsWarper (transformers 4.26.1 generation/flax_logits_process.py): keep the top_k scores, the rest -> -inf;
    lax.top_k resolves ties by the lower index."""
    k = min(max(int(top_k), 1), scores.shape[-1])
    order = np.argsort(-scores, kind="stable")          # descending, lower index first among equals
    out = np.full_like(scores, -np.inf)
    out[order[:k]] = scores[order[:k]]
    return out


def warp_top_p(scores: np.ndarray, top_p: float) -> np.ndarray:
    """FlaxTopPLogitsWarper: sort descending, keep a token iff the cumulative probability of the tokens before it is
    < top_p (so the token that crosses top_p is kept; min_tokens_to_keep = 1)."""
    order = np.argsort(-scores, kind="stable")
    s = scores[order].astype(np.float64)
    p = np.exp(s - s.max())
    p /= p.sum()
    before = np.concatenate(([0.0], np.cumsum(p)[:-1]))
    keep = before < top_p
    keep[0] = True
    out = np.full_like(scores, -np.inf)
    out[order[keep]] = scores[order[keep]]
    return out


def sample_scores(logits: np.ndarray, key: int, temperature: float = 1.0, do_sample: bool = True, top_k: int = 0,
                  top_p: float = 0.0) -> np.ndarray:
    s = logits.astype(np.float64) / (temperature if temperature and temperature > 0 else 1.0)
    if top_k and top_k > 0:
        s = warp_top_k(s, top_k)
    if top_p and 0.0 < top_p < 1.0:
        s = warp_top_p(s, top_p)
    if do_sample:
        s = s + gumbel_noise(key, len(logits))
    return s


def generate(params: Dict, prompt: Sequence[int], n_head: int, n_vocab: int, eos_id: int, max_new_tokens: int,
             base_seed: int, episode_key: int, turn: int, temperature: float = 1.0, do_sample: bool = True,
             max_input_length: Optional[int] = None, **kw) -> List[int]:
    """HF-style sampling loop for ONE sequence (prompt LEFT-truncated to max_input_length)."""
    toks = list(prompt)[-max_input_length:] if max_input_length else list(prompt)
    out: List[int] = []
    for k in range(max_new_tokens):
        logits = forward_last_logits(params, toks + out, n_head, n_vocab, **kw)
        t = int(np.argmax(sample_scores(logits, row_key(base_seed, episode_key, turn, k), temperature, do_sample)))
        out.append(t)
        if t == eos_id:
            break
    return out
