"""PPO round = device rollout + device PPO data path + gradient step (BASELINE.json configs[4]).

Mirrors the data flow of ``LLM_RL/algorithms/ppo/train.py::train_loop`` (:138-477) as driven by
``llm_rl_scripts/wordle/ppo/train_ppo_gpt2.py`` (:297-353): every round

    text_env_eval(env, policy)                        -> fused rollout on the engine (csrc/gpt2_engine.inl)
    episodes -> token trajectory chains               -> exported by the device (lmrl_rollout_chains_host); text path as fall-back
    get_ppo_data_from_token_trajectory_chain          -> device forward of policy / initial policy / value head + device GAE (ppo.py)
    kl_controller.update(mean_kl, train_bsize)        -> train_ppo_gpt2.py:352-353
    epochs x minibatches of PPOTrain.step             -> ``B200PPOTrain.step`` below
    policy.set_params(new params)                     -> ppo/train.py:244, 438 (device-to-device weight hand-over)

The gradient step restates ``GPT2PPOTrain._step`` (ppo/gpt2/interface.py:72-211: policy forward, LinearHead values, ``ppo_loss_fn``,
AdamW(b1 0.9, b2 0.95, eps 1e-8, weight decay masked off LayerNorm and biases, train_ppo_gpt2.py:144-165)) in two interchangeable ways:

  backward="b200"   HAND-WRITTEN: forward on the engine (``lmrl_forward_host``), the closed-form derivative of ``ppo_loss_fn`` with
                    respect to the per-token log-probs / values (``ppo_loss_grads``), the engine's backward pass through LM head, value
                    head and decoder stack (``lmrl_policy_grad_host``, csrc/train_kernels.inl: tcgen05 dgrad / wgrad GEMMs, LayerNorm /
                    GeLU / attention backward kernels), ``lmrl_adamw_dev``.  Checked against autograd in tests/test_gpu_backward.py.
  backward="torch"  LIBRARY path: torch autograd on torch/cuBLAS kernels — the comparison baseline (and the CPU test vehicle).

Gradients are summed over ranks with ``torch.distributed.all_reduce`` (NCCL; plumbing) where the reference relies on pjit data parallelism.
fp32 master weights live in torch tensors (device memory, plumbing) either way.
"""
from __future__ import annotations

import math
from dataclasses import replace
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import check
from .engine import Engine
from .environment import Text, TextTrajectory, TextTrajectoryChain, text_env_eval
from .gpt2 import GPT2Config, _LAYER_KEYS
from .ppo import B200PPOInference, PPOData, TokenTrajectory, TokenTrajectoryChain, _block_right
from .value_inference import upload_value_heads

import ctypes as C


def episodes_to_chains(raw_results, tokenizer, max_length: int, gamma: float) -> List[TextTrajectoryChain]:
    """train_ppo_gpt2.py:307-335: one TextTrajectory per episode over its final history, rewards [0, r_0, 0, r_1, ...]; histories
    too long for ``max_length`` drop their last (action, observation) pair — its discounted reward folded into the previous action —
    until they fit; episodes that still do not fit, or that have no full turn, are skipped."""
    chains = []
    for ep in raw_results:
        if len(ep) == 0:
            continue
        hist = tuple(ep[-1].post_transition_history)
        rewards = [0.0]
        for tr in ep:
            rewards += [float(tr.reward), 0.0]
        traj = TextTrajectory(hist, tuple(rewards), bool(ep[-1].done))
        n_tok = lambda t: TokenTrajectory.from_text_trajectory(t, tokenizer).tokens.shape[0]  # noqa: E731
        while len(traj.text_history) > 3 and n_tok(traj) >= max_length:
            r = list(traj.reward[:-2])
            r[-2] += sum(traj.reward[-2:]) * gamma
            traj = TextTrajectory(traj.text_history[:-2], tuple(r), False)
        if len(traj.text_history) < 3 or n_tok(traj) >= max_length:
            continue
        chains.append(TextTrajectoryChain(traj, None))
    return chains


def device_token_chains(engine: Engine, n: int, max_length: int) -> List[Optional["TokenTrajectoryChain"]]:
    """trajectory -> token chain WITHOUT a host tokeniser: after a fused Wordle rollout of ``n`` episodes the engine's canonical token
    streams are exported with their is_action / reward annotations (``lmrl_rollout_chains_host``) — exactly
    ``TokenTrajectory.from_text_trajectory`` of the trajectory ``episodes_to_chains`` builds.  Entries are None for episodes the device
    cannot represent (stream LEFT-truncated during the rollout, or >= max_length tokens: the reference drops trailing turns there,
    train_ppo_gpt2.py:317-331) — the caller converts those through the text path."""
    pitch = int(engine.max_input_length + 8)
    toks = np.zeros((n, pitch), np.int32)
    n_tok = np.zeros(n, np.int32)
    act = np.zeros((n, pitch), np.uint8)
    rew = np.zeros((n, pitch), np.float32)
    done, trunc = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    check(engine.lib.lmrl_rollout_chains_host(engine.handle, n, pitch, toks.ctypes.data, n_tok.ctypes.data, act.ctypes.data, rew.ctypes.data,
                                              done.ctypes.data, trunc.ctypes.data))
    out: List[Optional[TokenTrajectoryChain]] = []
    for i in range(n):
        k = int(n_tok[i])
        if trunc[i] or k >= max_length or not act[i, :k].any():
            out.append(None)
            continue
        out.append(TokenTrajectoryChain(TokenTrajectory(toks[i, :k].copy(), act[i, :k].astype(np.bool_), rew[i, :k].copy(), np.array(bool(done[i]), dtype=np.bool_)), None))
    return out


def block_ppo_data(datas: List[PPOData], pad_token_id: int, max_length: int) -> Dict[str, np.ndarray]:
    """``PPOData.block`` (ppo/data.py:18-60): RIGHT pad / RIGHT truncate to max_length (ids) and max_length - 1 (per-output arrays)."""
    return dict(
        input_ids=_block_right([d.input_ids for d in datas], pad_token_id, max_length, np.int32),
        should_take_action=_block_right([d.should_take_action for d in datas], False, max_length - 1, np.bool_),
        old_logprobs=_block_right([d.old_logprobs for d in datas], 0.0, max_length - 1, np.float32),
        old_values=_block_right([d.old_values for d in datas], 0.0, max_length - 1, np.float32),
        old_advantages=_block_right([d.old_advantages for d in datas], 0.0, max_length - 1, np.float32),
        old_returns=_block_right([d.old_returns for d in datas], 0.0, max_length - 1, np.float32))


def ppo_loss_grads(attention_mask, logprobs, values, should_take_action, old_logprobs, old_values, old_advantages, old_returns, *,
                   cliprange_value: float, cliprange: float, value_loss_coef: float):
    """``ppo_loss_fn`` (ppo/base_interface.py:71-142) and its closed-form derivative with respect to the per-token ``logprobs`` and
    ``values`` (what ``jax.value_and_grad`` propagates into the networks, ppo/gpt2/interface.py:106-164).  All arrays [b, t-1].
    Returns (loss, logs, dL/dlogprobs, dL/dvalues)."""
    f = lambda a: np.asarray(a, dtype=np.float64)  # noqa: E731
    lp, v, olp, ov, adv, ret = map(f, (logprobs, values, old_logprobs, old_values, old_advantages, old_returns))
    mask = np.asarray(should_take_action, dtype=np.float64) * f(attention_mask)
    n = mask.sum()
    vc = np.clip(v, ov - cliprange_value, ov + cliprange_value)
    vf1, vf2 = (v - ret) ** 2, (vc - ret) ** 2
    vf_loss = 0.5 * (np.maximum(vf1, vf2) * mask).sum() / n
    inside_v = (v >= ov - cliprange_value) & (v <= ov + cliprange_value)
    dvf = np.where(vf2 > vf1, 2.0 * (vc - ret) * inside_v, 2.0 * (v - ret))
    dvalues = value_loss_coef * 0.5 * dvf * mask / n
    log_ratio = (lp - olp) * mask
    ratio = np.exp(log_ratio)
    clipped = np.clip(ratio, 1.0 - cliprange, 1.0 + cliprange)
    pg1, pg2 = -adv * ratio, -adv * clipped
    pg_loss = (np.maximum(pg1, pg2) * mask).sum() / n
    inside_r = (ratio >= 1.0 - cliprange) & (ratio <= 1.0 + cliprange)
    dpg = np.where(pg2 > pg1, -adv * ratio * inside_r, -adv * ratio) * mask           # d ratio / d logprob = ratio * mask
    dlogprobs = dpg * mask / n
    loss = pg_loss + value_loss_coef * vf_loss
    logs = dict(total_loss=loss, policy_loss=pg_loss, value_loss=vf_loss, approx_kl=((ratio - 1) - log_ratio).sum() / n,
                policy_clipfrac=((pg2 > pg1) * mask).sum() / n, value_clipfrac=((vf2 > vf1) * mask).sum() / n, ratio=(ratio * mask).sum() / n)
    return float(loss), {k: float(x) for k, x in logs.items()}, dlogprobs.astype(np.float32), dvalues.astype(np.float32)


class B200PPOTrain:
    """fp32 master weights on the GPU as torch tensors (HF GPT-2 layout, the same dict layout as ``init_gpt2_params``) + LinearHead."""

    def __init__(self, engine: Engine, policy_params: Dict, value_head_params: Dict, config: Optional[GPT2Config] = None, *, lr: float = 1e-5,
                 weight_decay: float = 0.0, cliprange_value: float = 0.2, cliprange: float = 0.2, value_loss_coef: float = 1.0,
                 pad_token_id: Optional[int] = None, activation_dtype: str = "bf16", backward: Optional[str] = None, max_rows: int = 16384,
                 max_scored_rows: int = 8192, bc_loss_weight: float = 0.0):
        import torch
        self.torch = torch
        self.bc_loss_weight = float(bc_loss_weight)      # weight of the optional BC term (GPT2PPOTrain bc_loss_weight, ppo/gpt2/interface.py:36)
        self._lmrl_engine = engine                      # None: no rollout engine attached (CPU tests of the update itself)
        self.config = config or GPT2Config.small()
        self.device = torch.device("cuda", engine.device) if engine is not None else torch.device("cpu")
        self.pad_token_id = int(pad_token_id if pad_token_id is not None else self.config.n_vocab)
        self.cliprange_value, self.cliprange, self.value_loss_coef = cliprange_value, cliprange, value_loss_coef
        self.autocast = torch.bfloat16 if activation_dtype == "bf16" else None
        # fp32 master weights live in TWO flat buffers (weight-decayed matrices / embeddings; LayerNorm scales and every bias, which
        # train_ppo_gpt2.py:144-151 masks out of the decay); the per-tensor parameters are views, so autograd accumulates straight
        # into flat gradients: one all-reduce and one AdamW launch per buffer.
        entries = [("wte", policy_params["wte"]), ("wpe", policy_params["wpe"]), ("lnf_g", policy_params["lnf_g"]), ("lnf_b", policy_params["lnf_b"])]
        for i, L in enumerate(policy_params["layers"]):
            entries += [(f"h{i}.{k}", L[k]) for k in _LAYER_KEYS]
        entries += [("vh_w", np.asarray(value_head_params["dense"]["kernel"]).reshape(-1, 1)),
                    ("vh_b", np.asarray(value_head_params["dense"].get("bias", np.zeros(1))).reshape(1))]
        no_decay = lambda name: name.endswith("_b") or name.endswith("_g") or name == "vh_b"  # noqa: E731
        self.flat, self._views, self._names = [], {}, [[], []]
        for grp in (0, 1):
            chosen = [(n, np.ascontiguousarray(a, dtype=np.float32)) for n, a in entries if int(no_decay(n)) == grp]
            flat = torch.nn.Parameter(torch.cat([torch.from_numpy(a.reshape(-1)) for _, a in chosen]).to(self.device))
            self.flat.append(flat)
            o = 0
            for n, a in chosen:
                self._views[n] = (grp, o, a.shape)
                self._names[grp].append(n)
                o += a.size
        self.weight_decay, self.lr, self.betas, self.eps = float(weight_decay), float(lr), (0.9, 0.95), 1e-8
        self.use_device_adamw = engine is not None
        if self.use_device_adamw:
            self.m = [torch.zeros_like(f.data) for f in self.flat]
            self.v = [torch.zeros_like(f.data) for f in self.flat]
            self.opt = None
        else:
            self.opt = torch.optim.AdamW([dict(params=[self.flat[0]], weight_decay=weight_decay), dict(params=[self.flat[1]], weight_decay=0.0)],
                                         lr=lr, betas=self.betas, eps=self.eps)
        self.steps = 0
        # backward = "b200": gradients from the engine's hand-written backward pass (csrc/train_kernels.inl) instead of torch autograd
        backward = backward or ("b200" if engine is not None else "torch")
        self.backward = backward
        if backward == "b200":
            assert engine is not None, "the hand-written backward pass runs on the engine"
            self.gflat = [torch.zeros_like(f.data) for f in self.flat]
            self.push_to_engine()
            check(engine.lib.lmrl_train_enable(engine.handle, int(max_rows), int(max_scored_rows)))

    def _view(self, name):
        grp, o, shape = self._views[name]
        n = int(np.prod(shape))
        return self.flat[grp][o:o + n].view(*shape)

    @property
    def p(self) -> Dict:
        """The parameter dict (HF GPT-2 layout) as views of the flat master buffers — rebuilt per access so autograd sees fresh views."""
        return dict(wte=self._view("wte"), wpe=self._view("wpe"), lnf_g=self._view("lnf_g"), lnf_b=self._view("lnf_b"),
                    layers=[{k: self._view(f"h{i}.{k}") for k in _LAYER_KEYS} for i in range(self.config.n_layer)])

    @property
    def vh_w(self):
        return self._view("vh_w")

    @property
    def vh_b(self):
        return self._view("vh_b")

    def named_parameters(self):
        for grp in (0, 1):
            for n in self._names[grp]:
                yield n, self._view(n)

    # ---- model -------------------------------------------------------------------------------------------------------
    def _hidden(self, ids, attn_mask, params=None):
        torch, F = self.torch, self.torch.nn.functional
        cfg = self.config
        P = params if params is not None else self.p
        b, T = ids.shape
        pos = torch.arange(T, device=ids.device)
        x = P["wte"][ids.clamp(max=P["wte"].shape[0] - 1)] + P["wpe"][pos][None]
        causal = torch.tril(torch.ones(T, T, dtype=torch.bool, device=ids.device))
        mask = causal[None, None] & attn_mask[:, None, None, :]
        for L in P["layers"]:
            h = F.layer_norm(x, (cfg.d_model,), L["ln1_g"], L["ln1_b"], cfg.ln_eps)
            qkv = h @ L["attn_w"] + L["attn_b"]
            q, k, v = (z.view(b, T, cfg.n_head, -1).transpose(1, 2) for z in qkv.split(cfg.d_model, dim=-1))
            a = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
            x = x + a.transpose(1, 2).reshape(b, T, cfg.d_model) @ L["proj_w"] + L["proj_b"]
            h = F.layer_norm(x, (cfg.d_model,), L["ln2_g"], L["ln2_b"], cfg.ln_eps)
            x = x + F.gelu(h @ L["fc_w"] + L["fc_b"], approximate="tanh") @ L["fc2_w"] + L["fc2_b"]
        return F.layer_norm(x, (cfg.d_model,), P["lnf_g"], P["lnf_b"], cfg.ln_eps)

    def bc_loss(self, bc_batch: Dict[str, np.ndarray]):
        """The BC term of the PPO step (``bc_loss_fn``; every driver passes JaxSeq's ``loss_fn_mask``, train_ppo_gpt2.py:36, 288): masked
        next-token cross-entropy, sum(CE * m) / sum(m) with m = input_training_mask[:, 1:].  JaxSeq is not vendored by the reference: this
        restates its published definition (parity unpinned)."""
        torch = self.torch
        dev = self.device
        ids = torch.as_tensor(bc_batch["input_ids"], device=dev).long()
        attn = ids != self.pad_token_id
        m = torch.as_tensor(np.asarray(bc_batch["input_training_mask"], dtype=np.float32), device=dev)[:, 1:] * attn[:, 1:].float()
        with torch.autocast(self.device.type, dtype=self.autocast, enabled=self.autocast is not None):
            P = self.p
            logits = self._hidden(ids, attn, P) @ P["wte"][:self.config.n_vocab].T
        ce = torch.nn.functional.cross_entropy(logits[:, :-1].float().reshape(-1, logits.shape[-1]), ids[:, 1:].clamp(max=self.config.n_vocab - 1).reshape(-1),
                                               reduction="none").view(ids.shape[0], -1)
        return (ce * m).sum() / m.sum()

    def loss(self, batch: Dict[str, np.ndarray]):
        """ppo/gpt2/interface.py:106-146 + ppo_loss_fn (ppo/base_interface.py:71-142)."""
        torch = self.torch
        dev = self.device
        ids = torch.as_tensor(batch["input_ids"], device=dev).long()
        attn = ids != self.pad_token_id
        sta = torch.as_tensor(batch["should_take_action"], device=dev)
        old_lp, old_v, old_adv, old_ret = (torch.as_tensor(batch[k], device=dev).float() for k in ("old_logprobs", "old_values", "old_advantages", "old_returns"))
        with torch.autocast(self.device.type, dtype=self.autocast, enabled=self.autocast is not None):
            P = self.p
            h = self._hidden(ids, attn, P)
            logits = h @ P["wte"][:self.config.n_vocab].T
        values = (h.float() @ self.vh_w + self.vh_b)[:, :-1, 0]
        logprobs = -torch.nn.functional.cross_entropy(logits[:, :-1].float().reshape(-1, logits.shape[-1]), ids[:, 1:].clamp(max=self.config.n_vocab - 1).reshape(-1),
                                                      reduction="none").view(ids.shape[0], -1)
        mask = sta.float() * attn[:, 1:].float()
        n = mask.sum()
        v_clipped = torch.maximum(torch.minimum(values, old_v + self.cliprange_value), old_v - self.cliprange_value)
        vf1, vf2 = (values - old_ret) ** 2, (v_clipped - old_ret) ** 2
        vf_loss = 0.5 * (torch.maximum(vf1, vf2) * mask).sum() / n
        log_ratio = (logprobs - old_lp) * mask
        ratio = torch.exp(log_ratio)
        pg1, pg2 = -old_adv * ratio, -old_adv * torch.clamp(ratio, 1.0 - self.cliprange, 1.0 + self.cliprange)
        pg_loss = (torch.maximum(pg1, pg2) * mask).sum() / n
        loss = pg_loss + self.value_loss_coef * vf_loss
        logs = dict(total_loss=loss.detach(), policy_loss=pg_loss.detach(), value_loss=vf_loss.detach(),
                    approx_kl=(((ratio - 1) - log_ratio).sum() / n).detach(), policy_clipfrac=(((pg2 > pg1).float() * mask).sum() / n).detach(),
                    value_clipfrac=(((vf2 > vf1).float() * mask).sum() / n).detach(), ratio=((ratio * mask).sum() / n).detach())
        return loss, logs

    def step(self, batch: Dict[str, np.ndarray], bc_batch: Optional[Dict[str, np.ndarray]] = None) -> Tuple[float, Dict[str, float]]:
        """One optimizer step on one minibatch (``PPOTrain.step``, ppo/base_interface.py:183-228): torch autograd for the gradients
        (library path), a summing all-reduce of the two flat gradient buffers over the ranks, then AdamW — the hand-written streaming
        kernel ``lmrl_adamw_dev`` when an engine is attached, ``torch.optim.AdamW`` otherwise (CPU tests)."""
        torch = self.torch
        import os as _os
        import time as _time
        _timing = _os.environ.get("LMRL_STEP_TIMING") == "1"       # diagnostics: host wall clock of the step's phases on stderr
        _t = [_time.perf_counter()]
        # optional BC term (ppo/gpt2/interface.py:183-205): loss + w * bc_loss, policy gradients + w * bc gradients
        use_bc = bc_batch is not None and self.bc_loss_weight != 0.0
        if self.backward == "b200":
            loss, logs = self._b200_grads(batch)
            if use_bc:
                bc = self._b200_bc_grads(bc_batch)
                loss, logs = loss + self.bc_loss_weight * bc, dict(logs, bc_loss=bc, ppo_loss=loss, total_loss=loss + self.bc_loss_weight * bc)
            for f, g in zip(self.flat, self.gflat):
                f.grad = g
        else:
            for f in self.flat:
                f.grad = None
            loss, logs = self.loss(batch)
            if use_bc:
                bc = self.bc_loss(bc_batch)
                logs = dict(logs, bc_loss=bc.detach(), ppo_loss=loss.detach(), total_loss=(loss + self.bc_loss_weight * bc).detach())
                loss = loss + self.bc_loss_weight * bc
            loss.backward()
            loss = float(loss.detach())
        _t.append(_time.perf_counter())
        world = self._allreduce_grads()
        self.steps += 1
        if self.use_device_adamw:
            e = self._lmrl_engine
            torch.cuda.synchronize(self.device)
            for grp, f in enumerate(self.flat):
                check(e.lib.lmrl_adamw_dev(e.handle, f.data.data_ptr(), f.grad.data_ptr(), self.m[grp].data_ptr(), self.v[grp].data_ptr(), f.numel(),
                                           self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay if grp == 0 else 0.0, self.steps, 1.0 / world))
            e.sync()
        else:
            if world > 1:
                for f in self.flat:
                    f.grad /= world
            self.opt.step()
        _t.append(_time.perf_counter())
        if self.backward == "b200":
            self.push_to_engine()           # the next forward / backward (and the next rollout) run on the updated weights
        if _timing:
            import sys as _sys
            _t.append(_time.perf_counter())
            print("[B200PPOTrain.step] grads %.1f ms  allreduce+adamw %.1f ms  push_to_engine %.1f ms" % tuple(1e3 * (b - a) for a, b in zip(_t[:-1], _t[1:])), file=_sys.stderr, flush=True)
        return float(loss), {k: float(v) for k, v in logs.items()}

    def _allreduce_grads(self) -> int:
        """Summing all-reduce of the two flat gradient buffers over the data-parallel ranks (the one collective of the system:
        ppo/gpt2/interface.py:165-181 leaves it to pjit).  Returns the world size (AdamW divides by it).  Every call is bracketed by
        CUDA events so that ``allreduce_stats`` can report its device time and bytes (bench.py's ``ppo_round`` block)."""
        torch = self.torch
        if not (torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1):
            return 1
        timed = self.flat[0].is_cuda
        if timed:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        for f in self.flat:
            torch.distributed.all_reduce(f.grad)
        if timed:
            ev1.record()
            self._ar_events = getattr(self, "_ar_events", [])[-255:] + [(ev0, ev1)]
        self._ar_bytes = getattr(self, "_ar_bytes", 0) + sum(int(f.grad.numel()) * f.grad.element_size() for f in self.flat)
        self._ar_calls = getattr(self, "_ar_calls", 0) + 1
        return torch.distributed.get_world_size()

    def allreduce_stats(self) -> Dict[str, float]:
        """Mean device milliseconds and bytes of the gradient all-reduce per optimizer step since construction (zeros on one rank)."""
        calls = getattr(self, "_ar_calls", 0)
        if not calls:
            return dict(calls=0, ms_per_step=0.0, bytes_per_step=0)
        self.torch.cuda.synchronize(self.device)
        ev = getattr(self, "_ar_events", [])
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev])) if ev else 0.0
        return dict(calls=calls, ms_per_step=ms, bytes_per_step=self._ar_bytes // calls)

    def bc_step(self, bc_batch: Dict[str, np.ndarray]) -> float:
        """One optimizer step on the BC loss alone (masked next-token cross-entropy; what the reference's BC drivers minimise,
        wordle/bc/train_bc_gpt2.py with JaxSeq's ``loss_fn_mask``): same gradient paths, all-reduce and AdamW as ``step``; the value head
        gets a zero gradient.  Returns the loss."""
        torch = self.torch
        if self.backward == "b200":
            for g in self.gflat:
                g.zero_()
            w, self.bc_loss_weight = self.bc_loss_weight, 1.0          # _b200_bc_grads scales dL/dlogprob by the weight
            try:
                loss = self._b200_bc_grads(bc_batch)
            finally:
                self.bc_loss_weight = w
            for f, g in zip(self.flat, self.gflat):
                f.grad = g
        else:
            for f in self.flat:
                f.grad = None
            out = self.bc_loss(bc_batch)
            out.backward()
            loss = float(out.detach())
            for f in self.flat:                                        # parameters the loss does not reach (the value head)
                if f.grad is None:
                    f.grad = torch.zeros_like(f.data)
        world = self._allreduce_grads()
        self.steps += 1
        if self.use_device_adamw:
            e = self._lmrl_engine
            torch.cuda.synchronize(self.device)
            for grp, f in enumerate(self.flat):
                check(e.lib.lmrl_adamw_dev(e.handle, f.data.data_ptr(), f.grad.data_ptr(), self.m[grp].data_ptr(), self.v[grp].data_ptr(), f.numel(),
                                           self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay if grp == 0 else 0.0, self.steps, 1.0 / world))
            e.sync()
        else:
            if world > 1:
                for f in self.flat:
                    f.grad /= world
            self.opt.step()
        if self.backward == "b200":
            self.push_to_engine()
        return float(loss)

    def _grad_view(self, name):
        grp, o, shape = self._views[name]
        return self.gflat[grp][o:o + int(np.prod(shape))]

    def _grad_struct(self, keep: list):
        # built once per set of flat gradient buffers (the views' addresses do not change between steps)
        key = tuple(int(g.data_ptr()) for g in self.gflat)
        cached = getattr(self, "_grad_struct_cache", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        keep = []
        gs = self._grad_struct_build(keep)
        self._grad_struct_cache = (key, gs, keep)
        return gs

    def _grad_struct_build(self, keep: list):
        L = self.config.n_layer
        fp = lambda name: C.cast(C.c_void_p(self._grad_view(name).data_ptr()), _lib._PF)  # noqa: E731
        gs = _lib.lmrl_gpt2_grads()
        gs.wte, gs.n_vocab_rows, gs.wpe = fp("wte"), int(self._views["wte"][2][0]), fp("wpe")
        gs.lnf_g, gs.lnf_b, gs.vh_w, gs.vh_b = fp("lnf_g"), fp("lnf_b"), fp("vh_w"), fp("vh_b")
        for key in _LAYER_KEYS:
            ptrs = (_lib._PF * L)(*[fp(f"h{i}.{key}") for i in range(L)])
            keep.append(ptrs)
            setattr(gs, key, C.cast(ptrs, _lib._PPF))
        return gs

    def _b200_grads(self, batch: Dict[str, np.ndarray], fused: bool = True):
        """Gradients of the PPO minibatch loss into the flat gradient buffers, on the engine.

        fused=True   one call (``lmrl_ppo_grad_host``): a single forward with saved activations, ``ppo_loss_fn`` differentiated on the
                     device inside the LM-head / value kernels, backward pass.
        fused=False  the chain rule split on the host: ``lmrl_forward_host`` -> ``ppo_loss_grads`` (closed form, numpy float64) ->
                     ``lmrl_policy_grad_host``; kept as the cross-check of the fused kernels."""
        from .value_inference import forward_tokens
        e = self._lmrl_engine
        ids = np.asarray(batch["input_ids"])
        real = ids != self.pad_token_id
        lens = np.ascontiguousarray(real.sum(axis=1).astype(np.int32))
        n, T = ids.shape
        pitch = max(8, T)
        toks = np.zeros((n, pitch), np.int32)
        toks[:, :T] = np.where(real, ids, 0)
        for g in self.gflat:
            g.zero_()
        keep: list = []
        gs = self._grad_struct(keep)
        if fused:
            def pad(a, dtype):
                out = np.zeros((n, pitch), dtype)
                out[:, :T - 1] = a
                return out
            sta = pad(np.asarray(batch["should_take_action"], dtype=bool) & real[:, 1:], np.uint8)
            arrs = [pad(np.asarray(batch[k], dtype=np.float32), np.float32) for k in ("old_logprobs", "old_values", "old_advantages", "old_returns")]
            out = np.zeros(16, np.float64)
            pb = _lib.lmrl_ppo_batch(should_take_action=sta.ctypes.data, old_logprobs=arrs[0].ctypes.data, old_values=arrs[1].ctypes.data,
                                     old_advantages=arrs[2].ctypes.data, old_returns=arrs[3].ctypes.data, cliprange_value=self.cliprange_value,
                                     cliprange=self.cliprange, value_loss_coef=self.value_loss_coef, loss_out=out.ctypes.data)
            self.torch.cuda.synchronize(self.device)
            check(e.lib.lmrl_ppo_grad_host(e.handle, toks.ctypes.data, lens.ctypes.data, pitch, n, C.byref(pb), C.byref(gs)))
            logs = dict(total_loss=out[0], policy_loss=out[1], value_loss=out[2], approx_kl=out[5], policy_clipfrac=out[6], value_clipfrac=out[4], ratio=out[7])
            return float(out[0]), {k: float(v) for k, v in logs.items()}
        seqs = [ids[i, :lens[i]] for i in range(n)]
        fw = forward_tokens(e, 0, 0, seqs, logprobs=True, v=True, pitch=pitch)
        loss, logs, dlp, dv = ppo_loss_grads(real[:, 1:].astype(np.float32), fw["logprobs"][:, :T - 1], fw["v"][:, :T - 1], batch["should_take_action"],
                                             batch["old_logprobs"], batch["old_values"], batch["old_advantages"], batch["old_returns"],
                                             cliprange_value=self.cliprange_value, cliprange=self.cliprange, value_loss_coef=self.value_loss_coef)
        dlp_p, dv_p = np.zeros((n, pitch), np.float32), np.zeros((n, pitch), np.float32)
        dlp_p[:, :T - 1], dv_p[:, :T - 1] = dlp, dv
        self.torch.cuda.synchronize(self.device)
        check(e.lib.lmrl_policy_grad_host(e.handle, toks.ctypes.data, lens.ctypes.data, pitch, n, dlp_p.ctypes.data, dv_p.ctypes.data, C.c_float(1.0), C.byref(gs)))
        return loss, logs

    def _b200_bc_grads(self, bc_batch: Dict[str, np.ndarray]) -> float:
        """BC term on the engine: bc_loss = -sum(m * logp) / sum(m) from one teacher-forced forward, and its gradient through
        ``lmrl_policy_grad_host`` with dL/dlogprob = -w * m / sum(m), ACCUMULATED onto the PPO gradients already in the flat buffers."""
        from .value_inference import forward_tokens
        e = self._lmrl_engine
        ids = np.asarray(bc_batch["input_ids"])
        real = ids != self.pad_token_id
        lens = np.ascontiguousarray(real.sum(axis=1).astype(np.int32))
        assert (real == (np.arange(ids.shape[1])[None, :] < lens[:, None])).all(), "BC batches are RIGHT-padded (train_ppo_gpt2.py:139-143)"
        n, T = ids.shape
        pitch = max(8, T)
        toks = np.zeros((n, pitch), np.int32)
        toks[:, :T] = np.where(real, ids, 0)
        m = np.asarray(bc_batch["input_training_mask"], dtype=np.float64)[:, 1:] * real[:, 1:]
        fw = forward_tokens(e, 0, 0, [ids[i, :lens[i]] for i in range(n)], logprobs=True, pitch=pitch)
        total = m.sum()
        bc = float(-(fw["logprobs"][:, :T - 1].astype(np.float64) * m).sum() / total)
        dlp = np.zeros((n, pitch), np.float32)
        dlp[:, :T - 1] = -self.bc_loss_weight * m / total
        dv = np.zeros((n, pitch), np.float32)
        keep: list = []
        gs = self._grad_struct(keep)
        self.torch.cuda.synchronize(self.device)
        check(e.lib.lmrl_policy_grad_host(e.handle, toks.ctypes.data, lens.ctypes.data, pitch, n, dlp.ctypes.data, dv.ctypes.data, C.c_float(1.0), C.byref(gs)))
        return bc

    # ---- weight hand-over to the rollout engine (device to device; policy.set_params of ppo/train.py:244, 438) -------
    def push_to_engine(self, tower: int = 0, head_set: int = 0) -> None:
        torch = self.torch
        cfg, L = self.config, self.config.n_layer
        assert self._lmrl_engine is not None, "no engine attached"
        import os as _os
        import time as _time
        _timing = _os.environ.get("LMRL_STEP_TIMING") == "2"
        _t = [_time.perf_counter()]
        torch.cuda.synchronize(self.device)
        _t.append(_time.perf_counter())
        # the pointer tables of the hand-over are built once per set of flat master buffers (views of the flats: the addresses do not
        # change between optimizer steps; rebuilding ~200 views and their ctypes arrays was 1.3 ms of every 13 ms step)
        key = tuple(int(f.data_ptr()) for f in self.flat)
        cached = getattr(self, "_push_cache", None)
        if cached is None or cached[0] != key:
            keep = []
            fp = lambda t: C.cast(C.c_void_p(t.detach().contiguous().data_ptr()), _lib._PF)  # noqa: E731
            w = _lib.lmrl_gpt2_weights()
            P = self.p
            w.wte, w.n_vocab_rows, w.wpe = fp(P["wte"]), int(P["wte"].shape[0]), fp(P["wpe"])
            w.lnf_g, w.lnf_b = fp(P["lnf_g"]), fp(P["lnf_b"])
            for k in _LAYER_KEYS:
                ptrs = (_lib._PF * L)(*[fp(layer[k]) for layer in P["layers"]])
                keep.append(ptrs)
                setattr(w, k, C.cast(ptrs, _lib._PPF))
            dims = _lib.lmrl_gpt2_dims(n_layer=cfg.n_layer, n_head=cfg.n_head, d_model=cfg.d_model, d_ff=cfg.d_ff, n_vocab=cfg.n_vocab, n_pos=cfg.n_pos, ln_eps=cfg.ln_eps)
            self._push_cache = cached = (key, w, dims, keep)
        _, w, dims, _keep = cached
        e = self._lmrl_engine
        _t.append(_time.perf_counter())
        check(e.lib.lmrl_load_gpt2_tower(e.handle, int(tower), C.byref(dims), C.byref(w)))
        _t.append(_time.perf_counter())
        upload_value_heads(e, head_set, v_linear=dict(dense=dict(kernel=self.vh_w.detach().cpu().numpy(), bias=self.vh_b.detach().cpu().numpy())), role="PPO value head")
        if _timing:
            import sys as _sys
            _t.append(_time.perf_counter())
            print("[push_to_engine] sync %.2f ms  pointer tables %.2f ms  lmrl_load_gpt2_tower %.2f ms  value head %.2f ms" %
                  tuple(1e3 * (b - a) for a, b in zip(_t[:-1], _t[1:])), file=_sys.stderr, flush=True)

    def params_numpy(self) -> Dict:
        g = lambda t: t.detach().cpu().numpy().copy()  # noqa: E731
        P = self.p
        return dict(wte=g(P["wte"]), wpe=g(P["wpe"]), lnf_g=g(P["lnf_g"]), lnf_b=g(P["lnf_b"]), layers=[{k: g(L[k]) for k in _LAYER_KEYS} for L in P["layers"]])


def bc_batch_from_str_segments(items, tokenizer, max_length: int, pad_token_id: int) -> Dict[str, np.ndarray]:
    """One BC minibatch from ``[(text, mask_value), ...]`` segment lists — the data format of the PPO drivers' ``bc_data_path``
    (``MaskIterableDataset.blocked_from_str_segments_iterable``, train_ppo_gpt2.py:136-144): every segment is encoded on its own and
    the ids concatenated, each token carrying its segment's mask value; rows are LEFT-truncated and RIGHT-padded to ``max_length``."""
    ids = np.full((len(items), max_length), pad_token_id, np.int32)
    mask = np.zeros((len(items), max_length), np.float32)
    for i, segments in enumerate(items):
        t, m = [], []
        for text, w in segments:
            enc = tokenizer.encode(text)
            t += list(enc)
            m += [float(w)] * len(enc)
        t, m = t[-max_length:], m[-max_length:]
        ids[i, :len(t)], mask[i, :len(m)] = t, m
    return dict(input_ids=ids, input_training_mask=mask)


# ---- agreement across data-parallel ranks ---------------------------------------------------------------------------------
# The reference is one pjit controller: the PPO batch statistics (advantage whitening, mean KL for the adaptive controller) and the
# number of optimizer steps are functions of the GLOBAL batch (ppo/train.py:267-438).  With one process per GPU each rank collects its own
# episodes and may drop some (truncated chains), so these three quantities are agreed on explicitly: every rank must call all_reduce
# the same number of times, or NCCL hangs.

def _dist_group():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 else None


def _all_reduce_f64(values, op: str = "sum") -> np.ndarray:
    """All-reduce a small float64 vector over the ranks (identity without a process group); NCCL needs device tensors."""
    dist = _dist_group()
    v = np.asarray(values, dtype=np.float64)
    if dist is None:
        return v
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(v.copy()).to(dev)
    dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}[op])
    return t.cpu().numpy()


def global_mean_kl(kls: np.ndarray) -> float:
    """Mean per-action-token KL over every rank's PPO data (what ``kl_controller.update`` sees in the reference)."""
    s = _all_reduce_f64([float(np.sum(kls, dtype=np.float64)), float(len(kls))])
    return float(s[0] / s[1]) if s[1] > 0 else 0.0


def whiten_advantages_global(datas: List[PPOData]) -> List[PPOData]:
    """``whiten`` (ppo/base_interface.py:245-251: (x - mean) / sqrt(var + 1e-8)) of the action-token advantages with mean / variance taken
    over ALL ranks' data (three summed scalars); used instead of the per-rank device whitening when a process group is active."""
    acc = np.zeros(3, np.float64)
    for d in datas:
        a = np.asarray(d.old_advantages, np.float64)[np.asarray(d.should_take_action, bool)]
        acc += (a.sum(), (a * a).sum(), a.size)
    acc = _all_reduce_f64(acc)
    if acc[2] == 0:
        return datas
    mean = acc[0] / acc[2]
    var = max(acc[1] / acc[2] - mean * mean, 0.0)
    scale = 1.0 / np.sqrt(var + 1e-8)
    out = []
    for d in datas:
        sta = np.asarray(d.should_take_action, bool)
        adv = np.asarray(d.old_advantages, np.float32).copy()
        adv[sta] = ((adv[sta].astype(np.float64) - mean) * scale).astype(np.float32)
        out.append(d._replace(old_advantages=adv))
    return out


def agreed_steps_per_epoch(n_datas: int, train_bsize: int) -> int:
    """Optimizer steps every rank takes per epoch: the minimum over ranks of len(datas) // train_bsize (a rank with more data drops its
    surplus for the round rather than entering an all_reduce its peers never join)."""
    return int(_all_reduce_f64([float(n_datas // train_bsize)], "min")[0])


def ppo_train_loop(env, policy, inference: B200PPOInference, trainer: B200PPOTrain, kl_controller, *, n_rounds: int, n_rollouts: int, rollout_bsize: int,
                   epochs: int = 1, train_bsize: int = 32, max_length: int = 128, gamma: float = 1.0, lam: float = 0.95,
                   use_advantage_whitening: bool = True, seed_generator=None, rng: Optional[np.random.Generator] = None,
                   max_steps_per_round: Optional[int] = None, use_device_chains: bool = True, on_round: Optional[Callable[[int, Dict], None]] = None,
                   bc_batches=None) -> List[Dict]:
    """ppo/train.py:138-477 reduced to its data flow (no checkpointing / wandb): returns one log dict per round.  ``bc_batches``: an
    iterator of BC minibatches (``bc_batch_from_str_segments``), one drawn per optimizer step like the reference's cycled ``bc_d``
    (ppo/train.py:346-371); used when the trainer's ``bc_loss_weight`` is non-zero."""
    import time
    import torch
    rng = rng or np.random.default_rng(0)
    tok = inference.tokenizer
    history: List[Dict] = []
    for rnd in range(n_rounds):
        t0 = time.perf_counter()
        fused_wordle = getattr(env, "_lmrl_kind", None) == 1 and getattr(env, "_lmrl_engine", None) is inference._lmrl_engine and \
            getattr(env, "_lmrl_fusable", False) and getattr(policy, "fusable", False) and use_device_chains
        raw, token_chains, rollout_s = [], [], 0.0
        for lo in range(0, n_rollouts, rollout_bsize):
            nb = min(rollout_bsize, n_rollouts - lo)
            tb = time.perf_counter()
            r, _ = text_env_eval(env, policy, n_rollouts=nb, bsize=nb, seed_generator=seed_generator, verbose=False)
            rollout_s += time.perf_counter() - tb
            raw += list(r)
            if fused_wordle:
                # the engine still holds this batch's token streams: trajectory -> token chain without a host tokeniser
                dev = device_token_chains(inference._lmrl_engine, nb, max_length)
                for ep, ch in zip(r, dev):
                    if ch is None:
                        tc = episodes_to_chains([ep], tok, max_length, gamma)
                        ch = TokenTrajectoryChain.from_text_trajectory_chain(tc[0], tok) if tc else None
                    if ch is not None:
                        token_chains.append(ch)
        env_steps = int(sum(len(ep) for ep in raw))
        reward_mean = float(np.mean([float(sum(t.reward for t in ep)) for ep in raw])) if not fused_wordle else \
            float(np.mean([float(c.token_trajectory.reward.sum()) for c in token_chains])) if token_chains else 0.0
        t1 = t0 + rollout_s
        if not fused_wordle:
            token_chains = [TokenTrajectoryChain.from_text_trajectory_chain(c, tok) for c in episodes_to_chains(raw, tok, max_length, gamma)]
        multi = _dist_group() is not None
        datas, kls = inference.get_ppo_data_from_token_trajectory_chain(token_chains, bsize=1 << 20, max_length=max_length, gamma=gamma, lam=lam,
                                                                        kl_weight=kl_controller.value,
                                                                        use_advantage_whitening=use_advantage_whitening and not multi, verbose=False)
        if use_advantage_whitening and multi:
            datas = whiten_advantages_global(datas)          # statistics of the global batch, as the single-controller reference has them
        mean_kl = global_mean_kl(kls)                         # identical on every rank: kl_coef cannot diverge between replicas
        kl_controller.update(mean_kl, train_bsize)
        t2 = time.perf_counter()
        losses = []
        steps_per_epoch = agreed_steps_per_epoch(len(datas), train_bsize)     # same all_reduce count on every rank
        # the round's PPO data (thousands of small numpy-holding objects) is parked in the collector's permanent generation while the
        # optimizer steps run: a full collection traversing it in the middle of a 13 ms step showed up as 10-45 ms stalls
        import gc
        gc.collect()
        gc.freeze()
        for _ in range(epochs):
            order = rng.permutation(len(datas))
            for k in range(steps_per_epoch):
                if max_steps_per_round is not None and len(losses) >= max_steps_per_round:
                    break
                lo = k * train_bsize
                batch = block_ppo_data([datas[i] for i in order[lo:lo + train_bsize]], inference.pad_token_id, max_length)
                losses.append(trainer.step(batch, next(bc_batches) if bc_batches is not None else None)[0])
        gc.unfreeze()
        trainer.push_to_engine()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        log = dict(round=rnd, reward_mean=reward_mean, env_steps=env_steps, mean_kl=mean_kl,
                   kl_coef=float(kl_controller.value), n_ppo_data=len(datas), updates=len(losses), loss=float(np.mean(losses)) if losses else None,
                   rollout_s=t1 - t0, ppo_data_s=t2 - t1, update_s=t3 - t2)
        history.append(log)
        if on_round is not None:
            on_round(rnd, log)
    return history
