"""Engine: one GPU's worth of device-resident rollout state behind the C ABI (include/lmrl_b200.h)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import LmrlError, check, ptr
from .tokenizer import ByteTokenizer

ENV_WORDLE, ENV_MAZE = 1, 2


class Engine:
    """Owns an ``lmrl_engine`` handle.  Envs and policies bound to the same Engine can be fused by
    ``text_env_eval`` / ``interact_environment`` (lmrl_gym_b200/environment.py)."""

    def __init__(self, max_envs: int, device: int = 0, max_input_length: int = 256, max_new_tokens: int = 6,
                 max_turns: int = 6, tokenizer: Optional[ByteTokenizer] = None, kv_pages: int = 0):
        """``kv_pages``: size of the paged KV cache in pages of 32 positions (0 = every env at full length, the default)."""
        self.lib = _lib.load()
        self.cfg = _lib.lmrl_config(device=device, max_envs=max_envs, max_input_length=max_input_length,
                                    max_new_tokens=max_new_tokens, max_turns=max_turns, kv_pages=kv_pages)
        self.handle = C.c_void_p()
        check(self.lib.lmrl_create(C.byref(self.cfg), C.byref(self.handle)))
        self.max_envs = max_envs
        self.device = device
        self.max_input_length = max_input_length
        self.max_new_tokens = max_new_tokens
        self.max_turns = max_turns
        self.tokenizer: Optional[ByteTokenizer] = None
        self.env_kind = 0
        self._slots = {}          # (kind, index) -> role of what was loaded there (claim_slot)
        if tokenizer is not None:
            self.load_tokenizer(tokenizer)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.lmrl_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- tables ----
    def load_tokenizer(self, tok: ByteTokenizer, header: str = "Wordle:\n"):
        t = tok.device_tables()
        self._tok_tables = t
        hdr = np.asarray(tok.encode(header), dtype=np.int32)
        check(self.lib.lmrl_load_tokenizer(self.handle, ptr(t["token_bytes"]), ptr(t["offsets"]), len(tok),
                                           ptr(t["byte_tok"]), ptr(t["tok_first"]), ptr(t["tok_space"]),
                                           tok.encode("a b")[1] if False else int(t["byte_tok"][0x20]), tok.newline_id(),
                                           ptr(hdr), len(hdr)))
        self.tokenizer = tok

    @property
    def stream_ptr(self) -> int:
        return int(self.lib.lmrl_stream(self.handle) or 0)

    def profile_enable(self, enable: bool = True):
        check(self.lib.lmrl_profile_enable(self.handle, 1 if enable else 0))

    def profile_read(self):
        out = _lib.lmrl_profile_out()
        check(self.lib.lmrl_profile_read(self.handle, C.byref(out)))
        cats = ("gemm", "attention", "elementwise", "env")
        return dict(ms={c: out.ms[i] for i, c in enumerate(cats)}, work={c: out.work[i] for i, c in enumerate(cats)},
                    launches={c: int(out.launches[i]) for i, c in enumerate(cats)}, rows=int(out.rows), sampled=int(out.sampled),
                    tags={f"{k}_{m}": dict(ms=out.tag_ms[2 * i + j], work=out.tag_work[2 * i + j])
                          for i, k in enumerate(("gemm_qkv", "gemm_proj", "gemm_fc", "gemm_fc2", "gemm_head", "attention"))
                          for j, m in enumerate(("prefill", "decode"))})

    def rollout_dev(self, env_kind: int, seeds_dev_ptr: int, n: int, sampling, out_struct=None):
        """Enqueue one fused rollout with seeds already resident in HBM; no host synchronisation."""
        check(self.lib.lmrl_rollout_dev(self.handle, env_kind, seeds_dev_ptr, None, n, C.byref(sampling),
                                        None if out_struct is None else C.byref(out_struct)))

    def wordle_scripted_rollout_dev(self, seeds_dev_ptr: int, n: int, prob_smart: float, policy_seed: int, out_struct):
        """Env-only rollout with the device equivalent of ``RandomMixturePolicy(prob_smart)`` (scripted_policies.py:114-126);
        all pointers are device pointers; ``act_tokens[..., 0]`` receives the packed 5-letter guess of every turn."""
        check(self.lib.lmrl_wordle_scripted_rollout_dev(self.handle, seeds_dev_ptr, n, float(prob_smart), int(policy_seed) & ((1 << 64) - 1),
                                                        C.byref(out_struct)))

    def set_use_graphs(self, enable: bool):
        """CUDA-graph replay of fused rollouts (default on)."""
        check(self.lib.lmrl_set_use_graphs(self.handle, 1 if enable else 0))

    def set_two_lanes(self, enable: bool):
        """Two concurrent half-batch lanes inside fused rollouts (default off: measured slower on B200, results identical)."""
        check(self.lib.lmrl_set_two_lanes(self.handle, 1 if enable else 0))

    def set_use_pdl(self, enable: bool):
        """Programmatic dependent launch between the kernels of a pass (default off: measured slower on B200)."""
        check(self.lib.lmrl_set_use_pdl(self.handle, 1 if enable else 0))

    def claim_slot(self, kind: str, index: int, role: str):
        """The engine has three parameter towers and two value-head sets, and several host classes load into the same indices (tower 1:
        the frozen initial policy of PPO *or* the value base of ILQL / MC; head set 0: the PPO value head *or* the Q / V heads).  Loading a
        second role into an occupied slot would silently overwrite the first object's weights; refuse it instead."""
        cur = self._slots.get((kind, index))
        if cur is not None and cur != role:
            raise LmrlError(f"engine {kind} {index} already holds the {cur}: loading the {role} there would overwrite it "
                            f"(bind the two objects to separate engines)")
        self._slots[(kind, index)] = role

    def release_slots(self):
        """Forget what the towers / head sets hold (a caller that is done with the objects bound so far and reuses the engine)."""
        self._slots.clear()

    def set_option(self, name: str, value: int):
        """Named integer switch of the policy engine (include/lmrl_b200.h::lmrl_set_option): fuse_ln, attn_wide, gemv_max_rows."""
        check(self.lib.lmrl_set_option(self.handle, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        v = C.c_int32(0)
        check(self.lib.lmrl_get_option(self.handle, name.encode(), C.addressof(v)))
        return int(v.value)

    def set_gemm_impl(self, impl: int):
        """1 = tcgen05/TMEM/TMA (default), 0 = the mma.sync kernel (independent cross-check used by the tests)."""
        check(self.lib.lmrl_set_gemm_impl(self.handle, int(impl)))

    def test_gemm(self, impl: int, A_bf16, W_bf16, bias, resid):
        """Unit-test hook: returns resid + A @ W.T + bias computed by the chosen GEMM kernel (A, W as bf16 bit patterns)."""
        import numpy as np
        M, K = A_bf16.shape
        N = W_bf16.shape[0]
        out = np.ascontiguousarray(resid, dtype=np.float32).copy()
        check(self.lib.lmrl_test_gemm_host(self.handle, int(impl), M, N, K, ptr(np.ascontiguousarray(A_bf16)),
                                           ptr(np.ascontiguousarray(W_bf16)), None if bias is None else ptr(np.ascontiguousarray(bias, dtype=np.float32)),
                                           ptr(out)))
        return out

    def bench_kernels(self, M: int, ctx: int, iters: int = 20):
        """Mean device ms of the layer kernels on M synthetic rows (see include/lmrl_b200.h::lmrl_bench_kernels)."""
        out = np.zeros(8, dtype=np.float32)
        check(self.lib.lmrl_bench_kernels(self.handle, int(M), int(ctx), int(iters), ptr(out)))
        return dict(zip(("qkv", "proj", "fc", "fc2", "head", "attention", "layernorm", "embed"), out.tolist()))

    def kv_stats(self):
        """Paged KV cache: (pages in use since the current batch started, pool capacity, positions per page); raises if the pool overflowed."""
        a, b, c = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        check(self.lib.lmrl_kv_stats(self.handle, C.addressof(a), C.addressof(b), C.addressof(c)))
        return int(a.value), int(b.value), int(c.value)

    def sync(self):
        check(self.lib.lmrl_sync(self.handle))

    @property
    def device_bytes(self) -> int:
        return int(self.lib.lmrl_device_bytes(self.handle))

    def launch_count(self, reset: bool = False) -> int:
        return int(self.lib.lmrl_launch_count(self.handle, 1 if reset else 0))
