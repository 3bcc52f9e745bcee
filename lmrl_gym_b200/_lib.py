"""ctypes binding of csrc/liblmrl_b200.so (the C ABI declared in include/lmrl_b200.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails, ``LmrlError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
# LMRL_LIB_PATH: A/B aid — another build of the SAME C ABI (e.g. the previous commit's kernels) instead of the in-tree library
LIB_PATH = os.environ.get("LMRL_LIB_PATH") or os.path.join(HERE, "csrc", "liblmrl_b200.so")


class LmrlError(RuntimeError):
    pass


class lmrl_config(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_envs", C.c_int32), ("max_input_length", C.c_int32),
                ("max_new_tokens", C.c_int32), ("max_turns", C.c_int32), ("kv_pages", C.c_int32), ("reserved", C.c_int32 * 2)]


class lmrl_gpt2_dims(C.Structure):
    _fields_ = [("n_layer", C.c_int32), ("n_head", C.c_int32), ("d_model", C.c_int32), ("d_ff", C.c_int32),
                ("n_vocab", C.c_int32), ("n_pos", C.c_int32), ("ln_eps", C.c_float), ("reserved", C.c_int32)]


_PF = C.POINTER(C.c_float)
_PPF = C.POINTER(_PF)


class lmrl_gpt2_weights(C.Structure):
    _fields_ = [("wte", _PF), ("n_vocab_rows", C.c_int32), ("reserved", C.c_int32), ("wpe", _PF),
                ("ln1_g", _PPF), ("ln1_b", _PPF), ("attn_w", _PPF), ("attn_b", _PPF), ("proj_w", _PPF), ("proj_b", _PPF),
                ("ln2_g", _PPF), ("ln2_b", _PPF), ("fc_w", _PPF), ("fc_b", _PPF), ("fc2_w", _PPF), ("fc2_b", _PPF),
                ("lnf_g", _PF), ("lnf_b", _PF)]


class lmrl_mlp_head(C.Structure):
    _fields_ = [("w1", _PF), ("b1", _PF), ("w2", _PF), ("b2", _PF), ("hidden", C.c_int32), ("n_out", C.c_int32)]


class lmrl_value_heads(C.Structure):
    _fields_ = [("q1", C.POINTER(lmrl_mlp_head)), ("q2", C.POINTER(lmrl_mlp_head)), ("v", C.POINTER(lmrl_mlp_head)),
                ("v_linear_w", _PF), ("v_linear_b", _PF)]


class lmrl_forward_out(C.Structure):
    _fields_ = [("token_logprobs", C.c_void_p), ("v", C.c_void_p), ("q1", C.c_void_p), ("q2", C.c_void_p)]


class lmrl_gpt2_grads(C.Structure):
    _fields_ = [("wte", _PF), ("n_vocab_rows", C.c_int32), ("reserved", C.c_int32), ("wpe", _PF),
                ("ln1_g", _PPF), ("ln1_b", _PPF), ("attn_w", _PPF), ("attn_b", _PPF), ("proj_w", _PPF), ("proj_b", _PPF),
                ("ln2_g", _PPF), ("ln2_b", _PPF), ("fc_w", _PPF), ("fc_b", _PPF), ("fc2_w", _PPF), ("fc2_b", _PPF),
                ("lnf_g", _PF), ("lnf_b", _PF), ("vh_w", _PF), ("vh_b", _PF)]


class lmrl_ppo_batch(C.Structure):
    _fields_ = [("should_take_action", C.c_void_p), ("old_logprobs", C.c_void_p), ("old_values", C.c_void_p), ("old_advantages", C.c_void_p),
                ("old_returns", C.c_void_p), ("cliprange_value", C.c_double), ("cliprange", C.c_double), ("value_loss_coef", C.c_double),
                ("loss_out", C.c_void_p)]


class lmrl_sampling(C.Structure):
    _fields_ = [("temperature", C.c_float), ("do_sample", C.c_int32), ("seed", C.c_uint64), ("top_k", C.c_int32), ("top_p", C.c_float)]


class lmrl_profile_out(C.Structure):
    _fields_ = [("ms", C.c_double * 4), ("work", C.c_double * 4), ("launches", C.c_int64 * 4), ("tag_ms", C.c_double * 12),
                ("tag_work", C.c_double * 12), ("rows", C.c_int64),
                ("sampled", C.c_int64)]


class lmrl_rollout_out(C.Structure):
    _fields_ = [("ep_return", C.c_void_p), ("ep_length", C.c_void_p), ("ep_done", C.c_void_p),
                ("step_reward", C.c_void_p), ("step_valid", C.c_void_p), ("step_feedback", C.c_void_p),
                ("step_aux", C.c_void_p), ("act_tokens", C.c_void_p), ("act_len", C.c_void_p)]


# every symbol include/lmrl_b200.h declares: name -> (restype, argtypes)
_V = C.c_void_p
_I = C.c_int32
SYMBOLS = {
    "lmrl_last_error": (C.c_char_p, []),
    "lmrl_abi_version": (C.c_int, []),
    "lmrl_create": (C.c_int, [C.POINTER(lmrl_config), C.POINTER(_V)]),
    "lmrl_destroy": (C.c_int, [_V]),
    "lmrl_sync": (C.c_int, [_V]),
    "lmrl_device_bytes": (C.c_int64, [_V]),
    "lmrl_stream": (_V, [_V]),
    "lmrl_load_tokenizer": (C.c_int, [_V, _V, _V, _I, _V, _V, _V, _I, _I, _V, _I]),
    "lmrl_load_wordle": (C.c_int, [_V, _V, _I, _I, C.c_float]),
    "lmrl_load_maze": (C.c_int, [_V, _V, _I, _I, _I, _I, _I, _I, _V, _V, _V, _I, _V, _V]),
    "lmrl_load_gpt2": (C.c_int, [_V, C.POINTER(lmrl_gpt2_dims), C.POINTER(lmrl_gpt2_weights)]),
    "lmrl_load_gpt2_tower": (C.c_int, [_V, _I, C.POINTER(lmrl_gpt2_dims), C.POINTER(lmrl_gpt2_weights)]),
    "lmrl_load_q_heads": (C.c_int, [_V, C.POINTER(lmrl_mlp_head), C.POINTER(lmrl_mlp_head), C.c_float]),
    "lmrl_set_policy_kind": (C.c_int, [_V, _I]),
    "lmrl_wordle_reset_host": (C.c_int, [_V, _V, _I]),
    "lmrl_wordle_step_host": (C.c_int, [_V, _V, _V, _I, _V, _I, _V, _V, _V, _V, _V]),
    "lmrl_maze_reset_host": (C.c_int, [_V, _V, _V, _V, _I, _V, _V]),
    "lmrl_maze_step_host": (C.c_int, [_V, _V, _V, _I, _V, _I, _V, _V, _V, _V, _V]),
    "lmrl_wordle_peek_host": (C.c_int, [_V, _I, _V, _V, _V]),
    "lmrl_generate_host": (C.c_int, [_V, _V, _V, _I, _I, _V, _I, C.POINTER(lmrl_sampling), _V, _V]),
    "lmrl_logits_host": (C.c_int, [_V, _V, _V, _I, _I, _V]),
    "lmrl_score_host": (C.c_int, [_V, _V, _V, _I, _I, _V, C.c_float, _V, _V]),
    "lmrl_load_value_heads": (C.c_int, [_V, _I, C.POINTER(lmrl_value_heads)]),
    "lmrl_forward_host": (C.c_int, [_V, _I, _I, _V, _V, _I, _I, C.c_float, C.POINTER(lmrl_forward_out)]),
    "lmrl_chess_new": (_V, [C.c_char_p]),
    "lmrl_chess_free": (None, [_V]),
    "lmrl_chess_fen": (C.c_int, [_V, C.c_char_p, _I]),
    "lmrl_chess_legal_moves": (C.c_int, [_V, C.c_char_p, _I]),
    "lmrl_chess_push_san": (C.c_int, [_V, C.c_char_p, C.c_char_p, _I]),
    "lmrl_chess_push_uci": (C.c_int, [_V, C.c_char_p, C.c_char_p, _I]),
    "lmrl_chess_pop": (C.c_int, [_V]),
    "lmrl_chess_status": (C.c_int, [_V]),
    "lmrl_chess_turn": (C.c_int, [_V]),
    "lmrl_chess_is_valid": (C.c_int, [_V]),
    "lmrl_chess_perft": (C.c_int64, [_V, _I]),
    "lmrl_textnav_reset_host": (C.c_int, [_V, _V, _V, _I, _V, _V, _V]),
    "lmrl_textnav_step_host": (C.c_int, [_V, _V, _V, _I, _V, _I, _V, _V, _V, _V, _V, _V]),
    "lmrl_ppo_advantages_host": (C.c_int, [_V, _I, _V, _V, _V, _V, _V, _V, C.c_double, C.c_double, C.c_double, _I, _V, _V, _V]),
    "lmrl_ppo_loss_host": (C.c_int, [_V, C.c_int64, _V, _V, _V, _V, _V, _V, _V, C.c_double, C.c_double, C.c_double, _V]),
    "lmrl_rollout_chains_host": (C.c_int, [_V, _I, _I, _V, _V, _V, _V, _V, _V]),
    "lmrl_train_enable": (C.c_int, [_V, _I, _I]),
    "lmrl_policy_grad_host": (C.c_int, [_V, _V, _V, _I, _I, _V, _V, C.c_float, C.POINTER(lmrl_gpt2_grads)]),
    "lmrl_ppo_grad_host": (C.c_int, [_V, _V, _V, _I, _I, C.POINTER(lmrl_ppo_batch), C.POINTER(lmrl_gpt2_grads)]),
    "lmrl_adamw_dev": (C.c_int, [_V, _V, _V, _V, _V, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_double]),
    "lmrl_rollout_host": (C.c_int, [_V, _I, _V, _V, _I, C.POINTER(lmrl_sampling), C.POINTER(lmrl_rollout_out)]),
    "lmrl_rollout_dev": (C.c_int, [_V, _I, _V, _V, _I, C.POINTER(lmrl_sampling), C.POINTER(lmrl_rollout_out)]),
    "lmrl_wordle_scripted_rollout_dev": (C.c_int, [_V, _V, _I, C.c_float, C.c_uint64, C.POINTER(lmrl_rollout_out)]),
    "lmrl_set_use_graphs": (C.c_int, [_V, _I]),
    "lmrl_set_two_lanes": (C.c_int, [_V, _I]),
    "lmrl_set_use_pdl": (C.c_int, [_V, _I]),
    "lmrl_set_option": (C.c_int, [_V, C.c_char_p, _I]),
    "lmrl_get_option": (C.c_int, [_V, C.c_char_p, _V]),
    "lmrl_kv_stats": (C.c_int, [_V, _V, _V, _V]),
    "lmrl_launch_count": (C.c_int64, [_V, _I]),
    "lmrl_set_gemm_impl": (C.c_int, [_V, _I]),
    "lmrl_test_gemm_host": (C.c_int, [_V, _I, _I, _I, _I, _V, _V, _V, _V]),
    "lmrl_bench_kernels": (C.c_int, [_V, _I, _I, _I, _V]),
    "lmrl_profile_enable": (C.c_int, [_V, _I]),
    "lmrl_profile_read": (C.c_int, [_V, C.POINTER(lmrl_profile_out)]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the CUDA library; raise loudly if it is absent (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LmrlError(f"{LIB_PATH} is missing: build it with `python -m lmrl_gym_b200.build` "
                        "(nvcc, sm_100a). There is no CPU fallback for the rollout engine.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        msg = load().lmrl_last_error()
        raise LmrlError(msg.decode("utf-8", "replace") if msg else f"lmrl call failed with status {status}")


def ptr(a) -> int:
    """numpy array -> void* (None -> NULL)."""
    if a is None:
        return None
    return a.ctypes.data
