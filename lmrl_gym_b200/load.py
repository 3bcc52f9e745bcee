"""Weights in and out of the engine's parameter layout — what the task drivers call where the reference calls JaxSeq's
``load_params(model_load_mode, model_load_path, ...)`` (e.g. llm_rl_scripts/wordle/bc/eval_bc_gpt2.py:64-74; JaxSeq itself is not
vendored, so the file formats here are this repository's own).

``ModelLoadMode.CONFIG``  path = ``gpt2`` | ``gpt2-medium`` | a JSON file of ``GPT2Config`` fields → seeded random init (the
                          reference's CONFIG mode: architecture from a config, fresh weights)
``ModelLoadMode.HF``      path = a LOCAL directory ``transformers.GPT2LMHeadModel.from_pretrained`` can read (torch weights; there is
                          no network here, so hub names only work from a warm cache)
``ModelLoadMode.PARAMS``  path = an ``.npz`` written by ``save_params`` (fp32 arrays, flattened keys), or a DIRECTORY holding the reference's own
                          ``params.msgpack`` / ``train_state.msgpack`` + ``config.json`` (``flax_checkpoint.py``: format restated, unpinned)
"""
from __future__ import annotations

import enum
import json
import os
from typing import Dict, Tuple

import numpy as np

from .gpt2 import GPT2Config, _LAYER_KEYS, init_gpt2_params, params_from_hf_state_dict


class ModelLoadMode(enum.Enum):
    CONFIG = "config"
    HF = "hf"
    PARAMS = "params"


_NAMED = {"gpt2": GPT2Config.small, "gpt2-small": GPT2Config.small, "gpt2-medium": GPT2Config.medium}


def save_params(path: str, params: Dict, cfg: GPT2Config) -> None:
    flat = {k: np.asarray(params[k], np.float32) for k in ("wte", "wpe", "lnf_g", "lnf_b")}
    for l, layer in enumerate(params["layers"]):
        for k in _LAYER_KEYS:
            flat[f"h{l}.{k}"] = np.asarray(layer[k], np.float32)
    flat["config_json"] = np.frombuffer(json.dumps(cfg.__dict__).encode(), dtype=np.uint8)
    np.savez(path, **flat)


def load_params(model_load_mode: ModelLoadMode, model_load_path: str, seed: int = 2, force_pad_embeddings: bool = False) -> Tuple[Dict, GPT2Config]:
    """-> (params, config).  ``seed`` plays the part of the scripts' ``model_prng_key = PRNGKey(2)``.  ``force_pad_embeddings`` appends
    the ``<|pad|>`` row to a 50257-row embedding (the drivers add that token to the tokenizer, eval_bc_gpt2.py:52-53)."""
    mode = ModelLoadMode(model_load_mode) if not isinstance(model_load_mode, ModelLoadMode) else model_load_mode
    if mode is ModelLoadMode.CONFIG:
        if model_load_path in _NAMED:
            cfg = _NAMED[model_load_path]()
        else:
            with open(model_load_path) as f:
                cfg = GPT2Config(**json.load(f))
        return init_gpt2_params(cfg, seed=seed), cfg
    if mode is ModelLoadMode.PARAMS and os.path.isdir(model_load_path):
        from .flax_checkpoint import load_reference_checkpoint
        params, cfg = load_reference_checkpoint(model_load_path)
    elif mode is ModelLoadMode.PARAMS:
        z = np.load(model_load_path)
        cfg = GPT2Config(**json.loads(bytes(z["config_json"]).decode()))
        layers = [{k: z[f"h{l}.{k}"] for k in _LAYER_KEYS} for l in range(cfg.n_layer)]
        params = dict(wte=z["wte"], wpe=z["wpe"], lnf_g=z["lnf_g"], lnf_b=z["lnf_b"], layers=layers)
    else:
        if not os.path.isdir(model_load_path):
            raise FileNotFoundError(f"ModelLoadMode.HF needs a local model directory (no network): {model_load_path}")
        from transformers import GPT2LMHeadModel
        model = GPT2LMHeadModel.from_pretrained(model_load_path)
        hc = model.config
        cfg = GPT2Config(n_layer=hc.n_layer, n_head=hc.n_head, d_model=hc.n_embd, d_ff=hc.n_inner or 4 * hc.n_embd,
                         n_vocab=min(hc.vocab_size, 50257), n_vocab_rows=hc.vocab_size, n_pos=hc.n_positions, ln_eps=hc.layer_norm_epsilon)
        params = params_from_hf_state_dict(model.state_dict(), cfg)
    if params["wte"].shape[0] == cfg.n_vocab and (force_pad_embeddings or cfg.n_vocab == 50257):
        pad_row = params["wte"].mean(axis=0, keepdims=True).astype(np.float32)      # never sampled (ids >= n_vocab are masked), only embedded
        params["wte"] = np.concatenate([params["wte"], pad_row], axis=0)
        cfg.n_vocab_rows = params["wte"].shape[0]
    return params, cfg


def save_head_params(path: str, head: Dict) -> None:
    """An MLPHead (``{'dense1': ..., 'dense2': ...}``) or LinearHead (``{'dense': ...}``) parameter dict -> ``.npz``."""
    flat = {}
    for layer, p in head.items():
        flat[layer + ".kernel"] = np.asarray(p["kernel"], np.float32)
        if p.get("bias") is not None:
            flat[layer + ".bias"] = np.asarray(p["bias"], np.float32)
    np.savez(path, **flat)


def load_head_params(model_load_mode, model_load_path: str, cfg: GPT2Config, kind: str = "mlp", seed: int = 0, n_out: int = None) -> Dict:
    """What the value-RL drivers call where the reference calls ``LLM_RL.heads.{mlp,linear}_head.load_params`` (eval_ilql_gpt2.py:92-110,
    eval_mc_gpt2.py:93-101).  CONFIG: seeded init with the training scripts' head configs — MLPHead(d, d, vocab, layer2_bias_init=-4.4)
    (train_ilql_gpt2.py:224-231) or LinearHead(d, vocab, initializer_range=0.0, bias_init=-4.4)-style (train_mc_returns_gpt2.py:198-206);
    PARAMS: an ``.npz`` written by ``save_head_params``."""
    from .gpt2 import init_mlp_head_params
    mode = ModelLoadMode(model_load_mode) if not isinstance(model_load_mode, ModelLoadMode) else model_load_mode
    n_out = n_out or cfg.n_vocab_rows
    if mode is ModelLoadMode.PARAMS and (os.path.isdir(model_load_path) or model_load_path.endswith(".msgpack")):
        from .flax_checkpoint import head_params_from_flax_tree, read_msgpack_tree
        path = model_load_path if model_load_path.endswith(".msgpack") else os.path.join(model_load_path, "params.msgpack")
        return head_params_from_flax_tree(read_msgpack_tree(path))
    if mode is ModelLoadMode.PARAMS:
        z = np.load(model_load_path)
        layers = sorted({k.split(".")[0] for k in z.files})
        return {l: dict(kernel=z[l + ".kernel"], bias=z[l + ".bias"] if l + ".bias" in z.files else None) for l in layers}
    if mode is not ModelLoadMode.CONFIG:
        raise ValueError("heads load from 'config' (seeded init) or 'params' (.npz)")
    if kind == "mlp":
        return init_mlp_head_params(cfg, seed=seed, n_out=n_out)
    rng = np.random.default_rng(seed)
    return dict(dense=dict(kernel=(rng.standard_normal((cfg.d_model, n_out)) / np.sqrt(cfg.d_model)).astype(np.float32),
                           bias=np.full(n_out, -4.4, np.float32)))
