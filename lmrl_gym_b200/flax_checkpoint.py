"""Reading the reference's checkpoints (``policy/params.msgpack`` + ``policy/config.json``, ``value_head/params.msgpack``; written by
``dump_state``, LLM_RL/algorithms/ppo/train.py:28-85, through JaxSeq's ``save_pytree``) without flax or JaxSeq.

PARITY UNPINNED: neither flax nor JaxSeq is vendored by the reference or present in this image, so the byte format below is a restatement
of their published serialisation, accepted in both shapes it has had:
  * one msgpack document = the nested state dict (``flax.serialization.msgpack_serialize``);
  * a stream of ``(key_path, serialised_leaf)`` msgpack records, one per flattened leaf (the streaming writer JaxSeq's checkpointing uses).
Leaves are msgpack ext type 1 = ``(shape, dtype_name, raw_bytes)``; arrays too large for one msgpack bin are dicts
``{'__msgpack_chunked_array__': True, 'shape': ..., 'chunks': {'0': ..., ...}}``.  bfloat16 leaves (``save_bf16=True``, the drivers'
default) are widened to float32.  The tests write both shapes with a few lines of msgpack and read them back; a file produced by the real
stack should be tried before trusting this reader.

Parameter tree of HF ``FlaxGPT2LMHeadModel``: ``transformer/{wte,wpe}/embedding``, ``transformer/h/<i>/{ln_1,ln_2}/{scale,bias}``,
``.../attn/{c_attn,c_proj}/{kernel,bias}``, ``.../mlp/{c_fc,c_proj}/{kernel,bias}``, ``transformer/ln_f/{scale,bias}``.  ``FlaxConv1D``
stores its kernel as ``[features, in]`` — the transpose of the torch ``Conv1D`` layout this package uses — which ``c_attn`` (3d x d)
lets the loader verify.
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, Tuple

import numpy as np

from .gpt2 import GPT2Config


def _leaf_from_ext(code: int, data: bytes):
    import msgpack
    if code == 1:                                   # ndarray
        shape, dtype_name, buf = msgpack.unpackb(data, raw=True)
        name = dtype_name.decode() if isinstance(dtype_name, bytes) else str(dtype_name)
        if name == "bfloat16":
            u = np.frombuffer(buf, dtype=np.uint16).astype(np.uint32) << 16
            return u.view(np.float32).reshape(tuple(shape))
        return np.frombuffer(buf, dtype=np.dtype(name)).reshape(tuple(shape)).copy()
    if code == 3:                                   # numpy scalar: same triple, shape ()
        shape, dtype_name, buf = msgpack.unpackb(data, raw=True)
        name = dtype_name.decode() if isinstance(dtype_name, bytes) else str(dtype_name)
        return np.frombuffer(buf, dtype=np.dtype(name))[0]
    return msgpack.ExtType(code, data)


def _text_keys(x: Any) -> Any:
    if isinstance(x, dict):
        return {(k.decode() if isinstance(k, bytes) else str(k)): _text_keys(v) for k, v in x.items()}
    return x


def _unchunk(x: Any) -> Any:
    if isinstance(x, dict):
        if x.get("__msgpack_chunked_array__"):
            chunks = x["chunks"]
            flat = np.concatenate([np.asarray(chunks[str(i)]).reshape(-1) for i in range(len(chunks))])
            return flat.reshape(tuple(x["shape"]))
        return {k: _unchunk(v) for k, v in x.items()}
    return x


def read_msgpack_tree(path: str) -> Dict:
    """-> nested dict of numpy arrays, whichever of the two file shapes ``path`` has."""
    import msgpack
    with open(path, "rb") as f:
        unpacker = msgpack.Unpacker(f, ext_hook=_leaf_from_ext, raw=False, strict_map_key=False, max_buffer_size=0)
        first = next(unpacker)
        if isinstance(first, dict):                 # one document
            return _unchunk(_text_keys(first))
        tree: Dict = {}
        rec = first
        while True:                                 # stream of (key path, leaf bytes)
            key, value = rec
            leaf = msgpack.unpackb(value, ext_hook=_leaf_from_ext, raw=False, strict_map_key=False) if isinstance(value, (bytes, bytearray)) else value
            node = tree
            parts = [k.decode() if isinstance(k, bytes) else str(k) for k in (key if isinstance(key, (list, tuple)) else [key])]
            for p in parts[:-1]:
                node = node.setdefault(p, {})
            node[parts[-1]] = _unchunk(_text_keys(leaf))
            try:
                rec = next(unpacker)
            except StopIteration:
                return tree


def config_from_hf_json(path: str) -> GPT2Config:
    with open(path) as f:
        c = json.load(f)
    d = int(c["n_embd"])
    return GPT2Config(n_layer=int(c["n_layer"]), n_head=int(c["n_head"]), d_model=d, d_ff=int(c.get("n_inner") or 4 * d),
                      n_vocab=min(int(c["vocab_size"]), 50257), n_vocab_rows=int(c["vocab_size"]), n_pos=int(c["n_positions"]),
                      ln_eps=float(c.get("layer_norm_epsilon", 1e-5)))


def gpt2_params_from_flax_tree(tree: Dict, cfg: GPT2Config) -> Dict:
    """Flax GPT-2 parameter tree -> the dict layout of ``init_gpt2_params`` (fp32; Conv1D kernels as ``[in, out]``)."""
    t = tree.get("params", tree)
    t = t["transformer"]
    d = cfg.d_model
    f32 = lambda a: np.ascontiguousarray(np.asarray(a), dtype=np.float32)  # noqa: E731
    c_attn = np.asarray(t["h"]["0"]["attn"]["c_attn"]["kernel"])
    assert c_attn.shape in ((3 * d, d), (d, 3 * d)), f"unexpected c_attn kernel shape {c_attn.shape}"
    flip = c_attn.shape == (3 * d, d)               # FlaxConv1D layout [features, in]
    k = (lambda a: f32(np.asarray(a).T)) if flip else f32
    layers = []
    for i in range(cfg.n_layer):
        h = t["h"][str(i)]
        layers.append(dict(ln1_g=f32(h["ln_1"]["scale"]), ln1_b=f32(h["ln_1"]["bias"]),
                           attn_w=k(h["attn"]["c_attn"]["kernel"]), attn_b=f32(h["attn"]["c_attn"]["bias"]),
                           proj_w=k(h["attn"]["c_proj"]["kernel"]), proj_b=f32(h["attn"]["c_proj"]["bias"]),
                           ln2_g=f32(h["ln_2"]["scale"]), ln2_b=f32(h["ln_2"]["bias"]),
                           fc_w=k(h["mlp"]["c_fc"]["kernel"]), fc_b=f32(h["mlp"]["c_fc"]["bias"]),
                           fc2_w=k(h["mlp"]["c_proj"]["kernel"]), fc2_b=f32(h["mlp"]["c_proj"]["bias"])))
        assert layers[-1]["attn_w"].shape == (d, 3 * d) and layers[-1]["fc_w"].shape == (d, cfg.d_ff) and layers[-1]["fc2_w"].shape == (cfg.d_ff, d)
    return dict(wte=f32(t["wte"]["embedding"]), wpe=f32(t["wpe"]["embedding"]), lnf_g=f32(t["ln_f"]["scale"]), lnf_b=f32(t["ln_f"]["bias"]),
                layers=layers)


def head_params_from_flax_tree(tree: Dict) -> Dict:
    """``value_head/params.msgpack`` / ``q1_head/params.msgpack``: flax Dense kernels are ``[in, out]`` already."""
    t = tree.get("params", tree)
    return {layer: dict(kernel=np.ascontiguousarray(p["kernel"], dtype=np.float32),
                        bias=np.ascontiguousarray(p["bias"], dtype=np.float32) if "bias" in p else None) for layer, p in t.items()}


def load_reference_checkpoint(model_dir: str) -> Tuple[Dict, GPT2Config]:
    """``<model_dir>/params.msgpack`` (or ``train_state.msgpack``, whose ``params`` entry is taken) + ``<model_dir>/config.json``."""
    cfg = config_from_hf_json(os.path.join(model_dir, "config.json"))
    path = os.path.join(model_dir, "params.msgpack")
    if not os.path.exists(path):
        path = os.path.join(model_dir, "train_state.msgpack")
    return gpt2_params_from_flax_tree(read_msgpack_tree(path), cfg), cfg
