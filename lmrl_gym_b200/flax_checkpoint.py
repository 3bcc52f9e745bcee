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
            shape = x["shape"]
            if isinstance(shape, dict):               # flax.serialization._chunk stores the shape as _tuple_to_dict: {'0': d0, '1': d1, ...}
                shape = [shape[str(i)] for i in range(len(shape))]
            return flat.reshape(tuple(int(d) for d in shape))
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


def _ext(a) -> Any:
    import msgpack
    a = np.ascontiguousarray(a)
    return msgpack.ExtType(1, msgpack.packb((a.shape, a.dtype.name, a.tobytes()), use_bin_type=True))


def flax_tree_from_gpt2_params(params: Dict) -> Dict:
    """The inverse of ``gpt2_params_from_flax_tree``: numpy leaves, Conv1D kernels back in FlaxConv1D's ``[features, in]`` layout."""
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
    kT = lambda a: f32(np.asarray(a).T)  # noqa: E731
    h = {}
    for i, L in enumerate(params["layers"]):
        h[str(i)] = dict(ln_1=dict(scale=f32(L["ln1_g"]), bias=f32(L["ln1_b"])), ln_2=dict(scale=f32(L["ln2_g"]), bias=f32(L["ln2_b"])),
                         attn=dict(c_attn=dict(kernel=kT(L["attn_w"]), bias=f32(L["attn_b"])), c_proj=dict(kernel=kT(L["proj_w"]), bias=f32(L["proj_b"]))),
                         mlp=dict(c_fc=dict(kernel=kT(L["fc_w"]), bias=f32(L["fc_b"])), c_proj=dict(kernel=kT(L["fc2_w"]), bias=f32(L["fc2_b"]))))
    return dict(transformer=dict(wte=dict(embedding=f32(params["wte"])), wpe=dict(embedding=f32(params["wpe"])), h=h,
                                 ln_f=dict(scale=f32(params["lnf_g"]), bias=f32(params["lnf_b"]))))


def write_msgpack_tree(path: str, tree: Dict) -> None:
    """One msgpack document of the nested state dict, leaves as ext type 1 (the ``flax.serialization.msgpack_serialize`` shape)."""
    import msgpack

    def enc(x):
        return {k: enc(v) for k, v in x.items()} if isinstance(x, dict) else _ext(x)
    with open(path, "wb") as f:
        f.write(msgpack.packb(enc(tree), use_bin_type=True))


def save_reference_checkpoint(model_dir: str, params: Dict, cfg: GPT2Config) -> None:
    """``<model_dir>/params.msgpack`` + ``<model_dir>/config.json`` in the layout ``dump_state`` writes (ppo/train.py:28-85)."""
    os.makedirs(model_dir, exist_ok=True)
    write_msgpack_tree(os.path.join(model_dir, "params.msgpack"), flax_tree_from_gpt2_params(params))
    with open(os.path.join(model_dir, "config.json"), "w") as f:
        json.dump(dict(model_type="gpt2", n_layer=cfg.n_layer, n_head=cfg.n_head, n_embd=cfg.d_model, n_inner=cfg.d_ff, vocab_size=cfg.n_vocab_rows,
                       n_positions=cfg.n_pos, layer_norm_epsilon=cfg.ln_eps, activation_function="gelu_new"), f)
