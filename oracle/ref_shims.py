"""Import shims that let the reference's OWN algorithm code run in the build container (test infrastructure only).

The reference's PPO / ILQL data paths are plain array code (``LLM_RL/algorithms/ppo/base_interface.py:230-669``,
``ilql/gpt2/score_fn.py``, ``mc_returns/score_fn.py``) but the files import jax / flax / optax / JaxSeq at module top,
none of which exist here.  ``install()`` registers stand-in modules so that those files import unmodified from
/root/reference and their array code executes on numpy:

  * ``jax.numpy``  -> numpy, with results viewed as ``AtArray`` (adds the functional ``x.at[i].set(v) / .add(v)``);
    ``jax.nn.log_softmax``, ``jax.device_get``, ``jax.lax.stop_gradient``, ``jax.random.split`` trivial;
    ``pjit`` / ``jax.jit`` are identity decorators.
  * ``optax.softmax_cross_entropy_with_integer_labels`` -> logsumexp(logits) - logits[label] in float32.
  * ``flax.struct.PyTreeNode`` -> a plain base class (keyword constructor + ``replace``), ``struct.field`` -> its default.
  * ``JaxSeq.utils.block_sequences / BlockingStrategy / Padding / Truncation`` -> restated from JaxSEQ's published
    semantics (pad or truncate every sequence to ``max_length``, or to the longest when None, on the chosen side);
    ``multihost_device_get`` identity.  JaxSEQ itself (Sea-Snell/JaxSEQ @ HEAD, unpinned) is not vendored by the reference.

Only ``oracle/gen_golden_algos.py`` uses this, to write ``tests/golden/{score_fns,ppo_data}.json``.  Nothing here is
imported by the product, the tests or the bench.
"""
from __future__ import annotations

import enum
import importlib.machinery
import os
import sys
import types
from typing import Any, List, NamedTuple, Optional

import numpy as np

REF = os.environ.get("LMRL_REFERENCE", "/root/reference")


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v):
        out = np.array(self.arr, copy=True)
        out[self.idx] = v
        return out.view(AtArray)

    def add(self, v):
        out = np.array(self.arr, copy=True)
        out[self.idx] += v
        return out.view(AtArray)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class AtArray(np.ndarray):
    @property
    def at(self):
        return _At(self)


def _wrap(x):
    if isinstance(x, np.ndarray) and not isinstance(x, AtArray):
        return x.view(AtArray)
    if isinstance(x, tuple):
        return tuple(_wrap(y) for y in x)
    return x


class _NumpyAsJnp(types.ModuleType):
    def __getattr__(self, name):
        attr = getattr(np, name)
        if callable(attr) and not isinstance(attr, type):
            def fn(*a, **k):
                return _wrap(attr(*a, **k))
            fn.__name__ = name
            return fn
        return attr


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []  # behaves as a package for `import a.b`
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _identity_decorator(*dargs, **dkwargs):
    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]
    return lambda f: f


class _PyTreeNode:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def replace(self, **kw):
        out = self.__class__.__new__(self.__class__)
        out.__dict__.update(self.__dict__)
        out.__dict__.update(kw)
        return out


def _struct_field(pytree_node=True, default=None, **kw):
    return default


class Padding(enum.Enum):
    LEFT = 0
    RIGHT = 1


class Truncation(enum.Enum):
    LEFT = 0
    RIGHT = 1


class BlockingStrategy(NamedTuple):
    padding: Padding
    truncation: Truncation
    max_length: Optional[int]


def block_sequences(sequences: List[Any], pad_value: Any, dtype: Any, blocking_strategy: BlockingStrategy) -> np.ndarray:
    """JaxSeq.utils.block_sequences, restated: [n, max_length] with every sequence truncated / padded on the given sides."""
    max_len = blocking_strategy.max_length
    if max_len is None:
        max_len = max(len(s) for s in sequences)
    out = np.full((len(sequences), max_len), pad_value, dtype=dtype)
    for i, s in enumerate(sequences):
        s = list(s)
        if len(s) > max_len:
            s = s[-max_len:] if blocking_strategy.truncation == Truncation.LEFT else s[:max_len]
        if blocking_strategy.padding == Padding.LEFT:
            out[i, max_len - len(s):] = s
        else:
            out[i, :len(s)] = s
    return out


def softmax_cross_entropy_with_integer_labels(logits, labels):
    logits = np.asarray(logits, dtype=np.float32)
    m = logits.max(axis=-1, keepdims=True)
    lse = (np.log(np.exp(logits - m).sum(axis=-1, dtype=np.float32)) + m[..., 0]).astype(np.float32)
    picked = np.take_along_axis(logits, np.asarray(labels)[..., None], axis=-1)[..., 0]
    return _wrap((lse - picked).astype(np.float32))


def log_softmax(x, axis=-1):
    x = np.asarray(x, dtype=np.float32)
    m = x.max(axis=axis, keepdims=True)
    return _wrap((x - m - np.log(np.exp(x - m).sum(axis=axis, keepdims=True, dtype=np.float32))).astype(np.float32))


_installed = False


def install() -> None:
    global _installed
    if _installed:
        return
    _installed = True
    jnp = _NumpyAsJnp("jax.numpy")
    jnp.__spec__ = importlib.machinery.ModuleSpec("jax.numpy", None)
    sys.modules["jax.numpy"] = jnp
    jax = _module("jax", numpy=jnp, Array=np.ndarray, device_get=lambda x: np.asarray(x), jit=_identity_decorator,
                  process_index=lambda: 0)
    jax.nn = _module("jax.nn", log_softmax=log_softmax)
    jax.lax = _module("jax.lax", stop_gradient=lambda x: x)
    jax.random = _module("jax.random", split=lambda k, n=2: (None,) * n, PRNGKeyArray=Any, PRNGKey=lambda s: None)
    def tree_map(f, tree, *rest):        # dicts / lists / tuples of leaves: all the reference's data code maps over
        if isinstance(tree, dict):
            return {k: tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
        if isinstance(tree, (list, tuple)) and not hasattr(tree, "_fields"):
            return type(tree)(tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
        return f(tree, *rest)

    jax.tree_util = _module("jax.tree_util", tree_map=tree_map)
    jax.experimental = _module("jax.experimental")
    jax.experimental.pjit = _module("jax.experimental.pjit", pjit=_identity_decorator)
    jax.sharding = _module("jax.sharding", NamedSharding=object, PartitionSpec=object, Mesh=object)
    _module("jaxtyping", PyTree=Any)
    flax = _module("flax")
    flax.struct = _module("flax.struct", PyTreeNode=_PyTreeNode, field=_struct_field, dataclass=_identity_decorator)
    flax.training = _module("flax.training")
    flax.training.train_state = _module("flax.training.train_state", TrainState=object)
    flax.linen = _module("flax.linen", Module=object)
    flax.core = _module("flax.core", FrozenDict=dict, freeze=lambda x: x)
    _module("optax", softmax_cross_entropy_with_integer_labels=softmax_cross_entropy_with_integer_labels,
            l2_loss=lambda a, b=None: 0.5 * (a - (0 if b is None else b)) ** 2)
    js = _module("JaxSeq")
    js.utils = _module("JaxSeq.utils", BlockingStrategy=BlockingStrategy, block_sequences=block_sequences, Padding=Padding, Truncation=Truncation,
                       multihost_device_get=lambda x, mesh=None: np.asarray(x), ConvertPath=type("ConvertPath", (), {"convert": lambda self, p: p}),
                       Dataset=object, IterableDataset=object, with_named_sharding_constraint=lambda x, *a, **k: x,
                       match_partition_rules=None)
    js.models = _module("JaxSeq.models")
    js.models.base_interface = _module("JaxSeq.models.base_interface", initialize_attn_mask_pos_ids=lambda ids, pad, m, p: (m, p),
                                       Inference=object, Train=object)
    js.models.gpt2 = _module("JaxSeq.models.gpt2")
    js.models.gpt2.interface = _module("JaxSeq.models.gpt2.interface", GPT2Inference=object)
    import transformers.tokenization_utils as tu
    if not hasattr(tu, "PreTrainedTokenizerBase"):
        tu.PreTrainedTokenizerBase = object
    import transformers.generation as tg
    for name in ("FlaxBeamSearchOutput", "FlaxGreedySearchOutput", "FlaxSampleOutput"):
        if not hasattr(tg, name):
            setattr(tg, name, object)
    _module("transformers.modeling_flax_utils", FlaxPreTrainedModel=object)
    _module("transformers.modeling_flax_outputs", FlaxCausalLMOutput=object, FlaxCausalLMOutputWithCrossAttentions=object)
    ip = sys.modules.get("IPython") or _module("IPython")
    if not hasattr(ip, "embed"):
        ip.embed = lambda *a, **k: None
    if REF not in sys.path:
        sys.path.insert(0, REF)
    # the two score_fn files import these only for type annotations; the real modules pull in the whole JaxSeq training stack
    import LLM_RL.algorithms.ilql  # noqa: F401  (real, empty packages)
    import LLM_RL.algorithms.mc_returns  # noqa: F401
    for name, cls in (("LLM_RL.algorithms.ilql.base_interface", "ILQLInference"), ("LLM_RL.algorithms.mc_returns.base_interface", "MCInference")):
        m = types.ModuleType(name)
        setattr(m, cls, object)
        sys.modules[name] = m
