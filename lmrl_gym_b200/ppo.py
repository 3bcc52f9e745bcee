"""PPO data path on the device engine — SURVEY.md §8(f) rank 2 (the step either side of the rollout).

Host mirror of the reference's PPO inference surface with the arithmetic on the GPU:

    TokenHistory / TokenTrajectory / TokenTrajectoryChain        LLM_RL/environment.py:298-419
    CombinedTokenTrajectoryChain                                 LLM_RL/algorithms/ppo/base_interface.py:295-341
    PPOData                                                      LLM_RL/algorithms/ppo/data.py:9-16
    PPOInference.forward / get_ppo_data_from_token_trajectory_chain / get_ppo_data_from_text_trajectory_chain / eval_loss
                                                                 ppo/base_interface.py:413-440, 464-716
    AdaptiveKLController / FixedKLController                     ppo/base_interface.py:35-68

What runs where: token log-probabilities of the current policy (tower 0) and of the frozen initial policy (tower 1) and the
LinearHead values of every position come from ONE teacher-forced pass per tower (``lmrl_forward_host``; the [rows, vocab] logits are
never written); KL-shaped rewards, the GAE scan over action tokens, whitening and the clipped loss are device kernels
(``lmrl_ppo_advantages_host`` / ``lmrl_ppo_loss_host``, csrc/ppo_kernels.inl).  The host only slices chains into PPOData records.

Not built: the gradient step (PPOTrain.step: backward pass, AdamW, NCCL gradient all-reduce).
"""
from __future__ import annotations

import enum
from dataclasses import dataclass
from typing import Callable, Dict, List, NamedTuple, Optional, Sequence, Tuple, Union

import numpy as np

from ._lib import check, ptr
from .engine import Engine
from .environment import TextHistory, TextTrajectory, TextTrajectoryChain
from .gpt2 import GPT2Config, _upload_params
from .value_inference import forward_tokens, upload_value_heads

try:  # pragma: no cover - only when the reference is installed next to this package
    from LLM_RL.environment import TokenHistory, TokenTrajectory, TokenTrajectoryChain  # type: ignore
except Exception:

    @dataclass(frozen=True)
    class TokenHistory:
        tokens: np.ndarray
        is_action: np.ndarray

        def __post_init__(self):
            assert self.tokens.ndim == 1 and self.is_action.ndim == 1, "(tokens, is_action) must be 1 dimensional"
            assert self.tokens.shape == self.is_action.shape, "(tokens, is_action) must have the same shape"

        @classmethod
        def from_text_history(cls, text_history: TextHistory, tokenizer, token_process: Optional[Callable[[List[int]], List[int]]] = None):
            ids: List[int] = []
            act: List[bool] = []
            for item in text_history:
                new = tokenizer.encode(item.text)
                new = token_process(new) if token_process is not None else new
                ids += list(new)
                act += [item.is_action] * len(new)
            return cls(np.array(ids, dtype=np.int32), np.array(act, dtype=np.bool_))

    @dataclass(frozen=True)
    class TokenTrajectory:
        tokens: np.ndarray
        is_action: np.ndarray
        reward: np.ndarray
        done: np.ndarray

        def __post_init__(self):
            assert self.tokens.ndim == 1 and self.is_action.ndim == 1 and self.reward.ndim == 1, "tokens / is_action / reward must be 1 dimensional"
            assert np.ndim(self.done) == 0, "done must be scalar"
            assert self.is_action.shape == self.tokens.shape == self.reward.shape, "is_action and reward must have the same shape as tokens"
            assert not np.any(((1 - self.is_action.astype(np.float32)) * self.reward) != 0.0), "reward must be 0.0 if not an action"

        @classmethod
        def from_text_trajectory(cls, text_trajectory: TextTrajectory, tokenizer, token_process: Optional[Callable[[List[int]], List[int]]] = None):
            ids: List[int] = []
            act: List[bool] = []
            rew: List[float] = []
            for item, r in zip(text_trajectory.text_history, text_trajectory.reward):
                new = tokenizer.encode(item.text)
                new = token_process(new) if token_process is not None else new
                ids += list(new)
                act += [item.is_action] * len(new)
                rew += [0.0] * (len(new) - 1) + [r]            # the item's reward sits on its LAST token (environment.py:369-370)
            return cls(np.array(ids, dtype=np.int32), np.array(act, dtype=np.bool_), np.array(rew, dtype=np.float32),
                       np.array(text_trajectory.done, dtype=np.bool_))

    @dataclass(frozen=True)
    class TokenTrajectoryChain:
        token_trajectory: TokenTrajectory
        next: Optional["TokenTrajectoryChain"]

        def __post_init__(self):
            dones = [bool(t.done) for t in self.to_list()]
            assert not any(dones[:-1]), "token trajectory chain can only be done at the end"

        def to_list(self) -> List[TokenTrajectory]:
            out, cur = [], self
            while cur is not None:
                out.append(cur.token_trajectory)
                cur = cur.next
            return out

        @classmethod
        def from_text_trajectory_chain(cls, text_trajectory_chain: TextTrajectoryChain, tokenizer, token_process=None):
            items, cur = [], text_trajectory_chain
            while cur is not None:
                items.append(TokenTrajectory.from_text_trajectory(cur.text_trajectory, tokenizer, token_process=token_process))
                cur = cur.next
            chain = None
            for t in reversed(items):
                chain = cls(t, chain)
            return chain


class Padding(enum.Enum):
    LEFT = "left"
    RIGHT = "right"


class Truncation(enum.Enum):
    LEFT = "left"
    RIGHT = "right"


class BlockingStrategy(NamedTuple):
    """JaxSeq.utils.BlockingStrategy as the reference uses it (e.g. train_ppo_gpt2.py:356: ``BlockingStrategy(Padding.RIGHT,
    Truncation.RIGHT, max_length)``); JaxSeq is not vendored, the three fields and the two enums are its published surface."""
    padding: Padding
    truncation: Truncation
    max_length: Optional[int]


def block_sequences(sequences: Sequence[Sequence], pad_value, dtype, blocking_strategy: BlockingStrategy) -> np.ndarray:
    """``[n, max_length]`` (the longest sequence when max_length is None): each sequence truncated, then padded, on the given sides."""
    T = blocking_strategy.max_length if blocking_strategy.max_length is not None else max((len(s) for s in sequences), default=0)
    out = np.full((len(sequences), T), pad_value, dtype=dtype)
    for i, s in enumerate(sequences):
        s = np.asarray(s)
        if len(s) > T:
            s = s[len(s) - T:] if blocking_strategy.truncation == Truncation.LEFT else s[:T]
        if blocking_strategy.padding == Padding.LEFT:
            out[i, T - len(s):] = s
        else:
            out[i, :len(s)] = s
    return out


class PPOData(NamedTuple):
    input_ids: np.ndarray            # [t]
    should_take_action: np.ndarray   # [t-1]
    old_logprobs: np.ndarray         # [t-1]
    old_values: np.ndarray           # [t-1]
    old_advantages: np.ndarray       # [t-1]
    old_returns: np.ndarray          # [t-1]

    @staticmethod
    def block(data: List["PPOData"], blocking_strategy: BlockingStrategy, tokenizer) -> Dict[str, np.ndarray]:
        """ppo/data.py:17-60: ids blocked to max_length with the pad id, the five per-output arrays to max_length - 1 with False / 0."""
        shorter = blocking_strategy._replace(max_length=None if blocking_strategy.max_length is None else blocking_strategy.max_length - 1)
        out = dict(input_ids=block_sequences([d.input_ids for d in data], tokenizer.pad_token_id, np.int32, blocking_strategy),
                   should_take_action=block_sequences([d.should_take_action for d in data], False, np.bool_, shorter))
        for k in ("old_logprobs", "old_values", "old_advantages", "old_returns"):
            out[k] = block_sequences([getattr(d, k) for d in data], 0.0, np.float32, shorter)
        return out


_PPO_FIELDS = (("input_ids", np.int32), ("should_take_action", np.bool_), ("old_logprobs", np.float32), ("old_values", np.float32),
               ("old_advantages", np.float32), ("old_returns", np.float32))


class PPODataset:
    """ppo/data.py:62-113: the blocked arrays of one collection round, indexable by minibatch (``dataset[idxs]`` -> dict of arrays — host
    numpy here, where the reference converts to jnp; ``B200PPOTrain.step`` takes that dict)."""

    def __init__(self, input_ids, should_take_action, old_logprobs, old_values, old_advantages, old_returns):
        arrays = dict(input_ids=input_ids, should_take_action=should_take_action, old_logprobs=old_logprobs, old_values=old_values,
                      old_advantages=old_advantages, old_returns=old_returns)
        for k, a in arrays.items():
            assert a.shape[0] == input_ids.shape[0] and a.shape[1] == input_ids.shape[1] - (k != "input_ids"), k
        self.__dict__.update(arrays)

    def __getitem__(self, index) -> Dict[str, np.ndarray]:
        return {k: np.asarray(getattr(self, k)[index], dtype=dt) for k, dt in _PPO_FIELDS}

    def __len__(self) -> int:
        return self.input_ids.shape[0]

    @classmethod
    def from_ppo_data_list(cls, ppo_data_list: List[PPOData], tokenizer, blocking_strategy: BlockingStrategy) -> "PPODataset":
        return cls(**PPOData.block(ppo_data_list, blocking_strategy, tokenizer))


class PPOIterableDataset:
    """ppo/data.py:130-158: a stream of single blocked items."""

    def __init__(self, ppo_data):
        self.ppo_data = ppo_data

    def __iter__(self):
        for item in self.ppo_data:
            yield {k: np.asarray(item[k], dtype=dt) for k, dt in _PPO_FIELDS}

    @classmethod
    def from_ppo_data_iterable(cls, ppo_data, tokenizer, blocking_strategy: BlockingStrategy) -> "PPOIterableDataset":
        class _Blocked:
            def __iter__(self):
                for item in ppo_data:
                    yield {k: v[0] for k, v in PPOData.block([item], blocking_strategy, tokenizer).items()}
        return cls(_Blocked())


class CombinedTokenTrajectoryChain(NamedTuple):
    input_tokens: np.ndarray
    output_tokens: np.ndarray
    rewards: np.ndarray
    should_take_action: np.ndarray
    done: Union[bool, np.ndarray]
    chunk_lens: List[int]

    @classmethod
    def from_token_trajectory_chain(cls, token_trajectory_chain: TokenTrajectoryChain, max_length: Optional[int] = None):
        tts = token_trajectory_chain.to_list()
        assert len(tts) > 0, "token_trajectory_chain must have at least one token_trajectory"
        if max_length is None:
            max_length = max(tt.tokens.shape[0] for tt in tts) + 1
        assert not any(bool(tt.done) for tt in tts[:-1]), "done can only be true at the end of the chain"
        for i, tt in enumerate(tts):
            # an action may not be cut off by truncation, nor may a truncated-to-state chunk be followed by a chunk that opens with an action
            fits = (tt.tokens.shape[0] - 1) <= max_length
            tail_is_state = not np.any(tt.is_action[1:][max_length:])
            next_opens_with_action = i + 1 < len(tts) and bool(tts[i + 1].is_action[0])
            assert not (tail_is_state and next_opens_with_action), "trajectory truncation error"
            assert fits or tail_is_state, "trajectory truncation error"
        cat = lambda f: np.concatenate([f(tt)[:max_length] for tt in tts], axis=0)  # noqa: E731
        return cls(input_tokens=cat(lambda tt: tt.tokens[:-1]), output_tokens=cat(lambda tt: tt.tokens[1:]), rewards=cat(lambda tt: tt.reward[1:]),
                   should_take_action=cat(lambda tt: tt.is_action[1:]), done=tts[-1].done,
                   chunk_lens=[min(tt.tokens.shape[0] - 1, max_length) for tt in tts])

    def unroll_arr(self, arr: np.ndarray) -> List[np.ndarray]:
        assert arr.shape[0] == self.input_tokens.shape[0]
        return np.split(arr, np.cumsum(self.chunk_lens)[:-1], axis=0)


class AdaptiveKLController:
    """ppo/base_interface.py:35-56 (Ziegler et al.): value *= 1 + clip(kl / target - 1, +-0.2) * n_steps / horizon."""

    def __init__(self, init_kl_coef: float, target: float, horizon: int):
        self.value, self.target, self.horizon = init_kl_coef, target, horizon

    def update(self, current: float, n_steps: int):
        self.value *= 1 + float(np.clip(current / self.target - 1, -0.2, 0.2)) * n_steps / self.horizon


class FixedKLController:
    def __init__(self, kl_coef):
        self.value = kl_coef

    def update(self, current: float, n_steps: int):
        pass


# ---- device kernels ---------------------------------------------------------------------------------------------------

def ppo_advantages(engine: Engine, chain_off: np.ndarray, policy_logprobs: np.ndarray, initial_logprobs: np.ndarray, values: np.ndarray,
                   rewards: np.ndarray, should_take_action: np.ndarray, *, gamma: float, lam: float, kl_weight: float, whiten: bool):
    """Flattened chains -> (advantages [N], returns [N], kls [N]); see ``lmrl_ppo_advantages_host`` in include/lmrl_b200.h."""
    chain_off = np.ascontiguousarray(chain_off, dtype=np.int64)
    n_chains, N = chain_off.shape[0] - 1, int(chain_off[-1])
    f = lambda a, n: np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1)[:n])  # noqa: E731
    plp, ilp, val, rew = f(policy_logprobs, N), f(initial_logprobs, N), f(values, N + n_chains), f(rewards, N)
    assert plp.shape[0] == ilp.shape[0] == rew.shape[0] == N and val.shape[0] == N + n_chains
    sta = np.ascontiguousarray(np.asarray(should_take_action).reshape(-1)[:N].astype(np.uint8))
    adv, ret, kls = (np.zeros(max(N, 1), np.float32) for _ in range(3))
    check(engine.lib.lmrl_ppo_advantages_host(engine.handle, n_chains, ptr(chain_off), ptr(plp), ptr(ilp), ptr(val), ptr(rew), ptr(sta),
                                              float(gamma), float(lam), float(kl_weight), int(bool(whiten)), ptr(adv), ptr(ret), ptr(kls)))
    return adv[:N], ret[:N], kls[:N]


_LOSS_KEYS = ("total_loss", "policy_loss", "value_loss", "values_error", "value_clipfrac", "approx_kl", "policy_clipfrac", "ratio",
              "padding_percentage", "values_mean", "old_values_mean", "returns_mean", "n")


def ppo_loss(engine: Engine, attention_mask, logprobs, values, should_take_action, old_logprobs, old_values, old_advantages, old_returns, *,
             cliprange_value: float, cliprange: float, value_loss_coef: float) -> Tuple[float, Dict[str, float]]:
    """``ppo_loss_fn`` (ppo/base_interface.py:71-142) on the device; all arrays [batch, time-1].  Returns (loss, logs)."""
    mask = np.ascontiguousarray(np.asarray(should_take_action).astype(np.float32) * np.asarray(attention_mask, dtype=np.float32))
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (logprobs, values, old_logprobs, old_values, old_advantages, old_returns)]
    assert all(a.shape == mask.shape for a in arrs)
    out = np.zeros(16, np.float64)
    check(engine.lib.lmrl_ppo_loss_host(engine.handle, int(mask.size), ptr(mask), *[ptr(a) for a in arrs], float(cliprange_value), float(cliprange),
                                        float(value_loss_coef), ptr(out)))
    logs = {k: float(out[i]) for i, k in enumerate(_LOSS_KEYS)}
    return logs["total_loss"], logs


# ---- PPOInference -----------------------------------------------------------------------------------------------------

class PPOForwardOutput(NamedTuple):
    """Per-token results of ``PPOInference.forward`` after ``token_logprobs_from_logits`` (logits are never materialised)."""
    initial_policy_logprobs: Optional[np.ndarray]   # [b, t-1]
    policy_logprobs: np.ndarray                     # [b, t-1]
    values: np.ndarray                              # [b, t]


def _block_right(seqs: Sequence[np.ndarray], pad, max_length: Optional[int], dtype) -> np.ndarray:
    T = max_length if max_length is not None else max(len(s) for s in seqs)
    out = np.full((len(seqs), T), pad, dtype=dtype)
    for i, s in enumerate(seqs):
        s = np.asarray(s)[:T]
        out[i, :len(s)] = s
    return out


class B200PPOInference:
    """``PPOInference`` (ppo/base_interface.py:343-716) on one engine: tower 0 = the policy being trained (shared with the rollout
    policy bound to the same engine), tower 1 = the frozen initial policy, head set 0 = the LinearHead value head."""

    def __init__(self, engine: Engine, policy_params: Optional[Dict], value_head_params: Dict, initial_policy_params: Optional[Dict] = None,
                 config: Optional[GPT2Config] = None, pad_token_id: Optional[int] = None):
        self._lmrl_engine = engine
        self.config = config or GPT2Config.small()
        self.tokenizer = engine.tokenizer
        self.pad_token_id = int(pad_token_id if pad_token_id is not None else self.config.n_vocab)
        if policy_params is not None:
            _upload_params(engine, self.config, policy_params, tower=0)
        self.has_initial_policy = initial_policy_params is not None
        if self.has_initial_policy:
            _upload_params(engine, self.config, initial_policy_params, tower=1, role="frozen initial policy (PPO KL term)")
        upload_value_heads(engine, 0, v_linear=value_head_params, role="PPO value head")

    def set_params(self, policy_params: Dict, value_head_params: Optional[Dict] = None) -> None:
        _upload_params(self._lmrl_engine, self.config, policy_params, tower=0)
        if value_head_params is not None:
            upload_value_heads(self._lmrl_engine, 0, v_linear=value_head_params, role="PPO value head")

    def forward(self, input_ids: np.ndarray) -> PPOForwardOutput:
        """RIGHT-padded ``[b, t]`` ids -> per-token log-probs of both policies and the values of every position (0 on padding)."""
        input_ids = np.asarray(input_ids)
        real = input_ids != self.pad_token_id
        lens = real.sum(axis=1)
        seqs = [input_ids[i, :lens[i]] for i in range(input_ids.shape[0])]
        assert all(bool(real[i, :lens[i]].all()) for i in range(input_ids.shape[0])), "only RIGHT padding is supported"
        t = input_ids.shape[1]
        pitch = max(8, t)
        pol = forward_tokens(self._lmrl_engine, 0, 0, seqs, logprobs=True, v=True, pitch=pitch)
        ini = forward_tokens(self._lmrl_engine, 1, 0, seqs, logprobs=True, pitch=pitch)["logprobs"][:, :t - 1] if self.has_initial_policy else None
        return PPOForwardOutput(ini, pol["logprobs"][:, :t - 1], pol["v"][:, :t])

    @staticmethod
    def token_logprobs_from_logits(logits: np.ndarray, input_ids: np.ndarray) -> np.ndarray:
        """ppo/base_interface.py:397-403 for callers that hold logits ``[b, t, V]`` (the engine itself never materialises them)."""
        x = np.asarray(logits, dtype=np.float32)[:, :-1]
        m = x.max(axis=-1, keepdims=True)
        lse = np.log(np.exp(x - m).sum(axis=-1)) + m[..., 0]
        return (np.take_along_axis(x, np.asarray(input_ids)[:, 1:, None], axis=-1)[..., 0] - lse).astype(np.float32)

    def forward_from_str(self, input_strs: List[str], blocking_strategy: BlockingStrategy = BlockingStrategy(Padding.RIGHT, Truncation.RIGHT, None),
                         token_process: Optional[Callable[[List[int]], List[int]]] = None) -> PPOForwardOutput:
        """ppo/base_interface.py:437-462: encode every string on its own, block, forward.  The device forward takes RIGHT-padded rows."""
        assert blocking_strategy.padding == Padding.RIGHT, "the device forward takes RIGHT-padded rows (sequences start at position 0)"
        token_process = token_process or (lambda x: x)
        tokens = [token_process(list(self.tokenizer.encode(item))) for item in input_strs]
        return self.forward(block_sequences(tokens, self.pad_token_id, np.int32, blocking_strategy))

    def get_ppo_data_from_token_trajectory_chain(self, token_trajectory_chains: List[TokenTrajectoryChain], bsize: int, max_length: Optional[int] = None,
                                                 train: bool = False, prng_key=None, verbose: bool = True, *, gamma: float, lam: float, kl_weight: float,
                                                 use_advantage_whitening: bool = True, use_new_advantage_whitening: bool = False,
                                                 _forward: Optional[Callable] = None, _advantages: Optional[Callable] = None):
        """ppo/base_interface.py:464-669.  ``bsize`` only bounds the host-side batches handed to ``forward`` (the engine chunks by its own
        capacity); dropout (``train`` / ``prng_key``) is not modelled — the reference's drivers collect PPO data with train=False."""
        assert self.has_initial_policy or _forward is not None, "an initial policy is needed for the KL term"
        assert not use_new_advantage_whitening, "use_new_advantage_whitening re-whitens a list polluted by a stray loop (:617-633); not built"
        forward = _forward or self.forward
        combined = [CombinedTokenTrajectoryChain.from_token_trajectory_chain(c, None if max_length is None else max_length - 1)
                    for c in token_trajectory_chains]
        lists = [c.to_list() for c in token_trajectory_chains]
        tokens = _block_right([tt.tokens for l in lists for tt in l], self.pad_token_id, max_length, np.int32)
        ilp, plp, val = [], [], []
        for i in range(0, len(tokens), bsize):
            o = forward(tokens[i:i + bsize])
            ilp.append(o[0]); plp.append(o[1]); val.append(o[2])
        ilp, plp, val = (np.concatenate(x, axis=0) for x in (ilp, plp, val))
        n_real = (tokens != self.pad_token_id).sum(axis=1)
        # flatten: per chain the per-output-token arrays of its chunks back to back, values with one bootstrap slot per chain
        n_chains = len(combined)
        chain_off = np.zeros(n_chains + 1, np.int64)
        for c, cc in enumerate(combined):
            chain_off[c + 1] = chain_off[c] + cc.input_tokens.shape[0]
        N = int(chain_off[-1])
        f_plp, f_ilp, f_rew, f_val = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros(N + n_chains, np.float32)
        f_sta = np.zeros(N, np.uint8)
        row = 0
        for c, (cc, l) in enumerate(zip(combined, lists)):
            o = int(chain_off[c])
            for k, n_out in enumerate(cc.chunk_lens):
                r = row + k
                assert n_out == n_real[r] - 1
                f_plp[o:o + n_out] = plp[r, :n_out]
                f_ilp[o:o + n_out] = ilp[r, :n_out]
                f_val[o + c:o + c + n_out] = val[r, :n_out]
                o += n_out
            last = val[row + len(l) - 1, n_real[row + len(l) - 1] - 1]
            f_val[int(chain_off[c + 1]) + c] = last * np.float32(1.0 - float(cc.done))
            f_rew[int(chain_off[c]):int(chain_off[c + 1])] = cc.rewards
            f_sta[int(chain_off[c]):int(chain_off[c + 1])] = cc.should_take_action
            row += len(l)
        adv_fn = _advantages or (lambda *a, **k: ppo_advantages(self._lmrl_engine, *a, **k))
        adv, ret, kls = adv_fn(chain_off, f_plp, f_ilp, f_val, f_rew, f_sta, gamma=gamma, lam=lam, kl_weight=kl_weight, whiten=use_advantage_whitening)
        all_kls = kls[f_sta.astype(bool)]
        datas: List[PPOData] = []
        for c, (cc, l) in enumerate(zip(combined, lists)):
            o0, o1 = int(chain_off[c]), int(chain_off[c + 1])
            parts = [cc.unroll_arr(x) for x in (cc.should_take_action, f_plp[o0:o1], f_val[o0 + c:o1 + c], adv[o0:o1], ret[o0:o1])]
            for k, tt in enumerate(l):
                datas.append(PPOData(tt.tokens[:max_length], parts[0][k], parts[1][k], parts[2][k], parts[3][k], parts[4][k]))
        return datas, all_kls

    def get_ppo_data_from_text_trajectory_chain(self, text_trajectory_chains: List[TextTrajectoryChain], bsize: int, max_length: Optional[int] = None,
                                                train: bool = False, prng_key=None, token_process=None, verbose: bool = True, **kw):
        """ppo/base_interface.py:671-716: tokenise every item on its own, then the token path."""
        chains = [TokenTrajectoryChain.from_text_trajectory_chain(c, self.tokenizer, token_process=token_process) for c in text_trajectory_chains]
        return self.get_ppo_data_from_token_trajectory_chain(chains, bsize=bsize, max_length=max_length, train=train, prng_key=prng_key, verbose=verbose, **kw)

    get_ppo_data_from_text_trajectory_chain_iterble = get_ppo_data_from_text_trajectory_chain      # the reference's second name for it (:710-745)

    def eval_loss(self, input_ids, should_take_action, old_logprobs, old_values, old_advantages, old_returns, *, cliprange_value: float = 0.2,
                  cliprange: float = 0.2, value_loss_coef: float = 1.0, bc_data_input_ids=None, bc_data_input_training_mask=None,
                  bc_loss_weight: float = 0.0) -> Tuple[float, Dict[str, float]]:
        """``PPOInference.eval_loss`` (:747-820; ppo/gpt2/interface.py:381-470): forward the current policy, then ``ppo_loss_fn`` with
        ``values = forward.values[:, :-1]`` and ``attention_mask = (input_ids != pad)[:, 1:]``.  With a BC batch the masked next-token
        cross-entropy is added as ``loss + bc_loss * bc_loss_weight`` and the logs become ``{'ppo': ..., 'bc': {'loss': ...}, 'total_loss': ...}``
        (:451-464)."""
        input_ids = np.asarray(input_ids)
        o = self.forward(input_ids)
        mask = (input_ids != self.pad_token_id)[:, 1:].astype(np.float32)
        loss, logs = ppo_loss(self._lmrl_engine, mask, o.policy_logprobs, o.values[:, :-1], should_take_action, old_logprobs, old_values, old_advantages,
                              old_returns, cliprange_value=cliprange_value, cliprange=cliprange, value_loss_coef=value_loss_coef)
        if bc_data_input_ids is None:
            return loss, logs
        bc_ids = np.asarray(bc_data_input_ids)
        real = bc_ids != self.pad_token_id
        lens = real.sum(axis=1)
        lp = forward_tokens(self._lmrl_engine, 0, 0, [bc_ids[i, :lens[i]] for i in range(bc_ids.shape[0])], logprobs=True,
                            pitch=max(8, bc_ids.shape[1]))["logprobs"][:, :bc_ids.shape[1] - 1]
        m = np.asarray(bc_data_input_training_mask, dtype=np.float64)[:, 1:] * real[:, 1:]
        bc = float(-(lp.astype(np.float64) * m).sum() / m.sum())
        total = float(loss) + bc * float(bc_loss_weight)
        return total, dict(ppo=logs, bc=dict(loss=bc), total_loss=total)
