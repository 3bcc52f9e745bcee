"""PPO data path — numpy restatement of the reference (oracle; test infrastructure only).

Follows LLM_RL/algorithms/ppo/base_interface.py:
    ppo_loss_fn                                   :71-142
    get_action_state_next_state_idxs              :230-243
    whiten                                        :245-251
    get_advantages_and_returns                    :253-293
    CombinedTokenTrajectoryChain                  :295-341
    PPOInference.get_ppo_data_from_token_trajectory_chain   :464-669
    AdaptiveKLController / FixedKLController      :35-68
and LLM_RL/utils.py:36-40 (unpad_array), LLM_RL/algorithms/ppo/data.py:9-16 (PPOData).

PINNED: ``tests/golden/ppo_data.json`` is written by ``oracle/gen_golden_algos.py``, which executes the reference's own
functions (imported unmodified from /root/reference through ``oracle/ref_shims.py``: jax.numpy -> numpy) on seeded inputs;
``tests/test_oracle_ppo.py`` checks this restatement against those vectors.

A token trajectory here is a tuple ``(tokens int32[n], is_action bool[n], reward float32[n], done bool)`` and a chain a
list of them (the reference's linked ``TokenTrajectoryChain.to_list()``, environment.py:386-391).  The model enters
through ``forward_fn(tokens[n, T]) -> (initial_logprobs[n, T-1], policy_logprobs[n, T-1], values[n, T])``.
"""
from __future__ import annotations

from typing import Callable, Dict, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

Trajectory = Tuple[np.ndarray, np.ndarray, np.ndarray, bool]


class PPOData(NamedTuple):          # ppo/data.py:9-16
    input_ids: np.ndarray
    should_take_action: np.ndarray
    old_logprobs: np.ndarray
    old_values: np.ndarray
    old_advantages: np.ndarray
    old_returns: np.ndarray


def action_state_next_state_idxs(should_take_action: np.ndarray):
    """:230-243 — action and state indices are the action-token positions; the next state of an action token is the NEXT action-token
    position, and for the last one the extra slot appended after the chain (the bootstrap value)."""
    sta = np.asarray(should_take_action, dtype=bool)
    idx = np.flatnonzero(sta)
    nxt = np.concatenate((idx[1:], [sta.shape[0]])) if idx.size else idx
    return idx, idx.copy(), nxt.astype(idx.dtype)


def whiten(xs: np.ndarray, shift_mean: bool = True) -> np.ndarray:
    """:245-251."""
    xs = np.asarray(xs, dtype=np.float32)
    mean, var = np.mean(xs), np.var(xs)
    out = (xs - mean) * np.reciprocal(np.sqrt(var + np.float32(1e-8)))
    return out if shift_mean else out + mean


def advantages_and_returns(state_values, next_state_values, action_rewards, gamma: float, lam: float, use_whitening: bool = True):
    """:253-293 — GAE backwards over [b, n] arrays; every product with gamma / lam is a float32 product (python scalars are weak)."""
    sv, nv, r = (np.asarray(a, dtype=np.float32) for a in (state_values, next_state_values, action_rewards))
    assert sv.shape == nv.shape == r.shape
    n = sv.shape[1]
    adv = np.zeros_like(sv)
    last = np.zeros(sv.shape[0], np.float32)
    g, gl = np.float32(gamma), np.float32(gamma * lam)
    for t in range(n - 1, -1, -1):
        delta = r[:, t] + g * nv[:, t] - sv[:, t]
        last = delta + gl * last
        adv[:, t] = last
    ret = adv + sv
    if use_whitening:
        adv = whiten(adv)
    return adv, ret


def combine_chain(chain: Sequence[Trajectory], max_length: Optional[int] = None) -> Dict:
    """CombinedTokenTrajectoryChain.from_token_trajectory_chain (:303-334) incl. its truncation asserts."""
    assert len(chain) > 0
    if max_length is None:
        max_length = max(t[0].shape[0] for t in chain) + 1
    assert not any(t[3] for t in chain[:-1]), "done can only be true at the end of the chain"
    for i, (tok, act, _, _) in enumerate(chain):
        no_trunc = (tok.shape[0] - 1) <= max_length
        ends_with_state = not np.any(act[1:][max_length:])
        next_starts_with_action = i < len(chain) - 1 and bool(chain[i + 1][1][0])
        assert not (ends_with_state and next_starts_with_action), "trajectory truncation error"
        assert no_trunc or ends_with_state, "trajectory truncation error"
    return dict(
        input_tokens=np.concatenate([t[0][:-1][:max_length] for t in chain]),
        output_tokens=np.concatenate([t[0][1:][:max_length] for t in chain]),
        rewards=np.concatenate([t[2][1:][:max_length] for t in chain]).astype(np.float32),
        should_take_action=np.concatenate([t[1][1:][:max_length] for t in chain]).astype(bool),
        done=bool(chain[-1][3]),
        chunk_lens=[min(t[0].shape[0] - 1, max_length) for t in chain],
    )


def block_right(seqs: Sequence[np.ndarray], pad, max_length: Optional[int], dtype) -> np.ndarray:
    """JaxSeq block_sequences with Padding.RIGHT / Truncation.RIGHT (the strategy of :499-503)."""
    T = max_length if max_length is not None else max(len(s) for s in seqs)
    out = np.full((len(seqs), T), pad, dtype=dtype)
    for i, s in enumerate(seqs):
        s = np.asarray(s)[:T]
        out[i, :len(s)] = s
    return out


def _n_real(mask_row: np.ndarray) -> int:
    """unpad_array (utils.py:36-40): everything before the first masked-out position."""
    z = np.flatnonzero(~mask_row)
    return int(z[0]) if z.size else mask_row.shape[0]


def ppo_data_from_chains(chains: Sequence[Sequence[Trajectory]], forward_fn: Callable, pad_token_id: int, bsize: int,
                         max_length: Optional[int] = None, *, gamma: float, lam: float, kl_weight: float,
                         use_advantage_whitening: bool = True):
    """:464-669.  Returns (list of PPOData, all_kls)."""
    combined = [combine_chain(c, None if max_length is None else max_length - 1) for c in chains]
    tokens = block_right([t[0] for c in chains for t in c], pad_token_id, max_length, np.int32)
    ilp, plp, val = [], [], []
    for i in range(0, len(tokens), bsize):
        a, b, c = forward_fn(tokens[i:i + bsize])
        ilp.append(np.asarray(a, np.float32)); plp.append(np.asarray(b, np.float32)); val.append(np.asarray(c, np.float32))
    ilp, plp, val = np.concatenate(ilp), np.concatenate(plp), np.concatenate(val)
    mask = tokens != pad_token_id
    row = 0
    values_chains, plp_chains, log_ratio = [], [], []
    for c, cc in zip(chains, combined):
        rows = range(row, row + len(c))
        row += len(c)
        n_real = [_n_real(mask[r]) for r in rows]
        n_out = [_n_real(mask[r, 1:]) for r in rows]
        i_chain = np.concatenate([ilp[r, :n] for r, n in zip(rows, n_out)])
        p_chain = np.concatenate([plp[r, :n] for r, n in zip(rows, n_out)])
        v_chain = np.concatenate([val[r, :n][:-1] for r, n in zip(rows, n_real)])
        last_value = val[rows[-1], :n_real[-1]][-1]
        v_chain = np.concatenate((v_chain, last_value[None] * (1.0 - float(cc["done"])))).astype(np.float32)     # :555-570
        values_chains.append(v_chain)
        plp_chains.append(p_chain)
        log_ratio.append((p_chain - i_chain) * cc["should_take_action"].astype(np.float32))                         # :572-575
    sta_all = np.concatenate([cc["should_take_action"] for cc in combined])
    lr_all = np.concatenate(log_ratio)[np.flatnonzero(sta_all)]
    all_kls = np.exp(lr_all) - 1 - lr_all                                                                             # :577-579
    for cc, lr in zip(combined, log_ratio):
        cc["rewards"] = cc["rewards"] - kl_weight * lr                                                                 # :581-584
    all_adv, all_ret = [], []
    for cc, v_chain in zip(combined, values_chains):
        a_idx, s_idx, n_idx = action_state_next_state_idxs(cc["should_take_action"])
        adv, ret = advantages_and_returns(v_chain[s_idx][None], v_chain[n_idx][None], cc["rewards"][a_idx][None], gamma, lam, use_whitening=False)
        all_adv.append(adv[0]); all_ret.append(ret[0])
    if use_advantage_whitening:                                                                                        # :608-615
        w = whiten(np.concatenate(all_adv), shift_mean=True)
        pos = 0
        for i, a in enumerate(all_adv):
            all_adv[i] = w[pos:pos + a.shape[0]]
            pos += a.shape[0]
    out: List[PPOData] = []
    for c, cc, v_chain, p_chain, adv, ret in zip(chains, combined, values_chains, plp_chains, all_adv, all_ret):
        a_idx = np.flatnonzero(cc["should_take_action"])
        adv_chain = np.zeros(v_chain.shape[0] - 1, np.float32); adv_chain[a_idx] = adv
        ret_chain = np.zeros(v_chain.shape[0] - 1, np.float32); ret_chain[a_idx] = ret
        cuts = np.cumsum(cc["chunk_lens"])[:-1]
        parts = [np.split(x, cuts) for x in (cc["should_take_action"], p_chain, v_chain[:-1], adv_chain, ret_chain)]
        for k, traj in enumerate(c):
            out.append(PPOData(traj[0][:max_length], parts[0][k], parts[1][k], parts[2][k], parts[3][k], parts[4][k]))
    return out, all_kls


def ppo_loss(attention_mask, logprobs, values, should_take_action, old_logprobs, old_values, old_advantages, old_returns, *,
             cliprange_value: float, cliprange: float, value_loss_coef: float) -> Tuple[np.float32, Dict]:
    """ppo_loss_fn (:71-142); every array is [batch, time-1].  Returns (loss, flat logs)."""
    f = lambda a: np.asarray(a, dtype=np.float32)  # noqa: E731
    attention_mask, logprobs, values, old_logprobs, old_values, old_advantages, old_returns = map(
        f, (attention_mask, logprobs, values, old_logprobs, old_values, old_advantages, old_returns))
    mask = np.asarray(should_take_action).astype(np.float32) * attention_mask
    n = mask.sum()
    values_clipped = np.clip(values, old_values - np.float32(cliprange_value), old_values + np.float32(cliprange_value))
    vf1, vf2 = (values - old_returns) ** 2, (values_clipped - old_returns) ** 2
    vf_loss = np.float32(0.5) * np.sum(np.maximum(vf1, vf2) * mask) / n
    vf_clipfrac = np.sum((vf2 > vf1).astype(np.float32) * mask) / n
    log_ratio = (logprobs - old_logprobs) * mask
    ratio = np.exp(log_ratio)
    approx_kl = np.sum((ratio - 1) - log_ratio) / n
    pg1 = -old_advantages * ratio
    pg2 = -old_advantages * np.clip(ratio, np.float32(1.0 - cliprange), np.float32(1.0 + cliprange))
    pg_loss = np.sum(np.maximum(pg1, pg2) * mask) / n
    pg_clipfrac = np.sum((pg2 > pg1).astype(np.float32) * mask) / n
    loss = pg_loss + np.float32(value_loss_coef) * vf_loss
    logs = dict(total_loss=loss, policy_loss=pg_loss, value_loss=vf_loss, values_error=np.sum(((values - old_returns) * mask) ** 2) / n,
                vf_clipfrac=vf_clipfrac, approx_kl=approx_kl, pg_clipfrac=pg_clipfrac, ratio=(ratio * mask).sum() / n,
                padding_percentage=n / mask.size, values_mean=(values * mask).sum() / n, old_values_mean=(old_values * mask).sum() / n,
                returns_mean=(old_returns * mask).sum() / n)
    return loss, {k: float(v) for k, v in logs.items()}


class AdaptiveKLController:
    """:35-56."""

    def __init__(self, init_kl_coef: float, target: float, horizon: int):
        self.value, self.target, self.horizon = init_kl_coef, target, horizon

    def update(self, current: float, n_steps: int):
        err = float(np.clip(current / self.target - 1, -0.2, 0.2))
        self.value *= 1 + err * n_steps / self.horizon


class FixedKLController:
    """:58-68."""

    def __init__(self, kl_coef: float):
        self.value = kl_coef

    def update(self, current: float, n_steps: int):
        pass
