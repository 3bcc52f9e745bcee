"""Fused rollouts: ``interact_environment`` / ``text_env_eval`` for a (device env, device policy) pair.

The reference loop (LLM_RL/environment.py:176-206) alternates ``policy.act`` and ``env.step`` on the host with two
host<->device hops per turn.  Here one call to ``lmrl_rollout_host`` (include/lmrl_b200.h) runs every turn of every
episode of the batch on the GPU and returns compact logs; the ``InteractionTransition`` lists the reference returns
are rebuilt from those logs *lazily* (an episode's Text tuples are only materialised when indexed), because at
4096 concurrent episodes the ~10^6 Python objects of an eager rebuild would cost more than the rollout itself.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check, ptr
from .engine import ENV_MAZE, ENV_WORDLE
from .environment import InteractionTransition, Text, summarize_episodes
from .gpt2 import default_out_str_process
from .wordle import HEADER, normalise_action, render_action, seed_to_u64

_FB = "gyb"


class RolloutLogs:
    """Host copies of the per-episode / per-turn arrays of one fused batch."""

    def __init__(self, n: int, max_turns: int, max_new: int):
        self.n, self.T, self.K = n, max_turns, max_new
        self.ep_return = np.zeros(n, np.float32)
        self.ep_length = np.zeros(n, np.int32)
        self.ep_done = np.zeros(n, np.uint8)
        self.step_reward = np.zeros((n, max_turns), np.float32)
        self.step_valid = np.zeros((n, max_turns), np.uint8)
        self.step_feedback = np.zeros((n, max_turns, 5), np.uint8)
        self.step_aux = np.zeros((n, max_turns), np.int32)
        self.act_tokens = np.zeros((n, max_turns, max_new), np.int32)
        self.act_len = np.zeros((n, max_turns), np.int32)

    def as_struct(self) -> "_lib.lmrl_rollout_out":
        return _lib.lmrl_rollout_out(ptr(self.ep_return), ptr(self.ep_length), ptr(self.ep_done), ptr(self.step_reward),
                                     ptr(self.step_valid), ptr(self.step_feedback), ptr(self.step_aux), ptr(self.act_tokens),
                                     ptr(self.act_len))

    @property
    def h2d_bytes(self) -> int:
        return 8 * self.n  # seeds

    @property
    def d2h_bytes(self) -> int:
        return sum(a.nbytes for a in (self.ep_return, self.ep_length, self.ep_done, self.step_reward, self.step_valid,
                                      self.step_feedback, self.step_aux, self.act_tokens, self.act_len))


def run_fused_batch(engine, env, policy, seeds: Sequence[Optional[int]], options=None) -> RolloutLogs:
    n = len(seeds)
    if n > engine.max_envs:
        raise _lib.LmrlError(f"batch of {n} episodes exceeds the engine capacity {engine.max_envs}")
    logs = RolloutLogs(n, engine.max_turns, engine.max_new_tokens)
    seeds_u64 = np.asarray([seed_to_u64(s) for s in seeds], dtype=np.uint64)
    out = logs.as_struct()
    s = policy.sampling()
    opt_pos = env.fused_options(options) if hasattr(env, "fused_options") else None
    check(engine.lib.lmrl_rollout_host(engine.handle, env._lmrl_kind, ptr(seeds_u64), ptr(opt_pos), n, C.byref(s), C.byref(out)))
    return logs


class WordleEpisode(Sequence):
    """One episode's ``List[InteractionTransition]``, rebuilt on demand from the device logs."""

    def __init__(self, logs: RolloutLogs, i: int, tokenizer, env, out_str_process):
        self._logs, self._i, self._tok, self._env = logs, i, tokenizer, env
        self._out = out_str_process or default_out_str_process
        self._len = int(logs.ep_length[i])
        self._cache: Optional[List[InteractionTransition]] = None

    def __len__(self):
        return self._len

    def action_text(self, t: int) -> str:
        L, i = self._logs, self._i
        return self._out(self._tok.decode(L.act_tokens[i, t, :L.act_len[i, t]].tolist()))

    def reward(self, t: int):
        L, i, env = self._logs, self._i, self._env
        if L.step_valid[i, t] and (env.require_words_in_vocab or env._in_vocab(normalise_action(self.action_text(t)))):
            return int(L.step_reward[i, t])              # game.py:293 — Python int for a legal in-vocab guess
        return env.bad_word_reward                       # game.py:292 — the configured float

    def _build(self) -> List[InteractionTransition]:
        L, i = self._logs, self._i
        hist: Tuple[Text, ...] = (Text(HEADER, False),)
        out = []
        for t in range(self._len):
            action = self.action_text(t)
            post_action = hist + (Text(action, True),)
            fb = L.step_feedback[i, t]
            fb_text = (" ".join(_FB[c] for c in fb if c < 3) + "\n") if L.step_valid[i, t] else "\n"
            nxt = hist + (Text(render_action(normalise_action(action)), True), Text(fb_text, False))
            done = bool(L.ep_done[i]) if t == self._len - 1 else False
            out.append(InteractionTransition(hist, post_action, nxt, self.reward(t), done))
            hist = nxt
        return out

    def __getitem__(self, k):
        if self._cache is None:
            self._cache = self._build()
        return self._cache[k]

    def __iter__(self):
        if self._cache is None:
            self._cache = self._build()
        return iter(self._cache)


class MazeEpisode(Sequence):
    """One Maze episode's ``List[InteractionTransition]`` rebuilt from the device logs (maze/env/env.py:161-184):
    ``step_aux = outcome | cell_after << 8 | cell_before << 20``, ``step_valid`` = the action string was a known move."""

    def __init__(self, logs: RolloutLogs, i: int, tokenizer, env, out_str_process):
        self._logs, self._i, self._tok, self._env = logs, i, tokenizer, env
        self._out = out_str_process or default_out_str_process
        self._len = int(logs.ep_length[i])
        self._cache: Optional[List[InteractionTransition]] = None

    def __len__(self):
        return self._len

    def action_text(self, t: int) -> str:
        L, i = self._logs, self._i
        return self._out(self._tok.decode(L.act_tokens[i, t, :L.act_len[i, t]].tolist()))

    def _build(self) -> List[InteractionTransition]:
        L, i, env = self._logs, self._i, self._env
        out = []
        hist: Tuple[Text, ...] = ()
        for t in range(self._len):
            aux = int(L.step_aux[i, t])
            outcome, cell, cell0 = aux & 0xFF, (aux >> 8) & 0xFFF, (aux >> 20) & 0xFFF
            if t == 0:
                hist = (Text(env.describe_cell(cell0), False),)
            post_action = hist + (Text(self.action_text(t), True),)
            if outcome == 2:
                nxt: Tuple[Text, ...] = (Text("Failure\n", False),)
            elif outcome == 1:
                nxt = (Text("Success\n", False),)
            else:
                obs = Text(env.describe_cell(cell), False)
                if not L.step_valid[i, t]:
                    nxt = (obs,)
                else:
                    full = post_action + (obs,)
                    nxt = full[max(0, len(full) - env.last_k):]
            done = bool(L.ep_done[i]) if t == self._len - 1 else False
            out.append(InteractionTransition(hist, post_action, nxt, float(L.step_reward[i, t]), done))
            hist = nxt
        return out

    def __getitem__(self, k):
        if self._cache is None:
            self._cache = self._build()
        return self._cache[k]

    def __iter__(self):
        if self._cache is None:
            self._cache = self._build()
        return iter(self._cache)


def _episode_cls(env):
    return MazeEpisode if env._lmrl_kind == ENV_MAZE else WordleEpisode


def episode_returns(logs: RolloutLogs, env) -> np.ndarray:
    """Per-episode sum of rewards in step order, in float64 like the reference's Python ``sum`` (environment.py:238)."""
    n, T = logs.n, logs.T
    r = logs.step_reward.astype(np.float64)
    if env._lmrl_kind == ENV_MAZE:
        pass                                   # Maze rewards are floats as produced (env.py:109-131)
    elif env.require_words_in_vocab:
        r = np.where(logs.step_valid.astype(bool), r, env.bad_word_reward)
    else:
        # a legal guess outside the vocabulary earns bad_word_reward; the device already wrote that value
        r = np.where(r == np.float32(env.bad_word_reward), env.bad_word_reward, r)
    acc = np.zeros(n, np.float64)
    for t in range(T):
        acc = acc + np.where(t < logs.ep_length, r[:, t], 0.0)
    return acc


def fused_interact(engine, env, policy, env_seed, env_options, bsize: int):
    if not getattr(policy, "fusable", False):
        raise _lib.LmrlError("policy has custom in/out string processing; use the generic loop")
    seeds = [None] * bsize if env_seed is None else list(env_seed)
    assert len(seeds) == bsize
    logs = run_fused_batch(engine, env, policy, seeds, env_options)
    cls = _episode_cls(env)
    return [cls(logs, i, engine.tokenizer, env, policy.out_str_process) for i in range(bsize)]


def fused_text_env_eval(engine, env, policy, n_rollouts: int, seed_generator: Optional[Iterator[int]], env_options: Optional[Dict],
                        interaction_callback: Optional[Callable], bsize: int):
    interactions: List[Any] = []
    rewards: List[np.ndarray] = []
    dones: List[np.ndarray] = []
    lengths: List[np.ndarray] = []
    while len(interactions) < n_rollouts:
        actual = min(n_rollouts - len(interactions), bsize)
        seeds = [None] * actual if seed_generator is None else [next(seed_generator) for _ in range(actual)]
        logs = run_fused_batch(engine, env, policy, seeds, None if env_options is None else [env_options] * actual)
        cls = _episode_cls(env)
        eps = [cls(logs, i, engine.tokenizer, env, policy.out_str_process) for i in range(actual)]
        interactions.extend(eps)
        rewards.append(episode_returns(logs, env))
        dones.append(logs.ep_done.astype(bool))
        lengths.append(logs.ep_length.astype(np.int64))
        if interaction_callback is not None:
            for ep in eps:
                interaction_callback(ep)
    summary = summarize_episodes(np.concatenate(rewards), np.concatenate(dones), [int(x) for x in np.concatenate(lengths)])
    return interactions, summary
