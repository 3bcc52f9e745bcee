"""Host-side mirror of the reference's rollout surface (LLM_RL/environment.py).

Same names, argument meaning and results as the reference so that a task driver only swaps the env / policy
objects it constructs:

    Text, TextHistory, text_history_to_str, StepResult            environment.py:12-19
    TextTrajectory, TextTrajectoryChain                           environment.py:23-37
    TextEnv, BatchedTextEnv (+ the two adapters)                  environment.py:41-111
    TextPolicy, BatchedTextPolicy (+ the two adapters)            environment.py:115-143
    InteractionTransition, interact_environment, text_env_eval    environment.py:147-267

When the reference package itself is importable (``LLM_RL.environment``), its classes are re-exported so that
``isinstance`` checks inside the reference's own ``interact_environment`` (environment.py:164-167) accept the
device env / policy of this package; otherwise identical classes are defined here.

What is new: ``text_env_eval`` / ``interact_environment`` recognise a (device env, device policy) pair that
live on the same engine and run the whole ``while not all(done)`` loop (environment.py:180-206) on the GPU
through ``lmrl_rollout`` (include/lmrl_b200.h) — no host round-trip per turn or per token.  Any other pair goes
through the generic loop below, which follows the reference step by step (including the ``None`` skip rule of
environment.py:192-197 and the ``npad`` dummy rows of environment.py:182, 225-233).

Deliberate differences, all on paths no reference script exercises:
  * ``TextPolicyToBatchedTextPolicy.act`` does not print its inputs (the reference has three debug prints,
    environment.py:130-132).
  * ``initial_text_history`` works (the reference raises TypeError on it, environment.py:172).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from copy import deepcopy
from dataclasses import dataclass
from typing import Any, Callable, Dict, Iterator, List, NamedTuple, Optional, Sequence, Tuple, Union

import numpy as np

try:  # pragma: no cover - only when the reference is installed next to this package
    from LLM_RL.environment import (  # type: ignore
        Text, TextTrajectory, TextTrajectoryChain, TextEnv, BatchedTextEnv, TextPolicy, BatchedTextPolicy,
        InteractionTransition,
    )
    _USING_REFERENCE_TYPES = True
except Exception:
    _USING_REFERENCE_TYPES = False

    @dataclass(frozen=True)
    class Text:
        text: str
        is_action: bool

    @dataclass(frozen=True)
    class TextTrajectory:
        text_history: Tuple[Text, ...]
        reward: Tuple[float, ...]
        done: bool

        def __post_init__(self):
            assert len(self.reward) == len(self.text_history), "reward is needed for each text"
            assert all(r == 0.0 for r, t in zip(self.reward, self.text_history) if not t.is_action), \
                "reward for non-actions texts should be 0.0"

    @dataclass(frozen=True)
    class TextTrajectoryChain:
        text_trajectory: TextTrajectory
        next: Optional["TextTrajectoryChain"]

    class TextEnv(ABC):
        @abstractmethod
        def step(self, text_history):
            ...

        @abstractmethod
        def reset(self, seed: Optional[int] = None, options: Optional[Dict] = None):
            ...

        def close(self) -> None:
            pass

        def copy(self):
            return deepcopy(self)

    class BatchedTextEnv(ABC):
        @abstractmethod
        def step(self, text_history, done=None):
            ...

        @abstractmethod
        def reset(self, seed=None, options=None):
            ...

        def close(self) -> None:
            pass

        def copy(self):
            return deepcopy(self)

    class TextPolicy(ABC):
        @abstractmethod
        def act(self, text_history):
            ...

    class BatchedTextPolicy(ABC):
        @abstractmethod
        def act(self, text_history, done=None):
            ...

    class InteractionTransition(NamedTuple):
        pre_action_history: Tuple[Text, ...]
        post_action_history: Tuple[Text, ...]
        post_transition_history: Tuple[Text, ...]
        reward: float
        done: bool


TextHistory = Tuple[Text, ...]
StepResult = Tuple[TextHistory, float, bool]


def text_history_to_str(text_history: TextHistory) -> str:
    return "".join(t.text for t in text_history)


# --- single <-> batched adapters (environment.py:71-111, 125-143) ---------------------------------------

class TextEnvToBatchedTextEnv(BatchedTextEnv):
    def __init__(self, env: TextEnv):
        self.env = env
        self.batch_env_copies: Optional[List[TextEnv]] = None

    def step(self, text_history, done=None):
        assert self.batch_env_copies is not None, "reset must be called before step"
        assert len(text_history) == len(self.batch_env_copies), \
            "batch size must be the same as the number of environments initalized"
        if done is None:
            done = [False] * len(text_history)
        assert len(text_history) == len(done)
        return [None if d else e.step(h) for e, h, d in zip(self.batch_env_copies, text_history, done)]

    def reset(self, seed=None, options=None):
        if seed is None and options is None:
            seed, options = [None], [None]
        elif seed is None:
            seed = [None] * len(options)
        elif options is None:
            options = [None] * len(seed)
        assert len(seed) == len(options)
        self.batch_env_copies = [self.env.copy() for _ in seed]
        return [e.reset(seed=s, options=o) for e, s, o in zip(self.batch_env_copies, seed, options)]

    def close(self) -> None:
        for e in self.batch_env_copies or []:
            e.close()
        self.env.close()


class BatchedTextEnvToTextEnv(TextEnv):
    def __init__(self, env: BatchedTextEnv):
        self.env = env

    def step(self, text_history):
        return self.env.step([text_history])[0]

    def reset(self, seed=None, options=None):
        return self.env.reset(seed=[seed], options=[options])[0]

    def close(self) -> None:
        self.env.close()


class TextPolicyToBatchedTextPolicy(BatchedTextPolicy):
    def __init__(self, policy: TextPolicy):
        self.policy = policy

    def act(self, text_history, done=None):
        if done is None:
            done = [False] * len(text_history)
        assert len(text_history) == len(done)
        return [None if d else self.policy.act(h) for h, d in zip(text_history, done)]


class BatchedTextPolicyToTextPolicy(TextPolicy):
    def __init__(self, policy: BatchedTextPolicy):
        self.policy = policy

    def act(self, text_history):
        return self.policy.act([text_history])[0]


class UserPolicy(TextPolicy):
    """A human at the keyboard as the policy (LLM_RL/environment.py:271-288): prints the history, reads one line, returns the history plus
    ``initial_str + line`` as the action.  The reference appends with ``text_history+[Text(...)]`` — a TypeError on the tuple histories the
    loop passes; here the action is appended as a tuple."""

    def __init__(self, initial_str: str, postproc_print_f: Optional[Callable[[str], str]] = None, postproc_action_f: Optional[Callable[[str], str]] = None):
        self.initial_str = initial_str
        self.postproc_print_f = postproc_print_f if postproc_print_f is not None else (lambda x: x)
        self.postproc_action_f = postproc_action_f if postproc_action_f is not None else (lambda x: x)

    def act(self, text_history):
        print("=" * 25)
        print(self.postproc_print_f(text_history_to_str(text_history)))
        print("=" * 25)
        response = self.initial_str + input(self.initial_str)
        return tuple(text_history) + (Text(self.postproc_action_f(response), True),)


# --- the rollout loop -------------------------------------------------------------------------------------

def _fused_pair(env, policy):
    """(engine) if env and policy are device objects bound to the same engine, else None."""
    eng_e = getattr(env, "_lmrl_engine", None)
    eng_p = getattr(policy, "_lmrl_engine", None)
    if eng_e is not None and eng_e is eng_p and getattr(env, "_lmrl_fusable", False) and getattr(policy, "fusable", False):
        return eng_e
    return None


def interact_environment(
    env: Union[TextEnv, BatchedTextEnv],
    policy: Union[TextPolicy, BatchedTextPolicy],
    initial_text_history: Optional[Union[TextHistory, List[TextHistory]]] = None,
    env_seed: Union[Optional[int], Optional[List[Optional[int]]]] = None,
    env_options: Union[Optional[Dict], Optional[List[Optional[Dict]]]] = None,
    bsize: int = 1,
    npad: int = 0,
) -> List[List[InteractionTransition]]:
    """One batch of episodes run to completion (environment.py:154-207)."""
    assert bsize > 0
    if env_seed is not None and isinstance(env_seed, int):
        env_seed = [env_seed] * bsize
    if env_options is not None and isinstance(env_options, dict):
        env_options = [env_options] * bsize

    engine = _fused_pair(env, policy)
    if engine is not None and initial_text_history is None:
        from .rollout import fused_interact
        return fused_interact(engine, env, policy, env_seed, env_options, bsize)

    if isinstance(env, TextEnv):
        env = TextEnvToBatchedTextEnv(env)
    if isinstance(policy, TextPolicy):
        policy = TextPolicyToBatchedTextPolicy(policy)
    if initial_text_history is not None and isinstance(initial_text_history, tuple):
        initial_text_history = [initial_text_history] * bsize
    text_history = initial_text_history
    if text_history is None:
        text_history = env.reset(env_seed, env_options)  # positional, as the reference calls it (:176)

    transitions: List[List[InteractionTransition]] = [[] for _ in range(bsize)]
    done = [False] * bsize
    pad_rows = [(Text("", False),)] * npad
    while not all(done):
        pre = text_history
        acted = policy.act(list(text_history) + pad_rows, done=list(done) + [True] * npad)[:bsize]
        results = env.step(acted, done=done)
        results = [(None, None, True) if r is None else r for r in results]
        text_history = [r[0] for r in results]
        reward = [r[1] for r in results]
        done = [r[2] for r in results]
        for i in range(bsize):
            if done[i] and (pre[i] is None or acted[i] is None or text_history[i] is None or reward[i] is None):
                continue  # an episode that finished earlier (environment.py:192-197)
            transitions[i].append(InteractionTransition(pre[i], acted[i], text_history[i], reward[i], done[i]))
    return transitions


def summarize_episodes(rewards: Sequence[float], dones: Sequence[float], lengths: Sequence[int]) -> Dict[str, Any]:
    """environment.py:244-265 — reward/done in float32, lengths as given."""
    rewards = np.asarray(rewards, dtype=np.float32)
    dones = np.asarray(dones, dtype=np.float32)

    def stats(x):
        return dict(mean=np.mean(x), std=np.std(x), min=np.min(x), max=np.max(x))

    return dict(reward=stats(rewards), done=stats(dones), length=stats(lengths))


def text_env_eval(
    env: Union[TextEnv, BatchedTextEnv],
    policy: Union[TextPolicy, BatchedTextPolicy],
    n_rollouts: int,
    initial_text_history: Optional[TextHistory] = None,
    seed_generator: Optional[Iterator[int]] = None,
    env_options: Optional[Dict] = None,
    interaction_callback: Optional[Callable[[List[InteractionTransition]], None]] = None,
    bsize: int = 1,
    verbose: bool = True,
) -> Tuple[List[List[InteractionTransition]], Dict[str, Any]]:
    """n_rollouts episodes in ceil(n/bsize) batches + the summary dict (environment.py:211-267)."""
    engine = _fused_pair(env, policy)
    if engine is not None and initial_text_history is None:
        from .rollout import fused_text_env_eval
        return fused_text_env_eval(engine, env, policy, n_rollouts, seed_generator, env_options,
                                   interaction_callback, bsize)

    interactions: List[List[InteractionTransition]] = []
    rewards, dones, lengths = [], [], []
    n_batches = (n_rollouts + bsize - 1) // bsize
    it = range(n_batches)
    if verbose:
        try:
            from tqdm.auto import tqdm
            it = tqdm(it)
        except Exception:
            pass
    for _ in it:
        actual = min(n_rollouts - len(interactions), bsize)
        seeds = [None] * actual if seed_generator is None else [next(seed_generator) for _ in range(actual)]
        batch = interact_environment(env, policy, initial_text_history=initial_text_history, env_seed=seeds,
                                     env_options=[env_options] * actual, bsize=actual, npad=bsize - actual)
        for episode in batch:
            interactions.append(episode)
            rewards.append(sum(t.reward for t in episode))
            dones.append(episode[-1].done)
            lengths.append(len(episode))
            if interaction_callback is not None:
                interaction_callback(episode)
    return interactions, summarize_episodes(rewards, dones, lengths)
