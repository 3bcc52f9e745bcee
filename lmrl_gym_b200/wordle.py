"""Device-resident Wordle environment behind the reference's BatchedTextEnv interface.

Drop-in for ``ReformatWordleEnvironment(WordleEnvironment(Vocabulary.from_file(path, fill_cache=False),
require_words_in_vocab, bad_word_reward))`` (llm_rl_scripts/wordle/env/env.py:28-55) wrapped by
``TextEnvToBatchedTextEnv`` (LLM_RL/environment.py:71-98): same reset/step signatures, same texts, rewards
(int for legal guesses, float ``bad_word_reward`` otherwise, game.py:290-293) and done flags, with the game state
(knowledge table, CPython-compatible MT19937, action history) living in HBM and advanced by one warp per env.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import check, ptr
from .engine import Engine, ENV_WORDLE
from .environment import BatchedTextEnv, Text, TextHistory
from .vocab import load_packed

HEADER = "Wordle:\n"
_FB = "gyb"
MAX_ACTION_BYTES = 256


def seed_to_u64(seed: Optional[int]) -> int:
    """random.Random(seed) uses abs(seed) (CPython); None draws OS entropy like the reference (env.py:53)."""
    if seed is None:
        return int.from_bytes(os.urandom(8), "little")
    s = abs(int(seed))
    if s >> 64:
        raise ValueError("device envs take seeds below 2**64 (the reference's seed_generator yields 64-bit seeds)")
    return s


def normalise_action(text: str) -> str:
    return text.strip().replace(" ", "")  # wordle/env/env.py:23


def render_action(norm: str) -> str:
    return " ".join(norm) + "\n"  # env.py:11


def encode_actions(texts: Sequence[Optional[str]], cap: int = MAX_ACTION_BYTES) -> Tuple[np.ndarray, np.ndarray]:
    """Action strings -> [n, pitch] uint8 + lengths.  Normalisation is idempotent, so an over-long action is
    pre-normalised on the host before it is cut to the device row size (it cannot be a legal guess either way)."""
    rows = []
    for t in texts:
        b = b"" if t is None else t.encode("utf-8", "surrogatepass")
        if len(b) > cap:
            b = normalise_action(t).encode("utf-8", "surrogatepass")
            if len(b) > cap:
                b = b[:cap]
                while b and (b[-1] & 0xC0) == 0x80:
                    b = b[:-1]
                if b and b[-1] >= 0xC0:
                    b = b[:-1]
        rows.append(b)
    pitch = max(8, (max(len(b) for b in rows) + 7) // 8 * 8) if rows else 8
    arr = np.zeros((len(rows), pitch), dtype=np.uint8)
    lens = np.zeros(len(rows), dtype=np.int32)
    for i, b in enumerate(rows):
        arr[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
        lens[i] = len(b)
    return arr, lens


class B200WordleEnv(BatchedTextEnv):
    def __init__(self, engine: Engine, vocab: str = "wordle_official.txt", require_words_in_vocab: bool = True,
                 bad_word_reward: float = -1.0):
        self._lmrl_engine = engine
        self._lmrl_fusable = True
        self._lmrl_kind = ENV_WORDLE
        self.vocab_name = vocab
        self.packed = load_packed(vocab) if isinstance(vocab, str) else np.asarray(vocab, dtype=np.uint32)
        self.require_words_in_vocab = bool(require_words_in_vocab)
        self.bad_word_reward = float(bad_word_reward)
        check(engine.lib.lmrl_load_wordle(engine.handle, ptr(self.packed), len(self.packed),
                                          int(self.require_words_in_vocab), self.bad_word_reward))
        engine.env_kind = ENV_WORDLE
        self.n: Optional[int] = None
        self.last_n_filtered: Optional[np.ndarray] = None

    # BatchedTextEnv.reset(seed, options) — called positionally by interact_environment (environment.py:176)
    def reset(self, seed: Optional[List[Optional[int]]] = None, options: Optional[List[Optional[Dict]]] = None) -> List[TextHistory]:
        if seed is None and options is None:
            seed = [None]
        elif seed is None:
            seed = [None] * len(options)
        n = len(seed)
        seeds = np.asarray([seed_to_u64(s) for s in seed], dtype=np.uint64)
        check(self._lmrl_engine.lib.lmrl_wordle_reset_host(self._lmrl_engine.handle, ptr(seeds), n))
        self.n = n
        return [(Text(HEADER, False),) for _ in range(n)]

    def step(self, text_history: List[Optional[TextHistory]], done: Optional[List[bool]] = None):
        assert self.n is not None, "reset must be called before step"
        assert len(text_history) == self.n, "batch size must be the same as the number of environments initalized"
        if done is None:
            done = [False] * self.n
        assert len(done) == self.n
        n = self.n
        actions: List[Optional[str]] = []
        for h, d in zip(text_history, done):
            if d or h is None:
                actions.append(None)
            else:
                assert h[-1].is_action
                actions.append(h[-1].text)
        skip = np.asarray([a is None for a in actions], dtype=np.uint8)
        arr, lens = encode_actions(actions)
        reward = np.zeros(n, np.float32)
        dn = np.zeros(n, np.uint8)
        valid = np.zeros(n, np.uint8)
        fb = np.zeros((n, 5), np.uint8)
        nf = np.zeros(n, np.int32)
        e = self._lmrl_engine
        check(e.lib.lmrl_wordle_step_host(e.handle, ptr(arr), ptr(lens), arr.shape[1], ptr(skip), n, ptr(reward), ptr(dn),
                                          ptr(valid), ptr(fb), ptr(nf)))
        self.last_n_filtered = nf
        out = []
        for i in range(n):
            if actions[i] is None:
                out.append(None)
                continue
            out.append((self._render(text_history[i], valid[i], fb[i]),
                        int(reward[i]) if valid[i] and self._legal_reward_is_int(actions[i]) else float(reward[i]),
                        bool(dn[i])))
        return out

    def _legal_reward_is_int(self, action_text: str) -> bool:
        # game.py:291-293: the int branch needs `len == 5 and in vocab`; with require_words_in_vocab=False a legal
        # guess outside the vocabulary still earns the float bad_word_reward.
        return True if self.require_words_in_vocab else self._in_vocab(normalise_action(action_text))

    def _in_vocab(self, word: str) -> bool:
        if not hasattr(self, "_word_set"):
            from .vocab import unpack_word
            self._word_set = {unpack_word(v) for v in self.packed}
        return word in self._word_set

    @staticmethod
    def _render(history: TextHistory, valid, fb) -> TextHistory:
        """reformat_history(deformat_history(history) + (transition,)) of env.py:7-34, item by item."""
        out = [Text(HEADER, False)]
        for item in history[1:]:
            norm = item.text.strip().replace(" ", "")
            if item.is_action:
                out.append(Text(render_action(norm), True))
            else:
                out.append(Text(" ".join(norm) + "\n", False))
        if valid:
            out.append(Text(" ".join(_FB[c] for c in fb if c < 3) + "\n", False))
        else:
            out.append(Text("\n", False))
        return tuple(out)

    def copy(self):
        return self  # device state is per-slot already; TextEnvToBatchedTextEnv-style deep copies are not needed


def scripted_rollout_records(act_words: np.ndarray, feedback: np.ndarray, step_reward: np.ndarray, ep_len: np.ndarray, ep_done: np.ndarray) -> List[Dict]:
    """Logs of a scripted (``RandomMixturePolicy``-equivalent) device rollout — ``lmrl_wordle_scripted_rollout_dev``: packed action words
    ``[n, T]``, feedback codes ``[n, T, 5]`` (0 g, 1 y, 2 b), per-step rewards ``[n, T]``, episode lengths and done flags — as the records
    the reference's data generator writes (wordle/misc/data_gen.py:52-61 over ``PolicyDataGenerator``, wordle/env/data.py:23-34):
    ``sequence`` = ``[(text, float(is_action)), ...]`` of the reformatted final history (env.py:7-17), ``reward`` = ``[r_0, 0.0, r_1, 0.0, ...]``,
    ``done``.  These are the ``bc_data_path`` items of the PPO / BC drivers (``bc_batch_from_str_segments``)."""
    from .vocab import unpack_word
    out = []
    for i in range(len(ep_len)):
        seq, rew = [(HEADER, 0.0)], []
        for t in range(int(ep_len[i])):
            word = unpack_word(int(act_words[i, t]) & 0xFFFFFFFF)
            seq.append((" ".join(word) + "\n", 1.0))
            seq.append((" ".join("gyb"[int(c)] for c in feedback[i, t]) + "\n", 0.0))
            rew += [float(step_reward[i, t]), 0.0]
        out.append(dict(sequence=seq, reward=rew, done=bool(ep_done[i])))
    return out
