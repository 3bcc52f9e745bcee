"""Wordle env — CPU restatement of the reference (oracle; test infrastructure only).

Follows llm_rl_scripts/wordle/env/game.py and env.py of the reference:
  * knowledge table 26 letters x 5 positions, tri-state            game.py:17-51
  * per-letter word test                                            game.py:53-65
  * order-preserving vocabulary filter                              game.py:150-156
  * transition against a freshly sampled target                     game.py:82-92, 213-222
  * feedback from the *updated* table, reward, terminal             game.py:273-296
  * policy-facing text (re)formatting                               env.py:7-26
  * env reset/step                                                  env.py:28-55
Kept deliberately in the reference's cost profile (pure-Python loop over the
vocabulary per step) because ``bench.py`` times it as the CPU baseline "port".
Pinned by tests/golden/wordle_*.json (generated from the imported reference).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

from .mt19937 import MT19937

NOT_HERE, POSSIBLE, HERE = 0, 1, 2
ALPHABET = "abcdefghijklmnopqrstuvwxyz"
WORD_LEN = 5
MAX_TRIES = 6


def load_vocab(path: str) -> List[str]:
    """game.py:164-170 — stripped lines of length 5, file order."""
    out = []
    with open(path, "r") as f:
        for line in f:
            w = line.strip()
            if len(w) == WORD_LEN:
                out.append(w)
    return out


def letter_ok(know: Sequence[int], c: str, word: str) -> bool:
    """game.py:53-65."""
    if all(k == POSSIBLE for k in know):
        return True
    if all(k == NOT_HERE for k in know):
        return c not in word
    for i in range(WORD_LEN):
        if know[i] == HERE and word[i] != c:
            return False
        if know[i] == NOT_HERE and word[i] == c:
            return False
    return c in word


def word_ok(table: Sequence[Sequence[int]], word: str) -> bool:
    """game.py:76-80."""
    for li, c in enumerate(ALPHABET):
        if not letter_ok(table[li], c, word):
            return False
    return True


def apply_guess(table: List[List[int]], guess: str, target: str) -> List[List[int]]:
    """game.py:82-92 — sequential per-letter assignment, duplicates included."""
    new = [list(row) for row in table]
    for i, c in enumerate(guess):
        li = ord(c) - 97
        if c == target[i]:
            new[li][i] = HERE
        elif c in target:
            new[li][i] = NOT_HERE
        else:
            new[li] = [NOT_HERE] * WORD_LEN
    return new


def is_wellformed(action: str) -> bool:
    return len(action) == WORD_LEN and all(c in ALPHABET for c in action)


class WordleOracle:
    """One episode of WordleEnvironment (env.py:39-55) + WordleGame (game.py:193-296)."""

    def __init__(self, vocab: List[str], require_words_in_vocab: bool = True, bad_word_reward: float = -1.0):
        self.vocab = vocab
        self.vocab_set = set(vocab)
        self.require = require_words_in_vocab
        self.bad_word_reward = bad_word_reward
        self.reset(None)

    def reset(self, seed: Optional[int]):
        self.rng = MT19937(seed if seed is not None else 0)
        self.table = [[POSSIBLE] * WORD_LEN for _ in range(26)]
        self.filtered = [w for w in self.vocab if word_ok(self.table, w)]
        self.history: List[str] = []
        self.target: Optional[str] = None

    def _invalid(self, a: str) -> bool:
        return (not is_wellformed(a)) or (self.require and a not in self.vocab_set)

    def reward(self):
        """game.py:290-293."""
        if self.history and (len(self.history[-1]) != WORD_LEN or self.history[-1] not in self.vocab_set):
            return self.bad_word_reward
        return int(len(self.filtered) == 1 and self.filtered[0] in self.history) - 1

    def terminal(self) -> bool:
        """game.py:295-296."""
        return len(self.history) == MAX_TRIES or self.reward() == 0

    def feedback(self) -> str:
        """game.py:273-288, last action only (env.py:49)."""
        a = self.history[-1]
        if self._invalid(a):
            return ""
        s = ""
        for i, c in enumerate(a):
            know = self.table[ord(c) - 97]
            if know[i] == HERE:
                s += "<g>"
            elif all(k == NOT_HERE for k in know):
                s += "<b>"
            elif know[i] == NOT_HERE:
                s += "<y>"
        return s

    def step(self, action: str) -> Tuple[str, float, bool]:
        """game.py:213-222 then env.py:48-50.  Returns (feedback, reward, done)."""
        if self._invalid(action):
            self.history = self.history + [action]
        else:
            assert not self.terminal()
            self.target = self.filtered[self.rng.choice_index(len(self.filtered))]
            self.table = apply_guess(self.table, action, self.target)
            self.filtered = [w for w in self.vocab if word_ok(self.table, w)]
            self.history = self.history + [action]
        return self.feedback(), self.reward(), self.terminal()


# --- policy-facing formatting (env.py:7-26) ------------------------------------

HEADER = "Wordle:\n"


def normalise_action(text: str) -> str:
    """env.py:23."""
    return text.strip().replace(" ", "")


def render_action(norm: str) -> str:
    """env.py:11."""
    return " ".join(list(norm)) + "\n"


def render_feedback(fb: str) -> str:
    """env.py:13-16."""
    if len(fb) == 0:
        return "\n"
    return " ".join(fb[1:-1].split("><")) + "\n"


class ReformatWordleOracle:
    """ReformatWordleEnvironment(WordleEnvironment(vocab)) on (text, is_action) tuples."""

    def __init__(self, vocab, require_words_in_vocab=True, bad_word_reward=-1.0):
        self.inner = WordleOracle(vocab, require_words_in_vocab, bad_word_reward)

    def reset(self, seed=None, options=None):
        self.inner.reset(seed)
        return ((HEADER, False),)

    def step(self, history):
        assert history[-1][1], "last item must be an action"
        # deformat every item after the header, exactly like env.py:19-26
        inner_hist = []
        for text, is_action in history[1:]:
            if is_action:
                inner_hist.append((normalise_action(text), True))
            else:
                inner_hist.append(("<" + "><".join(list(text.strip().replace(" ", ""))) + ">", False))
        fb, r, done = self.inner.step(inner_hist[-1][0])
        inner_hist.append((fb, False))
        out = [(HEADER, False)]
        for text, is_action in inner_hist:
            out.append((render_action(text), True) if is_action else (render_feedback(text), False))
        return tuple(out), r, done
