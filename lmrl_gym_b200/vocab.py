"""Wordle vocabularies as device tables.

The reference keeps a vocabulary as a Python list of 5-letter strings read from a text file
(llm_rl_scripts/wordle/env/game.py:164-170: stripped lines of length 5, file order preserved).  The device
form is one uint32 per word, letter i in bits [5i, 5i+5) ('a' = 0), in the same order — order matters because
the target word is the k-th survivor of the filter (game.py:154, 179).  The word lists the reference ships
(llm_rl_scripts/wordle/vocab/*.txt) are stored here already packed as ``data/vocab/<name>.u32``.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "vocab")
WORD_LEN = 5


def pack_word(w: str) -> int:
    assert len(w) == WORD_LEN
    v = 0
    for i, c in enumerate(w):
        o = ord(c) - 97
        if not 0 <= o < 26:
            raise ValueError(f"device vocabularies hold lowercase a-z words only, got {w!r}")
        v |= o << (5 * i)
    return v


def unpack_word(v: int) -> str:
    return "".join(chr(97 + ((int(v) >> (5 * i)) & 31)) for i in range(WORD_LEN))


def pack_words(words: Sequence[str]) -> np.ndarray:
    return np.asarray([pack_word(w) for w in words], dtype=np.uint32)


def read_word_file(path: str) -> List[str]:
    """Same acceptance rule as the reference's Vocabulary.from_file (game.py:164-170)."""
    out = []
    with open(path, "r") as f:
        for line in f:
            w = line.strip()
            if len(w) == WORD_LEN:
                out.append(w)
    return out


def resolve_vocab_path(name_or_path: str) -> str:
    """Accepts a path to a .txt / .u32 file, or the bare name of a shipped list
    ('wordle_official.txt', 'wordle_official', ...)."""
    if os.path.exists(name_or_path):
        return name_or_path
    base = os.path.basename(name_or_path)
    stem = base[:-4] if base.endswith(".txt") or base.endswith(".u32") else base
    cand = os.path.join(DATA_DIR, stem + ".u32")
    if os.path.exists(cand):
        return cand
    raise FileNotFoundError(f"vocabulary {name_or_path!r} not found (shipped lists: {sorted(os.listdir(DATA_DIR))})")


def load_packed(name_or_path: str) -> np.ndarray:
    path = resolve_vocab_path(name_or_path)
    if path.endswith(".u32"):
        return np.fromfile(path, dtype="<u4").astype(np.uint32)
    return pack_words(read_word_file(path))


def load_vocab_words(name_or_path: str) -> List[str]:
    return [unpack_word(v) for v in load_packed(name_or_path)]
