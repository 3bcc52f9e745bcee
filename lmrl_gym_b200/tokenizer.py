"""Byte-level tokenizer tables for the rollout engine.

The reference tokenises on the host with the HF GPT-2 byte-level BPE tokenizer every turn
(``AutoTokenizer.from_pretrained('gpt2')`` + an added ``<|pad|>``, llm_rl_scripts/wordle/bc/eval_bc_gpt2.py:51-52;
generation EOS = ``tokenizer.encode('\\n')[0]``, wordle/ppo/train_ppo_gpt2.py:223).  The engine needs two things
from a tokenizer, both as flat tables it can keep in HBM:

  * id -> bytes for *every* id the policy can sample (to turn generated tokens into the action text the env
    validates, replacing ``tokenizer.batch_decode`` of value_rl_base/base_interface.py:126), and
  * the canonical ids of the closed set of strings the envs emit (header, letters, feedback, maze sentences).

No GPT-2 vocabulary files exist offline in this image, so ``ByteTokenizer.synthetic()`` builds a deterministic
stand-in with GPT-2's shape: 50257 ids, ids 0..255 = single bytes in GPT-2's ``bytes_to_unicode`` order (so
``'\\n'`` is id 198 and ``' '`` is id 220 exactly as in GPT-2), ``<|endoftext|>`` = 50256, ``<|pad|>`` = 50257,
GPT-2's pre-tokeniser regex, and greedy longest-match inside a pre-token in place of learned merges.
``ByteTokenizer.from_gpt2_files(vocab.json, merges.txt)`` loads the real thing (true BPE merges) when the files
are available; the engine only ever sees the tables.
"""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

try:
    import regex as _re
    _PRETOK = _re.compile(r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""")
except Exception:  # pragma: no cover
    import re as _re
    _PRETOK = _re.compile(r"""'s|'t|'re|'ve|'m|'ll|'d| ?[^\W\d_]+| ?\d+| ?[^\s\w]+|\s+(?!\S)|\s+""")

N_GPT2 = 50257
ENDOFTEXT_ID = 50256
PAD_ID = 50257


def gpt2_byte_order() -> List[int]:
    """Byte value of GPT-2 ids 0..255 (the published ``bytes_to_unicode`` ordering)."""
    printable = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    rest = [b for b in range(256) if b not in printable]
    return printable + rest


def gpt2_unicode_to_byte() -> Dict[str, int]:
    order = gpt2_byte_order()
    out, n = {}, 0
    for i, b in enumerate(order):
        if i < 188:
            out[chr(b)] = b
        else:
            out[chr(256 + n)] = b
            n += 1
    return out


ENV_WORDS = [
    "Wordle", "The", " goal", " is", " at", " position", " Your", " current", " There", " are", " walls", " wall",
    " to", " your", " left", " right", " above", " below", " you", " no", " near", " a", "move", " up", " down",
    "Success", "Failure", " move",
]


def _synthetic_token_bytes() -> List[bytes]:
    toks: List[bytes] = [bytes([b]) for b in gpt2_byte_order()]
    seen = set(toks)

    def add(t: bytes):
        if t not in seen and len(toks) < ENDOFTEXT_ID:
            seen.add(t)
            toks.append(t)

    for c in range(33, 127):
        add(b" " + bytes([c]))
    for w in ENV_WORDS:
        add(w.encode())
    letters = [bytes([c]) for c in range(97, 123)]
    for a in letters:
        for b in letters:
            add(a + b)
            add(b" " + a + b)
    for a in letters:
        for b in letters:
            for c in letters:
                add(a + b + c)
                add(b" " + a + b + c)
    # fill the remaining ids with 4..6-letter strings from a fixed LCG walk
    x = 12345
    while len(toks) < ENDOFTEXT_ID:
        x = (x * 1103515245 + 12345) & 0x7FFFFFFF
        n = 4 + (x >> 8) % 3
        y, s = x, b""
        for _ in range(n):
            y = (y * 1103515245 + 12345) & 0x7FFFFFFF
            s += letters[(y >> 16) % 26]
        add(s if (x >> 4) & 1 else b" " + s)
    toks.append(b"")  # <|endoftext|>  (special: decodes to nothing)
    toks.append(b"")  # <|pad|>
    return toks


class ByteTokenizer:
    def __init__(self, token_bytes: Sequence[bytes], merges: Optional[List[Tuple[bytes, bytes]]] = None,
                 special_ids: Iterable[int] = (ENDOFTEXT_ID, PAD_ID), name: str = "custom"):
        self.token_bytes = list(token_bytes)
        self.name = name
        self.special_ids = set(special_ids)
        self.vocab: Dict[bytes, int] = {}
        for i, t in enumerate(self.token_bytes):
            if i not in self.special_ids and t not in self.vocab:
                self.vocab[t] = i
        self.max_len = max(len(t) for t in self.token_bytes)
        self.ranks = None if merges is None else {m: r for r, m in enumerate(merges)}
        self.pad_token_id = PAD_ID if len(self.token_bytes) > PAD_ID else None
        self.eos_token_id = ENDOFTEXT_ID
        self._cache: Dict[str, List[int]] = {}
        for b in range(256):
            assert bytes([b]) in self.vocab, "byte-level vocabulary must contain every single byte"

    # ---- construction ----
    @classmethod
    def synthetic(cls) -> "ByteTokenizer":
        return cls(_synthetic_token_bytes(), None, name="synthetic-gpt2-shape-v1")

    @classmethod
    def from_gpt2_files(cls, vocab_json: str, merges_txt: str, add_pad: bool = True) -> "ByteTokenizer":
        u2b = gpt2_unicode_to_byte()
        vocab = json.load(open(vocab_json, "r", encoding="utf-8"))
        toks: List[bytes] = [b""] * len(vocab)
        specials = set()
        for s, i in vocab.items():
            if s == "<|endoftext|>":
                specials.add(i)
                continue
            toks[i] = bytes(u2b[ch] for ch in s)
        merges = []
        for line in open(merges_txt, "r", encoding="utf-8").read().split("\n")[1:]:
            if not line.strip():
                continue
            a, b = line.split()
            merges.append((bytes(u2b[ch] for ch in a), bytes(u2b[ch] for ch in b)))
        eot = vocab.get("<|endoftext|>")
        if add_pad:
            toks.append(b"")
            specials.add(len(toks) - 1)
        tok = cls(toks, merges, specials, name="gpt2-bpe")
        if eot is not None:                                   # vocabularies of any size: ids come from the file, not from GPT-2's constants
            tok.eos_token_id = int(eot)
        tok.pad_token_id = len(toks) - 1 if add_pad else None
        return tok

    def __len__(self):
        return len(self.token_bytes)

    # ---- encode / decode ----
    def _encode_pretoken(self, piece: bytes) -> List[int]:
        if self.ranks is None:  # greedy longest match
            out, i = [], 0
            while i < len(piece):
                for n in range(min(self.max_len, len(piece) - i), 0, -1):
                    t = self.vocab.get(piece[i:i + n])
                    if t is not None:
                        out.append(t)
                        i += n
                        break
            return out
        parts = [bytes([b]) for b in piece]
        while len(parts) > 1:
            best, best_rank = None, None
            for k in range(len(parts) - 1):
                r = self.ranks.get((parts[k], parts[k + 1]))
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = k, r
            if best is None:
                break
            a, b = parts[best], parts[best + 1]
            merged, k = [], 0
            while k < len(parts):
                if k < len(parts) - 1 and parts[k] == a and parts[k + 1] == b:
                    merged.append(a + b)
                    k += 2
                else:
                    merged.append(parts[k])
                    k += 1
            parts = merged
        return [self.vocab[p] for p in parts]

    def encode(self, text: str) -> List[int]:
        hit = self._cache.get(text)
        if hit is not None:
            return list(hit)
        out: List[int] = []
        for piece in _PRETOK.findall(text):
            out.extend(self._encode_pretoken(piece.encode("utf-8")))
        if len(text) <= 256 and len(self._cache) < 200000:
            self._cache[text] = list(out)
        return out

    def decode_bytes(self, ids: Iterable[int], skip_special_tokens: bool = True) -> bytes:
        tb = self.token_bytes
        return b"".join(tb[i] for i in ids if not (skip_special_tokens and i in self.special_ids))

    def decode(self, ids: Iterable[int], skip_special_tokens: bool = True) -> str:
        return self.decode_bytes(ids, skip_special_tokens).decode("utf-8", errors="replace")

    # ---- device tables ----
    def newline_id(self) -> int:
        return self.encode("\n")[0]

    def device_tables(self) -> Dict[str, np.ndarray]:
        """Flat arrays for ``lmrl_load_tokenizer`` (include/lmrl_b200.h)."""
        lens = np.asarray([0 if i in self.special_ids else len(t) for i, t in enumerate(self.token_bytes)], dtype=np.int32)
        offsets = np.zeros(len(lens) + 1, dtype=np.int32)
        np.cumsum(lens, out=offsets[1:])
        flat = np.frombuffer(b"".join(b"" if i in self.special_ids else t for i, t in enumerate(self.token_bytes)),
                             dtype=np.uint8).copy()
        byte_tok = np.asarray([self.vocab[bytes([b])] for b in range(256)], dtype=np.int32)
        tok_first = np.full(128, -1, dtype=np.int32)
        tok_space = np.full(128, -1, dtype=np.int32)
        for c in range(128):
            ch = chr(c)
            if ch.isspace():
                continue  # whitespace never survives normalisation at a line start / after a space as one token
            e = self.encode(ch)
            if len(e) == 1:
                tok_first[c] = e[0]
            e = self.encode("x " + ch)[1:]  # ' c' in mid-line position
            if len(e) == 1:
                tok_space[c] = e[0]
        return dict(token_bytes=flat, offsets=offsets, byte_tok=byte_tok, tok_first=tok_first, tok_space=tok_space)
