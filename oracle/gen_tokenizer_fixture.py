"""Write tests/golden/gpt2_format_tokenizer/{vocab.json, merges.txt}: a SMALL byte-level BPE vocabulary in GPT-2's file format.

No GPT-2 vocabulary exists offline, so `ByteTokenizer.from_gpt2_files` — true BPE merges instead of the synthetic
tokenizer's greedy longest match — had no fixture.  This file is this repository's own construction (not GPT-2's): the 256 byte
symbols in GPT-2's `bytes_to_unicode` order, then merges that reproduce the three properties of the real vocabulary the rollout
path depends on:
  * ' a' .. ' z' ('Ġa' ...) are single tokens, bare letters are single tokens: a spaced Wordle line is one token per letter;
  * 'Ċ' + 'Ċ' merges to 'ĊĊ' (id 628 in GPT-2): a prompt ENDING in two newlines ends in one token, two newlines inside a prompt
    stay two tokens (pre-tokeniser `\\s+(?!\\S)`), which is where joined-history and per-item tokenisation differ
    (ppo/gpt2/interface.py:519-522 vs environment.py:313-322);
  * multi-step merges ('W o' -> 'Wo', 'Wo r' -> 'Wor', ... 'Wordle') exercise merge ranks.
Run:  python -m oracle.gen_tokenizer_fixture      (deterministic; the output is committed)
"""
from __future__ import annotations

import json
import os

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gpt2_format_tokenizer")


def main():
    import sys
    sys.path.insert(0, os.path.dirname(OUT.rsplit("tests", 1)[0]))
    from lmrl_gym_b200.tokenizer import gpt2_byte_order, gpt2_unicode_to_byte
    b2u = {b: u for u, b in gpt2_unicode_to_byte().items()}
    sym = lambda bs: "".join(b2u[b] for b in bs)  # noqa: E731
    vocab = {sym(bytes([b])): i for i, b in enumerate(gpt2_byte_order())}
    merges = []

    def merge(a: bytes, b: bytes):
        merges.append((sym(a), sym(b)))
        vocab.setdefault(sym(a + b), len(vocab))

    for c in range(97, 123):
        merge(b" ", bytes([c]))                      # ' a' .. ' z'
    merge(b"\n", b"\n")                              # the trailing-newline-pair token
    for word in (b"Wordle", b"move", b" left", b" right", b" up", b" down"):
        cur = word[:1] if not word.startswith(b" ") else word[:2]
        start = len(cur)
        for i in range(start, len(word)):
            merge(cur, word[i:i + 1])
            cur = cur + word[i:i + 1]
    vocab["<|endoftext|>"] = len(vocab)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "vocab.json"), "w", encoding="utf-8") as f:
        json.dump(vocab, f, ensure_ascii=False)
    with open(os.path.join(OUT, "merges.txt"), "w", encoding="utf-8") as f:
        f.write("#version: 0.2 (fixture written by oracle/gen_tokenizer_fixture.py; not GPT-2's merges)\n")
        for a, b in merges:
            f.write(f"{a} {b}\n")
    print(f"{len(vocab)} tokens, {len(merges)} merges -> {OUT}")


if __name__ == "__main__":
    main()
