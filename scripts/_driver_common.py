"""Shared by the task drivers in this directory: tokenizer choice, engine sizing, output files."""
import json
import os
import pickle
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def load_tokenizer(tokenizer_path):
    """GPT-2 BPE from ``<dir>/vocab.json`` + ``<dir>/merges.txt`` plus the ``<|pad|>`` token the reference drivers add
    (eval_bc_gpt2.py:52-53); without a path, the synthetic byte-level table of the same shape (the image ships no tokenizer files)."""
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    if tokenizer_path is None:
        print("no --tokenizer-path: using the synthetic GPT-2-shaped tokenizer table", file=sys.stderr)
        return ByteTokenizer.synthetic()
    return ByteTokenizer.from_gpt2_files(os.path.join(tokenizer_path, "vocab.json"), os.path.join(tokenizer_path, "merges.txt"), add_pad=True)


def dump_outputs(outputs_path, raw, summary, name="interactions"):
    if outputs_path is None:
        return
    os.makedirs(outputs_path, exist_ok=True)
    with open(os.path.join(outputs_path, name + ".pkl"), "wb") as f:
        pickle.dump(raw, f)
    with open(os.path.join(outputs_path, name + "_summary.json"), "w") as f:
        json.dump(to_floats(summary), f)


def to_floats(tree):
    if isinstance(tree, dict):
        return {k: to_floats(v) for k, v in tree.items()}
    if isinstance(tree, (bool, str)) or tree is None:
        return tree
    return float(tree)
