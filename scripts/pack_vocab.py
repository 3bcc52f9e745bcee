"""Pack word lists (one word per line) into the device table format of lmrl_gym_b200/vocab.py.

    python scripts/pack_vocab.py <dir-with-txt-files> [out_dir]

Run once in the build container over the reference's llm_rl_scripts/wordle/vocab directory.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lmrl_gym_b200.vocab import DATA_DIR, pack_words, read_word_file  # noqa: E402


def main():
    src = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else DATA_DIR
    os.makedirs(out, exist_ok=True)
    for fn in sorted(os.listdir(src)):
        if not fn.endswith(".txt"):
            continue
        words = read_word_file(os.path.join(src, fn))
        pack_words(words).astype("<u4").tofile(os.path.join(out, fn[:-4] + ".u32"))
        print(fn, len(words))


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] in ("-h", "--help"):
        print(__doc__)
        sys.exit(0 if len(sys.argv) > 1 else 2)
    main()
