"""Extract the played Text-Nav transcripts the reference ships (build container only) -> tests/golden/textnav_transcripts.json.

    python -m oracle.gen_golden_textnav            # needs /root/reference (read-only)

Source: llm_rl_scripts/text_nav/notebooks/play_game.ipynb — cell outputs of two complete games played through the reference's
TextNavEnv (one won, one lost).  Kept per step: the command typed, the admissible-command list TextWorld printed after it
("Available actions: [...]", or the wrapped "Admissible commands: ..." line of the final screen), and score / done from the closing
banner.  Room prose is dropped (TextWorld grammar text, not reproducible).  Test infrastructure only.
"""
from __future__ import annotations

import ast
import json
import os
import re

REF = os.environ.get("LMRL_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "textnav_transcripts.json")


def admissible_of(block: str):
    m = re.search(r"Available actions: (\[.*?\])", block, re.S)
    if m:
        return ast.literal_eval(m.group(1))
    m = re.search(r"Admissible commands: (.*?)(?:\n\s*\n|\Z)", block, re.S)
    return [c.strip() for c in " ".join(m.group(1).split()).split(", ")] if m else None


def main():
    nb = json.load(open(os.path.join(REF, "llm_rl_scripts", "text_nav", "notebooks", "play_game.ipynb")))
    games = []
    for cell in nb["cells"]:
        text = "".join("".join(o.get("text", [])) for o in cell.get("outputs", []) if "text" in o)
        if "Your objective is to put the stale food" not in text:
            continue
        parts = re.split(r"\n> +([^\n]*)\n", text)
        first = dict(command=None, admissible=admissible_of(parts[0]), done=False, score=0)
        steps, last_cmd = [], None
        for i in range(1, len(parts), 2):
            cmd, block = parts[i].strip(), parts[i + 1]
            adm = admissible_of(block)
            if adm is None:            # the echo line "> cmd" printed before the rendered screen
                continue
            done = "*** The End ***" in block or "*** You lost! ***" in block
            score = 1 if "You scored 1 out of" in block else 0
            steps.append(dict(command=cmd, admissible=adm, done=done, score=score, known="You can't see any such thing" not in block))
        games.append(dict(first=first, steps=steps))
    json.dump(dict(source="llm_rl_scripts/text_nav/notebooks/play_game.ipynb", games=games), open(OUT, "w"), indent=1)
    for g in games:
        print(len(g["steps"]), "steps; first admissible:", g["first"]["admissible"], "; last:", g["steps"][-1])


if __name__ == "__main__":
    main()
