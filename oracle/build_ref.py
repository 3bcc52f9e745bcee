"""Recipe for oracle/_ref: the UNMODIFIED reference rollout loop + Wordle / Maze envs, staged for the CPU reference arm.

    python -m oracle.build_ref            # needs /root/reference (build container); writes only into oracle/_ref/

The reference is Python, so "building" it is staging the few source files of the path where a Python
process on the GPU box can import them: oracle/_ref/ is git-ignored (never committed — the repository holds no reference
source) but NOT gpurun-ignored, so it travels to the GPU box like a compiled .so would.  Files are copied byte for byte and
their sha256 is recorded in oracle/_ref/MANIFEST.json; `bench.py --impl reference` and the `cpu_baseline` leg import
`LLM_RL.environment.text_env_eval` / `TextEnvToBatchedTextEnv` and `llm_rl_scripts.wordle.env.env.*` from there
(`cpu_baseline.kind` = "reference") and fall back to the oracle port (`kind` = "port") when the directory is absent.

What is NOT staged: the reference's policy stack (JAX 0.4.7 / flax / optax / JaxSeq: absent from the image and the wheelhouse,
SURVEY.md §8c) — the policy of the reference arm is the torch-CPU stand-in `oracle.cpu_policy.TorchCPUGPT2Policy`, which
subclasses the reference's own `BatchedTextPolicy` once `LLM_RL` is importable.  One stub module (`termcolor`, imported by
game.py for its colour printing) is written next to the staged files; it is this repository's code, not the reference's.
Test infrastructure only: nothing under lmrl_gym_b200/ imports this.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LMRL_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")

# the path's files + the package markers they need
FILES = [
    "LLM_RL/environment.py",
    "llm_rl_scripts/wordle/env/env.py",
    "llm_rl_scripts/wordle/env/game.py",
    "llm_rl_scripts/wordle/env/utils.py",
    "llm_rl_scripts/wordle/env/scripted_policies.py",
    "llm_rl_scripts/wordle/vocab/wordle_official.txt",
    "llm_rl_scripts/maze/env/env.py",
    "llm_rl_scripts/maze/env/mazes.py",
    "llm_rl_scripts/maze/env/maze_utils.py",
    "llm_rl_scripts/maze/env/randomness.py",
]
PACKAGES = ["LLM_RL", "llm_rl_scripts", "llm_rl_scripts/wordle", "llm_rl_scripts/wordle/env", "llm_rl_scripts/maze", "llm_rl_scripts/maze/env"]

TERMCOLOR_STUB = '"""Stub written by oracle/build_ref.py (not reference code): game.py imports termcolor.colored for printing only."""\n\n\ndef colored(s, *a, **k):\n    return s\n'


def available() -> bool:
    return os.path.exists(os.path.join(DST, "MANIFEST.json"))


def build(force: bool = False) -> bool:
    """Stage the files; returns False (and changes nothing) when /root/reference is absent."""
    if not os.path.isdir(REF):
        return available()
    if available() and not force:
        return True
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for pkg in PACKAGES:
        os.makedirs(os.path.join(DST, pkg), exist_ok=True)
        src_init = os.path.join(REF, pkg, "__init__.py")
        dst_init = os.path.join(DST, pkg, "__init__.py")
        # LLM_RL/__init__.py of the reference is empty; package markers are written empty so nothing else is pulled in
        open(dst_init, "w").close()
        manifest[pkg + "/__init__.py"] = "empty marker (reference file %s)" % ("exists" if os.path.exists(src_init) else "absent")
    for rel in FILES:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    os.makedirs(os.path.join(DST, "_stubs"), exist_ok=True)
    with open(os.path.join(DST, "_stubs", "termcolor.py"), "w") as f:
        f.write(TERMCOLOR_STUB)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(dict(source=REF, files=manifest, note="byte-for-byte copies; sha256 of each source file"), f, indent=1)
    return True


def activate():
    """Put the staged reference (and the termcolor stub, only if termcolor is absent) on sys.path."""
    if not available():
        raise RuntimeError("oracle/_ref is absent: run `python -m oracle.build_ref` in the build container")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    try:
        import termcolor  # noqa: F401
    except Exception:
        stubs = os.path.join(DST, "_stubs")
        if stubs not in sys.path:
            sys.path.append(stubs)


if __name__ == "__main__":
    ok = build(force=True)
    print("oracle/_ref:", "staged" if ok else "reference not found at " + REF)
    if ok:
        print(json.dumps(json.load(open(os.path.join(DST, "MANIFEST.json")))["files"], indent=1))
