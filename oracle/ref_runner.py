"""CPU reference arm, kind "reference": the UNMODIFIED reference loop + Wordle env from oracle/_ref (staged by oracle/build_ref.py).

    loop    LLM_RL.environment.text_env_eval -> interact_environment            (environment.py:154-267, byte-for-byte copy)
    env     TextEnvToBatchedTextEnv(ReformatWordleEnvironment(WordleEnvironment(Vocabulary.from_file(wordle_official.txt,
            fill_cache=False))))                                                 (exactly wordle/bc/eval_bc_gpt2.py:59-63)
    policy  oracle.cpu_policy.TorchCPUGPT2Policy — a subclass of the reference's own BatchedTextPolicy here; HF torch-CPU GPT-2
            `generate` stands in for the JaxSeq policy that cannot be installed

Must run in a process that has NOT imported lmrl_gym_b200.environment yet (bench.py runs it as `--impl reference`, and spawns that
as a subprocess for the `cpu_baseline` leg), so that the mirror module adopts the reference's classes.  Test infrastructure only.
"""
from __future__ import annotations

import os
import sys
import time


def run_ref_wordle(params, cfg, tokenizer, n_rollouts: int, bsize: int, max_input_length: int = 256, max_new_tokens: int = 6,
                   pad_to_max_input: bool = False, seed: int = 0, threads: int = 0):
    assert "lmrl_gym_b200.environment" not in sys.modules or getattr(sys.modules["lmrl_gym_b200.environment"], "_USING_REFERENCE_TYPES", False), \
        "import oracle.ref_runner (and call build_ref.activate()) before anything imports lmrl_gym_b200.environment"
    from . import build_ref
    build_ref.activate()
    import torch
    if threads:
        torch.set_num_threads(threads)
    from LLM_RL.environment import text_env_eval                                   # the reference's own loop
    from llm_rl_scripts.wordle.env.env import ReformatWordleEnvironment, WordleEnvironment
    from llm_rl_scripts.wordle.env.game import Vocabulary
    import lmrl_gym_b200.environment as mirror
    assert mirror._USING_REFERENCE_TYPES
    from .cpu_policy import TorchCPUGPT2Policy
    from .mt19937 import seed_stream
    vocab = Vocabulary.from_file(vocab_file=os.path.join(build_ref.DST, "llm_rl_scripts", "wordle", "vocab", "wordle_official.txt"), fill_cache=False)
    env = ReformatWordleEnvironment(WordleEnvironment(vocab))
    pol = TorchCPUGPT2Policy(params, cfg, tokenizer, max_input_length, max_new_tokens, pad_to_max_input, seed)
    t0 = time.perf_counter()
    inter, summary = text_env_eval(env=env, policy=pol, n_rollouts=n_rollouts, seed_generator=seed_stream(seed), bsize=bsize, verbose=False)
    dt = time.perf_counter() - t0
    steps = sum(len(ep) for ep in inter)
    return dict(env_steps=steps, seconds=dt, env_steps_per_s=steps / dt, cores=torch.get_num_threads(), host_cpus=os.cpu_count(),
                summary={k: {kk: float(vv) for kk, vv in v.items()} for k, v in summary.items()})
