"""CPU reference arm: the reference's rollout loop + env (restated) + torch-CPU GPT-2 policy, timed on host cores.

Test infrastructure / bench baseline only.  ``kind`` = "port": the reference itself (/root/reference) is Python and
does not exist on the GPU box, and its policy stack (JAX) is not installable, so this runs the oracle restatement:
  loop    lmrl_gym_b200.environment.text_env_eval generic path (mirrors LLM_RL/environment.py:154-267; a (TextEnv,
          BatchedTextPolicy) pair never takes the fused path)
  env     oracle.wordle.ReformatWordleOracle inside TextEnvToBatchedTextEnv (deep copy per slot per reset, pure-Python
          vocabulary filter per step — the reference's cost profile)
  policy  oracle.cpu_policy.TorchCPUGPT2Policy (HF torch GPT-2 generate)
"""
from __future__ import annotations

import os
import time

import torch

from lmrl_gym_b200.environment import Text, TextEnv, text_env_eval
from lmrl_gym_b200.vocab import load_vocab_words
from .cpu_policy import TorchCPUGPT2Policy
from .mt19937 import seed_stream
from .wordle import ReformatWordleOracle


class OracleWordleTextEnv(TextEnv):
    def __init__(self, words, require_words_in_vocab=True, bad_word_reward=-1.0):
        self.o = ReformatWordleOracle(words, require_words_in_vocab, bad_word_reward)

    def reset(self, seed=None, options=None):
        return tuple(Text(t, a) for t, a in self.o.reset(seed))

    def step(self, text_history):
        h, r, d = self.o.step(tuple((t.text, t.is_action) for t in text_history))
        return tuple(Text(t, a) for t, a in h), r, d


def run_cpu_wordle(params, cfg, tokenizer, n_rollouts: int, bsize: int, max_input_length: int = 256, max_new_tokens: int = 6,
                   pad_to_max_input: bool = False, vocab: str = "wordle_official.txt", seed: int = 0, threads: int = 0):
    if threads:
        torch.set_num_threads(threads)
    env = OracleWordleTextEnv(load_vocab_words(vocab))
    pol = TorchCPUGPT2Policy(params, cfg, tokenizer, max_input_length, max_new_tokens, pad_to_max_input, seed)
    t0 = time.perf_counter()
    inter, summary = text_env_eval(env, pol, n_rollouts=n_rollouts, seed_generator=seed_stream(seed), bsize=bsize, verbose=False)
    dt = time.perf_counter() - t0
    steps = sum(len(ep) for ep in inter)
    return dict(env_steps=steps, seconds=dt, env_steps_per_s=steps / dt, cores=torch.get_num_threads(), host_cpus=os.cpu_count(),
                summary={k: {kk: float(vv) for kk, vv in v.items()} for k, v in summary.items()})
