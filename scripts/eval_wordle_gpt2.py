#!/usr/bin/env python
"""Wordle policy evaluation on the device engine — the driver a user of llm_rl_scripts/wordle/bc/eval_bc_gpt2.py switches to.
Same flags (mesh flags are gone: one process per GPU, launch under torchrun to shard the rollouts over ranks), same env construction,
same ``text_env_eval`` call, same outputs (``interactions.pkl``, ``interactions_summary.json``).

    python scripts/eval_wordle_gpt2.py config gpt2 wordle_official.txt --policy-n-rollouts 4096 --policy-bsize 4096 --policy-max-output-length 6

Differences, all flagged: weights come from ``lmrl_gym_b200.load`` (JaxSeq checkpoints are not readable here); ``policy_num_beams`` must
be 1 (the reference's default; beam search is not on the path); ``--seed`` feeds ``seed_generator`` so that runs are reproducible (the
reference passes no seeds → OS entropy).
"""
from typing import Optional

import tyro

from _driver_common import dump_outputs, load_tokenizer


def main(
    model_load_mode: str,
    model_load_path: str,
    vocab_file: str,

    /,

    outputs_path: Optional[str] = None,
    tokenizer_path: Optional[str] = None,

    bf16_activations: bool = True,

    policy_n_rollouts: int = 32,
    policy_bsize: int = 1,
    policy_max_input_length: int = 256,
    policy_max_output_length: int = 256,
    policy_do_sample: bool = True,
    policy_num_beams: int = 1,
    policy_temperature: Optional[float] = None,
    policy_top_p: Optional[float] = None,
    policy_top_k: Optional[int] = None,

    force_pad_embeddings: bool = False,
    seed: Optional[int] = None,
    print_episodes: int = 8,
):
    print(locals())
    assert policy_num_beams == 1, "beam search is not implemented (every shipped launch line uses num_beams=1)"
    assert bf16_activations, "the engine computes in bf16 with fp32 accumulation and an fp32 residual stream"
    from lmrl_gym_b200 import distributed as D
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.environment import text_env_eval, text_history_to_str
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    from lmrl_gym_b200.load import load_params
    from lmrl_gym_b200.randomness import seed_generator
    from lmrl_gym_b200.wordle import B200WordleEnv

    rank, world, device = D.init_from_env()
    tokenizer = load_tokenizer(tokenizer_path)
    params, cfg = load_params(model_load_mode, model_load_path, seed=2, force_pad_embeddings=force_pad_embeddings)
    engine = Engine(max_envs=policy_bsize, device=device, max_input_length=policy_max_input_length,
                    max_new_tokens=policy_max_output_length, tokenizer=tokenizer)
    env = B200WordleEnv(engine, vocab=vocab_file)                                     # ReformatWordleEnvironment(WordleEnvironment(vocab))
    policy = B200GPT2Policy(engine, params, cfg, do_sample=policy_do_sample, temperature=policy_temperature or 1.0,
                            top_k=policy_top_k, top_p=policy_top_p, seed=0)
    seeds = None if seed is None else seed_generator(seed)
    if world == 1:
        raw, summary = text_env_eval(env=env, policy=policy, n_rollouts=policy_n_rollouts, bsize=policy_bsize, seed_generator=seeds)
    else:                                                   # this rank's block of the one seed stream; the summary is the global one
        raw, summary = D.text_env_eval_sharded(env, policy, policy_n_rollouts, seed_generator=seeds, bsize=policy_bsize, rank=rank, world=world)
    if rank == 0:
        for item in raw[:print_episodes]:
            print("=" * 25)
            print(text_history_to_str(item[-1].post_transition_history))
            print("=" * 25)
        print(summary)
        dump_outputs(outputs_path, [list(ep) for ep in raw], summary)
    engine.close()
    D.finalize()
    return {"generation_metrics": summary}


if __name__ == "__main__":
    tyro.cli(main)
