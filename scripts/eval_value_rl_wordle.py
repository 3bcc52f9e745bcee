#!/usr/bin/env python
"""Wordle evaluation of a value-perturbed policy (logits = pi_beta + beta * min Q) on the device engine — the driver a user of
llm_rl_scripts/wordle/ilql/eval_ilql_gpt2.py (two MLP Q heads) or wordle/mc/eval_mc_gpt2.py (one LinearHead Q head) switches to.
Same positional arguments and flags; ``--algorithm ilql|mc`` selects the head layout the two reference files hard-code:
``<model_load_path>/base`` + ``q1_head`` + ``q2_head`` (ilql) or ``base`` + ``q_head`` (mc).

    python scripts/eval_value_rl_wordle.py config gpt2 config gpt2 wordle_official.txt --algorithm mc --policy-n-rollouts 4096 \
        --policy-bsize 4096 --policy-max-output-length 6

In ``config`` mode the paths name architectures (random init); in ``params`` mode ``<model_load_path>/base``, ``q1_head``, ... are either
the reference's checkpoint directories (``params.msgpack`` [+ ``config.json``]) or ``.npz`` files written by ``lmrl_gym_b200.load.save_params``
/ ``save_head_params``.
"""
import os
from typing import Optional

import tyro

from _driver_common import dump_outputs, load_tokenizer


def main(
    model_load_mode: str,
    model_load_path: str,
    pi_beta_load_mode: str,
    pi_beta_load_path: str,
    vocab_file: str,

    /,

    algorithm: str = "ilql",
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
    policy_beta: float = 16.0,

    force_pad_embeddings: bool = False,
    seed: Optional[int] = None,
    print_episodes: int = 8,
):
    print(locals())
    assert policy_num_beams == 1 and bf16_activations and algorithm in ("ilql", "mc")
    from lmrl_gym_b200 import distributed as D
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.environment import text_env_eval, text_history_to_str
    from lmrl_gym_b200.gpt2 import B200GPT2ValuePolicy
    from lmrl_gym_b200.load import ModelLoadMode, load_head_params, load_params
    from lmrl_gym_b200.randomness import seed_generator
    from lmrl_gym_b200.wordle import B200WordleEnv

    rank, world, device = D.init_from_env()
    tokenizer = load_tokenizer(tokenizer_path)
    from_files = ModelLoadMode(model_load_mode) is ModelLoadMode.PARAMS
    def sub(name):               # the reference's layout <path>/<name>/params.msgpack (eval_ilql_gpt2.py:78-110), or this repo's <path>/<name>.npz
        if not from_files:
            return model_load_path
        d = os.path.join(model_load_path, name)
        return d if os.path.isdir(d) else d + ".npz"
    pi_beta, cfg = load_params(pi_beta_load_mode, pi_beta_load_path, seed=0, force_pad_embeddings=force_pad_embeddings)
    base, base_cfg = load_params(model_load_mode, sub("base"), seed=1, force_pad_embeddings=force_pad_embeddings)
    assert (base_cfg.n_layer, base_cfg.d_model, base_cfg.n_head) == (cfg.n_layer, cfg.d_model, cfg.n_head), "pi_beta and the value base share one architecture"
    if algorithm == "ilql":
        q1 = load_head_params(model_load_mode, sub("q1_head"), cfg, kind="mlp", seed=2)
        q2 = load_head_params(model_load_mode, sub("q2_head"), cfg, kind="mlp", seed=3)
    else:
        q1, q2 = load_head_params(model_load_mode, sub("q_head"), cfg, kind="linear", seed=2), None
    engine = Engine(max_envs=policy_bsize, device=device, max_input_length=policy_max_input_length, max_new_tokens=policy_max_output_length,
                    tokenizer=tokenizer)
    env = B200WordleEnv(engine, vocab=vocab_file)
    policy = B200GPT2ValuePolicy(engine, (pi_beta, base, q1, q2), cfg, beta=policy_beta, do_sample=policy_do_sample,
                                 temperature=policy_temperature or 1.0, top_k=policy_top_k, top_p=policy_top_p, seed=0)
    seeds = None if seed is None else seed_generator(seed)
    if world == 1:
        raw, summary = text_env_eval(env=env, policy=policy, n_rollouts=policy_n_rollouts, bsize=policy_bsize, seed_generator=seeds)
    else:
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
