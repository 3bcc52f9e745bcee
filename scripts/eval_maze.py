#!/usr/bin/env python
"""Maze policy evaluation on the device engine — the driver a user of llm_rl_scripts/maze/bc/eval_bc.py switches to: reward evaluation
from every free start cell (sampling policy or re-ranker over the four moves) and "move accuracy" of the re-ranker against the BFS-optimal
move.  Same flags, same result tree (``reward_eval`` / ``move_accuracy``), same output files.

    python scripts/eval_maze.py config gpt2 --policy-n-rollouts 32 --policy-bsize 32 --policy-max-output-length 3

Differences, all flagged: weights come from ``lmrl_gym_b200.load``; mesh flags are gone; ``policy_num_beams`` must be 1; ``--seed`` makes
the rollouts reproducible; the move-accuracy pass scores every cell through ``policy.act`` exactly as the reference does.
"""
from collections import defaultdict
from typing import Optional

import numpy as np
import tyro

from _driver_common import dump_outputs, load_tokenizer, to_floats


def _flatten(d, prefix=()):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(_flatten(v, prefix + (k,)))
        else:
            out[prefix + (k,)] = v
    return out


def _unflatten(flat):
    out = {}
    for ks, v in flat.items():
        node = out
        for k in ks[:-1]:
            node = node.setdefault(k, {})
        node[ks[-1]] = v
    return out


def main(
    model_load_mode: str,
    model_load_path: str,

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

    maze_name: str = "double_t_maze",
    fully_observed: bool = True,
    maze_last_k: int = 1,
    maze_reward_function: str = "standard_reward",

    do_accuracy_eval: bool = True,
    do_reward_eval: bool = True,
    use_reranker_for_reward_eval: bool = False,

    force_pad_embeddings: bool = False,
    seed: Optional[int] = None,
):
    print(locals())
    assert policy_num_beams == 1, "beam search is not implemented (every shipped launch line uses num_beams=1)"
    assert bf16_activations, "the engine computes in bf16 with fp32 accumulation and an fp32 residual stream"
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.environment import Text, text_env_eval
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    from lmrl_gym_b200.load import load_params
    from lmrl_gym_b200.maze import B200MazeEnv, describe, maze_solver
    from lmrl_gym_b200.randomness import seed_generator
    from lmrl_gym_b200.reranker import ReRankerPolicy, ReRankerSamplePolicy, build_bc_score_fn, maze_proposal_function

    describe_function = "describe_observation_give_position" if fully_observed else "describe_observation_only_walls"
    tokenizer = load_tokenizer(tokenizer_path)
    params, cfg = load_params(model_load_mode, model_load_path, seed=2, force_pad_embeddings=force_pad_embeddings)
    engine = Engine(max_envs=max(policy_bsize, 4), max_input_length=policy_max_input_length, max_new_tokens=policy_max_output_length,
                    max_turns=101, tokenizer=tokenizer)
    env = B200MazeEnv(engine, maze_name=maze_name, describe_function=describe_function, reward_function=maze_reward_function, last_k=maze_last_k)
    goal = tuple(env.goal_default)
    possible_positions = [tuple(int(v) for v in p) for p in np.argwhere(env.grid == 0) if tuple(p) != goal]
    optimal_policy = maze_solver(env.grid, [goal])
    sampler = B200GPT2Policy(engine, params, cfg, do_sample=policy_do_sample, temperature=policy_temperature or 1.0,
                             top_k=policy_top_k, top_p=policy_top_p, seed=0)

    def reranker(sample: bool):
        score_fn = build_bc_score_fn(sampler, max_length=policy_max_input_length + policy_max_output_length, bsize=policy_bsize)
        return (ReRankerSamplePolicy if sample else ReRankerPolicy)(proposal_fn=maze_proposal_function, score_fn=score_fn)

    all_results, interactions = {}, {}
    seeds = None if seed is None else seed_generator(seed)
    if do_reward_eval:
        policy = reranker(policy_do_sample) if use_reranker_for_reward_eval else sampler
        results, avg = {}, defaultdict(float)
        for position in possible_positions:
            interactions[str(position)], results[str(position)] = text_env_eval(
                env=env, policy=policy, n_rollouts=policy_n_rollouts, verbose=False, seed_generator=seeds,
                env_options={"init_position": position}, bsize=policy_bsize)
            for k, v in _flatten(results[str(position)]).items():
                avg[k] += v
        results["avg_reward"] = _unflatten({k: v / len(possible_positions) for k, v in avg.items()})
        all_results["reward_eval"] = results
    if do_accuracy_eval:
        policy = reranker(False)
        results, num_correct = {}, 0
        for position in possible_positions:
            history = (Text(describe(1, env.grid, position, goal), False),)
            prediction = policy.act(history)[-1].text
            results[str(position)] = prediction.lower().strip() == optimal_policy[position].lower().strip()
            num_correct += int(results[str(position)])
        results["avg_accuracy"] = num_correct / len(possible_positions)
        all_results["move_accuracy"] = results
    print(to_floats(all_results.get("reward_eval", {}).get("avg_reward", {})), all_results.get("move_accuracy", {}).get("avg_accuracy"))
    if outputs_path is not None:
        dump_outputs(outputs_path, {k: [list(ep) for ep in v] for k, v in interactions.items()}, all_results, name="interactions")
    engine.close()
    return all_results


if __name__ == "__main__":
    tyro.cli(main)
