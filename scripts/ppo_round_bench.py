#!/usr/bin/env python
"""One PPO round at BASELINE.json configs[4] shapes (Wordle, GPT-2-small, 4096 rollouts per GPU): rollout / PPO-data / update seconds
reported SEPARATELY.  Not the headline bench.

    python scripts/ppo_round_bench.py [--rollouts 4096] [--rounds 2] [--train-bsize 32] [--max-updates 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/ppo_round_bench.py ...

rollout, PPO data AND the gradient step run on the engine (hand-written CUDA: --backward b200, the default); --backward torch selects the
torch-autograd library path for comparison.  Under torchrun the fp32 gradients are all-reduced over NCCL.  Prints one JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rollouts", type=int, default=4096)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--train-bsize", type=int, default=32)
    ap.add_argument("--max-updates", type=int, default=32, help="cap on optimizer steps per round (the reference's max_steps), to bound the run")
    ap.add_argument("--backward", choices=("b200", "torch"), default="b200", help="b200 = hand-written backward pass (csrc/train_kernels.inl); torch = autograd library path")
    ap.add_argument("--letters", action="store_true", help="steer the LM head to letter tokens (token shape of a trained policy)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params, steer_to_letters
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200.wordle import B200WordleEnv
    from lmrl_gym_b200 import ppo as P, ppo_train as T
    from lmrl_gym_b200.randomness import seed_generator
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config.small()
    B = args.rollouts
    eng = Engine(max_envs=B, device=local, max_input_length=256, max_new_tokens=6, max_turns=6, tokenizer=tok)
    params = init_gpt2_params(cfg, seed=0)
    if args.letters:
        params = steer_to_letters(params, tok)
    lin = dict(dense=dict(kernel=(np.random.default_rng(1).standard_normal((cfg.d_model, 1)) / 28).astype(np.float32), bias=np.zeros(1, np.float32)))
    env = B200WordleEnv(eng, "wordle_official.txt", bad_word_reward=-10.0)          # train_ppo_gpt2.py:211
    pol = B200GPT2Policy(eng, params, cfg, seed=rank)
    inf = P.B200PPOInference(eng, None, lin, initial_policy_params=params, config=cfg, pad_token_id=tok.pad_token_id)
    trainer = T.B200PPOTrain(eng, params, lin, cfg, lr=1e-5, pad_token_id=tok.pad_token_id, backward=args.backward, max_rows=args.train_bsize * 264, max_scored_rows=args.train_bsize * 264)
    seeds = seed_generator(0)
    for _ in range(rank * B * args.rounds):          # contiguous shards of one seed stream
        next(seeds)

    logs = []

    def on_round(r, log):
        logs.append(log)

    T.ppo_train_loop(env, pol, inf, trainer, P.AdaptiveKLController(0.001, 0.1, 10000), n_rounds=args.rounds, n_rollouts=B, rollout_bsize=B, epochs=1,
                     train_bsize=args.train_bsize, max_length=256 + 6, gamma=1.0, lam=0.95, seed_generator=seeds, max_steps_per_round=args.max_updates, on_round=on_round)
    if world > 1:
        dist.barrier()
    if rank == 0:
        last = logs[-1]
        upd = last["updates"]
        print(json.dumps(dict(workload="PPO round: Wordle x GPT-2-small, %d rollouts/GPU, train_bsize %d, %d optimizer steps/round (capped)" % (B, args.train_bsize, upd),
                              n_gpus=world, rounds=logs, env_steps_per_s_rollout=round(world * last["env_steps"] / last["rollout_s"], 1),
                              update_s_per_step=round(last["update_s"] / max(upd, 1), 4),
                              update_impl=("hand-written backward (train_kernels.inl) + lmrl_adamw_dev" if args.backward == "b200" else "torch autograd (library path) + lmrl_adamw_dev") +
                              (" + NCCL all-reduce" if world > 1 else ""))))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
