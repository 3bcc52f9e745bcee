#!/usr/bin/env python
"""BASELINE.json configs[3] shape (not the headline bench): chess endgames, GPT-2-medium shapes (random init), B concurrent games,
host-side rules + opponent, device policy.  Reports env-steps/s and where the time goes (policy.act on the GPU vs env.step on the host).

    python scripts/chess_bench.py [--batch 2048] [--max-moves 6] [--engine /path/to/stockfish]
Without --engine the opponent is the reference's random_opponent mode (Stockfish is not part of this image).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--max-moves", type=int, default=6)
    ap.add_argument("--engine", default=None)
    args = ap.parse_args()
    import numpy as np
    from lmrl_gym_b200 import chess_env as CE
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.environment import TextEnvToBatchedTextEnv, interact_environment
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config.medium()
    B = args.batch
    eng = Engine(max_envs=B, max_input_length=128, max_new_tokens=6, tokenizer=tok)
    pol = B200GPT2Policy(eng, init_gpt2_params(cfg, seed=0), cfg, seed=0)
    fen = CE.large_piece_random_endgame("kQK", np.random.default_rng(0))
    opp = (lambda: CE.UciOpponent(args.engine)) if args.engine else None
    env = CE.FenChessHistoryEnv(from_position=fen, random_opponent=args.engine is None, opponent=opp() if opp else None, max_moves=args.max_moves)
    t_act = [0.0]
    real_act = pol.act

    def timed_act(*a, **k):
        t0 = time.perf_counter()
        out = real_act(*a, **k)
        t_act[0] += time.perf_counter() - t0
        return out

    pol.act = timed_act
    interact_environment(env, pol, env_seed=[None] * 8, env_options=[None] * 8, bsize=8)          # warm-up
    t_act[0] = 0.0
    t0 = time.perf_counter()
    inter = interact_environment(env, pol, env_seed=[None] * B, env_options=[None] * B, bsize=B)
    wall = time.perf_counter() - t0
    steps = sum(len(ep) for ep in inter)
    print(json.dumps(dict(workload="chess endgame kQK x GPT-2-medium (random init), %d concurrent games, max_moves %d, %s opponent" %
                          (B, args.max_moves, "UCI engine" if args.engine else "random"), env_steps=steps, wall_s=round(wall, 3),
                          env_steps_per_s=round(steps / wall, 1), policy_act_s=round(t_act[0], 3), host_env_and_loop_s=round(wall - t_act[0], 3))))
    eng.close()


if __name__ == "__main__":
    main()
