#!/usr/bin/env python
"""A/B timing of the fused Wordle rollout (4096 episodes, GPT-2-small): CUDA graphs x two half-batch lanes (x programmatic dependent launch).

    python scripts/rollout_timing.py [--batch 4096] [--reps 4] [--letters]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--letters", action="store_true")
    ap.add_argument("--pdl", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    from lmrl_gym_b200.engine import Engine, ENV_WORDLE
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params, steer_to_letters
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200.wordle import B200WordleEnv
    B = args.batch
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config.small()
    params = init_gpt2_params(cfg, seed=0)
    if args.letters:
        params = steer_to_letters(params, tok)
    eng = Engine(max_envs=B, max_input_length=256, max_new_tokens=6, max_turns=6, tokenizer=tok)
    B200WordleEnv(eng, "wordle_official.txt")
    pol = B200GPT2Policy(eng, params, cfg, seed=0)
    s = pol.sampling()
    seeds = torch.from_numpy(np.random.default_rng(0).integers(0, 2**63, size=B, dtype=np.int64)).cuda()
    stream = torch.cuda.ExternalStream(eng.stream_ptr, device=torch.device("cuda", 0))
    for graphs, pdl, lanes in ((0, 0, 0), (0, 0, 1), (1, 0, 0), (1, 0, 1)) + (((1, 1, 1),) if args.pdl else ()):
        if True:
            eng.set_use_graphs(bool(graphs))
            eng.set_use_pdl(bool(pdl))
            eng.set_two_lanes(bool(lanes))
            for _ in range(3):
                eng.rollout_dev(ENV_WORDLE, seeds.data_ptr(), B, s, None)
            eng.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.reps):
                eng.rollout_dev(ENV_WORDLE, seeds.data_ptr(), B, s, None)
            e1.record(stream)
            eng.sync()
            print(json.dumps(dict(graphs=graphs, pdl=pdl, two_lanes=lanes, ms_per_rollout=e0.elapsed_time(e1) / args.reps)), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
