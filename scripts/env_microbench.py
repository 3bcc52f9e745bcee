#!/usr/bin/env python
"""Env kernels alone (bench variant 2b): scripted RandomMixturePolicy-equivalent Wordle rollouts, no policy forward.

    python scripts/env_microbench.py [--batch 4096] [--reps 20]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--vocab", default="wordle_official.txt")
    args = ap.parse_args()
    import numpy as np
    import torch
    from lmrl_gym_b200 import _lib
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200.wordle import B200WordleEnv
    B = args.batch
    eng = Engine(max_envs=B, max_new_tokens=6, max_turns=6, tokenizer=ByteTokenizer.synthetic())
    B200WordleEnv(eng, args.vocab)
    dev = torch.device("cuda", 0)
    seeds = torch.from_numpy(np.random.default_rng(0).integers(0, 2**63, size=B, dtype=np.int64)).to(dev)
    ep_ret = torch.zeros(B, dtype=torch.float32, device=dev)
    ep_len = torch.zeros(B, dtype=torch.int32, device=dev)
    ep_done = torch.zeros(B, dtype=torch.uint8, device=dev)
    out = _lib.lmrl_rollout_out(ep_ret.data_ptr(), ep_len.data_ptr(), ep_done.data_ptr(), None, None, None, None, None, None)
    stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    for _ in range(2):
        eng.wordle_scripted_rollout_dev(seeds.data_ptr(), B, 0.5, 1, out)
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for r in range(args.reps):
        eng.wordle_scripted_rollout_dev(seeds.data_ptr(), B, 0.5, 1 + r, out)
    e1.record(stream)
    eng.sync()
    ms = e0.elapsed_time(e1) / args.reps
    steps = int(ep_len.sum().item())
    print(json.dumps(dict(batch=B, ms_per_batch=ms, env_steps_per_batch=steps, env_steps_per_s=steps / ms * 1000.0,
                          mean_return=float(ep_ret.mean().item()), mean_length=steps / B)))
    eng.close()


if __name__ == "__main__":
    main()
