#!/usr/bin/env python
"""A/B timing of the fused Wordle rollout (GPT-2-small, graphs on) over the engine's named switches (GPU box only).

    python scripts/ab_options.py [--batch 4096] [--reps 4] [--combos fuse_ln=1,attn_wide=1;fuse_ln=0,attn_wide=1;...]
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
    ap.add_argument("--combos", default="fuse_ln=0,attn_wide=0;fuse_ln=1,attn_wide=0;fuse_ln=0,attn_wide=1;fuse_ln=1,attn_wide=1")
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
    ret = torch.zeros(B, dtype=torch.float32, device="cuda")
    ln = torch.zeros(B, dtype=torch.int32, device="cuda")
    from lmrl_gym_b200 import _lib
    out = _lib.lmrl_rollout_out(ret.data_ptr(), ln.data_ptr(), None, None, None, None, None, None, None)
    for combo in args.combos.split(";"):
        opts = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in combo.split(",") if kv)
        for k, v in opts.items():
            if k == "pdl":
                eng.set_use_pdl(bool(v))
            elif k == "graphs":
                eng.set_use_graphs(bool(v))
            elif k == "gemm_impl":
                eng.set_gemm_impl(v)
            else:
                eng.set_option(k, v)
        for _ in range(3):
            eng.rollout_dev(ENV_WORDLE, seeds.data_ptr(), B, s, out)
        eng.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.reps):
            eng.rollout_dev(ENV_WORDLE, seeds.data_ptr(), B, s, out)
        e1.record(stream)
        eng.sync()
        print(json.dumps(dict(opts, ms_per_rollout=round(e0.elapsed_time(e1) / args.reps, 3), env_steps=int(ln.sum().item()),
                              mean_return=float(ret.mean().item()))), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
