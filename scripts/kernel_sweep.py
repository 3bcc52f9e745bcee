#!/usr/bin/env python
"""Kernel microbenchmark over the shapes of the bench workload (GPU box only): TFLOP/s per GEMM, GB/s for attention.

    python scripts/kernel_sweep.py [--batch 4096] [--iters 20] [--medium]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--medium", action="store_true")
    ap.add_argument("--shapes", default="1:40,4:8,8:48,24:72")
    ap.add_argument("--gemm-impl", type=int, default=-1, help="0 mma.sync, 1 tcgen05 one CTA per tile, 2 tcgen05 CTA pairs on prefill shapes")
    args = ap.parse_args()
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    cfg = GPT2Config.medium() if args.medium else GPT2Config.small()
    eng = Engine(max_envs=args.batch, device=0, max_input_length=256, max_new_tokens=6, max_turns=6, tokenizer=ByteTokenizer.synthetic())
    B200GPT2Policy(eng, init_gpt2_params(cfg, seed=0), cfg, seed=0)
    if args.gemm_impl >= 0:
        eng.set_gemm_impl(args.gemm_impl)
    d, ff, V, H = cfg.d_model, cfg.d_ff, cfg.n_vocab, cfg.n_head
    for spec in args.shapes.split(","):
        rpe, ctx = (int(x) for x in spec.split(":"))
        M = args.batch * rpe
        ms = eng.bench_kernels(M, ctx, args.iters)
        n_slots = min(M, args.batch)
        fl = dict(qkv=2.0 * M * 3 * d * d, proj=2.0 * M * d * d, fc=2.0 * M * d * ff, fc2=2.0 * M * d * ff, head=2.0 * n_slots * d * V)
        # attention bytes: every (env, head) K and V prefix once + q/out rows
        att_bytes = n_slots * H * ctx * 64 * 2 * 2 + 2 * M * d * 2
        out = dict(M=M, rows_per_env=rpe, ctx=ctx)
        for k, f in fl.items():
            out[k] = dict(ms=round(ms[k], 4), tflops=round(f / ms[k] / 1e9, 1))
        out["attention"] = dict(ms=round(ms["attention"], 4), gbs=round(att_bytes / ms["attention"] / 1e6, 1))
        out["layernorm"] = dict(ms=round(ms["layernorm"], 4), gbs=round(M * d * 6 / ms["layernorm"] / 1e6, 1))
        out["embed"] = dict(ms=round(ms["embed"], 4))
        print(json.dumps(out), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
