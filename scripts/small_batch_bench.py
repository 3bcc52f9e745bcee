#!/usr/bin/env python
"""Decode throughput of small batches (GPU box only): tokens/s of `B200GPT2Policy.generate_tokens` at B = 1 .. 128 with the
weight-streaming GEMV path on and off, against the HBM roofline of one decode pass (every weight byte once per pass).

    python scripts/small_batch_bench.py [--batches 1,2,4,8,16,32,128] [--new 65] [--medium]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(batches=(1, 2, 4, 8, 16, 32, 128), n_new=65, medium=False, reps=3, peaks_gbs=None, pdl=False):
    import numpy as np
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    cfg = GPT2Config.medium() if medium else GPT2Config.small()
    params = init_gpt2_params(cfg, seed=0)
    params["wte"] = params["wte"].copy()
    tok = ByteTokenizer.synthetic()
    params["wte"][tok.newline_id()] *= 0.0                       # keep EOS improbable-but-possible as in a random-init model
    d, ff, L, V = cfg.d_model, cfg.d_ff, cfg.n_layer, cfg.n_vocab
    weight_bytes = 2.0 * (L * (3 * d * d + d * d + 2 * d * ff) + V * d)
    out = []
    for B in batches:
        eng = Engine(max_envs=B, max_input_length=32, max_new_tokens=n_new, tokenizer=tok)
        pol = B200GPT2Policy(eng, params, cfg, seed=0)
        if pdl:
            eng.set_use_pdl(True)
        rng = np.random.default_rng(B)
        prompts = [rng.integers(0, V, size=8).tolist() for _ in range(B)]
        row = dict(batch=B)
        for name, rows in (("gemv", 8), ("tensor_core_tiles", 0)):
            eng.set_option("gemv_max_rows", rows)
            for _ in range(3):                                      # eager, capture, replay
                gens = pol.generate_tokens(prompts)
            eng.sync()
            t0 = time.perf_counter()
            n_tok = 0
            for _ in range(reps):
                gens = pol.generate_tokens(prompts)
                n_tok += sum(len(g) for g in gens)
            dt = time.perf_counter() - t0
            passes = reps * n_new
            row[name] = dict(tok_per_s=round(n_tok / dt, 1), us_per_pass=round(1e6 * dt / passes, 2),
                             hbm_gbs=round(weight_bytes * passes / dt / 1e9, 1))
            if peaks_gbs:
                row[name]["hbm_frac"] = round(weight_bytes * passes / dt / 1e9 / peaks_gbs, 3)
            if B > 8:
                row.pop("gemv", None) if name == "gemv" else None
        row["weight_bytes_per_pass"] = weight_bytes
        out.append(row)
        eng.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,2,4,8,16,32,128")
    ap.add_argument("--new", type=int, default=65)
    ap.add_argument("--medium", action="store_true")
    ap.add_argument("--pdl", action="store_true", help="programmatic dependent launch on")
    args = ap.parse_args()
    peaks = None
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))["hbm_gbs"]
    for row in measure(tuple(int(x) for x in args.batches.split(",")), args.new, args.medium, peaks_gbs=peaks, pdl=args.pdl):
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
