#!/usr/bin/env python
"""Behaviour cloning of a Wordle policy on the device engine, from data the device itself generates — the in-repo stand-in for the
reference pipeline ``wordle/misc/data_gen.py`` (RandomMixturePolicy episodes) -> ``wordle/bc/train_bc_gpt2.py`` (masked next-token
cross-entropy on the action tokens).  Purpose: a policy with a realistic valid / invalid action mix for bench variant 2c.

    python scripts/train_bc_wordle.py --n-data 65536 --steps 2000 --out gpurun_out/bc_wordle

STATUS: written in round 1 after the GPU budget was spent — every call it makes is covered by a GPU test (scripted rollout:
tests/test_gpu_env.py; gradient path of bc_step: tests/test_gpu_backward.py::test_bc_term_on_the_engine_matches_autograd), but the script
as a whole has NOT been run on a GPU yet.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="gpt2", help="gpt2 | gpt2-medium | JSON of GPT2Config fields (random init)")
    ap.add_argument("--vocab-file", default="wordle_official.txt")
    ap.add_argument("--n-data", type=int, default=65536, help="episodes to generate")
    ap.add_argument("--gen-batch", type=int, default=4096)
    ap.add_argument("--prob-smart", type=float, default=0.5)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--bsize", type=int, default=64)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--weight-decay", type=float, default=0.0)
    ap.add_argument("--max-length", type=int, default=128)
    ap.add_argument("--eval-rollouts", type=int, default=4096)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import numpy as np
    import torch
    from lmrl_gym_b200 import _lib
    from lmrl_gym_b200 import ppo_train as T
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.environment import text_env_eval
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    from lmrl_gym_b200.load import load_params, save_params
    from lmrl_gym_b200.randomness import seed_generator
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200.wordle import B200WordleEnv, scripted_rollout_records

    tok = ByteTokenizer.synthetic()
    params, cfg = load_params("config", args.model, seed=2)
    B, Tn, K = args.gen_batch, 6, 6
    eng = Engine(max_envs=max(B, args.eval_rollouts, args.bsize), max_input_length=256, max_new_tokens=K, max_turns=Tn, tokenizer=tok)
    env = B200WordleEnv(eng, args.vocab_file)
    policy = B200GPT2Policy(eng, params, cfg, seed=args.seed)

    # ---- data: RandomMixturePolicy(prob_smart) episodes from the env kernels alone
    dev = torch.device("cuda", 0)
    seeds = seed_generator(args.seed)
    records, t0 = [], time.perf_counter()
    for r in range((args.n_data + B - 1) // B):
        s = np.asarray([next(seeds) for _ in range(B)], dtype=np.uint64)
        seeds_dev = torch.from_numpy(s.view(np.int64)).to(dev)
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
        ep_ret, ep_len, ep_done = z(B, torch.float32), z(B, torch.int32), z(B, torch.uint8)
        rew, valid, fb = z((B, Tn), torch.float32), z((B, Tn), torch.uint8), z((B, Tn, 5), torch.uint8)
        aux, act, alen = z((B, Tn), torch.int32), z((B, Tn, K), torch.int32), z((B, Tn), torch.int32)
        out = _lib.lmrl_rollout_out(ep_ret.data_ptr(), ep_len.data_ptr(), ep_done.data_ptr(), rew.data_ptr(), valid.data_ptr(), fb.data_ptr(),
                                    aux.data_ptr(), act.data_ptr(), alen.data_ptr())
        eng.wordle_scripted_rollout_dev(seeds_dev.data_ptr(), B, args.prob_smart, 1 + r, out)
        eng.sync()
        records += scripted_rollout_records(act[:, :, 0].cpu().numpy(), fb.cpu().numpy(), rew.cpu().numpy(), ep_len.cpu().numpy(), ep_done.cpu().numpy())
    records = records[:args.n_data]
    gen_s = time.perf_counter() - t0
    wins = sum(r["reward"][-2] == 0.0 for r in records)
    print(json.dumps(dict(episodes=len(records), gen_s=round(gen_s, 2), win_rate=round(wins / len(records), 4))), flush=True)

    # ---- BC: masked next-token cross-entropy on the action tokens, hand-written backward + AdamW
    lin = dict(dense=dict(kernel=np.zeros((cfg.d_model, 1), np.float32), bias=np.zeros(1, np.float32)))
    rows = args.bsize * args.max_length
    trainer = T.B200PPOTrain(eng, params, lin, cfg, lr=args.lr, weight_decay=args.weight_decay, pad_token_id=tok.pad_token_id, backward="b200",
                             max_rows=rows, max_scored_rows=rows)
    rng = np.random.default_rng(args.seed)
    t0 = time.perf_counter()
    for step in range(args.steps):
        idx = rng.integers(0, len(records), size=args.bsize)
        batch = T.bc_batch_from_str_segments([records[i]["sequence"] for i in idx], tok, args.max_length, tok.pad_token_id)
        loss = trainer.bc_step(batch)
        if step % 100 == 0 or step == args.steps - 1:
            print(json.dumps(dict(step=step, bc_loss=round(loss, 4), s_per_step=round((time.perf_counter() - t0) / (step + 1), 4))), flush=True)

    # ---- evaluation of the cloned policy (the engine already holds the trained weights: bc_step hands them over)
    raw, summary = text_env_eval(env, policy, n_rollouts=args.eval_rollouts, bsize=args.eval_rollouts, seed_generator=seed_generator(args.seed + 1), verbose=False)
    valid_frac = float(np.mean([t.post_transition_history[-1].text != "\n" for ep in raw[:512] for t in ep]))
    print(json.dumps(dict(eval={k: {m: float(v) for m, v in d.items()} for k, d in summary.items()}, valid_action_fraction=round(valid_frac, 4))), flush=True)
    if args.out:
        os.makedirs(args.out, exist_ok=True)
        save_params(os.path.join(args.out, "policy_params.npz"), trainer.params_numpy(), cfg)
    eng.close()


if __name__ == "__main__":
    main()
