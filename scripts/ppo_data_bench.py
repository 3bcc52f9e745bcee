#!/usr/bin/env python
"""Measure the rows either side of the rollout on one B200 — not the headline bench.

    python scripts/ppo_data_bench.py [--chains 4096]

  ppo_data   one PPO collection pass over `--chains` Wordle-shaped token-trajectory chains (6 turns, the token shape of the letters
             variant): per-token log-probs of the policy and the initial policy + LinearHead values (two teacher-forced GPT-2-small
             passes, lmrl_forward_host) -> KL rewards / GAE / whitening (lmrl_ppo_advantages_host) -> PPOData records.
             Reported: tokens/s through B200PPOInference.get_ppo_data_from_token_trajectory_chain (host chains in, host records out,
             wall clock) and the device share (kernels of the two forwards, CUDA events).
  score      ILQL advantage re-ranking (build_ilql_score_fn): four Maze proposals per history, histories/s.
Prints one JSON line.  The CPU port of the same pass is timed by tests/test_gpu_ppo.py::test_ppo_data_throughput_vs_cpu_port (the
oracle is test infrastructure and is not importable from scripts/).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def wordle_chains(tok, n, rng):
    """n chains of 6 trajectories; trajectory t = the whole history up to and including turn t's feedback (what
    text_env_eval -> TextTrajectoryChain conversion gives the PPO driver, wordle/ppo/train_ppo_gpt2.py)."""
    from lmrl_gym_b200 import ppo as P
    letters = "abcdefghijklmnopqrstuvwxyz"
    chains = []
    for _ in range(n):
        ids = list(tok.encode("Wordle:\n"))
        act = [False] * len(ids)
        rew = [0.0] * len(ids)
        trajs = []
        for t in range(6):
            a = tok.encode(" ".join(letters[i] for i in rng.integers(0, 26, 5)) + "\n")
            fb = tok.encode(" ".join("byg"[i] for i in rng.integers(0, 3, 5)) + "\n")
            ids += a; act += [True] * len(a); rew += [0.0] * (len(a) - 1) + [-1.0]
            ids += fb; act += [False] * len(fb); rew += [0.0] * len(fb)
            trajs.append(P.TokenTrajectory(np.asarray(ids, np.int32), np.asarray(act, np.bool_), np.asarray(rew, np.float32), np.asarray(t == 5, np.bool_)))
        chain = None
        for tr in reversed(trajs):
            chain = P.TokenTrajectoryChain(tr, chain)
        chains.append(chain)
    return chains


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=4096)
    args = ap.parse_args()
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params, init_mlp_head_params
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200 import ppo as P, reranker as RR, value_inference as VI
    from lmrl_gym_b200.environment import Text
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config.small()
    rng = np.random.default_rng(0)
    eng = Engine(max_envs=4096, max_input_length=128, max_new_tokens=6, tokenizer=tok)
    pol, ini = init_gpt2_params(cfg, seed=0), init_gpt2_params(cfg, seed=1)
    lin = dict(dense=dict(kernel=(rng.standard_normal((cfg.d_model, 1)) / 28).astype(np.float32), bias=np.zeros(1, np.float32)))
    B200GPT2Policy(eng, pol, cfg)
    inf = P.B200PPOInference(eng, None, lin, initial_policy_params=ini, config=cfg, pad_token_id=tok.pad_token_id)
    chains = wordle_chains(tok, args.chains, rng)
    n_tok = sum(t.tokens.shape[0] for c in chains for t in c.to_list())
    kw = dict(bsize=1 << 20, max_length=128, gamma=1.0, lam=0.95, kl_weight=0.001, use_advantage_whitening=True, verbose=False)
    inf.get_ppo_data_from_token_trajectory_chain(chains[:256], **kw)          # warm-up (allocations, tensor maps)
    eng.sync()
    t0 = time.perf_counter()
    datas, kls = inf.get_ppo_data_from_token_trajectory_chain(chains, **kw)
    wall = time.perf_counter() - t0
    # device share: the two forwards alone, CUDA-event timed by the engine's per-launch profiler
    seqs = [t.tokens for c in chains for t in c.to_list()]
    eng.profile_enable(True)
    VI.forward_tokens(eng, 0, 0, seqs, logprobs=True, v=True)
    VI.forward_tokens(eng, 1, 0, seqs, logprobs=True)
    prof = eng.profile_read()
    eng.profile_enable(False)
    dev_ms = float(sum(prof["ms"].values()))
    flops = float(prof["work"]["gemm"])
    out = dict(ppo_data=dict(chains=args.chains, trajectories=len(datas), tokens=n_tok, wall_s=round(wall, 3), tokens_per_s=round(n_tok / wall, 1),
                             forward_kernel_ms=round(dev_ms, 2), forward_tokens_per_s_device=round(2 * n_tok / (dev_ms / 1e3), 1),
                             gemm_tflops=round(flops / (float(prof["ms"]["gemm"]) / 1e3) / 1e12, 1) if prof["ms"]["gemm"] else None,
                             mean_kl=float(np.mean(kls)), note="two towers x %d rows; host wall clock includes chain flattening and PPOData slicing" % n_tok))
    # ILQL advantage re-ranking
    q1, q2 = init_mlp_head_params(cfg, seed=2), init_mlp_head_params(cfg, seed=3)
    v = init_mlp_head_params(cfg, seed=4, n_out=1, layer2_bias_init=0.0)
    eng.release_slots()                # the PPO objects above are finished: tower 1 / head set 0 are reused for the value networks
    ilql = VI.B200ILQLInference(VI.B200ValueRLInference(eng, ini, q1, q2, v, cfg, tower=1, head_set=0),
                                VI.B200ValueRLInference(eng, pol, q1, q2, None, cfg, tower=2, head_set=1))
    fn = RR.build_ilql_score_fn(ilql, None, tok, max_length=128, value_weight=1.0, logit_weight=None, bsize=4096)
    obs = "The goal is at position 8, 6. Your current position is at position 5, 3. There are walls to your left, below you.\n"
    hists = [(Text(obs, False),) for _ in range(1024)]
    props = [p for h in hists for p in RR.maze_proposal_function(h)]
    fn(props[:64])
    eng.sync()
    t0 = time.perf_counter()
    scores = fn(props)
    wall = time.perf_counter() - t0
    out["score"] = dict(histories=len(hists), proposals=len(props), wall_s=round(wall, 3), histories_per_s=round(len(hists) / wall, 1))
    eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
