#!/usr/bin/env python
"""bench.py — env-steps/sec of batched Wordle rollouts with a GPT-2-small policy (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B] [--workload wordle|maze_po_medium]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one batch of B concurrent episodes run to completion through the hot path
(obs tokens -> policy.generate -> env.step -> reward, up to 6 turns): exactly what one iteration of the reference's
``text_env_eval`` batch loop does (LLM_RL/environment.py:223-242).  Prints ONE JSON line (rank 0).

  value      whole-job env-steps/s over exactly K steps, device-timed (CUDA events on the engine stream, max over ranks),
             seeds resident in HBM, the production launch path (the fused rollout replayed from its CUDA graph)
  e2e        the same metric through the public API ``text_env_eval(env, policy, ...)`` with host seeds in and host
             logs out (H2D + D2H inside the timed region), wall clock
  roofline   the dominant kernel (bf16 GEMM, tensor-bound): algorithmic FLOPs (2*M*N*K of every launch, live M read back
             from the device) / summed launch durations, from a second pass over the same K steps launched eagerly with
             one CUDA event per launch boundary; ``decode_attention_hbm`` is the HBM fraction of the decode-pass KV
             attention (the path's batched GEMV)
  env_only   env kernels alone with scripted RandomMixturePolicy-equivalent actions (variant 2b)
  letters_variant  the same engine with the LM head steered to letter tokens: the token shape of a trained policy (2d)
  cpu_baseline  the CPU restatement of the reference path (oracle port) on a bounded sample, host cores stated,
             charitable (pad to longest) and faithful (LEFT-pad to 256) variants
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
WORKLOAD = "wordle_official(2315 words) x GPT-2-small(random init, seed 0), 6 turns, max_new_tokens=6, max_input_length=256"
# BASELINE.json configs[2] (not the headline; `--workload maze_po_medium`): partially observed double-T maze, GPT-2-medium shapes
WORKLOAD_ILQL = ("wordle_official(2315 words) x ILQL policy at GPT-2-small shapes (pi_beta tower + value tower + two MLP Q heads, beta = 16, random init), "
                 "6 turns, max_new_tokens=6, max_input_length=256")
WORKLOAD_MAZE = ("double_t_maze partially observed (describe_observation_only_walls, last_k=40, max_steps=100) x GPT-2-medium"
                 "(random init, seed 0), max_new_tokens=3, max_input_length=512")


def load_traffic():
    """DRAM bytes per launch of the kernel families, from the committed ncu launch list of this same command
    (profiles/r01_traffic.json, written by scripts/ncu_summarize.py traffic); None when no capture has been committed."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p))["families"]
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    power_w_max=max(power) if power else None, samples=len(sm))


def build_policy_params(medium: bool = False):
    from lmrl_gym_b200.gpt2 import GPT2Config, init_gpt2_params
    cfg = GPT2Config.medium() if medium else GPT2Config.small()
    return cfg, init_gpt2_params(cfg, seed=0)


def _ref_runner():
    """The CPU arm's rollout function: the unmodified reference loop + env staged under oracle/_ref when present (kind "reference"),
    else the oracle port (kind "port").  Only `--impl reference` processes call this (the product arm spawns one for cpu_baseline)."""
    from oracle import build_ref
    if build_ref.available():
        build_ref.activate()
        from oracle.ref_runner import run_ref_wordle
        return run_ref_wordle, "reference", ("LLM_RL.environment.text_env_eval + ReformatWordleEnvironment(WordleEnvironment) imported unmodified from oracle/_ref "
                                             "(oracle/build_ref.py); policy = HF torch-CPU GPT2LMHeadModel.generate (the JaxSeq policy is not installable)")
    from oracle.cpu_rollout import run_cpu_wordle
    return run_cpu_wordle, "port", "oracle port of the loop + env (oracle/_ref absent); policy = HF torch-CPU GPT2LMHeadModel.generate"


def run_reference(args):
    """--impl reference: the reference's own CPU rollout path on all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm must use the host's cores whatever launched it
    n_threads = os.cpu_count() or 1
    torch.set_num_threads(n_threads)
    run, kind, how = _ref_runner()
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    cfg, params = build_policy_params()
    tok = ByteTokenizer.synthetic()
    bs = args.ref_batch
    for _ in range(min(args.warmup, 1)):
        run(params, cfg, tok, n_rollouts=4, bsize=4, threads=n_threads)
    steps_total, secs = 0, 0.0
    for k in range(args.steps):
        r = run(params, cfg, tok, n_rollouts=bs, bsize=bs, seed=k, threads=n_threads)
        steps_total += r["env_steps"]; secs += r["seconds"]
    value = steps_total / secs
    faith = run(params, cfg, tok, n_rollouts=max(bs // 4, 4), bsize=max(bs // 4, 4), seed=0, pad_to_max_input=True, threads=n_threads)
    sample = (f"{args.steps} x text_env_eval(n_rollouts={bs}, bsize={bs}) episodes of the same workload ({steps_total} env-steps in {secs:.1f} s), prompts padded to the "
              f"longest in the batch (charitable); faithful variant (LEFT-padded to 256 tokens, eval_bc_gpt2.py:100-104) on {max(bs // 4, 4)} episodes: "
              f"{faith['env_steps_per_s']:.2f} env-steps/s; {how}")
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000.0 * secs / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=dict(workload=WORKLOAD, batch_per_step=bs, policy="transformers GPT2LMHeadModel.generate torch-CPU fp32"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=torch.get_num_threads(), kind=kind, sample=sample, host_cpus=os.cpu_count(),
                                  faithful_value=faith["env_steps_per_s"], faithful_env_steps=faith["env_steps"], faithful_seconds=faith["seconds"],
                                  env_steps=steps_total, seconds=secs),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def cpu_baseline_subprocess(ref_batch):
    """cpu_baseline leg of the product arm: the reference arm in its own process (it imports oracle/_ref, which the product process
    never does), all host threads, one bounded step."""
    env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "RANK", "LOCAL_RANK", "WORLD_SIZE")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-batch", str(ref_batch)],
                         capture_output=True, text=True, env=env, timeout=900)
    for ln in reversed(res.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    raise RuntimeError("cpu_baseline subprocess printed no JSON line: " + res.stderr[-2000:])



# ---- SURVEY.md §8(d) constants (GPT-2-small): matmul parameters, FLOPs per token, KV bytes per token per sequence ----------------
P_LAYER_SMALL = 7_077_888            # 768 x (2304 + 768 + 3072 + 3072)
P_HEAD_SMALL = 768 * 50_257
SURVEY_FLOP_PER_ENV_STEP = 3.09e9    # (6 + 6.5) tokens x 247.1 MFLOP: the survey's token counts for a trained policy
SURVEY_BYTES_PER_ENV_STEP = 10.0e6


def section_8d(env_steps_per_s, rows_per_env_step, sampled_per_env_step, att_bytes_per_env_step, peaks, batch):
    """The two whole-step fractions SURVEY.md §8(d) defines (decode-HBM and tensor), with the survey's constants AND with the work
    recomputed from what this run really did (decoder rows and sampled tokens per env-step read back from the device)."""
    flop = 2.0 * (12 * P_LAYER_SMALL * rows_per_env_step + P_HEAD_SMALL * sampled_per_env_step)
    w_bytes = 2.0 * (12 * P_LAYER_SMALL + P_HEAD_SMALL)
    passes_per_env_step = sampled_per_env_step               # one forward pass of the batch per sampled token
    byts = passes_per_env_step * w_bytes / batch + att_bytes_per_env_step
    return dict(
        tensor_frac_survey_constants=env_steps_per_s * SURVEY_FLOP_PER_ENV_STEP / (peaks["bf16_tflops_sustained"] * 1e12),
        hbm_frac_survey_constants=env_steps_per_s * SURVEY_BYTES_PER_ENV_STEP / (peaks["hbm_gbs"] * 1e9),
        survey_constants="3.09 GFLOP and 10.0 MB per env-step (12.5 decoder rows: the token counts of a trained policy)",
        decoder_rows_per_env_step=rows_per_env_step, sampled_tokens_per_env_step=sampled_per_env_step,
        flop_per_env_step=flop, bytes_per_env_step=byts,
        tensor_frac=env_steps_per_s * flop / (peaks["bf16_tflops_sustained"] * 1e12),
        hbm_frac=env_steps_per_s * byts / (peaks["hbm_gbs"] * 1e9),
        recomputed="flop = 2 x (84.9 M x decoder rows + 38.6 M x sampled tokens); bytes = weights once per pass / batch + K/V read by "
                   "attention (measured row and context counts of this run)")


def time_rollouts(eng, env_kind, seeds_dev, B, sampling, stream, ep_ret_dev, ep_len_dev, steps, first_seed_row, barrier):
    """K fused rollouts on the production path (graph replay), device timed; returns (ms, env-steps, mean return)."""
    import torch
    from lmrl_gym_b200 import _lib
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for k in range(steps):
        out = _lib.lmrl_rollout_out(ep_ret_dev[k].data_ptr(), ep_len_dev[k].data_ptr(), None, None, None, None, None, None, None)
        eng.rollout_dev(env_kind, seeds_dev[(first_seed_row + k) % seeds_dev.shape[0]].data_ptr(), B, sampling, out)
    ev1.record(stream)
    eng.sync()
    barrier()
    return ev0.elapsed_time(ev1), int(ep_len_dev[:steps].sum().item()), float(ep_ret_dev[:steps].mean().item())


def reduce_max_sum(ms, count, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    c = torch.tensor([count], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), int(c.item())


def bc_train_on_device(eng, tok, cfg, params, n_data, steps, bsize, lr, seed=0, log=None):
    """Bench variant 2c: behaviour-clone the Wordle policy on the device before the timed region — the reference pipeline
    wordle/misc/data_gen.py:16-27 (RandomMixturePolicy(prob_smart=0.5) episodes, here generated by the env kernels) ->
    wordle/bc/train_bc_gpt2.py (masked next-token cross-entropy on the action tokens; here the hand-written backward + AdamW).
    Leaves the trained weights in the engine's policy tower and returns (trainer, stats)."""
    import torch
    from lmrl_gym_b200 import _lib
    from lmrl_gym_b200 import ppo_train as T
    from lmrl_gym_b200.randomness import seed_generator
    from lmrl_gym_b200.wordle import scripted_rollout_records
    dev = torch.device("cuda", torch.cuda.current_device())
    B, Tn, K = min(4096, eng.max_envs), eng.max_turns, eng.max_new_tokens
    seeds = seed_generator(seed + 12345)
    records, t0 = [], time.perf_counter()
    for r in range((n_data + B - 1) // B):
        s_np = np.asarray([next(seeds) for _ in range(B)], dtype=np.uint64)
        seeds_dev = torch.from_numpy(s_np.view(np.int64)).to(dev)
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
        ep_ret, ep_len, ep_done = z(B, torch.float32), z(B, torch.int32), z(B, torch.uint8)
        rew, valid, fb = z((B, Tn), torch.float32), z((B, Tn), torch.uint8), z((B, Tn, 5), torch.uint8)
        aux, act, alen = z((B, Tn), torch.int32), z((B, Tn, K), torch.int32), z((B, Tn), torch.int32)
        out = _lib.lmrl_rollout_out(ep_ret.data_ptr(), ep_len.data_ptr(), ep_done.data_ptr(), rew.data_ptr(), valid.data_ptr(), fb.data_ptr(),
                                    aux.data_ptr(), act.data_ptr(), alen.data_ptr())
        eng.wordle_scripted_rollout_dev(seeds_dev.data_ptr(), B, 0.5, 1 + r, out)
        eng.sync()
        records += scripted_rollout_records(act[:, :, 0].cpu().numpy(), fb.cpu().numpy(), rew.cpu().numpy(), ep_len.cpu().numpy(), ep_done.cpu().numpy())
    records = records[:n_data]
    gen_s = time.perf_counter() - t0
    lin = dict(dense=dict(kernel=np.zeros((cfg.d_model, 1), np.float32), bias=np.zeros(1, np.float32)))
    max_length = 128
    rows = bsize * max_length
    trainer = T.B200PPOTrain(eng, params, lin, cfg, lr=lr, weight_decay=0.0, pad_token_id=tok.pad_token_id, backward="b200", max_rows=rows, max_scored_rows=rows)
    rng = np.random.default_rng(seed)
    import gc
    gc.collect()
    gc.freeze()          # 32 768 episode records (+ whatever earlier blocks left alive) stay out of the collector's passes during the 2 000 steps
    t0 = time.perf_counter()
    first = last = None
    warm = max(1, steps // 20)
    for step in range(steps):
        trainer.lr = lr * min(1.0, (step + 1) / warm) * (0.1 + 0.9 * 0.5 * (1 + np.cos(np.pi * step / steps)))     # linear warm-up, cosine decay to 10 %
        idx = rng.integers(0, len(records), size=bsize)
        batch = T.bc_batch_from_str_segments([records[i]["sequence"] for i in idx], tok, max_length, tok.pad_token_id)
        loss = trainer.bc_step(batch)
        first = loss if first is None else first
        last = loss
        if log and (step % 200 == 0 or step == steps - 1):
            log(f"bc step {step}: loss {loss:.4f}")
    train_s = time.perf_counter() - t0
    gc.unfreeze()
    return trainer, dict(episodes=len(records), data="device RandomMixturePolicy(prob_smart=0.5) episodes (wordle/misc/data_gen.py:16-27 equivalent)",
                         data_gen_s=gen_s, steps=steps, bsize=bsize, lr=lr, max_length=max_length, loss_first=first, loss_last=last, train_s=train_s,
                         s_per_step=train_s / max(steps, 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="concurrent episodes per GPU")
    ap.add_argument("--ref-batch", type=int, default=32, help="episodes per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="wordle", choices=["wordle", "maze_po_medium", "wordle_ilql"],
                    help="workload of the emitted line (default: the headline, BASELINE.json configs[1]); the other two also run as blocks of the default line")
    ap.add_argument("--no-profile", action="store_true", help="do not record per-launch CUDA events in the timed region")
    ap.add_argument("--skip", default="", help="comma list of optional blocks to skip: variant_2c,small_batch,maze_po_medium,wordle_ilql,ppo_round,letters,env_only,e2e_materialized")
    ap.add_argument("--bc-steps", type=int, default=2000, help="optimizer steps of the on-device BC training behind variant 2c")
    args = ap.parse_args()
    if args.workload == "maze_po_medium" and args.batch == 4096:
        args.batch = 1024
    if args.impl == "reference":
        return run_reference(args)
    skip = set(x for x in args.skip.split(",") if x)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def log(msg):
        if rank == 0:
            print(f"[bench] {msg}", file=sys.stderr, flush=True)

    line = run_workload(args, args.workload, rank, world, local, barrier, log, skip, headline=True)
    if args.workload == "wordle":
        # BASELINE.json configs[2] / the ILQL policy / configs[4] in front of the driver: same contract (barrier, device timing, max over
        # ranks), each on its own engine after the headline engine has been released
        for name in ("maze_po_medium", "wordle_ilql"):
            if name in skip:
                continue
            try:
                sub = run_workload(args, name, rank, world, local, barrier, log, skip | {"variant_2c", "small_batch", "letters", "env_only", "e2e_materialized"}, headline=False)
                if rank == 0:
                    line[name] = sub
            except Exception as ex:                     # a failing side block must not take the headline line with it
                if rank == 0:
                    line[name] = dict(error=repr(ex)[:500])
        if "ppo_round" not in skip:
            try:
                sub = run_ppo_round(rank, world, local, barrier, log)
                if rank == 0:
                    line["ppo_round"] = sub
            except Exception as ex:
                if rank == 0:
                    line["ppo_round"] = dict(error=repr(ex)[:500])
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline and args.workload == "wordle":
            line["cpu_baseline"] = cpu_baseline_subprocess(args.ref_batch)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_ppo_round(rank, world, local, barrier, log):
    """BASELINE.json configs[4]: one PPO round (rollout + PPO data + clipped-loss updates) on Wordle x GPT-2-small, 4096 rollouts per GPU;
    rollout / data / update seconds separately, and the gradient all-reduce (the only collective of the system) in ms and bytes."""
    import torch
    from lmrl_gym_b200 import ppo as P, ppo_train as T
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params, steer_to_letters
    from lmrl_gym_b200.randomness import seed_generator
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200.wordle import B200WordleEnv
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config.small()
    B, train_bsize, max_updates, rounds = 4096, 32, 16, 2
    eng = Engine(max_envs=B, device=local, max_input_length=256, max_new_tokens=6, max_turns=6, tokenizer=tok)
    params = steer_to_letters(init_gpt2_params(cfg, seed=0), tok)
    lin = dict(dense=dict(kernel=(np.random.default_rng(1).standard_normal((cfg.d_model, 1)) / 28).astype(np.float32), bias=np.zeros(1, np.float32)))
    env = B200WordleEnv(eng, "wordle_official.txt", bad_word_reward=-10.0)          # train_ppo_gpt2.py:211
    pol = B200GPT2Policy(eng, params, cfg, seed=rank)
    inf = P.B200PPOInference(eng, None, lin, initial_policy_params=params, config=cfg, pad_token_id=tok.pad_token_id)
    trainer = T.B200PPOTrain(eng, params, lin, cfg, lr=1e-5, pad_token_id=tok.pad_token_id, backward="b200", max_rows=train_bsize * 264, max_scored_rows=train_bsize * 264)
    seeds = seed_generator(0)
    for _ in range(rank * B * rounds):
        next(seeds)
    logs = []
    barrier()
    T.ppo_train_loop(env, pol, inf, trainer, P.AdaptiveKLController(0.001, 0.1, 10000), n_rounds=rounds, n_rollouts=B, rollout_bsize=B, epochs=1, train_bsize=train_bsize,
                     max_length=256 + 6, gamma=1.0, lam=0.95, seed_generator=seeds, max_steps_per_round=max_updates, on_round=lambda r, l: logs.append(l))
    barrier()
    ar = trainer.allreduce_stats()
    last = logs[-1]
    t = torch.tensor([last["rollout_s"], last["ppo_data_s"], last["update_s"]], dtype=torch.float64, device="cuda")
    c = torch.tensor([last["env_steps"]], dtype=torch.int64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    eng.close()
    upd = max(int(last["updates"]), 1)
    step_ms = 1000.0 * float(t[2]) / upd
    return dict(workload="BASELINE.json configs[4]: PPO round, Wordle x GPT-2-small (letters-steered random init), %d rollouts/GPU, train_bsize %d/GPU, %d optimizer steps "
                         "(capped), hand-written backward + lmrl_adamw_dev" % (B, train_bsize, upd),
                n_gpus=world, rollout_s=float(t[0]), ppo_data_s=float(t[1]), update_s=float(t[2]), update_ms_per_step=step_ms, updates=upd,
                env_steps_per_s_rollout=float(c[0]) / float(t[0]), kl_coef=last["kl_coef"], mean_kl=last["mean_kl"], loss=last["loss"],
                allreduce=dict(ms_per_step=ar["ms_per_step"], bytes_per_step=ar["bytes_per_step"], calls=ar["calls"],
                               share_of_update_step=(ar["ms_per_step"] / step_ms) if step_ms else None,
                               how="one summing NCCL all-reduce per flat fp32 gradient buffer (2 per step) after the backward pass, CUDA events on "
                                   "the issuing stream; not overlapped with the backward"),
                timing="max over ranks of each phase's host wall clock (every phase ends in a device synchronisation)")


def run_workload(args, workload, rank, world, local, barrier, log, skip, headline):
    import torch
    import torch.distributed as dist
    from lmrl_gym_b200 import _lib
    from lmrl_gym_b200.engine import Engine, ENV_MAZE, ENV_WORDLE
    from lmrl_gym_b200.environment import text_env_eval
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200.wordle import B200WordleEnv
    from lmrl_gym_b200.randomness import seed_generator as seed_stream   # the reference's seed_generator(0) stream (maze/env/randomness.py:35-41)

    maze, ilql = workload == "maze_po_medium", workload == "wordle_ilql"
    B = args.batch if headline else (1024 if maze else 4096)
    K, W = (args.steps, max(args.warmup, 0)) if headline else (max(1, min(args.steps, 2 if maze else 3)), 3)
    tok = ByteTokenizer.synthetic()
    cfg, params = build_policy_params(medium=maze)
    if maze:
        from lmrl_gym_b200.maze import B200MazeEnv
        eng = Engine(max_envs=B, device=local, max_input_length=512, max_new_tokens=3, max_turns=101, tokenizer=tok)
        env = B200MazeEnv(eng, "double_t_maze", "describe_observation_only_walls", "standard_reward", last_k=40, max_steps=100)
        env_kind, workload_name = ENV_MAZE, WORKLOAD_MAZE
    else:
        eng = Engine(max_envs=B, device=local, max_input_length=256, max_new_tokens=6, max_turns=6, tokenizer=tok)
        env = B200WordleEnv(eng, "wordle_official.txt")
        env_kind, workload_name = ENV_WORDLE, (WORKLOAD_ILQL if ilql else WORKLOAD)
    if ilql:
        from lmrl_gym_b200.gpt2 import B200GPT2ValuePolicy, init_gpt2_params, init_mlp_head_params
        value_params = (params, init_gpt2_params(cfg, seed=1), init_mlp_head_params(cfg, seed=2), init_mlp_head_params(cfg, seed=3))
        pol = B200GPT2ValuePolicy(eng, value_params, cfg, beta=16.0, seed=0)
    else:
        pol = B200GPT2Policy(eng, params, cfg, seed=0)
    sampling = pol.sampling()

    # per-episode seeds: the i-th value of ONE seed_generator(0) stream, sharded contiguously over ranks (no collective)
    n_steps_total = W + K
    gen = seed_stream(0)
    all_seeds = np.asarray([next(gen) for _ in range(n_steps_total * world * B)], dtype=np.uint64).reshape(n_steps_total, world, B)
    seeds_dev = torch.from_numpy(np.ascontiguousarray(all_seeds[:, rank]).view(np.int64)).cuda()           # resident in HBM before timing
    ep_len_dev = torch.zeros((K, B), dtype=torch.int32, device="cuda")
    ep_ret_dev = torch.zeros((K, B), dtype=torch.float32, device="cuda")
    stream = torch.cuda.ExternalStream(eng.stream_ptr, device=torch.device("cuda", local))

    def rollout(step_idx, out_k=None):
        out = None
        if out_k is not None:
            out = _lib.lmrl_rollout_out(ep_ret_dev[out_k].data_ptr(), ep_len_dev[out_k].data_ptr(), None, None, None, None, None, None, None)
        eng.rollout_dev(env_kind, seeds_dev[step_idx].data_ptr(), B, sampling, out)

    for w in range(W):
        rollout(w)
    eng.sync()
    barrier()
    # ---- timed region: exactly K steps on the production path (CUDA-graph replay, no per-launch events)
    eng.launch_count(reset=True)
    eng.profile_enable(False)
    clocks = ClockSampler(local)
    elapsed_ms, env_steps, mean_return = time_rollouts(eng, env_kind, seeds_dev, B, sampling, stream, ep_ret_dev, ep_len_dev, K, W, barrier)
    clock_info = clocks.stop()
    launches = eng.launch_count()
    kv_used, kv_total, kv_page = eng.kv_stats()
    # ---- the same K steps again with one CUDA event per launch boundary (eager launches): per-kernel times for the roofline
    prof, profiled_ms = None, None
    if not args.no_profile:
        eng.profile_enable(True)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for k in range(K):
            rollout(W + k, k)
        p1.record(stream)
        eng.sync()
        prof = eng.profile_read()
        profiled_ms = p0.elapsed_time(p1) / K
        eng.profile_enable(False)
    elapsed_ms, total_steps = reduce_max_sum(elapsed_ms, env_steps, world)
    _, total_launches = reduce_max_sum(0.0, launches, world)
    value = total_steps / (elapsed_ms / 1000.0)

    # ---- e2e: the public API with host buffers (seeds in, logs out), wall clock, max over ranks
    e2e_seeds = seed_stream(1000 + rank)
    text_env_eval(env, pol, n_rollouts=B, seed_generator=e2e_seeds, bsize=B, verbose=False)     # warm
    barrier()
    t0 = time.perf_counter()
    e2e_steps = 0
    d2h = h2d = 0
    inter = None
    for _ in range(K):
        inter, summ = text_env_eval(env, pol, n_rollouts=B, seed_generator=e2e_seeds, bsize=B, verbose=False)
        e2e_steps += int(round(float(summ["length"]["mean"]) * B))
        h2d, d2h = inter[0]._logs.h2d_bytes, inter[0]._logs.d2h_bytes
    barrier()
    e2e_s, e2e_total = reduce_max_sum(time.perf_counter() - t0, e2e_steps, world)
    e2e_value = e2e_total / e2e_s
    # ---- e2e with every InteractionTransition materialised (a caller that iterates the whole result, as the reference's drivers do when
    #      they print / pickle the raw interactions): the lazy episodes are expanded into Text / InteractionTransition objects
    e2e_mat = None
    if "e2e_materialized" not in skip and not maze:
        barrier()
        t0 = time.perf_counter()
        inter, summ = text_env_eval(env, pol, n_rollouts=B, seed_generator=e2e_seeds, bsize=B, verbose=False)
        n_tr = 0
        for ep in inter:
            for tr in ep:
                n_tr += 1 if tr.post_transition_history is not None else 0
        barrier()
        mat_s, mat_total = reduce_max_sum(time.perf_counter() - t0, n_tr, world)
        e2e_mat = dict(value=mat_total / mat_s, unit=UNIT, seconds_per_step=mat_s, transitions=mat_total,
                       note="text_env_eval + a full pass over every episode's InteractionTransition list (%d Python objects built on the host per step)" % (3 * n_tr))

    # ---- env kernels alone (variant 2b): device RandomMixturePolicy(prob_smart=0.5)-equivalent actions, no policy forward
    env_only = None
    if rank == 0 and not maze and "env_only" not in skip:
        out_env = _lib.lmrl_rollout_out(ep_ret_dev[0].data_ptr(), ep_len_dev[0].data_ptr(), torch.zeros(B, dtype=torch.uint8, device="cuda").data_ptr(),
                                        None, None, None, None, None, None)
        for _ in range(3):
            eng.wordle_scripted_rollout_dev(seeds_dev[0].data_ptr(), B, 0.5, 7, out_env)
        eng.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record(stream)
        for r in range(reps):
            eng.wordle_scripted_rollout_dev(seeds_dev[r % n_steps_total].data_ptr(), B, 0.5, 7 + r, out_env)
        e1.record(stream)
        eng.sync()
        env_steps_only = int(ep_len_dev[0].sum().item()) * reps          # last batch's count x reps (all batches are statistically alike)
        env_only = dict(value=env_steps_only / (e0.elapsed_time(e1) / 1000.0), unit=UNIT,
                        variant="2b: env kernels only, device RandomMixturePolicy(prob_smart=0.5) equivalent (every action a legal word)",
                        mean_episode_length=float(ep_len_dev[0].float().mean().item()), ms_per_batch=e0.elapsed_time(e1) / reps)

    def variant_run(tag):
        """K timed steps + one eager profiled step (row / token counts) of whatever weights the policy tower holds now."""
        for w in range(2):
            rollout(w % n_steps_total)
        eng.sync()
        ms, steps_v, ret_v = time_rollouts(eng, env_kind, seeds_dev, B, sampling, stream, ep_ret_dev, ep_len_dev, K, W, barrier)
        eng.profile_enable(True)
        rollout(W % n_steps_total, 0)
        vprof = eng.profile_read()
        eng.profile_enable(False)
        one_steps = int(ep_len_dev[0].sum().item())
        ms, steps_v = reduce_max_sum(ms, steps_v, world)
        att_b = sum(vprof["tags"].get(k, dict(work=0.0))["work"] for k in ("attention_decode", "attention_prefill"))
        return dict(value=steps_v / (ms / 1000.0), unit=UNIT, ms_per_step=ms / K, mean_return=ret_v, mean_episode_length=steps_v / (K * B * world),
                    decoder_rows_per_step=vprof["rows"], sampled_tokens_per_step=vprof["sampled"],
                    decoder_rows_per_env_step=vprof["rows"] / max(one_steps, 1), _att_bytes=att_b), vprof, one_steps

    # ---- same engine, policy steered to the token shape of a trained Wordle policy (~5 letter tokens + newline per action):
    #      prompts grow by ~7 tokens per turn instead of ~20 (SURVEY.md §8d's token counts); secondary number, not the headline
    letters_variant = None
    if not maze and not ilql and "letters" not in skip:
        from lmrl_gym_b200.gpt2 import steer_to_letters
        pol.set_params(steer_to_letters(params, tok))
        letters_variant, _, _ = variant_run("letters")
        letters_variant.pop("_att_bytes", None)
        letters_variant["variant"] = ("2d: random-init GPT-2-small with the LM head steered to letter tokens + newline (well-formed 5-letter-ish "
                                      "guesses, almost never vocabulary words): the token shape of a trained policy")
        pol.set_params(params)

    # ---- variant 2c: a policy that actually plays — behaviour-cloned ON THE DEVICE before its timed region
    variant_2c = None
    if not maze and not ilql and "variant_2c" not in skip:
        log(f"variant 2c: BC training, {args.bc_steps} steps")
        trainer, bc_stats = bc_train_on_device(eng, tok, cfg, params, n_data=32768, steps=args.bc_steps, bsize=64, lr=6e-4, log=log)
        variant_2c, vprof, one_steps = variant_run("2c")
        # valid-guess fraction and wins over one more batch through the public API (host logs)
        inter2, summ2 = text_env_eval(env, pol, n_rollouts=B, seed_generator=seed_stream(4242 + rank), bsize=B, verbose=False)
        lg = inter2[0]._logs
        n_act = int((lg.act_len > 0).sum())
        valid = int((lg.step_valid[lg.act_len > 0] != 0).sum())
        variant_2c.update(variant="2c: GPT-2-small behaviour-cloned on the device before the timed region (masked next-token cross-entropy on "
                                  "RandomMixturePolicy(prob_smart=0.5) episodes, hand-written backward + AdamW); target draw / transition / re-filter run "
                                  "inside the timed region for every valid guess",
                          valid_guess_fraction=valid / max(n_act, 1), mean_return_api=float(summ2["reward"]["mean"]),
                          episodes_won=float(np.mean(np.asarray(lg.ep_return) > -6.0)), bc_training=bc_stats)
        del trainer
        torch.cuda.empty_cache()
        pol.set_params(params)

    # ---- small batches (the reference's own eval batch sizes): decode tokens/s through policy.generate at B = 1 and 8
    small_batch = None
    if rank == 0 and headline and not maze and not ilql and "small_batch" not in skip:
        small_batch = small_batch_block(cfg, params, tok, local)

    line = None
    if rank == 0:
        peaks = load_peaks()
        if prof is None:
            prof = dict(ms=dict(gemm=0.0, attention=0.0, elementwise=0.0, env=0.0), work=dict(gemm=0.0), launches=dict(gemm=0), tags={}, rows=0, sampled=0)
        gemm_ms, gemm_flops = prof["ms"]["gemm"], prof["work"]["gemm"]
        att = prof["tags"].get("attention_decode", dict(ms=0.0, work=0.0))
        attp = prof["tags"].get("attention_prefill", dict(ms=0.0, work=0.0))
        att_gbs = att["work"] / (att["ms"] / 1000.0) / 1e9 if att["ms"] > 0 else 0.0
        achieved = gemm_flops / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
        peak = peaks["bf16_tflops_sustained"]
        total_kernel_ms = sum(prof["ms"].values())
        fam = load_traffic() or {}
        gemm_traffic = fam.get("gemm_bf16_tcgen05_kernel", {}).get("dram_bytes_per_launch")
        att_traffic = (fam.get("attention_hpw_kernel") or fam.get("attention_kernel") or {}).get("dram_bytes_per_launch")
        local_steps = max(env_steps, 1)
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=elapsed_ms / K, higher_is_better=True,
            scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
            config=dict(workload=workload_name, batch_per_gpu=B, global_batch=B * world, parallelism=f"dp{world} (episodes sharded, no collective)",
                        l2_policy="inputs larger than L2 (KV page pool + activations %.1f GB per GPU)" % (eng.device_bytes / 1e9), tokenizer=tok.name,
                        policy_variant=("random-init GPT-2-medium (every action is an unknown string: 101 turns, the history collapses to the observation)" if maze else
                                        "ILQL value-perturbed policy, beta = 16: random-init pi_beta tower + value tower + two MLP Q heads (eval_ilql_gpt2.py:47)" if ilql else
                                        "2a: random-init GPT-2-small (actions are malformed guesses; every episode runs 6 turns)"),
                        kv_cache="paged: %d of %d pages of %d positions in use after a step" % (kv_used, kv_total, kv_page),
                        timed_region="K fused rollouts replayed from a CUDA graph, no per-launch events" if eng.max_turns <= 8 else
                                     "K fused rollouts launched eagerly (rollouts longer than 8 turns poll the device for early exit)",
                        roofline_pass=("the same K steps re-run eagerly with one CUDA event per launch boundary: %.2f ms/step" % profiled_ms) if profiled_ms else None),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, api="lmrl_gym_b200.text_env_eval"),
            gpu_launches=total_launches,
            clocks=clock_info,
            roofline=dict(bound="tensor", kernel="gemm_bf16_tcgen05_kernel + gemm_bf16_tcgen05_2cta_kernel (tcgen05.mma UMMA 128x256x16 / cta_group::2 256x256x16 on the plain prefill GEMMs, TMA + TMEM, all epilogues)", achieved=achieved, peak=peak,
                          unit="TFLOP/s", frac=achieved / peak if peak else None, peak_source=f"{peaks['source']} bf16_tflops_sustained", traffic=gemm_traffic,
                          traffic_source="profiles/r02_traffic.json (r01 when absent): mean dram__bytes_read.sum + dram__bytes_write.sum per launch over every GEMM launch of one ncu-profiled step" if gemm_traffic else None,
                          flops_per_step=gemm_flops / K, launches_per_step=prof["launches"]["gemm"] / K,
                          share_of_kernel_time=gemm_ms / total_kernel_ms if total_kernel_ms else None,
                          # the HBM-bound batched GEMV of the path (north_star's "decode GEMV"): KV-cache attention of the decode passes
                          decode_attention_hbm=dict(kernel="attention_hpw_kernel<4> (decode passes, paged KV, 4 heads of a slot per warp)", achieved=att_gbs, peak=peaks["hbm_gbs"], unit="GB/s",
                                                    frac=att_gbs / peaks["hbm_gbs"], bytes_per_step=att["work"] / K, traffic=att_traffic,
                                                    peak_source=f"{peaks['source']} hbm_gbs",
                                                    algorithmic_bytes="every cached K and V position of the live envs once (bf16) + q / out rows")),
            kernel_ms_per_step={k: v / K for k, v in prof["ms"].items()},
            gemm_attention_split={k: dict(ms_per_step=round(v["ms"] / K, 3),
                                          **({"gbs": round(v["work"] / (v["ms"] / 1000.0) / 1e9, 1)} if k.startswith("attention") else
                                             {"tflops": round(v["work"] / (v["ms"] / 1000.0) / 1e12, 1)}) if v["ms"] > 0 and v["work"] > 0 else {})
                                  for k, v in prof["tags"].items()},
            decoder_rows_per_step=prof["rows"] / K, sampled_tokens_per_step=prof["sampled"] / K,
            decoder_rows_per_env_step=prof["rows"] / K / (local_steps / K), sampled_tokens_per_env_step=prof["sampled"] / K / (local_steps / K),
            mean_return=mean_return,
        )
        if not maze and not ilql and prof["rows"]:
            per_gpu = value / world
            line["section_8d"] = section_8d(per_gpu, prof["rows"] / local_steps, prof["sampled"] / local_steps,
                                            (att["work"] + attp["work"]) / local_steps, peaks, B)
        if e2e_mat is not None:
            line["e2e_materialized"] = e2e_mat
        if env_only is not None:
            line["env_only"] = env_only
        if letters_variant is not None:
            line["letters_variant"] = letters_variant
        if variant_2c is not None:
            if variant_2c.get("decoder_rows_per_env_step"):
                steps_1 = max(variant_2c["decoder_rows_per_step"] / variant_2c["decoder_rows_per_env_step"], 1)      # env-steps of the profiled batch
                variant_2c["section_8d"] = section_8d(variant_2c["value"] / world, variant_2c["decoder_rows_per_env_step"],
                                                      variant_2c["sampled_tokens_per_step"] / steps_1, variant_2c.pop("_att_bytes", 0.0) / steps_1, peaks, B)
            line["variant_2c"] = variant_2c
        if small_batch is not None:
            line["small_batch"] = small_batch
    eng.close()
    del eng
    torch.cuda.empty_cache()
    return line


def small_batch_block(cfg, params, tok, local):
    """Decode tokens/s of `policy.generate` at the reference's own eval batch sizes (1 and 8 episodes; eval_bc_gpt2.py:38-39): the
    weight-streaming GEMV path against the HBM roofline of one decode pass (every weight byte once), and the same batches through
    128-row tensor-core tiles for comparison."""
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    peaks = load_peaks()
    d, ff, L, V = cfg.d_model, cfg.d_ff, cfg.n_layer, cfg.n_vocab
    weight_bytes = 2.0 * (L * (3 * d * d + d * d + 2 * d * ff) + V * d)
    n_new, reps, out = 65, 3, {}
    for B in (1, 8):
        eng = Engine(max_envs=B, device=local, max_input_length=32, max_new_tokens=n_new, tokenizer=tok)
        pol = B200GPT2Policy(eng, params, cfg, seed=0)
        rng = np.random.default_rng(B)
        prompts = [rng.integers(0, V, size=8).tolist() for _ in range(B)]
        row = {}
        for name, rows in (("gemv", 8), ("tensor_core_tiles", 0)):
            eng.set_option("gemv_max_rows", rows)
            for _ in range(3):                                      # eager, capture of the decode pass, replay
                pol.generate_tokens(prompts)
            eng.sync()
            t0 = time.perf_counter()
            n_tok = 0
            for _ in range(reps):
                n_tok += sum(len(g) for g in pol.generate_tokens(prompts))
            dt = time.perf_counter() - t0
            gbs = weight_bytes * reps * n_new / dt / 1e9
            row[name] = dict(tokens_per_s=n_tok / dt, us_per_pass=1e6 * dt / (reps * n_new), weight_gbs=gbs, hbm_frac=gbs / peaks["hbm_gbs"])
        out[f"batch_{B}"] = row
        eng.close()
    out["note"] = ("policy.generate_tokens, 8-token prompts, 65 new tokens, host wall clock (prompt upload, prefill pass and token download included); "
                   "weight_gbs = bf16 weight bytes of one decode pass (%.1f MB) x passes / time; the B = 1 HBM roofline is %.0f tokens/s"
                   % (weight_bytes / 1e6, peaks["hbm_gbs"] * 1e9 / weight_bytes))
    return out


if __name__ == "__main__":
    main()
