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
WORKLOAD_MAZE = ("double_t_maze partially observed (describe_observation_only_walls, last_k=40, max_steps=100) x GPT-2-medium"
                 "(random init, seed 0), max_new_tokens=3, max_input_length=512")


def load_traffic():
    """DRAM bytes per launch of the kernel families, from the committed ncu launch list of this same command
    (profiles/r01_traffic.json, written by scripts/ncu_summarize.py traffic); None when no capture has been committed."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="concurrent episodes per GPU")
    ap.add_argument("--ref-batch", type=int, default=32, help="episodes per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="wordle", choices=["wordle", "maze_po_medium"])
    ap.add_argument("--no-profile", action="store_true", help="do not record per-launch CUDA events in the timed region")
    args = ap.parse_args()
    if args.workload == "maze_po_medium" and args.batch == 4096:
        args.batch = 1024
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from lmrl_gym_b200 import _lib
    from lmrl_gym_b200.engine import Engine, ENV_MAZE, ENV_WORDLE
    from lmrl_gym_b200.environment import text_env_eval
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200.wordle import B200WordleEnv
    from lmrl_gym_b200.randomness import seed_generator as seed_stream   # the reference's seed_generator(0) stream (maze/env/randomness.py:35-41)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, K, W = args.batch, args.steps, max(args.warmup, 0)
    tok = ByteTokenizer.synthetic()
    maze = args.workload == "maze_po_medium"
    cfg, params = build_policy_params(medium=maze)
    if maze:
        from lmrl_gym_b200.maze import B200MazeEnv
        eng = Engine(max_envs=B, device=local, max_input_length=512, max_new_tokens=3, max_turns=101, tokenizer=tok)
        env = B200MazeEnv(eng, "double_t_maze", "describe_observation_only_walls", "standard_reward", last_k=40, max_steps=100)
        env_kind, workload = ENV_MAZE, WORKLOAD_MAZE
    else:
        eng = Engine(max_envs=B, device=local, max_input_length=256, max_new_tokens=6, max_turns=6, tokenizer=tok)
        env = B200WordleEnv(eng, "wordle_official.txt")
        env_kind, workload = ENV_WORDLE, WORKLOAD
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(W):
        rollout(w)
    eng.sync()
    barrier()
    # ---- timed region: exactly K steps on the production path (CUDA-graph replay, no per-launch events)
    eng.launch_count(reset=True)
    eng.profile_enable(False)
    clocks = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for k in range(K):
        rollout(W + k, k)
    ev1.record(stream)
    eng.sync()
    barrier()
    clock_info = clocks.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count()
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
    env_steps = int(ep_len_dev.sum().item())
    mean_return = float(ep_ret_dev.mean().item())
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    s = torch.tensor([env_steps, launches], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    elapsed_ms = float(t.item())
    total_steps, total_launches = int(s[0].item()), int(s[1].item())
    value = total_steps / (elapsed_ms / 1000.0)

    # ---- e2e: the public API with host buffers (seeds in, logs out), wall clock, max over ranks
    e2e_seeds = seed_stream(1000 + rank)
    text_env_eval(env, pol, n_rollouts=B, seed_generator=e2e_seeds, bsize=B, verbose=False)     # warm
    barrier()
    t0 = time.perf_counter()
    e2e_steps = 0
    d2h = h2d = 0
    for _ in range(K):
        inter, summ = text_env_eval(env, pol, n_rollouts=B, seed_generator=e2e_seeds, bsize=B, verbose=False)
        e2e_steps += int(round(float(summ["length"]["mean"]) * B))
        h2d, d2h = inter[0]._logs.h2d_bytes, inter[0]._logs.d2h_bytes
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    s = torch.tensor([e2e_steps], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    e2e_value = float(s.item()) / float(t.item())

    # ---- env kernels alone (variant 2b): device RandomMixturePolicy(prob_smart=0.5)-equivalent actions, no policy forward
    env_only = None
    if rank == 0 and not maze:
        out_env = _lib.lmrl_rollout_out(ep_ret_dev[0].data_ptr(), ep_len_dev[0].data_ptr(), torch.zeros(B, dtype=torch.uint8, device="cuda").data_ptr(),
                                        None, None, None, None, None, None)
        for _ in range(3):
            eng.wordle_scripted_rollout_dev(seeds_dev[0].data_ptr(), B, 0.5, 7, out_env)
        eng.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps, env_steps_only = 20, 0
        e0.record(stream)
        for r in range(reps):
            eng.wordle_scripted_rollout_dev(seeds_dev[r % n_steps_total].data_ptr(), B, 0.5, 7 + r, out_env)
        e1.record(stream)
        eng.sync()
        env_steps_only = int(ep_len_dev[0].sum().item()) * reps          # last batch's count x reps (all batches are statistically alike)
        env_only = dict(value=env_steps_only / (e0.elapsed_time(e1) / 1000.0), unit=UNIT,
                        variant="2b: env kernels only, device RandomMixturePolicy(prob_smart=0.5) equivalent (every action a legal word)",
                        mean_episode_length=float(ep_len_dev[0].float().mean().item()), ms_per_batch=e0.elapsed_time(e1) / reps)

    # ---- same engine, policy steered to the token shape of a trained Wordle policy (~5 letter tokens + newline per action):
    #      prompts grow by ~7 tokens per turn instead of ~20 (SURVEY.md §8d's token counts); secondary number, not the headline
    letters_variant = None
    if not maze:
        from lmrl_gym_b200.gpt2 import steer_to_letters
        pol.set_params(steer_to_letters(params, tok))
        for w in range(2):
            rollout(w % n_steps_total)
        eng.sync()
        barrier()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record(stream)
        for k in range(K):
            rollout(W + k, k)
        l1.record(stream)
        eng.sync()
        barrier()
        letters_return = float(ep_ret_dev.mean().item())
        eng.profile_enable(True)                      # one more step, eagerly, only for the row / token counts
        rollout(W % n_steps_total, 0)
        lprof = eng.profile_read()
        eng.profile_enable(False)
        lt = torch.tensor([l0.elapsed_time(l1)], dtype=torch.float64, device="cuda")
        ls = torch.tensor([int(ep_len_dev.sum().item())], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(lt, op=dist.ReduceOp.MAX)
            dist.all_reduce(ls, op=dist.ReduceOp.SUM)
        letters_variant = dict(value=float(ls.item()) / (float(lt.item()) / 1000.0), unit=UNIT, ms_per_step=float(lt.item()) / K,
                               variant="2d: random-init GPT-2-small with the LM head steered to letter tokens + newline (well-formed 5-letter-ish "
                                       "guesses, almost never vocabulary words): the token shape of a trained policy",
                               decoder_rows_per_step=lprof["rows"], sampled_tokens_per_step=lprof["sampled"], mean_return=letters_return)
        pol.set_params(params)

    if rank == 0:
        peaks = load_peaks()
        if prof is None:
            prof = dict(ms=dict(gemm=0.0, attention=0.0, elementwise=0.0, env=0.0), work=dict(gemm=0.0), launches=dict(gemm=0), tags={}, rows=0, sampled=0)
        gemm_ms, gemm_flops = prof["ms"]["gemm"], prof["work"]["gemm"]
        att = prof["tags"].get("attention_decode", dict(ms=0.0, work=0.0))
        att_gbs = att["work"] / (att["ms"] / 1000.0) / 1e9 if att["ms"] > 0 else 0.0
        achieved = gemm_flops / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
        peak = peaks["bf16_tflops_sustained"]
        total_kernel_ms = sum(prof["ms"].values())
        fam = load_traffic() or {}
        gemm_traffic = fam.get("gemm_bf16_tcgen05_kernel", {}).get("dram_bytes_per_launch")
        att_traffic = fam.get("attention_kernel", {}).get("dram_bytes_per_launch")
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=elapsed_ms / K, higher_is_better=True,
            scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
            config=dict(workload=workload, batch_per_gpu=B, global_batch=B * world, parallelism=f"dp{world} (episodes sharded, no collective)",
                        l2_policy="inputs larger than L2 (KV cache %.1f GB per GPU)" % (eng.device_bytes / 1e9), tokenizer=tok.name,
                        policy_variant=("random-init GPT-2-medium (every action is an unknown string: 101 turns, the history collapses to the observation)" if maze else
                                        "2a: random-init GPT-2-small (actions are malformed guesses; every episode runs 6 turns)"),
                        timed_region="K fused rollouts replayed from a CUDA graph, no per-launch events",
                        roofline_pass=("the same K steps re-run eagerly with one CUDA event per launch boundary: %.2f ms/step" % profiled_ms) if prof else None),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, api="lmrl_gym_b200.text_env_eval"),
            gpu_launches=total_launches,
            clocks=clock_info,
            roofline=dict(bound="tensor", kernel="gemm_bf16_tcgen05_kernel (tcgen05.mma UMMA 128x256x16 + TMA + TMEM, all epilogues)", achieved=achieved, peak=peak,
                          unit="TFLOP/s", frac=achieved / peak if peak else None, peak_source=f"{peaks['source']} bf16_tflops_sustained", traffic=gemm_traffic,
                          traffic_source="profiles/r01_traffic.json: mean dram__bytes_read.sum + dram__bytes_write.sum per launch over every GEMM launch of one ncu-profiled step" if gemm_traffic else None,
                          flops_per_step=gemm_flops / K, launches_per_step=prof["launches"]["gemm"] / K,
                          share_of_kernel_time=gemm_ms / total_kernel_ms if total_kernel_ms else None,
                          # the HBM-bound batched GEMV of the path (north_star's "decode GEMV"): KV-cache attention of the decode passes
                          decode_attention_hbm=dict(kernel="attention_kernel (decode passes)", achieved=att_gbs, peak=peaks["hbm_gbs"], unit="GB/s",
                                                    frac=att_gbs / peaks["hbm_gbs"], bytes_per_step=att["work"] / K, traffic=att_traffic,
                                                    peak_source=f"{peaks['source']} hbm_gbs",
                                                    algorithmic_bytes="every cached K and V position of the live envs once (bf16) + q / out rows")),
            kernel_ms_per_step={k: v / K for k, v in prof["ms"].items()},
            gemm_attention_split={k: dict(ms_per_step=round(v["ms"] / K, 3),
                                          **({"gbs": round(v["work"] / (v["ms"] / 1000.0) / 1e9, 1)} if k.startswith("attention") else
                                             {"tflops": round(v["work"] / (v["ms"] / 1000.0) / 1e12, 1)}) if v["ms"] > 0 and v["work"] > 0 else {})
                                  for k, v in prof["tags"].items()},
            decoder_rows_per_step=prof["rows"] / K, sampled_tokens_per_step=prof["sampled"] / K,
            mean_return=mean_return,
        )
        if env_only is not None:
            line["env_only"] = env_only
        if letters_variant is not None:
            line["letters_variant"] = letters_variant
        if world == 1 and not args.no_cpu_baseline and not maze:
            line["cpu_baseline"] = cpu_baseline_subprocess(args.ref_batch)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
