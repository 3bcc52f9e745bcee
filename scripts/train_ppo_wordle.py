#!/usr/bin/env python
"""Online PPO on Wordle with the device engine — the driver a user of llm_rl_scripts/wordle/ppo/train_ppo_gpt2.py switches to.
Rollout (fused device loop), trajectory → token chains (on the device), PPO data (log-probs of policy and initial policy, values,
KL-shaped rewards, GAE, whitening), the clipped-loss gradient step (hand-written backward, optional BC term) and AdamW all run on the
engine; under torchrun every rank collects its own rollouts and the gradients are all-reduced over NCCL.

    python scripts/train_ppo_wordle.py config gpt2 none wordle_official.txt --n-rounds 2 --n-rollouts 4096 --rollout-bsize 4096 \
        --max-input-length 256 --max-output-length 6 --max-steps 32

Flags keep the reference's names and defaults.  Not carried over (out of this repository's scope): wandb, the checkpoint / evaluation
schedules (``save_at_end`` writes the final weights as an ``.npz``), gradient accumulation and the mesh flags.  ``bc_data_path``: a
JSON-lines file of ``{"sequence": [[text, mask], ...]}`` records as the reference's (train_ppo_gpt2.py:135-144), or ``none``.
"""
import itertools
import json
import os
from typing import Optional

import numpy as np
import tyro

from _driver_common import load_tokenizer, to_floats


def _bc_batches(path, tokenizer, bsize, max_length, pad_id, rank, world):
    from lmrl_gym_b200.ppo_train import bc_batch_from_str_segments
    with open(path) as f:
        items = [json.loads(line)["sequence"] for line in f if line.strip()]
    items = items[rank::world] or items
    for lo in itertools.cycle(range(0, max(len(items) - bsize + 1, 1), bsize)):
        yield bc_batch_from_str_segments(items[lo:lo + bsize], tokenizer, max_length, pad_id)


def main(
    model_load_mode: str,
    model_load_path: str,
    bc_data_path: str,
    vocab_file: str,

    /,

    outputs_path: Optional[str] = None,
    tokenizer_path: Optional[str] = None,

    n_rounds: int = 1,
    epochs: int = 1,
    max_steps: Optional[int] = None,

    lr: float = 1e-5,
    weight_decay: float = 0.0,

    train_bsize: int = 32,
    train_bc_bsize: Optional[int] = None,
    rollout_bsize: int = 32,
    n_rollouts: int = 128,

    bf16_activations: bool = True,

    max_input_length: int = 512,
    max_output_length: int = 512,

    save_at_end: bool = False,

    policy_do_sample: bool = True,
    policy_num_beams: int = 1,
    policy_temperature: Optional[float] = None,
    policy_top_p: Optional[float] = None,
    policy_top_k: Optional[int] = None,

    gamma: float = 1.0,
    lam: float = 0.95,
    use_advantage_whitening: bool = True,

    init_kl_coef: float = 0.001,
    kl_target: Optional[float] = None,
    kl_horizon: Optional[int] = None,

    cliprange_value: float = 0.2,
    cliprange: float = 0.2,
    value_loss_coef: float = 1.0,
    bc_loss_weight: float = 1.0,

    force_pad_embeddings: bool = False,
    seed: Optional[int] = None,
    backward: str = "b200",
):
    print(locals())
    assert policy_num_beams == 1 and bf16_activations
    use_adaptive_kl = kl_target is not None and kl_horizon is not None
    if not use_adaptive_kl:
        assert kl_target is None and kl_horizon is None
    from lmrl_gym_b200 import distributed as D, ppo as P, ppo_train as T
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    from lmrl_gym_b200.flax_checkpoint import save_reference_checkpoint
    from lmrl_gym_b200.load import load_params, save_params
    from lmrl_gym_b200.randomness import seed_generator
    from lmrl_gym_b200.wordle import B200WordleEnv

    rank, world, device = D.init_from_env()
    tokenizer = load_tokenizer(tokenizer_path)
    pad_id = tokenizer.pad_token_id
    params, cfg = load_params(model_load_mode, model_load_path, seed=2, force_pad_embeddings=force_pad_embeddings)
    # the teacher-forced passes of the update see whole trajectories (up to max_input_length + max_output_length tokens, :346-355): the
    # engine's row capacity covers them, within the model's position table
    engine_in = min(max(max_input_length, max_input_length + max_output_length), cfg.n_pos - max_output_length)
    max_length = min(max_input_length + max_output_length, engine_in)
    engine = Engine(max_envs=max(rollout_bsize, train_bsize), device=device, max_input_length=engine_in, max_new_tokens=max_output_length,
                    tokenizer=tokenizer)
    # LinearHeadConfig(input_dim=n_embd, output_dim=1, initializer_range=0.0, bias_init=-4.1)  (train_ppo_gpt2.py:248-259)
    value_head = dict(dense=dict(kernel=np.zeros((cfg.d_model, 1), np.float32), bias=np.asarray([-4.1], np.float32)))
    env = B200WordleEnv(engine, vocab=vocab_file, require_words_in_vocab=True, bad_word_reward=-10.0)      # :211
    policy = B200GPT2Policy(engine, params, cfg, do_sample=policy_do_sample, temperature=policy_temperature or 1.0, top_k=policy_top_k,
                            top_p=policy_top_p, seed=rank)
    inference = P.B200PPOInference(engine, None, value_head, initial_policy_params=params, config=cfg, pad_token_id=pad_id)
    have_bc = bc_data_path.lower() != "none" and bc_loss_weight != 0.0
    bc_bsize = train_bc_bsize or train_bsize
    rows = max(train_bsize, bc_bsize if have_bc else 0) * max_length
    trainer = T.B200PPOTrain(engine, params, value_head, cfg, lr=lr, weight_decay=weight_decay, cliprange_value=cliprange_value, cliprange=cliprange,
                             value_loss_coef=value_loss_coef, pad_token_id=pad_id, backward=backward, max_rows=rows, max_scored_rows=rows,
                             bc_loss_weight=bc_loss_weight if have_bc else 0.0)
    kl_controller = P.AdaptiveKLController(init_kl_coef, kl_target, kl_horizon) if use_adaptive_kl else P.FixedKLController(init_kl_coef)
    seeds = None
    if seed is not None:                                   # every rank takes its own contiguous blocks of the ONE seed stream
        stream = seed_generator(seed)
        seeds = (s for i, s in enumerate(stream) if (i // n_rollouts) % world == rank)
    bc = _bc_batches(bc_data_path, tokenizer, bc_bsize, max_length, pad_id, rank, world) if have_bc else None

    def on_round(r, log):
        if rank == 0:
            print(json.dumps(to_floats({k: v for k, v in log.items() if v is not None})))

    history = T.ppo_train_loop(env, policy, inference, trainer, kl_controller, n_rounds=n_rounds, n_rollouts=n_rollouts, rollout_bsize=rollout_bsize,
                               epochs=epochs, train_bsize=train_bsize, max_length=max_length, gamma=gamma, lam=lam,
                               use_advantage_whitening=use_advantage_whitening, seed_generator=seeds, max_steps_per_round=max_steps,
                               on_round=on_round, bc_batches=bc)
    if rank == 0 and outputs_path is not None:
        os.makedirs(outputs_path, exist_ok=True)
        with open(os.path.join(outputs_path, "rounds.json"), "w") as f:
            json.dump([to_floats({k: v for k, v in h.items() if v is not None}) for h in history], f)
        if save_at_end:
            save_params(os.path.join(outputs_path, "policy_params.npz"), trainer.params_numpy(), cfg)
            save_reference_checkpoint(os.path.join(outputs_path, "policy"), trainer.params_numpy(), cfg)      # the reference's layout too (policy/params.msgpack + config.json)
            np.savez(os.path.join(outputs_path, "value_head.npz"), kernel=trainer.vh_w.detach().cpu().numpy(), bias=trainer.vh_b.detach().cpu().numpy())
    engine.close()
    D.finalize()
    return history


if __name__ == "__main__":
    tyro.cli(main)
