"""The PPO gradient step (lmrl_gym_b200/ppo_train.py::B200PPOTrain — the torch-autograd LIBRARY path, see its docstring) on the CPU:
its forward equals the oracle GPT-2, its loss equals the reference-pinned oracle ppo_loss, a step lowers the loss, and with two gloo
ranks the all-reduced update keeps the replicas identical (the role of the NCCL gradient all-reduce in BASELINE.json configs[4])."""
import os
import sys

import numpy as np
import pytest
import torch

from lmrl_gym_b200.gpt2 import GPT2Config, init_gpt2_params
from lmrl_gym_b200 import ppo_train as T
from lmrl_gym_b200 import ppo as P
from oracle import gpt2 as og, ppo as O

CFG = GPT2Config(n_layer=2, n_head=2, d_model=64, d_ff=128, n_vocab=97, n_vocab_rows=98, n_pos=32)
PAD = 97


def make_batch(rng, b=6, t=12):
    ids = np.full((b, t), PAD, np.int32)
    sta = np.zeros((b, t - 1), bool)
    for i in range(b):
        n = int(rng.integers(4, t + 1))
        ids[i, :n] = rng.integers(0, CFG.n_vocab, n)
        sta[i, 1:n - 1] = rng.random(n - 2) < 0.6
    sta[0, 1] = True
    f = lambda: rng.standard_normal((b, t - 1)).astype(np.float32)  # noqa: E731
    return dict(input_ids=ids, should_take_action=sta, old_logprobs=-np.abs(f()) - 3.0, old_values=f(), old_advantages=f(), old_returns=f())


def make_trainer(seed=0, lr=1e-3):
    params = init_gpt2_params(CFG, seed=seed)
    lin = dict(dense=dict(kernel=(np.random.default_rng(seed + 1).standard_normal((CFG.d_model, 1)) / 8).astype(np.float32), bias=np.asarray([0.1], np.float32)))
    return T.B200PPOTrain(None, params, lin, CFG, lr=lr, pad_token_id=PAD, activation_dtype="f32"), params, lin


def test_forward_and_loss_match_oracle():
    tr, params, lin = make_trainer()
    batch = make_batch(np.random.default_rng(0))
    ids = torch.as_tensor(batch["input_ids"]).long()
    h = tr._hidden(ids, ids != PAD).detach().numpy()
    lp = np.zeros((ids.shape[0], ids.shape[1] - 1), np.float32)
    vals = np.zeros_like(lp)
    for i, row in enumerate(batch["input_ids"]):
        n = int((row != PAD).sum())
        s = row[:n].tolist()
        ho = og.forward_hidden(params, s, CFG.n_head, weights_bf16=False).numpy()
        np.testing.assert_allclose(h[i, :n], ho, rtol=2e-4, atol=2e-4)
        lp[i, :n - 1] = og.token_logprobs_all(params, s, CFG.n_head, CFG.n_vocab, weights_bf16=False)
        vals[i, :n - 1] = og.linear_value_forward(params, lin, s, CFG.n_head, weights_bf16=False)[:-1]
    loss, logs = tr.loss(batch)
    mask = (batch["input_ids"] != PAD)[:, 1:].astype(np.float32)
    o_loss, o_logs = O.ppo_loss(mask, lp, vals, batch["should_take_action"], batch["old_logprobs"], batch["old_values"], batch["old_advantages"], batch["old_returns"],
                                cliprange_value=0.2, cliprange=0.2, value_loss_coef=1.0)
    assert abs(float(loss) - float(o_loss)) <= 2e-4 * max(1.0, abs(float(o_loss)))
    for k in ("policy_loss", "value_loss", "approx_kl", "ratio"):
        assert abs(float(logs[k]) - o_logs[k]) <= 5e-4 * max(1.0, abs(o_logs[k])), k


def test_steps_lower_the_loss_and_respect_weight_decay_mask():
    tr, _, _ = make_trainer(lr=3e-3)
    batch = make_batch(np.random.default_rng(1))
    first = tr.step(batch)[0]
    for _ in range(15):
        last = tr.step(batch)[0]
    assert last < first
    groups = tr.opt.param_groups
    assert groups[1]["weight_decay"] == 0.0 and len(tr._names[1]) == 2 + 8 * CFG.n_layer + 1      # LN scales/biases + all biases + value bias
    assert set(tr._names[0]) == {"wte", "wpe", "vh_w"} | {f"h{i}.{k}" for i in range(CFG.n_layer) for k in ("attn_w", "proj_w", "fc_w", "fc2_w")}


def test_episodes_to_chains_and_block():
    from lmrl_gym_b200.environment import Text, InteractionTransition

    class Tok:
        @staticmethod
        def encode(s):
            return [ord(c) % 90 for c in s]
    h0 = (Text("Wordle:\n", False),)
    ep = []
    h = h0
    for k, r in enumerate((-1.0, -1.0, 0.0)):
        pa = h + (Text("c r a n e\n", True),)
        pt = pa + (Text("b y g b b\n", False),)
        ep.append(InteractionTransition(pre_action_history=h, post_action_history=pa, post_transition_history=pt, reward=r, done=k == 2))
        h = pt
    chains = T.episodes_to_chains([ep], Tok, max_length=200, gamma=0.5)
    tt = chains[0].text_trajectory
    assert tt.reward == (0.0, -1.0, 0.0, -1.0, 0.0, 0.0, 0.0) and tt.done and chains[0].next is None
    short = T.episodes_to_chains([ep], Tok, max_length=60, gamma=0.5)[0].text_trajectory          # 68 tokens -> drop the last turn
    assert len(short.text_history) == 5 and short.reward == (0.0, -1.0, 0.0, -1.0, 0.0) and not short.done
    assert T.episodes_to_chains([ep], Tok, max_length=20, gamma=0.5) == []
    d = P.PPOData(np.arange(5, dtype=np.int32), np.ones(4, bool), np.ones(4, np.float32), np.ones(4, np.float32), np.ones(4, np.float32), np.ones(4, np.float32))
    blk = T.block_ppo_data([d, d._replace(input_ids=np.arange(3, dtype=np.int32), should_take_action=np.ones(2, bool))], PAD, 8)
    assert blk["input_ids"].shape == (2, 8) and blk["old_returns"].shape == (2, 7) and blk["input_ids"][1, 3] == PAD and not blk["should_take_action"][1, 2]


def _rank_main(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tr, _, _ = make_trainer(lr=1e-3)
    for s in range(3):
        tr.step(make_batch(np.random.default_rng(100 + 10 * s + rank)))          # different data per rank
    flat = torch.cat([p.detach().reshape(-1) for _, p in tr.named_parameters()]).numpy()
    np.save(os.path.join(out_dir, f"r{rank}.npy"), flat)
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_keeps_replicas_identical(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    np.testing.assert_array_equal(a, b)
    single, _, _ = make_trainer(lr=1e-3)
    for s in range(3):
        single.step(make_batch(np.random.default_rng(100 + 10 * s)))
    c = torch.cat([p.detach().reshape(-1) for _, p in single.named_parameters()]).numpy()
    assert np.abs(a - c).max() > 0          # the other rank's gradients really entered the update


def test_ppo_loss_grads_match_autograd():
    """ppo_train.ppo_loss_grads (the closed-form derivative handed to the engine's backward pass) against torch autograd of the same
    ppo_loss_fn expression, and its loss value against the reference-pinned oracle."""
    rng = np.random.default_rng(0)
    b, t = 5, 11
    am = (rng.random((b, t)) < 0.9).astype(np.float32)
    sta = rng.random((b, t)) < 0.6
    lp = -np.abs(rng.standard_normal((b, t))).astype(np.float32)
    olp = (lp + 0.3 * rng.standard_normal((b, t))).astype(np.float32)
    v, ov, adv, ret = (rng.standard_normal((b, t)).astype(np.float32) for _ in range(4))
    loss, logs, dlp, dv = T.ppo_loss_grads(am, lp, v, sta, olp, ov, adv, ret, cliprange_value=0.2, cliprange=0.2, value_loss_coef=0.7)
    o_loss, _ = O.ppo_loss(am, lp, v, sta, olp, ov, adv, ret, cliprange_value=0.2, cliprange=0.2, value_loss_coef=0.7)
    assert abs(loss - float(o_loss)) < 1e-5
    LP = torch.tensor(lp, dtype=torch.float64, requires_grad=True)
    V = torch.tensor(v, dtype=torch.float64, requires_grad=True)
    m = torch.tensor(sta.astype(np.float64) * am)
    n = m.sum()
    OV, RET, ADV, OLP = (torch.tensor(x, dtype=torch.float64) for x in (ov, ret, adv, olp))
    vc = torch.maximum(torch.minimum(V, OV + 0.2), OV - 0.2)
    vf = 0.5 * (torch.maximum((V - RET) ** 2, (vc - RET) ** 2) * m).sum() / n
    r = torch.exp((LP - OLP) * m)
    pg = (torch.maximum(-ADV * r, -ADV * torch.clamp(r, 0.8, 1.2)) * m).sum() / n
    (pg + 0.7 * vf).backward()
    assert np.abs(LP.grad.numpy() - dlp).max() < 1e-7 and np.abs(V.grad.numpy() - dv).max() < 1e-7
    assert np.abs(dlp).max() > 0 and np.abs(dv).max() > 0


def make_bc_batch(rng, b=4, t=10):
    ids = np.full((b, t), PAD, np.int32)
    m = np.zeros((b, t), np.float32)
    for i in range(b):
        n = int(rng.integers(3, t + 1))
        ids[i, :n] = rng.integers(0, CFG.n_vocab, n)
        m[i, :n] = rng.random(n) < 0.5
    m[0, 1] = 1.0
    return dict(input_ids=ids, input_training_mask=m)


def test_bc_term_of_the_ppo_step():
    """loss + w * bc_loss with policy gradients + w * bc gradients (ppo/gpt2/interface.py:183-205); bc_loss = masked next-token
    cross-entropy (JaxSeq loss_fn_mask), checked here against the oracle's token log-probs."""
    rng = np.random.default_rng(5)
    batch, bc = make_batch(rng), make_bc_batch(rng)
    tr, params, lin = make_trainer()
    want, total = 0.0, 0.0
    for i, row in enumerate(bc["input_ids"]):
        n = int((row != PAD).sum())
        lp = og.token_logprobs_all(params, row[:n].tolist(), CFG.n_head, CFG.n_vocab, weights_bf16=False)
        want -= float((lp * bc["input_training_mask"][i, 1:n]).sum())
        total += float(bc["input_training_mask"][i, 1:n].sum())
    assert abs(float(tr.bc_loss(bc)) - want / total) <= 2e-4 * max(1.0, abs(want / total))

    def grads(weight, with_bc):
        tr, _, _ = make_trainer(lr=0.0)
        tr.bc_loss_weight = weight
        loss, logs = tr.step(batch, bc if with_bc else None)
        return loss, logs, [f.grad.clone() for f in tr.flat]
    l0, _, g0 = grads(0.0, False)
    l1, logs1, g1 = grads(1.0, True)
    l3, logs3, g3 = grads(3.0, True)
    assert "bc_loss" not in grads(0.0, True)[1]                         # weight 0: the BC batch is ignored
    assert abs(logs1["bc_loss"] - want / total) <= 2e-4 * max(1.0, abs(want / total))
    assert abs(l1 - (l0 + logs1["bc_loss"])) <= 1e-5 * max(1.0, abs(l1)) and abs(l3 - (l0 + 3.0 * logs3["bc_loss"])) <= 1e-5 * max(1.0, abs(l3))
    assert abs(logs1["ppo_loss"] - l0) <= 1e-6 * max(1.0, abs(l0)) and abs(logs3["total_loss"] - l3) <= 1e-6 * max(1.0, abs(l3))
    for a, b, c in zip(g0, g1, g3):                                     # gradients are affine in the weight: g3 - g0 = 3 (g1 - g0)
        assert float((b - a).abs().max()) > 0
        np.testing.assert_allclose((c - a).numpy(), 3.0 * (b - a).numpy(), rtol=2e-3, atol=1e-6)
    # the value head receives nothing from the BC term
    grp, o, shape = tr._views["vh_w"]
    n = int(np.prod(shape))
    np.testing.assert_allclose(g1[grp][o:o + n].numpy(), g0[grp][o:o + n].numpy(), rtol=0, atol=0)


def test_bc_batch_from_str_segments():
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    tok = ByteTokenizer.synthetic()
    items = [[("Wordle:\n", 0.0), ("c r a n e\n", 1.0), ("b y b b g\n", 0.0), ("s l a t e\n", 1.0)], [("Wordle:\n", 0.0), ("a b\n", 1.0)]]
    out = T.bc_batch_from_str_segments(items, tok, 24, tok.pad_token_id)
    for i, segs in enumerate(items):
        ids = sum((list(tok.encode(s)) for s, _ in segs), [])
        ms = sum(([w] * len(tok.encode(s)) for s, w in segs), [])
        ids, ms = ids[-24:], ms[-24:]                                    # LEFT truncation, RIGHT padding
        assert out["input_ids"][i, :len(ids)].tolist() == ids and (out["input_ids"][i, len(ids):] == tok.pad_token_id).all()
        assert out["input_training_mask"][i, :len(ms)].tolist() == ms and (out["input_training_mask"][i, len(ms):] == 0).all()
    short = T.bc_batch_from_str_segments(items[:1], tok, 8, tok.pad_token_id)
    full = sum((list(tok.encode(s)) for s, _ in items[0]), [])
    assert short["input_ids"][0].tolist() == full[-8:]


def test_bc_step_lowers_the_bc_loss():
    """``bc_step``: AdamW on the masked cross-entropy alone — the loss of a fixed batch falls, the value head does not move."""
    rng = np.random.default_rng(9)
    bc = make_bc_batch(rng, b=6)
    tr, _, _ = make_trainer(lr=3e-3)
    vh0 = tr.vh_w.detach().clone()
    losses = [tr.bc_step(bc) for _ in range(8)]
    assert losses[-1] < losses[0] - 0.05 and all(np.isfinite(losses))
    assert abs(losses[0] - float(tr.__class__.bc_loss(make_trainer()[0], bc))) < 1e-5        # the first loss is the untouched model's
    assert torch.equal(tr.vh_w.detach(), vh0)
    assert tr.steps == 8


def _agree_main(rank, world, port, out_dir):
    import torch.distributed as dist
    from lmrl_gym_b200.ppo import PPOData
    from lmrl_gym_b200.ppo_train import agreed_steps_per_epoch, global_mean_kl, whiten_advantages_global
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7 + rank)
    n = 11 if rank == 0 else 5                        # rank 1 dropped episodes (truncated chains): fewer PPO data than rank 0
    datas = []
    for _ in range(n):
        t = int(rng.integers(4, 9))
        sta = rng.random(t) < 0.6
        sta[0] = True
        adv = np.where(sta, rng.standard_normal(t) * (2.0 + rank) + rank, 0.0).astype(np.float32)
        datas.append(PPOData(np.arange(t + 1, dtype=np.int32), sta, np.zeros(t, np.float32), np.zeros(t, np.float32), adv, np.zeros(t, np.float32)))
    steps = agreed_steps_per_epoch(len(datas), 2)
    kl = global_mean_kl(np.full(3 + rank, float(rank + 1)))
    white = whiten_advantages_global(datas)
    acts = np.concatenate([np.asarray(d.old_advantages)[np.asarray(d.should_take_action, bool)] for d in white])
    rest = np.concatenate([np.asarray(d.old_advantages)[~np.asarray(d.should_take_action, bool)] for d in white])
    np.save(os.path.join(out_dir, f"agree{rank}.npy"), np.array([steps, kl, acts.sum(), (acts.astype(np.float64) ** 2).sum(), acts.size, np.abs(rest).max() if rest.size else 0.0]))
    dist.destroy_process_group()


def test_two_rank_ppo_round_agrees_on_steps_kl_and_whitening(tmp_path):
    """ranks with different numbers of PPO data must take the same number of optimizer steps (one all_reduce
    per step: otherwise NCCL hangs), see the same mean KL (so kl_coef stays equal) and whiten with the statistics of the global batch."""
    import torch.multiprocessing as mp
    port = 31500 + os.getpid() % 2000
    mp.spawn(_agree_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "agree0.npy"), np.load(tmp_path / "agree1.npy")
    assert a[0] == b[0] == 2                                    # min(11 // 2, 5 // 2)
    assert a[1] == b[1] == (3 * 1.0 + 4 * 2.0) / 7              # token-weighted global mean
    n = a[4] + b[4]
    mean, var = (a[2] + b[2]) / n, (a[3] + b[3]) / n - ((a[2] + b[2]) / n) ** 2
    assert abs(mean) < 1e-6 and abs(var - 1.0) < 1e-4           # the UNION of both ranks' action advantages is white, neither half alone is
    assert abs(a[2] / a[4]) > 1e-3 and a[5] == 0.0 and b[5] == 0.0
    from lmrl_gym_b200.ppo_train import agreed_steps_per_epoch, global_mean_kl
    assert agreed_steps_per_epoch(11, 2) == 5 and global_mean_kl(np.array([1.0, 3.0])) == 2.0        # no process group: local values
