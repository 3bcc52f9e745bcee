"""The hand-written backward pass of the policy + value head (csrc/train_kernels.inl, lmrl_policy_grad_host) against torch autograd of
the same network in fp32 on the GPU (the library path of ppo_train.B200PPOTrain, itself checked against the oracle GPT-2 on the CPU):
every parameter gradient of a PPO minibatch — embeddings, all four matrices + biases and both LayerNorms of every layer, ln_f, the tied
LM head (through wte) and the LinearHead value head.  The engine computes with bf16 operands and fp32 accumulation, so the bar is a
relative L2 error <= 2e-2 and a cosine >= 0.9995 per tensor (measured: worst 4e-3 / 0.99999) (weights are rounded to bf16 first so both sides hold the same numbers)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_batch(rng, cfg, pad, lp_fn, b=12, t=48):
    ids = np.full((b, t), pad, np.int32)
    sta = np.zeros((b, t - 1), bool)
    for i in range(b):
        n = int(rng.integers(6, t + 1))
        ids[i, :n] = rng.integers(0, cfg.n_vocab, n)
        sta[i, 1:n - 1] = rng.random(n - 2) < 0.5
    sta[0, 1] = True
    f = lambda s: (s * rng.standard_normal((b, t - 1))).astype(np.float32)  # noqa: E731
    return dict(input_ids=ids, should_take_action=sta, old_logprobs=None, old_values=f(0.3), old_advantages=f(1.0), old_returns=f(0.5))


@pytest.mark.parametrize("d_model,n_head,d_ff", [(256, 4, 1024), (1024, 16, 4096)], ids=["d256", "d1024-medium-width"])
def test_handwritten_backward_matches_autograd(d_model, n_head, d_ff):
    import torch
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200 import ppo_train as T
    from lmrl_gym_b200.value_inference import forward_tokens
    from oracle.gpt2 import bf16_round
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config(n_layer=2, n_head=n_head, d_model=d_model, d_ff=d_ff)
    rng = np.random.default_rng(0)
    params = init_gpt2_params(cfg, seed=3)
    for k in ("wte", "wpe"):
        params[k] = bf16_round(params[k])
    for L in params["layers"]:
        for k in ("attn_w", "proj_w", "fc_w", "fc2_w"):
            L[k] = bf16_round(L[k])
        for k in ("attn_b", "proj_b", "fc_b", "fc2_b", "ln1_b", "ln2_b"):
            L[k] = (0.02 * rng.standard_normal(L[k].shape)).astype(np.float32)
        for k in ("ln1_g", "ln2_g"):
            L[k] = (1.0 + 0.05 * rng.standard_normal(L[k].shape)).astype(np.float32)
    lin = dict(dense=dict(kernel=(rng.standard_normal((cfg.d_model, 1)) / 16).astype(np.float32), bias=np.asarray([0.1], np.float32)))
    eng = Engine(max_envs=32, max_input_length=64, max_new_tokens=4, tokenizer=tok)
    B200GPT2Policy(eng, params, cfg)
    ref = T.B200PPOTrain(eng, params, lin, cfg, pad_token_id=tok.pad_token_id, activation_dtype="f32", backward="torch")
    b2 = T.B200PPOTrain(eng, params, lin, cfg, pad_token_id=tok.pad_token_id, backward="b200", max_rows=2048, max_scored_rows=1024)
    batch = make_batch(rng, cfg, tok.pad_token_id, None)
    # old log-probs near the current ones so that both the clipped and the unclipped branch of the policy loss are taken
    ids = batch["input_ids"]
    lens = (ids != tok.pad_token_id).sum(1)
    cur = forward_tokens(eng, 0, 0, [ids[i, :lens[i]] for i in range(ids.shape[0])], logprobs=True, pitch=ids.shape[1])["logprobs"][:, :ids.shape[1] - 1]
    batch["old_logprobs"] = (cur + 0.25 * rng.standard_normal(cur.shape)).astype(np.float32)
    for f in ref.flat:
        f.grad = None
    loss_ref, logs_ref = ref.loss(batch)
    loss_ref.backward()
    # split path (host-side closed-form loss gradients) first, then the fused device path: same gradients, same loss
    loss_s, logs_s = b2._b200_grads(batch, fused=False)
    torch.cuda.synchronize()
    g_split = [g.clone() for g in b2.gflat]
    loss_b, logs_b = b2._b200_grads(batch)
    torch.cuda.synchronize()
    assert abs(loss_s - loss_b) <= 2e-3 * max(1.0, abs(loss_s)), (loss_s, loss_b)
    for k in ("policy_loss", "value_loss", "approx_kl", "policy_clipfrac", "value_clipfrac", "ratio"):
        assert abs(logs_s[k] - logs_b[k]) <= 2e-3 * max(1.0, abs(logs_s[k])), (k, logs_s[k], logs_b[k])
    for a, c in zip(g_split, b2.gflat):
        rel = float((a.double() - c.double()).norm() / a.double().norm())
        assert rel <= 1e-2, rel
    assert abs(float(loss_ref.detach()) - loss_b) <= 2e-2 * max(1.0, abs(loss_b)), (float(loss_ref.detach()), loss_b)
    assert 0.05 < logs_b["policy_clipfrac"] < 0.95
    worst = []
    for name in [n for grp in ref._names for n in grp]:
        grp, o, shape = ref._views[name]
        n = int(np.prod(shape))
        g_ref = ref.flat[grp].grad[o:o + n].double().cpu().numpy()
        g_b = b2.gflat[grp][o:o + n].double().cpu().numpy()
        if name == "wte":                       # only rows of tokens that occur (or are scored) carry signal worth comparing: all rows here
            pass
        nr = np.linalg.norm(g_ref)
        rel = np.linalg.norm(g_b - g_ref) / max(nr, 1e-12)
        cos = float(g_b @ g_ref / max(np.linalg.norm(g_b) * nr, 1e-20))
        worst.append((rel, cos, name, nr))
    worst.sort(reverse=True)
    # the two attention-backward kernels (tensor-core default, CUDA-core cross-check) give the same gradients
    import os
    g_mma = [g.clone() for g in b2.gflat]
    os.environ["LMRL_ATTN_BWD"] = "cuda"
    try:
        b2._b200_grads(batch)
        torch.cuda.synchronize()
    finally:
        del os.environ["LMRL_ATTN_BWD"]
    for a, c in zip(g_mma, b2.gflat):
        rel = float((a.double() - c.double()).norm() / c.double().norm())
        assert rel <= 5e-3, rel
    for g, saved in zip(b2.gflat, g_mma):
        g.copy_(saved)
    print("hand-written backward vs autograd (rel L2 err, cosine, tensor, |g|):")
    for rel, cos, name, nr in worst[:8]:
        print(f"  {rel:.4f}  {cos:.5f}  {name:12s} {nr:.3e}")
    for rel, cos, name, nr in worst:
        assert nr > 0, name
        assert rel <= 2e-2 and cos >= 0.9995, (name, rel, cos)
    eng.close()


def test_handwritten_backward_gpt2_small_and_training_step():
    """GPT-2-small shapes (12 layers, d = 768), a PPO minibatch of 32 sequences: gradients vs autograd (bf16-autocast library path is the
    comparison here, so the bar is looser: cosine >= 0.99 per tensor), then three optimizer steps on the hand-written path track the
    library path's loss and parameters and leave the engine's rollout policy on the updated weights."""
    import time
    import torch
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200 import ppo_train as T
    from lmrl_gym_b200.value_inference import forward_tokens
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config.small()
    rng = np.random.default_rng(1)
    params = init_gpt2_params(cfg, seed=5)
    lin = dict(dense=dict(kernel=(rng.standard_normal((cfg.d_model, 1)) / 28).astype(np.float32), bias=np.zeros(1, np.float32)))
    eng = Engine(max_envs=64, max_input_length=128, max_new_tokens=6, tokenizer=tok)
    pol = B200GPT2Policy(eng, params, cfg)
    ref = T.B200PPOTrain(eng, params, lin, cfg, pad_token_id=tok.pad_token_id, activation_dtype="f32", backward="torch", lr=1e-5)
    b2 = T.B200PPOTrain(eng, params, lin, cfg, pad_token_id=tok.pad_token_id, backward="b200", lr=1e-5, max_rows=8192, max_scored_rows=4096)
    batch = make_batch(rng, cfg, tok.pad_token_id, None, b=32, t=128)
    ids = batch["input_ids"]
    lens = (ids != tok.pad_token_id).sum(1)
    cur = forward_tokens(eng, 0, 0, [ids[i, :lens[i]] for i in range(ids.shape[0])], logprobs=True, pitch=ids.shape[1])["logprobs"][:, :ids.shape[1] - 1]
    batch["old_logprobs"] = (cur + 0.25 * rng.standard_normal(cur.shape)).astype(np.float32)
    for f in ref.flat:
        f.grad = None
    loss_ref, _ = ref.loss(batch)
    loss_ref.backward()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss_b, _ = b2._b200_grads(batch)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    worst = (0.0, 1.0, "")
    for name in [n for grp in ref._names for n in grp]:
        grp, o, shape = ref._views[name]
        n = int(np.prod(shape))
        g_ref = ref.flat[grp].grad[o:o + n].double()
        g_b = b2.gflat[grp][o:o + n].double()
        rel = float((g_b - g_ref).norm() / g_ref.norm().clamp_min(1e-12))
        cos = float((g_b @ g_ref) / (g_b.norm() * g_ref.norm()).clamp_min(1e-20))
        if rel > worst[0]:
            worst = (rel, cos, name)
        assert cos >= 0.99 and rel <= 0.1, (name, rel, cos)
    print(f"GPT-2-small backward: worst tensor {worst[2]} rel {worst[0]:.4f} cos {worst[1]:.5f}; forward + loss + backward of {int(lens.sum())} rows in {dt * 1e3:.1f} ms")
    # three optimizer steps on the same minibatch: the hand-written path tracks the library path (same AdamW kernel, same start)
    before = pol.logits([ids[0, :8].tolist()])
    traj_b, traj_r = [], []
    for _ in range(3):
        traj_b.append(b2.step(batch)[0])
        traj_r.append(ref.step(batch)[0])
    print("loss per step, hand-written backward:", [round(x, 4) for x in traj_b], " autograd:", [round(x, 4) for x in traj_r])
    for a, b in zip(traj_b, traj_r):
        assert abs(a - b) <= 0.05 * max(1.0, abs(b)), (traj_b, traj_r)
    pb = torch.cat([f.data.reshape(-1) for f in b2.flat]).double()
    pr = torch.cat([f.data.reshape(-1) for f in ref.flat]).double()
    assert float((pb - pr).abs().max()) <= 3 * 3 * 1e-5          # at most lr per step apart, and mostly identical signs:
    after = pol.logits([ids[0, :8].tolist()])
    assert np.abs(after - before).max() > 1e-5          # the rollout policy runs on the trained weights
    eng.close()


def test_policy_grad_edge_cases():
    """No scored token (value-only gradient), single-token sequences, a batch of one — the backward pass must run and give the
    gradients autograd gives; misuse is reported through LmrlError, never a crash."""
    import torch
    from lmrl_gym_b200._lib import LmrlError
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200 import ppo_train as T
    from lmrl_gym_b200.value_inference import forward_tokens
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)
    rng = np.random.default_rng(7)
    params = init_gpt2_params(cfg, seed=9)
    lin = dict(dense=dict(kernel=(rng.standard_normal((cfg.d_model, 1)) / 16).astype(np.float32), bias=np.zeros(1, np.float32)))
    eng = Engine(max_envs=8, max_input_length=32, max_new_tokens=4, tokenizer=tok)
    B200GPT2Policy(eng, params, cfg)
    ref = T.B200PPOTrain(eng, params, lin, cfg, pad_token_id=tok.pad_token_id, activation_dtype="f32", backward="torch")
    b2 = T.B200PPOTrain(eng, params, lin, cfg, pad_token_id=tok.pad_token_id, backward="b200", max_rows=256, max_scored_rows=128)
    pad = tok.pad_token_id
    # batch of one sequence of length 5, one action token; plus a one-token sequence (no output position at all)
    ids = np.full((2, 6), pad, np.int32)
    ids[0, :5] = rng.integers(0, cfg.n_vocab, 5)
    ids[1, :1] = rng.integers(0, cfg.n_vocab, 1)
    sta = np.zeros((2, 5), bool)
    sta[0, 3] = True
    f = lambda: rng.standard_normal((2, 5)).astype(np.float32)  # noqa: E731
    cur = forward_tokens(eng, 0, 0, [ids[0, :5], ids[1, :1]], logprobs=True, pitch=8)["logprobs"][:, :5]
    batch = dict(input_ids=ids, should_take_action=sta, old_logprobs=(cur + 0.05).astype(np.float32), old_values=f(), old_advantages=f(), old_returns=f())
    for fz in ref.flat:
        fz.grad = None
    loss_ref, _ = ref.loss(batch)
    loss_ref.backward()
    loss_b, _ = b2._b200_grads(batch)
    torch.cuda.synchronize()
    assert abs(float(loss_ref.detach()) - loss_b) <= 2e-2 * max(1.0, abs(loss_b))
    for grp in (0, 1):
        a, c = ref.flat[grp].grad.double(), b2.gflat[grp].double()
        assert float((a - c).norm() / a.norm()) <= 3e-2
    # value-only: upstream gradient on the values alone (lmrl_policy_grad_host with all-zero dlogprob)
    import ctypes as C
    from lmrl_gym_b200._lib import check
    toks = np.zeros((2, 8), np.int32)
    toks[:, :6] = np.where(ids != pad, ids, 0)
    lens = np.asarray([5, 1], np.int32)
    dlp, dv = np.zeros((2, 8), np.float32), np.zeros((2, 8), np.float32)
    dv[0, :5] = 0.1
    dv[1, 0] = -0.2
    for g in b2.gflat:
        g.zero_()
    keep = []
    gs = b2._grad_struct(keep)
    torch.cuda.synchronize()
    check(eng.lib.lmrl_policy_grad_host(eng.handle, toks.ctypes.data, lens.ctypes.data, 8, 2, dlp.ctypes.data, dv.ctypes.data, C.c_float(1.0), C.byref(gs)))
    torch.cuda.synchronize()
    idt = torch.as_tensor(ids, device=ref.device).long()
    for fz in ref.flat:
        fz.grad = None
    h = ref._hidden(idt, idt != pad)
    vals = (h.float() @ ref.vh_w + ref.vh_b)[..., 0]
    (vals[0, :5].sum() * 0.1 - 0.2 * vals[1, 0]).backward()
    for grp in (0, 1):
        a, c = ref.flat[grp].grad.double(), b2.gflat[grp].double()
        assert float((a - c).norm() / a.norm()) <= 3e-2, grp
    # misuse: more rows than the training capacity
    big = np.zeros((8, 40), np.int32)
    with pytest.raises(LmrlError):
        check(eng.lib.lmrl_policy_grad_host(eng.handle, big.ctypes.data, np.full(8, 40, np.int32).ctypes.data, 40, 8, np.zeros((8, 40), np.float32).ctypes.data,
                                            np.zeros((8, 40), np.float32).ctypes.data, C.c_float(1.0), C.byref(gs)))
    eng.close()


def test_bc_term_on_the_engine_matches_autograd():
    """PPO step with the BC term (loss + w * bc_loss, ppo/gpt2/interface.py:183-205): the engine path (lmrl_ppo_grad_host, then
    lmrl_policy_grad_host accumulating the masked cross-entropy gradient) against fp32 autograd of ``loss + w * bc_loss``."""
    import torch
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200 import ppo_train as T
    from lmrl_gym_b200.value_inference import forward_tokens
    from oracle.gpt2 import bf16_round
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)
    rng = np.random.default_rng(4)
    params = init_gpt2_params(cfg, seed=6)
    for k in ("wte", "wpe"):
        params[k] = bf16_round(params[k])
    for L in params["layers"]:
        for k in ("attn_w", "proj_w", "fc_w", "fc2_w"):
            L[k] = bf16_round(L[k])
    lin = dict(dense=dict(kernel=(rng.standard_normal((cfg.d_model, 1)) / 16).astype(np.float32), bias=np.asarray([0.1], np.float32)))
    eng = Engine(max_envs=32, max_input_length=64, max_new_tokens=4, tokenizer=tok)
    B200GPT2Policy(eng, params, cfg)
    W = 2.5
    ref = T.B200PPOTrain(eng, params, lin, cfg, lr=0.0, pad_token_id=tok.pad_token_id, activation_dtype="f32", backward="torch", bc_loss_weight=W)
    b2 = T.B200PPOTrain(eng, params, lin, cfg, lr=0.0, pad_token_id=tok.pad_token_id, backward="b200", max_rows=2048, max_scored_rows=1024, bc_loss_weight=W)
    batch = make_batch(rng, cfg, tok.pad_token_id, None)
    ids = batch["input_ids"]
    lens = (ids != tok.pad_token_id).sum(1)
    cur = forward_tokens(eng, 0, 0, [ids[i, :lens[i]] for i in range(ids.shape[0])], logprobs=True, pitch=ids.shape[1])["logprobs"][:, :ids.shape[1] - 1]
    batch["old_logprobs"] = (cur + 0.25 * rng.standard_normal(cur.shape)).astype(np.float32)
    items = [[("Wordle:\n", 0.0), ("c r a n e\n", 1.0), ("b y b b g\n", 0.0), ("s l a t e\n", 1.0)], [("Wordle:\n", 0.0), ("a d i e u\n", 1.0)],
             [("Wordle:\n", 0.0), ("s t o r k\n", 1.0), ("g b b y b\n", 0.0)]]
    bc = T.bc_batch_from_str_segments(items, tok, 40, tok.pad_token_id)
    loss_ref, logs_ref = ref.step(batch, bc)
    g_ref = [f.grad.clone() for f in ref.flat]
    loss_b, logs_b = b2.step(batch, bc)
    torch.cuda.synchronize()
    assert abs(logs_b["bc_loss"] - logs_ref["bc_loss"]) <= 2e-2 * max(1.0, abs(logs_ref["bc_loss"])), (logs_b["bc_loss"], logs_ref["bc_loss"])
    assert abs(loss_b - loss_ref) <= 2e-2 * max(1.0, abs(loss_ref)) and abs(loss_b - (logs_b["ppo_loss"] + W * logs_b["bc_loss"])) <= 1e-6 * abs(loss_b)
    assert logs_b["bc_loss"] > 5.0                                        # random-init model: about log(vocab)
    for name in [n for grp in ref._names for n in grp]:
        grp, o, shape = ref._views[name]
        n = int(np.prod(shape))
        a = g_ref[grp][o:o + n].double().cpu().numpy()
        c = b2.gflat[grp][o:o + n].double().cpu().numpy()
        rel = np.linalg.norm(c - a) / max(np.linalg.norm(a), 1e-12)
        cos = float(c @ a / max(np.linalg.norm(c) * np.linalg.norm(a), 1e-20))
        assert rel <= 2e-2 and cos >= 0.9995, (name, rel, cos)
    # the BC term really contributes: without the BC batch the gradients differ
    b2.step(batch, None)
    torch.cuda.synchronize()
    assert float((b2.gflat[0] - g_ref[0]).norm() / g_ref[0].norm()) > 0.05
    eng.close()
