"""One small PPO loop on the GPU (BASELINE.json configs[4] data flow, lmrl_gym_b200/ppo_train.py): fused Wordle rollouts -> trajectory
chains -> device PPO data -> optimizer steps (hand-written backward pass, or the torch-autograd library path for comparison) ->
device-to-device weight hand-over, after which the rollout engine must compute with the UPDATED weights (its logits equal the trainer's
own forward)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("backward", ["b200", "torch"])
def test_ppo_rounds_update_the_engine_policy(backward):
    import torch
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200.wordle import B200WordleEnv
    from lmrl_gym_b200 import ppo as P, ppo_train as T
    from lmrl_gym_b200.randomness import seed_generator
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)
    B = 64
    eng = Engine(max_envs=B, max_input_length=192, max_new_tokens=4, max_turns=6, tokenizer=tok)
    params = init_gpt2_params(cfg, seed=0)
    lin = dict(dense=dict(kernel=(np.random.default_rng(1).standard_normal((cfg.d_model, 1)) / 16).astype(np.float32), bias=np.zeros(1, np.float32)))
    env = B200WordleEnv(eng, "wordle_official.txt", bad_word_reward=-10.0)
    pol = B200GPT2Policy(eng, params, cfg, seed=0)
    inf = P.B200PPOInference(eng, None, lin, initial_policy_params=params, config=cfg, pad_token_id=tok.pad_token_id)
    trainer = T.B200PPOTrain(eng, params, lin, cfg, lr=1e-4, pad_token_id=tok.pad_token_id, activation_dtype="f32", backward=backward, max_rows=16 * 192,
                              max_scored_rows=16 * 192)
    prompts = [tok.encode("Wordle:\n"), tok.encode("Wordle:\nc r a n e\nb y g b b\n")]
    before = pol.logits(prompts)
    logs = T.ppo_train_loop(env, pol, inf, trainer, P.AdaptiveKLController(0.001, 0.1, 10000), n_rounds=2, n_rollouts=B, rollout_bsize=B, epochs=1,
                            train_bsize=16, max_length=192, gamma=1.0, lam=0.95, seed_generator=seed_generator(0))
    assert len(logs) == 2 and all(l["updates"] >= 1 and l["n_ppo_data"] > 0 and np.isfinite(l["loss"]) for l in logs)
    assert logs[0]["mean_kl"] < 1e-6            # round 0: the policy IS the initial policy
    assert logs[1]["mean_kl"] > 0.0             # round 1: it has moved
    after = pol.logits(prompts)
    assert np.abs(after - before).max() > 1e-3
    with torch.no_grad():
        for i, p in enumerate(prompts):
            ids = torch.as_tensor([p], device=trainer.device).long()
            h = trainer._hidden(ids, torch.ones_like(ids, dtype=torch.bool))[0, -1]
            want = (h @ trainer.p["wte"][:cfg.n_vocab].T).float().cpu().numpy()
            assert np.abs(after[i] - want).max() <= 6e-2, np.abs(after[i] - want).max()
    eng.close()


def test_device_adamw_matches_optax_formula_and_torch():
    """lmrl_adamw_dev against the optax.adamw update written out in float64 numpy (the formula of the reference's optimizer,
    train_ppo_gpt2.py:144-165) and against torch.optim.AdamW, over 5 steps on 3.3 M parameters; tolerance per step = 2e-4 of the
    update size (the kernel's fast division / sqrt) + one float32 rounding of the parameter."""
    import torch
    from lmrl_gym_b200._lib import check
    from lmrl_gym_b200.engine import Engine
    eng = Engine(max_envs=8, max_input_length=16, max_new_tokens=2)
    n = 3_300_001
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(n).astype(np.float32)
    lr, b1, b2, eps, wd = 1e-3, 0.9, 0.95, 1e-8, 0.01
    dev = torch.device("cuda", 0)
    p = torch.from_numpy(p0.copy()).to(dev)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    pt = torch.nn.Parameter(torch.from_numpy(p0.copy()).to(dev))
    opt = torch.optim.AdamW([pt], lr=lr, betas=(b1, b2), eps=eps, weight_decay=wd)
    p64, m64, v64 = p0.astype(np.float64), np.zeros(n), np.zeros(n)
    for step in range(1, 6):
        g = (rng.standard_normal(n) * 0.1).astype(np.float32)
        gd = torch.from_numpy(g).to(dev)
        torch.cuda.synchronize()
        check(eng.lib.lmrl_adamw_dev(eng.handle, p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, lr, b1, b2, eps, wd, step, 1.0))
        eng.sync()
        pt.grad = gd.clone()
        opt.step()
        m64 = b1 * m64 + (1 - b1) * g
        v64 = b2 * v64 + (1 - b2) * g.astype(np.float64) ** 2
        p64 = p64 - lr * ((m64 / (1 - b1 ** step)) / (np.sqrt(v64 / (1 - b2 ** step)) + eps) + wd * p64)
    got = p.cpu().numpy()
    tol = 5 * (lr * 2e-4 + np.spacing(np.float32(np.abs(p0).max())))          # per step: fast-maths error of an update of size ~lr + one fp32 rounding of p
    assert np.abs(got - p64).max() <= tol
    assert np.abs(got - pt.detach().cpu().numpy()).max() <= tol
    assert np.abs(got - p0).max() > lr                          # it moved
    # grad_scale: a summed gradient of 2 ranks scaled by 1/2 is the same update
    p2 = torch.from_numpy(p0.copy()).to(dev); m2, v2 = torch.zeros_like(p2), torch.zeros_like(p2)
    p3 = torch.from_numpy(p0.copy()).to(dev); m3, v3 = torch.zeros_like(p3), torch.zeros_like(p3)
    g1 = torch.from_numpy((rng.standard_normal(n) * 0.1).astype(np.float32)).to(dev)
    g2 = g1 * 2
    torch.cuda.synchronize()
    check(eng.lib.lmrl_adamw_dev(eng.handle, p2.data_ptr(), g1.data_ptr(), m2.data_ptr(), v2.data_ptr(), n, lr, b1, b2, eps, 0.0, 1, 1.0))
    check(eng.lib.lmrl_adamw_dev(eng.handle, p3.data_ptr(), g2.data_ptr(), m3.data_ptr(), v3.data_ptr(), n, lr, b1, b2, eps, 0.0, 1, 0.5))
    eng.sync()
    assert torch.equal(p2, p3)
    eng.close()


def test_device_token_chains_equal_the_text_path():
    """trajectory -> token chain on the device (lmrl_rollout_chains_host) == tokenising the episode's final history item by item
    (TokenTrajectory.from_text_trajectory of the trajectory train_ppo_gpt2.py:307-335 builds): ids, is_action, per-token rewards, done —
    for a policy that mixes legal guesses, illegal strings and wins, and with a short max_input_length so some streams get LEFT-truncated
    (those must come back as None)."""
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.environment import text_env_eval
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params, steer_to_letters
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200.wordle import B200WordleEnv
    from lmrl_gym_b200 import ppo as P, ppo_train as T
    from lmrl_gym_b200.randomness import seed_generator
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)
    params = steer_to_letters(init_gpt2_params(cfg, seed=0), tok)
    for max_in, expect_trunc in ((192, False), (48, True)):
        B = 128
        eng = Engine(max_envs=B, max_input_length=max_in, max_new_tokens=6, max_turns=6, tokenizer=tok)
        env = B200WordleEnv(eng, "wordle_official.txt", require_words_in_vocab=False, bad_word_reward=-10.0)
        pol = B200GPT2Policy(eng, params, cfg, seed=1)
        raw, _ = text_env_eval(env, pol, n_rollouts=B, bsize=B, seed_generator=seed_generator(3), verbose=False)
        dev = T.device_token_chains(eng, B, max_length=max_in + 6)
        n_none = 0
        for ep, ch in zip(raw, dev):
            hist = ep[-1].post_transition_history
            n_tok_text = sum(len(tok.encode(t.text)) for t in hist)
            if ch is None:
                n_none += 1
                assert n_tok_text > max_in or n_tok_text >= max_in + 6
                continue
            ref = T.episodes_to_chains([ep], tok, max_length=10 ** 6, gamma=1.0)[0]
            tt = P.TokenTrajectory.from_text_trajectory(ref.text_trajectory, tok)
            got = ch.token_trajectory
            assert got.tokens.tolist() == tt.tokens.tolist()
            assert got.is_action.tolist() == tt.is_action.tolist()
            assert got.reward.tolist() == tt.reward.tolist() and bool(got.done) == bool(tt.done)
        assert (n_none > 0) == expect_trunc
        eng.close()
