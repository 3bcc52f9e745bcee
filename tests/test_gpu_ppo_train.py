"""One small PPO loop on the GPU (BASELINE.json configs[4] data flow, lmrl_gym_b200/ppo_train.py): fused Wordle rollouts -> trajectory
chains -> device PPO data -> TorchPPOTrain steps (the torch-autograd LIBRARY path) -> device-to-device weight hand-over, after which
the rollout engine must compute with the UPDATED weights (its logits equal the trainer's own forward)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_ppo_rounds_update_the_engine_policy():
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
    trainer = T.TorchPPOTrain(eng, params, lin, cfg, lr=1e-4, pad_token_id=tok.pad_token_id, activation_dtype="f32")
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
