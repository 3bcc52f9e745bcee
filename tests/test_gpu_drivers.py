"""The task drivers end to end on the GPU with a 2-layer model: same result trees as the reference scripts' (eval_bc_gpt2.py:108-131,
maze/bc/eval_bc.py:160-225), reproducible under --seed, outputs written."""
import importlib.util
import json
import os
import pickle
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "scripts", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _tiny_config(tmp_path):
    p = str(tmp_path / "tiny.json")
    json.dump(dict(n_layer=2, n_head=4, d_model=256, d_ff=1024), open(p, "w"))
    return p


def test_wordle_driver(tmp_path):
    drv = _load("eval_wordle_gpt2")
    cfg = _tiny_config(tmp_path)
    out = str(tmp_path / "out")
    kw = dict(policy_n_rollouts=12, policy_bsize=8, policy_max_input_length=128, policy_max_output_length=6, seed=0, print_episodes=1)
    a = drv.main("config", cfg, "wordle_official.txt", outputs_path=out, **kw)["generation_metrics"]
    b = drv.main("config", cfg, "wordle_official.txt", **kw)["generation_metrics"]
    assert set(a) == {"reward", "done", "length"} and set(a["reward"]) == {"mean", "std", "min", "max"}
    assert a == b                                                        # seeded: reproducible
    assert a["done"]["mean"] == 1.0 and 1 <= a["length"]["min"] <= a["length"]["max"] <= 6
    raw = pickle.load(open(os.path.join(out, "interactions.pkl"), "rb"))
    assert len(raw) == 12 and all(ep[-1].done for ep in raw)
    summ = json.load(open(os.path.join(out, "interactions_summary.json")))
    assert summ["reward"]["mean"] == pytest.approx(float(a["reward"]["mean"]))
    assert float(np.mean([sum(t.reward for t in ep) for ep in raw])) == pytest.approx(float(a["reward"]["mean"]))


def test_maze_driver(tmp_path):
    drv = _load("eval_maze")
    cfg = _tiny_config(tmp_path)
    out = str(tmp_path / "out")
    res = drv.main("config", cfg, outputs_path=out, policy_n_rollouts=2, policy_bsize=2, policy_max_input_length=128,
                   policy_max_output_length=3, seed=1)
    reward, acc = res["reward_eval"], res["move_accuracy"]
    assert len(reward) == 25 + 1 and "avg_reward" in reward and "(1, 1)" in reward
    assert set(reward["avg_reward"]) == {"reward", "done", "length"}
    # a random-init model never says "move left\n": every episode runs into the step limit with the unknown-action reward
    assert reward["avg_reward"]["length"]["mean"] == 101.0 and reward["avg_reward"]["reward"]["mean"] == pytest.approx(-4.0 * 100 - 1.0)
    assert len(acc) == 25 + 1 and 0.0 <= acc["avg_accuracy"] <= 1.0
    assert acc["avg_accuracy"] == pytest.approx(sum(bool(v) for k, v in acc.items() if k != "avg_accuracy") / 25)
    assert os.path.exists(os.path.join(out, "interactions.pkl")) and os.path.exists(os.path.join(out, "interactions_summary.json"))
    # re-ranker for the reward evaluation: only legal move strings are ever sent to the env
    res2 = drv.main("config", cfg, policy_n_rollouts=1, policy_bsize=1, policy_max_input_length=128, policy_max_output_length=3, seed=1,
                    use_reranker_for_reward_eval=True, policy_do_sample=False, do_accuracy_eval=False)
    assert res2["reward_eval"]["avg_reward"]["reward"]["min"] >= -101.0


def test_ppo_training_driver(tmp_path):
    drv = _load("train_ppo_wordle")
    cfg = _tiny_config(tmp_path)
    bc_path = str(tmp_path / "bc.jsonl")
    with open(bc_path, "w") as f:
        for w in ("crane", "slate", "adieu", "stork", "pious", "mange", "blimp", "fudgy"):
            f.write(json.dumps({"sequence": [["Wordle:\n", 0.0], [" ".join(w) + "\n", 1.0]]}) + "\n")
    out = str(tmp_path / "out")
    hist = drv.main("config", cfg, bc_path, "wordle_official.txt", outputs_path=out, n_rounds=2, n_rollouts=16, rollout_bsize=16, train_bsize=4,
                    max_input_length=128, max_output_length=6, max_steps=2, lr=1e-4, seed=0, save_at_end=True)
    assert len(hist) == 2 and all(h["updates"] == 2 and np.isfinite(h["loss"]) for h in hist)
    assert all(h["env_steps"] == 16 * 6 for h in hist)                     # a random-init policy never says a word: six invalid tries each
    assert all(h["reward_mean"] == pytest.approx(-60.0) for h in hist)      # bad_word_reward = -10 in the PPO driver
    assert hist[0]["loss"] > 5.0                                            # the BC term (about log vocab) is part of the loss
    rounds = json.load(open(os.path.join(out, "rounds.json")))
    assert len(rounds) == 2 and rounds[1]["kl_coef"] == pytest.approx(0.001)
    from lmrl_gym_b200.load import load_params
    p2, c2 = load_params("params", os.path.join(out, "policy_params.npz"))
    assert c2.n_layer == 2 and p2["wte"].shape[1] == 256
    vh = np.load(os.path.join(out, "value_head.npz"))
    assert vh["kernel"].shape == (256, 1) and abs(float(vh["bias"][0]) + 4.1) < 0.1
    # without BC data the loss is the PPO loss alone
    hist2 = drv.main("config", cfg, "none", "wordle_official.txt", n_rounds=1, n_rollouts=8, rollout_bsize=8, train_bsize=4, max_input_length=128,
                     max_output_length=6, max_steps=1, seed=0)
    assert hist2[0]["updates"] == 1 and abs(hist2[0]["loss"]) < hist[0]["loss"]


@pytest.mark.parametrize("algorithm", ["ilql", "mc"])
def test_value_rl_driver(tmp_path, algorithm):
    """ILQL (two MLP Q heads) and MC returns (one LinearHead Q head) evaluation drivers: reproducible under --seed, weights from files
    (``params`` mode) give the same rollouts as the in-memory ones they were saved from."""
    drv = _load("eval_value_rl_wordle")
    cfg_path = _tiny_config(tmp_path)
    kw = dict(algorithm=algorithm, policy_n_rollouts=8, policy_bsize=8, policy_max_input_length=128, policy_max_output_length=6, seed=0,
              policy_beta=4.0, print_episodes=0)
    a = drv.main("config", cfg_path, "config", cfg_path, "wordle_official.txt", **kw)["generation_metrics"]
    assert set(a) == {"reward", "done", "length"} and a["done"]["mean"] == 1.0
    from lmrl_gym_b200.load import load_head_params, load_params, save_head_params, save_params
    d = tmp_path / "ckpt"
    d.mkdir()
    base, cfg = load_params("config", cfg_path, seed=1)
    save_params(str(d / "base.npz"), base, cfg)
    if algorithm == "ilql":
        save_head_params(str(d / "q1_head.npz"), load_head_params("config", "", cfg, kind="mlp", seed=2))
        save_head_params(str(d / "q2_head.npz"), load_head_params("config", "", cfg, kind="mlp", seed=3))
    else:
        save_head_params(str(d / "q_head.npz"), load_head_params("config", "", cfg, kind="linear", seed=2))
    b = drv.main("params", str(d), "config", cfg_path, "wordle_official.txt", **kw)["generation_metrics"]
    assert a == b
