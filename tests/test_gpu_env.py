"""GPU parity: device Wordle / Maze envs (through the C ABI) vs the reference goldens and the CPU oracle."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from lmrl_gym_b200.environment import Text, BatchedTextPolicy, interact_environment, text_env_eval  # noqa: E402
from lmrl_gym_b200.vocab import load_vocab_words  # noqa: E402


def _digest(hist):
    j = [[t.text, bool(t.is_action)] for t in hist]
    if len(j) <= 4:
        return dict(full=j)
    blob = json.dumps(j, ensure_ascii=True, separators=(",", ":")).encode()
    return dict(n=len(j), sha1=hashlib.sha1(blob).hexdigest(), tail=j[-2:])


class Replay(BatchedTextPolicy):
    def __init__(self, scripts):
        self.scripts, self.t = scripts, [0] * len(scripts)

    def act(self, text_history, done=None):
        out = []
        for i, (h, d) in enumerate(zip(text_history, done)):
            if d:
                out.append(None)
                continue
            out.append(h + (Text(self.scripts[i][self.t[i]], True),))
            self.t[i] += 1
        return out


def test_wordle_env_matches_reference_goldens(golden_dir):
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.wordle import B200WordleEnv
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    cases = json.load(open(os.path.join(golden_dir, "wordle_cases.json")))
    groups = {}
    for c in cases:
        groups.setdefault((c["vocab_file"], c["require_words_in_vocab"], c["bad_word_reward"]), []).append(c)
    eng = Engine(max_envs=64, tokenizer=ByteTokenizer.synthetic())
    for (vf, req, bad), group in groups.items():
        env = B200WordleEnv(eng, vf, req, bad)
        n = len(group)
        sizes = [[] for _ in range(n)]
        orig = env.step

        def step(text_history, done=None, _orig=orig, _sizes=sizes, _env=env):
            res = _orig(text_history, done=done)
            for i, r in enumerate(res):
                if r is not None:
                    _sizes[i].append(int(_env.last_n_filtered[i]))
            return res

        env.step = step
        inter = interact_environment(env, Replay([c["actions"] for c in group]), env_seed=[c["seed"] for c in group],
                                     env_options=[None] * n, bsize=n)
        for i, c in enumerate(group):
            assert len(inter[i]) == len(c["steps"])
            for k, (t, st) in enumerate(zip(inter[i], c["steps"])):
                assert [[x.text, x.is_action] for x in t.post_transition_history] == st["obs"]
                assert t.reward == st["reward"] and isinstance(t.reward, int) == st["reward_is_int"]
                assert t.done == st["done"]
                assert sizes[i][k] == st["n_filtered"]
    eng.close()


def test_maze_env_matches_reference_goldens(golden_dir):
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.maze import B200MazeEnv
    cases = json.load(open(os.path.join(golden_dir, "maze_cases.json")))
    groups = {}
    for c in cases:
        key = (c["maze_name"], c["describe_function"], c["reward_function"], c["last_k"], c["max_steps"])
        groups.setdefault(key, []).append(c)
    eng = Engine(max_envs=64)
    for key, group in groups.items():
        env = B200MazeEnv(eng, *key)
        n = len(group)
        first = {}
        orig_reset = env.reset

        def reset(seed=None, options=None, _orig=orig_reset, _first=first):
            _first["h"] = _orig(seed, options)
            return _first["h"]

        env.reset = reset
        inter = interact_environment(env, Replay([c["actions"] for c in group]), env_seed=[c["seed"] for c in group],
                                     env_options=[c["options"] for c in group], bsize=n)
        for i, c in enumerate(group):
            assert [[x.text, x.is_action] for x in first["h"][i]] == c["reset_obs"]
            assert len(inter[i]) == len(c["steps"])
            for t, st in zip(inter[i], c["steps"]):
                assert _digest(t.post_transition_history) == st["obs"]
                assert t.reward == st["reward"] and t.done == st["done"]
    eng.close()


def test_wordle_env_large_batch_vs_oracle():
    """2048 concurrent episodes of mixed legal / smart / garbage actions vs the numpy-free CPU oracle on a subset,
    plus size-independent invariants on all of them."""
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.wordle import B200WordleEnv
    from oracle.wordle import ReformatWordleOracle
    words = load_vocab_words("wordle_official.txt")
    B = 2048
    eng = Engine(max_envs=B)
    env = B200WordleEnv(eng, "wordle_official.txt")
    r = random.Random(11)
    seeds = [r.getrandbits(64) for _ in range(B)]
    scripts = []
    for i in range(B):
        s = []
        for _ in range(6):
            u = r.random()
            s.append(r.choice(words) if u < 0.7 else " ".join(r.choice(words)) + "\n" if u < 0.85 else r.choice(["", "abcde", "APPLE", "zz zz", "crané"]))
        scripts.append(s)
    inter = interact_environment(env, Replay(scripts), env_seed=seeds, env_options=[None] * B, bsize=B)
    n_valid = 0
    for ep in inter:
        assert 1 <= len(ep) <= 6 and ep[-1].done and all(not t.done for t in ep[:-1])
        for t in ep:
            assert t.reward in (0, -1, -1.0)
            n_valid += isinstance(t.reward, int)
        if len(ep) < 6:
            assert ep[-1].reward == 0 and ep[-1].post_transition_history[-1].text == "g g g g g\n"
    assert n_valid > B * 3
    for i in range(0, B, 16):   # 128 episodes through the pure-Python oracle
        o = ReformatWordleOracle(words)
        h = o.reset(seeds[i])
        for k, t in enumerate(inter[i]):
            h, rew, d = o.step(h + ((scripts[i][k], True),))
            assert [[x.text, x.is_action] for x in t.post_transition_history] == [[a, b] for a, b in h]
            assert t.reward == rew and type(t.reward) is type(rew) and t.done == d
    eng.close()


def test_scripted_random_mixture_rollout_replays_through_oracle():
    """Env-only rollout (variant 2b of the bench): device RandomMixturePolicy-equivalent actions (every one a legal word, half of
    them drawn from the filtered vocabulary) exercise target draw / transition / re-filter / win detection; every
    transition is replayed through the CPU oracle env bit-exactly."""
    import torch
    from lmrl_gym_b200 import _lib
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    from lmrl_gym_b200.vocab import load_vocab_words
    from lmrl_gym_b200.wordle import B200WordleEnv
    from oracle.mt19937 import seed_stream
    from oracle.wordle import ReformatWordleOracle
    n, T, K = 512, 6, 6
    eng = Engine(max_envs=n, max_new_tokens=K, max_turns=T, tokenizer=ByteTokenizer.synthetic())
    B200WordleEnv(eng, "wordle_official.txt")
    gen = seed_stream(9)
    seeds = np.asarray([next(gen) for _ in range(n)], dtype=np.uint64)
    dev = torch.device("cuda", 0)
    seeds_dev = torch.from_numpy(seeds.view(np.int64)).to(dev)
    ep_ret = torch.zeros(n, dtype=torch.float32, device=dev); ep_len = torch.zeros(n, dtype=torch.int32, device=dev)
    ep_done = torch.zeros(n, dtype=torch.uint8, device=dev); rew = torch.zeros((n, T), dtype=torch.float32, device=dev)
    valid = torch.zeros((n, T), dtype=torch.uint8, device=dev); fb = torch.zeros((n, T, 5), dtype=torch.uint8, device=dev)
    aux = torch.zeros((n, T), dtype=torch.int32, device=dev); act = torch.zeros((n, T, K), dtype=torch.int32, device=dev)
    alen = torch.zeros((n, T), dtype=torch.int32, device=dev)
    out = _lib.lmrl_rollout_out(ep_ret.data_ptr(), ep_len.data_ptr(), ep_done.data_ptr(), rew.data_ptr(), valid.data_ptr(), fb.data_ptr(),
                                aux.data_ptr(), act.data_ptr(), alen.data_ptr())
    eng.wordle_scripted_rollout_dev(seeds_dev.data_ptr(), n, 0.5, 1234, out)
    eng.sync()
    ep_len_h, ep_ret_h, rew_h, fb_h, aux_h, act_h = (t.cpu().numpy() for t in (ep_len, ep_ret, rew, fb, aux, act))
    words = load_vocab_words("wordle_official.txt")
    wins = 0
    for i in range(n):
        o = ReformatWordleOracle(words)
        h = o.reset(int(seeds[i]))
        total = 0.0
        for t in range(int(ep_len_h[i])):
            w = int(act_h[i, t, 0]) & 0xFFFFFFFF
            word = "".join(chr(ord("a") + ((w >> (5 * j)) & 31)) for j in range(5))
            assert word in words
            h, r, d = o.step(h + ((" ".join(word) + "\n", True),))
            assert float(r) == float(rew_h[i, t]) and len(o.inner.filtered) == int(aux_h[i, t]), (i, t)
            assert h[-1][0] == " ".join("gyb"[c] for c in fb_h[i, t]) + "\n"
            assert d == (t == ep_len_h[i] - 1)
            total += r
        assert total == float(ep_ret_h[i])
        wins += rew_h[i, ep_len_h[i] - 1] == 0.0
    assert 1 <= ep_len_h.min() and ep_len_h.max() == 6 and wins > 0
    eng.close()
