"""Generate tests/golden/*.json by RUNNING THE REFERENCE (build container only).

    python -m oracle.gen_golden            # needs /root/reference (read-only)

The reference's env code is imported unmodified from /root/reference with three
cosmetic stub modules (termcolor.colored, IPython.embed, skimage.draw), exactly
as SURVEY.md §8c describes.  The vectors are committed; nothing at test / bench /
smoke time reads /root/reference.  Test infrastructure only.
"""
from __future__ import annotations

import io
import json
import os
import random
import sys
import types
import contextlib

REF = os.environ.get("LMRL_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def install_stubs():
    tc = types.ModuleType("termcolor")
    tc.colored = lambda s, *a, **k: s
    sys.modules.setdefault("termcolor", tc)
    ip = types.ModuleType("IPython")
    ip.embed = lambda *a, **k: None
    sys.modules.setdefault("IPython", ip)
    sk = types.ModuleType("skimage")
    skd = types.ModuleType("skimage.draw")
    skd.disk = skd.random_shapes = skd.rectangle = None
    sk.draw = skd
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.draw", skd)
    if REF not in sys.path:
        sys.path.insert(0, REF)


def hist_to_json(h):
    return None if h is None else [[t.text, bool(t.is_action)] for t in h]


def hist_digest(h):
    """Long histories (last_k=40) are stored as length + sha1 + the last two items."""
    import hashlib
    j = hist_to_json(h)
    if len(j) <= 4:
        return dict(full=j)
    blob = json.dumps(j, ensure_ascii=True, separators=(",", ":")).encode()
    return dict(n=len(j), sha1=hashlib.sha1(blob).hexdigest(), tail=j[-2:])


GARBAGE = [
    "", "\n", " ", "xx", "qqqqq", "SLATE", "Crane", "cr ane", " c r a n e\n", "c r a n e", "crane\n\n",
    "\tcrane", "crane ", "cra\nne", "cr\tne", "c-ane", "cran3", "toolong", "abcd", "ééééé", "crané",
    "a b c d e f", " slate ", "s l a t e \n", "zzzzz", "aaaaa", "��", "c  r a n e",
]


def gen_wordle():
    from LLM_RL.environment import Text, BatchedTextPolicy, interact_environment, TextEnvToBatchedTextEnv
    from llm_rl_scripts.wordle.env.env import ReformatWordleEnvironment, WordleEnvironment
    from llm_rl_scripts.wordle.env.game import Vocabulary

    class Script(BatchedTextPolicy):
        """Chooses actions from a seeded host RNG; 'smart' peeks at the env's filtered list."""

        def __init__(self, benv, mode, seed):
            self.benv, self.mode, self.r = benv, mode, random.Random(seed)
            self.log = None

        def act(self, text_history, done=None):
            out = []
            for i, (h, d) in enumerate(zip(text_history, done)):
                if d:
                    out.append(None)
                    continue
                game = self.benv.batch_env_copies[i].env.state
                kind = self.mode
                if kind == "mix":
                    kind = self.r.choice(["smart", "smart", "vocab", "garbage", "spaced"])
                if kind == "smart":
                    a = self.r.choice(game.vocab.filtered_vocab)
                elif kind == "vocab":
                    a = self.r.choice(game.vocab.all_vocab)
                elif kind == "spaced":
                    a = " ".join(self.r.choice(game.vocab.filtered_vocab)) + "\n"
                else:
                    a = self.r.choice(GARBAGE)
                self.log[i].append(a)
                out.append(h + (Text(a, True),))
            return out

    cases = []
    vocab_dir = os.path.join(REF, "llm_rl_scripts", "wordle", "vocab")
    plan = [
        ("wordle_official.txt", True, -1.0, "mix", 24),
        ("wordle_official.txt", True, -1.0, "smart", 16),
        ("wordle_official.txt", True, -1.0, "vocab", 8),
        ("wordle_official.txt", True, -1.0, "garbage", 8),
        ("wordle_official_200.txt", True, -10.0, "mix", 16),
        ("wordle_official_200.txt", False, -1.0, "mix", 16),
        ("tweet_words.txt", True, -1.0, "smart", 8),
        ("wordle_official_800.txt", True, 0.0, "mix", 8),
    ]
    seed_src = random.Random(2024)
    for vocab_file, require, bad, mode, n in plan:
        vocab = Vocabulary.from_file(os.path.join(vocab_dir, vocab_file), fill_cache=False)
        env = ReformatWordleEnvironment(WordleEnvironment(vocab, require_words_in_vocab=require, bad_word_reward=bad))
        benv = TextEnvToBatchedTextEnv(env)
        seeds = [seed_src.choice([seed_src.getrandbits(64), seed_src.getrandbits(31), seed_src.randrange(0, 10)]) for _ in range(n)]
        pol = Script(benv, mode, seed_src.getrandbits(32))
        pol.log = [[] for _ in range(n)]
        sizes = [[] for _ in range(n)]
        orig_step = benv.step

        def step(text_history, done=None, _orig=orig_step, _benv=benv, _sizes=sizes):
            res = _orig(text_history, done=done)
            for i, r in enumerate(res):
                if r is not None:
                    _sizes[i].append(_benv.batch_env_copies[i].env.state.vocab.filtered_vocab_size())
            return res

        benv.step = step
        inter = interact_environment(benv, pol, env_seed=seeds, env_options=[None] * n, bsize=n)
        for i in range(n):
            cases.append(dict(
                vocab_file=vocab_file, require_words_in_vocab=require, bad_word_reward=bad, seed=seeds[i],
                actions=pol.log[i],
                steps=[dict(obs=hist_to_json(t.post_transition_history), reward=t.reward,
                            reward_is_int=isinstance(t.reward, int), done=bool(t.done), n_filtered=sizes[i][k])
                       for k, t in enumerate(inter[i])],
            ))
    return cases


def gen_maze():
    from LLM_RL.environment import Text, BatchedTextPolicy, interact_environment, TextEnvToBatchedTextEnv
    from llm_rl_scripts.maze.env.maze_utils import setup_maze_env
    from llm_rl_scripts.maze.env.mazes import double_t_maze_optimal_directions

    optimal = double_t_maze_optimal_directions()
    BAD = ["jump\n", "move left", "", "move  left\n", "Move left\n", "move top left\n", "move up\n\n"]
    MOVES = ["move left\n", "move right\n", "move up\n", "move down\n"]

    class Script(BatchedTextPolicy):
        def __init__(self, benv, mode, seed):
            self.benv, self.mode, self.r = benv, mode, random.Random(seed)
            self.log = None

        def act(self, text_history, done=None):
            out = []
            for i, (h, d) in enumerate(zip(text_history, done)):
                if d:
                    out.append(None)
                    continue
                e = self.benv.batch_env_copies[i]
                u = self.r.random()
                if self.mode == "optimal" and u < 0.8 and tuple(e.position) in optimal:
                    a = optimal[tuple(e.position)]
                elif u < 0.92:
                    a = self.r.choice(MOVES)
                else:
                    a = self.r.choice(BAD)
                self.log[i].append(a)
                out.append(h + (Text(a, True),))
            return out

    cases = []
    plan = [
        ("double_t_maze", "describe_observation_give_position", "standard_reward", 1, 100, "optimal", 12, False),
        ("double_t_maze", "describe_observation_give_position", "standard_reward", 1, 100, "optimal", 25, True),
        ("double_t_maze", "describe_observation_only_walls", "standard_reward", 40, 100, "random", 6, False),
        ("double_t_maze", "describe_observation", "illegal_penalty_reward", 4, 12, "optimal", 8, False),
        ("umaze", "describe_observation_give_position", "illegal_penalty_diff_scale", 3, 8, "random", 8, False),
        ("umaze", "describe_observation_only_walls", None, 40, 30, "random", 6, False),
    ]
    seed_src = random.Random(77)
    for name, desc, rew, last_k, max_steps, mode, n, use_init in plan:
        env = setup_maze_env(name, desc, rew, last_k=last_k, max_steps=max_steps)
        benv = TextEnvToBatchedTextEnv(env)
        seeds = [seed_src.choice([seed_src.getrandbits(64), seed_src.randrange(0, 100)]) for _ in range(n)]
        if use_init:
            free = [(y, x) for (y, x) in optimal.keys()]
            options = [{"init_position": free[i % len(free)]} for i in range(n)]
        else:
            options = [None] * n
        pol = Script(benv, mode, seed_src.getrandbits(32))
        pol.log = [[] for _ in range(n)]
        positions = [[] for _ in range(n)]
        orig_step = benv.step

        def step(text_history, done=None, _orig=orig_step, _benv=benv, _pos=positions):
            res = _orig(text_history, done=done)
            for i, r in enumerate(res):
                if r is not None:
                    _pos[i].append(list(map(int, _benv.batch_env_copies[i].position)))
            return res

        benv.step = step
        first = {}
        orig_reset = benv.reset

        def reset(seed=None, options=None, _orig=orig_reset, _first=first):
            h = _orig(seed, options)
            _first["h"] = h
            return h

        benv.reset = reset
        inter = interact_environment(benv, pol, env_seed=seeds, env_options=options, bsize=n)
        for i in range(n):
            cases.append(dict(
                maze_name=name, describe_function=desc, reward_function=rew, last_k=last_k, max_steps=max_steps,
                seed=seeds[i], options=None if options[i] is None else {k: list(v) for k, v in options[i].items()},
                reset_obs=hist_to_json(first["h"][i]), actions=pol.log[i],
                steps=[dict(obs=hist_digest(t.post_transition_history), reward=t.reward, done=bool(t.done),
                            position=positions[i][k]) for k, t in enumerate(inter[i])],
            ))
    return cases


def gen_eval_loop():
    """text_env_eval summary + per-episode stats with the reference loop, seeds from seed_generator."""
    from LLM_RL.environment import Text, BatchedTextPolicy, text_env_eval
    from llm_rl_scripts.wordle.env.env import ReformatWordleEnvironment, WordleEnvironment
    from llm_rl_scripts.wordle.env.game import Vocabulary
    from llm_rl_scripts.maze.env.randomness import seed_generator

    words = ["raise", "clout", "nymph", "befog", "waded", "jazzy", "SLATE", "crane", "slate", "x"]

    class Cyclic(BatchedTextPolicy):
        """Deterministic: action depends only on the visible history (stateless)."""

        def act(self, text_history, done=None):
            out = []
            for h, d in zip(text_history, done):
                if d:
                    out.append(None)
                    continue
                k = (len(h) * 7 + sum(len(t.text) for t in h)) % len(words)
                out.append(h + (Text(" ".join(words[k]) + "\n", True),))
            return out

    vocab = Vocabulary.from_file(os.path.join(REF, "llm_rl_scripts", "wordle", "vocab", "wordle_official.txt"), fill_cache=False)
    env = ReformatWordleEnvironment(WordleEnvironment(vocab))
    inter, summary = text_env_eval(env, Cyclic(), n_rollouts=10, seed_generator=seed_generator(0), bsize=4, verbose=False)
    return dict(
        words=words, n_rollouts=10, bsize=4, seed=0,
        summary={k: {kk: float(vv) for kk, vv in v.items()} for k, v in summary.items()},
        episodes=[dict(ret=float(sum(t.reward for t in ep)), length=len(ep), done=bool(ep[-1].done),
                       final=hist_to_json(ep[-1].post_transition_history)) for ep in inter],
    )


def gen_seed_streams():
    from llm_rl_scripts.maze.env.randomness import seed_generator
    out = {}
    for s in [0, 1, 12345, 2**40 + 17]:
        g = seed_generator(s)
        out[str(s)] = [next(g) for _ in range(8)]
    return out


def main():
    install_stubs()
    os.makedirs(OUT, exist_ok=True)
    with contextlib.redirect_stdout(io.StringIO()):
        wordle = gen_wordle()
        maze = gen_maze()
        loop = gen_eval_loop()
        seeds = gen_seed_streams()
    for name, obj in [("wordle_cases.json", wordle), ("maze_cases.json", maze), ("eval_loop.json", loop), ("seed_streams.json", seeds)]:
        with open(os.path.join(OUT, name), "w") as f:
            json.dump(obj, f, ensure_ascii=True, separators=(",", ":"))
        print(name, os.path.getsize(os.path.join(OUT, name)), "bytes")


if __name__ == "__main__":
    main()
