"""The per-env device logic (csrc/env_core.h), compiled for the host, against Python itself and the goldens."""
import ctypes as C
import json
import os
import random

import numpy as np
import pytest

from lmrl_gym_b200 import build
from lmrl_gym_b200.tokenizer import ByteTokenizer
from lmrl_gym_b200.vocab import load_packed, unpack_word
from lmrl_gym_b200.maze import maze_grid, GOALS, REWARDS


@pytest.fixture(scope="module")
def hc():
    return C.CDLL(build.build_host())


def _p(a):
    return C.c_void_p(a.ctypes.data)


def test_mt_matches_cpython(hc):
    for seed in [0, 1, 2**32 - 1, 2**32, 2**64 - 1, 987654321987654321]:
        out = np.zeros(1400, np.uint32)
        hc.hc_mt_u32_stream(C.c_uint64(seed), 1400, _p(out))
        r = random.Random(seed)
        assert out.tolist() == [r.getrandbits(32) for _ in range(1400)]
        ns = np.asarray([1, 2, 3, 24, 25, 2315, 12972, 1, 7, 2**31 + 5], np.uint32)
        got = np.zeros(len(ns), np.uint32)
        hc.hc_mt_randbelow_stream(C.c_uint64(seed), _p(ns), len(ns), _p(got))
        r = random.Random(seed)
        assert got.tolist() == [r.choice(range(int(n))) for n in ns]


def test_utf8_replace_matches_python(hc):
    rng = random.Random(5)
    interesting = [0x00, 0x20, 0x41, 0x7F, 0x80, 0xBF, 0xC0, 0xC1, 0xC2, 0xDF, 0xE0, 0xE1, 0xEC, 0xED, 0xEE, 0xEF, 0xF0, 0xF1,
                   0xF3, 0xF4, 0xF5, 0xFF, 0x90, 0x8F, 0xA0, 0x9F]
    for trial in range(4000):
        n = rng.randrange(0, 12)
        b = bytes(rng.choice(interesting) if rng.random() < 0.8 else rng.randrange(256) for _ in range(n))
        want = [ord(c) for c in b.decode("utf-8", errors="replace")]
        buf = np.frombuffer(b, np.uint8).copy() if n else np.zeros(1, np.uint8)
        out = np.zeros(64, np.uint32)
        m = hc.hc_utf8_decode(_p(buf), n, _p(out), 64)
        assert out[:m].tolist() == want, b


def test_normalise_matches_python(hc):
    rng = random.Random(6)
    alphabet = list("abc xyz\n\t\r\x0b\x0c\x1c\x1d\x1e\x1f\x85\xa0    ​    　﻿Zé")
    for trial in range(3000):
        s = "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 14)))
        want = [ord(c) for c in s.strip().replace(" ", "")]
        b = s.encode("utf-8")
        buf = np.frombuffer(b, np.uint8).copy() if b else np.zeros(1, np.uint8)
        out = np.zeros(64, np.uint32)
        m = hc.hc_normalise(_p(buf), len(b), _p(out), 64)
        assert out[:m].tolist() == want, repr(s)


def test_render_line_matches_tokenizer(hc):
    tok = ByteTokenizer.synthetic()
    t = tok.device_tables()
    rng = random.Random(7)
    alphabet = list("abcdefghijklmnopqrstuvwxyzABC019-_'!? \t\n\x1dé �")
    space_id, nl = int(t["byte_tok"][0x20]), tok.newline_id()
    for trial in range(2000):
        s = "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 10)))
        norm = s.strip().replace(" ", "")
        line = " ".join(norm) + "\n"
        want = tok.encode(line)
        b = s.encode("utf-8")
        buf = np.frombuffer(b, np.uint8).copy() if b else np.zeros(1, np.uint8)
        out = np.zeros(128, np.int32)
        m = hc.hc_render_action_line(_p(buf), len(b), _p(t["tok_first"]), _p(t["tok_space"]), _p(t["byte_tok"]), space_id, nl,
                                     _p(out), 128)
        assert out[:m].tolist() == want, repr(s)
        assert tok.decode(want) == line


def test_wordle_core_matches_reference_goldens(hc, golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "wordle_cases.json")))
    sym = {"g": 0, "y": 1, "b": 2}
    for case in cases:
        words = load_packed(case["vocab_file"])
        acts = [a.encode("utf-8") for a in case["actions"]]
        blob = b"".join(acts)
        buf = np.frombuffer(blob, np.uint8).copy() if blob else np.zeros(1, np.uint8)
        lens = np.asarray([len(a) for a in acts], np.int32)
        n = len(acts)
        reward = np.zeros(n, np.float32); done = np.zeros(n, np.uint8); valid = np.zeros(n, np.uint8)
        fb = np.zeros((n, 5), np.uint8); nf = np.zeros(n, np.int32); tg = np.zeros(n, np.uint32)
        hc.hc_wordle_episode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_uint64] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 6
        steps = hc.hc_wordle_episode(_p(words), len(words), int(case["require_words_in_vocab"]), case["bad_word_reward"],
                                     abs(case["seed"]), _p(buf), _p(lens), n, _p(reward), _p(done), _p(valid), _p(fb), _p(nf), _p(tg))
        assert steps == len(case["steps"])
        for k, st in enumerate(case["steps"]):
            assert float(reward[k]) == float(st["reward"])
            assert bool(done[k]) == st["done"]
            assert int(nf[k]) == st["n_filtered"]
            last = st["obs"][-1][0]
            want_fb = [sym[c] for c in last.strip().split(" ")] if last != "\n" else []
            assert [int(c) for c in fb[k] if c < 3] == want_fb
            assert bool(valid[k]) == (last != "\n")


def test_maze_core_matches_reference_goldens(hc, golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "maze_cases.json")))
    hc.hc_maze_episode.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                                 C.c_int] + [C.c_void_p] * 5
    for case in cases:
        grid = maze_grid(case["maze_name"])
        H, W = grid.shape
        gy, gx = GOALS[case["maze_name"]]
        rw = np.asarray(REWARDS[case["reward_function"] or "standard_reward"], np.float32)
        acts = [a.encode() for a in case["actions"]]
        blob = b"".join(acts)
        buf = np.frombuffer(blob, np.uint8).copy() if blob else np.zeros(1, np.uint8)
        lens = np.asarray([len(a) for a in acts], np.int32)
        n = len(acts)
        start = np.zeros(2, np.int32); reward = np.zeros(n, np.float32); done = np.zeros(n, np.uint8)
        outcome = np.zeros(n, np.int32); pos = np.zeros((n, 2), np.int32)
        oy, ox = (case["options"]["init_position"] if case["options"] else (-1, -1))
        steps = hc.hc_maze_episode(_p(np.ascontiguousarray(grid)), H, W, gy, gx, case["max_steps"], _p(rw), abs(case["seed"]), oy, ox,
                                   _p(buf), _p(lens), n, _p(start), _p(reward), _p(done), _p(outcome), _p(pos))
        assert steps == len(case["steps"])
        for k, st in enumerate(case["steps"]):
            assert float(reward[k]) == st["reward"] and bool(done[k]) == st["done"]
            assert pos[k].tolist() == st["position"]


@pytest.mark.parametrize("vocab,episodes", [("10k_words", 40), ("large_words", 5), ("wordle_official_guess", 5), ("wordle_official_400", 40)])
def test_wordle_core_randomised_against_oracle(hc, vocab, episodes):
    """csrc/env_core.h (the code the device kernels run) against the reference-pinned Python oracle on the vocabularies the goldens do
    not cover, with a mixture of smart guesses (drawn from the oracle's filtered list), random vocabulary words and malformed actions
    (upper case, wrong length, inner whitespace, non-ASCII, Unicode whitespace padding)."""
    from lmrl_gym_b200.vocab import load_vocab_words
    from oracle import wordle as ow
    words = load_vocab_words(vocab)
    packed = load_packed(vocab)
    rng = random.Random(sum(vocab.encode()) + 7)
    sym = {"g": 0, "y": 1, "b": 2}
    hc.hc_wordle_episode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_uint64] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 6
    bad = ["CRANE", "cran", "cranes", "cr\tne", "", "c r a n e s", "crané", "12345", "\n", "zzzzz"]
    for ep in range(episodes):
        require, bwr, seed = rng.random() < 0.8, rng.choice([-1.0, -10.0, 0.0]), rng.randrange(2 ** 62)
        o = ow.WordleOracle(words, require, bwr)
        o.reset(seed)
        actions, want = [], []
        while True:
            u = rng.random()
            if u < 0.55:
                w = rng.choice(o.filtered)
            elif u < 0.8:
                w = rng.choice(words)
            else:
                w = rng.choice(bad)
            text = " ".join(w) + "\n" if rng.random() < 0.7 else rng.choice(["", " ", " ", "\t"]) + w + rng.choice(["", "\n", " \n"])
            fb, r, done = o.step(ow.normalise_action(text))
            actions.append(text.encode("utf-8"))
            want.append((fb, float(r), done, len(o.filtered), o.target if fb != "" else None))
            if done:
                break
        blob = b"".join(actions)
        buf = np.frombuffer(blob, np.uint8).copy() if blob else np.zeros(1, np.uint8)
        lens = np.asarray([len(a) for a in actions], np.int32)
        n = len(actions)
        reward = np.zeros(n, np.float32); done_a = np.zeros(n, np.uint8); valid = np.zeros(n, np.uint8)
        fb_a = np.zeros((n, 5), np.uint8); nf = np.zeros(n, np.int32); tg = np.zeros(n, np.uint32)
        steps = hc.hc_wordle_episode(_p(packed), len(packed), int(require), bwr, seed, _p(buf), _p(lens), n, _p(reward), _p(done_a), _p(valid),
                                     _p(fb_a), _p(nf), _p(tg))
        assert steps == n, (vocab, ep, actions)
        for k, (fb, r, done, n_f, target) in enumerate(want):
            ctx = (vocab, ep, k, actions[k])
            assert float(reward[k]) == r and bool(done_a[k]) == done and int(nf[k]) == n_f, ctx
            assert [int(c) for c in fb_a[k] if c < 3] == [sym[c] for c in fb.replace("<", "").replace(">", "")], ctx
            assert bool(valid[k]) == (fb != ""), ctx
            if target is not None:                       # the word sampled from the filtered list BEFORE the feedback (game.py:219)
                assert unpack_word(int(tg[k])) == target, ctx


@pytest.mark.parametrize("maze_name", ["double_t_maze", "umaze"])
def test_maze_core_randomised_against_oracle(hc, maze_name):
    """csrc/env_core.h Maze reset / step against the reference-pinned oracle: every reward function, random seeds (goal draw + start draw
    of env.py:193-202) or an init_position option, random legal / blocked / unknown actions, step limits down to 3."""
    from oracle import maze as om
    hc.hc_maze_episode.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                                 C.c_int] + [C.c_void_p] * 5
    grid = maze_grid(maze_name)
    H, W = grid.shape
    gy, gx = GOALS[maze_name]
    rng = random.Random(len(maze_name))
    moves = ["move left\n", "move right\n", "move up\n", "move down\n"]
    junk = ["jump\n", "move left", "", "Move up\n", "move  down\n", "move up\n\n", "move north\n"]
    for ep in range(200):
        reward_function = rng.choice(list(REWARDS))
        max_steps = rng.choice([3, 7, 25, 100])
        o = om.MazeOracle(maze_name, "describe_observation_give_position", reward_function, last_k=1, max_steps=max_steps)
        seed = rng.randrange(2 ** 62)
        options = None
        if rng.random() < 0.4:
            free = [c for c in o.free_cells() if c != [gy, gx]]
            options = {"init_position": tuple(rng.choice(free))}
        hist = o.reset(seed, options)
        start0 = list(o.position)
        actions, want = [], []
        while True:
            a = rng.choice(moves) if rng.random() < 0.85 else rng.choice(junk)
            hist, r, done = o.step(tuple(hist) + ((a, True),))
            actions.append(a.encode())
            want.append((float(r), done, list(o.position)))
            if done:
                break
        blob = b"".join(actions)
        buf = np.frombuffer(blob, np.uint8).copy() if blob else np.zeros(1, np.uint8)
        lens = np.asarray([len(a) for a in actions], np.int32)
        n = len(actions)
        start = np.zeros(2, np.int32); reward = np.zeros(n, np.float32); done_a = np.zeros(n, np.uint8)
        outcome = np.zeros(n, np.int32); pos = np.zeros((n, 2), np.int32)
        oy, ox = options["init_position"] if options else (-1, -1)
        rw = np.asarray(REWARDS[reward_function], np.float32)
        steps = hc.hc_maze_episode(_p(np.ascontiguousarray(grid)), H, W, gy, gx, max_steps, _p(rw), seed, oy, ox, _p(buf), _p(lens), n,
                                   _p(start), _p(reward), _p(done_a), _p(outcome), _p(pos))
        assert steps == n and start.tolist() == start0, (maze_name, ep)
        for k, (r, done, p) in enumerate(want):
            assert float(reward[k]) == r and bool(done_a[k]) == done and pos[k].tolist() == p, (maze_name, ep, k, actions[k])
