"""Oracle env restatements vs golden vectors produced by the imported reference."""
import hashlib
import json
import os

import pytest

from oracle import maze as omaze
from oracle import wordle as owordle
from lmrl_gym_b200.vocab import load_vocab_words


def _digest(hist):
    j = [[t, bool(a)] for t, a in hist]
    if len(j) <= 4:
        return dict(full=j)
    blob = json.dumps(j, ensure_ascii=True, separators=(",", ":")).encode()
    return dict(n=len(j), sha1=hashlib.sha1(blob).hexdigest(), tail=j[-2:])


@pytest.fixture(scope="module")
def wordle_cases(golden_dir):
    return json.load(open(os.path.join(golden_dir, "wordle_cases.json")))


def test_wordle_oracle_matches_reference(wordle_cases):
    vocabs = {}
    n_steps = n_valid = n_win = 0
    for case in wordle_cases:
        vf = case["vocab_file"]
        if vf not in vocabs:
            vocabs[vf] = load_vocab_words(vf)
        env = owordle.ReformatWordleOracle(vocabs[vf], case["require_words_in_vocab"], case["bad_word_reward"])
        hist = env.reset(case["seed"])
        assert len(case["actions"]) == len(case["steps"])
        for action, step in zip(case["actions"], case["steps"]):
            hist, r, done = env.step(hist + ((action, True),))
            assert [[t, a] for t, a in hist] == step["obs"]
            assert r == step["reward"] and isinstance(r, int) == step["reward_is_int"]
            assert done == step["done"]
            assert len(env.inner.filtered) == step["n_filtered"]
            n_steps += 1
            n_valid += step["reward_is_int"]
            n_win += (step["reward"] == 0 and step["reward_is_int"])
        assert step["done"]
    assert n_steps > 400 and n_valid > 150 and n_win > 10  # the goldens exercise every branch


def test_maze_oracle_matches_reference(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "maze_cases.json")))
    outcomes = set()
    for case in cases:
        env = omaze.MazeOracle(case["maze_name"], case["describe_function"], case["reward_function"],
                               case["last_k"], case["max_steps"])
        hist = env.reset(case["seed"], case["options"])
        assert [[t, a] for t, a in hist] == case["reset_obs"]
        for action, step in zip(case["actions"], case["steps"]):
            hist, r, done = env.step(hist + ((action, True),))
            assert _digest(hist) == step["obs"]
            assert r == step["reward"] and done == step["done"]
            assert env.position == step["position"]
        assert step["done"]
        outcomes.add(hist[-1][0])
    assert outcomes == {"Success\n", "Failure\n"}


def test_optimal_table_is_bfs_optimal():
    """Known-answer material of the reference (mazes.py:20-48 vs maze_utils.py:91-116): following the
    restated layout greedily by BFS distance reaches the goal in the BFS-optimal number of moves."""
    from collections import deque
    grid = omaze.grid_of("double_t_maze")
    goal = omaze.GOALS["double_t_maze"]
    dist = {goal: 0}
    dq = deque([goal])
    while dq:
        y, x = dq.popleft()
        for dy, dx in omaze.MOVES.values():
            c = (y + dy, x + dx)
            if grid[c[0]][c[1]] == 0 and c not in dist:
                dist[c] = dist[(y, x)] + 1
                dq.append(c)
    assert len(dist) == 26 and max(dist.values()) == 12
