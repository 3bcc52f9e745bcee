"""Maze env — CPU restatement of the reference (oracle; test infrastructure only).

Follows llm_rl_scripts/maze/env/env.py:
  * observation renderers                    env.py:8-81
  * move table / update_position             env.py:94-107
  * reward functions                         env.py:109-131
  * step (step limit, goal, unknown action collapse, last_k window)   env.py:161-184
  * reset (goal draw, start draw, options)   env.py:186-214
and maze_utils.py:9-52 (named configurations) with layouts mazes.py:6-18, 50-58.
The layouts are restated as row strings ('#' wall, '.' free).  Pinned by
tests/golden/maze_*.json (generated from the imported reference).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

from .mt19937 import MT19937

LAYOUTS = {
    # mazes.py:6-18
    "double_t_maze": [
        "#############",
        "#.....#.....#",
        "###.#####.###",
        "###.#####.###",
        "###.#####.###",
        "###.......###",
        "######.######",
        "######.######",
        "######.######",
        "#############",
    ],
    # mazes.py:50-58
    "umaze": [
        "#####",
        "#...#",
        "#.#.#",
        "#.#.#",
        "#####",
    ],
}
GOALS = {"double_t_maze": (8, 6), "umaze": (3, 3)}  # maze_utils.py:11-18

MOVES = {  # env.py:94-99
    "move left\n": (0, -1),
    "move right\n": (0, 1),
    "move up\n": (-1, 0),
    "move down\n": (1, 0),
}
WALL_ORDER = (("to your right", (0, 1)), ("to your left", (0, -1)), ("above you", (-1, 0)), ("below you", (1, 0)))

REWARDS = {  # (goal, unknown action, otherwise)   env.py:109-131
    "standard_reward": (0.0, -4.0, -1.0),
    "illegal_penalty_reward": (1.0, -1.0, 0.0),
    "illegal_penalty_diff_scale": (1.0, -100.0, -1.0),
}


def grid_of(name: str) -> List[List[int]]:
    return [[1 if ch == "#" else 0 for ch in row] for row in LAYOUTS[name]]


def spaced(n: int) -> str:
    return " ".join(str(n))  # env.py:57-58


def walls_sentence(grid, pos) -> str:
    """env.py:8-13 with the fixed direction order of env.py:59."""
    rel = [name for name, (dy, dx) in WALL_ORDER if grid[pos[0] + dy][pos[1] + dx] == 1]
    if len(rel) == 0:
        return "There are no walls near you."
    if len(rel) == 1:
        return f"There is a wall {rel[0]}."
    return f"There are walls {', '.join(rel)}."


def describe(mode: str, grid, pos, goal) -> str:
    g = f"The goal is at position {spaced(goal[0])}, {spaced(goal[1])}."
    w = walls_sentence(grid, pos)
    if mode == "describe_observation":
        return f"{g} {w}\n"
    if mode == "describe_observation_give_position":
        p = f"Your current position is at position {spaced(pos[0])}, {spaced(pos[1])}."
        return f"{g} {p} {w}\n"
    if mode == "describe_observation_only_walls":
        return f"{w}\n"
    raise ValueError(mode)


class MazeOracle:
    def __init__(self, maze_name="double_t_maze", describe_function="describe_observation_give_position",
                 reward_function="standard_reward", last_k=1, max_steps=100):
        self.grid = grid_of(maze_name)
        self.valid_goals = [list(GOALS[maze_name])]
        self.mode = describe_function
        self.rw = REWARDS[reward_function or "standard_reward"]
        self.last_k = last_k
        self.max_steps = max_steps
        self.num_steps = 0
        self.goal = list(GOALS[maze_name])
        self.position = None

    def free_cells(self):
        return [[y, x] for y, row in enumerate(self.grid) for x, v in enumerate(row) if v == 0]

    def reset(self, seed: Optional[int] = None, options: Optional[Dict] = None):
        """env.py:186-214.  seed=None (global RNG state) is not reproducible and not modelled."""
        rng = MT19937(seed if seed is not None else 0)
        self.num_steps = 0
        if options is not None and "goal" in options:
            self.goal = list(options["goal"])
        else:
            self.goal = list(self.valid_goals[rng.choice_index(len(self.valid_goals))])
        cells = self.free_cells()
        cells.remove(self.goal)
        if options is not None and "init_position" in options:
            assert list(options["init_position"]) in cells
            self.position = list(options["init_position"])
        else:
            self.position = list(cells[rng.choice_index(len(cells))])
        return ((describe(self.mode, self.grid, self.position, self.goal), False),)

    def step(self, history):
        """env.py:161-184 on (text, is_action) tuples."""
        assert history[-1][1]
        if self.max_steps is not None and self.num_steps >= self.max_steps:
            return (("Failure\n", False),), -1.0, True
        action = history[-1][0]
        if action in MOVES:
            dy, dx = MOVES[action]
            if self.grid[self.position[0] + dy][self.position[1] + dx] == 0:
                self.position = [self.position[0] + dy, self.position[1] + dx]
        at_goal = self.position == self.goal
        reward = self.rw[0] if at_goal else (self.rw[1] if action not in MOVES else self.rw[2])
        if at_goal:
            return (("Success\n", False),), reward, True
        self.num_steps += 1
        obs = (describe(self.mode, self.grid, self.position, self.goal), False)
        if action not in MOVES:
            return (obs,), reward, False
        new_hist = list(history) + [obs]
        new_hist = new_hist[max(0, len(new_hist) - self.last_k):]
        return tuple(new_hist), reward, False
