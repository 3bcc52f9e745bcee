"""Device-resident Maze environment behind the reference's BatchedTextEnv interface.

Drop-in for ``TextEnvToBatchedTextEnv(setup_maze_env(maze_name, describe_function, reward_function, last_k,
max_steps))`` (llm_rl_scripts/maze/env/maze_utils.py:9-52, env.py:133-214).  Position, goal and step counter are
int32 tensors in HBM; the transition / reward kernel is bit-exact against the reference ``step()``; observation text
is rendered from the closed sentence set of env.py:8-81.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from ._lib import check, ptr
from .engine import Engine, ENV_MAZE
from .environment import BatchedTextEnv, Text, TextHistory
from .wordle import encode_actions, seed_to_u64

LAYOUTS = {
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
    "umaze": ["#####", "#...#", "#.#.#", "#.#.#", "#####"],
}
GOALS = {"double_t_maze": (8, 6), "umaze": (3, 3)}
REWARDS = {"standard_reward": (0.0, -4.0, -1.0), "illegal_penalty_reward": (1.0, -1.0, 0.0),
           "illegal_penalty_diff_scale": (1.0, -100.0, -1.0)}
DESCRIBE = {"describe_observation": 0, "describe_observation_give_position": 1, "describe_observation_only_walls": 2}
ACTIONS = ("move left\n", "move right\n", "move up\n", "move down\n")
_WALLS = (("to your right", 0, 1), ("to your left", 0, -1), ("above you", -1, 0), ("below you", 1, 0))


def maze_grid(name: str) -> np.ndarray:
    return np.asarray([[1 if c == "#" else 0 for c in row] for row in LAYOUTS[name]], dtype=np.uint8)


def _spaced(n: int) -> str:
    return " ".join(str(int(n)))


def describe(mode: int, grid: np.ndarray, pos, goal) -> str:
    rel = [nm for nm, dy, dx in _WALLS if grid[pos[0] + dy, pos[1] + dx] == 1]
    if not rel:
        walls = "There are no walls near you."
    elif len(rel) == 1:
        walls = f"There is a wall {rel[0]}."
    else:
        walls = f"There are walls {', '.join(rel)}."
    g = f"The goal is at position {_spaced(goal[0])}, {_spaced(goal[1])}."
    if mode == 0:
        return f"{g} {walls}\n"
    if mode == 1:
        return f"{g} Your current position is at position {_spaced(pos[0])}, {_spaced(pos[1])}. {walls}\n"
    return f"{walls}\n"


def maze_solver(grid: np.ndarray, goal_positions) -> Dict[Tuple[int, int], str]:
    """Optimal first move of every free cell: breadth-first from the goals (maze_utils.py:91-116, which takes the INVERTED grid; this
    one takes the wall grid, 1 = wall).  Neighbours are expanded down, right, up, left — the reference's order — so cells with two
    shortest paths get the same move."""
    H, W = grid.shape
    goals = [tuple(int(v) for v in g) for g in goal_positions]
    assert all(grid[g] == 0 for g in goals), "goal must be a free cell"
    moves: Dict[Tuple[int, int], str] = {}
    seen = set(goals)
    frontier = list(goals)
    while frontier:
        nxt = []
        for (y, x) in frontier:
            for dy, dx, back in ((1, 0, "move up\n"), (0, 1, "move left\n"), (-1, 0, "move down\n"), (0, -1, "move right\n")):
                c = (y + dy, x + dx)
                if c in seen or not (0 <= c[0] < H and 0 <= c[1] < W) or grid[c] == 1:
                    continue
                seen.add(c)
                moves[c] = back
                nxt.append(c)
        frontier = nxt
    return moves


def compute_move_accuracy(policy, reranker: bool = False, maze_name: str = "double_t_maze") -> float:
    """Percentage of free cells (goal excluded) on which ``policy``'s action for the fully-observed description is the optimal move
    (maze_utils.py:63-89; the reference hard-codes the double-T maze and its hand-written table, which equals ``maze_solver``'s)."""
    grid, goal = maze_grid(maze_name), GOALS[maze_name]
    best = maze_solver(grid, [goal])
    cells = [tuple(int(v) for v in c) for c in np.argwhere(grid == 0)]
    hists = [(Text(describe(1, grid, c, goal), False),) for c in cells if c != tuple(goal)]
    if reranker:
        preds = [policy.act(h)[-1].text for h in hists]
    else:                                                        # one batched call instead of the reference's one call per cell
        preds = [h[-1].text for h in policy.act(hists, done=[False] * len(hists))]
    correct = sum(p == best[c] for p, c in zip(preds, [c for c in cells if c != tuple(goal)]))
    return 100.0 * correct / len(hists)


class B200MazeEnv(BatchedTextEnv):
    def __init__(self, engine: Engine, maze_name: str = "double_t_maze",
                 describe_function: str = "describe_observation_give_position",
                 reward_function: Optional[str] = None, last_k: int = 1, max_steps: Optional[int] = 100):
        if maze_name not in LAYOUTS:
            raise ValueError(f"unknown maze name: {maze_name}")
        if describe_function not in DESCRIBE:
            raise ValueError(f"unknown describe function: {describe_function}")
        reward_function = reward_function or "standard_reward"
        if reward_function not in REWARDS:
            raise ValueError(f"unknown reward function: {reward_function}")
        self._lmrl_engine = engine
        self._lmrl_kind = ENV_MAZE
        self.grid = maze_grid(maze_name)
        self.goal_default = GOALS[maze_name]
        self.mode = DESCRIBE[describe_function]
        self.last_k = int(last_k)
        self.max_steps = max_steps
        rewards = np.asarray(REWARDS[reward_function], dtype=np.float32)
        H, W = self.grid.shape
        # fused rollout tables: canonical token ids of every cell's observation sentence (default goal) and of the 4 actions
        obs_tok = obs_len = act_tok = act_len = None
        pitch = 0
        tok = engine.tokenizer
        self._lmrl_fusable = tok is not None and 1 <= self.last_k <= 64
        if self._lmrl_fusable:
            sent = {}
            for y in range(H):
                for x in range(W):
                    if self.grid[y, x] == 0:
                        sent[y * W + x] = tok.encode(describe(self.mode, self.grid, (y, x), self.goal_default))
            pitch = max(len(v) for v in sent.values())
            obs_tok = np.zeros((H * W, pitch), np.int32)
            obs_len = np.zeros(H * W, np.int32)
            for c, v in sent.items():
                obs_tok[c, :len(v)] = v
                obs_len[c] = len(v)
            acts = [tok.encode(a) for a in ACTIONS]
            if max(len(a) for a in acts) > 8:
                self._lmrl_fusable = False
            else:
                act_tok = np.zeros((4, 8), np.int32)
                act_len = np.asarray([len(a) for a in acts], np.int32)
                for i, a in enumerate(acts):
                    act_tok[i, :len(a)] = a
        if not self._lmrl_fusable:
            obs_tok = obs_len = act_tok = act_len = None
        check(engine.lib.lmrl_load_maze(engine.handle, ptr(np.ascontiguousarray(self.grid)), H, W, self.goal_default[0],
                                        self.goal_default[1], -1 if max_steps is None else int(max_steps), self.last_k,
                                        ptr(rewards), ptr(obs_tok), ptr(obs_len), pitch, ptr(act_tok), ptr(act_len)))
        engine.env_kind = ENV_MAZE
        self.n: Optional[int] = None
        self.pos = None
        self.goal = None

    def reset(self, seed=None, options=None) -> List[TextHistory]:
        if seed is None and options is None:
            seed, options = [None], [None]
        elif seed is None:
            seed = [None] * len(options)
        elif options is None:
            options = [None] * len(seed)
        assert len(seed) == len(options)
        n = len(seed)
        seeds = np.asarray([seed_to_u64(s) for s in seed], dtype=np.uint64)
        opt_pos = np.full((n, 2), -1, np.int32)
        opt_goal = np.full((n, 2), -1, np.int32)
        for i, o in enumerate(options):
            if o is not None and "init_position" in o:
                p = list(o["init_position"])
                assert self.grid[p[0], p[1]] == 0 and tuple(p) != tuple(o.get("goal", self.goal_default))
                opt_pos[i] = p
            if o is not None and "goal" in o:
                opt_goal[i] = list(o["goal"])
        self.pos = np.zeros((n, 2), np.int32)
        self.goal = np.zeros((n, 2), np.int32)
        e = self._lmrl_engine
        check(e.lib.lmrl_maze_reset_host(e.handle, ptr(seeds), ptr(opt_pos), ptr(opt_goal), n, ptr(self.pos), ptr(self.goal)))
        self.n = n
        return [(Text(describe(self.mode, self.grid, self.pos[i], self.goal[i]), False),) for i in range(n)]

    def step(self, text_history, done=None):
        assert self.n is not None, "reset must be called before step"
        assert len(text_history) == self.n, "batch size must be the same as the number of environments initalized"
        if done is None:
            done = [False] * self.n
        n = self.n
        actions = []
        for h, d in zip(text_history, done):
            if d or h is None:
                actions.append(None)
            else:
                assert h[-1].is_action
                actions.append(h[-1].text)
        skip = np.asarray([a is None for a in actions], dtype=np.uint8)
        arr, lens = encode_actions(actions)
        reward = np.zeros(n, np.float32)
        dn = np.zeros(n, np.uint8)
        known = np.zeros(n, np.uint8)
        outcome = np.zeros(n, np.int32)
        e = self._lmrl_engine
        check(e.lib.lmrl_maze_step_host(e.handle, ptr(arr), ptr(lens), arr.shape[1], ptr(skip), n, ptr(reward), ptr(dn),
                                        ptr(known), ptr(outcome), ptr(self.pos)))
        out = []
        for i in range(n):
            if actions[i] is None:
                out.append(None)
                continue
            if outcome[i] == 2:
                hist = (Text("Failure\n", False),)
            elif outcome[i] == 1:
                hist = (Text("Success\n", False),)
            else:
                obs = Text(describe(self.mode, self.grid, self.pos[i], self.goal[i]), False)
                if not known[i]:
                    hist = (obs,)  # unknown action string collapses the history (env.py:179-180)
                else:
                    full = tuple(text_history[i]) + (obs,)
                    hist = full[max(0, len(full) - self.last_k):]
            out.append((hist, float(reward[i]), bool(dn[i])))
        return out

    def describe_cell(self, cell: int) -> str:
        W = self.grid.shape[1]
        return describe(self.mode, self.grid, (cell // W, cell % W), self.goal_default)

    def fused_options(self, options) -> Optional[np.ndarray]:
        """env_options of one batch -> [n][2] init positions for the fused loop (None if no option is given); raises if an
        option the fused loop does not implement ('goal') is present."""
        if options is None or all(o is None or len(o) == 0 for o in options):
            return None
        out = np.full((len(options), 2), -1, np.int32)
        for i, o in enumerate(options):
            if not o:
                continue
            if "goal" in o:
                raise NotImplementedError("the fused Maze loop uses the maze's default goal; pass goal options through the generic loop")
            if "init_position" in o:
                p = list(o["init_position"])
                assert self.grid[p[0], p[1]] == 0 and tuple(p) != tuple(self.goal_default)
                out[i] = p
        return out

    def copy(self):
        return self
