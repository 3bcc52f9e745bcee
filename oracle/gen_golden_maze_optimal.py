"""TEST INFRASTRUCTURE.  Writes tests/golden/maze_optimal.json by importing the reference (this container only; /root/reference does
not travel): the hand-written double-T table (llm_rl_scripts/maze/env/mazes.py:20-48) and ``maze_solver`` (maze_utils.py:91-116) on
both shipped mazes.      python -m oracle.gen_golden_maze_optimal
"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    from oracle.gen_golden import install_stubs
    install_stubs()
    from llm_rl_scripts.maze.env.maze_utils import maze_solver
    from llm_rl_scripts.maze.env.mazes import double_t_maze, double_t_maze_optimal_directions, maze2d_umaze
    out = {}
    table = double_t_maze_optimal_directions()
    out["double_t_maze_table"] = {f"{y},{x}": a for (y, x), a in table.items()}
    for name, fn, goal in (("double_t_maze", double_t_maze, (8, 6)), ("umaze", maze2d_umaze, (3, 3))):
        maze = fn()
        out[name + "_solver"] = {f"{y},{x}": a for (y, x), a in maze_solver(1 - maze, [goal]).items()}
    with open(os.path.join(ROOT, "tests", "golden", "maze_optimal.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
