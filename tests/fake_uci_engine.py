"""A minimal UCI engine for the tests of lmrl_gym_b200.chess_env.UciOpponent: answers `go` with the lexicographically smallest legal
move of the position it was given (legal moves come from the repository's own chess rules)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lmrl_gym_b200.chess_env import ChessBoard  # noqa: E402

board = ChessBoard()
for line in sys.stdin:
    w = line.split()
    if not w:
        continue
    if w[0] == "uci":
        print("id name fake\nuciok", flush=True)
    elif w[0] == "isready":
        print("readyok", flush=True)
    elif w[0] == "position":
        fen = " ".join(w[2:w.index("moves")] if "moves" in w else w[2:])
        board = ChessBoard(fen)
        for m in (w[w.index("moves") + 1:] if "moves" in w else []):
            board.push_uci(m)
    elif w[0] == "go":
        print("info depth 1\nbestmove " + sorted(board.legal_moves())[0], flush=True)
    elif w[0] == "quit":
        break
