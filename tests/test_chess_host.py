"""Chess rules (csrc/chess_core.h through the C ABI) and the Chess-endgames env mirror (lmrl_gym_b200/chess_env.py; reference
llm_rl_scripts/chess/env/env.py).  python-chess is not available, so the anchors are the published perft counts of the standard test
positions (move generation incl. castling, en passant, promotions, pins), and SAN / FEN / game-end facts that follow from the rules."""
import os
import sys

import numpy as np
import pytest

from lmrl_gym_b200 import chess_env as CE
from lmrl_gym_b200.environment import Text, TextPolicy, interact_environment

PERFT = [
    (None, [20, 400, 8902, 197281, 4865609]),
    ("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", [48, 2039, 97862, 4085603]),
    ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", [14, 191, 2812, 43238, 674624]),
    ("r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", [6, 264, 9467, 422333]),
    ("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", [44, 1486, 62379, 2103487]),
    ("r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10", [46, 2079, 89890, 3894594]),
]


@pytest.mark.parametrize("fen,counts", PERFT, ids=["start", "kiwipete", "pos3", "pos4", "pos5", "pos6"])
def test_perft_known_answers(fen, counts):
    b = CE.ChessBoard(fen)
    for d, want in enumerate(counts, start=1):
        assert b.perft(d) == want, (fen, d)


def play(board, sans):
    return [board.push_san(s) for s in sans]


def test_fen_follows_python_chess_conventions():
    b = CE.ChessBoard()
    assert b.fen() == CE.START_FEN
    play(b, ["e4"])
    assert b.fen() == "rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq - 0 1"        # no legal en-passant capture: '-' (not e3)
    play(b, ["d5", "e5", "f5"])
    assert b.fen() == "rnbqkbnr/ppp1p1pp/8/3pPp2/8/8/PPPP1PPP/RNBQKBNR w KQkq f6 0 3"    # exf6 is legal: the square is shown
    assert play(b, ["exf6"]) == ["e5f6"]
    assert b.fen() == "rnbqkbnr/ppp1p1pp/5P2/3p4/8/8/PPPP1PPP/RNBQKBNR b KQkq - 0 3"
    play(b, ["Nc6", "Ke2"])                                                               # the king moved: white's rights are gone
    assert b.fen().split()[2] == "kq"
    b.pop()
    assert b.fen().split()[2] == "KQkq"
    # castling rights are cleaned against the pieces actually standing on their squares
    assert CE.ChessBoard("4k3/8/8/8/8/8/8/4K2R w KQkq - 3 9").fen() == "4k3/8/8/8/8/8/8/4K2R w K - 3 9"
    with pytest.raises(ValueError):
        CE.ChessBoard("8/8/8/8 w - - 0 1")


def test_san_generation_and_parsing():
    two_knights = "4k3/8/8/8/8/5N2/8/1N2K3 w - - 0 1"
    b = CE.ChessBoard(two_knights)
    with pytest.raises(CE.IllegalMove):
        b.push_san("Nd2")                                   # ambiguous
    assert b.push_san("Nbd2") == "b1d2"
    b = CE.ChessBoard(two_knights)
    assert b.push_uci("f3d2") == "Nfd2"
    b = CE.ChessBoard("4k3/8/8/R7/8/8/8/R3K3 w - - 0 1")
    assert b.push_uci("a1a3") == "R1a3"                     # same file: disambiguate by rank
    b.pop()
    assert b.push_san("R5a3") == "a5a3"
    b = CE.ChessBoard("8/P7/8/8/8/8/8/k1K5 w - - 0 1")
    assert b.push_uci("a7a8q") == "a8=Q#"                   # a2 is covered by the new queen, b1 / b2 by the king: mate
    b.pop()
    assert b.push_uci("a7a8n") == "a8=N"
    b.pop()
    assert b.push_san("a8Q") == "a7a8q"                     # '=' is optional in the grammar
    b.pop()
    with pytest.raises(CE.IllegalMove):
        b.push_san("a8")                                    # a promotion needs its piece
    b = CE.ChessBoard()
    play(b, ["f3", "e5", "g4"])
    assert b.push_uci("d8h4") == "Qh4#"
    assert b.is_checkmate() and b.is_game_over() and b.legal_moves() == []
    b = CE.ChessBoard("r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1")
    assert b.push_san("O-O") == "e1g1"
    assert b.push_san("0-0-0") == "e8c8"
    assert b.fen() == "2kr3r/8/8/8/8/8/8/R4RK1 w - - 2 2"
    b = CE.ChessBoard()
    for bad in ["e5", "Ke2", "Nf4", "e4e5", "", "hello", "O-O", "exd5", "e9", "Pe4"]:
        with pytest.raises(CE.IllegalMove):
            b.push_san(bad)
    assert b.fen() == CE.START_FEN                           # nothing was pushed
    assert b.push_san("--") is None and b.fen() == CE.START_FEN
    assert b.push_san("Nf3+") == "g1f3"                      # a spurious check mark is tolerated by the grammar
    assert b.push_san("xe5") == "e7e5"                       # ... and so is a capture mark on a quiet move ([-x]? carries no meaning)
    assert sorted(CE.ChessBoard().legal_moves())[:3] == ["a2a3", "a2a4", "b1a3"]


def test_game_end_conditions():
    S = CE.ChessBoard
    st = S("7k/5Q2/6K1/8/8/8/8/8 b - - 0 1")
    assert st.status() & S.STALEMATE and st.is_game_over() and not st.is_checkmate()
    assert S("8/8/8/4k3/8/8/8/4K3 w - - 0 1").status() & S.INSUFFICIENT
    assert S("8/8/8/4k3/8/8/8/4KN2 w - - 0 1").status() & S.INSUFFICIENT
    assert S("8/8/8/4k1b1/8/8/8/2B1K3 w - - 0 1").status() & S.INSUFFICIENT            # bishops on the same colour
    assert not S("8/8/8/4kb2/8/8/8/2B1K3 w - - 0 1").status() & S.INSUFFICIENT         # opposite colours can still mate
    assert not S("8/8/8/4k3/8/8/8/4KR2 w - - 0 1").is_game_over()
    assert S("8/8/8/4k3/8/8/8/4KR2 w - - 150 90").status() & S.SEVENTYFIVE
    assert not S("8/8/8/4k3/8/8/8/4KR2 w - - 149 90").is_game_over()
    b = S()
    for _ in range(4):
        play(b, ["Nf3", "Nf6", "Ng1", "Ng8"])
        assert bool(b.status() & S.FIVEFOLD) == (_ == 3)                                 # the start position has now occurred five times
    assert b.is_game_over()


def test_chess_env_step_semantics():
    # mate in one: reward 1, done
    env = CE.ChessEnv(from_position="6k1/5ppp/8/8/8/8/8/R3K3 w - - 0 1", random_opponent=True, rng=np.random.default_rng(0))
    st, r, d, info = env.step("Qa8")
    assert (r, d) == (-1, 0) and st == env.starting_position                              # illegal: -1 and the game goes on
    st, r, d, info = env.step("--")
    assert (r, d) == (-1, True)                                                           # null move ends the episode
    st, r, d, info = env.step("Ra8#")
    assert (r, d) == (1, 1) and st.startswith("R5k1/")
    # a quiet move: the opponent answers, reward 0
    env = CE.ChessEnv(from_position="6k1/5ppp/8/8/8/8/8/R3K3 w - - 0 1", random_opponent=True, rng=np.random.default_rng(1))
    st, r, d, info = env.step("Ra2")
    assert (r, bool(d)) == (0, False) and info["opponent move"] and st.split()[1] == "w" and st.split()[5] == "2"
    # stalemating the opponent: a draw is reward 0, done
    env = CE.ChessEnv(from_position="7k/8/5QK1/8/8/8/8/8 w - - 0 1", random_opponent=True)
    st, r, d, info = env.step("Qf7")
    assert (r, d) == (0, 1)
    # the opponent mates the agent: reward -1
    class Mater:
        def best_move(self, board, start_fen, moves):
            return "d8h4"
    b_env = CE.ChessEnv(from_position="rnbqkbnr/pppp1ppp/8/4p3/8/5P2/PPPPP1PP/RNBQKBNR w KQkq - 0 2", opponent=Mater())
    st, r, d, info = b_env.step("g4")
    assert (r, d) == (-1, True) and info["opponent move"] == "Qh4#"


def test_fen_history_env_and_eval_loop():
    fen = "6k1/5ppp/8/8/8/8/8/R3K3 w - - 0 1"

    class MateInOne(TextPolicy):
        def act(self, text_history):
            assert CE.postprocess_state(text_history[-1].text) == fen
            return text_history + (Text(CE.preprocess_move("Ra8#"), True),)

    env = CE.FenChessHistoryEnv(from_position=fen, random_opponent=True)
    obs = env.reset()
    assert obs == (Text(" ".join(fen), False),)
    inter, summary = CE.text_env_eval_chess_positions([fen, fen], MateInOne(), n_rollouts=3, bsize=2, verbose=False, random_opponent=True)
    assert len(inter) == 6 and summary["victories"]["mean"] == 1.0 and summary["reward"]["mean"] == 1.0 and summary["episode_length"]["max"] == 1
    assert summary["percent_illegals"]["max"] == 0.0

    class Illegal(TextPolicy):
        def act(self, text_history):
            return text_history + (Text("z z\n", True),)
    inter, summary = CE.text_env_eval_chess_positions([fen], Illegal(), n_rollouts=1, bsize=1, verbose=False, random_opponent=True, max_moves=4)
    assert summary["episode_length"]["mean"] == 5 and summary["reward"]["mean"] == -5.0 and summary["victories"]["mean"] == 0
    st = CE.FenChessHistoryEnvSingleTurn((Text("Play white.\n", False),), from_position=fen, random_opponent=True)
    assert st.reset() == (Text("Play white.\n", False), Text(" ".join(fen) + "\n", False))


def test_random_endgames_are_valid():
    rng = np.random.default_rng(0)
    for pieces in ("kQK", "kRK", "kqKQ", "kPK"):
        for _ in range(10):
            fen = CE.large_piece_random_endgame(pieces, rng)
            b = CE.ChessBoard(fen)
            assert b.is_valid() and not b.is_check() and b.turn_white
            placed = [c for c in fen.split()[0] if c.isalpha()]
            assert sorted(placed) == sorted(pieces)


def test_uci_opponent_against_a_fake_engine():
    engine = [sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "fake_uci_engine.py")]
    opp = CE.UciOpponent(engine, nodes=1)
    try:
        env = CE.ChessEnv(from_position="6k1/5ppp/8/8/8/8/8/R3K3 w - - 0 1", opponent=opp)
        st, r, d, info = env.step("Ra2")
        assert (r, bool(d)) == (0, False) and info["opponent move"] == "f5"             # smallest legal uci move of black: f7f5
        st, r, d, info = env.step("Ra8+")
        assert info["opponent move"] == "Kf7" and st.split()[1] == "w"
    finally:
        opp.close()


def test_batched_env_with_engine_pool_equals_the_per_game_env():
    """BatchedFenChessHistoryEnv + UciOpponentPool (replies of a batch step fetched concurrently) gives, game by game, what
    TextEnvToBatchedTextEnv(FenChessHistoryEnv) with one engine gives: observations, rewards, done flags."""
    from lmrl_gym_b200.environment import BatchedTextPolicy
    engine = [sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "fake_uci_engine.py")]
    fen = "6k1/5ppp/8/8/8/8/8/R3K3 w - - 0 1"
    scripts = [["Ra2", "Ra8+", "zz", "Rb8"], ["Ra8#"], ["--"], ["Kd2", "Kc3", "Ra7", "Ra8+"], ["e4", "Ra3", "Ra4", "Ra5"]]

    class Scripted(BatchedTextPolicy):
        def __init__(self):
            self.t = 0

        def act(self, text_history, done=None):
            out = []
            for i, (h, d) in enumerate(zip(text_history, done)):
                out.append(None if d else tuple(h) + (Text(CE.preprocess_move(scripts[i][min(self.t, len(scripts[i]) - 1)]), True),))
            self.t += 1
            return out

    pool = CE.UciOpponentPool(engine, n_workers=3, nodes=1)
    single = CE.UciOpponent(engine, nodes=1)
    try:
        a = interact_environment(CE.BatchedFenChessHistoryEnv(max_moves=3, from_position=fen, opponent=pool), Scripted(), env_seed=[None] * 5, env_options=[None] * 5, bsize=5)
        b = interact_environment(CE.FenChessHistoryEnv(max_moves=3, from_position=fen, opponent=single), Scripted(), env_seed=[None] * 5, env_options=[None] * 5, bsize=5)
    finally:
        pool.close()
        single.close()
    assert [len(ep) for ep in a] == [len(ep) for ep in b] == [4, 1, 1, 4, 4]
    for ea, eb in zip(a, b):
        for ta, tb in zip(ea, eb):
            assert ta.post_transition_history == tb.post_transition_history and ta.reward == tb.reward and bool(ta.done) == bool(tb.done)
    assert a[1][0].reward == 1 and a[2][0].reward == -1 and a[4][0].reward == -1          # mate in one; null move; illegal first move


STOCKFISH = os.environ.get("LMRL_STOCKFISH", "/root/reference/stockfish/stockfish-ubuntu-20.04-x86-64-avx2")
needs_stockfish = pytest.mark.skipif(not os.path.exists(STOCKFISH), reason="the reference's vendored Stockfish binary is only present in the build container")


def _san(board, uci):
    """SAN of a legal move given in UCI notation (push, read ``board.san``, pop)."""
    san = board.push_uci(uci)
    board.pop()
    return san


class _StockfishProbe(CE.UciOpponent):
    """The real engine as an independent chess-rules oracle: `d` prints its FEN of the current position, `go perft n` its leaf count."""

    def fen_after(self, start_fen, moves):
        self._send(f"position fen {start_fen}" + (" moves " + " ".join(moves) if moves else ""))
        self._send("d")
        return self._wait("Fen:").split("Fen:")[1].strip()

    def perft(self, fen, depth):
        self._send(f"position fen {fen}")
        self._send(f"go perft {depth}")
        return int(self._wait("Nodes searched:").split(":")[1])


@needs_stockfish
def test_real_stockfish_opponent_and_rules_cross_check():
    """chess/env/env.py:157-170 with the engine the reference vendors (stockfish/stockfish-ubuntu-20.04-x86-64-avx2; round 1 only ever
    drove a fake engine).  Random-endgame games against it: every engine reply is legal under this repo's C++ rules, the FEN this repo
    emits after every ply equals Stockfish's own FEN of the same move list (piece placement, side, castling, en passant as python-chess /
    Stockfish print it, clocks), perft counts agree, and a K+Q vs K game is won by the engine's side when it plays both colours."""
    rng = np.random.default_rng(3)
    sf = _StockfishProbe(STOCKFISH, nodes=400, options={"Threads": 1})
    try:
        n_plies = 0
        for g in range(6):
            fen = CE.large_piece_random_endgame(["kQK", "kRK", "kRRK", "kQRK", "kBNK", "kQBK"][g], rng)
            assert sf.perft(fen, 3) == CE.ChessBoard(fen).perft(3), fen
            env = CE.ChessEnv(from_position=fen, opponent=sf)
            done = False
            for ply in range(12):
                legal = env.board.legal_moves()
                move = _san(env.board, legal[int(rng.integers(len(legal)))])
                st, r, done, info = env.step(move)
                n_plies += 1
                assert sf.fen_after(fen, env.uci_moves) == env.board.fen(), (fen, env.uci_moves)
                if done:
                    break
                assert info["opponent move"] is not None and st.split()[1] == "w"
        assert n_plies >= 20
        # the engine against itself from a won position: white (K+Q) mates within the move budget, this repo's rules see the mate
        b = CE.ChessBoard("8/8/8/4k3/8/8/3QK3/8 w - - 0 1")
        moves = []
        strong = CE.UciOpponent(STOCKFISH, nodes=20000, options={"Threads": 1})
        try:
            for _ in range(80):
                mv = strong.best_move(b, "8/8/8/4k3/8/8/3QK3/8 w - - 0 1", moves)
                b.push_uci(mv)
                moves.append(mv)
                if b.is_game_over():
                    break
        finally:
            strong.close()
        assert b.is_checkmate() and not b.turn_white, (len(moves), b.fen())
    finally:
        sf.close()


@needs_stockfish
def test_engine_pool_with_real_stockfish_overlaps_replies_with_the_policy():
    """BASELINE.json configs[3] on the host side: a pool of real Stockfish processes behind `BatchedFenChessHistoryEnv`, driven by the
    two-half-batch pipelined loop (`interact_environment_overlapped`): while the engines search the replies of one half, the policy acts
    on the other.  Node-limited search makes the games deterministic, so the pipelined loop must reproduce the plain loop's episodes; the
    overlap is observed directly (a policy call starts while an env step of the other half is still running)."""
    import threading
    import time
    from lmrl_gym_b200.environment import BatchedTextPolicy
    fen = "8/8/8/4k3/8/8/3QK3/8 w - - 0 1"
    events = []

    class FirstLegal(BatchedTextPolicy):
        """Stand-in for the GPU policy: the lexicographically first legal SAN move, 30 ms per call (a decode call's order of magnitude)."""

        def act(self, text_history, done=None):
            events.append(("act_start", time.perf_counter(), threading.get_ident()))
            time.sleep(0.03)
            out = []
            for h, d in zip(text_history, done):
                if d or h is None:
                    out.append(None)
                    continue
                b = CE.ChessBoard(CE.postprocess_state(h[-1].text))
                out.append(tuple(h) + (Text(CE.preprocess_move(_san(b, sorted(b.legal_moves())[0])), True),))
            events.append(("act_end", time.perf_counter(), threading.get_ident()))
            return out

    class Timed(CE.BatchedFenChessHistoryEnv):
        def step(self, text_history, done=None):
            events.append(("step_start", time.perf_counter(), threading.get_ident()))
            r = super().step(text_history, done)
            events.append(("step_end", time.perf_counter(), threading.get_ident()))
            return r

    pool = CE.UciOpponentPool(STOCKFISH, n_workers=4, nodes=3000, options={"Threads": 1}, clear_hash=True)
    try:
        plain = interact_environment(CE.BatchedFenChessHistoryEnv(max_moves=5, from_position=fen, opponent=pool), FirstLegal(), env_seed=[None] * 8, env_options=[None] * 8, bsize=8)
        events.clear()
        over = CE.interact_environment_overlapped(lambda: Timed(max_moves=5, from_position=fen, opponent=pool), FirstLegal(), n=8)
    finally:
        pool.close()
    assert len(over) == len(plain) == 8
    for ea, eb in zip(over, plain):
        assert len(ea) == len(eb)
        for ta, tb in zip(ea, eb):
            assert ta.post_transition_history == tb.post_transition_history and ta.reward == tb.reward and bool(ta.done) == bool(tb.done)
    # some policy call began while an env step (other half, other thread) was in flight
    open_steps, overlapped = 0, False
    for kind, t, _ in sorted(events, key=lambda e: e[1]):
        if kind == "step_start":
            open_steps += 1
        elif kind == "step_end":
            open_steps -= 1
        elif kind == "act_start" and open_steps > 0:
            overlapped = True
    assert overlapped
