"""Chess-endgames environment — SURVEY.md §8(f) rank 3, BASELINE.json configs[3].

Mirrors ``llm_rl_scripts/chess/env/env.py``: the text helpers (:13-27), ``ChessEnv`` (:29-186), ``FenChessHistoryEnv`` /
``FenChessHistoryEnvSingleTurn`` (:188-243), ``large_piece_random_endgame`` (:245-260) and ``text_env_eval_chess_positions``
(:262-345).  The reference leans on python-chess for the rules and on the ``stockfish`` package for the opponent; neither is a
dependency here: the rules are ``csrc/chess_core.h`` behind ``lmrl_chess_*`` (include/lmrl_b200.h; move generation pinned on the
published perft counts, tests/test_chess_host.py) and the opponent is anything with ``best_move(fen, moves) -> uci``:

  ``UciOpponent(path)``      a UCI engine process (Stockfish when one is installed: ``go movetime 100`` = the reference's
                             ``get_best_move_time(100)``, env.py:163; ``nodes=N`` makes it deterministic — a flagged deviation)
  ``RandomOpponent(rng)``    the reference's ``random_opponent=True`` (:172-186) with an explicit generator instead of the global numpy one

The env is host-side by construction (the opponent is a process) — what runs on the B200 is the policy (``B200GPT2Policy``), which the
generic ``interact_environment`` loop drives one batched ``act`` per ply.  Parity: rules pinned by perft known answers and by SAN / FEN
examples; python-chess itself is absent, so its corner cases (e.g. the move ORDER behind ``random_opponent``) are unpinned.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from typing import Callable, Dict, Iterator, List, Optional, Tuple, Union

import numpy as np

from . import _lib
from .environment import BatchedTextEnv, BatchedTextPolicy, InteractionTransition, Text, TextEnv, TextHistory, TextPolicy, interact_environment

START_FEN = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"


def preprocess_move(move: str):            # env.py:13-14
    return " ".join(move) + "\n"


def postprocess_move(move: str):           # env.py:16-17
    return move.replace(" ", "").strip()


def preprocess_state(state: str):          # env.py:19-20
    return " ".join(state) + "\n"


def preprocess_state_og(state: str):       # env.py:22-23
    return " ".join(state)


def postprocess_state(state: str):         # env.py:25-26
    return state.replace("  ", "__temp__").replace(" ", "").replace("__temp__", " ").strip()


class IllegalMove(ValueError):
    pass


class ChessBoard:
    """One board with its move stack (``chess.Board`` as far as the env uses it)."""

    CHECKMATE, STALEMATE, INSUFFICIENT, SEVENTYFIVE, FIVEFOLD, CHECK = 1, 2, 4, 8, 16, 32

    def __init__(self, fen: Optional[str] = None):
        self.lib = _lib.load()
        self._h = self.lib.lmrl_chess_new(None if fen is None else fen.encode())
        if not self._h:
            raise ValueError(self.lib.lmrl_last_error().decode())
        self._buf = C.create_string_buffer(4096)

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib.lmrl_chess_free(self._h)
            self._h = None

    def fen(self) -> str:
        _lib.check(min(self.lib.lmrl_chess_fen(self._h, self._buf, 4096), 0))
        return self._buf.value.decode()

    @property
    def turn_white(self) -> bool:
        return self.lib.lmrl_chess_turn(self._h) == 0

    def legal_moves(self) -> List[str]:
        n = self.lib.lmrl_chess_legal_moves(self._h, self._buf, 4096)
        _lib.check(min(n, 0))
        return self._buf.value.decode().split() if n else []

    def push_san(self, san: str) -> Optional[str]:
        """UCI string of the pushed move; None for a null move (reported, not pushed); IllegalMove otherwise (board unchanged)."""
        rc = self.lib.lmrl_chess_push_san(self._h, san.encode("utf-8", "replace"), self._buf, 4096)
        if rc < 0:
            raise IllegalMove(self.lib.lmrl_last_error().decode())
        return None if rc == 1 else self._buf.value.decode()

    def push_uci(self, uci: str) -> str:
        """SAN of the move (``board.san(move)``), then the move is pushed."""
        if self.lib.lmrl_chess_push_uci(self._h, uci.encode(), self._buf, 4096) < 0:
            raise IllegalMove(self.lib.lmrl_last_error().decode())
        return self._buf.value.decode()

    def pop(self):
        _lib.check(self.lib.lmrl_chess_pop(self._h))

    def status(self) -> int:
        return self.lib.lmrl_chess_status(self._h)

    def is_checkmate(self) -> bool:
        return bool(self.status() & self.CHECKMATE)

    def is_game_over(self) -> bool:
        return bool(self.status() & 31)

    def is_check(self) -> bool:
        return bool(self.status() & self.CHECK)

    def is_valid(self) -> bool:
        return self.lib.lmrl_chess_is_valid(self._h) == 1

    def perft(self, depth: int) -> int:
        return int(self.lib.lmrl_chess_perft(self._h, depth))


# ---- opponents -----------------------------------------------------------------------------------------------------------

class RandomOpponent:
    """``random_opponent=True`` (env.py:172-186): a uniformly random legal move."""

    def __init__(self, rng: Optional[np.random.Generator] = None):
        self.rng = rng or np.random.default_rng()

    def best_move(self, board: ChessBoard, start_fen: str, moves: List[str]) -> str:
        legal = board.legal_moves()
        return legal[int(self.rng.integers(0, len(legal)))]

    def close(self):
        pass


class UciOpponent:
    """A UCI engine process (``Stockfish(path, parameters={"Threads": 1, "UCI_Elo": elo})``, env.py:58-60; reply =
    ``get_best_move_time(100)``, :163).  ``nodes`` replaces the wall-clock budget by a node budget (deterministic; a deviation)."""

    def __init__(self, path: str, movetime_ms: int = 100, nodes: Optional[int] = None, options: Optional[Dict[str, object]] = None,
                 clear_hash: bool = False):
        """``clear_hash``: send ``ucinewgame`` before every search, so that with a node budget and one thread the reply is a function of
        the position alone (whichever process of a pool answers, whatever it searched before) — reproducible games for tests."""
        self.p = subprocess.Popen([path] if isinstance(path, str) else list(path), stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, bufsize=1)
        self.movetime_ms, self.nodes, self.clear_hash = movetime_ms, nodes, clear_hash
        self._send("uci")
        self._wait("uciok")
        for k, v in (options or {}).items():
            self._send(f"setoption name {k} value {v}")
        self._send("isready")
        self._wait("readyok")

    def _send(self, line: str):
        self.p.stdin.write(line + "\n")
        self.p.stdin.flush()

    def _wait(self, token: str) -> str:
        while True:
            line = self.p.stdout.readline()
            if not line:
                raise RuntimeError("UCI engine closed its output")
            if line.startswith(token):
                return line.strip()

    def best_move(self, board: ChessBoard, start_fen: str, moves: List[str]) -> str:
        if self.clear_hash:
            self._send("ucinewgame")
            self._send("isready")
            self._wait("readyok")
        self._send(f"position fen {start_fen}" + (" moves " + " ".join(moves) if moves else ""))
        self._send(f"go nodes {self.nodes}" if self.nodes else f"go movetime {self.movetime_ms}")
        return self._wait("bestmove").split()[1]

    def close(self):
        if self.p and self.p.poll() is None:
            try:
                self._send("quit")
                self.p.wait(timeout=2)
            except Exception:
                self.p.kill()
        self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ChessEnv:
    """env.py:29-186 — the agent plays White from ``from_position``; states are FEN strings."""

    def __init__(self, side="w", fen=True, from_position=None, opponent=None, random_opponent=False, rng: Optional[np.random.Generator] = None):
        assert side == "w" and fen, "the reference env only supports the FEN state for White (env.py:109, 147-148)"
        self.starting_position = from_position if from_position is not None else START_FEN
        self.opponent = opponent if opponent is not None else (RandomOpponent(rng) if random_opponent else None)
        self.reset()

    def reset(self):
        self.board = ChessBoard(self.starting_position)
        if self.starting_position == START_FEN:
            self.starting_position = self.board.fen()
        self.prev_moves: List[str] = []
        self.uci_moves: List[str] = []
        return self.starting_position, {}

    def step(self, action: str, opponent_move: bool = True):
        """Reward 0 for legal moves, -1 for illegal moves and defeat, 1 for victory, 0 for draws (env.py:93-140)."""
        self.prev_moves.append(action)
        assert self.board.turn_white
        reward, done, opponent = -1, 0, None
        try:
            uci = self.board.push_san(action)
        except IllegalMove:
            return self.board.fen(), reward, done, {"opponent move": opponent}          # illegal: -1, the game goes on
        if uci is None:                                                                  # null move: -1 and the episode ends (:116-118)
            return self.board.fen(), -1, True, {}
        self.uci_moves.append(uci)
        if self.board.is_game_over():
            reward, done = (1 if self.board.is_checkmate() else 0), 1
        elif opponent_move:
            assert self.opponent is not None, "no opponent configured (pass opponent=UciOpponent(path) or random_opponent=True)"
            reply = self.opponent.best_move(self.board, self.starting_position, self.uci_moves)
            opponent = self.board.push_uci(reply)
            self.uci_moves.append(reply)
            reward = -1 if self.board.is_checkmate() else 0
            done = self.board.is_game_over()
        else:
            reward, done = 0, 0
        return self.board.fen(), reward, done, {"opponent move": opponent}


class FenChessHistoryEnv(TextEnv):
    """env.py:215-243: the observation is the current FEN with its characters spaced out; the history is NOT accumulated."""

    def __init__(self, max_moves=400, from_position=None, random_opponent=False, opponent=None, rng: Optional[np.random.Generator] = None):
        self.from_position, self.max_moves, self.random_opponent, self._opponent, self._rng = from_position, max_moves, random_opponent, opponent, rng
        self.chess_env = ChessEnv(fen=True, from_position=from_position, random_opponent=random_opponent, opponent=opponent, rng=rng)

    def reset(self, seed: Optional[int] = None, options: Optional[Dict] = None):
        self.init_state, _ = self.chess_env.reset()
        self.num_moves_made = 0
        return (Text(preprocess_state_og(self.init_state), False),)

    def step(self, text_history: TextHistory):
        assert text_history[-1].is_action
        st, reward, done, _ = self.chess_env.step(postprocess_move(text_history[-1].text))
        self.num_moves_made += 1
        if self.num_moves_made > self.max_moves:
            done = 1
        return (Text(preprocess_state_og(st), False),), reward, done

    def copy(self):
        return FenChessHistoryEnv(self.max_moves, self.from_position, self.random_opponent, self._opponent, self._rng)

    def __deepcopy__(self, memo):          # TextEnvToBatchedTextEnv deep-copies the env per slot (environment.py:92); boards / processes are not copyable
        return self.copy()


class FenChessHistoryEnvSingleTurn(TextEnv):
    """env.py:188-213: every observation is ``initial_history`` + the current FEN line."""

    def __init__(self, initial_history: TextHistory, max_moves=400, from_position=None, opponent=None, random_opponent=False, rng=None):
        self.initial_history, self.max_moves, self.from_position = tuple(initial_history), max_moves, from_position
        self._opponent, self._random, self._rng = opponent, random_opponent, rng
        self.chess_env = ChessEnv(fen=True, from_position=from_position, opponent=opponent, random_opponent=random_opponent, rng=rng)

    def reset(self, seed: Optional[int] = None, options: Optional[Dict] = None):
        self.init_state, _ = self.chess_env.reset()
        self.num_moves_made = 0
        return self.initial_history + (Text(preprocess_state(self.init_state), False),)

    def step(self, text_history: TextHistory):
        assert text_history[-1].is_action
        st, reward, done, _ = self.chess_env.step(postprocess_move(text_history[-1].text))
        self.num_moves_made += 1
        if self.num_moves_made > self.max_moves:
            done = 1
        return self.initial_history + (Text(preprocess_state(st), False),), reward, done

    def copy(self):
        return FenChessHistoryEnvSingleTurn(self.initial_history, self.max_moves, self.from_position, self._opponent, self._random, self._rng)

    def __deepcopy__(self, memo):
        return self.copy()


_PIECE_ORDER = "PNBRQKpnbrqk"


def large_piece_random_endgame(pieces: str, rng: Optional[np.random.Generator] = None) -> str:
    """env.py:245-260: drop the pieces of e.g. 'kQK' on random distinct squares until the position is valid and not a check."""
    rng = rng or np.random.default_rng()
    while True:
        squares = rng.choice(64, size=len(pieces), replace=False)
        grid = [["" for _ in range(8)] for _ in range(8)]
        for p, s in zip(pieces, squares):
            grid[int(s) // 8][int(s) % 8] = p
        rows = []
        for r in range(7, -1, -1):
            row, empty = "", 0
            for f in range(8):
                if grid[r][f]:
                    row += (str(empty) if empty else "") + grid[r][f]
                    empty = 0
                else:
                    empty += 1
            rows.append(row + (str(empty) if empty else ""))
        fen = "/".join(rows) + " w - - 0 1"
        try:
            b = ChessBoard(fen)
        except ValueError:
            continue
        if b.is_valid() and not b.is_check():
            return b.fen()


def text_env_eval_chess_positions(positions: List[str], policy: Union[TextPolicy, BatchedTextPolicy], n_rollouts: int,
                                  initial_text_history: Optional[TextHistory] = None, seed_generator: Optional[Iterator[int]] = None,
                                  env_options: Optional[Dict] = None, interaction_callback: Optional[Callable] = None, bsize: int = 1, verbose: bool = True,
                                  random_opponent: bool = False, max_moves: int = 400, opponent_factory: Optional[Callable[[], object]] = None):
    """env.py:262-345: ``n_rollouts`` games from every position; summary of reward / done / victories / percent_illegals / episode_length."""
    interactions, rs, dones, victories, percent_illegals, episode_length = [], [], [], [], [], []
    for position in positions:
        env = FenChessHistoryEnv(from_position=position, random_opponent=random_opponent, max_moves=max_moves,
                                 opponent=opponent_factory() if opponent_factory else None)
        done_here = 0
        while done_here < n_rollouts:
            actual = min(n_rollouts - done_here, bsize)
            batch = interact_environment(env, policy, initial_text_history=initial_text_history,
                                         env_seed=[None] * actual if seed_generator is None else [next(seed_generator) for _ in range(actual)],
                                         env_options=[env_options] * actual, bsize=actual, npad=bsize - actual)
            for interaction in batch:
                interactions.append(interaction)
                rewards = [x.reward for x in interaction]
                victories.append(1 if 1 in rewards else 0)
                n_illegal = sum(1 if x.reward == -1 and i < len(rewards) - 1 else 0 for i, x in enumerate(interaction))
                percent_illegals.append(n_illegal / len(rewards) * 100)
                episode_length.append(len(rewards))
                rs.append(sum(rewards))
                dones.append(interaction[-1].done)
                if interaction_callback is not None:
                    interaction_callback(interaction)
            done_here += actual
    rs, dones = np.asarray(rs, dtype=np.float32), np.asarray(dones, dtype=np.float32)
    stats = lambda a: dict(mean=np.mean(a), std=np.std(a), min=np.min(a), max=np.max(a))  # noqa: E731
    return interactions, dict(reward=stats(rs), done=stats(dones), victories=stats(victories), percent_illegals=stats(percent_illegals),
                              episode_length=stats(episode_length))


# ---- batched env with an engine pool ----------------------------------------------------------------------------------------

class UciOpponentPool:
    """``n_workers`` UCI engine processes answering the replies of one batch step concurrently (threads block on the engines' pipes, so
    the engines search in parallel on the host cores while the caller — and the GPU, busy with other work — is not held per game)."""

    def __init__(self, path, n_workers: int, **kw):
        self.engines = [UciOpponent(path, **kw) for _ in range(n_workers)]

    def best_moves(self, requests: List[Tuple[str, List[str]]]) -> List[str]:
        """requests: (start_fen, uci moves so far) per game -> the engine's reply per game, in order."""
        import queue
        import threading
        out: List[Optional[str]] = [None] * len(requests)
        q: "queue.Queue[int]" = queue.Queue()
        for i in range(len(requests)):
            q.put(i)
        errors: List[BaseException] = []

        def work(engine: UciOpponent):
            while True:
                try:
                    i = q.get_nowait()
                except queue.Empty:
                    return
                try:
                    out[i] = engine.best_move(None, requests[i][0], requests[i][1])
                except BaseException as e:      # noqa: BLE001 - re-raised in the caller
                    errors.append(e)
                    return

        threads = [threading.Thread(target=work, args=(e,)) for e in self.engines[:max(1, min(len(self.engines), len(requests)))]]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return out  # type: ignore[return-value]

    def close(self):
        for e in self.engines:
            e.close()


class BatchedFenChessHistoryEnv(BatchedTextEnv):
    """``FenChessHistoryEnv`` (env.py:215-243) for a whole batch: same observations, rewards and done flags per game as
    ``TextEnvToBatchedTextEnv(FenChessHistoryEnv(...))``, but one ``step`` first plays every agent move (C rules), then asks the engine
    pool for all the replies at once.  ``opponent`` is a ``UciOpponentPool``, any object with ``best_move(board, fen, moves)``, or None with
    ``random_opponent=True``."""

    def __init__(self, max_moves=400, from_position=None, opponent=None, random_opponent=False, rng: Optional[np.random.Generator] = None):
        self.max_moves, self.from_position = max_moves, from_position
        self.opponent = opponent if opponent is not None else (RandomOpponent(rng) if random_opponent else None)
        self.envs: Optional[List[ChessEnv]] = None

    def reset(self, seed: Optional[List[Optional[int]]] = None, options: Optional[List[Optional[Dict]]] = None) -> List[TextHistory]:
        n = len(seed) if seed is not None else (len(options) if options is not None else 1)
        self.envs = [ChessEnv(from_position=self.from_position, opponent=self.opponent) for _ in range(n)]
        self.num_moves_made = [0] * n
        return [(Text(preprocess_state_og(e.starting_position), False),) for e in self.envs]

    def step(self, text_history: List[Optional[TextHistory]], done: Optional[List[bool]] = None):
        assert self.envs is not None, "reset must be called before step"
        assert len(text_history) == len(self.envs), "batch size must be the same as the number of environments initalized"
        if done is None:
            done = [False] * len(text_history)
        results: List[Optional[Tuple]] = [None] * len(self.envs)
        pending: List[int] = []
        for i, (h, d) in enumerate(zip(text_history, done)):
            if d or h is None:
                continue
            assert h[-1].is_action
            st, reward, dn, info = self.envs[i].step(postprocess_move(h[-1].text), opponent_move=False)
            legal_and_open = (reward, dn) == (0, 0) and self.envs[i].uci_moves and not self.envs[i].board.turn_white
            if legal_and_open:
                pending.append(i)
            results[i] = (st, reward, dn)
        if pending:
            assert self.opponent is not None, "no opponent configured"
            if isinstance(self.opponent, UciOpponentPool):
                replies = self.opponent.best_moves([(self.envs[i].starting_position, self.envs[i].uci_moves) for i in pending])
            else:
                replies = [self.opponent.best_move(self.envs[i].board, self.envs[i].starting_position, self.envs[i].uci_moves) for i in pending]
            for i, reply in zip(pending, replies):
                env = self.envs[i]
                env.board.push_uci(reply)
                env.uci_moves.append(reply)
                results[i] = (env.board.fen(), -1 if env.board.is_checkmate() else 0, env.board.is_game_over())
        out: List[Optional[Tuple]] = []
        for i, r in enumerate(results):
            if r is None:
                out.append(None)
                continue
            st, reward, dn = r
            self.num_moves_made[i] += 1
            if self.num_moves_made[i] > self.max_moves:
                dn = 1
            out.append(((Text(preprocess_state_og(st), False),), reward, dn))
        return out


def interact_environment_overlapped(env_factory, policy: BatchedTextPolicy, n: int, env_seed: Optional[List[Optional[int]]] = None,
                                    env_options: Optional[List[Optional[Dict]]] = None) -> List[List[InteractionTransition]]:
    """``interact_environment`` (LLM_RL/environment.py:154-207) for a host-bound env, as a two-stage pipeline over two half-batches: while
    the env step of one half runs on a worker thread (for chess: the Stockfish pool searching the opponent replies, chess/env/env.py:157-170),
    the policy acts on the other half — the GPU decode and the engine search overlap instead of alternating.  ``env_factory()`` builds one
    ``BatchedTextEnv`` per half (both may share one engine pool).  Per-episode results are those of the plain loop: every half follows
    the reference's loop (``None`` skip rule of :192-197 included); only the interleaving changes."""
    from concurrent.futures import ThreadPoolExecutor
    assert n >= 1
    n_a = (n + 1) // 2
    spans = [(0, n_a), (n_a, n)] if n > n_a else [(0, n_a)]
    halves = []
    for lo, hi in spans:
        env = env_factory()
        seeds = [None] * (hi - lo) if env_seed is None else list(env_seed[lo:hi])
        opts = [None] * (hi - lo) if env_options is None else list(env_options[lo:hi])
        halves.append(dict(env=env, hist=env.reset(seeds, opts), done=[False] * (hi - lo), tr=[[] for _ in range(hi - lo)], fut=None, pre=None, acted=None))

    def collect(h):
        results = [(None, None, True) if r is None else r for r in h["fut"].result()]
        h["fut"] = None
        pre, acted = h["pre"], h["acted"]
        h["hist"] = [r[0] for r in results]
        reward = [r[1] for r in results]
        h["done"] = [bool(r[2]) for r in results]
        for i in range(len(results)):
            if h["done"][i] and (pre[i] is None or acted[i] is None or h["hist"][i] is None or reward[i] is None):
                continue
            h["tr"][i].append(InteractionTransition(pre[i], acted[i], h["hist"][i], reward[i], results[i][2]))

    with ThreadPoolExecutor(max_workers=1) as worker:
        k = 0
        while any(not all(h["done"]) or h["fut"] is not None for h in halves):
            h = halves[k % len(halves)]
            k += 1
            if h["fut"] is not None:
                collect(h)                                   # the step submitted one round ago (it ran while the other half was in policy.act)
            if all(h["done"]):
                continue
            h["pre"] = h["hist"]
            h["acted"] = policy.act(list(h["hist"]), done=list(h["done"]))
            h["fut"] = worker.submit(h["env"].step, h["acted"], list(h["done"]))
    return [ep for h in halves for ep in h["tr"]]
