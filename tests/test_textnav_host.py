"""Text-Nav: the oracle world model and the engine's per-env logic (csrc/textnav_core.h compiled for the host)
against the transcripts the reference ships (tests/golden/textnav_transcripts.json, from text_nav/notebooks/play_game.ipynb), and
against each other on seeded random command walks (admissible lists, score, done after every command)."""
import ctypes as C
import json
import os
import random

import numpy as np
import pytest

from oracle.textnav import CHOICE_ORDER, ROOMS, TextNavOracle

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "textnav_transcripts.json")))


def infer_start(first_adm):
    hits = [r for r in ROOMS if TextNavOracle().reset(start_room=r) == first_adm]
    assert len(hits) == 1, hits
    return hits[0]


@pytest.fixture(scope="module")
def hostlib():
    from lmrl_gym_b200 import build
    lib = C.CDLL(build.build_host())
    lib.hc_textnav_episode.restype = C.c_int
    lib.hc_textnav_start_room.restype = C.c_int
    lib.hc_textnav_start_room.argtypes = [C.c_uint64]
    return lib


def host_episode(lib, start_room, cmds):
    blob = b"".join(c.encode() for c in cmds)
    lens = np.asarray([len(c.encode()) for c in cmds], np.int32)
    n = len(cmds)
    applied = np.zeros(max(n, 1), np.int32)
    states = np.zeros((n + 1, 12), np.int32)
    adm = np.full((n + 1, 32), -1, np.int32)
    n_adm = np.zeros(n + 1, np.int32)
    buf = np.frombuffer(blob + b"\0", np.uint8).copy()
    lib.hc_textnav_episode(C.c_int(ROOMS.index(start_room)), buf.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p), C.c_int(n),
                           applied.ctypes.data_as(C.c_void_p), states.ctypes.data_as(C.c_void_p), adm.ctypes.data_as(C.c_void_p), n_adm.ctypes.data_as(C.c_void_p))

    def strings(k):
        out = []
        for code in adm[k, :n_adm[k]]:
            b = np.zeros(64, np.uint8)
            m = lib.hc_textnav_cmd_string(C.c_int(int(code)), b.ctypes.data_as(C.c_void_p))
            out.append(bytes(b[:m]).decode())
        return out
    return applied, states, [strings(k) for k in range(n + 1)]


@pytest.mark.parametrize("gi", range(len(G["games"])))
def test_oracle_matches_reference_transcripts(gi):
    game = G["games"][gi]
    o = TextNavOracle()
    assert o.reset(start_room=infer_start(game["first"]["admissible"])) == game["first"]["admissible"]
    for st in game["steps"]:
        outcome, reward, done = o.step(st["command"])
        assert (outcome != "unknown") == st["known"], st
        assert o.admissible() == st["admissible"], (st["command"], o.admissible())
        assert done == st["done"] and reward == float(st["score"])


@pytest.mark.parametrize("gi", range(len(G["games"])))
def test_engine_logic_matches_reference_transcripts(hostlib, gi):
    game = G["games"][gi]
    start = infer_start(game["first"]["admissible"])
    applied, states, adm = host_episode(hostlib, start, [s["command"] for s in game["steps"]])
    assert adm[0] == game["first"]["admissible"]
    for k, st in enumerate(game["steps"]):
        assert adm[k + 1] == st["admissible"], (st["command"], adm[k + 1])
        assert bool(states[k + 1, 10]) == st["done"] and int(states[k + 1, 9]) == st["score"]
        assert (applied[k] != -1) == st["known"]


def random_walk(rng, n):
    o = TextNavOracle()
    cmds = []
    every = sorted(TextNavOracle.all_commands())
    junk = ["", " ", "go", "go up", "open  fridge", "Open fridge", "take stale food from the table", "look", "inventory", "examine fridge", "close gridge", "eat bowl",
            " go west \n", "\tgo east", "insert fruit into table", "take   plate"]
    return o, cmds, every, junk


@pytest.mark.parametrize("seed", range(6))
def test_engine_logic_matches_oracle_on_random_walks(hostlib, seed):
    rng = random.Random(seed)
    start = rng.choice(ROOMS)
    o = TextNavOracle()
    o.reset(start_room=start)
    every = sorted(TextNavOracle.all_commands())
    junk = ["", " ", "go", "go up", "open  fridge", "Open fridge", "take stale food from the table", "look", "inventory", "examine fridge", "close gridge", "eat bowl",
            " go west \n", "\tgo east", "insert fruit into table", "take   plate"]
    cmds, want = [], []
    for _ in range(400):
        r = rng.random()
        avoid_end = [c for c in o.admissible() if c not in ("eat stale food",) and not (c == "close fridge" and o.where["stale food"] == ("in", "fridge"))]
        cmd = rng.choice(avoid_end) if r < 0.7 and avoid_end else (rng.choice(every) if r < 0.9 else rng.choice(junk))
        if rng.random() < 0.01:
            cmd = rng.choice(o.admissible())          # now and then allow the game to end
        outcome, reward, done = o.step(cmd)
        cmds.append(cmd)
        want.append((outcome, reward, done, o.admissible(), o.room, o.moves))
    applied, states, adm = host_episode(hostlib, start, cmds)
    for k, (outcome, reward, done, oadm, room, moves) in enumerate(want):
        code = {"ok": None, "unknown": -1, "impossible": -2, "over": -3}[outcome]
        assert (applied[k] >= 0) if code is None else (applied[k] == code), (k, cmds[k], applied[k], outcome)
        assert adm[k + 1] == oadm, (k, cmds[k])
        assert ROOMS[states[k + 1, 0]] == room and int(states[k + 1, 9]) == int(reward) and bool(states[k + 1, 10]) == done and int(states[k + 1, 11]) == moves


def test_start_room_follows_cpython_random(hostlib):
    for seed in list(range(50)) + [2 ** 40 + 3, 2 ** 63 + 5]:
        assert ROOMS[hostlib.hc_textnav_start_room(C.c_uint64(seed))] == random.Random(seed).choice(CHOICE_ORDER)
