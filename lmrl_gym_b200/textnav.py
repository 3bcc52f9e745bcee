"""Text-Nav on the device engine — SURVEY.md §8(f) rank 4.

``B200TextNavEnv`` is a ``BatchedTextEnv`` with the interface of the reference's ``TextNavEnv``
(llm_rl_scripts/text_nav/env/env.py:13-72: ``display_location`` / ``display_inventory``, observation = game feedback +
``"\\nAdmissible commands: a, b, c\\n"`` with the examine / look / inventory commands removed, reward = the game's score, done)
over the game of env/make_game.py:30-107.  Rooms, furniture, objects, commands, quests and the admissible-command lists live on the
GPU as 12 int32 per environment (csrc/textnav_core.h, ``lmrl_textnav_reset_host`` / ``lmrl_textnav_step_host``); this module only
renders text from the codes the device returns.

Parity: the world logic is pinned on the transcripts the reference ships (tests/test_textnav_host.py).  The room PROSE of the
reference is TextWorld grammar output (random templates, engine not vendored): the sentences below are this repository's own,
deterministic, and follow the information content of TextWorld's (what is here, what is on / in it, the exits).  Two defects of the
reference's ``step`` are not reproduced: it raises NameError (``true``, env.py:71) and would append the command a second time.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from ._lib import check, ptr
from .engine import Engine
from .environment import BatchedTextEnv, Text, TextHistory
from .wordle import encode_actions, seed_to_u64

ROOMS = ["bedroom", "office", "bathroom", "living room", "kitchen", "dining room", "garden", "backyard"]          # make_game.py:33-40
OBJECTS = ["stale food", "fresh food", "bowl", "coffee cup", "plate", "utensils", "fruit"]                        # :66-72
DIRS = ["east", "north", "south", "west"]
CONTAINERS = ["chest", "fridge"]
_VERBS = ["close", "drop", "eat", "go", "insert", "open", "put", "take", "noop"]
_TAKE_FROM = ["", " from chest", " from fridge", " from table"]
_TABLE_ROOM = {0: 0, 3: 1, 5: 2}
_CONTAINER_ROOM = {1: 0, 4: 1}
OBJECTIVE = "Your objective is to put the stale food on the table in the living room into the fridge in the kitchen."   # make_game.py:118-119


def command_string(code: int) -> str:
    verb, a, b = _VERBS[code >> 8], (code >> 4) & 15, code & 15
    if verb == "go":
        return "go " + DIRS[a]
    if verb in ("open", "close"):
        return f"{verb} {CONTAINERS[a]}"
    if verb in ("drop", "eat"):
        return f"{verb} {OBJECTS[a]}"
    if verb == "put":
        return f"put {OBJECTS[a]} on table"
    if verb == "insert":
        return f"insert {OBJECTS[a]} into {CONTAINERS[b]}"
    if verb == "take":
        return f"take {OBJECTS[a]}" + _TAKE_FROM[b]
    return ""


def _listing(names: List[str]) -> str:
    names = [n if n == "utensils" else "a " + n for n in names]          # "a stale food and a fresh food", as TextWorld prints them
    return names[0] if len(names) == 1 else ", ".join(names[:-1]) + " and " + names[-1]


def describe_room(state: np.ndarray, display_location: bool) -> str:
    """This repository's room description (the reference's is TextWorld grammar text)."""
    room, loc, open_mask = int(state[0]), state[1:8], int(state[8])
    parts = []
    if display_location:
        parts.append(f"-= {ROOMS[room].title()} =-")
    if room in _TABLE_ROOM:
        on = [OBJECTS[o] for o in range(7) if loc[o] == 8 + _TABLE_ROOM[room]]
        parts.append("You see a table." + (f" On the table you see {_listing(on)}." if on else " The table is empty."))
    if room in _CONTAINER_ROOM:
        c = _CONTAINER_ROOM[room]
        if (open_mask >> c) & 1:
            inside = [OBJECTS[o] for o in range(7) if loc[o] == 11 + c]
            parts.append(f"You see a {CONTAINERS[c]}. The {CONTAINERS[c]} is open" + (f" and contains {_listing(inside)}." if inside else " and empty."))
        else:
            parts.append(f"You see a closed {CONTAINERS[c]}.")
    floor = [OBJECTS[o] for o in range(7) if loc[o] == room]
    if floor:
        parts.append(f"On the floor you see {_listing(floor)}.")
    if len(parts) == (1 if display_location else 0):
        parts.append("The room is empty.")
    return " ".join(parts)


def exits_sentence(adm: List[str]) -> str:
    ds = [c[3:] for c in adm if c.startswith("go ")]
    return "There are exits to the " + (ds[0] if len(ds) == 1 else ", ".join(ds[:-1]) + " and " + ds[-1]) + "."


def response_sentence(code: int, prev: np.ndarray, state: np.ndarray) -> str:
    """First line of the game's answer to a command, in the wording of the reference transcripts (notebooks/play_game.ipynb)."""
    if code == -1:
        return "You can't see any such thing."
    if code == -2:
        return "You can't do that."
    verb, a, b = _VERBS[code >> 8], (code >> 4) & 15, code & 15
    if verb == "open":
        inside = [OBJECTS[o] for o in range(7) if state[1 + o] == 11 + a]
        return f"You open the {CONTAINERS[a]}" + (f", revealing {_listing(inside)}." if inside else ".")
    if verb == "close":
        return f"You close the {CONTAINERS[a]}."
    if verb == "take":
        return f"You pick up the {OBJECTS[a]} from the ground." if b == 0 else f"You take the {OBJECTS[a]} from the {['', 'chest', 'fridge', 'table'][b]}."
    if verb == "drop":
        return f"You drop the {OBJECTS[a]} on the ground."
    if verb == "eat":
        return f"You eat the {OBJECTS[a]}. Not bad."
    if verb == "put":
        return f"You put the {OBJECTS[a]} on the table."
    if verb == "insert":
        return f"You put the {OBJECTS[a]} into the {CONTAINERS[b]}."
    return ""


class B200TextNavEnv(BatchedTextEnv):
    def __init__(self, engine: Engine, display_location: bool = False, display_inventory: bool = False):
        self._lmrl_engine = engine
        self._lmrl_fusable = False          # text is rendered on the host: this env is not part of the fused rollout
        self.display_location = bool(display_location)
        self.display_inventory = bool(display_inventory)
        self.n: Optional[int] = None
        self.state: Optional[np.ndarray] = None

    def _observation(self, body: str, state: np.ndarray, adm: List[str]) -> str:
        text = body
        if self.display_inventory:                                              # env.py:40-43
            inv = [OBJECTS[o] for o in range(7) if state[1 + o] == 13]
            text += "\n" + ("You are carrying: " + _listing(inv) + "." if inv else "You are carrying nothing.") + "\n"
        return text + "\nAdmissible commands: {}\n".format(", ".join(adm))     # env.py:44-45

    def reset(self, seed: Optional[List[Optional[int]]] = None, options: Optional[List[Optional[Dict]]] = None) -> List[TextHistory]:
        if seed is None and options is None:
            seed = [None]
        elif seed is None:
            seed = [None] * len(options)
        n = len(seed)
        seeds = np.asarray([seed_to_u64(s) for s in seed], dtype=np.uint64)
        start = np.full(n, -1, np.int32)
        for i, o in enumerate(options or []):
            if o and o.get("start_room") is not None:
                start[i] = ROOMS.index(o["start_room"]) if isinstance(o["start_room"], str) else int(o["start_room"])
        e = self._lmrl_engine
        self.state = np.zeros((n, 12), np.int32)
        adm, n_adm = np.zeros((n, 32), np.int32), np.zeros(n, np.int32)
        check(e.lib.lmrl_textnav_reset_host(e.handle, ptr(seeds), ptr(start), n, ptr(self.state), ptr(adm), ptr(n_adm)))
        self.n = n
        out = []
        for i in range(n):
            cmds = [command_string(int(c)) for c in adm[i, :n_adm[i]]]
            body = OBJECTIVE + "\n\n" + describe_room(self.state[i], self.display_location) + "\n\n" + exits_sentence(cmds) + "\n"
            out.append((Text(self._observation(body, self.state[i], cmds), False),))
        return out

    def step(self, text_history: List[Optional[TextHistory]], done: Optional[List[bool]] = None):
        assert self.n is not None, "reset must be called before step"
        assert len(text_history) == self.n, "batch size must be the same as the number of environments initalized"
        if done is None:
            done = [False] * self.n
        n = self.n
        actions: List[Optional[str]] = []
        for h, d in zip(text_history, done):
            if d or h is None:
                actions.append(None)
            else:
                assert h[-1].is_action
                actions.append(h[-1].text.strip())          # env.py:38 (the device strips too; doing it here keeps long padded commands intact)
        skip = np.asarray([a is None for a in actions], dtype=np.uint8)
        arr, lens = encode_actions(actions)
        reward, dn = np.zeros(n, np.float32), np.zeros(n, np.uint8)
        applied, adm, n_adm = np.zeros(n, np.int32), np.zeros((n, 32), np.int32), np.zeros(n, np.int32)
        prev = self.state.copy()
        e = self._lmrl_engine
        check(e.lib.lmrl_textnav_step_host(e.handle, ptr(arr), ptr(lens), arr.shape[1], ptr(skip), n, ptr(reward), ptr(dn), ptr(applied), ptr(self.state),
                                           ptr(adm), ptr(n_adm)))
        out = []
        for i in range(n):
            if actions[i] is None:
                out.append(None)
                continue
            cmds = [command_string(int(c)) for c in adm[i, :n_adm[i]]]
            code = int(applied[i])
            if code >= 0 and _VERBS[code >> 8] == "go":
                body = "\n" + describe_room(self.state[i], self.display_location) + "\n\n" + exits_sentence(cmds) + "\n"
            elif code >= 0 and _VERBS[code >> 8] == "noop":
                body = describe_room(self.state[i], self.display_location) + "\n"
            else:
                body = response_sentence(code, prev[i], self.state[i]) + "\n"
            if dn[i]:
                body += "\n" + ("Your score has just gone up by one point.\n\n*** The End ***" if reward[i] > 0 else "*** You lost! ***") + \
                        f"\n\nYou scored {int(reward[i])} out of a possible 1, in {int(self.state[i, 11])} turns.\n"
            out.append((tuple(text_history[i]) + (Text(self._observation(body, self.state[i], cmds), False),), float(reward[i]), bool(dn[i])))
        return out

    def copy(self):
        return self
