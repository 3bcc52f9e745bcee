"""Text-Nav world model — plain-Python restatement (oracle; test infrastructure only).

Follows llm_rl_scripts/text_nav/env/make_game.py:30-107 (rooms, connections, furniture, objects, the two quests) and
env/env.py:27-72 (strip the command, reward = score, done, the "Admissible commands: ..." line).  The engine underneath the
reference is TextWorld (jxihong/TextWorld @ main, unpinned, not vendored): its rules for which commands are admissible are restated
from its published house-theme logic (take/put/insert/drop/eat/open/close/go) and PINNED on the transcripts the reference ships
(tests/golden/textnav_transcripts.json from notebooks/play_game.ipynb): every admissible list, score and done flag of both games.
Room prose is not modelled.  Deliberately written with sets/dicts of names — nothing shared with csrc/textnav_core.h's encoding.
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional, Tuple

ROOMS = ["bedroom", "office", "bathroom", "living room", "kitchen", "dining room", "garden", "backyard"]
CHOICE_ORDER = ["bedroom", "bathroom", "office", "living room", "kitchen", "dining room", "garden", "backyard"]      # make_game.py:82-83
OPPOSITE = dict(north="south", south="north", east="west", west="east")
CONNECT = [("bedroom", "west", "living room"), ("living room", "west", "dining room"), ("dining room", "west", "kitchen"), ("kitchen", "north", "garden"),
           ("living room", "south", "office"), ("office", "east", "bathroom"), ("bedroom", "south", "bathroom"), ("kitchen", "west", "backyard")]
EXITS: Dict[str, Dict[str, str]] = {r: {} for r in ROOMS}
for a, d, b in CONNECT:
    EXITS[a][d] = b
    EXITS[b][OPPOSITE[d]] = a
TABLE_ROOMS = ["bedroom", "living room", "dining room"]
CONTAINER_ROOM = dict(chest="office", fridge="kitchen")
FOOD = {"stale food", "fresh food", "fruit"}
OBJECTS = ["stale food", "fresh food", "bowl", "coffee cup", "plate", "utensils", "fruit"]


class TextNavOracle:
    def __init__(self):
        self.reset(0)

    def reset(self, seed: Optional[int] = None, start_room: Optional[str] = None):
        self.room = start_room if start_room is not None else random.Random(seed).choice(CHOICE_ORDER)
        # where: ("table", room) | ("in", container) | ("floor", room) | ("inventory",) | ("gone",)
        self.where = {"stale food": ("table", "living room"), "fresh food": ("table", "living room"), "bowl": ("table", "bedroom"),
                      "coffee cup": ("table", "dining room"), "plate": ("table", "dining room"), "utensils": ("in", "chest"), "fruit": ("in", "fridge")}
        self.open = dict(chest=False, fridge=False)
        self.score, self.done, self.moves = 0, False, 0
        return self.admissible()

    def inventory(self) -> List[str]:
        return [o for o in OBJECTS if self.where[o] == ("inventory",)]

    def admissible(self) -> List[str]:
        cmds = ["go " + d for d in EXITS[self.room]]
        for c, r in CONTAINER_ROOM.items():
            if r == self.room:
                cmds.append(("close " if self.open[c] else "open ") + c)
                if self.open[c]:
                    cmds += [f"take {o} from {c}" for o in OBJECTS if self.where[o] == ("in", c)]
                    cmds += [f"insert {o} into {c}" for o in self.inventory()]
        if self.room in TABLE_ROOMS:
            cmds += [f"take {o} from table" for o in OBJECTS if self.where[o] == ("table", self.room)]
            cmds += [f"put {o} on table" for o in self.inventory()]
        cmds += [f"take {o}" for o in OBJECTS if self.where[o] == ("floor", self.room)]
        for o in self.inventory():
            cmds.append("drop " + o)
            if o in FOOD:
                cmds.append("eat " + o)
        return sorted(cmds)

    def step(self, command: str) -> Tuple[str, float, bool]:
        """-> (outcome in {'ok', 'unknown', 'impossible', 'over'}, reward = score, done)."""
        if self.done:
            return "over", float(self.score), True
        self.moves += 1
        cmd = command.strip()
        if cmd in ("look", "inventory") or cmd.startswith("examine "):
            return "ok", float(self.score), self.done
        if cmd not in self.all_commands():
            return "unknown", float(self.score), self.done
        if cmd not in self.admissible():
            return "impossible", float(self.score), self.done
        w = cmd.split(" ")
        if w[0] == "go":
            self.room = EXITS[self.room][w[1]]
        elif w[0] in ("open", "close"):
            self.open[w[1]] = w[0] == "open"
        else:
            obj = next(o for o in sorted(OBJECTS, key=len, reverse=True) if cmd[len(w[0]) + 1:].startswith(o))
            if w[0] == "take":
                self.where[obj] = ("inventory",)
            elif w[0] == "drop":
                self.where[obj] = ("floor", self.room)
            elif w[0] == "eat":
                self.where[obj] = ("gone",)
            elif w[0] == "put":
                self.where[obj] = ("table", self.room)
            elif w[0] == "insert":
                self.where[obj] = ("in", w[-1])
        if self.where["stale food"] == ("gone",):                                   # fail event (make_game.py:90-91)
            self.done = True
        if self.where["stale food"] == ("in", "fridge") and not self.open["fridge"]:  # win event (:93-98)
            self.score, self.done = 1, True
        return "ok", float(self.score), self.done

    @staticmethod
    def all_commands() -> set:
        out = {"go " + d for d in OPPOSITE} | {v + c for v in ("open ", "close ") for c in CONTAINER_ROOM}
        for o in OBJECTS:
            out |= {"take " + o, f"take {o} from table", f"take {o} from chest", f"take {o} from fridge", f"put {o} on table", f"insert {o} into chest",
                    f"insert {o} into fridge", "drop " + o, "eat " + o}
        return out
