"""Text-Nav on the GPU: the device state machine behind B200TextNavEnv against the oracle world model on seeded
random command walks (admissible lists, reward = score, done, room after every command, 96 concurrent games), the reference's two
shipped transcripts replayed through the device env, and a scripted walkthrough policy winning from every start room through the
reference-shaped interact_environment loop."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.textnav import CHOICE_ORDER, ROOMS, TextNavOracle  # noqa: E402

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "textnav_transcripts.json")))


def adm_of(text):
    line = [l for l in text.split("\n") if l.startswith("Admissible commands: ")][-1]
    return line[len("Admissible commands: "):].split(", ")


@pytest.fixture(scope="module")
def eng():
    from lmrl_gym_b200.engine import Engine
    e = Engine(max_envs=96, max_input_length=64, max_new_tokens=4)
    yield e
    e.close()


def test_device_env_matches_oracle_on_random_walks(eng):
    from lmrl_gym_b200.environment import Text
    from lmrl_gym_b200.textnav import B200TextNavEnv
    env = B200TextNavEnv(eng)
    B = 96
    seeds = [random.Random(s).getrandbits(64) for s in range(B)]
    hist = env.reset(seeds, None)
    oracles = [TextNavOracle() for _ in range(B)]
    rngs = [random.Random(1000 + i) for i in range(B)]
    every = sorted(TextNavOracle.all_commands())
    junk = ["", "go", "Open fridge", "look", "examine chest", "close gridge", " go west \n", "take   plate", "x" * 300, "é"]
    for i, o in enumerate(oracles):
        assert o.reset(seeds[i]) == adm_of(hist[i][-1].text)
        assert o.room == random.Random(seeds[i]).choice(CHOICE_ORDER)
    done = [False] * B
    for step in range(150):
        cmds = []
        for i, o in enumerate(oracles):
            r = rngs[i].random()
            cmds.append(rngs[i].choice(o.admissible()) if r < 0.75 else (rngs[i].choice(every) if r < 0.92 else rngs[i].choice(junk)))
        res = env.step([None if done[i] else hist[i] + (Text(cmds[i], True),) for i in range(B)], done)
        for i, o in enumerate(oracles):
            if done[i]:
                assert res[i] is None
                continue
            outcome, reward, d = o.step(cmds[i])
            h, r, dd = res[i]
            assert (r, dd) == (reward, d), (i, step, cmds[i])
            assert adm_of(h[-1].text) == o.admissible(), (i, step, cmds[i])
            assert ROOMS[env.state[i, 0]] == o.room and int(env.state[i, 11]) == o.moves
            assert ("You can't see any such thing." in h[-1].text) == (outcome == "unknown")
            hist[i] = h[-2:]            # keep histories short; the env only reads the last action
            done[i] = d
    assert any(done) and not all(done)


@pytest.mark.parametrize("gi", range(len(G["games"])))
def test_reference_transcripts_through_the_device_env(eng, gi):
    from lmrl_gym_b200.environment import Text
    from lmrl_gym_b200.textnav import B200TextNavEnv
    game = G["games"][gi]
    start = [r for r in ROOMS if TextNavOracle().reset(start_room=r) == game["first"]["admissible"]][0]
    env = B200TextNavEnv(eng)
    h = env.reset([0], [dict(start_room=start)])[0]
    assert adm_of(h[-1].text) == game["first"]["admissible"]
    assert h[-1].text.startswith("Your objective is to put the stale food on the table in the living room into the fridge in the kitchen.")
    for st in game["steps"]:
        h, r, d = env.step([h + (Text(st["command"], True),)])[0]
        assert adm_of(h[-1].text) == st["admissible"] and r == float(st["score"]) and d == st["done"]
        assert ("You can't see any such thing." in h[-1].text) == (not st["known"])
    if game["steps"][-1]["score"]:
        assert "Your score has just gone up by one point." in h[-1].text
    # the fixed response lines of the transcripts
    env.reset([0], [dict(start_room="kitchen")])
    h2, _, _ = env.step([(Text("x", False), Text("open fridge", True))])[0]
    assert h2[-1].text.startswith("You open the fridge, revealing a fruit.\n")


def test_walkthrough_policy_wins_from_every_room(eng):
    """The reference's data collector follows a per-room walkthrough (text_nav/dataset/collect_data.py:9-18); through the generic
    interact_environment loop it must reach score 1 from all eight start rooms."""
    from lmrl_gym_b200.environment import Text, TextPolicy, interact_environment
    from lmrl_gym_b200.textnav import B200TextNavEnv
    route = {"bedroom": "go west", "office": "go north", "bathroom": "go north", "garden": "go south", "backyard": "go east"}

    class Walkthrough(TextPolicy):
        def act(self, text_history):
            adm = adm_of(text_history[-1].text)
            has_food = "drop stale food" in adm
            if "take stale food from table" in adm:
                cmd = "take stale food from table"
            elif has_food and "insert stale food into fridge" in adm:
                cmd = "insert stale food into fridge"
            elif has_food and "open fridge" in adm:
                cmd = "open fridge"
            elif "close fridge" in adm and "take stale food from fridge" in adm:
                cmd = "close fridge"
            else:
                room = next(r for item in reversed(text_history) for r in ROOMS if f"-= {r.title()} =-" in item.text)
                if room in route:
                    cmd = route[room]
                elif room == "living room":
                    cmd = "go west"
                elif room == "dining room":
                    cmd = "go west" if has_food else "go east"
                else:
                    cmd = "go east"           # kitchen without the food
            return text_history + (Text(cmd, True),)

    env = B200TextNavEnv(eng, display_location=True)
    episodes = interact_environment(env, Walkthrough(), env_seed=[0] * 8, env_options=[dict(start_room=r) for r in ROOMS], bsize=8)
    for r, ep in zip(ROOMS, episodes):
        assert ep[-1].done and ep[-1].reward == 1.0 and sum(t.reward for t in ep) == 1.0, r
        assert len(ep) <= 12
