"""When the reference package is importable next to this one, the host mirror re-exports the reference's OWN value types and ABCs
(lmrl_gym_b200/environment.py), so device envs / policies pass the reference's isinstance checks (environment.py:164-167) and run under
the reference's own ``text_env_eval`` unchanged.  Needs /root/reference (build container only): skipped elsewhere.  Runs in a subprocess
so that the import order of the main test process is not disturbed."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
sys.path.insert(0, %r)
sys.path.insert(0, %r)
import LLM_RL.environment as R
import lmrl_gym_b200 as L
import lmrl_gym_b200.environment as E
assert E._USING_REFERENCE_TYPES
assert L.Text is R.Text and L.BatchedTextEnv is R.BatchedTextEnv and L.BatchedTextPolicy is R.BatchedTextPolicy and L.TextPolicy is R.TextPolicy
from lmrl_gym_b200.wordle import B200WordleEnv
from lmrl_gym_b200.maze import B200MazeEnv
from lmrl_gym_b200.textnav import B200TextNavEnv
from lmrl_gym_b200.gpt2 import B200GPT2Policy, B200GPT2ValuePolicy
from lmrl_gym_b200.reranker import ReRankerPolicy
assert all(issubclass(c, R.BatchedTextEnv) for c in (B200WordleEnv, B200MazeEnv, B200TextNavEnv))
assert issubclass(B200GPT2Policy, R.BatchedTextPolicy) and issubclass(B200GPT2ValuePolicy, R.BatchedTextPolicy) and issubclass(ReRankerPolicy, R.TextPolicy)

# the reference's own loop drives a product-side policy over a reference-side env
class Countdown(R.TextEnv):
    def __init__(self):
        self.n = 0
    def reset(self, seed=None, options=None):
        self.n = 3 if seed is None else 1 + seed %% 3
        return (R.Text("%%d left\n" %% self.n, False),)
    def step(self, text_history):
        assert text_history[-1].is_action
        self.n -= 1
        good = text_history[-1].text == "down\n"
        return text_history + (R.Text("%%d left\n" %% self.n, False),), (1.0 if good else -1.0), self.n == 0
    def copy(self):
        c = Countdown(); c.n = self.n
        return c

policy = ReRankerPolicy(proposal_fn=lambda h: [h + (R.Text(a, True),) for a in ("up\n", "down\n")],
                        score_fn=lambda hs: [1.0 if h[-1].text == "down\n" else 0.0 for h in hs])
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):          # the reference's single->batched adapter prints every call (environment.py:130-132)
    raw, summary = R.text_env_eval(Countdown(), policy, n_rollouts=5, seed_generator=iter(range(5)), bsize=2, verbose=False)
assert [len(ep) for ep in raw] == [1, 2, 3, 1, 2], [len(ep) for ep in raw]
assert abs(float(summary["reward"]["mean"]) - sum([1, 2, 3, 1, 2]) / 5) < 1e-6 and float(summary["done"]["mean"]) == 1.0       # float32 stats (:244-245)
# and the mirror's loop gives the same episodes with the same objects
raw2, summary2 = E.text_env_eval(Countdown(), policy, n_rollouts=5, seed_generator=iter(range(5)), bsize=2, verbose=False)
assert [[t.reward for t in ep] for ep in raw2] == [[t.reward for t in ep] for ep in raw]
assert {k: {m: float(v) for m, v in d.items()} for k, d in summary2.items()} == {k: {m: float(v) for m, v in d.items()} for k, d in summary.items()}
print("ok")
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "LLM_RL")), reason="the reference tree is only present in the build container")
def test_mirror_adopts_the_reference_types_and_runs_under_its_loop():
    out = subprocess.run([sys.executable, "-c", SCRIPT % (REF, ROOT)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-3000:]
