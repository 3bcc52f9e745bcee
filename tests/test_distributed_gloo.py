"""World-size-2 gloo test (CPU) of the episode sharding: the sharded ``text_env_eval`` gives the same episodes and the
same summary as one process, for an even and an uneven split.  The env is the CPU oracle (test infrastructure) behind
the reference's TextEnv interface; the policy is a deterministic scripted one."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, {root!r})
import torch.distributed as dist
from lmrl_gym_b200.distributed import text_env_eval_sharded, shard_bounds, reduce_throughput
from lmrl_gym_b200.environment import Text, TextEnv, TextPolicy
from lmrl_gym_b200.vocab import load_vocab_words
from oracle.mt19937 import seed_stream
from oracle.wordle import ReformatWordleOracle

WORDS = load_vocab_words("wordle_official_200")

class OracleEnv(TextEnv):
    def __init__(self): self.o = ReformatWordleOracle(WORDS)
    def reset(self, seed=None, options=None): return tuple(Text(a, b) for a, b in self.o.reset(seed))
    def step(self, h):
        nh, r, d = self.o.step(tuple((x.text, x.is_action) for x in h))
        return tuple(Text(a, b) for a, b in nh), r, d
    def copy(self): return OracleEnv()

class Scripted(TextPolicy):
    def act(self, h):
        k = sum(1 for x in h if x.is_action)
        w = WORDS[(7 * k + len(h[-1].text) * 13 + 3) % len(WORDS)] if k % 3 != 2 else "zzzzz"
        return h + (Text(" ".join(w) + "\n", True),)

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
if world > 1:
    dist.init_process_group("gloo")
out = {{}}
for n in (10, 7):
    inter, summ = text_env_eval_sharded(OracleEnv(), Scripted(), n_rollouts=n, seed_generator=seed_stream(5), bsize=3)
    lo, hi = shard_bounds(n, rank, world)
    assert len(inter) == hi - lo
    out[str(n)] = dict(summary={{k: {{m: float(v) for m, v in s.items()}} for k, s in summ.items()}},
                       episodes=[[[t.post_action_history[-1].text, t.reward, t.done] for t in ep] for ep in inter], lo=lo, hi=hi)
units, secs = reduce_throughput(100 * (rank + 1), 1.0 + rank)
out["reduce"] = [units, secs]
json.dump(out, open(os.path.join({outdir!r}, f"rank{{rank}}_of{{world}}.json"), "w"))
if world > 1:
    dist.destroy_process_group()
'''


def _run(world, outdir):
    script = os.path.join(outdir, "worker.py")
    with open(script, "w") as f:
        f.write(WORKER.format(root=ROOT, outdir=outdir))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    if world == 1:
        subprocess.run([sys.executable, script], check=True, env=env, timeout=300)
    else:
        subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", script], check=True, env=env, timeout=600)
    return [json.load(open(os.path.join(outdir, f"rank{r}_of{world}.json"))) for r in range(world)]


def test_sharded_eval_equals_single_process():
    with tempfile.TemporaryDirectory() as d:
        one = _run(1, d)[0]
        two = _run(2, d)
    for n in ("10", "7"):
        assert two[0][n]["summary"] == two[1][n]["summary"] == one[n]["summary"]
        assert two[0][n]["episodes"] + two[1][n]["episodes"] == one[n]["episodes"]
        assert two[0][n]["lo"] == 0 and two[0][n]["hi"] == two[1][n]["lo"] and two[1][n]["hi"] == int(n)
    assert two[0]["reduce"] == two[1]["reduce"] == [300, 2.0] and one["reduce"] == [100, 1.0]


def test_shard_bounds_partition():
    from lmrl_gym_b200.distributed import shard_bounds, sharded_seeds
    for n in (0, 1, 7, 8, 4096, 8191):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    allv = list(range(100, 111))
    got = sum((sharded_seeds(iter(allv), 11, r, 3) for r in range(3)), [])
    assert got == allv
