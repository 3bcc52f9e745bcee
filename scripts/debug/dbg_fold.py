"""compute-sanitizer target: the LayerNorm-fold launch sequence (ln_fold = 1) and the 4-heads-per-warp decode attention on a 2-layer toy model:
ragged prefill over several m-tiles (XB residual epilogue across tiles, partial last tile) + a short decode loop (128x192 decode tiles are
GPT-2-small only, so the decode passes here run the 256-wide XB instantiation), eager launches.
    LMRL_NO_GRAPHS=1 compute-sanitizer --tool racecheck python scripts/debug/dbg_fold.py [n_prompts]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from lmrl_gym_b200.engine import Engine
from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params
from lmrl_gym_b200.tokenizer import ByteTokenizer
tok = ByteTokenizer.synthetic()
cfg = GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)
params = init_gpt2_params(cfg, seed=7)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
eng = Engine(max_envs=max(n, 8), max_input_length=48, max_new_tokens=3, tokenizer=tok)
pol = B200GPT2Policy(eng, params, cfg, seed=1)
eng.set_use_graphs(False)
eng.set_option("gemv_max_rows", 0)
rng = np.random.default_rng(0)
prompts = [rng.integers(0, cfg.n_vocab, size=int(k)).tolist() for k in rng.integers(1, 48, size=n)]
eng.set_option("ln_fold", 0)
base = pol.logits(prompts)
eng.set_option("ln_fold", 1)
got = pol.logits(prompts)
gen = pol.generate_tokens(prompts, episode_keys=list(range(n)), turn=0)
print("max |logit diff| fold vs default", float(np.abs(got - base).max()), "generated", sum(len(g) for g in gen), flush=True)
eng.close()
