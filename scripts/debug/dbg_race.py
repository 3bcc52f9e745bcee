import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from lmrl_gym_b200.engine import Engine
from lmrl_gym_b200.environment import text_env_eval
from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params, steer_to_letters
from lmrl_gym_b200.tokenizer import ByteTokenizer
from lmrl_gym_b200.wordle import B200WordleEnv
from oracle.mt19937 import seed_stream
tok = ByteTokenizer.synthetic()
cfg = GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)
params = init_gpt2_params(cfg, seed=7)
mode = sys.argv[1]
turns, new, B = {"a": (1, 1, 4), "b": (1, 3, 4), "c": (3, 3, 4), "d": (6, 6, 200), "e": (6, 6, 4), "f": (3, 6, 4), "g": (6, 3, 4), "h": (4, 3, 4), "i": (3, 3, 64)}[mode]
print("mode", mode, turns, new, B, flush=True)
eng = Engine(max_envs=B, max_input_length=64, max_new_tokens=new, max_turns=turns, tokenizer=tok)
env = B200WordleEnv(eng, "wordle_official.txt")
pol = B200GPT2Policy(eng, params, cfg, seed=1)
eng.set_use_graphs(False)
print("built", flush=True)
inter, summ = text_env_eval(env, pol, n_rollouts=B, seed_generator=seed_stream(1), bsize=B, verbose=False)
print("done", mode, float(summ["reward"]["mean"]), flush=True)
eng.close()
