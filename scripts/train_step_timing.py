import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from lmrl_gym_b200.engine import Engine
from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params
from lmrl_gym_b200.tokenizer import ByteTokenizer
from lmrl_gym_b200 import ppo_train as T
tok = ByteTokenizer.synthetic(); cfg = GPT2Config.small(); rng = np.random.default_rng(1)
params = init_gpt2_params(cfg, seed=5)
lin = dict(dense=dict(kernel=(rng.standard_normal((cfg.d_model, 1)) / 28).astype(np.float32), bias=np.zeros(1, np.float32)))
eng = Engine(max_envs=64, max_input_length=264, max_new_tokens=6, tokenizer=tok)
B200GPT2Policy(eng, params, cfg)
tr = T.B200PPOTrain(eng, params, lin, cfg, pad_token_id=tok.pad_token_id, backward="b200", lr=1e-5, max_rows=32 * 264, max_scored_rows=32 * 264)
b, t = 32, 262
ids = rng.integers(0, cfg.n_vocab, (b, t)).astype(np.int32)
sta = rng.random((b, t - 1)) < 0.45
f = lambda s: (s * rng.standard_normal((b, t - 1))).astype(np.float32)
batch = dict(input_ids=ids, should_take_action=sta, old_logprobs=-10.8 + f(0.1), old_values=f(0.3), old_advantages=f(1.0), old_returns=f(0.5))
for i in range(3):
    t0 = time.perf_counter(); tr._b200_grads(batch); torch.cuda.synchronize(); t1 = time.perf_counter()
    tr.step(batch); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"grads {1e3*(t1-t0):.1f} ms   full step {1e3*(t2-t1):.1f} ms", flush=True)
t0 = time.perf_counter(); tr.push_to_engine(); torch.cuda.synchronize(); print(f"push_to_engine {1e3*(time.perf_counter()-t0):.1f} ms")
from lmrl_gym_b200.value_inference import forward_tokens
t0 = time.perf_counter(); forward_tokens(eng, 0, 0, [ids[i] for i in range(b)], logprobs=True, v=True, pitch=t); print(f"forward_tokens {1e3*(time.perf_counter()-t0):.1f} ms")
