import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from lmrl_gym_b200.engine import Engine
from lmrl_gym_b200.environment import text_env_eval
from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params, steer_to_letters, nlnl_token_id
from lmrl_gym_b200.tokenizer import ByteTokenizer
from lmrl_gym_b200.wordle import B200WordleEnv
from oracle import gpt2 as og
from oracle.mt19937 import seed_stream
toks = list(ByteTokenizer.synthetic().token_bytes); toks[45000] = b"\n\n"
tok = ByteTokenizer(toks, name="x")
cfg = GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)
params = steer_to_letters(init_gpt2_params(cfg, seed=7), tok)
B = 96
for fold in (nlnl_token_id(tok), -1):
    eng = Engine(max_envs=B, max_input_length=128, max_new_tokens=6, max_turns=6, tokenizer=tok)
    env = B200WordleEnv(eng, "wordle_official.txt", require_words_in_vocab=False)
    pol = B200GPT2Policy(eng, params, cfg, seed=11)
    eng.set_option("tok_nlnl_id", fold)
    eng.set_use_graphs(False)
    inter, _ = text_env_eval(env, pol, n_rollouts=B, seed_generator=seed_stream(2), bsize=B, verbose=False)
    seeds = seed_stream(2); seed_list = [next(seeds) for _ in range(B)]
    bad = 0
    for i in range(B):
        ep = inter[i]; logs, j = ep._logs, ep._i
        for t, tr in enumerate(ep):
            text = "".join(x.text for x in tr.pre_action_history)
            prompt = tok.encode(text)
            gen = logs.act_tokens[j, t, :logs.act_len[j, t]].tolist()
            for k, tk in enumerate(gen):
                logits = og.forward_last_logits(params, prompt + gen[:k], cfg.n_head, cfg.n_vocab, emulate_bf16=True)
                s = og.sample_scores(logits, og.row_key(11, seed_list[i], t, k), 1.0, True)
                if s[tk] < s.max() - 0.04:
                    bad += 1
                    if bad <= 6:
                        alt = [(tt, kk) for tt in range(0, 7) for kk in range(0, 7) if int(np.argmax(og.sample_scores(logits, og.row_key(11, seed_list[i], tt, kk), 1.0, True))) == tk]
                        print("fold", fold, "MISMATCH ep", i, "turn", t, "k", k, "tok", tk, repr(tok.token_bytes[tk]), "oracle argmax", int(np.argmax(s)), repr(tok.token_bytes[int(np.argmax(s))]),
                              "keys under which the device token wins:", alt, "text tail", repr(text[-12:]), "gen", gen, flush=True)
            if i < 2:
                print("fold", fold, "ep", i, "turn", t, repr(text[-14:]), "gen", [tok.token_bytes[g] for g in gen], flush=True)
    print("fold", fold, "mismatches", bad, flush=True)
    eng.close()
