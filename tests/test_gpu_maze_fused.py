"""GPU parity of the fused Maze rollout (policy tokens -> MazeEnv.step -> next prompt window, all on the device) against
the CPU oracle env (oracle/maze.py, pinned to the reference by tests/golden/maze_cases.json) and the oracle sampler.

The policy is steered so that most sampled tokens are whole moves: the test tokenizer carries four extra ids whose bytes
are 'move left' ... 'move down' (one sampled token = one action once out_str_process appends the newline), while the
canonical re-tokenisation of the action text ('move', ' left', '\\n') differs from the sampled id — so the KV roll-back,
the last_k window, LEFT truncation and the collapse on unknown actions are all exercised.  Bit-exact bar for texts,
rewards, done flags and returns; sampled ids are the oracle's arg-max within tolerance.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from lmrl_gym_b200.gpt2 import GPT2Config, init_gpt2_params  # noqa: E402
from lmrl_gym_b200.tokenizer import ByteTokenizer  # noqa: E402

TOL_EMU = 2e-2
MOVE_IDS = (40000, 40001, 40002, 40003)
MOVES = ("move left", "move right", "move up", "move down")


def _maze_tokenizer():
    base = ByteTokenizer.synthetic()
    toks = list(base.token_bytes)
    for i, m in zip(MOVE_IDS, MOVES):
        toks[i] = m.encode()
    return ByteTokenizer(toks, name="synthetic-gpt2-shape-v1+moves")


def _steered_params(cfg, tok, p_move=0.85, seed=3):
    """Context-free LM head: the final LN emits a constant vector, the four move ids share p_move, the rest of the
    vocabulary keeps N(0, 1) logits (=> unknown actions)."""
    params = init_gpt2_params(cfg, seed=seed)
    d = cfg.d_model
    u = np.random.default_rng(5).standard_normal(d)
    u /= np.linalg.norm(u)
    scale = 50.0
    # rest of the vocabulary: sum exp(N(0,1)) ~ 50257 * e^0.5; pick the move logit so that 4 e^L / (4 e^L + rest) = p_move
    rest = cfg.n_vocab * np.exp(0.5)
    L = np.log(p_move / (1 - p_move) * rest / 4)
    wte = params["wte"].copy()
    wte[list(MOVE_IDS)] = (L / scale) * u
    params["wte"] = wte.astype(np.float32)
    params["lnf_g"] = np.zeros_like(params["lnf_g"])
    params["lnf_b"] = (scale * u).astype(np.float32)
    return params


@pytest.mark.parametrize("describe_function,last_k,max_input,max_steps", [
    ("describe_observation_give_position", 1, 128, 20),      # config 1 (fully observed, last_k = 1)
    ("describe_observation_only_walls", 6, 128, 24),          # partially observed, window slides
    ("describe_observation_only_walls", 40, 64, 30),          # config 3's last_k = 40; LEFT truncation to 64 tokens kicks in
])
def test_fused_maze_rollout_matches_oracle(describe_function, last_k, max_input, max_steps):
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    from lmrl_gym_b200.maze import B200MazeEnv
    from lmrl_gym_b200.environment import text_env_eval
    from oracle import gpt2 as og
    from oracle.maze import MazeOracle
    from oracle.mt19937 import seed_stream
    tok = _maze_tokenizer()
    assert tok.encode("move left\n") != [MOVE_IDS[0], tok.newline_id()]
    cfg = GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)
    params = _steered_params(cfg, tok)
    B = 96
    eng = Engine(max_envs=B, max_input_length=max_input, max_new_tokens=1, max_turns=max_steps + 1, tokenizer=tok)
    env = B200MazeEnv(eng, "double_t_maze", describe_function, "standard_reward", last_k=last_k, max_steps=max_steps)
    assert env._lmrl_fusable
    pol = B200GPT2Policy(eng, params, cfg, seed=17)
    n0 = eng.launch_count(reset=True)
    inter, summary = text_env_eval(env, pol, n_rollouts=B + 10, seed_generator=seed_stream(4), bsize=B, verbose=False)
    assert eng.launch_count() > 0
    seeds = seed_stream(4)
    seed_list = [next(seeds) for _ in range(B + 10)]
    n_known = n_steps = n_success = 0
    rets = []
    for i, ep in enumerate(inter):
        o = MazeOracle("double_t_maze", describe_function, "standard_reward", last_k=last_k, max_steps=max_steps)
        h = o.reset(seed_list[i])
        assert 1 <= len(ep) <= max_steps + 1
        for t, tr in enumerate(ep):
            assert [[x.text, x.is_action] for x in tr.pre_action_history] == [[a, b] for a, b in h], (i, t)
            action = tr.post_action_history[-1].text
            h, r, d = o.step(h + ((action, True),))
            assert [[x.text, x.is_action] for x in tr.post_transition_history] == [[a, b] for a, b in h], (i, t)
            assert tr.reward == r and tr.done == d, (i, t)
            n_known += action in ("move left\n", "move right\n", "move up\n", "move down\n")
            n_steps += 1
        assert ep[-1].done
        n_success += ep[-1].post_transition_history[-1].text == "Success\n"
        rets.append(sum(x.reward for x in ep))
    assert n_known > 0.6 * n_steps and n_known < n_steps, (n_known, n_steps)
    assert float(summary["reward"]["mean"]) == float(np.mean(np.asarray(rets, np.float32)))
    print(f"maze {describe_function} last_k={last_k}: {n_steps} steps, {n_known} known moves, {n_success} successes")
    # sampler parity on the canonical (re-tokenised, LEFT-truncated) prompt of a few episodes
    for i in (0, 41, B + 5):
        ep = inter[i]
        logs, j = ep._logs, ep._i
        for t, tr in enumerate(ep):
            if t % 3:
                continue
            prompt = tok.encode("".join(x.text for x in tr.pre_action_history))[-max_input:]
            tk = int(logs.act_tokens[j, t, 0])
            logits = og.forward_last_logits(params, prompt, cfg.n_head, cfg.n_vocab, emulate_bf16=True)
            s = og.sample_scores(logits, og.row_key(17, seed_list[i], t, 0), 1.0, True)
            assert s[tk] >= s.max() - 2 * TOL_EMU, (i, t)
    eng.close()


def test_fused_maze_init_position_option_and_early_exit():
    """env_options={'init_position': p} (maze/bc/eval_bc.py:164-173) reaches the device reset; a batch whose episodes all end
    early stops enqueueing turns (the periodic live-count test)."""
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy
    from lmrl_gym_b200.maze import B200MazeEnv
    from lmrl_gym_b200.environment import text_env_eval
    from oracle.mt19937 import seed_stream
    tok = _maze_tokenizer()
    cfg = GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)
    params = _steered_params(cfg, tok, p_move=0.97)
    eng = Engine(max_envs=32, max_input_length=128, max_new_tokens=1, max_turns=101, tokenizer=tok)
    env = B200MazeEnv(eng, "double_t_maze", "describe_observation_give_position", "standard_reward", last_k=1, max_steps=100)
    pol = B200GPT2Policy(eng, params, cfg, seed=2)
    inter, summary = text_env_eval(env, pol, n_rollouts=32, seed_generator=seed_stream(1), env_options={"init_position": (7, 6)}, bsize=32, verbose=False)
    first = "The goal is at position 8, 6. Your current position is at position 7, 6. There are walls to your right, to your left.\n"
    assert all(ep[0].pre_action_history[0].text == first for ep in inter)
    # one 'move down' from (7, 6) reaches the goal: with 97% moves most episodes are short
    assert float(summary["length"]["mean"]) < 40
    launches_short = eng.launch_count(reset=True)
    assert launches_short < 101 * 40          # far fewer than 101 turns x ~40 launches: the loop stopped early
    eng.close()
