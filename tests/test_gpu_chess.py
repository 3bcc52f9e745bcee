"""Chess endgames with the device policy (BASELINE.json configs[3] data flow): FenChessHistoryEnv (host rules, random opponent) driven
by B200GPT2Policy through the reference-shaped interact_environment loop.  A random-init policy plays illegal moves: every ply costs
-1 and the episode runs to max_moves + 1 — what is checked is the plumbing (prompts = spaced FEN, actions de-spaced, rewards, lengths)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_policy_plays_chess_endgames():
    from lmrl_gym_b200 import chess_env as CE
    from lmrl_gym_b200.engine import Engine
    from lmrl_gym_b200.gpt2 import B200GPT2Policy, GPT2Config, init_gpt2_params
    from lmrl_gym_b200.tokenizer import ByteTokenizer
    tok = ByteTokenizer.synthetic()
    cfg = GPT2Config(n_layer=2, n_head=4, d_model=256, d_ff=1024)
    eng = Engine(max_envs=16, max_input_length=160, max_new_tokens=6, tokenizer=tok)
    pol = B200GPT2Policy(eng, init_gpt2_params(cfg, seed=0), cfg, seed=0)
    rng = np.random.default_rng(0)
    positions = [CE.large_piece_random_endgame("kQK", rng) for _ in range(2)]
    inter, summary = CE.text_env_eval_chess_positions(positions, pol, n_rollouts=12, bsize=8, verbose=False, random_opponent=True, max_moves=3)
    assert len(inter) == 24
    for ep in inter:
        assert 1 <= len(ep) <= 4 and ep[-1].done
        for tr in ep:
            fen = CE.postprocess_state(tr.pre_action_history[-1].text)
            assert CE.ChessBoard(fen).is_valid()
            assert tr.post_action_history[-1].is_action and tr.post_action_history[-1].text.endswith("\n")
            assert tr.reward in (-1, 0, 1)
    assert summary["episode_length"]["max"] <= 4 and -4.0 <= summary["reward"]["mean"] <= 1.0
    eng.close()
