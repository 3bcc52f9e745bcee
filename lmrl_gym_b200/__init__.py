"""lmrl_gym_b200 — B200-native batched rollout engine behind LMRL-Gym's TextEnv / TextPolicy / text_env_eval surface.

Host-side mirror of the reference interface (``environment``), device tables (``vocab``, ``tokenizer``) and the
ctypes binding of the CUDA library.  Importing the package never touches the GPU; constructing an ``Engine``,
env or policy loads ``csrc/liblmrl_b200.so`` and raises ``LmrlError`` if it is missing — there is no CPU fallback.
"""
from .environment import (  # noqa: F401
    Text, TextHistory, TextTrajectory, TextTrajectoryChain, TextEnv, BatchedTextEnv, TextPolicy, BatchedTextPolicy,
    TextEnvToBatchedTextEnv, BatchedTextEnvToTextEnv, TextPolicyToBatchedTextPolicy, BatchedTextPolicyToTextPolicy,
    InteractionTransition, interact_environment, text_env_eval, text_history_to_str, summarize_episodes, UserPolicy,
)
from .tokenizer import ByteTokenizer  # noqa: F401
from .randomness import seed_generator  # noqa: F401
from ._lib import LmrlError  # noqa: F401


def __getattr__(name):
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name == "B200WordleEnv":
        from .wordle import B200WordleEnv
        return B200WordleEnv
    if name == "B200MazeEnv":
        from .maze import B200MazeEnv
        return B200MazeEnv
    if name in ("B200GPT2Policy", "B200GPT2ValuePolicy", "GPT2Config", "init_gpt2_params", "init_mlp_head_params"):
        from . import gpt2
        return getattr(gpt2, name)
    lazy = {"B200TextNavEnv": "textnav", "B200PPOInference": "ppo", "AdaptiveKLController": "ppo", "FixedKLController": "ppo", "PPOData": "ppo", "PPODataset": "ppo", "PPOIterableDataset": "ppo", "BlockingStrategy": "ppo", "Padding": "ppo", "Truncation": "ppo",
            "B200PPOTrain": "ppo_train", "ppo_train_loop": "ppo_train", "B200ValueRLInference": "value_inference", "B200ILQLInference": "value_inference",
            "ReRankerPolicy": "reranker", "ReRankerSamplePolicy": "reranker", "build_bc_score_fn": "reranker", "build_ilql_score_fn": "reranker",
            "build_mc_score_fn": "reranker", "FenChessHistoryEnv": "chess_env", "BatchedFenChessHistoryEnv": "chess_env", "ChessEnv": "chess_env",
            "text_env_eval_chess_positions": "chess_env"}
    if name in lazy:
        import importlib
        return getattr(importlib.import_module("." + lazy[name], __name__), name)
    raise AttributeError(name)
