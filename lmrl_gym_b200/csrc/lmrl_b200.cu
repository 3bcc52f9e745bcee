// lmrl_b200.cu — single translation unit of liblmrl_b200.so (sm_100a).
//   engine.h          opaque engine, allocation bookkeeping
//   env_core.h        per-env integer logic (host+device)
//   env_kernels.cuh   Wordle / Maze kernels
//   lmrl_api.inl      C ABI: lifecycle, tables, text-path envs
//   gpt2_engine.inl   C ABI: GPT-2 decoder, sampler, fused rollout, per-position forward
//   textnav_core.h / textnav_kernels.inl   Text-Nav int32 state machine (host+device logic, kernels, C ABI)
//   train_kernels.inl C ABI: hand-written backward pass of the policy + value head (PPO gradient step)
//   chess_core.h / chess_api.inl   host-side chess rules (FEN, SAN, legal moves, game state) for the Chess-endgames env
//   ppo_kernels.inl   C ABI: PPO data path after the forward passes (KL rewards, GAE, whitening, clipped loss)
#include <stdarg.h>
#include <algorithm>
#include "engine.h"
#include "gpt2_engine.inl"
#include "train_kernels.inl"
#include "lmrl_api.inl"
#include "ppo_kernels.inl"
#include "textnav_kernels.inl"
#include "chess_api.inl"
