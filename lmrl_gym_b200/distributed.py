"""Episode sharding across GPUs: one process per GPU, no collective on the rollout path.

Episodes never interact (``TextEnvToBatchedTextEnv`` holds independent copies, LLM_RL/environment.py:92), so the
multi-GPU form of ``text_env_eval`` is a partition of the episode index range: rank r runs the r-th contiguous block of
the ONE ``seed_generator`` stream the reference would have consumed (environment.py:230), which makes every episode —
and therefore the summary — independent of the number of ranks.  The only communication is the end-of-eval gather of
(return, done, length) per episode, 16 bytes each.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from .environment import summarize_episodes, text_env_eval


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank ``rank`` out of ``n`` items; the first ``n % world`` ranks get one extra."""
    assert 0 <= rank < world
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_seeds(seed_generator: Optional[Iterator[int]], n_rollouts: int, rank: int, world: int) -> List[Optional[int]]:
    """This rank's seeds.  Every rank draws all ``n_rollouts`` values so that generators stay aligned across ranks and
    across successive calls, exactly as a single process would have advanced the stream."""
    lo, hi = shard_bounds(n_rollouts, rank, world)
    if seed_generator is None:
        return [None] * (hi - lo)
    allv = [next(seed_generator) for _ in range(n_rollouts)]
    return allv[lo:hi]


def _dist():
    import torch.distributed as dist
    return dist


def gather_episode_stats(rewards: Sequence[float], dones: Sequence[bool], lengths: Sequence[int], group=None):
    """All ranks receive the per-episode arrays of every rank, concatenated in rank order (== episode order)."""
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(rewards, np.float64), np.asarray(dones, bool), np.asarray(lengths, np.int64)
    import torch
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    n = torch.tensor([len(rewards)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    cap = int(max(int(c.item()) for c in counts))
    buf = torch.zeros((cap, 3), dtype=torch.float64, device=dev)
    if len(rewards):
        buf[:len(rewards), 0] = torch.as_tensor(np.asarray(rewards, np.float64), device=dev)
        buf[:len(rewards), 1] = torch.as_tensor(np.asarray(dones, np.float64), device=dev)
        buf[:len(rewards), 2] = torch.as_tensor(np.asarray(lengths, np.float64), device=dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    parts = [b[:int(c.item())].cpu().numpy() for b, c in zip(bufs, counts)]
    allv = np.concatenate(parts, axis=0) if parts else np.zeros((0, 3))
    return allv[:, 0], allv[:, 1].astype(bool), allv[:, 2].astype(np.int64)


def text_env_eval_sharded(env, policy, n_rollouts: int, seed_generator: Optional[Iterator[int]] = None,
                          env_options: Optional[Dict] = None, interaction_callback: Optional[Callable] = None, bsize: int = 1,
                          rank: Optional[int] = None, world: Optional[int] = None, group=None) -> Tuple[List[Any], Dict[str, Any]]:
    """``text_env_eval`` (LLM_RL/environment.py:211-267) over all ranks of the process group: returns this rank's
    interactions and the GLOBAL summary (identical on every rank and identical to a single-process run)."""
    dist = _dist()
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1
    seeds = sharded_seeds(seed_generator, n_rollouts, rank, world)
    if len(seeds) == 0:
        interactions: List[Any] = []
    else:
        interactions, _ = text_env_eval(env, policy, n_rollouts=len(seeds), seed_generator=iter(seeds) if seed_generator is not None else None,
                                        env_options=env_options, interaction_callback=interaction_callback, bsize=bsize, verbose=False)
    rewards = [sum(t.reward for t in ep) for ep in interactions]
    dones = [ep[-1].done for ep in interactions]
    lengths = [len(ep) for ep in interactions]
    r, d, l = gather_episode_stats(rewards, dones, lengths, group)
    return interactions, summarize_episodes(r, d, [int(x) for x in l])


def reduce_throughput(units: int, seconds: float, group=None) -> Tuple[int, float]:
    """Whole-job (units summed over ranks, time = max over ranks): the bench's weak-scaling aggregate."""
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(units), float(seconds)
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([seconds], dtype=torch.float64, device=dev)
    u = torch.tensor([units], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return int(u.item()), float(t.item())


def init_from_env() -> Tuple[int, int, int]:
    """(rank, world, local device) of a driver process: under torchrun (RANK / WORLD_SIZE / LOCAL_RANK set) joins the NCCL process
    group, one process per GPU; otherwise a single process on device 0."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1, 0
    import torch
    dist = _dist()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend="nccl" if torch.cuda.is_available() else "gloo")
    return dist.get_rank(), dist.get_world_size(), local


def finalize() -> None:
    """Leave the process group ``init_from_env`` joined (no-op in a single process)."""
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
