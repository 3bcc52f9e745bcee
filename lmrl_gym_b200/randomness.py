"""Host-side seed stream of the reference (llm_rl_scripts/maze/env/randomness.py:35-41).

``seed_generator(seed)`` yields successive ``getrandbits(64)`` of ONE private Mersenne-Twister state seeded with
``seed`` — the reference keeps that state by saving / restoring the global ``random`` module state around every draw
(``RandomState.unfreeze/freeze``), which is observationally the same as owning a ``random.Random(seed)``.  Every driver
passes ``seed_generator(...)`` to ``text_env_eval`` so that episode i receives the i-th value (environment.py:230).
"""
from __future__ import annotations

import random
from typing import Iterator, Optional


def seed_generator(seed: Optional[int]) -> Iterator[int]:
    rng = random.Random(seed)
    while True:
        yield rng.getrandbits(64)
