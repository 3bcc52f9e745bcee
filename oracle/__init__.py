"""CPU oracle for the LMRL-Gym batched-rollout path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of the reference's algorithm
for the hot path (``LLM_RL/environment.py::text_env_eval`` and what it calls).
It exists to *check* the CUDA path and to provide the CPU baseline number in
``bench.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing in
``lmrl_gym_b200/`` imports it, and the product fails loudly without its CUDA
library instead of falling back to this code.

Parity pinning
--------------
* Env side (Wordle, Maze, MT19937, interaction loop): PINNED.  The reference's
  own Python env imports in the build container (three cosmetic import stubs);
  ``oracle/gen_golden.py`` runs it and writes ``tests/golden/*.json``.  The
  restatements here are checked against those vectors by ``-m "not gpu"`` tests.
* Policy side (GPT-2 forward / sampling): PARITY UNPINNED by the reference.  The
  arithmetic lives in un-vendored third-party code (Sea-Snell/JaxSEQ @ HEAD,
  unpinned; transformers==4.26.1 Flax GPT-2; jax==0.4.7) that is absent from
  ``/root/reference`` and from this image.  ``oracle/gpt2.py`` restates the
  published GPT-2 architecture in fp32 torch and is itself checked against
  ``transformers.GPT2LMHeadModel`` (torch) with identical weights.
"""
