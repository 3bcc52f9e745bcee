"""GPU parity of the GEMM kernels (tcgen05/TMEM/TMA with one CTA per tile, its CTA-pair form, and mma.sync) against a float64 numpy product of the same
bf16 operands.  Tolerance: fp32 accumulation over K <= 3072 terms of magnitude <= 1 -> |err| <= 2e-3 absolute."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(1, 256, 64), (37, 256, 128), (300, 768, 768), (129, 768, 3072), (4096 + 17, 2304, 768), (1000, 3072, 768), (260, 50432, 768),
          (40000, 768, 768), (4096, 768, 3072), (77, 192, 64), (5000, 1152, 768)]


def _bf16_bits(a):
    t = torch.from_numpy(a).to(torch.bfloat16)
    return t.view(torch.int16).numpy().view(np.uint16), t.to(torch.float32).numpy()


@pytest.mark.parametrize("impl", [1, 2, 3, 0])
def test_gemm_resid_epilogue_matches_numpy(impl):
    from lmrl_gym_b200.engine import Engine
    eng = Engine(max_envs=8)
    rng = np.random.default_rng(0)
    for (M, N, K) in SHAPES:
        if (N > 4096 and impl == 0) or (impl == 3 and N % 192) or (impl in (1, 2) and N % 256) or (impl == 0 and N % 128):
            continue
        A_bits, A = _bf16_bits(rng.standard_normal((M, K), dtype=np.float32) * 0.5)
        W_bits, W = _bf16_bits(rng.standard_normal((N, K), dtype=np.float32) * 0.05)
        bias = rng.standard_normal(N).astype(np.float32)
        resid = rng.standard_normal((M, N)).astype(np.float32)
        got = eng.test_gemm(impl, A_bits, W_bits, bias, resid)
        want = resid.astype(np.float64) + A.astype(np.float64) @ W.astype(np.float64).T + bias
        err = np.abs(got - want).max()
        print(f"impl={impl} M={M} N={N} K={K}: max abs err {err:.3e}")
        assert err < 2e-3, (impl, M, N, K, err)
    eng.close()
