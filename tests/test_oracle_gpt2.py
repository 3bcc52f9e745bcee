"""The GPT-2 oracle restatement vs transformers' torch GPT2LMHeadModel with identical weights (CPU)."""
import numpy as np
import torch

from lmrl_gym_b200.gpt2 import GPT2Config, init_gpt2_params, params_from_hf_state_dict
from oracle import gpt2 as og


def test_oracle_forward_matches_hf_gpt2():
    from transformers import GPT2Config as HFConfig, GPT2LMHeadModel
    cfg = GPT2Config(n_layer=3, n_head=4, d_model=256, d_ff=1024, n_vocab=1000, n_vocab_rows=1000, n_pos=64)
    torch.manual_seed(0)
    hf = GPT2LMHeadModel(HFConfig(vocab_size=1000, n_positions=64, n_embd=256, n_layer=3, n_head=4, attn_pdrop=0.0, embd_pdrop=0.0,
                                  resid_pdrop=0.0)).eval()
    # give biases / LN params non-trivial values so they are exercised
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if "bias" in n or "ln" in n:
                p.add_(0.1 * torch.randn_like(p))
    params = params_from_hf_state_dict(hf.state_dict(), cfg)
    toks = [5, 17, 999, 0, 3, 3, 250, 77, 12]
    with torch.no_grad():
        want = hf(torch.tensor([toks])).logits[0, -1].numpy()
    got = og.forward_last_logits(params, toks, cfg.n_head, cfg.n_vocab, weights_bf16=False)
    assert np.abs(got - want).max() < 2e-4
    # bf16 weight rounding (what the engine stores) moves logits by a bounded amount
    got_bf = og.forward_last_logits(params, toks, cfg.n_head, cfg.n_vocab, weights_bf16=True)
    assert np.abs(got_bf - want).max() < 0.05


def test_sampler_restatement_properties():
    k = og.row_key(7, 12345, 2, 3)
    assert k == og.row_key(7, 12345, 2, 3) and k != og.row_key(7, 12345, 2, 4) and k != og.row_key(7, 12346, 2, 3)
    g = og.gumbel_noise(k, 50257)
    assert np.isfinite(g).all() and abs(g.mean() - 0.5772) < 0.02 and abs(g.var() - np.pi ** 2 / 6) < 0.05
    # Gumbel-max is a categorical sampler: empirical frequencies follow softmax(logits)
    logits = np.array([0.0, 1.0, 2.0, -1.0])
    p = np.exp(logits) / np.exp(logits).sum()
    counts = np.zeros(4)
    for ep in range(4000):
        counts[int(np.argmax(og.sample_scores(logits, og.row_key(0, ep, 0, 0))))] += 1
    assert np.abs(counts / 4000 - p).max() < 0.03


def test_init_params_shapes():
    cfg = GPT2Config.small()
    assert cfg.matmul_params == 123532032
    p = init_gpt2_params(GPT2Config(n_layer=2, n_head=2, d_model=128, d_ff=512, n_vocab=300, n_vocab_rows=301, n_pos=32), seed=1)
    assert p["wte"].shape == (301, 128) and p["layers"][1]["fc2_w"].shape == (512, 128)


def test_value_policy_restatement_properties():
    """generation.py:112-119: logits = pi_beta + beta * min(q1, q2); one head -> beta * q1; beta = 0 -> pi_beta."""
    from lmrl_gym_b200.gpt2 import init_mlp_head_params
    cfg = GPT2Config(n_layer=1, n_head=2, d_model=64, d_ff=128, n_vocab=97, n_vocab_rows=98, n_pos=16)
    pi, base = init_gpt2_params(cfg, seed=1), init_gpt2_params(cfg, seed=2)
    q1, q2 = init_mlp_head_params(cfg, seed=3), init_mlp_head_params(cfg, seed=4)
    toks = [5, 9, 2, 77]
    pi_logits = og.forward_last_logits(pi, toks, cfg.n_head, cfg.n_vocab, weights_bf16=False)
    h_v = og.forward_hidden(base, toks, cfg.n_head, weights_bf16=False)[-1]
    qa = og.mlp_head(h_v, q1, weights_bf16=False).numpy()[:cfg.n_vocab]
    qb = og.mlp_head(h_v, q2, weights_bf16=False).numpy()[:cfg.n_vocab]
    assert abs(qa.mean() + 4.4) < 0.5                      # layer2_bias_init = -4.4 (train_ilql_gpt2.py:224-231)
    both = og.value_policy_last_logits((pi, base, q1, q2), toks, cfg.n_head, cfg.n_vocab, 16.0, weights_bf16=False)
    one = og.value_policy_last_logits((pi, base, q1, None), toks, cfg.n_head, cfg.n_vocab, 16.0, weights_bf16=False)
    assert np.allclose(both, pi_logits + 16.0 * np.minimum(qa, qb), atol=1e-4)
    assert np.allclose(one, pi_logits + 16.0 * qa, atol=1e-4)
    zero = og.value_policy_last_logits((pi, base, q1, q2), toks, cfg.n_head, cfg.n_vocab, 0.0, weights_bf16=False)
    assert np.allclose(zero, pi_logits, atol=1e-5)
    # the bf16-emulating variant (min of the two per-head sums, beta folded into relu) is the same function up to rounding
    emu = og.value_policy_last_logits((pi, base, q1, q2), toks, cfg.n_head, cfg.n_vocab, 4.0, emulate_bf16=True)
    ref = og.value_policy_last_logits((pi, base, q1, q2), toks, cfg.n_head, cfg.n_vocab, 4.0)
    assert np.abs(emu - ref).max() < 0.2


def test_top_k_top_p_restatement_matches_hf_warpers():
    """The numpy restatement of the Flax warpers agrees with transformers' torch TopK / TopP warpers (same published
    semantics: the token that crosses top_p is kept, at least one token survives)."""
    import torch
    from transformers.generation.logits_process import TopKLogitsWarper, TopPLogitsWarper
    rng = np.random.default_rng(0)
    for trial in range(20):
        x = (rng.standard_normal(997) * 3).astype(np.float32)
        k = int(rng.integers(1, 50))
        p = float(rng.uniform(0.05, 0.99))
        tk = TopKLogitsWarper(top_k=k)(None, torch.from_numpy(x)[None].clone())[0].numpy()
        assert np.array_equal(np.isfinite(og.warp_top_k(x, k)), np.isfinite(tk))
        tp = TopPLogitsWarper(top_p=p)(None, torch.from_numpy(x)[None].clone())[0].numpy()
        mine = np.isfinite(og.warp_top_p(x, p))
        theirs = np.isfinite(tp)
        # identical up to the single boundary token when the cumulative mass lands within float32 rounding of top_p
        assert (mine != theirs).sum() <= 1, (trial, (mine != theirs).sum())
        assert mine.sum() >= 1
    # ties: the lower index wins in top_k; the tie that crosses top_p is kept, later ones are not
    x = np.array([1.0, 2.0, 2.0, 2.0, 0.0], np.float32)
    assert np.isfinite(og.warp_top_k(x, 2)).tolist() == [False, True, True, False, False]
    q = np.log(np.array([0.1, 0.3, 0.3, 0.3], np.float64)).astype(np.float32)
    assert np.isfinite(og.warp_top_p(q, 0.5)).tolist() == [False, True, True, False]


def test_value_forward_gathered_equals_full_heads():
    """q_k(h_t)[tok_{t+1}] computed from one dense2 column == the gather of the full MLPHead output
    (take_along_axis(qs[:, :-1], batch[:, 1:]), ilql/gpt2/score_fn.py:49); V and the linear PPO value head likewise."""
    import torch
    from lmrl_gym_b200.gpt2 import GPT2Config, init_gpt2_params, init_mlp_head_params
    from oracle import gpt2 as og
    cfg = GPT2Config(n_layer=2, n_head=2, d_model=64, d_ff=128, n_vocab=97, n_vocab_rows=98, n_pos=32)
    base = init_gpt2_params(cfg, seed=4)
    q1, q2 = init_mlp_head_params(cfg, seed=5), init_mlp_head_params(cfg, seed=6)
    v = init_mlp_head_params(cfg, seed=7, n_out=1, layer2_bias_init=0.3)
    toks = np.random.default_rng(0).integers(0, cfg.n_vocab, size=13).tolist()
    got = og.value_forward_gathered(base, q1, q2, v, toks, cfg.n_head)
    h = og.forward_hidden(base, toks, cfg.n_head)
    for name, q in (("q1", q1), ("q2", q2)):
        full = og.mlp_head(h, q).numpy()
        np.testing.assert_allclose(got[name], full[np.arange(12), toks[1:]], rtol=1e-5, atol=1e-5)
    # the engine keeps the V head's single dense2 column in fp32 (dense1 in bf16 like every GEMM operand)
    v_mixed = dict(dense1=dict(kernel=og.bf16_round(v["dense1"]["kernel"]), bias=v["dense1"]["bias"]), dense2=v["dense2"])
    np.testing.assert_allclose(got["v"], og.mlp_head(h, v_mixed, weights_bf16=False).numpy()[:, 0], rtol=1e-4, atol=1e-4)
    lin = dict(dense=dict(kernel=np.random.default_rng(1).standard_normal((64, 1)).astype(np.float32), bias=np.asarray([0.25], np.float32)))
    lv = og.linear_value_forward(base, lin, toks, cfg.n_head)
    np.testing.assert_allclose(lv, (h @ torch.from_numpy(lin["dense"]["kernel"]))[:, 0].numpy() + 0.25, rtol=1e-4, atol=1e-5)
    lp = og.token_logprobs_all(base, toks, cfg.n_head, cfg.n_vocab)
    assert lp.shape == (12,) and (lp < 0).all()


def test_linear_q_head_restatement():
    """MC returns: the Q head is a LinearHead (heads/linear_head.py:112-119: q = dense(h)); logits = pi_beta + beta * (h W + b)
    (generation.py:112-119 with q2 None), and the gathered per-position form equals the gather of the full head output."""
    cfg = GPT2Config(n_layer=1, n_head=2, d_model=64, d_ff=128, n_vocab=97, n_vocab_rows=98, n_pos=16)
    pi, base = init_gpt2_params(cfg, seed=1), init_gpt2_params(cfg, seed=2)
    rng = np.random.default_rng(3)
    W, b = (rng.standard_normal((64, 98)) / 8).astype(np.float32), rng.standard_normal(98).astype(np.float32)
    q, q_nobias = dict(dense=dict(kernel=W, bias=b)), dict(dense=dict(kernel=W, bias=None))
    toks = [5, 9, 2, 77, 30]
    pi_logits = og.forward_last_logits(pi, toks, cfg.n_head, cfg.n_vocab, weights_bf16=False)
    h = og.forward_hidden(base, toks, cfg.n_head, weights_bf16=False).numpy()
    want_q = h[-1] @ W + b
    got = og.value_policy_last_logits((pi, base, q, None), toks, cfg.n_head, cfg.n_vocab, 16.0, weights_bf16=False)
    assert np.allclose(got, pi_logits + 16.0 * want_q[:cfg.n_vocab], atol=2e-4)
    no_pi = og.value_policy_last_logits((None, base, q_nobias, None), toks, cfg.n_head, cfg.n_vocab, 2.0, weights_bf16=False)
    assert np.allclose(no_pi, 2.0 * (h[-1] @ W)[:cfg.n_vocab], atol=2e-4)
    two = og.value_policy_last_logits((pi, base, q, q_nobias), toks, cfg.n_head, cfg.n_vocab, 4.0, weights_bf16=False)
    assert np.allclose(two, pi_logits + 4.0 * np.minimum(want_q, h[-1] @ W)[:cfg.n_vocab], atol=2e-4)
    # the engine-emulating variant (beta folded into the bf16 kernel) is the same function up to rounding
    emu = og.value_policy_last_logits((pi, base, q, None), toks, cfg.n_head, cfg.n_vocab, 4.0, emulate_bf16=True)
    assert np.abs(emu - og.value_policy_last_logits((pi, base, q, None), toks, cfg.n_head, cfg.n_vocab, 4.0)).max() < 0.2
    g = og.value_forward_gathered(base, q, q_nobias, None, toks, cfg.n_head, weights_bf16=False)
    full = h @ W
    np.testing.assert_allclose(g["q1"], (full + b)[np.arange(4), toks[1:]], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(g["q2"], full[np.arange(4), toks[1:]], rtol=1e-5, atol=1e-5)


def test_layernorm_fold_identity_of_the_consumer_gemms():
    """The algebra behind ``ln_fold`` (csrc/gemm_tcgen05.cuh, fold_ln_weights_kernel): with W''[n][k] = gamma[k] W[k][n] - mean_k(gamma[k] W[k][n])
    and b'[n] = b[n] + sum_k beta[k] W[k][n], LayerNorm(x) @ W + b == rstd * (x @ W''^T) + b' for the RAW row x — the row mean drops out
    because every row of W'' sums to zero — and rstd follows from the row's (sum, sum of squares) alone."""
    rng = np.random.default_rng(0)
    K, N, M = 96, 40, 7
    x = (3.0 + 2.0 * rng.standard_normal((M, K))).astype(np.float64)            # non-zero mean on purpose
    gamma, beta = 1 + 0.2 * rng.standard_normal(K), 0.1 * rng.standard_normal(K)
    W, b = rng.standard_normal((K, N)) / 8, 0.1 * rng.standard_normal(N)         # HF Conv1D layout [K][N]
    eps = 1e-5
    mean, var = x.mean(axis=1, keepdims=True), x.var(axis=1, keepdims=True)
    ref = ((x - mean) / np.sqrt(var + eps) * gamma + beta) @ W + b
    gw = gamma[:, None] * W                                                        # [K][N]
    w_fold = (gw - gw.mean(axis=0, keepdims=True)).T                               # [N][K], rows sum to zero
    b_fold = b + beta @ W
    s1, s2 = x.sum(axis=1), (x * x).sum(axis=1)                                    # what the residual epilogue accumulates per row
    rstd = 1.0 / np.sqrt(np.maximum(s2 / K - (s1 / K) ** 2, 0.0) + eps)
    got = rstd[:, None] * (x @ w_fold.T) + b_fold
    assert np.abs(w_fold.sum(axis=1)).max() < 1e-12
    assert np.abs(got - ref).max() < 1e-10
