"""CPU check of the algebra behind the norm-folded GEMMs (csrc/gemm.cuh GemmArgs::row_stats,
csrc/ops.cu fold_norm_kernel / row_stats_kernel): with W' = W diag(g), bias' = bias + W b, colsum[n] = sum_k W'[n, k]

    Linear(LayerNorm(x)) = rstd * (x W'^T) + (-mean * rstd) * colsum + bias'
    Linear(RMSNorm(x))   = rstd * (x W'^T) + bias

evaluated in float64 so that only the identity itself is tested; and the pivoted one-pass variance the statistics
kernel uses stays exact where the textbook E[x^2] - mean^2 loses every digit."""
import numpy as np


def test_layernorm_fold_identity():
    rng = np.random.default_rng(0)
    rows, K, N, eps = 37, 96, 40, 1e-6
    x = rng.normal(size=(rows, K)) * 3 + 0.7
    g, b = rng.normal(size=K) * 0.3 + 1, rng.normal(size=K) * 0.2
    W, bias = rng.normal(size=(N, K)) / np.sqrt(K), rng.normal(size=N)
    mean = x.mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(x.var(-1, keepdims=True) + eps)
    ref = ((x - mean) * rstd * g + b) @ W.T + bias
    Wf, bias_f = W * g, bias + W @ b
    colsum = Wf.sum(-1)
    got = rstd * (x @ Wf.T) + (-mean * rstd) * colsum + bias_f
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)


def test_rmsnorm_fold_identity():
    rng = np.random.default_rng(1)
    rows, K, N, eps = 29, 64, 24, 1e-6
    x = rng.normal(size=(rows, K)) * 2
    g = rng.normal(size=K) * 0.1 + 1
    W, bias = rng.normal(size=(N, K)) / np.sqrt(K), rng.normal(size=N)
    rstd = 1.0 / np.sqrt((x * x).mean(-1, keepdims=True) + eps)
    ref = (x * rstd * g) @ W.T + bias
    got = rstd * (x @ (W * g).T) + bias
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)


def test_pivoted_variance_is_well_conditioned_in_fp32():
    """Rows with |mean| >> std: sums of (x - p) and (x - p)^2 around the row's first element keep fp32 accuracy."""
    rng = np.random.default_rng(2)
    x = (1000.0 + rng.normal(size=(16, 1024)) * 0.01).astype(np.float32)
    true_var = x.astype(np.float64).var(-1)
    p = x[:, :1]
    d = x - p
    m1 = d.sum(-1, dtype=np.float32) / np.float32(1024)
    var_pivot = (d * d).sum(-1, dtype=np.float32) / np.float32(1024) - m1 * m1
    var_naive = (x * x).sum(-1, dtype=np.float32) / np.float32(1024) - (x.sum(-1, dtype=np.float32) / np.float32(1024)) ** 2
    assert np.allclose(var_pivot, true_var, rtol=1e-3)
    assert not np.allclose(var_naive, true_var, rtol=0.5)   # the textbook form is off by orders of magnitude here
