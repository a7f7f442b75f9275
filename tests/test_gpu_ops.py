"""GPU parity of the isolated libdv3 ops against the NumPy oracle (bit-exact for indices)."""
import numpy as np
import pytest
import torch

from oracle import numerics as nx

pytestmark = pytest.mark.gpu


def dev(x, dtype=torch.float32):
    return torch.as_tensor(np.asarray(x), dtype=dtype).cuda()


def test_symlog_bit_exact_and_inverse():
    from dreamer_v3_b200 import ops
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(100000) * 10 ** rng.uniform(-4, 6, 100000),
                        [0.0, -0.0, 1.0, -1.0, 1e-8, 3.4e38, -3.4e38]]).astype(np.float32)
    y = ops.symlog(dev(x)).cpu().numpy()
    np.testing.assert_array_equal(y, nx.symlog(x))  # fp64 log rounded once on both sides
    xr = ops.inverse_symlog(dev(y[:100000])).cpu().numpy()
    np.testing.assert_allclose(xr, nx.inverse_symlog(y[:100000]), rtol=1e-5, atol=3e-7)


def test_two_hot_bit_exact_indices():
    from dreamer_v3_b200 import ops
    rng = np.random.default_rng(1)
    delta = np.float32(40.0 / 254.0)
    grid = (np.float32(-20.0) + np.arange(255, dtype=np.float32) * delta).astype(np.float32)
    v = np.concatenate([rng.uniform(-25, 25, 200000), grid, np.nextafter(grid, np.float32(100)),
                        np.nextafter(grid, np.float32(-100)), [20.1, 50.0, 150.0, -20.00001, 0.0, 1.0, -1.0]]).astype(np.float32)
    k, kp1, wk, wkp1, dense = ops.two_hot(dev(v))
    ek, ekp1, ewk, ewkp1 = nx.two_hot_indices(v)
    np.testing.assert_array_equal(k.cpu().numpy(), ek)
    np.testing.assert_array_equal(kp1.cpu().numpy(), ekp1)
    np.testing.assert_array_equal(wk.cpu().numpy(), ewk)      # IEEE ops only -> identical weights too
    np.testing.assert_array_equal(wkp1.cpu().numpy(), ewkp1)
    np.testing.assert_array_equal(dense.cpu().numpy(), nx.two_hot(v))
    # reference docstring vector in the 255-bucket setting: weights sum to one, decode recovers value
    d = dense.cpu().numpy()
    np.testing.assert_allclose(d.sum(-1), 1.0, atol=1e-5)


def test_sample_categorical_bit_exact():
    from dreamer_v3_b200 import ops
    rng = np.random.default_rng(2)
    for K in (32, 18, 6, 4, 2):
        logits = (rng.standard_normal((50000, K)) * 3).astype(np.float32)
        p = np.exp(logits - logits.max(-1, keepdims=True))
        p = (np.float32(0.99) * (p / p.sum(-1, keepdims=True)).astype(np.float32) + np.float32(0.01 / K)).astype(np.float32)
        u = rng.random(50000, dtype=np.float32)
        u[:5] = [0.0, np.float32(1) - np.float32(2 ** -24), 0.5, 0.25, 0.75]
        idx = ops.sample_categorical(dev(p), dev(u)).cpu().numpy()
        np.testing.assert_array_equal(idx, nx.sample_categorical(p, u))


def test_value_targets_and_percentiles():
    from dreamer_v3_b200 import ops
    rng = np.random.default_rng(3)
    H, N = 15, 1024
    r = rng.standard_normal((H + 1, N)).astype(np.float32)
    c = (rng.random((H + 1, N)) > 0.1).astype(np.float32)
    v = (rng.standard_normal((H + 1, N)) * 5).astype(np.float32)
    got = ops.value_targets(dev(r), dev(c), dev(v), 0.997, 0.95).cpu().numpy()
    np.testing.assert_allclose(got, nx.compute_value_targets(r, c, v, 0.997, 0.95), rtol=1e-5, atol=1e-5)
    # critic_loss.py:142-160 example
    r1 = np.array([[99.0], [1.0], [2.0], [3.0], [4.0], [5.0]], np.float32)
    c1 = np.array([[1.0], [1.0], [0.0], [1.0], [1.0], [1.0]], np.float32)
    v1 = np.array([[3.0], [2.0], [15.0], [12.0], [8.0], [3.0]], np.float32)
    np.testing.assert_allclose(ops.value_targets(dev(r1), dev(c1), dev(v1), 1.0, 1.0).cpu().numpy()[:, 0], [3, 2, 15, 12, 8], atol=1e-5)
    for n in (15360, 1000, 17, 8 * 15360):
        x = (rng.standard_normal(n) * 3).astype(np.float32)
        p = ops.percentiles(dev(x)).cpu().numpy()
        assert p[0] == nx.percentile_nearest(x, 5) and p[1] == nx.percentile_nearest(x, 95)


@pytest.mark.parametrize("M,N,K", [(16, 768, 256), (16, 1024, 1280), (1024, 256, 1280), (1024, 255, 256), (1, 256, 1024),
                                   (1024, 1, 256), (333, 77, 45), (4096, 512, 48), (128, 128, 16), (257, 129, 1031),
                                   # degenerate shapes of the heads: few outputs (skinny-N kernel) / rank-few updates (skinny-K)
                                   (1000, 18, 256), (16384, 1, 256), (1024, 256, 1), (1024, 1024, 6), (300, 32, 1280), (64, 40, 32)])
def test_gemm_f32_all_transposes(M, N, K):
    from dreamer_v3_b200 import ops
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    for ta in (False, True):
        for tb in (False, True):
            A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
            B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
            bias = rng.standard_normal(N).astype(np.float32)
            C0 = rng.standard_normal((M, N)).astype(np.float32)
            want = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64)
            got = ops.gemm(dev(A), dev(B), ta, tb, bias=dev(bias)).cpu().numpy()
            np.testing.assert_allclose(got, want + bias, rtol=0, atol=2e-5 * np.sqrt(K) * 4)
            got = ops.gemm(dev(A), dev(B), ta, tb, C_=dev(C0), beta=1.0).cpu().numpy()
            np.testing.assert_allclose(got, want + C0, rtol=0, atol=2e-5 * np.sqrt(K) * 4)


def test_gemm_strided_views():
    from dreamer_v3_b200 import ops
    rng = np.random.default_rng(9)
    big = dev(rng.standard_normal((64, 300)).astype(np.float32))
    W = dev(rng.standard_normal((100, 40)).astype(np.float32))
    A = big[:, 7:107]  # unaligned column slice with row stride 300
    got = ops.gemm(A, W).cpu().numpy()
    np.testing.assert_allclose(got, A.cpu().numpy().astype(np.float64) @ W.cpu().numpy().astype(np.float64), atol=1e-4)


def test_ln_silu():
    from dreamer_v3_b200 import ops
    rng = np.random.default_rng(4)
    # 255: scalar warp-per-row kernel; 4 ... 192 (D % 4 == 0): the several-rows-per-warp kernels of the conv channel counts;
    # multiples of 128: the vectorised warp-per-row kernel
    for D, rows in ((4, 7), (16, 300), (24, 300), (32, 4099), (48, 1001), (64, 300), (96, 4099), (192, 301), (255, 300), (256, 300),
                    (640, 300), (1024, 300)):
        x = (rng.standard_normal((rows, D)) * 3 + 1).astype(np.float32)
        g = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
        b = (0.1 * rng.standard_normal(D)).astype(np.float32)
        want = torch.nn.functional.silu(torch.nn.functional.layer_norm(
            torch.from_numpy(x).double(), (D,), torch.from_numpy(g).double(), torch.from_numpy(b).double(), 1e-3)).numpy()
        got = ops.ln_silu(dev(x), dev(g), dev(b)).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("tb", (False, True))
@pytest.mark.parametrize("M,N,K", ((16, 12288, 1024), (16, 1030, 1302), (5, 1024, 4099), (1, 4096, 517), (16, 96, 12288)))
@pytest.mark.parametrize("beta,with_bias", ((0.0, True), (1.0, False)))
@pytest.mark.parametrize("mode", (0, 1))
def test_weight_streaming_contraction_of_the_scan_rows(M, N, K, tb, beta, with_bias, mode):
    """dv3_gemv.cu: <= 16 rows against a large weight matrix (the per-op observe scan at the large sizes), both weight
    layouts, ragged N / K, unaligned leading dimensions, split-K accumulation; fp32 in every mode."""
    import torch
    from dreamer_v3_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    A = torch.randn((M, K), device="cuda", generator=g)
    B = torch.randn((N, K) if tb else (K, N), device="cuda", generator=g)
    bias = torch.randn((N,), device="cuda", generator=g) if with_bias else None
    C0 = torch.randn((M, N), device="cuda", generator=g)
    want = A.double() @ (B.double().t() if tb else B.double()) + (bias.double() if with_bias else 0.0) + beta * C0.double()
    got = ops.gemm(A, B, tb=tb, bias=bias, C_=C0.clone(), beta=beta, mode=mode)
    err = float((got.double() - want).abs().max() / want.abs().max())
    assert err < 2e-6, err


@pytest.mark.parametrize("D", (24, 32, 48, 96, 192, 256, 384, 512, 640, 768, 1024))
@pytest.mark.parametrize("rows", (16, 300, 4099))
def test_ln_silu_bwd_isolated(D, rows):
    """Backward of Dense -> LayerNorm(eps 1e-3) -> SiLU (components/mlp.py:69-84) in isolation at every layer width of the
    size table (dense 256...1024) and the conv channel counts (24...768): dx, dgamma, dbeta against torch autograd in fp64."""
    from dreamer_v3_b200 import ops
    rng = np.random.default_rng(D * 31 + rows)
    x = (rng.standard_normal((rows, D)) * 2 + 0.5).astype(np.float32)
    dy = rng.standard_normal((rows, D)).astype(np.float32)
    g = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
    b = (0.1 * rng.standard_normal(D)).astype(np.float32)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    gt = torch.from_numpy(g).double().requires_grad_(True)
    bt = torch.from_numpy(b).double().requires_grad_(True)
    y = torch.nn.functional.silu(torch.nn.functional.layer_norm(xt, (D,), gt, bt, 1e-3))
    y.backward(torch.from_numpy(dy).double())
    dg = torch.zeros(D, device="cuda")
    db = torch.zeros(D, device="cuda")
    dx = ops.ln_silu_bwd(dev(dy), dev(x), dev(g), dev(b), dg, db)()
    torch.cuda.synchronize()
    for name, got, want in (("dx", dx, xt.grad), ("dgamma", dg, gt.grad), ("dbeta", db, bt.grad)):
        err = float((got.cpu().double() - want).abs().max() / (want.abs().max() + 1e-30))
        assert err < 1e-4, (name, err)
