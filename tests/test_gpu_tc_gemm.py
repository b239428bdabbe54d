"""GPU: the tcgen05 3xTF32 GEMM kernel (csrc/tc_gemm.cu) in isolation against float64, and the per-step tensor-core
scans built on it (hidden sizes beyond the cluster-resident kernels) against the float64 oracle at batch sizes that
span several 128-row tiles.

Tolerance of the GEMM: 3xTF32 keeps ~21 mantissa bits per operand and accumulates in fp32; with unit-variance
operands the error of a K-term sum is well below 1e-5 * sqrt(K)."""
import numpy as np
import pytest

from oracle import sbr_oracle as O
from tests.test_gpu_parity import check_grads, _engine, _init, _batch_kwargs, _gpu_step, _okw
from tests.test_oracle import make_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from sbr_b200 import _capi
    e = _capi.Engine(n_items=16, cell="GRU", layers=(8,), max_length=4, batch_size=2)
    yield e
    e.close()


def _ref(A, B, ta, tb):
    a = A.astype(np.float64).T if ta else A.astype(np.float64)
    b = B.astype(np.float64).T if tb else B.astype(np.float64)
    return a @ b


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (130, 70, 45), (257, 384, 512), (64, 200, 3706), (1000, 16, 33),
                                   (128, 3706, 200), (300, 1024, 256)])
def test_gemm_matches_float64(eng, ta, tb, M, N, K):
    rng = np.random.RandomState(M + 7 * N + 13 * K + 2 * ta + tb)
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    C, _ = eng.debug_gemm(A, B, ta=ta, tb=tb, engine=1)
    ref = _ref(A, B, ta, tb)
    err = np.abs(C - ref).max()
    assert err <= 1e-5 * np.sqrt(K) + 1e-6, "max err %.3e (|C| max %.2f)" % (err, np.abs(ref).max())


def test_gemm_alpha_beta_bias_and_split_k(eng):
    rng = np.random.RandomState(3)
    # tall-K, small output: split over many CTAs, fp32 reductions into a pre-existing C (beta = 1)
    A = rng.standard_normal((9000, 200)).astype(np.float32)      # stored [K, M]
    B = rng.standard_normal((9000, 130)).astype(np.float32)      # stored [K, N]
    C0 = rng.standard_normal((200, 130)).astype(np.float32)
    C, _ = eng.debug_gemm(A, B, ta=True, tb=False, C0=C0, alpha=0.5, beta=1.0, engine=1)
    ref = 0.5 * (A.astype(np.float64).T @ B.astype(np.float64)) + C0
    assert np.abs(C - ref).max() <= 2e-3   # sums of 9000 products, magnitude ~100
    # beta = 0 with split-K (the output is cleared first), odd leading dimension
    C2, _ = eng.debug_gemm(A, B, ta=True, tb=False, engine=1)
    assert np.abs(C2 - A.astype(np.float64).T @ B.astype(np.float64)).max() <= 2e-3
    # bias broadcast over rows, fused in the epilogue
    X = rng.standard_normal((333, 96)).astype(np.float32)
    W = rng.standard_normal((96, 260)).astype(np.float32)
    b = rng.standard_normal(260).astype(np.float32)
    C3, _ = eng.debug_gemm(X, W, bias=b, engine=1)
    assert np.abs(C3 - (X.astype(np.float64) @ W.astype(np.float64) + b)).max() <= 1e-4


def test_gemm_small_magnitudes_keep_relative_accuracy(eng):
    """Gradients are tiny (1/global_batch scaling): the split must not lose them (tf32 keeps the fp32 exponent)."""
    rng = np.random.RandomState(5)
    A = (rng.standard_normal((256, 512)) * 1e-6).astype(np.float32)
    B = (rng.standard_normal((512, 256)) * 1e-3).astype(np.float32)
    C, _ = eng.debug_gemm(A, B, engine=1)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    assert np.abs(C - ref).max() <= 1e-5 * np.abs(ref).max()


def test_cp_async_loaders_match_the_tma_loaders(monkeypatch):
    """SBR_DISABLE_TMA_GEMM: the raw ring is filled by cp.async from 128 threads per operand instead of TMA tiled loads
    (the path unaligned operands take); same raw layouts, same results."""
    from sbr_b200 import _capi
    rng = np.random.RandomState(11)
    cases = []
    for ta, tb, M, N, K in [(False, True, 300, 260, 512), (True, False, 256, 96, 1000), (False, False, 200, 1024, 256), (True, True, 64, 48, 96)]:
        A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
        B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
        cases.append((ta, tb, A, B))
    outs = []
    for env in (None, "1"):
        if env:
            monkeypatch.setenv("SBR_DISABLE_TMA_GEMM", env)
        e = _capi.Engine(n_items=16, cell="GRU", layers=(8,), max_length=4, batch_size=2)
        try:
            outs.append([e.debug_gemm(A, B, ta=ta, tb=tb, engine=1)[0] for ta, tb, A, B in cases])
        finally:
            e.close()
    for (ta, tb, A, B), c_tma, c_cp in zip(cases, *outs):
        ref = _ref(A, B, ta, tb)
        assert np.abs(c_tma - ref).max() <= 1e-5 * np.sqrt(A.shape[0] if ta else A.shape[1]) + 1e-6
        np.testing.assert_allclose(c_tma, c_cp, rtol=0, atol=1e-4)       # same arithmetic (split-K sums land in any order)


def test_ffma_and_tensor_core_engines_agree(eng):
    rng = np.random.RandomState(9)
    A = rng.standard_normal((500, 300)).astype(np.float32)
    B = rng.standard_normal((700, 300)).astype(np.float32)
    C1, _ = eng.debug_gemm(A, B, tb=True, engine=1)
    C0, _ = eng.debug_gemm(A, B, tb=True, engine=0)
    assert np.abs(C1 - C0).max() <= 2e-4


# ---------------------------------------------------------------------------------------- per-step tensor-core scans
@pytest.mark.parametrize("cell,layers,B,T", [("LSTM", (256,), 160, 7), ("GRU", (512,), 130, 6), ("LSTM", (256, 256), 140, 6),
                                             ("GRU", (512, 512), 256, 5), ("Vanilla", (320,), 40, 6), ("LSTM", (512,), 33, 5),
                                             ("GRU", (240,), 64, 6)])
@pytest.mark.parametrize("persistent", [True, "no-split-k", False])
def test_step_scans_match_oracle(cell, layers, B, T, persistent, monkeypatch):
    """H > 224 (H % 16 == 0): the persistent cooperative tensor-core scans (tc_scan.cu, one launch per layer), or with
    SBR_DISABLE_PERSISTENT_SCAN one tcgen05 step kernel per time step (tc_gemm.cu); forward and BPTT; batch sizes that
    are not multiples of the 128-row / 32-column tiles."""
    if not persistent:
        monkeypatch.setenv("SBR_DISABLE_PERSISTENT_SCAN", "1")
    elif persistent == "no-split-k":
        monkeypatch.setenv("SBR_DISABLE_SPLITK_SCAN", "1")      # BPTT: one CTA per tile instead of a split-K cluster of 4
    spec = O.Spec(n_items=173, cell=cell, layers=layers, loss="CCE")
    check_grads(spec, B=B, T=T, seed=len(layers) + B)


def test_step_scans_really_run_and_agree_with_the_ffma_fallback(monkeypatch):
    """The step kernels are what runs (launch count grows with the number of time steps), and the FFMA cluster scans
    (SBR_DISABLE_STEP_SCAN) give the same cost and gradients on the same batch."""
    spec = O.Spec(n_items=97, cell="LSTM", layers=(256,), loss="CCE")
    B, T = 48, 9
    rng, vals = _init(spec, 4)
    X, mask, lens = make_batch(rng, B, T, spec.n_items, 1, 0)
    kw = _batch_kwargs(spec, rng, B, spec.n_items, X, lens)

    def run():
        e = _engine(spec, B, T)
        try:
            e.set_all_param_values(vals)
            e.set_skip_update(True)
            n0 = e.kernel_launches()
            c = _gpu_step(e, spec, X, mask, kw)
            return float(c), e.get_all_grads(), e.kernel_launches() - n0
        finally:
            e.close()

    monkeypatch.setenv("SBR_DISABLE_PERSISTENT_SCAN", "1")
    c1, g1, n1 = run()
    monkeypatch.setenv("SBR_DISABLE_STEP_SCAN", "1")
    c0, g0, n0 = run()
    t_max = int(lens.max())
    assert n1 >= n0 + 2 * t_max - 4, (n1, n0, t_max)      # one launch per step, forward and backward
    assert abs(c1 - c0) <= 2e-5
    for a, b in zip(g1, g0):
        assert np.abs(a - b).max() <= 2e-4 * np.abs(b).max() + 1e-7


def test_sampled_and_margin_losses_on_the_tensor_core_gemms():
    """The gathered-column (BPR) and full-catalog margin outputs run their GEMMs on tc_gemm too."""
    check_grads(O.Spec(n_items=301, cell="LSTM", layers=(256, 256), loss="BPR"), B=70, T=6, seed=21)
    check_grads(O.Spec(n_items=333, cell="LSTM", layers=(512,), loss="hinge"), B=36, T=5, seed=22)
    check_grads(O.Spec(n_items=120, cell="GRU", layers=(64,), loss="TOP1", embedding=24), B=20, T=7, seed=23)
