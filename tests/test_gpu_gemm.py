"""GPU parity of the GEMM building block: the tcgen05 3xTF32 kernel (engine 1) and the FFMA
kernel (engine 0) against a float64 matmul.  Tolerance: 2e-6 relative L2 (fp32 accumulate)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from deepconvsep_b200.engine import Context
    return Context(0)


@pytest.mark.parametrize("engine", [1, 0])
@pytest.mark.parametrize("M,N,K,lda_extra", [(128, 64, 32, 0), (1, 1, 1, 0), (300, 50, 750, 0), (257, 128, 800, 4),
                                             (1000, 150, 129, 3), (130, 2400, 128, 0), (4096, 50, 1025, 7)])
def test_gemm_matches_float64(ctx, engine, M, N, K, lda_extra):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K + lda_extra)).astype(np.float32)
    B = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32) * 0.1
    Ad = torch.tensor(A, device="cuda")[:, :K]
    for relu in (False, True):
        C = ctx.gemm(Ad, B, bias, relu=relu, engine=engine).cpu().numpy().astype(np.float64)
        ref = A[:, :K].astype(np.float64) @ B.astype(np.float64) + bias
        if relu:
            ref = np.maximum(ref, 0)
        err = np.linalg.norm(C - ref) / np.linalg.norm(ref)
        assert err < 2e-6, (engine, relu, err)
        assert np.abs(C - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("pitch", [50, 52])
def test_tensor_core_gemm_overlapping_rows(ctx, pitch):
    """A rows that overlap (row stride < K): the convolution-as-GEMM view (no im2col copy).
    pitch 52 (16-byte aligned rows) is what the TMA-fed kernel takes -- a tensor map whose row pitch
    is smaller than its row extent; pitch 50 stays on the register-staged kernel."""
    rng = np.random.default_rng(1)
    buf = rng.standard_normal(pitch * 400).astype(np.float32)
    K, M, N = 15 * pitch, 300, 50
    d = torch.tensor(buf, device="cuda")
    Ad = torch.as_strided(d, (M, K), (pitch, 1))
    B = (rng.standard_normal((K, N)) / 27.0).astype(np.float32)
    ref = np.lib.stride_tricks.as_strided(buf, (M, K), (4 * pitch, 4)).astype(np.float64) @ B.astype(np.float64)
    for engine in (1, 0):
        C = ctx.gemm(Ad, B, None, engine=engine).cpu().numpy().astype(np.float64)
        assert np.linalg.norm(C - ref) / np.linalg.norm(ref) < 2e-6


def test_tma_and_register_staged_kernels_agree(monkeypatch):
    """the copy-engine-fed kernel (raw fp32 tile as the high tf32 operand, low plane derived in shared
    memory) against the register-staged one (explicit hi/lo split by the producer warps): same
    products, same accumulator plan -> same result up to the order of the split-K partial sums"""
    from deepconvsep_b200.engine import Context
    rng = np.random.default_rng(5)
    cases = [(640, 50, 780, 52, 780), (3000, 128, 832, 260, 832), (500, 2496, 128, 128, 128), (1000, 50, 1025, 1032, 1025)]
    monkeypatch.setenv("DCS_DEBUG_TMA", "0")
    staged = Context(0)
    monkeypatch.setenv("DCS_DEBUG_TMA", "1")
    fed = Context(0)
    for M, N, K, pitch, width in cases:
        buf = torch.tensor(rng.standard_normal(pitch * (M - 1) + width + 8).astype(np.float32), device="cuda")
        A = torch.as_strided(buf, (M, K), (pitch, 1))
        B = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
        bias = rng.standard_normal(N).astype(np.float32)
        a = staged.gemm(A, B, bias, relu=True, engine=1).cpu().numpy().astype(np.float64)
        b = fed.gemm(A, B, bias, relu=True, engine=1).cpu().numpy().astype(np.float64)
        assert np.linalg.norm(a - b) <= 1e-6 * np.linalg.norm(a), (M, N, K)
