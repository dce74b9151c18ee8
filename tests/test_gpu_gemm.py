"""GPU: the tcgen05 + TMA contraction kernel (3xTF32 split) against a float64 numpy reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 256), (256, 384, 512), (200, 130, 100), (1026, 1026, 96), (37, 5, 9)])
def test_gemm_nt_tf32x3(engine, M, N, K):
    rng = np.random.RandomState(M * 7 + N * 3 + K)
    A = rng.randn(M, K).astype(np.float32)
    B = (rng.randn(N, K) * rng.lognormal(0, 2, (N, 1))).astype(np.float32)      # rows of very different scale
    C = engine.debug_gemm_nt(A, B)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T         # sum |a||b| bounds the rounding
    err = np.abs(C - ref) / np.maximum(scale, 1e-30)
    assert err.max() < 2e-6, err.max()            # fp32-faithful: ~2^-22 split error + fp32 accumulation
    # plain TF32 (10-bit mantissa) would sit at ~1e-3: make sure the split really is in effect
    assert np.median(err) < 5e-7
