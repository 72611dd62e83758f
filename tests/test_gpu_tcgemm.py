"""-m gpu: the tcgen05 GEMM engine (csrc/tcgemm.cuh) on plain matrices, all four operand-major combinations,
against float64 numpy.  TF32 x3 must be fp32-grade (<= 1e-5 of ||C||inf; measured ~2.5e-6 at K=256..1024); single-pass TF32 is ~1e-3."""
import numpy as np
import pytest

from tests.gpu_util import need_gpu, small_learner, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu():
    need_gpu()


@pytest.fixture(scope='module')
def learner():
    need_gpu()
    L = small_learner(64, batch=32, use_per=False)
    yield L
    L.close()


SHAPES = [(128, 32, 32), (128, 64, 64), (256, 64, 256), (300, 32, 96), (128, 128, 40), (1000, 192, 1024), (64, 32, 36)]


@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('ak,bk', [(True, True), (True, False), (False, True), (False, False)])
def test_tc3x_matches_fp64(learner, M, N, K, ak, bk):
    if not ak and M % 4:
        pytest.skip('MN-contiguous A needs M % 4 == 0')
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.normal(0, 1, (M, K)).astype(np.float32); B = rng.normal(0, 1, (N, K)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    c0 = learner.selftest_gemm('fp32', A, B, ak, bk)
    assert rel_err(c0, ref) < 2e-6
    c3 = learner.selftest_gemm('tc3x', A, B, ak, bk)
    assert rel_err(c3, ref) < 1e-5, rel_err(c3, ref)
    c1 = learner.selftest_gemm('tc1x', A, B, ak, bk)
    assert 1e-6 < rel_err(c1, ref) < 3e-3, rel_err(c1, ref)


def test_tc_identity_rows(learner):
    """Layout check with exactly representable data: C = A @ I must reproduce A's columns bit for bit."""
    M, K, N = 256, 64, 64
    A = np.arange(M * K, dtype=np.float32).reshape(M, K) % 251
    B = np.eye(N, K, dtype=np.float32)
    for ak in (True, False):
        for bk in (True, False):
            for mode in ('tc1x', 'tc3x'):
                c = learner.selftest_gemm(mode, A, B, ak, bk)
                assert np.array_equal(c, A[:, :N]), (ak, bk, mode)
