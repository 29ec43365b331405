"""tcgen05 building blocks: C = A B on the tensor cores (TF32 and error-compensated 3xTF32)
against a float64 matmul."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from nnx_ppo_b200 import _lib  # noqa: E402


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 256, 256), (300, 16, 64), (1000, 256, 64),
                                    (16896, 256, 256), (257, 64, 40), (64, 32, 8)])
def test_tc_gemm_3xtf32_matches_float64(cuda_device, M, N, K):
    from nnx_ppo_b200 import build
    build.build()
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(cuda_device)
    B = (torch.randn(K, N, generator=g) / K ** 0.5).to(cuda_device)
    ref = (A.double() @ B.double())
    scale = ref.abs().max().item()
    torch.backends.cuda.matmul.allow_tf32 = False
    err32 = ((A @ B).double() - ref).abs().max().item()       # cuBLAS fp32 (FFMA) error, for scale
    for split, tol in ((1, 4e-6), (0, 3e-3)):
        C = torch.full((M, N), float("nan"), device=cuda_device)
        _lib.check(lib.b200ppo_tc_gemm_test(_lib.current_stream(), A.data_ptr(), B.data_ptr(), C.data_ptr(),
                                            M, N, K, split), "tc_gemm_test")
        torch.cuda.synchronize()
        err = (C.double() - ref).abs().max().item()
        print(f"M={M} N={N} K={K} split={split} max|err|={err:.3e} fp32-cublas={err32:.3e} scale={scale:.2f}")
        assert np.isfinite(err) and err < tol * max(scale, 1.0), (split, err, scale)
        if split:   # error-compensated path stays within a small factor of true-fp32 FFMA accuracy
            assert err < 16 * max(err32, 1e-7), (err, err32)


def test_tc_microbench_runs(cuda_device):
    """Profiling aid: cycles per dependent tcgen05.mma and per bulk copy (numbers in profiles/)."""
    lib = _lib.load()
    src = torch.randn(1 << 18, device=cuda_device)
    out = torch.zeros(4, dtype=torch.int64, device=cuda_device)
    _lib.check(lib.b200ppo_tc_microbench(_lib.current_stream(), src.data_ptr(), out.data_ptr(), 64, 256, 16384, 4, 1))
    torch.cuda.synchronize()
    o = out.cpu().tolist()
    assert all(v > 0 for v in o) and o[0] / 256 < 1000
