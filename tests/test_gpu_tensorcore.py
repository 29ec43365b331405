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


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 256, 1024), (64, 16, 96), (132, 256, 2048)])
def test_tc_gemm_mn_major_operands(cuda_device, M, N, K):
    """C = At^T B with both operands MN-major in shared memory (the dW operand orientation)."""
    from nnx_ppo_b200 import build
    build.build()
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N + K)
    At = torch.randn(K, M, generator=g).to(cuda_device)
    B = (torch.randn(K, N, generator=g) / K ** 0.5).to(cuda_device)
    ref = At.double().t() @ B.double()
    scale = ref.abs().max().item()
    for variant in (1, 2, 3):
        C = torch.full((M, N), float("nan"), device=cuda_device)
        lib.b200ppo_tc_gemm_tn_test(_lib.current_stream(), At.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, 1 | (variant << 4))
        torch.cuda.synchronize()
        print(f"MN-major variant={variant} M={M} N={N} K={K} max|err|={(C.double() - ref).abs().max().item():.3e}")
    for split, tol in ((1, 6e-6), (0, 4e-3)):
        C = torch.full((M, N), float("nan"), device=cuda_device)
        _lib.check(lib.b200ppo_tc_gemm_tn_test(_lib.current_stream(), At.data_ptr(), B.data_ptr(), C.data_ptr(),
                                               M, N, K, split), "tc_gemm_tn_test")
        torch.cuda.synchronize()
        err = (C.double() - ref).abs().max().item()
        print(f"MN-major M={M} N={N} K={K} split={split} max|err|={err:.3e} scale={scale:.2f}")
        assert np.isfinite(err) and err < tol * max(scale, 1.0), (split, err, scale)
