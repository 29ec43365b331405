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


@pytest.mark.parametrize("rows,K,N", [(16, 128, 32), (37, 64, 16), (160, 64, 64), (1024, 64, 4), (1261, 128, 256)])
def test_tc_mn_major_operand_path(cuda_device, rows, K, N):
    """The dW kernel's operand path (csrc/tc.cuh make_desc_mn_sw128): C = H^T D with both operands fetched as
    swizzled {32 column, 16 row} TMA boxes and consumed MN-major, no software transposition.  Small integers are
    exact in tf32, so the plain-TF32 product must be exact (any layout / swizzle / descriptor mistake is an integer
    off); the 3xTF32 product (hi = the tile as it landed, lo = x - trunc(x)) stays at fp32-GEMM accuracy."""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(rows + K + N)
    H = torch.randint(-4, 5, (rows, K), generator=g).float().to(cuda_device)
    D = torch.randint(-4, 5, (rows, N), generator=g).float().to(cuda_device)
    C = torch.full((K, N), float("nan"), device=cuda_device)
    _lib.check(lib.b200ppo_tc_mn_test(_lib.current_stream(), H.data_ptr(), D.data_ptr(), C.data_ptr(), rows, K, N, 0, None))
    torch.cuda.synchronize()
    assert torch.equal(C.double(), H.double().T @ D.double())
    H = torch.randn(rows, K, generator=g).to(cuda_device)
    D = torch.randn(rows, N, generator=g).to(cuda_device)
    ref = H.double().T @ D.double()
    err = {}
    for split in (1, 0):
        C = torch.full((K, N), float("nan"), device=cuda_device)
        _lib.check(lib.b200ppo_tc_mn_test(_lib.current_stream(), H.data_ptr(), D.data_ptr(), C.data_ptr(), rows, K, N, split, None))
        torch.cuda.synchronize()
        err[split] = (C.double() - ref).abs().max().item()
    scale = max(ref.abs().max().item(), 1.0)
    print(f"rows={rows} K={K} N={N} max|err| 3xTF32 {err[1]:.3e} TF32 {err[0]:.3e} scale {scale:.1f}")
    assert err[1] < 2e-5 * scale and err[1] < err[0] / 20 and err[0] < 3e-3 * scale
