"""tcgen05 / TMEM layer (gpumd_b200/csrc/b2_tc.cuh): the 3xTF32 tensor-core GEMM the NEP hidden layer is
built on must reproduce an FP64 host GEMM to FP32 accuracy for every operand shape the MLP uses."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("N,K", [(16, 8), (32, 32), (32, 56), (48, 32), (64, 64), (96, 80), (128, 96),
                                 (256, 64)])
def test_tc_gemm_matches_host(b200md_lib, N, K, layout):
    import torch
    from gpumd_b200 import lib as L
    rng = np.random.default_rng(N * 1000 + K)
    A = rng.standard_normal((128, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    # entries with a wide dynamic range so that a TF32-only product would fail the tolerance
    A *= np.exp(rng.uniform(-3, 3, A.shape)).astype(np.float32)
    dA = torch.from_numpy(A).cuda()
    dB = torch.from_numpy(B).cuda()
    dD = torch.full((128, N), float("nan"), dtype=torch.float32, device="cuda")
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.b200md_tc_selftest(layout, N, K, dA.data_ptr(), dB.data_ptr(), dD.data_ptr(), st))
    torch.cuda.synchronize()
    D = dD.cpu().numpy().astype(np.float64)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    err = np.abs(D - ref) / scale
    # 3xTF32 drops only the lo*lo term (2^-22 relative) + FP32 accumulation
    assert np.isfinite(D).all()
    assert err.max() < 2e-6, f"max scaled error {err.max():.3e}"
