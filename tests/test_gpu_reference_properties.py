"""GPU twin of tests/test_reference_properties_cpu.py: the reference's property tests and oracle parity
on its four structure / model fixtures through the C-ABI: 17 / 13 basis functions with 100 neurons (carbon), the 4-body term
without the 5-body one (water), three types with ZBL (BaTiO3), all small boxes.  First run on a B200
in round 2 (gpurun_out/r02_a_pytest_ext_only.txt): 18 of 20 green; the two failures were test
conditioning, not kernels -- see check_nep (total-energy tolerance on a cancelling sum, force atol
tighter than the FP32 oracle's own distance from FP64) and scripts/diag_r02.py for the numbers.
"""
import pytest

import test_reference_properties_cpu as P
from conftest import GOLDEN
from test_kernel_bodies_cpu import check_nep

pytestmark = pytest.mark.gpu


class _Gpu:
    """Same surface as emu_py.Emu.nep over the real library."""

    def __init__(self, eng):
        self.eng = eng

    def nep(self, path, n):
        from test_gpu_parity import GpuNep
        return GpuNep(self.eng, path.name, n)


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from gpumd_b200 import engine
    return _Gpu(engine)


@pytest.mark.parametrize("name", list(P.PAIRS))
def test_matches_oracle(oracle, gpu, name):
    model, s = P.load(name)
    n = s["type"].shape[0]
    e_scale = {"C": 8.0}.get(name, 1.0)
    check_nep(oracle, gpu.nep(GOLDEN / model, n), model, s, n, energy_tol=1e-6 * e_scale)


@pytest.mark.parametrize("name", list(P.PAIRS))
def test_translation_and_lattice_shift_invariance(gpu, name):
    P.test_translation_and_lattice_shift_invariance(gpu, name)


@pytest.mark.parametrize("name", list(P.PAIRS))
def test_rotation_invariance(gpu, name):
    P.test_rotation_invariance(gpu, name)


@pytest.mark.parametrize("name", list(P.PAIRS))
def test_permutation_invariance(gpu, name):
    P.test_permutation_invariance(gpu, name)


@pytest.mark.parametrize("name", list(P.PAIRS))
def test_finite_difference_forces(gpu, name):
    P.test_finite_difference_forces(gpu, name)
