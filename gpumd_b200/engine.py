"""Python mirror of GPUMD's plugin surface for the hot path, over the libb200md C-ABI.

Class and method names follow the reference (paths relative to the GPUMD tree):
  Box            src/model/box.cuh:18-35
  Atom           src/model/atom.cuh:21-52        (SoA float64 device arrays)
  Potential      src/force/potential.cuh:20-113  (compute(box, type, position, potential, force, virial))
  NEP, LJ        src/force/nep.cuh:27-184, src/force/lj.cuh:31-49
  Force          src/force/force.cuh:27-85       (parse_potential, compute)
  Ensemble_NVE   src/integrate/ensemble_nve.cuh  (compute1, compute2)
The C++ adapters a GPUMD maintainer would compile into the reference live in gpumd_b200/host/.

Device arrays are torch CUDA tensors (torch = allocator + stream provider only); all arithmetic
happens in libb200md.so.  Without the library or without a GPU these classes raise.
"""
import ctypes as C

import numpy as np
import torch

from . import lib as _lib

K_B = 8.617343e-5                    # src/utilities/common.cuh:21
TIME_UNIT_CONVERSION = 1.018051e+1   # src/utilities/common.cuh:26


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.B200mdError("no CUDA device: gpumd_b200 has no CPU fallback")


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Box:
    """cpu_h[0..8] (row-major, lattice vectors as columns) + pbc flags."""

    def __init__(self, h, pbc=(1, 1, 1)):
        self.cpu_h = np.ascontiguousarray(np.asarray(h, dtype=np.float64).reshape(9))
        self.pbc = np.ascontiguousarray(np.asarray(pbc, dtype=np.int32).reshape(3))

    def get_volume(self):
        return abs(float(np.linalg.det(self.cpu_h.reshape(3, 3))))

    @property
    def _h(self):
        return self.cpu_h.ctypes.data_as(C.POINTER(C.c_double))

    @property
    def _p(self):
        return self.pbc.ctypes.data_as(C.POINTER(C.c_int))


class Atom:
    """Per-atom device state, same arrays and layouts as GPUMD's Atom."""

    def __init__(self, type_, position, mass, velocity=None, device="cuda"):
        _require_cuda()
        n = int(np.asarray(type_).shape[0])
        self.number_of_atoms = n
        dev = torch.device(device)
        self.type = torch.as_tensor(np.ascontiguousarray(type_, dtype=np.int32), device=dev)
        self.position_per_atom = torch.as_tensor(
            np.ascontiguousarray(position, dtype=np.float64).reshape(3 * n), device=dev)
        self.mass = torch.as_tensor(np.ascontiguousarray(mass, dtype=np.float64), device=dev)
        v = np.zeros(3 * n) if velocity is None else np.ascontiguousarray(
            velocity, dtype=np.float64).reshape(3 * n)
        self.velocity_per_atom = torch.as_tensor(v, device=dev)
        self.force_per_atom = torch.zeros(3 * n, dtype=torch.float64, device=dev)
        self.virial_per_atom = torch.zeros(9 * n, dtype=torch.float64, device=dev)
        self.potential_per_atom = torch.zeros(n, dtype=torch.float64, device=dev)


class Potential:
    """Abstract base, src/force/potential.cuh:20-113."""

    N1 = 0
    N2 = 0
    rc = 0.0

    def compute(self, box, type_, position, potential, force, virial):
        raise NotImplementedError


class NEP(Potential):
    """NEP(file_potential, num_atoms), src/force/nep.cu:100-395; compute = nep.cu:1356-1389."""

    def __init__(self, file_potential, num_atoms):
        _require_cuda()
        self._L = _lib.load()
        h = C.c_void_p()
        _lib.check(self._L.b200md_nep_create(str(file_potential).encode(), int(num_atoms), C.byref(h)))
        self._h = h
        self.N1, self.N2 = 0, int(num_atoms)
        self.rc = self._L.b200md_nep_rc(h)
        self.num_types = self._L.b200md_nep_info(h, 0)
        self.dim = self._L.b200md_nep_info(h, 1)
        self.MN_radial = self._L.b200md_nep_info(h, 3)
        self.MN_angular = self._L.b200md_nep_info(h, 4)
        self.symbols = [self._L.b200md_nep_symbol(h, t).decode() for t in range(self.num_types)]

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.b200md_nep_destroy(self._h)
            self._h = None

    def compute(self, box, type_, position, potential, force, virial):
        n = type_.shape[0]
        _lib.check(self._L.b200md_nep_compute(
            self._h, n, box._h, box._p, _ptr(type_), _ptr(position), _ptr(potential), _ptr(force),
            _ptr(virial), _stream()))

    def compute_host(self, box, type_, position, potential, force, virial):
        """Host (numpy or pinned torch CPU) buffers in and out; outputs are overwritten."""
        n = int(type_.shape[0])

        def p(a):
            return C.c_void_p(a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data)

        _lib.check(self._L.b200md_nep_compute_host(
            self._h, n, box._h, box._p, p(type_), p(position), p(potential), p(force), p(virial)))

    def check(self):
        _lib.check(self._L.b200md_nep_check(self._h, _stream()))

    def invalidate(self, n_new):
        _lib.check(self._L.b200md_nep_invalidate(self._h, int(n_new), _stream()))

    def set_accumulate(self, accumulate):
        """False: outputs are overwritten instead of added to (the caller may skip its zeroing pass)."""
        _lib.check(self._L.b200md_nep_set_accumulate(self._h, 1 if accumulate else 0))

    def set_owned(self, n_owned):
        """Spatial domains: only caller indices < n_owned get outputs (0 = all)."""
        _lib.check(self._L.b200md_nep_set_owned(self._h, int(n_owned)))

    @property
    def num_rebuilds(self):
        return self._L.b200md_nep_info(self._h, 6)

    def profile(self, enable=True):
        _lib.check(self._L.b200md_nep_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        """{stage name: (total ms, samples)} recorded since profile(True)."""
        ms = (C.c_float * 16)()
        cnt = (C.c_int * 16)()
        m = self._L.b200md_nep_profile_read(self._h, 16, ms, cnt)
        return {self._L.b200md_nep_stage_name(k).decode(): (float(ms[k]), int(cnt[k]))
                for k in range(m) if cnt[k] > 0}

    def mean_neighbors(self):
        out = (C.c_double * 3)()
        _lib.check(self._L.b200md_nep_mean_neighbors(self._h, out))
        return dict(skin=out[0], radial=out[1], angular=out[2])

    def export_neighbors(self, mn_r=None, mn_a=None):
        """(NN_radial, NL_radial[n,mn_r], NN_angular, NL_angular[n,mn_a]) of the last compute, in the
        caller's atom indices, ascending; -1 padded."""
        n = self.N2
        mn_r = mn_r or self.MN_radial
        mn_a = mn_a or self.MN_angular
        dev = torch.device("cuda")
        NNr = torch.zeros(n, dtype=torch.int32, device=dev)
        NLr = torch.full((n, mn_r), -1, dtype=torch.int32, device=dev)
        NNa = torch.zeros(n, dtype=torch.int32, device=dev)
        NLa = torch.full((n, mn_a), -1, dtype=torch.int32, device=dev)
        _lib.check(self._L.b200md_nep_export_neighbors(
            self._h, mn_r, _ptr(NNr), _ptr(NLr), mn_a, _ptr(NNa), _ptr(NLa), _stream()))
        self.check()
        return NNr.cpu().numpy(), NLr.cpu().numpy(), NNa.cpu().numpy(), NLa.cpu().numpy()

    def export_descriptors(self):
        q = torch.zeros(self.dim, self.N2, dtype=torch.float32, device="cuda")
        _lib.check(self._L.b200md_nep_export_descriptors(self._h, _ptr(q), _stream()))
        return q.cpu().numpy()


class LJ(Potential):
    """LJ potential; the file is the whole potential file ("lj Nt sym..." + Nt^2 lines),
    src/force/lj.cu:28-59; compute = lj.cu:184-219."""

    def __init__(self, file_potential, num_atoms):
        _require_cuda()
        self._L = _lib.load()
        h = C.c_void_p()
        _lib.check(self._L.b200md_lj_create(str(file_potential).encode(), int(num_atoms), C.byref(h)))
        self._h = h
        self.N1, self.N2 = 0, int(num_atoms)
        self.rc = self._L.b200md_lj_rc(h)
        self.num_types = self._L.b200md_lj_info(h, 0)
        self.symbols = [self._L.b200md_lj_symbol(h, t).decode() for t in range(self.num_types)]

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.b200md_lj_destroy(self._h)
            self._h = None

    def compute(self, box, type_, position, potential, force, virial):
        n = type_.shape[0]
        _lib.check(self._L.b200md_lj_compute(
            self._h, n, box._h, box._p, _ptr(type_), _ptr(position), _ptr(potential), _ptr(force),
            _ptr(virial), _stream()))

    def check(self):
        _lib.check(self._L.b200md_lj_check(self._h, _stream()))

    def invalidate(self, n_new):
        _lib.check(self._L.b200md_lj_invalidate(self._h, int(n_new), _stream()))

    @property
    def num_rebuilds(self):
        return self._L.b200md_lj_info(self._h, 6)


class Tersoff1989(Potential):
    """Tersoff-1989, FP64 (src/force/tersoff1989.cu:31-586); file = whole potential file."""

    def __init__(self, file_potential, num_atoms):
        _require_cuda()
        self._L = _lib.load()
        h = C.c_void_p()
        _lib.check(self._L.b200md_tersoff_create(str(file_potential).encode(), int(num_atoms), C.byref(h)))
        self._h = h
        self.N1, self.N2 = 0, int(num_atoms)
        self.rc = self._L.b200md_tersoff_rc(h)
        self.num_types = self._L.b200md_tersoff_info(h, 0)
        self.symbols = [self._L.b200md_tersoff_symbol(h, t).decode() for t in range(self.num_types)]

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.b200md_tersoff_destroy(self._h)
            self._h = None

    def compute(self, box, type_, position, potential, force, virial):
        n = type_.shape[0]
        _lib.check(self._L.b200md_tersoff_compute(
            self._h, n, box._h, box._p, _ptr(type_), _ptr(position), _ptr(potential), _ptr(force),
            _ptr(virial), _stream()))

    def check(self):
        _lib.check(self._L.b200md_tersoff_check(self._h, _stream()))

    def invalidate(self, n_new):
        _lib.check(self._L.b200md_tersoff_invalidate(self._h, int(n_new), _stream()))

    @property
    def num_rebuilds(self):
        return self._L.b200md_tersoff_info(self._h, 6)


class EAM(Potential):
    """Analytic EAM (eam_zhou_2004 / eam_dai_2006), src/force/eam.cu:28-580."""

    def __init__(self, file_potential, num_atoms):
        _require_cuda()
        self._L = _lib.load()
        h = C.c_void_p()
        _lib.check(self._L.b200md_eam_create(str(file_potential).encode(), int(num_atoms), C.byref(h)))
        self._h = h
        self.N1, self.N2 = 0, int(num_atoms)
        self.rc = self._L.b200md_eam_rc(h)
        self.num_types = self._L.b200md_eam_info(h, 0)
        self.symbols = [self._L.b200md_eam_symbol(h, t).decode() for t in range(self.num_types)]

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.b200md_eam_destroy(self._h)
            self._h = None

    def compute(self, box, type_, position, potential, force, virial):
        n = type_.shape[0]
        _lib.check(self._L.b200md_eam_compute(
            self._h, n, box._h, box._p, _ptr(type_), _ptr(position), _ptr(potential), _ptr(force),
            _ptr(virial), _stream()))

    def check(self):
        _lib.check(self._L.b200md_eam_check(self._h, _stream()))

    def invalidate(self, n_new):
        _lib.check(self._L.b200md_eam_invalidate(self._h, int(n_new), _stream()))

    @property
    def num_rebuilds(self):
        return self._L.b200md_eam_info(self._h, 6)


def compute_heat(atom, heat=None):
    """Per-atom heat current (compute_heat, src/measure/compute_heat.cu:66-90): heat[5N]."""
    n = atom.number_of_atoms
    if heat is None:
        heat = torch.zeros(5 * n, dtype=torch.float64, device=atom.velocity_per_atom.device)
    _lib.check(_lib.load().b200md_compute_heat(
        n, n, _ptr(atom.virial_per_atom), _ptr(atom.velocity_per_atom), _ptr(heat), n, _stream()))
    return heat


class HAC:
    """compute_hac, src/measure/hac.cu:32-280: sample the total heat current every sample_interval steps,
    then the autocorrelation hac[5, Nc] and the running thermal conductivity rtc[5, Nc] (W/mK)."""

    def __init__(self, number_of_steps, sample_interval, Nc):
        _require_cuda()
        self._L = _lib.load()
        self.Nc, self.Nd = int(Nc), int(number_of_steps) // int(sample_interval)
        self.sample_interval = int(sample_interval)
        h = C.c_void_p()
        _lib.check(self._L.b200md_hac_create(int(number_of_steps), int(sample_interval), int(Nc), C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.b200md_hac_destroy(self._h)
            self._h = None

    def process(self, step, atom):
        n = atom.number_of_atoms
        _lib.check(self._L.b200md_hac_sample(self._h, int(step), n, n, _ptr(atom.virial_per_atom),
                                             _ptr(atom.velocity_per_atom), _stream()))

    def series(self):
        out = np.zeros(5 * self.Nd)
        _lib.check(self._L.b200md_hac_series(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), _stream()))
        return out.reshape(5, self.Nd)

    def postprocess(self, time_step, temperature, volume):
        hac, rtc = np.zeros(5 * self.Nc), np.zeros(5 * self.Nc)
        dp = C.POINTER(C.c_double)
        _lib.check(self._L.b200md_hac_finish(self._h, float(time_step), float(temperature), float(volume),
                                             hac.ctypes.data_as(dp), rtc.ctypes.data_as(dp), _stream()))
        return hac.reshape(5, self.Nc), rtc.reshape(5, self.Nc)


class Force:
    """Force driver, src/force/force.cu: parse_potential (75-218) and compute (771-985)."""

    def __init__(self):
        self.potentials = []
        self._L = _lib.load()

    def parse_potential(self, file_potential, num_atoms):
        with open(file_potential) as f:
            first = f.readline().split()[0]
        if first.startswith("nep"):
            pot = NEP(file_potential, num_atoms)
        elif first == "lj":
            pot = LJ(file_potential, num_atoms)
        elif first == "tersoff_1989":
            pot = Tersoff1989(file_potential, num_atoms)
        elif first in ("eam_zhou_2004", "eam_dai_2006"):
            pot = EAM(file_potential, num_atoms)
        else:
            raise _lib.B200mdError(f"illegal potential model '{first}' for gpumd_b200")
        self.potentials = [pot]
        # a single NEP potential: let its final kernel store the outputs and drop the zeroing pass
        # (Force::compute zeroes because its potentials accumulate, force.cu:794-801)
        self._zero = True
        if isinstance(pot, NEP):
            pot.set_accumulate(False)
            self._zero = False
        return pot

    def compute(self, box, position, type_, potential, force, virial):
        n = type_.shape[0]
        st = _stream()
        _lib.check(self._L.b200md_apply_pbc(n, box._h, box._p, _ptr(position), st))
        if self._zero:
            _lib.check(self._L.b200md_zero_properties(n, _ptr(potential), _ptr(force), _ptr(virial), st))
        self.potentials[0].compute(box, type_, position, potential, force, virial)


class Ensemble_NVE:
    """NVE integrator, src/integrate/ensemble_nve.cu:31-95 (velocity-Verlet + thermo)."""

    def __init__(self, num_atoms):
        _require_cuda()
        self._L = _lib.load()
        nbytes = self._L.b200md_thermo_scratch_bytes(int(num_atoms))
        self._scratch = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")

    # `fix` / `move` keywords (integrate.cu:1362-1470): label = Group::label of the grouping method
    fixed_group = -1
    move_group = -1

    def set_groups(self, label, fixed_group=-1, move_group=-1, move_velocity=(0.0, 0.0, 0.0)):
        """label: int32 group label per atom (host array or device tensor).  Atoms of fixed_group stay
        put, atoms of move_group translate with move_velocity (natural units); neither enters the
        temperature (ensemble.cu:111-174, 645-651)."""
        lab = torch.as_tensor(np.ascontiguousarray(label, dtype=np.int32)) if not torch.is_tensor(label) else label
        self._label = lab.to(device="cuda", dtype=torch.int32).contiguous()
        self.fixed_group, self.move_group = int(fixed_group), int(move_group)
        self._move_velocity = (C.c_double * 3)(*[float(v) for v in move_velocity])
        excluded = 0
        for g in (self.fixed_group, self.move_group):
            if g >= 0:
                excluded += int((self._label == g).sum().item())
        self._n_excluded = excluded

    def _vv(self, step1, time_step, atom):
        n = atom.number_of_atoms
        if self.fixed_group < 0 and self.move_group < 0:
            _lib.check(self._L.b200md_velocity_verlet(
                step1, n, float(time_step), _ptr(atom.mass), _ptr(atom.position_per_atom),
                _ptr(atom.velocity_per_atom), _ptr(atom.force_per_atom), _stream()))
        else:
            _lib.check(self._L.b200md_velocity_verlet_groups(
                step1, n, n, float(time_step), _ptr(atom.mass), _ptr(atom.position_per_atom),
                _ptr(atom.velocity_per_atom), _ptr(atom.force_per_atom), _ptr(self._label),
                self.fixed_group, self.move_group, self._move_velocity, _stream()))

    def compute1(self, time_step, box, atom, thermo=None):
        self._vv(1, time_step, atom)

    def compute2(self, time_step, box, atom, thermo):
        self._vv(0, time_step, atom)
        self.find_thermo(box.get_volume(), atom, thermo)

    def find_thermo(self, volume, atom, thermo):
        n = atom.number_of_atoms
        n_t = n - getattr(self, "_n_excluded", 0)  # ensemble.cu:645-651
        _lib.check(self._L.b200md_find_thermo(
            n, n_t, float(volume), _ptr(atom.mass), _ptr(atom.potential_per_atom),
            _ptr(atom.velocity_per_atom), _ptr(atom.virial_per_atom), _ptr(thermo),
            _ptr(self._scratch), _stream()))


class Ensemble_BER(Ensemble_NVE):
    """nvt_ber, src/integrate/ensemble_ber.cu:178-233: velocity-Verlet, thermo, Berendsen scaling
    with the factor computed on the device from thermo[0]."""

    def __init__(self, num_atoms, temperature, temperature_coupling):
        super().__init__(num_atoms)
        self.temperature = float(temperature)
        self.temperature_coupling = float(temperature_coupling)

    def compute2(self, time_step, box, atom, thermo):
        super().compute2(time_step, box, atom, thermo)
        n = atom.number_of_atoms
        _lib.check(self._L.b200md_berendsen_temperature(
            n, n, self.temperature, self.temperature_coupling, _ptr(thermo),
            _ptr(atom.velocity_per_atom), _stream()))


K_B = 8.617343e-5  # common.cuh:21
PRESSURE_UNIT_CONVERSION = 1.602177e+2  # common.cuh:25


class Ensemble_NPT_BER(Ensemble_BER):
    """npt_ber, src/integrate/ensemble_ber.cu:38-64,88-172,237-285: Berendsen thermostat + barostat.
    target_pressure / elastic_modulus in GPa (1, 3 or 6 values, Voigt order for 6), tau_p in steps, as
    on the `ensemble npt_ber` line; converted like integrate.cu:690-700,1150-1153.  The box is updated
    in place on the host (Box.h), positions are rescaled on the device."""

    def __init__(self, num_atoms, temperature, temperature_coupling, target_pressure, elastic_modulus, tau_p):
        super().__init__(num_atoms, temperature, temperature_coupling)
        p = np.atleast_1d(np.asarray(target_pressure, dtype=np.float64))
        c = np.atleast_1d(np.asarray(elastic_modulus, dtype=np.float64))
        if p.shape[0] not in (1, 3, 6) or c.shape != p.shape:
            raise ValueError("npt_ber takes 1, 3 or 6 pressure components and as many elastic moduli")
        self.num_p = int(p.shape[0])
        pc = np.ones(6)
        pc[:self.num_p] = c
        coupling = np.where(pc > 2.0e3, 0.0, 1.0 / (float(tau_p) * 3.0 * pc))
        p6 = np.zeros(6)
        p6[:self.num_p] = p
        self._p0 = (C.c_double * 6)(*(p6 / PRESSURE_UNIT_CONVERSION))
        self._pc = (C.c_double * 6)(*(coupling * PRESSURE_UNIT_CONVERSION))
        self._zero3i = (C.c_int * 3)(0, 0, 0)
        self._zero3d = (C.c_double * 3)(0.0, 0.0, 0.0)

    def compute2(self, time_step, box, atom, thermo):
        super().compute2(time_step, box, atom, thermo)
        n = atom.number_of_atoms
        # box.cpu_h is updated in place (the next force call sees the new box)
        _lib.check(self._L.b200md_berendsen_pressure(
            n, n, self.num_p, self._p0, self._pc, self._zero3i, self._zero3d, box._p, box._h, _ptr(thermo),
            _ptr(atom.position_per_atom), _stream()))


class Ensemble_LAN(Ensemble_NVE):
    """nvt_lan, src/integrate/ensemble_lan.cu:29-41,92-124,190-269: half Langevin kick (c1 =
    exp(-0.5/Tc)) + momentum correction, velocity-Verlet, half kick.  cuRAND XORWOW per atom, seeded
    like the reference (seed 1804289383 = glibc's first rand(), what its -DDEBUG build passes)."""
    C1_EXPONENT = -0.5

    def __init__(self, num_atoms, temperature, temperature_coupling, seed=1804289383):
        super().__init__(num_atoms)
        self.temperature = float(temperature)
        self.c1 = float(np.exp(self.C1_EXPONENT / float(temperature_coupling)))
        h = C.c_void_p()
        _lib.check(self._L.b200md_langevin_create(int(num_atoms), int(seed), C.byref(h)))
        self._lan = h

    def __del__(self):
        if getattr(self, "_lan", None):
            self._L.b200md_langevin_destroy(self._lan)
            self._lan = None

    def _kick(self, atom):
        n = atom.number_of_atoms
        c2 = float(np.sqrt((1.0 - self.c1 * self.c1) * K_B * self.temperature))
        _lib.check(self._L.b200md_langevin_apply(self._lan, n, n, self.c1, c2, _ptr(atom.mass),
                                                 _ptr(atom.velocity_per_atom), _stream()))

    def compute1(self, time_step, box, atom, thermo=None):
        self._kick(atom)
        self._vv(1, time_step, atom)

    def compute2(self, time_step, box, atom, thermo):
        self._vv(0, time_step, atom)
        self._kick(atom)
        self.find_thermo(box.get_volume(), atom, thermo)


class Ensemble_BAO(Ensemble_LAN):
    """nvt_bao, src/integrate/ensemble_bao.cu:29-41,91-120,190-470: B A O A | force | B with the full
    Ornstein-Uhlenbeck step O (c1 = exp(-1/Tc))."""
    C1_EXPONENT = -1.0

    def _op(self, which, time_step, atom):
        n = atom.number_of_atoms
        label = _ptr(self._label) if self.fixed_group >= 0 else None
        _lib.check(self._L.b200md_baoab_operator(
            which, n, n, float(time_step), _ptr(atom.mass), _ptr(atom.position_per_atom),
            _ptr(atom.velocity_per_atom), _ptr(atom.force_per_atom), label, self.fixed_group, _stream()))

    def compute1(self, time_step, box, atom, thermo=None):
        self._op(1, time_step, atom)
        self._op(0, time_step, atom)
        self._kick(atom)
        self._op(0, time_step, atom)

    def compute2(self, time_step, box, atom, thermo):
        self._op(1, time_step, atom)
        self.find_thermo(box.get_volume(), atom, thermo)


class Ensemble_BDP(Ensemble_NVE):
    """nvt_bdp, src/integrate/ensemble_bdp.cu:69-101: velocity-Verlet, thermo, then the stochastic
    rescaling of Bussi et al.; generator and factor live on the device (seed 12345678 = the
    reference's -DDEBUG stream)."""

    def __init__(self, num_atoms, temperature, temperature_coupling, seed=12345678):
        super().__init__(num_atoms)
        h = C.c_void_p()
        _lib.check(self._L.b200md_bdp_create(int(num_atoms), float(temperature),
                                             float(temperature_coupling), int(seed), C.byref(h)))
        self._bdp = h

    def __del__(self):
        if getattr(self, "_bdp", None):
            self._L.b200md_bdp_destroy(self._bdp)
            self._bdp = None

    def compute2(self, time_step, box, atom, thermo):
        super().compute2(time_step, box, atom, thermo)
        n = atom.number_of_atoms
        _lib.check(self._L.b200md_bdp_step(
            self._bdp, n, n, _ptr(thermo), _ptr(atom.velocity_per_atom), _stream()))


class Ensemble_NHC(Ensemble_NVE):
    """nvt_nhc, src/integrate/ensemble_nhc.cu:173-237 (chain integrated on the device)."""

    def __init__(self, num_atoms, temperature, temperature_coupling, time_step):
        super().__init__(num_atoms)
        h = C.c_void_p()
        _lib.check(self._L.b200md_nhc_create(int(num_atoms), float(temperature),
                                             float(temperature_coupling), float(time_step), C.byref(h)))
        self._nhc = h

    def __del__(self):
        if getattr(self, "_nhc", None):
            self._L.b200md_nhc_destroy(self._nhc)
            self._nhc = None

    def _thermostat(self, time_step, box, atom, thermo):
        n = atom.number_of_atoms
        self.find_thermo(box.get_volume(), atom, thermo)
        _lib.check(self._L.b200md_nhc_half_step(
            self._nhc, n, n, float(time_step), _ptr(thermo), _ptr(atom.velocity_per_atom), _stream()))

    def compute1(self, time_step, box, atom, thermo):
        self._thermostat(time_step, box, atom, thermo)
        super().compute1(time_step, box, atom, thermo)

    def compute2(self, time_step, box, atom, thermo):
        self._vv(0, time_step, atom)
        self._thermostat(time_step, box, atom, thermo)


def launch_count():
    return _lib.load().b200md_launch_count()
