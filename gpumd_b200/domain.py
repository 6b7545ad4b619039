"""Spatial-domain sharding of the MD hot path: one process per GPU, 1-D slabs, ghost-atom halo
exchange over torch.distributed (NCCL on GPUs; gloo in the CPU tests).

Replaces the reference's single-process NEP_MULTIGPU scheme (src/force/nep_multigpu.cu:1416-1803:
GPU 0 owns all atoms, scatters positions and gathers forces with blocking peer cudaMemcpy every
step, and integrates everything itself).  Here every rank OWNS the atoms of its slab: they are
integrated where they live, only ghost POSITIONS travel (FP64 x,y,z of the atoms within
2*rc + skin of a face), and thermo is an 8-double all-reduce.  As in the reference
(nep_multigpu.cuh:42-53) the halo is two cutoffs wide so that the descriptors of the first ghost
layer are recomputed locally and no force has to be sent back; forces on ghosts are discarded.

Local frame of rank r (slab [r*w, (r+1)*w) along x, w = Lx/world): x_local = x - (r*w - halo).
Owned atoms sit in [halo, halo + w), ghosts in [0, halo) and [halo + w, w + 2*halo).  The local
box is open along x (pbc = 0,1,1).  Arrays are SoA with stride n_loc = n_own + n_ghost; the first
n_own entries are owned.  A ghost sent to the left neighbour appears there at x_local + w, one sent
to the right neighbour at x_local - w (periodic wrap included, since all slabs have width w).

The device work (force call, strided velocity-Verlet / wrap / thermo, halo pack) goes through the
libb200md C-ABI; this module is host-side plumbing.  `backend="torch"` swaps the device calls for
plain torch ops and a user-supplied force function so that the decomposition logic is testable on
CPU with gloo (tests/test_domain_cpu.py).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


class SlabDomain:
    def __init__(self, h, pbc, rc, rank, world, device, skin=1.0, backend="b200md"):
        h = np.asarray(h, dtype=np.float64).reshape(3, 3)
        if np.abs(h - np.diag(np.diag(h))).max() != 0:
            raise ValueError("slab decomposition needs an orthogonal box")
        if not pbc[0]:
            raise ValueError("slab decomposition is along x, which must be periodic")
        self.L = np.diag(h).copy()
        self.pbc = np.asarray(pbc, dtype=np.int32)
        self.rank, self.world = rank, world
        self.device = torch.device(device)
        self.rc, self.skin = float(rc), float(skin)
        self.halo = 2.0 * self.rc + self.skin
        self.w = self.L[0] / world
        if world > 1 and self.w < self.halo:
            raise ValueError("slab narrower than the halo: use fewer ranks")
        self.origin = rank * self.w - self.halo  # global x of local x = 0
        self.left = (rank - 1) % world
        self.right = (rank + 1) % world
        self.backend = backend
        # local box: open along x
        self.local_h = np.diag([self.w + 2 * self.halo, self.L[1], self.L[2]]).reshape(9)
        self.local_pbc = np.array([0, self.pbc[1], self.pbc[2]], np.int32)
        self.n_own = 0
        self.n_loc = 0
        self.volume = float(np.prod(self.L))

    # ------------------------------------------------------------------ setup
    def distribute(self, type_, pos, mass, vel=None):
        """Every rank holds the same global arrays (seeded generator); keep the owned part."""
        x = np.mod(pos[0], self.L[0])
        mine = np.nonzero((x >= self.rank * self.w) & (x < (self.rank + 1) * self.w))[0]
        self.n_global = int(type_.shape[0])
        p = pos[:, mine].copy()
        p[0] = x[mine] - self.origin
        dev = self.device
        self.own_type = torch.as_tensor(np.ascontiguousarray(type_[mine], dtype=np.int32), device=dev)
        self.own_mass = torch.as_tensor(np.ascontiguousarray(mass[mine], dtype=np.float64), device=dev)
        self.own_pos = torch.as_tensor(np.ascontiguousarray(p, dtype=np.float64), device=dev)
        v = np.zeros_like(p) if vel is None else vel[:, mine]
        self.own_vel = torch.as_tensor(np.ascontiguousarray(v, dtype=np.float64), device=dev)
        self.own_force = torch.zeros_like(self.own_vel)
        self.own_id = torch.as_tensor(mine.astype(np.int64), device=dev)
        self.exchange()

    # ------------------------------------------------------------------ migration + ghost lists
    def _sendrecv_var(self, to_left, to_right):
        """Variable-length exchange with both neighbours.  to_left / to_right: lists of tensors with
        the same leading length.  Returns (from_right, from_left) as lists of tensors."""
        dev = self.device
        cnt = torch.tensor([to_left[0].shape[-1], to_right[0].shape[-1]], dtype=torch.int64, device=dev)
        cnt_r = torch.zeros(2, dtype=torch.int64, device=dev)  # [from right nbr, from left nbr]
        ops = [dist.P2POp(dist.isend, cnt[0:1], self.left), dist.P2POp(dist.isend, cnt[1:2], self.right),
               dist.P2POp(dist.irecv, cnt_r[0:1], self.right), dist.P2POp(dist.irecv, cnt_r[1:2], self.left)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        n_from_right, n_from_left = int(cnt_r[0].item()), int(cnt_r[1].item())
        from_right = [torch.empty(t.shape[:-1] + (n_from_right,), dtype=t.dtype, device=dev) for t in to_left]
        from_left = [torch.empty(t.shape[:-1] + (n_from_left,), dtype=t.dtype, device=dev) for t in to_left]
        ops = []
        for t in to_left:
            if t.numel():
                ops.append(dist.P2POp(dist.isend, t.contiguous(), self.left))
        for t in to_right:
            if t.numel():
                ops.append(dist.P2POp(dist.isend, t.contiguous(), self.right))
        for t in from_right:
            if t.numel():
                ops.append(dist.P2POp(dist.irecv, t, self.right))
        for t in from_left:
            if t.numel():
                ops.append(dist.P2POp(dist.irecv, t, self.left))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        return from_right, from_left

    def exchange(self):
        """Migrate atoms that left the slab, rebuild the ghost send lists, size the local arrays.
        Rare (triggered by DomainMD.maybe_exchange when atoms have moved too far); host-orchestrated
        torch ops, tens to hundreds of ms per event."""
        dev = self.device
        if self.world == 1:
            raise RuntimeError("SlabDomain is for world_size > 1")
        x = self.own_pos[0]
        lo, hi = self.halo, self.halo + self.w
        go_left = x < lo
        go_right = x >= hi
        stay = ~(go_left | go_right)

        def pick(mask, shift):
            p = self.own_pos[:, mask].clone()
            p[0] += shift
            return [p, self.own_vel[:, mask], self.own_mass[mask], self.own_type[mask], self.own_id[mask],
                    self.own_force[:, mask]]

        fr, fl = self._sendrecv_var(pick(go_left, +self.w), pick(go_right, -self.w))
        self.own_pos = torch.cat([self.own_pos[:, stay], fr[0], fl[0]], dim=1).contiguous()
        self.own_vel = torch.cat([self.own_vel[:, stay], fr[1], fl[1]], dim=1).contiguous()
        self.own_mass = torch.cat([self.own_mass[stay], fr[2], fl[2]]).contiguous()
        self.own_type = torch.cat([self.own_type[stay], fr[3], fl[3]]).contiguous()
        self.own_id = torch.cat([self.own_id[stay], fr[4], fl[4]]).contiguous()
        self.own_force = torch.cat([self.own_force[:, stay], fr[5], fl[5]], dim=1).contiguous()
        self.n_own = int(self.own_type.shape[0])

        # ghost send lists (indices into the owned arrays)
        x = self.own_pos[0]
        self.idx_left = torch.nonzero(x < lo + self.halo).flatten().to(torch.int32).contiguous()
        self.idx_right = torch.nonzero(x >= hi - self.halo).flatten().to(torch.int32).contiguous()
        # static ghost data (type, mass) travels once per exchange
        il, ir = self.idx_left.long(), self.idx_right.long()
        fr, fl = self._sendrecv_var([self.own_type[il], self.own_mass[il]],
                                    [self.own_type[ir], self.own_mass[ir]])
        self.n_ghost_right = int(fr[0].shape[0])  # ghosts beyond my right face (from the right nbr)
        self.n_ghost_left = int(fl[0].shape[0])
        n_loc = self.n_own + self.n_ghost_right + self.n_ghost_left
        self.n_loc = n_loc
        # local SoA arrays [owned | right ghosts | left ghosts]
        self.type = torch.cat([self.own_type, fr[0], fl[0]]).contiguous()
        self.mass = torch.cat([self.own_mass, fr[1], fl[1]]).contiguous()
        self.pos = torch.zeros(3 * n_loc, dtype=torch.float64, device=dev)
        self.vel = torch.zeros(3 * n_loc, dtype=torch.float64, device=dev)
        self.force = torch.zeros(3 * n_loc, dtype=torch.float64, device=dev)
        self.virial = torch.zeros(9 * n_loc, dtype=torch.float64, device=dev)
        self.pe = torch.zeros(n_loc, dtype=torch.float64, device=dev)
        self.pos.view(3, n_loc)[:, :self.n_own] = self.own_pos
        self.vel.view(3, n_loc)[:, :self.n_own] = self.own_vel
        self.force.view(3, n_loc)[:, :self.n_own] = self.own_force  # forces travel with the atoms
        self.ref_pos = self.own_pos.clone()  # positions at this exchange (displacement trigger)
        self.send_left = torch.zeros(3 * self.idx_left.numel(), dtype=torch.float64, device=dev)
        self.send_right = torch.zeros(3 * self.idx_right.numel(), dtype=torch.float64, device=dev)
        self.recv_right = torch.zeros(3 * self.n_ghost_right, dtype=torch.float64, device=dev)
        self.recv_left = torch.zeros(3 * self.n_ghost_left, dtype=torch.float64, device=dev)
        self.halo_update()

    def sync_owned_views(self):
        """Copy the owned part of the local arrays back into own_* (before an exchange)."""
        n, m = self.n_own, self.n_loc
        self.own_pos = self.pos.view(3, m)[:, :n].clone()
        self.own_vel = self.vel.view(3, m)[:, :n].clone()
        self.own_force = self.force.view(3, m)[:, :n].clone()

    def needs_exchange(self, fraction=0.7):
        """True on every rank when any owned atom anywhere has moved further than
        fraction * skin/2 since the last exchange (one small all-reduce + one host read)."""
        n, m = self.n_own, self.n_loc
        d = self.pos.view(3, m)[:, :n] - self.ref_pos
        if self.pbc[1]:
            d[1] -= self.L[1] * torch.round(d[1] / self.L[1])
        if self.pbc[2]:
            d[2] -= self.L[2] * torch.round(d[2] / self.L[2])
        worst = (d * d).sum(dim=0).max().reshape(1)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        lim = fraction * 0.5 * self.skin
        return bool(worst.item() > lim * lim)

    # ------------------------------------------------------------------ per step
    def _pack(self, idx, out, shift_x):
        m = idx.numel()
        if m == 0:
            return
        if self.backend == "b200md":
            from . import lib as _lib
            L = _lib.load()
            sh = (C.c_double * 3)(shift_x, 0.0, 0.0)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(L.b200md_halo_pack(m, C.c_void_p(idx.data_ptr()), self.n_loc,
                                          C.c_void_p(self.pos.data_ptr()), sh,
                                          C.c_void_p(out.data_ptr()), st))
        else:
            p = self.pos.view(3, self.n_loc)[:, idx.long()].clone()
            p[0] += shift_x
            out.view(3, m).copy_(p)

    def halo_update(self):
        """Ghost positions for the current owned positions (every force evaluation)."""
        self._pack(self.idx_left, self.send_left, +self.w)
        self._pack(self.idx_right, self.send_right, -self.w)
        ops = []
        if self.send_left.numel():
            ops.append(dist.P2POp(dist.isend, self.send_left, self.left))
        if self.send_right.numel():
            ops.append(dist.P2POp(dist.isend, self.send_right, self.right))
        if self.recv_right.numel():
            ops.append(dist.P2POp(dist.irecv, self.recv_right, self.right))
        if self.recv_left.numel():
            ops.append(dist.P2POp(dist.irecv, self.recv_left, self.left))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        n, m = self.n_own, self.n_loc
        p = self.pos.view(3, m)
        p[:, n:n + self.n_ghost_right] = self.recv_right.view(3, self.n_ghost_right)
        p[:, n + self.n_ghost_right:] = self.recv_left.view(3, self.n_ghost_left)

    def allreduce_thermo(self, thermo, async_op=False):
        return dist.all_reduce(thermo, op=dist.ReduceOp.SUM, async_op=async_op)


class DomainMD:
    """NVE driver over a SlabDomain with the libb200md kernels (one instance per rank)."""

    def __init__(self, dom, potential_file, capacity_factor=1.35, ensemble="nve", temperature=300.0,
                 temperature_coupling=100.0, time_step=None):
        """ensemble: "nve", "nvt_ber" (Ensemble_BER, ensemble_ber.cu:178-233) or "nvt_nhc"
        (Ensemble_NHC, ensemble_nhc.cu:173-237); thermostats act on the GLOBAL temperature."""
        from . import engine, lib as _lib
        self.dom = dom
        self.eng = engine
        self.L = _lib.load()
        self._lib = _lib
        cap = int(dom.n_loc * capacity_factor) + 1024
        self.pot = engine.Force().parse_potential(potential_file, cap)
        self.capacity = cap
        self._set_owned()
        self.box = engine.Box(dom.local_h, dom.local_pbc)
        self.thermo = torch.zeros(8, dtype=torch.float64, device=dom.device)
        nbytes = self.L.b200md_thermo_scratch_bytes(cap)
        self._scratch = torch.zeros(nbytes, dtype=torch.uint8, device=dom.device)
        self.steps_since_exchange = 0
        self.ensemble = ensemble
        self.temperature = float(temperature)
        self.temperature_coupling = float(temperature_coupling)
        self._nhc = None
        self._bdp = None
        self.profile = None  # enable_profile(): {phase: [ms_sum, count]} from CUDA events
        self._pending = []
        if ensemble == "nvt_nhc":
            h = C.c_void_p()
            self._lib.check(self.L.b200md_nhc_create(
                int(dom.n_global), self.temperature, self.temperature_coupling, float(time_step), C.byref(h)))
            self._nhc = h
        elif ensemble == "nvt_bdp":
            h = C.c_void_p()
            self._lib.check(self.L.b200md_bdp_create(
                int(dom.n_global), self.temperature, self.temperature_coupling, 12345678, C.byref(h)))
            self._bdp = h  # every rank advances an identical generator from the all-reduced T
        elif ensemble not in ("nve", "nvt_ber"):
            raise ValueError(f"unsupported ensemble {ensemble}")

    def __del__(self):
        if getattr(self, "_nhc", None):
            self.L.b200md_nhc_destroy(self._nhc)
            self._nhc = None
        if getattr(self, "_bdp", None):
            self.L.b200md_bdp_destroy(self._bdp)
            self._bdp = None

    def _st(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _set_owned(self):
        # forces / virials of ghosts are never used here: let the potential skip them
        if hasattr(self.pot, "set_owned"):
            self.pot.set_owned(self.dom.n_own)

    # ---- optional per-phase device timing (CUDA events on the current stream) ----
    def enable_profile(self, on=True):
        self.profile = {} if on else None
        self._pending = []

    def _mark(self, name):
        """Close the phase opened by the previous _mark (if any) and open `name`."""
        if self.profile is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._pending.append((name, ev))

    def profile_read(self):
        """Mean milliseconds per occurrence of every phase since enable_profile()."""
        if self.profile is None:
            return None
        torch.cuda.synchronize()
        for (name, e0), (_, e1) in zip(self._pending[:-1], self._pending[1:]):
            if name is None:
                continue
            acc = self.profile.setdefault(name, [0.0, 0])
            acc[0] += e0.elapsed_time(e1)
            acc[1] += 1
        self._pending = []
        return {k: v[0] / max(v[1], 1) for k, v in self.profile.items()}

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def compute_force(self):
        d = self.dom
        if d.n_loc > self.capacity:
            raise RuntimeError("local atom count exceeds the capacity of the potential instance")
        self._lib.check(self.L.b200md_zero_properties(
            d.n_loc, self._p(d.pe), self._p(d.force), self._p(d.virial), self._st()))
        self.pot.compute(self.box, d.type, d.pos, d.pe, d.force, d.virial)

    def _nhc_half(self, dt):
        d = self.dom
        self.find_thermo(True, wait=True)
        self._lib.check(self.L.b200md_nhc_half_step(
            self._nhc, d.n_own, d.n_loc, float(dt), self._p(self.thermo), self._p(d.vel), self._st()))

    def heat_current(self):
        """Global heat current (5 components: jx_in, jx_out, jy_in, jy_out, jz) from the per-atom
        virial and velocity of the owned atoms (compute_heat.cu:32-90 + the 5-way sum of
        hac.cu:51,106), all-reduced over ranks."""
        d = self.dom
        if getattr(self, "_heat", None) is None or self._heat.numel() != 5 * d.n_own:
            self._heat = torch.zeros(5 * d.n_own, dtype=torch.float64, device=d.device)
        self._lib.check(self.L.b200md_compute_heat(
            d.n_own, d.n_loc, self._p(d.virial), self._p(d.vel), self._p(self._heat), d.n_own, self._st()))
        j = self._heat.view(5, d.n_own).sum(dim=1)
        dist.all_reduce(j, op=dist.ReduceOp.SUM)
        return j

    def step(self, dt, reduce_thermo=True):
        d = self.dom
        L, st = self.L, self._st()
        self._mark("vv1+wrap")
        if self._nhc is not None:
            self._nhc_half(dt)
        self._lib.check(L.b200md_velocity_verlet_strided(
            1, d.n_own, d.n_loc, float(dt), self._p(d.mass), self._p(d.pos), self._p(d.vel),
            self._p(d.force), st))
        self._lib.check(L.b200md_apply_pbc_strided(
            d.n_own, d.n_loc, self.box._h, self.box._p, self._p(d.pos), st))
        self._mark("halo")
        d.halo_update()
        self._mark("force")
        self.compute_force()
        self._mark("vv2+thermo")
        self._lib.check(L.b200md_velocity_verlet_strided(
            0, d.n_own, d.n_loc, float(dt), self._p(d.mass), self._p(d.pos), self._p(d.vel),
            self._p(d.force), st))
        if self._nhc is not None:
            self._nhc_half(dt)
        else:
            scaled = self.ensemble in ("nvt_ber", "nvt_bdp")
            self.find_thermo(reduce_thermo or scaled, wait=scaled)
            if self.ensemble == "nvt_bdp":
                self._lib.check(L.b200md_bdp_step(
                    self._bdp, d.n_own, d.n_loc, self._p(self.thermo), self._p(d.vel), st))
            if self.ensemble == "nvt_ber":
                self._lib.check(L.b200md_berendsen_temperature(
                    d.n_own, d.n_loc, self.temperature, self.temperature_coupling, self._p(self.thermo),
                    self._p(d.vel), st))
        self.steps_since_exchange += 1
        self._mark(None)

    def find_thermo(self, reduce=True, wait=False):
        d = self.dom
        self.thermo_wait()  # the previous all-reduce must be done before thermo is overwritten
        self._lib.check(self.L.b200md_find_thermo_strided(
            d.n_own, d.n_loc, d.n_global, d.volume, self._p(d.mass), self._p(d.pe), self._p(d.vel),
            self._p(d.virial), self._p(self.thermo), self._p(self._scratch), self._st()))
        if reduce:
            # asynchronous: the 8-double all-reduce runs on the communication stream while the next
            # step's kernels start; thermo_wait() (or the next find_thermo) joins it
            self._thermo_work = d.allreduce_thermo(self.thermo, async_op=True)
            if wait:
                self.thermo_wait()

    def read_thermo(self):
        """Global thermo[0..7] = T, U, sxx, syy, szz, sxy, sxz, syz as a host array."""
        self.thermo_wait()
        return self.thermo.cpu().numpy().copy()

    def thermo_wait(self):
        """Make the current stream wait for the pending thermo all-reduce (call before reading
        self.thermo)."""
        w = getattr(self, "_thermo_work", None)
        if w is not None:
            w.wait()
            self._thermo_work = None

    def exchange(self):
        """Migration + new ghost lists; invalidates the potential's cell order."""
        d = self.dom
        d.sync_owned_views()
        d.exchange()
        if d.n_loc > self.capacity:
            raise RuntimeError("local atom count exceeds the capacity of the potential instance")
        self.pot.invalidate(d.n_loc)
        self._set_owned()
        self.steps_since_exchange = 0

    def maybe_exchange(self, check_every=5):
        """Displacement-triggered migration: checked every `check_every` steps (the only host
        synchronisation of the multi-GPU loop)."""
        if self.steps_since_exchange and self.steps_since_exchange % check_every == 0:
            if self.dom.needs_exchange():
                self.exchange()
                return True
        return False
