"""Host-side mirror of the reference API for the hot path, over the C ABI of libcubezcuda.

Two layers:

* `BatchedWorld` — the new batched-world handle (many independent worlds, device resident).
* The reference's object API under its own names — `RigidBody`, `CollisionCube`,
  `CollisionSphere`, `CollisionPlane`, `Contact`, `CheckForCollisions`, `ResolveContacts`
  (rigidbody.go, colliders.go, contact.go) — implemented as batch-of-N shims: flatten the
  pointer graph to indices, call the C ABI, mirror the results back into the objects.  This is
  what the cgo package in go/cubez does; the Python spelling exists so the parity tests read
  like calls against the reference.

Everything computes on the GPU through libcubezcuda; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _abi
from ._abi import (Bodies, Colliders, Contacts, CubezError, CzStepStats, CzWorldDesc, Planes, SCHED_ALL_PAIRS_ORDERED,
                   SCHED_EXPLICIT, SHAPE_CUBE, SHAPE_NONE, SHAPE_SPHERE)


class Context:
    """cz_init / cz_shutdown for one device and one precision."""

    _cache = {}

    def __init__(self, device: int = 0, prec: str = "f64"):
        self.prec = _abi.precision(prec)
        self.lib = _abi.load(prec)
        self.h = C.c_void_p()
        rc = self.lib.cz_init(device, C.byref(self.h))
        if rc != 0:
            raise CubezError(rc, self.lib.cz_last_error(None).decode())
        self.device = device

    @classmethod
    def get(cls, device: int = 0, prec: str = "f64") -> "Context":
        key = (device, prec)
        if key not in cls._cache:
            cls._cache[key] = cls(device, prec)
        return cls._cache[key]

    def check(self, rc: int):
        if rc != 0:
            raise CubezError(rc, self.lib.cz_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.cz_shutdown(self.h)
            self.h = C.c_void_p()

    # --- math self-test entry (tests port the reference's math/*_test.go through this) ----
    def math_op(self, op: str, *values) -> np.ndarray:
        code = _abi.OPS[op]
        flat = np.zeros(24, dtype=self.prec.dtype)
        vals = np.concatenate([np.atleast_1d(np.asarray(v, dtype=self.prec.dtype)).ravel() for v in values])
        flat[: vals.size] = vals
        out = np.zeros(12, dtype=self.prec.dtype)
        PR = C.POINTER(self.prec.ctype)
        self.check(self.lib.cz_math_op(self.h, code, flat.ctypes.data_as(PR), out.ctypes.data_as(PR)))
        return out[: _abi.OP_OUT[code]].copy()

    # --- object-API shims on flat arrays ------------------------------------------------
    def integrate(self, bodies: Bodies, dt, lin_pow=None, ang_pow=None, bias=None):
        PR = C.POINTER(self.prec.ctype)
        st = bodies.struct()
        lp = None if lin_pow is None else np.ascontiguousarray(lin_pow, dtype=self.prec.dtype)
        ap = None if ang_pow is None else np.ascontiguousarray(ang_pow, dtype=self.prec.dtype)
        bs = None if bias is None else np.asarray([bias], dtype=self.prec.dtype)
        self.check(self.lib.cz_integrate(
            self.h, C.byref(st), self.prec.ctype(dt),
            None if lp is None else lp.ctypes.data_as(PR), None if ap is None else ap.ctypes.data_as(PR),
            None if bs is None else bs.ctypes.data_as(PR)))
        return bodies

    def calculate_derived_data(self, bodies: Bodies):
        st = bodies.struct()
        self.check(self.lib.cz_calculate_derived_data(self.h, C.byref(st)))
        return bodies

    def collider_derive(self, body_transform, offset):
        PR = C.POINTER(self.prec.ctype)
        t = np.ascontiguousarray(body_transform, dtype=self.prec.dtype).reshape(-1, 12)
        o = np.ascontiguousarray(offset, dtype=self.prec.dtype).reshape(-1, 12)
        out = np.zeros_like(t)
        self.check(self.lib.cz_collider_derive(self.h, t.shape[0], t.ctypes.data_as(PR), o.ctypes.data_as(PR), out.ctypes.data_as(PR)))
        return out

    def narrowphase(self, colliders: Colliders, planes: Optional[Planes], bodies: Optional[Bodies], one, two, capacity: int = 0):
        one = np.ascontiguousarray(one, dtype=np.int32)
        two = np.ascontiguousarray(two, dtype=np.int32)
        n = one.shape[0]
        cap = capacity or max(8 * n, 8)
        out = Contacts(cap, self.prec)
        found = np.zeros(max(n, 1), dtype=np.uint8)
        cst, ost = colliders.struct(), out.struct()
        pst = planes.struct() if planes is not None else None
        bst = bodies.struct() if bodies is not None else None
        P32 = C.POINTER(C.c_int32)
        self.check(self.lib.cz_narrowphase(
            self.h, C.byref(cst), None if pst is None else C.byref(pst), None if bst is None else C.byref(bst), n,
            one.ctypes.data_as(P32), two.ctypes.data_as(P32), C.byref(ost), found.ctypes.data_as(C.POINTER(C.c_uint8))))
        out.take(ost)
        return out, found[:n].astype(bool)

    def resolve_contacts(self, max_iterations: int, contacts: Contacts, bodies: Bodies, dt) -> Tuple[int, int]:
        cst, bst = contacts.struct(), bodies.struct()
        iters = (C.c_int32 * 2)()
        self.check(self.lib.cz_resolve_contacts(self.h, max_iterations, C.byref(cst), C.byref(bst), self.prec.ctype(dt), iters))
        contacts.take(cst)
        return int(iters[0]), int(iters[1])

    def pinned_array(self, shape, dtype=None) -> np.ndarray:
        """A page-locked host array (cz_host_alloc), e.g. the action buffers of BatchedWorld.step_rl."""
        dtype = np.dtype(self.prec.dtype if dtype is None else dtype)
        count = int(np.prod(shape))
        raw = C.c_void_p()
        self.check(self.lib.cz_host_alloc(self.h, max(1, count * dtype.itemsize), C.byref(raw)))
        buf = (C.c_uint8 * max(1, count * dtype.itemsize)).from_address(raw.value)
        arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
        self._pinned_keep = getattr(self, "_pinned_keep", []) + [(raw, buf)]
        arr[...] = 0
        return arr

    def pinned_bodies(self, n: int, fields=None) -> Bodies:
        """A Bodies record whose arrays live in page-locked host memory (cz_host_alloc): the
        zero-copy views a cgo caller would get with unsafe.Slice over C-allocated buffers.  Used
        with BatchedWorld.step_host / step_rl so the H2D / D2H copies are truly asynchronous.
        `fields` limits the record to the named arrays (the others stay NULL)."""
        b = Bodies.__new__(Bodies)
        b.n, b.prec, b.a = int(n), self.prec, {name: None for name, _ in _abi.BODY_FIELDS}
        isz = np.dtype(self.prec.dtype).itemsize
        wanted = [(name, comp) for name, comp in _abi.BODY_FIELDS if fields is None or name in fields]
        total = sum(max(1, n * max(comp, 1)) * (1 if comp == 0 else isz) + 64 for _, comp in wanted)
        raw = C.c_void_p()
        self.check(self.lib.cz_host_alloc(self.h, total, C.byref(raw)))
        buf = (C.c_uint8 * total).from_address(raw.value)
        b._pinned = (raw, buf)
        off = 0
        for name, comp in wanted:
            dt = np.uint8 if comp == 0 else self.prec.dtype
            cnt = n * max(comp, 1)
            arr = np.frombuffer(buf, dtype=dt, count=cnt, offset=off)
            b.a[name] = arr.reshape((n,) if comp in (0, 1) else (n, comp))
            off += (cnt * np.dtype(dt).itemsize + 63) // 64 * 64
        return b

    def bench_integrate(self, n: int, seed: int = 5, warmup: int = 3, steps: int = 100, dt: float = 1.0 / 60.0):
        ms = C.c_float()
        cks = C.c_uint64()
        self.check(self.lib.cz_bench_integrate(self.h, n, seed, warmup, steps, self.prec.ctype(dt), C.byref(ms), C.byref(cks)))
        return float(ms.value), int(cks.value)


class BatchedWorld:
    """Batched-world handle: n_worlds independent worlds stepped on one GPU.

    Each world follows the per-frame loop of examples/cubedrop.go:69-75 (Integrate every body,
    refresh collider transforms, generate contacts over the world's ordered check schedule,
    ResolveContacts(8*len(contacts)))."""

    def __init__(self, n_worlds: int, bodies_per_world: int, contacts_per_world: int, schedule: int = SCHED_ALL_PAIRS_ORDERED,
                 flags: int = 0, device: int = 0, prec: str = "f64", ctx: Optional[Context] = None):
        self.ctx = ctx or Context.get(device, prec)
        self.prec = self.ctx.prec
        self.lib = self.ctx.lib
        self.n_worlds, self.B, self.Cc = n_worlds, bodies_per_world, contacts_per_world
        self.desc = CzWorldDesc(n_worlds, bodies_per_world, contacts_per_world, schedule, flags)
        self.h = C.c_void_p()
        self.ctx.check(self.lib.cz_world_create(self.ctx.h, C.byref(self.desc), C.byref(self.h)))

    @classmethod
    def from_scene(cls, scene, device: int = 0, flags: int = 0, contacts_per_world: Optional[int] = None, ctx=None):
        w = cls(scene.n_worlds, scene.bodies_per_world, contacts_per_world or scene.contacts_per_world, scene.schedule,
                flags, device, scene.prec.name, ctx)
        w.upload_planes(scene.planes)
        if scene.schedule == SCHED_EXPLICIT:
            w.upload_schedule(scene.check_one, scene.check_two)
        w.upload_bodies(scene.bodies, derive=True)
        w.upload_colliders(scene.colliders, derive=True)
        if scene.active_from is not None or scene.integrate is not None:
            af = None if scene.active_from is None else np.tile(scene.active_from, scene.n_worlds) if scene.active_from.shape[0] == scene.bodies_per_world else scene.active_from
            ig = None if scene.integrate is None else np.tile(scene.integrate, scene.n_worlds) if scene.integrate.shape[0] == scene.bodies_per_world else scene.integrate
            w.set_activation(af, ig)
        if getattr(scene, "materials", None):
            w.set_materials(**scene.materials)
        return w

    # ---- uploads ---------------------------------------------------------------------
    def upload_bodies(self, bodies: Bodies, first_world: int = 0, derive: bool = False):
        n = bodies.n // self.B
        st = bodies.struct()
        self.ctx.check(self.lib.cz_world_upload_bodies(self.h, first_world, n, C.byref(st), int(derive)))

    def upload_colliders(self, colliders: Colliders, first_world: int = 0, derive: bool = False):
        n = colliders.n // self.B
        st = colliders.struct()
        self.ctx.check(self.lib.cz_world_upload_colliders(self.h, first_world, n, C.byref(st), int(derive)))

    def upload_planes(self, planes: Planes):
        st = planes.struct()
        self.ctx.check(self.lib.cz_world_upload_planes(self.h, C.byref(st)))

    def upload_schedule(self, one, two):
        one = np.ascontiguousarray(one, dtype=np.int32)
        two = np.ascontiguousarray(two, dtype=np.int32)
        P32 = C.POINTER(C.c_int32)
        self.ctx.check(self.lib.cz_world_upload_schedule(self.h, one.shape[0], one.ctypes.data_as(P32), two.ctypes.data_as(P32)))

    def set_activation(self, active_from=None, integrate=None, first_world: int = 0):
        af = None if active_from is None else np.ascontiguousarray(active_from, dtype=np.int32)
        ig = None if integrate is None else np.ascontiguousarray(integrate, dtype=np.uint8)
        n = (af if af is not None else ig).shape[0] // self.B
        self.ctx.check(self.lib.cz_world_set_activation(
            self.h, first_world, n, None if af is None else af.ctypes.data_as(C.POINTER(C.c_int32)),
            None if ig is None else ig.ctypes.data_as(C.POINTER(C.c_uint8))))

    def set_pow(self, dt, lin_pow, ang_pow, bias):
        PR = C.POINTER(self.prec.ctype)
        lp = np.ascontiguousarray(lin_pow, dtype=self.prec.dtype)
        ap = np.ascontiguousarray(ang_pow, dtype=self.prec.dtype)
        self.ctx.check(self.lib.cz_world_set_pow(self.h, self.prec.ctype(dt), lp.ctypes.data_as(PR), ap.ctypes.data_as(PR), self.prec.ctype(bias)))

    def add_forces(self, force=None, torque=None, first_world: int = 0):
        """forceAccum += force, torqueAccum += torque ([n_bodies, 3] arrays or None) — the writer the reference's
        accumulators never had (rigidbody.go:86-92); consumed and cleared by the next Integrate of an awake body."""
        PR = C.POINTER(self.prec.ctype)
        f = None if force is None else np.ascontiguousarray(force, dtype=self.prec.dtype)
        t = None if torque is None else np.ascontiguousarray(torque, dtype=self.prec.dtype)
        n = (f if f is not None else t).size // (3 * self.B)
        self.ctx.check(self.lib.cz_world_add_forces(self.h, first_world, n, None if f is None else f.ctypes.data_as(PR),
                                                    None if t is None else t.ctypes.data_as(PR)))

    def set_step_index(self, s: int):
        self.ctx.check(self.lib.cz_world_set_step_index(self.h, s))

    def set_episodes(self, length: int, phase0=None):
        """RL-style episodes: snapshot now; world k is at frame phase0[k] and resets on wrap."""
        ph = np.zeros(self.n_worlds, dtype=np.int32) if phase0 is None else np.ascontiguousarray(phase0, dtype=np.int32)
        self.ctx.check(self.lib.cz_world_set_episodes(self.h, length, ph.ctypes.data_as(C.POINTER(C.c_int32))))

    def set_materials(self, friction, restitution, body_material=None, plane_material=None, first_world: int = 0):
        """Per-pair surface materials (new API; replaces the hard-wired `c.Friction = 0.9` /
        `c.Restitution = 0.1` test constants, colliders.go:199-202 and five more sites).  friction /
        restitution: M x M tables indexed [material(one)][material(two)]; body_material: ids for the
        worlds from first_world on; plane_material: one id per plane.  friction=None restores the constants."""
        PR, P32 = C.POINTER(self.prec.ctype), C.POINTER(C.c_int32)
        if friction is None:
            self.ctx.check(self.lib.cz_world_set_materials(self.h, 0, None, None, 0, 0, None, None))
            return
        f = np.ascontiguousarray(friction, dtype=self.prec.dtype)
        r = np.ascontiguousarray(restitution, dtype=self.prec.dtype)
        m = f.shape[0]
        assert f.shape == (m, m) and r.shape == (m, m)
        bm = None if body_material is None else np.ascontiguousarray(body_material, dtype=np.int32).reshape(-1)
        pm = None if plane_material is None else np.ascontiguousarray(plane_material, dtype=np.int32).reshape(-1)
        n = 0 if bm is None else bm.shape[0] // self.B
        self.ctx.check(self.lib.cz_world_set_materials(
            self.h, m, f.ctypes.data_as(PR), r.ctypes.data_as(PR), first_world, n,
            None if bm is None else bm.ctypes.data_as(P32), None if pm is None else pm.ctypes.data_as(P32)))

    def export_gl(self, first_world: int = 0, n_worlds: Optional[int] = None, model: bool = False):
        """Renderer-side export (examples/cubedrop.go:35-37, SetGlVector3 / SetGlQuat of
        examples/exampleapp.go:146-159): float32 Location (n,3) and LocalRotation (n,4 as W,V0,V1,V2) of
        every body, converted on the device; model=True adds the body transform as a column-major 4x4."""
        n_worlds = self.n_worlds - first_world if n_worlds is None else n_worlds
        n = n_worlds * self.B
        loc = np.empty((n, 3), dtype=np.float32)
        rot = np.empty((n, 4), dtype=np.float32)
        mdl = np.empty((n, 16), dtype=np.float32) if model else None
        PF = C.POINTER(C.c_float)
        self.ctx.check(self.lib.cz_world_export_gl(self.h, first_world, n_worlds, loc.ctypes.data_as(PF), rot.ctypes.data_as(PF),
                                                   None if mdl is None else mdl.ctypes.data_as(PF), 0))
        return (loc, rot, mdl) if model else (loc, rot)

    # ---- stepping --------------------------------------------------------------------
    def step(self, dt, n_steps: int = 1, stats: bool = True) -> Optional[dict]:
        st = CzStepStats()
        self.ctx.check(self.lib.cz_world_step(self.h, self.prec.ctype(dt), n_steps, C.byref(st) if stats else None))
        return st.as_dict() if stats else None

    def step_host(self, bodies: Bodies, dt, n_steps: int = 1) -> dict:
        st = CzStepStats()
        bst = bodies.struct()
        self.ctx.check(self.lib.cz_world_step_host(self.h, C.byref(bst), self.prec.ctype(dt), n_steps, C.byref(st)))
        return st.as_dict()

    OBS_FIELDS = ("position", "orientation", "velocity", "rotation")

    def step_rl(self, add_velocity=None, add_rotation=None, obs: Optional[Bodies] = None, dt=None, n_steps: int = 1) -> dict:
        """RL-style step of device-resident worlds: batched AddVelocity / AddRotation in
        (rigidbody.go:195-202, arrays [n_bodies, 3] or None), n_steps frames, then the arrays
        present in `obs` filled (e.g. ctx.pinned_bodies(n, fields=BatchedWorld.OBS_FIELDS))."""
        PR = C.POINTER(self.prec.ctype)
        def ptr(a):
            if a is None:
                return None
            assert a.dtype == self.prec.dtype and a.flags["C_CONTIGUOUS"] and a.size == self.n_worlds * self.B * 3
            return a.ctypes.data_as(PR)
        st = CzStepStats()
        ost = obs.struct() if obs is not None else None
        self.ctx.check(self.lib.cz_world_step_rl(self.h, ptr(add_velocity), ptr(add_rotation), None if ost is None else C.byref(ost),
                                                 self.prec.ctype(dt), n_steps, C.byref(st)))
        return st.as_dict()

    def step_rl_async(self, add_velocity=None, add_rotation=None, obs: Optional[Bodies] = None, obs32: Optional[dict] = None, dt=None, n_steps: int = 1) -> int:
        """Pipelined RL step (cz_world_step_rl_async): returns a ticket at once; up to two steps in flight.  obs32: dict
        of pinned float32 arrays {"position": (n,3), "orientation": (n,4), "velocity": (n,3), "rotation": (n,3)}."""
        PR, PF = C.POINTER(self.prec.ctype), C.POINTER(C.c_float)
        def ptr(a):
            if a is None:
                return None
            assert a.dtype == self.prec.dtype and a.flags["C_CONTIGUOUS"] and a.size == self.n_worlds * self.B * 3
            return a.ctypes.data_as(PR)
        ost = obs.struct() if obs is not None else None
        o32 = None
        if obs32 is not None:
            o32 = _abi.CzObs32(self.n_worlds * self.B, *[None if obs32.get(k) is None else obs32[k].ctypes.data_as(PF)
                                                        for k in ("position", "orientation", "velocity", "rotation")])
        ticket = C.c_int32()
        self.ctx.check(self.lib.cz_world_step_rl_async(self.h, ptr(add_velocity), ptr(add_rotation), None if ost is None else C.byref(ost),
                                                       None if o32 is None else C.byref(o32), self.prec.ctype(dt), n_steps, C.byref(ticket)))
        return int(ticket.value)

    def rl_wait(self, ticket: int, stats: bool = True) -> Optional[dict]:
        st = CzStepStats()
        self.ctx.check(self.lib.cz_world_rl_wait(self.h, ticket, C.byref(st) if stats else None))
        return st.as_dict() if stats else None

    def synchronize(self):
        self.ctx.check(self.lib.cz_world_synchronize(self.h))

    # ---- downloads -------------------------------------------------------------------
    def download(self, first_world: int = 0, n_worlds: Optional[int] = None, fields=None, out: Optional[Bodies] = None) -> Bodies:
        n = self.n_worlds - first_world if n_worlds is None else n_worlds
        out = out if out is not None else Bodies(n * self.B, self.prec, fields=fields)
        st = out.struct()
        self.ctx.check(self.lib.cz_world_download_bodies(self.h, first_world, n, C.byref(st)))
        return out

    def download_colliders(self, first_world: int = 0, n_worlds: Optional[int] = None) -> Colliders:
        n = self.n_worlds - first_world if n_worlds is None else n_worlds
        out = Colliders(n * self.B, self.prec)
        st = out.struct()
        self.ctx.check(self.lib.cz_world_download_colliders(self.h, first_world, n, C.byref(st)))
        return out

    def contacts(self, world: int = 0) -> Contacts:
        out = Contacts(self.Cc, self.prec)
        st = out.struct()
        self.ctx.check(self.lib.cz_world_download_contacts(self.h, world, C.byref(st)))
        return out.take(st)

    def contact_pairs(self, world: int = 0) -> List[Tuple[int, int]]:
        c = self.contacts(world)
        return list(zip(c.valid("body0").tolist(), c.valid("body1").tolist()))

    def last_counts(self):
        nc = np.zeros(self.n_worlds, dtype=np.int32)
        pi = np.zeros(self.n_worlds, dtype=np.int32)
        vi = np.zeros(self.n_worlds, dtype=np.int32)
        P32 = C.POINTER(C.c_int32)
        self.ctx.check(self.lib.cz_world_last_step_counts(self.h, nc.ctypes.data_as(P32), pi.ctypes.data_as(P32), vi.ctypes.data_as(P32)))
        return nc, pi, vi

    def count_nonfinite(self) -> int:
        """Bodies with a NaN / infinite component in position, orientation, velocity or rotation (the reference propagates NaN silently)."""
        n = C.c_int64()
        self.ctx.check(self.lib.cz_world_count_nonfinite(self.h, C.byref(n)))
        return int(n.value)

    def island_stats(self) -> Tuple[int, int]:
        """(frames resolved as one CTA per contact island, of which re-run on the single-CTA path because the cap cut the loop)"""
        a, b = C.c_int64(), C.c_int64()
        self.ctx.check(self.lib.cz_world_island_stats(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def checksum_energy(self) -> Tuple[int, float]:
        cks, en = C.c_uint64(), C.c_double()
        self.ctx.check(self.lib.cz_world_checksum_energy(self.h, C.byref(cks), C.byref(en)))
        return int(cks.value), float(en.value)

    def close(self):
        if self.h:
            self.lib.cz_world_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Run:
    """Multi-GPU run of batched worlds inside ONE process (cz_run_*): the worlds of a scene are sharded contiguously
    over `devices`, stepped with no traffic between devices, and cz_run_finish reduces checksum / energy / counters
    with one grouped ncclAllReduce (host reduce when shards share a device)."""

    def __init__(self, scene, devices, contacts_per_world: Optional[int] = None, flags: int = 0):
        from ._abi import CzRunTotals, load
        self._Totals = CzRunTotals
        self.prec = scene.prec
        self.lib = load(scene.prec.name)
        self.scene = scene
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        self.desc = CzWorldDesc(scene.n_worlds, scene.bodies_per_world, contacts_per_world or scene.contacts_per_world, scene.schedule, flags)
        self.h = C.c_void_p()
        self._check(self.lib.cz_run_create(devs.shape[0], devs.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(self.desc), C.byref(self.h)), created=False)
        pst = scene.planes.struct()
        self._check(self.lib.cz_run_upload_planes(self.h, C.byref(pst)))
        bst = scene.bodies.struct()
        self._check(self.lib.cz_run_upload_bodies(self.h, C.byref(bst), 1))
        cst = scene.colliders.struct()
        self._check(self.lib.cz_run_upload_colliders(self.h, C.byref(cst), 1))

    def _check(self, rc, created=True):
        if rc != 0:
            msg = self.lib.cz_run_last_error(self.h if created else None)
            raise CubezError(rc, (msg or b"").decode())

    def shard(self, k: int):
        w, first, n = C.c_void_p(), C.c_int32(), C.c_int32()
        self._check(self.lib.cz_run_shard(self.h, k, C.byref(w), C.byref(first), C.byref(n)))
        return w, int(first.value), int(n.value)

    def set_episodes(self, length: int, phase0=None):
        ph = None if phase0 is None else np.ascontiguousarray(phase0, dtype=np.int32)
        self._check(self.lib.cz_run_set_episodes(self.h, length, None if ph is None else ph.ctypes.data_as(C.POINTER(C.c_int32))))

    def step(self, dt, n_steps: int = 1):
        self._check(self.lib.cz_run_step(self.h, self.prec.ctype(dt), n_steps))

    def finish(self) -> dict:
        t = self._Totals()
        self._check(self.lib.cz_run_finish(self.h, C.byref(t)))
        return t.as_dict()

    def close(self):
        if self.h:
            self.lib.cz_run_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ==========================================================================================
# The reference's object API (rigidbody.go / colliders.go / contact.go), same identifiers.
# ==========================================================================================
_DEFAULT = {"ctx": None}


def use_context(ctx: Context):
    """Select the context the object-API shims run on (default: device 0, f64)."""
    _DEFAULT["ctx"] = ctx


def _ctx() -> Context:
    if _DEFAULT["ctx"] is None:
        _DEFAULT["ctx"] = Context.get(0, "f64")
    return _DEFAULT["ctx"]


class RigidBody:
    """rigidbody.go:23-101.  Public fields keep the reference's names; private derived fields
    are reachable through the same Get* accessors."""

    def __init__(self):   # NewRigidBody, rigidbody.go:104-114
        R = _ctx().prec.dtype
        self.LinearDamping = R(0.95)
        self.AngularDamping = R(0.95)
        self.Position = np.zeros(3, dtype=R)
        self.Orientation = np.array([1, 0, 0, 0], dtype=R)
        self.Velocity = np.zeros(3, dtype=R)
        self.Acceleration = np.array([0.0, -9.78, 0.0], dtype=R)
        self.Rotation = np.zeros(3, dtype=R)
        self.InverseInertiaTensor = np.zeros(9, dtype=R)
        self.IsAwake = True
        self.CanSleep = True
        self._iit_world = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], dtype=R)
        self._inverse_mass = R(0)
        self._mass = R(0)
        self._transform = np.zeros(12, dtype=R)
        self._last_frame_acc = np.zeros(3, dtype=R)
        self._motion = R(0.6)

    def Clone(self) -> "RigidBody":
        import copy
        return copy.deepcopy(self)

    def SetMass(self, mass):
        R = _ctx().prec.dtype
        self._mass = R(mass)
        self._inverse_mass = R(1.0) / R(mass)

    def SetInfiniteMass(self):
        R = _ctx().prec.dtype
        self._mass = R(0)
        self._inverse_mass = R(0)

    def HasFiniteMass(self) -> bool:
        return bool(self._inverse_mass > 0)

    def GetMass(self):
        return np.finfo(_ctx().prec.dtype).max if self._inverse_mass == 0 else self._mass

    def GetInverseMass(self):
        return self._inverse_mass

    def GetTransform(self):
        return self._transform.copy()

    def GetLastFrameAccelleration(self):
        return self._last_frame_acc.copy()

    def GetInverseInertiaTensorWorld(self):
        return self._iit_world.copy()

    def SetInertiaTensor(self, m):
        from .hostmath import m3_invert
        self.InverseInertiaTensor = m3_invert(m, _ctx().prec.dtype)

    def SetAwake(self, awake: bool):   # rigidbody.go:182-192
        R = _ctx().prec.dtype
        if awake:
            self.IsAwake = True
            self._motion = R(0.6)
        else:
            self.IsAwake = False
            self.Velocity[:] = 0
            self.Rotation[:] = 0

    def AddVelocity(self, v):
        self.Velocity += np.asarray(v, dtype=self.Velocity.dtype)

    def AddRotation(self, v):
        self.Rotation += np.asarray(v, dtype=self.Rotation.dtype)

    def ClearAccumulators(self):
        pass   # forceAccum / torqueAccum have no writer in the reference (always zero)

    def Integrate(self, duration):
        integrate_bodies([self], duration)

    def CalculateDerivedData(self):
        b = _gather([self])
        _ctx().calculate_derived_data(b)
        _scatter([self], b)


def _gather(bodies: Sequence[RigidBody]) -> Bodies:
    ctx = _ctx()
    n = len(bodies)
    b = Bodies(n, ctx.prec)
    for i, r in enumerate(bodies):
        b.position[i] = r.Position; b.orientation[i] = r.Orientation; b.velocity[i] = r.Velocity
        b.rotation[i] = r.Rotation; b.acceleration[i] = r.Acceleration
        b.linear_damping[i] = r.LinearDamping; b.angular_damping[i] = r.AngularDamping
        b.inverse_inertia_tensor[i] = r.InverseInertiaTensor; b.inverse_mass[i] = r._inverse_mass
        b.motion[i] = r._motion; b.is_awake[i] = 1 if r.IsAwake else 0; b.can_sleep[i] = 1 if r.CanSleep else 0
        b.transform[i] = r._transform; b.inverse_inertia_tensor_world[i] = r._iit_world
        b.last_frame_acceleration[i] = r._last_frame_acc
    return b


def _scatter(bodies: Sequence[RigidBody], b: Bodies):
    for i, r in enumerate(bodies):
        r.Position = b.position[i].copy(); r.Orientation = b.orientation[i].copy(); r.Velocity = b.velocity[i].copy()
        r.Rotation = b.rotation[i].copy(); r._motion = b.motion[i]; r.IsAwake = bool(b.is_awake[i])
        r._transform = b.transform[i].copy(); r._iit_world = b.inverse_inertia_tensor_world[i].copy()
        r._last_frame_acc = b.last_frame_acceleration[i].copy()


def integrate_bodies(bodies: Sequence[RigidBody], duration):
    """Batch form of RigidBody.Integrate (one upload, one kernel, one download)."""
    b = _gather(bodies)
    _ctx().integrate(b, duration)
    _scatter(bodies, b)


class CollisionPlane:   # colliders.go:29-35, :81-113
    def __init__(self, normal, offset):
        R = _ctx().prec.dtype
        self.Normal = np.asarray(normal, dtype=R)
        self.Offset = R(offset)

    def Clone(self):
        return CollisionPlane(self.Normal.copy(), self.Offset)

    def CalculateDerivedData(self):
        pass

    def GetBody(self):
        return None

    def GetTransform(self):
        return np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=_ctx().prec.dtype)

    def CheckAgainstHalfSpace(self, plane, existing):
        return False, existing

    def CheckAgainstSphere(self, sphere, existing):
        return CheckForCollisions(self, sphere, existing)

    def CheckAgainstCube(self, cube, existing):
        return CheckForCollisions(self, cube, existing)


class _BodyCollider:
    SHAPE = SHAPE_NONE

    def __init__(self, optBody: Optional[RigidBody]):
        R = _ctx().prec.dtype
        self.Body = optBody if optBody is not None else RigidBody()
        self.Offset = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=R)
        self._transform = np.zeros(12, dtype=R)
        self.HalfSize = np.zeros(3, dtype=R)
        self.Radius = R(0)

    def GetBody(self):
        return self.Body

    def GetTransform(self):
        return self._transform.copy()

    def CalculateDerivedData(self):   # colliders.go:173-176 / :302-304
        self._transform = _ctx().collider_derive(self.Body._transform[None, :], self.Offset[None, :])[0]

    def CheckAgainstHalfSpace(self, plane, existing):
        return CheckForCollisions(self, plane, existing)

    def CheckAgainstSphere(self, sphere, existing):
        return CheckForCollisions(self, sphere, existing)

    def CheckAgainstCube(self, cube, existing):
        return CheckForCollisions(self, cube, existing)


class CollisionCube(_BodyCollider):   # colliders.go:39-53, :265-304
    SHAPE = SHAPE_CUBE

    def __init__(self, optBody, halfSize):
        super().__init__(optBody)
        self.HalfSize = np.asarray(halfSize, dtype=_ctx().prec.dtype)

    def Clone(self):
        c = CollisionCube(self.Body.Clone() if self.Body is not None else None, self.HalfSize.copy())
        c.Offset, c._transform = self.Offset.copy(), self._transform.copy()
        return c


class CollisionSphere(_BodyCollider):   # colliders.go:57-71, :136-176
    SHAPE = SHAPE_SPHERE

    def __init__(self, optBody, radius):
        super().__init__(optBody)
        self.Radius = _ctx().prec.dtype(radius)

    def Clone(self):
        c = CollisionSphere(self.Body.Clone() if self.Body is not None else None, self.Radius)
        c.Offset, c._transform = self.Offset.copy(), self._transform.copy()
        return c


class Contact:   # contact.go:17-57
    def __init__(self):
        R = _ctx().prec.dtype
        self.Bodies: List[Optional[RigidBody]] = [None, None]
        self.Friction = R(0)
        self.Restitution = R(0)
        self.ContactPoint = np.zeros(3, dtype=R)
        self.ContactNormal = np.zeros(3, dtype=R)
        self.Penetration = R(0)


def NewContact() -> Contact:
    return Contact()


def check_collision_list(checks: Sequence[Tuple[object, object]], existing: Optional[List[Contact]] = None):
    """Batch form of CheckForCollisions over an ordered list of (one, two) pairs
    (colliders.go:720-747).  Returns (found per check, contacts appended in the reference's
    append order)."""
    ctx = _ctx()
    contacts = list(existing) if existing else []
    colliders: List[_BodyCollider] = []
    planes: List[CollisionPlane] = []
    bodies: List[RigidBody] = []
    cidx, pidx, bidx = {}, {}, {}

    def reg(x):
        if isinstance(x, CollisionPlane):
            if id(x) not in pidx:
                pidx[id(x)] = len(planes)
                planes.append(x)
            return -(pidx[id(x)] + 1)
        if id(x) not in cidx:
            cidx[id(x)] = len(colliders)
            colliders.append(x)
            if id(x.Body) not in bidx:
                bidx[id(x.Body)] = len(bodies)
                bodies.append(x.Body)
        return cidx[id(x)]

    one = [reg(a) for a, _ in checks]
    two = [reg(b) for _, b in checks]
    if not colliders:
        return [False] * len(checks), contacts
    cs = Colliders(len(colliders), ctx.prec)
    for i, c in enumerate(colliders):
        cs.shape[i] = c.SHAPE; cs.body[i] = bidx[id(c.Body)]; cs.offset[i] = c.Offset; cs.transform[i] = c._transform
        cs.half_size[i] = c.HalfSize; cs.radius[i] = c.Radius
    ps = Planes([p.Normal for p in planes], [p.Offset for p in planes], ctx.prec) if planes else None
    out, found = ctx.narrowphase(cs, ps, _gather(bodies), one, two)
    for k in range(out.count):
        c = Contact()
        b0, b1 = int(out.body0[k]), int(out.body1[k])
        c.Bodies = [bodies[b0] if b0 >= 0 else None, bodies[b1] if b1 >= 0 else None]
        c.Friction, c.Restitution = out.friction[k], out.restitution[k]
        c.ContactPoint, c.ContactNormal = out.point[k].copy(), out.normal[k].copy()
        c.Penetration = out.penetration[k]
        contacts.append(c)
    return found.tolist(), contacts


def CheckForCollisions(one, two, existingContacts: Optional[List[Contact]]):
    """colliders.go:720-747 — returns (found, contacts)."""
    found, contacts = check_collision_list([(one, two)], existingContacts)
    return found[0], contacts


def ResolveContacts(maxIterations: int, contacts: Optional[List[Contact]], duration):
    """contact.go:208-222 — mutates the bodies and contacts in place."""
    ctx = _ctx()
    if not (duration > 0) or not contacts:
        return
    bodies: List[RigidBody] = []
    bidx = {}
    for c in contacts:
        for b in c.Bodies:
            if b is not None and id(b) not in bidx:
                bidx[id(b)] = len(bodies)
                bodies.append(b)
    n = len(contacts)
    cs = Contacts(n, ctx.prec)
    cs.count = n
    for i, c in enumerate(contacts):
        cs.body0[i] = bidx[id(c.Bodies[0])] if c.Bodies[0] is not None else -1
        cs.body1[i] = bidx[id(c.Bodies[1])] if c.Bodies[1] is not None else -1
        cs.friction[i], cs.restitution[i] = c.Friction, c.Restitution
        cs.point[i], cs.normal[i], cs.penetration[i] = c.ContactPoint, c.ContactNormal, c.Penetration
    b = _gather(bodies)
    ctx.resolve_contacts(maxIterations, cs, b, duration)
    _scatter(bodies, b)
    for i, c in enumerate(contacts):
        b0, b1 = int(cs.body0[i]), int(cs.body1[i])
        c.Bodies = [bodies[b0] if b0 >= 0 else None, bodies[b1] if b1 >= 0 else None]
        c.ContactNormal = cs.normal[i].copy()
        c.Penetration = cs.penetration[i]
