"""ctypes wrapper of the CPU oracle (oracle/_build/liboracle_*.so).  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  It mirrors the BatchedWorld interface of cubez_b200.api so parity tests drive
both sides with the same arrays."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Tuple

import numpy as np

from cubez_b200 import _abi
from cubez_b200._abi import Bodies, Colliders, Contacts, CzStepStats, CzWorldDesc, Planes, SCHED_EXPLICIT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIBS = {}


def load(prec: str = "f64") -> C.CDLL:
    if prec in _LIBS:
        return _LIBS[prec]
    path = os.path.join(ROOT, "oracle", "_build", f"liboracle_{prec}{os.environ.get('CUBEZ_ORACLE_SUFFIX', '')}.so")   # _asan: the sanitizer build
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-s"], cwd=os.path.join(ROOT, "oracle"))
    lib = C.CDLL(path)
    p = _abi.precision(prec)
    R, PR = p.ctype, C.POINTER(p.ctype)
    PB, PC, PP, PK = C.POINTER(p.Bodies), C.POINTER(p.Colliders), C.POINTER(p.Planes), C.POINTER(p.Contacts)
    P32, PU8, VP = C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.c_void_p
    sig = {
        "czo_real_size": ([], C.c_int),
        "czo_math_op": ([C.c_int32, PR, PR], C.c_int),
        "czo_integrate": ([PB, R, PR, PR, PR], C.c_int),
        "czo_calculate_derived_data": ([PB], C.c_int),
        "czo_collider_derive": ([C.c_int32, PR, PR, PR], C.c_int),
        "czo_narrowphase": ([PC, PP, PB, C.c_int32, P32, P32, PK, PU8], C.c_int),
        "czo_resolve_contacts": ([C.c_int32, PK, PB, R, P32], C.c_int),
        "czo_world_create": ([C.POINTER(CzWorldDesc)], VP),
        "czo_world_destroy": ([VP], C.c_int),
        "czo_world_upload_bodies": ([VP, C.c_int32, C.c_int32, PB, C.c_int32], C.c_int),
        "czo_world_upload_colliders": ([VP, C.c_int32, C.c_int32, PC, C.c_int32], C.c_int),
        "czo_world_upload_planes": ([VP, PP], C.c_int),
        "czo_world_upload_schedule": ([VP, C.c_int32, P32, P32], C.c_int),
        "czo_world_set_activation": ([VP, C.c_int32, C.c_int32, P32, PU8], C.c_int),
        "czo_world_set_step_index": ([VP, C.c_int64], C.c_int),
        "czo_world_add_forces": ([VP, C.c_int32, C.c_int32, PR, PR], C.c_int),
        "czo_world_set_episodes": ([VP, C.c_int32, P32], C.c_int),
        "czo_world_set_materials": ([VP, C.c_int32, PR, PR, C.c_int32, C.c_int32, P32, P32], C.c_int),
        "czo_world_step": ([VP, R, C.c_int32, C.c_int32, C.POINTER(CzStepStats)], C.c_int),
        "czo_world_download_bodies": ([VP, C.c_int32, C.c_int32, PB], C.c_int),
        "czo_world_download_colliders": ([VP, C.c_int32, C.c_int32, PC], C.c_int),
        "czo_world_download_contacts": ([VP, C.c_int32, PK], C.c_int),
        "czo_world_last_step_counts": ([VP, P32, P32, P32], C.c_int),
        "czo_world_checksum_energy": ([VP, C.POINTER(C.c_uint64), C.POINTER(C.c_double)], C.c_int),
        "czo_bench_integrate": ([PB, R, C.c_int32, C.c_int32, C.POINTER(C.c_double)], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, res
    assert lib.czo_real_size() == (8 if prec == "f64" else 4)
    _LIBS[prec] = lib
    return lib


class Oracle:
    """Object-API calls of the oracle on flat arrays (same signatures as api.Context)."""

    def __init__(self, prec: str = "f64"):
        self.prec = _abi.precision(prec)
        self.lib = load(prec)

    def math_op(self, op: str, *values) -> np.ndarray:
        code = _abi.OPS[op]
        flat = np.zeros(24, dtype=self.prec.dtype)
        vals = np.concatenate([np.atleast_1d(np.asarray(v, dtype=self.prec.dtype)).ravel() for v in values])
        flat[: vals.size] = vals
        out = np.zeros(12, dtype=self.prec.dtype)
        PR = C.POINTER(self.prec.ctype)
        assert self.lib.czo_math_op(code, flat.ctypes.data_as(PR), out.ctypes.data_as(PR)) == 0
        return out[: _abi.OP_OUT[code]].copy()

    def integrate(self, bodies: Bodies, dt, lin_pow=None, ang_pow=None, bias=None):
        PR = C.POINTER(self.prec.ctype)
        st = bodies.struct()
        lp = None if lin_pow is None else np.ascontiguousarray(lin_pow, dtype=self.prec.dtype)
        ap = None if ang_pow is None else np.ascontiguousarray(ang_pow, dtype=self.prec.dtype)
        bs = None if bias is None else np.asarray([bias], dtype=self.prec.dtype)
        assert self.lib.czo_integrate(C.byref(st), self.prec.ctype(dt), None if lp is None else lp.ctypes.data_as(PR),
                                      None if ap is None else ap.ctypes.data_as(PR), None if bs is None else bs.ctypes.data_as(PR)) == 0
        return bodies

    def calculate_derived_data(self, bodies: Bodies):
        st = bodies.struct()
        assert self.lib.czo_calculate_derived_data(C.byref(st)) == 0
        return bodies

    def collider_derive(self, body_transform, offset):
        PR = C.POINTER(self.prec.ctype)
        t = np.ascontiguousarray(body_transform, dtype=self.prec.dtype).reshape(-1, 12)
        o = np.ascontiguousarray(offset, dtype=self.prec.dtype).reshape(-1, 12)
        out = np.zeros_like(t)
        assert self.lib.czo_collider_derive(t.shape[0], t.ctypes.data_as(PR), o.ctypes.data_as(PR), out.ctypes.data_as(PR)) == 0
        return out

    def narrowphase(self, colliders: Colliders, planes: Optional[Planes], bodies: Optional[Bodies], one, two, capacity: int = 0):
        one = np.ascontiguousarray(one, dtype=np.int32)
        two = np.ascontiguousarray(two, dtype=np.int32)
        n = one.shape[0]
        out = Contacts(capacity or max(8 * n, 8), self.prec)
        found = np.zeros(max(n, 1), dtype=np.uint8)
        cst, ost = colliders.struct(), out.struct()
        pst = planes.struct() if planes is not None else None
        bst = bodies.struct() if bodies is not None else None
        P32 = C.POINTER(C.c_int32)
        rc = self.lib.czo_narrowphase(C.byref(cst), None if pst is None else C.byref(pst), None if bst is None else C.byref(bst), n,
                                      one.ctypes.data_as(P32), two.ctypes.data_as(P32), C.byref(ost), found.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert rc == 0, rc
        out.take(ost)
        return out, found[:n].astype(bool)

    def resolve_contacts(self, max_iterations: int, contacts: Contacts, bodies: Bodies, dt) -> Tuple[int, int]:
        cst, bst = contacts.struct(), bodies.struct()
        iters = (C.c_int32 * 2)()
        rc = self.lib.czo_resolve_contacts(max_iterations, C.byref(cst), C.byref(bst), self.prec.ctype(dt), iters)
        contacts.take(cst)
        self.last_rc = rc
        return int(iters[0]), int(iters[1])

    def bench_integrate(self, bodies: Bodies, dt, steps: int, n_threads: int = 1) -> float:
        st = bodies.struct()
        sec = C.c_double()
        assert self.lib.czo_bench_integrate(C.byref(st), self.prec.ctype(dt), steps, n_threads, C.byref(sec)) == 0
        return float(sec.value)


class OracleWorld:
    """Same surface as cubez_b200.api.BatchedWorld, computed by the CPU restatement."""

    def __init__(self, n_worlds: int, bodies_per_world: int, contacts_per_world: int, schedule: int = 0, prec: str = "f64"):
        self.prec = _abi.precision(prec)
        self.lib = load(prec)
        self.n_worlds, self.B, self.Cc = n_worlds, bodies_per_world, contacts_per_world
        self.desc = CzWorldDesc(n_worlds, bodies_per_world, contacts_per_world, schedule, 0)
        self.h = self.lib.czo_world_create(C.byref(self.desc))

    @classmethod
    def from_scene(cls, scene):
        w = cls(scene.n_worlds, scene.bodies_per_world, scene.contacts_per_world, scene.schedule, scene.prec.name)
        pst = scene.planes.struct()
        w.lib.czo_world_upload_planes(w.h, C.byref(pst))
        P32 = C.POINTER(C.c_int32)
        if scene.schedule == SCHED_EXPLICIT:
            w.lib.czo_world_upload_schedule(w.h, scene.check_one.shape[0], scene.check_one.ctypes.data_as(P32), scene.check_two.ctypes.data_as(P32))
        w.upload_bodies(scene.bodies, derive=True)
        w.upload_colliders(scene.colliders, derive=True)
        if scene.active_from is not None or scene.integrate is not None:
            af = None if scene.active_from is None else (np.tile(scene.active_from, scene.n_worlds) if scene.active_from.shape[0] == scene.bodies_per_world else scene.active_from)
            ig = None if scene.integrate is None else (np.tile(scene.integrate, scene.n_worlds) if scene.integrate.shape[0] == scene.bodies_per_world else scene.integrate)
            af = None if af is None else np.ascontiguousarray(af, dtype=np.int32)
            ig = None if ig is None else np.ascontiguousarray(ig, dtype=np.uint8)
            w.lib.czo_world_set_activation(w.h, 0, scene.n_worlds, None if af is None else af.ctypes.data_as(P32),
                                           None if ig is None else ig.ctypes.data_as(C.POINTER(C.c_uint8)))
        if getattr(scene, "materials", None):
            w.set_materials(**scene.materials)
        return w

    def upload_bodies(self, bodies: Bodies, first_world: int = 0, derive: bool = False):
        st = bodies.struct()
        self.lib.czo_world_upload_bodies(self.h, first_world, bodies.n // self.B, C.byref(st), int(derive))

    def upload_colliders(self, colliders: Colliders, first_world: int = 0, derive: bool = False):
        st = colliders.struct()
        self.lib.czo_world_upload_colliders(self.h, first_world, colliders.n // self.B, C.byref(st), int(derive))

    def add_forces(self, force=None, torque=None, first_world: int = 0):
        PR = C.POINTER(self.prec.ctype)
        f = None if force is None else np.ascontiguousarray(force, dtype=self.prec.dtype)
        t = None if torque is None else np.ascontiguousarray(torque, dtype=self.prec.dtype)
        n = (f if f is not None else t).size // (3 * self.B)
        self.lib.czo_world_add_forces(self.h, first_world, n, None if f is None else f.ctypes.data_as(PR), None if t is None else t.ctypes.data_as(PR))

    def set_step_index(self, s: int):
        self.lib.czo_world_set_step_index(self.h, s)

    def set_episodes(self, length: int, phase0=None):
        ph = np.zeros(self.n_worlds, dtype=np.int32) if phase0 is None else np.ascontiguousarray(phase0, dtype=np.int32)
        self.lib.czo_world_set_episodes(self.h, length, ph.ctypes.data_as(C.POINTER(C.c_int32)))

    def set_materials(self, friction, restitution, body_material=None, plane_material=None, first_world: int = 0):
        PR, P32 = C.POINTER(self.prec.ctype), C.POINTER(C.c_int32)
        if friction is None:
            self.lib.czo_world_set_materials(self.h, 0, None, None, 0, 0, None, None)
            return
        f = np.ascontiguousarray(friction, dtype=self.prec.dtype)
        r = np.ascontiguousarray(restitution, dtype=self.prec.dtype)
        bm = None if body_material is None else np.ascontiguousarray(body_material, dtype=np.int32).reshape(-1)
        pm = None if plane_material is None else np.ascontiguousarray(plane_material, dtype=np.int32).reshape(-1)
        n = 0 if bm is None else bm.shape[0] // self.B
        self.lib.czo_world_set_materials(self.h, f.shape[0], f.ctypes.data_as(PR), r.ctypes.data_as(PR), first_world, n,
                                         None if bm is None else bm.ctypes.data_as(P32), None if pm is None else pm.ctypes.data_as(P32))

    def step(self, dt, n_steps: int = 1, n_threads: int = 1) -> dict:
        st = CzStepStats()
        self.lib.czo_world_step(self.h, self.prec.ctype(dt), n_steps, n_threads, C.byref(st))
        return st.as_dict()

    def download(self, first_world: int = 0, n_worlds: Optional[int] = None) -> Bodies:
        n = self.n_worlds - first_world if n_worlds is None else n_worlds
        out = Bodies(n * self.B, self.prec)
        st = out.struct()
        self.lib.czo_world_download_bodies(self.h, first_world, n, C.byref(st))
        return out

    def download_colliders(self, first_world: int = 0, n_worlds: Optional[int] = None) -> Colliders:
        n = self.n_worlds - first_world if n_worlds is None else n_worlds
        out = Colliders(n * self.B, self.prec)
        st = out.struct()
        self.lib.czo_world_download_colliders(self.h, first_world, n, C.byref(st))
        return out

    def contacts(self, world: int = 0) -> Contacts:
        # the oracle never truncates (contacts_per_world is only a hint for it): size the arrays from the count
        probe = Contacts(1, self.prec)
        pst = probe.struct()
        self.lib.czo_world_download_contacts(self.h, world, C.byref(pst))
        out = Contacts(max(int(pst.n), 1), self.prec)
        st = out.struct()
        rc = self.lib.czo_world_download_contacts(self.h, world, C.byref(st))
        assert rc == 0, rc
        return out.take(st)

    def contact_pairs(self, world: int = 0) -> List[Tuple[int, int]]:
        c = self.contacts(world)
        return list(zip(c.valid("body0").tolist(), c.valid("body1").tolist()))

    def last_counts(self):
        nc = np.zeros(self.n_worlds, dtype=np.int32)
        pi = np.zeros(self.n_worlds, dtype=np.int32)
        vi = np.zeros(self.n_worlds, dtype=np.int32)
        P32 = C.POINTER(C.c_int32)
        self.lib.czo_world_last_step_counts(self.h, nc.ctypes.data_as(P32), pi.ctypes.data_as(P32), vi.ctypes.data_as(P32))
        return nc, pi, vi

    def checksum_energy(self) -> Tuple[int, float]:
        cks, en = C.c_uint64(), C.c_double()
        self.lib.czo_world_checksum_energy(self.h, C.byref(cks), C.byref(en))
        return int(cks.value), float(en.value)

    def close(self):
        if self.h:
            self.lib.czo_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
