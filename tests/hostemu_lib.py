"""ctypes wrapper of tests/hostemu (CPU run of the product's device functions).  TEST ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from cubez_b200 import _abi
from cubez_b200._abi import Bodies, Contacts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIBS = {}


def load(prec="f64"):
    if prec in _LIBS:
        return _LIBS[prec]
    path = os.path.join(ROOT, "tests", "hostemu", "_build", f"libhostemu_{prec}.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-s"], cwd=os.path.join(ROOT, "tests", "hostemu"))
    lib = C.CDLL(path)
    p = _abi.precision(prec)
    P32, PU8 = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    lib.cze_run.argtypes = [C.POINTER(p.Bodies), C.POINTER(p.Colliders), C.POINTER(p.Planes), C.c_int, C.c_int, P32, P32, P32, PU8,
                            p.ctype, C.c_int, C.c_int, P32, P32, P32, C.POINTER(C.c_uint64), C.POINTER(p.Contacts)]
    lib.cze_run.restype = C.c_int
    PR = C.POINTER(p.ctype)
    lib.cze_set_materials.argtypes = [C.c_int, PR, PR, C.c_int, P32, C.c_int, P32]
    lib.cze_set_materials.restype = C.c_int
    _LIBS[prec] = lib
    return lib


def run_scene(scene, n_steps, oracle_initial: Bodies, colliders_initial):
    """Step one-world `scene` n_steps frames on the CPU emulation.  oracle_initial /
    colliders_initial supply the derived fields (transform etc.) of step 0."""
    prec = scene.prec
    lib = load(prec.name)
    io = oracle_initial.copy()
    col = colliders_initial.copy()
    col.a["shape"] = scene.colliders.shape.copy()
    col.a["half_size"] = scene.colliders.half_size.copy()
    col.a["radius"] = scene.colliders.radius.copy()
    col.a["offset"] = scene.colliders.offset.copy()
    counts = np.zeros(n_steps, dtype=np.int32)
    pos = np.zeros(n_steps, dtype=np.int32)
    vel = np.zeros(n_steps, dtype=np.int32)
    ph = np.zeros(n_steps, dtype=np.uint64)
    last = Contacts(scene.contacts_per_world, prec)
    P32, PU8 = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    one = scene.check_one if scene.check_one is not None else np.zeros(1, dtype=np.int32)
    two = scene.check_two if scene.check_two is not None else np.zeros(1, dtype=np.int32)
    af = scene.active_from if scene.active_from is not None else np.zeros(scene.bodies_per_world, dtype=np.int32)
    ig = scene.integrate if scene.integrate is not None else np.ones(scene.bodies_per_world, dtype=np.uint8)
    ist, cst, pst, lst = io.struct(), col.struct(), scene.planes.struct(), last.struct()
    mat = getattr(scene, "materials", None)
    if mat:
        PR = C.POINTER(prec.ctype)
        f = np.ascontiguousarray(mat["friction"], dtype=prec.dtype)
        r = np.ascontiguousarray(mat["restitution"], dtype=prec.dtype)
        bm = np.ascontiguousarray(mat["body_material"], dtype=np.int32)
        pm = np.ascontiguousarray(mat["plane_material"], dtype=np.int32)
        lib.cze_set_materials(f.shape[0], f.ctypes.data_as(PR), r.ctypes.data_as(PR), bm.shape[0], bm.ctypes.data_as(P32), pm.shape[0], pm.ctypes.data_as(P32))
    else:
        lib.cze_set_materials(0, None, None, 0, None, 0, None)
    rc = lib.cze_run(C.byref(ist), C.byref(cst), C.byref(pst), scene.schedule, 0 if scene.check_one is None else one.shape[0],
                     one.ctypes.data_as(P32), two.ctypes.data_as(P32), af.ctypes.data_as(P32), ig.ctypes.data_as(PU8),
                     prec.ctype(scene.dt), n_steps, scene.contacts_per_world, counts.ctypes.data_as(P32), pos.ctypes.data_as(P32),
                     vel.ctypes.data_as(P32), ph.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(lst))
    assert rc == 0, rc
    last.take(lst)
    return io, counts, pos, vel, ph, last


def pair_hash(pairs):
    h = 0xcbf29ce484222325
    for a, b in pairs:
        for v in (a, b):
            h ^= v & 0xFFFFFFFF
            h = (h * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h
