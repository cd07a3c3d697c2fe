"""ctypes view of include/cubezcuda.h (the C ABI of libcubezcuda).

This module only declares structures and loads the shared library.  There is no CPU
fallback: `load()` raises if the CUDA library has not been built, and every compute entry
point returns CZ_ERR_CUDA (raised here as CubezError) when no device is present.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(HERE, "lib")

CZ_OK = 0
CZ_ERR_INVALID = -1
CZ_ERR_CUDA = -2
CZ_ERR_CAPACITY = -3
CZ_ERR_NIL_BODY = -4
CZ_ERR_NOMEM = -5

SHAPE_NONE, SHAPE_CUBE, SHAPE_SPHERE = 0, 1, 2
SCHED_ALL_PAIRS_ORDERED, SCHED_EXPLICIT = 0, 1
WORLD_BROADPHASE, WORLD_FUSED, WORLD_NO_FUSED = 1, 2, 4

OPS = dict(
    VEC_ADD=1, VEC_ADD_SCALED=2, VEC_COMPONENT_PRODUCT=3, VEC_CROSS=4, VEC_DOT=5, VEC_MAGNITUDE=6,
    VEC_SQUARE_MAGNITUDE=7, VEC_MUL_WITH=8, VEC_NORMALIZE=9, VEC_SUB=10, QUAT_MUL=20, QUAT_LEN=21,
    QUAT_NORMALIZE=22, QUAT_ROTATE=23, QUAT_ADD_SCALED_VECTOR=24, M3_MUL_M3=30, M3_INVERT=31, M3_MUL_V=32,
    M3_TRANSFORM_TRANSPOSE=33, M3_DETERMINANT=34, M34_MUL_M34=35, M34_MUL_V=36, M34_TRANSFORM_INVERSE=37,
    M34_SET_AS_TRANSFORM=38, REAL_EQUAL=40, TRANSFORM_INERTIA=41,
)
OP_OUT = {1: 3, 2: 3, 3: 3, 4: 3, 5: 1, 6: 1, 7: 1, 8: 3, 9: 3, 10: 3, 20: 4, 21: 1, 22: 4, 23: 3, 24: 4, 30: 9,
          31: 9, 32: 3, 33: 3, 34: 1, 35: 12, 36: 3, 37: 3, 38: 12, 40: 1, 41: 9}


class CubezError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libcubezcuda error {code}: {msg}")
        self.code = code


# field name -> (components, is_real)
BODY_FIELDS = (
    ("position", 3), ("orientation", 4), ("velocity", 3), ("rotation", 3), ("acceleration", 3),
    ("linear_damping", 1), ("angular_damping", 1), ("inverse_inertia_tensor", 9), ("inverse_mass", 1),
    ("motion", 1), ("is_awake", 0), ("can_sleep", 0), ("transform", 12), ("inverse_inertia_tensor_world", 9),
    ("last_frame_acceleration", 3),
)
COLLIDER_FIELDS = (("shape", -1), ("body", -1), ("offset", 12), ("transform", 12), ("half_size", 3), ("radius", 1))
CONTACT_FIELDS = (("body0", -1), ("body1", -1), ("friction", 1), ("restitution", 1), ("point", 3), ("normal", 3),
                  ("penetration", 1), ("check", -1))


def _ptr_type(comp: int, real):
    if comp == 0:
        return C.POINTER(C.c_uint8)
    if comp == -1:
        return C.POINTER(C.c_int32)
    return C.POINTER(real)


def make_structs(real):
    """Build the ctypes Structure classes for a given cz_real (c_double or c_float)."""

    class CzBodies(C.Structure):
        _fields_ = [("n", C.c_int32)] + [(name, _ptr_type(comp, real)) for name, comp in BODY_FIELDS]

    class CzColliders(C.Structure):
        _fields_ = [("n", C.c_int32)] + [(name, _ptr_type(comp, real)) for name, comp in COLLIDER_FIELDS]

    class CzPlanes(C.Structure):
        _fields_ = [("n", C.c_int32), ("normal", C.POINTER(real)), ("offset", C.POINTER(real))]

    class CzContacts(C.Structure):
        _fields_ = [("capacity", C.c_int32), ("n", C.c_int32)] + [(name, _ptr_type(comp, real)) for name, comp in CONTACT_FIELDS]

    return CzBodies, CzColliders, CzPlanes, CzContacts


class CzWorldDesc(C.Structure):
    _fields_ = [("n_worlds", C.c_int32), ("bodies_per_world", C.c_int32), ("contacts_per_world", C.c_int32),
                ("schedule", C.c_int32), ("flags", C.c_int32)]


class CzStepStats(C.Structure):
    _fields_ = [("steps", C.c_int64), ("contacts", C.c_int64), ("pos_iterations", C.c_int64),
                ("vel_iterations", C.c_int64), ("checks", C.c_int64), ("kernel_launches", C.c_int64),
                ("max_contacts", C.c_int32), ("status", C.c_int32), ("device_ms", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CzObs32(C.Structure):
    _fields_ = [("n", C.c_int32), ("position", C.POINTER(C.c_float)), ("orientation", C.POINTER(C.c_float)),
                ("velocity", C.POINTER(C.c_float)), ("rotation", C.POINTER(C.c_float))]


class CzRunTotals(C.Structure):
    _fields_ = [("checksum", C.c_uint64), ("energy", C.c_double), ("world_steps", C.c_int64), ("contacts", C.c_int64),
                ("pos_iterations", C.c_int64), ("vel_iterations", C.c_int64), ("max_device_ms", C.c_float), ("n_shards", C.c_int32),
                ("used_nccl", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Precision:
    """Everything that depends on cz_real: numpy dtype, ctypes type, struct classes."""

    def __init__(self, name: str):
        assert name in ("f64", "f32")
        self.name = name
        self.dtype = np.float64 if name == "f64" else np.float32
        self.ctype = C.c_double if name == "f64" else C.c_float
        self.Bodies, self.Colliders, self.Planes, self.Contacts = make_structs(self.ctype)

    def real(self, x):
        return self.dtype(x)


F64 = Precision("f64")
F32 = Precision("f32")


def precision(name: str) -> Precision:
    return F64 if name == "f64" else F32


def _np_dtype(comp: int, prec: Precision):
    if comp == 0:
        return np.uint8
    if comp == -1:
        return np.int32
    return prec.dtype


class ArrayRecord:
    """A set of named numpy arrays mirroring one of the cz_* SoA structs."""

    FIELDS = ()

    def __init__(self, n: int, prec: Precision, fields=None, **arrays):
        self.n = int(n)
        self.prec = prec
        self.a: Dict[str, Optional[np.ndarray]] = {}
        want = set(fields) if fields is not None else None
        for name, comp in self.FIELDS:
            if name in arrays:
                arr = np.ascontiguousarray(arrays[name], dtype=_np_dtype(comp, prec))
                self.a[name] = arr.reshape(self.shape_of(name))
            elif want is None or name in want:
                self.a[name] = np.zeros(self.shape_of(name), dtype=_np_dtype(comp, prec))
            else:
                self.a[name] = None

    def shape_of(self, name):
        comp = dict(self.FIELDS)[name]
        return (self.n,) if comp in (0, -1, 1) else (self.n, comp)

    def __getattr__(self, name):
        a = self.__dict__.get("a")
        if a is not None and name in a:
            return a[name]
        raise AttributeError(name)

    def _fill(self, st):
        for name, comp in self.FIELDS:
            arr = self.a[name]
            if arr is None:
                setattr(st, name, None)
            else:
                assert arr.flags["C_CONTIGUOUS"]
                setattr(st, name, arr.ctypes.data_as(_ptr_type(comp, self.prec.ctype)))
        return st

    def copy(self):
        out = type(self).__new__(type(self))
        out.__dict__.update({k: v for k, v in self.__dict__.items() if k != "a"})
        out.a = {k: (None if v is None else v.copy()) for k, v in self.a.items()}
        return out


class Bodies(ArrayRecord):
    FIELDS = BODY_FIELDS

    def struct(self):
        st = self.prec.Bodies()
        st.n = self.n
        return self._fill(st)

    @classmethod
    def defaults(cls, n: int, prec: Precision):
        """n bodies as NewRigidBody() leaves them (rigidbody.go:104-114)."""
        b = cls(n, prec)
        b.orientation[:, 0] = 1
        b.linear_damping[:] = prec.real(0.95)
        b.angular_damping[:] = prec.real(0.95)
        b.acceleration[:, 1] = prec.real(-9.78)
        b.inverse_inertia_tensor_world[:, (0, 4, 8)] = 1
        b.can_sleep[:] = 1
        b.is_awake[:] = 1
        b.motion[:] = prec.real(0.6)
        return b


class Colliders(ArrayRecord):
    FIELDS = COLLIDER_FIELDS

    def struct(self):
        st = self.prec.Colliders()
        st.n = self.n
        return self._fill(st)

    @classmethod
    def defaults(cls, n: int, prec: Precision):
        c = cls(n, prec)
        c.offset[:, (0, 4, 8)] = 1
        c.transform[:, (0, 4, 8)] = 1
        c.body[:] = np.arange(n, dtype=np.int32)
        return c


class Planes:
    def __init__(self, normals, offsets, prec: Precision):
        self.prec = prec
        self.normal = np.ascontiguousarray(normals, dtype=prec.dtype).reshape(-1, 3)
        self.offset = np.ascontiguousarray(offsets, dtype=prec.dtype).reshape(-1)
        self.n = self.offset.shape[0]

    def struct(self):
        st = self.prec.Planes()
        st.n = self.n
        st.normal = self.normal.ctypes.data_as(C.POINTER(self.prec.ctype))
        st.offset = self.offset.ctypes.data_as(C.POINTER(self.prec.ctype))
        return st


class Contacts(ArrayRecord):
    FIELDS = CONTACT_FIELDS

    def __init__(self, capacity: int, prec: Precision, **arrays):
        super().__init__(capacity, prec, **arrays)
        self.capacity = int(capacity)
        self.count = 0

    def struct(self):
        st = self.prec.Contacts()
        st.capacity = self.capacity
        st.n = self.count
        return self._fill(st)

    def take(self, st):
        self.count = int(st.n)
        return self

    def valid(self, name):
        return self.a[name][: self.count]


def lib_path(prec_name: str) -> str:
    return os.path.join(LIB_DIR, "libcubezcuda.so" if prec_name == "f64" else "libcubezcuda_f32.so")


_LIBS: Dict[str, C.CDLL] = {}


def declare(lib: C.CDLL, prec: Precision, prefix: str = "cz_"):
    """Attach argtypes/restype for every entry point declared in include/cubezcuda.h."""
    R = prec.ctype
    PR = C.POINTER(R)
    PB, PC, PP, PK = (C.POINTER(prec.Bodies), C.POINTER(prec.Colliders), C.POINTER(prec.Planes), C.POINTER(prec.Contacts))
    P32, PU8, VP = C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.c_void_p
    sig = {
        "init": ([C.c_int, C.POINTER(VP)], C.c_int),
        "shutdown": ([VP], C.c_int),
        "last_error": ([VP], C.c_char_p),
        "real_size": ([], C.c_int),
        "ctx_stream": ([VP], VP),
        "ctx_synchronize": ([VP], C.c_int),
        "host_alloc": ([VP, C.c_uint64, C.POINTER(VP)], C.c_int),
        "host_free": ([VP, VP], C.c_int),
        "integrate": ([VP, PB, R, PR, PR, PR], C.c_int),
        "calculate_derived_data": ([VP, PB], C.c_int),
        "collider_derive": ([VP, C.c_int32, PR, PR, PR], C.c_int),
        "narrowphase": ([VP, PC, PP, PB, C.c_int32, P32, P32, PK, PU8], C.c_int),
        "resolve_contacts": ([VP, C.c_int32, PK, PB, R, P32], C.c_int),
        "world_create": ([VP, C.POINTER(CzWorldDesc), C.POINTER(VP)], C.c_int),
        "world_destroy": ([VP], C.c_int),
        "world_upload_bodies": ([VP, C.c_int32, C.c_int32, PB, C.c_int32], C.c_int),
        "world_upload_colliders": ([VP, C.c_int32, C.c_int32, PC, C.c_int32], C.c_int),
        "world_upload_planes": ([VP, PP], C.c_int),
        "world_upload_schedule": ([VP, C.c_int32, P32, P32], C.c_int),
        "world_set_activation": ([VP, C.c_int32, C.c_int32, P32, PU8], C.c_int),
        "world_set_pow": ([VP, R, PR, PR, R], C.c_int),
        "world_set_step_index": ([VP, C.c_int64], C.c_int),
        "world_add_forces": ([VP, C.c_int32, C.c_int32, PR, PR], C.c_int),
        "world_set_episodes": ([VP, C.c_int32, P32], C.c_int),
        "world_set_materials": ([VP, C.c_int32, PR, PR, C.c_int32, C.c_int32, P32, P32], C.c_int),
        "world_export_gl": ([VP, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int32], C.c_int),
        "world_step": ([VP, R, C.c_int32, C.POINTER(CzStepStats)], C.c_int),
        "world_synchronize": ([VP], C.c_int),
        "world_download_bodies": ([VP, C.c_int32, C.c_int32, PB], C.c_int),
        "world_download_colliders": ([VP, C.c_int32, C.c_int32, PC], C.c_int),
        "world_download_contacts": ([VP, C.c_int32, PK], C.c_int),
        "world_last_step_counts": ([VP, P32, P32, P32], C.c_int),
        "world_count_nonfinite": ([VP, C.POINTER(C.c_int64)], C.c_int),
        "world_island_stats": ([VP, C.POINTER(C.c_int64), C.POINTER(C.c_int64)], C.c_int),
        "world_checksum_energy": ([VP, C.POINTER(C.c_uint64), C.POINTER(C.c_double)], C.c_int),
        "world_step_host": ([VP, PB, R, C.c_int32, C.POINTER(CzStepStats)], C.c_int),
        "world_step_rl": ([VP, PR, PR, PB, R, C.c_int32, C.POINTER(CzStepStats)], C.c_int),
        "world_step_rl_async": ([VP, PR, PR, PB, C.POINTER(CzObs32), R, C.c_int32, P32], C.c_int),
        "world_rl_wait": ([VP, C.c_int32, C.POINTER(CzStepStats)], C.c_int),
        "run_create": ([C.c_int32, P32, C.POINTER(CzWorldDesc), C.POINTER(VP)], C.c_int),
        "run_destroy": ([VP], C.c_int),
        "run_shard": ([VP, C.c_int32, C.POINTER(VP), P32, P32], C.c_int),
        "run_upload_bodies": ([VP, PB, C.c_int32], C.c_int),
        "run_upload_colliders": ([VP, PC, C.c_int32], C.c_int),
        "run_upload_planes": ([VP, PP], C.c_int),
        "run_set_episodes": ([VP, C.c_int32, P32], C.c_int),
        "run_step": ([VP, R, C.c_int32], C.c_int),
        "run_finish": ([VP, C.POINTER(CzRunTotals)], C.c_int),
        "run_last_error": ([VP], C.c_char_p),
        "bench_integrate": ([VP, C.c_int64, C.c_uint64, C.c_int32, C.c_int32, R, C.POINTER(C.c_float), C.POINTER(C.c_uint64)], C.c_int),
        "math_op": ([VP, C.c_int32, PR, PR], C.c_int),
        "bench_fp64_rate": ([VP, C.POINTER(C.c_double)], C.c_int),
        "broadphase_pairs": ([VP, C.c_int64, PR, PR, C.c_int64, P32, C.POINTER(C.c_int64)], C.c_int),
        "bench_broadphase": ([VP, C.c_int64, C.c_uint64, C.c_double, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_float)], C.c_int),
        "sort_pairs_u32": ([VP, C.c_int64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)], C.c_int),
        "sort_pairs_u64": ([VP, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_int32], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, prefix + name)
        fn.argtypes = args
        fn.restype = res
    return sig


EXPORTED = (
    "init shutdown last_error real_size ctx_stream ctx_synchronize host_alloc host_free integrate "
    "calculate_derived_data collider_derive narrowphase resolve_contacts world_create world_destroy "
    "world_upload_bodies world_upload_colliders world_upload_planes world_upload_schedule world_set_activation "
    "world_set_pow world_add_forces world_set_step_index world_set_episodes world_set_materials world_export_gl world_step world_synchronize world_download_bodies "
    "world_download_colliders world_download_contacts world_last_step_counts world_count_nonfinite world_island_stats world_checksum_energy "
    "world_step_host world_step_rl world_step_rl_async world_rl_wait run_create run_destroy run_shard run_upload_bodies run_upload_colliders run_upload_planes "
    "run_set_episodes run_step run_finish run_last_error bench_fp64_rate bench_integrate math_op broadphase_pairs bench_broadphase sort_pairs_u32 sort_pairs_u64"
).split()


def load(prec_name: str = "f64") -> C.CDLL:
    """Load libcubezcuda for the given precision.  Raises (never falls back) when missing."""
    if prec_name in _LIBS:
        return _LIBS[prec_name]
    path = lib_path(prec_name)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). cubez_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    declare(lib, precision(prec_name))
    if lib.cz_real_size() != (8 if prec_name == "f64" else 4):
        raise ImportError(f"{path}: cz_real size mismatch")
    _LIBS[prec_name] = lib
    return lib
