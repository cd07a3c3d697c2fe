"""Synthetic scenes of BASELINE.json's five configs (made concrete in SURVEY.md §8d).

Scene construction is host-side setup, exactly as in the reference's examples
(examples/cubedrop.go:146-177 `fire`, examples/ballistic.go:163-231); only primary state is
built here — derived data (transform, world inertia) is computed on the device at upload.
All random draws are splitmix64 + one multiply-add in float64 (no transcendental), so a Go or
C++ builder reproduces them bit for bit.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _abi
from .hostmath import block_inertia_tensor, inertia_tensor_coeffs, m3_invert, splitmix64_draws, uniform

INT32_MAX = np.int32(2**31 - 1)


@dataclass
class Scene:
    name: str
    prec: _abi.Precision
    n_worlds: int
    bodies_per_world: int
    bodies: _abi.Bodies
    colliders: _abi.Colliders
    planes: _abi.Planes
    schedule: int = _abi.SCHED_ALL_PAIRS_ORDERED
    check_one: Optional[np.ndarray] = None
    check_two: Optional[np.ndarray] = None
    active_from: Optional[np.ndarray] = None
    integrate: Optional[np.ndarray] = None
    contacts_per_world: int = 128
    dt: float = 1.0 / 60.0
    steps: int = 600
    notes: dict = field(default_factory=dict)
    # optional per-pair surface materials: dict(friction=MxM, restitution=MxM, body_material=[n_bodies], plane_material=[P])
    materials: Optional[dict] = None

    @property
    def n_bodies(self):
        return self.n_worlds * self.bodies_per_world


def ground_plane(prec):
    """examples/cubedrop.go:127"""
    return _abi.Planes([[0.0, 1.0, 0.0]], [0.0], prec)


def _cube_inverse_inertia(half, mass, prec):
    return m3_invert(block_inertia_tensor(half, mass, prec.dtype), prec.dtype)


def cubedrop(prec=_abi.F64, n_fire: int = 2, second_fire_step: int = 0) -> Scene:
    """cfg1: `fire()` n_fire times (examples/cubedrop.go:146-177), 4 cubes per call.
    second_fire_step > 0 gives the "staggered" variant (later calls spawn at that step)."""
    R = prec.dtype
    n = 4 * n_fire
    b = _abi.Bodies.defaults(n, prec)
    c = _abi.Colliders.defaults(n, prec)
    active = np.zeros(n, dtype=np.int32)
    iit = _cube_inverse_inertia((0.5, 0.5, 0.5), 8.0, prec)
    for f in range(n_fire):
        made = 4 * f
        offset = np.float32(0.75) if (made > 0 and (made // 4) % 2 >= 1) else np.float32(0.0)
        for i in range(4):
            k = made + i
            # m.Real(i*2.0-cubesToMake/2) - 0.5 + m.Real(offset)   (cubedrop.go:163)
            b.position[k] = (R(R(i * 2 - 2) - R(0.5)) + R(offset), R(10.0), R(0.0))
            b.inverse_mass[k] = R(1.0) / R(8.0)
            b.inverse_inertia_tensor[k] = iit
            c.shape[k] = _abi.SHAPE_CUBE
            c.half_size[k] = (0.5, 0.5, 0.5)
            active[k] = 0 if f == 0 else second_fire_step
    return Scene("cubedrop", prec, 1, n, b, c, ground_plane(prec), active_from=active, contacts_per_world=16 * n)


def ballistic(prec=_abi.F64, n_bullets: int = 64, first_step: int = 60, every: int = 8) -> Scene:
    """cfg2: examples/ballistic.go:163-231.  body 0 = cube, 1 = backboard (static, never
    integrated), 2.. = bullets; bullet k spawns at step first_step + every*k."""
    R = prec.dtype
    n = 2 + n_bullets
    b = _abi.Bodies.defaults(n, prec)
    c = _abi.Colliders.defaults(n, prec)
    # cube (ballistic.go:163-171)
    b.position[0] = (0.0, 5.0, 0.0)
    b.inverse_mass[0] = R(1.0) / R(8.0)
    b.inverse_inertia_tensor[0] = _cube_inverse_inertia((1.0, 1.0, 1.0), 8.0, prec)
    c.shape[0] = _abi.SHAPE_CUBE
    c.half_size[0] = (1.0, 1.0, 1.0)
    # backboard (ballistic.go:183-187): infinite mass, inverse inertia never set (zero)
    b.position[1] = (0.0, 2.0, -10.0)
    b.inverse_mass[1] = 0.0
    c.shape[1] = _abi.SHAPE_CUBE
    c.half_size[1] = (0.5, 2.0, 0.25)
    # bullets (ballistic.go:204-231)
    mass, radius = R(1.5), R(0.2)
    coeff = R(0.4) * mass * radius * radius
    iit = m3_invert(inertia_tensor_coeffs(coeff, coeff, coeff, 0.0, 0.0, 0.0, R), R)
    for k in range(n_bullets):
        i = 2 + k
        b.position[i] = (0.0, 1.5, 20.0)
        b.inverse_inertia_tensor[i] = iit
        b.inverse_mass[i] = R(1.0) / mass
        b.velocity[i] = (0.0, 0.0, -40.0)
        b.acceleration[i] = (0.0, -2.5, 0.0)
        c.shape[i] = _abi.SHAPE_SPHERE
        c.radius[i] = radius
    active = np.zeros(n, dtype=np.int32)
    active[2:] = first_step + every * np.arange(n_bullets, dtype=np.int32)
    integ = np.ones(n, dtype=np.uint8)
    integ[1] = 0
    # pair order of generateContacts (ballistic.go:47-97); plane 0 is encoded as -1
    one, two = [0, 0], [-1, 1]
    for k in range(n_bullets):
        bk = 2 + k
        one += [bk, 0, 1]
        two += [-1, bk, bk]
        for k2 in range(n_bullets):
            if k2 != k:
                one.append(2 + k2)
                two.append(bk)
    return Scene("ballistic", prec, 1, n, b, c, ground_plane(prec), schedule=_abi.SCHED_EXPLICIT,
                 check_one=np.asarray(one, dtype=np.int32), check_two=np.asarray(two, dtype=np.int32),
                 active_from=active, integrate=integ, contacts_per_world=max(64, 8 * n))


def pile(prec=_abi.F64, side: int = 16) -> Scene:
    """cfg3: side^3 jittered lattice of alternating cubes and spheres on the ground plane."""
    R = prec.dtype
    n = side ** 3
    b = _abi.Bodies.defaults(n, prec)
    c = _abi.Colliders.defaults(n, prec)
    idx = np.arange(n, dtype=np.int64)
    ix, iz, iy = idx % side, (idx // side) % side, idx // (side * side)
    u = splitmix64_draws((np.uint64(0xC0BE2) + idx.astype(np.uint64)), 3)
    j = uniform(u, -0.05, 0.05)
    half = (side - 1) / 2.0
    b.position[:, 0] = (1.25 * (ix - half) + j[:, 0]).astype(R)
    b.position[:, 1] = (0.75 + 1.25 * iy + j[:, 1]).astype(R)
    b.position[:, 2] = (1.25 * (iz - half) + j[:, 2]).astype(R)
    is_cube = ((ix + iy + iz) % 2) == 0
    cube_iit = _cube_inverse_inertia((0.5, 0.5, 0.5), 8.0, prec)
    coeff = R(0.4) * R(4.0) * R(0.5) * R(0.5)
    sph_iit = m3_invert(inertia_tensor_coeffs(coeff, coeff, coeff, 0.0, 0.0, 0.0, R), R)
    b.inverse_inertia_tensor[is_cube] = cube_iit
    b.inverse_inertia_tensor[~is_cube] = sph_iit
    b.inverse_mass[is_cube] = R(1.0) / R(8.0)
    b.inverse_mass[~is_cube] = R(1.0) / R(4.0)
    c.shape[is_cube] = _abi.SHAPE_CUBE
    c.shape[~is_cube] = _abi.SHAPE_SPHERE
    c.half_size[is_cube] = (0.5, 0.5, 0.5)
    c.radius[~is_cube] = R(0.5)
    return Scene("pile", prec, 1, n, b, c, ground_plane(prec), contacts_per_world=16 * n)


def archipelago(prec=_abi.F64, piles: int = 6, side: int = 2, spacing: float = 12.0) -> Scene:
    """ONE world made of piles x piles separate small piles (each a `side`^3 jittered lattice as in cfg3), far enough
    apart never to touch: the contact graph is a set of independent islands (the case the island resolver is for)."""
    R = prec.dtype
    unit = pile(prec, side)
    m = unit.bodies_per_world
    n = m * piles * piles
    b = _abi.Bodies.defaults(n, prec)
    c = _abi.Colliders.defaults(n, prec)
    for k in range(piles * piles):
        sl = slice(k * m, (k + 1) * m)
        off = np.array([spacing * (k % piles - (piles - 1) / 2.0), 0.0, spacing * (k // piles - (piles - 1) / 2.0)])
        b.position[sl] = (unit.bodies.position.astype(np.float64) + off).astype(R)
        b.position[sl, 1] += R(0.37 * (k % 5))          # piles land on different frames
        b.inverse_inertia_tensor[sl] = unit.bodies.inverse_inertia_tensor
        b.inverse_mass[sl] = unit.bodies.inverse_mass
        c.shape[sl] = unit.colliders.shape
        c.half_size[sl] = unit.colliders.half_size
        c.radius[sl] = unit.colliders.radius
    return Scene("archipelago", prec, 1, n, b, c, ground_plane(prec), contacts_per_world=16 * n)


def batched_cubedrop(prec=_abi.F64, n_worlds: int = 65536, first_world: int = 0) -> Scene:
    """cfg4: world w = cubedrop-8 with a per-world perturbation from splitmix64(1234 + w).
    `first_world` lets a rank build only its shard with the same global world ids."""
    R = prec.dtype
    base = cubedrop(prec)
    B = base.bodies_per_world
    n = n_worlds * B
    b = _abi.Bodies.defaults(n, prec)
    c = _abi.Colliders.defaults(n, prec)
    w = np.arange(first_world, first_world + n_worlds, dtype=np.uint64)
    u = splitmix64_draws(np.uint64(1234) + w, 6 * B).reshape(n_worlds, B, 6)
    pos = np.broadcast_to(base.bodies.position.astype(np.float64), (n_worlds, B, 3)).copy()
    pos[:, :, 0] += uniform(u[:, :, 0], -0.1, 0.1)
    pos[:, :, 2] += uniform(u[:, :, 1], -0.1, 0.1)
    pos[:, :, 1] += uniform(u[:, :, 2], -1.0, 1.0)
    q = np.empty((n_worlds, B, 4), dtype=np.float64)
    q[:, :, 0] = 1.0
    q[:, :, 1:] = uniform(u[:, :, 3:6], -0.1, 0.1)
    q = q.astype(R)
    # divide by its length: sqrt and divide only, in Real arithmetic, left to right
    ln = np.sqrt(((q[:, :, 0] * q[:, :, 0] + q[:, :, 1] * q[:, :, 1]) + q[:, :, 2] * q[:, :, 2]) + q[:, :, 3] * q[:, :, 3])
    q = q / ln[:, :, None]
    b.position[:] = pos.astype(R).reshape(n, 3)
    b.orientation[:] = q.reshape(n, 4)
    b.inverse_mass[:] = np.tile(base.bodies.inverse_mass, n_worlds)
    b.inverse_inertia_tensor[:] = np.tile(base.bodies.inverse_inertia_tensor, (n_worlds, 1))
    c.shape[:] = _abi.SHAPE_CUBE
    c.half_size[:] = (0.5, 0.5, 0.5)
    c.body[:] = np.tile(np.arange(B, dtype=np.int32), n_worlds)
    return Scene("batched_cubedrop", prec, n_worlds, B, b, c, ground_plane(prec), contacts_per_world=128,
                 notes={"first_world": first_world})


def free_bodies(prec=_abi.F64, n: int = 1 << 16, seed: int = 5) -> Scene:
    """cfg5 (host-built variant for parity tests; the 16M-body bench builds the same state on
    the device from the same splitmix64 stream, see cz_bench_integrate)."""
    R = prec.dtype
    b = _abi.Bodies.defaults(n, prec)
    c = _abi.Colliders.defaults(n, prec)
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        seeds = np.uint64(seed) * np.uint64(0x100000001B3) + idx * np.uint64(0x9E3779B97F4A7C15)
    u = splitmix64_draws(seeds, 24)
    b.position[:] = uniform(u[:, 0:3], -100, 100).astype(R)
    q = uniform(u[:, 3:7], -1, 1)
    ln = np.sqrt((q * q).sum(axis=1))
    q[ln < 0.1] = (1.0, 0.0, 0.0, 0.0)
    q = q.astype(R)
    ln = np.sqrt(((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2]) + q[:, 3] * q[:, 3])
    b.orientation[:] = q / ln[:, None]
    b.velocity[:] = uniform(u[:, 7:10], -5, 5).astype(R)
    b.rotation[:] = uniform(u[:, 10:13], -3, 3).astype(R)
    b.linear_damping[:] = uniform(u[:, 13], 0.90, 0.99).astype(R)
    b.angular_damping[:] = uniform(u[:, 14], 0.90, 0.99).astype(R)
    iit = np.zeros((n, 9), dtype=np.float64)
    iit[:, 0] = uniform(u[:, 15], 0.5, 2)
    iit[:, 4] = uniform(u[:, 16], 0.5, 2)
    iit[:, 8] = uniform(u[:, 17], 0.5, 2)
    iit[:, 1] = iit[:, 3] = uniform(u[:, 18], -0.1, 0.1)
    iit[:, 2] = iit[:, 6] = uniform(u[:, 19], -0.1, 0.1)
    iit[:, 5] = iit[:, 7] = uniform(u[:, 20], -0.1, 0.1)
    b.inverse_inertia_tensor[:] = iit.astype(R)
    b.inverse_mass[:] = R(1.0)
    return Scene("free_bodies", prec, 1, n, b, c, _abi.Planes(np.zeros((0, 3)), np.zeros(0), prec),
                 contacts_per_world=1)


def random_worlds(prec=_abi.F64, n_worlds: int = 64, bodies_per_world: int = 8, seed: int = 11, n_planes: int = 2,
                  extent: float = 1.8, height: float = 6.0) -> Scene:
    """Fuzz scene (no counterpart in the reference's examples): every world a random mix of cubes, spheres and
    collider-less bodies with their own sizes, masses, dampings, gravity, spin, sleep flags, activation steps and
    collider Offset matrices (rotation + translation), above one to three half-spaces (ground, a wall, a ramp).  It
    drives branches the five configs never reach together: sphere-sphere and cube-sphere contacts, non-identity
    Offsets, several planes, bodies that start asleep or cannot sleep, per-body damping.  Inputs only — the parity
    tests run the oracle and the CUDA path on the same arrays."""
    R = prec.dtype
    B, W = bodies_per_world, n_worlds
    n = W * B
    b = _abi.Bodies.defaults(n, prec)
    c = _abi.Colliders.defaults(n, prec)
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        seeds = np.uint64(seed) * np.uint64(0x100000001B3) + idx * np.uint64(0x9E3779B97F4A7C15)
    u = splitmix64_draws(seeds, 40)
    kind = uniform(u[:, 0], 0, 1)
    cube, sphere = kind < 0.47, (kind >= 0.47) & (kind < 0.92)
    half = uniform(u[:, 1:4], 0.25, 0.75)
    radius = uniform(u[:, 4], 0.25, 0.7)
    mass = uniform(u[:, 5], 1.0, 10.0).astype(R)
    c.shape[:] = _abi.SHAPE_NONE
    c.shape[cube] = _abi.SHAPE_CUBE
    c.shape[sphere] = _abi.SHAPE_SPHERE
    c.half_size[cube] = half[cube].astype(R)
    c.radius[sphere] = radius[sphere].astype(R)
    c.body[:] = np.tile(np.arange(B, dtype=np.int32), W)
    b.inverse_mass[:] = R(1.0) / mass
    for i in range(n):   # host-side setup arithmetic, as SetBlockInertiaTensor / the sphere coefficients of the examples do it
        if cube[i]:
            b.inverse_inertia_tensor[i] = _cube_inverse_inertia(tuple(c.half_size[i]), mass[i], prec)
        else:
            r = R(radius[i]) if sphere[i] else R(0.5)
            coeff = R(0.4) * mass[i] * r * r
            b.inverse_inertia_tensor[i] = m3_invert(inertia_tensor_coeffs(coeff, coeff, coeff, 0.0, 0.0, 0.0, R), R)
    b.position[:, 0] = uniform(u[:, 6], -extent, extent).astype(R)
    b.position[:, 2] = uniform(u[:, 7], -extent, extent).astype(R)
    b.position[:, 1] = uniform(u[:, 8], 0.9, height).astype(R)
    q = uniform(u[:, 9:13], -1, 1)
    q[np.sqrt((q * q).sum(axis=1)) < 0.1] = (1.0, 0.0, 0.0, 0.0)
    q = q.astype(R)
    ln = np.sqrt(((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2]) + q[:, 3] * q[:, 3])
    b.orientation[:] = q / ln[:, None]
    b.velocity[:] = uniform(u[:, 13:16], -2, 2).astype(R)
    b.rotation[:] = uniform(u[:, 16:19], -2, 2).astype(R)
    b.linear_damping[:] = uniform(u[:, 19], 0.90, 0.99).astype(R)
    b.angular_damping[:] = uniform(u[:, 20], 0.85, 0.99).astype(R)
    b.acceleration[uniform(u[:, 21], 0, 1) < 0.1] = (0.0, 0.0, 0.0)          # a few float
    b.can_sleep[:] = (uniform(u[:, 22], 0, 1) < 0.8).astype(np.uint8)
    asleep = uniform(u[:, 23], 0, 1) < 0.1
    b.is_awake[asleep] = 0
    b.velocity[asleep] = 0                                                      # SetAwake(false) zeroes them (rigidbody.go:188-190)
    b.rotation[asleep] = 0
    # collider Offset: a rotation about z and a small translation, for a third of the colliders.  The angle comes from a
    # table of Pythagorean pairs (cos, sin): no transcendental, so the Go harness builds the very same matrix
    off = uniform(u[:, 24], 0, 1) < 0.33
    table = np.array([[0.8, 0.6], [0.6, 0.8], [0.96, 0.28], [0.8, -0.6], [0.6, -0.8], [0.96, -0.28]])
    pick = np.minimum((uniform(u[:, 25], 0.0, 6.0)).astype(np.int64), 5)
    cs, sn = table[pick, 0], table[pick, 1]
    o = c.offset
    o[off, 0] = cs[off].astype(R); o[off, 1] = sn[off].astype(R); o[off, 3] = (-sn[off]).astype(R); o[off, 4] = cs[off].astype(R)
    o[off, 9:12] = uniform(u[:, 26:29], -0.2, 0.2)[off].astype(R)
    active = np.zeros(n, dtype=np.int32)
    late = uniform(u[:, 29], 0, 1) < 0.3
    active[late] = uniform(u[:, 30], 1, 40)[late].astype(np.int32)
    normals = [[0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [-0.6, 0.8, 0.0]][:max(1, min(3, n_planes))]
    offsets = [0.0, -(extent + 1.2), -1.5][:len(normals)]
    return Scene("random_worlds", prec, W, B, b, c, _abi.Planes(normals, offsets, prec), active_from=active,
                 contacts_per_world=max(64, 12 * B), notes={"seed": seed})


def with_materials(scene: Scene, seed: int = 7) -> Scene:
    """Three surface materials on any scene (0: the reference's 0.9 / 0.1, 1: slippery, 2: bouncy) with a
    deliberately asymmetric table, ids drawn per body from splitmix64; the first plane is slippery."""
    R = scene.prec.dtype
    fr = np.array([[0.9, 0.35, 0.8], [0.3, 0.05, 0.5], [0.8, 0.45, 1.1]], dtype=R)
    re = np.array([[0.1, 0.2, 0.55], [0.25, 0.0, 0.4], [0.6, 0.35, 0.7]], dtype=R)
    u = splitmix64_draws(np.uint64(seed) + np.arange(scene.n_bodies, dtype=np.uint64), 1).reshape(-1)
    ids = np.minimum((uniform(u, 0.0, 3.0)).astype(np.int32), 2)
    pm = np.zeros(scene.planes.n, dtype=np.int32)
    if pm.shape[0]:
        pm[0] = 1
    scene.materials = dict(friction=fr, restitution=re, body_material=ids, plane_material=pm)
    scene.name += "+materials"
    return scene
