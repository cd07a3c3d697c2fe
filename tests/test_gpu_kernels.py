"""GPU: per-kernel differential tests, CUDA path (through the C ABI) vs the CPU oracle on the
same seeded inputs.  Bit-exact: every op on the path is an IEEE add/mul/div/sqrt and both
sides are built without FMA contraction."""
import numpy as np
import pytest

from cubez_b200 import _abi, scenes
from cubez_b200._abi import Bodies, Colliders, Contacts, Planes
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu
ALL = ("position", "orientation", "velocity", "rotation", "motion", "is_awake", "transform", "inverse_inertia_tensor_world",
       "last_frame_acceleration")


@pytest.fixture(scope="module", params=["f64", "f32"])
def both(request):
    from cubez_b200.api import Context
    return Context.get(0, request.param), Oracle(request.param), _abi.precision(request.param)


def assert_bodies_equal(a, b, fields=ALL):
    for f in fields:
        assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True), f"{f} differs"


def test_integrate_free_bodies(both):
    """K1 over 2^16 random free bodies, 5 frames (cfg5 shape at a size the oracle finishes fast)."""
    gpu, cpu, prec = both
    sc = scenes.free_bodies(prec, n=1 << 16)
    a, b = sc.bodies.copy(), sc.bodies.copy()
    for _ in range(5):
        gpu.integrate(a, sc.dt)
        cpu.integrate(b, sc.dt)
    assert_bodies_equal(a, b)


def test_integrate_edge_cases(both):
    """sleeping bodies are skipped; CanSleep=false never sleeps; low motion puts a body to sleep
    and zeroes its velocities; motion is clamped at 3.0; zero / denormal quaternions."""
    gpu, cpu, prec = both
    n = 8
    b = Bodies.defaults(n, prec)
    b.inverse_inertia_tensor[:, (0, 4, 8)] = 1
    b.is_awake[0] = 0; b.velocity[0] = (1, 2, 3)
    b.can_sleep[1] = 0; b.acceleration[1] = 0; b.motion[1] = 0.0
    b.acceleration[2] = 0; b.motion[2] = 0.31; b.velocity[2] = (1e-3, 0, 0)          # falls asleep
    b.velocity[3] = (100, 0, 0); b.rotation[3] = (0, 50, 0)                           # clamp
    b.orientation[4] = (0, 0, 0, 0)                                                   # -> identity
    b.orientation[5] = (0, 1e-13, 0, 0)
    b.orientation[6] = (1, 1e-8, 0, 0)                                                # RealEqual skip
    b.velocity[7] = (-0.0, 0.0, -0.0); b.acceleration[7] = (-0.0, 0, 0)               # signed zeros
    a, c = b.copy(), b.copy()
    for _ in range(3):
        gpu.integrate(a, 1.0 / 60.0)
        cpu.integrate(c, 1.0 / 60.0)
    assert_bodies_equal(a, c)
    assert a.is_awake[2] == 0 and np.all(a.velocity[2] == 0) and a.is_awake[1] == 1
    assert a.motion[3] == prec.dtype(3.0)
    assert np.array_equal(a.velocity[0], [1, 2, 3])
    # raw bits (catches -0 vs +0)
    assert a.velocity.tobytes() == c.velocity.tobytes() and a.last_frame_acceleration.tobytes() == c.last_frame_acceleration.tobytes()


def test_integrate_uses_host_pow_when_given(both):
    gpu, cpu, prec = both
    sc = scenes.free_bodies(prec, n=257)
    dt = prec.dtype(0.01)
    lp = np.power(sc.bodies.linear_damping.astype(np.float64), float(dt)).astype(prec.dtype)
    ap = np.power(sc.bodies.angular_damping.astype(np.float64), float(dt)).astype(prec.dtype)
    bias = prec.dtype(np.float64(0.5) ** float(dt))
    a, b = sc.bodies.copy(), sc.bodies.copy()
    gpu.integrate(a, dt, lp, ap, bias)
    cpu.integrate(b, dt, lp, ap, bias)
    assert_bodies_equal(a, b)


def test_calculate_derived_data_and_collider_derive(both):
    gpu, cpu, prec = both
    sc = scenes.free_bodies(prec, n=1000)
    sc.bodies.orientation[:] *= prec.dtype(1.7)          # not normalised
    a, b = sc.bodies.copy(), sc.bodies.copy()
    gpu.calculate_derived_data(a)
    cpu.calculate_derived_data(b)
    assert_bodies_equal(a, b, ("orientation", "transform", "inverse_inertia_tensor_world"))
    rng = np.random.default_rng(3)
    off = rng.uniform(-1, 1, (1000, 12)).astype(prec.dtype)
    off[:10] = (1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0)
    assert np.array_equal(gpu.collider_derive(a.transform, off), cpu.collider_derive(b.transform, off))


def random_colliders(prec, n, rng, spread):
    """n cubes/spheres with random poses in a box of size `spread` (dense enough to overlap)."""
    b = Bodies.defaults(n, prec)
    b.position[:] = rng.uniform(-spread, spread, (n, 3))
    q = rng.normal(size=(n, 4))
    b.orientation[:] = q / np.linalg.norm(q, axis=1, keepdims=True)
    b.velocity[:] = rng.uniform(-1, 1, (n, 3))
    b.inverse_inertia_tensor[:, (0, 4, 8)] = 1
    Oracle(prec.name).calculate_derived_data(b)
    c = Colliders.defaults(n, prec)
    c.shape[:] = rng.integers(1, 3, n)
    c.half_size[:] = rng.uniform(0.3, 1.0, (n, 3))
    c.radius[:] = rng.uniform(0.3, 1.0, n)
    c.transform[:] = b.transform
    return b, c


@pytest.mark.parametrize("seed", [1, 2])
def test_narrowphase_random_pairs(both, seed):
    """All ordered pairs + plane checks of 48 random colliders: same contacts, same order."""
    gpu, cpu, prec = both
    rng = np.random.default_rng(seed)
    n = 48
    b, c = random_colliders(prec, n, rng, 2.5)
    planes = Planes([[0, 1, 0], [0.6, 0.8, 0]], [0.0, -1.0], prec)
    one, two = [], []
    for i in range(n):
        one += [i, i, -1, -2]; two += [-1, -2, i, -1]          # collider-plane, plane-collider, plane-plane
        for j in range(n):
            if j != i:
                one.append(i); two.append(j)
    ga, gf = gpu.narrowphase(c, planes, b, one, two, capacity=1 << 14)
    ca, cf = cpu.narrowphase(c, planes, b, one, two, capacity=1 << 14)
    assert ga.count == ca.count and ga.count > 50
    assert np.array_equal(gf, cf)
    for f in ("body0", "body1", "point", "normal", "penetration", "friction", "restitution"):
        assert np.array_equal(ga.valid(f), ca.valid(f), equal_nan=True), f


def test_narrowphase_touching_and_degenerate_cases(both):
    gpu, cpu, prec = both
    n = 8
    b = Bodies.defaults(n, prec)
    b.inverse_inertia_tensor[:, (0, 4, 8)] = 1
    pos = [(0, 0.5, 0), (1.0, 0.5, 0), (0, 0.5, 0), (5, 1.0, 0), (5, 1.0, 0), (9, 0.2, 0), (9, 0.2, 0.4), (20, 0.5, 0)]
    b.position[:] = pos
    b.velocity[4] = (0, 0, 2.0)          # cube-sphere coincident-centre fallback normal (colliders.go:417-421)
    Oracle(prec.name).calculate_derived_data(b)
    c = Colliders.defaults(n, prec)
    c.shape[:] = [1, 1, 1, 1, 2, 2, 2, 2]
    c.half_size[:] = 0.5
    c.radius[:] = 0.2
    c.transform[:] = b.transform
    planes = Planes([[0, 1, 0]], [0.0], prec)
    one = [0, 1, 0, 2, 3, 4, 5, 6, 7, 5, 0, 1, 2, 3, 4, 5, 6, 7]
    two = [1, 0, 2, 0, 4, 3, 6, 5, 7, 5, -1, -1, -1, -1, -1, -1, -1, -1]   # touching cubes, coincident cubes, coincident spheres...
    ga, gf = gpu.narrowphase(c, planes, b, one, two)
    ca, cf = cpu.narrowphase(c, planes, b, one, two)
    assert ga.count == ca.count and np.array_equal(gf, cf)
    for f in ("body0", "body1", "point", "normal", "penetration"):
        assert np.array_equal(ga.valid(f), ca.valid(f), equal_nan=True), f


def test_narrowphase_empty_and_capacity(both):
    gpu, cpu, prec = both
    rng = np.random.default_rng(0)
    b, c = random_colliders(prec, 4, rng, 50.0)      # far apart: no contacts
    ga, gf = gpu.narrowphase(c, None, b, [0, 1], [1, 2])
    assert ga.count == 0 and not gf.any()
    ga, gf = gpu.narrowphase(c, None, b, [], [])
    assert ga.count == 0
    b2, c2 = random_colliders(prec, 32, rng, 0.5)    # everything overlaps
    one = [i for i in range(32) for j in range(32) if i != j]
    two = [j for i in range(32) for j in range(32) if i != j]
    from cubez_b200._abi import CubezError
    with pytest.raises(CubezError) as e:
        gpu.narrowphase(c2, None, b2, one, two, capacity=16)
    assert e.value.code == _abi.CZ_ERR_CAPACITY


@pytest.mark.parametrize("seed,n_bodies,n_contacts", [(1, 6, 20), (2, 40, 300), (3, 3, 5)])
def test_resolve_contacts_random(both, seed, n_bodies, n_contacts):
    """ResolveContacts on random contact soups (one-body and two-body contacts, a static body,
    a sleeping body that gets woken): bodies and contacts identical to the oracle."""
    gpu, cpu, prec = both
    rng = np.random.default_rng(seed)
    b = Bodies.defaults(n_bodies, prec)
    b.position[:] = rng.uniform(-2, 2, (n_bodies, 3))
    q = rng.normal(size=(n_bodies, 4))
    b.orientation[:] = q / np.linalg.norm(q, axis=1, keepdims=True)
    b.velocity[:] = rng.uniform(-2, 2, (n_bodies, 3))
    b.rotation[:] = rng.uniform(-1, 1, (n_bodies, 3))
    b.inverse_mass[:] = rng.uniform(0.1, 1.0, n_bodies)
    b.inverse_inertia_tensor[:, (0, 4, 8)] = rng.uniform(0.5, 2.0, (n_bodies, 3))
    b.last_frame_acceleration[:] = (0, -9.78, 0)
    b.inverse_mass[0] = 0; b.inverse_inertia_tensor[0] = 0; b.last_frame_acceleration[0] = 0       # static
    b.is_awake[1] = 0; b.velocity[1] = 0; b.rotation[1] = 0                                         # asleep
    cpu.calculate_derived_data(b)
    cs = Contacts(n_contacts, prec)
    cs.count = n_contacts
    cs.body0[:] = rng.integers(0, n_bodies, n_contacts)
    other = rng.integers(-1, n_bodies, n_contacts)
    other[other == cs.body0] = -1
    cs.body1[:] = other
    swap = rng.random(n_contacts) < 0.1
    swap &= cs.body1 >= 0
    nil_first = rng.random(n_contacts) < 0.1
    nil_first &= cs.body1 < 0
    cs.body1[nil_first] = cs.body0[nil_first]; cs.body0[nil_first] = -1                              # Bodies[0] == nil path
    nrm = rng.normal(size=(n_contacts, 3))
    cs.normal[:] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    cs.point[:] = rng.uniform(-2, 2, (n_contacts, 3))
    cs.penetration[:] = rng.uniform(-0.05, 0.3, n_contacts)
    cs.friction[:] = 0.9
    cs.restitution[:] = 0.1
    cs.friction[rng.random(n_contacts) < 0.1] = 0.5
    ga, gb, ca, cb = cs.copy(), b.copy(), cs.copy(), b.copy()
    ga.capacity = ca.capacity = n_contacts; ga.count = ca.count = n_contacts
    gi = gpu.resolve_contacts(8 * n_contacts, ga, gb, 1.0 / 60.0)
    ci = cpu.resolve_contacts(8 * n_contacts, ca, cb, 1.0 / 60.0)
    assert gi == ci and gi[0] + gi[1] > 0
    assert_bodies_equal(gb, cb)
    for f in ("body0", "body1", "normal", "penetration"):
        assert np.array_equal(ga.valid(f), ca.valid(f), equal_nan=True), f


def test_resolve_contacts_iteration_cap_and_noops(both):
    gpu, cpu, prec = both
    sc = scenes.cubedrop(prec)
    b = sc.bodies.copy()
    cpu.calculate_derived_data(b)
    b.position[:, 1] = 0.45
    cpu.calculate_derived_data(b)
    col = sc.colliders.copy(); col.a["transform"] = b.transform.copy()
    contacts, _ = cpu.narrowphase(col, sc.planes, b, list(range(8)), [-1] * 8)
    assert contacts.count == 32
    for cap in (0, 1, 5, 8 * 32):
        ga, gb, ca, cb = contacts.copy(), b.copy(), contacts.copy(), b.copy()
        for x in (ga, ca):
            x.capacity = contacts.capacity; x.count = contacts.count
        assert gpu.resolve_contacts(cap, ga, gb, 1.0 / 60.0) == cpu.resolve_contacts(cap, ca, cb, 1.0 / 60.0)
        assert_bodies_equal(gb, cb)
    # duration <= 0 and empty contact list are no-ops (contact.go:210-212)
    gb = b.copy()
    assert gpu.resolve_contacts(10, contacts.copy(), gb, 0.0) == (0, 0)
    assert_bodies_equal(gb, b)


def test_frictionless_one_body_contact_reports_nil_dereference(both):
    """contact.go:512-523: the reference dereferences Bodies[1] == nil (a Go panic)."""
    gpu, cpu, prec = both
    b = Bodies.defaults(1, prec)
    b.inverse_mass[:] = 1; b.inverse_inertia_tensor[:, (0, 4, 8)] = 1; b.velocity[0] = (0, -1, 0)
    cpu.calculate_derived_data(b)
    cs = Contacts(1, prec); cs.count = 1
    cs.body0[0] = 0; cs.body1[0] = -1; cs.normal[0] = (0, 1, 0); cs.penetration[0] = 0.0; cs.friction[0] = 0.0; cs.restitution[0] = 0.1
    from cubez_b200._abi import CubezError
    with pytest.raises(CubezError) as e:
        gpu.resolve_contacts(8, cs, b, 1.0 / 60.0)
    assert e.value.code == _abi.CZ_ERR_NIL_BODY


def test_resolve_static_static_contact_gives_nan_like_the_reference(both):
    """Two infinite-mass bodies in contact: totalInertia = 0 -> 0/0 = NaN moves (contact.go:329-330).
    The reference propagates the NaN silently and never selects a NaN contact again; so do we."""
    gpu, cpu, prec = both
    b = Bodies.defaults(3, prec)
    b.inverse_mass[:] = (0, 0, 1); b.inverse_inertia_tensor[2, (0, 4, 8)] = 1
    b.position[:] = ((0, 0, 0), (0, 0.9, 0), (3, 0, 0))
    cpu.calculate_derived_data(b)
    cs = Contacts(2, prec); cs.count = 2
    cs.body0[:] = (0, 2); cs.body1[:] = (1, -1)
    cs.normal[:] = ((0, -1, 0), (0, 1, 0)); cs.point[:] = ((0, 0.45, 0), (3, -0.5, 0)); cs.penetration[:] = (0.1, 0.05)
    cs.friction[:] = 0.9; cs.restitution[:] = 0.1
    ga, gb, ca, cb = cs.copy(), b.copy(), cs.copy(), b.copy()
    for x in (ga, ca):
        x.capacity = 2; x.count = 2
    with np.errstate(all="ignore"):
        assert gpu.resolve_contacts(16, ga, gb, 1.0 / 60.0) == cpu.resolve_contacts(16, ca, cb, 1.0 / 60.0)
    assert_bodies_equal(gb, cb)
    assert np.isnan(gb.position[0]).any() and not np.isnan(gb.position[2]).any()
    assert np.array_equal(ga.valid("penetration"), ca.valid("penetration"), equal_nan=True)
