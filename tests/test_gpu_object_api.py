"""GPU: the reference's object API under its own names (RigidBody, CollisionCube,
CheckForCollisions, ResolveContacts) driven exactly like examples/cubedrop.go:29-75, compared
with the oracle's world loop.  Reads like a test written against the reference."""
import numpy as np
import pytest

from cubez_b200 import scenes
from cubez_b200.hostmath import block_inertia_tensor
from oracle_lib import OracleWorld

pytestmark = pytest.mark.gpu


def fire(cubes, cubez):
    """examples/cubedrop.go:146-177"""
    offset = 0.75 if (len(cubes) > 0 and (len(cubes) // 4) % 2 >= 1) else 0.0
    for i in range(4):
        c = cubez.CollisionCube(None, [0.5, 0.5, 0.5])
        c.Body.Position = np.array([float(i * 2 - 2) - 0.5 + offset, 10.0, 0.0])
        c.Body.SetMass(8.0)
        c.Body.CanSleep = True
        c.Body.SetInertiaTensor(block_inertia_tensor(c.HalfSize, 8.0))
        c.Body.CalculateDerivedData()
        c.CalculateDerivedData()
        cubes.append(c)


def test_cubedrop_loop_through_object_api():
    from cubez_b200 import api as cubez
    cubes = []
    fire(cubes, cubez)
    fire(cubes, cubez)
    ground = cubez.CollisionPlane([0.0, 1.0, 0.0], 0.0)
    scene = scenes.cubedrop()
    oracle = OracleWorld.from_scene(scene)
    delta = scene.dt
    for frame in range(100):
        # updateObjects (cubedrop.go:29-39)
        for cube in cubes:
            cube.GetBody().Integrate(delta)
            cube.CalculateDerivedData()
        # generateContacts (cubedrop.go:42-67)
        returnFound, contacts = False, []
        for cube in cubes:
            found, contacts = cube.CheckAgainstHalfSpace(ground, contacts)
            returnFound |= found
            for other in cubes:
                if other is cube:
                    continue
                found, contacts = cubez.CheckForCollisions(cube, other, contacts)
                returnFound |= found
        if returnFound:
            cubez.ResolveContacts(len(contacts) * 8, contacts, delta)
        oracle.step(delta, 1)
        assert len(contacts) == oracle.last_counts()[0][0], frame
        if frame in (0, 1, 50, 86, 99):
            ref = oracle.download()
            for i, cube in enumerate(cubes):
                assert np.array_equal(cube.Body.Position, ref.position[i]), (frame, i)
                assert np.array_equal(cube.Body.Orientation, ref.orientation[i])
                assert np.array_equal(cube.Body.Velocity, ref.velocity[i])
                assert np.array_equal(cube.Body.GetTransform(), ref.transform[i])


def test_batched_check_list_equals_single_calls():
    from cubez_b200 import api as cubez
    a = cubez.CollisionSphere(None, 0.5); a.Body.Position = np.array([0.0, 0.4, 0.0]); a.Body.SetMass(1.0)
    b = cubez.CollisionCube(None, [0.5, 0.5, 0.5]); b.Body.Position = np.array([0.6, 0.45, 0.0]); b.Body.SetMass(1.0)
    for c in (a, b):
        c.Body.CalculateDerivedData(); c.CalculateDerivedData()
    plane = cubez.CollisionPlane([0.0, 1.0, 0.0], 0.0)
    found, contacts = cubez.check_collision_list([(a, plane), (a, b), (b, a), (b, plane), (plane, a), (plane, plane)])
    assert found == [True, True, True, True, True, False]
    # sphere-vs-cube is canonicalised to Bodies (cube, sphere) whatever the call direction (colliders.go:210-213)
    two_body = [c for c in contacts if c.Bodies[1] is not None]
    assert all(c.Bodies[0] is b.Body and c.Bodies[1] is a.Body for c in two_body) and len(two_body) == 2
    single = []
    for one, two in [(a, plane), (a, b), (b, a), (b, plane), (plane, a)]:
        _, single = cubez.CheckForCollisions(one, two, single)
    assert len(single) == len(contacts)
    for x, y in zip(single, contacts):
        assert np.array_equal(x.ContactPoint, y.ContactPoint) and x.Penetration == y.Penetration


def test_random_scene_loop_through_object_api():
    """A random scene (cubes and spheres of different sizes and masses, rotated and shifted collider Offsets, two planes,
    bodies that start asleep or cannot sleep) driven through the reference-named objects with the batch entry points —
    integrate_bodies, check_collision_list in the all-pairs order of examples/cubedrop.go:42-67, ResolveContacts —
    against the oracle's world loop, state bits compared every few frames."""
    from cubez_b200 import _abi, api as cubez
    scene = scenes.random_worlds(_abi.F64, n_worlds=1, bodies_per_world=11, seed=91, n_planes=2)
    scene.active_from[:] = 0                     # the object loop has no activation steps
    b, c = scene.bodies, scene.colliders
    objs = []
    for i in range(scene.bodies_per_world):
        if c.shape[i] == _abi.SHAPE_CUBE:
            col = cubez.CollisionCube(None, c.half_size[i])
        elif c.shape[i] == _abi.SHAPE_SPHERE:
            col = cubez.CollisionSphere(None, c.radius[i])
        else:
            col = None
        body = col.Body if col is not None else cubez.RigidBody()
        body.Position, body.Orientation = b.position[i].copy(), b.orientation[i].copy()
        body.Velocity, body.Rotation, body.Acceleration = b.velocity[i].copy(), b.rotation[i].copy(), b.acceleration[i].copy()
        body.LinearDamping, body.AngularDamping = b.linear_damping[i], b.angular_damping[i]
        body.SetMass(1.0)
        body._inverse_mass = b.inverse_mass[i]
        body._mass = type(b.inverse_mass[i])(1.0) / b.inverse_mass[i]
        body.InverseInertiaTensor = b.inverse_inertia_tensor[i].copy()
        body.CanSleep, body.IsAwake = bool(b.can_sleep[i]), bool(b.is_awake[i])
        body.CalculateDerivedData()
        if col is not None:
            col.Offset = c.offset[i].copy()
            col.CalculateDerivedData()
        objs.append((body, col))
    planes = [cubez.CollisionPlane(scene.planes.normal[k], scene.planes.offset[k]) for k in range(scene.planes.n)]
    oracle = OracleWorld.from_scene(scene)
    colliders = [col for _, col in objs if col is not None]
    delta = scene.dt
    for frame in range(45):
        cubez.integrate_bodies([body for body, _ in objs], delta)
        for col in colliders:
            col.CalculateDerivedData()
        checks = []
        for col in colliders:                    # cubedrop.go:42-67: every collider against the planes, then against every other
            for pl in planes:
                checks.append((col, pl))
            for other in colliders:
                if other is not col:
                    checks.append((col, other))
        found, contacts = cubez.check_collision_list(checks)
        if any(found):
            cubez.ResolveContacts(len(contacts) * 8, contacts, delta)
        oracle.step(delta, 1)
        assert len(contacts) == oracle.last_counts()[0][0], frame
        if frame % 6 == 0 or frame == 44:
            ref = oracle.download()
            for i, (body, _) in enumerate(objs):
                assert np.array_equal(body.Position, ref.position[i]), (frame, i)
                assert np.array_equal(body.Orientation, ref.orientation[i]), (frame, i)
                assert np.array_equal(body.Velocity, ref.velocity[i]), (frame, i)
                assert np.array_equal(body.Rotation, ref.rotation[i]), (frame, i)
                assert bool(body.IsAwake) == bool(ref.is_awake[i]), (frame, i)
