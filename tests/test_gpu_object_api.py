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
