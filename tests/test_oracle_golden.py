"""CPU: the oracle reproduces the committed golden fixtures bit for bit (regression pin of
the checker itself), and the product's device functions — compiled for the host by
tests/hostemu — agree with the oracle bit for bit on the same scenes."""
import numpy as np
import pytest

import hostemu_lib as he
from golden_cases import CASES, STATE_FIELDS, check_against_golden, load_golden
from oracle_lib import OracleWorld


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name):
    make, n = CASES[name]
    scene = make()
    check_against_golden(OracleWorld.from_scene(scene), scene, load_golden(name), n, he.pair_hash)


@pytest.mark.parametrize("name", ["cubedrop_f64", "cubedrop_f32", "cubedrop_staggered_f64", "ballistic16_f64", "pile27_f64",
                                  "cubedrop_materials_f64", "pile27_materials_f32"])
def test_device_functions_on_host_match_golden(name):
    """cz_math/cz_body/cz_narrow/cz_resolve .cuh run on the CPU == golden (no GPU involved)."""
    make, n = CASES[name]
    scene = make()
    gold = load_golden(name)
    w = OracleWorld.from_scene(scene)     # only used to obtain derived initial data (transform, world inertia)
    io, counts, pos, vel, ph, _ = he.run_scene(scene, n, w.download(), w.download_colliders())
    assert np.array_equal(counts, gold["counts"][:, 0])
    assert np.array_equal(pos, gold["pos_iters"][:, 0])
    assert np.array_equal(vel, gold["vel_iters"][:, 0])
    assert np.array_equal(ph, gold["pair_hash"][:, 0])
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(io, f), gold[f]), f


def test_oracle_object_api_matches_world_loop():
    """Driving the oracle through Integrate / narrowphase / ResolveContacts calls in the order of
    examples/cubedrop.go:29-75 gives the same trajectory as its world handle."""
    from cubez_b200 import _abi, scenes
    from cubez_b200._abi import Contacts
    from oracle_lib import Oracle
    sc = scenes.cubedrop()
    orc = Oracle("f64")
    w = OracleWorld.from_scene(sc)
    b = w.download()
    col = sc.colliders.copy()
    n = sc.bodies_per_world
    one, two = [], []
    for i in range(n):
        one.append(i); two.append(-1)
        for j in range(n):
            if j != i:
                one.append(i); two.append(j)
    for step in range(130):
        orc.integrate(b, sc.dt)
        col.a["transform"] = orc.collider_derive(b.transform, col.offset)
        contacts, found = orc.narrowphase(col, sc.planes, b, one, two)
        if contacts.count:
            cs = Contacts(contacts.count, _abi.F64, **{k: contacts.valid(k) for k in ("body0", "body1", "friction", "restitution", "point", "normal", "penetration")})
            cs.count = contacts.count
            orc.resolve_contacts(8 * contacts.count, cs, b, sc.dt)
        w.step(sc.dt, 1)
        assert contacts.count == w.last_counts()[0][0]
    ref = w.download()
    for f in STATE_FIELDS + ("transform", "inverse_inertia_tensor_world", "last_frame_acceleration"):
        assert np.array_equal(getattr(b, f), getattr(ref, f)), f


def test_oracle_rl_loop_matches_golden():
    """The RL-style loop (AddVelocity / AddRotation on every body, then frames) through the oracle
    reproduces tests/golden/rl12_f64.npz (made by tests/golden/make_golden_rl.py)."""
    from golden import make_golden_rl as rl
    gold = load_golden("rl12_f64")
    out = rl.run_oracle()
    assert np.array_equal(out["counts"], gold["counts"]) and int(out["checksum"]) == int(gold["checksum"])
    for f in rl.OBS:
        assert np.array_equal(out[f], gold[f]), f
