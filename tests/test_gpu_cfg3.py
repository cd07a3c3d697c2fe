"""GPU: BASELINE config 3 at its REAL size — the 4 096-body pile (16^3 jittered lattice of cubes and
spheres, all-pairs-ordered schedule = 16.8 M ordered checks per frame in the reference's loop,
examples/cubedrop.go:42-67) stepped through the sort-based broadphase (K2) and the large-world
resolver, against the CPU oracle's O(n^2) loop: contact pair sequence, contact counts, the two
iteration counters (contact.go:233, :390) every frame, and bit-identical state.

The oracle costs ~0.1 s per frame while bodies fall and seconds per frame once thousands of
contacts exist (its resolver is O(iterations x contacts)), so the window is the first 60 frames
(11.7 k contacts, both loops at their 8*len(contacts) iteration cap from frame ~35 on) plus a
3-frame window restarted from a GPU snapshot of the piled state."""
import numpy as np
import pytest

from cubez_b200 import _abi, scenes
from golden_cases import STATE_FIELDS
from oracle_lib import OracleWorld

pytestmark = pytest.mark.gpu
ALL_FIELDS = STATE_FIELDS + ("transform", "inverse_inertia_tensor_world", "last_frame_acceleration")


def test_cfg3_pile_4096_first_60_frames_vs_oracle():
    from cubez_b200.api import BatchedWorld
    scene = scenes.pile(side=16)
    assert scene.bodies_per_world == 4096
    gpu = BatchedWorld.from_scene(scene, flags=_abi.WORLD_BROADPHASE)
    cpu = OracleWorld.from_scene(scene)
    peak = 0
    for s in range(60):
        gs, cs = gpu.step(scene.dt, 1), cpu.step(scene.dt, 1)
        for k in ("contacts", "pos_iterations", "vel_iterations", "max_contacts"):
            assert gs[k] == cs[k], (s, k, gs[k], cs[k])
        assert gpu.contact_pairs(0) == cpu.contact_pairs(0), s
        peak = max(peak, gs["contacts"])
        if s % 10 == 9:      # generated contact geometry, bit for bit
            gc, cc = gpu.contacts(0), cpu.contacts(0)
            for f in ("point", "normal", "penetration"):
                assert np.array_equal(gc.valid(f), cc.valid(f)), (s, f)
    assert peak > 8000                     # the pile did form: thousands of simultaneous contacts
    g, c = gpu.download(), cpu.download()
    for f in ALL_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    assert gpu.checksum_energy()[0] == cpu.checksum_energy()[0]
    gpu.close()


def test_cfg3_piled_snapshot_three_frames_vs_oracle():
    """Settled regime: the GPU runs the pile to frame 44 (thousands of contacts, iteration caps hit), its
    state is downloaded and handed to the oracle (snapshot / restore through the ABI's upload and
    download), and both step 3 more frames."""
    from cubez_b200.api import BatchedWorld
    scene = scenes.pile(side=16)
    gpu = BatchedWorld.from_scene(scene, flags=_abi.WORLD_BROADPHASE)
    gpu.step(scene.dt, 44)
    snap_b, snap_c = gpu.download(), gpu.download_colliders()
    cpu = OracleWorld.from_scene(scene)
    cpu.upload_bodies(snap_b, derive=False)
    cpu.upload_colliders(snap_c, derive=False)
    cpu.set_step_index(44)
    for s in range(3):
        gs, cs = gpu.step(scene.dt, 1), cpu.step(scene.dt, 1)
        assert gs["contacts"] >= 1000
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (s, k, gs[k], cs[k])
        assert gs["vel_iterations"] == 8 * gs["contacts"]      # the reference's cap (examples/cubedrop.go:73) is what ends the loop here
        assert gpu.contact_pairs(0) == cpu.contact_pairs(0), s
    g, c = gpu.download(), cpu.download()
    for f in ALL_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    gpu.close()
