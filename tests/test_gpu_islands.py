"""GPU: contact islands of one large world (k_resolve_islands: one CTA per connected component of the contact graph).
Exact while the reference's worst-first loop converges inside its cap of 8*len(contacts) iterations per phase
(examples/cubedrop.go:73); when the cap would cut the loop the library detects it (sum of the islands' iterations
against the cap) and re-runs the frame's resolve on the single-CTA path.  Both branches are compared with the CPU
oracle's single global loop: contact sequence, iteration counts, bit-identical state."""
import numpy as np
import pytest

from cubez_b200 import _abi, scenes
from golden_cases import STATE_FIELDS
from oracle_lib import OracleWorld

pytestmark = pytest.mark.gpu


def _compare(scene, frames, every, monkeypatch, mode):
    from cubez_b200.api import BatchedWorld
    monkeypatch.setenv("CUBEZ_RESOLVE_ISLANDS", mode)
    gpu = BatchedWorld.from_scene(scene, flags=_abi.WORLD_BROADPHASE)
    cpu = OracleWorld.from_scene(scene)
    for s in range(0, frames, every):
        gs, cs = gpu.step(scene.dt, every), cpu.step(scene.dt, every)
        for k in ("contacts", "pos_iterations", "vel_iterations"):     # (max_contacts: the oracle reports the last frame's, the library the call's maximum)
            assert gs[k] == cs[k], (s, k, gs[k], cs[k])
        assert gpu.contact_pairs(0) == cpu.contact_pairs(0), s
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS + ("transform", "inverse_inertia_tensor_world"):
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    stats = gpu.island_stats()
    gpu.close()
    return stats


@pytest.mark.parametrize("mode", ["2", "1"])
def test_archipelago_of_separate_piles_resolves_per_island(monkeypatch, mode):
    """36 separate 8-body piles in one world: 36+ islands, no cap hit -> every frame with contacts is kept."""
    scene = scenes.archipelago(piles=6, side=2)
    frames, fallbacks = _compare(scene, 160, 1, monkeypatch, mode)
    assert frames > (100 if mode == "2" else 20)
    assert fallbacks <= frames // 4


def test_one_big_pile_hits_the_cap_and_falls_back_exactly(monkeypatch):
    """A single 512-body pile: one giant island, both loops end at the cap -> the detected fallback must reproduce the
    reference's cut-off loop (iteration counts equal the oracle's 8*len(contacts))."""
    scene = scenes.pile(side=8)
    frames, fallbacks = _compare(scene, 60, 5, monkeypatch, "2")
    assert frames > 0 and fallbacks > 0


def test_islands_off_is_the_same(monkeypatch):
    scene = scenes.archipelago(piles=3, side=2)
    assert _compare(scene, 120, 4, monkeypatch, "0") == (0, 0)
