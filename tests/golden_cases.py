"""Shared table of golden cases (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import os

import numpy as np

from golden.make_golden import CASES  # noqa: F401

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STATE_FIELDS = ("position", "orientation", "velocity", "rotation", "motion", "is_awake")


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def check_against_golden(world, scene, gold, n_steps, pair_hash, worlds_checked=None):
    """Step `world` and compare per-step contact counts / iteration counts / contact-pair
    sequences and the final state with the golden record.  Bit-exact."""
    nw = scene.n_worlds
    for s in range(n_steps):
        world.step(scene.dt, 1)
        c, p, v = world.last_counts()
        assert np.array_equal(c, gold["counts"][s]), f"contact counts differ at step {s}: {c} vs {gold['counts'][s]}"
        assert np.array_equal(p, gold["pos_iters"][s]), f"position iterations differ at step {s}"
        assert np.array_equal(v, gold["vel_iters"][s]), f"velocity iterations differ at step {s}"
        for k in (range(nw) if worlds_checked is None else worlds_checked):
            assert pair_hash(world.contact_pairs(k)) == int(gold["pair_hash"][s, k]), f"contact pair sequence differs at step {s} world {k}"
    b = world.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(b, f), gold[f]), f"final {f} differs from golden"
    cks, en = world.checksum_energy()
    assert cks == int(gold["checksum"])
    assert abs(en - float(gold["energy"])) <= 1e-9 * max(1.0, abs(float(gold["energy"])))
