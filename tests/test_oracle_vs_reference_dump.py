"""CPU: the hand-written restatement (oracle/cubez_oracle.hpp) against dumps printed by the REFERENCE ITSELF.

tests/golden/ref/*.txt were produced by tbogdala/cubez's own Go sources — rigidbody.go, colliders.go, contact.go,
math/*.go, read where they lie under /root/reference — translated statement by statement to C++ by oracle/go2cpp.py and
driven by the headless harness mains of go/harness/ (oracle/Makefile target `ref`, oracle/make_ref_golden.py).  The
translator knows Go syntax, not physics, so agreement here is not a shared reading of the source by one author: every
frame's contact count, (body, body) sequence, as-generated contact geometry and the raw bits of every body's state
must be identical — in float64 and, through the reference's own `type Real float32` switch (math/math.go:23, applied
by go2cpp.py --real=float32), in float32.  What remains outside: the Go compiler itself, and math.Pow (C pow() on both sides; the library
takes the Pow factors as host inputs)."""
import os
import subprocess

import pytest

import refdump
from oracle_lib import OracleWorld
from ref_cases import REF_CASES, REF_DIR, ref_text

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPU_CASES = [n for n in sorted(REF_CASES) if n != "pile4096_80"]     # the 4 096-body pile costs the oracle minutes: GPU suite only


@pytest.mark.parametrize("name", CPU_CASES)
def test_oracle_equals_reference_dump(name):
    make, frames = REF_CASES[name]
    scene = make()
    lines = refdump.run_dump(OracleWorld.from_scene(scene), scene, frames)
    assert refdump.first_difference(ref_text(name), lines) is None


def test_reference_own_unit_tests_pass_when_translated():
    """math/vector_test.go, quaternion_test.go, matrix_test.go (22 tests) run against the translated math package —
    among them TestVectorGoCopies, which pins Go's value-copy semantics of array types in the translation."""
    out = ref_text("math_tests")
    assert "22 tests, 0 failed" in out
    assert out.count("ok  ") == 22


HAVE_REF = os.path.exists("/root/reference/rigidbody.go")


@pytest.mark.skipif(not HAVE_REF, reason="the reference sources are only present in the build container")
@pytest.mark.parametrize("binary,args,name,lines", [("cubedrop_headless", ["120"], "cubedrop_600", 120), ("ballistic_headless", ["200"], "ballistic_600", 200),
                                                    ("pile_headless", ["60", "6"], "pile216_120", 60),
                                                    ("cubedrop_headless", ["100", "256", "0"], "batched256_600", 100)])
def test_committed_dumps_are_what_the_translated_reference_prints(binary, args, name, lines):
    """Regenerates a prefix of a committed dump from /root/reference (build container only)."""
    subprocess.run(["make", "-s", "ref"], cwd=os.path.join(ROOT, "oracle"), check=True)
    out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", binary)] + args, capture_output=True, text=True, check=True).stdout
    _, fresh, _ = refdump.parse(out)
    _, committed, _ = refdump.parse(ref_text(name))
    assert len(fresh) == lines
    for s in range(lines):
        assert fresh[s] == committed[s], s


@pytest.mark.skipif(not HAVE_REF, reason="the reference sources are only present in the build container")
def test_translated_reference_math_tests_run_here():
    subprocess.run(["make", "-s", "ref"], cwd=os.path.join(ROOT, "oracle"), check=True)
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "math_tests")], capture_output=True, text=True)
    assert r.returncode == 0 and "22 tests, 0 failed" in r.stdout
