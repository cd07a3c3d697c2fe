"""CPU: the oracle under AddressSanitizer + UndefinedBehaviorSanitizer: the restatement is the checker
of everything else, so it must not rest on an out-of-bounds read or undefined behaviour.  A subprocess loads
oracle/_build/liboracle_f64_asan.so (oracle/Makefile target `asan`) with libasan preloaded and steps three scenes —
cubedrop (cube-cube SAT, sleep), ballistic (spheres, static body, explicit schedule, late spawns) and a small pile
(cube-sphere, multi-contact resolver) — against the committed reference dumps; any sanitizer report fails the test."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import refdump
from oracle_lib import OracleWorld
from ref_cases import REF_CASES, ref_text
for name, frames in (("cubedrop_600", 200), ("ballistic16_300", 150), ("pile27_150", 120), ("batched64_from1000_300", 60)):
    make, _ = REF_CASES[name]
    scene = make()
    lines = refdump.run_dump(OracleWorld.from_scene(scene), scene, frames)
    assert refdump.first_difference(ref_text(name), lines[:frames]) is None, name
print("SANITIZED-OK")
"""


def test_oracle_is_clean_under_asan_and_ubsan():
    subprocess.run(["make", "-s", "asan"], cwd=os.path.join(ROOT, "oracle"), check=True)
    libasan = subprocess.run(["/usr/bin/gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    libubsan = subprocess.run(["/usr/bin/gcc", "-print-file-name=libubsan.so"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ, CUBEZ_ORACLE_SUFFIX="_asan", LD_PRELOAD=f"{libasan} {libubsan}",
               ASAN_OPTIONS="detect_leaks=0:abort_on_error=0:halt_on_error=1", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], capture_output=True, text=True, env=env, timeout=900)
    assert "SANITIZED-OK" in r.stdout, r.stderr[-3000:]
    assert "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
    assert r.returncode == 0
