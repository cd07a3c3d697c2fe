"""Scene builders of the reference dumps under tests/golden/ref/ (written by oracle/make_ref_golden.py from the
reference's own Go sources, mechanically translated — see oracle/go2cpp.py).  name -> (scene factory, frames)."""
import os

from cubez_b200 import scenes
from cubez_b200._abi import F32

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref")
REF_CASES = {
    "cubedrop_600": (lambda: scenes.cubedrop(), 600),
    "cubedrop_staggered_260": (lambda: scenes.cubedrop(second_fire_step=120), 260),
    "ballistic_600": (lambda: scenes.ballistic(), 600),
    "ballistic16_300": (lambda: scenes.ballistic(n_bullets=16), 300),
    "batched256_600": (lambda: scenes.batched_cubedrop(n_worlds=256), 600),
    "batched64_from1000_300": (lambda: scenes.batched_cubedrop(n_worlds=64, first_world=1000), 300),
    "pile27_150": (lambda: scenes.pile(side=3), 150),
    "pile216_120": (lambda: scenes.pile(side=6), 120),
    "pile4096_80": (lambda: scenes.pile(side=16), 80),
    "free65536_16": (lambda: scenes.free_bodies(n=65536), 16),
    # the fuzz scene through the reference's own sources (go/harness/random_headless.go): cubes, spheres and collider-less
    # bodies of random sizes and masses, rotated and shifted collider Offsets, one to three planes, sleeping and
    # late-activated bodies, per-body damping and gravity
    "random8x8_120": (lambda: scenes.random_worlds(n_worlds=8, bodies_per_world=8, seed=11, n_planes=2), 120),
    "random6x13_3planes_150": (lambda: scenes.random_worlds(n_worlds=6, bodies_per_world=13, seed=23, n_planes=3), 150),
    "random4x24_100": (lambda: scenes.random_worlds(n_worlds=4, bodies_per_world=24, seed=5, n_planes=1), 100),
    "random1x300_big_40": (lambda: scenes.random_worlds(n_worlds=1, bodies_per_world=300, seed=7, n_planes=3, extent=5.0, height=9.0), 40),
    # ... and with per-pair surface materials: the harness paints Friction / Restitution onto the contacts each check
    # appends, as a host of the reference would; the library does it through cz_world_set_materials
    "random8x10_materials_150": (lambda: scenes.with_materials(scenes.random_worlds(n_worlds=8, bodies_per_world=10, seed=31, n_planes=2), seed=5), 150),
    # the reference in SINGLE precision: `type Real float64` -> float32 (math/math.go:23), the reference's own switch, applied by
    # the translator to the text it reads (oracle/go2cpp.py --real=float32).  The float32 library and oracle must print the same.
    "cubedrop_f32_600": (lambda: scenes.cubedrop(F32), 600),
    "batched64_f32_300": (lambda: scenes.batched_cubedrop(F32, n_worlds=64), 300),
    "ballistic16_f32_300": (lambda: scenes.ballistic(F32, n_bullets=16), 300),
    "pile216_f32_100": (lambda: scenes.pile(F32, side=6), 100),
    "free65536_f32_16": (lambda: scenes.free_bodies(F32, n=65536), 16),
    "random8x8_f32_120": (lambda: scenes.random_worlds(F32, n_worlds=8, bodies_per_world=8, seed=11, n_planes=2), 120),
    "random8x10_materials_f32_150": (lambda: scenes.with_materials(scenes.random_worlds(F32, n_worlds=8, bodies_per_world=10, seed=31, n_planes=2), seed=5), 150),
}


def ref_text(name: str) -> str:
    with open(os.path.join(REF_DIR, name + ".txt")) as f:
        return f.read()
