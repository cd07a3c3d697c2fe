#!/usr/bin/env python
"""Generates tests/golden/ref/*.txt — dumps printed by the REFERENCE ITSELF (tbogdala/cubez's Go sources, translated
mechanically by oracle/go2cpp.py and driven by the harness mains of go/harness/, see oracle/Makefile target `ref`).
The vectors travel to the GPU box (where /root/reference does not exist); tests compare the CPU oracle and the CUDA
path with them bit for bit.  Re-run after changing a harness:   python oracle/make_ref_golden.py [name ...]
(pile4096_80 takes ~10 minutes: the reference's all-pairs loop and O(contacts^2) resolver)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref")
CASES = {   # name -> (binary, args)          scene builder of the same case: tests/ref_cases.py
    "cubedrop_600": ("cubedrop_headless", ["600"]),
    "cubedrop_staggered_260": ("cubedrop_headless", ["260", "0", "0", "120"]),
    "ballistic_600": ("ballistic_headless", ["600"]),
    "ballistic16_300": ("ballistic_headless", ["300", "16"]),
    "batched256_600": ("cubedrop_headless", ["600", "256", "0"]),
    "batched64_from1000_300": ("cubedrop_headless", ["300", "64", "1000"]),
    "pile27_150": ("pile_headless", ["150", "3"]),
    "pile216_120": ("pile_headless", ["120", "6"]),
    "pile4096_80": ("pile_headless", ["80", "16"]),
    "free65536_16": ("integrate_bench_headless", ["16", "65536"]),
    # the fuzz scene (scenes.random_worlds): <steps> <worlds> <bodies per world> <seed> <planes> [big 0|1] [material seed]
    "random8x8_120": ("random_headless", ["120", "8", "8", "11", "2"]),
    "random6x13_3planes_150": ("random_headless", ["150", "6", "13", "23", "3"]),
    "random4x24_100": ("random_headless", ["100", "4", "24", "5", "1"]),
    "random1x300_big_40": ("random_headless", ["40", "1", "300", "7", "3", "1"]),
    "random8x10_materials_150": ("random_headless", ["150", "8", "10", "31", "2", "0", "5"]),   # + per-pair surface materials painted by the host
    # the reference in SINGLE precision (`type Real float32`, math/math.go:23 — go2cpp.py --real=float32)
    "cubedrop_f32_600": ("cubedrop_headless_f32", ["600"]),
    "batched64_f32_300": ("cubedrop_headless_f32", ["300", "64", "0"]),
    "ballistic16_f32_300": ("ballistic_headless_f32", ["300", "16"]),
    "pile216_f32_100": ("pile_headless_f32", ["100", "6"]),
    "free65536_f32_16": ("integrate_bench_headless_f32", ["16", "65536"]),
    "random8x8_f32_120": ("random_headless_f32", ["120", "8", "8", "11", "2"]),
    "random8x10_materials_f32_150": ("random_headless_f32", ["150", "8", "10", "31", "2", "0", "5"]),
}

if __name__ == "__main__":
    subprocess.run(["make", "-s", "ref"], cwd=HERE, check=True)
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, (binary, args) in CASES.items():
        if only and name not in only:
            continue
        r = subprocess.run([os.path.join(HERE, "_ref", binary)] + args, capture_output=True, text=True, check=True)
        with open(os.path.join(OUT, name + ".txt"), "w") as f:
            f.write(r.stdout)
        print(name, len(r.stdout.splitlines()), "lines;", r.stderr.strip().splitlines()[-1])
    if not only or "math_tests" in only:
        r = subprocess.run([os.path.join(HERE, "_ref", "math_tests")], capture_output=True, text=True)
        with open(os.path.join(OUT, "math_tests.txt"), "w") as f:
            f.write(r.stdout)
        print("math_tests:", r.stdout.strip().splitlines()[-1])
