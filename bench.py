#!/usr/bin/env python
"""bench.py — headline benchmark of the cubez B200 hot path (contract: see README/DESIGN.md).

Workload (BASELINE.json configs[3], "batched RL-style"): 65 536 independent, perturbed 8-cube
cubedrop worlds PER GPU (weak scaling; worlds never exchange data), 600-frame episodes at
dt = 1/60, float64.  World k starts at frame (k mod 600) of its episode and is reset to its
initial state when the episode ends, so every timed frame sees the same stationary mix of
free fall / impact / settling / sleeping worlds, whatever --steps is.  One "step" = one frame
of updateCallback (examples/cubedrop.go:69-75) for every world of the job.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torchrun, one rank per GPU; NCCL only for the end-of-run reduce)

Prints ONE JSON line on rank 0.  `value` = world-steps/s with state resident on the GPU;
`e2e` = the same through cz_world_step_host with pinned HOST buffers (H2D + D2H every step);
`roofline` = the dominant kernel of this workload (fused world step); `roofline_k1` = the
HBM-bound integrate+derive kernel on 16 Mi free bodies (cfg5); `cpu_baseline` = the CPU oracle
("port" of the Go loops; Go itself can not run in this image) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

EPISODE = 600
DT = 1.0 / 60.0
WORLDS_PER_GPU = int(os.environ.get("CUBEZ_BENCH_WORLDS", 65536))
BODIES_PER_WORLD = 8
WORKLOAD = "cfg4 batched RL-style: independent perturbed 8-cube cubedrop worlds, 600-frame episodes, phases staggered uniformly, reset at episode end"
K1_BODIES = 1 << 24
K1_BYTES_F64 = 531          # algorithmic bytes per awake body-step


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 7 for i in range(4) if s[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


CPU_SAMPLE_WORLDS = 2048   # bounded cpu_baseline sample: 2048 whole episodes, about 10 s of one core


def cpu_oracle_sample(n_worlds: int, n_threads: int):
    """World-steps/s of the CPU oracle on full 600-frame episodes of the first n_worlds worlds
    (the stationary population's average cost per frame equals the episode average)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cubez_b200 import scenes
    from oracle_lib import OracleWorld     # the checker; allowed here as the CPU baseline only
    sc = scenes.batched_cubedrop(n_worlds=n_worlds)
    w = OracleWorld.from_scene(sc)
    t0 = time.perf_counter()
    w.step(DT, EPISODE, n_threads=n_threads)
    dt = time.perf_counter() - t0
    w.close()
    return n_worlds * EPISODE / dt, dt


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores.
    Go can not be built here, so this is the line-by-line C++ restatement (oracle/), AoS
    structs and per-contact heap allocation kept, worlds partitioned over all host threads."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step_worlds = threads * 8
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    for _ in range(args.warmup):
        cpu_oracle_sample(per_step_worlds, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_sample(per_step_worlds, threads)
    el = time.perf_counter() - t0
    value = per_step_worlds * EPISODE * args.steps / el
    line = {
        "impl": "reference", "metric": "world_steps_per_s", "value": value, "unit": "world-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the arm's own config (same workload as --impl ours); what one step samples of it is in cpu_baseline.sample
        "config": {"workload": WORKLOAD, "worlds_per_gpu": args.worlds, "bodies_per_world": BODIES_PER_WORLD, "dt": DT,
                   "contact_capacity": int(os.environ.get("CUBEZ_BENCH_CONTACT_CAP", 64)),
                   "parallelism": f"worlds partitioned over {threads} host threads (the reference is single-threaded; no GPU on this arm)"},
        "cpu_baseline": {"value": value, "unit": "world-steps/s", "cores": threads, "kind": "port",
                         "sample": f"each step = {per_step_worlds} worlds x one whole {EPISODE}-frame episode (the same frame mix as the staggered phases) on {threads} threads (C++ restatement of the Go loops; no Go toolchain in this image)"},
        "e2e": {"value": value, "unit": "world-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "body_steps_per_s": value * BODIES_PER_WORLD,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--worlds", type=int, default=WORLDS_PER_GPU, help="worlds per GPU")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-k1", action="store_true", help="skip the cfg5 integrate roofline run")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world_size = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libcubezcuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from cubez_b200 import _abi, scenes
    from cubez_b200.api import BatchedWorld, Context
    from cubez_b200.sharding import reduce_run

    W = args.worlds
    first_world = rank * W                      # weak scaling: every rank owns W worlds with distinct global ids
    ctx = Context.get(local_rank, "f64")
    scene = scenes.batched_cubedrop(_abi.F64, n_worlds=W, first_world=first_world)
    contact_cap = int(os.environ.get("CUBEZ_BENCH_CONTACT_CAP", 64))
    world = BatchedWorld.from_scene(scene, device=local_rank, contacts_per_world=contact_cap, ctx=ctx)
    phase0 = ((first_world + np.arange(W, dtype=np.int64)) % EPISODE).astype(np.int32)
    world.set_episodes(EPISODE, phase0)
    world.step(DT, EPISODE, stats=True)         # pre-roll (setup, untimed): the population becomes stationary

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        world.synchronize()

    # ---- device-resident arm --------------------------------------------------------------
    world.step(DT, args.warmup, stats=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    st = world.step(DT, args.steps, stats=True)         # K frames, CUDA events on the launching stream inside
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = float(st["device_ms"])
    cks, energy = world.checksum_energy()
    red = reduce_run(cks, energy, {"contacts": st["contacts"], "pos_iterations": st["pos_iterations"],
                                   "vel_iterations": st["vel_iterations"], "launches": st["kernel_launches"]}, dev_ms, dev)
    max_ms = red["max_ms"]
    total_world_steps = W * world_size * args.steps
    value = total_world_steps / (max_ms * 1e-3)

    # ---- end-to-end arm: host buffers through cz_world_step_host ------------------------------
    host = world.download(out=ctx.pinned_bodies(W * BODIES_PER_WORLD))      # page-locked host arrays (cz_host_alloc)
    nb = W * BODIES_PER_WORLD
    h2d = nb * (28 * 8 + 2)                      # primary state the host may have edited (K1 read set)
    d2h = nb * (38 * 8 + 1)                      # everything the frame writes
    world.step_host(host, DT, 1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        world.step_host(host, DT, 1)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.e2e_steps
    if world_size > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = W * world_size / (e2e_ms * 1e-3)

    # ---- RL-style loop: worlds stay resident, actions in, observations out ------
    # every frame: batched AddVelocity for every body from pinned host memory (48 B/body H2D), one frame,
    # position + orientation + velocity + rotation back (104 B/body D2H)
    act = ctx.pinned_array((nb, 3))
    act[...] = np.random.default_rng(1 + rank).uniform(-1e-3, 1e-3, (nb, 3))
    obs = ctx.pinned_bodies(nb, fields=BatchedWorld.OBS_FIELDS)
    world.step_rl(act, None, obs, DT, 1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        world.step_rl(act, None, obs, DT, 1)
    barrier()
    rl_ms = (time.perf_counter() - t0) * 1e3 / args.e2e_steps
    if world_size > 1:
        t = torch.tensor([rl_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rl_ms = float(t.item())
    e2e_rl = {"value": W * world_size / (rl_ms * 1e-3), "unit": "world-steps/s", "h2d_bytes_per_step": nb * 24 * world_size,
              "d2h_bytes_per_step": nb * 13 * 8 * world_size, "ms_per_step": rl_ms,
              "api": "cz_world_step_rl: device-resident worlds; batched AddVelocity in, position/orientation/velocity/rotation out, pinned host arrays, 1 frame per call"}
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- roofline of the dominant kernel (fused world step) ---------------------------------
    peak, peak_kind = measured_peaks()
    # Algorithmic HBM bytes of ONE FRAME of the fused step in split mode (three launches per frame at this batch size:
    # A integrate + narrowphase + prepare | B position loop | C velocity loop):
    #   A stages the whole body + collider record (41 chunks of 16 B + 8 flag bytes in, 25 chunks + 1 flag out);
    #   B and C each stage the loops' body record (15 chunks + 8 flag bytes in, 20 chunks + 1 flag out);
    #   every contact is written once by A (as-generated 64 B, cold record 144 B, hot fields 24 B) and its hot
    #   fields cross twice more (B: 16 B in, 8 B out; C: 16 B in) — the loops' cold-record reads hit L2.
    contacts_per_frame = st["contacts"] / max(1, args.steps)
    fused_bytes = nb * ((41 * 16 + 8 + 25 * 16 + 1) + 2 * (15 * 16 + 8 + 20 * 16 + 1)) + contacts_per_frame * (232 + 24 + 16)
    frame_ms = dev_ms / max(1, args.steps)
    roofline = {"bound": "hbm", "kernel": "k_world_fused, split mode: phases A | B | C, one launch each per frame",
                "achieved": fused_bytes / (frame_ms * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "frac": fused_bytes / (frame_ms * 1e-3) / 1e9 / peak,
                "traffic": 1.979e9,   # dram__bytes_read+write of the three launches of one frame, ncu --set full (profiles/r01_fused_split_phases_v5.txt)
                "algorithmic_bytes_per_frame": fused_bytes, "peak_kind": peak_kind,
                "note": "not HBM-bound by design: a frame's working state is shared-memory resident inside each launch; the kernels are bound by FP64 dependency latency and L2 latency of the cold contact records at 12 warps per SM (see profiles/)",
                # what does bound it, from ncu --set full of one frame at 65 536 worlds (profiles/r01_fused_split_phases_v5.txt):
                "ncu": {"phases_ms": {"A integrate+narrowphase+prepare": 0.642, "B position loop": 0.382, "C velocity loop": 1.014},
                        "fp64_pipe_active_pct": {"A": 21.6, "B": 14.8, "C": 31.6}, "issue_active_pct": {"A": 34.4, "B": 28.4, "C": 39.0},
                        "active_lanes_of_32": {"A": 16.3, "B": 18.4, "C": 20.0}, "warps_per_sm": 12, "registers_per_thread": 168,
                        "top_stalls": ["long_scoreboard (cold contact records and staged state in L2)", "wait (fixed-latency FP64 dependency at 3 warps per scheduler)"]}}
    roofline_k1 = None
    if rank == 0 and not args.no_k1:
        world.close()
        ms, _ = ctx.bench_integrate(K1_BODIES, warmup=3, steps=50, dt=DT)
        ach = K1_BODIES * K1_BYTES_F64 / (ms * 1e-3) / 1e9
        roofline_k1 = {"bound": "hbm", "kernel": "k_integrate<false> (Integrate+CalculateDerivedData, cfg5: 16Mi free bodies f64)",
                       "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                       "traffic": 8.837e9,   # dram__bytes_read+write per launch, ncu --set full (profiles/r01_k1_integrate_f64.txt); algorithmic 8.909e9
                       "peak_kind": peak_kind,
                       "ms_per_launch": ms, "body_steps_per_s": K1_BODIES / (ms * 1e-3), "bytes_per_body": K1_BYTES_F64}

    roofline_k1_f32 = roofline_k2_f32 = None
    if rank == 0 and not args.no_k1:
        import ctypes as C
        # the float32 build (the reference's tunable Real): 267 B per body-step for K1, 116 n + 8 q for K2
        ctx32 = Context.get(local_rank, "f32")
        ms32, _ = ctx32.bench_integrate(K1_BODIES, warmup=3, steps=50, dt=DT)
        ach32 = K1_BODIES * 267 / (ms32 * 1e-3) / 1e9
        roofline_k1_f32 = {"bound": "hbm", "kernel": "k_integrate<false>, float32 build", "achieved": ach32, "peak": peak, "unit": "GB/s",
                           "frac": ach32 / peak, "traffic": None, "peak_kind": peak_kind, "ms_per_launch": ms32,
                           "body_steps_per_s": K1_BODIES / (ms32 * 1e-3), "bytes_per_body": 267}
        ms3, pairs3, sms3 = C.c_float(), C.c_int64(), C.c_float()
        ctx32.check(ctx32.lib.cz_bench_broadphase(ctx32.h, K1_BODIES, 7, 0.05, 2, 5, C.byref(ms3), C.byref(pairs3), C.byref(sms3)))
        alg3 = 116 * K1_BODIES + 8 * pairs3.value
        ach3 = alg3 / (ms3.value * 1e-3) / 1e9
        roofline_k2_f32 = {"bound": "hbm", "kernel": "K2 sort-based broadphase, float32 bounds", "achieved": ach3, "peak": peak, "unit": "GB/s",
                           "frac": ach3 / peak, "traffic": None, "peak_kind": peak_kind, "ms_per_frame": ms3.value, "candidate_pairs": pairs3.value}

    roofline_k2 = None
    if rank == 0 and not args.no_k1:
        import ctypes as C
        ms2, pairs, sms = C.c_float(), C.c_int64(), C.c_float()
        ctx.check(ctx.lib.cz_bench_broadphase(ctx.h, K1_BODIES, 7, 0.05, 2, 5, C.byref(ms2), C.byref(pairs), C.byref(sms)))
        alg = 148 * K1_BODIES + 8 * pairs.value          # 148 n + 8 q bytes (f64 bounds, 4-pass radix sort)
        ach2 = alg / (ms2.value * 1e-3) / 1e9
        roofline_k2 = {"bound": "hbm", "kernel": "K2 sort-based broadphase (one-pass counting sort by cell key: count, scan, resolve, place; neighbour sweep): 16Mi unit spheres, 5% fill",
                       "achieved": ach2, "peak": peak, "unit": "GB/s", "frac": ach2 / peak,
                       "traffic": 3.34e9,   # dram read+write of memset, count, scan, resolve, place, sweep per frame (ncu --set full, profiles/r01_k2_counting_sort.txt); algorithmic 2.51e9
                       "peak_kind": peak_kind,
                       "ms_per_frame": ms2.value, "lsd_radix_sort_alone_ms": sms.value, "candidate_pairs": pairs.value}

    if rank == 0:
        cpu_value, cpu_s = cpu_oracle_sample(CPU_SAMPLE_WORLDS, 1)   # ~10 s of one host core
        line = {
            "metric": "world_steps_per_s", "value": value, "unit": "world-steps/s", "n_gpus": world_size, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "worlds_per_gpu": W, "bodies_per_world": BODIES_PER_WORLD, "dt": DT, "contact_capacity": contact_cap,
                       "parallelism": f"worlds sharded over {world_size} GPU(s), no per-step collective",
                       "l2": "state (%.0f MB/GPU) larger than L2; frames of one call run from shared memory" % (nb * 84 * 16 / 1e6)},
            "body_steps_per_s": value * BODIES_PER_WORLD,
            "e2e": {"value": e2e_value, "unit": "world-steps/s", "h2d_bytes_per_step": h2d * world_size, "d2h_bytes_per_step": d2h * world_size,
                    "ms_per_step": e2e_ms, "api": "cz_world_step_host: pinned host arrays (full body state in, full body state out), 1 frame per call, 6-chunk H2D | pack+step+unpack | D2H pipeline, 3 compute streams"},
            "e2e_rl": e2e_rl,
            "gpu_launches": int(red["counters"]["launches"]),
            "clocks": sampler.summary(),
            "roofline": roofline,
            "roofline_k1": roofline_k1,
            "roofline_k2": roofline_k2,
            "roofline_k1_f32": roofline_k1_f32,
            "roofline_k2_f32": roofline_k2_f32,
            "cpu_baseline": {"value": cpu_value, "unit": "world-steps/s", "cores": 1, "kind": "port",
                             "sample": f"{CPU_SAMPLE_WORLDS} worlds x {EPISODE} frames, 1 thread, {cpu_s:.1f} s (C++ restatement of the Go loops)"},
            "checksum": hex(red["checksum"]), "energy": red["energy"],
            "contacts_per_world_step": red["counters"]["contacts"] / total_world_steps,
            "vel_iterations_per_world_step": red["counters"]["vel_iterations"] / total_world_steps,
            "pos_iterations_per_world_step": red["counters"]["pos_iterations"] / total_world_steps,
            "wall_ms_timed_region": wall_ms,
        }
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
