#!/usr/bin/env python
"""bench.py — headline benchmark of the cubez B200 hot path (contract: DESIGN.md section 5).

Workload of the headline line (BASELINE.json configs[3], "batched RL-style"): independent, perturbed 8-cube cubedrop
worlds, 600-frame episodes at dt = 1/60, float64, 65 536 worlds PER GPU (weak scaling; worlds never exchange data).
World k starts at frame (k mod 600) of its episode and is reset when the episode ends, so every timed frame sees the
same stationary mix of free fall / impact / settling / sleeping worlds, whatever --steps is.  One "step" = one frame
of updateCallback (examples/cubedrop.go:69-75) for every world of the job.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torchrun, one rank per GPU; NCCL only for the end-of-run reduce)

Prints ONE JSON line on rank 0:
  value        world-steps/s, state resident on the GPU (CUDA events on the launching stream, max over ranks)
  e2e          the same through cz_world_step_host with pinned HOST buffers (H2D + D2H every step)
  e2e_rl       the RL loop: actions in, observations out, worlds resident (cz_world_step_rl)
  strong       BASELINE config 4 as defined: 65 536 worlds IN TOTAL sharded over the N ranks, 600 frames from t = 0;
               the reduced checksum must equal the CPU oracle's (tests/golden/cfg4_checksum.json) for every N
  roofline     dominant kernel of the headline workload (fused world step) against HBM, plus roofline_fp64 (what bounds it)
  roofline_k1 / roofline_k2 (+ _f32)   the HBM-bound kernels: integrate (cfg5, 16 Mi bodies), sort-based broadphase (16 Mi spheres)
  cfg1 / cfg2 / cfg3 / cfg5            the other BASELINE configs: GPU time per frame next to the CPU restatement on ONE host thread
  cpu_baseline the CPU oracle ("port" of the Go loops; Go itself cannot run in this image) on a bounded sample of the headline workload
"""
from __future__ import annotations

import argparse
import hashlib
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
STRONG_WORLDS = 65536           # BASELINE config 4: 65 536 worlds in total
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


def csrc_sha() -> str:
    """Hash of the CUDA sources: ncu-derived numbers in profiles/ncu_counters.json are only valid for the code they were
    captured from (they are emitted as null when this hash differs)."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "cubez_b200", "csrc")
    for f in sorted(os.listdir(d)):
        with open(os.path.join(d, f), "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_counters():
    """profiles/ncu_counters.json: {"csrc_sha": ..., "fused_frame_dram_bytes": ..., "fused_fp64_ops_per_world_step": ...,
    "k1_dram_bytes": ..., "k2_dram_bytes": ...} written by tools/ncu_counters.py from ncu captures of THIS code."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_counters.json")) as f:
            d = json.load(f)
        return d if d.get("csrc_sha") == csrc_sha() else {"stale": d.get("csrc_sha")}
    except Exception:
        return {}


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


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib    # the checker; allowed here as the CPU baseline only
    return oracle_lib


def cpu_oracle_sample(n_worlds: int, n_threads: int):
    """World-steps/s of the CPU oracle on full 600-frame episodes of the first n_worlds worlds
    (the stationary population's average cost per frame equals the episode average)."""
    from cubez_b200 import scenes
    sc = scenes.batched_cubedrop(n_worlds=n_worlds)
    w = _oracle().OracleWorld.from_scene(sc)
    t0 = time.perf_counter()
    w.step(DT, EPISODE, n_threads=n_threads)
    dt = time.perf_counter() - t0
    w.close()
    return n_worlds * EPISODE / dt, dt


def physical_cores() -> int:
    try:
        import psutil
        return psutil.cpu_count(logical=False) or (os.cpu_count() or 1)
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  Go cannot be built here, so
    this is the line-by-line C++ restatement (oracle/; pinned bit for bit against the mechanically translated Go
    sources, tests/golden/ref/), AoS structs and per-contact heap allocation kept, worlds handed out dynamically to
    all host threads."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step_worlds = threads * 16
    for _ in range(args.warmup):
        cpu_oracle_sample(per_step_worlds, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_sample(per_step_worlds, threads)
    el = time.perf_counter() - t0
    value = per_step_worlds * EPISODE * args.steps / el
    one, _ = cpu_oracle_sample(256, 1)
    line = {
        "impl": "reference", "metric": "world_steps_per_s", "value": value, "unit": "world-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the arm's own config (same workload as --impl ours); what one step samples of it is in cpu_baseline.sample
        "config": {"workload": WORKLOAD, "worlds_per_gpu": args.worlds, "bodies_per_world": BODIES_PER_WORLD, "dt": DT,
                   "contact_capacity": int(os.environ.get("CUBEZ_BENCH_CONTACT_CAP", 64)),
                   "parallelism": f"worlds over {threads} host threads (the reference itself is single-threaded; no GPU on this arm)"},
        "cpu_baseline": {"value": value, "unit": "world-steps/s", "cores": threads, "kind": "port",
                         "physical_cores": physical_cores(), "one_thread_value": one,
                         "thread_scaling_efficiency": value / (one * threads),
                         "sample": f"each step = {per_step_worlds} worlds x one whole {EPISODE}-frame episode (the same frame mix as the staggered phases) on {threads} hardware threads ({physical_cores()} physical cores); C++ restatement of the Go loops (no Go toolchain in this image); the reference as shipped is single-threaded: one thread gives {one:.0f} world-steps/s"},
        "e2e": {"value": value, "unit": "world-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "body_steps_per_s": value * BODIES_PER_WORLD,
    }
    print(json.dumps(line), flush=True)


# ---- the other BASELINE configs: GPU next to ONE host thread of the CPU restatement (rank 0, N = 1 only) ----------
def time_cpu(scene, frames, repeat=1, upload=None, first_step=0):
    ol = _oracle()
    best = None
    for _ in range(repeat):
        w = ol.OracleWorld.from_scene(scene)
        if upload is not None:
            w.upload_bodies(upload[0], derive=False)
            w.upload_colliders(upload[1], derive=False)
            w.set_step_index(first_step)
        t0 = time.perf_counter()
        st = w.step(scene.dt, frames)
        el = time.perf_counter() - t0
        w.close()
        best = el if best is None else min(best, el)
    return best, st


def bench_small_configs(ctx):
    from cubez_b200 import _abi, scenes
    from cubez_b200.api import BatchedWorld
    out = {}
    for name, scene, note in (("cfg1", scenes.cubedrop(), "cubedrop: 8 cubes, 600 frames, one world (fused persistent kernel, one launch for all frames)"),
                              ("cfg2", scenes.ballistic(), "ballistic: cube + backboard + 64 bullets, 600 frames, explicit 4 226-check schedule (multi-kernel path)")):
        g = BatchedWorld.from_scene(scene, ctx=ctx)
        g.step(scene.dt, 5)
        g.close()
        g = BatchedWorld.from_scene(scene, ctx=ctx)
        t0 = time.perf_counter()
        st = g.step(scene.dt, 600)
        wall = time.perf_counter() - t0
        cks = g.checksum_energy()[0]
        g.close()
        cpu_s, cst = time_cpu(scene, 600, repeat=5)
        assert (st["contacts"], st["pos_iterations"], st["vel_iterations"]) == (cst["contacts"], cst["pos_iterations"], cst["vel_iterations"]), name
        out[name] = {"workload": note, "frames": 600, "gpu_ms_per_frame": st["device_ms"] / 600, "gpu_wall_ms_per_frame": wall * 1e3 / 600,
                     "gpu_world_steps_per_s": 600 / (st["device_ms"] * 1e-3), "gpu_launches": st["kernel_launches"],
                     "cpu_1thread_ms_per_frame": cpu_s * 1e3 / 600, "cpu_1thread_world_steps_per_s": 600 / cpu_s,
                     "gpu_over_cpu": cpu_s * 1e3 / st["device_ms"], "checksum": hex(cks),
                     "note": "ONE small world is a single serial chain: the GPU runs it on one warp and is slower than one host core here; the GPU path exists for batches (cfg4) and large worlds (cfg3)"}
    # cfg3: the 4 096-body pile through the broadphase; falling window from t = 0, settled window from the committed snapshot
    scene = scenes.pile(side=16)
    g = BatchedWorld.from_scene(scene, flags=_abi.WORLD_BROADPHASE, ctx=ctx)
    st = g.step(scene.dt, 10)
    cpu_fall, _ = time_cpu(scene, 10)
    snap = np.load(os.path.join(ROOT, "tests", "golden", "pile4096_f100.npz"))
    b, c = scene.bodies, scene.colliders
    for f in ("position", "orientation", "velocity", "rotation", "motion", "is_awake", "transform", "inverse_inertia_tensor_world", "last_frame_acceleration"):
        getattr(b, f)[...] = snap[f]
    c.transform[...] = snap["collider_transform"]
    g.upload_bodies(b, derive=False)
    g.upload_colliders(c, derive=False)
    g.set_step_index(int(snap["frame"]))
    g.step(scene.dt, 1)
    st2 = g.step(scene.dt, 3)
    g.close()
    cpu_settled, cst2 = time_cpu(scene, 1, upload=(b, c), first_step=int(snap["frame"]))
    out["cfg3"] = {"workload": "pile: 4 096 mixed cubes and spheres on a plane, all-pairs-ordered schedule (16.8 M ordered checks per frame in the reference), one world on one GPU through the sort-based broadphase and the single-CTA worst-first resolver",
                   "falling_window": {"frames": "0-9", "gpu_ms_per_frame": st["device_ms"] / 10, "cpu_1thread_ms_per_frame": cpu_fall * 1e3 / 10,
                                      "gpu_body_steps_per_s": 4096 * 10 / (st["device_ms"] * 1e-3), "gpu_over_cpu": cpu_fall * 1e3 / st["device_ms"]},
                   "settled_window": {"frames": "101-103 (GPU), 100 (CPU) from tests/golden/pile4096_f100.npz", "contacts_per_frame": st2["contacts"] / 3,
                                      "resolver_iterations_per_frame": (st2["pos_iterations"] + st2["vel_iterations"]) / 3,
                                      "gpu_ms_per_frame": st2["device_ms"] / 3, "cpu_1thread_ms_per_frame": cpu_settled * 1e3,
                                      "gpu_us_per_resolver_iteration": st2["device_ms"] * 1e3 / max(1, st2["pos_iterations"] + st2["vel_iterations"]),
                                      "gpu_body_steps_per_s": 4096 * 3 / (st2["device_ms"] * 1e-3), "gpu_over_cpu": cpu_settled * 1e3 / (st2["device_ms"] / 3),
                                      "cpu_contacts": cst2["contacts"],
                                      "note": "both loops end at the reference's iteration cap 8*len(contacts) (examples/cubedrop.go:73): ~120 k strictly sequential iterations per frame; the time is their latency, not the broadphase"}}
    return out


def bench_cfg5_cpu():
    """One host thread of the CPU restatement on a 1 Mi-body sample of cfg5 (AoS structs as in the reference)."""
    from cubez_b200 import scenes
    n, steps = 1 << 20, 4
    sc = scenes.free_bodies(n=n)
    o = _oracle().Oracle("f64")
    w = _oracle().OracleWorld.from_scene(sc)       # derived data
    b = w.download()
    w.close()
    sec = o.bench_integrate(b, sc.dt, steps, 1)
    return {"value": n * steps / sec, "unit": "body-steps/s", "cores": 1, "kind": "port", "sample": f"{n} bodies x {steps} steps, 1 thread, {sec:.2f} s (AoS restatement of RigidBody.Integrate)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--worlds", type=int, default=WORLDS_PER_GPU, help="worlds per GPU (weak-scaling arm)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-k1", action="store_true", help="skip the K1 / K2 roofline microbenchmarks")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg1 / cfg2 / cfg3 / cfg5 sub-benchmarks and their CPU baselines")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling arm")
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
    from cubez_b200.sharding import pin_to_gpu_numa_node, reduce_run, shard_range

    numa = pin_to_gpu_numa_node(local_rank)     # host threads and pinned buffers next to this rank's GPU
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

    def max_over_ranks(x):
        if world_size > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

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
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.e2e_steps)
    e2e_value = W * world_size / (e2e_ms * 1e-3)

    # ---- RL-style loop: worlds stay resident, actions in, observations out ------
    # every frame: batched AddVelocity for every body from pinned host memory (24 B/body H2D), one frame,
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
    rl_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.e2e_steps)
    e2e_rl = {"value": W * world_size / (rl_ms * 1e-3), "unit": "world-steps/s", "h2d_bytes_per_step": nb * 24 * world_size,
              "d2h_bytes_per_step": nb * 13 * 8 * world_size, "ms_per_step": rl_ms,
              "api": "cz_world_step_rl: device-resident worlds; batched AddVelocity in, position/orientation/velocity/rotation out, pinned host arrays, 1 frame per call"}
    # the same loop pipelined (cz_world_step_rl_async, two steps in flight) with float32 observations converted on the device
    acts2 = [act, ctx.pinned_array((nb, 3))]
    acts2[1][...] = act
    obs32 = [{k: ctx.pinned_array((nb, c), dtype=np.float32) for k, c in (("position", 3), ("orientation", 4), ("velocity", 3), ("rotation", 3))} for _ in range(2)]
    def rl_pipelined(n):
        tickets = []
        for f in range(n):
            if f >= 2:
                world.rl_wait(tickets[f - 2], stats=False)
            tickets.append(world.step_rl_async(acts2[f & 1], None, None, obs32[f & 1], DT, 1))
        for t in tickets[-2:]:
            world.rl_wait(t, stats=False)
    rl_pipelined(4)
    barrier()
    t0 = time.perf_counter()
    rl_pipelined(2 * args.e2e_steps)
    barrier()
    rlp_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / (2 * args.e2e_steps))
    e2e_rl_pipelined = {"value": W * world_size / (rlp_ms * 1e-3), "unit": "world-steps/s", "h2d_bytes_per_step": nb * 24 * world_size,
                        "d2h_bytes_per_step": nb * 13 * 4 * world_size, "ms_per_step": rlp_ms,
                        "api": "cz_world_step_rl_async + cz_world_rl_wait: two steps in flight on alternating pinned buffers, batched AddVelocity in, float32 position/orientation/velocity/rotation out"}
    sampler.stop_flag = True
    sampler.join(timeout=2)
    world.close()

    # ---- strong scaling: BASELINE config 4 as defined — 65 536 worlds IN TOTAL, sharded contiguously over the ranks ----
    strong = None
    if not args.no_strong:
        lo, n_s = shard_range(STRONG_WORLDS, rank, world_size)
        hi = lo + n_s
        sc_s = scenes.batched_cubedrop(_abi.F64, n_worlds=n_s, first_world=lo)
        ws = BatchedWorld.from_scene(sc_s, device=local_rank, contacts_per_world=contact_cap, ctx=ctx)
        ws.step(DT, 3, stats=True)                # warm the kernels (then restart from t = 0)
        ws.close()
        ws = BatchedWorld.from_scene(sc_s, device=local_rank, contacts_per_world=contact_cap, ctx=ctx)
        world = ws
        barrier()
        sst = ws.step(DT, EPISODE, stats=True)
        barrier()
        scks, sen = ws.checksum_energy()
        sred = reduce_run(scks, sen, {"contacts": sst["contacts"], "pos_iterations": sst["pos_iterations"], "vel_iterations": sst["vel_iterations"]},
                          float(sst["device_ms"]), dev)
        ws.close()
        n1_ms = None
        if rank == 0 and world_size > 1:          # the same job on ONE GPU of this box, for the efficiency (the other ranks wait)
            sc1 = scenes.batched_cubedrop(_abi.F64, n_worlds=STRONG_WORLDS)
            w1 = BatchedWorld.from_scene(sc1, device=local_rank, contacts_per_world=contact_cap, ctx=ctx)
            n1_ms = float(w1.step(DT, EPISODE, stats=True)["device_ms"])
            w1.close()
        if world_size > 1:
            dist.barrier()
        expected = None
        try:
            with open(os.path.join(ROOT, "tests", "golden", "cfg4_checksum.json")) as f:
                expected = json.load(f)
        except Exception:
            pass
        s_value = STRONG_WORLDS * EPISODE / (sred["max_ms"] * 1e-3)
        strong = {"workload": "BASELINE cfg4 as defined: 65 536 perturbed cubedrop-8 worlds in total, 600 frames from t = 0, worlds sharded contiguously over the ranks, no per-step collective",
                  "worlds_total": STRONG_WORLDS, "worlds_per_rank": hi - lo, "frames": EPISODE, "value": s_value, "unit": "world-steps/s",
                  "ms": sred["max_ms"], "checksum": hex(sred["checksum"]), "energy": sred["energy"],
                  "checksum_expected": expected["checksum"] if expected else None,
                  "checksum_equal": (hex(sred["checksum"]) == expected["checksum"]) if expected else None,
                  "counters_equal": ((sred["counters"]["contacts"], sred["counters"]["pos_iterations"], sred["counters"]["vel_iterations"])
                                     == (expected["contacts"], expected["pos_iterations"], expected["vel_iterations"])) if expected else None,
                  "checksum_source": "CPU oracle over all 65 536 worlds (tests/golden/make_cfg4_checksum.py)",
                  "one_gpu_value_same_box": None if n1_ms is None else STRONG_WORLDS * EPISODE / (n1_ms * 1e-3),
                  "efficiency": 1.0 if world_size == 1 else (None if n1_ms is None else (n1_ms / sred["max_ms"]) / world_size)}

    # ---- roofline of the dominant kernel (fused world step) ---------------------------------
    peak, peak_kind = measured_peaks()
    counters = ncu_counters()
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
                "traffic": counters.get("fused_frame_dram_bytes"),   # dram__bytes_read+write of one frame's launches (ncu --set full), null when the capture is stale
                "algorithmic_bytes_per_frame": fused_bytes, "peak_kind": peak_kind,
                "note": "not HBM-bound by design: a frame's working state is shared-memory resident inside each launch; what bounds it is in roofline_fp64"}
    roofline_fp64 = None
    roofline_k1 = roofline_k1_f32 = roofline_k2 = roofline_k2_f32 = None
    configs = cfg5_cpu = None
    if rank == 0 and not args.no_k1:
        import ctypes as C
        rate = C.c_double()
        ctx.check(ctx.lib.cz_bench_fp64_rate(ctx.h, C.byref(rate)))
        ops = counters.get("fused_fp64_ops_per_world_step")
        roofline_fp64 = {"bound": "fp64", "kernel": "k_world_fused phases A | B | C", "peak": rate.value / 1e12, "peak_kind": "measured here (cz_bench_fp64_rate: independent DMUL/DADD chains, every SM busy)",
                         "unit": "T thread-level FP64 instructions/s", "ops_per_world_step": ops,
                         "achieved": None if ops is None else ops * (value / world_size) / 1e12,
                         "frac": None if ops is None else ops * (value / world_size) / rate.value,
                         "note": "ops_per_world_step = executed DADD + DMUL + DFMA thread instructions of one frame's three launches / worlds (ncu, profiles/ncu_counters.json; null when that capture is stale). The loops execute the winner's scalar resolve redundantly in the 8 lanes of a world's group, so `achieved` counts issued work; the dependent chain, not the pipe, is the limiter (profiles/)"}
        ms, _ = ctx.bench_integrate(K1_BODIES, warmup=3, steps=50, dt=DT)
        ach = K1_BODIES * K1_BYTES_F64 / (ms * 1e-3) / 1e9
        roofline_k1 = {"bound": "hbm", "kernel": "k_integrate<false> (Integrate+CalculateDerivedData, cfg5: 16Mi free bodies f64)",
                       "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                       "traffic": counters.get("k1_dram_bytes"), "peak_kind": peak_kind,
                       "ms_per_launch": ms, "body_steps_per_s": K1_BODIES / (ms * 1e-3), "bytes_per_body": K1_BYTES_F64}
        ms2, pairs, sms = C.c_float(), C.c_int64(), C.c_float()
        ctx.check(ctx.lib.cz_bench_broadphase(ctx.h, K1_BODIES, 7, 0.05, 2, 5, C.byref(ms2), C.byref(pairs), C.byref(sms)))
        alg = 148 * K1_BODIES + 8 * pairs.value          # 148 n + 8 q bytes (f64 bounds, 4-pass radix sort)
        ach2 = alg / (ms2.value * 1e-3) / 1e9
        roofline_k2 = {"bound": "hbm", "kernel": "K2 sort-based broadphase (one-pass counting sort by cell key: count, scan, resolve, place; neighbour sweep): 16Mi unit spheres, 5% fill",
                       "achieved": ach2, "peak": peak, "unit": "GB/s", "frac": ach2 / peak,
                       "traffic": counters.get("k2_dram_bytes"), "peak_kind": peak_kind,
                       "ms_per_frame": ms2.value, "lsd_radix_sort_alone_ms": sms.value, "candidate_pairs": pairs.value}
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
    if rank == 0 and world_size == 1 and not args.no_configs:
        configs = bench_small_configs(ctx)
        cfg5_cpu = bench_cfg5_cpu()

    if rank == 0:
        cpu_value, cpu_s = cpu_oracle_sample(CPU_SAMPLE_WORLDS, 1) if world_size == 1 else (None, None)   # ~10 s of one host core
        line = {
            "metric": "world_steps_per_s", "value": value, "unit": "world-steps/s", "n_gpus": world_size, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "worlds_per_gpu": W, "bodies_per_world": BODIES_PER_WORLD, "dt": DT, "contact_capacity": contact_cap,
                       "parallelism": f"worlds sharded over {world_size} GPU(s), no per-step collective",
                       "host_numa": numa,
                       "l2": "state (%.0f MB/GPU) larger than L2; frames of one call run from shared memory" % (nb * 84 * 16 / 1e6)},
            "body_steps_per_s": value * BODIES_PER_WORLD,
            "e2e": {"value": e2e_value, "unit": "world-steps/s", "h2d_bytes_per_step": h2d * world_size, "d2h_bytes_per_step": d2h * world_size,
                    "ms_per_step": e2e_ms, "api": "cz_world_step_host: pinned host arrays (full body state in, full body state out), 1 frame per call, 6-chunk H2D | pack+step+unpack | D2H pipeline, 3 compute streams"},
            "e2e_rl": e2e_rl,
            "e2e_rl_pipelined": e2e_rl_pipelined,
            "gpu_launches": int(red["counters"]["launches"]),
            "clocks": sampler.summary(),
            "strong": strong,
            "roofline": roofline,
            "roofline_fp64": roofline_fp64,
            "cpu_baseline": None if cpu_value is None else {"value": cpu_value, "unit": "world-steps/s", "cores": 1, "kind": "port",
                                                            "sample": f"{CPU_SAMPLE_WORLDS} worlds x {EPISODE} frames, 1 thread, {cpu_s:.1f} s (C++ restatement of the Go loops, pinned bit for bit against the mechanically translated Go sources)"},
            "roofline_k1": roofline_k1,
            "roofline_k2": roofline_k2,
            "roofline_k1_f32": roofline_k1_f32,
            "roofline_k2_f32": roofline_k2_f32,
            "cfg1": configs["cfg1"] if configs else None,
            "cfg2": configs["cfg2"] if configs else None,
            "cfg3": configs["cfg3"] if configs else None,
            "cfg5": None if roofline_k1 is None else {"gpu_body_steps_per_s": roofline_k1["body_steps_per_s"], "gpu_body_steps_per_s_f32": roofline_k1_f32["body_steps_per_s"],
                                                       "cpu_baseline": cfg5_cpu,
                                                       "gpu_over_cpu": None if cfg5_cpu is None else roofline_k1["body_steps_per_s"] / cfg5_cpu["value"]},
            "checksum": hex(red["checksum"]), "energy": red["energy"],
            "contacts_per_world_step": red["counters"]["contacts"] / total_world_steps,
            "vel_iterations_per_world_step": red["counters"]["vel_iterations"] / total_world_steps,
            "pos_iterations_per_world_step": red["counters"]["pos_iterations"] / total_world_steps,
            "wall_ms_timed_region": wall_ms,
            "ncu_counters": "profiles/ncu_counters.json" if counters.get("csrc_sha") else ("stale or absent: traffic / FP64-op counts reported as null" ),
        }
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
