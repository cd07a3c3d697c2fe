"""Developer probe: run on the GPU box (gpurun) to shake out the CUDA path section by section."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cubez_b200 import scenes, _abi
from cubez_b200.api import BatchedWorld, Context
from oracle_lib import OracleWorld, Oracle

FIELDS = ("position", "orientation", "velocity", "rotation", "motion", "is_awake", "transform", "inverse_inertia_tensor_world", "last_frame_acceleration")

def section(name):
    def deco(fn):
        t = time.time()
        try:
            fn()
            print(f"[ok] {name} ({time.time()-t:.2f}s)", flush=True)
        except Exception:
            print(f"[FAIL] {name}", flush=True)
            traceback.print_exc()
        return fn
    return deco

def compare_world(scene, n_steps, flags=0, env=None, every=1, label=""):
    for k, v in (env or {}).items():
        os.environ[k] = v
    gpu = BatchedWorld.from_scene(scene, flags=flags)
    cpu = OracleWorld.from_scene(scene)
    for k in (env or {}):
        os.environ.pop(k)
    bad = None
    for s in range(0, n_steps, every):
        gpu.step(scene.dt, every); cpu.step(scene.dt, every)
        gc, gp, gv = gpu.last_counts(); cc, cp, cv = cpu.last_counts()
        if not (np.array_equal(gc, cc) and np.array_equal(gp, cp) and np.array_equal(gv, cv)):
            w = int(np.argmax((gc != cc) | (gp != cp) | (gv != cv)))
            bad = f"step {s+every-1} world {w}: contacts {gc[w]} vs {cc[w]}, pos {gp[w]} vs {cp[w]}, vel {gv[w]} vs {cv[w]}"
            break
        for w in range(min(scene.n_worlds, 4)):
            if gpu.contact_pairs(w) != cpu.contact_pairs(w):
                bad = f"step {s} world {w}: contact pairs differ"
                break
        if bad: break
    g, c = gpu.download(), cpu.download()
    mism = [f for f in FIELDS if not np.array_equal(getattr(g, f), getattr(c, f))]
    ck_g, ck_c = gpu.checksum_energy(), cpu.checksum_energy()
    print(f"   {label or scene.name}: first divergence: {bad}; state mismatch: {mism}; checksum {ck_g[0]==ck_c[0]} energy {ck_g[1]:.9g} vs {ck_c[1]:.9g}", flush=True)
    gpu.close(); cpu.close()
    assert bad is None and not mism and ck_g[0] == ck_c[0]

@section("math op")
def _():
    ctx = Context.get(0, "f64")
    print("   cross", ctx.math_op("VEC_CROSS", [1, 2, 3], [10, 11, 12]), "dot", ctx.math_op("VEC_DOT", [1, 2, 3], [10, 11, 12]))

@section("integrate shim f64 vs oracle")
def _():
    sc = scenes.free_bodies(_abi.F64, n=1 << 14)
    ctx, orc = Context.get(0, "f64"), Oracle("f64")
    a, b = sc.bodies.copy(), sc.bodies.copy()
    for _ in range(3):
        ctx.integrate(a, sc.dt); orc.integrate(b, sc.dt)
    mism = [f for f in FIELDS if not np.array_equal(getattr(a, f), getattr(b, f))]
    print("   mismatch:", mism)
    assert not mism

@section("cubedrop multi-kernel")
def _():
    compare_world(scenes.cubedrop(), 200, flags=_abi.WORLD_NO_FUSED, label="cubedrop/multi")

for G in ("32", "16", "8"):
    @section(f"cubedrop fused G={G}")
    def _(G=G):
        compare_world(scenes.cubedrop(), 200, env={"CUBEZ_FUSED_G": G}, label=f"cubedrop/fused{G}")

@section("ballistic multi-kernel")
def _():
    compare_world(scenes.ballistic(n_bullets=16), 300, flags=_abi.WORLD_NO_FUSED, label="ballistic/multi")

@section("ballistic fused")
def _():
    compare_world(scenes.ballistic(n_bullets=16), 300, label="ballistic/fused")

@section("batched 256 worlds fused, 10-step calls")
def _():
    compare_world(scenes.batched_cubedrop(n_worlds=256), 300, every=10, label="batched256/fused")

@section("batched 64 worlds multi")
def _():
    compare_world(scenes.batched_cubedrop(n_worlds=64), 150, every=5, flags=_abi.WORLD_NO_FUSED, label="batched64/multi")

@section("f32 cubedrop fused + multi")
def _():
    compare_world(scenes.cubedrop(_abi.F32), 200, label="cubedrop32/fused")
    compare_world(scenes.cubedrop(_abi.F32), 200, flags=_abi.WORLD_NO_FUSED, label="cubedrop32/multi")

@section("bench integrate 16M f64")
def _():
    ctx = Context.get(0, "f64")
    ms, ck = ctx.bench_integrate(1 << 24, warmup=3, steps=20)
    print(f"   {ms:.3f} ms/step -> {(1<<24)/ms/1e6:.2f} G body-steps/s, {(1<<24)*531/ms/1e6:.0f} GB/s algorithmic")

@section("bench integrate 16M f32")
def _():
    ctx = Context.get(0, "f32")
    ms, ck = ctx.bench_integrate(1 << 24, warmup=3, steps=20)
    print(f"   {ms:.3f} ms/step -> {(1<<24)/ms/1e6:.2f} G body-steps/s, {(1<<24)*267/ms/1e6:.0f} GB/s algorithmic")

@section("episodes parity (fused8)")
def _():
    sc = scenes.batched_cubedrop(n_worlds=24)
    gpu, cpu = BatchedWorld.from_scene(sc), OracleWorld.from_scene(sc)
    ph = (np.arange(24) * 7) % 130
    gpu.set_episodes(130, ph); cpu.set_episodes(130, ph)
    for s_ in range(0, 300, 25):
        gs, cs = gpu.step(sc.dt, 25), cpu.step(sc.dt, 25)
        assert all(gs[k] == cs[k] for k in ("contacts", "pos_iterations", "vel_iterations")), (s_, gs, cs)
    g, c = gpu.download(), cpu.download()
    assert not [f for f in FIELDS if not np.array_equal(getattr(g, f), getattr(c, f))]

@section("cfg4 throughput sweep (65536 worlds, stationary episodes)")
def _():
    W = 65536
    sc = scenes.batched_cubedrop(n_worlds=W)
    ph = (np.arange(W) % 600).astype(np.int32)
    ref = None
    for G, MB, LS, TH in (("8", "2", "1", "128"), ("8", "2", "1", "64"), ("8", "2", "1", "32"), ("8", "2", "0", "64"), ("16", "2", "1", "64")):
        os.environ["CUBEZ_FUSED_G"] = G; os.environ["CUBEZ_FUSED_MINB"] = MB; os.environ["CUBEZ_FUSED_LOCKSTEP"] = LS; os.environ["CUBEZ_FUSED_THREADS"] = TH
        gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
        gpu.set_episodes(600, ph)
        t = time.time(); gpu.step(sc.dt, 600); pre = time.time() - t
        st = gpu.step(sc.dt, 60)
        ck = gpu.checksum_energy()[0]
        ref = ref or ck
        print(f"   G={G} MINB={MB} LOCKSTEP={LS} THREADS={TH}: pre-roll {pre:.2f}s; {st['device_ms']/60:.3f} ms/frame -> {W*60/st['device_ms']/1e3:.2f} M world-steps/s; checksum same {ck == ref}; vel it/ws {st['vel_iterations']/(W*60):.2f}", flush=True)
        gpu.close()
    for k in ("CUBEZ_FUSED_G", "CUBEZ_FUSED_MINB", "CUBEZ_FUSED_LOCKSTEP", "CUBEZ_FUSED_THREADS"):
        os.environ.pop(k)

@section("step_host e2e timing (65536 worlds)")
def _():
    W = 65536
    sc = scenes.batched_cubedrop(n_worlds=W)
    ctx = Context.get(0, "f64")
    gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
    gpu.set_episodes(600, (np.arange(W) % 600).astype(np.int32))
    gpu.step(sc.dt, 600)
    ref = BatchedWorld.from_scene(sc, contacts_per_world=64)
    ref.set_episodes(600, (np.arange(W) % 600).astype(np.int32))
    ref.step(sc.dt, 600)
    for pinned in (True, False):
        host = gpu.download(out=ctx.pinned_bodies(W * 8)) if pinned else gpu.download()
        gpu.step_host(host, sc.dt, 1); ref.step(sc.dt, 1)
        t = time.perf_counter()
        for _ in range(10):
            gpu.step_host(host, sc.dt, 1)
        el = (time.perf_counter() - t) / 10
        ref.step(sc.dt, 10)
        r = ref.download()
        same = all(np.array_equal(getattr(host, f), getattr(r, f)) for f in FIELDS)
        print(f"   pinned={pinned}: {el*1e3:.2f} ms/frame e2e -> {W/el/1e6:.2f} M world-steps/s; equals device-resident run: {same}", flush=True)
