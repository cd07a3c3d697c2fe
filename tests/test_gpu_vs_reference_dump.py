"""GPU: the CUDA path (through the C ABI) against dumps printed by the REFERENCE ITSELF — tbogdala/cubez's Go sources
translated mechanically by oracle/go2cpp.py (see tests/test_oracle_vs_reference_dump.py and oracle/make_ref_golden.py).
Per frame: contact count, (body, body) sequence, as-generated contact geometry and the raw bits of every body's state.
Covers every BASELINE config: cfg1 (600 frames), cfg2 (600 frames, 64 bullets), cfg3 at its real size (4 096 bodies, 80
frames, 13 k contacts, iteration caps hit), cfg4 (worlds 0..255, 600 frames), cfg5 (65 536 free bodies)."""
import pytest

import refdump
from cubez_b200 import _abi
from ref_cases import REF_CASES, ref_text

pytestmark = pytest.mark.gpu


def _flags(name):
    if name.startswith("pile4096") or name.startswith("pile216") or name.startswith("random1x300"):     # (pile216_f32 too)
        return _abi.WORLD_BROADPHASE
    return 0


@pytest.mark.parametrize("name", sorted(REF_CASES))
def test_gpu_equals_reference_dump(name):
    from cubez_b200.api import BatchedWorld
    make, frames = REF_CASES[name]
    scene = make()
    gpu = BatchedWorld.from_scene(scene, flags=_flags(name))
    lines = refdump.run_dump(gpu, scene, frames)
    gpu.close()
    assert refdump.first_difference(ref_text(name), lines) is None


@pytest.mark.parametrize("name", ["cubedrop_600", "ballistic16_300", "batched64_from1000_300"])
def test_gpu_multi_kernel_path_equals_reference_dump(name):
    from cubez_b200.api import BatchedWorld
    make, frames = REF_CASES[name]
    scene = make()
    gpu = BatchedWorld.from_scene(scene, flags=_abi.WORLD_NO_FUSED)
    lines = refdump.run_dump(gpu, scene, frames)
    gpu.close()
    assert refdump.first_difference(ref_text(name), lines) is None
