"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/particles_b200.h declares, rejects bad arguments, and refuses to compute without a
GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import particles_b200 as pb
from particles_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "particles_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", txt))
    return names - {"pf_mix64", "pf_uniform"}        # static inline helpers


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    declared = _declared()
    assert declared, "no declarations parsed"
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound
    for name in declared:
        assert getattr(L, name) is not None
    assert b"sm_100a" in L.pf_version()


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.pf_config) == 144
    assert C.sizeof(_lib.pf_step_stats) == 8 * 8 + 4 * 8 + 2 * 4 + 8 + 4 * 8 + 8 + 2 * 4


def test_bad_arguments_are_rejected_before_touching_the_gpu():
    with pytest.raises(ValueError):
        pb.FearnheadParticles(0, pb.NormalInverseChisq(0., 1., 0.1, 1.), pb.ChineseRestaurantProcess(1.))
    with pytest.raises(ValueError):
        pb.FearnheadParticles(10, pb.NormalInverseChisq(0., 1., -0.1, 1.), pb.ChineseRestaurantProcess(1.))
    with pytest.raises(ValueError):
        pb.FearnheadParticles(10, pb.NormalInverseWishart(np.zeros(5), 0.1, np.eye(5), 7.), pb.ChineseRestaurantProcess(1.))
    with pytest.raises(ValueError):
        pb.FearnheadParticles(10, pb.NormalInverseChisq(0., 1., 0.1, 1.), pb.ChineseRestaurantProcess(1.), k_max=64)
    with pytest.raises(ValueError):
        pb.FearnheadParticles(10, pb.NormalInverseChisq(0., 1., 0.1, 1.), "crp")


def test_device_uniform_stream_is_reproducible_on_the_host():
    # pf_uniform of the header restated in numpy (uint64 wrap-around arithmetic)
    def mix(z):
        z = (z + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))
    with np.errstate(over="ignore"):
        k = mix(np.uint64(7) + np.uint64(3) * np.uint64(0xD1342543DE82EF95) + np.uint64(1))
        u = (mix(k + np.arange(1000, dtype=np.uint64)) >> np.uint64(11)).astype(np.float64) / 2.0 ** 53
    assert np.all((u >= 0) & (u < 1))
    assert abs(u.mean() - 0.5) < 0.05


@pytest.mark.skipif(_lib.load() is None, reason="library missing")
def test_no_cpu_fallback_without_a_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(pb.PFError) as e:
        pb.FearnheadParticles(100, pb.NormalInverseChisq(0., 1., 0.1, 1.), pb.ChineseRestaurantProcess(1.))
    assert e.value.code == _lib.PF_ERR_CUDA
    with pytest.raises(pb.PFError):
        pb.select(np.ones(20), 10, 0.5)
