"""CPU tests: the C-ABI shared library loads and exports every symbol include/radsearch_b200.h declares, and the
ctypes structures match the C layout.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

from radiation_ppo_b200 import _lib as L, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return L.load()


def declared_functions():
    src = open(os.path.join(ROOT, "include", "radsearch_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rs_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions()
    assert {"rs_step", "rs_reset", "rs_load_scenarios", "rs_gae", "rs_adv_stats", "rs_adv_normalize", "rs_last_error",
            "rs_version", "rs_prepare", "rs_bump_ctr", "rs_query_shortest_path"} <= set(names)
    for n in names:
        assert hasattr(lib, n), n


def test_struct_layout_and_version(lib):
    assert lib.rs_version() == 3
    assert lib.rs_sizeof_config() == C.sizeof(L.RsConfig) == 52
    assert lib.rs_sizeof_state() == C.sizeof(L.RsState) == 232
    assert lib.rs_sizeof_maps_config() == C.sizeof(L.RsMapsConfig) == 40
    assert lib.rs_sizeof_maps_state() == C.sizeof(L.RsMapsState) == 88


def test_argument_errors_are_reported_without_a_gpu(lib):
    cfg, st = L.RsConfig(), L.RsState()
    rc = lib.rs_step(C.byref(cfg), C.byref(st), None, None, None, None, None, None, None, None, 0, 0, 0, 0, None, 0, 0, None)
    assert rc < 0 and b"n_env" in lib.rs_last_error()
    cfg.n_agents = 99
    rc = lib.rs_reset(C.byref(cfg), C.byref(st), None, None, None, 4, 0, 0, 0, None, 0, 0, None)
    assert rc < 0 and b"n_agents" in lib.rs_last_error()
    assert lib.rs_gae(None, None, None, None, None, None, 1, 1, 0.99, 0.9, None, 0, None) < 0
    mc, ms = L.RsMapsConfig(), L.RsMapsState()
    assert lib.rs_maps_update(C.byref(mc), C.byref(ms), None, None, None, 0, 8, None) < 0 and b"n_agents" in lib.rs_last_error()
    mc.n_agents, mc.dim_x, mc.dim_y, mc.base, mc.log_cap, mc.resolution_accuracy, mc.scale = 2, 27, 27, 242, 246, 22.0, 1 / 2200
    assert lib.rs_maps_reset(C.byref(mc), C.byref(ms), None, 0, 8, None) < 0 and b"NULL" in lib.rs_last_error()


def test_product_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import radiation_ppo_b200 as rp

    with pytest.raises(rp.RadSearchLibraryError):
        rp.RadSearch(num_envs=4)
    with pytest.raises(rp.RadSearchLibraryError):
        rp.PPOBuffer(11, 4, 4, 1)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "radiation_ppo_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f
