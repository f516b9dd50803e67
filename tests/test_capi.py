"""The C-ABI library: loads, exports every symbol include/asif_b200.h declares, validates arguments,
and refuses to compute without a device (no CPU fallback).  No compute calls here; CPU only."""
import ctypes as C
import os
import re
import subprocess

import pytest

import conftest as cf

ROOT = cf.ROOT


@pytest.fixture(scope="module")
def lib():
    import asif_b200
    from asif_b200 import _build
    if _build.needs_build():
        _build.build()
    return asif_b200.load_library()


def declared_functions():
    src = open(os.path.join(ROOT, "include", "asif_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asif_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(lib):
    names = declared_functions()
    assert len(names) >= 11
    for n in names:
        assert hasattr(lib, n), "libasif_b200.so does not export %s" % n
    assert lib.asif_b200_abi_version() == 3


def test_struct_layout_matches(lib):
    import asif_b200
    cfg = asif_b200.EngineConfig()
    assert lib.asif_engine_config_init(C.byref(cfg), asif_b200.FILTER_IMPLICIT_TB, asif_b200.MODEL_DOUBLE_INTEGRATOR_TB) == 0
    assert cfg.struct_size == C.sizeof(asif_b200.EngineConfig)
    # defaults of include/asif_implicit_tb.h:19-33 and the example's input bounds
    assert (cfg.relaxCost, cfg.relaxLb, cfg.relaxTTS, cfg.relaxMinOrtho) == (50.0, 5.0, 5.0, 5.0)
    assert (cfg.backTrajHorizon, cfg.backTrajExtend, cfg.backTrajDt, cfg.backTrajMinOrtho) == (1.0, 0.05, 0.01, 0.01)
    assert (cfg.satSharpness, cfg.inf, cfg.npBTSS) == (0.1, 1e20, 4)
    assert (cfg.lb[0], cfg.ub[0]) == (-1.0, 1.0)
    assert lib.asif_engine_config_init(C.byref(cfg), asif_b200.FILTER_EXPLICIT, asif_b200.MODEL_DOUBLE_INTEGRATOR) == 0
    assert (cfg.relaxLb, cfg.relaxCost) == (5.0, 50.0)
    assert lib.asif_engine_config_init(C.byref(cfg), 99, 1) < 0
    assert b"unknown filter" in lib.asif_last_error()


def test_argument_validation_and_no_fallback(lib):
    import asif_b200
    cfg = asif_b200.EngineConfig()
    lib.asif_engine_config_init(C.byref(cfg), asif_b200.FILTER_IMPLICIT_TB, asif_b200.MODEL_DOUBLE_INTEGRATOR_TB)
    h = C.c_void_p()
    cfg.struct_size = 12
    assert lib.asif_engine_create(C.byref(cfg), C.byref(h)) == -101  # ASIF_ERR_INVALID_ARGUMENT
    assert lib.asif_engine_create(None, C.byref(h)) == -101
    assert lib.asif_engine_filter_batch(None, 1, None, None, None, None, None, None, 0, None) == -101
    n_dev = lib.asif_device_count()
    if n_dev <= 0:
        # without a device nothing computes: create fails with ASIF_ERR_NO_DEVICE and says why
        lib.asif_engine_config_init(C.byref(cfg), asif_b200.FILTER_IMPLICIT_TB, asif_b200.MODEL_DOUBLE_INTEGRATOR_TB)
        assert lib.asif_engine_create(C.byref(cfg), C.byref(h)) == -104
        assert b"no CPU fallback" in lib.asif_last_error()
        with pytest.raises(asif_b200.AsifError):
            asif_b200.Engine(asif_b200.FILTER_EXPLICIT, asif_b200.MODEL_DOUBLE_INTEGRATOR)


def test_product_does_not_touch_oracle():
    """Nothing under asif_b200/ or include/ may import, include, link or load anything from oracle/."""
    bad = []
    for base in ("asif_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".so", ".pyc", ".o")) or f == "host_check":
                    continue
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"oracle[/.]|liboracle|libasif_ref|pyref", txt):
                    bad.append(os.path.join(d, f))
                # nor may it include, import or load the test-only host team of the QP solver (tests/hostemu)
                if re.search(r"(#include|import|CDLL|dlopen)[^\n]*hostemu", txt):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
    so = os.path.join(ROOT, "asif_b200", "libasif_b200.so")
    if os.path.exists(so):
        out = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
        assert "oracle" not in out and "asif_ref" not in out


def test_host_layer_compiles():
    """The C++ host layer (QPWrapperB200, FilterBatch*) builds against the C ABI with a plain C++11 compiler."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "asif_b200", "host"), "-s"])
    assert os.path.exists(os.path.join(ROOT, "asif_b200", "host", "host_check"))


def test_loop_config_defaults_and_argument_checks(lib):
    """asif_loop_config mirrors the example main loops' constants; argument errors are reported without a device."""
    from asif_b200 import capi
    L = lib
    c = capi.LoopConfig()
    assert L.asif_loop_config_init(C.byref(c)) == 0
    assert c.struct_size == C.sizeof(capi.LoopConfig)
    assert c.dt == 1e-3 and c.steps_per_sample == 1 and c.smooth_bounds == 0 and c.log_after_step == 1
    assert (c.smooth_lb, c.smooth_ub) == (-20.0, 20.0) and c.plant_gain == 1.0
    assert L.asif_engine_closed_loop(None, 1, C.byref(c), None, None, None, None, None, None, None, 0, None) == -101
    assert b"NULL" in L.asif_last_error()
    f = capi.loop_log_fields(2, 1, 2)
    assert f["t"] == 0 and f["x0"] == 1 and f["xEstim0"] == 3 and f["relax1"] == f["rc"] - 1 and len(f) == 16
