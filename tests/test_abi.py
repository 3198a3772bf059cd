"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports
every symbol include/pba.h declares, and fails loudly (no fallback) without a
GPU.  No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pba_b200 as pb
from pba_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "pba.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pba_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_ffi_list_agree():
    assert declared_symbols() == sorted(_ffi.PBA_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_ffi.lib_path())
    for name in declared_symbols():
        assert hasattr(lib, name), "libpba_b200.so does not export %s" % name
    assert lib.pba_abi_version() == 2


def test_struct_layouts_match_header():
    # sizes the C side computes for the same structs (compiled into the synth lib's translation unit)
    assert C.sizeof(_ffi.pba_problem) == 4 * 4 + 8 + 5 * 8 + 6 * 8 + 2 * 8 + 8 + 3 * 4 + 4 + 8
    assert C.sizeof(_ffi.pba_iteration) == 4 * 4 + 10 * 8
    assert C.sizeof(_ffi.pba_kernel_stat) == 48 + 8 + 8


def test_options_defaults_are_the_references():
    o = pb.BundleAdjustmentOptions()
    # include/visnav/map_utils.h:304-319
    assert (o.verbosity_level, o.optimize_intrinsics, o.use_huber, o.huber_parameter, o.max_num_iterations) == \
        (1, False, True, 1.0, 20)
    c = o.to_c()
    # Ceres 2.0.0 Solver::Options (include/ceres/solver.h:277-322)
    assert c.initial_trust_region_radius == 1e4 and c.max_trust_region_radius == 1e16
    assert c.min_relative_decrease == 1e-3 and c.min_lm_diagonal == 1e-6 and c.max_lm_diagonal == 1e32
    assert c.function_tolerance == 1e-6 and c.gradient_tolerance == 1e-10 and c.parameter_tolerance == 1e-8
    assert c.max_num_consecutive_invalid_steps == 5 and c.jacobi_scaling == 1


def test_status_strings():
    lib = _ffi.load_lib()
    assert lib.pba_status_string(0) == b"PBA_OK"
    assert lib.pba_status_string(2) == b"PBA_ERR_NO_DEVICE"


@pytest.mark.skipif(pb.device_count() > 0, reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    prob, _ = pb.make_scene(pb.MODE_GEOMETRIC, 6, 40, "pinhole")
    before = prob.poses.copy()
    with pytest.raises(RuntimeError, match="PBA_ERR_NO_DEVICE"):
        pb.bundle_adjustment(prob, pb.BundleAdjustmentOptions(verbosity_level=0))
    assert np.array_equal(prob.poses, before)
    with pytest.raises(RuntimeError, match="PBA_ERR_NO_DEVICE"):
        pb.Engine(prob)


def test_product_never_touches_the_oracle():
    """The product tree must not import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "photometric-bundle-adjustment_b200")
    banned = ("oracle/", "libpba_oracle", "libpba_ref", "pba_oracle_", "pba_ref_", "oracle_ffi", "libceres", "#include <ceres")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                code = "\n".join(line for line in open(os.path.join(dirpath, f)).read().splitlines()
                                 if not line.lstrip().startswith(("//", "#", "*", "/*")))
                for b in banned:
                    assert b not in code, "%s references %r" % (os.path.join(dirpath, f), b)
    out = os.popen("ldd %s" % _ffi.lib_path()).read()
    assert "oracle" not in out and "ceres" not in out
