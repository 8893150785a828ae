"""The C-ABI library loads and exports every symbol include/ppe.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

import path_planner_b200 as ppb
from path_planner_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ppe.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ppe_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    for name in ("ppe_create", "ppe_destroy", "ppe_set_config", "ppe_set_map_bitmap", "ppe_set_obstacles_gaussian",
                 "ppe_put_ribbon_set", "ppe_dubins_batch", "ppe_true_cost_batch", "ppe_get_ribbons_after", "ppe_best",
                 "ppe_true_cost_batch_device", "ppe_best_copy_device"):
        assert name in syms


def test_library_exports_every_declared_symbol():
    lib = ppb.load_library()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_struct_layouts_match_the_compiled_library():
    lib = ppb.load_library()
    assert lib.ppe_abi_version() == abi.PPE_ABI_VERSION
    assert lib.ppe_abi_sizeof_edge() == abi.EDGE_DTYPE.itemsize == 176
    assert lib.ppe_abi_sizeof_edge_result() == abi.RESULT_DTYPE.itemsize == 208
    assert lib.ppe_abi_sizeof_config() == C.sizeof(abi.PpeConfig)
    # offsets the bench relies on (int32 columns of the result record)
    assert abi.RESULT_DTYPE.fields["infeasible"][1] == 45 * 4
    assert abi.RESULT_DTYPE.fields["n_samples"][1] == 47 * 4
    assert abi.RESULT_DTYPE.fields["n_checkpoints"][1] == 48 * 4


def test_no_cpu_fallback():
    """Without a CUDA device the engine refuses to exist; with one it must come up."""
    import torch

    if torch.cuda.is_available():
        ppb.EdgeEngine(0).close()
    else:
        with pytest.raises(ppb.PpeError):
            ppb.EdgeEngine(0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under path_planner_b200/ may reference it."""
    pkg = os.path.join(ROOT, "path_planner_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"libppe_oracle|ppe_oracle\.c|oracle_true_cost|from tests|import tests", text):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_harness_library_exports_every_declared_symbol():
    """include/ppe_harness.h (the standalone harness's C ABI) against path_planner_b200/libppe_harness.so; skipped where the
    harness was not built (it compiles against the reference's headers)."""
    from path_planner_b200 import harness as ph
    if not ph.available():
        pytest.skip("libppe_harness.so not built (needs /root/reference at build time)")
    text = open(os.path.join(ROOT, "include", "ppe_harness.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    syms = sorted(set(re.findall(r"\b(pph_[a-z0-9_]+)\s*\(", text)))
    assert "pph_plan" in syms and "pph_create" in syms
    lib = C.CDLL(ph.HARNESS_PATH)
    assert not [s for s in syms if not hasattr(lib, s)]
    assert C.sizeof(ph.PlanStats) == 24 * 8 and C.sizeof(ph.PlanOptions) == 64 and ph.PATH_DTYPE.itemsize == 88
