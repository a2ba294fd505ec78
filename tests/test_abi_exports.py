"""CPU: the C-ABI library loads and exports every symbol include/vihds_b200.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

from vihds_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "vihds_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vh_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    syms = declared_symbols()
    assert "vh_elbo_terms_fwd" in syms and "vh_elbo_terms_bwd" in syms and len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "libvihds_b200.so does not export %s" % s
    assert lib.vh_abi_version() == 4


def test_registry_helpers():
    lib = L.load()
    assert L.model_id("dr_constant") == 0 and L.model_id("relay_constant_precisions") == 5
    assert L.solver_id("midpoint") == 1 and lib.vh_solver_id(b"dopri5") < 0
    names = L.slot_names(0)
    assert names[0] == "r" and "KGR_76" in names and "prec_x" in names
    assert "init_prec_x" in L.slot_names(2)
    p = L.vh_problem()
    p.model = 5
    assert lib.vh_state_width(C.byref(p)) == 16 and lib.vh_num_species(5) == 12
    assert lib.vh_num_weights(C.byref(p)) == 2 * (4 * 13 + 4)


def test_invalid_arguments_are_reported_not_crashed():
    lib = L.load()
    p = L.vh_problem()
    io = L.vh_fwd_io()
    p.B, p.IW, p.T = 0, 1, 10
    assert lib.vh_elbo_terms_fwd(C.byref(p), C.byref(io), None) < 0
    assert b"B, IW" in lib.vh_last_error()
