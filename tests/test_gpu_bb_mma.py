"""-m gpu: the tensor-core dr_blackbox kernels (csrc/vh_bb_mma.cuh: warp-level mma.sync 3xTF32, forward and reverse)
against the scalar kernels on the same inputs -- every output of the forward launch and every gradient of the reverse
launch -- over all five solvers, ragged / tiny batches and the BASELINE size (B = 36 x IW = 200).  (The golden and oracle
comparisons of test_gpu_parity.py / test_gpu_properties.py run the tensor-core kernels too: they are the default.)"""
import numpy as np
import pytest

from conftest import load_case
from test_gpu_parity import _rel, run_case_on_gpu
from test_gpu_properties import sub_case

pytestmark = pytest.mark.gpu

KEYS = ["theta", "x_states", "x_predict", "logp_by_species", "logp_theta", "logq_theta", "cost", "d_q_mu", "d_q_prec",
        "d_weights", "d_extra"]


def _both(case, monkeypatch):
    monkeypatch.setenv("VIHDS_BB_IMPL", "scalar")
    ref = run_case_on_gpu(case)
    monkeypatch.setenv("VIHDS_BB_IMPL", "mma")
    return ref, run_case_on_gpu(case)


@pytest.mark.parametrize("solver,t,rows,iw", [("euler", 2, None, None), ("midpoint", 3, None, None), ("midpoint", None, None, None),
                                              ("rk4", 9, None, None), ("modeuler", 9, None, None), ("modeulerwhile", 9, None, None),
                                              ("midpoint", 9, [0, 1, 2, 3, 4], 7), ("midpoint", 4, [3], 1)])
def test_tensor_core_kernels_match_scalar_kernels(solver, t, rows, iw, monkeypatch):
    case = load_case("dr_blackbox_icml_midpoint_f32_iw8")
    c = sub_case(case, rows if rows is not None else list(range(case["u"].shape[0])), iw, t)
    c["solver"] = np.array(solver)
    ref, got = _both(c, monkeypatch)
    for k in KEYS:
        assert _rel(got[k], ref[k]) < 5e-5, k  # measured 1e-7 .. 2e-5 (3xTF32 products, SFU sigmoid)


def test_tensor_core_kernels_at_the_baseline_size(monkeypatch):
    case = load_case("dr_blackbox_icml_midpoint_f32_iw200")
    ref, got = _both(case, monkeypatch)
    for k in KEYS:
        assert _rel(got[k], ref[k]) < 1e-4, k  # measured 3e-5 on the q gradients
    B, IW, P, T, S = got["dims"]
    xs = got["x_states"].reshape(T, S, B, IW).transpose(2, 3, 1, 0)
    assert _rel(xs[:, :, :6, -1], case["x_states_last"]) < 1e-4
    assert abs(float(got["cost"][0]) - float(case["loss"])) <= 1e-4 * abs(float(case["loss"]))
