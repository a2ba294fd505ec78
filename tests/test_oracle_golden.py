"""Pin the CPU oracle (oracle/vihds_oracle.py) to the golden vectors minted from the reference itself."""
import numpy as np
import pytest

from conftest import golden_cases, load_case
import vihds_oracle as O


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_reference(name):
    case = load_case(name)
    out = O.elbo_step(case)
    f64 = str(case["dtype"]) == "float64"
    tol = 1e-9 if f64 else 2e-5
    gtol = 1e-7 if f64 else 2e-3
    assert _rel(out["theta"], case["theta"]) < tol
    if "x_states" in case:
        # per-timepoint states, relative to each species' peak
        xs, ref = out["x_states"].numpy(), case["x_states"]
        for s in range(ref.shape[2]):
            assert _rel(xs[:, :, s], ref[:, :, s]) < tol * 5, "species %d" % s
        assert _rel(out["x_predict"], case["x_predict"]) < tol * 5
        assert _rel(out["precisions"], case["precisions"]) < tol * 5
    else:
        assert _rel(out["x_states"][:, :, :, -1], case["x_states_last"]) < tol * 5
    assert _rel(out["log_p_by_species"], case["log_p_by_species"]) < tol * 10
    assert _rel(out["log_p_theta"], case["log_p_theta"]) < tol * 10
    assert _rel(out["log_q_theta"], case["log_q_theta"]) < tol * 10
    assert abs(float(out["loss"]) - float(case["loss"])) <= tol * 10 * abs(float(case["loss"]))
    # gradient: per-individual parameters row by row, global parameters as the total over individuals
    per_ind = case["per_individual"].astype(bool)
    g_mu, g_prec = out["grad_q_mu"].numpy(), out["grad_q_prec"].numpy()
    kinds = case["kinds"]
    for g, ref in ((g_mu, case["grad_q_mu"]), (g_prec, case["grad_q_prec"])):
        tot = g.sum(0)
        ref_tot = np.where(per_ind, ref.sum(0), ref[0])
        sel = kinds != 0
        assert _rel(tot[sel], ref_tot[sel]) < gtol
        if per_ind.any():
            assert _rel(g[:, per_ind], ref[:, per_ind]) < gtol
    for k in case:
        if k.startswith("gw:"):
            assert _rel(out[k], case[k]) < gtol, k
