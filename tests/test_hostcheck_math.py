"""The hand-written RHS / stepper / discrete-adjoint maths of vihds_b200/csrc, compiled for the HOST, against the
golden vectors minted from the reference (no GPU needed).  The CUDA kernels instantiate the same headers; the
``-m gpu`` tests then only have to establish that the device build computes what the host build computes."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden_cases, load_case
import helpers as H
from vihds_b200 import _lib as L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


DR_CASES = golden_cases()


@pytest.fixture(scope="module")
def hc():
    return H.build_hostcheck()


@pytest.mark.parametrize("name", DR_CASES)
def test_host_math_matches_reference(hc, name):
    case = load_case(name)
    f64 = str(case["dtype"]) == "float64"
    dt = np.float64 if f64 else np.float32
    model = H.MODEL_IDS[str(case["model"])]
    slot_names = [hc.hc_slot_name(model, s).decode() for s in range(hc.hc_num_slots())]
    src, extra = H.slot_map(case, slot_names)
    w, gw_ref = H.flat_weights(case)
    p = H.make_problem(case, src, 0 if extra is None else extra.shape[0])
    B, IW, P, T = p.B, p.IW, p.P, p.T
    N = B * IW
    S = {0: 8, 1: 8, 2: 12, 3: 12, 4: 12, 5: 16, 6: 10, 7: 4, 8: 8, 9: 6, 10: 10, 11: 5, 12: 9, 13: 11, 14: 15}[model]
    lo, hi = H.clip_bounds(case)
    keep = dict(
        times=case["times"].astype(dt), u=np.ascontiguousarray(case["u"].reshape(N, P)), q_mu=case["q_mu"].astype(dt),
        q_prec=case["q_prec"].astype(dt), p_mu=case["p_mu"].astype(dt), p_prec=case["p_prec"].astype(dt), clip_lo=lo,
        clip_hi=hi, kind=case["kinds"].astype(np.int32), extra=extra, treatments=case["inputs"].astype(dt),
        dev_1hot=case["dev_1hot"].astype(dt), observations=np.ascontiguousarray(case["observations"].astype(dt)),
        weights=None if w is None else w.astype(dt),
        theta=np.zeros((P, N), dt), x_states=np.zeros((T, S, N), dt), x_predict=np.zeros((T, 4, N), dt),
        logp_by_species=np.zeros((N, 4), dt), logp_theta=np.zeros(N, dt), logq_theta=np.zeros(N, dt))
    io = L.vh_fwd_io(**{k: _ptr(v) for k, v in keep.items()})
    assert hc.hc_fwd(C.byref(p), C.byref(io)) == 0

    tol = 1e-9 if f64 else 3e-5
    assert _rel(keep["theta"].reshape(P, B, IW), case["theta"]) < tol
    xs = keep["x_states"].reshape(T, S, B, IW).transpose(2, 3, 1, 0)
    ns = case["x_states_last"].shape[2] if "x_states_last" in case else case["x_states"].shape[2]
    if "x_states" in case:
        for s in range(ns):
            assert _rel(xs[:, :, s], case["x_states"][:, :, s]) < tol * 5, "species %d" % s
        xp = keep["x_predict"].reshape(T, 4, B, IW).transpose(2, 3, 1, 0)
        assert _rel(xp, case["x_predict"]) < tol * 5
        if S > ns:
            assert _rel(xs[:, :, ns:], case["precisions"]) < tol * 5
    else:
        assert _rel(xs[:, :, :ns, -1], case["x_states_last"]) < tol * 5
    lpx = keep["logp_by_species"].reshape(B, IW, 4)
    assert _rel(lpx, case["log_p_by_species"]) < tol * 10
    assert _rel(keep["logp_theta"].reshape(B, IW), case["log_p_theta"]) < tol * 10
    assert _rel(keep["logq_theta"].reshape(B, IW), case["log_q_theta"]) < tol * 10
    cost, g = H.iwae_upstream(keep["logp_by_species"], keep["logp_theta"], keep["logq_theta"], B, IW)
    assert abs(cost - float(case["loss"])) <= tol * 10 * abs(float(case["loss"]))

    # reverse pass seeded with the IWAE cost gradient
    g = g.astype(dt)
    bkeep = dict(g_logp_by_species=np.ascontiguousarray(np.repeat(g[:, None], 4, 1)), g_logp_theta=g, g_logq_theta=(-g).astype(dt),
                 d_q_mu=np.zeros((B, P), dt), d_q_prec=np.zeros((B, P), dt),
                 d_weights=None if w is None else np.zeros_like(w, dtype=dt),
                 d_extra=None if extra is None else np.zeros_like(extra))
    bio = L.vh_bwd_io(fwd=io, **{k: _ptr(v) for k, v in bkeep.items()})
    assert hc.hc_bwd(C.byref(p), C.byref(bio)) == 0  # fwd.theta handed back: the reverse sweep re-reads theta
    first = {k: v.copy() for k, v in bkeep.items() if k.startswith("d_") and v is not None}
    io2 = L.vh_fwd_io(**{k: _ptr(v) for k, v in keep.items() if k != "theta"})
    bio2 = L.vh_bwd_io(fwd=io2, **{k: _ptr(v) for k, v in bkeep.items()})
    assert hc.hc_bwd(C.byref(p), C.byref(bio2)) == 0  # fwd.theta == NULL: re-sampled from u; same gradients
    for k, v in first.items():
        assert _rel(bkeep[k], v) < (1e-12 if f64 else 2e-5), k
    gtol = 1e-6 if f64 else 3e-3
    per_ind = case["per_individual"].astype(bool)
    sel = case["kinds"] != 0
    for got, ref in ((bkeep["d_q_mu"], case["grad_q_mu"]), (bkeep["d_q_prec"], case["grad_q_prec"])):
        ref_tot = np.where(per_ind, ref.sum(0), ref[0])
        assert _rel(got.sum(0)[sel], ref_tot[sel]) < gtol
        for k in np.nonzero(sel)[0]:  # every parameter against its own scale, not just the largest one
            assert abs(got.sum(0)[k] - ref_tot[k]) <= 10 * gtol * abs(ref_tot[k]) + 1e-6 * np.max(np.abs(ref_tot)) * gtol, (
                "column %d (%s)" % (k, case["names"][k]))
        if per_ind.any():
            assert _rel(got[:, per_ind], ref[:, per_ind]) < gtol
    if w is not None:
        assert _rel(bkeep["d_weights"], gw_ref) < gtol
    if str(case["model"]) == "dr_blackbox":
        dW, db = H.offset_layer_grads(case, bkeep["d_extra"])
        assert _rel(dW, case["gw:ode_model.offset_layer.weight"]) < gtol
        assert _rel(db, case["gw:ode_model.offset_layer.bias"]) < gtol


@pytest.mark.parametrize("version", [1, 2])
def test_matrix_form_of_the_midpoint_adjoint_equals_the_vjp_form(version):
    """vh_mx_math.cuh (what elbo_bwd_mx_kernel's producers and consumer compute) against rk_step_adjoint (what every other
    reverse kernel computes) on random states, parameters and cotangents, in fp64 on the host: the two are the same linear
    map, lambda0 = (I + h A + h^2/2 A B)^T lambda1."""
    import ctypes as C

    lib = H.build_hostcheck()
    dp = C.POINTER(C.c_double)
    lib.hc_mx_step.argtypes = [C.c_int, dp, dp, dp, dp, C.c_double, C.c_double, dp, dp]
    lib.hc_mx_step.restype = None
    rng = np.random.RandomState(7 + version)
    nslot = lib.hc_num_slots()
    worst = 0.0
    for trial in range(50):
        th = np.ascontiguousarray(rng.uniform(0.05, 2.5, size=nslot))
        tc = np.ascontiguousarray(rng.uniform(0.1, 20.0, size=3))
        x = np.ascontiguousarray(rng.uniform(0.01, 3.0, size=8))
        lam1 = np.ascontiguousarray(rng.randn(8))
        t0 = float(rng.uniform(0.0, 10.0))
        t1 = t0 + float(rng.uniform(0.05, 0.4))
        ref, mx = np.zeros(8), np.zeros(8)
        lib.hc_mx_step(version, th.ctypes.data_as(dp), tc.ctypes.data_as(dp), x.ctypes.data_as(dp), lam1.ctypes.data_as(dp),
                       t0, t1, ref.ctypes.data_as(dp), mx.ctypes.data_as(dp))
        assert np.all(np.isfinite(ref))
        worst = max(worst, float(np.max(np.abs(ref - mx)) / max(1e-300, np.max(np.abs(ref)))))
    assert worst < 1e-12, worst
