"""ctypes binding of libvihds_b200.so (C ABI declared in include/vihds_b200.h).

The library is the product: there is NO fallback.  If it is missing or was built without CUDA kernels the import of
anything that needs it raises immediately (``load()``), so a GPU test can never pass on a silent PyTorch path.
"""
import ctypes as C
import os

VH_MAX_SLOTS = 64
VH_SLOT_UNUSED = -1000000
VH_F32, VH_F64 = 0, 1
KIND_CONSTANT, KIND_NORMAL, KIND_LOGNORMAL = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VIHDS_B200_LIB") or os.path.join(_HERE, "csrc", "libvihds_b200.so")  # env: kernel experiments


class vh_problem(C.Structure):
    _fields_ = [
        ("model", C.c_int), ("solver", C.c_int), ("dtype", C.c_int),
        ("B", C.c_int), ("IW", C.c_int), ("T", C.c_int), ("P", C.c_int), ("C", C.c_int), ("D", C.c_int), ("E", C.c_int),
        ("n_hidden", C.c_int), ("n_hidden_states", C.c_int), ("n_latent", C.c_int),
        ("n_z", C.c_int), ("n_x", C.c_int), ("n_y", C.c_int),
        ("slot_src", C.c_int * VH_MAX_SLOTS),
        ("init_latent_species", C.c_double), ("init_prec", C.c_double),
    ]


_FWD_FIELDS = ["times", "u", "q_mu", "q_prec", "p_mu", "p_prec", "clip_lo", "clip_hi", "kind", "extra", "treatments",
               "dev_1hot", "observations", "weights", "theta", "x_states", "x_predict", "logp_by_species", "logp_theta",
               "logq_theta"]
_BWD_FIELDS = ["g_logp_by_species", "g_logp_theta", "g_logq_theta", "g_theta", "g_x_states", "g_x_predict", "d_q_mu",
               "d_q_prec", "d_extra", "d_weights"]


class vh_fwd_io(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _FWD_FIELDS]


class vh_bwd_io(C.Structure):
    _fields_ = [("fwd", vh_fwd_io)] + [(n, C.c_void_p) for n in _BWD_FIELDS] + [("iwae_cost", C.c_void_p), ("iwae_b_total", C.c_int),
                                                                                      ("outputs_cleared", C.c_int)]


class vh_encoder_desc(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "dtype", "B", "T", "n_signals", "n_filters", "filter_size", "pool_size", "n_hidden", "C", "D", "n_local", "n_gcond",
        "n_global", "n_const", "local_cond_treatments", "local_cond_devices", "gcond_cond_treatments", "gcond_cond_devices")]


class vh_encoder_io(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "observations", "inputs", "dev_1hot", "conv_w", "conv_b", "lin_w", "lin_b", "local_w", "local_b", "gcond_w",
        "global_free", "const_values", "q_mu", "q_prec", "pooled", "enc")]


class vh_encoder_grads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "d_q_mu", "d_q_prec", "g_conv_w", "g_conv_b", "g_lin_w", "g_lin_b", "g_local_w", "g_local_b", "g_gcond_w",
        "g_global_free", "d_pre")] + [("skip_lin_wgrad", C.c_int), ("dpool", C.c_void_p)]


class vh_lin_wgrad(C.Structure):
    _fields_ = [("d_pre", C.c_void_p), ("pooled", C.c_void_p), ("B", C.c_int), ("H", C.c_int), ("NLIN", C.c_int),
                ("offset", C.c_longlong)]


_lib = None


def load():
    """Load libvihds_b200.so (built in-tree by ``__graft_entry__.build()`` / ``make -C vihds_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "vihds_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU/PyTorch fallback for the ODE+ELBO hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.vh_abi_version.restype = C.c_int
    lib.vh_last_error.restype = C.c_char_p
    lib.vh_model_id.argtypes = [C.c_char_p]
    lib.vh_solver_id.argtypes = [C.c_char_p]
    lib.vh_num_slots.argtypes = [C.c_int]
    lib.vh_slot_name.argtypes = [C.c_int, C.c_int]
    lib.vh_slot_name.restype = C.c_char_p
    lib.vh_num_species.argtypes = [C.c_int]
    lib.vh_state_width.argtypes = [C.POINTER(vh_problem)]
    lib.vh_num_weights.argtypes = [C.POINTER(vh_problem)]
    lib.vh_num_weights.restype = C.c_size_t
    for name, io in (("vh_elbo_terms_fwd", vh_fwd_io), ("vh_elbo_terms_bwd", vh_bwd_io), ("vh_simulate", vh_fwd_io),
                     ("vh_simulate_bwd", vh_bwd_io)):
        fn = getattr(lib, name)
        fn.argtypes = [C.POINTER(vh_problem), C.POINTER(io), C.c_void_p]
        fn.restype = C.c_int
    lib.vh_iwae_fwd.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7
    lib.vh_iwae_bwd.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
    lib.vh_iw_moments.argtypes = [C.POINTER(vh_problem)] + [C.c_void_p] * 9
    lib.vh_adam_step.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                 C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p]
    lib.vh_adam_step_dev.argtypes = [C.c_int, C.c_size_t] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p, C.c_void_p]
    lib.vh_copy_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.vh_zero_async.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    lib.vh_peer_buffer_bytes.restype = C.c_size_t
    lib.vh_peer_buffer_bytes.argtypes = [C.c_int, C.c_size_t, C.c_int]
    lib.vh_peer_buffer_create.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]
    lib.vh_peer_buffer_open.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.vh_peer_buffer_close.argtypes = [C.c_void_p]
    lib.vh_peer_buffer_destroy.argtypes = [C.c_void_p]
    lib.vh_adam_allreduce_step.argtypes = ([C.c_int, C.c_size_t] + [C.c_void_p] * 7 +
                                           [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p])
    lib.vh_adam_allreduce_step_wgrad.argtypes = ([C.c_int, C.c_size_t] + [C.c_void_p] * 7 +
                                                 [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.POINTER(vh_lin_wgrad), C.c_void_p])
    lib.vh_iwae_fwd_bwd.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 10
    lib.vh_device_conditioner.argtypes = [C.c_int] * 7 + [C.c_void_p] * 6
    lib.vh_encoder_fwd.argtypes = [C.POINTER(vh_encoder_desc), C.POINTER(vh_encoder_io), C.c_void_p]
    lib.vh_encoder_bwd.argtypes = [C.POINTER(vh_encoder_desc), C.POINTER(vh_encoder_io), C.POINTER(vh_encoder_grads), C.c_void_p]
    lib.vh_encoder_bwd_adam.argtypes = ([C.POINTER(vh_encoder_desc), C.POINTER(vh_encoder_io), C.POINTER(vh_encoder_grads), C.c_size_t] +
                                        [C.c_void_p] * 8)
    if lib.vh_abi_version() != 4:
        raise RuntimeError("vihds_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(status):
    if status:
        raise RuntimeError("libvihds_b200: %s (status %d)" % (load().vh_last_error().decode(), status))


def model_id(name):
    mid = load().vh_model_id(name.encode())
    if mid < 0:
        raise NotImplementedError("vihds_b200: no kernel for model '%s'" % name)
    return mid


def solver_id(name):
    sid = load().vh_solver_id(name.encode())
    if sid < 0:
        raise NotImplementedError(
            "vihds_b200: solver '%s' has no fixed-step kernel (adaptive solvers are not supported; there is no CPU fallback)" % name)
    return sid


def slot_names(model):
    lib = load()
    return [lib.vh_slot_name(model, s).decode() for s in range(lib.vh_num_slots(model))]
