"""ctypes binding of ``libgleipnir_b200.so`` (C ABI: ``include/gleipnir_b200.h``).

The library is built in-tree by ``gleipnir_b200.build`` and loaded lazily.  There is no
CPU fallback anywhere in this package: if the library is missing, or no CUDA device is
usable, the first compute call raises.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_uint32, c_uint64,
                    c_void_p)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgleipnir_b200.so")

GL_MAX_DIM = 128
GL_MAX_TUNING_CYCLES = 8

GL_PRIOR_UNIFORM, GL_PRIOR_NORM, GL_PRIOR_LOGUNIFORM = 0, 1, 2
GL_MODEL_IID_GAUSSIAN, GL_MODEL_EGGCARTON, GL_MODEL_MIXTURE1D, GL_MODEL_DENSE_GAUSSIAN = 0, 1, 2, 3
GL_PROPOSAL_UNIFORM, GL_PROPOSAL_NORMAL = 0, 1
GL_TAG_INIT, GL_TAG_WIDTH, GL_TAG_PICK, GL_TAG_MOVE, GL_TAG_USER = 0, 1, 2, 3, 4


class GleipnirB200Error(RuntimeError):
    pass


class gl_prior(Structure):
    _fields_ = [("kind", c_int32), ("reserved", c_int32), ("p0", c_double), ("p1", c_double)]


class gl_ns_config(Structure):
    _fields_ = [("population_size", c_int64), ("batch_k", c_int32), ("iterations", c_int32),
                ("burn_in", c_int32), ("tuning_cycles", c_int32), ("proposal", c_int32), ("walker_lanes", c_int32),
                ("seed", c_uint64), ("dead_capacity", c_int64), ("row_offset", c_int64),
                ("global_population", c_int64)]


class gl_ns_buffers(Structure):
    _fields_ = [("live", c_void_p), ("logl", c_void_p), ("dead_x", c_void_p), ("dead_logl", c_void_p),
                ("dead_w", c_void_p), ("dead_idx", c_void_p), ("dead_step", c_void_p), ("workspace", c_void_p)]


class gl_ns_state(Structure):
    _fields_ = [("n_dead", c_int64), ("n_rows", c_int64), ("n_evals", c_int64), ("n_launches", c_int64),
                ("step", c_int32),
                ("finalized", c_int32), ("log_x", c_double), ("lstar", c_double), ("z_linear", c_double),
                ("h_linear", c_double), ("lse_m", c_double), ("lse_z", c_double), ("lse_h", c_double),
                ("log_z", c_double), ("information", c_double), ("widths", c_double * GL_MAX_DIM),
                ("shard_overflow", c_int32), ("reserved", c_int32)]


_P = c_void_p            # device / opaque pointers
_SIGNATURES = {
    "gl_abi_version": (c_int, []),
    "gl_last_error": (c_char_p, []),
    "gl_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "gl_prior_draw": (c_int, [POINTER(gl_prior), c_int32, c_uint64, c_uint32, c_int64, c_int64, _P, _P]),
    "gl_prior_ppf": (c_int, [POINTER(gl_prior), _P, _P, c_int64, _P]),
    "gl_prior_pdf": (c_int, [POINTER(gl_prior), _P, _P, c_int64, _P]),
    "gl_prior_logpdf": (c_int, [POINTER(gl_prior), _P, _P, c_int64, _P]),
    "gl_prior_transform": (c_int, [POINTER(gl_prior), c_int32, _P, _P, c_int64, _P]),
    "gl_prior_ppf_host": (c_int, [POINTER(gl_prior), POINTER(c_double), POINTER(c_double), c_int64]),
    "gl_prior_pdf_host": (c_int, [POINTER(gl_prior), POINTER(c_double), POINTER(c_double), c_int64]),
    "gl_prior_logpdf_host": (c_int, [POINTER(gl_prior), POINTER(c_double), POINTER(c_double), c_int64]),
    "gl_model_create_iid_gaussian": (c_int, [c_int32, c_double, c_double, POINTER(_P)]),
    "gl_model_create_eggcarton": (c_int, [c_int32, POINTER(_P)]),
    "gl_model_create_mixture1d": (c_int, [POINTER(c_double), c_int32, c_double, POINTER(_P)]),
    "gl_model_create_dense_gaussian": (c_int, [c_int32, POINTER(c_double), POINTER(c_double), c_double, POINTER(_P)]),
    "gl_model_create_external": (c_int, [c_int32, _P, _P, _P, c_int64, POINTER(_P)]),
    "gl_model_set_tensor_path": (c_int, [_P, c_int32]),
    "gl_model_dim": (c_int, [_P]),
    "gl_model_kind_of": (c_int, [_P]),
    "gl_model_destroy": (c_int, [_P]),
    "gl_loglike_eval": (c_int, [_P, _P, c_int64, _P, _P]),
    "gl_loglike_eval_host": (c_int, [_P, POINTER(c_double), c_int64, POINTER(c_double)]),
    "gl_select_workspace_bytes": (c_int64, [c_int64]),
    "gl_select_kworst": (c_int, [_P, c_int64, c_int32, _P, _P, _P, _P]),
    "gl_ns_workspace_bytes": (c_int64, [POINTER(gl_ns_config), c_int32]),
    "gl_ns_create": (c_int, [POINTER(gl_ns_config), POINTER(gl_prior), c_int32, _P, POINTER(gl_ns_buffers),
                             POINTER(_P)]),
    "gl_ns_destroy": (c_int, [_P]),
    "gl_ns_init": (c_int, [_P, c_int32, _P]),
    "gl_ns_advance": (c_int, [_P, c_int64, _P]),
    "gl_ns_kill": (c_int, [_P, c_int32, _P]),
    "gl_ns_refill": (c_int, [_P, _P]),
    "gl_ns_inject": (c_int, [_P, _P, _P, c_int32, _P]),
    "gl_ns_finalize": (c_int, [_P, _P]),
    "gl_ns_export_dead": (c_int, [_P, _P, _P]),
    "gl_ns_read_state": (c_int, [_P, POINTER(gl_ns_state), _P]),
    "gl_ns_shard_candidates": (c_int, [_P, c_int32, _P, _P]),
    "gl_ns_shard_commit": (c_int, [_P, _P, c_int32, c_int32, c_int32, _P]),
    "gl_ns_shard_commit_scratch": (c_int, [_P, _P, c_int32, c_int32, c_int32, _P, _P]),
    "gl_ns_shard_final_partial": (c_int, [_P, _P, _P]),
    "gl_ns_shard_finalize": (c_int, [_P, _P, c_int32, c_int32, _P]),
    "gl_comm_unique_id_bytes": (c_int, []),
    "gl_comm_unique_id": (c_int, [_P]),
    "gl_comm_create": (c_int, [c_int32, c_int32, _P, POINTER(_P)]),
    "gl_comm_destroy": (c_int, [_P]),
    "gl_comm_all_gather_f64": (c_int, [_P, _P, _P, c_int64, _P]),
    "gl_ns_advance_sharded": (c_int, [_P, _P, c_int64, c_int32, c_int64, c_double, _P]),
    "gl_ns_advance_many": (c_int, [_P, c_int32, _P, _P]),
    "gl_ns_advance_many_graphs": (c_int, [_P, c_int32, _P, _P]),
    "gl_ns_advance_graph": (c_int, [_P, c_int64, _P]),
    "gl_ns_step_sharded": (c_int, [_P, _P, c_int32, c_int32, c_int64, c_double, _P]),
    "gl_ns_shard_candidate_count": (c_int32, [c_int32, c_int32, c_int64, c_double]),
    "gl_ns_shard_commit_truncated": (c_int, [_P, _P, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "gl_ns_local_population": (c_int64, [_P]),
    "gl_comm_p2p_handle_bytes": (c_int, []),
    "gl_comm_p2p_alloc": (c_int, [_P, c_int64, POINTER(ctypes.c_uint8)]),
    "gl_comm_p2p_open": (c_int, [_P, POINTER(ctypes.c_uint8)]),
    "gl_comm_p2p_ready": (c_int, [_P]),
    "gl_comm_p2p_disable": (c_int, [_P]),
    "gl_microbench_dfma_latency": (c_int, [c_int32, _P, POINTER(c_double)]),
    "gl_export_gathered_table": (c_int, [_P, _P, c_int32, c_int64, c_int32, c_int32, c_int32, c_int64, _P, _P]),
    "gl_ns_force_smem_walker": (c_int, [_P, c_int32]),
    "gl_ns_finalize_sharded": (c_int, [_P, _P, _P]),
    "gl_ns_force_multipass_select": (c_int, [_P, c_int32]),
    "gl_ns_set_select_path": (c_int, [_P, c_int32]),
    "gl_ns_profile_enable": (c_int, [_P, c_int32]),
    "gl_ns_profile_read": (c_int, [_P, POINTER(c_double), POINTER(c_int64), POINTER(c_double), POINTER(c_int64)]),
    "gl_posterior_stats_bytes": (c_int64, []),
    "gl_posterior_weights": (c_int, [_P, _P, _P, c_int64, c_int32, c_double, _P, _P, POINTER(c_int64), _P]),
    "gl_posterior_partial_doubles": (c_int64, [c_int64, c_int32, c_int32]),
    "gl_posterior_histograms": (c_int, [_P, _P, c_int64, c_int32, c_int32, _P, _P, _P, _P, _P, _P]),
    "gl_sampler_call_host": (c_int, [POINTER(gl_prior), c_int32, _P, c_int32, c_int32, c_int32, c_int32, c_uint64, c_int32,
                                     POINTER(c_double), c_double, POINTER(c_double), c_int32, POINTER(c_double),
                                     POINTER(c_double)]),
    "gl_reference_exact_run_host": (c_int, [c_uint32, c_int32, c_int32, POINTER(c_double), POINTER(c_double), c_double,
                                            c_int32, c_int32, c_int32, c_int32, POINTER(c_double), POINTER(c_double),
                                            POINTER(c_double), POINTER(c_double)]),
    "gl_microbench_dmma": (c_int, [c_int32, c_int32, _P, POINTER(c_int64), _P]),
    "gl_microbench_dfma": (c_int, [c_int32, c_int32, _P, POINTER(c_int64), _P]),
    "gl_ns_run_host": (c_int, [POINTER(gl_ns_config), POINTER(gl_prior), c_int32, _P, c_int64, POINTER(c_double),
                               POINTER(gl_ns_state)]),
}

_lib = None


def exported_symbols():
    """Names this binding expects the shared library to export."""
    return sorted(_SIGNATURES)


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GleipnirB200Error(
            "%s is missing: build it with `python -m gleipnir_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.gl_abi_version() != 1:
        raise GleipnirB200Error("ABI version mismatch: library %d, binding 1" % lib.gl_abi_version())
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().gl_last_error()
        raise GleipnirB200Error("gleipnir_b200 call failed (%d): %s" % (rc, msg.decode() if msg else "?"))
    return rc


def as_double_ptr(arr):
    return arr.ctypes.data_as(POINTER(c_double))


def prior_array(table):
    """[(kind, p0, p1), ...] -> ctypes array of gl_prior."""
    arr = (gl_prior * len(table))()
    for i, (kind, p0, p1) in enumerate(table):
        arr[i].kind, arr[i].p0, arr[i].p1 = int(kind), float(p0), float(p1)
    return arr


def device_info():
    sm, maj, mnr = c_int(0), c_int(0), c_int(0)
    check(load().gl_device_info(byref(sm), byref(maj), byref(mnr)))
    return sm.value, maj.value, mnr.value
