"""ctypes binding of the C ABI declared in include/fsp_b200.h (libpacmensl_b200.so).

The product path is the CUDA library; there is NO CPU fallback: importing this module raises if the
shared object is missing (build it with `make` or `python -c "import __graft_entry__ as g; g.build()"`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpacmensl_b200.so")

vp, ci, cl, cd = C.c_void_p, C.c_int, C.c_long, C.c_double
ip, dp, lp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_long)
vpp = C.POINTER(C.c_void_p)

CONSTR_FN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, C.c_void_p)
PROP_FN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.c_int, ip, dp, C.c_void_p)
TCOEF_FN = C.CFUNCTYPE(C.c_int, C.c_double, C.c_int, dp, C.c_void_p)


class FspMatDesc(C.Structure):
    _fields_ = [
        ("n_states", ci), ("n_rows", ci), ("n_reactions", ci), ("n_tv", ci), ("n_ti", ci),
        ("tv_reactions", ip), ("ti_reactions", ip),
        ("col", vp), ("off", vp), ("diag", vp), ("ld", cl),
        ("arrays_on_device", ci), ("n_constr", ci),
        ("sink_ptr", lp), ("sink_idx", vp), ("sink_val", vp),
        ("owns_sinks", ci), ("n_ghost", cl),
    ]


class FspMatEpilogue(C.Structure):
    _fields_ = [("alpha", cd), ("beta", cd), ("scale_dev", vp), ("n_dots", ci), ("dot_vec_dev", vp * 2), ("dot_out_dev", vp)]


# name -> (restype, argtypes); every symbol include/fsp_b200.h declares
SIGNATURES = {
    "fsp_device_count": (ci, [ip]),
    "fsp_device_set": (ci, [ci]),
    "fsp_device_get": (ci, [ip]),
    "fsp_device_sm_count": (ci, [ip]),
    "fsp_last_error": (C.c_char_p, []),
    "fsp_malloc": (ci, [vpp, C.c_size_t]),
    "fsp_free": (ci, [vp]),
    "fsp_malloc_host": (ci, [vpp, C.c_size_t]),
    "fsp_free_host": (ci, [vp]),
    "fsp_memcpy_h2d": (ci, [vp, vp, C.c_size_t, vp]),
    "fsp_memcpy_d2h": (ci, [vp, vp, C.c_size_t, vp]),
    "fsp_memcpy_d2d": (ci, [vp, vp, C.c_size_t, vp]),
    "fsp_memcpy_h2d_async": (ci, [vp, vp, C.c_size_t, vp]),
    "fsp_memcpy_d2h_async": (ci, [vp, vp, C.c_size_t, vp]),
    "fsp_memset": (ci, [vp, ci, C.c_size_t, vp]),
    "fsp_stream_create": (ci, [vpp]),
    "fsp_stream_destroy": (ci, [vp]),
    "fsp_stream_sync": (ci, [vp]),
    "fsp_device_sync": (ci, []),
    "fsp_event_create": (ci, [vpp]),
    "fsp_event_destroy": (ci, [vp]),
    "fsp_event_record": (ci, [vp, vp]),
    "fsp_stream_wait_event": (ci, [vp, vp]),
    "fsp_event_sync": (ci, [vp]),
    "fsp_event_elapsed_ms": (ci, [vp, vp, C.POINTER(C.c_float)]),
    "fsp_graph_begin_capture": (ci, [vp]),
    "fsp_graph_end_capture": (ci, [vp, vpp]),
    "fsp_graph_abort_capture": (ci, [vp]),
    "fsp_graph_launch": (ci, [vp, vp]),
    "fsp_graph_num_kernels": (ci, [vp, lp]),
    "fsp_graph_destroy": (ci, [vp]),
    "fsp_launch_count": (C.c_longlong, []),
    "fspvec_set": (ci, [vp, cd, cl, vp]),
    "fspvec_copy": (ci, [vp, vp, cl, vp]),
    "fspvec_scale": (ci, [vp, cd, cl, vp]),
    "fspvec_axpy": (ci, [vp, cd, vp, cl, vp]),
    "fspvec_linear_sum": (ci, [vp, cd, vp, cd, vp, cl, vp]),
    "fspvec_wlincomb": (ci, [vp, vp, cd, vp, cd, vp, cl, vp]),
    "fspvec_div": (ci, [vp, vp, vp, cl, vp]),
    "fspvec_prod": (ci, [vp, vp, vp, cl, vp]),
    "fspvec_lincomb3": (ci, [vp, cd, vp, cd, vp, cd, vp, cl, vp]),
    "fspvec_maxpy": (ci, [vp, cd, ci, dp, vpp, cl, vp]),
    "fspvec_mdot": (ci, [vp, vp, ci, vpp, cl, vp]),
    "fspvec_dot": (ci, [vp, vp, vp, cl, vp]),
    "fspvec_norm2sq": (ci, [vp, vp, cl, vp]),
    "fspvec_sum": (ci, [vp, vp, cl, vp]),
    "fspvec_norm1": (ci, [vp, vp, cl, vp]),
    "fspvec_wsqsum": (ci, [vp, vp, vp, cl, vp]),
    "fspvec_ewt": (ci, [vp, vp, cd, cd, cl, vp, vp]),
    "fspvec_ratio_absmax": (ci, [vp, vp, vp, cd, cd, cl, vp]),
    "fspvec_axpy_dot": (ci, [vp, vp, cd, vp, vp, vp, cl, vp]),
    "fspvec_scale_rsqrt": (ci, [vp, vp, cl, vp]),
    "fspvec_lincomb3_wprod_sqsum": (ci, [vp, vp, cd, vp, cd, vp, cd, vp, vp, vp, cl, vp]),
    "fspvec_scale_div": (ci, [vp, cd, vp, vp, cl, vp]),
    "fspvec_scale_mul": (ci, [vp, cd, vp, vp, cl, vp]),
    "fspvec_newton_update_mul": (ci, [vp, vp, vp, vp, vp, vp, cl, vp]),
    "fspvec_ewt_pair": (ci, [vp, vp, vp, cd, cd, cl, vp, vp]),
    "fspvec_newton_update": (ci, [vp, vp, vp, vp, vp, vp, vp, cl, vp]),
    "fspvec_nordsieck": (ci, [vp, ci, vp, ci, cl, vp]),
    "fspvec_multi_axpy": (ci, [vp, ci, vp, vp, cl, vp]),
    "fspvec_iop_orth": (ci, [vp, ci, vp, vp, vp, cl, vp]),
    "fspvec_marginal": (ci, [vp, ci, vp, vp, ci, ci, cl, vp]),
    "fspvec_max_species": (ci, [vp, vp, ci, ci, cl, vp]),
    "fspvec_clamp_min": (ci, [vp, vp, cd, cl, vp]),
    "fspvec_wdiv_dot": (ci, [vp, vp, vp, vp, cl, vp]),
    "fspvec_dot_h": (ci, [dp, vp, vp, cl, vp]),
    "fspvec_norm2_h": (ci, [dp, vp, cl, vp]),
    "fspvec_sum_h": (ci, [dp, vp, cl, vp]),
    "fspvec_norm1_h": (ci, [dp, vp, cl, vp]),
    "fspvec_scatter": (ci, [vp, cl, vp, vp, cl, vp]),
    "fspvec_gather": (ci, [vp, vp, vp, cl, vp]),
    "fspvec_route_by_owner": (ci, [vp, vp, cl, lp, ci, vp, vp, lp, vp]),
    "fspvec_scatter_range": (ci, [vp, cl, vp, vp, cl, cl, vp]),
    "fspset_create": (ci, [vpp, ci, ci, ip]),
    "fspset_destroy": (ci, [vp]),
    "fspset_set_shape": (ci, [vp, ci, vp, ip, vp]),
    "fspset_set_bounds": (ci, [vp, ci, ip]),
    "fspset_add_states": (ci, [vp, ci, cl, vp, ci]),
    "fspset_expand": (ci, [vp]),
    "fspset_num_states": (ci, [vp, ip]),
    "fspset_state2index": (ci, [vp, cl, vp, ci, vp, ci]),
    "fspset_lookup_shifted": (ci, [vp, ip, ci, cl, cl, vp]),
    "fspset_check_constraints_shifted": (ci, [vp, ip, cl, cl, vp]),
    "fspset_sink_lists": (ci, [vp, ip, cl, cl, vp, cl, lp]),
    "fspset_num_boundary_states": (ci, [vp, cl, cl, lp]),
    "fspset_states_dev": (ci, [vp, vpp]),
    "fspset_copy_states": (ci, [vp, cl, cl, ip]),
    "fspset_copy_status": (ci, [vp, cl, cl, C.POINTER(C.c_byte)]),
    "fspset_add_box_lattice": (ci, [vp, ip]),
    "fspset_set_sharded": (ci, [vp, vp]),
    "fspset_is_sharded": (ci, [vp]),
    "fspset_layout": (ci, [vp, lp, lp]),
    "fspset_rebalance_plan": (ci, [ci, lp, ci, lp, lp, lp, lp, lp, ip, ip]),
    "fspset_remember_local": (ci, [vp]),
    "fspset_remembered_indices": (ci, [vp, ip, cl]),
    "fspset_eval_mass_action": (ci, [vp, cd, ip, ip, ci, cl, cl, vp]),
    "fspset_eval_separable": (ci, [vp, cd, ip, vp, ip, ip, ip, ci, cl, cl, vp]),
    "fspmat_create": (ci, [vpp]),
    "fspmat_destroy": (ci, [vp]),
    "fspmat_generate": (ci, [vp, C.POINTER(FspMatDesc)]),
    "fspmat_clear": (ci, [vp]),
    "fspmat_action": (ci, [vp, dp, vp, vp, vp, vp, vp]),
    "fspmat_action_phase": (ci, [vp, dp, vp, vp, vp, vp, ci, vp]),
    "fspmat_action_rows": (ci, [vp, dp, vp, vp, cl, cl, ci, vp]),
    "fspmat_chunk_max_columns": (ci, [vp, cl, ci, ip]),
    "fspmat_num_boundary_rows": (ci, [vp, lp]),
    "fspmat_fused_supported": (ci, [vp]),
    "fspmat_action_fused": (ci, [vp, dp, vp, vp, vp, vp]),
    "fspmat_action_sinks_p2p": (ci, [vp, dp, vp, vp, vp]),
    "fspmat_action_boundary_p2p": (ci, [vp, dp, vp, vp, vp, vp]),
    "fspmat_action_p2p": (ci, [vp, dp, vp, vp, vp, vp]),
    "fspmat_p2p_cta_counts": (ci, [vp, lp, lp]),
    "fspmat_flops": (ci, [vp, lp]),
    "fspmat_num_rows": (ci, [vp, ip]),
    "fspmat_action_bytes": (ci, [vp, dp]),
    "fspmat_set_ti_coef": (ci, [vp, cd]),
    "fspmat_set_variant": (ci, [vp, ci]),
    "fspmat_dense": (ci, [vp, dp, dp]),
    "fspmat_csr_size": (ci, [vp, lp, ip]),
    "fspmat_csr_export": (ci, [vp, dp, ci, vp, vp, vp, vp]),
    "fspmat_csr_spmv": (ci, [ci, vp, vp, vp, vp, vp, vp]),
    "fspmat_build_ghosts": (ci, [vp, cl, ci, ci, vpp, lp]),
    "fspmat_shift_indices": (ci, [vp, cl, ci]),
    "fspcomm_unique_id": (ci, [C.c_char_p]),
    "fspcomm_create": (ci, [vpp, C.c_char_p, ci, ci]),
    "fspcomm_destroy": (ci, [vp]),
    "fspcomm_rank": (ci, [vp, ip, ip]),
    "fspcomm_allreduce_sum": (ci, [vp, vp, cl, vp]),
    "fspcomm_allreduce_max": (ci, [vp, vp, cl, vp]),
    "fspcomm_reduce_sum": (ci, [vp, vp, cl, ci, vp]),
    "fspcomm_allgather_int": (ci, [vp, vp, vp, cl, vp]),
    "fspcomm_allgather_f64": (ci, [vp, vp, vp, cl, vp]),
    "fspcomm_halo_exchange": (ci, [vp, vp, lp, vp, lp, vp]),
    "fspcomm_alltoall_counts": (ci, [vp, lp, lp, vp]),
    "fspcomm_exchange_int": (ci, [vp, vp, lp, vp, lp, vp]),
    "fspcomm_p2p_enabled": (ci, [vp]),
    "fsphalo_create": (ci, [vp, vpp, vp, lp, lp, ci]),
    "fsphalo_destroy": (ci, [vp]),
    "fsphalo_begin": (ci, [vp, vp, vp, vp]),
    "fsphalo_next": (ci, [vp, vp, vp]),
    "fsphalo_check": (ci, [vp]),
    "fspcomm_check": (ci, [vp]),
    "fspcomm_alive": (ci, [vp]),
    "fspcomm_window_create": (ci, [vp, C.c_size_t, vpp]),
    "fspcomm_window_destroy": (ci, [vp, vpp]),
    "fspcomm_window_retire": (ci, [vp, vpp, C.c_size_t]),
    "fspcomm_alltoallv": (ci, [vp, vp, lp, vp, lp, ci, vp]),
    "fspcomm_barrier": (ci, [vp, vp]),
    "fspcomm_barrier_sync": (ci, [vp]),
    "fspcomm_gather_long": (ci, [vp, cl, lp]),
    "fspmat_action_halo": (ci, [vp, dp, vp, vp, vp, vp, vp]),
    "fspmat_halo_fused_supported": (ci, [vp]),
    "fspmat_action_halo_part": (ci, [vp, dp, vp, vp, vp, vp, ci, cl, cl, ci, vp, vp]),
    "fspmat_chunk_has_ghost": (ci, [vp, cl, ci, ip]),
}

_LIB = None


def lib():
    """Load libpacmensl_b200.so; raises (no fallback) if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "pacmensl_b200: %s is missing -- the CUDA library must be built (run `make` in the repo "
                "root); there is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


class FspError(RuntimeError):
    pass


def check(ierr, what=""):
    if ierr != 0:
        msg = lib().fsp_last_error()
        raise FspError("%s failed (%d): %s" % (what, ierr, msg.decode() if msg else ""))
