"""ctypes binding of include/gvm_b200.h (no torch types cross this boundary)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib_path():
    return os.path.join(_HERE, "libgvmb200.so")


class gvm_config(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int64), ("DELTAX", C.c_double), ("DELTAY", C.c_double),
                ("nu_0", C.c_float), ("eta", C.c_float), ("minpix", C.c_float),
                ("noise_cut", C.c_float), ("threshold", C.c_float), ("fg_scale", C.c_float),
                ("device", C.c_int), ("grad_mode", C.c_int), ("keep_vm", C.c_int)]


class gvm_channel_desc(C.Structure):
    _fields_ = [("freq", C.c_float), ("antenna_diameter", C.c_float), ("pb_factor", C.c_float),
                ("pb_cutoff", C.c_float), ("primary_beam", C.c_int),
                ("ref_xobs_pix", C.c_float), ("ref_yobs_pix", C.c_float),
                ("phs_xobs_pix", C.c_float), ("phs_yobs_pix", C.c_float)]


class gvm_prior_params(C.Structure):
    _fields_ = [("prior_value", C.c_float), ("eta", C.c_float), ("epsilon", C.c_float),
                ("epsilon_b", C.c_float), ("prior_image_dev", C.c_void_p)]


class gvm_taper(C.Structure):
    _fields_ = [("enabled", C.c_int), ("sigma_maj", C.c_float), ("sigma_min", C.c_float),
                ("bpa", C.c_float), ("amplitude", C.c_float), ("u_0", C.c_double), ("v_0", C.c_double)]


# every symbol include/gvm_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "gvm_last_error": (C.c_char_p, []),
    "gvm_version": (C.c_int, []),
    "gvm_create": (C.c_int, [C.POINTER(gvm_config), C.POINTER(_P)]),
    "gvm_destroy": (C.c_int, [_P]),
    "gvm_set_stream": (C.c_int, [_P, _P]),
    "gvm_get_stream": (_P, [_P]),
    "gvm_synchronize": (C.c_int, [_P]),
    "gvm_set_scalars": (C.c_int, [_P, C.c_float, C.c_float, C.c_float]),
    "gvm_set_grad_mode": (C.c_int, [_P, C.c_int]),
    "gvm_set_flag_opt": (C.c_int, [_P, C.c_int]),
    "gvm_set_noise_image": (C.c_int, [_P, _P, C.c_int]),
    "gvm_build_noise_image": (C.c_int, [_P, C.c_float, C.POINTER(C.c_float)]),
    "gvm_build_noise_image_fields": (C.c_int, [_P, C.c_float, C.c_int, _P, C.POINTER(C.c_float)]),
    "gvm_set_block_nvis": (C.c_int, [_P, C.c_int, C.c_int64]),
    "gvm_get_noise_image": (C.c_int, [_P, _P]),
    "gvm_set_gcf": (C.c_int, [_P, _P]),
    "gvm_set_degrid_kernel": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gvm_get_model_grid": (C.c_int, [_P, _P]),
    "gvm_set_forward_mode": (C.c_int, [_P, C.c_int]),
    "gvm_last_forward_mode": (C.c_int, [_P]),
    "gvm_add_channel": (C.c_int, [_P, C.POINTER(gvm_channel_desc), C.c_int64, _P, _P, _P, C.POINTER(C.c_int)]),
    "gvm_clear_channels": (C.c_int, [_P]),
    "gvm_num_channels": (C.c_int, [_P]),
    "gvm_channel_nvis": (C.c_int64, [_P, C.c_int]),
    "gvm_get_vis": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P]),
    "gvm_chi2": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_float)]),
    "gvm_chi2_async": (C.c_int, [_P, _P, C.c_int, _P]),
    "gvm_dchi2": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "gvm_error_maps": (C.c_int, [_P, _P, C.c_int, _P]),
    "gvm_eval_host": (C.c_int, [_P, _P, C.c_int, C.c_int, C.POINTER(C.c_float), _P]),
    "gvm_prior_value": (C.c_int, [_P, C.c_int, _P, C.c_int, C.POINTER(gvm_prior_params), C.POINTER(C.c_float)]),
    "gvm_chi2_to_slot": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "gvm_prior_value_to_slot": (C.c_int, [_P, C.c_int, _P, C.c_int, C.POINTER(gvm_prior_params), C.c_int]),
    "gvm_fetch_slots": (C.c_int, [_P, C.c_int, _P]),
    "gvm_prior_grad": (C.c_int, [_P, C.c_int, _P, C.c_int, C.POINTER(gvm_prior_params), C.c_float, _P]),
    "gvm_add_to_dphi": (C.c_int, [_P, _P, _P, C.c_int]),
    "gvm_vec_evaluate_xt": (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_int, C.c_int]),
    "gvm_vec_new_p": (C.c_int, [_P, _P, _P, C.c_float, C.c_int, C.c_int]),
    "gvm_vec_dot": (C.c_int, [_P, _P, _P, C.c_int64, C.POINTER(C.c_float)]),
    "gvm_vec_gg_dgg": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "gvm_vec_grad_condition": (C.c_int, [_P, _P, _P, C.c_float, C.c_int, C.POINTER(C.c_float)]),
    "gvm_vec_new_xi": (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_int]),
    "gvm_vec_axpby": (C.c_int, [_P, C.c_float, _P, C.c_float, _P, C.c_int64]),
    "gvm_vec_absmax": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_float)]),
    "gvm_vec_scale": (C.c_int, [_P, _P, C.c_float, C.c_int64]),
    "gvm_vec_lbfgs_sy": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int64]),
    "gvm_dev_alloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "gvm_dev_free": (C.c_int, [_P, _P]),
    "gvm_dev_memset": (C.c_int, [_P, _P, C.c_int, C.c_size_t]),
    "gvm_dev_copy": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_int]),
    "gvm_dist_unique_id": (C.c_int, [C.c_char_p, C.c_size_t]),
    "gvm_dist_init": (C.c_int, [_P, C.c_int, C.c_int, C.c_char_p, C.c_size_t]),
    "gvm_dist_rank": (C.c_int, [_P]),
    "gvm_dist_world": (C.c_int, [_P]),
    "gvm_dist_allreduce": (C.c_int, [_P, _P, C.c_int64]),
    "gvm_dist_collectives": (C.c_int64, [_P]),
    "gvm_fetch_slots_enqueue": (C.c_int, [_P, C.c_int]),
    "gvm_fetch_slots_wait": (C.c_int, [_P, C.c_int, _P]),
    "gvm_graph_begin": (C.c_int, [_P]),
    "gvm_graph_end": (C.c_int, [_P, C.POINTER(_P)]),
    "gvm_graph_launch": (C.c_int, [_P, _P]),
    "gvm_graph_destroy": (C.c_int, [_P, _P]),
    "gvm_state_epoch": (C.c_int64, [_P]),
    "gvm_weights_dist": (C.c_int, [_P, C.c_int, C.c_float, C.c_int, _P, _P, _P, _P, _P]),
    "gvm_grid_block_dist": (C.c_int, [_P, C.c_float, C.c_int64, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "gvm_prior_grad_add": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, C.c_float, _P, C.c_int]),
    "gvm_sort_pairs_host": (C.c_int, [C.c_int, _P, _P, C.c_int64, C.c_int]),
    "gvm_dist_abort": (C.c_int, [_P]),
    "gvm_dist_set_replicated": (C.c_int, [_P, C.c_int]),
    "gvm_dist_broadcast": (C.c_int, [_P, _P, C.c_int64, C.c_int]),
    "gvm_weights": (C.c_int, [C.c_int, C.c_int, C.c_float, C.c_int64, C.c_int64, C.c_double, C.c_double,
                              C.c_int, _P, _P, _P, _P, C.POINTER(gvm_taper)]),
    "gvm_grid_block": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_float,
                                 C.c_int64, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                                 _P, _P, _P, C.POINTER(C.c_int64)]),
    "gvm_grid_fetch": (C.c_int, [_P, _P, _P]),
    "gvm_grid_release": (C.c_int, []),
    "gvm_grid_reserve": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "gvm_launch_count": (C.c_int64, [_P]),
    "gvm_last_grad_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "gvm_last_grad_mode": (C.c_int, [_P]),
    "gvm_grad_plan": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
}


def load_library():
    """Load libgvmb200.so and attach the prototypes. Raises if it is not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build the CUDA library first (python -c \"import __graft_entry__ as g; "
            "g.build()\"). gpuvmem_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib
