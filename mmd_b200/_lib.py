"""ctypes binding of libmmdk.so (include/mmdk.h).  There is NO fallback: if the CUDA library is missing or no
sm_100 device is visible, every compute entry point raises."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmdk.so")

MMDK_OK, MMDK_EINVAL, MMDK_ECUDA, MMDK_ENOMEM = 0, 1, 2, 3
MAX_LEVELS, MAX_HARD_ROWS, STATE_DIM, MAX_RANKS = 4, 4, 4, 8
UNET_FP32, UNET_F16X3, UNET_F16X3_LAYERS = 0, 1, 2
UNET_MODES = {"fp32": UNET_FP32, "f16x3": UNET_F16X3, "tc": UNET_F16X3, "f16x3_layers": UNET_F16X3_LAYERS}

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


class UnetConfig(C.Structure):
    _fields_ = [("state_dim", C.c_int), ("horizon", C.c_int), ("unet_input_dim", C.c_int), ("n_levels", C.c_int),
                ("dim_mults", C.c_int * MAX_LEVELS), ("time_emb_dim", C.c_int), ("n_diffusion_steps", C.c_int),
                ("self_attention", C.c_int)]


class GuideEnv(C.Structure):
    _fields_ = [("norm_min", C.c_float * 4), ("norm_range", C.c_float * 4), ("grid_dev", C.c_void_p),
                ("nx", C.c_int), ("ny", C.c_int), ("grid_lo", C.c_float * 2), ("grid_map_dim", C.c_float * 2),
                ("margin", C.c_float), ("ws_min", C.c_float * 2), ("ws_max", C.c_float * 2),
                ("w_collision", C.c_float), ("w_border", C.c_float), ("w_smooth", C.c_float),
                ("max_grad_norm", C.c_float), ("dt", C.c_float), ("gp_q11", C.c_float), ("gp_q12", C.c_float),
                ("gp_q22", C.c_float), ("coll_inv_sigma2", C.c_float)]


class Groups(C.Structure):
    _fields_ = [("n_groups", C.c_int), ("K", C.c_int), ("hard_vals_dev", C.c_void_p), ("hard_rows_dev", C.c_void_p),
                ("obj_ptr_dev", C.c_void_p), ("obj_weight_dev", C.c_void_p), ("bucket_ptr_dev", C.c_void_p),
                ("cons_dev", C.c_void_p), ("peers_dev", C.c_void_p), ("peer_self_dev", C.c_void_p),
                ("n_peers", C.c_int), ("peer_radius", C.c_float), ("peer_weight", C.c_float),
                ("peer_cell_start_dev", C.c_void_p), ("peer_sorted_dev", C.c_void_p), ("peer_grid", C.c_int),
                ("peer_grid_lo", C.c_float), ("peer_grid_inv_cell", C.c_float), ("peer_seq_dev", C.c_void_p)]


class StepScalars(C.Structure):
    _fields_ = [("do_posterior", C.c_int), ("sqrt_recip_alphas_cumprod", C.c_float),
                ("sqrt_recipm1_alphas_cumprod", C.c_float), ("posterior_mean_coef1", C.c_float),
                ("posterior_mean_coef2", C.c_float), ("clip_denoised", C.c_int), ("predict_epsilon", C.c_int),
                ("n_guide_steps", C.c_int), ("add_noise", C.c_int), ("model_std", C.c_float), ("noise_std", C.c_float),
                ("final_hard_conds", C.c_int)]


class PeerExchange(C.Structure):
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("row_offset", C.c_int), ("n_rows", C.c_int), ("rep_index", C.c_int),
                ("tables_dev", C.c_void_p * MAX_RANKS), ("flags_dev", C.c_void_p * MAX_RANKS), ("state_dev", C.c_void_p)]


class ChainDesc(C.Structure):
    _fields_ = [("n_steps", C.c_int), ("t_index", C.POINTER(C.c_int)), ("scalars", C.POINTER(StepScalars)),
                ("lockstep", C.c_int), ("rep_index", C.c_int), ("peers_local_dev", C.c_void_p),
                ("exchange", C.POINTER(PeerExchange))]


class EnsembleTile(C.Structure):
    _fields_ = [("net", C.c_void_p), ("unet_mode", C.c_int), ("env", C.POINTER(GuideEnv)), ("groups", C.POINTER(Groups)),
                ("scalars", C.POINTER(StepScalars)), ("x_dev", C.c_void_p), ("eps_dev", C.c_void_p), ("noise_dev", C.c_void_p),
                ("chain_out_dev", C.c_void_p), ("chain_init_dev", C.c_void_p)]


class CrossCond(C.Structure):
    _fields_ = [("m1", C.c_int), ("m2", C.c_int), ("ind1", C.c_int), ("ind2", C.c_int), ("row_lo", C.c_int), ("row_hi", C.c_int),
                ("rel", C.c_float * 4), ("bnd", C.c_float * 4)]


class EnsembleDesc(C.Structure):
    _fields_ = [("n_tiles", C.c_int), ("tiles", C.POINTER(EnsembleTile)), ("n_steps", C.c_int), ("t_index", C.POINTER(C.c_int)),
                ("n_cross", C.c_int), ("cross", C.POINTER(CrossCond))]


# every symbol declared in include/mmdk.h (checked by tests/test_abi.py)
EXPORTS = [
    "mmdk_last_error", "mmdk_device_info", "mmdk_unet_create", "mmdk_unet_destroy", "mmdk_unet_forward",
    "mmdk_unet_cond_row", "mmdk_unet_debug_tap", "mmdk_unet_debug_timeline", "mmdk_unet_debug_stamps", "mmdk_unet_debug_keep_activations", "mmdk_debug_mma_calibrate", "mmdk_guide_grad", "mmdk_ddpm_step", "mmdk_ddpm_step_publish", "mmdk_wait_peers", "mmdk_p2p_alloc", "mmdk_p2p_free", "mmdk_p2p_export",
    "mmdk_p2p_open", "mmdk_p2p_close", "mmdk_run_chain", "mmdk_run_chain_ensemble", "mmdk_publish_peers", "mmdk_build_peer_hash", "mmdk_cross_condition",
    "mmdk_q_sample", "mmdk_cell_index", "mmdk_check_rr_collisions", "mmdk_classify_trajs", "mmdk_unnormalize", "mmdk_get_conflicts", "mmdk_smooth_trajs",
]

_lib = None


class MMDKError(RuntimeError):
    pass


def load():
    """dlopen libmmdk.so (no GPU needed for this; used by the ABI test)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MMDKError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        f"(mmd_b200/csrc/build.sh). There is no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.mmdk_last_error.restype = C.c_char_p
    vp, i, f, i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
    lib.mmdk_device_info.argtypes = [i, c_int_p, c_int_p, c_int_p]
    lib.mmdk_unet_create.argtypes = [C.POINTER(UnetConfig), i, C.POINTER(C.c_char_p), C.POINTER(vp),
                                     C.POINTER(i64), vp, C.POINTER(vp)]
    lib.mmdk_unet_destroy.argtypes = [vp]
    lib.mmdk_unet_destroy.restype = None
    lib.mmdk_unet_forward.argtypes = [vp, i, vp, i, i, vp, vp]
    lib.mmdk_unet_cond_row.argtypes = [vp, i, vp, c_int_p, vp]
    lib.mmdk_unet_debug_tap.argtypes = [vp, i, vp, c_int_p, c_int_p, c_int_p, vp]
    lib.mmdk_unet_debug_timeline.argtypes = [vp, i, vp, vp]
    lib.mmdk_unet_debug_keep_activations.argtypes = [vp, i]
    lib.mmdk_unet_debug_stamps.argtypes = [vp, vp, i, i]
    lib.mmdk_debug_mma_calibrate.argtypes = [i, i, i, i, vp, vp]
    lib.mmdk_guide_grad.argtypes = [C.POINTER(GuideEnv), C.POINTER(Groups), i, vp, vp, vp, i, vp]
    lib.mmdk_ddpm_step.argtypes = [C.POINTER(GuideEnv), C.POINTER(Groups), C.POINTER(StepScalars), i, vp, vp, vp, vp, vp]
    lib.mmdk_run_chain.argtypes = [vp, i, C.POINTER(GuideEnv), C.POINTER(Groups), C.POINTER(ChainDesc), i, vp, vp, vp, vp, i, vp]
    lib.mmdk_run_chain_ensemble.argtypes = [C.POINTER(EnsembleDesc), i, i, vp]
    lib.mmdk_publish_peers.argtypes = [C.POINTER(GuideEnv), i, i, i, i, vp, vp, vp]
    lib.mmdk_build_peer_hash.argtypes = [vp, vp, i, i, i, f, f, vp, vp, vp]
    lib.mmdk_ddpm_step_publish.argtypes = [C.POINTER(GuideEnv), C.POINTER(Groups), C.POINTER(StepScalars), i, vp, vp, vp, vp,
                                           C.POINTER(PeerExchange), vp]
    lib.mmdk_wait_peers.argtypes = [C.POINTER(PeerExchange), vp]
    lib.mmdk_p2p_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    lib.mmdk_p2p_free.argtypes = [vp]
    lib.mmdk_p2p_export.argtypes = [vp, C.c_char_p]
    lib.mmdk_p2p_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.mmdk_p2p_close.argtypes = [vp]
    lib.mmdk_cross_condition.argtypes = [vp, vp, i, i, i, i, C.c_float * 4, C.c_float * 4, vp]
    lib.mmdk_q_sample.argtypes = [vp, vp, f, f, i64, vp, vp]
    lib.mmdk_cell_index.argtypes = [C.POINTER(GuideEnv), vp, i64, vp, vp]
    lib.mmdk_check_rr_collisions.argtypes = [vp, i64, i, f, vp, vp, vp]
    lib.mmdk_classify_trajs.argtypes = [C.POINTER(GuideEnv), vp, i, i, i, f, C.c_float * 2, C.c_float * 2, vp, vp, vp, vp]
    lib.mmdk_get_conflicts.argtypes = [vp, vp, i, i, i, i, i, f, vp, vp, vp, vp]
    lib.mmdk_smooth_trajs.argtypes = [vp, vp, i, i, i, vp, vp]
    lib.mmdk_unnormalize.argtypes = [C.POINTER(GuideEnv), vp, i64, i, vp, vp]
    _lib = lib
    return lib


_checked_device = False


def lib():
    """Library handle for compute calls: also insists on a visible sm_100 device."""
    global _checked_device
    l = load()
    if not _checked_device:
        if not torch.cuda.is_available():
            raise MMDKError("mmd_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no fallback")
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        check(l.mmdk_device_info(torch.cuda.current_device(), C.byref(sm), C.byref(ma), C.byref(mi)))
        _checked_device = True
    return l


def check(rc):
    if rc == MMDK_OK:
        return
    msg = load().mmdk_last_error().decode()
    if rc == MMDK_EINVAL:
        raise ValueError(msg)
    raise MMDKError(msg)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
