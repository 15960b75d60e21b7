"""ctypes binding of include/hdpgpc_b200.h.  No CPU fallback: a missing library is an error."""
import ctypes
import os

from .build import LIBPATH

_c = ctypes
_p = _c.c_void_p
_i64 = _c.c_int64
_int = _c.c_int
_dbl = _c.c_double

# name -> (restype, argtypes); must list every symbol declared in include/hdpgpc_b200.h
PROTOTYPES = {
    "hgp_version": (_int, []),
    "hgp_build_info": (_c.c_char_p, []),
    "hgp_last_error": (_c.c_char_p, []),
    "hgp_launch_count": (_i64, []),
    "hgp_pack_leads": (_int, [_p, _i64, _int, _int, _p, _p]),
    "hgp_pack_leads_slice": (_int, [_p, _i64, _int, _int, _p, _i64, _p]),
    "hgp_chol_batched": (_int, [_p, _i64, _int, _p, _dbl, _p, _p, _p, _p]),
    "hgp_tri_inverse_batched": (_int, [_p, _i64, _int, _p, _p]),
    "hgp_cholinv_batched": (_int, [_p, _i64, _int, _p, _dbl, _p, _p, _p, _p, _p]),
    "hgp_packed_factor_bytes": (_i64, [_int]),
    "hgp_pack_factors": (_int, [_p, _i64, _int, _p, _p]),
    "hgp_tile_beats": (_int, []),
    "hgp_tile_uniform_states": (_int, [_p, _i64, _int, _p, _p]),
    "hgp_whiten_means": (_int, [_p, _p, _p, _i64, _int, _p, _p]),
    "hgp_whiten_means_tiles": (_int, [_p, _p, _p, _p, _i64, _int, _p, _i64, _p, _i64, _p, _p]),
    "hgp_score_tiles": (_int, [_p, _i64, _int, _p, _p, _p, _p, _p, _int, _p, _p, _p, _p, _p]),
    "hgp_score_blocks": (_int, [_p, _i64, _int, _p, _p, _p, _p, _int, _p, _p]),
    "hgp_score_pairs": (_int, [_p, _i64, _int, _p, _p, _p, _p, _int, _p, _p, _i64, _p, _p]),
    "hgp_score_groups_max_pairs": (_int, []),
    "hgp_score_groups": (_int, [_p, _i64, _int, _p, _p, _p, _p, _int, _p, _p, _p, _i64, _int, _p, _p]),
    "hgp_snr_states": (_int, [_p, _i64, _int, _p, _p, _int, _p, _p]),
    "hgp_mean_beat_work_doubles": (_i64, [_int]),
    "hgp_mean_beat": (_int, [_p, _i64, _int, _p, _p, _p]),
    "hgp_lead_weights": (_int, [_p, _p, _p, _i64, _int, _int, _p, _p, _p, _p, _p]),
    "hgp_hmm_workspace_bytes": (_i64, [_i64, _int]),
    "hgp_hmm_smooth": (_int, [_p, _i64, _int, _p, _p, _p, _p, _p, _int, _int, _p, _p, _p, _p, _p, _p, _p, _i64,
                              _c.POINTER(_int), _p]),
    "hgp_hmm_resmooth": (_int, [_p, _i64, _int, _p, _p, _p, _p, _p, _int, _int, _p, _p, _p, _p, _p, _p, _p, _i64,
                              _c.POINTER(_int), _p]),
    "hgp_suffstats_workspace_bytes": (_i64, [_i64, _int]),
    "hgp_suffstats": (_int, [_p, _p, _p, _i64, _int, _int, _p, _p, _p, _p, _p, _i64, _p]),
    "hgp_qlat_workspace_bytes": (_i64, [_i64, _int]),
    "hgp_qlat_batched": (_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _int, _p, _p, _p, _i64, _p]),
    "hgp_gemm_batched": (_int, [_p, _p, _p, _p, _p, _i64, _int, _int, _int, _p]),
    "hgp_chain_desc_bytes": (_i64, []),
    "hgp_chain_work_doubles": (_i64, [_int]),
    "hgp_chain_rts_cache_doubles": (_i64, [_int, _int]),
    "hgp_chain_small_path": (_int, [_int]),
    "hgp_chain_run": (_int, [_p, _int, _int, _p]),
    "hgp_chain_pipeline_ctas": (_int, []),
    "hgp_chain_run_ex": (_int, [_p, _int, _int, _int, _p]),
    "hgp_la_op": (_int, [_int, _p, _p, _p, _p, _int, _p, _p]),
    "hgp_pred_dist_work_doubles": (_i64, [_i64, _int, _int]),
    "hgp_pred_dist_inducing": (_int, [_p, _int, _p, _i64, _int, _p, _p, _p, _p, _i64, _dbl, _dbl, _dbl, _p, _p, _p, _p, _p]),
    "hgp_emission_means": (_int, [_p, _p, _p, _p, _i64, _int, _p, _p]),
    "hgp_warp_fit_batched": (_int, [_p, _int, _p, _i64, _p, _int, _p, _int, _int, _dbl, _dbl, _dbl, _dbl, _p, _p, _p, _p,
                                    _p, _p]),
    "hgp_rbf_kernel_matrix": (_int, [_p, _int, _p, _int, _dbl, _dbl, _dbl, _p, _p]),
    "hgp_hyperfit_work_doubles": (_i64, [_int, _int]),
    "hgp_hyperfit_batched": (_int, [_p, _p, _int, _int, _dbl, _dbl, _dbl, _int, _int, _dbl, _p, _p, _p]),
    "hgp_mniw_workspace_bytes": (_i64, [_i64, _int]),
    "hgp_mniw_loglik_batched": (_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _int, _p, _p, _p, _i64, _p]),
    "hgp_warp_prior_cov": (_int, [_p, _int, _dbl, _dbl, _dbl, _int, _p, _p]),
}


class HgpError(RuntimeError):
    pass


_lib = None


def load():
    """Load libhdpgpc_b200.so (building is the job of __graft_entry__.build / hdpgpc_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("HGP_LIB", LIBPATH)     # HGP_LIB: A/B-test another build of the same ABI
    if not os.path.exists(path):
        raise HgpError(f"{path} is missing: run `python -m hdpgpc_b200.build` (needs nvcc). "
                       "hdpgpc_b200 has no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if int(lib.hgp_chain_desc_bytes()) != ctypes.sizeof(ChainDesc):
        raise HgpError(f"hgp_chain_desc is {int(lib.hgp_chain_desc_bytes())} bytes in the library but "
                       f"{ctypes.sizeof(ChainDesc)} in the ctypes mirror (_lib.ChainDesc): header and binding are out of step")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().hgp_last_error().decode()
        raise HgpError(f"{what} failed (rc={rc}): {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise HgpError("hdpgpc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


class ChainDesc(ctypes.Structure):
    """Mirror of hgp_chain_desc (include/hdpgpc_b200.h)."""
    _fields_ = ([("n_members", _int), ("first_is_prior", _int), ("annealing", _int), ("estimation_limit", _int),
                 ("r_first", _dbl), ("member_beats", _p), ("Y", _p)] +
                [(n, _p) for n in ("f_star", "f_star_sm", "cov_f", "cov_f_sm", "A", "Gamma", "C", "Sigma",
                                   "int_m_mean", "int_m_r_cov", "int_scale", "int_n0",
                                   "obs_m_mean", "obs_m_r_cov", "obs_scale", "obs_n0", "work", "piv", "status")] +
                [("start_members", _int), ("start_params", _int), ("phases", _int), ("reserved_", _int),
                 ("rts_cache", _p)])
