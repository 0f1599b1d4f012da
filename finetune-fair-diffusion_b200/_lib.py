"""ctypes binding of csrc/libfairguide.so (the C ABI declared in include/fairguide.h).

There is no fallback: if the library is missing the import of any op raises, and every
non-zero return code becomes a RuntimeError (the only place this package raises).
"""
import collections
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("FG_LIB", os.path.join(CSRC, "libfairguide.so"))   # FG_LIB: kernel-tuning builds only

F32, BF16, F16 = 0, 1, 2

_p = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float
_d = ctypes.c_double
_z = ctypes.c_size_t
_i64 = ctypes.c_int64

# name -> (restype, argtypes); mirrors include/fairguide.h one to one
SIGNATURES = {
    "fg_abi_version": (_i, []),
    "fg_error_string": (ctypes.c_char_p, [_i]),
    "fg_select_expand_boxes": (_i, [_p, _p, _i, _i, _i, _d, _d, _i64, _p, _p, _p]),
    "fg_crop_resize_fwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _i, _i, _p, _i, _i, _f, _i, _p]),
    "fg_guidance_factors": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "fg_image_grad": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "fg_region_scale": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "fg_head_workspace_bytes": (_z, [_i, _i, _i, _i, _i]),
    "fg_head_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _z, _i, _p]),
    "fg_head_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _z, _i, _p]),
    "fg_head_attributes": (_i, [_p, _i, _i, _p, _p, _i, _i, _p, _p, _f, _p, _p, _p, _i, _p]),
    "fg_head_attributes_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _i, _p]),
    "fg_fair_ce_fwd": (_i, [_p, _p, _p, _i, _i, _f, _p, _i, _p]),
    "fg_fair_ce_bwd": (_i, [_p, _p, _p, _p, _i, _i, _p, _i, _p]),
    "fg_fair_loss_fused": (_i, [_p, _p, _p, _p, _i, _p, _i, _i, _f, _f, _p, _p, _p, _p, _f, _f, _p, _p, _p, _i, _p]),
    "fg_rank_binom_workspace_bytes": (_z, [_i]),
    "fg_assign_rank_binom": (_i, [_p, _i, _d, _f, _p, _p, _p, _z, _i, _p]),
    "fg_ot_workspace_bytes": (_z, [_i, _i, _i]),
    "fg_ot_plan_counts": (_i, [_p, _p, _p, _i, _p, _p, _p, _i, _i, _p, _p, _z, _i, _p]),
    "fg_ot_targets": (_i, [_p, _p, _p, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p, _p, _z, _i, _p]),
    "fg_ot_solve_single": (_i, [_p, _i, _i, _p, _p, _p, _z, _p]),
    "fg_ot_cost_matrix": (_i, [_p, _p, _p, _i, _i, _p, _p, _z, _i, _p]),
    "fg_stage_detector_input": (_i, [_p, _i, _i, _i, _i, _p, _i, _p]),
    "fg_bias_metrics": (_i, [_p, _p, _p, _i, _p, _i, _p]),
    "fg_race_workspace_bytes": (_z, [_i, _i]),
    "fg_assign_race_enumerated": (_i, [_p, _i, _i, _p, _p, _i, _f, _p, _p, _p, _z, _i, _p]),
    "fg_race_cost_matrix": (_i, [_p, _i, _i, _p, _p, _z, _i, _p]),
    "fg_align_params_bytes": (_z, [_i]),
    "fg_align_matrices": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "fg_aligned_warp_fwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _i, _i, _f, _i, _p]),
    "fg_aligned_warp_bwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _i, _i, _i, _i, _p]),
    "fg_feats_normalize_fwd": (_i, [_p, _i, _i, _p, _p, _i, _p]),
    "fg_feats_normalize_bwd": (_i, [_p, _p, _p, _i, _i, _p, _i, _p]),
    "fg_face_search_workspace_bytes": (_z, [_i]),
    "fg_face_search_top1": (_i, [_p, _p, _i, _p, _i, _i, _p, _p, _p, _z, _p]),
    "fg_face_search_tc_workspace_bytes": (_z, [_i, _i]),
    "fg_face_search_top1_tc": (_i, [_p, _p, _i, _p, _i, _i, _f, _p, _p, _p, _z, _p]),
    "fg_face_loss_workspace_bytes": (_z, [_i, _i]),
    "fg_face_loss_fwd": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _i, _f, _i, _f, _p, _p, _z, _i, _p]),
    "fg_face_loss_bwd": (_i, [_p, _p, _p, _i, _i, _p, _p, _z, _i, _p]),
    "fg_face_loss_target_rows": (_i, [_p, _i, _i, _p, _p]),
    "fg_grad_bucket_pack": (_i, [_p, _p, _i, _i64, _p, _p, _i, _p]),
    "fg_grad_bucket_unpack": (_i, [_p, _p, _i, _i64, _p, _d, _d, _i, _p]),
    "fg_peer_epoch_advance": (_i, [_p, _p]),
    "fg_peer_push": (_i, [_p, _z, _p, _z, _z, _z, _i, _z, _i, _i, _i, _p, _p, _p]),
    "fg_peer_wait_copy": (_i, [_p, _i, _i, _z, _i, _z, _z, _p, _p, _z, _p, _p]),
    "fg_peer_wait_sum": (_i, [_p, _i, _i, _z, _i, _z, _z, _p, _p, _z, _p, _p]),
    "fg_peer_push_rows": (_i, [_p, _p, _p, _i, _i, _p, _z, _z, _z, _z, _i, _i, _p, _p, _i, _p]),
    "fg_peer_wait_unpack": (_i, [_p, _i, _i, _z, _z, _z, _p, _p, _p, _p, _i, _i, _p, _i, _p]),
}

_lib = None


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into csrc/libfairguide.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libfairguide.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(fairguide has no CPU or PyTorch fallback)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)           # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        if l.fg_abi_version() != 1:
            raise RuntimeError("libfairguide.so ABI version mismatch")
        _lib = l
    return _lib


CALLS = collections.Counter()      # C-ABI calls made so far, by entry point (bench.py reports launches from it)

# kernels launched per successful call (fg_ot_plan_counts adds one per coarse-to-fine level, see ot_levels)
KERNELS_PER_CALL = {
    "fg_select_expand_boxes": 1, "fg_crop_resize_fwd": 1, "fg_guidance_factors": 1, "fg_image_grad": 1,
    "fg_region_scale": 1, "fg_head_fwd": 2, "fg_head_bwd": 2, "fg_head_attributes": 1, "fg_head_attributes_bwd": 1, "fg_fair_ce_fwd": 1,
    "fg_fair_ce_bwd": 1, "fg_fair_loss_fused": 1, "fg_assign_rank_binom": 2, "fg_ot_plan_counts": 3, "fg_ot_targets": 1,
    "fg_ot_solve_single": 2, "fg_ot_cost_matrix": 2, "fg_stage_detector_input": 1, "fg_bias_metrics": 1,
    "fg_assign_race_enumerated": 4, "fg_race_cost_matrix": 2, "fg_align_matrices": 1, "fg_aligned_warp_fwd": 1,
    "fg_aligned_warp_bwd": 1, "fg_feats_normalize_fwd": 1, "fg_feats_normalize_bwd": 1, "fg_face_search_top1": 2, "fg_face_search_top1_tc": 6,
    "fg_face_loss_fwd": 4, "fg_face_loss_bwd": 1, "fg_face_loss_target_rows": 0,
    "fg_grad_bucket_pack": 2, "fg_grad_bucket_unpack": 1,
    "fg_peer_epoch_advance": 1, "fg_peer_push": 1, "fg_peer_wait_copy": 1, "fg_peer_wait_sum": 1, "fg_peer_push_rows": 1, "fg_peer_wait_unpack": 1,
}


def ot_levels(n_valid):
    """Extra solver launches of fg_ot_plan_counts besides compact / cost+hist / draws (none: every draw searches
    from zero prices in its own CTA)."""
    return 0


def check(rc, what):
    CALLS[what] += 1
    if rc != 0:
        msg = lib().fg_error_string(rc)
        raise RuntimeError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
