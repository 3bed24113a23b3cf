"""ctypes mirror of ``include/w2t_types.h`` / ``include/w2t.h``.

Only declarations live here: the structs, the status codes and the argument
types of every exported symbol.  ``_lib.py`` binds them to ``libw2t.so``.
"""
import ctypes as C
import os

W2T_MAX_CLASSES = 8

W2T_OK = 0
W2T_ERR_ARG = 1
W2T_ERR_CAPACITY = 2
W2T_ERR_CUDA = 3
W2T_ERR_NONFINITE = 4
W2T_ERR_UNSUPPORTED = 5
W2T_LAYOUT_ENSEMBLE, W2T_LAYOUT_STREAMS = 0, 1

STATUS_NAMES = {
    W2T_OK: "ok",
    W2T_ERR_ARG: "bad argument / unsupported input",
    W2T_ERR_CAPACITY: "plan capacity exceeded",
    W2T_ERR_CUDA: "CUDA error",
    W2T_ERR_NONFINITE: "tracker box became infinite",
    W2T_ERR_UNSUPPORTED: "input outside the fast host path",
}

_p = C.c_void_p


class SortPlan(C.Structure):
    _fields_ = [
        ("order", _p),
        ("track_cap", _p),
        ("det_cap", _p),
        ("ws_offset", _p),
        ("ws_bytes", C.c_int64),
        ("chunk_of", _p),
        ("chunk_done", _p),
        ("n_wide", C.c_int32),
        ("n_mid", C.c_int32),
        ("aux_offset", C.c_int64),
        ("narrow_cap", C.c_int32),
    ]


class SortProblem(C.Structure):
    _fields_ = [
        ("n_streams", C.c_int32),
        ("n_classes", C.c_int32),
        ("stream_img_offsets", _p),
        ("det_start", _p),
        ("det_count", _p),
        ("det_box", _p),
        ("img_exists", _p),
        ("cam_wh", _p),
        ("iou_thr", C.c_double * W2T_MAX_CLASSES),
        ("max_age", C.c_int32),
        ("min_hits", C.c_int32),
        ("promotion", C.c_int32),
    ]


class SortResult(C.Structure):
    _fields_ = [
        ("out_box", _p),
        ("out_score", _p),
        ("out_birth", _p),
        ("out_count", _p),
        ("created", _p),
        ("first_img", _p),
        ("final_count", _p),
        ("final_state", _p),
        ("final_cap", C.c_int32),
    ]


class Rows(C.Structure):
    _fields_ = [
        ("box", _p),
        ("score", _p),
        ("object_id", _p),
        ("image", _p),
        ("category", _p),
        ("totals", _p),
        ("id_base_device", _p),
        ("capacity", C.c_int64),
        ("image_base", C.c_int32),
        ("birth_group_base", C.c_int64),
        ("compact", _p),
    ]


class NmsProblem(C.Structure):
    _fields_ = [
        ("n_groups", C.c_int32),
        ("group_offsets", _p),
        ("rows", _p),
        ("iou_thresh", C.c_double),
        ("soft_nms_cut", C.c_double),
        ("min_score", C.c_double),
        ("n_classes", C.c_int32),
        ("score_thr", _p),
        ("box_format", C.c_int32),
        ("top_k", C.c_int32),
        ("conf_thresh", C.c_double),
        ("compute_f32", C.c_int32),
    ]


W2T_WIDE_DETS = 320          # include/w2t_types.h
W2T_NARROW_DETS = 96
W2T_PROMOTION_LEGACY, W2T_PROMOTION_NEP50 = 0, 1
# NumPy promotion regime the tracker reproduces unless a call says otherwise: "legacy" = NumPy 1.x value-based
# casting, the reference's pinned environment (python 3.7, /root/reference/environment.yml:7 — scikit-learn 0.22.2
# and filterpy do not even import under NumPy 2); "nep50" = NumPy 2.  Environment variable W2T_PROMOTION overrides.
DEFAULT_PROMOTION = os.environ.get("W2T_PROMOTION", "legacy")


def promotion_code(promotion=None):
    name = DEFAULT_PROMOTION if promotion is None else promotion
    if name in (W2T_PROMOTION_LEGACY, W2T_PROMOTION_NEP50):
        return int(name)
    try:
        return {"legacy": W2T_PROMOTION_LEGACY, "nep50": W2T_PROMOTION_NEP50}[str(name).lower()]
    except KeyError:
        raise ValueError("promotion must be 'legacy' or 'nep50', not %r" % (name,))


def sort_aux_bytes(n_substreams):
    """W2T_SORT_AUX_BYTES of include/w2t_types.h: queue area (rounded to 256) + the warp kernel's spill areas."""
    return (64 + 12 * int(n_substreams) + 255) // 256 * 256 + 160 * 8 * 65536
W2T_BOX_LTWH, W2T_BOX_CXCYWH, W2T_BOX_XYXY, W2T_BOX_LTWH_I16, W2T_BOX_LTWH_P64 = 0, 1, 2, 3, 4


class NmsResult(C.Structure):
    _fields_ = [
        ("merged", _p),
        ("src_index", _p),
        ("kept_count", _p),
        ("ens_count", _p),
        ("ens_box", _p),
        ("ens_score", _p),
        ("trk_count", _p),
        ("trk_box", _p),
        ("img_exists", _p),
    ]


# name -> (restype, argtypes); every symbol include/w2t.h declares
EXPORTS = {
    "w2t_version": (C.c_char_p, []),
    "w2t_last_error": (C.c_char_p, []),
    "w2t_clear_error": (None, []),
    "w2t_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "w2t_stream_wait_value32": (C.c_int, [_p, _p, C.c_int32]),
    "w2t_softnms_groups": (C.c_int, [C.POINTER(NmsProblem), C.POINTER(NmsResult), C.c_int, _p, _p]),
    "w2t_hardnms_groups": (C.c_int, [C.POINTER(NmsProblem), C.POINTER(NmsResult), C.c_int, _p, _p]),
    "w2t_softnms_max_group": (C.c_int, []),
    "w2t_fusion_groups": (C.c_int, [C.POINTER(NmsProblem), _p, C.c_int32, C.POINTER(NmsResult), C.c_int, _p, _p]),
    "w2t_fusion_max_group": (C.c_int, []),
    "w2t_sort_plan": (C.c_int, [C.c_int32, C.c_int32, _p, _p, _p, C.c_int32, C.POINTER(SortPlan)]),
    "w2t_sort_plan_offsets": (C.c_int, [C.c_int32, C.c_int32, _p, _p, C.c_int32, C.POINTER(SortPlan)]),
    "w2t_sort_track": (C.c_int, [C.POINTER(SortProblem), C.POINTER(SortPlan), C.POINTER(SortResult), _p, _p, _p]),
    "w2t_sort_step": (C.c_int, [C.POINTER(SortProblem), C.POINTER(SortPlan), C.POINTER(SortResult), _p, _p, C.c_int32,
                                _p, _p]),
    "w2t_sort_slab_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "w2t_pb_write_submission": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32, C.c_char_p, C.c_char_p, _p, C.c_int32,
                                          C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_int64, _p, C.c_int64,
                                          _p, _p, _p, _p, _p]),
    "w2t_sort_finalize_workspace": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64]),
    "w2t_sort_finalize": (C.c_int, [C.POINTER(SortProblem), C.POINTER(SortResult), _p, C.c_int64, C.c_int64, _p,
                                    C.POINTER(Rows), _p]),
    "w2t_assign_ids": (C.c_int, [C.c_int32, C.c_int32, _p, _p, _p, _p, _p, _p, _p, C.c_int64, _p,
                                 C.POINTER(C.c_int64)]),
    "w2t_iou_matrix": (C.c_int, [_p, C.c_int32, _p, C.c_int32, _p, _p]),
    "w2t_linear_assignment_workspace": (C.c_size_t, [C.c_int32, C.c_int32]),
    "w2t_linear_assignment": (C.c_int, [_p, C.c_int32, C.c_int32, _p, _p, _p, _p]),
    "w2t_kf_init": (C.c_int, [_p, _p, _p, C.c_int32, C.c_int32, _p]),
    "w2t_kf_predict": (C.c_int, [_p, _p, _p, C.c_int32, _p]),
    "w2t_kf_update": (C.c_int, [_p, _p, _p, _p, C.c_int32, C.c_int32, _p]),
    "w2t_json_load": (C.c_int, [C.c_char_p, C.POINTER(_p)]),
    "w2t_json_count": (C.c_int64, [_p]),
    "w2t_json_n_images": (C.c_int64, [_p]),
    "w2t_json_copy": (C.c_int, [_p, _p, _p, _p, _p, _p]),
    "w2t_json_image_ids": (_p, [_p, C.POINTER(C.c_int64)]),
    "w2t_json_free": (None, [_p]),
    "w2t_json_group_files": (C.c_int, [_p, C.c_int32, _p, C.c_double, C.c_int32, C.c_int32, C.POINTER(_p)]),
    "w2t_json_groups_info": (C.c_int, [_p, _p]),
    "w2t_json_groups_copy": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "w2t_json_groups_image_ids": (_p, [_p, C.POINTER(C.c_int64)]),
    "w2t_json_groups_free": (None, [_p]),
    "w2t_json_pack_tracks": (C.c_int, [C.c_char_p, _p, C.c_int32, C.c_int32, C.c_char_p, C.c_int32, C.c_int32,
                                       C.POINTER(_p)]),
    "w2t_json_tracks_info": (C.c_int, [_p, _p]),
    "w2t_json_tracks_copy": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p]),
    "w2t_json_tracks_free": (None, [_p]),
    "w2t_json_write_tracks": (C.c_int, [C.c_char_p, C.c_int64, _p, _p, _p, _p, _p, _p]),
    "w2t_json_write_detections": (C.c_int, [C.c_char_p, C.c_int64, _p, _p, _p, _p, _p]),
    "w2t_bbox_to_z": (C.c_int, [_p, _p, C.c_int32, C.c_int32, _p]),
    "w2t_bbox_vote": (C.c_int, [_p, C.c_int32, _p, _p, C.c_int32, C.c_double, C.c_int32, _p, _p]),
    "w2t_x_to_bbox": (C.c_int, [_p, C.c_int32, _p, C.c_int32, _p]),
}


def bind(lib, exports=EXPORTS):
    """Attach restype/argtypes; raises AttributeError if a symbol is missing."""
    for name, (restype, argtypes) in exports.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    return lib
