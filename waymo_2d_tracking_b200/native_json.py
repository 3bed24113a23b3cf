"""Native JSON reader / writer (``csrc/json_io.cpp`` behind ``w2t_json_*``) for the two schemas on
either side of the path: submission / annotation files in, ensemble rows / tracker rows out.

At test scale a file holds 15 M detections; ``json.load`` plus the per-dict loops of
``convert_submission`` / ``read_data_file`` then cost minutes, and so does ``json.dump`` of 12 M
row dicts.  These functions go straight between the files and flat NumPy arrays; the files
written are byte-identical to what the reference's ``json.dump`` produces.
"""
import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _abi
from ._lib import check, lib


@dataclass
class Detections:
    """One submission / annotation file as flat arrays (rows in file order)."""
    image_ids: List[str]        # distinct image ids, first-appearance order
    image_index: np.ndarray     # [n] int32 into image_ids
    category: np.ndarray        # [n] int32
    bbox: np.ndarray            # [n,4] float64 x, y, w, h
    score: np.ndarray           # [n] float64 (1.0 where the row had none)
    has_score: np.ndarray       # [n] uint8

    def __len__(self):
        return len(self.score)


def load(path) -> Detections:
    """Parse a submission (list) or annotation file (``{'annotations': [...]}``)."""
    handle = C.c_void_p()
    check(lib().w2t_json_load(os.fsencode(str(path)), C.byref(handle)), "w2t_json_load")
    try:
        n = int(lib().w2t_json_count(handle))
        out = Detections([], np.empty(n, np.int32), np.empty(n, np.int32), np.empty((n, 4), np.float64),
                         np.empty(n, np.float64), np.empty(n, np.uint8))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib().w2t_json_copy(handle, vp(out.image_index), vp(out.category), vp(out.bbox), vp(out.score),
                                  vp(out.has_score)), "w2t_json_copy")
        nbytes = C.c_int64(0)
        ptr = lib().w2t_json_image_ids(handle, C.byref(nbytes))
        text = C.string_at(ptr, nbytes.value).decode("utf-8") if nbytes.value else ""
        out.image_ids = text.split("\n")[:-1] if text else []
        assert len(out.image_ids) == int(lib().w2t_json_n_images(handle))
        return out
    finally:
        lib().w2t_json_free(handle)


def _id_table(image_ids):
    encoded = [s.encode("utf-8") for s in image_ids]
    table = (C.c_char_p * max(len(encoded), 1))(*encoded)
    return table, encoded


def write_tracks(path, image_ids, image, bbox, score, category, object_id):
    """``json.dump`` of the tracker rows (tracking/utils.py:52-58): ``image`` indexes ``image_ids``."""
    table, keep = _id_table(image_ids)
    image = np.ascontiguousarray(image, np.int32)
    bbox = np.ascontiguousarray(bbox, np.float64).reshape(-1, 4)
    score = np.ascontiguousarray(score, np.float64)
    category = np.ascontiguousarray(category, np.int32)
    object_id = np.ascontiguousarray(object_id, np.int64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib().w2t_json_write_tracks(os.fsencode(str(path)), len(image), C.cast(table, C.c_void_p), vp(image), vp(bbox),
                                      vp(score), vp(category), vp(object_id)), "w2t_json_write_tracks")


def write_detections(path, image_ids, image, category, bbox, score):
    """``json.dump`` of the ensemble rows (detnet/ensemble.py:61-62): int boxes, 5-decimal scores."""
    table, keep = _id_table(image_ids)
    image = np.ascontiguousarray(image, np.int32)
    category = np.ascontiguousarray(category, np.int32)
    bbox = np.ascontiguousarray(bbox, np.int32).reshape(-1, 4)
    score = np.ascontiguousarray(score, np.float64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib().w2t_json_write_detections(os.fsencode(str(path)), len(image), C.cast(table, C.c_void_p), vp(image),
                                          vp(category), vp(bbox), vp(score)), "w2t_json_write_detections")


@dataclass
class FileGroups:
    """``w2t_json_group_files``: the submissions of an ensemble parsed and grouped by (image, category) natively."""
    image_ids: List[str]            # images that keep a row, sorted
    category_ids: List[int]         # every category id of the files, ascending
    image_order: np.ndarray         # [n_img] layout position -> index into image_ids
    group_offsets: np.ndarray       # [n_img * columns + 1] int32
    sub_counts: np.ndarray          # [n_img * columns, files] int32
    rows: Optional[np.ndarray]      # [N,5] float64 (None: not asked for, see ``packed``)
    packed: Optional[np.ndarray]    # [N] uint64 (``packing.packed_rows`` format) or None
    max_group: int
    columns: int
    stream_img_offsets: Optional[np.ndarray] = None   # W2T_LAYOUT_STREAMS only
    frame_ids: Optional[np.ndarray] = None


def group_files(paths, weights, min_score, layout=_abi.W2T_LAYOUT_ENSEMBLE, n_classes=0, want_rows=True):
    """Parse and group submission files in one native call; ``None`` when the input needs the general packer
    (``packing.pack_detection_files``: odd frame spellings, categories outside the tracker's list, ...).
    ``want_rows=False`` skips the float64 rows when the 8-byte packed rows hold them exactly."""
    encoded = [os.fsencode(str(q)) for q in paths]
    table = (C.c_char_p * len(encoded))(*encoded)
    w = np.ascontiguousarray(weights, np.float64)
    handle = C.c_void_p()
    status = lib().w2t_json_group_files(C.cast(table, C.c_void_p), len(encoded), w.ctypes.data_as(C.c_void_p),
                                        float(min_score), int(layout), int(n_classes), C.byref(handle))
    if status == _abi.W2T_ERR_UNSUPPORTED:
        lib().w2t_clear_error()
        return None
    check(status, "w2t_json_group_files")
    try:
        info = (C.c_int64 * 8)()
        check(lib().w2t_json_groups_info(handle, info), "w2t_json_groups_info")
        n_img, cols, n_rows, max_group, n_streams, packable, n_cat, n_files = (int(v) for v in info)
        G = n_img * cols
        streams = layout == _abi.W2T_LAYOUT_STREAMS
        out = FileGroups([], [], np.empty(n_img, np.int32), np.empty(G + 1, np.int32), np.empty((G, n_files), np.int32),
                         np.empty((n_rows, 5), np.float64) if (want_rows or not packable) else None,
                         np.empty(n_rows, np.uint64) if packable else None,
                         max_group, cols, np.empty(n_streams + 1, np.int32) if streams else None,
                         np.empty(n_img, np.int64) if streams else None)
        cats = np.empty(n_cat, np.int32)
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        check(lib().w2t_json_groups_copy(handle, vp(cats), vp(out.image_order), vp(out.stream_img_offsets),
                                         vp(out.frame_ids), vp(out.group_offsets), vp(out.sub_counts), vp(out.rows),
                                         vp(out.packed)), "w2t_json_groups_copy")
        out.category_ids = [int(c) for c in cats]
        nbytes = C.c_int64(0)
        ptr = lib().w2t_json_groups_image_ids(handle, C.byref(nbytes))
        text = C.string_at(ptr, nbytes.value).decode("utf-8") if nbytes.value else ""
        out.image_ids = text.split("\n")[:-1] if text else []
        return out
    finally:
        lib().w2t_json_groups_free(handle)


def pack_tracks(path, score_threshold, n_classes, segment_id=None, segment_block=None):
    """``w2t_json_pack_tracks``: parse one detection file and lay it out for the tracker in one native call.
    Returns a dict of arrays (``streams``: list of (segment, camera)), or ``None`` when the input needs the general
    packer (``load`` + ``packing.pack_detections``: odd frame spellings, categories outside the threshold list, more
    than W2T_MAX_CLASSES categories)."""
    thr = np.ascontiguousarray(score_threshold, np.float64)
    if not 1 <= int(n_classes) <= _abi.W2T_MAX_CLASSES:
        return None
    rank, world = (0, 0) if segment_block is None else (int(segment_block[0]), int(segment_block[1]))
    handle = C.c_void_p()
    status = lib().w2t_json_pack_tracks(os.fsencode(str(path)), thr.ctypes.data_as(C.c_void_p), len(thr), int(n_classes),
                                        None if segment_id is None else str(segment_id).encode("utf-8"), rank, world,
                                        C.byref(handle))
    if status == _abi.W2T_ERR_UNSUPPORTED:
        lib().w2t_clear_error()
        return None
    check(status, "w2t_json_pack_tracks")
    try:
        info = (C.c_int64 * 4)()
        check(lib().w2t_json_tracks_info(handle, info), "w2t_json_tracks_info")
        S, n_img, n_rows, name_bytes = (int(v) for v in info)
        G = n_img * int(n_classes)
        out = {"stream_img_offsets": np.empty(S + 1, np.int32), "frame_ids": np.empty(n_img, np.int64),
               "det_start": np.empty(G, np.int32), "det_count": np.empty(G, np.int32),
               "det_box": np.empty((n_rows, 4), np.float32), "class_rank": np.empty(S * int(n_classes), np.int32)}
        names = C.create_string_buffer(max(name_bytes, 1))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib().w2t_json_tracks_copy(handle, vp(out["stream_img_offsets"]), vp(out["frame_ids"]), vp(out["det_start"]),
                                         vp(out["det_count"]), vp(out["det_box"]), vp(out["class_rank"]),
                                         C.cast(names, C.c_void_p)), "w2t_json_tracks_copy")
        text = names.raw[:name_bytes].decode("utf-8")
        out["streams"] = [tuple(line.split("\t")) for line in text.split("\n")[:-1]] if text else []
        return out
    finally:
        lib().w2t_json_tracks_free(handle)
