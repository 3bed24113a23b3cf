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
from typing import List

import numpy as np

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
