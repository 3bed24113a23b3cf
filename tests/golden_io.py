"""Encoding of golden fixtures: reference inputs/outputs as exact NumPy arrays (npz)."""
import os

import numpy as np

from waymo_2d_tracking_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def scene_from_cfg(cfg_kwargs):
    return synth.make_scene(synth.SynthConfig(**cfg_kwargs))


def tracks_to_arrays(rows, image_ids):
    """list of tracker output dicts -> arrays (image index, category, bbox, score, object id)."""
    index = {iid: i for i, iid in enumerate(image_ids)}
    n = len(rows)
    out = dict(img=np.zeros(n, np.int32), cat=np.zeros(n, np.int32), bbox=np.zeros((n, 4), np.float64),
               score=np.zeros(n, np.float64), oid=np.zeros(n, np.int64))
    for i, r in enumerate(rows):
        out["img"][i] = index[r["image_id"]]
        out["cat"][i] = r["category_id"]
        out["bbox"][i] = [float(v) for v in r["bbox"]]
        out["score"][i] = float(r["score"])
        out["oid"][i] = int(r["object_id"])
    return out


def dets_to_arrays(rows, image_ids):
    """list of ensemble output dicts -> arrays."""
    index = {iid: i for i, iid in enumerate(image_ids)}
    n = len(rows)
    out = dict(img=np.zeros(n, np.int32), cat=np.zeros(n, np.int32), bbox=np.zeros((n, 4), np.int64),
               score=np.zeros(n, np.float64))
    for i, r in enumerate(rows):
        out["img"][i] = index[r["image_id"]]
        out["cat"][i] = r["category_id"]
        out["bbox"][i] = r["bbox"]
        out["score"][i] = float(r["score"])
    return out


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True))
