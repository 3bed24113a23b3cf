"""Seeded synthetic Waymo-shaped detection streams.

No dataset or checkpoint is reachable, so benchmarks and parity tests run on
synthetic submissions with the reference's shapes: segments x 5 cameras x 200
frames, ``image_id = 'segment/timestamp/camera'`` (tracking/utils.py:70), int
pixel ``[x, y, w, h]`` boxes and 5-decimal scores (detnet/data/coco.py:249-250),
category mix vehicle/pedestrian/sign/cyclist = .60/.25/.05/.10.  Scores are made
distinct inside every (image, category) across all submissions, because the
reference's sort order among equal scores is implementation-defined
(SURVEY.md §8c).

Everything is vectorised NumPy and produces flat arrays; ``to_json_list``
converts a (small) submission to the reference's list-of-dicts form.
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .packing import IMAGE_SIZES, PackedGroups, PackedTracks

CAMERAS = ['FRONT', 'FRONT_LEFT', 'FRONT_RIGHT', 'SIDE_LEFT', 'SIDE_RIGHT']
N_CLASSES = 4
DEFAULT_SCORE_THR = (0.95, 0.6, 1.0, 0.9)     # README.md:54 / track.py:23-24
DEFAULT_IOU_THR = (0.01, 0.01, 1.0, 0.0)      # track.py:25-26


@dataclass
class SynthConfig:
    n_segments: int = 1
    cameras: Sequence[str] = tuple(CAMERAS)
    n_frames: int = 200
    objects_per_frame: float = 110.0     # mean live objects per camera frame
    class_mix: Sequence[float] = (0.60, 0.25, 0.05, 0.10)
    mean_life: float = 40.0
    p_miss: float = 0.1
    jitter: float = 1.5
    fp_per_frame: float = 6.0            # Poisson mean of false positives per frame and submission
    n_submissions: int = 3
    sub_jitter: float = 3.0
    sub_drop: float = 0.1
    score_noise: float = 0.05
    confident: float = 0.88              # share of objects scored above their class threshold
    size_range: Sequence[float] = (20.0, 200.0)
    seed: int = 0


def preset(name, **overrides):
    """Configurations of BASELINE.json (C1..C5), full size unless overridden."""
    presets = {
        # SORT on one segment, ~100 dets/frame surviving the thresholds
        "c1": dict(n_segments=1, n_submissions=1, objects_per_frame=125.0),
        # soft-NMS ensemble of 3 submissions on one segment
        "c2": dict(n_segments=1, n_submissions=3),
        # full test scale: 150 segments
        "c3": dict(n_segments=150, n_submissions=3),
        # crowded scene: ~1000 dets/frame, ~300 live tracks per class
        "c4": dict(n_segments=1, n_submissions=1, objects_per_frame=1100.0, mean_life=80.0,
                   size_range=(12.0, 90.0), fp_per_frame=20.0),
        # 5-way TTA ensemble, ~3000 boxes/frame
        "c5": dict(n_segments=1, n_submissions=5, objects_per_frame=620.0, fp_per_frame=40.0,
                   size_range=(12.0, 120.0)),
    }
    kw = dict(presets[name])
    kw.update(overrides)
    return SynthConfig(**kw)


@dataclass
class Submission:
    """One detector output file as flat arrays (rows in JSON order)."""
    image_index: np.ndarray   # [n] int32
    category: np.ndarray      # [n] int32 (1-based)
    bbox: np.ndarray          # [n,4] int32 x, y, w, h
    score: np.ndarray         # [n] float64, 5 decimals


@dataclass
class Scene:
    cfg: SynthConfig
    segments: List[str]
    cameras: List[str]
    frame_ids: np.ndarray             # [n_img] int64 timestamps
    stream_img_offsets: np.ndarray    # [S+1] int32 (stream = segment-major, then camera)
    submissions: List[Submission]

    @property
    def n_streams(self):
        return len(self.segments) * len(self.cameras)

    @property
    def n_img(self):
        return int(self.stream_img_offsets[-1])

    def streams(self):
        return [(seg, cam) for seg in self.segments for cam in self.cameras]

    def cam_wh(self):
        wh = np.asarray([IMAGE_SIZES[c] for c in self.cameras], np.float64)
        return np.tile(wh, (len(self.segments), 1))

    def image_id(self, img):
        s = int(np.searchsorted(self.stream_img_offsets, img, side='right') - 1)
        seg, cam = self.segments[s // len(self.cameras)], self.cameras[s % len(self.cameras)]
        return '%s/%i/%s' % (seg, self.frame_ids[img], cam)

    def image_ids(self):
        out = []
        nc = len(self.cameras)
        for s in range(self.n_streams):
            seg, cam = self.segments[s // nc], self.cameras[s % nc]
            for img in range(int(self.stream_img_offsets[s]), int(self.stream_img_offsets[s + 1])):
                out.append('%s/%i/%s' % (seg, self.frame_ids[img], cam))
        return out


def _segment_name(rng):
    a = int(rng.integers(10 ** 18, 9 * 10 ** 18))
    t = int(rng.integers(0, 500)) * 20
    return '%d_%d_000_%d_000' % (a, t, t + 20)


def make_scene(cfg: SynthConfig) -> Scene:
    rng = np.random.default_rng(cfg.seed)
    cams = list(cfg.cameras)
    S, F = cfg.n_segments * len(cams), cfg.n_frames
    segments = [_segment_name(rng) for _ in range(cfg.n_segments)]
    t0 = rng.integers(15 * 10 ** 14, 16 * 10 ** 14, size=cfg.n_segments)
    frame_ids = (np.repeat(t0, len(cams))[:, None] + 100000 * np.arange(F)[None, :]).reshape(-1).astype(np.int64)
    offsets = (np.arange(S + 1) * F).astype(np.int32)
    wh = np.tile(np.asarray([IMAGE_SIZES[c] for c in cams], np.float64), (cfg.n_segments, 1))

    # ---- persistent objects -------------------------------------------------
    L = cfg.mean_life
    n_obj_per_stream = int(np.ceil(cfg.objects_per_frame * (F + L) / L))
    n_obj = S * n_obj_per_stream
    o_stream = np.repeat(np.arange(S), n_obj_per_stream)
    birth = rng.integers(-int(L), F, size=n_obj)
    life = rng.geometric(1.0 / L, size=n_obj)
    cls = rng.choice(N_CLASSES, size=n_obj, p=np.asarray(cfg.class_mix) / np.sum(cfg.class_mix))
    ow = rng.uniform(cfg.size_range[0], cfg.size_range[1], n_obj)
    oh = rng.uniform(cfg.size_range[0], cfg.size_range[1], n_obj)
    ocx = rng.uniform(0, 1, n_obj) * wh[o_stream, 0]
    ocy = rng.uniform(0, 1, n_obj) * wh[o_stream, 1]
    ovx = rng.normal(0, 4.0, n_obj)
    ovy = rng.normal(0, 2.0, n_obj)
    thr = np.asarray(DEFAULT_SCORE_THR)[cls]
    confident = rng.uniform(0, 1, n_obj) < cfg.confident
    hi = np.where(thr >= 1.0, 1.0, rng.uniform(np.minimum(thr + 0.004, 0.9999), 0.99999))
    lo = rng.uniform(0.05, np.maximum(thr - 0.02, 0.06))
    quality = np.where(confident, hi, lo)

    first = np.maximum(birth, 0)
    last = np.minimum(birth + life, F)
    span = np.maximum(last - first, 0)
    p_obj = np.repeat(np.arange(n_obj), span)
    starts = np.cumsum(span) - span
    p_frame = (np.arange(int(span.sum())) - np.repeat(starts, span)) + np.repeat(first, span)
    age = p_frame - birth[p_obj]
    tcx = ocx[p_obj] + ovx[p_obj] * age
    tcy = ocy[p_obj] + ovy[p_obj] * age
    n_pair = len(p_obj)
    p_img = o_stream[p_obj] * F + p_frame

    subs = []
    for k in range(cfg.n_submissions):
        keep = rng.uniform(0, 1, n_pair) >= (1.0 - (1.0 - cfg.p_miss) * (1.0 - (cfg.sub_drop if cfg.n_submissions > 1 else 0.0)))
        idx = np.nonzero(keep)[0]
        sd = np.hypot(cfg.jitter, cfg.sub_jitter if cfg.n_submissions > 1 else 0.0)
        cx = tcx[idx] + rng.normal(0, sd, len(idx))
        cy = tcy[idx] + rng.normal(0, sd, len(idx))
        w = np.maximum(ow[p_obj[idx]] + rng.normal(0, sd, len(idx)), 2.0)
        h = np.maximum(oh[p_obj[idx]] + rng.normal(0, sd, len(idx)), 2.0)
        q = quality[p_obj[idx]]
        noise = rng.normal(0, cfg.score_noise, len(idx))
        sc = np.where(q >= 1.0, np.where(noise > -0.05, 1.0, 0.97 + noise), q + np.where(confident[p_obj[idx]], np.abs(noise) * 0.2, noise))
        img = p_img[idx]
        cat = cls[p_obj[idx]] + 1
        # false positives
        n_fp = rng.poisson(cfg.fp_per_frame * S * F)
        f_img = rng.integers(0, S * F, n_fp)
        f_w = rng.uniform(cfg.size_range[0], cfg.size_range[1], n_fp)
        f_h = rng.uniform(cfg.size_range[0], cfg.size_range[1], n_fp)
        f_stream = f_img // F
        f_cx = rng.uniform(0, 1, n_fp) * wh[f_stream, 0]
        f_cy = rng.uniform(0, 1, n_fp) * wh[f_stream, 1]
        f_sc = rng.uniform(0.011, 0.4, n_fp)
        f_cat = rng.choice(N_CLASSES, size=n_fp, p=np.asarray(cfg.class_mix) / np.sum(cfg.class_mix)) + 1
        img = np.concatenate([img, f_img])
        cat = np.concatenate([cat, f_cat])
        cx = np.concatenate([cx, f_cx]); cy = np.concatenate([cy, f_cy])
        w = np.concatenate([w, f_w]); h = np.concatenate([h, f_h])
        sc = np.clip(np.concatenate([sc, f_sc]), 0.011, 1.0)
        bw = np.maximum(np.rint(w), 1).astype(np.int64)
        bh = np.maximum(np.rint(h), 1).astype(np.int64)
        bx = np.floor(cx - w / 2).astype(np.int64)
        by = np.floor(cy - h / 2).astype(np.int64)
        subs.append(dict(img=img.astype(np.int64), cat=cat.astype(np.int64),
                         bbox=np.stack([bx, by, bw, bh], 1), score_int=np.rint(sc * 1e5).astype(np.int64)))

    _make_scores_distinct(subs)
    out = []
    for d in subs:
        score = d['score_int'] / 1e5
        score = np.round(score, 5)
        order = np.lexsort((-d['score_int'], d['img']))          # detector order: image, then score desc
        out.append(Submission(d['img'][order].astype(np.int32), d['cat'][order].astype(np.int32),
                              d['bbox'][order].astype(np.int32), score[order]))
    return Scene(cfg, segments, cams, frame_ids, offsets, out)


def _make_scores_distinct(subs):
    """Strictly decreasing integer scores inside each (image, category) over all submissions."""
    img = np.concatenate([d['img'] for d in subs])
    cat = np.concatenate([d['cat'] for d in subs])
    s = np.concatenate([d['score_int'] for d in subs])
    group = img * (N_CLASSES + 1) + cat
    order = np.lexsort((-s, group))
    g_sorted, s_sorted = group[order], s[order]
    new_group = np.r_[True, g_sorted[1:] != g_sorted[:-1]]
    gid = np.cumsum(new_group) - 1
    start = np.nonzero(new_group)[0]
    j = np.arange(len(s_sorted)) - start[gid]
    BIG = 10 ** 7
    v = s_sorted + j - gid * BIG
    v = np.minimum.accumulate(v) + gid * BIG - j
    assert v.min() >= 1, "score collision pushed a score to <= 0; lower the density or raise scores"
    s_new = np.empty_like(s)
    s_new[order] = v
    pos = 0
    for d in subs:
        n = len(d['score_int'])
        d['score_int'] = s_new[pos:pos + n]
        pos += n


def to_json_list(scene: Scene, sub: Submission):
    """The reference's submission JSON (list of dicts, detnet/data/coco.py:229-252)."""
    ids = scene.image_ids()
    return [{'image_id': ids[int(i)], 'category_id': int(c), 'bbox': [int(v) for v in b], 'score': float(s)}
            for i, c, b, s in zip(sub.image_index, sub.category, sub.bbox, sub.score)]


# ---------------------------------------------------------------------------
# vectorised packers (what load_input_submissions / read_data_file + pack_* do, without dicts)
# ---------------------------------------------------------------------------

def groups_from_scene(scene: Scene, weights: Optional[Sequence[float]] = None, min_score: float = 0.0) -> PackedGroups:
    """convert_submission (ensemble.py:31-47) for every submission + grouping by
    (image, category): group g = img * 4 + (category - 1), rows in submission order then JSON order."""
    subs = scene.submissions
    if weights is None:
        weights = [1.0] * len(subs)
    top = max(weights)
    weights = [w / top for w in weights]
    rows, keys = [], []
    for k, (sub, w) in enumerate(zip(subs, weights)):
        sw = sub.score * w
        ok = (sub.bbox[:, 2] > 0) & (sub.bbox[:, 3] > 0) & (sw >= min_score)
        r = np.empty((int(ok.sum()), 5), np.float64)
        r[:, 0] = sw[ok]
        r[:, 1:] = sub.bbox[ok]
        rows.append(r)
        keys.append(sub.image_index[ok].astype(np.int64) * N_CLASSES + (sub.category[ok] - 1))
    per_file_keys = list(keys)
    rows = np.concatenate(rows) if rows else np.zeros((0, 5))
    keys = np.concatenate(keys) if keys else np.zeros(0, np.int64)
    order = np.argsort(keys, kind='stable')          # stable: keeps (submission, JSON) order inside a group
    G = scene.n_img * N_CLASSES
    counts = np.bincount(keys, minlength=G)
    offsets = np.zeros(G + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    return PackedGroups(image_ids=None, category_ids=[1, 2, 3, 4], group_offsets=offsets.astype(np.int32),
                        rows=np.ascontiguousarray(rows[order]), max_group=int(counts.max()) if G else 0,
                        sub_counts=np.stack([np.bincount(k, minlength=G) for k in per_file_keys], axis=1).astype(np.int32)
                        if per_file_keys else None)


def tracks_from_submission(scene: Scene, sub: Submission, score_thr=DEFAULT_SCORE_THR) -> PackedTracks:
    """read_data_file (utils.py:63-96) + the row building of track_sort (utils.py:32-35),
    vectorised: group g = img * 4 + (category - 1), rows in JSON order."""
    thr = np.asarray(score_thr, np.float64)
    n_img = scene.n_img
    exists = np.zeros(n_img, np.uint8)
    exists[sub.image_index] = 1                       # the frame entry is created before the filters
    ok = ~((sub.bbox[:, 2] < 1) | (sub.bbox[:, 3] < 1)) & ~(sub.score < thr[sub.category - 1])
    key = sub.image_index[ok].astype(np.int64) * N_CLASSES + (sub.category[ok] - 1)
    order = np.argsort(key, kind='stable')
    b = sub.bbox[ok][order].astype(np.int64)
    box = np.stack([b[:, 0], b[:, 1], b[:, 0] + b[:, 2], b[:, 1] + b[:, 3]], 1).astype(np.float32)
    counts = np.bincount(key, minlength=n_img * N_CLASSES).astype(np.int32)
    start = (np.cumsum(counts, dtype=np.int64) - counts).astype(np.int32)
    # first-appearance rank of categories: position of the first surviving row per (stream, category)
    S = scene.n_streams
    pos_first = np.full(S * N_CLASSES, np.iinfo(np.int64).max, np.int64)
    img_ok = sub.image_index[ok].astype(np.int64)
    stream = np.searchsorted(scene.stream_img_offsets, img_ok, side='right') - 1
    q = stream * N_CLASSES + (sub.category[ok] - 1)
    # key: (image, JSON position) -> images of a stream are in frame order
    poskey = img_ok * (len(sub.score) + 1) + np.nonzero(ok)[0]
    np.minimum.at(pos_first, q, poskey)
    rank = np.argsort(np.argsort(pos_first.reshape(S, N_CLASSES), axis=1, kind='stable'), axis=1, kind='stable')
    return PackedTracks(
        n_streams=S, n_classes=N_CLASSES, streams=scene.streams(), frame_ids=scene.frame_ids,
        stream_img_offsets=scene.stream_img_offsets, det_start=start, det_count=counts,
        det_box=np.ascontiguousarray(box), cam_wh=scene.cam_wh(), img_exists=exists,
        class_rank=rank.astype(np.int32).reshape(-1), n_rows=int(box.shape[0]))
