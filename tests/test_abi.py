"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol
include/w2t.h declares; host-only entry points (plan, id assignment) are exercised; compute
entry points must fail loudly without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from waymo_2d_tracking_b200 import _abi, _lib, packing, runtime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "w2t.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(w2t_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_abi.EXPORTS) == names          # the ctypes mirror binds exactly the header


def test_version_and_struct_sizes():
    lib = _lib.lib()
    assert b"sm_100a" in lib.w2t_version()
    assert ctypes.sizeof(_abi.SortProblem) == 8 + 6 * 8 + 8 * 8 + 16     # max_age, min_hits, promotion + padding
    assert ctypes.sizeof(_abi.SortPlan) == 10 * 8     # ... n_wide + n_mid share a word, aux_offset, narrow_cap + padding


def test_sort_plan_bounds_and_order():
    offs = np.array([0, 4, 8], np.int32)
    cnt = np.zeros((8, 2), np.int32)
    cnt[:4, 0] = [3, 5, 2, 7]
    cnt[4:, 1] = [1, 1, 9, 1]
    plan = runtime.make_plan(2, 2, offs, cnt.reshape(-1), None, max_age=1)
    # window = max_age + 2 = 3 images
    assert plan["track_cap"].tolist() == [14, 1, 1, 11]
    assert plan["det_cap"].tolist() == [7, 1, 1, 9]
    assert plan["order"].tolist()[:2] == [0, 3]
    assert plan["ws_offset"][0] == 0 and np.all(np.diff(plan["ws_offset"]) > 0) and plan["ws_bytes"] > 0
    assert plan["n_wide"] == 0
    # crowded sub-streams (more than W2T_WIDE_DETS detections in some image) lead the launch order
    cnt[5, 1] = 400
    cnt[1, 0] = 321
    cnt[2, 0] = 5000 // 100          # heavy by total work, but not crowded
    plan = runtime.make_plan(2, 2, offs, cnt.reshape(-1), None, max_age=1)
    assert plan["n_wide"] == 2 and plan["n_mid"] == 0 and sorted(plan["order"].tolist()[:2]) == [0, 3]
    # sub-streams above W2T_NARROW_DETS come next (CTAs), everything else is tracked by warps
    cnt[7, 1] = 0
    cnt[5, 1] = 200
    plan = runtime.make_plan(2, 2, offs, cnt.reshape(-1), None, max_age=1)
    assert plan["n_wide"] == 1 and plan["n_mid"] == 1 and plan["order"].tolist()[:2] == [0, 3]
    # the auxiliary area of the warp kernel follows the slabs
    assert plan["aux_offset"] > plan["ws_offset"][-1] and plan["ws_bytes"] >= plan["aux_offset"] + _abi.sort_aux_bytes(4)


def _plan_numpy(n_streams, NC, offs, cnt, exists, max_age):
    """Plain restatement of w2t_sort_plan's capacities and work (one sub-stream at a time)."""
    window = max_age + 2
    cnt = cnt.reshape(-1, NC)
    tcap, dcap, work = [], [], []
    for s in range(n_streams):
        imgs = [i for i in range(offs[s], offs[s + 1]) if exists is None or exists[i]]
        for c in range(NC):
            d = cnt[imgs, c].astype(np.int64) if imgs else np.zeros(0, np.int64)
            best = max([int(d[max(0, j - window + 1):j + 1].sum()) for j in range(len(d))], default=0)
            tcap.append(max(best, 1))
            dcap.append(max(int(d.max(initial=0)), 1))
            work.append(int((d * d + d).sum()))
    return np.array(tcap), np.array(dcap), np.array(work)


@pytest.mark.parametrize("n_streams,frames,max_age", [(12, 4000, 2), (3, 50, 0), (9, 37, 5)])
def test_sort_plan_threaded_pass_matches_restatement(n_streams, frames, max_age):
    """Large jobs (>= 40 000 images, >= 8 streams) are planned by four host threads, one pass per stream for
    all categories; small ones by the calling thread: same capacities, same heaviest-first order either way."""
    rng = np.random.default_rng(n_streams)
    NC = 4
    lens = rng.integers(frames // 2, frames + 1, n_streams)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    n_img = int(offs[-1])
    cnt = rng.poisson([55, 25, 5, 10], (n_img, NC)).astype(np.int32)
    cnt[rng.random((n_img, NC)) < 0.05] = 0
    exists = (rng.random(n_img) > 0.03).astype(np.uint8)
    plan = runtime.make_plan(n_streams, NC, offs, cnt.reshape(-1), exists, max_age)
    tcap, dcap, work = _plan_numpy(n_streams, NC, offs, cnt, exists, max_age)
    np.testing.assert_array_equal(plan["track_cap"], tcap)
    np.testing.assert_array_equal(plan["det_cap"], dcap)
    order = plan["order"]
    assert sorted(order.tolist()) == list(range(n_streams * NC)) and plan["n_wide"] == 0
    assert np.all(np.diff(work[order]) <= 0)                       # heaviest first ...
    ties = np.diff(work[order]) == 0
    assert np.all(np.diff(order)[ties] > 0)                        # ... ties in sub-stream order (stable)
    assert np.all(np.diff(plan["ws_offset"]) > 0) and plan["ws_bytes"] > plan["ws_offset"][-1]


def test_chunk_bounds_counts_and_fractions():
    offs = np.arange(0, 11, dtype=np.int32) * 3          # 10 streams x 3 images
    NC = 2
    sizes = np.tile([4, 1], 30)
    go = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    for n in (1, 3, 4, 10, 25):
        ch = runtime._chunk_bounds(offs, go, NC, n)
        assert ch[0][0] == 0 and ch[-1][1] == 10 and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
        assert len(ch) == min(n, 10) and all(b > a for a, b in ch)
    ch = runtime._chunk_bounds(offs, go, NC, [0.2, 0.3, 0.5])
    assert ch == [(0, 2), (2, 5), (5, 10)]


def test_assign_ids_c_matches_numpy_statement():
    rng = np.random.default_rng(0)
    S, NC, F = 3, 4, 5
    offs = (np.arange(S + 1) * F).astype(np.int32)
    cnt = rng.integers(0, 4, S * F * NC).astype(np.int32)
    start = (np.cumsum(cnt) - cnt).astype(np.int32)
    created = np.minimum(cnt, rng.integers(0, 3, len(cnt))).astype(np.int32)
    out_count = cnt.copy()
    birth = np.zeros((int(cnt.sum()), 2), np.int32)
    for g in range(len(cnt)):
        for k in range(cnt[g]):
            # any earlier-or-equal group of the same sub-stream that created something
            s, c = (g // NC) // F, g % NC
            cands = [h for h in range(s * F * NC + c, g + 1, NC) if created[h] > 0]
            if cands:
                h = cands[rng.integers(len(cands))]
                birth[start[g] + k] = (h, rng.integers(created[h]))
            else:
                out_count[g] = 0
    first = np.full(S * NC, -1, np.int32)
    for s in range(S):
        for c in range(NC):
            nz = np.nonzero(cnt.reshape(S, F, NC)[s, :, c])[0]
            if len(nz):
                first[s * NC + c] = nz[0]
    rank = packing.default_class_rank(first, S, NC)
    a, na = packing.assign_ids(offs, NC, start, out_count, created, first, rank, birth, id_base=7)
    b, nb = runtime.assign_ids(S, NC, offs, start, out_count, created, first, rank, birth, id_base=7)
    np.testing.assert_array_equal(a, b)
    c, nc = runtime.assign_ids(S, NC, offs, start, out_count, created, first, None, birth, id_base=7)
    np.testing.assert_array_equal(a, c)
    assert na == nb == nc == 7 + created.sum()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_entry_points_fail_loudly_without_gpu():
    with pytest.raises(_lib.W2TError):
        runtime.softnms_groups(np.array([0, 1], np.int32), np.zeros((1, 5)))
    with pytest.raises(_lib.W2TError):
        runtime.iou_matrix(np.zeros((1, 4), np.float32), np.zeros((1, 4)))


def test_header_is_plain_c_and_struct_sizes_match_the_ctypes_mirror(tmp_path):
    # include/w2t.h must compile as C99 (the boundary is a C ABI) and every struct of include/w2t_types.h must have
    # the size its ctypes mirror in _abi.py has: a field added on one side only shows up here, not on the GPU
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    inc = os.path.join(ROOT, "include")
    pairs = [("w2t_sort_problem_t", _abi.SortProblem), ("w2t_sort_plan_t", _abi.SortPlan),
             ("w2t_sort_result_t", _abi.SortResult), ("w2t_rows_t", _abi.Rows),
             ("w2t_nms_problem_t", _abi.NmsProblem), ("w2t_nms_result_t", _abi.NmsResult)]
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "w2t.h"\nint main(void) {\n' +
                   "".join('  printf("%%zu\\n", sizeof(%s));\n' % c for c, _ in pairs) + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    for (cname, mirror), size in zip(pairs, sizes):
        assert ctypes.sizeof(mirror) == size, cname
