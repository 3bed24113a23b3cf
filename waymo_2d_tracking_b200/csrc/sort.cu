// sort.cu — host entry points of the SORT stage: plan, launch, id assignment.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>
#include <vector>

#include "sort_kernel.cuh"
#include "sort_warp.cuh"
#include "sort_crowd.cuh"

using namespace w2t;

static int sort_plan_impl(int32_t n_streams, int32_t n_classes, const int32_t *stream_img_offsets,
                          const int32_t *det_count, const int32_t *group_offsets, const uint8_t *img_exists,
                          int32_t max_age, w2t_sort_plan_t *plan);

extern "C" int w2t_sort_plan(int32_t n_streams, int32_t n_classes, const int32_t *stream_img_offsets,
                             const int32_t *det_count, const uint8_t *img_exists, int32_t max_age,
                             w2t_sort_plan_t *plan) {
  return sort_plan_impl(n_streams, n_classes, stream_img_offsets, det_count, nullptr, img_exists, max_age, plan);
}

extern "C" int w2t_sort_plan_offsets(int32_t n_streams, int32_t n_classes, const int32_t *stream_img_offsets,
                                     const int32_t *group_offsets, int32_t max_age, w2t_sort_plan_t *plan) {
  return sort_plan_impl(n_streams, n_classes, stream_img_offsets, nullptr, group_offsets, nullptr, max_age, plan);
}

// det_count[g], or — when only the group offsets of the rows are known (upper bounds, e.g. the group sizes before
// the ensemble) — group_offsets[g + 1] - group_offsets[g], with an image counted as present when it has any row
static int sort_plan_impl(int32_t n_streams, int32_t n_classes, const int32_t *stream_img_offsets,
                          const int32_t *det_count, const int32_t *group_offsets, const uint8_t *img_exists,
                          int32_t max_age, w2t_sort_plan_t *plan) {
  if (!stream_img_offsets || (!det_count && !group_offsets) || !plan || !plan->order || !plan->track_cap || !plan->det_cap ||
      !plan->ws_offset || n_streams < 0 || n_classes < 1 || n_classes > W2T_MAX_CLASSES || max_age < 0) {
    set_last_error("w2t_sort_plan: bad argument");
    return W2T_ERR_ARG;
  }
  const int NC = n_classes;
  const int nq = n_streams * NC;
  const int window = max_age + 2;  // images whose detections can still own a live tracker
  std::vector<int64_t> work(nq, 0);
  // one pass over the images of a stream for all categories at once (det_count is image-major);
  // streams are independent, so the pass is split over a few host threads when the job is large
  auto plan_streams = [&](int s_begin, int s_end) {
    std::vector<int> ring((size_t)window * NC);
    for (int s = s_begin; s < s_end; s++) {
      std::fill(ring.begin(), ring.end(), 0);
      int64_t sum[W2T_MAX_CLASSES] = {0}, best[W2T_MAX_CLASSES] = {0}, w[W2T_MAX_CLASSES] = {0};
      int dmax[W2T_MAX_CLASSES] = {0};
      int pos = 0;
      for (int img = stream_img_offsets[s]; img < stream_img_offsets[s + 1]; img++) {
        if (img_exists && !img_exists[img]) continue;
        int32_t from_offsets[W2T_MAX_CLASSES];
        const int32_t *cnt = det_count ? det_count + (size_t)img * NC : from_offsets;
        if (!det_count) {
          const int32_t *go = group_offsets + (size_t)img * NC;
          if (go[NC] == go[0]) continue;  // no row at all: the image does not exist for the tracker
          for (int c = 0; c < NC; c++) from_offsets[c] = go[c + 1] - go[c];
        }
        int *slot = ring.data() + (size_t)pos * NC;
        for (int c = 0; c < NC; c++) {
          const int d = cnt[c];
          sum[c] += d - slot[c];
          slot[c] = d;
          if (sum[c] > best[c]) best[c] = sum[c];
          if (d > dmax[c]) dmax[c] = d;
          w[c] += (int64_t)d * d + d;
        }
        if (++pos == window) pos = 0;
      }
      for (int c = 0; c < NC; c++) {
        const int q = s * NC + c;
        plan->track_cap[q] = (int32_t)std::max<int64_t>(best[c], 1);
        plan->det_cap[q] = std::max(dmax[c], 1);
        work[q] = w[c];
      }
    }
  };
  const int n_img_total = n_streams > 0 ? stream_img_offsets[n_streams] : 0;
  // a few host threads for large jobs; W2T_PLAN_THREADS caps them (one process per GPU: cores / ranks)
  int n_threads = (n_img_total >= 40000 && n_streams >= 8) ? 4 : 1;
  if (const char *e = getenv("W2T_PLAN_THREADS")) n_threads = std::max(1, std::min(n_threads, atoi(e)));
  if (n_threads == 1) {
    plan_streams(0, n_streams);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++)
      pool.emplace_back(plan_streams, (int)((int64_t)n_streams * t / n_threads), (int)((int64_t)n_streams * (t + 1) / n_threads));
    for (auto &th : pool) th.join();
  }
  // launch classes [wide | mid | narrow] (w2t_sort_plan_t), each heaviest first
  std::iota(plan->order, plan->order + nq, 0);
  auto cls = [&](int q) { return plan->det_cap[q] > W2T_WIDE_DETS ? 0 : plan->det_cap[q] > W2T_NARROW_DETS ? 1 : 2; };
  std::stable_sort(plan->order, plan->order + nq, [&](int a, int b) {
    if (cls(a) != cls(b)) return cls(a) < cls(b);
    return work[a] > work[b];
  });
  plan->n_wide = 0;
  plan->n_mid = 0;
  plan->narrow_cap = 0;
  for (int q = 0; q < nq; q++) {
    plan->n_wide += cls(q) == 0 ? 1 : 0;
    plan->n_mid += cls(q) == 1 ? 1 : 0;
    if (cls(q) == 2) plan->narrow_cap = std::max(plan->narrow_cap, plan->det_cap[q]);
  }
  int64_t off = 0;
  for (int q = 0; q < nq; q++) {
    plan->ws_offset[q] = off;
    off += (int64_t)slab_layout(plan->track_cap[q], plan->det_cap[q]).total;
  }
  plan->aux_offset = off;
  plan->ws_bytes = off + (int64_t)align_up((size_t)W2T_SORT_AUX_BYTES(nq), 256);
  return W2T_OK;
}

static int launch_sort(const char *who, const w2t_sort_problem_t *problem, const w2t_sort_plan_t *plan,
                       w2t_sort_result_t *result, void *workspace, int32_t *sub_state, int32_t group_base,
                       int32_t *status, w2t_stream_t stream) {
  const bool step = sub_state != nullptr;
  if (!problem || !plan || !result || !status || problem->n_classes < 1 ||
      problem->n_classes > W2T_MAX_CLASSES || problem->n_streams < 0) {
    set_last_error("%s: bad argument", who);
    return W2T_ERR_ARG;
  }
  const int nq = problem->n_streams * problem->n_classes;
  if (nq == 0) return W2T_OK;
  if (!workspace || !result->out_box || !result->out_score || !result->out_birth || !result->out_count ||
      !result->created || (!step && !result->first_img)) {
    set_last_error("%s: null buffer", who);
    return W2T_ERR_ARG;
  }
  SortParams P;
  P.sub_state = sub_state;
  P.group_base = group_base;
  P.p = *problem;
  P.r = *result;
  P.order = plan->order;
  P.track_cap = plan->track_cap;
  P.det_cap = plan->det_cap;
  P.ws_offset = plan->ws_offset;
  P.ws = static_cast<char *>(workspace);
  P.chunk_of = plan->chunk_of;
  P.chunk_done = (plan->chunk_of != nullptr) ? plan->chunk_done : nullptr;
  P.status = status;
  // debug aid: W2T_SORT_TIMERS=<device pointer as decimal> receives [n_substreams,16] phase cycle
  // counters from an instrumented instantiation of the same kernel (scripts/phase_timers.py)
  P.timers = getenv("W2T_SORT_TIMERS") ? reinterpret_cast<long long *>(strtoull(getenv("W2T_SORT_TIMERS"), nullptr, 10))
                                        : nullptr;
  P.nep50 = problem->promotion == W2T_PROMOTION_NEP50 ? 1 : 0;
  P.queue = nullptr;
  P.bail = nullptr;
  P.n_items = 0;
  P.bail_want = 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int n_wide = std::min(std::max(plan->n_wide, 0), nq);
  const int n_mid = std::min(std::max(plan->n_mid, 0), nq - n_wide);
  if (step) {
    if (n_wide > 0) sort_track_kernel<512, 1, false, kSmemC, true><<<n_wide, 512, 0, st>>>(P);
    if (nq > n_wide) {
      P.order = plan->order + n_wide;
      sort_track_kernel<kSortBlock, kSortMinBlocks, false, kSmemC, true><<<nq - n_wide, kSortBlock, 0, st>>>(P);
    }
  } else {
    // The plan's classes [wide | mid | narrow] say what a sub-stream MAY need (its counts can be upper
    // bounds); sort_classify_kernel decides from the actual counts.  Crowded sub-streams: 512 threads walk the
    // big cost matrices of the global-memory solver; the middle class: 128 threads, 4 CTAs per SM; both run
    // on a side stream next to the warp kernel (sort_warp.cuh: one warp per sub-stream, persistent warps
    // pulling from a queue), which tracks everything else; what outgrows it is tracked again by CTAs.
    // Without an aux area (hand-made plans) or with W2T_SORT_CTA_ONLY (A/B timing) CTAs track everything.
    const bool warp_path = plan->aux_offset >= 0 && getenv("W2T_SORT_CTA_ONLY") == nullptr;
    const bool tm = P.timers != nullptr;
    if (!warp_path) {
      if (n_wide > 0) {
        if (tm) sort_track_kernel<512, 1, true><<<n_wide, 512, 0, st>>>(P);
        else sort_track_kernel<512, 1, false><<<n_wide, 512, 0, st>>>(P);
      }
      if (nq > n_wide) {
        P.order = plan->order + n_wide;
        if (tm) sort_track_kernel<kSortBlock, kSortMinBlocks, true><<<nq - n_wide, kSortBlock, 0, st>>>(P);
        else sort_track_kernel<kSortBlock, kSortMinBlocks, false><<<nq - n_wide, kSortBlock, 0, st>>>(P);
      }
    } else {
      static int sm_count = 0;
      static cudaStream_t side[16] = {nullptr};
      static cudaEvent_t ev_fork[16] = {nullptr}, ev_join[16] = {nullptr};
      int dev = 0;
      W2T_CUDA_TRY(cudaGetDevice(&dev));
      if (sm_count == 0) W2T_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
      char *aux = P.ws + plan->aux_offset;
      W2T_CUDA_TRY(cudaMemsetAsync(aux, 0, (size_t)W2T_SORT_QUEUE_BYTES(nq), st));
      const WarpQueues Q = warp_queues(aux, nq);
      static_assert((int64_t)kSpillCtas * kTeamsPerCta * kSpillFloats * 4 == W2T_SORT_SPILL_BYTES, "spill area size");
      P.spill = reinterpret_cast<float *>(aux + W2T_SORT_QUEUE_BYTES(nq));
      P.queue = Q.hdr;
      P.bail = Q.cls;
      P.n_items = nq;
      sort_dmax_kernel<<<(nq + 63) / 64, 64, 0, st>>>(P.p, Q.cls);
      sort_classify_kernel<<<1, 1024, 0, st>>>(P.p, plan->order, Q);
      const int n_big = n_wide + n_mid;  // leading entries of the order that may be too crowded for a warp
      const bool fork = n_big > 0 && dev < 16;
      cudaStream_t sb = st;
      if (fork) {
        if (side[dev] == nullptr) {
          W2T_CUDA_TRY(cudaStreamCreateWithFlags(&side[dev], cudaStreamNonBlocking));
          W2T_CUDA_TRY(cudaEventCreateWithFlags(&ev_fork[dev], cudaEventDisableTiming));
          W2T_CUDA_TRY(cudaEventCreateWithFlags(&ev_join[dev], cudaEventDisableTiming));
        }
        sb = side[dev];
        W2T_CUDA_TRY(cudaEventRecord(ev_fork[dev], st));
        W2T_CUDA_TRY(cudaStreamWaitEvent(sb, ev_fork[dev], 0));
      }
      static bool crowd_attr[16] = {false};
      const size_t crowd_smem = (size_t)kCrowdSmemBytes;
      if (dev >= 16 || !crowd_attr[dev]) {
        W2T_CUDA_TRY(cudaFuncSetAttribute(sort_crowd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)crowd_smem));
        W2T_CUDA_TRY(cudaFuncSetAttribute(sort_crowd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        if (dev < 16) crowd_attr[dev] = true;
      }
      // clusters of `ctas` CTAs over the first `entries` sub-streams of the launch order, serving two class flags
      auto launch_crowd = [&](int entries, int ctas, int want_a, int want_b, int overflow, cudaStream_t s_) -> cudaError_t {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(entries * ctas));
        cfg.blockDim = dim3(kCrowdWarps * 32);
        cfg.dynamicSmemBytes = crowd_smem;
        cfg.stream = s_;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)ctas;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, sort_crowd_kernel, P, want_a, want_b, overflow);
      };
      const bool dbg_no16 = getenv("W2T_DBG_NO_CROWD16") != nullptr, dbg_no8 = getenv("W2T_DBG_NO_CROWD8") != nullptr;
      if (n_wide > 0 && !dbg_no16) W2T_CUDA_TRY(launch_crowd(n_wide, 16, kClsWide, kClsWide, kClsHuge, sb));
      if (n_big > 0 && !dbg_no8) W2T_CUDA_TRY(launch_crowd(n_big, 8, kClsMid, kClsMid, kClsOver8, sb));
      if (fork) W2T_CUDA_TRY(cudaEventRecord(ev_join[dev], sb));
      {
        static bool attr_set[16] = {false};
        const size_t smem = sizeof(WarpShared) * kTeamsPerCta;
        if (dev >= 16 || !attr_set[dev]) {
          W2T_CUDA_TRY(cudaFuncSetAttribute(sort_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          W2T_CUDA_TRY(cudaFuncSetAttribute(sort_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          if (dev < 16) attr_set[dev] = true;
        }
        const int ctas = std::min(std::min(sm_count, kSpillCtas), nq);  // one CTA per SM; the first round is dealt across them
        if (tm) sort_warp_kernel<true><<<ctas, kWarpsPerCta * 32, smem, st>>>(P);
        else sort_warp_kernel<false><<<ctas, kWarpsPerCta * 32, smem, st>>>(P);
      }
      // second pass: sub-streams that outgrew the warp kernel are tracked again by clusters ...
      if (!dbg_no16) W2T_CUDA_TRY(launch_crowd(nq, 16, kClsBailed, kClsOver8, kClsHuge, st));
      if (fork) W2T_CUDA_TRY(cudaStreamWaitEvent(st, ev_join[dev], 0));
      // ... and what outgrew those, by CTAs that keep the cost matrix in global memory
      P.bail_want = kClsHuge;
      sort_track_kernel<512, 1, false><<<nq, 512, 0, st>>>(P);
    }
  }
  W2T_CUDA_TRY(cudaGetLastError());
  return W2T_OK;
}

extern "C" int w2t_sort_track(const w2t_sort_problem_t *problem, const w2t_sort_plan_t *plan,
                              w2t_sort_result_t *result, void *workspace, int32_t *status, w2t_stream_t stream) {
  return launch_sort("w2t_sort_track", problem, plan, result, workspace, nullptr, 0, status, stream);
}

extern "C" int w2t_sort_step(const w2t_sort_problem_t *problem, const w2t_sort_plan_t *plan,
                             w2t_sort_result_t *result, void *workspace, int32_t *sub_state, int32_t group_base,
                             int32_t *status, w2t_stream_t stream) {
  if (!sub_state) {
    set_last_error("w2t_sort_step: sub_state is required");
    return W2T_ERR_ARG;
  }
  return launch_sort("w2t_sort_step", problem, plan, result, workspace, sub_state, group_base, status, stream);
}

extern "C" size_t w2t_sort_slab_bytes(int32_t track_cap, int32_t det_cap) {
  return slab_layout(std::max(track_cap, 1), std::max(det_cap, 1)).total;
}

extern "C" int w2t_assign_ids(int32_t n_streams, int32_t n_classes, const int32_t *stream_img_offsets,
                              const int32_t *det_start, const int32_t *out_count, const int32_t *created,
                              const int32_t *first_img, const int32_t *class_rank, const int32_t *out_birth,
                              int64_t id_base, int64_t *out_id, int64_t *id_next) {
  if (!stream_img_offsets || !det_start || !out_count || !created || !out_birth || !out_id ||
      (!first_img && !class_rank) || n_classes < 1 || n_classes > W2T_MAX_CLASSES) {
    set_last_error("w2t_assign_ids: bad argument");
    return W2T_ERR_ARG;
  }
  const int NC = n_classes;
  const int n_img = stream_img_offsets[n_streams];
  std::vector<int64_t> base((size_t)n_img * NC);
  int64_t next = id_base;
  for (int s = 0; s < n_streams; s++) {
    // categories in the order the reference's tracker dict holds them
    int cats[W2T_MAX_CLASSES];
    std::iota(cats, cats + NC, 0);
    if (class_rank) {
      const int32_t *rk = class_rank + (size_t)s * NC;
      std::stable_sort(cats, cats + NC, [&](int a, int b) { return rk[a] < rk[b]; });
    } else {
      const int32_t *fi = first_img + (size_t)s * NC;
      std::stable_sort(cats, cats + NC, [&](int a, int b) {
        const int64_t fa = fi[a] < 0 ? INT64_MAX : fi[a], fb = fi[b] < 0 ? INT64_MAX : fi[b];
        return fa < fb;
      });
    }
    for (int img = stream_img_offsets[s]; img < stream_img_offsets[s + 1]; img++)
      for (int k = 0; k < NC; k++) {
        const size_t g = (size_t)img * NC + cats[k];
        base[g] = next;
        next += created[g];
      }
  }
  for (size_t g = 0; g < (size_t)n_img * NC; g++) {
    const size_t r0 = (size_t)det_start[g];
    for (int k = 0; k < out_count[g]; k++) {
      const size_t r = r0 + k;
      out_id[r] = base[out_birth[2 * r]] + out_birth[2 * r + 1] + 1;  // id + 1, sort.py:288
    }
  }
  if (id_next) *id_next = next;
  return W2T_OK;
}
