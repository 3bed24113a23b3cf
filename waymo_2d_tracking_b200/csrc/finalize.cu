// finalize.cu — global object ids and the dense output list of the SORT stage, on the device.
//
// The reference numbers trackers with one process-global counter (KalmanBoxTracker.count,
// tracking/sort/sort.py:86,140-141) in creation order — streams in order, images in order,
// categories in tracker-dict order (tracker_sort.py:32-33,41), new trackers in unmatched
// order — and appends output rows in the same nesting, each (image, category) block walked
// in reverse list order (sort.py:280; tracking/utils.py:37-58).  Both are exclusive scans
// over per-(image, category) counts taken in that processing order:
//   1. order kernel : per stream, rank the categories and permute the `created` and
//                     `out_count` arrays into processing order;
//   2. two scans    : id base per group, dense row offset per group (three small kernels for both);
//   3. rows kernel  : one warp per group copies its rows to their dense position (reversed)
//                     and resolves object_id = base[birth group] + k + 1 (sort.py:288).
#include <algorithm>

#include "common.cuh"

using namespace w2t;

namespace {

struct FinParams {
  int32_t n_streams, n_classes;
  const int32_t *stream_img_offsets, *det_start, *out_count, *created, *first_img, *class_rank;
  const double *out_box, *out_score;
  const int32_t *out_birth;
  int64_t id_base;
  const int64_t *id_base_dev;
  int32_t *perm_created, *perm_count, *scan_created, *scan_count, *stream_rank;
  int64_t *totals;  // [0] ids created, [1] dense rows, [2] next id base
  double *rows_box, *rows_score;
  int64_t *rows_id;
  int32_t *rows_img, *rows_cat;
  int2 *rows_compact;  // optional: (id - id_base, image * 8 + category - 1) instead of the three arrays above
  int64_t rows_cap;
  int32_t image_base;
  int64_t birth_base;
};

// One block per stream.
__global__ void order_kernel(const FinParams P) {
  const int s = blockIdx.x, NC = P.n_classes;
  int rank[W2T_MAX_CLASSES];
  for (int c = 0; c < NC; c++) {
    int r = 0;
    if (P.class_rank) {
      r = P.class_rank[s * NC + c];
    } else {
      // position in the tracker dict = order of first appearance; same image -> category order
      // (what detnet.ensemble's output gives: it emits an image's rows category by category)
      const long long fc = P.first_img[s * NC + c] < 0 ? (1ll << 40) : P.first_img[s * NC + c];
      for (int o = 0; o < NC; o++) {
        if (o == c) continue;
        const long long fo = P.first_img[s * NC + o] < 0 ? (1ll << 40) : P.first_img[s * NC + o];
        r += (fo < fc || (fo == fc && o < c)) ? 1 : 0;
      }
    }
    rank[c] = r;
    if (threadIdx.x == 0) P.stream_rank[s * NC + c] = r;
  }
  const int img0 = P.stream_img_offsets[s], img1 = P.stream_img_offsets[s + 1];
  for (int i = img0 * NC + threadIdx.x; i < img1 * NC; i += blockDim.x) {
    const int img = i / NC, c = i % NC;
    P.perm_created[img * NC + rank[c]] = P.created[i];
    P.perm_count[img * NC + rank[c]] = P.out_count[i];
  }
}

// Exclusive scans of the two permuted count arrays (ids created, rows emitted) over all groups,
// in three small kernels: per-block sums of contiguous chunks, a one-block scan of those sums, and
// the per-block scan of each chunk from its offset.  kScanBlocks chunks keep every SM busy.
constexpr int kScanBlocks = 592;
constexpr int kScanThreads = 256;

__device__ __forceinline__ int64_t block_sum(int64_t v, int64_t *s_red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  int64_t t = 0;
  for (int i = 0; i < kScanThreads / 32; i++) t += s_red[i];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(const int32_t *a, const int32_t *b, int64_t n,
                                                                 int64_t *sums) {
  __shared__ int64_t s_red[kScanThreads / 32];
  const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = min((int64_t)blockIdx.x * chunk, n), hi = min(lo + chunk, n);
  int64_t sa = 0, sb = 0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += kScanThreads) { sa += a[i]; sb += b[i]; }
  sa = block_sum(sa, s_red);
  sb = block_sum(sb, s_red);
  if (threadIdx.x == 0) { sums[2 * blockIdx.x] = sa; sums[2 * blockIdx.x + 1] = sb; }
}

// one block: exclusive scan of the per-chunk sums in place; totals[0] = all of a, totals[1] = all of b
__global__ void scan_offsets_kernel(int64_t *sums, int nb, int64_t *totals) {
  if (threadIdx.x == 0) {
    int64_t ra = 0, rb = 0;
    for (int i = 0; i < nb; i++) {
      const int64_t va = sums[2 * i], vb = sums[2 * i + 1];
      sums[2 * i] = ra;
      sums[2 * i + 1] = rb;
      ra += va;
      rb += vb;
    }
    totals[0] = ra;
    totals[1] = rb;
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int32_t *a, const int32_t *b, int64_t n,
                                                                  const int64_t *sums, int32_t *out_a, int32_t *out_b) {
  __shared__ int64_t s_red[kScanThreads / 32];
  __shared__ int s_wa[kScanThreads / 32], s_wb[kScanThreads / 32];
  const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = min((int64_t)blockIdx.x * chunk, n), hi = min(lo + chunk, n);
  int64_t run_a = sums[2 * blockIdx.x], run_b = sums[2 * blockIdx.x + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t i0 = lo; i0 < hi; i0 += kScanThreads) {
    const int64_t i = i0 + threadIdx.x;
    const int va = (i < hi) ? a[i] : 0, vb = (i < hi) ? b[i] : 0;
    int ia = va, ib = vb;  // inclusive warp scans
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
      if (lane >= o) { ia += ta; ib += tb; }
    }
    if (lane == 31) { s_wa[warp] = ia; s_wb[warp] = ib; }
    __syncthreads();
    int pa = 0, pb = 0, ta = 0, tb = 0;
    for (int w = 0; w < kScanThreads / 32; w++) {
      if (w < warp) { pa += s_wa[w]; pb += s_wb[w]; }
      ta += s_wa[w];
      tb += s_wb[w];
    }
    if (i < hi) {
      out_a[i] = (int32_t)(run_a + pa + ia - va);
      out_b[i] = (int32_t)(run_b + pb + ib - vb);
    }
    run_a += ta;
    run_b += tb;
    __syncthreads();
  }
  (void)s_red;
}

// One block per stream, one warp per (image, category) group at a time.
__global__ void rows_kernel(const FinParams P) {
  const int s = blockIdx.x, NC = P.n_classes;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int img0 = P.stream_img_offsets[s], img1 = P.stream_img_offsets[s + 1];
  const int64_t id_base = P.id_base + (P.id_base_dev ? *P.id_base_dev : 0);
  if (s == 0 && threadIdx.x == 0) P.totals[2] = id_base + P.totals[0];
  for (int g = img0 * NC + warp; g < img1 * NC; g += nwarp) {
    const int cnt = P.out_count[g];
    if (cnt == 0) continue;
    const int img = g / NC, c = g % NC;
    const int k = P.stream_rank[s * NC + c];
    const int64_t off = P.scan_count[img * NC + k];
    const int64_t r0 = P.det_start[g];
    for (int j = lane; j < cnt; j += 32) {
      const int64_t src = r0 + j;
      const int64_t dst = off + (cnt - 1 - j);  // the reference walks the tracker list reversed
      if (dst >= P.rows_cap) continue;
      const double4 b = reinterpret_cast<const double4 *>(P.out_box)[src];
      reinterpret_cast<double4 *>(P.rows_box)[dst] = b;
      P.rows_score[dst] = P.out_score[src];
      const int bg = (int)(P.out_birth[2 * src] - P.birth_base), bk = P.out_birth[2 * src + 1];
      const int64_t id = id_base + P.scan_created[(bg / NC) * NC + k] + bk + 1;
      if (P.rows_compact != nullptr) {
        P.rows_compact[dst] = make_int2((int)(id - P.id_base), (img + P.image_base) * 8 + c);
      } else {
        P.rows_id[dst] = id;
        P.rows_img[dst] = img + P.image_base;
        P.rows_cat[dst] = c + 1;
      }
    }
  }
}

}  // namespace

extern "C" size_t w2t_sort_finalize_workspace(int32_t n_streams, int32_t n_classes, int64_t n_groups) {
  return (size_t)(4 * n_groups + (int64_t)n_streams * n_classes) * sizeof(int32_t) + 256 +
         2 * kScanBlocks * sizeof(int64_t);
}

extern "C" int w2t_sort_finalize(const w2t_sort_problem_t *problem, const w2t_sort_result_t *result,
                                 const int32_t *class_rank, int64_t id_base, int64_t n_groups, void *workspace,
                                 w2t_rows_t *rows, w2t_stream_t stream) {
  if (!problem || !result || !rows || problem->n_classes < 1 || problem->n_classes > W2T_MAX_CLASSES ||
      problem->n_streams < 0 || n_groups < 0) {
    set_last_error("w2t_sort_finalize: bad argument");
    return W2T_ERR_ARG;
  }
  if (!rows->totals) {
    set_last_error("w2t_sort_finalize: rows->totals is required");
    return W2T_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (problem->n_streams == 0 || n_groups == 0) {
    // nothing created: the next id base is the incoming one
    W2T_CUDA_TRY(cudaMemsetAsync(rows->totals, 0, 2 * sizeof(int64_t), st));
    if (rows->id_base_device)
      W2T_CUDA_TRY(cudaMemcpyAsync(rows->totals + 2, rows->id_base_device, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    else
      W2T_CUDA_TRY(cudaMemcpyAsync(rows->totals + 2, &id_base, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    return W2T_OK;
  }
  if (!workspace || !rows->box || !rows->score ||
      (!rows->compact && (!rows->object_id || !rows->image || !rows->category))) {
    set_last_error("w2t_sort_finalize: null buffer");
    return W2T_ERR_ARG;
  }
  FinParams P;
  P.n_streams = problem->n_streams;
  P.n_classes = problem->n_classes;
  P.stream_img_offsets = problem->stream_img_offsets;
  P.det_start = problem->det_start;
  P.out_count = result->out_count;
  P.created = result->created;
  P.first_img = result->first_img;
  P.class_rank = class_rank;
  P.out_box = result->out_box;
  P.out_score = result->out_score;
  P.out_birth = result->out_birth;
  P.id_base = id_base;
  P.id_base_dev = rows->id_base_device;
  int32_t *w = static_cast<int32_t *>(workspace);
  P.perm_created = w;
  P.perm_count = w + n_groups;
  P.scan_created = w + 2 * n_groups;
  P.scan_count = w + 3 * n_groups;
  P.stream_rank = w + 4 * n_groups;
  P.totals = rows->totals;
  P.rows_box = rows->box;
  P.rows_score = rows->score;
  P.rows_id = rows->object_id;
  P.rows_img = rows->image;
  P.rows_cat = rows->category;
  P.rows_compact = reinterpret_cast<int2 *>(rows->compact);
  P.rows_cap = rows->capacity;
  P.image_base = rows->image_base;
  P.birth_base = rows->birth_group_base;
  order_kernel<<<P.n_streams, 256, 0, st>>>(P);
  // 8-byte aligned scratch for the per-chunk sums, behind the int32 arrays
  const size_t ints = (size_t)(4 * n_groups + (int64_t)P.n_streams * P.n_classes);
  int64_t *sums = reinterpret_cast<int64_t *>(static_cast<char *>(workspace) + ((ints * sizeof(int32_t) + 7) / 8) * 8);
  const int nb = (int)std::min<int64_t>(kScanBlocks, (n_groups + kScanThreads - 1) / kScanThreads);
  scan_sums_kernel<<<nb, kScanThreads, 0, st>>>(P.perm_created, P.perm_count, n_groups, sums);
  scan_offsets_kernel<<<1, 32, 0, st>>>(sums, nb, P.totals);
  scan_apply_kernel<<<nb, kScanThreads, 0, st>>>(P.perm_created, P.perm_count, n_groups, sums, P.scan_created, P.scan_count);
  rows_kernel<<<P.n_streams, 256, 0, st>>>(P);
  W2T_CUDA_TRY(cudaGetLastError());
  return W2T_OK;
}
