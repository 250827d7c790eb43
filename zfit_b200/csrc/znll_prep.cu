// znll_prep.cu -- bind-time preprocessing on the device: (1) a stable bucket partition of
// the resident event columns, so that the step kernel's warps are uniform with respect to piecewise pdfs, and (2) the
// inclusive cut to the observable space with a stable compaction (znll_cut_events, at the end of this file).
//
// Why: events are static for a whole fit and the NLL is a sum over them, so their order is free.  A CrystalBall (and its
// two-sided relatives) evaluates a Gaussian core or a power-law tail depending on which side of mu - |alpha| sigma the
// event lies; on shuffled data 91 % of the warps contain at least one tail event and pay for both (+20 % per step on
// cfg 2: 0.673 vs 0.562 ms, measured).  A full sort (round 1: torch.sort, 14 ms for 1e8 events -- more than a whole
// fit's kernel time) is not needed: warps only have to be uniform, and the threshold moves during the fit.  One stable
// counting pass into ZNLL_PREP_BUCKETS = 32 equal-width buckets of the key observable leaves only the bucket that
// contains the threshold mixed (~1/32 of the range) and costs two reads of the key column plus one read and one write
// of every column.
//
// Layout: the n events are cut into rows of 32 consecutive events; every warp of the grid owns a contiguous range of
// rows and walks it twice with the same code -- pass 1 counts its events per bucket, a single-block scan turns the
// [bucket][warp] counts into exclusive offsets (bucket-major, so bucket b of warp w starts where bucket b of warp w-1
// ended), pass 2 moves the events.  Inside a row the rank of an event among the row's events of the same bucket comes
// from __match_any_sync, so the partition is STABLE and therefore deterministic: the same input gives the same order,
// and the NLL's fixed-order reduction stays bitwise reproducible across binds.
//
// The reference has no counterpart (tf.where evaluates both branches for every event, models/physics.py:31-45); this
// replaces the torch.sort call of round 1's CudaEngine._sort_for_coherence.
#include <cstdint>
#include <cstdio>

#include "../../include/znll.h"
#include "znll_internal.h"

#define ZNLL_PREP_BUCKETS 32
#define ZNLL_PREP_BLOCK 256
#define ZNLL_PREP_WARPS (ZNLL_PREP_BLOCK / 32)

namespace {

struct PrepArgs {
  const double* src[ZNLL_MAX_COLS + 1];  // columns, then the weights (if any)
  double* dst[ZNLL_MAX_COLS + 1];
  int n_streams, key;
  long long n, rows_per_warp;
  double lo, scale;      // bucket = clamp(floor((x - lo) * scale), 0, 31)
  unsigned int* counts;  // [32][n_warps] counts, then exclusive offsets after the scan (n <= 2^32 - 1 per bind)
  int n_warps;
};

__device__ __forceinline__ int bucket_of(double x, double lo, double scale) {
  const int b = __double2int_rd((x - lo) * scale);  // NaN -> 0
  return min(max(b, 0), ZNLL_PREP_BUCKETS - 1);
}

// one row of (up to) 32 consecutive events: rank every event among the row's events of its bucket, advance the warp's
// per-bucket cursors, and (MOVE) write all streams of the event to its position
template <bool MOVE>
__device__ __forceinline__ void do_row(const PrepArgs& a, unsigned int* cur, unsigned mask, long long i, double x, int lane) {
  const int b = bucket_of(x, a.lo, a.scale);
  const unsigned same = __match_any_sync(mask, b);
  const int leader = __ffs(same) - 1;
  unsigned base = 0;
  if (lane == leader) {
    base = cur[b];
    cur[b] = base + __popc(same);
  }
  __syncwarp(mask);
  if (MOVE) {
    base = __shfl_sync(mask, base, leader);
    const size_t pos = (size_t)base + __popc(same & ((1u << lane) - 1u));
#pragma unroll 1
    for (int s = 0; s < a.n_streams; ++s) a.dst[s][pos] = (s == a.key) ? x : __ldg(a.src[s] + i);
  }
}

// MOVE = false: count events per (warp, bucket);  MOVE = true: scatter every stream to its bucket position
template <bool MOVE>
__global__ void __launch_bounds__(ZNLL_PREP_BLOCK) partition_kernel(const __grid_constant__ PrepArgs a) {
  __shared__ unsigned int off[ZNLL_PREP_WARPS][ZNLL_PREP_BUCKETS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = blockIdx.x * ZNLL_PREP_WARPS + wib;
  unsigned int* cur = off[wib];
  cur[lane] = MOVE ? a.counts[(size_t)lane * a.n_warps + warp] : 0u;
  __syncwarp();
  const long long full_rows = a.n >> 5;  // rows whose 32 events all exist
  long long r = (long long)warp * a.rows_per_warp;
  const long long r_stop = r + a.rows_per_warp;
  const long long r_end = min(full_rows, r_stop);
  const double* __restrict__ keycol = a.src[a.key];
  for (; r + 4 <= r_end; r += 4) {  // four key loads in flight per lane
    const long long i = (r << 5) + lane;
    const double x0 = __ldg(keycol + i), x1 = __ldg(keycol + i + 32), x2 = __ldg(keycol + i + 64), x3 = __ldg(keycol + i + 96);
    do_row<MOVE>(a, cur, 0xffffffffu, i, x0, lane);
    do_row<MOVE>(a, cur, 0xffffffffu, i + 32, x1, lane);
    do_row<MOVE>(a, cur, 0xffffffffu, i + 64, x2, lane);
    do_row<MOVE>(a, cur, 0xffffffffu, i + 96, x3, lane);
  }
  for (; r < r_end; ++r) {
    const long long i = (r << 5) + lane;
    do_row<MOVE>(a, cur, 0xffffffffu, i, __ldg(keycol + i), lane);
  }
  if ((a.n & 31) && full_rows >= (long long)warp * a.rows_per_warp && full_rows < r_stop) {  // the partial last row is this warp's
    const long long i = (full_rows << 5) + lane;
    const bool live = i < a.n;
    const unsigned mask = __ballot_sync(0xffffffffu, live);
    if (live) do_row<MOVE>(a, cur, mask, i, __ldg(keycol + i), lane);
  }
  if (!MOVE) {
    __syncwarp();
    a.counts[(size_t)lane * a.n_warps + warp] = cur[lane];
  }
}

// exclusive scan of m = 32 * n_warps counts (a multiple of 4), in place, by one block of 1024 threads: tiles of 4096
// entries, four per thread (one 16-byte access), warp-shuffle scans, a running carry
__global__ void __launch_bounds__(1024) scan_kernel(unsigned int* c, int m, unsigned int* total /* nullable */) {
  __shared__ unsigned int wsum[32];
  __shared__ unsigned int tile_total;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  unsigned int carry = 0;
  for (int base = 0; base < m; base += 4096) {
    const int i = base + 4 * t;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (i < m) v = *reinterpret_cast<const uint4*>(c + i);
    const unsigned int s = v.x + v.y + v.z + v.w;
    unsigned int incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const unsigned int w = wsum[lane];
      unsigned int wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned int o = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += o;
      }
      wsum[lane] = wi - w;  // exclusive prefix of the warp totals
      if (lane == 31) tile_total = wi;
    }
    __syncthreads();
    const unsigned int excl = carry + wsum[warp] + (incl - s);
    if (i < m) *reinterpret_cast<uint4*>(c + i) = make_uint4(excl, excl + v.x, excl + v.x + v.y, excl + v.x + v.y + v.z);
    carry += tile_total;
    __syncthreads();
  }
  if (total && t == 0) *total = carry;
}

// ---------------------------------------------------------------------------------------------------------------
// Inclusive cut + stable compaction (the reference's one-time cut of Data to its Space: check_cut_data_weights /
// Data._cut, core/data.py:1203-1244, an inclusive `lower <= x <= upper` on every observable followed by a boolean mask).
// Same two-pass shape as the partition with a single bucket: every warp owns a contiguous range of rows of 32 events,
// pass 1 counts its survivors, the scan turns the per-warp counts into offsets (and the grand total), pass 2 moves the
// survivors -- rank inside a row from the row's ballot, so the order of the kept events is their original order.
// The optional pair (less_a, less_b) adds the predicate less_a[i] < less_b[i]: the accept step of accept-reject
// sampling (u < p, core/sample.py:440-505) compacts with the same kernel.
struct CutArgs {
  const double* src[ZNLL_MAX_COLS + 1];  // columns, then the weights (if any)
  double* dst[ZNLL_MAX_COLS + 1];
  double lo[ZNLL_MAX_COLS], hi[ZNLL_MAX_COLS];
  const double *less_a, *less_b;
  int n_cols, n_streams;
  long long n, rows_per_warp;
  unsigned int* counts;  // [n_warps] survivors, then exclusive offsets after the scan
  int n_warps;
};

template <bool MOVE>
__device__ __forceinline__ void cut_row(const CutArgs& a, unsigned int& cursor, long long i, bool live, int lane) {
  double x[ZNLL_MAX_COLS];
  bool keep = live;
#pragma unroll
  for (int c = 0; c < ZNLL_MAX_COLS; ++c)
    if (c < a.n_cols) {
      x[c] = live ? __ldg(a.src[c] + i) : 0.0;
      keep = keep && x[c] >= a.lo[c] && x[c] <= a.hi[c];  // NaN fails both comparisons
    }
  if (a.less_a && live) keep = keep && __ldg(a.less_a + i) < __ldg(a.less_b + i);
  const unsigned kept = __ballot_sync(0xffffffffu, keep);
  if (MOVE && keep) {
    const size_t pos = (size_t)cursor + __popc(kept & ((1u << lane) - 1u));
#pragma unroll
    for (int c = 0; c < ZNLL_MAX_COLS; ++c)
      if (c < a.n_cols) a.dst[c][pos] = x[c];
    if (a.n_streams > a.n_cols) a.dst[a.n_cols][pos] = __ldg(a.src[a.n_cols] + i);
  }
  cursor += __popc(kept);
}

template <bool MOVE>
__global__ void __launch_bounds__(ZNLL_PREP_BLOCK) cut_kernel(const __grid_constant__ CutArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * ZNLL_PREP_WARPS + (threadIdx.x >> 5);
  unsigned int cursor = MOVE ? a.counts[warp] : 0u;
  const long long rows = (a.n + 31) >> 5;
  long long r = (long long)warp * a.rows_per_warp;
  const long long r_end = min(rows, r + a.rows_per_warp);
#pragma unroll 2
  for (; r < r_end; ++r) {
    const long long i = (r << 5) + lane;
    cut_row<MOVE>(a, cursor, i, i < a.n, lane);
  }
  if (!MOVE && lane == 0) a.counts[warp] = cursor;
}

thread_local char g_prep_err[256] = "";

}  // namespace

// One block of counters per device, kept for the life of the process (the stream-ordered allocator hands memory back at
// every synchronisation and re-mapping it costs more than the kernels); sized for the largest grid used here plus the
// cut's total.
static cudaError_t counts_buffer(int device, size_t m, int sms, unsigned int** out) {
  static unsigned int* counts_of[64] = {nullptr};
  static size_t counts_cap[64] = {0};
  if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
  if (counts_cap[device] < m) {
    if (counts_of[device]) cudaFree(counts_of[device]);
    counts_of[device] = nullptr;
    counts_cap[device] = 0;
    const size_t cap = (size_t)ZNLL_PREP_BUCKETS * (size_t)sms * 8 * ZNLL_PREP_WARPS + 4;
    const cudaError_t e = cudaMalloc(&counts_of[device], (cap > m ? cap : m) * sizeof(unsigned int));
    if (e != cudaSuccess) return e;
    counts_cap[device] = cap > m ? cap : m;
  }
  *out = counts_of[device];
  return cudaSuccess;
}

static int sm_count_of(int device, cudaError_t* e) {
  static int sms_of[64] = {0};
  if (*e == cudaSuccess && device >= 0 && device < 64 && sms_of[device] == 0)
    *e = cudaDeviceGetAttribute(&sms_of[device], cudaDevAttrMultiProcessorCount, device);
  return (device >= 0 && device < 64 && sms_of[device] > 0) ? sms_of[device] : 148;
}

extern "C" const char* znll_prep_last_error(void) { return g_prep_err; }

extern "C" int znll_partition_events(int device, int n_cols, const double* const* src_cols, double* const* dst_cols,
                                     const double* src_w, double* dst_w, int key_col, double lo, double hi, int64_t n,
                                     void* stream) {
  if (n_cols < 1 || n_cols > ZNLL_MAX_COLS || !src_cols || !dst_cols || key_col < 0 || key_col >= n_cols || n < 0 ||
      !(hi > lo) || (src_w != nullptr) != (dst_w != nullptr) || n > 0xfffffff0ll) {
    snprintf(g_prep_err, sizeof g_prep_err, "znll_partition_events: bad arguments");
    return 1;
  }
  if (n == 0) return 0;
  cudaError_t e = cudaSetDevice(device);
  const int sms = sm_count_of(device, &e);
  PrepArgs a;
  a.n_streams = n_cols + (src_w ? 1 : 0);
  for (int i = 0; i < n_cols; ++i) {
    a.src[i] = src_cols[i];
    a.dst[i] = dst_cols[i];
    if (!a.src[i] || !a.dst[i] || a.src[i] == a.dst[i]) {
      snprintf(g_prep_err, sizeof g_prep_err, "znll_partition_events: column %d null or aliased", i);
      return 1;
    }
  }
  if (src_w) {
    a.src[n_cols] = src_w;
    a.dst[n_cols] = dst_w;
  }
  a.key = key_col;
  a.n = n;
  a.lo = lo;
  a.scale = (double)ZNLL_PREP_BUCKETS / (hi - lo);
  const long long rows = (n + 31) >> 5;
  long long blocks = (long long)sms * 8;  // 8 resident CTAs of 256 threads per SM: the pass is latency-bound on its loads
  if (blocks * ZNLL_PREP_WARPS > rows) blocks = (rows + ZNLL_PREP_WARPS - 1) / ZNLL_PREP_WARPS;
  a.n_warps = (int)(blocks * ZNLL_PREP_WARPS);
  a.rows_per_warp = (rows + a.n_warps - 1) / a.n_warps;
  cudaStream_t st = (cudaStream_t)stream;
  // the [bucket][warp] counters: see counts_buffer.  Calls on one device must be ordered by the caller (same stream, or a
  // synchronisation in between); the engine does both.
  const size_t m = (size_t)ZNLL_PREP_BUCKETS * a.n_warps;
  unsigned int* counts = nullptr;
  if (e == cudaSuccess) e = counts_buffer(device, m, sms, &counts);
  if (e == cudaSuccess) {
    a.counts = counts;
    partition_kernel<false><<<(unsigned)blocks, ZNLL_PREP_BLOCK, 0, st>>>(a);
    scan_kernel<<<1, 1024, 0, st>>>(a.counts, (int)m, nullptr);
    partition_kernel<true><<<(unsigned)blocks, ZNLL_PREP_BLOCK, 0, st>>>(a);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    snprintf(g_prep_err, sizeof g_prep_err, "znll_partition_events: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

extern "C" int znll_cut_events(int device, int n_cols, const double* const* src_cols, double* const* dst_cols,
                               const double* src_w, double* dst_w, const double* lower, const double* upper,
                               const double* less_a, const double* less_b, int64_t n, int64_t* n_kept, void* stream) {
  if (n_cols < 1 || n_cols > ZNLL_MAX_COLS || !src_cols || !dst_cols || !lower || !upper || !n_kept || n < 0 ||
      (src_w != nullptr) != (dst_w != nullptr) || (less_a != nullptr) != (less_b != nullptr) || n > 0xfffffff0ll) {
    snprintf(g_prep_err, sizeof g_prep_err, "znll_cut_events: bad arguments");
    return 1;
  }
  *n_kept = 0;
  if (n == 0) return 0;
  cudaError_t e = cudaSetDevice(device);
  const int sms = sm_count_of(device, &e);
  CutArgs a;
  a.n_cols = n_cols;
  a.n_streams = n_cols + (src_w ? 1 : 0);
  for (int i = 0; i < n_cols; ++i) {
    a.src[i] = src_cols[i];
    a.dst[i] = dst_cols[i];
    a.lo[i] = lower[i];
    a.hi[i] = upper[i];
    if (!a.src[i] || !a.dst[i] || a.src[i] == a.dst[i]) {
      snprintf(g_prep_err, sizeof g_prep_err, "znll_cut_events: column %d null or aliased", i);
      return 1;
    }
  }
  if (src_w) {
    a.src[n_cols] = src_w;
    a.dst[n_cols] = dst_w;
  }
  a.less_a = less_a;
  a.less_b = less_b;
  a.n = n;
  const long long rows = (n + 31) >> 5;
  long long blocks = (long long)sms * 8;
  if (blocks * ZNLL_PREP_WARPS > rows) blocks = (rows + ZNLL_PREP_WARPS - 1) / ZNLL_PREP_WARPS;
  a.n_warps = (int)(blocks * ZNLL_PREP_WARPS);  // a multiple of 8: the scan reads 16 bytes at a time
  a.rows_per_warp = (rows + a.n_warps - 1) / a.n_warps;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* counts = nullptr;
  if (e == cudaSuccess) e = counts_buffer(device, (size_t)a.n_warps + 4, sms, &counts);
  unsigned int total = 0;
  if (e == cudaSuccess) {
    a.counts = counts;
    cut_kernel<false><<<(unsigned)blocks, ZNLL_PREP_BLOCK, 0, st>>>(a);
    scan_kernel<<<1, 1024, 0, st>>>(a.counts, a.n_warps, a.counts + a.n_warps);
    cut_kernel<true><<<(unsigned)blocks, ZNLL_PREP_BLOCK, 0, st>>>(a);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(&total, counts + a.n_warps, sizeof(total), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    snprintf(g_prep_err, sizeof g_prep_err, "znll_cut_events: %s", cudaGetErrorString(e));
    return 1;
  }
  *n_kept = (int64_t)total;
  return 0;
}
