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
// rows and walks it twice -- pass 1 (count_kernel) counts its events per bucket, a single-block scan turns the
// [bucket][warp] counts into exclusive offsets (bucket-major, so bucket b of warp w starts where bucket b of warp w-1
// ended), pass 2 (move_kernel) moves the events, a group of rows at a time through shared memory so that the stores
// are runs of consecutive addresses.  The rank of an event among the events of its bucket comes from five ballots (one
// per bucket bit) and running per-bucket counts, so the partition is STABLE and therefore deterministic: the same input
// gives the same order, and the NLL's fixed-order reduction stays bitwise reproducible across binds.
//
// The reference has no counterpart (tf.where evaluates both branches for every event, models/physics.py:31-45); this
// replaces the torch.sort call of round 1's CudaEngine._sort_for_coherence.
#include <cstdint>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

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

// Lanes of a row that fall into bucket `b` (a 5-bit number), among the `live` ones: five ballots, one per bit.  (The
// first version used __match_any_sync: MATCH.ANY kept the count pass at 96 % SM utilisation and 1.7 TB/s.)
__device__ __forceinline__ unsigned lanes_in_bucket(const unsigned (&bit)[5], unsigned live, int b) {
  unsigned m = live;
#pragma unroll
  for (int k = 0; k < 5; ++k) m &= ((b >> k) & 1) ? bit[k] : ~bit[k];
  return m;
}
__device__ __forceinline__ void bucket_ballots(unsigned (&bit)[5], int b) {
#pragma unroll
  for (int k = 0; k < 5; ++k) bit[k] = __ballot_sync(0xffffffffu, (b >> k) & 1);
}

// Pass 1: events per (warp, bucket).  Lane L of a warp keeps the count of bucket L in a register.
__global__ void __launch_bounds__(ZNLL_PREP_BLOCK) count_kernel(const __grid_constant__ PrepArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * ZNLL_PREP_WARPS + (threadIdx.x >> 5);
  const long long rows = (a.n + 31) >> 5;
  long long r = (long long)warp * a.rows_per_warp;
  const long long r_end = min(rows, r + a.rows_per_warp);
  const double* __restrict__ keycol = a.src[a.key];
  unsigned int cnt = 0;
  for (; r < r_end; r += 8) {  // eight key loads in flight per lane (the pass shares the move's two CTAs per SM)
    double x[8];
    bool live[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long i = ((r + k) << 5) + lane;
      live[k] = (r + k) < r_end && i < a.n;
      x[k] = live[k] ? __ldg(keycol + i) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      unsigned bit[5];
      bucket_ballots(bit, bucket_of(x[k], a.lo, a.scale));
      cnt += __popc(lanes_in_bucket(bit, __ballot_sync(0xffffffffu, live[k]), lane));
    }
  }
  a.counts[(size_t)lane * a.n_warps + warp] = cnt;
}

// Pass 2: move.  A warp ranks R rows (R = 16: 512 events) at a time, lays them out bucket-major in shared memory and
// writes each bucket's run with consecutive lanes on consecutive addresses: full 32-byte sectors instead of the ~23
// partially written sectors per 32-lane store of the first version (every lane its own bucket cursor; ncu: 1.25 GB read
// and 1.18 GB written for 0.8 GB of payload, 1.6 ms per stream -- partial-sector writes are read-modify-writes in L2).
// Lane L holds the global cursor of bucket L in a register; ranks inside the 512 events come from the ballots and the
// running per-bucket counts, so the order inside a bucket is the original order (stable) as before.
template <int R>
struct MoveStage {
  double vals[R * 32];
  unsigned int gbase[ZNLL_PREP_BUCKETS];  // bucket's global position minus its local offset: position = gbase[b] + slot (mod 2^32)
  unsigned char sb[R * 32];  // bucket of the event staged in each slot
};

template <int R>
__global__ void __launch_bounds__(ZNLL_PREP_BLOCK, 32 / R) move_kernel(const __grid_constant__ PrepArgs a) {
  __shared__ MoveStage<R> stage_of[ZNLL_PREP_WARPS];
  MoveStage<R>& st = stage_of[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int warp = blockIdx.x * ZNLL_PREP_WARPS + (threadIdx.x >> 5);
  const long long rows = (a.n + 31) >> 5;
  long long r = (long long)warp * a.rows_per_warp;
  const long long r_end = min(rows, r + a.rows_per_warp);
  const double* __restrict__ keycol = a.src[a.key];
  unsigned int cur = a.counts[(size_t)lane * a.n_warps + warp];  // where bucket `lane` of this warp continues
  for (; r < r_end; r += R) {
    // per event of the group: rank among the group's events of its bucket (bits 0-15), bucket (16-20), live (31);
    // after the scan: its staging slot, or ~0 for a lane without event
    unsigned int slot[R];
    unsigned int cnt = 0;  // lane L: events of bucket L in this group so far
    {
      double x[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const long long i = ((r + k) << 5) + lane;
        const bool live = (r + k) < r_end && i < a.n;
        x[k] = live ? __ldg(keycol + i) : 0.0;
        slot[k] = live ? 0x80000000u : 0u;
      }
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int b = bucket_of(x[k], a.lo, a.scale);
        unsigned bit[5];
        bucket_ballots(bit, b);
        const unsigned livemask = __ballot_sync(0xffffffffu, (slot[k] >> 31) != 0u);
        const unsigned before = __shfl_sync(0xffffffffu, cnt, b);  // earlier rows of the group, same bucket
        slot[k] |= (before + __popc(lanes_in_bucket(bit, livemask, b) & lt)) | ((unsigned)b << 16);
        cnt += __popc(lanes_in_bucket(bit, livemask, lane));
      }
    }
    unsigned int incl = cnt;  // bucket-major offsets inside the group
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const unsigned int excl = incl - cnt;
    const int total = (int)__shfl_sync(0xffffffffu, incl, 31);
    st.gbase[lane] = cur - excl;
    cur += cnt;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int b = (int)((slot[k] >> 16) & 31u);
      const unsigned int off = __shfl_sync(0xffffffffu, excl, b);
      if (slot[k] >> 31) {
        slot[k] = (slot[k] & 0xffffu) + off;
        st.sb[slot[k]] = (unsigned char)b;
      } else {
        slot[k] = 0xffffffffu;
      }
    }
    // every stream, the key among them (its second read comes from L1/L2: the group was loaded a moment ago)
#pragma unroll 1
    for (int s = 0; s < a.n_streams; ++s) {
      const double* __restrict__ sc = a.src[s];
      double v[R];
#pragma unroll
      for (int k = 0; k < R; ++k)
        v[k] = (slot[k] != 0xffffffffu) ? __ldg(sc + ((r + k) << 5) + lane) : 0.0;
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (slot[k] != 0xffffffffu) st.vals[slot[k]] = v[k];
      __syncwarp();
      double* __restrict__ sd = a.dst[s];
      for (int j = lane; j < total; j += 32) sd[st.gbase[st.sb[j]] + (unsigned int)j] = st.vals[j];
      __syncwarp();
    }
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

// NC: columns held in registers (n_cols <= NC; instantiated for 1, 2, 4 and ZNLL_MAX_COLS)
template <bool MOVE, int NC>
__global__ void __launch_bounds__(ZNLL_PREP_BLOCK) cut_kernel(const __grid_constant__ CutArgs a) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int warp = blockIdx.x * ZNLL_PREP_WARPS + (threadIdx.x >> 5);
  unsigned int cursor = MOVE ? a.counts[warp] : 0u;
  const long long rows = (a.n + 31) >> 5;
  long long r = (long long)warp * a.rows_per_warp;
  const long long r_end = min(rows, r + a.rows_per_warp);
  for (; r < r_end; r += 4) {  // four rows of every column in flight per lane
    double x[4][NC];
    bool keep[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = ((r + k) << 5) + lane;
      keep[k] = (r + k) < r_end && i < a.n;
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c < a.n_cols) x[k][c] = keep[k] ? __ldg(a.src[c] + i) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = ((r + k) << 5) + lane;
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c < a.n_cols) keep[k] = keep[k] && x[k][c] >= a.lo[c] && x[k][c] <= a.hi[c];  // NaN fails both comparisons
      if (a.less_a && keep[k]) keep[k] = __ldg(a.less_a + i) < __ldg(a.less_b + i);
      const unsigned kept = __ballot_sync(0xffffffffu, keep[k]);
      if (MOVE && keep[k]) {
        const size_t pos = (size_t)cursor + __popc(kept & lt);
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if (c < a.n_cols) a.dst[c][pos] = x[k][c];
        if (a.n_streams > a.n_cols) a.dst[a.n_cols][pos] = __ldg(a.src[a.n_cols] + i);
      }
      cursor += __popc(kept);
    }
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

// resident CTAs per SM of a kernel of this file (registers and static shared memory decide): the grids are exactly one
// wave, a larger grid would leave a partial second wave (measured on the first version of the cut: 1184 CTAs for 888 slots)
template <class K>
static int resident_blocks(K kernel) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, ZNLL_PREP_BLOCK, 0) != cudaSuccess || nb < 1) nb = 1;
  return nb;
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
  // rows a warp moves at a time: 16 (measured 0.87 / 1.30 / 1.97 ms for 1e8 events of 1 / 2 / 3 streams against 0.94 / 1.88 /
  // 3.11 ms with 8, whose shorter bucket runs leave more partial sectors; ZNLL_PREP_R=8 keeps the A/B possible)
  static const int R = (getenv("ZNLL_PREP_R") && atoi(getenv("ZNLL_PREP_R")) == 8) ? 8 : 16;
  static const int per_sm = std::min(resident_blocks(count_kernel), R == 16 ? resident_blocks(move_kernel<16>) : resident_blocks(move_kernel<8>));  // both passes share the layout
  long long blocks = (long long)sms * per_sm;
  const long long groups = (rows + R - 1) / R;
  if (blocks * ZNLL_PREP_WARPS > groups) blocks = (groups + ZNLL_PREP_WARPS - 1) / ZNLL_PREP_WARPS;
  a.n_warps = (int)(blocks * ZNLL_PREP_WARPS);
  a.rows_per_warp = ((groups + a.n_warps - 1) / a.n_warps) * R;
  cudaStream_t st = (cudaStream_t)stream;
  // the [bucket][warp] counters: see counts_buffer.  Calls on one device must be ordered by the caller (same stream, or a
  // synchronisation in between); the engine does both.
  const size_t m = (size_t)ZNLL_PREP_BUCKETS * a.n_warps;
  unsigned int* counts = nullptr;
  if (e == cudaSuccess) e = counts_buffer(device, m, sms, &counts);
  if (e == cudaSuccess) {
    a.counts = counts;
    count_kernel<<<(unsigned)blocks, ZNLL_PREP_BLOCK, 0, st>>>(a);
    scan_kernel<<<1, 1024, 0, st>>>(a.counts, (int)m, nullptr);
    if (R == 16) move_kernel<16><<<(unsigned)blocks, ZNLL_PREP_BLOCK, 0, st>>>(a);
    else move_kernel<8><<<(unsigned)blocks, ZNLL_PREP_BLOCK, 0, st>>>(a);
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
  typedef void (*CutFn)(const CutArgs);
  struct Variant {
    CutFn count, move;
    int per_sm;
  };
  static Variant variants[4] = {{cut_kernel<false, 1>, cut_kernel<true, 1>, 0},
                                {cut_kernel<false, 2>, cut_kernel<true, 2>, 0},
                                {cut_kernel<false, 4>, cut_kernel<true, 4>, 0},
                                {cut_kernel<false, ZNLL_MAX_COLS>, cut_kernel<true, ZNLL_MAX_COLS>, 0}};
  Variant& v = variants[n_cols <= 1 ? 0 : n_cols <= 2 ? 1 : n_cols <= 4 ? 2 : 3];
  if (v.per_sm == 0) v.per_sm = std::min(resident_blocks(v.count), resident_blocks(v.move));
  long long blocks = (long long)sms * v.per_sm;
  const long long groups = (rows + 3) / 4;  // a warp handles four rows at a time
  if (blocks * ZNLL_PREP_WARPS > groups) blocks = (groups + ZNLL_PREP_WARPS - 1) / ZNLL_PREP_WARPS;
  a.n_warps = (int)(blocks * ZNLL_PREP_WARPS);  // a multiple of 8: the scan reads 16 bytes at a time
  a.rows_per_warp = ((groups + a.n_warps - 1) / a.n_warps) * 4;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* counts = nullptr;
  if (e == cudaSuccess) e = counts_buffer(device, (size_t)a.n_warps + 4, sms, &counts);
  unsigned int total = 0;
  if (e == cudaSuccess) {
    a.counts = counts;
    v.count<<<(unsigned)blocks, ZNLL_PREP_BLOCK, 0, st>>>(a);
    scan_kernel<<<1, 1024, 0, st>>>(a.counts, a.n_warps, a.counts + a.n_warps);
    v.move<<<(unsigned)blocks, ZNLL_PREP_BLOCK, 0, st>>>(a);
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
