// znll_internal.h -- shared between the host side (znll_host.cu) and the kernel catalogue TUs.
#pragma once
#include <cuda_runtime.h>

struct ZnllXchg {  // == znll::Xchg
  double* buf[8];
  double* acc;
  unsigned long long seq;
  int rank, world, total, stride;
  double* hout;
  unsigned long long* hflag;
};

struct ZnllLaunchArgs {  // == the prefix of znll::KArgs<NC> (the NVRTC path copies it verbatim)
  const double* cols[8];
  const double* w;
  long long n;
  double log_offset;
  double log_offset2;  // 2 * log_offset: the pair loop subtracts it as a constant-bank operand
  double* partials;
  unsigned int* ticket;
  double* out;
  ZnllXchg xc;
};

// launches nll_kernel<M, WEIGHTED, GRAD> with `grid` blocks of 256 threads on `stream`
typedef cudaError_t (*znll_launcher_t)(const ZnllLaunchArgs&, const double* consts, int grid, cudaStream_t stream);

struct ZnllCatalogEntry {
  const char* name;        // canonical type name, e.g. "Sum<CBall<0>,Expo<0>>"
  int n_slots, n_consts;
  znll_launcher_t fn[2][2];  // [weighted][grad]
  znll_launcher_t pdf;       // per-event pdf values into args.partials
  znll_launcher_t cov[2];    // weighted-covariance pass [weighted]
  znll_launcher_t hess[2];   // exact Hessian pass [weighted]; null when the tree is not flat / a node lacks Node::second
  int n_hess;                // second-order leaf accumulators of that pass (M::NH)
};

// each catalogue TU registers its instantiations at static-init time
void znll_catalog_register(const ZnllCatalogEntry* entries, int n);
const ZnllCatalogEntry* znll_catalog_find(const char* name);
int znll_catalog_size();
const ZnllCatalogEntry* znll_catalog_at(int i);

cudaError_t znll_launch_sumw(const double* w, long long n, double* partials, unsigned int* ticket, double* out,
                             int grid, cudaStream_t stream);

#ifndef ZNLL_BLOCKS_PER_SM
#define ZNLL_BLOCKS_PER_SM 2  // must match ZNLL_MIN_BLOCKS in znll_device.cuh
#endif

// test hook: out[i] = f(in[i]) with f = fast_exp (0), fast_log (1), fast_rcp (2) of znll_device.cuh
cudaError_t znll_launch_test_math(int kind, const double* in, double* out, long long n, cudaStream_t stream);
