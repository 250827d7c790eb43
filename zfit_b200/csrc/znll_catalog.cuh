// znll_catalog.cuh -- helper to instantiate and register kernels ahead of time (nvcc, sm_100a).
#pragma once
#include <cstring>

#include "znll_device.cuh"
#include "znll_internal.h"

namespace znll {

template <class M, bool W, bool G>
cudaError_t launch_nll(const ZnllLaunchArgs& a, const double* consts, int grid, cudaStream_t stream) {
  KArgs<M::NC> k;
  for (int i = 0; i < ZNLL_MAXCOL; ++i) k.cols[i] = a.cols[i];
  k.w = a.w;
  k.n = a.n;
  k.log_offset = a.log_offset;
  k.log_offset2 = 2.0 * a.log_offset;
  k.partials = a.partials;
  k.ticket = a.ticket;
  k.out = a.out;
  static_assert(sizeof(Xchg) == sizeof(ZnllXchg), "Xchg layout");
  memcpy(&k.xc, &a.xc, sizeof(Xchg));
  for (int i = 0; i < M::NC; ++i) k.c[i] = consts[i];
  nll_kernel<M, W, G><<<grid, ZNLL_BLOCK, 0, stream>>>(k);
  return cudaGetLastError();
}

template <class M>
cudaError_t launch_pdf(const ZnllLaunchArgs& a, const double* consts, int grid, cudaStream_t stream) {
  KArgs<M::NC> k;
  for (int i = 0; i < ZNLL_MAXCOL; ++i) k.cols[i] = a.cols[i];
  k.w = a.w;
  k.n = a.n;
  k.log_offset = 0.0;
  k.log_offset2 = 0.0;
  k.partials = a.partials;
  k.ticket = a.ticket;
  k.out = a.out;
  memset(&k.xc, 0, sizeof(Xchg));
  for (int i = 0; i < M::NC; ++i) k.c[i] = consts[i];
  pdf_kernel<M><<<grid, ZNLL_BLOCK, 0, stream>>>(k);
  return cudaGetLastError();
}

template <class M, bool W>
cudaError_t launch_cov(const ZnllLaunchArgs& a, const double* consts, int grid, cudaStream_t stream) {
  KArgs<M::NC> k;
  for (int i = 0; i < ZNLL_MAXCOL; ++i) k.cols[i] = a.cols[i];
  k.w = a.w;
  k.n = a.n;
  k.log_offset = a.log_offset;
  k.log_offset2 = 2.0 * a.log_offset;
  k.partials = a.partials;
  k.ticket = a.ticket;
  k.out = a.out;
  memset(&k.xc, 0, sizeof(Xchg));
  for (int i = 0; i < M::NC; ++i) k.c[i] = consts[i];
  cov_kernel<M, W><<<grid, ZNLL_BLOCK, 0, stream>>>(k);
  return cudaGetLastError();
}

template <class M, bool W>
cudaError_t launch_hess(const ZnllLaunchArgs& a, const double* consts, int grid, cudaStream_t stream) {
  KArgs<M::NC> k;
  for (int i = 0; i < ZNLL_MAXCOL; ++i) k.cols[i] = a.cols[i];
  k.w = a.w;
  k.n = a.n;
  k.log_offset = a.log_offset;
  k.log_offset2 = 2.0 * a.log_offset;
  k.partials = a.partials;
  k.ticket = a.ticket;
  k.out = a.out;
  memset(&k.xc, 0, sizeof(Xchg));
  for (int i = 0; i < M::NC; ++i) k.c[i] = consts[i];
  hess_kernel<M, W><<<grid, ZNLL_BLOCK, 0, stream>>>(k);
  return cudaGetLastError();
}
// the Hessian pass exists for flat trees of nodes that implement Node::second, within the reduction's 256 accumulators
template <class M>
constexpr bool has_hess() {
  return M::HESS && 2 + (M::NS + 1) * (M::NS + 2) / 2 + M::NH <= 256;
}

template <class M>
ZnllCatalogEntry make_entry(const char* name) {
  ZnllCatalogEntry e;
  e.name = name;
  e.n_slots = M::NS;
  e.n_consts = M::NC;
  e.fn[0][0] = &launch_nll<M, false, false>;
  e.fn[0][1] = &launch_nll<M, false, true>;
  e.fn[1][0] = &launch_nll<M, true, false>;
  e.fn[1][1] = &launch_nll<M, true, true>;
  e.pdf = &launch_pdf<M>;
  e.cov[0] = &launch_cov<M, false>;
  e.cov[1] = &launch_cov<M, true>;
  e.hess[0] = e.hess[1] = nullptr;
  e.n_hess = 0;
  if constexpr (has_hess<M>()) {
    e.hess[0] = &launch_hess<M, false>;
    e.hess[1] = &launch_hess<M, true>;
    e.n_hess = M::NH;
  }
  return e;
}

struct Registrar {
  Registrar(const ZnllCatalogEntry* e, int n) { znll_catalog_register(e, n); }
};

}  // namespace znll

// the stringified type must be the canonical name the host generates (no spaces)
#define ZNLL_ENTRY(...) znll::make_entry<znll::__VA_ARGS__>(#__VA_ARGS__)
