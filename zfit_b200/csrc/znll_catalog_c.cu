// Ahead-of-time kernel catalogue, part C: multi-dimensional products (sm_100a).
#include "znll_catalog.cuh"
using namespace znll;

static const ZnllCatalogEntry kEntries[] = {
    ZNLL_ENTRY(Prod<Gauss<0>,Expo<1>,Cheb<2,4>>),
    ZNLL_ENTRY(Prod<Gauss<0>,Gauss<1>>),
    ZNLL_ENTRY(Prod<Gauss<0>,Gauss<1>,Gauss<2>>),
    ZNLL_ENTRY(Prod<Gauss<0>,Expo<1>>),
    ZNLL_ENTRY(Prod<Gauss<0>,Cheb<1,2>>),
    ZNLL_ENTRY(Sum<Prod<Gauss<0>,Expo<1>>,Prod<Cheb<0,1>,Expo<1>>>),
};
static Registrar reg_c(kEntries, sizeof(kEntries) / sizeof(kEntries[0]));

// sum of weights
cudaError_t znll_launch_sumw(const double* w, long long n, double* partials, unsigned int* ticket, double* out,
                             int grid, cudaStream_t stream) {
  sumw_kernel<0><<<grid, ZNLL_BLOCK, 0, stream>>>(w, n, partials, ticket, out);
  return cudaGetLastError();
}
