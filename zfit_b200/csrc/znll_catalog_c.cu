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
    ZNLL_ENTRY(GCBall<0>),
    ZNLL_ENTRY(Sum<GCBall<0>,Expo<0>>),
    ZNLL_ENTRY(Sum<Gauss<0>,Legd<0,2>>),
    ZNLL_ENTRY(Sum<Gauss<0>,Chb2<0,2>>),
    ZNLL_ENTRY(Sum<GETail<0>,Expo<0>>),
    ZNLL_ENTRY(Sum<GGETail<0>,Expo<0>>),
    ZNLL_ENTRY(Sum<Gauss<0>,Bern<0,3>>),
    ZNLL_ENTRY(Sum<TGauss<0>,Expo<0>>),
};
static Registrar reg_c(kEntries, sizeof(kEntries) / sizeof(kEntries[0]));

// sum of weights
cudaError_t znll_launch_sumw(const double* w, long long n, double* partials, unsigned int* ticket, double* out,
                             int grid, cudaStream_t stream) {
  sumw_kernel<0><<<grid, ZNLL_BLOCK, 0, stream>>>(w, n, partials, ticket, out);
  return cudaGetLastError();
}

__global__ void test_math_kernel(int kind, const double* __restrict__ in, double* __restrict__ out, long long n) {
  __shared__ MathTabs tabs;
  init_tabs(tabs);
  __syncthreads();
  const Ctx cx = ctx_of(tabs);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double x = in[i];
    out[i] = kind == 0 ? fast_exp(x, cx) : (kind == 1 ? fast_log(x, cx) : fast_rcp(x));
  }
}
cudaError_t znll_launch_test_math(int kind, const double* in, double* out, long long n, cudaStream_t stream) {
  test_math_kernel<<<296, 256, 0, stream>>>(kind, in, out, n);
  return cudaGetLastError();
}
