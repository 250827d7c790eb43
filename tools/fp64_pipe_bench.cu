// fp64_pipe_bench.cu -- what the B200 FP64 pipe sustains for different operand mixes / ILP / occupancy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe_bench tools/fp64_pipe_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int ILP>
__global__ void __launch_bounds__(256) k(double* out, const double* in, int iters, double a, double b) {
  double x[ILP], y[ILP], z[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    x[i] = in[(threadIdx.x + i) & 255];
    y[i] = in[(threadIdx.x + 2 * i + 1) & 255] * 0.5;
    z[i] = in[(threadIdx.x + 3 * i + 2) & 255] * 1e-7;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (MODE == 0) x[i] = fma(x[i], a, b);           // 1 register operand + 2 constants(1 const + 1 reg in practice)
        if (MODE == 1) x[i] = fma(x[i], y[i], b);        // 2 distinct register operands + constant
        if (MODE == 2) x[i] = fma(x[i], y[i], z[i]);     // 3 distinct register operands
        if (MODE == 3) x[i] = x[i] * y[i];               // DMUL 2 regs
        if (MODE == 4) x[i] = x[i] + y[i];               // DADD 2 regs
        if (MODE == 5) x[i] = fma(y[i], z[i], x[i]);     // accumulate form, 3 regs
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int ILP>
double run(int bps, double* d_out, double* d_in) {
  int dev = 0;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  const int grid = prop.multiProcessorCount * bps;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int iters = 4000;
  k<MODE, ILP><<<grid, 256>>>(d_out, d_in, 100, 0.999999, 1e-7);
  cudaEventRecord(e0);
  k<MODE, ILP><<<grid, 256>>>(d_out, d_in, iters, 0.999999, 1e-7);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = 16.0 * ILP * iters * (double)grid * 256;
  return ops / (ms * 1e-3) / 1e12;  // tera DP-instructions per second (x2 = TFLOP/s for FMA)
}

int main() {
  double *d_out, *d_in;
  cudaMalloc(&d_out, sizeof(double) * 148 * 8 * 256 * 2);
  cudaMalloc(&d_in, sizeof(double) * 256);
  double h[256];
  for (int i = 0; i < 256; ++i) h[i] = 1.0 + i * 1e-3;
  cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice);
  const char* names[] = {"fma(x,c,c)", "fma(x,y,c)", "fma(x,y,z)", "mul(x,y)", "add(x,y)", "fma(y,z,x)"};
  printf("# tera DP instr/s (peak = 148 SM * 64 lanes * clk); columns: blocks/SM (x8 warps) \n");
  printf("%-12s %4s %8s %8s %8s %8s\n", "mode", "ILP", "bps=1", "bps=2", "bps=3", "bps=4");
#define ROW(M, I)                                                                                          \
  printf("%-12s %4d %8.2f %8.2f %8.2f %8.2f\n", names[M], I, run<M, I>(1, d_out, d_in), run<M, I>(2, d_out, d_in), \
         run<M, I>(3, d_out, d_in), run<M, I>(4, d_out, d_in));
  ROW(0, 1) ROW(0, 2) ROW(0, 4) ROW(0, 8)
  ROW(1, 1) ROW(1, 2) ROW(1, 4) ROW(1, 8)
  ROW(2, 1) ROW(2, 2) ROW(2, 4) ROW(2, 8)
  ROW(5, 2) ROW(5, 4)
  ROW(3, 2) ROW(3, 4)
  ROW(4, 2) ROW(4, 4)
  return 0;
}
