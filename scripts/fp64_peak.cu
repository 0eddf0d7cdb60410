// fp64 peak microbenchmark (dev tool): DFMA vs DMMA (mma.sync f64) on sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a[i] * b + c;
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m8n8k4: A 8x4 (1 double/thread), B 4x8 (1), C 8x8 (2)
__global__ void k_dmma884(double *out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k16 f64 (sm_90+): A 16x16 (8 doubles/thread), B 16x8 (4), C 16x8 (4)
__global__ void k_dmma16816(double *out, int iters) {
  double c[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = 1.0 + i * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static double timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e-3;
}

int main() {
  double *out; cudaMalloc(&out, 8 * 148 * 16 * 256);
  const int iters = 20000;
  for (int bps : {2, 4, 8}) {
    const int grid = 148 * bps, block = 256;
    double t = timeit([&] { k_dfma<<<grid, block>>>(out, iters); });
    printf("DFMA      bps=%d: %.2f TFLOP/s\n", bps, 2.0 * 8 * iters * (double)grid * block / t / 1e12);
    t = timeit([&] { k_dmma884<<<grid, block>>>(out, iters); });
    printf("DMMA884   bps=%d: %.2f TFLOP/s\n", bps, 2.0 * 8 * 8 * 4 * 8 * iters * (double)grid * (block / 32) / t / 1e12);
    t = timeit([&] { k_dmma16816<<<grid, block>>>(out, iters); });
    printf("DMMA16816 bps=%d: %.2f TFLOP/s\n", bps, 2.0 * 16 * 8 * 16 * 4 * iters * (double)grid * (block / 32) / t / 1e12);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
