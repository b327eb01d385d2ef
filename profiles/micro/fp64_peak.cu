// FP64 peaks of one B200, measured (north_star's tensor-core clause):
//   * DFMA: 8 independent fused multiply-add chains per thread (vector FP64 pipe)
//   * DMMA: mma.sync.m8n8k4.f64 (FP64 tensor path), 4 independent accumulator tiles per warp
//   * the 1-D sum-factorisation contraction of the GLS kernel as DMMA: a [n x n] matrix applied to 8 lines
//     per instruction needs k = n padded to 4 / 8 and m = n padded to 8, i.e. only n^2 / (8 * 4 * ceil(n/4))
//     of the issued flops are useful.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters, double a, double b)
{
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        x[i] = fma(x[i], a, b);
    }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double *out, int iters, double a0, double b0)
{
  double c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
  double a = a0 + (threadIdx.x & 7) * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1])
                     : "d"(a), "d"(b));
    }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f, int reps)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r)
    {
      cudaEventRecord(e0);
      f();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
  return best;
}

int main()
{
  int dev = 0, sms = 0, clk = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  double *out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  const int iters = 1 << 15;
  for (int bps = 1; bps <= 8; bps *= 2)
    {
      const int    grid = sms * bps;
      const float  t1   = time_ms([&] { k_dfma<<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      const double f1   = 2.0 * 8 * (double)iters * grid * 256 / (t1 * 1e-3) / 1e12;
      const float  t2   = time_ms([&] { k_dmma<<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      const double f2   = 2.0 * 256 * 4 * (double)iters * grid * 8 / (t2 * 1e-3) / 1e12;
      printf("warps/SM %2d : DFMA %.2f TFLOP/s   DMMA m8n8k4 %.2f TFLOP/s\n", bps * 8, f1, f2);
    }
  printf("sms %d  clock %.0f MHz (max)\n", sms, clk / 1e3);
  printf("useful share of a DMMA-issued 1-D contraction (n x n applied to 8 lines): ");
  for (int n = 3; n <= 5; ++n)
    printf("n=%d: %.0f%%  ", n, 100.0 * n * n / (8.0 * 4 * ((n + 3) / 4)));
  printf("\n");
  if (cudaDeviceSynchronize() != cudaSuccess)
    return 1;
  return 0;
}
