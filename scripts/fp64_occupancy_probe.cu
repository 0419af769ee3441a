// FP64 pipe throughput of one B200 as a function of resident warps per SM and independent DFMA chains per thread (ILP):
// the practical ceiling for kernels that can only keep 8-12 warps per SM resident (255-168 registers per thread).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/fp64_probe scripts/fp64_occupancy_probe.cu && build/fp64_probe
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
double run(int warpsPerSM, int sms, double* out) {
  // one CTA of 32 threads per resident warp; dynamic shared memory limits residency to exactly warpsPerSM CTAs per SM
  const int iters = 2000;
  size_t smem = (size_t)(227 * 1024) / warpsPerSM - 1024;
  smem = smem / 1024 * 1024;
  cudaFuncSetAttribute(k<ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    k<ILP><<<sms * warpsPerSM, 32, smem>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * 16 * ILP * (double)iters * sms * warpsPerSM * 32 / (ms * 1e-3) / 1e12;
    if (rep && tf > best) best = tf;
  }
  return best;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 64 * 32);
  const int ws[] = {4, 8, 9, 10, 11, 12, 16, 32};
  printf("TFLOP/s (DFMA = 2) on %s, %d SMs; rows = resident warps per SM, columns = independent DFMA chains per thread\n", p.name, p.multiProcessorCount);
  printf("warps/SM      ILP1      ILP2      ILP3      ILP4      ILP6      ILP8\n");
  for (int w : ws) {
    printf("%8d  %8.2f  %8.2f  %8.2f  %8.2f  %8.2f  %8.2f\n", w, run<1>(w, p.multiProcessorCount, out), run<2>(w, p.multiProcessorCount, out),
           run<3>(w, p.multiProcessorCount, out), run<4>(w, p.multiProcessorCount, out), run<6>(w, p.multiProcessorCount, out), run<8>(w, p.multiProcessorCount, out));
  }
  return 0;
}
