// fp64_peak.cu -- measures the FP64 pipe of the GPU the benchmark runs on: the roofline
// denominator of the edge-check kernels (K4/K5 are FP64-issue bound; MEASURED_PEAKS.json
// carries HBM and bf16 figures only).  Not product code: built into benchmarks/libfp64peak.so by
// __graft_entry__.build() / benchmarks/fp64_peak.py and called from bench.py.
//
//   kind 0: DFMA   -- 8 independent a = a * b + c chains per thread (2 flop per instruction)
//   kind 1: DMUL + DADD, not contracted -- what the library issues (it is built with -fmad=false
//           for bit parity with the reference's x86-64 build: 1 flop per instruction)
// Returns the best of `reps` timed launches in TFLOP/s (CUDA events on the launching stream,
// after a warm-up launch); instructions/s = TFLOP/s / (2 or 1).
#include <cuda_runtime.h>
#include <stdint.h>

template <int KIND>
__global__ void __launch_bounds__(256) fp64_chain_kernel(double *out, int iters, double b, double c) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (KIND == 0) a[k] = __fma_rn(a[k], b, c);
      else a[k] = __dadd_rn(__dmul_rn(a[k], b), c);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k];
  if (s == 123.456) out[0] = s;  // never true: keeps the chains alive
}

extern "C" __attribute__((visibility("default")))
double fp64_peak_tflops(int device, int kind, int reps) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  double *out = nullptr;
  if (cudaMalloc(&out, 8) != cudaSuccess) return -1.0;
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int iters = 1 << 15, blocks = sms * 8, threads = 256;
  double best = 0.0;
  for (int r = 0; r < reps + 1; r++) {
    cudaEventRecord(a, st);
    if (kind == 0) fp64_chain_kernel<0><<<blocks, threads, 0, st>>>(out, iters, 0.999999999, 1e-9);
    else fp64_chain_kernel<1><<<blocks, threads, 0, st>>>(out, iters, 0.999999999, 1e-9);
    cudaEventRecord(b, st);
    if (cudaEventSynchronize(b) != cudaSuccess) { best = -1.0; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    const double flop = 2.0 * 8.0 * (double)iters * (double)blocks * threads;  // mul + add per chain step
    const double tf = flop / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaStreamDestroy(st);
  cudaFree(out);
  return best;
}
