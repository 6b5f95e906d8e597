// scan.cu -- exclusive prefix sum over int32 (cell histogram -> cell_start).
// Hand-written three-phase scan: per-block scan of 2048 elements, recursive
// scan of the block totals, uniform add.  In-place safe.
#include "store.cuh"

namespace drrt {

static constexpr int SCAN_T = 256;
static constexpr int SCAN_E = 8;  // elements per thread
static constexpr int SCAN_B = SCAN_T * SCAN_E;

__global__ void scan_block_kernel(const int32_t *in, int32_t *out, int64_t n, int32_t *totals) {
  __shared__ int32_t warp_sum[SCAN_T / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_E;
  int32_t v[SCAN_E];
  int32_t local = 0;
#pragma unroll
  for (int e = 0; e < SCAN_E; e++) {
    int64_t i = base + e;
    v[e] = (i < n) ? in[i] : 0;
    local += v[e];
  }
  // inclusive scan of per-thread sums across the block
  int32_t inc = local;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int32_t ws = (lane < SCAN_T / 32) ? warp_sum[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, ws, o);
      if (lane >= o) ws += t;
    }
    if (lane < SCAN_T / 32) warp_sum[lane] = ws;
  }
  __syncthreads();
  int32_t excl = inc - local + (w ? warp_sum[w - 1] : 0);
#pragma unroll
  for (int e = 0; e < SCAN_E; e++) {
    int64_t i = base + e;
    if (i < n) out[i] = excl;
    excl += v[e];
  }
  if (threadIdx.x == SCAN_T - 1 && totals) totals[blockIdx.x] = excl;
}

__global__ void scan_add_kernel(int32_t *out, int64_t n, const int32_t *block_off) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_E;
  const int32_t add = block_off[blockIdx.x];
#pragma unroll
  for (int e = 0; e < SCAN_E; e++) {
    int64_t i = base + e;
    if (i < n) out[i] += add;
  }
}

static int scan_rec(const int32_t *in, int32_t *out, int64_t n, int32_t *tmp, cudaStream_t st) {
  int64_t nb = (n + SCAN_B - 1) / SCAN_B;
  scan_block_kernel<<<(unsigned)nb, SCAN_T, 0, st>>>(in, out, n, nb > 1 ? tmp : nullptr);
  DRRT_LAUNCHED();
  if (nb > 1) {
    DRRT_CHECK(scan_rec(tmp, tmp, nb, tmp + nb, st));
    scan_add_kernel<<<(unsigned)nb, SCAN_T, 0, st>>>(out, n, tmp);
    DRRT_LAUNCHED();
  }
  return DRRT_OK;
}

int exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, DevBuf<int32_t> &tmp,
                       cudaStream_t st) {
  if (n <= 0) return DRRT_OK;
  // block totals of every recursion level: n/2048 + n/2048^2 + ... (+ slack)
  size_t need = 0;
  for (int64_t m = (n + SCAN_B - 1) / SCAN_B; m > 1; m = (m + SCAN_B - 1) / SCAN_B) need += m;
  need += 16;
  DRRT_CHECK(tmp.reserve(need, st));
  DRRT_CHECK(scan_rec(in, out, n, tmp.p, st));
  DRRT_CUDA(cudaGetLastError());
  return DRRT_OK;
}

}  // namespace drrt
