// Inverse-CDF hierarchical sampling.  Compiled with -fmad=false.
//
//   nm_sample_pdf          <- utils/ray_utils.py:164-194 (sample_pdf)
//   nm_importance_samples  <- utils/ray_utils.py:138-160 (ray_to_importance_samples, det=True)
//
// One warp per ray.  The cdf lives in shared memory; each lane inverts it for its own u values by
// binary search (searchsorted right=True); the union with the old samples is sorted by a warp
// bitonic network (torch.sort at ray_utils.py:152).
#include "nm_internal.cuh"
#include "nm_sort.cuh"

#define FULL 0xffffffffu
#define RS_WARPS 4

// cdf[0..B-1] for one ray from weights[0..B-2]; ATen's CPU cumsum accumulates float input in
// double (acc_type<float,false>) and rounds each prefix to float, reproduced here.
__device__ __forceinline__ void build_cdf(const float* __restrict__ w_in, int B, float* cdf, int lane) {
  const int nw = B - 1;
  double part = 0.0;
  for (int i = lane; i < nw; i += 32) part += (double)(w_in[i] + 1e-5f);      // weights + 1e-5 (:166)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
  const float total = (float)part;                                           // torch.sum (:167)
  double carry = 0.0;
  if (lane == 0) cdf[0] = 0.f;                                               // leading zero (:169)
  for (int base = 0; base < nw; base += 32) {
    int i = base + lane;
    double v = (i < nw) ? (double)((w_in[i] + 1e-5f) / total) : 0.0;         // pdf (:167)
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(FULL, inc, o);
      if (lane >= o) inc += t;
    }
    if (i < nw) cdf[i + 1] = (float)(carry + inc);                           // cumsum (:168)
    carry += __shfl_sync(FULL, inc, 31);
  }
  __syncwarp();
}

__device__ __forceinline__ float invert_cdf(const float* cdf, const float* bins, int B, float u) {
  // searchsorted(cdf, u, right=True): number of entries <= u   (:181)
  int lo = 0, hi = B;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  int below = max(0, lo - 1), above = min(B - 1, lo);                         // (:182-183)
  float c0 = cdf[below], c1 = cdf[above];
  float b0 = bins[below], b1 = bins[above];
  float den = c1 - c0;
  if (den < 1e-5f) den = 1.f;                                                 // (:190-191)
  float t = (u - c0) / den;
  return b0 + t * (b1 - b0);                                                  // (:193)
}

__global__ void __launch_bounds__(32 * RS_WARPS) k_sample_pdf(const float* __restrict__ bins,
                                                                const float* __restrict__ weights, long long R, int B,
                                                                int N, const float* __restrict__ u_in,
                                                                float* __restrict__ out) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* cdf = sm + (size_t)wid * 2 * B;
  float* sb = cdf + B;
  long long r = (long long)blockIdx.x * RS_WARPS + wid;
  if (r >= R) return;
  build_cdf(weights + r * (B - 1), B, cdf, lane);
  for (int i = lane; i < B; i += 32) sb[i] = bins[r * B + i];
  __syncwarp();
  for (int j = lane; j < N; j += 32) {
    float u = u_in ? u_in[r * N + j] : nm_linspace01(j, N);                   // det=True (:173-174)
    out[r * N + j] = invert_cdf(cdf, sb, B, u);
  }
}

extern "C" int nm_sample_pdf(nm_ctx* ctx, const float* bins, const float* weights, int64_t R, int32_t B, int32_t N,
                             const float* u, float* out, void* stream) {
  NM_ENTER(ctx);
  if (R == 0) return NM_OK;
  if (!bins || !weights || !out || R < 0 || B < 2 || N <= 0) NM_FAIL(ctx, NM_ERR_INVALID, "nm_sample_pdf: bad argument");
  size_t smem = (size_t)RS_WARPS * 2 * B * sizeof(float);
  if (smem > 48 * 1024) NM_FAIL(ctx, NM_ERR_UNSUPPORTED, "nm_sample_pdf: too many bins");
  unsigned blocks = (unsigned)((R + RS_WARPS - 1) / RS_WARPS);
  k_sample_pdf<<<blocks, 32 * RS_WARPS, smem, (cudaStream_t)stream>>>(bins, weights, R, B, N, u, out);
  NM_CHECK_LAUNCH(ctx);
  return NM_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * RS_WARPS) k_importance(
    const float* __restrict__ origins, const float* __restrict__ dirs, const float* __restrict__ z,
    const float* __restrict__ weights, long long R, int S, int N, int including_old, int pow2,
    float* __restrict__ pts, float* __restrict__ dirs_out, float* __restrict__ z_out) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int B = S - 1;
  // per-warp layout: cdf[B] | mids[B] | keys[pow2] | idx[pow2]
  float* cdf = sm + (size_t)wid * (2 * B + 2 * pow2);
  float* mids = cdf + B;
  float* keys = mids + B;
  int* idx = reinterpret_cast<int*>(keys + pow2);
  long long r = (long long)blockIdx.x * RS_WARPS + wid;
  if (r >= R) return;
  const float* zr = z + r * S;
  for (int i = lane; i < B; i += 32) mids[i] = 0.5f * (zr[i + 1] + zr[i]);    // z_vals_mid (:147)
  build_cdf(weights + r * S + 1, B, cdf, lane);                               // weights[..., 1:-1] (:148)
  const int total = including_old ? S + N : N;
  if (including_old)
    for (int i = lane; i < S; i += 32) { keys[i] = zr[i]; idx[i] = i; }
  const int off = including_old ? S : 0;
  for (int j = lane; j < N; j += 32) {
    keys[off + j] = invert_cdf(cdf, mids, B, nm_linspace01(j, N));
    idx[off + j] = off + j;
  }
  for (int i = total + lane; i < pow2; i += 32) { keys[i] = INFINITY; idx[i] = i; }
  __syncwarp();
  if (including_old) {                                                         // torch.sort (:152)
    // Both lists are normally already ascending (z_vals by construction, the inverse-CDF samples because u
    // ascends): then the sorted union is a rank merge -- rank(old i) = i + #{new < z_i}, rank(new j) = j +
    // #{old <= z_j} -- instead of a full sort.  Any inversion (1-ulp rounding) falls back to the bitonic network.
    bool sorted = true;
    for (int i = lane; i + 1 < S; i += 32) sorted = sorted && (keys[i] <= keys[i + 1]);
    for (int j = lane; j + 1 < N; j += 32) sorted = sorted && (keys[S + j] <= keys[S + j + 1]);
    if (__all_sync(FULL, sorted)) {
      float* merged = reinterpret_cast<float*>(idx);
      for (int i = lane; i < S; i += 32) {
        const float v = keys[i];
        int lo = 0, hi = N;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (keys[S + mid] < v) lo = mid + 1; else hi = mid; }
        merged[i + lo] = v;
      }
      for (int j = lane; j < N; j += 32) {
        const float v = keys[S + j];
        int lo = 0, hi = S;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (keys[mid] <= v) lo = mid + 1; else hi = mid; }
        merged[j + lo] = v;
      }
      __syncwarp();
      for (int i = lane; i < total; i += 32) keys[i] = merged[i];
      __syncwarp();
    } else {
      nm_warp_bitonic_sort(keys, idx, pow2, lane);
    }
  }
  float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0;
  if (pts || dirs_out) {
    dx = dirs[3 * r]; dy = dirs[3 * r + 1]; dz = dirs[3 * r + 2];
    if (pts) { ox = origins[3 * r]; oy = origins[3 * r + 1]; oz = origins[3 * r + 2]; }
  }
  for (int i = lane; i < total; i += 32) {
    float zv = keys[i];
    long long o = r * total + i;
    z_out[o] = zv;
    if (pts) { pts[3 * o] = ox + dx * zv; pts[3 * o + 1] = oy + dy * zv; pts[3 * o + 2] = oz + dz * zv; }   // (:155)
    if (dirs_out) { dirs_out[3 * o] = dx; dirs_out[3 * o + 1] = dy; dirs_out[3 * o + 2] = dz; }
  }
}

extern "C" int nm_importance_samples(nm_ctx* ctx, const float* origins, const float* dirs, const float* z,
                                     const float* weights, int64_t R, int32_t S, int32_t N, int32_t including_old,
                                     float* pts, float* dirs_out, float* z_out, void* stream) {
  NM_ENTER(ctx);
  if (R == 0) return NM_OK;
  if (!z || !weights || !z_out || R < 0 || S < 3 || N <= 0 || ((pts || dirs_out) && (!origins || !dirs)))
    NM_FAIL(ctx, NM_ERR_INVALID, "nm_importance_samples: bad argument (needs S >= 3)");
  int total = including_old ? S + N : N;
  int pow2 = 32;
  while (pow2 < total) pow2 <<= 1;
  size_t smem = (size_t)RS_WARPS * (2 * (S - 1) + 2 * pow2) * sizeof(float);
  if (smem > 200 * 1024) NM_FAIL(ctx, NM_ERR_UNSUPPORTED, "nm_importance_samples: too many samples per ray");
  if (smem > 48 * 1024)
    NM_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_importance, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned blocks = (unsigned)((R + RS_WARPS - 1) / RS_WARPS);
  k_importance<<<blocks, 32 * RS_WARPS, smem, (cudaStream_t)stream>>>(origins, dirs, z, weights, R, S, N, including_old,
                                                                        pow2, pts, dirs_out, z_out);
  NM_CHECK_LAUNCH(ctx);
  return NM_OK;
}
