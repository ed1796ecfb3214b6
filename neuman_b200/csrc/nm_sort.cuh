// Per-warp bitonic sort of (key, index) pairs held in shared memory; index is the secondary key,
// so the order is total and stable.
#pragma once

__device__ __forceinline__ bool key_less(float za, int ia, float zb, int ib) {
  return (za < zb) || (za == zb && ia < ib);
}

// keys/idx: per-warp shared arrays of n (power of two) entries
__device__ __forceinline__ void nm_warp_bitonic_sort(float* keys, int* idx, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n >> 1); t += 32) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j cleared
        int l = i | j;
        bool up = ((i & k) == 0);
        float a = keys[i], b = keys[l];
        int ia = idx[i], ib = idx[l];
        bool swap = up ? key_less(b, ib, a, ia) : key_less(a, ia, b, ib);
        if (swap) { keys[i] = b; keys[l] = a; idx[i] = ib; idx[l] = ia; }
      }
      __syncwarp();
    }
  }
}

