// Micro-benchmark of the accumulator drain (tcgen05.ld -> registers -> f16 -> shared memory) on sm_100a:
// how many cycles does it take NW warps of one CTA to read 128 lanes x 256 columns of TMEM, by load shape,
// loads in flight and amount of per-element work?  Bounds the epilogue of k_mlp_tc (DESIGN.md "two coincident rooflines").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_bench tools/tmem_bench.cu && ./tmem_bench
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

struct BiasTab { __align__(16) float b[2560]; };

// WORK 0: xor only.  1: + bias (register constant), relu-pack, swizzled STS.128.  2: as 1 with the bias from shared memory
// (4 x LDS.128 broadcast per 16 columns).  3: as 1 plus one HMNMX2 per packed register (range tracking).
// 4: bias from the kernel-parameter constant bank at a warp-uniform runtime offset (LDCU -> uniform registers, FADD R, R, UR).
// 5: no bias at all (bias folded into the MMAs).
template <int WORK>
__device__ __forceinline__ uint32_t consume16(const uint32_t (&v)[16], int c0, uint8_t* act, int row, const float* sbias, uint32_t& rng,
                                          const BiasTab& tab, int crow) {
  if (WORK == 0) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) x ^= v[i];
    return x;
  }
  uint32_t packed[8];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float4 b = make_float4(0.25f, 0.5f, 0.75f, 1.f);
    if (WORK == 2) b = *reinterpret_cast<const float4*>(sbias + c0 + 4 * g);
    if (WORK == 4) b = make_float4(tab.b[crow + c0 + 4 * g], tab.b[crow + c0 + 4 * g + 1], tab.b[crow + c0 + 4 * g + 2], tab.b[crow + c0 + 4 * g + 3]);
    float x0 = __uint_as_float(v[4 * g]), x1 = __uint_as_float(v[4 * g + 1]);
    float x2 = __uint_as_float(v[4 * g + 2]), x3 = __uint_as_float(v[4 * g + 3]);
    if (WORK != 5) { x0 += b.x; x1 += b.y; x2 += b.z; x3 += b.w; }
    packed[2 * g] = pack2(x0, x1);
    packed[2 * g + 1] = pack2(x2, x3);
  }
  if (WORK == 3) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&rng), *reinterpret_cast<const __half2*>(&packed[j]));
      rng = *reinterpret_cast<const uint32_t*>(&m);
    }
  }
  uint8_t* blk = act + (c0 >> 6) * 16384 + row * 128;
  const int ch0 = (c0 & 63) >> 3;
  *reinterpret_cast<uint4*>(blk + ((ch0 ^ (row & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  *reinterpret_cast<uint4*>(blk + (((ch0 + 1) ^ (row & 7)) << 4)) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
  return packed[0];
}

// NW epilogue warps (8 or 16): warp w reads TMEM lane quadrant w % 4, column slice (w / 4) of width 256 / (NW / 4).
template <int NW, int WORK, int DEPTH, int SHAPE>
__global__ void __launch_bounds__(NW * 32 + 32, 1) k_drain(const __grid_constant__ BiasTab tab, int iters, long long* cycles, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(16) float sbias[256];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  if (threadIdx.x < 256) sbias[threadIdx.x] = 0.001f * threadIdx.x;
  if (warp == NW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_ptr;
  uint32_t acc = 0, rng = 0;
  long long t0 = 0, t1 = 0;
  if (warp < NW) {
    const int quad = warp & 3, slice = warp >> 2;
    constexpr int NC = 256 / (NW / 4);
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tbase + ((uint32_t)(quad * 32) << 16);
    asm volatile("bar.sync 1, %0;" ::"r"(NW * 32) : "memory");
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t tl = t_lane + (it & 1) * 256;
      uint8_t* act = smem + (it & 1) * 65536;
      const int cb = slice * NC;
      const int crow = __shfl_sync(0xffffffffu, (it % 10) * 256, 0);
      if (SHAPE == 16) {
        uint32_t v[DEPTH][16];
#pragma unroll
        for (int d = 0; d < DEPTH - 1; ++d) tmem_ld16(tl + cb + 16 * d, v[d]);
#pragma unroll
        for (int q = 0; q < NC / 16; ++q) {
          tmem_wait_ld();        // everything issued so far has landed (loads complete in order)
          if (q + DEPTH - 1 < NC / 16) tmem_ld16(tl + cb + 16 * (q + DEPTH - 1), v[(q + DEPTH - 1) % DEPTH]);
          acc ^= consume16<WORK>(v[q % DEPTH], cb + 16 * q, act, row, sbias, rng, tab, crow);
        }
      } else {
        uint32_t v[2][32];
        tmem_ld32(tl + cb, v[0]);
#pragma unroll
        for (int q = 0; q < NC / 32; ++q) {
          tmem_wait_ld();
          if (q + 1 < NC / 32) tmem_ld32(tl + cb + 32 * (q + 1), v[(q + 1) & 1]);
          uint32_t lo[16], hi[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { lo[i] = v[q & 1][i]; hi[i] = v[q & 1][16 + i]; }
          acc ^= consume16<WORK>(lo, cb + 32 * q, act, row, sbias, rng, tab, crow);
          acc ^= consume16<WORK>(hi, cb + 32 * q + 16, act, row, sbias, rng, tab, crow);
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(NW * 32) : "memory");
    t1 = clock64();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (acc == 0x12345678u || rng == 0x7777u) sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == NW) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
  }
}

template <int NW, int WORK, int DEPTH, int SHAPE>
void run(const char* name, long long* d_cyc, uint32_t* d_sink) {
  const int iters = 400;
  auto kern = k_drain<NW, WORK, DEPTH, SHAPE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536);
  static BiasTab tab;
  for (int i = 0; i < 2560; ++i) tab.b[i] = 0.001f * (i % 97);
  for (int rep = 0; rep < 2; ++rep) kern<<<148, NW * 32 + 32, 2 * 65536>>>(tab, iters, d_cyc, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
  printf("%-58s %8.0f cycles per 128x256 tile  (%.1f B/clk)  %s\n", name, (double)cyc / iters, 131072.0 * iters / cyc,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_cyc; uint32_t* d_sink;
  cudaMalloc(&d_cyc, 8); cudaMalloc(&d_sink, 4096);
  run<8, 0, 2, 16>("8 warps  x16 depth 2  xor only", d_cyc, d_sink);
  run<8, 0, 4, 16>("8 warps  x16 depth 4  xor only", d_cyc, d_sink);
  run<8, 0, 2, 32>("8 warps  x32 depth 2  xor only", d_cyc, d_sink);
  run<16, 0, 2, 16>("16 warps x16 depth 2  xor only", d_cyc, d_sink);
  run<8, 1, 2, 16>("8 warps  x16 depth 2  +bias(reg) relu f16 STS", d_cyc, d_sink);
  run<8, 1, 3, 16>("8 warps  x16 depth 3  +bias(reg) relu f16 STS", d_cyc, d_sink);
  run<8, 1, 4, 16>("8 warps  x16 depth 4  +bias(reg) relu f16 STS", d_cyc, d_sink);
  run<8, 1, 2, 32>("8 warps  x32 depth 2  +bias(reg) relu f16 STS", d_cyc, d_sink);
  run<8, 2, 2, 16>("8 warps  x16 depth 2  +bias(LDS) relu f16 STS", d_cyc, d_sink);
  run<8, 2, 3, 16>("8 warps  x16 depth 3  +bias(LDS) relu f16 STS", d_cyc, d_sink);
  run<8, 3, 2, 16>("8 warps  x16 depth 2  +bias(reg) relu f16 STS + range", d_cyc, d_sink);
  run<8, 4, 2, 16>("8 warps  x16 depth 2  +bias(param bank, uniform) relu f16 STS", d_cyc, d_sink);
  run<8, 5, 2, 16>("8 warps  x16 depth 2  no bias (folded into MMA) relu f16 STS", d_cyc, d_sink);
  run<16, 1, 2, 16>("16 warps x16 depth 2  +bias(reg) relu f16 STS", d_cyc, d_sink);
  run<16, 2, 2, 16>("16 warps x16 depth 2  +bias(LDS) relu f16 STS", d_cyc, d_sink);
  run<16, 3, 2, 16>("16 warps x16 depth 2  +bias(reg) relu f16 STS + range", d_cyc, d_sink);
  return 0;
}
