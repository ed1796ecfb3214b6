// Backward (input-gradient) chain of NeRF.forward (models/vanilla.py:120-152) on the tensor cores:
// the adjoint of mlp_tc.cu's training forward, same machinery (tcgen05.mma cta_group::2 kind::f16, fp32
// TMEM accumulators, bulk-TMA weight ring, persistent CTA pairs, activations kept on chip).
//
// For every sample row, with g = dL/d raw (4 values) scaled by the loss scale S:
//   dVpre = (g_rgb @ Wrgb) * [V > 0]                         (epilogue threads, K = 3; [V > 0] from sign words)
//   b0: dF   = dVpre @ Wviews[:, :256]                        (K = 128)
//   b1: dX8  = dF @ Wfeature + g_alpha * w_alpha ;  dpre7 = dX8 * [X8 > 0]
//   b2..b8 (l = 7..1): dX_{l-1} = dpre_l @ W_l[:, -256:] ;   dpre_{l-1} = dX_{l-1} * [X_{l-1} > 0]
// Every dpre_l, dF and dVpre is written to HBM in fp16 by TMA stores from the A buffer (they are the left operands of the weight-gradient
// GEMMs dW_l = dpre_l^T @ X_{l-1}, done by the caller with the forward's activation stash), and stays in shared
// memory as the next step's A operand.  The ReLU masks come from the forward's 256-bit sign words.
// Gradients w.r.t. the sample positions are not produced (the trainers treat samples as constants).
#include "nm_internal.cuh"
#include "nm_pe.cuh"
#include "tc_common.cuh"

#define BW_STEPS 9
__host__ __device__ constexpr int bw_nkb(int b) { return b == 0 ? 2 : 4; }
#define BW_SLABS (2 + 8 * 4)

// rgb_linear.weight as [3][128] for the head (dV = g_rgb @ Wrgb): every thread needs all 384 values, with indices
// that are compile-time constants after unrolling, so they are folded into the FFMAs as constant-bank operands
// (read from shared memory this cost 384 LDS per thread and made the head the slowest step of a round: 17k cycles).
// Copied from the net's packed image before each launch (stream-ordered, device to device).
__constant__ float c_bw_wrgb[384];

template <int kPair>
struct BwCfg {
  static constexpr int NT = 2;
  static constexpr int NSLOT = kPair == 2 ? 5 : 3;
  static constexpr int SLOT_BYTES = 32768 / kPair;
  static constexpr int THREADS = 64 + 256;
  static constexpr int OFF_ACT = 0;
  static constexpr int OFF_RING = OFF_ACT + NT * 4 * TC_KB_BYTES;
  static constexpr int OFF_BAR = OFF_RING + NSLOT * SLOT_BYTES;
  static constexpr int N_BAR = 3 * NSLOT + 2 * NT;          // full peer_full empty | tmem_full act_ready
  static constexpr int OFF_TMEMPTR = OFF_BAR + 8 * N_BAR;
  static constexpr int OFF_CONST = (OFF_TMEMPTR + 16 + 127) & ~127;   // w_alpha[256] fp32
  static constexpr int SMEM_USED = OFF_CONST + 1024;
  static constexpr int SMEM_BYTES = SMEM_USED + 1024;
};

struct BwParams {
  const uint8_t* wimg;      // transposed weight slabs, kPair images back to back
  uint32_t image_bytes;
  const float* d_raw;       // [n][4] fp32 dL/d(r,g,b,sigma)
  const float* scale;       // device scalar: loss scale S (a power of two)
  const float* w_alpha;     // [256] fp32 alpha_linear.weight
  const uint32_t* st_m;     // [9][n][8] forward stash: ReLU sign words of pts_linears 0..7 and (plane 8, 4 words) of the views layer
  __half* g_pre;            // [8][n][256] out: S * dL/d(pre-activation of pts_linears l)
  __half* g_f;              // [n][256]    out: S * dL/d feature
  __half* g_v;              // [n][128]    out: S * dL/d(pre-activation of views_linears.0)
  long long n, n_tiles;
  CUtensorMap map_pre, map_f, map_v;   // TMA store maps of g_pre / g_f / g_v
  long long* trace;             // optional debug timeline (tools/tc_trace.py bwd): same layout as mlp_tc.cu's
};

// 16 accumulator columns [c0, c0+16) of one row: (+ rank-1 alpha term), ReLU mask from sign bits, pack,
// swizzled store into the next A operand (from where a TMA store takes it to the HBM gradient plane)
template <bool ALPHA, bool MASK>
__device__ __forceinline__ void bw_sub16(const uint32_t (&v)[16], int c0, float da, const float* s_walpha, uint32_t mbits,
                                         uint8_t* act, int row) {
  float x[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    x[e] = __uint_as_float(v[e]);
    if (ALPHA) x[e] = fmaf(da, s_walpha[c0 + e], x[e]);          // same address in every lane: a broadcast
  }
  if (MASK) {
    // sign word layout (mlp_tc.cu epi_sub16): bit k < 8 = column 2k, bit 8+k = column 2k+1; tested in bit order so
    // that the compiler moves them to predicates wholesale (R2P)
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (!((mbits >> k) & 1u)) x[k < 8 ? 2 * k : 2 * (k - 8) + 1] = 0.f;
  }
  uint32_t packed[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) packed[j] = pack_f16x2(x[2 * j], x[2 * j + 1], false);
  uint8_t* blk = act + (c0 >> 6) * TC_KB_BYTES + row * 128;
  const int ch0 = (c0 & 63) >> 3;
  *reinterpret_cast<uint4*>(blk + ((ch0 ^ (row & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  *reinterpret_cast<uint4*>(blk + (((ch0 + 1) ^ (row & 7)) << 4)) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
}

// drains 128 accumulator columns [cbase, cbase+128) of this thread's TMEM lane; mask = 128 sign bits
template <bool ALPHA, bool MASK>
__device__ __forceinline__ void bw_step(uint32_t t_lane, int cbase, float da, const float* s_walpha, const uint4& mask,
                                        uint8_t* act, int row) {
  uint32_t v0[16], v1[16];
  const uint32_t mw[4] = {mask.x, mask.y, mask.z, mask.w};
  tmem_ld16(t_lane + cbase, v0);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = cbase + 32 * q;
    tmem_wait_ld();
    tmem_ld16(t_lane + c + 16, v1);
    bw_sub16<ALPHA, MASK>(v0, c, da, s_walpha, mw[q] & 0xffffu, act, row);
    tmem_wait_ld();
    if (q < 3) tmem_ld16(t_lane + c + 32, v0);
    bw_sub16<ALPHA, MASK>(v1, c + 16, da, s_walpha, mw[q] >> 16, act, row);
  }
}

// debug timeline, as in mlp_tc.cu: role 0 = MMA issuer (0: operand ready seen, 1: step issued + committed), role 1 =
// epilogue warp 2 lane 0 (0: accumulator ready seen, 1: drained, 2: handed over); tile 0 of CTAs 0 / 1 only
#define BW_TRACE(role, ev, idx)                                                                             \
  do {                                                                                                      \
    if (P.trace && blockIdx.x < 2 && (idx) < 256) P.trace[((blockIdx.x * 2 + (role)) * 4 + (ev)) * 256 + (idx)] = clock64(); \
  } while (0)

template <int kPair>
__global__ void __launch_bounds__(BwCfg<kPair>::THREADS, 1) k_mlp_tc_bwd(const __grid_constant__ BwParams P) {
  using C = BwCfg<kPair>;
  constexpr int NT = C::NT, NSLOT = C::NSLOT;
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t raw_addr = smem_u32(smem_dyn);
  const uint32_t pad = (1024 - (raw_addr & 1023)) & 1023;          // SWIZZLE_128B atoms: 1024-byte aligned base
  uint8_t* smem = smem_dyn + pad;
  const uint32_t sbase = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kPair == 2 ? cluster_ctarank() : 0;
  const long long pair_id = blockIdx.x / kPair;
  const long long n_pairs = gridDim.x / kPair;

  auto bar_full = [&](int i) { return sbase + C::OFF_BAR + 8 * i; };
  auto bar_peer = [&](int i) { return sbase + C::OFF_BAR + 8 * (NSLOT + i); };
  auto bar_empty = [&](int i) { return sbase + C::OFF_BAR + 8 * (2 * NSLOT + i); };
  auto bar_tfull = [&](int t) { return sbase + C::OFF_BAR + 8 * (3 * NSLOT + t); };
  auto bar_aready = [&](int t) { return sbase + C::OFF_BAR + 8 * (3 * NSLOT + NT + t); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEMPTR);
  float* s_walpha = reinterpret_cast<float*>(smem + C::OFF_CONST);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(bar_full(i), 1); mbar_init(bar_peer(i), 1); mbar_init(bar_empty(i), 1); }
    for (int t = 0; t < NT; ++t) { mbar_init(bar_tfull(t), 1); mbar_init(bar_aready(t), 8 * kPair); }
    fence_mbar_init();
  }
  if (threadIdx.x >= 64) {
    const int e = threadIdx.x - 64;
    s_walpha[e] = __ldg(P.w_alpha + e);
  }
  if (warp == 1) {
    tmem_alloc<kPair>(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish<kPair>();
  }
  tc_fence_before();
  __syncthreads();
  if (kPair == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const long long tiles_per_round = n_pairs * NT;
  const long long n_rounds = (P.n_tiles + tiles_per_round - 1) / tiles_per_round;

  if (warp == 0) {
    // =============================== bulk-TMA producer ===============================
    if (lane == 0) {
      const uint8_t* img = P.wimg + (size_t)rank * P.image_bytes;
      uint32_t q = 0;
      for (long long round = 0; round < n_rounds; ++round)
        for (int k = 0; k < BW_SLABS; ++k, ++q) {
          const uint32_t slot = q % NSLOT, gen = q / NSLOT;
          mbar_wait(bar_empty(slot), (gen & 1) ^ 1);
          mbar_arrive_expect_tx(bar_full(slot), C::SLOT_BYTES);
          bulk_g2s(sbase + C::OFF_RING + slot * C::SLOT_BYTES, img + (size_t)k * C::SLOT_BYTES, C::SLOT_BYTES, bar_full(slot));
        }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ======================= MMA issuer (leader CTA): whole warp, warp-uniform, one elected lane issues =======================
      uint32_t issuer = 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\t.reg .b32 r;\n\telect.sync r|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(issuer));
      constexpr uint32_t idesc = make_idesc(128 * kPair, 256);
      uint32_t q0 = 0, nstep = 0;
      for (long long round = 0; round < n_rounds; ++round) {
        for (int b = 0; b < BW_STEPS; ++b, ++nstep) {
          const int nkb = bw_nkb(b);
          for (int t = 0; t < NT; ++t) {
            mbar_wait(bar_aready(t), nstep & 1);
            tc_fence_after();
            if (t == 0 && issuer) BW_TRACE(0, 0, nstep);
            const uint32_t d_tmem = tmem_base + t * 256;
            for (int kb = 0; kb < nkb; ++kb) {
              const uint32_t q = q0 + kb, slot = q % NSLOT, gen = q / NSLOT;
              if (t == 0) {
                mbar_wait(bar_full(slot), gen & 1);
                if (kPair == 2) mbar_wait(bar_peer(slot), gen & 1);
                tc_fence_after();
              }
              const uint64_t a_desc = make_desc(sbase + C::OFF_ACT + (t * 4 + kb) * TC_KB_BYTES);
              const uint64_t b_desc = make_desc(sbase + C::OFF_RING + slot * C::SLOT_BYTES);
              if (issuer) {
                umma_f16<kPair>(d_tmem, a_desc, b_desc, idesc, kb != 0);
                umma_f16<kPair>(d_tmem, a_desc + 2, b_desc + 2, idesc, 1);
                umma_f16<kPair>(d_tmem, a_desc + 4, b_desc + 4, idesc, 1);
                umma_f16<kPair>(d_tmem, a_desc + 6, b_desc + 6, idesc, 1);
                if (t == NT - 1) umma_commit<kPair>(bar_empty(slot));
              }
              __syncwarp();
            }
            if (issuer) {
              umma_commit<kPair>(bar_tfull(t));
              if (t == 0) BW_TRACE(0, 1, nstep);
            }
            __syncwarp();
          }
          q0 += nkb;
        }
      }
    } else {
      // =============================== relay (peer CTA of a pair) ===============================
      if (lane == 0) {
        uint32_t q = 0;
        for (long long round = 0; round < n_rounds; ++round)
          for (int k = 0; k < BW_SLABS; ++k, ++q) {
            const uint32_t slot = q % NSLOT, gen = q / NSLOT;
            mbar_wait(bar_full(slot), gen & 1);
            mbar_arrive_cluster(bar_peer(slot), 0);
          }
      }
    }
  } else {
    // ========================= epilogue: 8 warps, both warpgroups drain every tile (column halves) =========================
    const int ew = warp - 2;
    const int g = ew >> 2;                         // column half; also: the tile whose head (dVpre row) this thread builds
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t nstep = 0;
    const float S = __ldg(P.scale);

    auto publish = [&](int t) {
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar_aready(t), 0);
    };
    auto sample_index = [&](long long round, int t) { return (((round * n_pairs + pair_id) * NT + t) * kPair + rank) * 128 + row; };
    auto tile_valid = [&](long long round, int t) {
      return ((round * n_pairs + pair_id) * NT + t) < P.n_tiles && sample_index(round, t) < P.n;
    };
    // head of tile g, round r: dVpre row (128 channels) in registers, computed one round ahead
    uint32_t head[64];
    // The head needs, per row, dL/d raw (16 B) and the 128 ReLU sign bits of the views layer (16 B): both are
    // fetched a few steps before they are used (prefetch_head), so building the head is pure arithmetic.
    float4 gr_next = make_float4(0.f, 0.f, 0.f, 0.f);
    uint4 vm_next = make_uint4(0, 0, 0, 0);
    auto prefetch_head = [&](long long round) {
      gr_next = make_float4(0.f, 0.f, 0.f, 0.f);
      vm_next = make_uint4(0, 0, 0, 0);
      if (tile_valid(round, g)) {
        const long long i = sample_index(round, g);
        gr_next = __ldg(reinterpret_cast<const float4*>(P.d_raw) + i);
        vm_next = __ldg(reinterpret_cast<const uint4*>(P.st_m + ((size_t)8 * P.n + i) * 8));
      }
    };
    auto build_head = [&]() {
      const float gx = gr_next.x * S, gy = gr_next.y * S, gz = gr_next.z * S;
      const uint32_t vw[4] = {vm_next.x, vm_next.y, vm_next.z, vm_next.w};
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        // columns 8j .. 8j+7 live in word j/4, 16-bit half (j/2)%2; pair p of that group: bit p = even column, bit 8+p = odd
        const uint32_t bits16 = vw[j >> 2] >> (16 * ((j >> 1) & 1));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = 8 * j + 2 * k, pr = 4 * (j & 1) + k;
          float x0 = fmaf(gx, c_bw_wrgb[c], fmaf(gy, c_bw_wrgb[128 + c], gz * c_bw_wrgb[256 + c]));
          float x1 = fmaf(gx, c_bw_wrgb[c + 1], fmaf(gy, c_bw_wrgb[129 + c], gz * c_bw_wrgb[257 + c]));
          if (!((bits16 >> pr) & 1u)) x0 = 0.f;
          if (!((bits16 >> (8 + pr)) & 1u)) x1 = 0.f;
          head[4 * j + k] = pack_f16x2(x0, x1, false);
        }
      }
    };
    auto store_head = [&](int t) {                  // 128 channels -> A k-blocks 0 and 1 of tile t
      uint8_t* act = smem + C::OFF_ACT + t * 4 * TC_KB_BYTES;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(act + kb * TC_KB_BYTES + row * 128 + ((j ^ (row & 7)) << 4)) =
              make_uint4(head[32 * kb + 4 * j], head[32 * kb + 4 * j + 1], head[32 * kb + 4 * j + 2], head[32 * kb + 4 * j + 3]);
    };
    auto load_mask = [&](long long round, int t, int plane) {
      if (!tile_valid(round, t)) return make_uint4(0, 0, 0, 0);
      return __ldg(reinterpret_cast<const uint4*>(P.st_m + ((size_t)plane * P.n + sample_index(round, t)) * 8) + (g ^ t));
    };
    if (n_rounds > 0) { prefetch_head(0); build_head(); }

    // Column ownership: for tile t this thread drains columns [(g ^ t) * 128, +128) = A k-blocks 2(g^t), 2(g^t)+1 of its
    // row in every step.  The warpgroup that builds the head of tile t (g == t) is therefore the one that owns k-blocks
    // 0 and 1 of that tile, and every write / TMA read of a slice of the A buffer is ordered inside one warp.
    for (long long round = 0; round < n_rounds; ++round) {
      float da[NT];
      uint4 mnext[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        if (g == t) {
          // the last step of the previous round stored this slice by TMA: it must have been read
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          store_head(t);
          fence_async_smem();
          __syncwarp();
          const long long i0 = sample_index(round, t) - lane;
          if (lane == 0 && i0 < P.n) {                  // dL/d(views pre-activation) leaves from the A buffer as well
            const uint32_t src = sbase + C::OFF_ACT + t * 4 * TC_KB_BYTES + quad * 32 * 128;
            tma_store_3d(&P.map_v, src, 0, (int)i0, 0);
            tma_store_3d(&P.map_v, src + TC_KB_BYTES, 64, (int)i0, 0);
            tma_store_commit();
          }
        }
        publish(t);
        da[t] = tile_valid(round, t) ? S * __ldg(P.d_raw + 4 * sample_index(round, t) + 3) : 0.f;
        mnext[t] = load_mask(round, t, 7);          // step b1 masks with [X8 > 0] = plane 7
      }
      for (int b = 0; b < BW_STEPS; ++b, ++nstep) {
        uint4 mcur[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) mcur[t] = mnext[t];
        if (b >= 1 && b < BW_STEPS - 1) {           // masks of step b+1: plane 8-(b+1), one full step ahead
#pragma unroll
          for (int t = 0; t < NT; ++t) mnext[t] = load_mask(round, t, 7 - b);
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          uint8_t* act = smem + C::OFF_ACT + t * 4 * TC_KB_BYTES;
          const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + t * 256;
          const int gc = g ^ t;                       // column half of this thread for tile t
          mbar_wait(bar_tfull(t), nstep & 1);
          tc_fence_after();
          if (t == 0 && ew == 0 && lane == 0) BW_TRACE(1, 0, nstep);
          // this warp's slice is the source of the TMA store it issued one step ago (the other tile's may still
          // fly); in step 0 the head's store of this round may be the most recent one: wait for everything
          if (lane == 0) { if (b == 0) tma_store_wait_read<0>(); else tma_store_wait_read<1>(); }
          __syncwarp();
          if (b == 0) bw_step<false, false>(t_lane, gc * 128, 0.f, s_walpha, mcur[t], act, row);
          else if (b == 1) bw_step<true, true>(t_lane, gc * 128, da[t], s_walpha, mcur[t], act, row);
          else bw_step<false, true>(t_lane, gc * 128, 0.f, s_walpha, mcur[t], act, row);
          if (t == 0 && ew == 0 && lane == 0) BW_TRACE(1, 1, nstep);
          const long long i0 = sample_index(round, t) - lane;              // first row of this warp
          const bool issue = lane == 0 && i0 < P.n;
          const uint32_t src = sbase + C::OFF_ACT + (t * 4 + 2 * gc) * TC_KB_BYTES + quad * 32 * 128;
          const CUtensorMap* m = b == 0 ? &P.map_f : &P.map_pre;
          const int plane = b == 0 ? 0 : 8 - b;
          if (b < BW_STEPS - 1) {
            publish(t);                               // the MMA thread first: the store and the next MMAs only read
          } else {
            fence_async_smem();                       // last step: nothing to hand over, only the store
            tc_fence_before();
            __syncwarp();
          }
          if (issue) {
            tma_store_3d(m, src, 128 * gc, (int)i0, plane);
            tma_store_3d(m, src + TC_KB_BYTES, 128 * gc + 64, (int)i0, plane);
            tma_store_commit();
          }
          if (t == 0 && ew == 0 && lane == 0) BW_TRACE(1, 2, nstep);
        }
        if (b == 1 && round + 1 < n_rounds) prefetch_head(round + 1);
        if (b == 4 && round + 1 < n_rounds) build_head();
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (kPair == 2) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kPair>(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// Packing: transposed weights -> fp16 slabs [N = input channel][K = output channel], 128B-swizzled
// ---------------------------------------------------------------------------------------------
// Sources are the context-owned fp32 copies of the weights in the "Wt[k = input][n = output]" layout (api.cu)
struct BwPackSrc {
  const float* wt[8]; const float* feat_t; const float* views_t;
};

// slab k of the image: (step b, k-block kb); element (n = input channel, kk = output channel inside the k-block)
__device__ __forceinline__ float bw_src_weight(const BwPackSrc& S, int b, int n, int kb, int kk) {
  const int out = kb * 64 + kk;
  if (b == 0) return S.views_t[(size_t)n * NM_VIEWS_HID + out];                 // views input = [feature(256), dir PE]
  if (b == 1) return S.feat_t[(size_t)n * 256 + out];
  const int l = 9 - b;                                                          // b = 2..8 -> layer 7..1
  return S.wt[l][(size_t)((l == 5 ? NM_POS_PE : 0) + n) * 256 + out];           // layer 5 input = [PE(63), hidden]
}

__global__ void k_bw_pack(BwPackSrc S, int kpair, uint32_t image_bytes, __half* __restrict__ out, const float* __restrict__ rgb_t,
                          float* __restrict__ wrgb) {
  const int rank = blockIdx.y;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (rank == 0 && e < 384) wrgb[e] = rgb_t[(e & 127) * 3 + (e >> 7)];      // [128][3] (api.cu layout) -> [3][128]
  if (e * 2 >= image_bytes) return;
  const uint32_t slot_bytes = 32768u / kpair;
  const uint32_t byte = (uint32_t)(e * 2);
  const int k = byte / slot_bytes;                     // slab index
  const int b = k < 2 ? 0 : 1 + (k - 2) / 4;
  const int kb = k < 2 ? k : (k - 2) % 4;
  const uint32_t in_slab = byte - k * slot_bytes;
  const int n_local = in_slab >> 7;
  const int chunk = ((in_slab & 127) >> 4) ^ (n_local & 7);
  const int kk = chunk * 8 + ((in_slab & 15) >> 1);
  const int n = rank * (256 / kpair) + n_local;
  out[(size_t)rank * (image_bytes / 2) + e] = __float2half_rn(bw_src_weight(S, b, n, kb, kk));
}

// ---------------------------------------------------------------------------------------------
// Bias gradients: column sums of fp16 gradient planes, fp32 accumulation.  HBM-bound (reads each plane once).
// grid = (row chunks, planes); a block owns `width` columns (2 per thread) and strides over its rows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_colsum_f16(const __half* __restrict__ src, long long n, int width,
                                                    long long rows_per_block, float* __restrict__ out) {
  const int plane = blockIdx.y;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(n, r0 + rows_per_block);
  const int c2 = threadIdx.x;                       // column pair
  if (2 * c2 >= width) return;
  const __half2* p = reinterpret_cast<const __half2*>(src + ((size_t)plane * n + r0) * width) + c2;
  const size_t stride = (size_t)width / 2;
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  long long r = r0;
  for (; r + 4 <= r1; r += 4) {                     // 4 independent loads in flight per thread
    const float2 v0 = __half22float2(p[0]), v1 = __half22float2(p[stride]);
    const float2 v2 = __half22float2(p[2 * stride]), v3 = __half22float2(p[3 * stride]);
    a0 += v0.x; a1 += v0.y; b0 += v1.x; b1 += v1.y;
    a0 += v2.x; a1 += v2.y; b0 += v3.x; b1 += v3.y;
    p += 4 * stride;
  }
  for (; r < r1; ++r) { const float2 v = __half22float2(p[0]); a0 += v.x; a1 += v.y; p += stride; }
  atomicAdd(out + (size_t)plane * width + 2 * c2, a0 + b0);
  atomicAdd(out + (size_t)plane * width + 2 * c2 + 1, a1 + b1);
}

int nm_impl_colsum_f16(nm_ctx* ctx, const __half* src, int planes, int64_t n, int width, float* out, cudaStream_t st) {
  NM_CHECK_CUDA(ctx, cudaMemsetAsync(out, 0, (size_t)planes * width * sizeof(float), st));
  const int chunks = max(1, (ctx->sm_count * 16) / max(planes, 1));
  long long rows_per_block = (n + chunks - 1) / chunks;
  if (rows_per_block < 64) rows_per_block = 64;
  dim3 grid((unsigned)((n + rows_per_block - 1) / rows_per_block), planes);
  k_colsum_f16<<<grid, 128, 0, st>>>(src, n, width, rows_per_block, out);
  NM_CHECK_LAUNCH(ctx);
  return NM_OK;
}

// ---------------------------------------------------------------------------------------------
// Adjoint of Embedder.forward (models/vanilla.py:82-92): d_x = J^T d_enc with
//   posenc: enc = [x, sin(f_k x), cos(f_k x)]_k          rotate: enc = [x, sin(x B^T), cos(x B^T)]
// One thread per sample; sin/cos recomputed in fp32 exactly as the fp32 forward does (nm_pe_pair).
// d_enc rows are `ld` floats apart and carry a scale whose inverse is *inv_scale (device scalar, may be null).
// ---------------------------------------------------------------------------------------------
__global__ void k_pe_backward(NmPeSpec pe, const float* __restrict__ x, long long group, const float* __restrict__ d_enc,
                              int ld, const float* __restrict__ inv_scale, long long n, float* __restrict__ d_x) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long xi = group > 0 ? i / group : i;
  const float xv[3] = {x[3 * xi], x[3 * xi + 1], x[3 * xi + 2]};
  const float* g = d_enc + (size_t)i * ld;
  float dx[3] = {g[0], g[1], g[2]};
  const int nq = 3 * pe.n_freqs;
  for (int q = 0; q < nq; ++q) {
    float sn, cs;
    int c_sin, c_cos;
    nm_pe_pair(pe, xv, q, sn, cs, c_sin, c_cos);
    const float w = cs * g[c_sin] - sn * g[c_cos];
    if (pe.kind == NM_PE_ROTATE) {
      const float* b = pe.table + 3 * q;
      dx[0] = fmaf(w, b[0], dx[0]); dx[1] = fmaf(w, b[1], dx[1]); dx[2] = fmaf(w, b[2], dx[2]);
    } else {
      const int k = q / 3, d = q - 3 * k;
      dx[d] = fmaf(w, pe.table[k], dx[d]);
    }
  }
  const float sc = inv_scale ? *inv_scale : 1.f;
  d_x[3 * i] = dx[0] * sc; d_x[3 * i + 1] = dx[1] * sc; d_x[3 * i + 2] = dx[2] * sc;
}

int nm_impl_pe_backward(nm_ctx* ctx, const NmNet& net, int which, const float* x, int64_t group, const float* d_enc, int ld,
                        const float* inv_scale, int64_t n, float* d_x, cudaStream_t st) {
  NmPeSpec pe = which == 0 ? NmPeSpec{net.desc.pos_pe_kind, net.desc.pos_n_freqs, net.f32 + net.o_pos_bv}
                           : NmPeSpec{net.desc.dir_pe_kind, net.desc.dir_n_freqs, net.f32 + net.o_dir_bv};
  k_pe_backward<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pe, x, group, d_enc, ld, inv_scale, n, d_x);
  NM_CHECK_LAUNCH(ctx);
  return NM_OK;
}

static int bw_pair_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("NEUMAN_TC_PAIR");
    mode = (e && e[0] == '1') ? 1 : 2;
  }
  return mode;
}

int nm_tc_pack_bwd(nm_ctx* ctx, NmNet& net, cudaStream_t st) {
  const int kpair = bw_pair_mode();
  const uint32_t image_bytes = (uint32_t)BW_SLABS * (32768u / kpair);
  const size_t halfs = (size_t)kpair * image_bytes / 2;
  if (!net.f16_bwd) NM_CHECK_CUDA(ctx, cudaMalloc(&net.f16_bwd, halfs * sizeof(__half)));
  if (!net.bw_wrgb) NM_CHECK_CUDA(ctx, cudaMalloc(&net.bw_wrgb, 384 * sizeof(float)));
  BwPackSrc S;
  for (int l = 0; l < 8; ++l) S.wt[l] = net.f32 + net.o_pts_w[l];
  S.feat_t = net.f32 + net.o_feat_w; S.views_t = net.f32 + net.o_views_w;
  dim3 grid((unsigned)((image_bytes / 2 + 255) / 256), kpair);
  k_bw_pack<<<grid, 256, 0, st>>>(S, kpair, image_bytes, net.f16_bwd, net.f32 + net.o_rgb_w, net.bw_wrgb);
  NM_CHECK_LAUNCH(ctx);
  net.bwd_packed = true;
  return NM_OK;
}

template <int kPair>
static int launch_bwd(nm_ctx* ctx, const BwParams& P, cudaStream_t st) {
  using C = BwCfg<kPair>;
  NM_SET_SMEM_ONCE(ctx, (k_mlp_tc_bwd<kPair>), C::SMEM_BYTES);
  int ctas = ctx->sm_count - (ctx->sm_count % kPair);
  long long need = (P.n_tiles * kPair + C::NT - 1) / C::NT;
  if (need < ctas) ctas = (int)((need + kPair - 1) / kPair * kPair);
  if (ctas < kPair) ctas = kPair;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NM_CHECK_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_mlp_tc_bwd<kPair>, P));
  NM_LAUNCHED(ctx);
  return NM_OK;
}

int nm_tc_backward(nm_ctx* ctx, NmNet& net, const float* d_raw, const float* scale, int64_t n, const __half* st_v,
                   const uint32_t* st_m, __half* g_pre, __half* g_f, __half* g_v, cudaStream_t st) {
  const int kpair = bw_pair_mode();
  if (!net.bwd_packed) {
    int rc = nm_tc_pack_bwd(ctx, net, st);
    if (rc != NM_OK) return rc;
  }
  BwParams P;
  P.wimg = reinterpret_cast<const uint8_t*>(net.f16_bwd);
  P.image_bytes = (uint32_t)BW_SLABS * (32768u / kpair);
  P.d_raw = d_raw; P.scale = scale;
  P.w_alpha = net.f32 + net.o_alpha_w;
  NM_CHECK_CUDA(ctx, cudaMemcpyToSymbolAsync(c_bw_wrgb, net.bw_wrgb, 384 * sizeof(float), 0, cudaMemcpyDeviceToDevice, st));
  P.st_m = st_m;
  P.g_pre = g_pre; P.g_f = g_f; P.g_v = g_v;
  P.n = n;
  P.n_tiles = (n + 128 * kpair - 1) / (128 * kpair);
  P.trace = nullptr;
  if (const char* e = getenv("NEUMAN_TC_TRACE_BWD")) P.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  if (n >= (int64_t)0x7fff0000) NM_FAIL(ctx, NM_ERR_INVALID, "nm_mlp_backward: n too large for one call");
  if (tc_make_store_map(&P.map_pre, g_pre, 8, (uint64_t)n, 256) || tc_make_store_map(&P.map_f, g_f, 1, (uint64_t)n, 256) ||
      tc_make_store_map(&P.map_v, g_v, 1, (uint64_t)n, 128))
    NM_FAIL(ctx, NM_ERR_CUDA, "nm_mlp_backward: cuTensorMapEncodeTiled failed");
  return kpair == 2 ? launch_bwd<2>(ctx, P, st) : launch_bwd<1>(ctx, P, st);
}
