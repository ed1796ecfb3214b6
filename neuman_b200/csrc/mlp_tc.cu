// Tensor-core implementation of Joiner.forward (positional encoding + 8x256 NeRF MLP,
// models/vanilla.py:82-92,:120-152,:162-166) for sm_100a: tcgen05.mma (kind::f16, fp16 operands,
// fp32 accumulation in TMEM), weights streamed by bulk-TMA (cp.async.bulk) through a shared-memory
// ring, activations kept on chip between layers, warp-specialised roles, persistent CTAs.
//
// fp16 operands carry the same 11-bit significand as TF32, at twice the tensor rate; accumulation,
// bias, ReLU, the alpha head and the encodings are fp32 (DESIGN.md "Numerics").
//
// Work decomposition
//   tile      = 128 consecutive samples per CTA (one TMEM lane per sample).  With kPair == 2 two
//               CTAs of a cluster form a cta_group::2 pair: one 256-sample pair-tile, each CTA owns
//               128 rows of A / D and HALF of every weight slab (N/2 rows of B).
//   step      = one GEMM of the network: 0: L0 (K=64 PE) | 1-4: L1-4 | 5: L5 (PE block + 4 act
//               blocks, "input first", :131) | 6,7: L6,L7 (+alpha head in the epilogue of 7, :135) |
//               8: feature (:136) | 9: views layer (4 feature blocks + dir-PE block, N=128, :137-141) |
//               10: rgb (N=16, 3 used, :143).
//   slab      = one 64-wide K block of one step's weights for this CTA: [N_cta rows][128 B],
//               128B-swizzled, K-major -- exactly the UMMA canonical layout, pre-packed in HBM so one
//               cp.async.bulk moves it.  Slabs flow through an NSLOT-deep ring; with two tiles in
//               flight a slab is consumed by tile A then tile B before its slot is released.
//   warps     : 0 = bulk-TMA producer, 1 = TMEM allocator + MMA issuer (leader CTA) / relay (peer),
//               2.. = epilogue warpgroups (one per tile in flight; thread == sample row).
//   epilogue  : tcgen05.ld 32 columns -> +bias -> ReLU -> cvt to f16x2 -> swizzled st.shared into the
//               tile's activation buffer, which is the next step's A operand (in place).
//   on-chip   : act[NT][4 kblocks][128][128B] + one PE block shared by the tiles (the encodings live
//               in the epilogue threads' registers and are written to it just before steps 0/5/9).
#include "nm_internal.cuh"
#include "nm_pe.cuh"
#include "tc_common.cuh"
#include <string.h>
// ---------------------------------------------------------------------------------------------
// Plan: which slabs a step consumes, where they live in the packed image.
// ---------------------------------------------------------------------------------------------
struct TcPlan {
  uint32_t slab_off[TC_STEPS][5];   // byte offset inside one CTA-rank image
  uint32_t slab_bytes[TC_STEPS];    // bytes per slab of this step (per CTA)
  uint32_t image_bytes;             // size of one CTA-rank image
};

// Every layer's bias rides in the MMAs.  The last channel of each encoding is the constant 1 (channel 63 of the position
// encoding, 27 of the direction encoding) and the weight column that multiplies it holds the bias (fp16, like every other
// weight): steps 0, 5 and 9 read the PE block anyway, so there the bias costs nothing.  The K = 256 steps (1-4, 6-8) get
// one extra k-block, a "bias slab" that is zero except for column 63, consumed by ONE K = 16 MMA against the last K slice
// (channels 48..63) of whatever encoding the PE block holds at that moment: channels 48..62 meet zero weights, channel 63
// is 1 in every position encoding, and the direction-encoding stores never touch that half of the block.  The epilogue
// then has no bias loads and no adds at all -- they were its largest cost (4 LDS.128 broadcasts + 16 FADD per 16 columns
// on the pipe that also feeds the UMMA operands and takes the activation stores: profiles/r02_mlp_tc_experiments.md); the
// extra MMA costs 1/16 of a layer's tensor time.  (A variant with a 256-byte constant-ones A operand -- K-major without
// swizzle, zero stride between row groups -- and a hi + lo bias pair in a 4 KB slab computed the right values but ran 10 %
// slower and dead-locked in multi-round launches; with five slabs per step and a five-slot ring shared by both tiles there
// is no room for a sixth slab at steps 5 and 9 either.  Not kept.)
__host__ __device__ constexpr int step_nkb(int s) { return s == 0 ? 1 : (s == 10 ? 2 : 5); }
__host__ __device__ constexpr bool kb_is_bias(int s, int kb) { return kb == 4 && s != 5 && s != 9 && s >= 1 && s <= 8; }
__host__ __device__ constexpr int step_N(int s) { return s <= 8 ? 256 : (s == 9 ? 128 : 16); }
// k-block kb of step s reads the PE buffer (else activation block `act_kb`)
__host__ __device__ constexpr bool kb_is_pe(int s, int kb) { return (s == 0) || (s == 5 && kb == 0) || (s == 9 && kb == 4); }
__host__ __device__ constexpr int kb_act_index(int s, int kb) { return s == 5 ? kb - 1 : kb; }
__host__ __device__ constexpr int kb_ksteps(int s, int kb) { return (s == 9 && kb == 4) ? 2 : 4; }   // dir PE: K=32

template <int kPair>
struct TcCfg {
  static constexpr int NT = kPair == 2 ? 2 : 1;
  static constexpr int NSLOT = kPair == 2 ? 5 : 4;
  static constexpr int SLOT_BYTES = 32768 / kPair;
  static constexpr int THREADS = 64 + 256 + 128;  // producer + MMA warps, 8 epilogue warps, 4 encoding warps
  static constexpr int OFF_ACT = 0;
  static constexpr int OFF_PE = OFF_ACT + NT * 4 * TC_KB_BYTES;
  static constexpr int OFF_RING = OFF_PE + TC_KB_BYTES;
  static constexpr int OFF_BAR = OFF_RING + NSLOT * SLOT_BYTES;
  // barriers: full[NSLOT] peer_full[NSLOT] empty[NSLOT] tmem_full[NT] act_ready[NT] pe_free pe_ready
  static constexpr int N_BAR = 3 * NSLOT + 2 * NT + 2;
  static constexpr int OFF_TMEMPTR = OFF_BAR + 8 * N_BAR;
  // alpha-head partials of the upper column half, one float per row and tile in flight
  static constexpr int OFF_ALPHA = (OFF_TMEMPTR + 16 + 127) & ~127;
  static constexpr int SMEM_USED = OFF_ALPHA + NT * 512;
  static constexpr int SMEM_SLACK = (232448 - SMEM_USED) < 1024 ? (232448 - SMEM_USED) : 1024;   // alignment slack that still fits 227 KB
  static constexpr int SMEM_BYTES = SMEM_USED + SMEM_SLACK;
};

// What the epilogue threads still multiply or add with, the same for every row -- the alpha_linear weights (the alpha
// head runs in fp32 on the unrounded post-ReLU layer-7 accumulators, models/vanilla.py:135) and the four output biases --
// is read from a CONSTANT BANK through the uniform datapath (LDCU.128 into uniform registers, then `FFMA R, R, UR, R`): no
// shared-memory loads (the LSU / shared-memory pipe is the busiest unit of this kernel: UMMA operand fetches + activation
// stores) and no vector registers.  Inference launches carry the table as kernel parameters (bank 0); the training forward, whose weights are
// re-packed on the device every optimiser step, reads a __constant__ copy refreshed by a stream-ordered
// device-to-device copy in front of the launch (filling kernel parameters would need a host round trip per step).
//   [0, 256) alpha_linear.weight | 256..258 rgb bias | 259 alpha bias
#define TC_CONST_ALPHA 0
#define TC_CONST_OUT 256
__constant__ __align__(16) float c_tc_consts[TC_CONST_FLOATS];

// with two tiles sharing every slab, a step's slabs must all fit the ring at once (see step_nkb)
__host__ __device__ constexpr bool ring_holds_a_step(int nslot) {
  for (int s = 0; s < TC_STEPS; ++s)
    if (step_nkb(s) > nslot) return false;
  return true;
}
static_assert(ring_holds_a_step(TcCfg<2>::NSLOT), "a step has more slabs than the ring has slots: tiles A and B would deadlock");

struct TcParams {
  const uint8_t* wimg;      // packed slabs, kPair images back to back
  TcPlan plan;
  NmMlpInput in;
  NmPeSpec pos_pe, dir_pe;
  float* raw;
  long long n_tiles;        // number of (pair-)tiles
  int32_t* range_flag;      // device word: bit 0 is set when an activation reached the fp16 range limit (saturated)
  int range_phase;          // sampled range check (kRange == 1): the rounds r with r % 64 == range_phase % 64 are checked
  // training forward (kTrain): fp16 activation stash for the backward pass, planes of n rows each
  __half* st_x;             // [8][n][256] post-ReLU outputs of layers 0..7
  __half* st_f;             // [n][256]    feature_linear output
  __half* st_v;             // [n][128]    views layer post-ReLU
  uint32_t* st_m;           // [9][n][8]   sign words, 16 bits per 16 columns: bit j = [col 2j > 0], bit 8+j = [col 2j+1 > 0];
                            //             planes 0..7 = pts_linears, plane 8 = views layer (words 0..3 used)
  CUtensorMap map_x, map_f, map_v;   // TMA store maps of st_x / st_f / st_v (kTrain only)
  int dbg;                  // debug (NEUMAN_TC_DEBUG): bit 0 = training kernel skips its TMA stash stores (timing experiments only)
  long long* trace;         // optional debug timeline (tools/tc_trace.py): [cta<2][role<2][event<4][256] clock64 stamps
  __align__(16) float consts[TC_CONST_FLOATS];   // inference: the constant table as kernel parameters
};

template <bool kParam>
__device__ __forceinline__ float4 cst4(const TcParams& P, int i) {      // i % 4 == 0
  return kParam ? *reinterpret_cast<const float4*>(&P.consts[i]) : *reinterpret_cast<const float4*>(&c_tc_consts[i]);
}
template <bool kParam>
__device__ __forceinline__ float cst(const TcParams& P, int i) { return kParam ? P.consts[i] : c_tc_consts[i]; }

// sin/cos of 2*pi*f (f in cycles).  The operands of the tensor-core path are fp16 (quantisation 2.4e-4 on a
// [-1,1] value), so the encodings only need ~1e-5: range reduction is done exactly in "cycles" with a
// two-float 1/(2*pi) scale (error-free products via fma), then MUFU sin/cos on [-pi, pi] (abs err ~4e-7).
__device__ __forceinline__ void sincos_cycles(float f, float& s, float& c) {
  f = f - rintf(f);
  const float a = f * 6.283185307179586f;
  s = __sinf(a);
  c = __cosf(a);
}

// Encodes x (3) into 64 f16 channels packed as 32 x f16x2; unused channels are zero.
// pe.table for this path: posenc [n_freqs][2] = (hi, lo) of f_k/(2*pi); rotate [3 n_freqs][6] = (hi[3], lo[3])
// of bvals[q]/(2*pi)  (api.cu: pe_cycles_table).
__device__ __forceinline__ void encode_f16(const NmPeSpec& pe, const float x[3], uint32_t (&out)[32], int nq) {
  float ch[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) ch[i] = 0.f;
  ch[0] = x[0]; ch[1] = x[1]; ch[2] = x[2];
  if (pe.kind == NM_PE_ROTATE) {
#pragma unroll
    for (int q = 0; q < 30; ++q) {
      if (q < nq) {
        const float* b = pe.table + 6 * q;
        float frac = 0.f, low = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float bh = __ldg(b + i), bl = __ldg(b + 3 + i);
          const float pr = x[i] * bh;
          low += fmaf(x[i], bl, fmaf(x[i], bh, -pr));        // exact product residual + low part
          frac += pr - rintf(pr);                            // exact
        }
        float s, c;
        sincos_cycles(frac + low, s, c);
        if (nq == 30) { ch[3 + q] = s; ch[33 + q] = c; }
        else { ch[3 + q] = s; ch[15 + q] = c; }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      if (3 * k < nq) {
        const float fh = __ldg(pe.table + 2 * k), fl = __ldg(pe.table + 2 * k + 1);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float pr = x[d] * fh;
          const float low = fmaf(x[d], fl, fmaf(x[d], fh, -pr));
          float s, c;
          sincos_cycles((pr - rintf(pr)) + low, s, c);
          ch[3 + 6 * k + d] = s; ch[6 + 6 * k + d] = c;
        }
      }
    }
  }
  ch[nq == 30 ? 63 : 27] = 1.f;          // the constant channel that multiplies the bias column of the weight slabs
#pragma unroll
  for (int i = 0; i < 32; ++i) out[i] = pack_f16x2(ch[2 * i], ch[2 * i + 1], false);
}

// write 64 f16 (one 128-byte row of a K block) with the 128B swizzle
__device__ __forceinline__ void store_row_swizzled(uint8_t* blk, int row, const uint32_t* v, int nchunks) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j < nchunks) {
      uint4 q = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      *reinterpret_cast<uint4*>(blk + row * 128 + ((j ^ (row & 7)) << 4)) = q;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Epilogue of 16 accumulator columns [c0, c0+16) of one row: + bias, the alpha head on step 7 (fp32 FFMAs on the
// ReLU of the unrounded accumulators), ReLU, saturating f16x2 pack, two swizzled 16-byte stores into the activation
// block that is the next step's A operand.  c0 is a compile-time constant and `crow` (= step * 256) warp-uniform, so
// the bias and the alpha weights are uniform-register operands.
// ---------------------------------------------------------------------------------------------
template <bool RELU>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t d;
  if (RELU) asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// running maximum of |activation| as packed halves (range flag)
template <bool RELU>
__device__ __forceinline__ void track_range(uint32_t& rng, uint32_t packed) {
  __half2 h = *reinterpret_cast<const __half2*>(&packed);
  if (!RELU) h = __habs2(h);
  const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&rng), h);
  rng = *reinterpret_cast<const uint32_t*>(&m);
}

template <bool RELU, bool ALPHA, bool kParam, bool kRange>
__device__ __forceinline__ uint32_t epi_sub16(const TcParams& P, const uint32_t (&v)[16], int c0, float (&alpha)[4],
                                          uint8_t* act, int row, uint32_t& rng) {
  uint32_t packed[8];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float x0 = __uint_as_float(v[4 * g + 0]), x1 = __uint_as_float(v[4 * g + 1]);      // bias included by the MMAs
    const float x2 = __uint_as_float(v[4 * g + 2]), x3 = __uint_as_float(v[4 * g + 3]);
    if (ALPHA) {                                // alpha_linear on the fp32 ReLU output (:135)
      const float4 w = cst4<kParam>(P, TC_CONST_ALPHA + c0 + 4 * g);
      alpha[0] = fmaf(fmaxf(x0, 0.f), w.x, alpha[0]);
      alpha[1] = fmaf(fmaxf(x1, 0.f), w.y, alpha[1]);
      alpha[2] = fmaf(fmaxf(x2, 0.f), w.z, alpha[2]);
      alpha[3] = fmaf(fmaxf(x3, 0.f), w.w, alpha[3]);
    }
    packed[2 * g] = pack2<RELU>(x0, x1);
    packed[2 * g + 1] = pack2<RELU>(x2, x3);
  }
  if (kRange) {
#pragma unroll
    for (int j = 0; j < 8; ++j) track_range<RELU>(rng, packed[j]);
  }
  uint8_t* blk = act + (c0 >> 6) * TC_KB_BYTES + row * 128;
  const int ch0 = (c0 & 63) >> 3;
  *reinterpret_cast<uint4*>(blk + ((ch0 ^ (row & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  *reinterpret_cast<uint4*>(blk + (((ch0 + 1) ^ (row & 7)) << 4)) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
  // ReLU sign word of these 16 outputs (training kernel only): bit j = [column c0+2j > 0], bit 8+j = [column
  // c0+2j+1 > 0].  The INT32 pipe runs at half the FP rate and this epilogue is short of issue slots in the training
  // kernel, so the compare is a packed-half HSET2 (one per register, FP16 pipe) that yields 0xFFFF per positive half,
  // and only one LOP3 per register (pick bit j of each half, OR into the word) plus one PRMT remain on the INT pipe.
  uint32_t acc = 0;
  const __half2 zero2 = __float2half2_rn(0.f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const __half2 h = *reinterpret_cast<const __half2*>(&packed[j]);
    acc |= __hgt2_mask(h, zero2) & ((1u << j) | (1u << (16 + j)));
  }
  return __byte_perm(acc, 0u, 0x4420);          // byte 0 = low halves, byte 1 = high halves
}

// Drains NC accumulator columns [CB, CB + NC) of this thread's TMEM lane into the activation block, fully unrolled
// (every column offset is a compile-time constant) and software pipelined over 16-column sub-chunks: the tcgen05.ld
// of sub-chunk i+1 is in flight while sub-chunk i is converted and stored.
template <bool RELU, bool ALPHA, int CB, int NC, bool kParam, bool kRange>
__device__ __forceinline__ void epi_step(const TcParams& P, uint32_t t_lane, float (&alpha)[4], uint8_t* act, int row,
                                         uint4& signs, uint32_t& rng) {
  uint32_t v0[16], v1[16];
  tmem_ld16(t_lane + CB, v0);
#pragma unroll
  for (int q = 0; q < NC / 32; ++q) {
    const int c = CB + 32 * q;
    tmem_wait_ld();
    tmem_ld16(t_lane + c + 16, v1);
    const uint32_t m0 = epi_sub16<RELU, ALPHA, kParam, kRange>(P, v0, c, alpha, act, row, rng);
    tmem_wait_ld();
    if (q + 1 < NC / 32) tmem_ld16(t_lane + c + 32, v0);
    const uint32_t m1 = epi_sub16<RELU, ALPHA, kParam, kRange>(P, v1, c + 16, alpha, act, row, rng);
    const uint32_t w = m0 | (m1 << 16);
    if (q == 0) signs.x = w; else if (q == 1) signs.y = w; else if (q == 2) signs.z = w; else signs.w = w;
  }
}

// ---------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------
// debug timeline: role 0 = MMA issuer (events 0: operand ready seen, 1: step issued+committed),
// role 1 = epilogue warp 2 lane 0 (events 0: accumulator ready seen, 1: drained, 2: published); tile 0 only
#define TC_TRACE(role, ev, idx)                                                                             \
  do {                                                                                                      \
    if (P.trace && blockIdx.x < 2 && (idx) < 256) P.trace[((blockIdx.x * 2 + (role)) * 4 + (ev)) * 256 + (idx)] = clock64(); \
  } while (0)

template <int kPair, bool kTrain, int kRange>
__global__ void __launch_bounds__(TcCfg<kPair>::THREADS, 1) k_mlp_tc(const __grid_constant__ TcParams P) {
  using C = TcCfg<kPair>;
  constexpr int NT = C::NT, NSLOT = C::NSLOT;
  constexpr bool kParam = !kTrain;               // where the constant table comes from (see TcParams::consts)
  extern __shared__ uint8_t smem_dyn[];
  // SWIZZLE_128B atoms need a 1024-byte aligned base (identical in both CTAs of a pair)
  const uint32_t raw_addr = smem_u32(smem_dyn);
  const uint32_t pad = (1024 - (raw_addr & 1023)) & 1023;
  if (pad > C::SMEM_SLACK) __trap();
  uint8_t* smem = smem_dyn + pad;
  const uint32_t sbase = smem_u32(smem);

  // warp index broadcast from lane 0: tells the compiler it is warp-uniform, so role branches are uniform branches and
  // step-dependent constant-bank addresses can live in uniform registers (LDCU / FADD R, R, UR in the epilogue)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t rank = kPair == 2 ? cluster_ctarank() : 0;
  const long long pair_id = blockIdx.x / kPair;
  const long long n_pairs = gridDim.x / kPair;

  auto bar_full = [&](int i) { return sbase + C::OFF_BAR + 8 * i; };
  auto bar_peer = [&](int i) { return sbase + C::OFF_BAR + 8 * (NSLOT + i); };
  auto bar_empty = [&](int i) { return sbase + C::OFF_BAR + 8 * (2 * NSLOT + i); };
  auto bar_tfull = [&](int t) { return sbase + C::OFF_BAR + 8 * (3 * NSLOT + t); };
  auto bar_aready = [&](int t) { return sbase + C::OFF_BAR + 8 * (3 * NSLOT + NT + t); };
  const uint32_t bar_pefree = sbase + C::OFF_BAR + 8 * (3 * NSLOT + 2 * NT);
  const uint32_t bar_peready = sbase + C::OFF_BAR + 8 * (3 * NSLOT + 2 * NT + 1);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEMPTR);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(bar_full(i), 1); mbar_init(bar_peer(i), 1); mbar_init(bar_empty(i), 1); }
    for (int t = 0; t < NT; ++t) { mbar_init(bar_tfull(t), 1); mbar_init(bar_aready(t), 8 * kPair); }
    mbar_init(bar_pefree, 1);
    mbar_init(bar_peready, 4 * kPair);
    fence_mbar_init();
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
  // sample handled by row `r` of this CTA in tile t of a round
  auto sample_of = [&](long long round, int t, int r) { return (((round * n_pairs + pair_id) * NT + t) * kPair + rank) * 128 + r; };
  auto valid_of = [&](long long round, int t, int r) {
    return ((round * n_pairs + pair_id) * NT + t) < P.n_tiles && sample_of(round, t, r) < P.in.n;
  };
  // PE-buffer use index (order of the MMA stream: per round tile0/tile1 at step 0, then at step 5, then at step 9)
  auto pe_use = [&](long long round, int k, int t) { return (round * 3 + k) * NT + t; };

  if (warp == 0) {
    // =============================== bulk-TMA producer ===============================
    if (lane == 0) {
      const uint8_t* img = P.wimg + (size_t)rank * P.plan.image_bytes;
      uint32_t q = 0;
      for (long long round = 0; round < n_rounds; ++round) {
        for (int s = 0; s < TC_STEPS; ++s) {
          for (int kb = 0; kb < step_nkb(s); ++kb, ++q) {
            const uint32_t bytes = P.plan.slab_bytes[s];
            const uint32_t slot = q % NSLOT, gen = q / NSLOT;
            mbar_wait(bar_empty(slot), (gen & 1) ^ 1);
            mbar_arrive_expect_tx(bar_full(slot), bytes);
            bulk_g2s(sbase + C::OFF_RING + slot * C::SLOT_BYTES, img + P.plan.slab_off[s][kb], bytes, bar_full(slot));
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // =============================== MMA issuer (leader CTA) ===============================
      // The whole warp runs this loop with warp-uniform values (so descriptors stay in uniform registers and
      // ptxas needs no divergence "waterfall" around UTCHMMA); one elected lane issues the tcgen05 instructions.
      uint32_t issuer = 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\t.reg .b32 r;\n\telect.sync r|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(issuer));
      uint32_t q0 = 0, nstep = 0;
      for (long long round = 0; round < n_rounds; ++round) {
        for (int s = 0; s < TC_STEPS; ++s, ++nstep) {
          const uint32_t idesc = make_idesc(128 * kPair, step_N(s));
          const int nkb = step_nkb(s);
          const int pe_k = s == 0 ? 0 : (s == 5 ? 1 : 2);
          for (int t = 0; t < NT; ++t) {
            mbar_wait(bar_aready(t), nstep & 1);          // A operand written, accumulator drained
            tc_fence_after();
            if (t == 0 && issuer) TC_TRACE(0, 0, nstep);
            const uint32_t d_tmem = tmem_base + t * 256 + (s == 10 ? 128 : 0);
            for (int kb = 0; kb < nkb; ++kb) {
              const uint32_t q = q0 + kb, slot = q % NSLOT, gen = q / NSLOT;
              if (t == 0) {
                mbar_wait(bar_full(slot), gen & 1);
                if (kPair == 2) mbar_wait(bar_peer(slot), gen & 1);
                tc_fence_after();
              }
              const bool is_pe = kb_is_pe(s, kb), is_bias = kb_is_bias(s, kb);
              if (is_pe) {                                // encodings written by the encoding warps of both CTAs
                mbar_wait(bar_peready, (uint32_t)(pe_use(round, pe_k, t) & 1));
                tc_fence_after();
              }
              // (descriptors are computed by the whole warp, outside the elected-lane region: they stay in uniform registers)
              const uint32_t a_addr = (is_pe || is_bias) ? sbase + C::OFF_PE
                                                         : sbase + C::OFF_ACT + (t * 4 + kb_act_index(s, kb)) * TC_KB_BYTES;
              // K advances by 32 B (= 2 in descriptor address units) inside the 128-byte swizzle atom; a bias slab is one
              // K = 16 MMA on the last K slice (channels 48..63 of the PE block x columns 48..63 of the slab)
              const uint64_t a_desc = make_desc(a_addr) + (is_bias ? 6 : 0);
              const uint64_t b_desc = make_desc(sbase + C::OFF_RING + slot * C::SLOT_BYTES) + (is_bias ? 6 : 0);
              if (issuer) {
                umma_f16<kPair>(d_tmem, a_desc, b_desc, idesc, kb != 0);
                if (!is_bias) {
                  umma_f16<kPair>(d_tmem, a_desc + 2, b_desc + 2, idesc, 1);
                  if (kb_ksteps(s, kb) == 4) {
                    umma_f16<kPair>(d_tmem, a_desc + 4, b_desc + 4, idesc, 1);
                    umma_f16<kPair>(d_tmem, a_desc + 6, b_desc + 6, idesc, 1);
                  }
                }
                if (t == NT - 1) umma_commit<kPair>(bar_empty(slot));      // slab consumed by every tile
                if (is_pe) umma_commit<kPair>(bar_pefree);                 // PE block may be rewritten
              }
              __syncwarp();
            }
            if (issuer) {
              umma_commit<kPair>(bar_tfull(t));                            // accumulator complete
              if (t == 0) TC_TRACE(0, 1, nstep);
            }
            __syncwarp();
          }
          q0 += nkb;
        }
      }
    } else {
      // =============================== relay (peer CTA of a pair) ===============================
      // tells the leader's MMA thread that this CTA's half of a slab has landed
      if (lane == 0) {
        uint32_t q = 0;
        for (long long round = 0; round < n_rounds; ++round)
          for (int s = 0; s < TC_STEPS; ++s)
            for (int kb = 0; kb < step_nkb(s); ++kb, ++q) {
              const uint32_t slot = q % NSLOT, gen = q / NSLOT;
              mbar_wait(bar_full(slot), gen & 1);
              mbar_arrive_cluster(bar_peer(slot), 0);
            }
      }
    }
  } else if (warp >= 10) {
    // ========================= encoding warps: one thread per sample row =========================
    // Embedder.forward (models/vanilla.py:82-92) of the samples, well ahead of the MMAs that read it: the fp16
    // encodings of both tiles of a round live in these threads' registers and are stored into the (single) PE block
    // in the order the MMA stream uses it -- tile 0 / tile 1 at step 0, again at step 5 (skip connection, :131), the
    // direction encoding at step 9 (:137) -- each store waiting for the MMAs of the previous use to retire.
    const int prow = (warp - 10) * 32 + lane;
    uint8_t* pebuf = smem + C::OFF_PE;
    uint32_t pos[NT][32], dir[NT][16];
    uint32_t rng = 0;
    auto encode_pos = [&](long long round) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        float p[3] = {0.f, 0.f, 0.f}, v[3] = {0.f, 0.f, 0.f};
        if (valid_of(round, t, prow)) nm_fetch_sample(P.in, sample_of(round, t, prow), p, v);
        encode_f16(P.pos_pe, p, pos[t], 30);
        if (kRange) track_range<false>(rng, pos[t][0]), track_range<false>(rng, pos[t][1]);   // raw x, y, z (+ one sine)
      }
    };
    auto encode_dir = [&](long long round) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        float p[3] = {0.f, 0.f, 0.f}, v[3] = {0.f, 0.f, 0.f};
        if (valid_of(round, t, prow)) nm_fetch_sample(P.in, sample_of(round, t, prow), p, v);
        uint32_t tmp[32];
        encode_f16(P.dir_pe, v, tmp, 12);
#pragma unroll
        for (int j = 0; j < 16; ++j) dir[t][j] = tmp[j];
        if (kRange) track_range<false>(rng, tmp[0]), track_range<false>(rng, tmp[1]);
      }
    };
    auto put = [&](long long u, const uint32_t* v, int nchunks) {
      if (u > 0) mbar_wait_backoff(bar_pefree, (uint32_t)((u - 1) & 1), 32);
      store_row_swizzled(pebuf, prow, v, nchunks);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar_peready, 0);
    };
    if (n_rounds > 0) encode_pos(0);
    for (long long round = 0; round < n_rounds; ++round) {
#pragma unroll
      for (int t = 0; t < NT; ++t) put(pe_use(round, 0, t), pos[t], 8);
      encode_dir(round);
#pragma unroll
      for (int t = 0; t < NT; ++t) put(pe_use(round, 1, t), pos[t], 8);
#pragma unroll
      for (int t = 0; t < NT; ++t) put(pe_use(round, 2, t), dir[t], 4);
      if (round + 1 < n_rounds) encode_pos(round + 1);
    }
    if (kRange && (((rng & 0xFFFFu) >= 0x7BFFu) || ((rng >> 16) >= 0x7BFFu))) atomicOr(P.range_flag, 1);
  } else {
    // ========================= epilogue: 8 warps serve the tiles in flight in turn =========================
    // Both warpgroups drain every tile: warp (2+q) and warp (6+q) share TMEM lane quadrant q and split the
    // accumulator columns in halves (g = 0 / 1), so a step's epilogue takes half as long.
    const int ew = warp - 2;                       // 0..7
    const int g = ew >> 2;                         // column half
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;              // sample row inside the CTA tile
    const int etid = ew * 32 + lane;               // 0..255
    float* s_alpha = reinterpret_cast<float*>(smem + C::OFF_ALPHA);        // [NT][128] alpha partial of the g==1 half
    uint32_t nstep = 0;
    uint32_t rng = 0;

    auto publish = [&](int t) {                     // tile t: A operand ready + accumulator drained
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar_aready(t), 0);
    };

    for (long long round = 0; round < n_rounds; ++round) {
      // ---- step 0 reads only the PE block: nothing to drain, the tiles' accumulators are free ----
#pragma unroll
      for (int t = 0; t < NT; ++t) publish(t);
      float alpha[NT][4];
#pragma unroll
      for (int t = 0; t < NT; ++t) alpha[t][0] = alpha[t][1] = alpha[t][2] = alpha[t][3] = 0.f;
      const long long rperiod = n_rounds < 64 ? n_rounds : 64;
      const bool track = kRange == 2 || (kRange == 1 && (round % rperiod) == (P.range_phase % rperiod));
      for (int s = 0; s < TC_STEPS; ++s, ++nstep) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          uint8_t* act = smem + C::OFF_ACT + t * 4 * TC_KB_BYTES;
          const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + t * 256;
          mbar_wait(bar_tfull(t), nstep & 1);
          tc_fence_after();
          if (t == 0 && etid == 0) TC_TRACE(1, 0, nstep);
          if (s < 10) {
            // training: this warp's slice of the tile's activation buffer is the source of the TMA store issued one
            // step ago: it must have been read before the slice is overwritten (the other tile's store may still fly)
            if (kTrain) { if (lane == 0) tma_store_wait_read<1>(); __syncwarp(); }
            uint4 signs = make_uint4(0, 0, 0, 0);
            // range flag (kRange 0: off, 2: every sample, 1: every 64th round of the launch, the phase rotating from
            // launch to launch -- the packed-half max costs ~4 % of the kernel when it runs on every sample)
#define NM_EPI(RELU, ALPHA, CB, NC)                                                                         \
  do {                                                                                                       \
    if (track) epi_step<RELU, ALPHA, CB, NC, kParam, true>(P, t_lane, alpha[t], act, row, signs, rng);         \
    else epi_step<RELU, ALPHA, CB, NC, kParam, false>(P, t_lane, alpha[t], act, row, signs, rng);              \
  } while (0)
            if (g == 0) {
              if (s == 7) NM_EPI(true, true, 0, 128);
              else if (s == 8) NM_EPI(false, false, 0, 128);
              else if (s == 9) NM_EPI(true, false, 0, 64);
              else NM_EPI(true, false, 0, 128);
            } else {
              if (s == 7) NM_EPI(true, true, 128, 128);
              else if (s == 8) NM_EPI(false, false, 128, 128);
              else if (s == 9) NM_EPI(true, false, 64, 64);
              else NM_EPI(true, false, 128, 128);
            }
#undef NM_EPI
            if (kTrain && s < 8 && valid_of(round, t, row))
              reinterpret_cast<uint4*>(P.st_m + ((size_t)s * P.in.n + sample_of(round, t, row)) * 8)[g] = signs;
            if (kTrain && s == 9 && valid_of(round, t, row))   // views layer: 64 columns per thread -> words 2g, 2g+1 of plane 8
              reinterpret_cast<uint2*>(P.st_m + ((size_t)8 * P.in.n + sample_of(round, t, row)) * 8)[g] = make_uint2(signs.x, signs.y);
            if (s == 7 && g == 1) s_alpha[t * 128 + row] = (alpha[t][0] + alpha[t][1]) + (alpha[t][2] + alpha[t][3]);
            if (t == 0 && etid == 0) TC_TRACE(1, 1, nstep);
            if (kTrain) {
              // activation stash: this warp's 32 rows x (128 | 64) columns leave by TMA straight from the swizzled
              // A buffer.  Steps 0..7: the MMA thread is told first (the store and the next step's MMAs only read the
              // slice).  Steps 8 and 9 hand the slice to another warp (the column split changes from 128 to 64 per
              // warpgroup and back), so there the store must have finished reading before anyone goes on.
              const long long i0 = sample_of(round, t, row) - lane;            // first row of this warp
              const bool issue = lane == 0 && i0 < P.in.n && !(P.dbg & 1);
              const uint32_t src = sbase + C::OFF_ACT + t * 4 * TC_KB_BYTES + quad * 32 * 128;
              if (s < 8) {
                publish(t);
                if (issue) {
                  tma_store_3d(&P.map_x, src + (2 * g) * TC_KB_BYTES, 128 * g, (int)i0, s);
                  tma_store_3d(&P.map_x, src + (2 * g + 1) * TC_KB_BYTES, 128 * g + 64, (int)i0, s);
                  tma_store_commit();
                }
              } else {
                fence_async_smem();
                __syncwarp();
                if (issue) {
                  if (s == 8) {
                    tma_store_3d(&P.map_f, src + (2 * g) * TC_KB_BYTES, 128 * g, (int)i0, 0);
                    tma_store_3d(&P.map_f, src + (2 * g + 1) * TC_KB_BYTES, 128 * g + 64, (int)i0, 0);
                  } else {
                    tma_store_3d(&P.map_v, src + g * TC_KB_BYTES, 64 * g, (int)i0, 0);
                  }
                  tma_store_commit();
                  tma_store_wait_read<0>();
                }
                publish(t);
              }
            } else {
              publish(t);
            }
            if (t == 0 && etid == 0) TC_TRACE(1, 2, nstep);
          } else {
            if (g == 0) {
              uint32_t v[4];
              tmem_ld4(t_lane + 128, v);
              tmem_wait_ld();
              if (valid_of(round, t, row)) {
                const float a = (alpha[t][0] + alpha[t][1]) + (alpha[t][2] + alpha[t][3]) + s_alpha[t * 128 + row];
                const float4 ob = cst4<kParam>(P, TC_CONST_OUT);
                float4 o = make_float4(__uint_as_float(v[0]) + ob.x, __uint_as_float(v[1]) + ob.y,
                                       __uint_as_float(v[2]) + ob.z, a + ob.w);
                reinterpret_cast<float4*>(P.raw)[sample_of(round, t, row)] = o;   // [r,g,b,sigma] (:144)
              }
            }
            tc_fence_before();
          }
        }
      }
    }
    if (kRange && (((rng & 0xFFFFu) >= 0x7BFFu) || ((rng >> 16) >= 0x7BFFu))) atomicOr(P.range_flag, 1);
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
// Packing: fp32 nn.Linear weights -> fp16 slabs in the swizzled UMMA layout.
// ---------------------------------------------------------------------------------------------
struct PackSrc {
  const float* w[8]; const float* feat; const float* views; const float* rgb;
  const float* b[8]; const float* feat_b; const float* views_b;      // biases: they ride in the slabs (column of the constant-1 channel)
};

// weight of (step s, output n, k-block kb, kk in [0,64)) or 0 for padding
__device__ __forceinline__ float src_weight(const PackSrc& S, int s, int n, int kb, int kk) {
  if (kb_is_bias(s, kb)) {                       // bias slab of a K = 256 step: only the column of PE channel 63 is non-zero
    if (kk != 63) return 0.f;
    return s == 8 ? S.feat_b[n] : S.b[s][n];
  }
  if (s == 0) return kk < NM_POS_PE ? S.w[0][(size_t)n * NM_POS_PE + kk] : S.b[0][n];           // kk == 63: bias
  if (s >= 1 && s <= 7 && s != 5) return S.w[s][(size_t)n * 256 + kb * 64 + kk];
  if (s == 5) {
    const int ld = NM_POS_PE + 256;
    if (kb == 0) return kk < NM_POS_PE ? S.w[5][(size_t)n * ld + kk] : S.b[5][n];
    return S.w[5][(size_t)n * ld + NM_POS_PE + (kb - 1) * 64 + kk];
  }
  if (s == 8) return S.feat[(size_t)n * 256 + kb * 64 + kk];
  if (s == 9) {
    const int ld = 256 + NM_DIR_PE;
    if (kb < 4) return S.views[(size_t)n * ld + kb * 64 + kk];
    return kk < NM_DIR_PE ? S.views[(size_t)n * ld + 256 + kk] : (kk == NM_DIR_PE ? S.views_b[n] : 0.f);    // kk == 27: bias
  }
  // s == 10: rgb, N padded 3 -> 16 (its bias is added with the alpha bias when the output is written)
  return n < 3 ? S.rgb[(size_t)n * 128 + kb * 64 + kk] : 0.f;
}

__global__ void k_tc_pack(PackSrc S, TcPlan plan, int kpair, __half* __restrict__ out) {
  // one thread per packed element of one CTA-rank image; grid.y = rank
  const int rank = blockIdx.y;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // half index inside the image
  if (e * 2 >= plan.image_bytes) return;
  const uint32_t byte = (uint32_t)(e * 2);
  // locate (s, kb)
  int s = 0, kb = 0;
  for (int ss = 0; ss < TC_STEPS; ++ss)
    for (int k = 0; k < step_nkb(ss); ++k)
      if (byte >= plan.slab_off[ss][k]) { s = ss; kb = k; }
  const uint32_t in_slab = byte - plan.slab_off[s][kb];
  const int n_cta = step_N(s) / kpair;
  const int n_local = in_slab >> 7;
  const int chunk_phys = (in_slab & 127) >> 4;
  const int chunk = chunk_phys ^ (n_local & 7);                         // undo the 128B swizzle
  const int kk = chunk * 8 + ((in_slab & 15) >> 1);
  const int n = rank * n_cta + n_local;
  out[(size_t)rank * (plan.image_bytes / 2) + e] = __float2half_rn(src_weight(S, s, n, kb, kk));
}

// the constant table of the epilogue (layout: TcParams::consts)
__global__ void k_tc_consts(const float* rgb_b, const float* alpha_w, const float* alpha_b, float* __restrict__ out) {
  const int i = threadIdx.x;      // 256 threads
  out[TC_CONST_ALPHA + i] = alpha_w[i];
  if (i < TC_CONST_FLOATS - TC_CONST_OUT) out[TC_CONST_OUT + i] = i < 3 ? rgb_b[i] : (i == 3 ? alpha_b[0] : 0.f);
}

// ---------------------------------------------------------------------------------------------
// The encodings as the tensor-core kernel feeds them to the MMAs (same code, same fp16 values), written out as
// planes for the weight-gradient GEMMs of the training step: [n][64] (position, channel 63 = 1.0) or [n][32]
// (direction, channel 27 = 1.0).  The constant channel multiplies a zero weight in the forward and returns the bias
// gradient as an extra column of  g^T @ plane.  One thread per sample; 10-20 us per step, cheaper than stashing the
// encodings from inside the MLP kernel (per-thread 128-byte rows from the epilogue warps: ~10 % of that kernel).
// ---------------------------------------------------------------------------------------------
__global__ void k_encode_f16(NmPeSpec pe, int which, const float* __restrict__ x, long long group, long long n,
                             __half* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long xi = group > 0 ? i / group : i;
  const float xv[3] = {x[3 * xi], x[3 * xi + 1], x[3 * xi + 2]};
  uint32_t e[32];
  encode_f16(pe, xv, e, which == 0 ? 30 : 12);
  if (which == 0) {
    e[31] |= 0x3C000000u;                                                  // channel 63 := 1.0
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)i * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = make_uint4(e[4 * j], e[4 * j + 1], e[4 * j + 2], e[4 * j + 3]);
  } else {
    e[13] |= 0x3C000000u;                                                  // channel 27 := 1.0
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)i * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = make_uint4(e[4 * j], e[4 * j + 1], e[4 * j + 2], e[4 * j + 3]);
  }
}

int nm_tc_encode(nm_ctx* ctx, const NmNet& net, int which, const float* x, int64_t group, int64_t n, __half* out, cudaStream_t st) {
  NmPeSpec pe = which == 0 ? NmPeSpec{net.desc.pos_pe_kind, net.desc.pos_n_freqs, net.f32 + net.o_pos_cyc}
                           : NmPeSpec{net.desc.dir_pe_kind, net.desc.dir_n_freqs, net.f32 + net.o_dir_cyc};
  k_encode_f16<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pe, which, x, group, n, out);
  NM_CHECK_LAUNCH(ctx);
  return NM_OK;
}

static int tc_pair_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("NEUMAN_TC_PAIR");
    mode = (e && e[0] == '1') ? 1 : 2;
  }
  return mode;
}

static TcPlan make_plan(int kpair) {
  TcPlan p{};
  uint32_t off = 0;
  for (int s = 0; s < TC_STEPS; ++s) {
    p.slab_bytes[s] = (uint32_t)(step_N(s) / kpair) * 128u;
    for (int kb = 0; kb < step_nkb(s); ++kb) { p.slab_off[s][kb] = off; off += p.slab_bytes[s]; }
  }
  p.image_bytes = off;
  return p;
}

bool nm_tc_available() { return true; }

int nm_tc_pack(nm_ctx* ctx, NmNet& net, cudaStream_t st) {
  const int kpair = tc_pair_mode();
  TcPlan plan = make_plan(kpair);
  const size_t halfs = (size_t)kpair * plan.image_bytes / 2;
  if (!net.f16 || net.f16_halfs != halfs) {
    if (net.f16) { NM_CHECK_CUDA(ctx, cudaDeviceSynchronize()); NM_CHECK_CUDA(ctx, cudaFree(net.f16)); net.f16 = nullptr; }
    NM_CHECK_CUDA(ctx, cudaMalloc(&net.f16, halfs * sizeof(__half)));
    net.f16_halfs = halfs;
  }
  if (!net.tc_bias) NM_CHECK_CUDA(ctx, cudaMalloc(&net.tc_bias, TC_CONST_FLOATS * sizeof(float)));
  const nm_nerf_desc& d = net.desc;
  PackSrc S;
  for (int l = 0; l < 8; ++l) S.w[l] = d.pts_w[l];
  S.feat = d.feature_w; S.views = d.views_w; S.rgb = d.rgb_w;
  for (int l = 0; l < 8; ++l) S.b[l] = d.pts_b[l];
  S.feat_b = d.feature_b; S.views_b = d.views_b;
  dim3 grid((unsigned)((plan.image_bytes / 2 + 255) / 256), kpair);
  k_tc_pack<<<grid, 256, 0, st>>>(S, plan, kpair, net.f16);
  NM_CHECK_LAUNCH(ctx);
  k_tc_consts<<<1, 256, 0, st>>>(d.rgb_b, d.alpha_w, d.alpha_b, net.tc_bias);
  NM_CHECK_LAUNCH(ctx);
  net.consts_host_valid = false;        // the host copy (kernel parameters of inference launches) is refreshed lazily
  return NM_OK;
}

template <int kPair, bool kTrain, int kRange>
static int launch_tc(nm_ctx* ctx, const TcParams& P, cudaStream_t st) {
  using C = TcCfg<kPair>;
  NM_SET_SMEM_ONCE(ctx, (k_mlp_tc<kPair, kTrain, kRange>), C::SMEM_BYTES);
  int ctas = ctx->sm_count - (ctx->sm_count % kPair);
  long long need = P.n_tiles * kPair;                       // CTAs that have work in the first round
  need = (need + C::NT - 1) / C::NT;
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
  NM_CHECK_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_mlp_tc<kPair, kTrain, kRange>, P));
  NM_LAUNCHED(ctx);
  return NM_OK;
}

int nm_tc_forward(nm_ctx* ctx, NmNet& net, const float* pts, const float* views, const float* origins,
                  const float* dirs, const float* z, int64_t n, int32_t group, float* raw, cudaStream_t st,
                  const NmTrainStash* stash) {
  const int kpair = tc_pair_mode();
  if (!net.f16 || !net.tc_bias) NM_FAIL(ctx, NM_ERR_STATE, "nm_tc_forward: weights not packed");
  TcParams P;
  P.wimg = reinterpret_cast<const uint8_t*>(net.f16);
  P.plan = make_plan(kpair);
  P.in = NmMlpInput{pts, views, origins, dirs, z, (long long)n, group};
  P.pos_pe = NmPeSpec{net.desc.pos_pe_kind, net.desc.pos_n_freqs, net.f32 + net.o_pos_cyc};
  P.dir_pe = NmPeSpec{net.desc.dir_pe_kind, net.desc.dir_n_freqs, net.f32 + net.o_dir_cyc};
  P.raw = raw;
  P.n_tiles = (n + 128 * kpair - 1) / (128 * kpair);
  P.trace = nullptr;
  P.dbg = 0;
  P.range_flag = ctx->d_counter + NM_RANGE_FLAG_WORD;
  P.range_phase = (int)(ctx->range_seq++ % 64);
  if (const char* e = getenv("NEUMAN_TC_DEBUG")) P.dbg = atoi(e);
  P.st_x = stash ? stash->x : nullptr; P.st_f = stash ? stash->f : nullptr; P.st_v = stash ? stash->v : nullptr;
  P.st_m = stash ? stash->m : nullptr;
  memset(&P.map_x, 0, 3 * sizeof(CUtensorMap));
  if (stash) {
    if (n >= (int64_t)0x7fff0000) NM_FAIL(ctx, NM_ERR_INVALID, "nm_mlp_forward_train: n too large for one call");
    if (tc_make_store_map(&P.map_x, stash->x, 8, (uint64_t)n, 256) || tc_make_store_map(&P.map_f, stash->f, 1, (uint64_t)n, 256) ||
        tc_make_store_map(&P.map_v, stash->v, 1, (uint64_t)n, 128))
      NM_FAIL(ctx, NM_ERR_CUDA, "nm_mlp_forward_train: cuTensorMapEncodeTiled failed");
    // training forward: the constant table stays on the device (stream-ordered copy into the __constant__ bank)
    NM_CHECK_CUDA(ctx, cudaMemcpyToSymbolAsync(c_tc_consts, net.tc_bias, TC_CONST_FLOATS * sizeof(float), 0,
                                               cudaMemcpyDeviceToDevice, st));
  } else {
    if (!net.consts_host_valid) {       // once per (re)pack: the constant table becomes kernel parameters
      if (!net.consts_host) NM_CHECK_CUDA(ctx, cudaMallocHost(&net.consts_host, TC_CONST_FLOATS * sizeof(float)));
      NM_CHECK_CUDA(ctx, cudaMemcpyAsync(net.consts_host, net.tc_bias, TC_CONST_FLOATS * sizeof(float), cudaMemcpyDeviceToHost, st));
      NM_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
      net.consts_host_valid = true;
    }
    memcpy(P.consts, net.consts_host, sizeof(P.consts));
  }
  if (const char* e = getenv("NEUMAN_TC_TRACE")) P.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  // NEUMAN_TC_RANGE: 0 = no range flag, 1 (default) = sampled (every 64th round, rotating phase), 2 = every sample
  static const int range_mode = [] { const char* e = getenv("NEUMAN_TC_RANGE"); return e ? atoi(e) : 1; }();
#define NM_LAUNCH(TRAIN)                                                                                                  \
  do {                                                                                                                     \
    if (range_mode <= 0) return kpair == 2 ? launch_tc<2, TRAIN, 0>(ctx, P, st) : launch_tc<1, TRAIN, 0>(ctx, P, st);       \
    if (range_mode == 1) return kpair == 2 ? launch_tc<2, TRAIN, 1>(ctx, P, st) : launch_tc<1, TRAIN, 1>(ctx, P, st);       \
    return kpair == 2 ? launch_tc<2, TRAIN, 2>(ctx, P, st) : launch_tc<1, TRAIN, 2>(ctx, P, st);                            \
  } while (0)
  if (P.st_x) NM_LAUNCH(true);
  NM_LAUNCH(false);
#undef NM_LAUNCH
}
