// Frame drivers: the per-batch loop bodies of the reference renderers, run as chains of the stage
// kernels over large ray chunks that stay resident in HBM.
//
//   nm_render_vanilla    <- utils/render_utils.py:108-161 (render_vanilla)
//   nm_render_smpl_nerf  <- utils/render_utils.py:164-246 (render_smpl_nerf)
//   nm_render_hybrid     <- utils/render_utils.py:249-362 (render_hybrid_nerf) and
//                           :365-461 (render_hybrid_nerf_multi_persons)
//
// The reference's `rays_per_batch` only bounds its memory; results do not depend on it, so the
// drivers pick their own chunk (opt->rays_per_batch, default 32768 rays) sized for HBM.
#include "nm_internal.cuh"

namespace {

struct Arena {
  char* base = nullptr;
  size_t off = 0, cap = 0;
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

// general torch.linspace(start, end, steps)[i] in float32 (see nm_linspace01)
__device__ __forceinline__ float linspace_f(float start, float end, int i, int steps) {
  if (steps <= 1) return start;
  float step = (end - start) / (float)(steps - 1);
  if (i < steps / 2) return start + step * (float)i;
  return end - step * (float)(steps - 1 - i);
}

__global__ void k_fill_placeholder(float* __restrict__ z, float4* __restrict__ raw, long long R, int S, float start,
                                   float end) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * S) return;
  int s = (int)(idx % S);
  z[idx] = linspace_f(start, end, s, S);          // utils/render_utils.py:419
  raw[idx] = make_float4(0.f, 0.f, 0.f, 0.f);     // :418
}

// hit test near < far (utils/render_utils.py:206 / :313) + compaction of hit-ray indices
__global__ void k_compact_hits(const float* __restrict__ near_v, const float* __restrict__ far_v, long long R,
                               int32_t* __restrict__ hit_idx, int32_t* __restrict__ counter) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool hit = r < R && near_v[r] < far_v[r];
  unsigned m = __ballot_sync(0xffffffffu, hit);
  int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0 && m) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (hit) hit_idx[base + __popc(m & ((1u << lane) - 1))] = (int32_t)r;
}

__global__ void k_gather_rays(const int32_t* __restrict__ idx, int n, const float* __restrict__ o,
                              const float* __restrict__ d, const float* __restrict__ near_v,
                              const float* __restrict__ far_v, float* __restrict__ oh, float* __restrict__ dh,
                              float* __restrict__ nh, float* __restrict__ fh) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = idx[i];
#pragma unroll
  for (int c = 0; c < 3; ++c) { oh[3 * i + c] = o[3 * (size_t)r + c]; dh[3 * i + c] = d[3 * (size_t)r + c]; }
  nh[i] = near_v[r]; fh[i] = far_v[r];
}

// rows of `width` floats: dst[i] = src[idx[i]] (gather) or dst[idx[i]] = src[i] (scatter)
__global__ void k_move_rows(const int32_t* __restrict__ idx, long long n, int width, const float* __restrict__ src,
                            float* __restrict__ dst, int scatter) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * width) return;
  long long i = t / width;
  int c = (int)(t - i * width);
  long long r = idx[i];
  if (scatter) dst[r * width + c] = src[t];
  else dst[t] = src[r * width + c];
}

__global__ void k_mark_rows(const int32_t* __restrict__ idx, long long n, float* __restrict__ flag) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) flag[idx[t]] = 1.f;
}

// flag[r] = 1 for the rays of a near/far pair that hit (near < far)
__global__ void k_mark_hits(const float* __restrict__ near_v, const float* __restrict__ far_v, long long R,
                            float* __restrict__ flag) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R && near_v[r] < far_v[r]) flag[r] = 1.f;
}

__global__ void k_fill(float* __restrict__ p, long long n, float v) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) p[t] = v;
}

#define LAUNCH1D(kernel, n, st, ...)                                              \
  do {                                                                            \
    long long _n = (n);                                                           \
    if (_n > 0) {                                                                 \
      kernel<<<(unsigned)((_n + 255) / 256), 256, 0, st>>>(__VA_ARGS__);          \
      NM_CHECK_LAUNCH(ctx);                                                       \
    }                                                                             \
  } while (0)

#define TRY(call)                \
  do {                           \
    int _rc = (call);            \
    if (_rc != NM_OK) return _rc; \
  } while (0)

int check_common(nm_ctx* ctx, const nm_camera* cam, const nm_render_opts* opt, int64_t pix0, int64_t n, const int32_t* pixels,
                 const char* who) {
  if (!cam || !opt || n < 0 || (!pixels && (pix0 < 0 || pix0 + n > (int64_t)cam->H * cam->W)))
    NM_FAIL(ctx, NM_ERR_INVALID, std::string(who) + ": bad camera/options/pixel range");
  if (opt->samples_per_ray <= 0 || opt->importance_samples_per_ray < 0)
    NM_FAIL(ctx, NM_ERR_INVALID, std::string(who) + ": bad sample counts");
  return NM_OK;
}

int chunk_of(const nm_render_opts* opt) { return opt->rays_per_batch > 0 ? opt->rays_per_batch : 32768; }

int slot_ok(nm_ctx* ctx, int slot) { return slot >= 0 && slot < NM_MAX_NET_SLOTS && ctx->nets[slot].packed; }

// background coarse (+fine) pass over C rays already in o/d: leaves raw/z of the last pass in
// (*raw_out, *z_out) with *S_out samples.  utils/render_utils.py:141-153 / :287-298 / :398-409
int bkg_pass(nm_ctx* ctx, int coarse, int fine, const nm_render_opts* opt, const float* o, const float* d, int64_t C,
             float* z_c, float* raw_c, float* w_c, float* z_f, float* raw_f, float** raw_out, float** z_out,
             int* S_out, cudaStream_t st) {
  const int S = opt->samples_per_ray, N = opt->importance_samples_per_ray;
  TRY(nm_ray_to_samples(ctx, o, d, nullptr, nullptr, opt->near_bkg, opt->far_bkg, C, S, 0, nullptr, nullptr, nullptr,
                        z_c, st));
  TRY(nm_mlp_forward_rays(ctx, coarse, opt->mlp_mode, o, d, z_c, C, S, raw_c, st));
  ctx->last_mlp_evals += C * S;
  if (fine >= 0 && N > 0) {
    TRY(nm_raw2outputs(ctx, raw_c, z_c, d, C, S, nullptr, 1.f, opt->white_bkg, nullptr, nullptr, nullptr, w_c, nullptr, st));
    TRY(nm_importance_samples(ctx, o, d, z_c, w_c, C, S, N, 1, nullptr, nullptr, z_f, st));
    TRY(nm_mlp_forward_rays(ctx, fine, opt->mlp_mode, o, d, z_f, C, S + N, raw_f, st));
    ctx->last_mlp_evals += C * (S + N);
    *raw_out = raw_f; *z_out = z_f; *S_out = S + N;
  } else {
    *raw_out = raw_c; *z_out = z_c; *S_out = S;
  }
  return NM_OK;
}

int copy_out(nm_ctx* ctx, float* dst, const float* src, size_t n, int host_out, cudaStream_t st) {
  if (!dst || dst == src) return NM_OK;
  NM_CHECK_CUDA(ctx, cudaMemcpyAsync(dst, src, n * sizeof(float), host_out ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
  return NM_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" int nm_render_vanilla(nm_ctx* ctx, int coarse_slot, int fine_slot, const nm_camera* cam,
                                 const nm_render_opts* opt, int64_t pix0, int64_t n, const int32_t* pixels, float* rgb,
                                 float* depth, int32_t host_out, void* stream) {
  NM_ENTER(ctx);
  TRY(check_common(ctx, cam, opt, pix0, n, pixels, "nm_render_vanilla"));
  if (!slot_ok(ctx, coarse_slot) || (fine_slot >= 0 && !slot_ok(ctx, fine_slot)))
    NM_FAIL(ctx, NM_ERR_STATE, "nm_render_vanilla: net slot not packed");
  if (!rgb) NM_FAIL(ctx, NM_ERR_INVALID, "nm_render_vanilla: null rgb");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = opt->samples_per_ray, N = fine_slot >= 0 ? opt->importance_samples_per_ray : 0;
  const int64_t C = std::min<int64_t>(chunk_of(opt), n > 0 ? n : 1);
  ctx->last_mlp_evals = 0; ctx->last_hit_rays = 0;
  size_t bytes = (size_t)C * (6 + S + 4 * S + S + (S + N) + 4 * (S + N) + 4) * sizeof(float) + 16 * 256;
  Arena A;
  TRY(nm_impl_workspace(ctx, bytes, &A.base));
  float* o = A.take<float>(3 * C); float* d = A.take<float>(3 * C);
  float* z_c = A.take<float>(C * S); float* raw_c = A.take<float>(4 * C * S); float* w_c = A.take<float>(C * S);
  float* z_f = A.take<float>(C * (S + N)); float* raw_f = A.take<float>(4 * C * (S + N));
  float* rgb_s = A.take<float>(3 * C); float* dep_s = A.take<float>(C);
  if (A.off > ctx->ws_bytes) NM_FAIL(ctx, NM_ERR_STATE, "render: workspace arena overflow (internal sizing bug)");
  for (int64_t i = 0; i < n; i += C) {
    int64_t c = std::min<int64_t>(C, n - i);
    TRY(nm_impl_raygen(ctx, cam, 1, pix0 + i, c, nullptr, pixels ? pixels + i : nullptr, o, d, st));                       // shot_all_rays (:122)
    float *raw, *z; int St;
    TRY(bkg_pass(ctx, coarse_slot, fine_slot, opt, o, d, c, z_c, raw_c, w_c, z_f, raw_f, &raw, &z, &St, st));
    float* rgb_dst = host_out ? rgb_s : rgb + 3 * i;
    float* dep_dst = host_out ? dep_s : (depth ? depth + i : nullptr);
    TRY(nm_raw2outputs(ctx, raw, z, d, c, St, nullptr, 1.f, opt->white_bkg, rgb_dst, nullptr, nullptr, nullptr, dep_dst, st));
    if (host_out) {
      TRY(copy_out(ctx, rgb + 3 * i, rgb_s, 3 * c, 1, st));
      if (depth) TRY(copy_out(ctx, depth + i, dep_s, c, 1, st));
      NM_CHECK_CUDA(ctx, cudaStreamSynchronize(st));   // staging buffers are reused by the next chunk
    }
  }
  return NM_OK;
}

// ---------------------------------------------------------------------------------------------
// human branch for the hit rays of one chunk: samples, warp, net, (optional) own composite.
// Produces raw_h [Rh,S,4], z_h [Rh,S] for the compacted hit rays.
static int human_branch(nm_ctx* ctx, int slot, int actor, const nm_render_opts* opt, const float* oh, const float* dh,
                        const float* nh, const float* fh, int64_t Rh, float* pts, float* can_pts, float* can_dirs,
                        float* z_h, float* raw_h, bool render_can, cudaStream_t st) {
  const int S = opt->samples_per_ray;
  if (render_can) {                                                          // (:214-216)
    TRY(nm_ray_to_samples(ctx, oh, dh, nh, fh, 0, 0, Rh, S, 0, nullptr, nullptr, nullptr, z_h, st));
    TRY(nm_mlp_forward_rays(ctx, slot, opt->mlp_mode, oh, dh, z_h, Rh, S, raw_h, st));
  } else {
    TRY(nm_ray_to_samples(ctx, oh, dh, nh, fh, 0, 0, Rh, S, 0, nullptr, pts, nullptr, z_h, st));
    TRY(nm_warp_to_canonical(ctx, actor, pts, Rh, S, can_pts, can_dirs, nullptr, nullptr, st));   // (:218-225)
    TRY(nm_mlp_forward(ctx, slot, opt->mlp_mode, can_pts, can_dirs, Rh * S, 0, raw_h, st));
  }
  ctx->last_mlp_evals += Rh * S;
  return NM_OK;
}

static int compact(nm_ctx* ctx, const float* near_v, const float* far_v, int64_t c, int32_t* hit_idx, int64_t* Rh,
                   cudaStream_t st) {
  NM_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->d_counter, 0, sizeof(int32_t), st));
  k_compact_hits<<<(unsigned)((c + 255) / 256), 256, 0, st>>>(near_v, far_v, c, hit_idx, ctx->d_counter);
  NM_CHECK_LAUNCH(ctx);
  NM_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_counter, ctx->d_counter, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  NM_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
  *Rh = ctx->h_counter[0];
  return NM_OK;
}

extern "C" int nm_render_smpl_nerf(nm_ctx* ctx, int human_slot, int actor, const nm_camera* cam,
                                   const nm_render_opts* opt, int64_t pix0, int64_t n, const int32_t* pixels, float* rgb,
                                   float* depth, float* acc, int32_t host_out, void* stream) {
  NM_ENTER(ctx);
  TRY(check_common(ctx, cam, opt, pix0, n, pixels, "nm_render_smpl_nerf"));
  if (!slot_ok(ctx, human_slot)) NM_FAIL(ctx, NM_ERR_STATE, "nm_render_smpl_nerf: net slot not packed");
  if (actor < 0 || actor >= NM_MAX_ACTORS || !ctx->meshes[actor].set)
    NM_FAIL(ctx, NM_ERR_STATE, "nm_render_smpl_nerf: mesh not set");
  if (!rgb) NM_FAIL(ctx, NM_ERR_INVALID, "nm_render_smpl_nerf: null rgb");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = opt->samples_per_ray;
  const int64_t C = std::min<int64_t>(chunk_of(opt), n > 0 ? n : 1);
  ctx->last_mlp_evals = 0; ctx->last_hit_rays = 0;
  size_t bytes = (size_t)C * (6 + 2 + 1 + 8 + 9 * S + S + 4 * S + 5 + 5) * sizeof(float) + 32 * 256;
  Arena A;
  TRY(nm_impl_workspace(ctx, bytes, &A.base));
  float* o = A.take<float>(3 * C); float* d = A.take<float>(3 * C);
  float* nr = A.take<float>(C); float* fr = A.take<float>(C);
  int32_t* hit = A.take<int32_t>(C);
  float* oh = A.take<float>(3 * C); float* dh = A.take<float>(3 * C); float* nh = A.take<float>(C); float* fh = A.take<float>(C);
  float* pts = A.take<float>(3 * C * S); float* cpts = A.take<float>(3 * C * S); float* cdirs = A.take<float>(3 * C * S);
  float* z_h = A.take<float>(C * S); float* raw_h = A.take<float>(4 * C * S);
  float* rgb_h = A.take<float>(3 * C); float* dep_h = A.take<float>(C); float* acc_h = A.take<float>(C);
  float* rgb_s = A.take<float>(3 * C); float* dep_s = A.take<float>(C); float* acc_s = A.take<float>(C);
  const NmMesh& mesh = ctx->meshes[actor];
  if (A.off > ctx->ws_bytes) NM_FAIL(ctx, NM_ERR_STATE, "render: workspace arena overflow (internal sizing bug)");
  for (int64_t i = 0; i < n; i += C) {
    int64_t c = std::min<int64_t>(C, n - i);
    TRY(nm_impl_raygen(ctx, cam, 0, pix0 + i, c, nullptr, pixels ? pixels + i : nullptr, o, d, st));                        // shot_rays (:186)
    TRY(nm_impl_near_far_mesh(ctx, mesh, o, d, c, opt->geo_threshold, nr, fr, st));   // (:198)
    int64_t Rh = 0;
    TRY(compact(ctx, nr, fr, c, hit, &Rh, st));
    ctx->last_hit_rays += Rh;
    float* rgb_dst = host_out ? rgb_s : rgb + 3 * i;
    float* dep_dst = host_out ? dep_s : (depth ? depth + i : dep_s);
    float* acc_dst = host_out ? acc_s : (acc ? acc + i : acc_s);
    LAUNCH1D(k_fill, 3 * c, st, rgb_dst, 3 * c, opt->white_bkg ? 1.f : 0.f);             // miss rays (:199-205)
    LAUNCH1D(k_fill, c, st, dep_dst, c, 0.f);
    LAUNCH1D(k_fill, c, st, acc_dst, c, 0.f);
    if (Rh > 0) {
      LAUNCH1D(k_gather_rays, Rh, st, hit, (int)Rh, o, d, nr, fr, oh, dh, nh, fh);
      TRY(human_branch(ctx, human_slot, actor, opt, oh, dh, nh, fh, Rh, pts, cpts, cdirs, z_h, raw_h, opt->render_can != 0, st));
      TRY(nm_raw2outputs(ctx, raw_h, z_h, dh, Rh, S, nullptr, opt->interval_comp, opt->white_bkg, rgb_h, nullptr, acc_h,
                         nullptr, dep_h, st));                                          // (:229-230)
      LAUNCH1D(k_move_rows, Rh * 3, st, hit, Rh, 3, rgb_h, rgb_dst, 1);                  // (:231-233)
      LAUNCH1D(k_move_rows, Rh, st, hit, Rh, 1, dep_h, dep_dst, 1);
      LAUNCH1D(k_move_rows, Rh, st, hit, Rh, 1, acc_h, acc_dst, 1);
    }
    if (host_out) {
      TRY(copy_out(ctx, rgb + 3 * i, rgb_s, 3 * c, 1, st));
      if (depth) TRY(copy_out(ctx, depth + i, dep_s, c, 1, st));
      if (acc) TRY(copy_out(ctx, acc + i, acc_s, c, 1, st));
      NM_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    }
  }
  return NM_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" int nm_render_hybrid(nm_ctx* ctx, int coarse_slot, int fine_slot, int32_t n_actors,
                                const int32_t* human_slots, const int32_t* actors, int32_t multi_person,
                                const nm_camera* cam, const nm_render_opts* opt, int64_t pix0, int64_t n,
                                const int32_t* pixels, float* rgb, float* depth, float* acc, int32_t host_out, void* stream) {
  NM_ENTER(ctx);
  TRY(check_common(ctx, cam, opt, pix0, n, pixels, "nm_render_hybrid"));
  if (!slot_ok(ctx, coarse_slot) || (fine_slot >= 0 && !slot_ok(ctx, fine_slot)))
    NM_FAIL(ctx, NM_ERR_STATE, "nm_render_hybrid: bkg net slot not packed");
  if (n_actors < 1 || n_actors > NM_MAX_ACTORS || !human_slots || !actors || (!multi_person && n_actors != 1))
    NM_FAIL(ctx, NM_ERR_INVALID, "nm_render_hybrid: bad actor list");
  for (int a = 0; a < n_actors; ++a)
    if (!slot_ok(ctx, human_slots[a]) || actors[a] < 0 || actors[a] >= NM_MAX_ACTORS || !ctx->meshes[actors[a]].set)
      NM_FAIL(ctx, NM_ERR_STATE, "nm_render_hybrid: human net or mesh not set");
  if (!rgb) NM_FAIL(ctx, NM_ERR_INVALID, "nm_render_hybrid: null rgb");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = opt->samples_per_ray, N = fine_slot >= 0 ? opt->importance_samples_per_ray : 0;
  const int Sb = S + N;
  const int n_lists = 1 + n_actors;
  const int Sm = Sb + n_actors * S;
  const int64_t C = std::min<int64_t>(chunk_of(opt), n > 0 ? n : 1);
  ctx->last_mlp_evals = 0; ctx->last_hit_rays = 0;
  size_t per_ray = 6 + 2 + 1 + 8 + 9 * S + (size_t)5 * S + 6 * S + 5 * Sb + 5 * Sb + 5 * Sm + 8 + 5 + 3 * (size_t)n_actors +
                   (multi_person ? (size_t)10 * S * n_actors + 2 : 0);
  size_t bytes = (size_t)C * per_ray * sizeof(float) + 96 * 256;
  Arena A;
  TRY(nm_impl_workspace(ctx, bytes, &A.base));
  float* o = A.take<float>(3 * C); float* d = A.take<float>(3 * C);
  float* nr_a[NM_MAX_ACTORS]; float* fr_a[NM_MAX_ACTORS]; int32_t* hit_a[NM_MAX_ACTORS];
  for (int a = 0; a < n_actors; ++a) { nr_a[a] = A.take<float>(C); fr_a[a] = A.take<float>(C); hit_a[a] = A.take<int32_t>(C); }
  int32_t* hit = A.take<int32_t>(C);                                                // rays any actor hits (multi-person)
  float* oh = A.take<float>(3 * C); float* dh = A.take<float>(3 * C); float* nh = A.take<float>(C); float* fh = A.take<float>(C);
  float* pts = A.take<float>(3 * C * S); float* cpts = A.take<float>(3 * C * S); float* cdirs = A.take<float>(3 * C * S);
  float* z_h = A.take<float>(C * S); float* raw_h = A.take<float>(4 * C * S);
  float* z_c = A.take<float>(C * S); float* raw_c = A.take<float>(4 * C * S); float* w_c = A.take<float>(C * S);
  float* z_f = A.take<float>(C * Sb); float* raw_f = A.take<float>(4 * C * Sb);
  float* z_bh = A.take<float>(C * Sb); float* raw_bh = A.take<float>(4 * C * Sb);   // bkg rows of the hit rays
  float* z_m = A.take<float>(C * Sm); float* raw_m = A.take<float>(4 * C * Sm);
  float* rgb_h = A.take<float>(3 * C); float* dep_h = A.take<float>(C); float* acc_h = A.take<float>(C);
  float* rgb_s = A.take<float>(3 * C); float* dep_s = A.take<float>(C); float* acc_s = A.take<float>(C);
  float* z_a[NM_MAX_ACTORS]; float* raw_a[NM_MAX_ACTORS];
  float* zu_a[NM_MAX_ACTORS]; float* rawu_a[NM_MAX_ACTORS];      // the same rows, compacted to the rays any actor hits
  float *flag = nullptr, *zeros = nullptr;
  if (multi_person) {
    for (int a = 0; a < n_actors; ++a) {
      z_a[a] = A.take<float>(C * S); raw_a[a] = A.take<float>(4 * C * S);
      zu_a[a] = A.take<float>(C * S); rawu_a[a] = A.take<float>(4 * C * S);
    }
    flag = A.take<float>(C); zeros = A.take<float>(C);
  }

  if (A.off > ctx->ws_bytes) NM_FAIL(ctx, NM_ERR_STATE, "render: workspace arena overflow (internal sizing bug)");
  if (!ctx->ev_counts) NM_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_counts, cudaEventDisableTiming));
  for (int64_t i = 0; i < n; i += C) {
    int64_t c = std::min<int64_t>(C, n - i);
    TRY(nm_impl_raygen(ctx, cam, 0, pix0 + i, c, nullptr, pixels ? pixels + i : nullptr, o, d, st));                        // shot_rays (:271 / :386)
    // Hit lists of every actor (:299 / :415) and, for several actors, of their union FIRST: the counts travel to the host
    // while the background networks run, so the host never waits for them with the GPU idle.
    int64_t Rh_a[NM_MAX_ACTORS]; int64_t Ru = 0;
    NM_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->d_counter, 0, sizeof(int32_t) * (n_actors + 1), st));
    if (multi_person) { LAUNCH1D(k_fill, c, st, flag, c, 0.f); LAUNCH1D(k_fill, c, st, zeros, c, 0.f); }
    for (int a = 0; a < n_actors; ++a) {
      TRY(nm_impl_near_far_mesh(ctx, ctx->meshes[actors[a]], o, d, c, opt->geo_threshold, nr_a[a], fr_a[a], st));
      LAUNCH1D(k_compact_hits, c, st, nr_a[a], fr_a[a], c, hit_a[a], ctx->d_counter + a);
      if (multi_person) LAUNCH1D(k_mark_hits, c, st, nr_a[a], fr_a[a], c, flag);
    }
    if (multi_person) LAUNCH1D(k_compact_hits, c, st, zeros, flag, c, hit, ctx->d_counter + n_actors);
    NM_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_counter, ctx->d_counter, sizeof(int32_t) * (n_actors + 1), cudaMemcpyDeviceToHost, st));
    NM_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_counts, st));
    float *raw_b, *z_b; int St;
    TRY(bkg_pass(ctx, coarse_slot, fine_slot, opt, o, d, c, z_c, raw_c, w_c, z_f, raw_f, &raw_b, &z_b, &St, st));
    NM_CHECK_CUDA(ctx, cudaEventSynchronize(ctx->ev_counts));
    for (int a = 0; a < n_actors; ++a) { Rh_a[a] = ctx->h_counter[a]; ctx->last_hit_rays += Rh_a[a]; }
    Ru = ctx->h_counter[n_actors];
    float* rgb_dst = host_out ? rgb_s : rgb + 3 * i;
    float* dep_dst = host_out ? dep_s : (depth ? depth + i : dep_s);
    float* acc_dst = host_out ? acc_s : (acc ? acc + i : acc_s);
    if (!multi_person) {
      // all rays first get the background-only composite (miss rays keep it, :301-311)
      TRY(nm_raw2outputs(ctx, raw_b, z_b, d, c, St, nullptr, 1.f, opt->white_bkg, rgb_dst, nullptr, nullptr, nullptr, dep_dst, st));
      LAUNCH1D(k_fill, c, st, acc_dst, c, 0.f);
      const int64_t Rh = Rh_a[0];
      const int32_t* hit = hit_a[0];
      const float *nr = nr_a[0], *fr = fr_a[0];
      if (Rh > 0) {
        LAUNCH1D(k_gather_rays, Rh, st, hit, (int)Rh, o, d, nr, fr, oh, dh, nh, fh);
        TRY(human_branch(ctx, human_slots[0], actors[0], opt, oh, dh, nh, fh, Rh, pts, cpts, cdirs, z_h, raw_h, false, st));
        LAUNCH1D(k_move_rows, Rh * St, st, hit, Rh, St, z_b, z_bh, 0);
        LAUNCH1D(k_move_rows, Rh * St * 4, st, hit, Rh, St * 4, raw_b, raw_bh, 0);
        const float* zl[2] = {z_bh, z_h};
        const float* rl[2] = {raw_bh, raw_h};
        const int32_t Sl[2] = {St, S};
        TRY(nm_merge_samples(ctx, 2, zl, rl, Sl, Rh, z_m, raw_m, st));                  // (:330-337)
        TRY(nm_raw2outputs(ctx, raw_m, z_m, dh, Rh, St + S, nullptr, 1.f, opt->white_bkg, rgb_h, nullptr, nullptr, nullptr,
                           dep_h, st));                                                // (:338-343)
        TRY(nm_raw2outputs(ctx, raw_h, z_h, dh, Rh, S, nullptr, 1.f, opt->white_bkg, nullptr, nullptr, acc_h, nullptr,
                           nullptr, st));                                              // (:345-350)
        LAUNCH1D(k_move_rows, Rh * 3, st, hit, Rh, 3, rgb_h, rgb_dst, 1);
        LAUNCH1D(k_move_rows, Rh, st, hit, Rh, 1, dep_h, dep_dst, 1);
        LAUNCH1D(k_move_rows, Rh, st, hit, Rh, 1, acc_h, acc_dst, 1);
      }
    } else {
      // Rays no actor hits carry only zero-density placeholders behind the background samples (:418-419): their
      // composite is the background composite whose last interval ends at the first placeholder (z = 2 far), no
      // sort needed.  Rays at least one actor hits go through the full z-sorted merge (:441-448), compacted.
      TRY(nm_impl_raw2outputs_zend(ctx, raw_b, z_b, d, c, St, opt->white_bkg, opt->far_bkg * 2.f, rgb_dst, dep_dst, st));
      for (int a = 0; a < n_actors; ++a) {
        LAUNCH1D(k_fill_placeholder, c * S, st, z_a[a], (float4*)raw_a[a], c, S, opt->far_bkg * 2.f, opt->far_bkg * 3.f);
        const int64_t Rh = Rh_a[a];
        if (Rh > 0) {
          LAUNCH1D(k_gather_rays, Rh, st, hit_a[a], (int)Rh, o, d, nr_a[a], fr_a[a], oh, dh, nh, fh);
          TRY(human_branch(ctx, human_slots[a], actors[a], opt, oh, dh, nh, fh, Rh, pts, cpts, cdirs, z_h, raw_h, false, st));
          LAUNCH1D(k_move_rows, Rh * S, st, hit_a[a], Rh, S, z_h, z_a[a], 1);              // (:438-439)
          LAUNCH1D(k_move_rows, Rh * S * 4, st, hit_a[a], Rh, S * 4, raw_h, raw_a[a], 1);
        }
      }
      if (Ru > 0) {
        const float* zl[1 + NM_MAX_ACTORS]; const float* rl[1 + NM_MAX_ACTORS]; int32_t Sl[1 + NM_MAX_ACTORS];
        LAUNCH1D(k_gather_rays, Ru, st, hit, (int)Ru, o, d, zeros, flag, oh, dh, nh, fh);
        LAUNCH1D(k_move_rows, Ru * St, st, hit, Ru, St, z_b, z_bh, 0);
        LAUNCH1D(k_move_rows, Ru * St * 4, st, hit, Ru, St * 4, raw_b, raw_bh, 0);
        zl[0] = z_bh; rl[0] = raw_bh; Sl[0] = St;
        for (int a = 0; a < n_actors; ++a) {
          LAUNCH1D(k_move_rows, Ru * S, st, hit, Ru, S, z_a[a], zu_a[a], 0);
          LAUNCH1D(k_move_rows, Ru * S * 4, st, hit, Ru, S * 4, raw_a[a], rawu_a[a], 0);
          zl[1 + a] = zu_a[a]; rl[1 + a] = rawu_a[a]; Sl[1 + a] = S;
        }
        TRY(nm_merge_samples(ctx, n_lists, zl, rl, Sl, Ru, z_m, raw_m, st));              // (:441-448)
        TRY(nm_raw2outputs(ctx, raw_m, z_m, dh, Ru, Sm, nullptr, 1.f, opt->white_bkg, rgb_h, nullptr, nullptr, nullptr,
                           dep_h, st));                                                // (:449-454)
        LAUNCH1D(k_move_rows, Ru * 3, st, hit, Ru, 3, rgb_h, rgb_dst, 1);
        LAUNCH1D(k_move_rows, Ru, st, hit, Ru, 1, dep_h, dep_dst, 1);
      }
      LAUNCH1D(k_fill, c, st, acc_dst, c, 0.f);
    }
    if (host_out) {
      TRY(copy_out(ctx, rgb + 3 * i, rgb_s, 3 * c, 1, st));
      if (depth) TRY(copy_out(ctx, depth + i, dep_s, c, 1, st));
      if (acc) TRY(copy_out(ctx, acc + i, acc_s, c, 1, st));
      NM_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    }
  }
  return NM_OK;
}

// ---------------------------------------------------------------------------------------------
// frame reassembly after the gather of the ranks' pixel-list shards
__global__ void k_assemble_frame(const float* __restrict__ shards, long long per, int planes, long long total,
                                 const int32_t* __restrict__ pixels_all, float* __restrict__ rgb, float* __restrict__ depth,
                                 float* __restrict__ acc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long pix = pixels_all[i];
  if (pix < 0) return;
  const long long r = i / per, j = i - r * per;
  const float* base = shards + r * planes * per;
  rgb[3 * pix + 0] = base[3 * j + 0];
  rgb[3 * pix + 1] = base[3 * j + 1];
  rgb[3 * pix + 2] = base[3 * j + 2];
  if (depth) depth[pix] = base[3 * per + j];
  if (acc && planes > 4) acc[pix] = base[4 * per + j];
}

extern "C" int nm_assemble_frame(nm_ctx* ctx, const float* shards, int32_t world, int64_t per, int32_t planes,
                                 const int32_t* pixels_all, float* rgb, float* depth, float* acc, void* stream) {
  NM_ENTER(ctx);
  if (!shards || !pixels_all || !rgb || world < 1 || per < 0 || (planes != 4 && planes != 5))
    NM_FAIL(ctx, NM_ERR_INVALID, "nm_assemble_frame: bad argument");
  const long long total = (long long)world * per;
  cudaStream_t st = (cudaStream_t)stream;
  LAUNCH1D(k_assemble_frame, total, st, shards, (long long)per, (int)planes, total, pixels_all, rgb, depth, acc);
  return NM_OK;
}
