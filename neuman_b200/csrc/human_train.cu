// Entry points of the human trainer's differentiable observation->canonical map; the kernels are in
// human_train_kernels.cuh (shared with the host emulation under tests/emu/).
//
//   nm_warp_diff_forward / nm_warp_diff_backward           <- utils/ray_utils.py:69-93 (warp_samples_to_canonical_diff) and
//                                                             its autograd: the form the reference's trainer consumes
//                                                             (T_interp_inv per sample, trainers/human_nerf_trainer.py:266-272)
//   nm_human_canonicalize / nm_human_canonicalize_backward <- the same plus trainers/human_nerf_trainer.py:272-276 fused:
//                                                             canonical points (+ offset) and directions without ever
//                                                             writing the n x 4 x 4 matrices to HBM
#include "nm_internal.cuh"
#include "human_train_kernels.cuh"

static inline unsigned wd_blocks(long long n, int b) { return (unsigned)((n + b - 1) / b); }

extern "C" int nm_warp_diff_forward(nm_ctx* ctx, const int32_t* f_id, const double* closest, const float* verts,
                                    const int32_t* faces, const float* T, int64_t n, float* Tinv, void* stream) {
  NM_ENTER(ctx);
  if (n == 0) return NM_OK;
  if (n < 0 || !f_id || !closest || !verts || !faces || !T || !Tinv)
    NM_FAIL(ctx, NM_ERR_INVALID, "nm_warp_diff_forward: null argument");
  k_wd_forward<<<wd_blocks(n, 128), 128, 0, (cudaStream_t)stream>>>(f_id, closest, verts, faces, T, nullptr, nullptr,
                                                                    (long long)n, Tinv, nullptr);
  NM_CHECK_LAUNCH(ctx);
  return NM_OK;
}

extern "C" int nm_warp_diff_backward(nm_ctx* ctx, const int32_t* f_id, const double* closest, const float* verts,
                                     const int32_t* faces, const float* T, int64_t n, const float* g_Tinv, int32_t n_verts,
                                     float* g_T, float* g_verts, void* stream) {
  NM_ENTER(ctx);
  if (n < 0 || n_verts <= 0 || (!g_T && !g_verts)) NM_FAIL(ctx, NM_ERR_INVALID, "nm_warp_diff_backward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (g_T) NM_CHECK_CUDA(ctx, cudaMemsetAsync(g_T, 0, (size_t)n_verts * 16 * sizeof(float), st));
  if (g_verts) NM_CHECK_CUDA(ctx, cudaMemsetAsync(g_verts, 0, (size_t)n_verts * 3 * sizeof(float), st));
  if (n == 0) return NM_OK;
  if (!f_id || !closest || !verts || !faces || !T || !g_Tinv) NM_FAIL(ctx, NM_ERR_INVALID, "nm_warp_diff_backward: null argument");
  k_wd_backward<<<wd_blocks(n, 128), 128, 0, st>>>(f_id, closest, verts, faces, T, nullptr, g_Tinv, nullptr, (long long)n, g_T,
                                                   g_verts);
  NM_CHECK_LAUNCH(ctx);
  return NM_OK;
}

extern "C" int nm_human_canonicalize(nm_ctx* ctx, const int32_t* f_id, const double* closest, const float* verts,
                                     const int32_t* faces, const float* T, const float* pts, const float* offset, int64_t R,
                                     int32_t S, float* can_pts, float* can_dirs, void* stream) {
  NM_ENTER(ctx);
  if (R == 0) return NM_OK;
  if (R < 0 || S < 2) NM_FAIL(ctx, NM_ERR_INVALID, "nm_human_canonicalize: need R >= 0 and S >= 2 samples per ray");
  if (!f_id || !closest || !verts || !faces || !T || !pts || !can_pts)
    NM_FAIL(ctx, NM_ERR_INVALID, "nm_human_canonicalize: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)R * S;
  k_wd_forward<<<wd_blocks(n, 128), 128, 0, st>>>(f_id, closest, verts, faces, T, pts, offset, n, nullptr, can_pts);
  NM_CHECK_LAUNCH(ctx);
  if (can_dirs) {
    k_wd_dirs<<<wd_blocks(n, 256), 256, 0, st>>>(can_pts, (long long)R, S, can_dirs);
    NM_CHECK_LAUNCH(ctx);
  }
  return NM_OK;
}

extern "C" int nm_human_canonicalize_backward(nm_ctx* ctx, const int32_t* f_id, const double* closest, const float* verts,
                                              const int32_t* faces, const float* T, const float* pts, const float* can_pts,
                                              const float* g_can_pts, const float* g_can_dirs, int64_t R, int32_t S,
                                              int32_t n_verts, float* g_offset, float* g_T, float* g_verts, void* stream) {
  NM_ENTER(ctx);
  if (R < 0 || S < 2 || n_verts <= 0 || !g_offset)
    NM_FAIL(ctx, NM_ERR_INVALID, "nm_human_canonicalize_backward: bad argument (g_offset is required: it is the total dL/dcan_pts)");
  cudaStream_t st = (cudaStream_t)stream;
  if (g_T) NM_CHECK_CUDA(ctx, cudaMemsetAsync(g_T, 0, (size_t)n_verts * 16 * sizeof(float), st));
  if (g_verts) NM_CHECK_CUDA(ctx, cudaMemsetAsync(g_verts, 0, (size_t)n_verts * 3 * sizeof(float), st));
  if (R == 0) return NM_OK;
  if (!f_id || !closest || !verts || !faces || !T || !pts || !can_pts || (!g_can_pts && !g_can_dirs))
    NM_FAIL(ctx, NM_ERR_INVALID, "nm_human_canonicalize_backward: null argument");
  const long long n = (long long)R * S;
  k_wd_dirs_backward<<<wd_blocks(n, 256), 256, 0, st>>>(can_pts, g_can_pts, g_can_dirs, (long long)R, S, g_offset);
  NM_CHECK_LAUNCH(ctx);
  if (g_T || g_verts) {
    k_wd_backward<<<wd_blocks(n, 128), 128, 0, st>>>(f_id, closest, verts, faces, T, pts, nullptr, g_offset, n, g_T, g_verts);
    NM_CHECK_LAUNCH(ctx);
  }
  return NM_OK;
}
