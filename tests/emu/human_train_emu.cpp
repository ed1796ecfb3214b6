// TEST INFRASTRUCTURE ONLY -- C entry points that run the kernels of neuman_b200/csrc/human_train_kernels.cuh on host
// arrays through the serial emulation in cuda_emu.h, with the launch shapes human_train.cu uses.
#include "cuda_emu.h"
#include "human_train_kernels.cuh"
#include "smpl_train_kernels.cuh"

static inline unsigned blocks_for(long long n, int b) { return (unsigned)((n + b - 1) / b); }

extern "C" {

void emu_wd_forward(const int* f_id, const double* closest, const float* verts, const int* faces, const float* T,
                    const float* pts, const float* offset, long long n, float* Tinv, float* can_pts) {
  EMU_LAUNCH(k_wd_forward, blocks_for(n, 128), 128, f_id, closest, verts, faces, T, pts, offset, n, Tinv, can_pts);
}

void emu_wd_dirs(const float* can_pts, long long R, int S, float* can_dirs) {
  EMU_LAUNCH(k_wd_dirs, blocks_for(R * S, 128), 128, can_pts, R, S, can_dirs);
}

void emu_wd_dirs_backward(const float* can_pts, const float* g_pts, const float* g_dirs, long long R, int S, float* g_total) {
  EMU_LAUNCH(k_wd_dirs_backward, blocks_for(R * S, 128), 128, can_pts, g_pts, g_dirs, R, S, g_total);
}

void emu_wd_backward(const int* f_id, const double* closest, const float* verts, const int* faces, const float* T,
                     const float* pts, const float* g_Tinv, const float* g_can, long long n, float* g_T, float* g_verts) {
  EMU_LAUNCH(k_wd_backward, blocks_for(n, 128), 128, f_id, closest, verts, faces, T, pts, g_Tinv, g_can, n, g_T, g_verts);
}

// forward of HumanNeRF.vertex_forward from the LBS intermediates (T_pose, T_da [V,16], rest = v_shaped [V,3])
void emu_smplt_scene_forward(const float* T_pose, const float* T_da, const float* rest, const float* alignment, float scale,
                             int nv, float* T_out, float* world) {
  EMU_LAUNCH(k_smplt_scene_forward, blocks_for(nv, 128), 128, T_pose, T_da, rest, alignment, scale, nv, T_out, world);
}

// the backward pipeline exactly as nm_smpl_scene_backward launches it (smpl.cu); scratch buffers are the caller's
void emu_smplt_scene_backward(const float* T_pose, const float* T_da, const float* rest, const float* J, const float* pose,
                              const float* da_pose, const float* alignment, float scale, const float* W, const float* Jreg,
                              const float* shapedirs, const int* parents, int nv, int nj, int nb, const float* gT,
                              const float* gworld, float* gP, float* gD, float* grest, float* gpre, float* gA_pose,
                              float* gA_da, float* gJ, float* g_pose, float* g_betas, float* g_alignment) {
  memset(gpre, 0, 16 * sizeof(float));
  memset(gA_pose, 0, (size_t)nj * 16 * sizeof(float));
  memset(gA_da, 0, (size_t)nj * 16 * sizeof(float));
  memset(g_betas, 0, (size_t)nb * sizeof(float));
  const int chunks = (nv + SMPLT_VPT - 1) / SMPLT_VPT;
  EMU_LAUNCH(k_smplt_scene_backward, blocks_for(chunks, 64), 64, T_pose, T_da, rest, alignment, scale, gT, gworld, nv, gP, gD,
             grest, gpre);
  EMU_LAUNCH(k_smplt_blend_backward, blocks_for(nv, 128), 128, W, gP, gD, nv, nj, gA_pose, gA_da);
  SmpltParents par;
  for (int j = 0; j < nj; ++j) par.p[j] = parents[j];
  EMU_LAUNCH(k_smplt_chain_backward, 1, 32, pose, da_pose, J, par, nj, gA_pose, gA_da, gpre, scale, g_pose, gJ, g_alignment);
  EMU_LAUNCH(k_smplt_vshaped_backward, blocks_for(nv, 128), 128, Jreg, gJ, nv, nj, grest);
  EMU_LAUNCH(k_smplt_betas_backward, blocks_for((nv * 3 + 63) / 64, 64), 64, shapedirs, grest, nv * 3, nb, g_betas);
}

}  // extern "C"
