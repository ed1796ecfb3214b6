// Training-time SMPL scene transforms and their adjoint:
//
//   forward   HumanNeRF.vertex_forward models/human_nerf.py:92-122 (float32 torch in the reference):
//             T_da2scene[v] = S . alignment^T . T_t2pose[v] . inv(T_t2da[v]),  world[v] = T_da2scene[v] . (T_t2da[v] . [v_shaped;1])
//   backward  what loss.backward() (trainers/human_nerf_trainer.py:205) propagates from the warp's dL/dT_da2scene and
//             dL/dworld_verts to the trainer's parameters `poses`, `betas`, `alignments` (models/human_nerf.py:36-38):
//             through the blend T[v] = sum_j W[v,j] A_j (models/smpl.py:344-345), the relative transforms / kinematic chain
//             (batch_rigid_transform :454-505), Rodrigues (batch_rodrigues :407-438), the joint regressor (:363) and the
//             shape blend shapes (:383).
//
// Same restricted CUDA subset as human_train_kernels.cuh (tests/emu/ compiles these bodies for the host): one thread per
// element or per small chunk, atomics for the reductions -- the whole problem is 6890 vertices and 24 joints, latency
// matters here, not throughput.
#pragma once
#include "train_common_kernels.cuh"

#define SMPLT_MAX_J 64
#define SMPLT_VPT 8            // vertices per thread in the per-vertex backward (bounds the atomics on the 16 `pre` sums)

struct SmpltParents { int p[SMPLT_MAX_J]; };

NM_DEV void smplt_mm4(const float* a, const float* b, float* o) {          // o = a b
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float acc = 0.f;
      for (int k = 0; k < 4; ++k) acc += a[4 * i + k] * b[4 * k + j];
      o[4 * i + j] = acc;
    }
}
NM_DEV void smplt_mm4_tn(const float* a, const float* b, float* o) {       // o = a^T b
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float acc = 0.f;
      for (int k = 0; k < 4; ++k) acc += a[4 * k + i] * b[4 * k + j];
      o[4 * i + j] = acc;
    }
}
NM_DEV void smplt_mm4_nt(const float* a, const float* b, float* o) {       // o = a b^T
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float acc = 0.f;
      for (int k = 0; k < 4; ++k) acc += a[4 * i + k] * b[4 * j + k];
      o[4 * i + j] = acc;
    }
}

// pre = S . alignment^T (models/human_nerf.py:108-111); alignment [16] row-major on the device
NM_DEV void smplt_pre(const float* __restrict__ alignment, float scale, float* pre) {
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) pre[4 * a + b] = alignment[4 * b + a] * (a < 3 ? scale : 1.f);
}

// per-vertex forward pieces shared by the forward and the backward kernels
struct SmpltVert {
  float Di[16], M[16], R[16], dv[4];
};
NM_DEV void smplt_vertex(const float* __restrict__ P, const float* __restrict__ D, const float* __restrict__ rest,
                         const float* pre, SmpltVert& o) {
  double m[16], inv[16];
  for (int k = 0; k < 16; ++k) m[k] = (double)D[k];
  if (!wd_inv4(m, inv))
    for (int k = 0; k < 16; ++k) inv[k] = (double)NAN;
  for (int k = 0; k < 16; ++k) o.Di[k] = (float)inv[k];
  smplt_mm4(P, o.Di, o.M);
  smplt_mm4(pre, o.M, o.R);
  for (int a = 0; a < 3; ++a) o.dv[a] = ((D[4 * a] * rest[0] + D[4 * a + 1] * rest[1]) + D[4 * a + 2] * rest[2]) + D[4 * a + 3];
  o.dv[3] = 1.f;
}

// forward: T_pose, T_da [V,16] (nm LBS of the pose / the da pose), rest = v_shaped [V,3] -> T_out [V,16], world [V,3]
NM_KERNEL void k_smplt_scene_forward(const float* __restrict__ T_pose, const float* __restrict__ T_da,
                                     const float* __restrict__ rest, const float* __restrict__ alignment, float scale,
                                     int nv, float* __restrict__ T_out, float* __restrict__ world) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  float pre[16];
  smplt_pre(alignment, scale, pre);
  SmpltVert s;
  smplt_vertex(T_pose + (size_t)16 * v, T_da + (size_t)16 * v, rest + 3 * v, pre, s);
  for (int k = 0; k < 16; ++k) T_out[(size_t)16 * v + k] = s.R[k];
  if (world)
    for (int a = 0; a < 3; ++a)
      world[3 * v + a] = ((s.R[4 * a] * s.dv[0] + s.R[4 * a + 1] * s.dv[1]) + s.R[4 * a + 2] * s.dv[2]) + s.R[4 * a + 3];
}

// backward, per vertex: gT [V,16] = dL/dT_da2scene, gworld [V,3] = dL/dworld (either may be null)
//   -> gP [V,16] = dL/dT_pose[v], gD [V,16] = dL/dT_da[v], grest [V,3] = dL/dv_shaped[v] (direct part),
//      gpre [16] += dL/d(S . alignment^T)   (atomics, SMPLT_VPT vertices per thread)
NM_KERNEL void k_smplt_scene_backward(const float* __restrict__ T_pose, const float* __restrict__ T_da,
                                      const float* __restrict__ rest, const float* __restrict__ alignment, float scale,
                                      const float* __restrict__ gT, const float* __restrict__ gworld, int nv,
                                      float* __restrict__ gP, float* __restrict__ gD, float* __restrict__ grest,
                                      float* __restrict__ gpre) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int v0 = t * SMPLT_VPT;
  if (v0 >= nv) return;
  float pre[16], gpre_acc[16];
  smplt_pre(alignment, scale, pre);
  for (int k = 0; k < 16; ++k) gpre_acc[k] = 0.f;
  for (int v = v0; v < v0 + SMPLT_VPT && v < nv; ++v) {
    const float* P = T_pose + (size_t)16 * v;
    const float* D = T_da + (size_t)16 * v;
    SmpltVert s;
    smplt_vertex(P, D, rest + 3 * v, pre, s);
    float gR[16], gdv[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < 16; ++k) gR[k] = gT ? gT[(size_t)16 * v + k] : 0.f;
    if (gworld) {
      for (int a = 0; a < 3; ++a) {
        float g = gworld[3 * v + a];
        for (int b = 0; b < 4; ++b) gR[4 * a + b] += g * s.dv[b];
        for (int b = 0; b < 3; ++b) gdv[b] += s.R[4 * a + b] * g;
      }
    }
    float gM[16], tmp[16], gDi[16], gDm[16];
    smplt_mm4_tn(pre, gR, gM);                      // R = pre M
    smplt_mm4_nt(gR, s.M, tmp);
    for (int k = 0; k < 16; ++k) gpre_acc[k] += tmp[k];
    smplt_mm4_nt(gM, s.Di, tmp);                    // M = P Di: gP = gM Di^T
    for (int k = 0; k < 16; ++k) gP[(size_t)16 * v + k] = tmp[k];
    smplt_mm4_tn(P, gM, gDi);                       // gDi = P^T gM
    smplt_mm4_tn(s.Di, gDi, tmp);                   // gD = -Di^T gDi Di^T
    smplt_mm4_nt(tmp, s.Di, gDm);
    for (int k = 0; k < 16; ++k) gDm[k] = -gDm[k];
    const float rh[4] = {rest[3 * v], rest[3 * v + 1], rest[3 * v + 2], 1.f};
    float gr[3] = {0.f, 0.f, 0.f};
    for (int a = 0; a < 3; ++a) {                   // dv = D[:3,:] [rest;1]
      for (int b = 0; b < 4; ++b) gDm[4 * a + b] += gdv[a] * rh[b];
      for (int b = 0; b < 3; ++b) gr[b] += D[4 * a + b] * gdv[a];
    }
    for (int k = 0; k < 16; ++k) gD[(size_t)16 * v + k] = gDm[k];
    for (int b = 0; b < 3; ++b) grest[3 * v + b] = gr[b];
  }
  for (int k = 0; k < 16; ++k) NM_ATOMIC_ADD(gpre + k, gpre_acc[k]);
}

// T[v] = sum_j W[v,j] A_j  ->  gA_pose[j] += W[v,j] gP[v], gA_da[j] += W[v,j] gD[v]   (skinning weights are sparse:
// 4 non-zeros per vertex in SMPL)
NM_KERNEL void k_smplt_blend_backward(const float* __restrict__ W, const float* __restrict__ gP,
                                      const float* __restrict__ gD, int nv, int nj, float* __restrict__ gA_pose,
                                      float* __restrict__ gA_da) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  for (int j = 0; j < nj; ++j) {
    float w = W[(size_t)v * nj + j];
    if (w == 0.f) continue;
    for (int k = 0; k < 16; ++k) {
      NM_ATOMIC_ADD(gA_pose + 16 * j + k, w * gP[(size_t)16 * v + k]);
      NM_ATOMIC_ADD(gA_da + 16 * j + k, w * gD[(size_t)16 * v + k]);
    }
  }
}

// Rodrigues forward as models/smpl.py:407-438 (note the 1e-8 inside the norm) with the pieces the adjoint needs
struct SmpltRod { float r[3], angle, s, c, K[9], K2[9]; };
NM_DEV void smplt_rodrigues(const float* __restrict__ pose3, SmpltRod& o, float* L /* 4x4, rotation part written */) {
  for (int k = 0; k < 3; ++k) o.r[k] = pose3[k];
  float ax = o.r[0] + 1e-8f, ay = o.r[1] + 1e-8f, az = o.r[2] + 1e-8f;
  o.angle = sqrtf(ax * ax + ay * ay + az * az);
  float dx = o.r[0] / o.angle, dy = o.r[1] / o.angle, dz = o.r[2] / o.angle;
  o.s = sinf(o.angle);
  o.c = cosf(o.angle);
  const float K[9] = {0.f, -dz, dy, dz, 0.f, -dx, -dy, dx, 0.f};
  for (int k = 0; k < 9; ++k) o.K[k] = K[k];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) o.K2[3 * a + b] = K[3 * a] * K[b] + K[3 * a + 1] * K[3 + b] + K[3 * a + 2] * K[6 + b];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) L[4 * a + b] = (a == b ? 1.f : 0.f) + o.s * K[3 * a + b] + (1.f - o.c) * o.K2[3 * a + b];
}
// gRot [3x3 inside a 4x4, row stride 4] -> g_pose3 [3]
NM_DEV void smplt_rodrigues_backward(const SmpltRod& o, const float* gL, float* g_pose3) {
  float gs = 0.f, g1c = 0.f, gR[9];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      gR[3 * a + b] = gL[4 * a + b];
      gs += gR[3 * a + b] * o.K[3 * a + b];
      g1c += gR[3 * a + b] * o.K2[3 * a + b];
    }
  float gK[9];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      float acc = 0.f;
      for (int k = 0; k < 3; ++k) acc += gR[3 * a + k] * o.K[3 * b + k] + o.K[3 * k + a] * gR[3 * k + b];   // gR K^T + K^T gR
      gK[3 * a + b] = o.s * gR[3 * a + b] + (1.f - o.c) * acc;
    }
  float g_angle = gs * o.c + g1c * o.s;
  const float gd[3] = {gK[7] - gK[5], gK[2] - gK[6], gK[3] - gK[1]};
  float dot = gd[0] * o.r[0] + gd[1] * o.r[1] + gd[2] * o.r[2];
  g_angle -= dot / (o.angle * o.angle);
  for (int k = 0; k < 3; ++k) g_pose3[k] = gd[k] / o.angle + g_angle * (o.r[k] + 1e-8f) / o.angle;
}

// Adjoint of the kinematic chain + relative transforms for ONE pose vector (one thread): gA [nj,16] -> g_pose [nj*3]
// (may be null: the da pose is a constant), gJ [nj*3] += (joint positions feed both chains)
NM_DEV void smplt_chain_backward(const float* __restrict__ pose, const float* __restrict__ J, const SmpltParents& par,
                                 int nj, const float* __restrict__ gA, float* __restrict__ g_pose,
                                 float* __restrict__ gJ) {
  float G[SMPLT_MAX_J][16], L[SMPLT_MAX_J][16], gG[SMPLT_MAX_J][16];
  for (int j = 0; j < nj; ++j) {                                   // forward (models/smpl.py:479-493)
    SmpltRod rod;
    smplt_rodrigues(pose + 3 * j, rod, L[j]);
    int p = par.p[j];
    for (int a = 0; a < 3; ++a) L[j][4 * a + 3] = J[3 * j + a] - (j > 0 ? J[3 * p + a] : 0.f);
    L[j][12] = L[j][13] = L[j][14] = 0.f; L[j][15] = 1.f;
    if (j == 0) for (int k = 0; k < 16; ++k) G[0][k] = L[0][k];
    else smplt_mm4(G[p], L[j], G[j]);
  }
  for (int j = 0; j < nj; ++j) {                                   // A = G - [0 | G [J;0]]  (:500-503)
    for (int a = 0; a < 4; ++a) {
      float g3 = gA[16 * j + 4 * a + 3];
      for (int b = 0; b < 3; ++b) {
        gG[j][4 * a + b] = gA[16 * j + 4 * a + b] - g3 * J[3 * j + b];
        gJ[3 * j + b] -= g3 * G[j][4 * a + b];
      }
      gG[j][4 * a + 3] = g3;
    }
  }
  for (int j = nj - 1; j >= 0; --j) {
    float gL[16];
    int p = par.p[j];
    if (j == 0) {
      for (int k = 0; k < 16; ++k) gL[k] = gG[0][k];
    } else {
      float tmp[16];
      smplt_mm4_tn(G[p], gG[j], gL);                               // G_j = G_p L_j
      smplt_mm4_nt(gG[j], L[j], tmp);
      for (int k = 0; k < 16; ++k) gG[p][k] += tmp[k];
    }
    for (int a = 0; a < 3; ++a) {
      gJ[3 * j + a] += gL[4 * a + 3];
      if (j > 0) gJ[3 * p + a] -= gL[4 * a + 3];
    }
    if (g_pose) {
      SmpltRod rod;
      float scratch[16];
      smplt_rodrigues(pose + 3 * j, rod, scratch);
      smplt_rodrigues_backward(rod, gL, g_pose + 3 * j);
    }
  }
}

// one thread: both chains, then the alignment gradient from gpre (pre[a][b] = alignment[b][a] * (a<3 ? scale : 1))
NM_KERNEL void k_smplt_chain_backward(const float* __restrict__ pose, const float* __restrict__ da_pose,
                                      const float* __restrict__ J, SmpltParents par, int nj,
                                      const float* __restrict__ gA_pose, const float* __restrict__ gA_da,
                                      const float* __restrict__ gpre, float scale, float* __restrict__ g_pose,
                                      float* __restrict__ gJ, float* __restrict__ g_alignment) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int k = 0; k < 3 * nj; ++k) gJ[k] = 0.f;
  smplt_chain_backward(pose, J, par, nj, gA_pose, g_pose, gJ);
  smplt_chain_backward(da_pose, J, par, nj, gA_da, (float*)0, gJ);
  if (g_alignment)
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) g_alignment[4 * b + a] = gpre[4 * a + b] * (a < 3 ? scale : 1.f);
}

// g v_shaped[v] = grest[v] + sum_j J_regressor[j,v] gJ[j]  (vertices2joints, models/smpl.py:363); in place on grest
NM_KERNEL void k_smplt_vshaped_backward(const float* __restrict__ Jreg, const float* __restrict__ gJ, int nv, int nj,
                                        float* __restrict__ grest) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  float acc[3] = {grest[3 * v], grest[3 * v + 1], grest[3 * v + 2]};
  for (int j = 0; j < nj; ++j) {
    float w = Jreg[(size_t)j * nv + v];
    if (w == 0.f) continue;
    for (int c = 0; c < 3; ++c) acc[c] += w * gJ[3 * j + c];
  }
  for (int c = 0; c < 3; ++c) grest[3 * v + c] = acc[c];
}

// g betas[l] = sum_i shapedirs[i,l] g v_shaped[i], i over nv*3 (blend_shapes, models/smpl.py:383): 64 rows per thread
NM_KERNEL void k_smplt_betas_backward(const float* __restrict__ shapedirs, const float* __restrict__ gvs, int n3, int nb,
                                      float* __restrict__ g_betas) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int i0 = t * 64;
  if (i0 >= n3) return;
  for (int l = 0; l < nb; ++l) {
    float acc = 0.f;
    for (int i = i0; i < i0 + 64 && i < n3; ++i) acc += shapedirs[(size_t)i * nb + l] * gvs[i];
    NM_ATOMIC_ADD(g_betas + l, acc);
  }
}
