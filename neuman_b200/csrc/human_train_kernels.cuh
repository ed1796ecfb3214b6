// Kernels of the human trainer's differentiable observation->canonical map:
//
//   forward   utils/ray_utils.py:69-93 (warp_samples_to_canonical_diff: barycentric coordinates of the closest point from
//             cross products, blend of the three per-vertex transforms, inverse) and
//             trainers/human_nerf_trainer.py:263-276 (apply to the samples, add the offset, finite-difference directions)
//   backward  what torch autograd computes for those lines inside loss.backward() (trainers/human_nerf_trainer.py:205):
//             gradients to the per-vertex transforms T (-> SMPL pose / shape / alignment), to the posed vertices (through
//             the barycentric coordinates) and to the offset.
//
// The closest face / closest point of every sample come from nm_signed_distance (the reference's igl.signed_distance,
// :70) and are constants of the step, as in the reference (numpy arrays re-wrapped with torch.from_numpy).
//
// This header is written in a subset of CUDA C++ (one thread per element, atomics for the scatter, no shared memory or
// warp intrinsics) so that tests/emu/ can compile the SAME kernel bodies for the host with a few macros and check the
// algebra against torch autograd on the CPU (this container has no GPU); libneuman_b200.so gets them through
// human_train.cu.
#pragma once
#include "train_common_kernels.cuh"

NM_DEV void wd_cross(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
NM_DEV float wd_dot(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Everything the forward and the backward need about one sample's closest triangle (utils/ray_utils.py:72-88).
struct WdTri {
  int vi[3];
  float e1[3], e2[3], e[3], ep[3], f[3], fp[3];   // v0v1, v0v2, v1v2, v2v0, v1p, v2p
  float N[3], C1[3], C2[3], D, u, v, w;
};

NM_DEV void wd_triangle(const float* __restrict__ verts, const int* __restrict__ faces, int face,
                        const double* __restrict__ closest, WdTri& t) {
  float a[3], b[3], c[3], q[3];
  for (int k = 0; k < 3; ++k) t.vi[k] = faces[3 * face + k];
  for (int k = 0; k < 3; ++k) {
    a[k] = verts[3 * t.vi[0] + k]; b[k] = verts[3 * t.vi[1] + k]; c[k] = verts[3 * t.vi[2] + k];
    q[k] = (float)closest[k];                                              // torch.from_numpy(closest).float() (:74)
  }
  for (int k = 0; k < 3; ++k) {
    t.e1[k] = b[k] - a[k]; t.e2[k] = c[k] - a[k]; t.e[k] = c[k] - b[k]; t.ep[k] = a[k] - c[k];
    t.f[k] = q[k] - b[k]; t.fp[k] = q[k] - c[k];
  }
  wd_cross(t.e1, t.e2, t.N);
  t.D = wd_dot(t.N, t.N);
  wd_cross(t.e, t.f, t.C1);
  t.u = wd_dot(t.N, t.C1) / t.D;
  wd_cross(t.ep, t.fp, t.C2);
  t.v = wd_dot(t.N, t.C2) / t.D;
  t.w = 1.f - t.u - t.v;
}

// T_interp = sum_k bary_k T[v_k] (float32, :90) and its inverse (:91)
NM_DEV bool wd_blend_inverse(const float* __restrict__ T, const WdTri& t, float* Ti, double* Tinv) {
  const float* T0 = T + (size_t)16 * t.vi[0];
  const float* T1 = T + (size_t)16 * t.vi[1];
  const float* T2 = T + (size_t)16 * t.vi[2];
  double m[16];
  for (int k = 0; k < 16; ++k) {
    Ti[k] = (T0[k] * t.u + T1[k] * t.v) + T2[k] * t.w;
    m[k] = (double)Ti[k];
  }
  return wd_inv4(m, Tinv);
}

// ---- forward: one thread per sample ----------------------------------------------------------------------------------
// f_id [n] closest face, closest [n,3] f64 (nm_signed_distance), verts [V,3], faces [F,3], T [V,16] f32.
// Tinv_out [n,16] (may be null) = T_interp_inv of :91; can_pts [n,3] (may be null) = (T_interp_inv @ [p;1])[:3] + offset
// (trainers/human_nerf_trainer.py:272-273), pts [n,3] and offset [n,3] (may be null) read only when can_pts is wanted.
NM_KERNEL void k_wd_forward(const int* __restrict__ f_id, const double* __restrict__ closest,
                            const float* __restrict__ verts, const int* __restrict__ faces,
                            const float* __restrict__ T, const float* __restrict__ pts,
                            const float* __restrict__ offset, long long n, float* __restrict__ Tinv_out,
                            float* __restrict__ can_pts) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  WdTri t;
  wd_triangle(verts, faces, f_id[i], closest + 3 * i, t);
  float Ti[16];
  double inv[16];
  if (!wd_blend_inverse(T, t, Ti, inv))
    for (int k = 0; k < 16; ++k) inv[k] = (double)NAN;                    // singular blend: NaN, which the trainer's loss check catches
  float fi[16];
  for (int k = 0; k < 16; ++k) fi[k] = (float)inv[k];
  if (Tinv_out)
    for (int k = 0; k < 16; ++k) Tinv_out[16 * i + k] = fi[k];
  if (can_pts) {
    float p0 = pts[3 * i], p1 = pts[3 * i + 1], p2 = pts[3 * i + 2];
    for (int a = 0; a < 3; ++a) {
      float r = ((fi[4 * a] * p0 + fi[4 * a + 1] * p1) + fi[4 * a + 2] * p2) + fi[4 * a + 3];
      can_pts[3 * i + a] = offset ? r + offset[3 * i + a] : r;
    }
  }
}

// can_dirs (trainers/human_nerf_trainer.py:274-276): differences of consecutive canonical points along the ray, the last
// sample repeats the previous direction, unit length.  One thread per sample; S >= 2.
NM_DEV void wd_raw_dir(const float* __restrict__ cp, long long ray, int S, int s, float* d) {
  int s0 = s < S - 1 ? s : S - 2;
  const float* p = cp + 3 * (ray * S + s0);
  for (int k = 0; k < 3; ++k) d[k] = p[3 + k] - p[k];
}

NM_KERNEL void k_wd_dirs(const float* __restrict__ can_pts, long long R, int S, float* __restrict__ can_dirs) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * S) return;
  long long ray = i / S;
  int s = (int)(i - ray * S);
  float d[3];
  wd_raw_dir(can_pts, ray, S, s, d);
  float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  for (int k = 0; k < 3; ++k) can_dirs[3 * i + k] = d[k] / nrm;
}

// ---- backward of the directions: total dL/d can_pts ------------------------------------------------------------------
// g_pts_in [R,S,3] (may be null) direct gradient of can_pts (from the network's position input), g_dirs [R,S,3] (may be
// null) gradient of can_dirs; g_total [R,S,3] out = g_pts_in + the directions' contribution.
NM_DEV void wd_dir_grad(const float* __restrict__ cp, const float* __restrict__ g_dirs, long long ray, int S, int s,
                        float* g) {
  // gradient w.r.t. the raw difference d_s = cp[s+1] - cp[s], s in [0, S-2]: the normalisation's Jacobian applied to
  // g_dirs[s] (+ g_dirs[S-1] for s = S-2, whose direction is the copy)
  float d[3];
  wd_raw_dir(cp, ray, S, s, d);
  float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  float gs[3];
  const float* g0 = g_dirs + 3 * (ray * S + s);
  for (int k = 0; k < 3; ++k) gs[k] = g0[k] + (s == S - 2 ? g0[3 + k] : 0.f);
  float inv = 1.f / nrm;
  float nd = (d[0] * gs[0] + d[1] * gs[1] + d[2] * gs[2]) * inv * inv;       // (n . g) / |d|
  for (int k = 0; k < 3; ++k) g[k] = (gs[k] - d[k] * nd) * inv;
}

NM_KERNEL void k_wd_dirs_backward(const float* __restrict__ can_pts, const float* __restrict__ g_pts_in,
                                  const float* __restrict__ g_dirs, long long R, int S, float* __restrict__ g_total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * S) return;
  long long ray = i / S;
  int s = (int)(i - ray * S);
  float acc[3] = {0.f, 0.f, 0.f};
  if (g_pts_in)
    for (int k = 0; k < 3; ++k) acc[k] = g_pts_in[3 * i + k];
  if (g_dirs) {
    float g[3];
    if (s >= 1) {                                   // d_{s-1} = cp[s] - cp[s-1]
      wd_dir_grad(can_pts, g_dirs, ray, S, s - 1, g);
      for (int k = 0; k < 3; ++k) acc[k] += g[k];
    }
    if (s <= S - 2) {                               // d_s = cp[s+1] - cp[s]
      wd_dir_grad(can_pts, g_dirs, ray, S, s, g);
      for (int k = 0; k < 3; ++k) acc[k] -= g[k];
    }
  }
  for (int k = 0; k < 3; ++k) g_total[3 * i + k] = acc[k];
}

// ---- backward of the blend / inverse / barycentrics: one thread per sample, atomics into the per-vertex gradients ----
// g_Tinv [n,16] (may be null): dL/d T_interp_inv (the drop-in form, where the trainer applies the matrices itself);
// g_can [n,3] (may be null): total dL/d can_pts of the fused form, pts [n,3] then required.
// g_T [V,16], g_verts [V,3]: accumulated (caller zeroes them).
NM_KERNEL void k_wd_backward(const int* __restrict__ f_id, const double* __restrict__ closest,
                             const float* __restrict__ verts, const int* __restrict__ faces,
                             const float* __restrict__ T, const float* __restrict__ pts,
                             const float* __restrict__ g_Tinv, const float* __restrict__ g_can, long long n,
                             float* __restrict__ g_T, float* __restrict__ g_verts) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  WdTri t;
  wd_triangle(verts, faces, f_id[i], closest + 3 * i, t);
  float Ti[16];
  double inv[16];
  if (!wd_blend_inverse(T, t, Ti, inv)) return;
  float Y[16], G[16];
  for (int k = 0; k < 16; ++k) { Y[k] = (float)inv[k]; G[k] = g_Tinv ? g_Tinv[16 * i + k] : 0.f; }
  if (g_can) {
    float ph[4] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], 1.f};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 4; ++b) G[4 * a + b] += g_can[3 * i + a] * ph[b];
  }
  // dL/dT_interp = -Y^T G Y^T
  float M[16], gTi[16];
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      float acc = 0.f;
      for (int k = 0; k < 4; ++k) acc += Y[4 * k + a] * G[4 * k + b];        // (Y^T G)[a][b]
      M[4 * a + b] = acc;
    }
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      float acc = 0.f;
      for (int k = 0; k < 4; ++k) acc += M[4 * a + k] * Y[4 * b + k];        // (M Y^T)[a][b]
      gTi[4 * a + b] = -acc;
    }
  const float bary[3] = {t.u, t.v, t.w};
  float gb[3];
  for (int j = 0; j < 3; ++j) {
    const float* Tj = T + (size_t)16 * t.vi[j];
    float acc = 0.f;
    for (int k = 0; k < 16; ++k) {
      acc += Tj[k] * gTi[k];
      if (g_T) NM_ATOMIC_ADD(g_T + (size_t)16 * t.vi[j] + k, bary[j] * gTi[k]);
    }
    gb[j] = acc;
  }
  if (!g_verts) return;
  // bary = (u, v, 1 - u - v); u = N.C1 / D, v = N.C2 / D  (utils/ray_utils.py:82-88)
  const float gu = gb[0] - gb[2], gv = gb[1] - gb[2], iD = 1.f / t.D;
  float gN[3], gC1[3], gC2[3];
  for (int k = 0; k < 3; ++k) {
    gN[k] = (gu * (t.C1[k] - 2.f * t.u * t.N[k]) + gv * (t.C2[k] - 2.f * t.v * t.N[k])) * iD;
    gC1[k] = gu * t.N[k] * iD;
    gC2[k] = gv * t.N[k] * iD;
  }
  float ge[3], gf[3], gep[3], gfp[3], ge1[3], ge2[3];
  wd_cross(t.f, gC1, ge);      // C1 = e x f
  wd_cross(gC1, t.e, gf);
  wd_cross(t.fp, gC2, gep);    // C2 = e' x f'
  wd_cross(gC2, t.ep, gfp);
  wd_cross(t.e2, gN, ge1);     // N = e1 x e2
  wd_cross(gN, t.e1, ge2);
  for (int k = 0; k < 3; ++k) {
    NM_ATOMIC_ADD(g_verts + 3 * t.vi[0] + k, -ge1[k] - ge2[k] + gep[k]);
    NM_ATOMIC_ADD(g_verts + 3 * t.vi[1] + k, ge1[k] - ge[k] - gf[k]);
    NM_ATOMIC_ADD(g_verts + 3 * t.vi[2] + k, ge2[k] + ge[k] - gep[k] - gfp[k]);
  }
}
