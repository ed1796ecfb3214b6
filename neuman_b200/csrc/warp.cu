// Observation -> canonical warp: closest point on the posed SMPL mesh, barycentric blend of the
// three per-vertex 4x4 transforms, inverse, apply; canonical view directions by finite differences.
//
//   nm_mesh_set            per-frame inputs of warp_samples_to_canonical (verts, faces, T) + grid
//   nm_warp_to_canonical   <- utils/ray_utils.py:48-66 (igl.point_mesh_squared_distance :53,
//                             igl.barycentric_coordinates_tri :55, blend :56, inverse :57, apply :58,
//                             finite-difference directions :62-64)
//
// The reference does this stage on the CPU in float64 (libigl AABB tree) with a device->host->device
// round trip per batch (utils/render_utils.py:218-227).  Here a per-frame uniform grid stores, for
// every cell, the conservative list of triangles that can be the closest one for ANY point of the
// cell (sphere bounds), so a query is one short exact scan.  The arg-min runs in fp32; the winning
// triangle is then re-evaluated in float64 (closest point, barycentrics, blend, inverse, apply), which
// is what the reference's float64 chain produces before `.float()` (utils/render_utils.py:226).
#include <cub/device/device_scan.cuh>
#include <float.h>
#include <math.h>

#include "nm_internal.cuh"

#define GRID_MAX_DIM 64

struct GridDesc {
  float3 gmin;
  float cell, inv_cell;
  int3 dims;
  int ncell;
};

// ---------------------------------------------------------------------------------------------
__global__ void k_tri_prepare(const float* __restrict__ verts, const int32_t* __restrict__ faces, int nf,
                              float4* __restrict__ sphere, float* __restrict__ tri9) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  float v[9];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int vi = faces[3 * f + k];
    v[3 * k] = verts[3 * vi]; v[3 * k + 1] = verts[3 * vi + 1]; v[3 * k + 2] = verts[3 * vi + 2];
  }
  float cx = (v[0] + v[3] + v[6]) / 3.f, cy = (v[1] + v[4] + v[7]) / 3.f, cz = (v[2] + v[5] + v[8]) / 3.f;
  float r2 = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float dx = v[3 * k] - cx, dy = v[3 * k + 1] - cy, dz = v[3 * k + 2] - cz;
    r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz);
  }
  sphere[f] = make_float4(cx, cy, cz, sqrtf(r2) * 1.0001f + 1e-7f);
#pragma unroll
  for (int k = 0; k < 9; ++k) tri9[9 * f + k] = v[k];
}

#define CB_TILE 1024
// mode 0: count[cell] = #candidates; mode 1: fill lists (ascending face index)
__global__ void __launch_bounds__(128) k_cell_lists(GridDesc g, const float4* __restrict__ sphere, int nf, int mode,
                                                     int32_t* __restrict__ count, const int32_t* __restrict__ start,
                                                     int32_t* __restrict__ lists) {
  __shared__ float4 ss[CB_TILE];
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = c < g.ncell;
  int cz = c / (g.dims.x * g.dims.y), rem = c - cz * g.dims.x * g.dims.y, cy = rem / g.dims.x, cx = rem - cy * g.dims.x;
  float qx = g.gmin.x + (cx + 0.5f) * g.cell, qy = g.gmin.y + (cy + 0.5f) * g.cell, qz = g.gmin.z + (cz + 0.5f) * g.cell;
  const float h = g.cell * 0.8660255f * 1.001f;     // half diagonal
  // pass A: upper bound U on the distance from the cell centre to the mesh (centroids lie on it)
  float U = FLT_MAX;
  for (int base = 0; base < nf; base += CB_TILE) {
    int cnt = min(CB_TILE, nf - base);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) ss[j] = sphere[base + j];
    __syncthreads();
    for (int j = 0; j < cnt; ++j) {
      float dx = ss[j].x - qx, dy = ss[j].y - qy, dz = ss[j].z - qz;
      U = fminf(U, dx * dx + dy * dy + dz * dz);
    }
  }
  U = sqrtf(U);
  const float lim = (U + 2.f * h) * 1.0001f + 1e-6f;
  // pass B: a triangle can be the closest for some point of the cell only if |q-c|-r <= U + 2h
  int n = 0;
  int32_t* out = (mode == 1 && live) ? lists + start[c] : nullptr;
  for (int base = 0; base < nf; base += CB_TILE) {
    int cnt = min(CB_TILE, nf - base);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) ss[j] = sphere[base + j];
    __syncthreads();
    for (int j = 0; j < cnt; ++j) {
      float dx = ss[j].x - qx, dy = ss[j].y - qy, dz = ss[j].z - qz;
      float d = sqrtf(dx * dx + dy * dy + dz * dz) - ss[j].w;
      if (d <= lim) {
        if (out) out[n] = base + j;
        ++n;
      }
    }
  }
  if (live && mode == 0) count[c] = n;
}

// ---------------------------------------------------------------------------------------------
template <typename T>
struct V3 { T x, y, z; };
template <typename T> __device__ __forceinline__ V3<T> sub(V3<T> a, V3<T> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> __device__ __forceinline__ T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> __device__ __forceinline__ V3<T> madd(V3<T> a, V3<T> d, T t) { return {a.x + d.x * t, a.y + d.y * t, a.z + d.z * t}; }
template <typename T> __device__ __forceinline__ V3<T> cross(V3<T> a, V3<T> b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Exact closest point on triangle (a,b,c) to p: vertex / edge / face Voronoi regions (Ericson 5.1.5)
template <typename T>
__device__ __forceinline__ V3<T> closest_on_tri(V3<T> p, V3<T> a, V3<T> b, V3<T> c) {
  V3<T> ab = sub(b, a), ac = sub(c, a), ap = sub(p, a);
  T d1 = dot(ab, ap), d2 = dot(ac, ap);
  if (d1 <= 0 && d2 <= 0) return a;
  V3<T> bp = sub(p, b);
  T d3 = dot(ab, bp), d4 = dot(ac, bp);
  if (d3 >= 0 && d4 <= d3) return b;
  T vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) return madd(a, ab, d1 / (d1 - d3));
  V3<T> cp = sub(p, c);
  T d5 = dot(ab, cp), d6 = dot(ac, cp);
  if (d6 >= 0 && d5 <= d6) return c;
  T vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) return madd(a, ac, d2 / (d2 - d6));
  T va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) return madd(b, sub(c, b), (d4 - d3) / ((d4 - d3) + (d5 - d6)));
  T den = T(1) / (va + vb + vc);
  return madd(madd(a, ab, vb * den), ac, vc * den);
}

__device__ __forceinline__ void test_tri(int f, V3<float> p, const float4* __restrict__ sphere,
                                         const float* __restrict__ tri9, float& best, int& best_f) {
  float4 s = __ldg(sphere + f);
  float dx = s.x - p.x, dy = s.y - p.y, dz = s.z - p.z;
  float lb = sqrtf(dx * dx + dy * dy + dz * dz) - s.w;
  if (lb > 0.f && lb * lb > best) return;
  const float* t = tri9 + 9 * (size_t)f;
  V3<float> a{__ldg(t), __ldg(t + 1), __ldg(t + 2)}, b{__ldg(t + 3), __ldg(t + 4), __ldg(t + 5)},
      c{__ldg(t + 6), __ldg(t + 7), __ldg(t + 8)};
  V3<float> q = closest_on_tri<float>(p, a, b, c);
  V3<float> e = sub(q, p);
  float d2 = dot(e, e);
  if (d2 < best) { best = d2; best_f = f; }
}

// 4x4 inverse (general, cofactor expansion) in double; returns false if singular
__host__ __device__ __forceinline__ bool inv4(const double* m, double* o) {
  double s0 = m[0] * m[5] - m[4] * m[1], s1 = m[0] * m[6] - m[4] * m[2], s2 = m[0] * m[7] - m[4] * m[3];
  double s3 = m[1] * m[6] - m[5] * m[2], s4 = m[1] * m[7] - m[5] * m[3], s5 = m[2] * m[7] - m[6] * m[3];
  double c5 = m[10] * m[15] - m[14] * m[11], c4 = m[9] * m[15] - m[13] * m[11], c3 = m[9] * m[14] - m[13] * m[10];
  double c2 = m[8] * m[15] - m[12] * m[11], c1 = m[8] * m[14] - m[12] * m[10], c0 = m[8] * m[13] - m[12] * m[9];
  double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
  if (det == 0.0) return false;
  double id = 1.0 / det;
  o[0] = (m[5] * c5 - m[6] * c4 + m[7] * c3) * id;
  o[1] = (-m[1] * c5 + m[2] * c4 - m[3] * c3) * id;
  o[2] = (m[13] * s5 - m[14] * s4 + m[15] * s3) * id;
  o[3] = (-m[9] * s5 + m[10] * s4 - m[11] * s3) * id;
  o[4] = (-m[4] * c5 + m[6] * c2 - m[7] * c1) * id;
  o[5] = (m[0] * c5 - m[2] * c2 + m[3] * c1) * id;
  o[6] = (-m[12] * s5 + m[14] * s2 - m[15] * s1) * id;
  o[7] = (m[8] * s5 - m[10] * s2 + m[11] * s1) * id;
  o[8] = (m[4] * c4 - m[5] * c2 + m[7] * c0) * id;
  o[9] = (-m[0] * c4 + m[1] * c2 - m[3] * c0) * id;
  o[10] = (m[12] * s4 - m[13] * s2 + m[15] * s0) * id;
  o[11] = (-m[8] * s4 + m[9] * s2 - m[11] * s0) * id;
  o[12] = (-m[4] * c3 + m[5] * c1 - m[6] * c0) * id;
  o[13] = (m[0] * c3 - m[1] * c1 + m[2] * c0) * id;
  o[14] = (-m[12] * s3 + m[13] * s1 - m[14] * s0) * id;
  o[15] = (m[8] * s3 - m[9] * s1 + m[10] * s0) * id;
  return true;
}

__global__ void __launch_bounds__(128) k_warp_points(GridDesc g, const int32_t* __restrict__ cell_start,
                                                      const int32_t* __restrict__ cell_tris,
                                                      const float4* __restrict__ sphere, const float* __restrict__ tri9,
                                                      const float* __restrict__ verts, const int32_t* __restrict__ faces,
                                                      int nf, const double* __restrict__ T, const float* __restrict__ pts,
                                                      long long n, double* __restrict__ can64,
                                                      float* __restrict__ closest_out, int32_t* __restrict__ face_out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  V3<float> p{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  float best = FLT_MAX;
  int best_f = -1;
  float fx = (p.x - g.gmin.x) * g.inv_cell, fy = (p.y - g.gmin.y) * g.inv_cell, fz = (p.z - g.gmin.z) * g.inv_cell;
  bool inside = fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)g.dims.x && fy < (float)g.dims.y && fz < (float)g.dims.z;
  if (inside) {
    int c = ((int)fz * g.dims.y + (int)fy) * g.dims.x + (int)fx;
    int s = cell_start[c], e = cell_start[c + 1];
    for (int k = s; k < e; ++k) test_tri(__ldg(cell_tris + k), p, sphere, tri9, best, best_f);
  } else {
    for (int f = 0; f < nf; ++f) test_tri(f, p, sphere, tri9, best, best_f);    // exact fallback
  }
  // ---- float64 re-evaluation on the winning triangle (utils/ray_utils.py:53-58) ----
  int i0 = faces[3 * best_f], i1 = faces[3 * best_f + 1], i2 = faces[3 * best_f + 2];
  V3<double> P{(double)p.x, (double)p.y, (double)p.z};
  V3<double> A{(double)verts[3 * i0], (double)verts[3 * i0 + 1], (double)verts[3 * i0 + 2]};
  V3<double> B{(double)verts[3 * i1], (double)verts[3 * i1 + 1], (double)verts[3 * i1 + 2]};
  V3<double> C{(double)verts[3 * i2], (double)verts[3 * i2 + 1], (double)verts[3 * i2 + 2]};
  V3<double> Q = closest_on_tri<double>(P, A, B, C);
  // barycentric coordinates of Q w.r.t. (A,B,C): signed sub-areas over the area
  V3<double> nrm = cross(sub(B, A), sub(C, A));
  double nn = dot(nrm, nrm);
  double la = dot(nrm, cross(sub(C, B), sub(Q, B))) / nn;
  double lb = dot(nrm, cross(sub(A, C), sub(Q, C))) / nn;
  double lc = 1.0 - la - lb;
  double M[16], Mi[16];
  const double* Ta = T + 16 * (size_t)i0;
  const double* Tb = T + 16 * (size_t)i1;
  const double* Tc = T + 16 * (size_t)i2;
#pragma unroll
  for (int k = 0; k < 16; ++k) M[k] = Ta[k] * la + Tb[k] * lb + Tc[k] * lc;          // (:56)
  inv4(M, Mi);                                                                      // (:57)
  double cx = Mi[0] * P.x + Mi[1] * P.y + Mi[2] * P.z + Mi[3];                       // (:58)
  double cy = Mi[4] * P.x + Mi[5] * P.y + Mi[6] * P.z + Mi[7];
  double cz = Mi[8] * P.x + Mi[9] * P.y + Mi[10] * P.z + Mi[11];
  can64[3 * i] = cx; can64[3 * i + 1] = cy; can64[3 * i + 2] = cz;
  if (closest_out) { closest_out[3 * i] = (float)Q.x; closest_out[3 * i + 1] = (float)Q.y; closest_out[3 * i + 2] = (float)Q.z; }
  if (face_out) face_out[i] = best_f;
}

// can_dirs: normalised forward difference along the ray, last one duplicated (:62-64), in float64
__global__ void __launch_bounds__(256) k_warp_dirs(const double* __restrict__ can64, long long R, int S,
                                                    float* __restrict__ can_pts, float* __restrict__ can_dirs) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * S) return;
  int s = (int)(i % S);
  double x = can64[3 * i], y = can64[3 * i + 1], z = can64[3 * i + 2];
  can_pts[3 * i] = (float)x; can_pts[3 * i + 1] = (float)y; can_pts[3 * i + 2] = (float)z;
  if (!can_dirs) return;
  long long a = (s < S - 1) ? i : i - 1;      // difference (a+1) - a
  if (S == 1) { can_dirs[3 * i] = can_dirs[3 * i + 1] = can_dirs[3 * i + 2] = NAN; return; }
  double dx = can64[3 * (a + 1)] - can64[3 * a], dy = can64[3 * (a + 1) + 1] - can64[3 * a + 1],
         dz = can64[3 * (a + 1) + 2] - can64[3 * a + 2];
  double nrm = sqrt(dx * dx + dy * dy + dz * dz);
  can_dirs[3 * i] = (float)(dx / nrm); can_dirs[3 * i + 1] = (float)(dy / nrm); can_dirs[3 * i + 2] = (float)(dz / nrm);
}

// ---------------------------------------------------------------------------------------------
template <typename T>
static int ensure(nm_ctx* ctx, T** p, size_t* cap, size_t need) {
  if (need <= *cap && *p) return NM_OK;
  if (*p) { NM_CHECK_CUDA(ctx, cudaDeviceSynchronize()); NM_CHECK_CUDA(ctx, cudaFree(*p)); *p = nullptr; }
  size_t want = need + (need >> 2) + 16;
  NM_CHECK_CUDA(ctx, cudaMalloc(p, want * sizeof(T)));
  *cap = want;
  return NM_OK;
}

static GridDesc grid_of(const NmMesh& m) {
  GridDesc g;
  g.gmin = m.grid_min; g.cell = m.cell; g.inv_cell = 1.f / m.cell; g.dims = m.dims;
  g.ncell = m.dims.x * m.dims.y * m.dims.z;
  return g;
}

extern "C" int nm_mesh_set(nm_ctx* ctx, int actor, const float* verts, int32_t n_verts, const int32_t* faces,
                           int32_t n_faces, const double* T, int32_t n_T, int32_t on_device, void* stream) {
  if (!ctx) return NM_ERR_INVALID;
  if (actor < 0 || actor >= NM_MAX_ACTORS || !verts || !faces || !T || n_verts <= 0 || n_faces <= 0 || n_T < n_verts)
    NM_FAIL(ctx, NM_ERR_INVALID, "nm_mesh_set: bad argument (need n_T >= n_verts)");
  cudaStream_t st = (cudaStream_t)stream;
  NmMesh& m = ctx->meshes[actor];
  int rc;
  if ((rc = ensure(ctx, &m.verts, &m.cap_verts, (size_t)n_verts * 3))) return rc;
  if ((rc = ensure(ctx, &m.T, &m.cap_T, (size_t)n_T * 16))) return rc;
  {
    size_t capf = m.cap_faces;
    if ((rc = ensure(ctx, &m.faces, &capf, (size_t)n_faces * 3))) return rc;
    if (capf != m.cap_faces || !m.tri_sphere) {
      // tri_sphere holds [F] float4 followed by the packed triangles [F][9] floats
      if (m.tri_sphere) { NM_CHECK_CUDA(ctx, cudaFree(m.tri_sphere)); m.tri_sphere = nullptr; }
      NM_CHECK_CUDA(ctx, cudaMalloc(&m.tri_sphere, capf / 3 * (sizeof(float4) + 9 * sizeof(float)) + 64));
      m.cap_faces = capf;
    }
  }
  cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  NM_CHECK_CUDA(ctx, cudaMemcpyAsync(m.verts, verts, (size_t)n_verts * 3 * sizeof(float), kind, st));
  NM_CHECK_CUDA(ctx, cudaMemcpyAsync(m.faces, faces, (size_t)n_faces * 3 * sizeof(int32_t), kind, st));
  NM_CHECK_CUDA(ctx, cudaMemcpyAsync(m.T, T, (size_t)n_T * 16 * sizeof(double), kind, st));
  m.n_verts = n_verts; m.n_faces = n_faces; m.n_T = n_T;
  // bounding box on the host (82 KB; once per frame)
  std::vector<float> hv((size_t)n_verts * 3);
  if (on_device) {
    NM_CHECK_CUDA(ctx, cudaMemcpyAsync(hv.data(), verts, hv.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    NM_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
  } else {
    memcpy(hv.data(), verts, hv.size() * sizeof(float));
  }
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int v = 0; v < n_verts; ++v)
    for (int c = 0; c < 3; ++c) { lo[c] = fminf(lo[c], hv[3 * v + c]); hi[c] = fmaxf(hi[c], hv[3 * v + c]); }
  // grid covers the bbox grown by 35% of its longest side (every sample of a geometry-guided ray
  // lies within geo_threshold of the bbox; points outside take the exact brute-force path)
  float ext = fmaxf(hi[0] - lo[0], fmaxf(hi[1] - lo[1], hi[2] - lo[2]));
  float pad = 0.35f * ext;
  m.cell = (ext + 2 * pad) / GRID_MAX_DIM;
  m.grid_min = make_float3(lo[0] - pad, lo[1] - pad, lo[2] - pad);
  m.dims = make_int3(max(1, (int)ceilf((hi[0] - lo[0] + 2 * pad) / m.cell)), max(1, (int)ceilf((hi[1] - lo[1] + 2 * pad) / m.cell)),
                     max(1, (int)ceilf((hi[2] - lo[2] + 2 * pad) / m.cell)));
  GridDesc g = grid_of(m);
  float* tri9 = reinterpret_cast<float*>(m.tri_sphere + n_faces);
  k_tri_prepare<<<(n_faces + 127) / 128, 128, 0, st>>>(m.verts, m.faces, n_faces, m.tri_sphere, tri9);
  NM_CHECK_LAUNCH(ctx);
  if ((rc = ensure(ctx, &m.cell_start, &m.cap_cells, (size_t)2 * (g.ncell + 1)))) return rc;
  int32_t* counts = m.cell_start + (g.ncell + 1);
  NM_CHECK_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)(g.ncell + 1) * sizeof(int32_t), st));
  k_cell_lists<<<(g.ncell + 127) / 128, 128, 0, st>>>(g, m.tri_sphere, n_faces, 0, counts, nullptr, nullptr);
  NM_CHECK_LAUNCH(ctx);
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, m.cell_start, g.ncell + 1, st);
  char* tmp = nullptr;
  if ((rc = nm_impl_workspace(ctx, tmp_bytes, &tmp))) return rc;
  NM_CHECK_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, m.cell_start, g.ncell + 1, st));
  NM_LAUNCHED(ctx);
  NM_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_counter, m.cell_start + g.ncell, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  NM_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
  m.n_refs = ctx->h_counter[0];
  if ((rc = ensure(ctx, &m.cell_tris, &m.cap_refs, (size_t)m.n_refs + 1))) return rc;
  k_cell_lists<<<(g.ncell + 127) / 128, 128, 0, st>>>(g, m.tri_sphere, n_faces, 1, nullptr, m.cell_start, m.cell_tris);
  NM_CHECK_LAUNCH(ctx);
  m.set = true;
  return NM_OK;
}

extern "C" int nm_warp_to_canonical(nm_ctx* ctx, int actor, const float* pts, int64_t R, int32_t S, float* can_pts,
                                    float* can_dirs, float* closest, int32_t* face_id, void* stream) {
  if (!ctx) return NM_ERR_INVALID;
  if (R == 0) return NM_OK;
  if (actor < 0 || actor >= NM_MAX_ACTORS || !ctx->meshes[actor].set) NM_FAIL(ctx, NM_ERR_STATE, "nm_warp_to_canonical: mesh not set");
  if (!pts || !can_pts || R < 0 || S <= 0) NM_FAIL(ctx, NM_ERR_INVALID, "nm_warp_to_canonical: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  NmMesh& m = ctx->meshes[actor];
  long long n = (long long)R * S;
  // float64 canonical points: private scratch sized on demand (the frame drivers own ctx->ws)
  if ((size_t)n * 3 > ctx->can64_cap) {
    if (ctx->can64) { NM_CHECK_CUDA(ctx, cudaDeviceSynchronize()); NM_CHECK_CUDA(ctx, cudaFree(ctx->can64)); ctx->can64 = nullptr; }
    size_t want = (size_t)n * 3 + ((size_t)n * 3 >> 3);
    NM_CHECK_CUDA(ctx, cudaMalloc(&ctx->can64, want * sizeof(double)));
    ctx->can64_cap = want;
  }
  double* can64 = ctx->can64;
  GridDesc g = grid_of(m);
  const float* tri9 = reinterpret_cast<const float*>(m.tri_sphere + m.n_faces);
  k_warp_points<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(g, m.cell_start, m.cell_tris, m.tri_sphere, tri9, m.verts,
                                                               m.faces, m.n_faces, m.T, pts, n, can64, closest, face_id);
  NM_CHECK_LAUNCH(ctx);
  k_warp_dirs<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(can64, R, S, can_pts, can_dirs);
  NM_CHECK_LAUNCH(ctx);
  return NM_OK;
}
