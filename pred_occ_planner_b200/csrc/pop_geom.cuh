// pop_geom.cuh — device-side grid geometry and cell access.
//
// Index math restates plan_env/include/plan_env/map.h:153-215 of the reference in fp32 with
// explicit round-to-nearest intrinsics (never contracted to FMA, division kept as a division)
// so voxel indices are bit-identical to the CPU path.
#pragma once
#include "pop_internal.h"

__device__ __forceinline__ bool geom_in_range_f(const PopGeom& g, float x, float y, float z) {
  return x > -g.rx && x < g.rx && y > -g.ry && y < g.ry && z > -g.rz && z < g.rz;
}
__device__ __forceinline__ bool geom_in_range_i(const PopGeom& g, int x, int y, int z) {
  return x >= 0 && x < g.nx && y >= 0 && y < g.ny && z >= 0 && z < g.nz;
}
// getVoxelRelIndex: (int)((p + range) / res), truncation toward zero
__device__ __forceinline__ int geom_axis_index(float p, float range, float res) {
  return static_cast<int>(__fdiv_rn(__fadd_rn(p, range), res));
}
__device__ __forceinline__ int geom_voxel_index(const PopGeom& g, float x, float y, float z) {
  const int ix = geom_axis_index(x, g.rx, g.res);
  const int iy = geom_axis_index(y, g.ry, g.res);
  const int iz = geom_axis_index(z, g.rz, g.res);
  return iz * g.nx * g.ny + iy * g.nx + ix;
}
// getVoxelPosition: voxel corner in the world frame, (i*res - range) + pose
__device__ __forceinline__ float geom_axis_position(int i, float range, float res, float pose) {
  return __fadd_rn(__fsub_rn(__fmul_rn(static_cast<float>(i), res), range), pose);
}

// ---- ground-truth obstacle ownership (fake_particle_risk_voxel.cpp:127-152) --------------------------------------
// fp32 3-vector arithmetic in the order Eigen 3.3's scalar code path evaluates the reference's Vector3f / Quaternionf
// / Hyperplane<float, 3> expressions: reductions of a 3-vector pair c0 + (c1 + c2) (Core/Redux.h), cross is the plain
// formula, q * v = v + w (2 vec x v) + vec x (2 vec x v) (Geometry/Quaternion.h).  POP_F* are the round-to-nearest
// intrinsics on the device and the plain operators in host code (no FMA contraction there: -ffp-contract=off).
#if defined(__CUDA_ARCH__) || !defined(__CUDACC__)
#define POP_FADD(a, b) __fadd_rn(a, b)
#define POP_FSUB(a, b) __fsub_rn(a, b)
#define POP_FMUL(a, b) __fmul_rn(a, b)
#define POP_FDIV(a, b) __fdiv_rn(a, b)
#define POP_FSQRT(a) __fsqrt_rn(a)
#else
#define POP_FADD(a, b) ((a) + (b))
#define POP_FSUB(a, b) ((a) - (b))
#define POP_FMUL(a, b) ((a) * (b))
#define POP_FDIV(a, b) ((a) / (b))
#define POP_FSQRT(a) sqrtf(a)
#endif
__host__ __device__ __forceinline__ float geom_dot3(const float* a, const float* b) {
  return POP_FADD(POP_FMUL(a[0], b[0]), POP_FADD(POP_FMUL(a[1], b[1]), POP_FMUL(a[2], b[2])));
}
__host__ __device__ __forceinline__ void geom_cross3(const float* a, const float* b, float* o) {
  o[0] = POP_FSUB(POP_FMUL(a[1], b[2]), POP_FMUL(a[2], b[1]));
  o[1] = POP_FSUB(POP_FMUL(a[2], b[0]), POP_FMUL(a[0], b[2]));
  o[2] = POP_FSUB(POP_FMUL(a[0], b[1]), POP_FMUL(a[1], b[0]));
}
// q = (w, x, y, z)
__host__ __device__ __forceinline__ void geom_quat_rotate(const float* q, const float* v, float* o) {
  float uv[3], c[3];
  geom_cross3(q + 1, v, uv);
  for (int k = 0; k < 3; k++) uv[k] = POP_FADD(uv[k], uv[k]);
  geom_cross3(q + 1, uv, c);
  for (int k = 0; k < 3; k++) o[k] = POP_FADD(POP_FADD(v[k], POP_FMUL(q[0], uv[k])), c[k]);
}
// Hyperplane::Through(p0, p0 + q*(0,1,0), p0 + q*(1,0,0)) (fake_particle_risk_voxel.cpp:141-143): normal and offset
__host__ __device__ __forceinline__ void geom_ring_plane(float cx, float cy, float cz, const float* q, float* n, float* off) {
  const float pc[3] = {cx, cy, cz}, ey[3] = {0.f, 1.f, 0.f}, ex[3] = {1.f, 0.f, 0.f};
  float r1[3], r2[3], v0[3], v1[3];
  geom_quat_rotate(q, ey, r1);
  geom_quat_rotate(q, ex, r2);
  for (int k = 0; k < 3; k++) {
    v0[k] = POP_FSUB(POP_FADD(pc[k], r2[k]), pc[k]);   // p2 - p0
    v1[k] = POP_FSUB(POP_FADD(pc[k], r1[k]), pc[k]);   // p1 - p0
  }
  geom_cross3(v0, v1, n);
  const float nn = POP_FSQRT(geom_dot3(n, n));
  for (int k = 0; k < 3; k++) n[k] = POP_FDIV(n[k], nn);
  *off = -geom_dot3(pc, n);
}
// pop_cylinder -> what the voxel update reads: the casts of fake_particle_risk_voxel.cpp:133,137,140-141 and, for a
// ring, its plane
__host__ __device__ __forceinline__ void geom_cyl_prepare(const pop_cylinder& c, PopCylDev* o) {
  o->x = static_cast<float>(c.x);
  o->y = static_cast<float>(c.y);
  o->z = static_cast<float>(c.z);
  o->vx = static_cast<float>(c.vx);
  o->vy = static_cast<float>(c.vy);
  o->w = c.w;
  o->type = c.type;
  o->n[0] = o->n[1] = o->n[2] = o->off = 0.f;
  if (c.type == 2) {
    const float q[4] = {static_cast<float>(c.qw), static_cast<float>(c.qx), static_cast<float>(c.qy), static_cast<float>(c.qz)};
    geom_ring_plane(o->x, o->y, o->z, q, o->n, &o->off);
  }
}
// The obstacle that owns the occupied voxel whose corner is (ptx, pty, ptz): the first, in id order, of
//   type 3 (cylinder): |corner - (x, y, corner.z)| <= w + clearance, compared in double (:131-138);
//   type 2 (ring): |w / 2 - dist(centre, projection on the ring's plane)| < 2 res (a double against an fp32 product)
//                  and distance to the plane < 2 res (:139-150).
// Returns its index or -1 (velocity 0).
__device__ __forceinline__ int geom_gt_owner(const PopGeom& g, const PopCylDev* __restrict__ cyl, int n_cyl, float ptx,
                                             float pty, float ptz) {
  const float pt[3] = {ptx, pty, ptz};
  const float two_res = POP_FMUL(2.0f, g.res);
  for (int c = 0; c < n_cyl; c++) {
    if (cyl[c].type == 3) {
      const float d[3] = {POP_FSUB(ptx, cyl[c].x), POP_FSUB(pty, cyl[c].y), POP_FSUB(ptz, ptz)};
      const float dist = POP_FSQRT(geom_dot3(d, d));
      if (static_cast<double>(dist) <= cyl[c].w + static_cast<double>(g.clearance)) return c;
    } else if (cyl[c].type == 2) {
      const float sd = POP_FADD(geom_dot3(cyl[c].n, pt), cyl[c].off);
      const float pc[3] = {cyl[c].x, cyl[c].y, cyl[c].z};
      float d[3];
      for (int k = 0; k < 3; k++) d[k] = POP_FSUB(pc[k], POP_FSUB(pt[k], POP_FMUL(sd, cyl[c].n[k])));
      const float dist = POP_FSQRT(geom_dot3(d, d));
      if (fabs(cyl[c].w / 2 - static_cast<double>(dist)) < static_cast<double>(two_res) && fabsf(sd) < two_res) return c;
    }
  }
  return -1;
}

// ---- cell access -----------------------------------------------------------------------------
// Layout: [agent][t][z][y][x], x fastest, one slice = V cells.  u8 cells hold exact small
// integers (0/1 ground truth + unit peer risks); f32 cells hold the reference's float risks.
template <typename C> struct Cell;

template <> struct Cell<uint8_t> {
  static constexpr int kBytes = 1;
  // set slice cell to 1; returns true iff this call changed it from 0
  static __device__ __forceinline__ bool set_one_first(uint8_t* base, size_t off) {
    unsigned* w = reinterpret_cast<unsigned*>(reinterpret_cast<uintptr_t>(base + off) & ~uintptr_t(3));
    const unsigned sh = 8u * static_cast<unsigned>(reinterpret_cast<uintptr_t>(base + off) & 3u);
    const unsigned old = atomicOr(w, 1u << sh);
    return ((old >> sh) & 0xffu) == 0u;
  }
  static __device__ __forceinline__ void store_one(uint8_t* base, size_t off) { base[off] = 1; }
  // += risk (unit risks only): saturating byte increment via CAS on the containing word
  static __device__ __forceinline__ void add(uint8_t* base, size_t off, float risk) {
    unsigned* w = reinterpret_cast<unsigned*>(reinterpret_cast<uintptr_t>(base + off) & ~uintptr_t(3));
    const unsigned sh = 8u * static_cast<unsigned>(reinterpret_cast<uintptr_t>(base + off) & 3u);
    const unsigned inc = static_cast<unsigned>(risk);
    unsigned old = *w, assumed;
    do {
      assumed = old;
      unsigned b = (assumed >> sh) & 0xffu;
      unsigned nb = b + inc > 255u ? 255u : b + inc;
      unsigned nw = (assumed & ~(0xffu << sh)) | (nb << sh);
      old = atomicCAS(w, assumed, nw);
    } while (old != assumed);
  }
  static __device__ __forceinline__ float load(const uint8_t* base, size_t off) {
    return static_cast<float>(base[off]);
  }
};

template <> struct Cell<float> {
  static constexpr int kBytes = 4;
  static __device__ __forceinline__ bool set_one_first(uint8_t* base, size_t off) {
    float old = atomicExch(reinterpret_cast<float*>(base) + off, 1.0f);
    return old == 0.0f;
  }
  static __device__ __forceinline__ void store_one(uint8_t* base, size_t off) {
    reinterpret_cast<float*>(base)[off] = 1.0f;
  }
  static __device__ __forceinline__ void add(uint8_t* base, size_t off, float risk) {
    atomicAdd(reinterpret_cast<float*>(base) + off, risk);
  }
  static __device__ __forceinline__ float load(const uint8_t* base, size_t off) {
    return reinterpret_cast<const float*>(base)[off];
  }
};

// Sum of the (2S+1) x (2S+1) x (2SZ+1) collision footprint in the reference's order (x outer, y, z inner), compared
// with the threshold once at the end; see geom_query for why that equals the reference's early exit.  Without an
// exit between them, all loads of the footprint are independent and go out together.
template <typename C, int S, int SZ>
__device__ __forceinline__ int geom_region_sum(const PopGeom& g, const uint8_t* grid, size_t slice, int ix, int iy,
                                               int iz, float thr) {
  constexpr int NY = 2 * S + 1, NZ = 2 * SZ + 1;
  float sum = 0.0f;
#pragma unroll
  for (int dx = -S; dx <= S; dx++) {
    const int x = ix + dx;
    const bool xin = x >= 0 && x < g.nx;
    float v[NY * NZ];
#pragma unroll
    for (int dy = -S; dy <= S; dy++) {
      const int y = iy + dy;
#pragma unroll
      for (int dz = -SZ; dz <= SZ; dz++) {
        const int z = iz + dz;
        const bool in = xin && y >= 0 && y < g.ny && z >= 0 && z < g.nz;
        v[(dy + S) * NZ + dz + SZ] =
            in ? Cell<C>::load(grid, slice + static_cast<size_t>(z) * g.nx * g.ny + static_cast<size_t>(y) * g.nx + x) : 0.0f;
      }
    }
#pragma unroll
    for (int k = 0; k < NY * NZ; k++) sum = __fadd_rn(sum, v[k]);
  }
  return sum > thr ? 1 : 0;
}

// Byte cells hold small integers (saturating unit risks), so the footprint's float sum is exact in any order and the
// 2S+1 <= 9 x-contiguous cells of a row can be read as three aligned words instead of nine byte loads: the window is
// shifted into place with funnel shifts and its bytes are added with dp4a.  Only for footprints that lie inside the
// map in x (the clipped ones keep the per-cell path); rows outside in y or z contribute nothing, as in the reference.
template <int S, int SZ>
__device__ __forceinline__ int geom_region_sum_u8_rows(const PopGeom& g, const uint8_t* grid, size_t slice, int ix,
                                                       int iy, int iz, float thr) {
  constexpr int W = 2 * S + 1;
  static_assert(W <= 9, "window of at most 9 cells: 12 bytes cover it at any alignment");
  unsigned sum = 0u;
#pragma unroll
  for (int dz = -SZ; dz <= SZ; dz++) {
    const int z = iz + dz;
#pragma unroll
    for (int dy = -S; dy <= S; dy++) {
      const int y = iy + dy;
      if (y < 0 || y >= g.ny || z < 0 || z >= g.nz) continue;
      const uint8_t* row = grid + slice + static_cast<size_t>(z) * g.nx * g.ny + static_cast<size_t>(y) * g.nx + (ix - S);
      const unsigned* w = reinterpret_cast<const unsigned*>(reinterpret_cast<uintptr_t>(row) & ~uintptr_t(3));
      const unsigned sh = 8u * static_cast<unsigned>(reinterpret_cast<uintptr_t>(row) & 3u);
      const unsigned w0 = w[0], w1 = w[1], w2 = W + 3 > 8 ? w[2] : 0u;
      unsigned b0 = __funnelshift_r(w0, w1, sh);           // window bytes 0..3
      unsigned b1 = __funnelshift_r(w1, w2, sh);           // window bytes 4..7
      unsigned b2 = w2 >> sh;                              // window byte 8
      if (W < 4) b0 &= (1u << (8 * (W & 3))) - 1u;
      if (W <= 4) b1 = 0u;
      else if (W < 8) b1 &= (1u << (8 * (W & 3))) - 1u;
      b2 = W > 8 ? (b2 & 0xffu) : 0u;
      sum = __dp4a(b0, 0x01010101u, sum);
      sum = __dp4a(b1, 0x01010101u, sum);
      sum += b2;
    }
  }
  return static_cast<float>(sum) > thr ? 1 : 0;
}

template <typename C, int S, int SZ>
__device__ __forceinline__ int geom_region(const PopGeom& g, const uint8_t* grid, size_t slice, int ix, int iy, int iz,
                                           float thr) {
  if constexpr (sizeof(C) == 1) {
    if (ix - S >= 0 && ix + S < g.nx) return geom_region_sum_u8_rows<S, SZ>(g, grid, slice, ix, iy, iz, thr);
  }
  return geom_region_sum<C, S, SZ>(g, grid, slice, ix, iy, iz, thr);
}

// getClearOcccupancy(pos, int t): fake_particle_risk_voxel.cpp:309-330 / risk_base.cpp:228-250.
// Kernel order x outer, y, z inner with early exit, as in the reference.
template <typename C>
__device__ __forceinline__ int geom_query(const PopGeom& g, const uint8_t* grid, float posex,
                                          float posey, float posez, double px, double py,
                                          double pz, int t) {
  if (g.query_variant == POP_QUERY_FAKE) {
    if (pz < static_cast<double>(g.ground) || pz > static_cast<double>(g.ceiling)) return -1;
  } else {
    if (pz < static_cast<double>(g.ground)) return 1;
    if (pz > static_cast<double>(g.ceiling)) return 1;
  }
  const float fx = __fsub_rn(static_cast<float>(px), posex);
  const float fy = __fsub_rn(static_cast<float>(py), posey);
  const float fz = __fsub_rn(static_cast<float>(pz), posez);
  const int ix = geom_axis_index(fx, g.rx, g.res);
  const int iy = geom_axis_index(fy, g.ry, g.res);
  const int iz = geom_axis_index(fz, g.rz, g.res);
  if (!geom_in_range_i(g, ix, iy, iz)) return -1;
  if (t < 0 || t >= g.nt) return -1;  // the reference would index risk_maps_ out of bounds
  const float thr = g.query_variant == POP_QUERY_FAKE
                        ? g.risk_threshold
                        : __fsub_rn(g.thr_region, __fmul_rn(static_cast<float>(t), g.decay_region));
  const size_t slice = static_cast<size_t>(t) * g.V;
  const int s = g.inf_step;
  const int sz = g.kernel_shape == POP_KERNEL_DISC_Z0 ? 0 : s;
  // Cells are non-negative and round-to-nearest addition is monotone, so the running sum never decreases and the
  // reference's early exit ("return 1 as soon as the sum exceeds the threshold") equals comparing the sum of the
  // whole footprint, taken in the same order, once.  That lets all its loads be issued together instead of one
  // dependent load per cell (the A* inner loop is bound by exactly this latency).  Out-of-range cells contribute 0,
  // which changes neither the sum nor the comparison.
  if (s == 2 && sz == 0) return geom_region<C, 2, 0>(g, grid, slice, ix, iy, iz, thr);
  if (s == 1 && sz == 1) return geom_region<C, 1, 1>(g, grid, slice, ix, iy, iz, thr);
  if (s == 2 && sz == 2) return geom_region<C, 2, 2>(g, grid, slice, ix, iy, iz, thr);
  if (s == 1 && sz == 0) return geom_region<C, 1, 0>(g, grid, slice, ix, iy, iz, thr);
  if (s == 3 && sz == 0) return geom_region<C, 3, 0>(g, grid, slice, ix, iy, iz, thr);
  if (s == 4 && sz == 0) return geom_region<C, 4, 0>(g, grid, slice, ix, iy, iz, thr);   // 0.45 m clearance at 0.1 m
  float sum = 0.0f;
  for (int dx = -s; dx <= s; dx++) {
    const int x = ix + dx;
    if (x < 0 || x >= g.nx) continue;
    for (int dy = -s; dy <= s; dy++) {
      const int y = iy + dy;
      if (y < 0 || y >= g.ny) continue;
      for (int dz = -sz; dz <= sz; dz++) {
        const int z = iz + dz;
        if (z < 0 || z >= g.nz) continue;
        sum = __fadd_rn(sum, Cell<C>::load(grid, slice + static_cast<size_t>(z) * g.nx * g.ny +
                                                     static_cast<size_t>(y) * g.nx + x));
        if (sum > thr) return 1;
      }
    }
  }
  return 0;
}

// MapBase::getInflateOccupancy (map.cpp:345-395): single-voxel test.  mode 0 (pos): slice 0; mode 1 (pos, int t);
// mode 2 (pos, double t): slices ceil(t / dt) and floor(t / dt).  A slice index equal to PREDICTION_TIMES, which the
// reference lets through to read one slice past the end, answers -1 (mode 1) or is clamped to nt - 1 (mode 2).
template <typename C>
__device__ __forceinline__ int geom_inflate_query(const PopGeom& g, const uint8_t* grid, float posex, float posey,
                                                  float posez, double px, double py, double pz, int mode, double t) {
  const float fx = __fsub_rn(static_cast<float>(px), posex);
  const float fy = __fsub_rn(static_cast<float>(py), posey);
  const float fz = __fsub_rn(static_cast<float>(pz), posez);
  const int idx = geom_voxel_index(g, fx, fy, fz);
  if (!geom_in_range_f(g, fx, fy, fz)) return -1;
  if (idx < 0 || idx >= g.V) return -1;
  if (mode == 0) return Cell<C>::load(grid, idx) > g.risk_threshold ? 1 : 0;
  if (mode == 1) {
    const int ti = static_cast<int>(t);
    if (ti < 0 || ti >= g.nt) return -1;
    return Cell<C>::load(grid, static_cast<size_t>(ti) * g.V + idx) > g.risk_threshold ? 1 : 0;
  }
  const double q = t / static_cast<double>(g.dt);
  int tc = static_cast<int>(ceil(q)), tf = static_cast<int>(floor(q));
  tc = tc > g.nt - 1 ? g.nt - 1 : tc;
  tf = tf > g.nt - 1 ? g.nt - 1 : tf;
  if (tc < 0 || tf < 0) return -1;
  if (Cell<C>::load(grid, static_cast<size_t>(tc) * g.V + idx) > g.risk_threshold) return 1;
  if (Cell<C>::load(grid, static_cast<size_t>(tf) * g.V + idx) > g.risk_threshold) return 1;
  return 0;
}
