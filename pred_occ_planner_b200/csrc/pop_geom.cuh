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

// Sum of the (2S+1) x (2S+1) x (2SZ+1) collision footprint in the reference's order (x outer, y, z inner), one
// x-plane at a time: the plane's cells are loaded together, then added in order; see geom_query.
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
    if (sum > thr) return 1;
  }
  return 0;
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
  // Cells are non-negative, so the running sum never decreases and the reference's early exit ("return 1 as soon
  // as the sum exceeds the threshold") equals comparing the sum of each x-plane, taken in the same order, once per
  // plane.  That lets the loads of a plane be issued together instead of one dependent load per cell (the A*
  // inner loop is bound by exactly this latency).  Out-of-range cells contribute 0, which changes neither the sum
  // nor the comparison.
  if (s == 2 && sz == 0) return geom_region_sum<C, 2, 0>(g, grid, slice, ix, iy, iz, thr);
  if (s == 1 && sz == 1) return geom_region_sum<C, 1, 1>(g, grid, slice, ix, iy, iz, thr);
  if (s == 2 && sz == 2) return geom_region_sum<C, 2, 2>(g, grid, slice, ix, iy, iz, thr);
  if (s == 1 && sz == 0) return geom_region_sum<C, 1, 0>(g, grid, slice, ix, iy, iz, thr);
  if (s == 3 && sz == 0) return geom_region_sum<C, 3, 0>(g, grid, slice, ix, iy, iz, thr);
  if (s == 4 && sz == 0) return geom_region_sum<C, 4, 0>(g, grid, slice, ix, iy, iz, thr);   // 0.45 m clearance at 0.1 m
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
