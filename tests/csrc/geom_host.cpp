// geom_host.cpp — TEST INFRASTRUCTURE: the product's device grid geometry and occupancy query
// (pred_occ_planner_b200/csrc/pop_geom.cuh: index math of map.h:153-215, getClearOcccupancy of
// fake_particle_risk_voxel.cpp:309-330 / risk_base.cpp:228-250) compiled for the host.  CUDA's fp32 round-to-nearest
// intrinsics are the plain operators on the host with contraction off; the atomics only have to compile (the query
// path does not use them).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p |= v; return o; }
static inline unsigned atomicCAS(unsigned* p, unsigned c, unsigned v) { unsigned o = *p; if (o == c) *p = v; return o; }
static inline float atomicExch(float* p, float v) { float o = *p; *p = v; return o; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p += v; return o; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
  return static_cast<unsigned>(((static_cast<unsigned long long>(hi) << 32) | lo) >> (sh & 31u));
}
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
  for (int k = 0; k < 4; k++) c += ((a >> (8 * k)) & 0xffu) * ((b >> (8 * k)) & 0xffu);
  return c;
}

#include "../../pred_occ_planner_b200/csrc/pop_geom.cuh"

// PopGeom from the configuration, as derive_geometry does (pop_capi.cu:78-103)
static PopGeom geom_of(const pop_config& c) {
  PopGeom g{};
  g.nx = c.nx; g.ny = c.ny; g.nz = c.nz; g.nt = c.nt;
  g.V = c.nx * c.ny * c.nz;
  g.res = c.resolution;
  g.dt = c.time_resolution;
  g.rx = static_cast<float>(c.nx / 2) * g.res;
  g.ry = static_cast<float>(c.ny / 2) * g.res;
  g.rz = static_cast<float>(c.nz / 2) * g.res;
  g.risk_threshold = c.risk_threshold;
  g.clearance = c.clearance;
  g.ground = c.ground_height;
  g.ceiling = c.ceiling_height;
  g.thr_region = c.risk_threshold_region;
  g.decay_region = c.risk_region_decay;
  g.decay_voxel = c.risk_voxel_decay;
  g.inf_step = static_cast<int>(c.clearance / c.resolution);
  g.kernel_shape = c.kernel_shape;
  g.query_variant = c.query_variant;
  return g;
}

extern "C" {
// grid: one agent's SOGM in the device layout [t][z][y][x] (u8 or f32 cells per cfg->cell_dtype)
void dev_query(const pop_config* cfg, const void* grid, const float* pose, const double* pos, const int32_t* t, int n,
               int8_t* out) {
  const PopGeom g = geom_of(*cfg);
  // the byte-cell query reads whole aligned words around a row (the device grid is allocated with slack for that)
  std::vector<uint32_t> padded;
  const uint8_t* base = static_cast<const uint8_t*>(grid);
  if (cfg->cell_dtype == POP_CELL_U8) {
    const size_t bytes = static_cast<size_t>(g.nt) * g.V;
    padded.assign(bytes / 4 + 8, 0u);
    memcpy(padded.data(), grid, bytes);
    base = reinterpret_cast<const uint8_t*>(padded.data());
  }
  for (int i = 0; i < n; i++)
    out[i] = static_cast<int8_t>(cfg->cell_dtype == POP_CELL_U8
                                     ? geom_query<uint8_t>(g, base, pose[0], pose[1], pose[2], pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], t[i])
                                     : geom_query<float>(g, base, pose[0], pose[1], pose[2], pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], t[i]));
}
// voxel index of a map-frame point and the world position of a voxel (getVoxelIndex / getVoxelPosition)
int dev_voxel_index(const pop_config* cfg, float x, float y, float z) { return geom_voxel_index(geom_of(*cfg), x, y, z); }
float dev_axis_position(const pop_config* cfg, int i, int axis, float pose) {
  const PopGeom g = geom_of(*cfg);
  return geom_axis_position(i, axis == 0 ? g.rx : axis == 1 ? g.ry : g.rz, g.res, pose);
}
// owner (index into cyl, -1 none) of each voxel corner, as the ground-truth voxel update decides it
void dev_gt_owner(const pop_config* cfg, const pop_cylinder* cyl, int n_cyl, const float* corners, int n, int32_t* out) {
  const PopGeom g = geom_of(*cfg);
  PopCylDev* d = new PopCylDev[n_cyl > 0 ? n_cyl : 1];
  for (int c = 0; c < n_cyl; c++) geom_cyl_prepare(cyl[c], &d[c]);
  for (int i = 0; i < n; i++) out[i] = geom_gt_owner(g, d, n_cyl, corners[3 * i], corners[3 * i + 1], corners[3 * i + 2]);
  delete[] d;
}
}
