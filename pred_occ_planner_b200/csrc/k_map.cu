// k_map.cu — SOGM kernels: ground-truth voxel update (K1f), peer projection (K5), occupancy
// queries (A13), obstacle-point extraction (A15), reference-layout download.
//
// Reference semantics: plan_env/src/fake_particle_risk_voxel.cpp:80-222,281-348,
// plan_env/src/map.cpp:463-531, plan_env/src/risk_base.cpp:228-337,
// traj_coordinator/src/particles.cpp:316-422.
#include <cstdio>
#include <cstdlib>

#include "pop_bezier.cuh"
#include "pop_geom.cuh"

// =============================================================================================
// K1f.a  clear — streaming zero fill of every local agent's [t][z][y][x] grid.
// One elected thread per CTA issues TMA bulk stores (cp.async.bulk shared→global, SASS UBLKCP)
// from a zeroed shared-memory tile; the other variant uses 16-byte vector stores.
// Algorithmic bytes: V*T*s per agent, written once.
// =============================================================================================
constexpr int kClearTile = 32 * 1024;  // bytes per bulk store

__global__ void __launch_bounds__(128) k_clear_tma(uint8_t* __restrict__ base, unsigned long long total) {
  extern __shared__ __align__(128) uint8_t ztile[];
  for (int i = threadIdx.x * 16; i < kClearTile; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(ztile + i) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned smem = static_cast<unsigned>(__cvta_generic_to_shared(ztile));
    const unsigned long long step = static_cast<unsigned long long>(gridDim.x) * kClearTile;
    for (unsigned long long off = static_cast<unsigned long long>(blockIdx.x) * kClearTile; off < total;
         off += step) {
      const unsigned long long rem = total - off;
      const unsigned n = rem < kClearTile ? static_cast<unsigned>(rem) : kClearTile;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off),
                   "r"(smem), "r"(n)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

__global__ void __launch_bounds__(256) k_clear_vec(uint4* __restrict__ base, unsigned long long n16) {
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  const uint4 z = make_uint4(0, 0, 0, 0);
  unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  // 4 independent stores per iteration
  for (; i + 3 * stride < n16; i += 4 * stride) {
    base[i] = z; base[i + stride] = z; base[i + 2 * stride] = z; base[i + 3 * stride] = z;
  }
  for (; i < n16; i += stride) base[i] = z;
}

// =============================================================================================
// K1f.b  slice 0 from the world cloud (fake_particle_risk_voxel.cpp:88-115) + compaction of the
// occupied voxels (:118-125).  Thread per (agent, point).
// =============================================================================================
template <typename C>
__global__ void __launch_bounds__(256) k_mark_slice0(PopGeom g, uint8_t* __restrict__ grid,
                                                     const float* __restrict__ cloud, int n_cloud,
                                                     const float* __restrict__ poses,
                                                     int* __restrict__ occ_list,
                                                     int* __restrict__ occ_count, int occ_cap) {
  const int a = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cloud) return;
  const float posex = poses[3 * a], posey = poses[3 * a + 1], posez = poses[3 * a + 2];
  const float px = cloud[3 * i], py = cloud[3 * i + 1], pz = cloud[3 * i + 2];
  if (!isfinite(px) || !isfinite(py) || !isfinite(pz)) return;
  // pcl::PassThrough limits [pose - range, pose + range] per axis (:88-104)
  if (px < __fsub_rn(posex, g.rx) || px > __fadd_rn(posex, g.rx)) return;
  if (py < __fsub_rn(posey, g.ry) || py > __fadd_rn(posey, g.ry)) return;
  if (pz < __fsub_rn(posez, g.rz) || pz > __fadd_rn(posez, g.rz)) return;
  const float x = __fsub_rn(px, posex), y = __fsub_rn(py, posey), z = __fsub_rn(pz, posez);
  if (!geom_in_range_f(g, x, y, z)) return;
  const int idx = geom_voxel_index(g, x, y, z);
  if (idx < 0 || idx >= g.V) return;
  if (!(1.0f > g.risk_threshold)) {  // tmp[i][0] > risk_threshold_ never true → no future slices
    Cell<C>::store_one(grid + a * g.agent_stride, idx);
    return;
  }
  if (Cell<C>::set_one_first(grid + a * g.agent_stride, idx)) {
    const int slot = atomicAdd(&occ_count[a], 1);
    if (slot < occ_cap) occ_list[static_cast<size_t>(a) * occ_cap + slot] = idx;
  }
}

// =============================================================================================
// K1f.c  future slices: owning cylinder → velocity → deposit voxel(pt + v*dt*k - pose), k=1..T-1
// (fake_particle_risk_voxel.cpp:127-162).  Thread per (agent, occupied voxel).
// =============================================================================================
template <typename C>
__global__ void __launch_bounds__(128) k_future_slices(PopGeom g, uint8_t* __restrict__ grid,
                                                       const float* __restrict__ poses,
                                                       const PopCylDev* __restrict__ cyl, int n_cyl,
                                                       const int* __restrict__ occ_list,
                                                       const int* __restrict__ occ_count,
                                                       int occ_cap) {
  extern __shared__ PopCylDev scyl[];
  for (int c = threadIdx.x; c < n_cyl; c += blockDim.x) scyl[c] = cyl[c];
  __syncthreads();
  const int a = blockIdx.y;
  int count = occ_count[a];
  if (count > occ_cap) count = occ_cap;
  const float posex = poses[3 * a], posey = poses[3 * a + 1], posez = poses[3 * a + 2];
  uint8_t* ag = grid + a * g.agent_stride;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
    const int idx = occ_list[static_cast<size_t>(a) * occ_cap + e];
    const int ix = idx % g.nx, iy = (idx / g.nx) % g.ny, iz = idx / (g.nx * g.ny);
    const float ptx = geom_axis_position(ix, g.rx, g.res, posex);
    const float pty = geom_axis_position(iy, g.ry, g.res, posey);
    const float ptz = geom_axis_position(iz, g.rz, g.res, posez);
    float vx = 0.f, vy = 0.f;
    const int own = geom_gt_owner(g, scyl, n_cyl, ptx, pty, ptz);
    if (own >= 0) { vx = scyl[own].vx; vy = scyl[own].vy; }
    for (int k = 1; k < g.nt; k++) {
      const float kf = static_cast<float>(k);
      const float qx = __fsub_rn(__fadd_rn(ptx, __fmul_rn(__fmul_rn(vx, g.dt), kf)), posex);
      const float qy = __fsub_rn(__fadd_rn(pty, __fmul_rn(__fmul_rn(vy, g.dt), kf)), posey);
      const float qz = __fsub_rn(__fadd_rn(ptz, __fmul_rn(__fmul_rn(0.0f, g.dt), kf)), posez);
      if (!geom_in_range_f(g, qx, qy, qz)) continue;
      const int q = geom_voxel_index(g, qx, qy, qz);
      if (q < 0 || q >= g.V) continue;
      Cell<C>::store_one(ag, static_cast<size_t>(k) * g.V + q);
    }
  }
}

// =============================================================================================
// K5  peer projection.  (a) world position + validity of every peer at every slice time
// (ParticleATC::getWaypoints / getParticlesWithRisk, particles.cpp:316-422), shared by all
// local agents; (b) scatter body particles into each ego grid (addParticlesToRiskMap,
// fake_particle_risk_voxel.cpp:178-218,291-301).
// =============================================================================================
__global__ void k_peer_positions(int n_total, int nt, float dt, double t_map, float risk_rate,
                                 const pop_traj_record* __restrict__ trajs,
                                 const int* __restrict__ body_count, double* __restrict__ pos,
                                 int8_t* __restrict__ valid, float* __restrict__ sigma) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n_total * nt) return;
  const int i = tid / nt, k = tid % nt;
  const pop_traj_record& r = trajs[i];
  int8_t ok = 0;
  if (r.n_pieces > 0 && r.n_pieces <= POP_MAX_PIECES && body_count[i] > 0) {
    double t_end = r.start_time;
    for (int p = 0; p < r.n_pieces; p++) t_end = __dadd_rn(t_end, static_cast<double>(r.duration[p]));
    const double t0 = __dadd_rn(t_map, static_cast<double>(__fmul_rn(dt, 0.0f)));
    const double tk = __dadd_rn(t_map, static_cast<double>(__fmul_rn(dt, static_cast<float>(k))));
    // valid at slice k iff the first slice was inside (t_start, t_end) and t_k still is
    if (r.start_time < t0 && t_end > t0 && r.start_time < tk && t_end > tk) {
      ok = 1;
      double p3[3];
      pop_bezier_eval(r.duration, r.n_pieces, r.cpts, __dsub_rn(tk, r.start_time), 0, p3);
      pos[3 * tid] = p3[0]; pos[3 * tid + 1] = p3[1]; pos[3 * tid + 2] = p3[2];
      // pos_stddev = replan_risk_rate_ * static_cast<float>(t0 - ptr->time_start) (particles.cpp:372)
      sigma[tid] = __fmul_rn(risk_rate, static_cast<float>(__dsub_rn(tk, r.start_time)));
    }
  }
  valid[tid] = ok;
}

template <typename C>
__global__ void __launch_bounds__(256) k_peer_scatter(PopGeom g, uint8_t* __restrict__ grid,
                                                      const float* __restrict__ poses, int n_local,
                                                      int n_total, int id_base,
                                                      const double* __restrict__ peer_pos,
                                                      const int8_t* __restrict__ peer_valid,
                                                      const double* __restrict__ body,
                                                      const int* __restrict__ body_count,
                                                      const int* __restrict__ agent_ok,
                                                      const float* __restrict__ peer_sigma, int n_resample,
                                                      const float* __restrict__ noise, int n_noise) {
  // thread per (ego a, peer i, slice k, body particle j); j fastest
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per_agent = static_cast<long long>(n_total) * g.nt * POP_MAX_BODY;
  if (tid >= per_agent * n_local) return;
  const int a = static_cast<int>(tid / per_agent);
  if (agent_ok && !agent_ok[a]) return;   // this agent's map was not refreshed (rejected particle-map update)
  long long r = tid % per_agent;
  const int i = static_cast<int>(r / (g.nt * POP_MAX_BODY));
  r %= g.nt * POP_MAX_BODY;
  const int k = static_cast<int>(r / POP_MAX_BODY), j = static_cast<int>(r % POP_MAX_BODY);
  if (i == id_base + a) return;  // ego
  if (j >= body_count[i]) return;
  if (!peer_valid[i * g.nt + k]) return;
  const double* p = peer_pos + 3 * (static_cast<size_t>(i) * g.nt + k);
  const double* e = body + 3 * (static_cast<size_t>(i) * POP_MAX_BODY + j);
  const double pose3[3] = {static_cast<double>(poses[3 * a]), static_cast<double>(poses[3 * a + 1]), static_cast<double>(poses[3 * a + 2])};
  const float sg = peer_sigma[i * g.nt + k];
  if (static_cast<double>(sg) < 1e-3 || n_noise <= 0) {   // `if (pos_stddev < 1e-3)`: the particle itself, risk 1 (particles.cpp:378-384)
    // (pt + e) - (double)pose, then float cast
    const float x = static_cast<float>(__dsub_rn(__dadd_rn(p[0], e[0]), pose3[0]));
    const float y = static_cast<float>(__dsub_rn(__dadd_rn(p[1], e[1]), pose3[1]));
    const float z = static_cast<float>(__dsub_rn(__dadd_rn(p[2], e[2]), pose3[2]));
    if (!geom_in_range_f(g, x, y, z)) return;
    const int idx = geom_voxel_index(g, x, y, z);
    if (idx < 0 || idx >= g.V) return;
    Cell<C>::add(grid + a * g.agent_stride, static_cast<size_t>(k) * g.V + idx, 1.0f);
    return;
  }
  // Gaussian resampling (particles.cpp:385-408): n_resample copies displaced by (nx, ny, nz) = noise triple * stddev,
  // weighted exp(-|n|^2 / (2 stddev^2)), the weights of one particle normalised to sum n_resample.  The reference
  // re-seeds its engine in every call, so every (peer, slice) replays the noise sequence from its start: particle j
  // takes triples j * n_resample ... (wrapping around the table).  exp is evaluated in double and rounded once, which
  // is what a correctly rounded expf (glibc's) returns.
  float risk[16];
  int cell[16];
  float sum = 0.0f;
  const float s2 = __fmul_rn(sg, sg);
  for (int r = 0; r < n_resample; r++) {
    const float* zt = noise + 3 * ((static_cast<long long>(j) * n_resample + r) % n_noise);
    const float nx = __fadd_rn(__fmul_rn(zt[0], sg), 0.0f), ny = __fadd_rn(__fmul_rn(zt[1], sg), 0.0f),
                nz = __fadd_rn(__fmul_rn(zt[2], sg), 0.0f);
    const float q = __fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz));
    const float arg = __fdiv_rn(__fmul_rn(-0.5f, q), s2);
    risk[r] = static_cast<float>(exp(static_cast<double>(arg)));
    sum = __fadd_rn(sum, risk[r]);
    const float x = static_cast<float>(__dsub_rn(__dadd_rn(__dadd_rn(p[0], e[0]), static_cast<double>(nx)), pose3[0]));
    const float y = static_cast<float>(__dsub_rn(__dadd_rn(__dadd_rn(p[1], e[1]), static_cast<double>(ny)), pose3[1]));
    const float z = static_cast<float>(__dsub_rn(__dadd_rn(__dadd_rn(p[2], e[2]), static_cast<double>(nz)), pose3[2]));
    int idx = -1;
    if (geom_in_range_f(g, x, y, z)) {
      idx = geom_voxel_index(g, x, y, z);
      if (idx < 0 || idx >= g.V) idx = -1;
    }
    cell[r] = idx;
  }
  const float rn = static_cast<float>(n_resample);
  for (int r = 0; r < n_resample; r++)
    if (cell[r] >= 0)
      Cell<C>::add(grid + a * g.agent_stride, static_cast<size_t>(k) * g.V + cell[r], __fdiv_rn(__fmul_rn(risk[r], rn), sum));
}

// addParticlesToRiskMap for caller-provided particles (map-frame doubles)
template <typename C>
__global__ void k_add_particles(PopGeom g, uint8_t* __restrict__ ag, const double* __restrict__ pts,
                                const float* __restrict__ risks, int n, int t_index) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = static_cast<float>(pts[3 * i]), y = static_cast<float>(pts[3 * i + 1]),
              z = static_cast<float>(pts[3 * i + 2]);
  if (!geom_in_range_f(g, x, y, z)) return;
  const int idx = geom_voxel_index(g, x, y, z);
  if (idx < 0 || idx >= g.V) return;
  Cell<C>::add(ag, static_cast<size_t>(t_index) * g.V + idx, risks[i]);
}

// =============================================================================================
// A13 batched queries
// =============================================================================================
template <typename C>
__global__ void k_query(PopGeom g, const uint8_t* __restrict__ ag, const float* __restrict__ pose,
                        const double* __restrict__ pos, const int32_t* __restrict__ t_idx,
                        const double* __restrict__ dts, int n, int8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int t;
  if (dts) {  // getClearOcccupancy(pos, double dt): floor(dt / time_resolution_) clamped above
    t = static_cast<int>(floor(__ddiv_rn(dts[i], static_cast<double>(g.dt))));
    t = t > g.nt - 1 ? g.nt - 1 : t;
  } else {
    t = t_idx[i];
  }
  out[i] = static_cast<int8_t>(
      geom_query<C>(g, ag, pose[0], pose[1], pose[2], pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], t));
}

// getInflateOccupancy batched: mode 0 (pos), 1 (pos, int t), 2 (pos, double t); ts = n doubles (ignored for mode 0)
template <typename C>
__global__ void k_query_inflate(PopGeom g, const uint8_t* __restrict__ ag, const float* __restrict__ pose,
                                const double* __restrict__ pos, const double* __restrict__ ts, int mode, int n,
                                int8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = static_cast<int8_t>(geom_inflate_query<C>(g, ag, pose[0], pose[1], pose[2], pos[3 * i], pos[3 * i + 1],
                                                     pos[3 * i + 2], mode, mode ? ts[i] : 0.0));
}

// =============================================================================================
// A15 obstacle points in a box × slice range, emitted in the reference's order (z, y, x, slice)
// (map.cpp:463-531).  One CTA; each thread owns a contiguous run of box voxels, block scan of
// counts, then ordered writes.
// =============================================================================================
struct BoxSpec { int lx, ly, lz, hx, hy, hz, j0, j1; };

template <typename C>
__device__ __forceinline__ int obst_count_or_emit(const PopGeom& g, const uint8_t* ag, const BoxSpec& b,
                                                  float posex, float posey, float posez, int v,
                                                  double* out, int base, int cap) {
  const int bx = b.hx - b.lx + 1, by = b.hy - b.ly + 1;
  const int x = b.lx + v % bx, y = b.ly + (v / bx) % by, z = b.lz + v / (bx * by);
  const size_t i = static_cast<size_t>(x) + static_cast<size_t>(y) * g.nx + static_cast<size_t>(z) * g.nx * g.ny;
  int c = 0;
  for (int j = b.j0; j <= b.j1; j++) {
    const float thr = g.query_variant == POP_QUERY_FAKE
                          ? g.risk_threshold
                          : __fsub_rn(g.risk_threshold, __fmul_rn(g.decay_voxel, static_cast<float>(j)));
    if (Cell<C>::load(ag, static_cast<size_t>(j) * g.V + i) > thr) {
      if (out && base + c < cap) {
        out[3 * (base + c)] = static_cast<double>(geom_axis_position(x, g.rx, g.res, posex));
        out[3 * (base + c) + 1] = static_cast<double>(geom_axis_position(y, g.ry, g.res, posey));
        out[3 * (base + c) + 2] = static_cast<double>(geom_axis_position(z, g.rz, g.res, posez));
      }
      c++;
    }
  }
  return c;
}

template <typename C>
__global__ void __launch_bounds__(1024) k_obstacle_points(PopGeom g, const uint8_t* __restrict__ ag,
                                                          const float* __restrict__ pose, BoxSpec b,
                                                          double* __restrict__ out, int cap,
                                                          int* __restrict__ n_out) {
  __shared__ int warp_sums[32];
  __shared__ int total_s;
  const int bx = b.hx - b.lx + 1, by = b.hy - b.ly + 1, bz = b.hz - b.lz + 1;
  const int nvox = (bx > 0 && by > 0 && bz > 0) ? bx * by * bz : 0;
  const int per = (nvox + blockDim.x - 1) / blockDim.x;
  const int v0 = threadIdx.x * per, v1 = min(nvox, v0 + per);
  const float posex = pose[0], posey = pose[1], posez = pose[2];
  int cnt = 0;
  for (int v = v0; v < v1; v++) cnt += obst_count_or_emit<C>(g, ag, b, posex, posey, posez, v, nullptr, 0, 0);
  // block exclusive scan
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
    int wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += n;
    }
    warp_sums[lane] = wi - w;
    if (lane == 31) total_s = wi;
  }
  __syncthreads();
  int base = warp_sums[warp] + incl - cnt;
  for (int v = v0; v < v1; v++) base += obst_count_or_emit<C>(g, ag, b, posex, posey, posez, v, out, base, cap);
  if (threadIdx.x == 0) *n_out = total_s;
}

// reference-layout export: out[v*T + t] (float) from [t][v] cells
template <typename C>
__global__ void k_download(PopGeom g, const uint8_t* __restrict__ ag, float* __restrict__ out) {
  const size_t n = static_cast<size_t>(g.V) * g.nt;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t v = i / g.nt, t = i % g.nt;
    out[i] = Cell<C>::load(ag, t * g.V + v);
  }
}

// RiskBase::futureRiskCallback's copy (risk_base.cpp:60-68): in[v*T + t] (the "map/future_risk" layout) -> [t][v] cells
template <typename C>
__global__ void k_upload(PopGeom g, uint8_t* __restrict__ ag, const float* __restrict__ in) {
  const size_t n = static_cast<size_t>(g.V) * g.nt;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t t = i / g.V, v = i % g.V;   // coalesced writes; the strided reads hit L2
    if (sizeof(C) == 4) reinterpret_cast<float*>(ag)[i] = in[v * g.nt + t];
    else ag[i] = static_cast<uint8_t>(in[v * g.nt + t]);
  }
}

// =============================================================================================
// launchers
// =============================================================================================
// =============================================================================================
// K1f (tiled)  u8 grids: every SOGM byte is written exactly once, by a TMA bulk store of a
// finished shared-memory tile; nothing is read back and no partial-sector read-modify-write
// reaches L2/HBM.
//
// A tile is a band of H rows of one z-layer of one slice of one agent (nx*H bytes <= 96 KB, so
// two CTAs per SM alternate between building and draining; at BASELINE config 2 a band is the
// whole 300x300 layer).
//   1. k_vox_count / k_vox_scan / k_vox_fill: counting sort of every agent's cloud points by
//      (z-layer, band) — one list for slice 0 (voxel of pt - pose, keyed by its own row) and one
//      for the future slices (voxel corner + owner cylinder id, keyed by the SOURCE row).  The z
//      of pt + v*dt*k - pose does not depend on k (vz = 0) but is recomputed in fp32 exactly as the
//      reference does, so it may differ from the slice-0 layer by rounding.
//   2. k_vox_tiles: persistent CTAs walk the (agent, slice, layer, band) units; a unit zeroes its
//      tile, scatters the entries of the source bands within reach of the largest |vy|*dt*k
//      displaced by v*dt*k (same fp32 operation order as fake_particle_risk_voxel.cpp:127-162),
//      and hands the tile to cp.async.bulk (shared → global).
// Duplicate points in a voxel are idempotent stores of 1, so the per-voxel compaction of the
// scatter path is not needed.  Algorithmic bytes: V*T written once + N_pts*12 read per agent.
// =============================================================================================
struct VoxDev {
  int cap = 0, nl = 0, nkeys = 0, H = 0, nbands = 0;
  int* cnt = nullptr;            // [agent][2][nkeys]   entries per (layer, band); list 0: slice 0, 1: future
  int* off = nullptr;            // [agent][2][nkeys+1] exclusive prefix
  int* cur = nullptr;            // [agent][2][nkeys]   fill cursors
  unsigned* ent0 = nullptr;      // [agent][cap]        iy*nx + ix of the slice-0 voxel
  unsigned long long* entf = nullptr;  // [agent][cap]  ix | iy<<16 | owner<<32 (owner 0xffff: static)
};

// slice-0 voxel of a cloud point as the reference computes it, or false
__device__ __forceinline__ bool vox_point(const PopGeom& g, const float* __restrict__ cloud, int i, float posex,
                                          float posey, float posez, int* ix, int* iy, int* iz) {
  const float px = cloud[3 * i], py = cloud[3 * i + 1], pz = cloud[3 * i + 2];
  if (!isfinite(px) || !isfinite(py) || !isfinite(pz)) return false;
  if (px < __fsub_rn(posex, g.rx) || px > __fadd_rn(posex, g.rx)) return false;
  if (py < __fsub_rn(posey, g.ry) || py > __fadd_rn(posey, g.ry)) return false;
  if (pz < __fsub_rn(posez, g.rz) || pz > __fadd_rn(posez, g.rz)) return false;
  const float x = __fsub_rn(px, posex), y = __fsub_rn(py, posey), z = __fsub_rn(pz, posez);
  if (!geom_in_range_f(g, x, y, z)) return false;
  *ix = geom_axis_index(x, g.rx, g.res);
  *iy = geom_axis_index(y, g.ry, g.res);
  *iz = geom_axis_index(z, g.rz, g.res);
  return geom_in_range_i(g, *ix, *iy, *iz);
}
// z-layer of the future slices of the voxel in layer iz: voxel of (corner_z + 0*dt*k) - pose_z
__device__ __forceinline__ int vox_future_layer(const PopGeom& g, int iz, float posez) {
  const float ptz = geom_axis_position(iz, g.rz, g.res, posez);
  const float qz = __fsub_rn(__fadd_rn(ptz, __fmul_rn(__fmul_rn(0.0f, g.dt), 1.0f)), posez);
  if (!(qz > -g.rz && qz < g.rz)) return -1;
  const int l = geom_axis_index(qz, g.rz, g.res);
  return (l >= 0 && l < g.nz) ? l : -1;
}

__global__ void __launch_bounds__(256) k_vox_count(PopGeom g, int S, int nstrips, const float* __restrict__ cloud,
                                                   int n_cloud, const float* __restrict__ poses,
                                                   int* __restrict__ cnt) {
  const int a = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cloud) return;
  const float posez = poses[3 * a + 2];
  const int nkeys = g.nz * nstrips;
  int ix, iy, iz;
  if (!vox_point(g, cloud, i, poses[3 * a], poses[3 * a + 1], posez, &ix, &iy, &iz)) return;
  const int strip = iy / S;
  atomicAdd(&cnt[(a * 2 + 0) * nkeys + iz * nstrips + strip], 1);
  if (1.0f > g.risk_threshold) {   // tmp[i][0] > risk_threshold_ (fake_particle_risk_voxel.cpp:121)
    const int lf = vox_future_layer(g, iz, posez);
    if (lf >= 0) atomicAdd(&cnt[(a * 2 + 1) * nkeys + lf * nstrips + strip], 1);
  }
}

// exclusive prefix over the keys of one (agent, list) per CTA
__global__ void __launch_bounds__(256) k_vox_scan(int nkeys, int cap, const int* __restrict__ cnt,
                                                  int* __restrict__ off, int* __restrict__ cur) {
  __shared__ int part[256];
  const int l = blockIdx.x, tid = threadIdx.x;
  const int per = (nkeys + 255) / 256;
  const int k0 = tid * per, k1 = min(nkeys, k0 + per);
  int s = 0;
  for (int k = k0; k < k1; k++) s += cnt[l * nkeys + k];
  part[tid] = s;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < 256; t++) { const int v = part[t]; part[t] = run; run += v; }
    off[l * (nkeys + 1) + nkeys] = run < cap ? run : cap;
  }
  __syncthreads();
  int run = part[tid];
  for (int k = k0; k < k1; k++) {
    off[l * (nkeys + 1) + k] = run < cap ? run : cap;
    cur[l * nkeys + k] = 0;
    run += cnt[l * nkeys + k];
  }
}

__global__ void __launch_bounds__(256) k_vox_fill(PopGeom g, int S, int nstrips, const float* __restrict__ cloud,
                                                  int n_cloud, const float* __restrict__ poses,
                                                  const PopCylDev* __restrict__ cyl, int n_cyl,
                                                  const int* __restrict__ off, int* __restrict__ cur,
                                                  unsigned* __restrict__ ent0,
                                                  unsigned long long* __restrict__ entf, int cap) {
  extern __shared__ PopCylDev scyl[];
  for (int c = threadIdx.x; c < n_cyl; c += blockDim.x) scyl[c] = cyl[c];
  __syncthreads();
  const int a = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cloud) return;
  const float posex = poses[3 * a], posey = poses[3 * a + 1], posez = poses[3 * a + 2];
  const int nkeys = g.nz * nstrips;
  int ix, iy, iz;
  if (!vox_point(g, cloud, i, posex, posey, posez, &ix, &iy, &iz)) return;
  const int strip = iy / S;
  {
    const int l = a * 2 + 0, key = iz * nstrips + strip;
    const int pos = off[l * (nkeys + 1) + key] + atomicAdd(&cur[l * nkeys + key], 1);
    if (pos < cap) ent0[static_cast<size_t>(a) * cap + pos] = static_cast<unsigned>(iy * g.nx + ix);
  }
  if (!(1.0f > g.risk_threshold)) return;
  const int lf = vox_future_layer(g, iz, posez);
  if (lf < 0) return;
  // owner: first cylinder or ring, in id order, whose test the voxel corner passes (geom_gt_owner)
  const float ptx = geom_axis_position(ix, g.rx, g.res, posex);
  const float pty = geom_axis_position(iy, g.ry, g.res, posey);
  const float ptz = geom_axis_position(iz, g.rz, g.res, posez);
  const int own = geom_gt_owner(g, scyl, n_cyl, ptx, pty, ptz);
  const unsigned owner = own >= 0 ? static_cast<unsigned>(own) : 0xffffu;
  const int l = a * 2 + 1, key = lf * nstrips + strip;
  const int pos = off[l * (nkeys + 1) + key] + atomicAdd(&cur[l * nkeys + key], 1);
  if (pos < cap)
    entf[static_cast<size_t>(a) * cap + pos] = static_cast<unsigned long long>(ix) |
                                               (static_cast<unsigned long long>(iy) << 16) |
                                               (static_cast<unsigned long long>(owner) << 32);
}

constexpr int kTileThreads = 256;
constexpr int kTileChunk = 32 * 1024;   // bytes per bulk store
constexpr int kTileMaxDefault = 96 * 1024;   // tile bytes: two CTAs per SM (POP_TILE_KB overrides: smaller tiles, more CTAs per SM)
static int tile_max_bytes() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("POP_TILE_KB");
    v = e ? atoi(e) * 1024 : kTileMaxDefault;
    if (v < 16 * 1024) v = 16 * 1024;
    if (v > 200 * 1024) v = 200 * 1024;
  }
  return v;
}

__global__ void __launch_bounds__(kTileThreads) k_vox_tiles(PopGeom g, int H, int nbands, int S, int nstrips,
                                                               float reach_per_slice,
                                                               uint8_t* __restrict__ grid, int n_agents,
                                                               const float* __restrict__ poses,
                                                               const PopCylDev* __restrict__ cyl, int n_cyl,
                                                               const int* __restrict__ off,
                                                               const unsigned* __restrict__ ent0,
                                                               const unsigned long long* __restrict__ entf,
                                                               int cap) {
  extern __shared__ __align__(128) uint8_t tile[];
  const int tile_max = ((g.nx * H + 127) & ~127);
  float2* svel = reinterpret_cast<float2*>(tile + tile_max);   // cylinder velocities
  for (int c = threadIdx.x; c < n_cyl; c += blockDim.x) svel[c] = make_float2(cyl[c].vx, cyl[c].vy);
  const unsigned smem = static_cast<unsigned>(__cvta_generic_to_shared(tile));
  const int nkeys = g.nz * nstrips, nunits = g.nz * nbands, spb = H / S;   // strips per band (S divides H)
  const long long units = static_cast<long long>(n_agents) * g.nt * nunits;
  for (long long u = blockIdx.x; u < units; u += gridDim.x) {
    const int band = static_cast<int>(u % nbands);
    const int iz = static_cast<int>((u / nbands) % g.nz);
    const int k = static_cast<int>((u / nunits) % g.nt);
    const int a = static_cast<int>(u / (static_cast<long long>(nunits) * g.nt));
    const int y0 = band * H, rows = min(H, g.ny - y0), bytes = rows * g.nx;
    // the previous unit's bulk stores must have finished READING the tile before it is rebuilt
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
    for (int i = threadIdx.x * 16; i < bytes; i += kTileThreads * 16)
      *reinterpret_cast<uint4*>(tile + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (k == 0) {
      const int* o = off + (a * 2 + 0) * (nkeys + 1) + iz * nstrips;   // the strips of one layer are consecutive keys
      const unsigned* e = ent0 + static_cast<size_t>(a) * cap;
      const int e1 = o[min(nstrips, (band + 1) * spb)], base = y0 * g.nx;
      for (int q = o[band * spb] + threadIdx.x; q < e1; q += kTileThreads) tile[e[q] - base] = 1;
    } else {
      // source strips whose voxels can land in this band after k slices: |dy| <= reach*k rows (+2 for rounding)
      const int dr = static_cast<int>(reach_per_slice * static_cast<float>(k) + 2.0f) + 1;
      const int s0 = max(0, (y0 - dr) / S), s1 = min(nstrips - 1, (y0 + rows - 1 + dr) / S);
      const int* o = off + (a * 2 + 1) * (nkeys + 1) + iz * nstrips;
      const int e0 = o[s0], e1 = o[s1 + 1];   // the keys of one layer are consecutive
      const unsigned long long* e = entf + static_cast<size_t>(a) * cap;
      const float posex = poses[3 * a], posey = poses[3 * a + 1];
      const float kf = static_cast<float>(k);
      auto place = [&](unsigned long long v) {
        const int ix = static_cast<int>(v & 0xffffu), iy = static_cast<int>((v >> 16) & 0xffffu);
        const unsigned owner = static_cast<unsigned>(v >> 32) & 0xffffu;
        float vx = 0.f, vy = 0.f;
        if (owner != 0xffffu) { const float2 w = svel[owner]; vx = w.x; vy = w.y; }
        const float ptx = geom_axis_position(ix, g.rx, g.res, posex);
        const float pty = geom_axis_position(iy, g.ry, g.res, posey);
        const float qx = __fsub_rn(__fadd_rn(ptx, __fmul_rn(__fmul_rn(vx, g.dt), kf)), posex);
        const float qy = __fsub_rn(__fadd_rn(pty, __fmul_rn(__fmul_rn(vy, g.dt), kf)), posey);
        if (!(qx > -g.rx && qx < g.rx && qy > -g.ry && qy < g.ry)) return;
        const int jx = geom_axis_index(qx, g.rx, g.res), jy = geom_axis_index(qy, g.ry, g.res) - y0;
        if (jx < 0 || jx >= g.nx || jy < 0 || jy >= rows) return;
        tile[jy * g.nx + jx] = 1;
      };
      // four entries in flight per thread: the list comes from L2/HBM and one dependent load per iteration
      // left the tile build, not the bulk store, as the longer half of a unit on dense clouds (config 3)
      int q = e0 + threadIdx.x;
      for (; q + 3 * kTileThreads < e1; q += 4 * kTileThreads) {
        const unsigned long long v0 = e[q], v1 = e[q + kTileThreads], v2 = e[q + 2 * kTileThreads], v3 = e[q + 3 * kTileThreads];
        place(v0); place(v1); place(v2); place(v3);
      }
      for (; q < e1; q += kTileThreads) place(e[q]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      uint8_t* dst = grid + a * g.agent_stride + ((static_cast<size_t>(k) * g.nz + iz) * g.ny + y0) * g.nx;
      for (int b = 0; b < bytes; b += kTileChunk) {
        const unsigned n = bytes - b < kTileChunk ? bytes - b : kTileChunk;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + b),
                     "r"(smem + b), "r"(n)
                     : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

void popk_vox_free(pop_handle* h) {
  VoxDev* d = h->vox;
  if (!d) return;
  void* ptrs[] = {d->cnt, d->off, d->cur, d->ent0, d->entf};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete d;
  h->vox = nullptr;
}

// band height: the largest H with nx*H <= kTileMax whose tiles start and end on 16-byte boundaries
// (bulk-copy alignment); 0 if the grid cannot be tiled that way
static int vox_band_rows(const PopGeom& g) {
  const int kTileMax = tile_max_bytes();
  if (static_cast<long long>(g.nx) * g.ny <= kTileMax) return (g.nx * g.ny) % 16 == 0 ? g.ny : 0;
  int step = 1;
  while ((g.nx * step) % 16) step++;        // <= 16
  if ((static_cast<long long>(g.nx) * g.ny) % 16) return 0;   // layer starts must stay aligned too
  const int H = (kTileMax / g.nx) / step * step;
  return H > 0 ? H : 0;
}

// the tiled path applies to u8 grids (the scatter kernels cover f32 cells and unaligned shapes)
static bool vox_tiles_ok(const pop_handle* h) {
  const PopGeom& g = h->geom;
  const char* e = getenv("POP_VOXEL_MODE");   // "tiles" (default) | "scatter"
  if (e && e[0] == 's') return false;
  return h->cfg.cell_dtype == POP_CELL_U8 && vox_band_rows(g) > 0 && g.nx < 65536 && g.ny < 65536 &&
         h->cfg.max_cylinders < 0xffff;
}

static int vox_update_tiles(pop_handle* h, int sms, int* launches) {
  const PopGeom& g = h->geom;
  const int nl = h->cfg.n_agents_local, cap = h->cfg.max_cloud_points;
  const int H = vox_band_rows(g), nbands = (g.ny + H - 1) / H;
  // the entry lists are keyed by (layer, strip of S rows); S divides H so a band is a whole number of strips, and a
  // band's unit reads only the strips within reach of its rows instead of whole neighbouring bands
  int S = H;
  if (nbands > 1) for (int c = 32; c >= 1; c--) if (H % c == 0) { S = c; break; }
  const int nstrips = (g.ny + S - 1) / S, nkeys = g.nz * nstrips;
  const int smem = ((g.nx * H + 127) & ~127) + static_cast<int>(sizeof(float2)) * (h->cfg.max_cylinders + 1);
  if (!h->vox) {
    VoxDev* d = new VoxDev();
    h->vox = d;
    d->cap = cap; d->nl = nl; d->nkeys = nkeys; d->H = H; d->nbands = nbands;
    bool ok = true;
    ok &= cudaMalloc(&d->cnt, sizeof(int) * nl * 2 * nkeys) == cudaSuccess;
    ok &= cudaMalloc(&d->off, sizeof(int) * nl * 2 * (nkeys + 1)) == cudaSuccess;
    ok &= cudaMalloc(&d->cur, sizeof(int) * nl * 2 * nkeys) == cudaSuccess;
    ok &= cudaMalloc(&d->ent0, sizeof(unsigned) * static_cast<size_t>(nl) * cap) == cudaSuccess;
    ok &= cudaMalloc(&d->entf, sizeof(unsigned long long) * static_cast<size_t>(nl) * cap) == cudaSuccess;
    if (!ok) { cudaGetLastError(); popk_vox_free(h); return pop_fail(h, POP_E_CUDA, "voxel-update entry lists allocation failed"); }
    POP_CUDA(h, cudaFuncSetAttribute(k_vox_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  VoxDev* d = h->vox;
  POP_CUDA(h, cudaMemsetAsync(d->cnt, 0, sizeof(int) * nl * 2 * nkeys, h->stream));
  if (h->n_cloud > 0) {
    dim3 grid((h->n_cloud + 255) / 256, nl);
    k_vox_count<<<grid, 256, 0, h->stream>>>(g, S, nstrips, h->d_cloud, h->n_cloud, h->d_poses, d->cnt);
    POP_LAUNCH_CHECK(h);
    (*launches)++;
  }
  k_vox_scan<<<nl * 2, 256, 0, h->stream>>>(nkeys, cap, d->cnt, d->off, d->cur);
  POP_LAUNCH_CHECK(h);
  (*launches)++;
  if (h->n_cloud > 0) {
    dim3 grid((h->n_cloud + 255) / 256, nl);
    const size_t sm = sizeof(PopCylDev) * (h->n_cyl > 0 ? h->n_cyl : 1);
    k_vox_fill<<<grid, 256, sm, h->stream>>>(g, S, nstrips, h->d_cloud, h->n_cloud, h->d_poses, h->d_cyl, h->n_cyl,
                                             d->off, d->cur, d->ent0, d->entf, cap);
    POP_LAUNCH_CHECK(h);
    (*launches)++;
  }
  // rows a voxel can move per slice: |vy|*dt/res (the kernel adds a 2-row rounding margin)
  const float reach = h->max_abs_vy * g.dt / g.res;
  // persistent CTAs: as many as fit per SM with this tile size
  static int per_sm = 0;
  if (per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_vox_tiles, kTileThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
  }
  k_vox_tiles<<<per_sm * sms, kTileThreads, smem, h->stream>>>(g, H, nbands, S, nstrips, reach, h->d_grid, nl, h->d_poses, h->d_cyl,
                                                         h->n_cyl, d->off, d->ent0, d->entf, cap);
  POP_LAUNCH_CHECK(h);
  (*launches)++;
  return POP_OK;
}

static int clear_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("POP_CLEAR_MODE");  // "tma" (default) | "vec" | "memset"
    mode = 0;
    if (e && e[0] == 'v') mode = 1;
    if (e && e[0] == 'm') mode = 2;
  }
  return mode;
}

// zero the SOGM grids of agents [a0, a0+na) with the TMA bulk-store kernel (the particle map's
// accumulators are cleared once per update, dsp_dynamic.h:461-465)
int popk_grid_clear(pop_handle* h, int a0, int na) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_clear_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, kClearTile);
    attr_set = true;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  const unsigned long long total = h->geom.agent_stride * static_cast<unsigned long long>(na);
  const unsigned long long tiles = (total + kClearTile - 1) / kClearTile;
  const unsigned grid = static_cast<unsigned>(tiles < static_cast<unsigned long long>(sms) * 4 ? tiles : sms * 4);
  k_clear_tma<<<grid, 128, kClearTile, h->stream>>>(h->d_grid + h->geom.agent_stride * a0, total);
  POP_LAUNCH_CHECK(h);
  return POP_OK;
}

int popk_voxel_update(pop_handle* h) {
  const PopGeom& g = h->geom;
  const int nl = h->cfg.n_agents_local;
  const unsigned long long total = g.agent_stride * nl;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  pop_prof_begin(h, POP_STAGE_VOXEL_UPDATE);
  int launches = 0;
  if (vox_tiles_ok(h)) {
    const int rc = vox_update_tiles(h, sms, &launches);
    if (rc) return rc;
    pop_prof_end(h, POP_STAGE_VOXEL_UPDATE, launches);
    return POP_OK;
  }
  POP_CUDA(h, cudaMemsetAsync(h->d_occ_count, 0, sizeof(int) * nl, h->stream));
  const int mode = clear_mode();
  if (mode == 0) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(k_clear_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, kClearTile);
      attr_set = true;
    }
    k_clear_tma<<<sms * 4, 128, kClearTile, h->stream>>>(h->d_grid, total);
  } else if (mode == 1) {
    k_clear_vec<<<sms * 8, 256, 0, h->stream>>>(reinterpret_cast<uint4*>(h->d_grid), total / 16);
  } else {
    POP_CUDA(h, cudaMemsetAsync(h->d_grid, 0, total, h->stream));
  }
  POP_LAUNCH_CHECK(h);
  launches++;
  if (h->n_cloud > 0) {
    dim3 grid((h->n_cloud + 255) / 256, nl);
    if (h->cfg.cell_dtype == POP_CELL_U8)
      k_mark_slice0<uint8_t><<<grid, 256, 0, h->stream>>>(g, h->d_grid, h->d_cloud, h->n_cloud, h->d_poses,
                                                          h->d_occ_list, h->d_occ_count,
                                                          h->cfg.max_cloud_points);
    else
      k_mark_slice0<float><<<grid, 256, 0, h->stream>>>(g, h->d_grid, h->d_cloud, h->n_cloud, h->d_poses,
                                                        h->d_occ_list, h->d_occ_count,
                                                        h->cfg.max_cloud_points);
    POP_LAUNCH_CHECK(h);
    launches++;
    // occupied voxels <= cloud points; size the grid for the worst case
    dim3 grid2(min((h->n_cloud + 127) / 128, sms * 8), nl);
    const size_t smem = sizeof(PopCylDev) * (h->n_cyl > 0 ? h->n_cyl : 1);
    if (h->cfg.cell_dtype == POP_CELL_U8)
      k_future_slices<uint8_t><<<grid2, 128, smem, h->stream>>>(g, h->d_grid, h->d_poses, h->d_cyl, h->n_cyl,
                                                                h->d_occ_list, h->d_occ_count,
                                                                h->cfg.max_cloud_points);
    else
      k_future_slices<float><<<grid2, 128, smem, h->stream>>>(g, h->d_grid, h->d_poses, h->d_cyl, h->n_cyl,
                                                              h->d_occ_list, h->d_occ_count,
                                                              h->cfg.max_cloud_points);
    POP_LAUNCH_CHECK(h);
    launches++;
  }
  pop_prof_end(h, POP_STAGE_VOXEL_UPDATE, launches);
  return POP_OK;
}

int popk_peer_project(pop_handle* h) {
  if (!h->have_body || !h->have_traj) return POP_OK;  // is_multi_agents_ == false
  const PopGeom& g = h->geom;
  const int nl = h->cfg.n_agents_local, ntot = h->cfg.n_agents_total;
  pop_prof_begin(h, POP_STAGE_PEERS);
  k_peer_positions<<<(ntot * g.nt + 127) / 128, 128, 0, h->stream>>>(
      ntot, g.nt, g.dt, h->t_map, h->cfg.replan_risk_rate, h->d_traj_all, h->d_body_count, h->d_peer_pos, h->d_peer_valid,
      h->d_peer_sigma);
  POP_LAUNCH_CHECK(h);
  const long long n = static_cast<long long>(nl) * ntot * g.nt * POP_MAX_BODY;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  if (h->cfg.cell_dtype == POP_CELL_U8)
    k_peer_scatter<uint8_t><<<blocks, 256, 0, h->stream>>>(g, h->d_grid, h->d_poses, nl, ntot,
                                                          h->cfg.agent_id_base, h->d_peer_pos,
                                                          h->d_peer_valid, h->d_body, h->d_body_count, h->d_agent_ok,
                                                          h->d_peer_sigma, h->cfg.num_resample, h->d_noise, h->n_noise);
  else
    k_peer_scatter<float><<<blocks, 256, 0, h->stream>>>(g, h->d_grid, h->d_poses, nl, ntot,
                                                        h->cfg.agent_id_base, h->d_peer_pos,
                                                        h->d_peer_valid, h->d_body, h->d_body_count, h->d_agent_ok,
                                                          h->d_peer_sigma, h->cfg.num_resample, h->d_noise, h->n_noise);
  POP_LAUNCH_CHECK(h);
  pop_prof_end(h, POP_STAGE_PEERS, 2);
  return POP_OK;
}

int popk_query(pop_handle* h, int agent, const double* d_pos, const int32_t* d_t, const double* d_dt,
               int n, int8_t* d_out) {
  const PopGeom& g = h->geom;
  const uint8_t* ag = h->d_grid + agent * g.agent_stride;
  const float* pose = h->d_poses + 3 * agent;
  if (h->cfg.cell_dtype == POP_CELL_U8)
    k_query<uint8_t><<<(n + 127) / 128, 128, 0, h->stream>>>(g, ag, pose, d_pos, d_t, d_dt, n, d_out);
  else
    k_query<float><<<(n + 127) / 128, 128, 0, h->stream>>>(g, ag, pose, d_pos, d_t, d_dt, n, d_out);
  POP_LAUNCH_CHECK(h);
  h->n_launches++;
  return POP_OK;
}

int popk_query_inflate(pop_handle* h, int agent, const double* d_pos, const double* d_t, int mode, int n, int8_t* d_out) {
  const PopGeom& g = h->geom;
  const uint8_t* ag = h->d_grid + agent * g.agent_stride;
  const float* pose = h->d_poses + 3 * agent;
  if (h->cfg.cell_dtype == POP_CELL_U8)
    k_query_inflate<uint8_t><<<(n + 127) / 128, 128, 0, h->stream>>>(g, ag, pose, d_pos, d_t, mode, n, d_out);
  else
    k_query_inflate<float><<<(n + 127) / 128, 128, 0, h->stream>>>(g, ag, pose, d_pos, d_t, mode, n, d_out);
  POP_LAUNCH_CHECK(h);
  h->n_launches++;
  return POP_OK;
}

// slice range and voxel box of getObstaclePoints, in the reference's double arithmetic
// (map.cpp:469-492); slice indices clamped to nt-1 (the reference reads one slice past the end)
static BoxSpec make_box(const PopGeom& g, const float pose[3], double t_map, double t_start,
                        double t_end, const double* lc, const double* hc) {
  BoxSpec b;
  int i0 = static_cast<int>(std::floor((t_start - t_map) / g.dt));
  int i1 = static_cast<int>(std::ceil((t_end - t_map) / g.dt));
  i0 = i0 < 0 ? 0 : i0;
  i0 = i0 > g.nt - 1 ? g.nt - 1 : i0;
  i1 = i1 > g.nt - 1 ? g.nt - 1 : i1;
  i1 = i1 < 0 ? 0 : i1;
  b.j0 = i0; b.j1 = i1;
  b.lx = static_cast<int>((lc[0] - pose[0] + g.rx) / g.res);
  b.ly = static_cast<int>((lc[1] - pose[1] + g.ry) / g.res);
  b.lz = static_cast<int>((lc[2] - pose[2] + g.rz) / g.res);
  b.hx = static_cast<int>((hc[0] - pose[0] + g.rx) / g.res);
  b.hy = static_cast<int>((hc[1] - pose[1] + g.ry) / g.res);
  b.hz = static_cast<int>((hc[2] - pose[2] + g.rz) / g.res);
  b.hx = std::min(b.hx, g.nx - 1); b.hy = std::min(b.hy, g.ny - 1); b.hz = std::min(b.hz, g.nz - 1);
  b.lx = std::max(b.lx, 0); b.ly = std::max(b.ly, 0); b.lz = std::max(b.lz, 0);
  return b;
}

int popk_obstacle_points(pop_handle* h, int agent, double t_start, double t_end, const double* lc,
                         const double* hc, double* d_out, int cap, int* d_count) {
  const PopGeom& g = h->geom;
  float pose[3];
  POP_CUDA(h, cudaMemcpyAsync(pose, h->d_poses + 3 * agent, sizeof(pose), cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  BoxSpec b = make_box(g, pose, h->t_map, t_start, t_end, lc, hc);
  const uint8_t* ag = h->d_grid + agent * g.agent_stride;
  if (h->cfg.cell_dtype == POP_CELL_U8)
    k_obstacle_points<uint8_t><<<1, 1024, 0, h->stream>>>(g, ag, h->d_poses + 3 * agent, b, d_out, cap, d_count);
  else
    k_obstacle_points<float><<<1, 1024, 0, h->stream>>>(g, ag, h->d_poses + 3 * agent, b, d_out, cap, d_count);
  POP_LAUNCH_CHECK(h);
  h->n_launches++;
  return POP_OK;
}

int popk_upload(pop_handle* h, int agent, const float* d_in) {
  const PopGeom& g = h->geom;
  uint8_t* ag = h->d_grid + agent * g.agent_stride;
  if (h->cfg.cell_dtype == POP_CELL_U8)
    k_upload<uint8_t><<<1184, 256, 0, h->stream>>>(g, ag, d_in);
  else
    k_upload<float><<<1184, 256, 0, h->stream>>>(g, ag, d_in);
  POP_LAUNCH_CHECK(h);
  h->n_launches++;
  return POP_OK;
}

int popk_add_particles(pop_handle* h, int agent, const double* d_pts, const float* d_risks, int n,
                       int t_index) {
  const PopGeom& g = h->geom;
  uint8_t* ag = h->d_grid + agent * g.agent_stride;
  if (n <= 0) return POP_OK;
  if (h->cfg.cell_dtype == POP_CELL_U8)
    k_add_particles<uint8_t><<<(n + 127) / 128, 128, 0, h->stream>>>(g, ag, d_pts, d_risks, n, t_index);
  else
    k_add_particles<float><<<(n + 127) / 128, 128, 0, h->stream>>>(g, ag, d_pts, d_risks, n, t_index);
  POP_LAUNCH_CHECK(h);
  h->n_launches++;
  return POP_OK;
}

int popk_download(pop_handle* h, int agent, float* d_out) {
  const PopGeom& g = h->geom;
  const uint8_t* ag = h->d_grid + agent * g.agent_stride;
  if (h->cfg.cell_dtype == POP_CELL_U8)
    k_download<uint8_t><<<1184, 256, 0, h->stream>>>(g, ag, d_out);
  else
    k_download<float><<<1184, 256, 0, h->stream>>>(g, ag, d_out);
  POP_LAUNCH_CHECK(h);
  h->n_launches++;
  return POP_OK;
}
