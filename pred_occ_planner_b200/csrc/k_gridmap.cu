// k_gridmap.cu — the log-odds raycast map, batched over maps.
//
// Reference: plan_env/src/grid_map.cpp:192-208 (setCacheOccupancy), :313-456 (raycastProcess), :458-475
// (closetPointInMap), :477-583 (clearAndInflateLocalMap); plan_env/src/raycast.cpp:17-31, 242-335 (RayCaster);
// plan_env/include/plan_env/grid_map.h:261-420.
//
// raycastProcess is sequential in the projected points: only the FIRST point ending in a voxel casts a ray
// (flag_rayend_), and a ray stops at the first voxel an EARLIER ray of the same frame traversed
// (flag_traverse_) — after counting a miss there.  Both are reproduced exactly:
//   * ray ends: atomicMin of the point index per end voxel;
//   * traversal: one CTA per map walks the rays in index order.  The casting points are compacted in order and 128
//     rays at a time are traced in parallel
//     (Amanatides-Woo DDA in the reference's double arithmetic) into shared-memory path buffers; one warp
//     then commits them in order against a shared-memory bitmap of the voxels traversed so far (the rays of a
//     frame live inside a (2R+1)^3 box around the camera, R = max_ray_length / resolution), 32 path voxels
//     per step with a ballot for the first already-traversed one.
//   * the reference keeps both flags in vector<char> and its frame counter raycast_num_ is a char as well
//     (grid_map.h:139-140): the counter wraps 127 -> -128, and a stamp written 256 frames ago — or the initial -1 of a
//     never-touched voxel, at frame 255 — reads as "already done in this frame".  Kept: persistent per-voxel stamps
//     are preloaded into the bitmap when such a match is possible and written back after every frame.
// Hit/miss counters are integer atomics; every touched voxel receives exactly one log-odds update per frame,
// so occupancy_buffer_ is bit-identical to the sequential result.
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/pop.h"

namespace {

struct GmGeo {
  double res, res_inv, origin[3], bmin[3], bmax[3], lur[3];
  double prob_hit, prob_miss, clamp_min, clamp_max, min_occ, unk;
  double max_ray, ceil_h, ground;
  int n[3], N, margin, inf_step, max_pts, R, B;   // B = 2R+1 bitmap side
};

constexpr int kChunk = 64;      // rays traced in parallel per commit round
constexpr int kMaxPath = 192;   // voxels per ray (<= 3 R + 3)

__device__ __forceinline__ unsigned long long d2ord(double d) {
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(d));
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord2d(unsigned long long u) {
  return __longlong_as_double(static_cast<long long>((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u));
}
__device__ __forceinline__ int gm_addr(const GmGeo& G, const int* id) { return id[0] * G.n[1] * G.n[2] + id[1] * G.n[2] + id[2]; }
__device__ __forceinline__ void gm_pos_to_index(const GmGeo& G, const double* p, int* id) {
  for (int i = 0; i < 3; i++) id[i] = static_cast<int>(floor((p[i] - G.origin[i]) * G.res_inv));
}
__device__ __forceinline__ void gm_bound(const GmGeo& G, int* id) {
  for (int i = 0; i < 3; i++) id[i] = max(min(id[i], G.n[i] - 1), 0);
}
__device__ __forceinline__ double gm_mod(double v, double m) { return fmod(fmod(v, m) + m, m); }
__device__ __forceinline__ double gm_intbound(double s, double ds) {
  if (ds < 0) { s = -s; ds = -ds; }
  s = gm_mod(s, 1.0);
  return (1 - s) / ds;
}

// setCacheOccupancy :192-208: count, remember first touches
__device__ __forceinline__ void gm_count(int a, int occ, int* cnt_hm, int* cnt_hit, int* touched, int* ntouched) {
  if (atomicAdd(&cnt_hm[a], 1) == 0) touched[atomicAdd(ntouched, 1)] = a;
  if (occ) atomicAdd(&cnt_hit[a], 1);
}

// per map state pointers
struct GmMap {
  double* occ; signed char* inflate; int* cnt_hm; int* cnt_hit; int* first_end; int* touched; int* ntouched;
  double* pw; int* endv; unsigned long long* bbox;   // bbox: min xyz, max xyz as ordered keys
  int* lb;   // local_bound_min (3), local_bound_max (3)
  int* cast; // indices of the points that cast a ray this frame, in order
  signed char* flag_rayend;    // the reference's per-voxel frame stamps (grid_map.h:139), persistent, initial -1
  signed char* flag_traverse;
};

// K1: clip every projected point, count its end voxel, elect the casting point per end voxel (:337-377)
__global__ void __launch_bounds__(256) k_gm_points(GmGeo G, const GmMap* __restrict__ maps, const double* __restrict__ pts_all,
                                                   const int* __restrict__ counts, const double* __restrict__ cam_all,
                                                   const int* __restrict__ flags_work, int* __restrict__ err) {
  const int m = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= counts[m]) return;
  const GmMap M = maps[m];
  const double* cam = cam_all + 3 * m;
  const double* p = pts_all + (static_cast<size_t>(m) * G.max_pts + i) * 3;
  double pw[3] = {p[0], p[1], p[2]};
  bool inmap = !(pw[0] < G.bmin[0] + 1e-4 || pw[1] < G.bmin[1] + 1e-4 || pw[2] < G.bmin[2] + 1e-4 ||
                 pw[0] > G.bmax[0] - 1e-4 || pw[1] > G.bmax[1] - 1e-4 || pw[2] > G.bmax[2] - 1e-4);
  int occ = 1;
  if (!inmap) {   // closetPointInMap :458-475
    double diff[3];
    double min_t = 1000000;
    for (int k = 0; k < 3; ++k) {
      diff[k] = pw[k] - cam[k];
      if (fabs(diff[k]) > 0) {
        const double t1 = (G.bmax[k] - cam[k]) / diff[k];
        if (t1 > 0 && t1 < min_t) min_t = t1;
        const double t2 = (G.bmin[k] - cam[k]) / diff[k];
        if (t2 > 0 && t2 < min_t) min_t = t2;
      }
    }
    for (int k = 0; k < 3; k++) pw[k] = cam[k] + (min_t - 1e-3) * diff[k];
    occ = 0;
  }
  {
    const double d0 = pw[0] - cam[0], d1 = pw[1] - cam[1], d2 = pw[2] - cam[2];
    const double len = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    if (len > G.max_ray) {
      pw[0] = d0 / len * G.max_ray + cam[0];
      pw[1] = d1 / len * G.max_ray + cam[1];
      pw[2] = d2 / len * G.max_ray + cam[2];
      occ = 0;
    }
  }
  if (!(isfinite(pw[0]) && isfinite(pw[1]) && isfinite(pw[2]))) {   // depth-derived points are finite; refuse others
    atomicOr(err, 4);
    M.endv[i] = -2;
    return;
  }
  int id[3];
  gm_pos_to_index(G, pw, id);
  int a = gm_addr(G, id);
  // unchecked linear addressing as in the reference (an index out of range along one axis aliases a neighbouring
  // row); an address outside the buffers is undefined behaviour there and is skipped
  if (a < 0 || a >= G.N) a = -1;
  if (a >= 0) gm_count(a, occ, M.cnt_hm, M.cnt_hit, M.touched, M.ntouched);
  for (int k = 0; k < 3; k++) {
    M.pw[(static_cast<size_t>(i)) * 3 + k] = pw[k];
    atomicMin(&M.bbox[k], d2ord(pw[k]));
    atomicMax(&M.bbox[3 + k], d2ord(pw[k]));
  }
  M.endv[i] = a;
  if (a >= 0) atomicMin(&M.first_end[a], i);   // the first point of an end voxel is the one that may cast (:370-377)
}

// DDA of one ray into a path of voxel coordinates relative to the bitmap box (packed 3 x 10 bits); returns length
__device__ int gm_trace(const GmGeo& G, const double* pw, const double* cam, const int* box0, unsigned* path, int cap, bool* bad) {
  double s[3], e[3], d[3], tmax[3], tdelta[3];
  int x[3], ex[3], st[3];
  for (int k = 0; k < 3; k++) {
    s[k] = pw[k] / G.res; e[k] = cam[k] / G.res;
    x[k] = static_cast<int>(floor(s[k]));
    ex[k] = static_cast<int>(floor(e[k]));
    d[k] = ex[k] - x[k];
    const int di = static_cast<int>(d[k]);
    st[k] = di == 0 ? 0 : di < 0 ? -1 : 1;
    tmax[k] = gm_intbound(s[k], d[k]);
    tdelta[k] = static_cast<double>(st[k]) / d[k];
  }
  int len = 0, steps = 0;
  while (!(x[0] == ex[0] && x[1] == ex[1] && x[2] == ex[2])) {
    if (++steps > 4 * cap) { *bad = true; return len; }   // non-finite input: the reference's loop would not end
    // the voxel whose centre ((ray + 0.5) * res) the reference hands to setCacheOccupancy
    int id[3];
    for (int k = 0; k < 3; k++) {
      const double c = (static_cast<double>(x[k]) + 0.5) * G.res;
      id[k] = static_cast<int>(floor((c - G.origin[k]) * G.res_inv));
    }
    // the reference addresses linearly without range checks: take the voxel that address really is
    const int a = gm_addr(G, id);
    if (a >= 0 && a < G.N) {
      // an index inside the map is its own address decomposition (no divisions on the common path)
      const bool inside = id[0] >= 0 && id[0] < G.n[0] && id[1] >= 0 && id[1] < G.n[1] && id[2] >= 0 && id[2] < G.n[2];
      const int rx = (inside ? id[0] : a / (G.n[1] * G.n[2])) - box0[0];
      const int ry = (inside ? id[1] : (a / G.n[2]) % G.n[1]) - box0[1];
      const int rz = (inside ? id[2] : a % G.n[2]) - box0[2];
      if (len >= cap || rx < 0 || ry < 0 || rz < 0 || rx >= G.B || ry >= G.B || rz >= G.B) { *bad = true; return len; }
      path[len++] = static_cast<unsigned>(rx) | (static_cast<unsigned>(ry) << 10) | (static_cast<unsigned>(rz) << 20);
    }
    if (tmax[0] < tmax[1]) {
      if (tmax[0] < tmax[2]) { x[0] += st[0]; tmax[0] += tdelta[0]; } else { x[2] += st[2]; tmax[2] += tdelta[2]; }
    } else {
      if (tmax[1] < tmax[2]) { x[1] += st[1]; tmax[1] += tdelta[1]; } else { x[2] += st[2]; tmax[2] += tdelta[2]; }
    }
  }
  return len;
}

// K2: rays in index order (one CTA per map)
__global__ void __launch_bounds__(256) k_gm_rays(GmGeo G, const GmMap* __restrict__ maps, const int* __restrict__ counts,
                                                 const double* __restrict__ cam_all, const int* __restrict__ flags_work,
                                                 int* __restrict__ err, int chunk) {
  extern __shared__ unsigned smem[];
  const int m = blockIdx.x, tid = threadIdx.x;
  // `char raycast_num_` (grid_map.h:140): the frame counter wraps 127 -> -128, and the per-voxel stamps it is compared
  // with are chars too, so a stamp written 256 frames ago — or the initial -1 of a never-touched voxel, at frame 255 —
  // reads as "already done in this frame".  Within a frame the stamps are the bitmap below; the persistent copies are
  // consulted when such a stale match is possible (frame >= 255) and updated at the end of every frame.
  const int frame = flags_work[m];
  const signed char num = static_cast<signed char>(static_cast<unsigned char>(frame & 0xff));
  const GmMap M = maps[m];
  const int np = counts[m];
  if (np == 0) return;
  const double cam[3] = {cam_all[3 * m], cam_all[3 * m + 1], cam_all[3 * m + 2]};
  int camv[3], box0[3];
  gm_pos_to_index(G, cam, camv);
  for (int k = 0; k < 3; k++) box0[k] = camv[k] - G.R;
  const int nwords = (G.B * G.B * G.B + 31) / 32;
  unsigned* bitmap = smem;                       // nwords
  unsigned* paths = smem + nwords;               // chunk * kMaxPath
  int* plen = reinterpret_cast<int*>(paths + chunk * kMaxPath);   // chunk
  __shared__ int s_wsum[8], s_ncast;
  for (int i = tid; i < nwords; i += blockDim.x) bitmap[i] = 0u;
  __syncthreads();
  if (frame >= 255) {   // stale traverse stamps inside the box count as traversed
    for (int q = tid; q < G.B * G.B; q += blockDim.x) {
      const int rx = q / G.B, ry = q % G.B, x = rx + box0[0], y = ry + box0[1];
      if (x < 0 || x >= G.n[0] || y < 0 || y >= G.n[1]) continue;
      for (int rz = 0; rz < G.B; rz++) {
        const int z = rz + box0[2];
        if (z < 0 || z >= G.n[2]) continue;
        if (M.flag_traverse[(x * G.n[1] + y) * G.n[2] + z] == num) {
          const int lin = (rx * G.B + ry) * G.B + rz;
          atomicOr(&bitmap[lin >> 5], 1u << (lin & 31));
        }
      }
    }
  }
  // the points that cast a ray (not clipped away; only the first point of every end voxel, and none if the voxel's
  // ray-end stamp already equals this frame's number), compacted
  // in index order, so that every trace thread of a round below has a ray to walk
  if (tid == 0) s_ncast = 0;
  __syncthreads();
  for (int w0 = 0; w0 < np; w0 += blockDim.x) {
    const int i = w0 + tid;
    bool casts = i < np && M.endv[i] != -2;
    if (casts && M.endv[i] >= 0) {
      const int a = M.endv[i];
      casts = M.first_end[a] == i && M.flag_rayend[a] != num;
      if (casts) M.flag_rayend[a] = num;   // only the elected point of a voxel reads and writes its stamp
    }
    const unsigned bal = __ballot_sync(0xffffffffu, casts);
    if ((tid & 31) == 0) s_wsum[tid >> 5] = __popc(bal);
    __syncthreads();
    int off = s_ncast;
    for (int w = 0; w < (tid >> 5); w++) off += s_wsum[w];
    if (casts) M.cast[off + __popc(bal & ((1u << (tid & 31)) - 1u))] = i;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < 8; w++) t += s_wsum[w]; s_ncast += t; }
    __syncthreads();
  }
  const int nc = s_ncast;
  for (int base = 0; base < nc; base += chunk) {
    // trace: thread r < chunk walks the r-th casting ray of this round
    if (tid < chunk) {
      int len = -1;
      if (base + tid < nc) {
        const int i = M.cast[base + tid];
        bool bad = false;
        len = gm_trace(G, M.pw + static_cast<size_t>(i) * 3, cam, box0, paths + tid * kMaxPath, kMaxPath, &bad);
        if (bad) atomicOr(err, 1);
      }
      plen[tid] = len;
    }
    __syncthreads();
    {
      // commit in index order: warp 0, 32 path voxels per step.  Only the bitmap is order-dependent, so the walk
      // touches shared memory alone and leaves, per ray, how many of its voxels count (up to and including the one
      // that stops it); the counters in global memory are then updated by all threads at once — with the atomics
      // inside the walk every step waited for a global round trip.
      if (tid < 32) {
        for (int r = 0; r < chunk; r++) {
          const int len = plen[r];
          int keep = len < 0 ? 0 : len;
          bool stop = false;
          for (int s0 = 0; s0 < len && !stop; s0 += 32) {
            const int s = s0 + tid;
            int bit = 0, lin = 0;
            if (s < len) {
              const unsigned pk = paths[r * kMaxPath + s];
              const int rx = pk & 1023, ry = (pk >> 10) & 1023, rz = pk >> 20;
              lin = (rx * G.B + ry) * G.B + rz;
              bit = (bitmap[lin >> 5] >> (lin & 31)) & 1;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, bit);
            const int f = hit ? __ffs(hit) - 1 : 32;      // first already-traversed voxel of this step
            if (s < len && tid < f) atomicOr(&bitmap[lin >> 5], 1u << (lin & 31));
            __syncwarp();
            if (f < 32) { stop = true; keep = s0 + f + 1; }   // a miss also on the voxel that stops the ray
          }
          if (tid == 0) plen[r] = keep;
        }
      }
      __syncthreads();
      for (int e = tid; e < chunk * kMaxPath; e += blockDim.x) {
        const int r = e / kMaxPath, s = e % kMaxPath;
        if (s < plen[r]) {
          const unsigned pk = paths[e];
          const int rx = pk & 1023, ry = (pk >> 10) & 1023, rz = pk >> 20;
          const int a = ((rx + box0[0]) * G.n[1] + (ry + box0[1])) * G.n[2] + (rz + box0[2]);
          gm_count(a, 0, M.cnt_hm, M.cnt_hit, M.touched, M.ntouched);
        }
      }
    }
    __syncthreads();
  }
  // every voxel traversed in this frame carries the frame's number from now on
  for (int w = tid; w < nwords; w += blockDim.x) {
    unsigned bits = bitmap[w];
    while (bits) {
      const int lin = w * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
      const int rz = lin % G.B, ry = (lin / G.B) % G.B, rx = lin / (G.B * G.B);
      M.flag_traverse[((rx + box0[0]) * G.n[1] + (ry + box0[1])) * G.n[2] + (rz + box0[2])] = num;
    }
  }
}

// K3: local bounds from the bounding box (:400-412), one log-odds update per touched voxel (:414-455)
__global__ void __launch_bounds__(256) k_gm_update(GmGeo G, const GmMap* __restrict__ maps, const int* __restrict__ counts,
                                                   const double* __restrict__ cam_all) {
  const int m = blockIdx.y;
  if (counts[m] == 0) return;
  const GmMap M = maps[m];
  const double* cam = cam_all + 3 * m;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double mn[3], mx[3];
    for (int k = 0; k < 3; k++) { mn[k] = fmin(ord2d(M.bbox[k]), cam[k]); mx[k] = fmax(ord2d(M.bbox[3 + k]), cam[k]); }
    mx[2] = fmax(mx[2], G.ground);
    int a[3], b[3];
    gm_pos_to_index(G, mn, a); gm_pos_to_index(G, mx, b);
    gm_bound(G, a); gm_bound(G, b);
    for (int k = 0; k < 3; k++) { M.lb[k] = a[k]; M.lb[3 + k] = b[k]; }
  }
  double lmin[3], lmax[3];
  int min_id[3], max_id[3];
  for (int k = 0; k < 3; k++) { lmin[k] = cam[k] - G.lur[k]; lmax[k] = cam[k] + G.lur[k]; }
  gm_pos_to_index(G, lmin, min_id); gm_pos_to_index(G, lmax, max_id);
  gm_bound(G, min_id); gm_bound(G, max_id);
  const int nt = *M.ntouched;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    const int a = M.touched[t];
    const int hm = M.cnt_hm[a], hit = M.cnt_hit[a];
    M.cnt_hm[a] = 0; M.cnt_hit[a] = 0; M.first_end[a] = INT_MAX;
    const double upd = hit >= hm - hit ? G.prob_hit : G.prob_miss;
    double o = M.occ[a];
    if (upd >= 0 && o >= G.clamp_max) continue;
    if (upd <= 0 && o <= G.clamp_min) { M.occ[a] = G.clamp_min; continue; }
    const int ix = a / (G.n[1] * G.n[2]), iy = (a / G.n[2]) % G.n[1], iz = a % G.n[2];
    const bool in_local = ix >= min_id[0] && ix <= max_id[0] && iy >= min_id[1] && iy <= max_id[1] && iz >= min_id[2] && iz <= max_id[2];
    if (!in_local) o = G.clamp_min;
    M.occ[a] = fmin(fmax(o + upd, G.clamp_min), G.clamp_max);
  }
}

// K4a: reset to unknown the shell between the cut box and its 5-voxel margin; clear the inflated buffer in the
// local bounds (:477-553)
__global__ void __launch_bounds__(256) k_gm_clear(GmGeo G, const GmMap* __restrict__ maps, const int* __restrict__ counts) {
  const int m = blockIdx.y;
  if (counts[m] == 0) return;
  const GmMap M = maps[m];
  int cut0[3], cut1[3], m0[3], m1[3];
  for (int k = 0; k < 3; k++) { cut0[k] = M.lb[k] - G.margin; cut1[k] = M.lb[3 + k] + G.margin; }
  gm_bound(G, cut0); gm_bound(G, cut1);
  for (int k = 0; k < 3; k++) { m0[k] = cut0[k] - 5; m1[k] = cut1[k] + 5; }
  gm_bound(G, m0); gm_bound(G, m1);
  const int sx = m1[0] - m0[0] + 1, sy = m1[1] - m0[1] + 1, sz = m1[2] - m0[2] + 1;
  const long long total = static_cast<long long>(sx) * sy * sz;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = m0[0] + static_cast<int>(e / (static_cast<long long>(sy) * sz));
    const int y = m0[1] + static_cast<int>((e / sz) % sy), z = m0[2] + static_cast<int>(e % sz);
    const int a = (x * G.n[1] + y) * G.n[2] + z;
    const bool inside_cut = x >= cut0[0] && x <= cut1[0] && y >= cut0[1] && y <= cut1[1] && z >= cut0[2] && z <= cut1[2];
    if (!inside_cut) M.occ[a] = G.unk;
    if (x >= M.lb[0] && x <= M.lb[3] && y >= M.lb[1] && y <= M.lb[4] && z >= M.lb[2] && z <= M.lb[5]) M.inflate[a] = 0;
  }
}

// K4b: inflate the occupied voxels of the local bounds, virtual ceiling (:555-583)
__global__ void __launch_bounds__(256) k_gm_inflate(GmGeo G, const GmMap* __restrict__ maps, const int* __restrict__ counts) {
  const int m = blockIdx.y;
  if (counts[m] == 0) return;
  const GmMap M = maps[m];
  const int sx = M.lb[3] - M.lb[0] + 1, sy = M.lb[4] - M.lb[1] + 1, sz = M.lb[5] - M.lb[2] + 1;
  const long long total = static_cast<long long>(sx) * sy * sz;
  const int ceil_id = static_cast<int>(floor((G.ceil_h - G.origin[2]) * G.res_inv)) - 1;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = M.lb[0] + static_cast<int>(e / (static_cast<long long>(sy) * sz));
    const int y = M.lb[1] + static_cast<int>((e / sz) % sy), z = M.lb[2] + static_cast<int>(e % sz);
    if (M.occ[(x * G.n[1] + y) * G.n[2] + z] > G.min_occ) {
      for (int dx = -G.inf_step; dx <= G.inf_step; ++dx)
        for (int dy = -G.inf_step; dy <= G.inf_step; ++dy)
          for (int dz = -G.inf_step; dz <= G.inf_step; ++dz) {
            const int a = ((x + dx) * G.n[1] + (y + dy)) * G.n[2] + (z + dz);   // only the linear address is checked
            if (a < 0 || a >= G.N) continue;
            M.inflate[a] = 1;
          }
    }
    if (G.ceil_h > -0.5 && z == M.lb[2]) {
      const int a = (x * G.n[1] + y) * G.n[2] + ceil_id;
      if (a >= 0 && a < G.N) M.inflate[a] = 1;
    }
  }
}

// projectDepthImage :210-311 for every map: one CTA walks the sampled pixels in the reference's (v, u) order and
// compacts the projected points in that order (the order is the order the rays are cast in).  Quirks kept: without
// the filter the row pointer advances by one per sampled pixel whatever skip_pixel is; with the filter the zero
// test reads the pixel after the pointer advanced by skip_pixel (non-zero past the end of the image).
__global__ void __launch_bounds__(1024) k_gm_project(pop_depth_params P, double max_ray, const unsigned short* __restrict__ img_all,
                                                     const double* __restrict__ R_all, const double* __restrict__ cam_all,
                                                     const int* __restrict__ active, double* __restrict__ pts_all,
                                                     int max_pts, int* __restrict__ counts, int* __restrict__ err) {
  __shared__ int s_w[33];
  const int m = blockIdx.x, tid = threadIdx.x;
  if (!active[m]) { if (tid == 0) counts[m] = 0; return; }
  const long long total = static_cast<long long>(P.rows) * P.cols;
  const unsigned short* img = img_all + static_cast<size_t>(m) * total;
  const double* R = R_all + 9 * m;
  const double* cam = cam_all + 3 * m;
  double* out = pts_all + static_cast<size_t>(m) * max_pts * 3;
  const int mg = P.use_depth_filter ? P.depth_filter_margin : 0;
  const int nu = P.cols - 2 * mg > 0 ? (P.cols - 2 * mg + P.skip_pixel - 1) / P.skip_pixel : 0;
  const int nv = P.rows - 2 * mg > 0 ? (P.rows - 2 * mg + P.skip_pixel - 1) / P.skip_pixel : 0;
  const int nsamp = nu * nv;
  const double inv_factor = 1.0 / P.k_depth_scaling_factor;
  int carry = 0;
  for (int base = 0; base < nsamp; base += blockDim.x) {
    const int s = base + tid;
    bool keep = false;
    double x = 0, y = 0, z = 0;
    if (s < nsamp) {
      const int jv = s / nu, ju = s % nu;
      const int v = mg + jv * P.skip_pixel, u = mg + ju * P.skip_pixel;
      double depth;
      if (!P.use_depth_filter) {
        depth = img[static_cast<long long>(v) * P.cols + ju] / P.k_depth_scaling_factor;
        keep = true;
      } else {
        const long long ptr = static_cast<long long>(v) * P.cols + u;
        depth = img[ptr] * inv_factor;
        const unsigned next = ptr + P.skip_pixel < total ? img[ptr + P.skip_pixel] : 1u;
        keep = true;
        if (next == 0) depth = max_ray + 0.1;
        else if (depth < P.depth_filter_mindist) keep = false;
        else if (depth > P.depth_filter_maxdist) depth = max_ray + 0.1;
      }
      x = (u - P.cx) * depth / P.fx; y = (v - P.cy) * depth / P.fy; z = depth;
    }
    // ordered compaction (block-wide exclusive scan)
    const int lane = tid & 31, w = tid >> 5;
    int inc = keep ? 1 : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
      int xv = s_w[lane], xi = xv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
      s_w[lane] = xi - xv;
      if (lane == 31) s_w[32] = xi;
    }
    __syncthreads();
    const int pos = carry + inc - (keep ? 1 : 0) + s_w[w];
    if (keep && pos < max_pts)
      for (int i = 0; i < 3; i++) out[3 * pos + i] = ((R[3 * i] * x + R[3 * i + 1] * y) + R[3 * i + 2] * z) + cam[i];
    carry += s_w[32];
  }
  if (tid == 0) {
    if (carry > max_pts) atomicOr(err, 2);
    counts[m] = carry < max_pts ? carry : max_pts;
  }
}

__global__ void k_gm_frame_reset(const GmMap* __restrict__ maps, int n_maps) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_maps) return;
  *maps[m].ntouched = 0;
  for (int k = 0; k < 3; k++) { maps[m].bbox[k] = ~0ull; maps[m].bbox[3 + k] = 0ull; }
}

__global__ void k_gm_fill(double* occ, size_t n, double v, int* first_end) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) { occ[i] = v; first_end[i] = INT_MAX; }
}

__global__ void k_gm_query(GmGeo G, GmMap M, const double* __restrict__ pos, int n, signed char* __restrict__ out_occ,
                           signed char* __restrict__ out_inf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double p[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
  const bool inmap = !(p[0] < G.bmin[0] + 1e-4 || p[1] < G.bmin[1] + 1e-4 || p[2] < G.bmin[2] + 1e-4 ||
                       p[0] > G.bmax[0] - 1e-4 || p[1] > G.bmax[1] - 1e-4 || p[2] > G.bmax[2] - 1e-4);
  if (!inmap) { out_occ[i] = -1; out_inf[i] = -1; return; }
  int id[3];
  gm_pos_to_index(G, p, id);
  const int a = gm_addr(G, id);
  out_occ[i] = M.occ[a] > G.min_occ ? 1 : 0;
  out_inf[i] = M.inflate[a];
}

}  // namespace

struct pop_gridmap {
  pop_gridmap_config cfg;
  GmGeo G;
  int n_maps = 0, device = 0;
  std::vector<int> frame;        // md_.raycast_num_ of every map (advances only on frames with points, :316-320)
  pop_depth_params depth;        // projectDepthImage parameters
  bool have_depth = false;
  std::vector<int> has_first;    // md_.has_first_depth_ per map
  std::vector<int32_t> last_counts;
  unsigned short* d_img = nullptr;
  double* d_R = nullptr;
  int* d_flags = nullptr;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  std::vector<GmMap> h_maps;
  GmMap* d_maps = nullptr;
  double* d_pts = nullptr; double* d_cam = nullptr; int* d_counts = nullptr; int* d_err = nullptr;
  void* d_q = nullptr; size_t d_q_bytes = 0;
  std::vector<void*> allocs;
  size_t rays_smem = 0;
  int rays_chunk = 0;
};

// see PopDeviceGuard (pop_internal.h): the map's device is current inside every entry point, the caller's afterwards
struct GmDeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit GmDeviceGuard(const pop_gridmap* g) {
    if (g && cudaGetDevice(&prev) == cudaSuccess && prev != g->device) switched = cudaSetDevice(g->device) == cudaSuccess;
  }
  ~GmDeviceGuard() { if (switched) cudaSetDevice(prev); }
};

static int gm_fail(pop_gridmap* g, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (g) g->err = buf;
  return code;
}
#define GM_CUDA(g, expr)                                                                                           \
  do {                                                                                                             \
    cudaError_t _e = (expr);                                                                                       \
    if (_e != cudaSuccess) return gm_fail((g), POP_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                          __FILE__, __LINE__);                                                     \
  } while (0)

extern "C" {

void pop_gridmap_default_config(pop_gridmap_config* c) {
  memset(c, 0, sizeof(*c));
  // plan_env/launch + grid_map.cpp:20-63 defaults as used by the reference's simulator launch files
  c->resolution = 0.1;
  c->map_size[0] = 40.0; c->map_size[1] = 40.0; c->map_size[2] = 5.0;
  c->local_update_range[0] = 5.5; c->local_update_range[1] = 5.5; c->local_update_range[2] = 4.5;
  c->obstacles_inflation = 0.099;
  c->p_hit = 0.70; c->p_miss = 0.35; c->p_min = 0.12; c->p_max = 0.97; c->p_occ = 0.80;
  c->max_ray_length = 4.5;
  c->virtual_ceil_height = 2.5;
  c->ground_height = -0.01;
  c->local_map_margin = 30;
  c->max_points = 20000;
}

const char* pop_gridmap_last_error(const pop_gridmap* g) { return g ? g->err.c_str() : "null gridmap"; }

int pop_gridmap_destroy(pop_gridmap* g) {
  if (!g) return POP_OK;
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  cudaSetDevice(g->device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  for (void* p : g->allocs) cudaFree(p);
  if (g->d_q) cudaFree(g->d_q);
  if (g->d_img) cudaFree(g->d_img);
  if (g->d_R) cudaFree(g->d_R);
  if (g->own_stream && g->stream) cudaStreamDestroy(g->stream);
  delete g;
  if (prev >= 0) cudaSetDevice(prev);
  return POP_OK;
}

int pop_gridmap_dims(const pop_gridmap* g, int32_t* out3) {
  if (!g || !out3) return POP_E_INVALID;
  for (int k = 0; k < 3; k++) out3[k] = g->G.n[k];
  return POP_OK;
}

int pop_gridmap_create(const pop_gridmap_config* c, int n_maps, int device, void* stream, pop_gridmap** out) {
  if (!c || !out || n_maps < 1 || !(c->resolution > 0) || c->max_points < 1) return POP_E_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) { cudaGetLastError(); return POP_E_NO_DEVICE; }
  pop_gridmap* g = new pop_gridmap();
  g->cfg = *c; g->n_maps = n_maps; g->device = device;
  struct Restore {   // leave the caller's current device as it was
    int prev = -1;
    Restore() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~Restore() { if (prev >= 0) cudaSetDevice(prev); }
  } restore;
  auto fail = [&](int code) { pop_gridmap_destroy(g); return code; };
  if (cudaSetDevice(device) != cudaSuccess) return fail(POP_E_NO_DEVICE);
  if (stream) g->stream = static_cast<cudaStream_t>(stream);
  else { if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(POP_E_CUDA); g->own_stream = true; }
  GmGeo& G = g->G;
  auto logit = [](double x) { return std::log(x / (1 - x)); };
  G.res = c->resolution; G.res_inv = 1 / G.res;
  G.origin[0] = -c->map_size[0] / 2.0; G.origin[1] = -c->map_size[1] / 2.0; G.origin[2] = c->ground_height;
  long long N = 1;
  for (int k = 0; k < 3; k++) {
    G.n[k] = static_cast<int>(std::ceil(c->map_size[k] / G.res));
    G.bmin[k] = G.origin[k]; G.bmax[k] = G.origin[k] + c->map_size[k];
    G.lur[k] = c->local_update_range[k];
    N *= G.n[k];
  }
  if (N <= 0 || N >= (1ll << 31)) return fail(POP_E_INVALID);
  G.N = static_cast<int>(N);
  G.prob_hit = logit(c->p_hit); G.prob_miss = logit(c->p_miss); G.clamp_min = logit(c->p_min); G.clamp_max = logit(c->p_max);
  G.min_occ = logit(c->p_occ); G.unk = G.clamp_min - 0.01;
  G.max_ray = c->max_ray_length; G.ceil_h = c->virtual_ceil_height; G.ground = c->ground_height;
  if (G.ceil_h - G.ground > c->map_size[2]) G.ceil_h = G.ground + c->map_size[2];   // :65-67
  G.margin = c->local_map_margin; G.inf_step = static_cast<int>(std::ceil(c->obstacles_inflation / G.res));
  G.max_pts = c->max_points;
  G.R = static_cast<int>(std::ceil(c->max_ray_length / G.res)) + 2;
  G.B = 2 * G.R + 1;
  if (G.B > 1023 || 3 * G.R + 3 > kMaxPath) return fail(POP_E_UNSUPPORTED);
  const size_t nwords = (static_cast<size_t>(G.B) * G.B * G.B + 31) / 32;
  // rays traced per round: 128 when the bitmap leaves room for their paths, else 64
  g->rays_chunk = 2 * kChunk;
  g->rays_smem = sizeof(unsigned) * (nwords + g->rays_chunk * kMaxPath) + sizeof(int) * g->rays_chunk;
  if (g->rays_smem > 227 * 1024) {
    g->rays_chunk = kChunk;
    g->rays_smem = sizeof(unsigned) * (nwords + kChunk * kMaxPath) + sizeof(int) * kChunk;
  }
  if (g->rays_smem > 227 * 1024) return fail(POP_E_UNSUPPORTED);
  if (cudaFuncSetAttribute(k_gm_rays, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(g->rays_smem)) != cudaSuccess)
    return fail(POP_E_CUDA);
  auto alloc = [&](size_t bytes) -> void* {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    cudaMemsetAsync(p, 0, bytes, g->stream);
    g->allocs.push_back(p);
    return p;
  };
  g->h_maps.resize(n_maps);
  bool ok = true;
  for (int m = 0; m < n_maps && ok; m++) {
    GmMap& M = g->h_maps[m];
    M.occ = static_cast<double*>(alloc(sizeof(double) * N));
    M.inflate = static_cast<signed char*>(alloc(N));
    M.cnt_hm = static_cast<int*>(alloc(sizeof(int) * N));
    M.cnt_hit = static_cast<int*>(alloc(sizeof(int) * N));
    M.first_end = static_cast<int*>(alloc(sizeof(int) * N));
    M.touched = static_cast<int*>(alloc(sizeof(int) * N));
    M.ntouched = static_cast<int*>(alloc(sizeof(int)));
    M.pw = static_cast<double*>(alloc(sizeof(double) * 3 * c->max_points));
    M.endv = static_cast<int*>(alloc(sizeof(int) * c->max_points));
    M.bbox = static_cast<unsigned long long*>(alloc(sizeof(unsigned long long) * 6));
    M.lb = static_cast<int*>(alloc(sizeof(int) * 6));
    M.cast = static_cast<int*>(alloc(sizeof(int) * c->max_points));
    M.flag_rayend = static_cast<signed char*>(alloc(N));
    M.flag_traverse = static_cast<signed char*>(alloc(N));
    if (M.flag_rayend && M.flag_traverse) {
      cudaMemsetAsync(M.flag_rayend, 0xff, N, g->stream);
      cudaMemsetAsync(M.flag_traverse, 0xff, N, g->stream);
    }
    ok = M.occ && M.inflate && M.cnt_hm && M.cnt_hit && M.first_end && M.touched && M.ntouched && M.pw && M.endv && M.bbox && M.lb && M.cast && M.flag_rayend && M.flag_traverse;
    if (ok) k_gm_fill<<<static_cast<unsigned>((N + 255) / 256), 256, 0, g->stream>>>(M.occ, N, G.unk, M.first_end);
  }
  g->d_maps = static_cast<GmMap*>(alloc(sizeof(GmMap) * n_maps));
  g->d_pts = static_cast<double*>(alloc(sizeof(double) * 3 * static_cast<size_t>(c->max_points) * n_maps));
  g->d_cam = static_cast<double*>(alloc(sizeof(double) * 3 * n_maps));
  g->d_counts = static_cast<int*>(alloc(sizeof(int) * n_maps));
  g->d_err = static_cast<int*>(alloc(sizeof(int)));
  g->d_flags = static_cast<int*>(alloc(sizeof(int) * n_maps));
  g->frame.assign(n_maps, 0);
  if (!ok || !g->d_maps || !g->d_pts || !g->d_cam || !g->d_counts || !g->d_err || !g->d_flags) { cudaGetLastError(); return fail(POP_E_CUDA); }
  if (cudaMemcpyAsync(g->d_maps, g->h_maps.data(), sizeof(GmMap) * n_maps, cudaMemcpyHostToDevice, g->stream) != cudaSuccess ||
      cudaStreamSynchronize(g->stream) != cudaSuccess)
    return fail(POP_E_CUDA);
  *out = g;
  return POP_OK;
}

// raycastProcess + clearAndInflateLocalMap on the points resident in d_pts; counts = host copy of d_counts
static int gm_run(pop_gridmap* g, const int32_t* counts) {
  const GmGeo& G = g->G;
  const int A = g->n_maps;
  cudaStream_t st = g->stream;
  int maxn = 0;
  for (int m = 0; m < A; m++) maxn = counts[m] > maxn ? counts[m] : maxn;
  if (maxn == 0) { GM_CUDA(g, cudaStreamSynchronize(st)); return POP_OK; }
  std::vector<int> flags_work(A);
  for (int m = 0; m < A; m++) {
    if (counts[m] > 0) g->frame[m] += 1;           // md_.raycast_num_
    flags_work[m] = g->frame[m];                   // the kernels take the char it wraps to
  }
  GM_CUDA(g, cudaMemcpyAsync(g->d_flags, flags_work.data(), sizeof(int) * A, cudaMemcpyHostToDevice, st));
  GM_CUDA(g, cudaStreamSynchronize(st));           // flags_work is stack-owned
  k_gm_frame_reset<<<(A + 63) / 64, 64, 0, st>>>(g->d_maps, A);
  k_gm_points<<<dim3((maxn + 255) / 256, A), 256, 0, st>>>(G, g->d_maps, g->d_pts, g->d_counts, g->d_cam, g->d_flags, g->d_err);
  k_gm_rays<<<A, 256, g->rays_smem, st>>>(G, g->d_maps, g->d_counts, g->d_cam, g->d_flags, g->d_err, g->rays_chunk);
  k_gm_update<<<dim3(64, A), 256, 0, st>>>(G, g->d_maps, g->d_counts, g->d_cam);
  k_gm_clear<<<dim3(128, A), 256, 0, st>>>(G, g->d_maps, g->d_counts);
  k_gm_inflate<<<dim3(128, A), 256, 0, st>>>(G, g->d_maps, g->d_counts);
  GM_CUDA(g, cudaGetLastError());
  int e = 0;
  GM_CUDA(g, cudaMemcpyAsync(&e, g->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  GM_CUDA(g, cudaStreamSynchronize(st));
  if (e) {
    cudaMemsetAsync(g->d_err, 0, sizeof(int), st);
    if (e & 4) return gm_fail(g, POP_E_INVALID, "non-finite projected points");
    if (e & 2) return gm_fail(g, POP_E_CAPACITY, "the depth image projects to more than max_points=%d points", G.max_pts);
    return gm_fail(g, POP_E_CAPACITY, "a ray left the %d^3 voxel box around the camera or exceeded %d voxels", G.B, kMaxPath);
  }
  return POP_OK;
}

static int gm_check_cameras(pop_gridmap* g, const double* cam_pos) {
  const GmGeo& G = g->G;
  for (int m = 0; m < g->n_maps; m++) {
    const double* c = cam_pos + 3 * m;
    for (int k = 0; k < 3; k++)
      if (!(c[k] >= G.bmin[k] + 1e-4 && c[k] <= G.bmax[k] - 1e-4))
        return gm_fail(g, POP_E_INVALID, "map %d: the camera must be inside the map (the reference indexes out of bounds otherwise)", m);
  }
  return POP_OK;
}

int pop_gridmap_update(pop_gridmap* g, const double* pts, const int32_t* counts, const double* cam_pos) {
  if (!g || !pts || !counts || !cam_pos) return POP_E_INVALID;
  GmDeviceGuard guard(g);
  const GmGeo& G = g->G;
  const int A = g->n_maps;
  for (int m = 0; m < A; m++)
    if (counts[m] < 0 || counts[m] > G.max_pts) return gm_fail(g, POP_E_CAPACITY, "map %d: %d points exceed max_points=%d", m, counts[m], G.max_pts);
  int rc = gm_check_cameras(g, cam_pos);
  if (rc) return rc;
  cudaStream_t st = g->stream;
  for (int m = 0; m < A; m++)
    if (counts[m] > 0)
      GM_CUDA(g, cudaMemcpyAsync(g->d_pts + static_cast<size_t>(m) * G.max_pts * 3, pts + static_cast<size_t>(m) * G.max_pts * 3,
                                 sizeof(double) * 3 * counts[m], cudaMemcpyHostToDevice, st));
  GM_CUDA(g, cudaMemcpyAsync(g->d_counts, counts, sizeof(int) * A, cudaMemcpyHostToDevice, st));
  GM_CUDA(g, cudaMemcpyAsync(g->d_cam, cam_pos, sizeof(double) * 3 * A, cudaMemcpyHostToDevice, st));
  return gm_run(g, counts);
}

int pop_gridmap_set_depth_params(pop_gridmap* g, const pop_depth_params* p) {
  if (!g || !p) return POP_E_INVALID;
  GmDeviceGuard guard(g);
  if (p->rows < 1 || p->cols < 1 || p->skip_pixel < 1 || p->depth_filter_margin < 0 || !(p->fx > 0) || !(p->fy > 0) ||
      !(p->k_depth_scaling_factor > 0))
    return gm_fail(g, POP_E_INVALID, "bad depth parameters");
  const size_t npix = static_cast<size_t>(p->rows) * p->cols;
  if (g->d_img) { cudaFree(g->d_img); g->d_img = nullptr; }
  if (g->d_R) { cudaFree(g->d_R); g->d_R = nullptr; }
  GM_CUDA(g, cudaMalloc(&g->d_img, sizeof(unsigned short) * npix * g->n_maps));
  GM_CUDA(g, cudaMalloc(&g->d_R, sizeof(double) * 9 * g->n_maps));
  g->depth = *p;
  g->have_depth = true;
  g->has_first.assign(g->n_maps, 0);
  return POP_OK;
}

int pop_gridmap_update_depth(pop_gridmap* g, const uint16_t* depth, const double* camera_r, const double* cam_pos) {
  if (!g || !depth || !camera_r || !cam_pos) return POP_E_INVALID;
  GmDeviceGuard guard(g);
  if (!g->have_depth) return gm_fail(g, POP_E_STATE, "no depth parameters (pop_gridmap_set_depth_params)");
  int rc = gm_check_cameras(g, cam_pos);
  if (rc) return rc;
  const int A = g->n_maps;
  cudaStream_t st = g->stream;
  const size_t npix = static_cast<size_t>(g->depth.rows) * g->depth.cols;
  // with the depth filter the first frame of a map only arms has_first_depth_ (:244-246)
  std::vector<int> active(A, 1);
  if (g->depth.use_depth_filter)
    for (int m = 0; m < A; m++) { active[m] = g->has_first[m]; g->has_first[m] = 1; }
  GM_CUDA(g, cudaMemcpyAsync(g->d_img, depth, sizeof(unsigned short) * npix * A, cudaMemcpyHostToDevice, st));
  GM_CUDA(g, cudaMemcpyAsync(g->d_R, camera_r, sizeof(double) * 9 * A, cudaMemcpyHostToDevice, st));
  GM_CUDA(g, cudaMemcpyAsync(g->d_cam, cam_pos, sizeof(double) * 3 * A, cudaMemcpyHostToDevice, st));
  GM_CUDA(g, cudaMemcpyAsync(g->d_flags, active.data(), sizeof(int) * A, cudaMemcpyHostToDevice, st));
  k_gm_project<<<A, 1024, 0, st>>>(g->depth, g->G.max_ray, g->d_img, g->d_R, g->d_cam, g->d_flags, g->d_pts, g->G.max_pts,
                                   g->d_counts, g->d_err);
  GM_CUDA(g, cudaGetLastError());
  std::vector<int32_t> counts(A);
  GM_CUDA(g, cudaMemcpyAsync(counts.data(), g->d_counts, sizeof(int) * A, cudaMemcpyDeviceToHost, st));
  GM_CUDA(g, cudaStreamSynchronize(st));
  g->last_counts = counts;
  return gm_run(g, counts.data());
}

int pop_gridmap_projected_points(pop_gridmap* g, int map, double* out, int cap, int32_t* n_out) {
  if (!g || map < 0 || map >= g->n_maps || !out || !n_out || cap < 0) return POP_E_INVALID;
  GmDeviceGuard guard(g);
  if (g->last_counts.empty()) return gm_fail(g, POP_E_STATE, "no depth frame has been projected yet");
  const int n = g->last_counts[map];
  *n_out = n;
  const int m = n < cap ? n : cap;
  if (m > 0)
    GM_CUDA(g, cudaMemcpy(out, g->d_pts + static_cast<size_t>(map) * g->G.max_pts * 3, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost));
  return n > cap ? POP_E_CAPACITY : POP_OK;
}

int pop_gridmap_query(pop_gridmap* g, int map, const double* pos, int n, int8_t* out_occ, int8_t* out_inflate) {
  if (!g || map < 0 || map >= g->n_maps || !pos || n < 0) return POP_E_INVALID;
  GmDeviceGuard guard(g);
  if (n == 0) return POP_OK;
  const size_t bytes = sizeof(double) * 3 * n + 2 * static_cast<size_t>(n) + 64;
  if (bytes > g->d_q_bytes) {
    if (g->d_q) cudaFree(g->d_q);
    g->d_q = nullptr;
    GM_CUDA(g, cudaMalloc(&g->d_q, bytes * 2));
    g->d_q_bytes = bytes * 2;
  }
  double* d_pos = static_cast<double*>(g->d_q);
  signed char* d_o = reinterpret_cast<signed char*>(d_pos + 3 * static_cast<size_t>(n));
  signed char* d_i = d_o + n;
  GM_CUDA(g, cudaMemcpyAsync(d_pos, pos, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, g->stream));
  k_gm_query<<<(n + 127) / 128, 128, 0, g->stream>>>(g->G, g->h_maps[map], d_pos, n, d_o, d_i);
  GM_CUDA(g, cudaGetLastError());
  if (out_occ) GM_CUDA(g, cudaMemcpyAsync(out_occ, d_o, n, cudaMemcpyDeviceToHost, g->stream));
  if (out_inflate) GM_CUDA(g, cudaMemcpyAsync(out_inflate, d_i, n, cudaMemcpyDeviceToHost, g->stream));
  GM_CUDA(g, cudaStreamSynchronize(g->stream));
  return POP_OK;
}

int pop_gridmap_download(pop_gridmap* g, int map, double* occ, int8_t* inflate, int32_t* bounds) {
  if (!g || map < 0 || map >= g->n_maps) return POP_E_INVALID;
  GmDeviceGuard guard(g);
  const GmMap& M = g->h_maps[map];
  if (occ) GM_CUDA(g, cudaMemcpyAsync(occ, M.occ, sizeof(double) * g->G.N, cudaMemcpyDeviceToHost, g->stream));
  if (inflate) GM_CUDA(g, cudaMemcpyAsync(inflate, M.inflate, g->G.N, cudaMemcpyDeviceToHost, g->stream));
  if (bounds) GM_CUDA(g, cudaMemcpyAsync(bounds, M.lb, sizeof(int) * 6, cudaMemcpyDeviceToHost, g->stream));
  GM_CUDA(g, cudaStreamSynchronize(g->stream));
  return POP_OK;
}

}  // extern "C"
