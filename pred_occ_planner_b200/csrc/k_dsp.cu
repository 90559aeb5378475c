// k_dsp.cu — K1d: the DSP particle map (SMC-PHD filter over particles binned in voxels and in
// field-of-view pyramids) that produces the SOGM on the real-perception path.
//
// Reference: plan_env/include/plan_env/dsp_dynamic.h (DSPMap::update :165-364 and the steps it calls),
// plan_env/src/risk_mapping_node.cpp:74-188 (filterPointCloud, cloudCallback, publishMap),
// plan_env/src/risk_base.cpp:60-80 (futureRiskCallback).
//
// The reference is one sequential sweep per step; its results depend on that order (which free slot a
// moved particle takes, which particles a full voxel or a full pyramid drops, the order fp32 sums are
// taken in).  Every kernel here reproduces the sequential result exactly:
//   * a particle is identified by its sweep time (source voxel * S + slot); arrivals into a voxel and
//     candidates of a pyramid are processed in ascending sweep time by the ONE thread/warp that owns the
//     destination, so slot assignment and overflow drops equal the sequential ones;
//   * voxel placement and pyramid admission depend on each other (a particle dropped by a full pyramid
//     frees its voxel slot; a particle dropped by a full voxel never reaches its pyramid).  Both are
//     functions of EARLIER particles only, so iterating placement <-> admission from "all admitted"
//     reaches the unique causal fixed point — the sequential result — after 2 rounds in practice; the
//     rounds stop when admission no longer changes and the update fails loudly if kDspMaxIter rounds
//     were not enough;
//   * per-particle and per-observation sums run in one thread in the reference's loop order;
//   * the noise tables are consumed through prefix sums of the per-point draw counts, which gives every
//     draw the sequence index the sequential loop would have used.
// The only order-dependent quantity left is the fp32 accumulation of future-occupancy deposits into the
// SOGM (atomicAdd): equal to the sequential sum up to fp32 rounding (tests: 1e-5 relative).
//
// No FMA contraction (-fmad=false), IEEE division and square root: fp32 expressions round as on the CPU.
#include <cfloat>
#include <math_constants.h>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "k_dsp.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// helpers
__device__ __forceinline__ bool dsp_vidx(const DspGeo& G, float px, float py, float pz, int* idx) {
  // getParticleVoxelsIndex :1148-1163
  if (px >= G.hx || px <= -G.hx || py >= G.hy || py <= -G.hy || pz >= G.hz || pz <= -G.hz) return false;
  const int x = static_cast<int>((px + G.hx) / G.res);
  const int y = static_cast<int>((py + G.hy) / G.res);
  const int z = static_cast<int>((pz + G.hz) / G.res);
  const int i = z * G.ny * G.nx + y * G.nx + x;
  if (i < 0 || i >= G.V) return false;
  *idx = i;
  return true;
}

// rotateVectorByQuaternion :1379-1401 (att * (0, v) * att.inverse())
__device__ __forceinline__ void dsp_rotate(const float* v, const float* q, float* out) {
  // Eigen 3.3 on x86-64 compiles Quaternionf * Quaternionf to its SSE kernel (Geometry/arch/Geometry_SSE.h,
  // quat_product<Architecture::SSE, ., ., float>), whose lanes pair the eight products as written here; the squared norm of
  // inverse() is a packet reduction, (x^2 + z^2) + (y^2 + w^2) (predux<Packet4f> without SSE3), and conjugate / n2 divides
  // every coefficient.  This is the arithmetic the reference's binary runs on the hosts this repo is measured on.
  const float aw = q[0], ax = q[1], ay = q[2], az = q[3];
  const float bw = 0.f, bx = v[0], by = v[1], bz = v[2];
  const float tx = (ax * bw - az * by) + (ay * bz + aw * bx);
  const float ty = (ay * bw - ax * bz) + (az * bx + aw * by);
  const float tz = (az * bw - ay * bx) + (ax * by + aw * bz);
  const float tw = (aw * bw - ax * bx) - (az * bz + ay * by);
  const float n2 = (ax * ax + az * az) + (ay * ay + aw * aw);
  float iw = 0.f, ix = 0.f, iy = 0.f, iz = 0.f;
  if (n2 > 0.f) { iw = aw / n2; ix = -ax / n2; iy = -ay / n2; iz = -az / n2; }
  out[0] = (tx * iw - tz * iy) + (ty * iz + tw * ix);
  out[1] = (ty * iw - tx * iz) + (tz * ix + tw * iy);
  out[2] = (tz * iw - ty * ix) + (tx * iy + tw * iz);
}

__device__ __forceinline__ float dot3(float x1, float y1, float z1, const float* b) {
  return x1 * b[0] + y1 * b[1] + z1 * b[2];
}
// pyramid index of a point from the rotated plane normals (ifInPyramidsArea :1411-1427,
// findPointPyramidHorizontalIndex / VerticalIndex :1429-1471); -1 outside the field of view
__device__ __forceinline__ int dsp_pyramid(const DspGeo& G, const float* __restrict__ bph,
                                           const float* __restrict__ bpv, float x, float y, float z) {
  if (!(dot3(x, y, z, bph) >= 0.f && dot3(x, y, z, bph + 3 * G.nph) <= 0.f && dot3(x, y, z, bpv) <= 0.f &&
        dot3(x, y, z, bpv + 3 * G.npv) >= 0.f))
    return -1;
  int ih = -1, iv = -1;
  float last = 1.f;
  for (int i = 0; i < G.nph; i++) {
    const float d = dot3(x, y, z, bph + 3 * (i + 1));
    if (last * d <= 0.f) { ih = i; break; }
    last = d;
  }
  last = -1.f;
  for (int j = 0; j < G.npv; j++) {
    const float d = dot3(x, y, z, bpv + 3 * (j + 1));
    if (last * d <= 0.f) { iv = j; break; }
    last = d;
  }
  // the reference returns -1 from the index search "which should not happen" and then indexes with it
  if (ih < 0 || iv < 0) return -1;
  return ih * G.npv + iv;
}

// queryNormalPDF :1369-1377
__device__ __forceinline__ float dsp_pdf(const float* __restrict__ pdf, float x, float mu, float sigma) {
  float cx = (x - mu) / sigma;
  if (cx > 9.9f) cx = 9.9f; else if (cx < -9.9f) cx = -9.9f;
  return __ldg(pdf + static_cast<int>(cx * 1000 + 10000));
}

// generateRandomFloat :1687-1689 on the supplied rand() stream
__device__ __forceinline__ float dsp_rand_float(int r, float mn, float mx) {
  return mn + static_cast<float>(r) / (static_cast<float>(2147483647 / (mx - mn)));
}

__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// exclusive scan over the block (blockDim.x a multiple of 32, <= 1024); s_w: 33 ints
__device__ __forceinline__ int block_scan_excl(int v, int* s_w, int* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protects s_w from the previous use
  if (lane == 31) s_w[w] = inc;
  __syncthreads();
  if (w == 0) {
    int x = lane < nw ? s_w[lane] : 0;
    int xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) xi += t;
    }
    if (lane < nw) s_w[lane] = xi - x;
    if (lane == 31) s_w[32] = xi;
  }
  __syncthreads();
  *total = s_w[32];
  return inc - v + s_w[w];
}

__device__ __forceinline__ unsigned flag_byte(const uint4& f, int p) {
  const unsigned w = p < 4 ? f.x : p < 8 ? f.y : p < 12 ? f.z : f.w;
  return (w >> (8 * (p & 3))) & 0xffu;
}
__device__ __forceinline__ void set_flag_byte(uint4& f, int p, unsigned v) {
  unsigned& w = p < 4 ? f.x : p < 8 ? f.y : p < 12 ? f.z : f.w;
  const int sh = 8 * (p & 3);
  w = (w & ~(0xffu << sh)) | (v << sh);
}

constexpr int kFilterSmemKeys = 16384;

// ---------------------------------------------------------------------------------------------
// A2  filterPointCloud (risk_mapping_node.cpp:74-99): pcl::VoxelGrid centroid per leaf in ascending leaf
// index, camera -> body axes, range test in double, first max_points kept.  One CTA per agent.
__global__ void __launch_bounds__(1024) k_dsp_filter(DspGeo G, const DspFrame* __restrict__ fr,
                                                     const float* __restrict__ cam_all,
                                                     unsigned long long* __restrict__ keys_all,
                                                     float* __restrict__ pts_all, int* __restrict__ npts,
                                                     int* __restrict__ err, const int* __restrict__ cam_counts) {
  extern __shared__ unsigned long long s_keys[];   // kFilterSmemKeys sort keys
  const int a = blockIdx.x, tid = threadIdx.x;
  __shared__ unsigned s_mn[3], s_mx[3];
  __shared__ int s_w[33];
  const int n = cam_counts ? cam_counts[a] : fr[a].n_in;   // rendered clouds carry their size on the device
  const float* cam = cam_all + static_cast<size_t>(a) * G.MAXIN * 3;
  float* pts = pts_all + static_cast<size_t>(a) * G.MAXPTS * 3;
  if (tid < 3) { s_mn[tid] = 0xffffffffu; s_mx[tid] = 0u; }
  __syncthreads();
  const float inv = 1.0f / G.filt;
  for (int i = tid; i < n; i += blockDim.x) {
    const float x = cam[3 * i], y = cam[3 * i + 1], z = cam[3 * i + 2];
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
      atomicMin(&s_mn[0], f2ord(x)); atomicMax(&s_mx[0], f2ord(x));
      atomicMin(&s_mn[1], f2ord(y)); atomicMax(&s_mx[1], f2ord(y));
      atomicMin(&s_mn[2], f2ord(z)); atomicMax(&s_mx[2], f2ord(z));
    }
  }
  __syncthreads();
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  // clouds of up to kFilterSmemKeys points are sorted in shared memory, larger ones in their (L2-resident) global segment
  unsigned long long* keys = np2 <= kFilterSmemKeys ? s_keys : keys_all + static_cast<size_t>(a) * G.NPOW2;
  if (s_mx[0] == 0u || n == 0) {  // no finite point
    if (tid == 0) npts[a] = 0;
    return;
  }
  int minb[3];
  long long div[3];
  for (int k = 0; k < 3; k++) {
    minb[k] = static_cast<int>(floorf(ord2f(s_mn[k]) * inv));
    const int maxb = static_cast<int>(floorf(ord2f(s_mx[k]) * inv));
    div[k] = static_cast<long long>(maxb) - minb[k] + 1;
  }
  if (div[0] * div[1] * div[2] >= (1ll << 43)) {
    if (tid == 0) { atomicOr(err, 1); npts[a] = 0; }
    return;
  }
  for (int i = tid; i < np2; i += blockDim.x) {
    unsigned long long key = ~0ull;
    if (i < n) {
      const float x = cam[3 * i], y = cam[3 * i + 1], z = cam[3 * i + 2];
      if (isfinite(x) && isfinite(y) && isfinite(z)) {
        const long long i0 = static_cast<int>(floorf(x * inv)) - minb[0];
        const long long i1 = static_cast<int>(floorf(y * inv)) - minb[1];
        const long long i2 = static_cast<int>(floorf(z * inv)) - minb[2];
        key = (static_cast<unsigned long long>(i0 + i1 * div[0] + i2 * div[0] * div[1]) << 20) |
              static_cast<unsigned long long>(i);
      }
    }
    keys[i] = key;
  }
  __syncthreads();
  // bitonic sort of (leaf, input order) keys; the segment stays L2-resident
  for (int k = 2; k <= np2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < np2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long x = keys[i], y = keys[l];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { keys[i] = y; keys[l] = x; }
        }
      }
      __syncthreads();
    }
  // leaf heads in ascending leaf order -> centroid -> body frame -> range -> ordered compaction
  int carry = 0;
  for (int base = 0; base < np2 && carry < G.MAXPTS; base += blockDim.x) {
    const int i = base + tid;
    bool keep = false;
    float bx = 0.f, by = 0.f, bz = 0.f;
    if (i < np2) {
      const unsigned long long key = keys[i];
      if (key != ~0ull && (i == 0 || (keys[i - 1] >> 20) != (key >> 20))) {
        float sx = 0.f, sy = 0.f, sz = 0.f;
        int cnt = 0;
        for (int r = i; r < np2; r++) {
          const unsigned long long kr = keys[r];
          if (kr == ~0ull || (kr >> 20) != (key >> 20)) break;
          const int src = static_cast<int>(kr & 0xfffffull);
          sx += cam[3 * src]; sy += cam[3 * src + 1]; sz += cam[3 * src + 2];
          cnt++;
        }
        const float c = static_cast<float>(cnt);
        const float cx = sx / c, cy = sy / c, cz = sz / c;
        bx = cz; by = -cx; bz = -cy;
        keep = static_cast<double>(bx) > -G.RX && static_cast<double>(bx) < G.RX && static_cast<double>(by) > -G.RY &&
               static_cast<double>(by) < G.RY && static_cast<double>(bz) > -G.RZ && static_cast<double>(bz) < G.RZ;
      }
    }
    int total;
    const int pos = carry + block_scan_excl(keep ? 1 : 0, s_w, &total);
    if (keep && pos < G.MAXPTS) { pts[3 * pos] = bx; pts[3 * pos + 1] = by; pts[3 * pos + 2] = bz; }
    carry += total;
  }
  if (tid == 0) npts[a] = carry < G.MAXPTS ? carry : G.MAXPTS;
}

// ---------------------------------------------------------------------------------------------
// A3  observation binning (DSPMap::update :216-300).  One CTA per agent: rotate the plane normals and
// every point, pyramid index, slot = number of earlier points of the same pyramid (the reference's
// saturating counter keeps OCAP-1 of them).
__global__ void __launch_bounds__(1024) k_dsp_bin(DspGeo G, const DspFrame* __restrict__ fr,
                                                  const float* __restrict__ bp_ori, float* __restrict__ bp_all,
                                                  const float* __restrict__ pts_all, const int* __restrict__ npts,
                                                  float* __restrict__ rot_all, int* __restrict__ pobs_all,
                                                  float4* __restrict__ obs_all, int* __restrict__ obs_num_all,
                                                  float* __restrict__ maxlen_all, int* __restrict__ valid) {
  extern __shared__ unsigned char smem_raw[];
  const int a = blockIdx.x, tid = threadIdx.x;
  if (!fr[a].ok) return;
  const int nplanes = G.nph + 1 + G.npv + 1;
  float* s_bp = reinterpret_cast<float*>(smem_raw);                // nplanes*3
  int* s_cnt = reinterpret_cast<int*>(s_bp + ((nplanes * 3 + 3) & ~3));  // NP
  unsigned* s_ml = reinterpret_cast<unsigned*>(s_cnt + G.NP);      // NP (ordered float bits)
  int* s_p = reinterpret_cast<int*>(s_ml + G.NP);                  // blockDim
  __shared__ int s_valid;
  float q[4] = {fr[a].q[0], fr[a].q[1], fr[a].q[2], fr[a].q[3]};
  float* bp = bp_all + static_cast<size_t>(a) * nplanes * 3;
  for (int i = tid; i < nplanes; i += blockDim.x) {
    float o[3];
    dsp_rotate(bp_ori + 3 * i, q, o);
    s_bp[3 * i] = o[0]; s_bp[3 * i + 1] = o[1]; s_bp[3 * i + 2] = o[2];
    bp[3 * i] = o[0]; bp[3 * i + 1] = o[1]; bp[3 * i + 2] = o[2];
  }
  for (int i = tid; i < G.NP; i += blockDim.x) { s_cnt[i] = 0; s_ml[i] = f2ord(-1.f); }
  if (tid == 0) s_valid = 0;
  __syncthreads();
  const float* bph = s_bp;
  const float* bpv = s_bp + 3 * (G.nph + 1);
  const int n = npts[a];
  const float* pts = pts_all + static_cast<size_t>(a) * G.MAXPTS * 3;
  float* rot = rot_all + static_cast<size_t>(a) * G.MAXPTS * 3;
  int* pobs = pobs_all + static_cast<size_t>(a) * G.MAXPTS;
  float4* obs = obs_all + static_cast<size_t>(a) * G.NP * G.OCAP;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + tid;
    int p = -1;
    float r[3] = {0.f, 0.f, 0.f};
    if (i < n) {
      dsp_rotate(pts + 3 * i, q, r);
      rot[3 * i] = r[0]; rot[3 * i + 1] = r[1]; rot[3 * i + 2] = r[2];
      p = dsp_pyramid(G, bph, bpv, r[0], r[1], r[2]);
    }
    s_p[tid] = p;
    __syncthreads();
    if (p >= 0) {
      int rank = 0;
      for (int j = 0; j < tid; j++) rank += (s_p[j] == p);
      const int slot = s_cnt[p] + rank;
      const float len = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
      atomicMax(&s_ml[p], f2ord(len));
      if (slot < G.OCAP - 1) {
        obs[static_cast<size_t>(p) * G.OCAP + slot] = make_float4(r[0], r[1], r[2], 0.f);
        pobs[i] = p * G.OCAP + slot;
      } else {
        pobs[i] = -1;
      }
      atomicAdd(&s_valid, 1);
    } else if (i < n) {
      pobs[i] = -1;
    }
    __syncthreads();
    if (p >= 0) atomicAdd(&s_cnt[p], 1);
    __syncthreads();
  }
  int* obs_num = obs_num_all + static_cast<size_t>(a) * G.NP;
  float* maxlen = maxlen_all + static_cast<size_t>(a) * G.NP;
  for (int i = tid; i < G.NP; i += blockDim.x) {
    obs_num[i] = s_cnt[i] < G.OCAP - 1 ? s_cnt[i] : G.OCAP - 1;
    maxlen[i] = ord2f(s_ml[i]);
  }
  if (tid == 0) valid[a] = s_valid;
}

// ---------------------------------------------------------------------------------------------
// A6  velocityEstimationThread (:1487-1683).  One CTA per agent: ground split, Euclidean clustering
// (connected components of the 2*filter_res radius graph by lock-free union-find, root = smallest member),
// clusters ordered by size, centroid sums in member order, minimum-cost matching with the previous
// frame's clusters (shortest augmenting paths, warp-parallel column scans), per-point velocities written
// in the reference's output order.
struct VelSmem {
  float* gx; float* gy; float* gz;  // non-ground points, world frame, cloud order
  int* parent;                      // union-find / root
  int* aux;                         // component size -> cluster rank of a root (-1: dropped)
  int* mrank;                       // member rank of a point inside its cluster
  int* croot; int* csize;           // clusters in ascending root order, then indexed by rank
  float* ccx; float* ccy; float* ccz; float* cint; float* cvx; float* cvy; float* cvz;
  int* cdyn; int* coff;
};

__device__ __forceinline__ int uf_find(int* parent_, int x) {
  volatile int* parent = parent_;
  int p = parent[x];
  while (p != x) { x = p; p = parent[x]; }
  return x;
}

__global__ void __launch_bounds__(1024) k_dsp_velocity(DspGeo G, const DspFrame* __restrict__ fr,
                                                       const float* __restrict__ rot_all,
                                                       const int* __restrict__ npts, const int* __restrict__ ur,
                                                       int* __restrict__ seq_all, float* __restrict__ cv_all,
                                                       int* __restrict__ ncv, float* __restrict__ cl_last_all,
                                                       int* __restrict__ ncl_last, double* __restrict__ cost_all,
                                                       int* __restrict__ err) {
  extern __shared__ unsigned char smem_raw[];
  const int a = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  if (!fr[a].ok) return;
  const int n = npts[a];
  if (n == 0) return;  // :1488: returns before clearing input_cloud_with_velocity (stale points stay)
  __shared__ int s_w[33];
  __shared__ int s_nng, s_ngr, s_nc, s_ndyn, s_nstat_c;
  VelSmem S;
  {
    float* f = reinterpret_cast<float*>(smem_raw);
    S.gx = f; S.gy = f + G.MAXPTS; S.gz = f + 2 * G.MAXPTS;
    int* ip = reinterpret_cast<int*>(f + 3 * G.MAXPTS);
    S.parent = ip; S.aux = ip + G.MAXPTS; S.mrank = ip + 2 * G.MAXPTS;
    int* c = ip + 3 * G.MAXPTS;
    S.croot = c; S.csize = c + kDspMaxClusters; S.cdyn = c + 2 * kDspMaxClusters; S.coff = c + 3 * kDspMaxClusters;
    float* cf = reinterpret_cast<float*>(c + 4 * kDspMaxClusters);
    S.ccx = cf; S.ccy = cf + kDspMaxClusters; S.ccz = cf + 2 * kDspMaxClusters; S.cint = cf + 3 * kDspMaxClusters;
    S.cvx = cf + 4 * kDspMaxClusters; S.cvy = cf + 5 * kDspMaxClusters; S.cvz = cf + 6 * kDspMaxClusters;
  }
  const float* rot = rot_all + static_cast<size_t>(a) * G.MAXPTS * 3;
  float* cv = cv_all + static_cast<size_t>(a) * G.MAXPTS * 7;
  const float px0 = fr[a].pos[0], py0 = fr[a].pos[1], pz0 = fr[a].pos[2];
  int* seq = seq_all + 4 * a;
  const int useq0 = seq[2];

  // 1. ground split with ordered compaction (non-ground -> smem, ground -> tail of the output later)
  //    ground rank of point i is kept in mrank[] of a scratch region: reuse aux[] indexed by cloud index
  int carry_ng = 0, carry_gr = 0;
  for (int base = 0; base < n; base += nth) {
    const int i = base + tid;
    float x = 0.f, y = 0.f, z = 0.f;
    bool ng = false, gr = false;
    if (i < n) {
      x = rot[3 * i] + px0; y = rot[3 * i + 1] + py0; z = rot[3 * i + 2] + pz0;
      ng = z > G.filt;
      gr = !ng;
    }
    int tot_ng, tot_gr;
    const int r_ng = carry_ng + block_scan_excl(ng ? 1 : 0, s_w, &tot_ng);
    const int r_gr = carry_gr + block_scan_excl(gr ? 1 : 0, s_w, &tot_gr);
    if (ng) { S.gx[r_ng] = x; S.gy[r_ng] = y; S.gz[r_ng] = z; }
    (void)r_gr;  // ground points are written in step 7 (their offset needs the dynamic point count)
    carry_ng += tot_ng; carry_gr += tot_gr;
  }
  if (tid == 0) { s_nng = carry_ng; s_ngr = carry_gr; s_nc = 0; s_ndyn = 0; s_nstat_c = 0; }
  __syncthreads();
  const int m = s_nng;

  // 2. union-find over pairs within tolerance (hook the larger root under the smaller one)
  for (int i = tid; i < m; i += nth) { S.parent[i] = i; S.aux[i] = 0; }
  __syncthreads();
  const float tol = 2 * G.filt;
  const float tol2 = tol * tol;
  for (int i = tid; i < m; i += nth) {
    const float xi = S.gx[i], yi = S.gy[i], zi = S.gz[i];
    for (int j = i + 1; j < m; j++) {
      const float dx = xi - S.gx[j], dy = yi - S.gy[j], dz = zi - S.gz[j];
      if (dx * dx + dy * dy + dz * dz <= tol2) {
        int ra = uf_find(S.parent, i), rb = uf_find(S.parent, j);
        while (ra != rb) {
          const int hi = ra > rb ? ra : rb, lo = ra > rb ? rb : ra;
          const int old = atomicCAS(&S.parent[hi], hi, lo);
          if (old == hi) break;
          ra = uf_find(S.parent, hi); rb = uf_find(S.parent, lo);
        }
      }
    }
  }
  __syncthreads();
  // 3. roots and component sizes
  int myroot[8];
  {
    int k = 0;
    for (int i = tid; i < m; i += nth, k++) myroot[k] = uf_find(S.parent, i);
    __syncthreads();
    k = 0;
    for (int i = tid; i < m; i += nth, k++) { S.parent[i] = myroot[k]; atomicAdd(&S.aux[myroot[k]], 1); }
  }
  __syncthreads();
  // 4. clusters = roots with 5 <= size <= 10000, compacted in ascending root order
  int carry_c = 0;
  for (int base = 0; base < m; base += nth) {
    const int i = base + tid;
    const bool isc = i < m && S.parent[i] == i && S.aux[i] >= 5 && S.aux[i] <= 10000;
    int tot;
    const int r = carry_c + block_scan_excl(isc ? 1 : 0, s_w, &tot);
    if (isc) {
      if (r < kDspMaxClusters) { S.croot[r] = i; S.csize[r] = S.aux[i]; } else atomicOr(err, 2);
    }
    carry_c += tot;
  }
  __syncthreads();
  const int nc = carry_c < kDspMaxClusters ? carry_c : kDspMaxClusters;
  //    order: size descending, ties by smaller root; rank -> permuted arrays
  int my_c_root = -1, my_c_size = 0, my_c_rank = -1;
  if (tid < nc) {
    my_c_root = S.croot[tid]; my_c_size = S.csize[tid];
    int rank = 0;
    for (int j = 0; j < nc; j++) {
      const int sj = S.csize[j];
      rank += (sj > my_c_size) || (sj == my_c_size && j < tid);
    }
    my_c_rank = rank;
  }
  __syncthreads();
  for (int i = tid; i < m; i += nth) S.aux[i] = -1;   // aux: root -> cluster rank
  __syncthreads();
  if (tid < nc) { S.croot[my_c_rank] = my_c_root; S.csize[my_c_rank] = my_c_size; S.aux[my_c_root] = my_c_rank; }
  __syncthreads();
  // 5. per cluster (thread = rank): members in ascending index -> centroid, classification, intensity
  if (tid < nc) {
    const int root = S.croot[tid];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    int cnt = 0;
    for (int i = root; i < m; i++)
      if (S.parent[i] == root) { sx += S.gx[i]; sy += S.gy[i]; sz += S.gz[i]; S.mrank[i] = cnt; cnt++; }
    const float c = static_cast<float>(cnt);
    const float cx = sx / c, cy = sy / c, cz = sz / c;
    S.ccx[tid] = cx; S.ccy[tid] = cy; S.ccz[tid] = cz;
    S.cint[tid] = dsp_rand_float(ur[(useq0 + tid) % G.nrand], 0.1f, 1.f);
    S.cvx[tid] = -10000.f; S.cvy[tid] = -10000.f; S.cvz[tid] = -10000.f;
    S.cdyn[tid] = !(cnt > G.cl_maxn || cz > G.cl_maxh);
  }
  __syncthreads();
  // dynamic index / offsets in rank order (sequential, <= 1024 clusters)
  if (tid == 0) {
    int nd = 0, off_d = 0, off_s = 0;
    for (int c = 0; c < nc; c++) {
      if (S.cdyn[c]) { S.coff[c] = off_d; off_d += S.csize[c]; S.cdyn[c] = 1 + nd; nd++; }
      else { S.coff[c] = off_s; off_s += S.csize[c]; }
    }
    s_ndyn = nd;
    s_nc = off_d;       // points in dynamic clusters
    s_nstat_c = off_s;  // points in static clusters
    seq[2] = (useq0 + nc) % G.nrand;
  }
  __syncthreads();
  const int ndyn = s_ndyn;
  if (ndyn > kDspMaxDyn) { if (tid == 0) atomicOr(err, 4); }
  // 6. matching with the previous frame (:1566-1633)
  const int nlast = ncl_last[a];
  float* cl_last = cl_last_all + static_cast<size_t>(a) * kDspMaxDyn * 5;
  const float dtl = fr[a].dt;
  if (nlast > 0 && ndyn > 0 && ndyn <= kDspMaxDyn && dtl > 0.00001 && dtl < 10.0) {
    double* cost = cost_all + static_cast<size_t>(a) * kDspMaxDyn * kDspMaxDyn;  // [ndyn][nlast]
    __shared__ int s_dyn2c[kDspMaxDyn];
    __shared__ int s_asg[kDspMaxDyn];
    for (int c = tid; c < nc; c += nth) if (S.cdyn[c]) s_dyn2c[S.cdyn[c] - 1] = c;
    __syncthreads();
    const float distance_gate = 1.5f;
    for (int e = tid; e < ndyn * nlast; e += nth) {
      const int r = e / nlast, q = e % nlast, c = s_dyn2c[r];
      const float ddx = S.ccx[c] - cl_last[5 * q], ddy = S.ccy[c] - cl_last[5 * q + 1], ddz = S.ccz[c] - cl_last[5 * q + 2];
      const float d = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
      const int dn = S.csize[c] - static_cast<int>(cl_last[5 * q + 3]);
      const bool gated = (dn < 0 ? -dn : dn) > 100 || d >= distance_gate;
      // matrix_gate is re-derived after the assignment from the same test
      cost[e] = gated ? static_cast<double>(distance_gate * 5000.f) : static_cast<double>(d / distance_gate * 1000.f);
    }
    __syncthreads();
    if (tid < 32) {
      // shortest augmenting paths on the n x mm matrix (n <= mm), rows = smaller side
      const bool tr = ndyn > nlast;
      const int nn = tr ? nlast : ndyn, mm = tr ? ndyn : nlast;
      __shared__ double s_u[kDspMaxDyn + 1], s_v[kDspMaxDyn + 1], s_minv[kDspMaxDyn + 1];
      __shared__ int s_p[kDspMaxDyn + 1], s_way[kDspMaxDyn + 1];
      __shared__ unsigned char s_used[kDspMaxDyn + 1];
      const int lane = tid;
      for (int j = lane; j <= mm; j += 32) { s_v[j] = 0.0; s_p[j] = 0; s_way[j] = 0; }
      for (int i = lane; i <= nn; i += 32) s_u[i] = 0.0;
      __syncwarp();
      for (int i = 1; i <= nn; ++i) {
        if (lane == 0) s_p[0] = i;
        for (int j = lane; j <= mm; j += 32) { s_minv[j] = CUDART_INF; s_used[j] = 0; }
        __syncwarp();
        int j0 = 0;
        while (true) {
          if (lane == 0) s_used[j0] = 1;
          __syncwarp();
          const int i0 = s_p[j0];
          double best = CUDART_INF;
          int bj = 0x7fffffff;
          for (int j = 1 + lane; j <= mm; j += 32)
            if (!s_used[j]) {
              const double aij = tr ? cost[static_cast<size_t>(j - 1) * nlast + (i0 - 1)]
                                    : cost[static_cast<size_t>(i0 - 1) * nlast + (j - 1)];
              const double cur = aij - s_u[i0] - s_v[j];
              if (cur < s_minv[j]) { s_minv[j] = cur; s_way[j] = j0; }
              if (s_minv[j] < best) { best = s_minv[j]; bj = j; }
            }
          for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
          }
          const double delta = best;
          const int j1 = bj;
          __syncwarp();
          for (int j = lane; j <= mm; j += 32) {
            if (s_used[j]) { s_u[s_p[j]] += delta; s_v[j] -= delta; } else { s_minv[j] -= delta; }
          }
          __syncwarp();
          j0 = j1;
          if (s_p[j0] == 0) break;
        }
        if (lane == 0) {
          do { const int j1 = s_way[j0]; s_p[j0] = s_p[j1]; j0 = j1; } while (j0);
        }
        __syncwarp();
      }
      for (int r = lane; r < ndyn; r += 32) s_asg[r] = -1;
      __syncwarp();
      for (int j = 1 + lane; j <= mm; j += 32)
        if (s_p[j] > 0) { if (tr) s_asg[j - 1] = s_p[j] - 1; else s_asg[s_p[j] - 1] = j - 1; }
      __syncwarp();
    }
    __syncthreads();
    if (tid < ndyn) {
      const int q = s_asg[tid], c = s_dyn2c[tid];
      if (q >= 0) {
        const float ddx = S.ccx[c] - cl_last[5 * q], ddy = S.ccy[c] - cl_last[5 * q + 1], ddz = S.ccz[c] - cl_last[5 * q + 2];
        const float d = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
        const int dn = S.csize[c] - static_cast<int>(cl_last[5 * q + 3]);
        const bool gated = (dn < 0 ? -dn : dn) > 100 || d >= distance_gate;
        if (!gated) {
          float vx = ddx / dtl, vy = ddy / dtl, vz = ddz / dtl;
          const float v = sqrtf(vx * vx + vy * vy + vz * vz);
          if (v > 5.f) { vx = 0.f; vy = 0.f; vz = 0.f; }
          S.cvx[c] = vx; S.cvy[c] = vy; S.cvz[c] = vz;
          S.cint[c] = cl_last[5 * q + 4];
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  // 7. output in the reference's order: dynamic-cluster points (cluster order, member order), ground
  //    points (cloud order), static-cluster points (cluster order, member order); small clusters vanish
  const int n_dyn_pts = s_nc, n_ground = s_ngr;
  for (int i = tid; i < m; i += nth) {
    const int c = S.aux[S.parent[i]];
    if (c < 0) continue;
    const bool dyn = S.cdyn[c] != 0;
    const int pos = dyn ? S.coff[c] + S.mrank[i] : n_dyn_pts + n_ground + S.coff[c] + S.mrank[i];
    float* o = cv + 7 * pos;
    o[0] = S.gx[i]; o[1] = S.gy[i]; o[2] = S.gz[i];
    if (dyn) { o[3] = S.cint[c]; o[4] = S.cvx[c]; o[5] = S.cvy[c]; o[6] = S.cvz[c]; }
    else { o[3] = 0.f; o[4] = 0.f; o[5] = 0.f; o[6] = 0.f; }
  }
  int carry = 0;
  for (int base = 0; base < n; base += nth) {
    const int i = base + tid;
    float x = 0.f, y = 0.f, z = 0.f;
    bool gr = false;
    if (i < n) {
      x = rot[3 * i] + px0; y = rot[3 * i + 1] + py0; z = rot[3 * i + 2] + pz0;
      gr = !(z > G.filt);
    }
    int tot;
    const int r = carry + block_scan_excl(gr ? 1 : 0, s_w, &tot);
    if (gr) {
      float* o = cv + 7 * (n_dyn_pts + r);
      o[0] = x; o[1] = y; o[2] = z; o[3] = 0.f; o[4] = 0.f; o[5] = 0.f; o[6] = 0.f;
    }
    carry += tot;
  }
  if (tid == 0) ncv[a] = n_dyn_pts + n_ground + s_nstat_c;
  // 8. clusters_feature_vector_dynamic_last = this frame's possibly-dynamic clusters
  for (int c = tid; c < nc; c += nth)
    if (S.cdyn[c] && S.cdyn[c] - 1 < kDspMaxDyn) {
      float* o = cl_last + 5 * (S.cdyn[c] - 1);
      o[0] = S.ccx[c]; o[1] = S.ccy[c]; o[2] = S.ccz[c]; o[3] = static_cast<float>(S.csize[c]); o[4] = S.cint[c];
    }
  if (tid == 0) ncl_last[a] = ndyn < kDspMaxDyn ? ndyn : kDspMaxDyn;
}

// ---------------------------------------------------------------------------------------------
// A4  mapPrediction (:663-748) + moveParticle (:1281-1353) as a re-binning of all live particles.
// P1: move every live particle, emit a record (sweep time, destination voxel, pyramid, payload), link it
//     into the destination voxel's list, count it in its pyramid.
__global__ void __launch_bounds__(256) k_dsp_pred_move(DspGeo G, const DspFrame* __restrict__ fr,
                                                       const float* __restrict__ bp_all,
                                                       const uint8_t* __restrict__ flags_all,
                                                       const float4* __restrict__ pa_all,
                                                       const float4* __restrict__ pb_all, int* __restrict__ nrec,
                                                       unsigned* __restrict__ r_key, int* __restrict__ r_dest,
                                                       int* __restrict__ r_next, int* __restrict__ r_pyr,
                                                       uint8_t* __restrict__ r_slot, uint8_t* __restrict__ r_state,
                                                       float4* __restrict__ r_a, float4* __restrict__ r_b,
                                                       int* __restrict__ vhead_all, int* __restrict__ pcount_all,
                                                       int* __restrict__ stats, int* __restrict__ err,
                                                       const int* __restrict__ act_all, const int* __restrict__ nact,
                                                       int* __restrict__ touch_all, int* __restrict__ ntouch) {
  const int a = blockIdx.y;
  if (!fr[a].ok) return;
  const int* act = act_all + static_cast<size_t>(a) * G.V;
  int* touch = touch_all + static_cast<size_t>(a) * G.V;
  const int na = nact[a];
  // 16 lanes per active voxel, one slot each
  const int lane16 = threadIdx.x & 15;
  for (int ai = (blockIdx.x * blockDim.x + threadIdx.x) >> 4; ai < na; ai += (gridDim.x * blockDim.x) >> 4) {
  const int v = act[ai];
  const uint4 f = *reinterpret_cast<const uint4*>(flags_all + (static_cast<size_t>(a) * G.V + v) * kDspSlotStride);
  if ((f.x | f.y | f.z | f.w) == 0u) continue;
  const int nplanes = G.nph + 1 + G.npv + 1;
  const float* bph = bp_all + static_cast<size_t>(a) * nplanes * 3;
  const float* bpv = bph + 3 * (G.nph + 1);
  const float dt = fr[a].dt, dx = fr[a].d[0], dy = fr[a].d[1], dz = fr[a].d[2];
  const size_t rbase = static_cast<size_t>(a) * G.RCAP;
  int* vhead = vhead_all + static_cast<size_t>(a) * G.V;
  int* pcount = pcount_all + static_cast<size_t>(a) * (G.NP + 1);
  for (int p = lane16; p < G.S; p += 16) {
    const unsigned fl = flag_byte(f, p);
    if (fl != DSP_F_VALID && fl != DSP_F_COPY) continue;  // 0.1 < flag < 6
    const size_t si = (static_cast<size_t>(a) * G.V + v) * G.S + p;
    float4 A = pa_all[si], B = pb_all[si];
    // velocity noise is added only when |vx*vy*vz| >= 1e-6; vz is always 0 here (LIMIT_MOVEMENT_IN_XY_PLANE),
    // so the branch is never taken (:689-695) — a non-zero vz would need the sequential noise counter
    if (!(fabsf(A.x * A.y * A.z) < 1e-6)) atomicOr(err, 8);
    A.z = 0.f;
    B.x += dt * A.x + dx;
    B.y += dt * A.y + dy;
    B.z += dt * A.z + dz;
    int dest;
    if (!dsp_vidx(G, B.x, B.y, B.z, &dest)) { atomicAdd(&stats[8 * a + 2], 1); continue; }
    const int r = atomicAdd(&nrec[a], 1);
    if (r >= G.RCAP) { atomicOr(err, 16); continue; }
    const int py = dsp_pyramid(G, bph, bpv, B.x, B.y, B.z);
    r_key[rbase + r] = static_cast<unsigned>(v) * G.S + p;
    r_dest[rbase + r] = dest;
    r_pyr[rbase + r] = py;
    r_slot[rbase + r] = static_cast<uint8_t>(p);
    r_state[rbase + r] = 3;  // bit0 placed in its voxel, bit1 admitted to its pyramid
    r_a[rbase + r] = A;
    r_b[rbase + r] = B;
    const int old = atomicExch(&vhead[dest], r);
    r_next[rbase + r] = old;
    if (old < 0) touch[atomicAdd(&ntouch[a], 1)] = dest;   // first arrival: the voxel joins the touched list (<= V entries)
    if (py >= 0) atomicAdd(&pcount[py], 1);
  }
  }
}

// P2: exclusive scan of the pyramid candidate counts (one CTA per agent)
__global__ void __launch_bounds__(1024) k_dsp_pyr_scan(DspGeo G, const DspFrame* __restrict__ fr,
                                                       const int* __restrict__ pcount_all, int* __restrict__ poff_all,
                                                       int* __restrict__ pcur_all) {
  const int a = blockIdx.x, tid = threadIdx.x;
  if (!fr[a].ok) return;
  __shared__ int s_w[33];
  const int* pcount = pcount_all + static_cast<size_t>(a) * (G.NP + 1);
  int* poff = poff_all + static_cast<size_t>(a) * (G.NP + 1);
  int* pcur = pcur_all + static_cast<size_t>(a) * (G.NP + 1);
  int carry = 0;
  for (int base = 0; base < G.NP; base += blockDim.x) {
    const int i = base + tid;
    const int c = i < G.NP ? pcount[i] : 0;
    int tot;
    const int e = carry + block_scan_excl(c, s_w, &tot);
    if (i < G.NP) { poff[i] = e; pcur[i] = 0; }
    carry += tot;
  }
  if (tid == 0) poff[G.NP] = carry;
}

// P3: group the records by pyramid
__global__ void __launch_bounds__(256) k_dsp_pyr_fill(DspGeo G, const DspFrame* __restrict__ fr,
                                                      const int* __restrict__ nrec, const int* __restrict__ r_pyr,
                                                      const int* __restrict__ poff_all, int* __restrict__ pcur_all,
                                                      int* __restrict__ plist) {
  const int a = blockIdx.y;
  if (!fr[a].ok) return;
  const int n = nrec[a] < G.RCAP ? nrec[a] : G.RCAP;
  const size_t rbase = static_cast<size_t>(a) * G.RCAP;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const int py = r_pyr[rbase + r];
    if (py < 0) continue;
    const int pos = poff_all[static_cast<size_t>(a) * (G.NP + 1) + py] +
                    atomicAdd(&pcur_all[static_cast<size_t>(a) * (G.NP + 1) + py], 1);
    plist[rbase + pos] = r;
  }
}

// P4: order every pyramid's candidates by sweep time (warp per pyramid, rank by counting)
__global__ void __launch_bounds__(256) k_dsp_pyr_sort(DspGeo G, const DspFrame* __restrict__ fr,
                                                      const unsigned* __restrict__ r_key,
                                                      const int* __restrict__ poff_all, const int* __restrict__ plist,
                                                      int* __restrict__ plist2) {
  const int a = blockIdx.y;
  if (!fr[a].ok) return;
  const int lane = threadIdx.x & 31;
  const size_t rbase = static_cast<size_t>(a) * G.RCAP;
  const int* poff = poff_all + static_cast<size_t>(a) * (G.NP + 1);
  for (int py = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; py < G.NP; py += (gridDim.x * blockDim.x) >> 5) {
    const int b = poff[py], e = poff[py + 1];
    for (int i = b + lane; i < e; i += 32) {
      const int r = plist[rbase + i];
      const unsigned key = r_key[rbase + r];
      int rank = 0;
      for (int j = b; j < e; j++) rank += (r_key[rbase + plist[rbase + j]] < key);
      plist2[rbase + b + rank] = r;
    }
  }
}

// PV: voxel placement of one round.  Thread per destination voxel: arrivals from earlier voxels, then the
// voxel's own sweep (leavers and pyramid-dropped stayers free their slots), then arrivals from later voxels;
// each arrival takes the lowest free slot at its time.
__global__ void __launch_bounds__(256) k_dsp_pred_voxel(DspGeo G, const DspFrame* __restrict__ fr, int round,
                                                        const int* __restrict__ changed_all,
                                                        const uint8_t* __restrict__ flags_all,
                                                        const unsigned* __restrict__ r_key,
                                                        const int* __restrict__ r_next, const int* __restrict__ r_pyr,
                                                        uint8_t* __restrict__ r_slot, uint8_t* __restrict__ r_state,
                                                        const int* __restrict__ vhead_all,
                                                        const int* __restrict__ touch_all,
                                                        const int* __restrict__ ntouch) {
  const int a = blockIdx.y;
  if (!fr[a].ok) return;
  if (round > 1 && changed_all[a * (kDspMaxIter + 2) + round - 1] == 0) return;
  const int nt = ntouch[a];
  for (int ti = blockIdx.x * blockDim.x + threadIdx.x; ti < nt; ti += gridDim.x * blockDim.x) {
  const int u = touch_all[static_cast<size_t>(a) * G.V + ti];
  const int head = vhead_all[static_cast<size_t>(a) * G.V + u];
  if (head < 0) continue;
  const size_t rbase = static_cast<size_t>(a) * G.RCAP;
  const uint4 f = *reinterpret_cast<const uint4*>(flags_all + (static_cast<size_t>(a) * G.V + u) * kDspSlotStride);
  unsigned mask = 0, proc = 0;
  for (int p = 0; p < G.S; p++) {
    const unsigned fl = flag_byte(f, p);
    if (fl != DSP_F_EMPTY) mask |= 1u << p;
    if (fl == DSP_F_VALID || fl == DSP_F_COPY) proc |= 1u << p;
  }
  const unsigned full = (1u << G.S) - 1u;
  const unsigned own_lo = static_cast<unsigned>(u) * G.S;
  bool swept = false;
  long long last = -1;
  while (true) {
    // next record of this voxel's list in ascending sweep time
    int best = -1;
    unsigned bkey = 0xffffffffu;
    for (int r = head; r >= 0; r = r_next[rbase + r]) {
      const unsigned k = r_key[rbase + r];
      if (static_cast<long long>(k) > last && k <= bkey) { bkey = k; best = r; }
    }
    if (best < 0) break;
    last = bkey;
    const bool own = bkey >= own_lo && bkey < own_lo + G.S;
    if (!swept && bkey >= own_lo) {  // the voxel's own sweep starts: every processed particle leaves its slot
      mask &= ~proc;
      swept = true;
    }
    const unsigned st = r_state[rbase + best];
    const bool admitted = r_pyr[rbase + best] < 0 || (st & 2u);
    if (own) {
      // stayer: keeps its slot unless its pyramid was full
      if (admitted) mask |= 1u << (bkey - own_lo);
      r_state[rbase + best] = static_cast<uint8_t>((st & 2u) | 1u);
    } else {
      const unsigned freeb = ~mask & full;
      if (freeb == 0u) {
        r_state[rbase + best] = static_cast<uint8_t>(st & 2u);
      } else {
        const int s = __ffs(freeb) - 1;
        r_slot[rbase + best] = static_cast<uint8_t>(s);
        r_state[rbase + best] = static_cast<uint8_t>((st & 2u) | 1u);
        if (admitted) mask |= 1u << s;
      }
    }
  }
  }
}

// PP: pyramid admission of one round (warp per pyramid): the first PCAP placed candidates in sweep order
__global__ void __launch_bounds__(256) k_dsp_pred_pyr(DspGeo G, const DspFrame* __restrict__ fr, int round,
                                                      int* __restrict__ changed_all, const int* __restrict__ poff_all,
                                                      const int* __restrict__ plist2, uint8_t* __restrict__ r_state) {
  const int a = blockIdx.y;
  if (!fr[a].ok) return;
  int* changed = changed_all + a * (kDspMaxIter + 2);
  if (round > 1 && changed[round - 1] == 0) return;
  const int lane = threadIdx.x & 31;
  const size_t rbase = static_cast<size_t>(a) * G.RCAP;
  const int* poff = poff_all + static_cast<size_t>(a) * (G.NP + 1);
  int diff = 0;
  for (int py = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; py < G.NP; py += (gridDim.x * blockDim.x) >> 5) {
  const int b = poff[py], e = poff[py + 1];
  int running = 0;
  for (int i0 = b; i0 < e; i0 += 32) {
    const int i = i0 + lane;
    int r = -1;
    unsigned st = 0;
    if (i < e) { r = plist2[rbase + i]; st = r_state[rbase + r]; }
    const bool placed = (st & 1u) != 0u;
    const unsigned bal = __ballot_sync(0xffffffffu, placed);
    const int rank = running + __popc(bal & ((1u << lane) - 1u));
    if (r >= 0) {
      const unsigned adm = (!placed || rank < G.PCAP) ? 2u : 0u;
      if (adm != (st & 2u)) { diff++; r_state[rbase + r] = static_cast<uint8_t>((st & 1u) | adm); }
    }
    running += __popc(bal);
  }
  }
  for (int o = 16; o > 0; o >>= 1) diff += __shfl_xor_sync(0xffffffffu, diff, o);
  if (lane == 0 && diff) atomicAdd(&changed[round], diff);
}

// PF: commit.  Thread per voxel rewrites its own flags and payloads from the records.
__global__ void __launch_bounds__(256) k_dsp_pred_commit(DspGeo G, const DspFrame* __restrict__ fr,
                                                         const int* __restrict__ changed_all,
                                                         uint8_t* __restrict__ flags_all, float4* __restrict__ pa_all,
                                                         float4* __restrict__ pb_all,
                                                         const unsigned* __restrict__ r_key,
                                                         const int* __restrict__ r_next, const int* __restrict__ r_pyr,
                                                         const uint8_t* __restrict__ r_slot,
                                                         const uint8_t* __restrict__ r_state,
                                                         const float4* __restrict__ r_a, const float4* __restrict__ r_b,
                                                         int* __restrict__ vhead_all, int* __restrict__ stats,
                                                         int* __restrict__ err, const int* __restrict__ list_all,
                                                         const int* __restrict__ nlist, int touched_pass) {
  const int a = blockIdx.y;
  if (!fr[a].ok) return;
  if (blockIdx.x == 0 && threadIdx.x == 0 && changed_all[a * (kDspMaxIter + 2) + kDspMaxIter] != 0)
    atomicOr(err, 32);  // no fixed point
  const int nl = nlist[a];
  for (int li = blockIdx.x * blockDim.x + threadIdx.x; li < nl; li += gridDim.x * blockDim.x) {
  const int u = list_all[static_cast<size_t>(a) * G.V + li];
  uint4* fp = reinterpret_cast<uint4*>(flags_all + (static_cast<size_t>(a) * G.V + u) * kDspSlotStride);
  uint4 f = *fp;
  const int head = vhead_all[static_cast<size_t>(a) * G.V + u];
  // pass 1 (touched list, runs first): voxels that held nothing; pass 2 (active list): every voxel that held something
  if (touched_pass && (f.x | f.y | f.z | f.w) != 0u) continue;
  if (head < 0 && (f.x | f.y | f.z | f.w) == 0u) continue;
  bool dirty = false;
  for (int p = 0; p < G.S; p++) {
    const unsigned fl = flag_byte(f, p);
    if (fl == DSP_F_VALID || fl == DSP_F_COPY) { set_flag_byte(f, p, DSP_F_EMPTY); dirty = true; }
  }
  const size_t rbase = static_cast<size_t>(a) * G.RCAP;
  const unsigned own_lo = static_cast<unsigned>(u) * G.S;
  int n_vox = 0, n_pyr = 0;
  for (int r = head; r >= 0; r = r_next[rbase + r]) {
    const unsigned st = r_state[rbase + r];
    if (!(st & 1u)) { n_vox++; continue; }
    if (r_pyr[rbase + r] >= 0 && !(st & 2u)) { n_pyr++; continue; }
    const unsigned k = r_key[rbase + r];
    const bool own = k >= own_lo && k < own_lo + G.S;
    const int s = r_slot[rbase + r];
    set_flag_byte(f, s, own ? DSP_F_VALID : DSP_F_MOVED);
    const size_t si = (static_cast<size_t>(a) * G.V + u) * G.S + s;
    pa_all[si] = r_a[rbase + r];
    pb_all[si] = r_b[rbase + r];
    dirty = true;
  }
  if (dirty) *fp = f;
  if (head >= 0) vhead_all[static_cast<size_t>(a) * G.V + u] = -1;
  if (n_vox) atomicAdd(&stats[8 * a + 0], n_vox);
  if (n_pyr) atomicAdd(&stats[8 * a + 1], n_pyr);
  }
}

// pyramids_in_fov: the admitted particles of every pyramid in sweep order, with a copy of their state
__global__ void __launch_bounds__(256) k_dsp_pyr_commit(DspGeo G, const DspFrame* __restrict__ fr,
                                                        const int* __restrict__ poff_all,
                                                        const int* __restrict__ plist2,
                                                        const uint8_t* __restrict__ r_state,
                                                        const int* __restrict__ r_dest,
                                                        const uint8_t* __restrict__ r_slot,
                                                        const float4* __restrict__ r_a, const float4* __restrict__ r_b,
                                                        int* __restrict__ pyr_n, int2* __restrict__ pyr_ref,
                                                        float4* __restrict__ pyr_p) {
  const int a = blockIdx.y;
  if (!fr[a].ok) return;
  const int lane = threadIdx.x & 31;
  const size_t rbase = static_cast<size_t>(a) * G.RCAP;
  const int* poff = poff_all + static_cast<size_t>(a) * (G.NP + 1);
  for (int py = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; py < G.NP; py += (gridDim.x * blockDim.x) >> 5) {
  const int b = poff[py], e = poff[py + 1];
  const size_t tb = (static_cast<size_t>(a) * G.NP + py) * G.PCAP;
  int running = 0;
  for (int i0 = b; i0 < e && running < G.PCAP; i0 += 32) {
    const int i = i0 + lane;
    int r = -1;
    unsigned st = 0;
    if (i < e) { r = plist2[rbase + i]; st = r_state[rbase + r]; }
    const bool in = (st & 3u) == 3u;
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    const int rank = running + __popc(bal & ((1u << lane) - 1u));
    if (in && rank < G.PCAP) {
      pyr_ref[tb + rank] = make_int2(r_dest[rbase + r], r_slot[rbase + r]);
      const float4 A = r_a[rbase + r], B = r_b[rbase + r];
      pyr_p[tb + rank] = make_float4(B.x, B.y, B.z, A.w);
    }
    running += __popc(bal);
  }
  if (lane == 0) pyr_n[static_cast<size_t>(a) * G.NP + py] = running < G.PCAP ? running : G.PCAP;
  }
}

// ---------------------------------------------------------------------------------------------
// A5  mapUpdate (:750-850)
// U1: C_k of every stored observation: kappa + expected newborn + sum over the particles of the 3x3
// neighbouring pyramids, in the reference's loop order (one thread per observation).
__global__ void __launch_bounds__(128) k_dsp_update_ck(DspGeo G, const DspFrame* __restrict__ fr,
                                                       const int* __restrict__ npts, const int* __restrict__ pobs_all,
                                                       float4* __restrict__ obs_all, const int* __restrict__ valid,
                                                       const int* __restrict__ pyr_n, const float4* __restrict__ pyr_p,
                                                       const float* __restrict__ pdf) {
  const int a = blockIdx.y;
  if (!fr[a].ok) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts[a]) return;
  const int po = pobs_all[static_cast<size_t>(a) * G.MAXPTS + i];
  if (po < 0) return;
  const int py = po / G.OCAP;
  float4* z = obs_all + static_cast<size_t>(a) * G.NP * G.OCAP + po;
  const float zx = z->x, zy = z->y, zz = z->z;
  float ck = 0.f;
  const int h0 = py / G.npv, v0 = py % G.npv;
  for (int di = -1; di <= 1; ++di)
    for (int dj = -1; dj <= 1; ++dj) {
      const int hh = h0 + di, vv = v0 + dj;
      if (!(hh >= 0 && hh < G.nph && vv >= 0 && vv < G.npv)) continue;
      const int pi = hh * G.npv + vv;
      const int cnt = pyr_n[static_cast<size_t>(a) * G.NP + pi];
      const float4* pp = pyr_p + (static_cast<size_t>(a) * G.NP + pi) * G.PCAP;
      for (int s = 0; s < cnt; ++s) {
        const float4 q = pp[s];
        const float gk = dsp_pdf(pdf, q.x, zx, G.sigma) * dsp_pdf(pdf, q.y, zy, G.sigma) * dsp_pdf(pdf, q.z, zz, G.sigma);
        ck += G.pd * q.w * gk;
      }
    }
  const float expected = G.nbw * static_cast<float>(valid[a]) * static_cast<float>(G.NB);
  ck += (expected + G.kappa);
  z->w = ck;
}

// U2: weight update of every particle in a pyramid (one thread per pyramid entry, sums in loop order)
__global__ void __launch_bounds__(128) k_dsp_update_w(DspGeo G, const DspFrame* __restrict__ fr,
                                                      const float4* __restrict__ obs_all,
                                                      const int* __restrict__ obs_num_all,
                                                      const float* __restrict__ maxlen_all,
                                                      const int* __restrict__ pyr_n, const int2* __restrict__ pyr_ref,
                                                      const float4* __restrict__ pyr_p, float4* __restrict__ pa_all,
                                                      const float* __restrict__ pdf) {
  const int a = blockIdx.y;
  if (!fr[a].ok) return;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= G.NP * G.PCAP) return;
  const int py = e / G.PCAP, s = e % G.PCAP;
  if (s >= pyr_n[static_cast<size_t>(a) * G.NP + py]) return;
  const float4 q = pyr_p[static_cast<size_t>(a) * G.NP * G.PCAP + e];
  const float dist = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
  const float ml = maxlen_all[static_cast<size_t>(a) * G.NP + py];
  if (ml > 0.f && dist > ml + G.thick) return;  // occluded
  const float4* obs = obs_all + static_cast<size_t>(a) * G.NP * G.OCAP;
  const int* obs_num = obs_num_all + static_cast<size_t>(a) * G.NP;
  float sum = 0.f;
  const int h0 = py / G.npv, v0 = py % G.npv;
  for (int di = -1; di <= 1; ++di)
    for (int dj = -1; dj <= 1; ++dj) {
      const int hh = h0 + di, vv = v0 + dj;
      if (!(hh >= 0 && hh < G.nph && vv >= 0 && vv < G.npv)) continue;
      const int ni = hh * G.npv + vv;
      const int cnt = obs_num[ni];
      const float4* zz = obs + static_cast<size_t>(ni) * G.OCAP;
      for (int k = 0; k < cnt; ++k) {
        const float4 z = zz[k];
        const float gk = dsp_pdf(pdf, q.x, z.x, G.sigma) * dsp_pdf(pdf, q.y, z.y, G.sigma) * dsp_pdf(pdf, q.z, z.z, G.sigma);
        sum += G.pd * gk / z.w;
      }
    }
  const int2 ref = pyr_ref[static_cast<size_t>(a) * G.NP * G.PCAP + e];
  float4* A = pa_all + (static_cast<size_t>(a) * G.V + ref.x) * G.S + ref.y;
  A->w = q.w * ((1 - G.pd) + sum);
}

// ---------------------------------------------------------------------------------------------
// A7  mapAddNewBornParticlesByObservation (:853-991).  One CTA per agent computes, for every point of
// input_cloud_with_velocity in order, its Dempster-Shafer static share, the noise-table indices its 20
// particles consume (prefix sums of the draw counts) and the newborn records, linked per destination voxel.
__global__ void __launch_bounds__(1024) k_dsp_born(DspGeo G, const DspFrame* __restrict__ fr,
                                                   const float* __restrict__ cv_all, const int* __restrict__ ncv,
                                                   const float4* __restrict__ obs_all,
                                                   const int* __restrict__ obs_num_all,
                                                   const uint8_t* __restrict__ flags_all,
                                                   const float4* __restrict__ pa_all, const float* __restrict__ pg,
                                                   const float* __restrict__ vg, const int* __restrict__ ur,
                                                   int* __restrict__ seq_all, int* __restrict__ b_dest,
                                                   int* __restrict__ b_next, float4* __restrict__ b_a,
                                                   float4* __restrict__ b_b, int* __restrict__ vhead_all,
                                                   int* __restrict__ err, int* __restrict__ touch2_all,
                                                   int* __restrict__ ntouch2) {
  extern __shared__ unsigned char smem_raw[];
  const int a = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  if (!fr[a].ok) return;
  __shared__ int s_w[33];
  __shared__ float s_norm;
  float* s_inv = reinterpret_cast<float*>(smem_raw);   // MAXPTS: 1/C_k in (pyramid, slot) order
  // normalisation: sum of 1/C_k over pyramids and slots in order (:855-861)
  const float4* obs = obs_all + static_cast<size_t>(a) * G.NP * G.OCAP;
  const int* obs_num = obs_num_all + static_cast<size_t>(a) * G.NP;
  int carry = 0;
  for (int base = 0; base < G.NP; base += nth) {
    const int i = base + tid;
    const int c = i < G.NP ? obs_num[i] : 0;
    int tot;
    const int off = carry + block_scan_excl(c, s_w, &tot);
    for (int j = 0; j < c; j++)
      if (off + j < G.MAXPTS) s_inv[off + j] = 1.f / obs[static_cast<size_t>(i) * G.OCAP + j].w;
    carry += tot;
  }
  __syncthreads();
  if (tid == 0) {
    float norm = 0.f;
    const int tot = carry < G.MAXPTS ? carry : G.MAXPTS;
    for (int k = 0; k < tot; k++) norm += s_inv[k];
    s_norm = norm;
  }
  __syncthreads();
  const float w_born = G.nbw * s_norm;
  const int n = ncv[a];
  const float* cv = cv_all + static_cast<size_t>(a) * G.MAXPTS * 7;
  int* seq = seq_all + 4 * a;
  const int pseq0 = seq[0], vseq0 = seq[1], useq0 = seq[2];
  const float c0 = fr[a].pos[0], c1 = fr[a].pos[1], c2 = fr[a].pos[2];
  const int min_static = static_cast<int>(static_cast<float>(G.NB) * 0.15f);
  const int model_n = static_cast<int>(static_cast<float>(G.NB) * 0.8f);
  const size_t bbase = static_cast<size_t>(a) * G.MAXPTS * G.NB;
  int* vhead = vhead_all + static_cast<size_t>(a) * G.V;
  int carry_in = 0, carry_v = 0, carry_u = 0;
  for (int base = 0; base < n; base += nth) {
    const int q = base + tid;
    bool in = false;
    float cx = 0.f, cy = 0.f, cz = 0.f, inten = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
    int n_static = 0;
    if (q < n) {
      const float* p = cv + 7 * q;
      cx = p[0] - c0; cy = p[1] - c1; cz = p[2] - c2;
      inten = p[3]; nx = p[4]; ny = p[5]; nz = p[6];
      int pv;
      if (dsp_vidx(G, cx, cy, cz, &pv)) {
        in = true;
        const uint4 f = *reinterpret_cast<const uint4*>(flags_all + (static_cast<size_t>(a) * G.V + pv) * kDspSlotStride);
        float ws = 0.f, wd = 0.f, wsd = 0.f;
        for (int kk = 0; kk < G.S; ++kk) {
          const unsigned fl = flag_byte(f, kk);
          if (fl == DSP_F_VALID || fl == DSP_F_MOVED) {  // 0.9 < flag < 14
            const float4 A = pa_all[(static_cast<size_t>(a) * G.V + pv) * G.S + kk];
            const float vabs = fabsf(A.x) + fabsf(A.y) + fabsf(A.z);
            if (vabs < 0.1f) ws += A.w; else if (vabs < 0.5f) wsd += A.w; else wd += A.w;
          }
        }
        const float tot = ws + wd + wsd;
        const float ms = ws / tot, md = wd / tot, msd = wsd / tot;
        const float p_s = (ms + ms + msd) * 0.5f, p_d = (md + md + msd) * 0.5f;
        const float psn = p_s / (p_s + p_d);
        const float fs = static_cast<float>(model_n) * psn;
        n_static = isnan(fs) ? min_static : max(min_static, static_cast<int>(fs));  // (int)NaN = INT_MIN on x86-64
      }
    }
    int tot_in;
    const int rank = carry_in + block_scan_excl(in ? 1 : 0, s_w, &tot_in);
    // pass 1: perturbed positions, destination voxels, draw counts
    int nv = 0, nu = 0;
    unsigned inmask = 0;
    if (in) {
      for (int p = 0; p < G.NB; p++) {
        const long long pi = static_cast<long long>(pseq0) + 3ll * (static_cast<long long>(G.NB) * rank + p);
        const float px = cx + pg[pi % G.nrand], py = cy + pg[(pi + 1) % G.nrand], pz = cz + pg[(pi + 2) % G.nrand];
        int vi;
        if (dsp_vidx(G, px, py, pz, &vi)) {
          inmask |= 1u << p;
          if (p < n_static) {
          } else if (nx > -100.f && p < model_n) {
            if (inten > 0.01f) nv += 3;
          } else {
            if (inten > 0.01f) nu += 3;
          }
        }
      }
    }
    int tot_v, tot_u;
    int ov = carry_v + block_scan_excl(nv, s_w, &tot_v);
    int ou = carry_u + block_scan_excl(nu, s_w, &tot_u);
    // pass 2: velocities and records
    if (in) {
      for (int p = 0; p < G.NB; p++) {
        const size_t bi = bbase + static_cast<size_t>(q) * G.NB + p;
        if (!(inmask & (1u << p))) { b_dest[bi] = -1; continue; }
        const long long pi = static_cast<long long>(pseq0) + 3ll * (static_cast<long long>(G.NB) * rank + p);
        const float px = cx + pg[pi % G.nrand], py = cy + pg[(pi + 1) % G.nrand], pz = cz + pg[(pi + 2) % G.nrand];
        int vi = 0;
        dsp_vidx(G, px, py, pz, &vi);
        float vx = 0.f, vy = 0.f, vz = 0.f;
        if (p < n_static) {
        } else if (nx > -100.f && p < model_n) {
          if (inten > 0.01f) {
            vx = nx + 4 * vg[(static_cast<long long>(vseq0) + ov) % G.nrand];
            vy = ny + 4 * vg[(static_cast<long long>(vseq0) + ov + 1) % G.nrand];
            vz = nz + 4 * vg[(static_cast<long long>(vseq0) + ov + 2) % G.nrand];
            ov += 3;
          }
        } else {
          if (inten > 0.01f) {
            vx = dsp_rand_float(ur[(static_cast<long long>(useq0) + ou) % G.nrand], -1.5f, 1.5f);
            vy = dsp_rand_float(ur[(static_cast<long long>(useq0) + ou + 1) % G.nrand], -1.5f, 1.5f);
            vz = dsp_rand_float(ur[(static_cast<long long>(useq0) + ou + 2) % G.nrand], -0.5f, 0.5f);
            ou += 3;
          }
        }
        vz = 0.f;  // LIMIT_MOVEMENT_IN_XY_PLANE
        b_dest[bi] = vi;
        b_a[bi] = make_float4(vx, vy, vz, w_born);
        b_b[bi] = make_float4(px, py, pz, 0.f);
        const int old = atomicExch(&vhead[vi], static_cast<int>(static_cast<size_t>(q) * G.NB + p));
        b_next[bi] = old;
        if (old < 0) touch2_all[static_cast<size_t>(a) * G.V + atomicAdd(&ntouch2[a], 1)] = vi;
      }
    } else if (q < n) {
      for (int p = 0; p < G.NB; p++) b_dest[bbase + static_cast<size_t>(q) * G.NB + p] = -1;
    }
    carry_in += tot_in; carry_v += tot_v; carry_u += tot_u;
  }
  if (G.NB > 32 && tid == 0) atomicOr(err, 64);
  __syncthreads();
  if (tid == 0) {
    seq[0] = static_cast<int>((static_cast<long long>(pseq0) + 3ll * G.NB * carry_in) % G.nrand);
    seq[1] = static_cast<int>((static_cast<long long>(vseq0) + carry_v) % G.nrand);
    seq[2] = static_cast<int>((static_cast<long long>(useq0) + carry_u) % G.nrand);
  }
}

// addAParticle (:1257-1276) for every newborn, in (point, p) order per destination voxel
__global__ void __launch_bounds__(256) k_dsp_born_place(DspGeo G, const DspFrame* __restrict__ fr,
                                                        uint8_t* __restrict__ flags_all, float4* __restrict__ pa_all,
                                                        float4* __restrict__ pb_all, const int* __restrict__ b_next,
                                                        const float4* __restrict__ b_a, const float4* __restrict__ b_b,
                                                        int* __restrict__ vhead_all, int* __restrict__ stats,
                                                        const int* __restrict__ touch2_all,
                                                        const int* __restrict__ ntouch2) {
  const int a = blockIdx.y;
  if (!fr[a].ok) return;
  const int nt = ntouch2[a];
  for (int ti = blockIdx.x * blockDim.x + threadIdx.x; ti < nt; ti += gridDim.x * blockDim.x) {
  const int u = touch2_all[static_cast<size_t>(a) * G.V + ti];
  const int head = vhead_all[static_cast<size_t>(a) * G.V + u];
  if (head < 0) continue;
  vhead_all[static_cast<size_t>(a) * G.V + u] = -1;
  const size_t bbase = static_cast<size_t>(a) * G.MAXPTS * G.NB;
  uint4* fp = reinterpret_cast<uint4*>(flags_all + (static_cast<size_t>(a) * G.V + u) * kDspSlotStride);
  uint4 f = *fp;
  int slot = 0, born = 0;
  int last = -1;
  while (slot < G.S) {
    int best = 0x7fffffff;
    for (int r = head; r >= 0; r = b_next[bbase + r])
      if (r > last && r < best) best = r;
    if (best == 0x7fffffff) break;
    last = best;
    while (slot < G.S && flag_byte(f, slot) != DSP_F_EMPTY) slot++;
    if (slot >= G.S) break;
    set_flag_byte(f, slot, DSP_F_BORN);
    const size_t si = (static_cast<size_t>(a) * G.V + u) * G.S + slot;
    pa_all[si] = b_a[bbase + best];
    pb_all[si] = b_b[bbase + best];
    born++;
  }
  *fp = f;
  if (born) atomicAdd(&stats[8 * a + 3], born);
  }
}

// ---------------------------------------------------------------------------------------------
// A8  mapOccupancyCalculationAndResample (:994-1130): the particle voxel-update.  Thread per voxel:
// weight sum, mean velocity, future-occupancy deposits of every non-newborn particle into the SOGM slices
// (voxels_objects_number[.][4+k] = the [t][z][y][x] f32 grid here), systematic resampling in slot order.
__global__ void __launch_bounds__(256) k_dsp_occupancy(DspGeo G, const DspFrame* __restrict__ fr,
                                                       uint8_t* __restrict__ flags_all, float4* __restrict__ pa_all,
                                                       float4* __restrict__ pb_all, float4* __restrict__ vobj_all,
                                                       uint8_t* __restrict__ grid, unsigned long long agent_stride,
                                                       const int* __restrict__ act_all, const int* __restrict__ nact,
                                                       const int* __restrict__ touch_all, const int* __restrict__ ntouch,
                                                       const int* __restrict__ touch2_all,
                                                       const int* __restrict__ ntouch2, int* __restrict__ vstamp_all,
                                                       int stamp, int* __restrict__ act_next_all,
                                                       int* __restrict__ nact_next) {
  const int a = blockIdx.y;
  const int n0 = nact[a];
  int* act_next = act_next_all + static_cast<size_t>(a) * G.V;
  if (!fr[a].ok) {  // rejected update: the map is untouched, its active list carries over
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0; i += gridDim.x * blockDim.x)
      act_next[i] = act_all[static_cast<size_t>(a) * G.V + i];
    if (blockIdx.x == 0 && threadIdx.x == 0) nact_next[a] = n0;
    return;
  }
  const int n1 = ntouch[a], n2 = ntouch2[a];
  for (int li = blockIdx.x * blockDim.x + threadIdx.x; li < n0 + n1 + n2; li += gridDim.x * blockDim.x) {
  const int v = li < n0 ? act_all[static_cast<size_t>(a) * G.V + li]
                        : li < n0 + n1 ? touch_all[static_cast<size_t>(a) * G.V + li - n0]
                                       : touch2_all[static_cast<size_t>(a) * G.V + li - n0 - n1];
  if (atomicExch(&vstamp_all[static_cast<size_t>(a) * G.V + v], stamp) == stamp) continue;  // listed more than once
  uint4* fp = reinterpret_cast<uint4*>(flags_all + (static_cast<size_t>(a) * G.V + v) * kDspSlotStride);
  uint4 f = *fp;
  float4* vo = vobj_all + static_cast<size_t>(a) * G.V + v;
  if ((f.x | f.y | f.z | f.w) == 0u) { *vo = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
  float* sogm = reinterpret_cast<float*>(grid + a * agent_stride);
  float4* pa = pa_all + (static_cast<size_t>(a) * G.V + v) * G.S;
  float4* pb = pb_all + (static_cast<size_t>(a) * G.V + v) * G.S;
  float wsum = 0.f, vxs = 0.f, vys = 0.f, vzs = 0.f;
  int np = 0, nold = 0;
  float w[kDspSlotStride];
  for (int p = 0; p < G.S; p++) {
    w[p] = 0.f;
    const unsigned fl = flag_byte(f, p);
    if (fl == DSP_F_EMPTY) continue;
    const float4 A = pa[p];
    if (A.w < 1e-3) { set_flag_byte(f, p, DSP_F_EMPTY); continue; }
    if (fl != DSP_F_BORN) {
      ++nold;
      vxs += A.x; vys += A.y; vzs += A.z;
      const float4 B = pb[p];
      for (int k = 0; k < G.T; ++k) {
        const float pt = G.tf[k];
        const float fx = B.x + A.x * pt, fy = B.y + A.y * pt, fz = B.z + A.z * pt;
        int pi;
        if (dsp_vidx(G, fx, fy, fz, &pi)) atomicAdd(sogm + static_cast<size_t>(k) * G.V + pi, A.w);
      }
    }
    set_flag_byte(f, p, DSP_F_VALID);
    ++np;
    wsum += A.w;
    w[p] = A.w;
  }
  float4 o = make_float4(wsum, 0.f, 0.f, 0.f);
  if (nold > 0) { o.y = vxs / static_cast<float>(nold); o.z = vys / static_cast<float>(nold); o.w = vzs / static_cast<float>(nold); }
  *vo = o;
  if (np >= 5) {
    const int n_after = np > G.MAXP ? G.MAXP : np;
    const float w_after = wsum / static_cast<float>(n_after);
    float acc_ori = 0.f, acc_new = w_after * 0.5f;
    for (int p = 0; p < G.S; ++p) {
      if (flag_byte(f, p) != DSP_F_VALID) continue;   // flag > 0.7: excludes empty and resample copies
      acc_ori += w[p];
      if (acc_ori > acc_new) {
        float wp = w_after;
        acc_new += w_after;
        bool fullv = false;
        int p_i = 0;
        float4 A = pa[p];
        const float4 B = pb[p];
        while (acc_ori > acc_new) {
          bool found = false;
          if (!fullv) {
            for (; p_i < G.S; ++p_i) {
              if (flag_byte(f, p_i) == DSP_F_EMPTY) {
                set_flag_byte(f, p_i, DSP_F_COPY);
                // copies the particle as it is at this moment (weight = w_after plus any weight already folded back)
                pa[p_i] = make_float4(A.x, A.y, A.z, wp);
                pb[p_i] = B;
                found = true;
                break;
              }
            }
          }
          if (!found) { wp += w_after; fullv = true; }
          acc_new += w_after;
        }
        A.w = wp;
        pa[p] = A;
      } else {
        set_flag_byte(f, p, DSP_F_EMPTY);
      }
    }
  }
  *fp = f;
  if ((f.x | f.y | f.z | f.w) != 0u) act_next[atomicAdd(&nact_next[a], 1)] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// parity/debug views
__global__ void k_dsp_download_particles(DspGeo G, int a, const uint8_t* __restrict__ flags_all,
                                         const float4* __restrict__ pa_all, const float4* __restrict__ pb_all,
                                         float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(G.V) * G.S) return;
  const int v = static_cast<int>(i / G.S), p = static_cast<int>(i % G.S);
  const unsigned fl = flags_all[(static_cast<size_t>(a) * G.V + v) * kDspSlotStride + p];
  const float ff = fl == DSP_F_VALID ? 1.f : fl == DSP_F_COPY ? 0.6f : fl == DSP_F_MOVED ? 7.f : fl == DSP_F_BORN ? 15.f : 0.f;
  const float4 A = pa_all[(static_cast<size_t>(a) * G.V + v) * G.S + p];
  const float4 B = pb_all[(static_cast<size_t>(a) * G.V + v) * G.S + p];
  float* o = out + i * 8;
  o[0] = ff; o[1] = A.x; o[2] = A.y; o[3] = A.z; o[4] = B.x; o[5] = B.y; o[6] = B.z; o[7] = A.w;
}

// getOccupancyMap (:402-420): centres of the voxels over the threshold, ascending index (one CTA)
__global__ void __launch_bounds__(1024) k_dsp_occupied(DspGeo G, int a, const float4* __restrict__ vobj_all,
                                                       float* __restrict__ out, int cap, int* __restrict__ count) {
  __shared__ int s_w[33];
  const float4* vo = vobj_all + static_cast<size_t>(a) * G.V;
  const float corr_x = -G.hx + G.res * 0.5f, corr_y = -G.hy + G.res * 0.5f, corr_z = -G.hz + G.res * 0.5f;
  int carry = 0;
  for (int base = 0; base < G.V; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const bool occ = i < G.V && vo[i].x > G.occ_thr;
    int tot;
    const int pos = carry + block_scan_excl(occ ? 1 : 0, s_w, &tot);
    if (occ && pos < cap) {
      const int zi = i / (G.nx * G.ny), rem = i - zi * G.nx * G.ny, yi = rem / G.nx, xi = rem - yi * G.nx;
      out[3 * pos] = static_cast<float>(xi) * G.res + corr_x;
      out[3 * pos + 1] = static_cast<float>(yi) * G.res + corr_y;
      out[3 * pos + 2] = static_cast<float>(zi) * G.res + corr_z;
    }
    carry += tot;
  }
  if (threadIdx.x == 0) *count = carry;
}

__global__ void k_dsp_fill_int(int* p, size_t n, int v) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

size_t vel_smem_bytes(const DspGeo& G) {
  return sizeof(float) * 3 * G.MAXPTS + sizeof(int) * 3 * G.MAXPTS + sizeof(int) * 4 * kDspMaxClusters +
         sizeof(float) * 7 * kDspMaxClusters;
}
size_t filter_smem_bytes(const DspGeo& G) {
  return sizeof(unsigned long long) * static_cast<size_t>(G.NPOW2 < kFilterSmemKeys ? G.NPOW2 : kFilterSmemKeys);
}
size_t bin_smem_bytes(const DspGeo& G) {
  const int nplanes = G.nph + 1 + G.npv + 1;
  return sizeof(float) * ((nplanes * 3 + 3) & ~3) + sizeof(int) * 2 * G.NP + sizeof(int) * 1024;
}

}  // namespace

// =============================================================================================
// host side
#define DSP_ALLOC(ptr, count)                                                                          \
  do {                                                                                                 \
    POP_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&(ptr)), sizeof(*(ptr)) * static_cast<size_t>(count))); \
    POP_CUDA(h, cudaMemsetAsync((ptr), 0, sizeof(*(ptr)) * static_cast<size_t>(count), h->stream));     \
  } while (0)

void popk_dsp_free(pop_handle* h) {
  DspDev* d = h->dsp;
  if (!d) return;
  void* ptrs[] = {d->flags, d->pa, d->pb, d->vobj, d->pg, d->vg, d->ur, d->pdf, d->bp_ori, d->frame, d->bp, d->cam,
                  d->fkeys, d->pts, d->npts, d->rot, d->pobs, d->obs, d->obs_num, d->obs_maxlen, d->valid, d->cv,
                  d->ncv, d->cl_last, d->ncl_last, d->cost, d->nrec, d->r_key, d->r_dest, d->r_next, d->r_pyr,
                  d->r_slot, d->r_state, d->r_a, d->r_b, d->vhead, d->pcount, d->poff, d->pcur, d->plist,
                  d->plist2, d->changed, d->pyr_n, d->pyr_ref, d->pyr_p, d->b_dest, d->b_next, d->b_a, d->b_b,
                  d->seq, d->stats, d->err, d->act[0], d->act[1], d->nact[0], d->nact[1], d->touch, d->ntouch,
                  d->touch2, d->ntouch2, d->vstamp, d->ok_dev, d->depth, d->cam_counts, d->rpose, d->rpose_d};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (d->side) cudaStreamDestroy(d->side);
  if (d->ev_fork) cudaEventDestroy(d->ev_fork);
  if (d->ev_join) cudaEventDestroy(d->ev_join);
  delete d;
  h->dsp = nullptr;
}

int popk_dsp_alloc(pop_handle* h, const pop_dsp_config* c) {
  popk_dsp_free(h);
  DspDev* d = new DspDev();
  h->dsp = d;
  d->c = *c;
  const pop_config& g = h->cfg;
  DspGeo& G = d->G;
  const int A = g.n_agents_local;
  d->nl = A;
  G.nx = g.nx; G.ny = g.ny; G.nz = g.nz; G.V = g.nx * g.ny * g.nz; G.T = g.nt;
  G.S = 2 * c->max_particle_num_voxel;
  G.MAXP = c->max_particle_num_voxel;
  G.nph = c->half_fov_h * 2 / c->angle_resolution;
  G.npv = c->half_fov_v * 2 / c->angle_resolution;
  G.NP = G.nph * G.npv;
  G.PCAP = c->pyramid_capacity;
  G.OCAP = c->max_points_per_pyramid;
  G.MAXPTS = c->max_points;
  G.MAXIN = c->max_input_points;
  int np2 = 1;
  while (np2 < G.MAXIN) np2 <<= 1;
  G.NPOW2 = np2;
  G.NB = c->new_born_particle_number_each_point;
  G.nrand = c->n_randoms;
  G.cl_maxn = c->dynamic_cluster_max_point_num;
  G.res = g.resolution;
  // setInitParameters :546-548
  G.hx = (G.res * static_cast<float>(G.nx)) * 0.5f;
  G.hy = (G.res * static_cast<float>(G.ny)) * 0.5f;
  G.hz = (G.res * static_cast<float>(G.nz)) * 0.5f;
  G.sigma = c->sigma_observation; G.kappa = c->kappa; G.pd = c->p_detection; G.nbw = c->new_born_particle_weight;
  G.filt = c->voxel_filter_resolution; G.thick = c->obstacle_thickness_for_occlusion;
  G.cl_maxh = c->dynamic_cluster_max_center_height; G.occ_thr = c->occupancy_threshold;
  // risk_mapping_node.cpp:34-36: LOCAL_RANGE = N / 2.f * VOXEL_RESOLUTION in double
  G.RX = g.nx / 2.f * static_cast<double>(g.resolution);
  G.RY = g.ny / 2.f * static_cast<double>(g.resolution);
  G.RZ = g.nz / 2.f * static_cast<double>(g.resolution);
  for (int k = 0; k < POP_DSP_MAX_T; k++) G.tf[k] = c->prediction_future_time[k];
  const size_t VS = static_cast<size_t>(G.V) * G.S;
  const size_t rcap = VS < (1u << 20) ? VS : (VS / 4 > (1u << 20) ? VS / 4 : (1u << 20));
  G.RCAP = static_cast<int>(rcap);
  const int nplanes = G.nph + 1 + G.npv + 1;

  DSP_ALLOC(d->flags, static_cast<size_t>(A) * G.V * kDspSlotStride);
  DSP_ALLOC(d->pa, A * VS);
  DSP_ALLOC(d->pb, A * VS);
  DSP_ALLOC(d->vobj, static_cast<size_t>(A) * G.V);
  DSP_ALLOC(d->pg, G.nrand); DSP_ALLOC(d->vg, G.nrand); DSP_ALLOC(d->ur, G.nrand);
  DSP_ALLOC(d->pdf, 20000);
  DSP_ALLOC(d->bp_ori, nplanes * 3);
  DSP_ALLOC(d->frame, A);
  DSP_ALLOC(d->bp, static_cast<size_t>(A) * nplanes * 3);
  DSP_ALLOC(d->cam, static_cast<size_t>(A) * G.MAXIN * 3);
  DSP_ALLOC(d->fkeys, static_cast<size_t>(A) * G.NPOW2);
  DSP_ALLOC(d->pts, static_cast<size_t>(A) * G.MAXPTS * 3);
  DSP_ALLOC(d->npts, A);
  DSP_ALLOC(d->rot, static_cast<size_t>(A) * G.MAXPTS * 3);
  DSP_ALLOC(d->pobs, static_cast<size_t>(A) * G.MAXPTS);
  DSP_ALLOC(d->obs, static_cast<size_t>(A) * G.NP * G.OCAP);
  DSP_ALLOC(d->obs_num, static_cast<size_t>(A) * G.NP);
  DSP_ALLOC(d->obs_maxlen, static_cast<size_t>(A) * G.NP);
  DSP_ALLOC(d->valid, A);
  DSP_ALLOC(d->cv, static_cast<size_t>(A) * G.MAXPTS * 7);
  DSP_ALLOC(d->ncv, A);
  DSP_ALLOC(d->cl_last, static_cast<size_t>(A) * kDspMaxDyn * 5);
  DSP_ALLOC(d->ncl_last, A);
  DSP_ALLOC(d->cost, static_cast<size_t>(A) * kDspMaxDyn * kDspMaxDyn);
  DSP_ALLOC(d->nrec, A);
  DSP_ALLOC(d->r_key, A * rcap); DSP_ALLOC(d->r_dest, A * rcap); DSP_ALLOC(d->r_next, A * rcap);
  DSP_ALLOC(d->r_pyr, A * rcap); DSP_ALLOC(d->r_slot, A * rcap); DSP_ALLOC(d->r_state, A * rcap);
  DSP_ALLOC(d->r_a, A * rcap); DSP_ALLOC(d->r_b, A * rcap);
  DSP_ALLOC(d->vhead, static_cast<size_t>(A) * G.V);
  DSP_ALLOC(d->pcount, static_cast<size_t>(A) * (G.NP + 1));
  DSP_ALLOC(d->poff, static_cast<size_t>(A) * (G.NP + 1));
  DSP_ALLOC(d->pcur, static_cast<size_t>(A) * (G.NP + 1));
  DSP_ALLOC(d->plist, A * rcap); DSP_ALLOC(d->plist2, A * rcap);
  DSP_ALLOC(d->changed, A * (kDspMaxIter + 2));
  DSP_ALLOC(d->pyr_n, static_cast<size_t>(A) * G.NP);
  DSP_ALLOC(d->pyr_ref, static_cast<size_t>(A) * G.NP * G.PCAP);
  DSP_ALLOC(d->pyr_p, static_cast<size_t>(A) * G.NP * G.PCAP);
  const size_t bcap = static_cast<size_t>(A) * G.MAXPTS * G.NB;
  DSP_ALLOC(d->b_dest, bcap); DSP_ALLOC(d->b_next, bcap); DSP_ALLOC(d->b_a, bcap); DSP_ALLOC(d->b_b, bcap);
  const size_t AV = static_cast<size_t>(A) * G.V;
  DSP_ALLOC(d->act[0], AV); DSP_ALLOC(d->act[1], AV); DSP_ALLOC(d->nact[0], A); DSP_ALLOC(d->nact[1], A);
  DSP_ALLOC(d->touch, AV); DSP_ALLOC(d->ntouch, A); DSP_ALLOC(d->touch2, AV); DSP_ALLOC(d->ntouch2, A);
  DSP_ALLOC(d->vstamp, AV);
  DSP_ALLOC(d->ok_dev, A);
  POP_CUDA(h, cudaStreamCreateWithFlags(&d->side, cudaStreamNonBlocking));
  POP_CUDA(h, cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
  POP_CUDA(h, cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming));
  DSP_ALLOC(d->seq, A * 4);
  DSP_ALLOC(d->stats, A * 8);
  DSP_ALLOC(d->err, 1);
  {
    const size_t n = static_cast<size_t>(A) * G.V;
    k_dsp_fill_int<<<static_cast<unsigned>((n + 255) / 256), 256, 0, h->stream>>>(d->vhead, n, -1);
    POP_LAUNCH_CHECK(h);
  }
  // calculateNormalPDFBuffer :1362-1367 (host libm, like the reference); plane normals :575-589
  std::vector<float> pdf(20000), bp(nplanes * 3, 0.f);
  for (int i = 0; i < 20000; ++i) {
    const float value = static_cast<float>(i - 10000) * 0.001f;
    pdf[i] = (1.f / (std::sqrt(2.f * 1.57079632679489661923f))) * std::exp(-std::pow(value, 2.f) / (2));
  }
  const float ar = static_cast<float>(c->angle_resolution) / 180.f * 3.14159265358979323846f;
  const int h_end = c->half_fov_h / c->angle_resolution, v_end = c->half_fov_v / c->angle_resolution;
  for (int i = -h_end; i <= h_end; i++) {
    bp[(i + h_end) * 3 + 0] = -std::sin(static_cast<float>(i) * ar);
    bp[(i + h_end) * 3 + 1] = std::cos(static_cast<float>(i) * ar);
  }
  float* bv = bp.data() + 3 * (G.nph + 1);
  for (int i = -v_end; i <= v_end; i++) {
    bv[(i + v_end) * 3 + 0] = std::sin(static_cast<float>(i) * ar);
    bv[(i + v_end) * 3 + 2] = std::cos(static_cast<float>(i) * ar);
  }
  POP_CUDA(h, cudaMemcpyAsync(d->pdf, pdf.data(), sizeof(float) * 20000, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d->bp_ori, bp.data(), sizeof(float) * bp.size(), cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  d->last_pos.assign(3 * A, 0.f);
  d->last_stamp.assign(A, 0.0);
  d->first.assign(A, 1);
  d->h_frame.resize(A);
  POP_CUDA(h, cudaFuncSetAttribute(k_dsp_velocity, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(vel_smem_bytes(G))));
  POP_CUDA(h, cudaFuncSetAttribute(k_dsp_filter, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(filter_smem_bytes(G))));
  POP_CUDA(h, cudaFuncSetAttribute(k_dsp_bin, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(bin_smem_bytes(G))));
  POP_CUDA(h, cudaFuncSetAttribute(k_dsp_born, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(sizeof(float) * G.MAXPTS)));
  return POP_OK;
}

int popk_dsp_filter_only(pop_handle* h) {
  DspDev* d = h->dsp;
  k_dsp_filter<<<d->nl, 1024, filter_smem_bytes(d->G), h->stream>>>(d->G, d->frame, d->cam, d->fkeys, d->pts, d->npts, d->err,
                                                                    d->use_cam_counts ? d->cam_counts : nullptr);
  POP_LAUNCH_CHECK(h);
  h->n_launches++;
  return POP_OK;
}

int popk_dsp_update(pop_handle* h, int filtered) {
  DspDev* d = h->dsp;
  const DspGeo& G = d->G;
  const int A = d->nl;
  cudaStream_t st = h->stream;
  int launches = 0;
  pop_prof_begin(h, POP_STAGE_DSP);
  if (!filtered) {
    k_dsp_filter<<<A, 1024, filter_smem_bytes(G), st>>>(G, d->frame, d->cam, d->fkeys, d->pts, d->npts, d->err,
                                                                d->use_cam_counts ? d->cam_counts : nullptr);
    POP_LAUNCH_CHECK(h); launches++;
  }
  k_dsp_bin<<<A, 1024, bin_smem_bytes(G), st>>>(G, d->frame, d->bp_ori, d->bp, d->pts, d->npts, d->rot, d->pobs,
                                                d->obs, d->obs_num, d->obs_maxlen, d->valid);
  POP_LAUNCH_CHECK(h); launches++;
  // velocityEstimationThread runs beside prediction + update, as the reference's side thread does (:305-320);
  // it shares only the rand() sequence counter with the newborn step, which waits for it
  POP_CUDA(h, cudaEventRecord(d->ev_fork, st));
  POP_CUDA(h, cudaStreamWaitEvent(d->side, d->ev_fork, 0));
  k_dsp_velocity<<<A, 1024, vel_smem_bytes(G), d->side>>>(G, d->frame, d->rot, d->npts, d->ur, d->seq, d->cv, d->ncv,
                                                          d->cl_last, d->ncl_last, d->cost, d->err);
  POP_LAUNCH_CHECK(h); launches++;
  POP_CUDA(h, cudaEventRecord(d->ev_join, d->side));
  // --- prediction ---
  POP_CUDA(h, cudaMemsetAsync(d->nrec, 0, sizeof(int) * A, st));
  POP_CUDA(h, cudaMemsetAsync(d->pcount, 0, sizeof(int) * static_cast<size_t>(A) * (G.NP + 1), st));
  POP_CUDA(h, cudaMemsetAsync(d->changed, 0, sizeof(int) * A * (kDspMaxIter + 2), st));
  POP_CUDA(h, cudaMemsetAsync(d->ntouch, 0, sizeof(int) * A, st));
  POP_CUDA(h, cudaMemsetAsync(d->ntouch2, 0, sizeof(int) * A, st));
  int* act = d->act[d->cur];
  int* nact = d->nact[d->cur];
  int* act_next = d->act[d->cur ^ 1];
  int* nact_next = d->nact[d->cur ^ 1];
  POP_CUDA(h, cudaMemsetAsync(nact_next, 0, sizeof(int) * A, st));
  // the per-voxel passes walk work lists (grid-stride): enough CTAs to fill the GPU, no more
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  const int per_agent = (sms * 8 + A - 1) / A;
  const int vblocks = (G.V + 255) / 256;
  const dim3 gv(per_agent < vblocks ? per_agent : vblocks, A);
  const int pblocks = (G.NP * 32 + 255) / 256;
  const dim3 gp(per_agent < pblocks ? per_agent : pblocks, A);
  k_dsp_pred_move<<<gv, 256, 0, st>>>(G, d->frame, d->bp, d->flags, d->pa, d->pb, d->nrec, d->r_key, d->r_dest,
                                      d->r_next, d->r_pyr, d->r_slot, d->r_state, d->r_a, d->r_b, d->vhead, d->pcount,
                                      d->stats, d->err, act, nact, d->touch, d->ntouch);
  POP_LAUNCH_CHECK(h); launches++;
  k_dsp_pyr_scan<<<A, 1024, 0, st>>>(G, d->frame, d->pcount, d->poff, d->pcur);
  POP_LAUNCH_CHECK(h); launches++;
  k_dsp_pyr_fill<<<gv, 256, 0, st>>>(G, d->frame, d->nrec, d->r_pyr, d->poff, d->pcur, d->plist);
  POP_LAUNCH_CHECK(h); launches++;
  k_dsp_pyr_sort<<<gp, 256, 0, st>>>(G, d->frame, d->r_key, d->poff, d->plist, d->plist2);
  POP_LAUNCH_CHECK(h); launches++;
  for (int round = 1; round <= kDspMaxIter; round++) {
    k_dsp_pred_voxel<<<gv, 256, 0, st>>>(G, d->frame, round, d->changed, d->flags, d->r_key, d->r_next, d->r_pyr,
                                         d->r_slot, d->r_state, d->vhead, d->touch, d->ntouch);
    POP_LAUNCH_CHECK(h); launches++;
    k_dsp_pred_pyr<<<gp, 256, 0, st>>>(G, d->frame, round, d->changed, d->poff, d->plist2, d->r_state);
    POP_LAUNCH_CHECK(h); launches++;
  }
  k_dsp_pyr_commit<<<gp, 256, 0, st>>>(G, d->frame, d->poff, d->plist2, d->r_state, d->r_dest, d->r_slot, d->r_a,
                                       d->r_b, d->pyr_n, d->pyr_ref, d->pyr_p);
  POP_LAUNCH_CHECK(h); launches++;
  k_dsp_pred_commit<<<gv, 256, 0, st>>>(G, d->frame, d->changed, d->flags, d->pa, d->pb, d->r_key, d->r_next,
                                        d->r_pyr, d->r_slot, d->r_state, d->r_a, d->r_b, d->vhead, d->stats, d->err,
                                        d->touch, d->ntouch, 1);
  POP_LAUNCH_CHECK(h); launches++;
  k_dsp_pred_commit<<<gv, 256, 0, st>>>(G, d->frame, d->changed, d->flags, d->pa, d->pb, d->r_key, d->r_next,
                                        d->r_pyr, d->r_slot, d->r_state, d->r_a, d->r_b, d->vhead, d->stats, d->err,
                                        act, nact, 0);
  POP_LAUNCH_CHECK(h); launches++;
  // --- update ---
  k_dsp_update_ck<<<dim3((G.MAXPTS + 127) / 128, A), 128, 0, st>>>(G, d->frame, d->npts, d->pobs, d->obs, d->valid,
                                                                   d->pyr_n, d->pyr_p, d->pdf);
  POP_LAUNCH_CHECK(h); launches++;
  k_dsp_update_w<<<dim3((G.NP * G.PCAP + 127) / 128, A), 128, 0, st>>>(G, d->frame, d->obs, d->obs_num,
                                                                       d->obs_maxlen, d->pyr_n, d->pyr_ref, d->pyr_p,
                                                                       d->pa, d->pdf);
  POP_LAUNCH_CHECK(h); launches++;
  // --- newborn ---
  POP_CUDA(h, cudaStreamWaitEvent(st, d->ev_join, 0));
  k_dsp_born<<<A, 1024, sizeof(float) * G.MAXPTS, st>>>(G, d->frame, d->cv, d->ncv, d->obs, d->obs_num, d->flags,
                                                        d->pa, d->pg, d->vg, d->ur, d->seq, d->b_dest, d->b_next,
                                                        d->b_a, d->b_b, d->vhead, d->err, d->touch2, d->ntouch2);
  POP_LAUNCH_CHECK(h); launches++;
  k_dsp_born_place<<<gv, 256, 0, st>>>(G, d->frame, d->flags, d->pa, d->pb, d->b_next, d->b_a, d->b_b, d->vhead,
                                       d->stats, d->touch2, d->ntouch2);
  POP_LAUNCH_CHECK(h); launches++;
  // --- occupancy, future deposits into the SOGM, resampling ---
  for (int a = 0; a < A;) {   // runs of accepted agents: one bulk-store launch per run (one launch in all if every update is accepted)
    if (!d->h_frame[a].ok) { a++; continue; }
    int b = a;
    while (b < A && d->h_frame[b].ok) b++;
    const int rc = popk_grid_clear(h, a, b - a);
    if (rc) return rc;
    launches++;
    a = b;
  }
  d->updates++;
  k_dsp_occupancy<<<gv, 256, 0, st>>>(G, d->frame, d->flags, d->pa, d->pb, d->vobj, h->d_grid, h->geom.agent_stride,
                                      act, nact, d->touch, d->ntouch, d->touch2, d->ntouch2, d->vstamp, d->updates,
                                      act_next, nact_next);
  POP_LAUNCH_CHECK(h); launches++;
  d->cur ^= 1;
  pop_prof_end(h, POP_STAGE_DSP, launches);
  return POP_OK;
}

int popk_dsp_download_particles(pop_handle* h, int agent, float* d_out) {
  DspDev* d = h->dsp;
  const size_t n = static_cast<size_t>(d->G.V) * d->G.S;
  k_dsp_download_particles<<<static_cast<unsigned>((n + 255) / 256), 256, 0, h->stream>>>(d->G, agent, d->flags, d->pa,
                                                                                          d->pb, d_out);
  POP_LAUNCH_CHECK(h);
  h->n_launches++;
  return POP_OK;
}

int popk_dsp_occupied(pop_handle* h, int agent, float* d_out, int cap, int* d_count) {
  DspDev* d = h->dsp;
  k_dsp_occupied<<<1, 1024, 0, h->stream>>>(d->G, agent, d->vobj, d_out, cap, d_count);
  POP_LAUNCH_CHECK(h);
  h->n_launches++;
  return POP_OK;
}
