// o_map.cpp — oracle: SOGM grid, ground-truth update, occupancy queries (TEST INFRASTRUCTURE).
// Follows plan_env/src/fake_particle_risk_voxel.cpp, plan_env/src/risk_base.cpp,
// plan_env/src/map.cpp and plan_env/include/plan_env/map.h of the reference.
// PINNED: plan_env/src/map.cpp, fake_particle_risk_voxel.cpp and risk_base.cpp are compiled unmodified into
// oracle/_ref/libref_map_*.so (oracle/Makefile, ref_glue_map.cpp; stand-ins for ROS, PCL and Eigen's fp32 geometry) and
// tests/test_ref_map_pins.py compares every function of this file with them bit for bit (live and through
// tests/golden/ref_map_pins.npz).
#include <algorithm>

#include "oracle.hpp"

namespace orc {

// fake_particle_risk_voxel.cpp:27-46 (Fake) / risk_base.cpp:25-39 (RiskBase)
void Grid::init(const pop_config& c) {
  nx = c.nx; ny = c.ny; nz = c.nz; nt = c.nt;
  res = c.resolution;
  dt = c.time_resolution;
  rx = nx / 2 * res;  // integer half, then fp32 multiply
  ry = ny / 2 * res;
  rz = nz / 2 * res;
  risk_threshold = c.risk_threshold;
  clearance = c.clearance;
  ground = c.ground_height;
  ceiling = c.ceiling_height;
  thr_region = c.risk_threshold_region;
  decay_region = c.risk_region_decay;
  decay_voxel = c.risk_voxel_decay;
  query_variant = c.query_variant;
  kernel_shape = c.kernel_shape;
  inf_step = clearance / res;  // fp32 divide, truncation (0.45f/0.15f -> 2)
  kernel.clear();
  for (int x = -inf_step; x <= inf_step; x++)
    for (int y = -inf_step; y <= inf_step; y++) {
      if (kernel_shape == POP_KERNEL_DISC_Z0) {
        // `for (int z = -0.3; z <= 0.3; z++)` collapses to z = 0
        kernel.push_back(V3i{x, y, 0});
      } else {
        for (int z = -inf_step; z <= inf_step; z++) kernel.push_back(V3i{x, y, z});
      }
    }
  pose = V3f{0.f, 0.f, 0.f};
  t_map = 0.0;
  risk.assign(static_cast<size_t>(V()) * nt, 0.f);
}

// fp32 3-vector helpers in the evaluation order of Eigen 3.3's scalar code path, which is what the reference's
// Vector3f / Quaternionf / Hyperplane<float, 3> expressions compile to: reductions of a fixed-size 3-vector pair
// c0 + (c1 + c2) (Core/Redux.h, redux_novec_unroller), cross is the plain formula (Geometry/OrthoMethods.h),
// q * v = v + w * (2 vec x v) + vec x (2 vec x v) (Geometry/Quaternion.h, _transformVector).
static inline V3f add3(V3f a, V3f b) { return V3f{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3f sub3(V3f a, V3f b) { return V3f{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline float dot3(V3f a, V3f b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
static inline float norm3(V3f a) { return std::sqrt(dot3(a, a)); }
static inline V3f cross3(V3f a, V3f b) { return V3f{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline V3f quat_rotate(float w, V3f q, V3f v) {
  V3f uv = cross3(q, v);
  uv = add3(uv, uv);
  const V3f c = cross3(q, uv);
  return V3f{(v.x + w * uv.x) + c.x, (v.y + w * uv.y) + c.y, (v.z + w * uv.z) + c.z};
}

// fake_particle_risk_voxel.cpp:80-170
void gt_update(Grid& g, const float* cloud, int n, const pop_cylinder* cyl, int ncyl,
               double t_map) {
  g.t_map = t_map;
  const int V = g.V(), T = g.nt;
  // PassThrough x, y, z (:88-104): keep finite points with field in [pose-range, pose+range]
  const float xl = g.pose.x - g.rx, xh = g.pose.x + g.rx;
  const float yl = g.pose.y - g.ry, yh = g.pose.y + g.ry;
  const float zl = g.pose.z - g.rz, zh = g.pose.z + g.rz;

  std::vector<float> tmp(static_cast<size_t>(V) * T, 0.0F);  // :106-107
  for (int i = 0; i < n; i++) {
    const float px = cloud[3 * i], py = cloud[3 * i + 1], pz = cloud[3 * i + 2];
    if (!std::isfinite(px) || !std::isfinite(py) || !std::isfinite(pz)) continue;
    if (px < xl || px > xh) continue;
    if (py < yl || py > yh) continue;
    if (pz < zl || pz > zh) continue;
    V3f pt{px - g.pose.x, py - g.pose.y, pz - g.pose.z};  // :111
    if (g.in_range(pt)) {
      // fp32 rounding of (p + range) / res can reach nx (ny, nz) for points a hair inside the
      // border; the reference then writes the linear index as is (bleeding into the next
      // row/layer, out of bounds past the last layer — UB). Kept, minus the out-of-bounds write.
      const int idx = g.index(pt);
      if (idx >= 0 && idx < V) tmp[static_cast<size_t>(idx) * T + 0] = 1.0F;
    }
  }
  std::vector<int> obs;  // :118-125
  for (int i = 0; i < V; i++)
    if (tmp[static_cast<size_t>(i) * T] > g.risk_threshold) obs.push_back(i);

  for (int i : obs) {  // :127-162
    V3f pt = g.voxel_position(i);
    V3f vel{0, 0, 0};
    for (int c = 0; c < ncyl; c++) {
      if (cyl[c].type == 3) {
        V3f pc{static_cast<float>(cyl[c].x), static_cast<float>(cyl[c].y), pt.z};
        float dist = norm3(V3f{pt.x - pc.x, pt.y - pc.y, pt.z - pc.z});
        if (dist <= cyl[c].w + g.clearance) {  // evaluated in double
          vel = V3f{static_cast<float>(cyl[c].vx), static_cast<float>(cyl[c].vy), 0.0F};
          break;
        }
      } else if (cyl[c].type == 2) {  // :139-150 vertical ring (MovingCircle), fp32 Eigen geometry
        const V3f pc{static_cast<float>(cyl[c].x), static_cast<float>(cyl[c].y), static_cast<float>(cyl[c].z)};
        const float qw = static_cast<float>(cyl[c].qw);
        const V3f qv{static_cast<float>(cyl[c].qx), static_cast<float>(cyl[c].qy), static_cast<float>(cyl[c].qz)};
        // Hyperplane::Through(p0, p0 + q*(0,1,0), p0 + q*(1,0,0)): normal = (p2 - p0) x (p1 - p0), normalised
        const V3f p1 = add3(pc, quat_rotate(qw, qv, V3f{0, 1, 0}));
        const V3f p2 = add3(pc, quat_rotate(qw, qv, V3f{1, 0, 0}));
        V3f nrm = cross3(sub3(p2, pc), sub3(p1, pc));
        const float nn = norm3(nrm);
        nrm = V3f{nrm.x / nn, nrm.y / nn, nrm.z / nn};
        const float off = -dot3(pc, nrm);
        const float sd = dot3(nrm, pt) + off;                                  // signedDistance(pt)
        const V3f b{pt.x - sd * nrm.x, pt.y - sd * nrm.y, pt.z - sd * nrm.z};  // projection(pt)
        const float dist_to_plane = std::fabs(sd);
        const float dist = norm3(sub3(pc, b));
        // `abs(cyl.w / 2 - dist) < 2 * resolution_`: the left side is a double, 2 * res an fp32 product
        if (std::fabs(cyl[c].w / 2 - dist) < 2 * g.res && dist_to_plane < 2 * g.res) {
          vel = V3f{static_cast<float>(cyl[c].vx), static_cast<float>(cyl[c].vy), 0.0F};
          break;
        }
      }
      // other types only make the reference print "unknown type"
    }
    for (int k = 1; k < T; k++) {
      const float kf = static_cast<float>(k);
      V3f pp{(pt.x + vel.x * g.dt * kf) - g.pose.x, (pt.y + vel.y * g.dt * kf) - g.pose.y,
             (pt.z + vel.z * g.dt * kf) - g.pose.z};
      if (g.in_range(pp)) {
        const int idx = g.index(pp);
        if (idx >= 0 && idx < V) tmp[static_cast<size_t>(idx) * T + k] = 1.0F;
      }
    }
  }
  g.risk.swap(tmp);  // :164-170
}

// fake_particle_risk_voxel.cpp:291-301
void add_particles(Grid& g, const V3d* pts, const float* risks, int n, int t_index) {
  for (int i = 0; i < n; i++) {
    V3f p{static_cast<float>(pts[i].x), static_cast<float>(pts[i].y), static_cast<float>(pts[i].z)};
    if (!g.in_range(p)) continue;
    const int idx = g.index(p);
    if (idx < 0 || idx >= g.V()) continue;  // see gt_update
    g.at(idx, t_index) += risks[i];
  }
}

// fake_particle_risk_voxel.cpp:309-330 / risk_base.cpp:228-250
int query(const Grid& g, V3d pos, int t) {
  if (g.query_variant == POP_QUERY_FAKE) {
    if (pos.z < g.ground || pos.z > g.ceiling) return -1;
  } else {
    if (pos.z < g.ground) return 1;
    if (pos.z > g.ceiling) return 1;
  }
  V3f pf{static_cast<float>(pos.x) - g.pose.x, static_cast<float>(pos.y) - g.pose.y,
         static_cast<float>(pos.z) - g.pose.z};
  V3i pi = g.rel_index(pf);
  if (!g.in_range(pi)) return -1;
  if (t < 0 || t >= g.nt) return -1;  // the reference would index risk_maps_ out of bounds
  float sum = 0.0;
  const float thr = g.query_variant == POP_QUERY_FAKE ? g.risk_threshold
                                                      : g.thr_region - t * g.decay_region;
  for (const V3i& k : g.kernel) {
    V3i p{pi.x + k.x, pi.y + k.y, pi.z + k.z};
    if (!g.in_range(p)) continue;
    sum += g.at(g.index(p), t);
    if (sum > thr) return 1;
  }
  return 0;
}

// fake_particle_risk_voxel.cpp:343-348 / risk_base.cpp:256-260
int query_dt(const Grid& g, V3d pos, double dt) {
  int tf = std::floor(dt / g.dt);
  tf = tf > (g.nt - 1) ? g.nt - 1 : tf;
  return query(g, pos, tf);
}

// MapBase::getInflateOccupancy (map.cpp:345-395): the single-voxel test of the kept map interface (map.h:120-122).
// mode 0 (pos) = slice 0, mode 1 (pos, int t), mode 2 (pos, double t) = slices ceil(t / dt) and floor(t / dt).
// Deviations: the reference lets t == PREDICTION_TIMES through (`t > PREDICTION_TIMES` / clamps to PREDICTION_TIMES)
// and then reads one slice past the end (the aliasing of SURVEY Appendix A Q2); here such a slice index answers -1
// (mode 1) or is clamped to nt - 1 (mode 2).  Negative indices answer -1.
int inflate_query(const Grid& g, V3d pos, int mode, double t) {
  const V3f pf{static_cast<float>(pos.x) - g.pose.x, static_cast<float>(pos.y) - g.pose.y,
               static_cast<float>(pos.z) - g.pose.z};
  const int idx = g.index(pf);
  if (!g.in_range(pf)) return -1;
  if (idx < 0 || idx >= g.V()) return -1;  // see gt_update: rounding at the border
  if (mode == 0) return g.at(idx, 0) > g.risk_threshold ? 1 : 0;
  if (mode == 1) {
    const int ti = static_cast<int>(t);
    if (ti < 0 || ti >= g.nt) return -1;
    return g.at(idx, ti) > g.risk_threshold ? 1 : 0;
  }
  int tc = std::ceil(t / g.dt);
  tc = tc > g.nt - 1 ? g.nt - 1 : tc;
  int tf = std::floor(t / g.dt);
  tf = tf > g.nt - 1 ? g.nt - 1 : tf;
  if (tc < 0 || tf < 0) return -1;
  if (g.at(idx, tc) > g.risk_threshold) return 1;
  if (g.at(idx, tf) > g.risk_threshold) return 1;
  return 0;
}

// map.cpp:463-531 (Fake inherits MapBase's) / risk_base.cpp:295-337.
// Deviation: the reference clamps slice indices to PREDICTION_TIMES and reads
// one slice past the end; here they are clamped to nt-1.
void obstacle_points(const Grid& g, double t_start, double t_end, V3d lc, V3d hc,
                     std::vector<V3d>& out) {
  const double t0 = g.t_map;
  int idx_start = std::floor((t_start - t0) / g.dt);
  int idx_end = std::ceil((t_end - t0) / g.dt);
  idx_start = idx_start < 0 ? 0 : idx_start;
  idx_start = idx_start > g.nt - 1 ? g.nt - 1 : idx_start;
  idx_end = idx_end > g.nt - 1 ? g.nt - 1 : idx_end;
  idx_end = idx_end < 0 ? 0 : idx_end;

  int lx = (lc.x - g.pose.x + g.rx) / g.res;  // double arithmetic, truncation
  int ly = (lc.y - g.pose.y + g.ry) / g.res;
  int lz = (lc.z - g.pose.z + g.rz) / g.res;
  int hx = (hc.x - g.pose.x + g.rx) / g.res;
  int hy = (hc.y - g.pose.y + g.ry) / g.res;
  int hz = (hc.z - g.pose.z + g.rz) / g.res;
  hx = std::min(hx, g.nx - 1); hy = std::min(hy, g.ny - 1); hz = std::min(hz, g.nz - 1);
  lx = std::max(lx, 0); ly = std::max(ly, 0); lz = std::max(lz, 0);

  for (int z = lz; z <= hz; z++)
    for (int y = ly; y <= hy; y++)
      for (int x = lx; x <= hx; x++) {
        int i = x + y * g.nx + z * g.nx * g.ny;
        for (int j = idx_start; j <= idx_end; j++) {
          const float thr = g.query_variant == POP_QUERY_FAKE
                                ? g.risk_threshold
                                : g.risk_threshold - g.decay_voxel * j;
          if (g.at(i, j) > thr) {
            V3f p = g.voxel_position(i);
            out.push_back(V3d{p.x, p.y, p.z});
          }
        }
      }
}

}  // namespace orc
