// o_gridmap.cpp — oracle: the log-odds raycast map (TEST INFRASTRUCTURE).  Follows
//   plan_env/src/grid_map.cpp:15-131 (initMap), :192-208 (setCacheOccupancy), :313-456 (raycastProcess),
//   :458-475 (closetPointInMap), :477-583 (clearAndInflateLocalMap)
//   plan_env/src/raycast.cpp:17-31, 242-335 (RayCaster::setInput / step)
//   plan_env/include/plan_env/grid_map.h:261-420 (index helpers, occupancy getters)
// PINNED against the reference's own sources: the DDA (RayDda below) against raycast.cpp in oracle/_ref/libref.so
// (identical voxel sequences on 25 000 rays, tests/test_ref_pins.py) and everything else — projectDepthImage, the
// log-odds update, the shell reset, inflation, the virtual ceiling, both queries — against grid_map.cpp compiled
// unmodified into oracle/_ref/libref_gridmap.so (ROS, PCL, cv_bridge and Eigen replaced by stand-ins): projected
// points, occupancy_buffer_ doubles, inflated map and bounds bit for bit over hundreds of frames, live and through
// committed vectors (tests/test_ref_gridmap_pins.py).  Quirk kept: the frame counter raycast_num_ is a `char`
// (grid_map.h:140) like the flags it is compared with, so it wraps 127 -> -128 and a flag written 256 frames ago — or
// the initial -1 of a never-touched voxel, at frame 255 — reads as "already done in this frame".  (Round 1 restated
// the counter as an int, i.e. no de-duplication after frame 127; the pin against the compiled reference found it.)
#include <algorithm>
#include <cmath>
#include <queue>

#include "oracle.hpp"

namespace orc {

struct GridMapO {
  pop_gridmap_config c;
  double res, res_inv, origin[3], size[3], bmin[3], bmax[3];
  int n[3];
  double prob_hit, prob_miss, clamp_min, clamp_max, min_occ, unknown_flag;
  std::vector<double> occ;
  std::vector<char> inflate;
  std::vector<short> cnt_hm, cnt_hit;
  std::vector<signed char> flag_traverse, flag_rayend;
  signed char raycast_num = 0;   // `char raycast_num_` (grid_map.h:140), signed on the reference's platforms
  int lb_min[3] = {0, 0, 0}, lb_max[3] = {0, 0, 0};
  std::queue<int> cache;   // addresses (the reference queues index triples)

  static double logit(double x) { return std::log(x / (1 - x)); }
  void init(const pop_gridmap_config& cfg) {
    c = cfg;
    res = cfg.resolution;
    res_inv = 1 / res;
    origin[0] = -cfg.map_size[0] / 2.0; origin[1] = -cfg.map_size[1] / 2.0; origin[2] = cfg.ground_height;
    for (int i = 0; i < 3; i++) {
      size[i] = cfg.map_size[i];
      n[i] = static_cast<int>(std::ceil(size[i] / res));
      bmin[i] = origin[i];
      bmax[i] = origin[i] + size[i];
    }
    prob_hit = logit(cfg.p_hit); prob_miss = logit(cfg.p_miss);
    clamp_min = logit(cfg.p_min); clamp_max = logit(cfg.p_max); min_occ = logit(cfg.p_occ);
    unknown_flag = 0.01;
    const size_t N = static_cast<size_t>(n[0]) * n[1] * n[2];
    occ.assign(N, clamp_min - unknown_flag);
    inflate.assign(N, 0);
    cnt_hm.assign(N, 0); cnt_hit.assign(N, 0);
    flag_traverse.assign(N, -1); flag_rayend.assign(N, -1);
  }
  int addr(const int* id) const { return id[0] * n[1] * n[2] + id[1] * n[2] + id[2]; }
  void pos_to_index(const double* p, int* id) const {
    for (int i = 0; i < 3; i++) id[i] = static_cast<int>(std::floor((p[i] - origin[i]) * res_inv));
  }
  void bound(int* id) const { for (int i = 0; i < 3; i++) id[i] = std::max(std::min(id[i], n[i] - 1), 0); }
  bool in_map(const double* p) const {
    if (p[0] < bmin[0] + 1e-4 || p[1] < bmin[1] + 1e-4 || p[2] < bmin[2] + 1e-4) return false;
    if (p[0] > bmax[0] - 1e-4 || p[1] > bmax[1] - 1e-4 || p[2] > bmax[2] - 1e-4) return false;
    return true;
  }
  int set_cache(const double* pos, int occv) {   // setCacheOccupancy :192-208
    int id[3];
    pos_to_index(pos, id);
    const int a = addr(id);
    // the reference indexes with the linear address unchecked: an index that is out of range along one axis
    // (a ground point a hair below the map floor) aliases a neighbouring row; an address outside the buffers is
    // undefined behaviour there and is skipped here
    if (a < 0 || a >= n[0] * n[1] * n[2]) return -1;
    cnt_hm[a] += 1;
    if (cnt_hm[a] == 1) cache.push(a);
    if (occv == 1) cnt_hit[a] += 1;
    return a;
  }
  void closest_in_map(const double* pt, const double* cam, double* out) const {   // :458-475
    double diff[3], max_tc[3], min_tc[3];
    for (int i = 0; i < 3; i++) { diff[i] = pt[i] - cam[i]; max_tc[i] = bmax[i] - cam[i]; min_tc[i] = bmin[i] - cam[i]; }
    double min_t = 1000000;
    for (int i = 0; i < 3; ++i)
      if (std::fabs(diff[i]) > 0) {
        const double t1 = max_tc[i] / diff[i];
        if (t1 > 0 && t1 < min_t) min_t = t1;
        const double t2 = min_tc[i] / diff[i];
        if (t2 > 0 && t2 < min_t) min_t = t2;
      }
    for (int i = 0; i < 3; i++) out[i] = cam[i] + (min_t - 1e-3) * diff[i];
  }
  void raycast(const double* pts, int np, const double* cam);
  void clear_and_inflate();
};

static double rc_mod(double v, double m) { return std::fmod(std::fmod(v, m) + m, m); }
static double intbound(double s, double ds) {
  if (ds < 0) return intbound(-s, -ds);
  s = rc_mod(s, 1);
  return (1 - s) / ds;
}
static int signum(int x) { return x == 0 ? 0 : x < 0 ? -1 : 1; }

// RayCaster (plan_env/src/raycast.cpp:242-335): setInput floors both ends to voxel coordinates and prepares the
// per-axis crossing times; step hands out the current voxel, stops at the end voxel, else advances along the axis
// with the smallest crossing time (ties: y before x, z before both).
struct RayDda {
  int x[3], ex[3], st[3];
  double tmax[3], tdelta[3];
  bool set_input(const double* s, const double* e) {
    for (int k = 0; k < 3; k++) {
      x[k] = static_cast<int>(std::floor(s[k]));
      ex[k] = static_cast<int>(std::floor(e[k]));
      const double d = ex[k] - x[k];
      st[k] = signum(static_cast<int>(d));
      tmax[k] = intbound(s[k], d);
      tdelta[k] = static_cast<double>(st[k]) / d;
    }
    return !(st[0] == 0 && st[1] == 0 && st[2] == 0);
  }
  bool step(double* ray) {
    for (int k = 0; k < 3; k++) ray[k] = static_cast<double>(x[k]);
    if (x[0] == ex[0] && x[1] == ex[1] && x[2] == ex[2]) return false;
    if (tmax[0] < tmax[1]) {
      if (tmax[0] < tmax[2]) { x[0] += st[0]; tmax[0] += tdelta[0]; } else { x[2] += st[2]; tmax[2] += tdelta[2]; }
    } else {
      if (tmax[1] < tmax[2]) { x[1] += st[1]; tmax[1] += tdelta[1]; } else { x[2] += st[2]; tmax[2] += tdelta[2]; }
    }
    return true;
  }
};

void GridMapO::raycast(const double* pts, int np, const double* cam) {
  if (np == 0) return;
  raycast_num = static_cast<signed char>(raycast_num + 1);   // wraps 127 -> -128
  double mn[3] = {bmax[0], bmax[1], bmax[2]}, mx[3] = {bmin[0], bmin[1], bmin[2]};
  for (int i = 0; i < np; ++i) {
    double pw[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    int vox;
    auto clip = [&](double* p) {
      double d[3] = {p[0] - cam[0], p[1] - cam[1], p[2] - cam[2]};
      const double len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      if (len > c.max_ray_length) {
        for (int k = 0; k < 3; k++) p[k] = d[k] / len * c.max_ray_length + cam[k];
        return true;
      }
      return false;
    };
    if (!in_map(pw)) {
      double q[3];
      closest_in_map(pw, cam, q);
      for (int k = 0; k < 3; k++) pw[k] = q[k];
      clip(pw);
      vox = set_cache(pw, 0);
    } else {
      vox = clip(pw) ? set_cache(pw, 0) : set_cache(pw, 1);
    }
    for (int k = 0; k < 3; k++) { mx[k] = std::max(mx[k], pw[k]); mn[k] = std::min(mn[k], pw[k]); }
    if (vox >= 0) {
      if (flag_rayend[vox] == raycast_num) continue;
      flag_rayend[vox] = raycast_num;
    }
    // RayCaster::setInput(pt_w / res, camera / res) + step loop (:379-395)
    double s[3], e[3], ray[3];
    for (int k = 0; k < 3; k++) { s[k] = pw[k] / res; e[k] = cam[k] / res; }
    RayDda dda;
    dda.set_input(s, e);
    while (dda.step(ray)) {
      const double tmp[3] = {(ray[0] + 0.5) * res, (ray[1] + 0.5) * res, (ray[2] + 0.5) * res};
      const int v = set_cache(tmp, 0);
      if (v < 0) continue;
      if (flag_traverse[v] == raycast_num) break;
      flag_traverse[v] = raycast_num;
    }
  }
  for (int k = 0; k < 3; k++) { mn[k] = std::min(mn[k], cam[k]); mx[k] = std::max(mx[k], cam[k]); }
  mx[2] = std::max(mx[2], c.ground_height);
  pos_to_index(mx, lb_max);
  pos_to_index(mn, lb_min);
  bound(lb_min);
  bound(lb_max);
  double lmin[3], lmax[3];
  for (int k = 0; k < 3; k++) { lmin[k] = cam[k] - c.local_update_range[k]; lmax[k] = cam[k] + c.local_update_range[k]; }
  int min_id[3], max_id[3];
  pos_to_index(lmin, min_id); pos_to_index(lmax, max_id);
  bound(min_id); bound(max_id);
  while (!cache.empty()) {
    const int a = cache.front();
    cache.pop();
    const int ix = a / (n[1] * n[2]), iy = (a / n[2]) % n[1], iz = a % n[2];
    const double upd = cnt_hit[a] >= cnt_hm[a] - cnt_hit[a] ? prob_hit : prob_miss;
    cnt_hit[a] = cnt_hm[a] = 0;
    if (upd >= 0 && occ[a] >= clamp_max) continue;
    else if (upd <= 0 && occ[a] <= clamp_min) { occ[a] = clamp_min; continue; }
    const bool in_local = ix >= min_id[0] && ix <= max_id[0] && iy >= min_id[1] && iy <= max_id[1] && iz >= min_id[2] && iz <= max_id[2];
    if (!in_local) occ[a] = clamp_min;
    occ[a] = std::min(std::max(occ[a] + upd, clamp_min), clamp_max);
  }
}

void GridMapO::clear_and_inflate() {   // :477-583
  const int vec_margin = 5;
  int min_cut[3], max_cut[3], min_cut_m[3], max_cut_m[3];
  for (int k = 0; k < 3; k++) { min_cut[k] = lb_min[k] - c.local_map_margin; max_cut[k] = lb_max[k] + c.local_map_margin; }
  bound(min_cut); bound(max_cut);
  for (int k = 0; k < 3; k++) { min_cut_m[k] = min_cut[k] - vec_margin; max_cut_m[k] = max_cut[k] + vec_margin; }
  bound(min_cut_m); bound(max_cut_m);
  const double unk = clamp_min - unknown_flag;
  auto A = [&](int x, int y, int z) { return x * n[1] * n[2] + y * n[2] + z; };
  for (int x = min_cut_m[0]; x <= max_cut_m[0]; ++x)
    for (int y = min_cut_m[1]; y <= max_cut_m[1]; ++y) {
      for (int z = min_cut_m[2]; z < min_cut[2]; ++z) occ[A(x, y, z)] = unk;
      for (int z = max_cut[2] + 1; z <= max_cut_m[2]; ++z) occ[A(x, y, z)] = unk;
    }
  for (int z = min_cut_m[2]; z <= max_cut_m[2]; ++z)
    for (int x = min_cut_m[0]; x <= max_cut_m[0]; ++x) {
      for (int y = min_cut_m[1]; y < min_cut[1]; ++y) occ[A(x, y, z)] = unk;
      for (int y = max_cut[1] + 1; y <= max_cut_m[1]; ++y) occ[A(x, y, z)] = unk;
    }
  for (int y = min_cut_m[1]; y <= max_cut_m[1]; ++y)
    for (int z = min_cut_m[2]; z <= max_cut_m[2]; ++z) {
      for (int x = min_cut_m[0]; x < min_cut[0]; ++x) occ[A(x, y, z)] = unk;
      for (int x = max_cut[0] + 1; x <= max_cut_m[0]; ++x) occ[A(x, y, z)] = unk;
    }
  const int inf_step = static_cast<int>(std::ceil(c.obstacles_inflation / res));
  const int N = n[0] * n[1] * n[2];
  for (int x = lb_min[0]; x <= lb_max[0]; ++x)
    for (int y = lb_min[1]; y <= lb_max[1]; ++y)
      for (int z = lb_min[2]; z <= lb_max[2]; ++z) inflate[A(x, y, z)] = 0;
  for (int x = lb_min[0]; x <= lb_max[0]; ++x)
    for (int y = lb_min[1]; y <= lb_max[1]; ++y)
      for (int z = lb_min[2]; z <= lb_max[2]; ++z)
        if (occ[A(x, y, z)] > min_occ)
          for (int dx = -inf_step; dx <= inf_step; ++dx)
            for (int dy = -inf_step; dy <= inf_step; ++dy)
              for (int dz = -inf_step; dz <= inf_step; ++dz) {
                const int a = A(x + dx, y + dy, z + dz);   // the reference only checks the linear address
                if (a < 0 || a >= N) continue;
                inflate[a] = 1;
              }
  if (c.virtual_ceil_height > -0.5) {
    const int ceil_id = static_cast<int>(std::floor((c.virtual_ceil_height - origin[2]) * res_inv)) - 1;
    for (int x = lb_min[0]; x <= lb_max[0]; ++x)
      for (int y = lb_min[1]; y <= lb_max[1]; ++y) {
        const int a = A(x, y, ceil_id);
        if (a >= 0 && a < N) inflate[a] = 1;   // (the reference writes unchecked)
      }
  }
}

// projectDepthImage :210-311.  Quirks kept: without the filter the row pointer advances by ONE per sampled pixel
// whatever skip_pixel is; with the filter the zero test reads the pixel AFTER the pointer advanced by skip_pixel
// (the next sample; past the end of the image it is undefined in the reference and read as non-zero here), and the
// very first frame only arms has_first_depth_.
int project_depth(const pop_depth_params& P, bool& has_first, const uint16_t* img, const double* R, const double* cam,
                  double max_ray, double* out, int cap) {
  int n = 0;
  auto emit = [&](int u, int v, double depth) {
    const double x = (u - P.cx) * depth / P.fx, y = (v - P.cy) * depth / P.fy, z = depth;
    if (n < cap)
      for (int i = 0; i < 3; i++) out[3 * n + i] = ((R[3 * i] * x + R[3 * i + 1] * y) + R[3 * i + 2] * z) + cam[i];
    n++;
  };
  const long long total = static_cast<long long>(P.rows) * P.cols;
  if (!P.use_depth_filter) {
    for (int v = 0; v < P.rows; v += P.skip_pixel) {
      long long ptr = static_cast<long long>(v) * P.cols;
      for (int u = 0; u < P.cols; u += P.skip_pixel) {
        const double depth = img[ptr++] / P.k_depth_scaling_factor;
        emit(u, v, depth);
      }
    }
  } else {
    if (!has_first) { has_first = true; return 0; }
    const double inv_factor = 1.0 / P.k_depth_scaling_factor;
    for (int v = P.depth_filter_margin; v < P.rows - P.depth_filter_margin; v += P.skip_pixel) {
      long long ptr = static_cast<long long>(v) * P.cols + P.depth_filter_margin;
      for (int u = P.depth_filter_margin; u < P.cols - P.depth_filter_margin; u += P.skip_pixel) {
        double depth = img[ptr] * inv_factor;
        ptr += P.skip_pixel;
        const unsigned next = ptr < total ? img[ptr] : 1u;
        if (next == 0) depth = max_ray + 0.1;
        else if (depth < P.depth_filter_mindist) continue;
        else if (depth > P.depth_filter_maxdist) depth = max_ray + 0.1;
        emit(u, v, depth);
      }
    }
  }
  return n;
}

}  // namespace orc

using namespace orc;
extern "C" {
// the DDA alone, driven as raycastProcess drives it: every voxel for which step() returned true
int orc_raycaster(const double* start, const double* end, double* out, int cap, int* set_ok) {
  RayDda dda;
  const bool ok = dda.set_input(start, end);
  if (set_ok) *set_ok = ok ? 1 : 0;
  double ray[3];
  int n = 0;
  while (dda.step(ray)) {
    if (n < cap) for (int k = 0; k < 3; k++) out[3 * n + k] = ray[k];
    if (++n > (1 << 22)) break;
  }
  return n;
}
void* orc_gridmap_create(const pop_gridmap_config* c) { GridMapO* m = new GridMapO(); m->init(*c); return m; }
void orc_gridmap_destroy(void* h) { delete static_cast<GridMapO*>(h); }
void orc_gridmap_dims(void* h, int32_t* out) { GridMapO* m = static_cast<GridMapO*>(h); out[0] = m->n[0]; out[1] = m->n[1]; out[2] = m->n[2]; }
// updateOccupancyCallback body after projectDepthImage: raycastProcess + clearAndInflateLocalMap
void orc_gridmap_update(void* h, const double* pts, int n, const double* cam) {
  GridMapO* m = static_cast<GridMapO*>(h);
  m->raycast(pts, n, cam);
  if (n > 0) m->clear_and_inflate();
}
// projectDepthImage alone; has_first is the map's has_first_depth_ flag (in/out)
int orc_gridmap_project(const pop_depth_params* P, int* has_first, const uint16_t* img, const double* R, const double* cam,
                        double max_ray, double* out, int cap) {
  bool hf = *has_first != 0;
  const int n = project_depth(*P, hf, img, R, cam, max_ray, out, cap);
  *has_first = hf ? 1 : 0;
  return n;
}
void orc_gridmap_download(void* h, double* occ, int8_t* inflate, int32_t* bounds) {
  GridMapO* m = static_cast<GridMapO*>(h);
  std::copy(m->occ.begin(), m->occ.end(), occ);
  for (size_t i = 0; i < m->inflate.size(); i++) inflate[i] = m->inflate[i];
  for (int k = 0; k < 3; k++) { bounds[k] = m->lb_min[k]; bounds[3 + k] = m->lb_max[k]; }
}
}
