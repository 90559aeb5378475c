// o_dsp.cpp — oracle: the DSP particle map (SMC-PHD dual-structure particle map) that feeds the SOGM on
// the real-perception path (TEST INFRASTRUCTURE).  Follows, line by line and in the reference's own
// sequential order:
//   plan_env/include/plan_env/dsp_dynamic.h   (DSPMap: update, mapPrediction, mapUpdate,
//                                              mapAddNewBornParticlesByObservation,
//                                              mapOccupancyCalculationAndResample, velocityEstimationThread)
//   plan_env/include/plan_env/map_parameters.h
//   plan_env/src/risk_mapping_node.cpp:61-188 (filterPointCloud, cloudCallback, publishMap)
//   plan_env/src/risk_base.cpp:60-80          (futureRiskCallback: copy into risk_maps_)
//
// PARITY UNPINNED: the reference holds no test for this map, cannot be compiled here (ROS, Eigen, PCL,
// Munkres absent) and is itself non-deterministic (noise tables and rand() seeded with time(NULL),
// dsp_dynamic.h:628,1220).  Determinisation chosen here (and mirrored by the CUDA path):
//   * the Gaussian tables p_gaussian_randoms / v_gaussian_randoms and the rand() stream are INPUTS
//     (tables of configurable length, consumed through the reference's wrapping sequence counters);
//   * Eigen::Quaternionf products are evaluated as Eigen 3.3's SSE kernel pairs them (see DspMap::rotate);
//   * PCL VoxelGrid = per-leaf centroid, leaves in ascending linear leaf index, points summed in input order;
//   * PCL EuclideanClusterExtraction = connected components of the radius graph (tolerance 2*filter res),
//     members in ascending index, clusters of >= 5 points ordered by size descending, ties by smallest member;
//   * Munkres = minimum-cost assignment (Jonker-Volgenant style shortest augmenting paths); ties are broken
//     by lowest column index;
//   * column 8 of voxels_with_particle (update time stamp) is write-only in the reference and not kept.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>
#include <map>
#include <numeric>

#include "oracle.hpp"

namespace orc {

struct VelPoint { float x, y, z, intensity, nx, ny, nz; };  // pcl::PointXYZINormal as used
struct Cluster {
  float cx = 0.f, cy = 0.f, cz = 0.f;
  int n = 0, match = -1;
  float vx = -10000.f, vy = -10000.f, vz = -10000.f, v = 0.f, intensity = 0.f;
};

struct DspMap {
  pop_dsp_config c;
  int nx, ny, nz, V, T, S;        // S = SAFE_PARTICLE_NUM_VOXEL
  int nph, npv, NP, PCAP, OCAP;   // pyramids: h x v, particle capacity, observation capacity
  float res, hx, hy, hz;          // map_length_{x,y,z}_half
  std::vector<float> part;        // voxels_with_particle[V][S][8] (flag vx vy vz px py pz w)
  std::vector<float> vobj;        // voxels_objects_number[V][4+T]
  std::vector<int> pyr;           // pyramids_in_fov[NP][PCAP][3]
  std::vector<int> nbr;           // observation_pyramid_neighbors[NP][10]
  std::vector<float> pc;          // point_cloud[NP][OCAP][5]
  std::vector<int> obs_num;
  std::vector<float> pc_maxlen;
  std::vector<float> bp_ori_h, bp_ori_v, bp_h, bp_v;
  float quat[4];
  std::vector<float> pg, vg;      // Gaussian tables
  std::vector<int32_t> ur;        // rand() stream
  int pseq = 0, vseq = 0, useq = 0;
  std::vector<float> pdf;
  float sigma_update, kappa, P_detection, update_time = 0.f;
  int update_counter = 0;
  float expected_new_born_objects = 0.f, new_born_particle_weight, new_born_each_object_weight = 0.f;
  int nb_each;
  float cur_pos[3] = {0, 0, 0};
  float delt_t_last = 0.f;
  std::vector<V3f> cloud_rot;
  std::vector<VelPoint> cloud_vel;
  std::vector<Cluster> clusters_last;
  bool first = true;
  float last_pos[3];
  double last_stamp;
  int stat_vox_full = 0, stat_pyr_full = 0, stat_out = 0, stat_born = 0;

  float* P(int v, int p) { return &part[(static_cast<size_t>(v) * S + p) * 8]; }

  void init(const pop_config& g, const pop_dsp_config& d);
  int voxel_index(float px, float py, float pz, int& index) const;
  int rand_next() { int r = ur[useq]; if (++useq >= (int)ur.size()) useq = 0; return r; }
  float rand_float(float mn, float mx) {  // generateRandomFloat :1687-1689
    return mn + static_cast<float>(rand_next()) / (static_cast<float>(2147483647 / (mx - mn)));
  }
  float pos_gauss() { float v = pg[pseq]; if (++pseq >= (int)pg.size()) pseq = 0; return v; }
  float vel_gauss() { float v = vg[vseq]; if (++vseq >= (int)vg.size()) vseq = 0; return v; }
  float query_pdf(float x, float mu, float sigma) const {  // :1369-1377
    float cx = (x - mu) / sigma;
    if (cx > 9.9f) cx = 9.9f; else if (cx < -9.9f) cx = -9.9f;
    return pdf[static_cast<int>(cx * 1000 + 10000)];
  }
  static void rotate(const float* v, const float* q, float* out);
  bool in_pyramids(float x, float y, float z) const;
  int pyr_h(float x, float y, float z) const;
  int pyr_v(float x, float y, float z) const;
  int add_particle(float px, float py, float pz, float vx, float vy, float vz, float w, int vidx);
  int move_particle(int new_v, int cur_v, int cur_p, float* flag_ptr);
  int update(int n, const float* pts, const float* pos, double stamp, const float* q);
  void prediction(float dx, float dy, float dz, float dt);
  void map_update();
  void velocity_estimation();
  void add_newborn();
  void occupancy_resample();
};

static inline float dot3(float x1, float y1, float z1, float x2, float y2, float z2) {
  return x1 * x2 + y1 * y2 + z1 * z2;  // vectorMultiply :1407-1409
}

void DspMap::init(const pop_config& g, const pop_dsp_config& d) {
  c = d;
  nx = g.nx; ny = g.ny; nz = g.nz; T = g.nt; V = nx * ny * nz;
  res = g.resolution;
  S = 2 * d.max_particle_num_voxel;
  nph = d.half_fov_h * 2 / d.angle_resolution;
  npv = d.half_fov_v * 2 / d.angle_resolution;
  NP = nph * npv;
  PCAP = d.pyramid_capacity;
  OCAP = d.max_points_per_pyramid;
  // setInitParameters :542-597
  hx = (res * static_cast<float>(nx)) * 0.5f;
  hy = (res * static_cast<float>(ny)) * 0.5f;
  hz = (res * static_cast<float>(nz)) * 0.5f;
  sigma_update = d.sigma_observation;
  kappa = d.kappa;
  P_detection = d.p_detection;
  new_born_particle_weight = d.new_born_particle_weight;
  nb_each = d.new_born_particle_number_each_point;
  part.assign(static_cast<size_t>(V) * S * 8, 0.f);
  vobj.assign(static_cast<size_t>(V) * (4 + T), 0.f);
  pyr.assign(static_cast<size_t>(NP) * PCAP * 3, 0);
  pc.assign(static_cast<size_t>(NP) * OCAP * 5, 0.f);
  obs_num.assign(NP, 0);
  pc_maxlen.assign(NP, -1.f);
  const float ar = static_cast<float>(d.angle_resolution) / 180.f * 3.14159265358979323846f;
  bp_ori_h.assign((nph + 1) * 3, 0.f); bp_h = bp_ori_h;
  bp_ori_v.assign((npv + 1) * 3, 0.f); bp_v = bp_ori_v;
  const int h_end = d.half_fov_h / d.angle_resolution;
  for (int i = -h_end; i <= h_end; i++) {
    bp_ori_h[(i + h_end) * 3 + 0] = -std::sin(static_cast<float>(i) * ar);
    bp_ori_h[(i + h_end) * 3 + 1] = std::cos(static_cast<float>(i) * ar);
    bp_ori_h[(i + h_end) * 3 + 2] = 0.f;
  }
  const int v_end = d.half_fov_v / d.angle_resolution;
  for (int i = -v_end; i <= v_end; i++) {
    bp_ori_v[(i + v_end) * 3 + 0] = std::sin(static_cast<float>(i) * ar);
    bp_ori_v[(i + v_end) * 3 + 1] = 0.f;
    bp_ori_v[(i + v_end) * 3 + 2] = std::cos(static_cast<float>(i) * ar);
  }
  // findPyramidNeighborIndexInFOV :1201-1218
  nbr.assign(static_cast<size_t>(NP) * 10, 0);
  for (int idx = 0; idx < NP; idx++) {
    const int h0 = idx / npv, v0 = idx % npv;
    int n = 0;
    for (int i = -1; i <= 1; ++i)
      for (int j = -1; j <= 1; ++j) {
        const int h = h0 + i, v = v0 + j;
        if (h >= 0 && h < nph && v >= 0 && v < npv) nbr[idx * 10 + 1 + n++] = h * npv + v;
      }
    nbr[idx * 10] = n;
  }
  // calculateNormalPDFBuffer :1362-1367 with standardNormalPDF :1357-1360 (2*pi/2: the reference's
  // normalisation constant is 1/sqrt(pi))
  pdf.resize(20000);
  for (int i = 0; i < 20000; ++i) {
    const float value = static_cast<float>(i - 10000) * 0.001f;
    pdf[i] = (1.f / (std::sqrt(2.f * 1.57079632679489661923f))) * std::exp(-std::pow(value, 2.f) / (2));
  }
}

int DspMap::voxel_index(float px, float py, float pz, int& index) const {  // :1148-1163
  if (px >= hx || px <= -hx || py >= hy || py <= -hy || pz >= hz || pz <= -hz) return 0;
  const int x = static_cast<int>((px + hx) / res);
  const int y = static_cast<int>((py + hy) / res);
  const int z = static_cast<int>((pz + hz) / res);
  index = z * ny * nx + y * nx + x;
  if (index < 0 || index >= V) return 0;
  return 1;
}

// rotateVectorByQuaternion :1379-1401: att * (0,v) * att.inverse()
void DspMap::rotate(const float* v, const float* q, float* out) {
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

bool DspMap::in_pyramids(float x, float y, float z) const {  // :1411-1427
  const float* h0 = &bp_h[0];
  const float* h1 = &bp_h[nph * 3];
  const float* v0 = &bp_v[0];
  const float* v1 = &bp_v[npv * 3];
  return dot3(x, y, z, h0[0], h0[1], h0[2]) >= 0.f && dot3(x, y, z, h1[0], h1[1], h1[2]) <= 0.f &&
         dot3(x, y, z, v0[0], v0[1], v0[2]) <= 0.f && dot3(x, y, z, v1[0], v1[1], v1[2]) >= 0.f;
}
int DspMap::pyr_h(float x, float y, float z) const {  // :1429-1449
  float last = 1.f;
  for (int i = 0; i < nph; i++) {
    const float* b = &bp_h[(i + 1) * 3];
    const float d = dot3(x, y, z, b[0], b[1], b[2]);
    if (last * d <= 0.f) return i;
    last = d;
  }
  return -1;
}
int DspMap::pyr_v(float x, float y, float z) const {  // :1451-1471
  float last = -1.f;
  for (int j = 0; j < npv; j++) {
    const float* b = &bp_v[(j + 1) * 3];
    const float d = dot3(x, y, z, b[0], b[1], b[2]);
    if (last * d <= 0.f) return j;
    last = d;
  }
  return -1;
}

int DspMap::add_particle(float px, float py, float pz, float vx, float vy, float vz, float w, int vidx) {
  for (int i = 0; i < S; i++) {  // addAParticle :1257-1276
    float* q = P(vidx, i);
    if (q[0] < 0.1f) {
      q[0] = 15.f; q[1] = vx; q[2] = vy; q[3] = vz; q[4] = px; q[5] = py; q[6] = pz; q[7] = w;
      return 1;
    }
  }
  return 0;
}

// moveParticle :1281-1353
int DspMap::move_particle(int new_v, int cur_v, int cur_p, float* flag_ptr) {
  int new_p = cur_p;
  if (new_v != cur_v) {
    *flag_ptr = 0.f;
    int ok = 0;
    for (int i = 0; i < S; ++i) {
      float* q = P(new_v, i);
      if (q[0] < 0.1f) {
        new_p = i; ok = 1;
        q[0] = 7.f;
        for (int k = 1; k < 8; ++k) q[k] = flag_ptr[k];
        break;
      }
    }
    if (!ok) return -1;
  }
  float* q = P(new_v, new_p);
  if (in_pyramids(q[4], q[5], q[6])) {
    const int h = pyr_h(q[4], q[5], q[6]);
    const int v = pyr_v(q[4], q[5], q[6]);
    const int pi = h * npv + v;
    int ok = 0;
    for (int j = 0; j < PCAP; j++) {
      int* e = &pyr[(static_cast<size_t>(pi) * PCAP + j) * 3];
      if (e[0] == 0) { e[0] |= 1; e[1] = new_v; e[2] = new_p; ok = 1; break; }
    }
    if (!ok) { q[0] = 0.f; return -2; }
    if (std::fabs(q[1] * q[2] * q[3]) < 1e-6) {
    } else {
      q[1] += vel_gauss();
      q[2] += vel_gauss();
      q[3] = 0.f;
    }
  }
  return 1;
}

// mapPrediction :663-748 (LIMIT_MOVEMENT_IN_XY_PLANE = 1, CONSIDER_LOCALIZATION_UNCERTAINTY undefined)
void DspMap::prediction(float dx, float dy, float dz, float dt) {
  update_time += dt;
  update_counter += 1;
  for (size_t i = 0; i < pyr.size(); i += 3) pyr[i] &= 0;
  for (int v = 0; v < V; ++v)
    for (int p = 0; p < S; p++) {
      float* q = P(v, p);
      if (q[0] > 0.1f && q[0] < 6.f) {
        q[0] = 1.f;
        if (std::fabs(q[1] * q[2] * q[3]) < 1e-6) {
        } else {
          q[1] += vel_gauss(); q[2] += vel_gauss(); q[3] += vel_gauss();
        }
        q[3] = 0.f;
        q[4] += dt * q[1] + dx;
        q[5] += dt * q[2] + dy;
        q[6] += dt * q[3] + dz;
        int nv;
        if (voxel_index(q[4], q[5], q[6], nv)) {
          const int f = move_particle(nv, v, p, q);
          if (f == -2) { ++stat_pyr_full; continue; }
          if (f == -1) { ++stat_vox_full; continue; }
        } else {
          q[0] = 0.f;
          ++stat_out;
        }
      }
    }
}

// mapUpdate :750-850
void DspMap::map_update() {
  for (int i = 0; i < NP; ++i)
    for (int j = 0; j < obs_num[i]; ++j) {
      float* z = &pc[(static_cast<size_t>(i) * OCAP + j) * 5];
      for (int n = 0; n < nbr[i * 10]; ++n) {
        const int pi = nbr[i * 10 + 1 + n];
        for (int s = 0; s < PCAP; ++s) {
          const int* e = &pyr[(static_cast<size_t>(pi) * PCAP + s) * 3];
          if (e[0] & 1) {
            const float* q = P(e[1], e[2]);
            const float gk = query_pdf(q[4], z[0], sigma_update) * query_pdf(q[5], z[1], sigma_update) *
                             query_pdf(q[6], z[2], sigma_update);
            z[3] += P_detection * q[7] * gk;
          }
        }
      }
      z[3] += (expected_new_born_objects + kappa);
    }
  for (int i = 0; i < NP; i++)
    for (int s = 0; s < PCAP; s++) {
      const int* e = &pyr[(static_cast<size_t>(i) * PCAP + s) * 3];
      if (!(e[0] & 1)) continue;
      float* q = P(e[1], e[2]);
      const float px = q[4], py = q[5], pz = q[6];
      const float dist = std::sqrt(px * px + py * py + pz * pz);
      if (pc_maxlen[i] > 0.f && dist > pc_maxlen[i] + c.obstacle_thickness_for_occlusion) continue;
      float sum = 0.f;
      for (int n = 0; n < nbr[i * 10]; ++n) {
        const int ni = nbr[i * 10 + 1 + n];
        for (int zs = 0; zs < obs_num[ni]; ++zs) {
          const float* z = &pc[(static_cast<size_t>(ni) * OCAP + zs) * 5];
          const float gk = query_pdf(px, z[0], sigma_update) * query_pdf(py, z[1], sigma_update) *
                           query_pdf(pz, z[2], sigma_update);
          sum += P_detection * gk / z[3];
        }
      }
      q[7] *= ((1 - P_detection) + sum);
    }
}

// minimum-cost assignment on a rows x cols matrix (rows <= cols after padding); returns col of each row
static std::vector<int> assign_min_cost(const std::vector<double>& cost, int nr, int nc) {
  // shortest augmenting path (e-maxx formulation) on the matrix padded to n x m with n <= m
  const bool transpose = nr > nc;
  const int n = transpose ? nc : nr, m = transpose ? nr : nc;
  auto a = [&](int i, int j) { return transpose ? cost[static_cast<size_t>(j) * nc + i] : cost[static_cast<size_t>(i) * nc + j]; };
  const double INF = std::numeric_limits<double>::infinity();
  std::vector<double> u(n + 1, 0), v(m + 1, 0);
  std::vector<int> p(m + 1, 0), way(m + 1, 0);
  for (int i = 1; i <= n; ++i) {
    p[0] = i;
    int j0 = 0;
    std::vector<double> minv(m + 1, INF);
    std::vector<char> used(m + 1, 0);
    do {
      used[j0] = 1;
      const int i0 = p[j0];
      double delta = INF;
      int j1 = 0;
      for (int j = 1; j <= m; ++j)
        if (!used[j]) {
          const double cur = a(i0 - 1, j - 1) - u[i0] - v[j];
          if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
          if (minv[j] < delta) { delta = minv[j]; j1 = j; }
        }
      for (int j = 0; j <= m; ++j)
        if (used[j]) { u[p[j]] += delta; v[j] -= delta; } else { minv[j] -= delta; }
      j0 = j1;
    } while (p[j0] != 0);
    do { const int j1 = way[j0]; p[j0] = p[j1]; j0 = j1; } while (j0);
  }
  std::vector<int> res(nr, -1);
  for (int j = 1; j <= m; ++j)
    if (p[j] > 0) { if (transpose) res[j - 1] = p[j] - 1; else res[p[j] - 1] = j - 1; }
  return res;
}

// velocityEstimationThread :1487-1683
void DspMap::velocity_estimation() {
  if (cloud_rot.empty()) return;
  cloud_vel.clear();
  std::vector<V3f> stat, ngr;
  for (const V3f& p : cloud_rot) {
    V3f t{p.x + cur_pos[0], p.y + cur_pos[1], p.z + cur_pos[2]};
    if (t.z > c.voxel_filter_resolution) ngr.push_back(t); else stat.push_back(t);
  }
  std::vector<Cluster> dyn;
  std::vector<std::vector<int>> clusters;
  std::vector<char> possibly_dynamic;
  if (!ngr.empty()) {
    // EuclideanClusterExtraction(tolerance 2*res, min 5, max 10000): components of the radius graph
    const int n = static_cast<int>(ngr.size());
    const float tol = 2 * c.voxel_filter_resolution;
    const float tol2 = tol * tol;
    std::vector<int> parent(n);
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int a) { while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; } return a; };
    for (int i = 0; i < n; i++)
      for (int j = i + 1; j < n; j++) {
        const float dx = ngr[i].x - ngr[j].x, dy = ngr[i].y - ngr[j].y, dz = ngr[i].z - ngr[j].z;
        if (dx * dx + dy * dy + dz * dz <= tol2) {
          const int a = find(i), b = find(j);
          if (a != b) parent[std::max(a, b)] = std::min(a, b);
        }
      }
    std::map<int, std::vector<int>> comp;
    for (int i = 0; i < n; i++) comp[find(i)].push_back(i);
    for (auto& kv : comp)
      if (kv.second.size() >= 5 && kv.second.size() <= 10000) clusters.push_back(kv.second);
    std::stable_sort(clusters.begin(), clusters.end(),
                     [](const std::vector<int>& a, const std::vector<int>& b) { return a.size() > b.size(); });
    for (const auto& ci : clusters) {
      Cluster cl;
      cl.intensity = rand_float(0.1f, 1.f);
      for (int k : ci) { cl.cx += ngr[k].x; cl.cy += ngr[k].y; cl.cz += ngr[k].z; ++cl.n; }
      cl.cx /= static_cast<float>(cl.n); cl.cy /= static_cast<float>(cl.n); cl.cz /= static_cast<float>(cl.n);
      if (static_cast<int>(ci.size()) > c.dynamic_cluster_max_point_num || cl.cz > c.dynamic_cluster_max_center_height) {
        for (int k : ci) stat.push_back(ngr[k]);
        possibly_dynamic.push_back(0);
      } else {
        dyn.push_back(cl);
        possibly_dynamic.push_back(1);
      }
    }
    const float distance_gate = 1.5f, maximum_velocity = 5.f;
    const int point_num_gate = 100;
    if (!clusters_last.empty() && !dyn.empty() && delt_t_last > 0.00001 && delt_t_last < 10.0) {
      const int nr = static_cast<int>(dyn.size()), nc = static_cast<int>(clusters_last.size());
      std::vector<double> cost(static_cast<size_t>(nr) * nc);
      std::vector<float> gate(static_cast<size_t>(nr) * nc);
      for (int r = 0; r < nr; ++r)
        for (int q = 0; q < nc; ++q) {
          const float ddx = dyn[r].cx - clusters_last[q].cx, ddy = dyn[r].cy - clusters_last[q].cy,
                      ddz = dyn[r].cz - clusters_last[q].cz;
          const float d = std::sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
          if (std::abs(dyn[r].n - clusters_last[q].n) > point_num_gate || d >= distance_gate) {
            gate[r * nc + q] = 0.f; cost[r * nc + q] = distance_gate * 5000.f;
          } else {
            gate[r * nc + q] = 1.f; cost[r * nc + q] = d / distance_gate * 1000.f;
          }
        }
      const std::vector<int> asg = assign_min_cost(cost, nr, nc);  // Munkres marks the assignment with zeros
      for (int r = 0; r < nr; ++r) {
        const int q = asg[r];
        if (q < 0 || !(gate[r * nc + q] > 0.01f)) continue;
        Cluster& d = dyn[r];
        d.match = q;
        d.vx = (d.cx - clusters_last[q].cx) / delt_t_last;
        d.vy = (d.cy - clusters_last[q].cy) / delt_t_last;
        d.vz = (d.cz - clusters_last[q].cz) / delt_t_last;
        d.v = std::sqrt(d.vx * d.vx + d.vy * d.vy + d.vz * d.vz);
        d.intensity = clusters_last[q].intensity;
        if (d.v > maximum_velocity) { d.v = 0.f; d.vx = d.vy = d.vz = 0.f; }
      }
    }
    int seq = 0, dseq = 0;
    for (const auto& ci : clusters) {
      if (possibly_dynamic[seq]) {
        for (int k : ci)
          cloud_vel.push_back(VelPoint{ngr[k].x, ngr[k].y, ngr[k].z, dyn[dseq].intensity, dyn[dseq].vx, dyn[dseq].vy, dyn[dseq].vz});
        ++dseq;
      }
      ++seq;
    }
  }
  for (const V3f& s : stat) cloud_vel.push_back(VelPoint{s.x, s.y, s.z, 0.f, 0.f, 0.f, 0.f});
  clusters_last = dyn;
}

// mapAddNewBornParticlesByObservation :853-991
void DspMap::add_newborn() {
  float norm = 0.f;
  for (int i = 0; i < NP; i++)
    for (int j = 0; j < obs_num[i]; j++) norm += 1.f / pc[(static_cast<size_t>(i) * OCAP + j) * 5 + 3];
  const float w_born = new_born_particle_weight * norm;
  const int min_static = static_cast<int>(static_cast<float>(nb_each) * 0.15f);
  const int model_n = static_cast<int>(static_cast<float>(nb_each) * 0.8f);
  for (const VelPoint& pt : cloud_vel) {
    const float cx = pt.x - cur_pos[0], cy = pt.y - cur_pos[1], cz = pt.z - cur_pos[2];
    int pv;
    float ws = 0.f, wd = 0.f, wsd = 0.f;
    if (voxel_index(cx, cy, cz, pv)) {
      for (int kk = 0; kk < S; ++kk) {
        const float* q = P(pv, kk);
        if (q[0] > 0.9f && q[0] < 14.f) {
          const float vabs = std::fabs(q[1]) + std::fabs(q[2]) + std::fabs(q[3]);
          if (vabs < 0.1f) ws += q[7]; else if (vabs < 0.5f) wsd += q[7]; else wd += q[7];
        }
      }
    } else {
      continue;
    }
    const float tot = ws + wd + wsd;
    const float ms = ws / tot, md = wd / tot, msd = wsd / tot;
    const float p_s = (ms + ms + msd) * 0.5f, p_d = (md + md + msd) * 0.5f;
    const float np_ = p_s + p_d;
    const float psn = p_s / np_;
    // (int)NaN is INT_MIN on x86-64 (cvttss2si); either way max() below yields min_static
    const float fs = static_cast<float>(model_n) * psn;
    int n_static = std::isnan(fs) ? std::numeric_limits<int>::min() : static_cast<int>(fs);
    n_static = std::max(min_static, n_static);
    for (int p = 0; p < nb_each; p++) {
      const float px = cx + pos_gauss();
      const float py = cy + pos_gauss();
      const float pz = cz + pos_gauss();
      int vi;
      if (voxel_index(px, py, pz, vi)) {
        float vx, vy, vz;
        if (p < n_static) {
          vx = vy = vz = 0.f;
        } else if (pt.nx > -100.f && p < model_n) {
          if (pt.intensity > 0.01f) {
            vx = pt.nx + 4 * vel_gauss(); vy = pt.ny + 4 * vel_gauss(); vz = pt.nz + 4 * vel_gauss();
          } else { vx = vy = vz = 0.f; }
        } else {
          if (pt.intensity > 0.01f) {
            vx = rand_float(-1.5f, 1.5f); vy = rand_float(-1.5f, 1.5f); vz = rand_float(-0.5f, 0.5f);
          } else { vx = vy = vz = 0.f; }
        }
        vz = 0.f;
        if (add_particle(px, py, pz, vx, vy, vz, w_born, vi) > 0) ++stat_born;
      }
    }
  }
}

// mapOccupancyCalculationAndResample :994-1130
void DspMap::occupancy_resample() {
  const int D = 4 + T, MAXP = c.max_particle_num_voxel;
  for (int v = 0; v < V; ++v) {
    float wsum = 0.f, vxs = 0.f, vys = 0.f, vzs = 0.f;
    int np = 0, nold = 0;
    for (int p = 0; p < S; p++) {
      float* q = P(v, p);
      if (q[0] > 0.1f) {
        if (q[7] < 1e-3) {
          q[0] = 0.f;
        } else {
          if (q[0] < 10.f) {
            ++nold;
            vxs += q[1]; vys += q[2]; vzs += q[3];
            for (int k = 0; k < T; ++k) {
              const float pt = c.prediction_future_time[k];
              const float fx = q[4] + q[1] * pt, fy = q[5] + q[2] * pt, fz = q[6] + q[3] * pt;
              int pi;
              if (voxel_index(fx, fy, fz, pi)) vobj[static_cast<size_t>(pi) * D + 4 + k] += q[7];
            }
          }
          q[0] = 1.f;
          ++np;
          wsum += q[7];
        }
      }
    }
    float* o = &vobj[static_cast<size_t>(v) * D];
    o[0] = wsum;
    if (nold > 0) {
      o[1] = vxs / static_cast<float>(nold); o[2] = vys / static_cast<float>(nold); o[3] = vzs / static_cast<float>(nold);
    } else {
      o[1] = o[2] = o[3] = 0.f;
    }
    if (np < 5) continue;
    const int n_after = np > MAXP ? MAXP : np;
    const float w_after = wsum / static_cast<float>(n_after);
    float acc_ori = 0.f, acc_new = w_after * 0.5f;
    for (int p = 0; p < S; ++p) {
      float* q = P(v, p);
      if (q[0] > 0.7f) {
        acc_ori += q[7];
        if (acc_ori > acc_new) {
          q[7] = w_after;
          acc_new += w_after;
          int full = 0, p_i = 0;
          while (acc_ori > acc_new) {
            int found = 0;
            if (!full) {
              for (; p_i < S; ++p_i) {
                float* e = P(v, p_i);
                if (e[0] < 0.1f) {
                  e[0] = 0.6f;
                  for (int k = 1; k < 8; k++) e[k] = q[k];
                  found = 1;
                  break;
                }
              }
            }
            if (!found) { q[7] += w_after; full = 1; }
            acc_new += w_after;
          }
        } else {
          q[0] = 0.f;
        }
      }
    }
  }
}

// DSPMap::update :165-364
int DspMap::update(int n, const float* pts, const float* pos, double stamp, const float* q) {
  if (first) { last_pos[0] = pos[0]; last_pos[1] = pos[1]; last_pos[2] = pos[2]; last_stamp = stamp; first = false; }
  if (std::fabs(q[0]) > 1.001f || std::fabs(q[1]) > 1.001f || std::fabs(q[2]) > 1.001f || std::fabs(q[3]) > 1.001f) return 0;
  // non-finite odometry passes the reference's comparisons; rejected here and in the CUDA path
  for (int k = 0; k < 4; k++) if (!std::isfinite(q[k])) return 0;
  for (int k = 0; k < 3; k++) if (!std::isfinite(pos[k])) return 0;
  if (!std::isfinite(stamp)) return 0;
  const float dx = pos[0] - last_pos[0], dy = pos[1] - last_pos[1], dz = pos[2] - last_pos[2];
  const float dt = static_cast<float>(stamp - last_stamp);
  if (std::fabs(dx) > 10.f || std::fabs(dy) > 10.f || std::fabs(dz) > 10.f || dt < 0.f || dt > 10.f) return 0;
  for (int k = 0; k < 3; k++) cur_pos[k] = last_pos[k] = pos[k];
  last_stamp = stamp;
  delt_t_last = dt;
  for (int k = 0; k < 4; k++) quat[k] = q[k];
  for (int i = 0; i < nph + 1; i++) rotate(&bp_ori_h[i * 3], quat, &bp_h[i * 3]);
  for (int j = 0; j < npv + 1; j++) rotate(&bp_ori_v[j * 3], quat, &bp_v[j * 3]);
  for (int i = 0; i < NP; i++) { obs_num[i] = 0; pc_maxlen[i] = -1.f; }
  cloud_rot.clear();
  int valid = 0;
  for (int s = 0; s < n; ++s) {
    float r[3];
    rotate(&pts[3 * s], quat, r);
    cloud_rot.push_back(V3f{r[0], r[1], r[2]});
    if (in_pyramids(r[0], r[1], r[2])) {
      const int ih = pyr_h(r[0], r[1], r[2]), iv = pyr_v(r[0], r[1], r[2]);
      const int pi = ih * npv + iv;
      const int k = obs_num[pi];
      const float len = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
      float* z = &pc[(static_cast<size_t>(pi) * OCAP + k) * 5];
      z[0] = r[0]; z[1] = r[1]; z[2] = r[2]; z[3] = 0.f; z[4] = len;
      if (pc_maxlen[pi] < len) pc_maxlen[pi] = len;
      obs_num[pi] += 1;
      if (obs_num[pi] >= OCAP) obs_num[pi] = OCAP - 1;
      ++valid;
    }
  }
  expected_new_born_objects = new_born_particle_weight * static_cast<float>(valid) * static_cast<float>(nb_each);
  new_born_each_object_weight = new_born_particle_weight * static_cast<float>(nb_each);
  velocity_estimation();  // a side thread in the reference; it shares only rand() with the steps below,
                          // which do not draw from it until the thread is joined
  prediction(-dx, -dy, -dz, dt);
  if (n >= 0) map_update();
  if (n >= 0) add_newborn();
  occupancy_resample();
  return 1;
}

// risk_mapping_node.cpp:74-99 filterPointCloud: pcl::VoxelGrid at filter_res, camera -> body axis swap, range
// filter in double (LOCAL_RANGE_* are constexpr double, :34-36), cap max_points
static int filter_cloud(const pop_config& g, const pop_dsp_config& d, const float* cam, int n, float* out) {
  const float leaf = d.voxel_filter_resolution;
  const float inv = 1.0f / leaf;
  std::vector<int> fin;
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-mn[0], -mn[1], -mn[2]};
  for (int i = 0; i < n; i++) {
    const float* p = cam + 3 * i;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    fin.push_back(i);
    for (int k = 0; k < 3; k++) { mn[k] = std::min(mn[k], p[k]); mx[k] = std::max(mx[k], p[k]); }
  }
  int nout = 0;
  if (fin.empty()) return 0;
  int minb[3], maxb[3], div[3];
  for (int k = 0; k < 3; k++) {
    minb[k] = static_cast<int>(std::floor(mn[k] * inv));
    maxb[k] = static_cast<int>(std::floor(mx[k] * inv));
    div[k] = maxb[k] - minb[k] + 1;
  }
  std::vector<std::pair<int64_t, int>> idx;
  idx.reserve(fin.size());
  for (int i : fin) {
    const float* p = cam + 3 * i;
    const int i0 = static_cast<int>(std::floor(p[0] * inv)) - minb[0];
    const int i1 = static_cast<int>(std::floor(p[1] * inv)) - minb[1];
    const int i2 = static_cast<int>(std::floor(p[2] * inv)) - minb[2];
    idx.push_back({static_cast<int64_t>(i0) + static_cast<int64_t>(i1) * div[0] + static_cast<int64_t>(i2) * div[0] * div[1], i});
  }
  std::stable_sort(idx.begin(), idx.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  const double RX = g.nx / 2.f * static_cast<double>(g.resolution), RY = g.ny / 2.f * static_cast<double>(g.resolution),
               RZ = g.nz / 2.f * static_cast<double>(g.resolution);
  size_t a = 0;
  while (a < idx.size()) {
    size_t b = a;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    while (b < idx.size() && idx[b].first == idx[a].first) {
      const float* p = cam + 3 * idx[b].second;
      sx += p[0]; sy += p[1]; sz += p[2];
      ++b;
    }
    const float cnt = static_cast<float>(b - a);
    const float cx = sx / cnt, cy = sy / cnt, cz = sz / cnt;
    const float x = cz, y = -cx, z = -cy;
    if (x > -RX && x < RX && y > -RY && y < RY && z > -RZ && z < RZ) {
      out[nout * 3 + 0] = x; out[nout * 3 + 1] = y; out[nout * 3 + 2] = z;
      nout++;
      if (nout >= d.max_points) break;
    }
    a = b;
  }
  return nout;
}

}  // namespace orc

using namespace orc;

struct OrcDsp {
  pop_config g;
  DspMap m;
  std::vector<float> filtered;
};

extern "C" {

void* orc_dsp_create(const pop_config* g, const pop_dsp_config* d) {
  OrcDsp* o = new OrcDsp();
  o->g = *g;
  o->m.init(*g, *d);
  o->filtered.resize(static_cast<size_t>(d->max_points) * 3);
  return o;
}
void orc_dsp_destroy(void* h) { delete static_cast<OrcDsp*>(h); }
void orc_dsp_set_random_tables(void* h, const float* pg, const float* vg, const int32_t* ur, int n) {
  DspMap& m = static_cast<OrcDsp*>(h)->m;
  m.pg.assign(pg, pg + n); m.vg.assign(vg, vg + n); m.ur.assign(ur, ur + n);
  m.pseq = m.vseq = m.useq = 0;
}
// filterPointCloud only: returns the number of valid body-frame points written to out (cap max_points)
int orc_dsp_filter(void* h, const float* cam, int n, float* out) {
  OrcDsp* o = static_cast<OrcDsp*>(h);
  return filter_cloud(o->g, o->m.c, cam, n, out);
}
// DSPMap::update on already filtered body-frame points (valid_clouds_)
int orc_dsp_update_filtered(void* h, const float* pts, int n, const float* pos, const float* quat, double stamp) {
  return static_cast<OrcDsp*>(h)->m.update(n, pts, pos, stamp, quat);
}
// cloudCallback: filterPointCloud + update
int orc_dsp_update(void* h, const float* cam, int n, const float* pos, const float* quat, double stamp) {
  OrcDsp* o = static_cast<OrcDsp*>(h);
  const int nv = filter_cloud(o->g, o->m.c, cam, n, o->filtered.data());
  return o->m.update(nv, o->filtered.data(), pos, stamp, quat);
}
// getOccupancyMapWithFutureStatus (:445-468) + futureRiskCallback copy: out = [V][T] future risk; clears
// the accumulators; returns the number of voxels whose occupancy weight exceeds thr
int orc_dsp_export(void* h, float* out, float thr) {
  DspMap& m = static_cast<OrcDsp*>(h)->m;
  const int D = 4 + m.T;
  int occ = 0;
  for (int i = 0; i < m.V; i++) {
    if (m.vobj[static_cast<size_t>(i) * D] > thr) ++occ;
    for (int k = 0; k < m.T; k++) {
      out[static_cast<size_t>(i) * m.T + k] = m.vobj[static_cast<size_t>(i) * D + 4 + k];
      m.vobj[static_cast<size_t>(i) * D + 4 + k] = 0.f;
    }
  }
  return occ;
}
void orc_dsp_particles(void* h, float* out) {  // [V][S][8]
  DspMap& m = static_cast<OrcDsp*>(h)->m;
  std::copy(m.part.begin(), m.part.end(), out);
}
void orc_dsp_voxels(void* h, float* out) {  // [V][4]: occupancy weight, mean velocity
  DspMap& m = static_cast<OrcDsp*>(h)->m;
  const int D = 4 + m.T;
  for (int i = 0; i < m.V; i++)
    for (int k = 0; k < 4; k++) out[static_cast<size_t>(i) * 4 + k] = m.vobj[static_cast<size_t>(i) * D + k];
}
void orc_dsp_observations(void* h, float* pc, int32_t* counts, float* maxlen) {
  DspMap& m = static_cast<OrcDsp*>(h)->m;
  std::copy(m.pc.begin(), m.pc.end(), pc);
  std::copy(m.obs_num.begin(), m.obs_num.end(), counts);
  std::copy(m.pc_maxlen.begin(), m.pc_maxlen.end(), maxlen);
}
void orc_dsp_pyramids(void* h, int32_t* out) {
  DspMap& m = static_cast<OrcDsp*>(h)->m;
  std::copy(m.pyr.begin(), m.pyr.end(), out);
}
int orc_dsp_points_with_velocity(void* h, float* out, int cap) {
  DspMap& m = static_cast<OrcDsp*>(h)->m;
  const int n = static_cast<int>(m.cloud_vel.size());
  for (int i = 0; i < std::min(n, cap); i++) {
    const VelPoint& p = m.cloud_vel[i];
    float* o = out + 7 * i;
    o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = p.intensity; o[4] = p.nx; o[5] = p.ny; o[6] = p.nz;
  }
  return n;
}
// the assignment step alone (cross-checked against scipy in the CPU tests): out[r] = column or -1
void orc_dsp_assign(const double* cost, int nr, int nc, int32_t* out) {
  const std::vector<int> a = assign_min_cost(std::vector<double>(cost, cost + static_cast<size_t>(nr) * nc), nr, nc);
  std::copy(a.begin(), a.end(), out);
}
void orc_dsp_stats(void* h, int32_t* out) {  // voxel-full, pyramid-full, moved out, newborn, seq p/v/u
  DspMap& m = static_cast<OrcDsp*>(h)->m;
  out[0] = m.stat_vox_full; out[1] = m.stat_pyr_full; out[2] = m.stat_out; out[3] = m.stat_born;
  out[4] = m.pseq; out[5] = m.vseq; out[6] = m.useq; out[7] = m.update_counter;
}

}  // extern "C"
