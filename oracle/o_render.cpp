// o_render.cpp — oracle: the simulator's depth sensor (TEST INFRASTRUCTURE).  Follows
//   uav_simulator/local_sensing/src/depth_render.cu:2-58   (render / depth_initial kernels, restated on the CPU)
//   uav_simulator/local_sensing/src/pcl_render_node.cpp:108-136 (cam2world from the odometry), :219-254
//   (render_pcl_world: depth image -> points, sensing horizon), :263-285 (render_currentpose: mm -> metres)
// render_depth is PINNED against the reference's own kernel: depth_render.cu compiled unmodified by nvcc into
// oracle/_ref/libref_depth.so and run on the GPU gives the same depth images, pixel for pixel
// (tests/test_gpu_zz_reference_depth.py).  The node around it (pose handling, image -> cloud) is restated without a
// pin; every step is integer or a fixed sequence of IEEE operations, so the restatement is deterministic.  Deviations, documented: the camera pose
// is inverted as a rigid transform (Rᵀ, −Rᵀt) instead of Eigen's general 4x4 inverse; points are returned in
// the CAMERA frame (what filterPointCloud consumes) as well as in the world frame the node publishes.
#include <algorithm>
#include <cmath>

#include "oracle.hpp"

namespace orc {

// Eigen::Quaterniond::toRotationMatrix
static void quat_to_rot(const double* q, double R[9]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

// cam2world = body_pose * cam02body (pcl_render_node.cpp:120-128), row-major 4x4
void render_cam2world(const pop_camera& c, const double* pos, const double* quat, double out[16]) {
  double R[9];
  quat_to_rot(quat, R);
  double B[16] = {R[0], R[1], R[2], pos[0], R[3], R[4], R[5], pos[1], R[6], R[7], R[8], pos[2], 0, 0, 0, 1};
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double s = 0.0;
      for (int k = 0; k < 4; k++) s += B[4 * i + k] * c.cam2body[4 * k + j];
      out[4 * i + j] = s;
    }
}

// depth image in millimetres (999999 = nothing), depth_render.cu:2-58
void render_depth(const pop_camera& c, const double cam2world[16], const float* cloud, int n, int32_t* depth) {
  const int W = c.width, H = c.height;
  std::fill(depth, depth + static_cast<size_t>(W) * H, 999999);
  // cam_pose = cam2world^-1 (rigid), handed to the kernel as floats (Parameter r, t)
  float r[3][3], t[3];
  for (int i = 0; i < 3; i++) {
    double s = 0.0;
    for (int j = 0; j < 3; j++) {
      r[i][j] = static_cast<float>(cam2world[4 * j + i]);
      s += cam2world[4 * j + i] * cam2world[4 * j + 3];
    }
    t[i] = static_cast<float>(-s);
  }
  for (int k = 0; k < n; k++) {
    const float px = cloud[3 * k], py = cloud[3 * k + 1], pz = cloud[3 * k + 2];
    const float tx = px * r[0][0] + py * r[0][1] + pz * r[0][2] + t[0];
    const float ty = px * r[1][0] + py * r[1][1] + pz * r[1][2] + t[1];
    const float tz = px * r[2][0] + py * r[2][1] + pz * r[2][2] + t[2];
    if (tz <= 0.0f) continue;
    const int ux = static_cast<int>(tx / tz * c.fx + c.cx + 0.5);
    const int uy = static_cast<int>(ty / tz * c.fy + c.cy + 0.5);
    if (ux < 0 || ux >= W || uy < 0 || uy >= H) continue;
    const float dist = tz;
    const int dist_mm = static_cast<int>(dist * 1000.0f + 0.5f);
    // a radius beyond the image covers the whole image: capped before the int conversion (out of range in the
    // reference for a point millimetres from the lens)
    const double rd = 0.0573 * c.fx / dist + 0.5f;
    const int rad = rd > 8192.0 ? 8192 : static_cast<int>(rd);
    for (int i = std::max(-rad, -uy); i <= std::min(rad, H - 1 - uy); i++)
      for (int j = std::max(-rad, -ux); j <= std::min(rad, W - 1 - ux); j++) {
        const int x = ux + j, y = uy + i;
        if (x < 0 || x >= W || y < 0 || y >= H) continue;
        int32_t& d = depth[static_cast<size_t>(y) * W + x];
        d = std::min(d, dist_mm);
      }
  }
}

// render_currentpose mm -> m (:270-276) + render_pcl_world (:219-254); out_cam / out_world = n x 3 floats
int render_cloud(const pop_camera& c, const double cam2world[16], const double* body_pos, const int32_t* depth,
                 float* out_cam, float* out_world, int cap) {
  int n = 0;
  for (int u = 0; u < c.width; u++)
    for (int v = 0; v < c.height; v++) {
      float d = static_cast<float>(depth[static_cast<size_t>(v) * c.width + u]) / 1000.0f;
      d = d < 500.0f ? d : 0;
      if (d == 0.0) continue;
      const double pc[4] = {(u - static_cast<double>(c.cx)) * d / c.fx, (v - static_cast<double>(c.cy)) * d / c.fy, d, 1.0};
      double pw[3];
      for (int i = 0; i < 3; i++)  // Eigen matrix * vector: columns accumulated left to right
        pw[i] = ((cam2world[4 * i] * pc[0] + cam2world[4 * i + 1] * pc[1]) + cam2world[4 * i + 2] * pc[2]) + cam2world[4 * i + 3] * pc[3];
      const double dx = pw[0] - body_pos[0], dy = pw[1] - body_pos[1], dz = pw[2] - body_pos[2];
      if (std::sqrt(dx * dx + dy * dy + dz * dz) > c.sensing_horizon) continue;
      if (n < cap) {
        for (int i = 0; i < 3; i++) {
          if (out_cam) out_cam[3 * n + i] = static_cast<float>(pc[i]);
          if (out_world) out_world[3 * n + i] = static_cast<float>(pw[i]);
        }
      }
      n++;
    }
  return n;
}

}  // namespace orc

extern "C" {
// one agent: depth image (W*H int32 mm) and the sensed cloud; returns the number of points
int orc_render(const pop_camera* cam, const double* pos, const double* quat, const float* cloud, int n,
               int32_t* depth, float* out_cam, float* out_world, int cap) {
  double c2w[16];
  orc::render_cam2world(*cam, pos, quat, c2w);
  orc::render_depth(*cam, c2w, cloud, n, depth);
  return orc::render_cloud(*cam, c2w, pos, depth, out_cam, out_world, cap);
}
}
