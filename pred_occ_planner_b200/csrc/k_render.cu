// k_render.cu — K0: the simulator's depth sensor, batched over agents, feeding the particle map on the device.
//
// Reference: uav_simulator/local_sensing/src/depth_render.cu:2-58 (the reference's only CUDA kernel: point
// splatting into an integer z-buffer with atomicMin), pcl_render_node.cpp:108-136 (camera pose from the
// odometry), :219-254 (depth image -> points inside the sensing horizon), :263-285 (millimetres -> metres).
// Differences from the reference kernel: all agents share one launch (blockIdx.y = agent) over the world
// cloud already resident for the ground-truth map; a point's footprint is written by a half-warp instead
// of a serial double loop; the back-projection compacts the pixels in the reference's (u, v) order with a
// block scan and leaves the camera-frame cloud where the voxel filter reads it — no image or cloud crosses
// PCIe.  Depth values are integers combined by atomicMin, so the image is independent of thread order.
#include <cmath>
#include <cstring>

#include "k_dsp.cuh"

namespace {

__global__ void __launch_bounds__(256) k_depth_init(int* __restrict__ depth, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) depth[i] = 999999;   // depth_initial
}

// 16 lanes per world point: the lanes stride over the (2r+1)^2 footprint
__global__ void __launch_bounds__(256) k_depth_splat(pop_camera cam, const float* __restrict__ cloud, int n_cloud,
                                                     const float* __restrict__ rpose_all, int* __restrict__ depth_all) {
  const int a = blockIdx.y;
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, lane = threadIdx.x & 15;
  if (k >= n_cloud) return;
  const float* P = rpose_all + 12 * a;   // r row-major, t
  const float px = cloud[3 * k], py = cloud[3 * k + 1], pz = cloud[3 * k + 2];
  const float tx = px * P[0] + py * P[1] + pz * P[2] + P[9];
  const float ty = px * P[3] + py * P[4] + pz * P[5] + P[10];
  const float tz = px * P[6] + py * P[7] + pz * P[8] + P[11];
  if (tz <= 0.0f) return;
  const int ux = static_cast<int>(tx / tz * cam.fx + cam.cx + 0.5);
  const int uy = static_cast<int>(ty / tz * cam.fy + cam.cy + 0.5);
  if (ux < 0 || ux >= cam.width || uy < 0 || uy >= cam.height) return;
  const int dist_mm = static_cast<int>(tz * 1000.0f + 0.5f);
  // footprint radius; a radius beyond the image covers the whole image, so it is capped before the int conversion
  // (a point millimetres from the lens would otherwise give an astronomically large loop — and an out-of-range
  // double -> int conversion in the reference)
  const double rd = 0.0573 * cam.fx / tz + 0.5f;
  const int rad = rd > 8192.0 ? 8192 : static_cast<int>(rd);
  // only the part of the footprint inside the image is visited (the reference skips the rest pixel by pixel)
  const int x0 = max(ux - rad, 0), x1 = min(ux + rad, cam.width - 1);
  const int y0 = max(uy - rad, 0), y1 = min(uy + rad, cam.height - 1);
  const int sx = x1 - x0 + 1, sy = y1 - y0 + 1;
  int* depth = depth_all + static_cast<size_t>(a) * cam.width * cam.height;
  for (int e = lane; e < sx * sy; e += 16) {
    const int x = x0 + e % sx, y = y0 + e / sx;
    atomicMin(depth + static_cast<size_t>(y) * cam.width + x, dist_mm);
  }
}

__device__ __forceinline__ int scan_excl_1024(int v, int* s_w, int* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) s_w[w] = inc;
  __syncthreads();
  if (w == 0) {
    int x = lane < nw ? s_w[lane] : 0, xi = x;
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

// render_currentpose + render_pcl_world: one CTA per agent walks the pixels u-major and compacts in order
__global__ void __launch_bounds__(1024) k_depth_to_cloud(pop_camera cam, const int* __restrict__ depth_all,
                                                         const double* __restrict__ rpose_d_all,
                                                         float* __restrict__ cam_all, int max_in,
                                                         int* __restrict__ counts, int* __restrict__ err) {
  __shared__ int s_w[33];
  const int a = blockIdx.x;
  const int* depth = depth_all + static_cast<size_t>(a) * cam.width * cam.height;
  const double* M = rpose_d_all + 20 * a;   // cam2world row-major, then the body position
  float* out = cam_all + static_cast<size_t>(a) * max_in * 3;
  const int npix = cam.width * cam.height;
  int carry = 0;
  for (int base = 0; base < npix; base += blockDim.x) {
    const int i = base + threadIdx.x;
    bool keep = false;
    double pc[3] = {0, 0, 0};
    if (i < npix) {
      const int u = i / cam.height, v = i % cam.height;
      float d = static_cast<float>(depth[static_cast<size_t>(v) * cam.width + u]) / 1000.0f;
      d = d < 500.0f ? d : 0;
      if (d != 0.0f) {
        pc[0] = (u - static_cast<double>(cam.cx)) * d / cam.fx;
        pc[1] = (v - static_cast<double>(cam.cy)) * d / cam.fy;
        pc[2] = d;
        double pw[3];
        for (int r = 0; r < 3; r++)
          pw[r] = ((M[4 * r] * pc[0] + M[4 * r + 1] * pc[1]) + M[4 * r + 2] * pc[2]) + M[4 * r + 3] * 1.0;
        const double dx = pw[0] - M[16], dy = pw[1] - M[17], dz = pw[2] - M[18];
        keep = !(sqrt(dx * dx + dy * dy + dz * dz) > cam.sensing_horizon);
      }
    }
    int total;
    const int pos = carry + scan_excl_1024(keep ? 1 : 0, s_w, &total);
    if (keep && pos < max_in) {
      out[3 * pos] = static_cast<float>(pc[0]);
      out[3 * pos + 1] = static_cast<float>(pc[1]);
      out[3 * pos + 2] = static_cast<float>(pc[2]);
    }
    carry += total;
  }
  if (threadIdx.x == 0) {
    if (carry > max_in) atomicOr(err, 128);
    counts[a] = carry < max_in ? carry : max_in;
  }
}

// Eigen::Quaterniond::toRotationMatrix
void quat_to_rot(const double* q, double R[9]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

}  // namespace

extern "C" {

void pop_camera_default(pop_camera* c, int width, int height) {
  memset(c, 0, sizeof(*c));
  // uav_simulator/local_sensing/params/camera.yaml scaled from 640x480
  const double sx = width / 640.0, sy = height / 480.0;
  c->width = width; c->height = height;
  c->fx = static_cast<float>(387.229248046875 * sx);
  c->fy = static_cast<float>(387.229248046875 * sy);
  c->cx = static_cast<float>(321.04638671875 * sx);
  c->cy = static_cast<float>(243.44969177246094 * sy);
  c->sensing_horizon = 6.0f;   // plan_manager/launch/sim_baseline_plan.launch:128
  const double m[16] = {0.0, 0.0, 1.0, 0.0, -1.0, 0.0, 0.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0};
  memcpy(c->cam2body, m, sizeof(m));   // pcl_render_node.cpp:338
}

int pop_render_set_camera(pop_handle* h, const pop_camera* cam) {
  if (!h || !cam) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->dsp) return pop_fail(h, POP_E_STATE, "attach the particle map first (pop_dsp_attach): the sensor renders into its input buffers");
  if (cam->width < 1 || cam->height < 1 || !(cam->fx > 0) || !(cam->fy > 0)) return pop_fail(h, POP_E_INVALID, "bad camera");
  DspDev* d = h->dsp;
  const size_t npix = static_cast<size_t>(cam->width) * cam->height;
  if (d->depth) { cudaFree(d->depth); d->depth = nullptr; }
  if (!d->cam_counts) POP_CUDA(h, cudaMalloc(&d->cam_counts, sizeof(int) * d->nl));
  if (!d->rpose) POP_CUDA(h, cudaMalloc(&d->rpose, sizeof(float) * 12 * d->nl));
  if (!d->rpose_d) POP_CUDA(h, cudaMalloc(&d->rpose_d, sizeof(double) * 20 * d->nl));
  POP_CUDA(h, cudaMalloc(&d->depth, sizeof(int) * npix * d->nl));
  d->camera = *cam;
  d->have_camera = true;
  return POP_OK;
}

int pop_render_depth(pop_handle* h, const double* pos, const double* quat, int32_t* out_depth_mm, int32_t* out_counts) {
  if (!h || !pos || !quat) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->dsp || !h->dsp->have_camera) return pop_fail(h, POP_E_STATE, "no camera (pop_render_set_camera)");
  DspDev* d = h->dsp;
  const pop_camera& c = d->camera;
  const int A = d->nl;
  std::vector<float> rp(12 * A);
  std::vector<double> rd(20 * A);
  for (int a = 0; a < A; a++) {
    double R[9];
    quat_to_rot(quat + 4 * a, R);
    const double* p = pos + 3 * a;
    const double B[16] = {R[0], R[1], R[2], p[0], R[3], R[4], R[5], p[1], R[6], R[7], R[8], p[2], 0, 0, 0, 1};
    double* M = &rd[20 * a];   // cam2world = body_pose * cam02body
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        double s = 0.0;
        for (int k = 0; k < 4; k++) s += B[4 * i + k] * c.cam2body[4 * k + j];
        M[4 * i + j] = s;
      }
    M[16] = p[0]; M[17] = p[1]; M[18] = p[2]; M[19] = 0.0;
    // cam_pose = cam2world^-1 as a rigid transform, narrowed to the kernel's float parameters
    float* P = &rp[12 * a];
    for (int i = 0; i < 3; i++) {
      double s = 0.0;
      for (int j = 0; j < 3; j++) {
        P[3 * i + j] = static_cast<float>(M[4 * j + i]);
        s += M[4 * j + i] * M[4 * j + 3];
      }
      P[9 + i] = static_cast<float>(-s);
    }
  }
  cudaStream_t st = h->stream;
  POP_CUDA(h, cudaMemcpyAsync(d->rpose, rp.data(), sizeof(float) * rp.size(), cudaMemcpyHostToDevice, st));
  POP_CUDA(h, cudaMemcpyAsync(d->rpose_d, rd.data(), sizeof(double) * rd.size(), cudaMemcpyHostToDevice, st));
  POP_CUDA(h, cudaStreamSynchronize(st));   // rp / rd are stack-owned
  const size_t npix = static_cast<size_t>(c.width) * c.height;
  k_depth_init<<<static_cast<unsigned>((npix * A + 255) / 256), 256, 0, st>>>(d->depth, npix * A);
  POP_LAUNCH_CHECK(h);
  if (h->n_cloud > 0) {
    const dim3 g((static_cast<unsigned>(h->n_cloud) * 16 + 255) / 256, A);
    k_depth_splat<<<g, 256, 0, st>>>(c, h->d_cloud, h->n_cloud, d->rpose, d->depth);
    POP_LAUNCH_CHECK(h);
  }
  k_depth_to_cloud<<<A, 1024, 0, st>>>(c, d->depth, d->rpose_d, d->cam, d->G.MAXIN, d->cam_counts, d->err);
  POP_LAUNCH_CHECK(h);
  h->n_launches += 3;
  d->use_cam_counts = true;
  if (out_depth_mm) POP_CUDA(h, cudaMemcpyAsync(out_depth_mm, d->depth, sizeof(int) * npix * A, cudaMemcpyDeviceToHost, st));
  if (out_counts) POP_CUDA(h, cudaMemcpyAsync(out_counts, d->cam_counts, sizeof(int) * A, cudaMemcpyDeviceToHost, st));
  if (out_depth_mm || out_counts) POP_CUDA(h, cudaStreamSynchronize(st));
  return POP_OK;
}

int pop_render_download_cloud(pop_handle* h, int agent, float* out, int cap, int32_t* n_out) {
  if (!h || !out || !n_out || cap < 0) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->dsp || !h->dsp->have_camera) return pop_fail(h, POP_E_STATE, "no camera (pop_render_set_camera)");
  if (agent < 0 || agent >= h->dsp->nl) return POP_E_INVALID;
  DspDev* d = h->dsp;
  int n = 0;
  POP_CUDA(h, cudaMemcpyAsync(&n, d->cam_counts + agent, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  *n_out = n;
  const int m = n < cap ? n : cap;
  if (m > 0)
    POP_CUDA(h, cudaMemcpy(out, d->cam + static_cast<size_t>(agent) * d->G.MAXIN * 3, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost));
  return n > cap ? POP_E_CAPACITY : POP_OK;
}

}  // extern "C"
