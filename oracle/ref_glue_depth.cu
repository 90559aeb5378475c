// ref_glue_depth.cu — TEST INFRASTRUCTURE: the reference simulator's depth sensor kernel,
// uav_simulator/local_sensing/src/depth_render.cu (with its own depth_render.cuh, device_image.cuh, cuda_exception.cuh,
// helper_math.h), compiled UNMODIFIED by nvcc for sm_100a into oracle/_ref/libref_depth.so, with one C entry point.
// It needs a GPU to run: tests/test_gpu_zz_reference_depth.py compares its depth images with the oracle's restatement
// (oracle/o_render.cpp: render_depth) and with the product's k_depth_splat on the B200.
#include <depth_render.cu>

extern "C" int refdepth_render(const float* cloud_xyz, int n, float fx, float fy, float cx, float cy, int width, int height,
                               const double* transformation16, int* out_depth) {
  try {
    DepthRender r;
    r.set_para(fx, fy, cx, cy, width, height);
    std::vector<float> data(cloud_xyz, cloud_xyz + 3 * static_cast<size_t>(n));
    r.set_data(data);
    double T[16];
    for (int k = 0; k < 16; k++) T[k] = transformation16[k];
    r.render_pose(T, out_depth);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 2;
  } catch (...) {
    return 1;
  }
}
