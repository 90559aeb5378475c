// k_dsp.cuh — device state of the DSP particle map (K1d), shared by k_dsp.cu (kernels) and
// pop_dsp.cu (C ABI).  One dsp_map::DSPMap per local agent
// (plan_env/include/plan_env/dsp_dynamic.h); all agents of a handle are updated by the same launches.
#pragma once
#include "pop_internal.h"

constexpr int kDspSlotStride = 16;   // bytes of flags per voxel (>= SAFE_PARTICLE_NUM_VOXEL = 14)
constexpr int kDspMaxClusters = 1024;
constexpr int kDspMaxDyn = 256;      // possibly-dynamic clusters matched by the assignment step
constexpr int kDspMaxIter = 6;       // fixed-point rounds of the re-binning (2 in practice)

// particle flags: voxels_with_particle[.][.][0] of the reference (dsp_dynamic.h:88-92)
enum : uint8_t { DSP_F_EMPTY = 0, DSP_F_VALID = 1 /*1.f*/, DSP_F_COPY = 2 /*0.6f*/, DSP_F_MOVED = 3 /*7.f*/,
                 DSP_F_BORN = 4 /*15.f*/ };

struct DspGeo {
  int nx, ny, nz, V, T, S;
  int nph, npv, NP, PCAP, OCAP, MAXPTS, MAXIN, NPOW2, RCAP, NB, MAXP;
  int nrand, cl_maxn;
  float res, hx, hy, hz;
  float sigma, kappa, pd, nbw, filt, thick, cl_maxh, occ_thr;
  double RX, RY, RZ;
  float tf[POP_DSP_MAX_T];
};

struct DspFrame {   // one agent, one update (host -> device)
  float q[4];
  float pos[3];
  float d[3];       // -(odometry delta): particles move against the ego motion
  float dt;
  int n_in;         // raw (camera) or filtered points supplied
  int ok;           // 0: update rejected (dsp_dynamic.h:183-204)
  int _pad;
};

struct DspDev {
  pop_dsp_config c;
  DspGeo G;
  int nl = 0;
  // persistent particle store
  uint8_t* flags = nullptr;   // [A][V][16]
  float4* pa = nullptr;       // [A][V][S] (vx, vy, vz, weight)
  float4* pb = nullptr;       // [A][V][S] (px, py, pz, -)
  float4* vobj = nullptr;     // [A][V] (occupancy weight, mean vx, vy, vz)
  // tables
  float* pg = nullptr; float* vg = nullptr; int* ur = nullptr; float* pdf = nullptr;
  float* bp_ori = nullptr;    // [(nph+1) + (npv+1)][3]
  // per update
  DspFrame* frame = nullptr;  // [A]
  float* bp = nullptr;        // [A][(nph+1)+(npv+1)][3] rotated plane normals
  float* cam = nullptr;       // [A][MAXIN][3]
  unsigned long long* fkeys = nullptr;  // [A][NPOW2]
  float* pts = nullptr;       // [A][MAXPTS][3] filtered body-frame points
  int* npts = nullptr;        // [A]
  float* rot = nullptr;       // [A][MAXPTS][3]
  int* pobs = nullptr;        // [A][MAXPTS] pyramid*OCAP + slot of each point, -1 if not stored
  float4* obs = nullptr;      // [A][NP][OCAP] (x, y, z, C_k)
  int* obs_num = nullptr;     // [A][NP]
  float* obs_maxlen = nullptr;  // [A][NP]
  int* valid = nullptr;       // [A] points inside the pyramid area
  float* cv = nullptr;        // [A][MAXPTS][7] input_cloud_with_velocity
  int* ncv = nullptr;         // [A]
  float* cl_last = nullptr;   // [A][kDspMaxDyn][5] cx cy cz n intensity
  int* ncl_last = nullptr;    // [A]
  double* cost = nullptr;     // [A][kDspMaxDyn][kDspMaxDyn]
  // re-binning records
  int* nrec = nullptr;        // [A]
  unsigned* r_key = nullptr; int* r_dest = nullptr; int* r_next = nullptr; int* r_pyr = nullptr;
  uint8_t* r_slot = nullptr; uint8_t* r_state = nullptr; float4* r_a = nullptr; float4* r_b = nullptr;  // [A][RCAP]
  int* vhead = nullptr;       // [A][V]
  int* pcount = nullptr; int* poff = nullptr; int* pcur = nullptr;   // [A][NP+1]
  int* plist = nullptr; int* plist2 = nullptr;                       // [A][RCAP]
  int* changed = nullptr;     // [A][kDspMaxIter+2]
  int* pyr_n = nullptr;       // [A][NP]
  int2* pyr_ref = nullptr;    // [A][NP][PCAP] (voxel, slot)
  float4* pyr_p = nullptr;    // [A][NP][PCAP] (px, py, pz, weight)
  // newborn records, index = point*NB + p
  int* b_dest = nullptr; int* b_next = nullptr; float4* b_a = nullptr; float4* b_b = nullptr;  // [A][MAXPTS*NB]
  // voxel work lists (a voxel appears at most once in each): the per-voxel passes never sweep the grid
  int* act[2] = {nullptr, nullptr};   // [A][V] voxels holding particles, ping-pong (written by the occupancy pass)
  int* nact[2] = {nullptr, nullptr};  // [A]
  int* touch = nullptr; int* ntouch = nullptr;     // [A][V] voxels that received a moved particle this update
  int* touch2 = nullptr; int* ntouch2 = nullptr;   // [A][V] voxels that received a newborn particle
  int* ok_dev = nullptr;      // [A] accepted flags of the current update (mask of the peer projection)
  int* vstamp = nullptr;      // [A][V] update id of the last occupancy pass over the voxel (de-duplication)
  int cur = 0;                // which act[] is current
  cudaStream_t side = nullptr;         // clustering/assignment runs beside prediction+update (the reference's side thread)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // depth sensor (k_render.cu): rendered clouds land in `cam`, their sizes in `cam_counts`
  pop_camera camera;
  bool have_camera = false;
  int* depth = nullptr;       // [A][height*width] millimetres
  int* cam_counts = nullptr;  // [A]
  float* rpose = nullptr;     // [A][12]: cam_pose = cam2world^-1 as the splat kernel's float parameters, r (9) then t (3)
  double* rpose_d = nullptr;  // [A][20]: cam2world (16, row-major) + body position (3) + pad, for the back-projection
  bool use_cam_counts = false;  // the next update reads cloud sizes from cam_counts (rendered input)
  int* seq = nullptr;         // [A][4] sequence counters of the three tables
  int* stats = nullptr;       // [A][8]
  int* err = nullptr;         // [1] device-side error flags
  // host-side per-agent odometry memory (static locals of DSPMap::update)
  std::vector<float> last_pos;
  std::vector<double> last_stamp;
  std::vector<char> first;
  std::vector<DspFrame> h_frame;
  int updates = 0;
};

int popk_grid_clear(pop_handle* h, int a0, int na);   // k_map.cu
int popk_dsp_alloc(pop_handle* h, const pop_dsp_config* c);
void popk_dsp_free(pop_handle* h);
// filtered = 0: d->cam holds camera clouds (frame.n_in points) -> filter; 1: d->pts / d->npts already set
int popk_dsp_update(pop_handle* h, int filtered);
int popk_dsp_filter_only(pop_handle* h);
int popk_dsp_download_particles(pop_handle* h, int agent, float* d_out);
int popk_dsp_occupied(pop_handle* h, int agent, float* d_out, int cap, int* d_count);
