// plan_env/map_parameters.h — a WRAPPER around the reference's compile-time map configuration.  TEST INFRASTRUCTURE
// (oracle/_ref only).  The reference fixes the grid in #defines (plan_env/include/plan_env/map_parameters.h:5-13) that
// its users edit per deployment; to pin the oracle at more than one size the map sources are compiled once per size:
// without -DREF_MAP_NX this header forwards to the reference's own file, with it the same names are defined from the
// command line (every derived constant restated from map_parameters.h:18-53 with the reference's expressions).
// FakeParticleRiskVoxel re-#defines PREDICTION_TIMES to 9 after MapBase::risk_maps_[VOXEL_NUM][6] was declared
// (fake_particle_risk_voxel.h:28-29, SURVEY Appendix A Q1: every [i][j >= 6] access aliases the next voxel).  Builds of
// that class pass -DREF_MAP_NT=9 (or the T under test) so the array and the loops agree — the intended semantics.
// fake_particle_risk_voxel.cpp includes this header a second time, after the class header's re-#define: the tail
// below (outside the include guard) then restores the T under test, standing for a user who edits both #defines.
#if !defined(REF_MAP_NX)
#include_next <plan_env/map_parameters.h>
#else
#ifndef DSP_DYNAMIC_PARMETERS_H
#define DSP_DYNAMIC_PARMETERS_H
#define MAP_LENGTH_VOXEL_NUM       REF_MAP_NX
#define MAP_WIDTH_VOXEL_NUM        REF_MAP_NY
#define MAP_HEIGHT_VOXEL_NUM       REF_MAP_NZ
#define VOXEL_RESOLUTION           REF_MAP_RES
#define ANGLE_RESOLUTION           1
#define MAX_PARTICLE_NUM_VOXEL     7
#define LIMIT_MOVEMENT_IN_XY_PLANE 1
#define PREDICTION_TIMES REF_MAP_NT
#ifdef REF_MAP_FUTURE_TIMES
static const float prediction_future_time[PREDICTION_TIMES] = REF_MAP_FUTURE_TIMES;
#endif
const int half_fov_h = 43;
const int half_fov_v = 29;
#define DYNAMIC_CLUSTER_MAX_POINT_NUM 200
#define DYNAMIC_CLUSTER_MAX_CENTER_HEIGHT 1.5
#ifndef GAUSSIAN_RANDOMS_NUM
#define GAUSSIAN_RANDOMS_NUM 10000000
#endif
#define O_MAKE_VALID   1
#define O_MAKE_INVALID 0
static const int observation_pyramid_num_h = (int)half_fov_h * 2 / ANGLE_RESOLUTION;
static const int observation_pyramid_num_v = (int)half_fov_v * 2 / ANGLE_RESOLUTION;
static const int observation_pyramid_num   = observation_pyramid_num_h * observation_pyramid_num_v;
static const int VOXEL_NUM   = MAP_LENGTH_VOXEL_NUM * MAP_WIDTH_VOXEL_NUM * MAP_HEIGHT_VOXEL_NUM;
static const int PYRAMID_NUM = 360 * 180 / ANGLE_RESOLUTION / ANGLE_RESOLUTION;
static const int SAFE_PARTICLE_NUM         = VOXEL_NUM * MAX_PARTICLE_NUM_VOXEL + 1e5;
static const int SAFE_PARTICLE_NUM_VOXEL   = MAX_PARTICLE_NUM_VOXEL * 2;
static const int SAFE_PARTICLE_NUM_PYRAMID = SAFE_PARTICLE_NUM / PYRAMID_NUM * 2;
static const int   observation_max_points_num_one_pyramid = 100;
static const float obstacle_thickness_for_occlusion       = 0.3;
#endif
#endif
#if defined(REF_MAP_NX)
#undef PREDICTION_TIMES
#define PREDICTION_TIMES REF_MAP_NT
#endif
