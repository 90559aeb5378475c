// cpp_host.cpp — the C++ host side (include/pop.hpp) in use, as a plan_manager translation unit would use it.
//   g++ -std=c++17 -Iinclude examples/cpp_host.cpp -o cpp_host pred_occ_planner_b200/libpop.so -Wl,-rpath,$PWD/pred_occ_planner_b200
// Without a CUDA device it checks the host-only parts (defaults, wire format, the no-fallback error) and exits 0;
// with POP_CPP_GPU=1 it also runs BASELINE config 1 (one agent, pillars on a ring) through map update, queries,
// hybrid A*, and the fused replan, printing what it got.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "pop.hpp"

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "REQUIRE failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); return 1; } } while (0)

int main() {
  // defaults are the reference's simulator parameters
  pop_config cfg = pop::Engine::defaultConfig();
  REQUIRE(cfg.abi_version == POP_ABI_VERSION && cfg.nx == 66 && cfg.ny == 66 && cfg.nz == 20 && cfg.nt == 9);
  REQUIRE(std::fabs(cfg.resolution - 0.15f) < 1e-7f && cfg.allocate_num == 10000);
  pop_gridmap_config gc = pop::GridMap::defaultConfig();
  REQUIRE(gc.p_hit == 0.70 && gc.p_occ == 0.80);

  // BezierTraj wire round trip (host only)
  pop_traj_record rec;
  std::memset(&rec, 0, sizeof(rec));
  rec.drone_id = 3; rec.traj_id = 9; rec.n_pieces = 2; rec.valid = 1; rec.start_time = 41.5;
  for (int k = 0; k < 2; k++) rec.duration[k] = 0.3f;
  for (int k = 0; k < 30; k++) rec.cpts[k] = 0.125 * k;
  double pub = 0.0;
  const std::vector<uint8_t> wire = pop::toWire(rec, 41.75);
  const pop_traj_record back = pop::fromWire(wire, &pub);
  REQUIRE(wire.size() == 4 + 4 + 8 + 8 + 1 + 4 + 8 + 4 + 240);
  REQUIRE(back.drone_id == 3 && back.traj_id == 9 && back.n_pieces == 2 && back.start_time == 41.5 && pub == 41.75);
  REQUIRE(std::memcmp(back.cpts, rec.cpts, sizeof(double) * 30) == 0);

  const char* want_gpu = std::getenv("POP_CPP_GPU");
  if (!want_gpu || want_gpu[0] != '1') {
    // no CPU fallback: without a device the engine refuses to exist (on a GPU box this simply succeeds)
    try {
      pop::Engine e(cfg);
      std::printf("cpp_host: device present, engine created\n");
    } catch (const pop::Error& err) {
      REQUIRE(err.code == POP_E_NO_DEVICE);
      std::printf("cpp_host: no device -> %s\n", err.what());
    }
    std::printf("cpp_host: host-only checks ok\n");
    return 0;
  }

  // ---- GPU part: config 1 ----
  cfg.cell_dtype = POP_CELL_U8;
  pop::Engine eng(cfg);
  std::vector<float> cloud;
  std::vector<pop_cylinder> cyl;
  for (int k = 0; k < 12; k++) {                       // pillars on a ring of radius 2.5 m
    const double ang = 2.0 * M_PI * k / 12.0, cx = 2.5 * std::cos(ang), cy = 2.5 * std::sin(ang);
    pop_cylinder c;
    std::memset(&c, 0, sizeof(c));
    c.type = 3; c.x = cx; c.y = cy; c.w = 0.5; c.h = 3.0; c.qw = 1.0;
    cyl.push_back(c);
    for (double z = 0.05; z < 3.0; z += 0.1)
      for (double a = 0; a < 2 * M_PI; a += 0.4) {
        cloud.push_back(static_cast<float>(cx + 0.2 * std::cos(a)));
        cloud.push_back(static_cast<float>(cy + 0.2 * std::sin(a)));
        cloud.push_back(static_cast<float>(z));
      }
  }
  eng.setWorldCloud(cloud);
  eng.setCylinders(cyl);
  const std::vector<double> start = {-4.0, 0.0, 1.0, 0, 0, 0, 0, 0, 0}, goal = {4.0, 0.0, 1.0};
  const std::vector<float> poses = {-4.0f, 0.0f, 1.0f};
  eng.setPoses(poses);
  eng.updateMap(10.0);
  pop::RiskVoxelMap map(eng, 0);
  const int at_pillar = map.getClearOcccupancy(pop::Vec3{-2.5, 0.0, 1.0}, 0);
  const int in_free = map.getClearOcccupancy(pop::Vec3{-3.6, 0.0, 1.0}, 0);
  const int outside = map.getClearOcccupancy(pop::Vec3{40.0, 0.0, 1.0}, 0);
  std::printf("cpp_host: occupancy at a pillar %d, in free space %d, outside the map %d\n", at_pillar, in_free, outside);
  REQUIRE(at_pillar == 1 && in_free == 0 && outside == -1);
  pop::HybridAstar astar(eng);
  const pop::HybridAstar::Result r = astar.search(start, goal, {0.02});
  std::printf("cpp_host: A* ret %d, %d path samples\n", r.ret[0], r.len[0]);
  REQUIRE(r.ret[0] >= POP_ASTAR_REACH_HORIZON && r.len[0] >= 2);
  std::vector<pop_traj_record> traj;
  const std::vector<int32_t> status = eng.replan(start, goal, poses, 10.0, 10.02, &traj);
  std::printf("cpp_host: replan status %d, %d pieces, first control point (%.3f, %.3f, %.3f)\n", status[0], traj[0].n_pieces,
              traj[0].cpts[0], traj[0].cpts[1], traj[0].cpts[2]);
  REQUIRE(status[0] == POP_REPLAN_OK && traj[0].n_pieces >= 2 && std::fabs(traj[0].cpts[0] + 4.0) < 1e-3);
  std::printf("cpp_host: GPU checks ok\n");
  return 0;
}
