// cpp_host.cpp — the C++ host side (include/pop.hpp) in use, as a plan_manager translation unit would use it.
//   g++ -std=c++17 -Iinclude examples/cpp_host.cpp -o cpp_host pred_occ_planner_b200/libpop.so -Wl,-rpath,$PWD/pred_occ_planner_b200
// Without a CUDA device it checks the host-only parts (defaults, wire format, the no-fallback error) and exits 0;
// with POP_CPP_GPU=1 it also runs BASELINE config 1 (one agent, pillars on a ring) through map update, queries,
// hybrid A*, and the fused replan, printing what it got; with POP_CPP_GPU=2 (two GPUs) it then flies a four-agent swarm
// as two ranks, one host thread and one engine per GPU, exchanging the trajectories through the engines' own NCCL
// communicator (add -pthread to the command above).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <future>
#include <thread>

#include "pop.hpp"

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "REQUIRE failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); return 1; } } while (0)

// One rank of a two-GPU swarm: agents [2 rank, 2 rank + 2) of four flying across a circle; three cycles, an exchange
// after each.  Returns 0 when, after every exchange, the table holds what both ranks planned in that cycle.
static int swarm_rank(int rank, std::shared_future<std::vector<uint8_t>> id, std::promise<std::vector<uint8_t>>* make_id,
                      std::vector<pop_traj_record>* last_table) {
  const int world = 2, n_local = 2, n_total = 4;
  pop_config cfg = pop::Engine::defaultConfig();
  cfg.cell_dtype = POP_CELL_U8;
  cfg.n_agents_local = n_local; cfg.n_agents_total = n_total; cfg.agent_id_base = rank * n_local;
  pop::Engine eng(cfg, /*device*/ rank);
  if (make_id) make_id->set_value(pop::Engine::commUniqueId());   // rank 0 creates the id, the other rank waits for it
  eng.commInit(id.get(), rank, world);
  std::vector<double> start, goal;
  std::vector<float> poses;
  for (int a = rank * n_local; a < (rank + 1) * n_local; a++) {
    const double ang = 2.0 * M_PI * a / n_total, x = 3.5 * std::cos(ang), y = 3.5 * std::sin(ang);
    const double s9[9] = {x, y, 1.0, 0, 0, 0, 0, 0, 0};
    start.insert(start.end(), s9, s9 + 9);
    goal.insert(goal.end(), {-x, -y, 1.0});
    poses.insert(poses.end(), {static_cast<float>(x), static_cast<float>(y), 1.0f});
  }
  eng.setWorldCloud({});
  eng.setCylinders({});
  std::vector<double> offs(static_cast<size_t>(n_total) * POP_MAX_BODY * 3, 0.0);
  std::vector<int32_t> counts(n_total, 1);                        // one body particle at the centre of every drone
  eng.setBodyParticles(offs, counts);
  eng.setTrajectories(std::vector<pop_traj_record>(n_total));
  for (int k = 0; k < 3; k++) {
    std::vector<pop_traj_record> mine;
    const std::vector<int32_t> st = eng.replan(start, goal, poses, 10.0 + 0.1 * k, 10.02 + 0.1 * k, &mine);
    eng.exchange();
    const std::vector<pop_traj_record> table = eng.trajectories();
    for (int a = 0; a < n_local; a++) {
      if (st[a] != POP_REPLAN_OK) return 10 + k;
      if (std::memcmp(&table[rank * n_local + a], &mine[a], sizeof(pop_traj_record)) != 0) return 20 + k;
    }
    for (int a = 0; a < n_total; a++)
      if (table[a].drone_id != a || table[a].n_pieces <= 0) return 30 + k;   // the other rank's records arrived too
    *last_table = table;
  }
  return 0;
}

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
  if (!want_gpu || (want_gpu[0] != '1' && want_gpu[0] != '2')) {
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
  if (want_gpu[0] == '1') return 0;

  // ---- two GPUs, two host threads, one engine each ----
  std::promise<std::vector<uint8_t>> id_promise;
  std::shared_future<std::vector<uint8_t>> id = id_promise.get_future().share();
  std::vector<pop_traj_record> tab[2];
  int rc[2] = {-1, -1};
  std::thread t1([&] { try { rc[1] = swarm_rank(1, id, nullptr, &tab[1]); } catch (const pop::Error& e) { std::fprintf(stderr, "rank 1: %s\n", e.what()); } });
  try { rc[0] = swarm_rank(0, id, &id_promise, &tab[0]); } catch (const pop::Error& e) { std::fprintf(stderr, "rank 0: %s\n", e.what()); }
  t1.join();
  std::printf("cpp_host: two ranks finished with %d / %d\n", rc[0], rc[1]);
  REQUIRE(rc[0] == 0 && rc[1] == 0);
  REQUIRE(tab[0].size() == 4 && std::memcmp(tab[0].data(), tab[1].data(), 4 * sizeof(pop_traj_record)) == 0);   // same table on both
  std::printf("cpp_host: two-rank exchange ok\n");
  return 0;
}
