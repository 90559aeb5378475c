// o_astar.cpp — oracle: kinodynamic hybrid A* over the SOGM (TEST INFRASTRUCTURE).
// Follows path_searching/src/fake_risk_hybrid_a_star.cpp:85-107 (reset), :114-421 (search),
// :423-457 (estimateHeuristic), :459-519 (computeShotTraj), :521-585 (cubic/quartic),
// :663-694 (getPathWithVel), :795-836 and the call site plan_manager/src/baseline_fake.cpp:264-296.
// path_searching/src/risk_hybrid_a_star.cpp (the planner over RiskBase maps) is the same search except for the
// one-shot collision check, which it runs against slice 0 (:514); selected by pop_config.query_variant.
// Quirks kept on purpose: hash INSERT keys use (int)time while lookups use timeToIndex()
// (:387 vs :269-271); open-set nodes are re-scored in place without re-heapifying (:345-362,
// :401-412); the heuristic is scaled by (1 + tie_breaker) (:456).
// PINNED against the reference's own fake_risk_hybrid_a_star.cpp, compiled unmodified into oracle/_ref/libref.so
// (oracle/Makefile; ROS and the map class replaced by stand-ins, occupancy answered by this oracle's grid):
// return code, expansion / node counts, retry and every sampled path state identical — tests/test_ref_pins.py,
// committed vectors tests/golden/ref_pins.npz.
#include <algorithm>
#include <array>
#include <queue>
#include <unordered_map>

#include "oracle.hpp"

namespace orc {
namespace {

enum { IN_CLOSE_SET = 1, IN_OPEN_SET = 2, NOT_EXPAND = 3 };

struct Node {
  Node* parent = nullptr;
  double input[3] = {0, 0, 0};
  double state[6] = {0, 0, 0, 0, 0, 0};
  double duration = 0, time = 0;
  int time_idx = 0;
  double g = 0, f = 0;
  int node_state = NOT_EXPAND;
  int index[3] = {0, 0, 0};
};
struct NodeCmp {
  bool operator()(const Node* a, const Node* b) const { return a->f > b->f; }
};
struct Key4 {
  int x, y, z, t;
  bool operator==(const Key4& o) const { return x == o.x && y == o.y && z == o.z && t == o.t; }
};
struct Key4Hash {
  size_t operator()(const Key4& k) const {
    size_t seed = 0;
    const int e[4] = {k.x, k.y, k.z, k.t};
    for (int i = 0; i < 4; i++) seed ^= std::hash<int>()(e[i]) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    return seed;
  }
};

struct Search {
  const Grid& g;
  const pop_config& c;
  std::vector<Node> pool;
  std::priority_queue<Node*, std::vector<Node*>, NodeCmp> open;
  std::unordered_map<Key4, Node*, Key4Hash> expanded;
  std::vector<Node*> node_path;
  double center[3];
  double inv_res, inv_tres, time_origin = 0, tie_breaker;
  int use_node_num = 0, iter_num = 0;
  bool is_shot_succ = false;

  Search(const Grid& g_, const pop_config& c_) : g(g_), c(c_), pool(c_.allocate_num) {
    inv_res = 1.0 / c.resolution_astar;
    inv_tres = 1.0 / c.time_resolution_astar;
    tie_breaker = 1.0 + 1.0 / 10000;
  }
  void reset() {
    expanded.clear();
    node_path.clear();
    open = decltype(open)();
    for (int i = 0; i < use_node_num; i++) { pool[i].parent = nullptr; pool[i].node_state = NOT_EXPAND; }
    use_node_num = 0;
    iter_num = 0;
    is_shot_succ = false;
  }
  void pos_to_index(const double* p, int* idx) const {
    for (int i = 0; i < 3; i++) idx[i] = static_cast<int>(std::floor((p[i] - center[i]) * inv_res));
  }
  int time_to_index(double t) const { return static_cast<int>(std::floor((t - time_origin) * inv_tres)); }
  static void transit(const double* s0, double* s1, const double* um, double tau) {
    // state1 = phi*state0 + [0.5*tau^2*um; tau*um]
    const double h = 0.5 * std::pow(tau, 2);
    for (int i = 0; i < 3; i++) {
      s1[i] = (s0[i] + tau * s0[i + 3]) + h * um[i];
      s1[i + 3] = s0[i + 3] + tau * um[i];
    }
  }
  static std::vector<double> cubic(double a, double b, double c_, double d) {
    std::vector<double> dts;
    double a2 = b / a, a1 = c_ / a, a0 = d / a;
    double Q = (3 * a1 - a2 * a2) / 9;
    double R = (9 * a1 * a2 - 27 * a0 - 2 * a2 * a2 * a2) / 54;
    double D = Q * Q * Q + R * R;
    if (D > 0) {
      double S = std::cbrt(R + std::sqrt(D));
      double T = std::cbrt(R - std::sqrt(D));
      dts.push_back(-a2 / 3 + (S + T));
    } else if (D == 0) {
      double S = std::cbrt(R);
      dts.push_back(-a2 / 3 + S + S);
      dts.push_back(-a2 / 3 - S);
    } else {
      double theta = std::acos(R / std::sqrt(-Q * Q * Q));
      dts.push_back(2 * std::sqrt(-Q) * std::cos(theta / 3) - a2 / 3);
      dts.push_back(2 * std::sqrt(-Q) * std::cos((theta + 2 * M_PI) / 3) - a2 / 3);
      dts.push_back(2 * std::sqrt(-Q) * std::cos((theta + 4 * M_PI) / 3) - a2 / 3);
    }
    return dts;
  }
  static std::vector<double> quartic(double a, double b, double c_, double d, double e) {
    std::vector<double> dts;
    double a3 = b / a, a2 = c_ / a, a1 = d / a, a0 = e / a;
    std::vector<double> ys = cubic(1, -a2, a1 * a3 - 4 * a0, 4 * a2 * a0 - a1 * a1 - a3 * a3 * a0);
    double y1 = ys.front();
    double r = a3 * a3 / 4 - a2 + y1;
    if (r < 0) return dts;
    double R = std::sqrt(r), D, E;
    if (R != 0) {
      D = std::sqrt(0.75 * a3 * a3 - R * R - 2 * a2 + 0.25 * (4 * a3 * a2 - 8 * a1 - a3 * a3 * a3) / R);
      E = std::sqrt(0.75 * a3 * a3 - R * R - 2 * a2 - 0.25 * (4 * a3 * a2 - 8 * a1 - a3 * a3 * a3) / R);
    } else {
      D = std::sqrt(0.75 * a3 * a3 - 2 * a2 + 2 * std::sqrt(y1 * y1 - 4 * a0));
      E = std::sqrt(0.75 * a3 * a3 - 2 * a2 - 2 * std::sqrt(y1 * y1 - 4 * a0));
    }
    if (!std::isnan(D)) { dts.push_back(-a3 / 4 + R / 2 + D / 2); dts.push_back(-a3 / 4 + R / 2 - D / 2); }
    if (!std::isnan(E)) { dts.push_back(-a3 / 4 - R / 2 + E / 2); dts.push_back(-a3 / 4 - R / 2 - E / 2); }
    return dts;
  }
  double heuristic(const double* x1, const double* x2, double& optimal_time) const {
    double dp[3], v0[3], v1[3];
    for (int i = 0; i < 3; i++) { dp[i] = x2[i] - x1[i]; v0[i] = x1[i + 3]; v1[i] = x2[i + 3]; }
    auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    double vs[3] = {v0[0] + v1[0], v0[1] + v1[1], v0[2] + v1[2]};
    double c1 = -36 * dot(dp, dp);
    double c2 = 24 * dot(vs, dp);
    double c3 = -4 * (dot(v0, v0) + dot(v0, v1) + dot(v1, v1));
    double c4 = 0, c5 = c.w_time;
    std::vector<double> ts = quartic(c5, c4, c3, c2, c1);
    double v_max = c.max_vel * 0.5;
    double linf = std::max(std::max(std::fabs(x1[0] - x2[0]), std::fabs(x1[1] - x2[1])), std::fabs(x1[2] - x2[2]));
    double t_bar = linf / v_max;
    ts.push_back(t_bar);
    double cost = 100000000, t_d = t_bar;
    for (double t : ts) {
      if (t < t_bar) continue;
      double cc = -c1 / (3 * t * t * t) - c2 / (2 * t * t) - c3 / t + c.w_time * t;
      if (cc < cost) { cost = cc; t_d = t; }
    }
    optimal_time = t_d;
    return 1.0 * (1 + tie_breaker) * cost;
  }
  bool shot(const double* s1, const double* s2, double t_d) {
    double coef[3][4];
    for (int i = 0; i < 3; i++) {
      double p0 = s1[i], dp = s2[i] - p0, v0 = s1[i + 3], v1 = s2[i + 3], dv = v1 - v0;
      coef[i][3] = 1.0 / 6.0 * (-12.0 / (t_d * t_d * t_d) * (dp - v0 * t_d) + 6 / (t_d * t_d) * dv);
      coef[i][2] = 0.5 * (6.0 / (t_d * t_d) * (dp - v0 * t_d) - 2 / t_d * dv);
      coef[i][1] = v0;
      coef[i][0] = p0;
    }
    double t_delta = t_d / 10;
    // the reference loops forever when t_d == 0 (goal reached before the first expansion: `time += 0`) and takes
    // a NaN t_d for a successful shot; both are treated as "no shot" here and in the CUDA path
    if (!(t_delta > 0.0) || !std::isfinite(t_d)) return false;
    for (double time = t_delta; time <= t_d; time += t_delta) {
      double tp[4];
      for (int j = 0; j < 4; j++) tp[j] = std::pow(time, j);
      V3d coord;
      double* cp = &coord.x;
      // Eigen dot of dynamic 4-vectors, SSE2 packets of two: (c0 t0 + c2 t2) + (c1 t1 + c3 t3)
      for (int d = 0; d < 3; d++)
        cp[d] = (coef[d][0] * tp[0] + coef[d][2] * tp[2]) + (coef[d][1] * tp[1] + coef[d][3] * tp[3]);
      // FakeRiskHybridAstar checks the shot at its own time (fake_risk_hybrid_a_star.cpp:521); RiskHybridAstar, the
      // planner of the real-perception path, checks it against the current slice only (risk_hybrid_a_star.cpp:514)
      if (query_dt(g, coord, g.query_variant == POP_QUERY_FAKE ? time : 0.0) != 0) return false;
    }
    is_shot_succ = true;
    return true;
  }
  void retrieve(Node* end) {
    Node* cur = end;
    node_path.push_back(cur);
    while (cur->parent) { cur = cur->parent; node_path.push_back(cur); }
    std::reverse(node_path.begin(), node_path.end());
  }

  int search(const double* start, const double* goal, bool init, double time_start) {
    const double* start_pt = start;
    const double* start_a = start + 6;
    center[0] = static_cast<double>(g.pose.x); center[1] = static_cast<double>(g.pose.y);
    center[2] = static_cast<double>(g.pose.z);
    Node* cur = &pool[0];
    cur->parent = nullptr;
    for (int i = 0; i < 6; i++) cur->state[i] = start[i];
    pos_to_index(start_pt, cur->index);
    cur->g = 0.0;
    double end_state[6] = {goal[0], goal[1], goal[2], 0, 0, 0};
    int end_index[3];
    pos_to_index(goal, end_index);
    double time_to_goal;
    cur->f = c.lambda_heu * heuristic(cur->state, end_state, time_to_goal);
    cur->node_state = IN_OPEN_SET;
    open.push(cur);
    use_node_num += 1;
    time_origin = time_start;
    cur->time = time_start;
    cur->time_idx = time_to_index(time_start);
    expanded.insert({Key4{cur->index[0], cur->index[1], cur->index[2], cur->time_idx}, cur});
    bool init_search = init;

    while (!open.empty()) {
      cur = open.top();
      double dx = cur->state[0] - start_pt[0], dy = cur->state[1] - start_pt[1], dz = cur->state[2] - start_pt[2];
      bool reach_horizon = std::sqrt(dx * dx + dy * dy + dz * dz) >= c.horizon;
      bool near_end = std::abs(cur->index[0] - end_index[0]) <= c.tolerance &&
                      std::abs(cur->index[1] - end_index[1]) <= c.tolerance &&
                      std::abs(cur->index[2] - end_index[2]) <= c.tolerance;
      bool exceed_time = cur->time >= c.max_tau;
      if (reach_horizon || near_end || exceed_time) {
        retrieve(cur);
        if (near_end) {
          heuristic(cur->state, end_state, time_to_goal);
          shot(cur->state, end_state, time_to_goal);
        }
      }
      if (reach_horizon) return is_shot_succ ? POP_ASTAR_REACH_END : POP_ASTAR_REACH_HORIZON;
      if (near_end) {
        if (is_shot_succ) return POP_ASTAR_REACH_END;
        if (cur->parent) return POP_ASTAR_NEAR_END;
        return POP_ASTAR_NO_PATH;
      }
      if (exceed_time) return POP_ASTAR_REACH_HORIZON;

      open.pop();
      cur->node_state = IN_CLOSE_SET;
      iter_num += 1;

      const double res = 1 / 2.0;
      double cur_state[6];
      for (int i = 0; i < 6; i++) cur_state[i] = cur->state[i];
      std::vector<Node*> tmp_expand;
      std::vector<std::array<double, 3>> inputs;
      if (init_search) {
        inputs.push_back({start_a[0], start_a[1], start_a[2]});
        init_search = false;
      } else {
        for (double ax = -c.max_acc; ax <= c.max_acc + 1e-3; ax += c.max_acc * res)
          for (double ay = -c.max_acc; ay <= c.max_acc + 1e-3; ay += c.max_acc * res)
            for (double az = -0.5 * c.max_acc; az <= 0.5 * c.max_acc + 1e-3; az += c.max_acc * res)
              inputs.push_back({ax, ay, az});
      }
      const double tau = c.time_resolution_astar;
      for (size_t i = 0; i < inputs.size(); ++i) {
        const double* um = inputs[i].data();
        double pro_state[6];
        transit(cur_state, pro_state, um, tau);
        double pro_t = cur->time + tau;
        int pro_id[3];
        pos_to_index(pro_state, pro_id);
        int pro_t_id = time_to_index(pro_t);
        Node* pro_node = nullptr;
        auto it = expanded.find(Key4{pro_id[0], pro_id[1], pro_id[2], pro_t_id});
        if (it != expanded.end()) pro_node = it->second;
        if (pro_node && pro_node->node_state == IN_CLOSE_SET) continue;
        if (std::fabs(pro_state[3]) > c.max_vel || std::fabs(pro_state[4]) > c.max_vel ||
            std::fabs(pro_state[5]) > c.max_vel)
          continue;
        int diff_time = pro_t_id - cur->time_idx;
        if (pro_id[0] == cur->index[0] && pro_id[1] == cur->index[1] && pro_id[2] == cur->index[2] &&
            diff_time == 0)
          continue;
        bool is_occ = false;
        for (int k = 1; k <= c.check_num; ++k) {
          double dt = tau * static_cast<double>(k) / static_cast<double>(c.check_num);
          double xt[6];
          transit(cur_state, xt, um, dt);
          double t = cur->time + dt;
          if (query_dt(g, V3d{xt[0], xt[1], xt[2]}, t) != 0) { is_occ = true; break; }
        }
        if (is_occ) continue;
        double ttg;
        double tmp_g = (um[0] * um[0] + um[1] * um[1] + um[2] * um[2] + c.w_time) * tau + cur->g;
        double tmp_f = tmp_g + c.lambda_heu * heuristic(pro_state, end_state, ttg);
        bool prune = false;
        for (Node* en : tmp_expand) {
          if (pro_id[0] == en->index[0] && pro_id[1] == en->index[1] && pro_id[2] == en->index[2] &&
              pro_t_id == en->time_idx) {
            prune = true;
            if (tmp_f < en->f) {
              en->f = tmp_f;
              en->g = tmp_g;
              for (int q = 0; q < 6; q++) en->state[q] = pro_state[q];
              for (int q = 0; q < 3; q++) en->input[q] = um[q];
              en->duration = tau;
              en->time = cur->time + tau;
            }
            break;
          }
        }
        if (prune) continue;
        if (!pro_node) {
          pro_node = &pool[use_node_num];
          for (int q = 0; q < 3; q++) pro_node->index[q] = pro_id[q];
          for (int q = 0; q < 6; q++) pro_node->state[q] = pro_state[q];
          pro_node->f = tmp_f;
          pro_node->g = tmp_g;
          for (int q = 0; q < 3; q++) pro_node->input[q] = um[q];
          pro_node->duration = tau;
          pro_node->parent = cur;
          pro_node->node_state = IN_OPEN_SET;
          pro_node->time = cur->time + tau;
          pro_node->time_idx = time_to_index(pro_node->time);
          open.push(pro_node);
          // insert(pro_id, pro_node->time, pro_node): the double is truncated to int by the call
          expanded.insert({Key4{pro_id[0], pro_id[1], pro_id[2], static_cast<int>(pro_node->time)}, pro_node});
          tmp_expand.push_back(pro_node);
          use_node_num += 1;
          if (use_node_num == c.allocate_num) return POP_ASTAR_NO_PATH;
        } else if (pro_node->node_state == IN_OPEN_SET) {
          if (tmp_g < pro_node->g) {
            for (int q = 0; q < 6; q++) pro_node->state[q] = pro_state[q];
            pro_node->f = tmp_f;
            pro_node->g = tmp_g;
            for (int q = 0; q < 3; q++) pro_node->input[q] = um[q];
            pro_node->duration = tau;
            pro_node->parent = cur;
            pro_node->time = cur->time + tau;
          }
        } else {
          return POP_ASTAR_SEARCH_ERR;
        }
      }
    }
    return POP_ASTAR_NO_PATH;
  }

  // getPathWithVel (:663-694)
  std::vector<double> path_with_vel(double delta_t) const {
    std::vector<std::array<double, 6>> list;
    Node* node = node_path.back();
    double t_node = 0, t_sample = delta_t;
    std::array<double, 6> s;
    for (int i = 0; i < 6; i++) s[i] = node->state[i];
    list.push_back(s);
    while (node->parent) {
      const double* ut = node->input;
      double duration = node->duration;
      const double* x0 = node->parent->state;
      t_node = duration;
      while (true) {
        if (t_sample > t_node) {
          node = node->parent;
          t_sample -= t_node;
          break;
        }
        t_node -= t_sample;
        transit(x0, s.data(), ut, t_node);
        list.push_back(s);
        t_sample = delta_t;
      }
    }
    std::reverse(list.begin(), list.end());
    std::vector<double> out;
    for (auto& e : list) out.insert(out.end(), e.begin(), e.end());
    return out;
  }
};

}  // namespace

// baseline_fake.cpp:264-296: reset, search(init=true), retry with init=false when NO_PATH
AstarResult astar_replan(const Grid& g, const pop_config& c, const double start[9],
                         const double goal[3], double t_start, int /*math_mode*/) {
  AstarResult r;
  Search s(g, c);
  s.reset();
  int rst = s.search(start, goal, true, t_start);
  if (rst == 0) {
    s.reset();
    rst = s.search(start, goal, false, t_start);
    r.retried = 1;
  }
  r.ret = rst;
  r.iters = s.iter_num;
  r.nodes = s.use_node_num;
  if (rst != POP_ASTAR_NO_PATH && !s.node_path.empty()) r.path = s.path_with_vel(c.corridor_tau);
  return r;
}

}  // namespace orc
