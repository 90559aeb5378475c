// k_corridor.cu — K3: safe-flight-corridor stage, batched over (agent, path segment).
//
//   k_corridor_points : per segment, the local box and the SOGM cells above threshold inside it
//                       (MapBase::getObstaclePoints, plan_env/src/map.cpp:463-531), compacted in
//                       the reference's emission order (z, y, x, slice).
//   k_corridor_firi   : one warp per segment: firi::firi with 2 iterations
//                       (plan_manager/include/sfc_gen/firi.hpp:238-365), the MVIE step
//                       (:146-236: Seidel LP for the deepest point, 9-variable L-BFGS with
//                       Lewis-Overton line search, plan_manager/include/sfc_gen/lbfgs.hpp),
//                       ShrinkCorridor(hPoly, path) and checkCorridorValidity
//                       (plan_manager/src/baseline_fake.cpp:191-228).
//   k_corridor_finish : per agent: first-invalid cut, adjacent-intersection LPs, goal
//                       reachability / projection (baseline_fake.cpp:143-182, 362-411).
// The point loops are lane-parallel with first-index argmin reductions so face selection is
// identical to the sequential reference order.
#include <cfloat>

#include "pop_geom.cuh"
#include "pop_lp.cuh"
#include "pop_corridor_finish.cuh"

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kLpRows = 2 * POP_MAX_FACES + 8;  // largest LP: two corridors concatenated

struct SegInfo {       // one path segment of one agent
  double a[3], b[3];   // wpts[i], wpts[i+1]
  double lhc[3], llc[3];
  int n_pts;           // obstacle points emitted (may exceed cap → overflow)
  int active;          // segment exists (i < len-1)
  int j0, j1;
  int lx, ly, lz, hx, hy, hz;
};

__device__ __forceinline__ double dot3(const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
__device__ __forceinline__ double norm3(const double* a) { return sqrt(dot3(a, a)); }

// =================================================================================================
// obstacle points
// =================================================================================================
template <typename C>
__device__ __forceinline__ int seg_count_or_emit(const PopGeom& g, const uint8_t* ag, const SegInfo& s, int v,
                                                 int* out, int base, int cap) {
  const int bx = s.hx - s.lx + 1, by = s.hy - s.ly + 1;
  const int x = s.lx + v % bx, y = s.ly + (v / bx) % by, z = s.lz + v / (bx * by);
  const size_t i = static_cast<size_t>(x) + static_cast<size_t>(y) * g.nx + static_cast<size_t>(z) * g.nx * g.ny;
  // The reference emits the voxel once per slice above threshold (duplicates adjacent in its
  // order).  Duplicates cannot change FIRI's output — an identical point is eliminated by the face
  // its first copy generates — so each voxel is emitted once.
  for (int j = s.j0; j <= s.j1; j++) {
    const float thr = g.query_variant == POP_QUERY_FAKE
                          ? g.risk_threshold
                          : __fsub_rn(g.risk_threshold, __fmul_rn(g.decay_voxel, static_cast<float>(j)));
    if (Cell<C>::load(ag, static_cast<size_t>(j) * g.V + i) > thr) {
      if (out && base < cap) out[base] = static_cast<int>(i);
      return 1;
    }
  }
  return 0;
}

template <typename C>
__global__ void __launch_bounds__(256) k_corridor_points(PopGeom g, const uint8_t* __restrict__ grid,
                                                         const float* __restrict__ poses, double t_map,
                                                         const double* __restrict__ path,
                                                         const int* __restrict__ len,
                                                         const double* __restrict__ start,
                                                         const double* __restrict__ tst, double tau,
                                                         double init_range, SegInfo* __restrict__ segs,
                                                         int* __restrict__ pts, int cap, int seg0) {
  __shared__ SegInfo s;
  __shared__ int warp_sums[8];
  __shared__ int total_s;
  const int bid = blockIdx.x + seg0;   // global segment index
  const int a = bid / POP_MAX_PIECES, i = bid % POP_MAX_PIECES;
  const int n = len[a];
  const float* pose = poses + 3 * a;
  if (threadIdx.x == 0) {
    s.active = (n >= 2 && i < n - 1) ? 1 : 0;
    s.n_pts = 0;
    if (s.active) {
      const double* w0 = path + (static_cast<size_t>(a) * POP_MAX_PATH + i) * 6;
      const double* w1 = w0 + 6;
      for (int k = 0; k < 3; k++) { s.a[k] = w0[k]; s.b[k] = w1[k]; }
      if (s.a[2] < 0) s.a[2] = 0.1;
      if (s.b[2] < 0) s.b[2] = 0.1;
      const double* sp = start + a * 9;
      double lower[3] = {-4 + sp[0], -4 + sp[1], -1 + sp[2]}, higher[3] = {4 + sp[0], 4 + sp[1], 1 + sp[2]};
      if (lower[2] < 0) lower[2] = 0;
      if (higher[2] > 4) higher[2] = 4;
      for (int k = 0; k < 3; k++) {
        s.lhc[k] = fmin(fmax(s.a[k], s.b[k]) + init_range, higher[k]);
        s.llc[k] = fmax(fmin(s.a[k], s.b[k]) - init_range, lower[k]);
      }
      const double t1 = tst[a] + i * tau, t2 = tst[a] + (i + 1) * tau;
      int i0 = static_cast<int>(floor((t1 - t_map) / g.dt));
      int i1 = static_cast<int>(ceil((t2 - t_map) / g.dt));
      i0 = i0 < 0 ? 0 : i0;
      i0 = i0 > g.nt - 1 ? g.nt - 1 : i0;
      i1 = i1 > g.nt - 1 ? g.nt - 1 : i1;
      i1 = i1 < 0 ? 0 : i1;
      s.j0 = i0; s.j1 = i1;
      s.lx = static_cast<int>((s.llc[0] - pose[0] + g.rx) / g.res);
      s.ly = static_cast<int>((s.llc[1] - pose[1] + g.ry) / g.res);
      s.lz = static_cast<int>((s.llc[2] - pose[2] + g.rz) / g.res);
      s.hx = static_cast<int>((s.lhc[0] - pose[0] + g.rx) / g.res);
      s.hy = static_cast<int>((s.lhc[1] - pose[1] + g.ry) / g.res);
      s.hz = static_cast<int>((s.lhc[2] - pose[2] + g.rz) / g.res);
      s.hx = min(s.hx, g.nx - 1); s.hy = min(s.hy, g.ny - 1); s.hz = min(s.hz, g.nz - 1);
      s.lx = max(s.lx, 0); s.ly = max(s.ly, 0); s.lz = max(s.lz, 0);
    }
  }
  __syncthreads();
  if (!s.active) {
    if (threadIdx.x == 0) segs[bid] = s;
    return;
  }
  const uint8_t* ag = grid + a * g.agent_stride;
  int* out = pts + static_cast<size_t>(bid) * cap;
  const int bx = s.hx - s.lx + 1, by = s.hy - s.ly + 1, bz = s.hz - s.lz + 1;
  const int nvox = (bx > 0 && by > 0 && bz > 0) ? bx * by * bz : 0;
  const int per = (nvox + blockDim.x - 1) / blockDim.x;
  const int v0 = threadIdx.x * per, v1 = min(nvox, v0 + per);
  int cnt = 0;
  for (int v = v0; v < v1; v++) cnt += seg_count_or_emit<C>(g, ag, s, v, nullptr, 0, 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    const int nbr = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += nbr;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 0; w < 8; w++) { const int t = warp_sums[w]; warp_sums[w] = run; run += t; }
    total_s = run;
  }
  __syncthreads();
  int base = warp_sums[warp] + incl - cnt;
  for (int v = v0; v < v1; v++) base += seg_count_or_emit<C>(g, ag, s, v, out, base, cap);
  if (threadIdx.x == 0) {
    s.n_pts = total_s;
    segs[bid] = s;
  }
}

// =================================================================================================
// MVIE: cost, L-BFGS (n = 9), ellipsoid update — single-thread routines run by lane 0
// =================================================================================================
__device__ __forceinline__ bool smoothed_l1(double mu, double x, double& f, double& df) {
  if (x < 0.0) return false;
  if (x > mu) { f = x - 0.5 * mu; df = 1.0; return true; }
  const double xdmu = x / mu, sqr = xdmu * xdmu, mumxd2 = mu - 0.5 * x;
  f = mumxd2 * sqr * xdmu;
  df = sqr * ((-0.5) * xdmu + 3.0 * mumxd2 / mu);
  return true;
}

// costMVIE (firi.hpp:73-142); A = M x 3 row-major.  Warp-cooperative: every lane evaluates the faces
// lane, lane+32 (square root, divisions, smoothed penalty — the expensive part), leaves the face's
// ten contributions in shared memory, and lanes 0..9 then add one quantity each over the faces IN
// FACE ORDER, so the sums round exactly like the sequential loop (an inactive face contributes +0.0,
// which never changes a sum that started at +0.0).  All lanes hold the same x and receive the same
// cost and gradient.
constexpr int kTerms = 10;
__device__ double cost_mvie(const double* A, int M, double smooth_eps, double penalty, const double* x, double* grad,
                            double (*terms)[kTerms]) {
  const int lane = threadIdx.x & 31;
  const double* p = x; const double* rtd = x + 3; const double* cde = x + 6;
  const double L00 = rtd[0] * rtd[0] + DBL_EPSILON, L10 = cde[0], L11 = rtd[1] * rtd[1] + DBL_EPSILON,
               L20 = cde[2], L21 = cde[1], L22 = rtd[2] * rtd[2] + DBL_EPSILON;
  for (int i = lane; i < M; i += 32) {
    const double* a = A + i * 3;
    const double al0 = (a[0] * L00 + a[1] * L10) + a[2] * L20;
    const double al1 = (a[0] * 0.0 + a[1] * L11) + a[2] * L21;
    const double al2 = (a[0] * 0.0 + a[1] * 0.0) + a[2] * L22;
    const double nrm = sqrt((al0 * al0 + al1 * al1) + al2 * al2);
    const double adj0 = al0 / nrm, adj1 = al1 / nrm, adj2 = al2 / nrm;
    const double viol = (nrm + ((a[0] * p[0] + a[1] * p[1]) + a[2] * p[2])) - 1.0;
    double c, dc;
    double* t = terms[i];
    if (smoothed_l1(smooth_eps, viol, c, dc)) {
      const double v0 = dc * a[0], v1 = dc * a[1], v2 = dc * a[2];
      t[0] = c; t[1] = v0; t[2] = v1; t[3] = v2;
      t[4] = adj0 * v0; t[5] = adj1 * v1; t[6] = adj2 * v2;
      t[7] = adj0 * v1; t[8] = adj1 * v2; t[9] = adj0 * v2;
    } else {
#pragma unroll
      for (int q = 0; q < kTerms; q++) t[q] = 0.0;
    }
  }
  __syncwarp();
  double acc = 0;
  if (lane < kTerms)
    for (int i = 0; i < M; ++i) acc += terms[i][lane];
  __syncwarp();
  double cost = __shfl_sync(0xffffffffu, acc, 0);
  double gdp[3], gdrtd[3], gdcde[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    gdp[k] = __shfl_sync(0xffffffffu, acc, 1 + k);
    gdrtd[k] = __shfl_sync(0xffffffffu, acc, 4 + k);
    gdcde[k] = __shfl_sync(0xffffffffu, acc, 7 + k);
  }
  cost *= penalty;
  for (int k = 0; k < 3; k++) { gdp[k] *= penalty; gdrtd[k] *= penalty; gdcde[k] *= penalty; }
  cost -= log(L00) + log(L11) + log(L22);
  gdrtd[0] -= 1.0 / L00; gdrtd[1] -= 1.0 / L11; gdrtd[2] -= 1.0 / L22;
  gdrtd[0] *= 2.0 * rtd[0]; gdrtd[1] *= 2.0 * rtd[1]; gdrtd[2] *= 2.0 * rtd[2];
  for (int k = 0; k < 3; k++) { grad[k] = gdp[k]; grad[3 + k] = gdrtd[k]; grad[6 + k] = gdcde[k]; }
  return cost;
}

constexpr int N9 = 9, kMem = 18;
__device__ __forceinline__ double dot9(const double* a, const double* b) {
  double s = 0;
  for (int i = 0; i < N9; i++) s += a[i] * b[i];
  return s;
}
__device__ __forceinline__ double maxabs9(const double* a) {
  double m = 0;
  for (int i = 0; i < N9; i++) m = fmax(m, fabs(a[i]));
  return m;
}

// line_search_lewisoverton (lbfgs.hpp:96-190)
// Run by ALL lanes of the warp with identical arguments (uniform control flow).
__device__ int line_search(const double* A, int M, double* x, double& f, double* g, double& stp, const double* s,
                           const double* xp, const double* gp, double stpmin, double stpmax,
                           double (*terms)[kTerms]) {
  const double f_dec = 1.0e-4, s_curv = 0.9, machine_prec = 1.0e-16;
  const int max_ls = 64;
  int count = 0;
  bool brackt = false, touched = false;
  double mu = 0.0, nu = stpmax;
  if (!(stp > 0.0)) return -1;
  const double dginit = dot9(gp, s);
  if (0.0 < dginit) return -2;
  const double finit = f, dgtest = f_dec * dginit, dstest = s_curv * dginit;
  while (true) {
    for (int i = 0; i < N9; i++) x[i] = xp[i] + stp * s[i];
    f = cost_mvie(A, M, 1.0e-2, 1.0e+3, x, g, terms);
    ++count;
    if (isinf(f) || isnan(f)) return -3;
    if (f > finit + stp * dgtest) {
      nu = stp;
      brackt = true;
    } else {
      if (dot9(g, s) < dstest) mu = stp;
      else return count;
    }
    if (max_ls <= count) return -4;
    if (brackt && (nu - mu) < machine_prec * nu) return -5;
    if (brackt) stp = 0.5 * (mu + nu);
    else stp *= 2.0;
    if (stp < stpmin) return -6;
    if (stp > stpmax) {
      if (touched) return -7;
      touched = true;
      stp = stpmax;
    }
  }
}

// lbfgs_optimize (lbfgs.hpp:192-430) with mem_size 18, past 3, delta 1e-7, g_epsilon 0,
// min_step 1e-32 (firi.hpp:191-199); lm = 2*9*18 + 2*18 doubles of per-warp shared memory.
// Run by ALL lanes of the warp with identical arguments: the scalar algorithm is replicated, only
// the cost evaluation is split over the lanes.
__device__ int lbfgs9(const double* A, int M, double* x, double* lm, double (*terms)[kTerms]) {
  const int m = kMem, past = 3;
  const double delta = 1.0e-7, g_epsilon = 0.0, min_step = 1.0e-32, max_step = 1.0e+20, cautious = 1.0e-6;
  double* lm_s = lm;
  double* lm_y = lm + N9 * kMem;
  double* lm_ys = lm + 2 * N9 * kMem;
  double* lm_alpha = lm_ys + kMem;
  double xp[N9], g[N9], gp[N9], d[N9], pf[3];
  int ret = 0, k = 1, end = 0, bound = 0;
  double fx = cost_mvie(A, M, 1.0e-2, 1.0e+3, x, g, terms);
  pf[0] = fx;
  for (int i = 0; i < N9; i++) d[i] = -g[i];
  double gnorm = maxabs9(g), xnorm = maxabs9(x);
  if (gnorm / fmax(1.0, xnorm) < g_epsilon) return 0;
  double step = 1.0 / sqrt(dot9(d, d));
  while (true) {
    for (int i = 0; i < N9; i++) { xp[i] = x[i]; gp[i] = g[i]; }
    const int ls = line_search(A, M, x, fx, g, step, d, xp, gp, min_step, max_step, terms);
    if (ls < 0) {
      for (int i = 0; i < N9; i++) { x[i] = xp[i]; g[i] = gp[i]; }
      ret = ls;
      break;
    }
    gnorm = maxabs9(g); xnorm = maxabs9(x);
    if (gnorm / fmax(1.0, xnorm) < g_epsilon) { ret = 0; break; }
    if (past <= k) {
      const double rate = fabs(pf[k % past] - fx) / fmax(1.0, fabs(fx));
      if (rate < delta) { ret = 1; break; }
    }
    pf[k % past] = fx;
    ++k;
    double* sc = lm_s + end * N9;
    double* yc = lm_y + end * N9;
    for (int i = 0; i < N9; i++) { sc[i] = x[i] - xp[i]; yc[i] = g[i] - gp[i]; }
    __syncwarp();
    const double ys = dot9(yc, sc), yy = dot9(yc, yc);
    lm_ys[end] = ys;
    for (int i = 0; i < N9; i++) d[i] = -g[i];
    const double cau = dot9(sc, sc) * sqrt(dot9(gp, gp)) * cautious;
    if (ys > cau) {
      ++bound;
      bound = m < bound ? m : bound;
      end = (end + 1) % m;
      int j = end;
      for (int i = 0; i < bound; ++i) {
        j = (j + m - 1) % m;
        lm_alpha[j] = dot9(lm_s + j * N9, d) / lm_ys[j];
        for (int q = 0; q < N9; q++) d[q] += (-lm_alpha[j]) * lm_y[j * N9 + q];
      }
      for (int q = 0; q < N9; q++) d[q] *= ys / yy;
      for (int i = 0; i < bound; ++i) {
        const double beta = dot9(lm_y + j * N9, d) / lm_ys[j];
        for (int q = 0; q < N9; q++) d[q] += (lm_alpha[j] - beta) * lm_s[j * N9 + q];
        j = (j + 1) % m;
      }
    }
    step = 1.0;
  }
  return ret;
}

__device__ void chol3(const double* A, double* L) {  // firi.hpp:44-55
  L[0] = sqrt(A[0]); L[1] = 0; L[2] = 0;
  L[3] = 0.5 * (A[1] + A[3]) / L[0];
  L[4] = sqrt(A[4] - L[3] * L[3]); L[5] = 0;
  L[6] = 0.5 * (A[2] + A[6]) / L[0];
  L[7] = (0.5 * (A[5] + A[7]) - L[6] * L[3]) / L[4];
  L[8] = sqrt(A[8] - L[6] * L[6] - L[7] * L[7]);
}

// U and singular values of a 3x3 matrix from a cyclic Jacobi eigen-solve of L L^T (descending)
__device__ void svd3_us(const double* Lm, double* U, double* S) {
  double B[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += Lm[i * 3 + k] * Lm[j * 3 + k];
      B[i * 3 + j] = s;
    }
  double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int sweep = 0; sweep < 60; sweep++) {
    const double off = fabs(B[1]) + fabs(B[2]) + fabs(B[5]);
    const double diag = fabs(B[0]) + fabs(B[4]) + fabs(B[8]);
    if (off <= 1e-300 || off <= 1e-17 * diag) break;
    for (int e = 0; e < 3; e++) {
      const int p = e == 2 ? 1 : 0, q = e == 0 ? 1 : 2;
      const double apq = B[p * 3 + q];
      if (apq == 0.0) continue;
      const double theta = (B[q * 3 + q] - B[p * 3 + p]) / (2.0 * apq);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; k++) {
        const double bkp = B[k * 3 + p], bkq = B[k * 3 + q];
        B[k * 3 + p] = c * bkp - s * bkq;
        B[k * 3 + q] = s * bkp + c * bkq;
      }
      for (int k = 0; k < 3; k++) {
        const double bpk = B[p * 3 + k], bqk = B[q * 3 + k];
        B[p * 3 + k] = c * bpk - s * bqk;
        B[q * 3 + k] = s * bpk + c * bqk;
      }
      for (int k = 0; k < 3; k++) {
        const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
        V[k * 3 + p] = c * vkp - s * vkq;
        V[k * 3 + q] = s * vkp + c * vkq;
      }
    }
  }
  int order[3] = {0, 1, 2};
  const double ev[3] = {B[0], B[4], B[8]};
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2 - i; j++)
      if (ev[order[j]] < ev[order[j + 1]]) { const int t = order[j]; order[j] = order[j + 1]; order[j + 1] = t; }
  for (int j = 0; j < 3; j++) {
    S[j] = sqrt(fmax(ev[order[j]], 0.0));
    for (int k = 0; k < 3; k++) U[k * 3 + j] = V[k * 3 + order[j]];
  }
}

__device__ double det3(const double* M) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// scratch layout per segment (doubles): [Alp 4*F][blp F][A3 3*F][lm 2*9*18+36][lp work]
constexpr int kScrAlp = 0;
constexpr int kScrBlp = kScrAlp + 4 * POP_MAX_FACES;
constexpr int kScrA3 = kScrBlp + POP_MAX_FACES;
constexpr int kScrLm = kScrA3 + 3 * POP_MAX_FACES;
constexpr int kScrLp = kScrLm + 2 * N9 * kMem + 2 * kMem;
constexpr int kScrDoubles = kScrLp + poplp::work_doubles(kLpRows, 4);
constexpr int kScrInts = poplp::work_ints(kLpRows);

// maxVolInsEllipsoid (firi.hpp:146-236); hPoly = M x 4.  Three steps: mvie_prepare (lane 0: deepest
// point by LP, normalised faces A3, Cholesky start point), lbfgs9 (whole warp), mvie_finish (lane 0).
__device__ bool mvie_prepare(const double* hPoly, int M, const double* R, const double* p, const double* r,
                             unsigned long long* rng, double* scr, int* scri, double* x, double* interior) {
  double* Alp = scr + kScrAlp;
  double* blp = scr + kScrBlp;
  double* A3 = scr + kScrA3;
  for (int i = 0; i < M; i++) {
    const double* h = hPoly + i * 4;
    const double hn = sqrt((h[0] * h[0] + h[1] * h[1]) + h[2] * h[2]);
    for (int k = 0; k < 3; k++) Alp[i * 4 + k] = h[k] / hn;
    Alp[i * 4 + 3] = 1.0;
    blp[i] = -h[3] / hn;
  }
  const double clp[4] = {0, 0, 0, -1.0};
  double xlp[4];
  const double maxdepth = -poplp::linprog<4>(clp, Alp, 4, blp, 1, M, xlp, rng, scr + kScrLp, scri);
  if (!(maxdepth > 0.0) || isinf(maxdepth)) return false;
  for (int k = 0; k < 3; k++) interior[k] = xlp[k];
  for (int i = 0; i < M; i++) {
    const double* a = Alp + i * 4;
    const double den = blp[i] - ((a[0] * interior[0] + a[1] * interior[1]) + a[2] * interior[2]);
    for (int k = 0; k < 3; k++) A3[i * 3 + k] = a[k] / den;
  }
  double RD[9], Q[9], L[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) RD[i * 3 + j] = R[i * 3 + j] * (r[j] * r[j]);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += RD[i * 3 + k] * R[j * 3 + k];
      Q[i * 3 + j] = s;
    }
  chol3(Q, L);
  for (int k = 0; k < 3; k++) x[k] = p[k] - interior[k];
  x[3] = sqrt(L[0]); x[4] = sqrt(L[4]); x[5] = sqrt(L[8]);
  x[6] = L[3]; x[7] = L[7]; x[8] = L[6];
  return true;
}

__device__ void mvie_finish(const double* x, const double* interior, double* R, double* p, double* r) {
  for (int k = 0; k < 3; k++) p[k] = x[k] + interior[k];
  const double Lf[9] = {x[3] * x[3], 0, 0, x[6], x[4] * x[4], 0, x[8], x[7], x[5] * x[5]};
  double U[9], S[3];
  svd3_us(Lf, U, S);
  if (det3(U) < 0.0) {
    for (int k = 0; k < 3; k++) { R[k * 3 + 0] = U[k * 3 + 1]; R[k * 3 + 1] = U[k * 3 + 0]; R[k * 3 + 2] = U[k * 3 + 2]; }
    r[0] = S[1]; r[1] = S[0]; r[2] = S[2];
  } else {
    for (int k = 0; k < 9; k++) R[k] = U[k];
    for (int k = 0; k < 3; k++) r[k] = S[k];
  }
}

// checkCorridorValidity (baseline_fake.cpp:191-204) on m rows of (h0,h1,h2,h3)
__device__ bool corridor_valid(const double* h, int m, unsigned long long* rng, double* lpw, int* lpi,
                               double* ab /* 4*m scratch */) {
  for (int i = 0; i < m; i++) {
    for (int k = 0; k < 3; k++) ab[i * 4 + k] = h[i * 4 + k];
    ab[i * 4 + 3] = -h[i * 4 + 3];
  }
  const double c[3] = {0, 0, 0};
  double x[3];
  const double rst = poplp::linprog<3>(c, ab, 4, ab + 3, 4, m, x, rng, lpw, lpi);
  return !isinf(rst);
}

// =================================================================================================
// FIRI: one warp per (agent, segment)
// =================================================================================================
struct FiriWarp {           // shared memory per warp
  double fH[POP_MAX_FACES][4];
  double hP[POP_MAX_FACES][4];
  double R[9], p[3], r[3];
  double lm[2 * N9 * kMem + 2 * kMem];   // L-BFGS memory of the MVIE step
  double terms[POP_MAX_FACES][kTerms];   // per-face cost / gradient contributions
  double x9[N9], interior[3];
  int mvie_ok;
};

__device__ __forceinline__ void voxel_pos(const PopGeom& g, const float* pose, int idx, double* out) {
  const int x = idx % g.nx, y = (idx / g.nx) % g.ny, z = idx / (g.nx * g.ny);
  out[0] = static_cast<double>(geom_axis_position(x, g.rx, g.res, pose[0]));
  out[1] = static_cast<double>(geom_axis_position(y, g.ry, g.res, pose[1]));
  out[2] = static_cast<double>(geom_axis_position(z, g.rz, g.res, pose[2]));
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_corridor_firi(
    PopGeom g, const float* __restrict__ poses, const SegInfo* __restrict__ segs, const int* __restrict__ pts,
    int cap, double shrink_size, int id_base, double* __restrict__ distRs_all, double* __restrict__ tang_all,
    uint8_t* __restrict__ flags_all, double* __restrict__ scr_all, int* __restrict__ scri_all,
    double* __restrict__ out_h, int* __restrict__ out_rows, int* __restrict__ out_valid, int seg0, int n_segs,
    int* __restrict__ err_flag) {
  __shared__ FiriWarp sw[kWarpsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seg = seg0 + blockIdx.x * kWarpsPerBlock + warp;   // n_segs = end of this launch's range
  if (seg >= n_segs) return;
  const SegInfo& S = segs[seg];
  const int a = seg / POP_MAX_PIECES, si = seg % POP_MAX_PIECES;
  double* oh = out_h + static_cast<size_t>(seg) * POP_MAX_FACES * 4;
  if (!S.active) {
    if (lane == 0) { out_rows[seg] = 0; out_valid[seg] = 0; }
    return;
  }
  FiriWarp& W = sw[warp];
  const float* pose = poses + 3 * a;
  const int* pidx = pts + static_cast<size_t>(seg) * cap;
  double* distRs = distRs_all + static_cast<size_t>(seg) * cap;
  double* tang = tang_all + static_cast<size_t>(seg) * cap * 4;
  uint8_t* flags = flags_all + static_cast<size_t>(seg) * cap;
  double* scr = scr_all + static_cast<size_t>(seg) * kScrDoubles;
  int* scri = scri_all + static_cast<size_t>(seg) * kScrInts;
  const unsigned long long seed_base = static_cast<unsigned long long>(id_base + a);
  const unsigned FULL = 0xffffffffu;
  if (S.n_pts > cap) { if (lane == 0) atomicExch(err_flag, 1); }
  const int N = min(S.n_pts, cap);
  const double epsilon = 1.0e-6;
  const int M = 6;
  double bd[24] = {1, 0, 0, -S.lhc[0], 0, 1, 0, -S.lhc[1], 0, 0, 1, -S.lhc[2],
                   -1, 0, 0, S.llc[0], 0, -1, 0, S.llc[1], 0, 0, -1, S.llc[2]};
  int nH = 0;
  bool firi_ok = true;
  for (int i = 0; i < M && firi_ok; i++) if (dot3(&bd[i * 4], S.a) + bd[i * 4 + 3] > 0.0) firi_ok = false;
  for (int i = 0; i < M && firi_ok; i++) if (dot3(&bd[i * 4], S.b) + bd[i * 4 + 3] > 0.0) firi_ok = false;

  if (firi_ok) {
    if (lane == 0) {
      const double I9[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      for (int k = 0; k < 9; k++) W.R[k] = I9[k];
      for (int k = 0; k < 3; k++) { W.p[k] = 0.5 * (S.a[k] + S.b[k]); W.r[k] = 1.0; }
    }
    __syncwarp();
    for (int loop = 0; loop < 2; ++loop) {
      double R[9], p[3], r[3], fwd[9], bwd[9];
      for (int k = 0; k < 9; k++) R[k] = W.R[k];
      for (int k = 0; k < 3; k++) { p[k] = W.p[k]; r[k] = W.r[k]; }
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
          fwd[i * 3 + j] = (1.0 / r[i]) * R[j * 3 + i];
          bwd[i * 3 + j] = R[i * 3 + j] * r[j];
        }
      double fB[18], fD[6], distDs[6];
      for (int i = 0; i < M; i++) {
        for (int j = 0; j < 3; j++)
          fB[i * 3 + j] = (bd[i * 4] * bwd[j] + bd[i * 4 + 1] * bwd[3 + j]) + bd[i * 4 + 2] * bwd[6 + j];
        fD[i] = bd[i * 4 + 3] + dot3(&bd[i * 4], p);
        distDs[i] = fabs(fD[i]) / norm3(&fB[i * 3]);
      }
      auto fwd_mul = [&](const double* v, double* o) {
        for (int i = 0; i < 3; i++) o[i] = (fwd[i * 3] * v[0] + fwd[i * 3 + 1] * v[1]) + fwd[i * 3 + 2] * v[2];
      };
      double fa[3], fb[3], tmp[3];
      for (int k = 0; k < 3; k++) tmp[k] = S.a[k] - p[k];
      fwd_mul(tmp, fa);
      for (int k = 0; k < 3; k++) tmp[k] = S.b[k] - p[k];
      fwd_mul(tmp, fb);
      // tangent planes of every obstacle point (firi.hpp:268-297), lane-parallel
      double bestR = INFINITY;
      int bestI = 0x7fffffff;
      for (int i = lane; i < N; i += 32) {
        double pos[3], q[3], tg[4];
        voxel_pos(g, pose, pidx[i], pos);
        const double v[3] = {pos[0] - p[0], pos[1] - p[1], pos[2] - p[2]};
        fwd_mul(v, q);
        double dr = norm3(q);
        tg[3] = -dr;
        for (int k = 0; k < 3; k++) tg[k] = q[k] / dr;
        if (dot3(tg, fa) + tg[3] > epsilon) {
          const double delta[3] = {q[0] - fa[0], q[1] - fa[1], q[2] - fa[2]};
          const double coef = dot3(delta, fa) / dot3(delta, delta);
          for (int k = 0; k < 3; k++) tg[k] = fa[k] - coef * delta[k];
          dr = norm3(tg);
          tg[3] = -dr;
          for (int k = 0; k < 3; k++) tg[k] /= dr;
        }
        if (dot3(tg, fb) + tg[3] > epsilon) {
          const double delta[3] = {q[0] - fb[0], q[1] - fb[1], q[2] - fb[2]};
          const double coef = dot3(delta, fb) / dot3(delta, delta);
          for (int k = 0; k < 3; k++) tg[k] = fb[k] - coef * delta[k];
          dr = norm3(tg);
          tg[3] = -dr;
          for (int k = 0; k < 3; k++) tg[k] /= dr;
        }
        if (dot3(tg, fa) + tg[3] > epsilon) {
          const double u[3] = {fa[0] - q[0], fa[1] - q[1], fa[2] - q[2]};
          const double w[3] = {fb[0] - q[0], fb[1] - q[1], fb[2] - q[2]};
          double c[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
          const double n2 = dot3(c, c);
          if (n2 > 0.0) { const double nn = sqrt(n2); for (int k = 0; k < 3; k++) c[k] /= nn; }
          for (int k = 0; k < 3; k++) tg[k] = c[k];
          tg[3] = -dot3(tg, fa);
          const double sgn = tg[3] > 0.0 ? -1.0 : 1.0;
          for (int k = 0; k < 4; k++) tg[k] *= sgn;
        }
        distRs[i] = dr;
        for (int k = 0; k < 4; k++) tang[i * 4 + k] = tg[k];
        flags[i] = 1;
        if (dr < bestR) { bestR = dr; bestI = i; }
      }
      __syncwarp();
      // first-index argmin across lanes
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(FULL, bestR, o);
        const int oi = __shfl_xor_sync(FULL, bestI, o);
        if (ov < bestR || (ov == bestR && oi < bestI)) { bestR = ov; bestI = oi; }
      }
      uint8_t bdFlags[6] = {1, 1, 1, 1, 1, 1};
      nH = 0;
      bool completed = false;
      int bdMinId = 0, pcMinId = N > 0 ? bestI : 0;
      double minSqrD = distDs[0];
      for (int j = 1; j < M; j++) if (distDs[j] < minSqrD) { minSqrD = distDs[j]; bdMinId = j; }
      double minSqrR = N != 0 ? bestR : INFINITY;
      bool overflow = false;
      for (int i = 0; !completed && i < (M + N); ++i) {
        if (nH >= POP_MAX_FACES) { overflow = true; break; }
        double fh[4];
        if (minSqrD < minSqrR) {
          for (int k = 0; k < 3; k++) fh[k] = fB[bdMinId * 3 + k];
          fh[3] = fD[bdMinId];
          bdFlags[bdMinId] = 0;
        } else {
          for (int k = 0; k < 4; k++) fh[k] = tang[pcMinId * 4 + k];
          if (lane == 0) flags[pcMinId] = 0;
        }
        if (lane == 0) for (int k = 0; k < 4; k++) W.fH[nH][k] = fh[k];
        __syncwarp();
        completed = true;
        minSqrD = INFINITY;
        for (int j = 0; j < M; ++j)
          if (bdFlags[j]) {
            completed = false;
            if (minSqrD > distDs[j]) { bdMinId = j; minSqrD = distDs[j]; }
          }
        double lr = INFINITY;
        int li = 0x7fffffff;
        bool lane_open = false;
        for (int j = lane; j < N; j += 32) {
          if (flags[j]) {
            double pos[3], q[3];
            voxel_pos(g, pose, pidx[j], pos);
            const double v[3] = {pos[0] - p[0], pos[1] - p[1], pos[2] - p[2]};
            fwd_mul(v, q);
            if (dot3(fh, q) + fh[3] > -epsilon) {
              flags[j] = 0;
            } else {
              lane_open = true;
              const double dr = distRs[j];
              if (lr > dr) { li = j; lr = dr; }
            }
          }
        }
        __syncwarp();
        if (__any_sync(FULL, lane_open)) completed = false;
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_xor_sync(FULL, lr, o);
          const int oi = __shfl_xor_sync(FULL, li, o);
          if (ov < lr || (ov == lr && oi < li)) { lr = ov; li = oi; }
        }
        minSqrR = lr;
        pcMinId = li == 0x7fffffff ? 0 : li;
        ++nH;
      }
      if (overflow) { if (lane == 0) atomicExch(err_flag, 2); }
      __syncwarp();
      for (int i = lane; i < nH; i += 32) {
        const double* fh = W.fH[i];
        double h[4];
        for (int j = 0; j < 3; j++) h[j] = (fh[0] * fwd[j] + fh[1] * fwd[3 + j]) + fh[2] * fwd[6 + j];
        h[3] = fh[3] - dot3(h, p);
        for (int k = 0; k < 4; k++) W.hP[i][k] = h[k];
      }
      __syncwarp();
      if (loop == 1) break;
      if (lane == 0) {
        unsigned long long rng = lp_seed(seed_base, si, 0);
        W.mvie_ok = mvie_prepare(&W.hP[0][0], nH, W.R, W.p, W.r, &rng, scr, scri, W.x9, W.interior) ? 1 : 0;
      }
      __syncwarp();
      if (W.mvie_ok) {   // uniform: the whole warp runs the 9-variable L-BFGS, lanes share the face loop
        __threadfence_block();
        double x9[N9];
        for (int k = 0; k < N9; k++) x9[k] = W.x9[k];
        lbfgs9(scr + kScrA3, nH, x9, W.lm, W.terms);
        __syncwarp();
        if (lane == 0) {
          double Rl[9], pl[3], rl[3];
          mvie_finish(x9, W.interior, Rl, pl, rl);
          for (int k = 0; k < 9; k++) W.R[k] = Rl[k];
          for (int k = 0; k < 3; k++) { W.p[k] = pl[k]; W.r[k] = rl[k]; }
        }
      }
      __syncwarp();
    }
  }
  // ShrinkCorridor(hPoly, path): the simulator planner leaves front / back and top / bottom faces alone
  // (baseline_fake.cpp:215-228); the real-perception planner shrinks every face (baseline.cpp:213-226)
  const double path[3] = {S.b[0] - S.a[0], S.b[1] - S.a[1], S.b[2] - S.a[2]};
  const bool keep_end_faces = g.query_variant == POP_QUERY_FAKE;
  for (int i = lane; i < nH; i += 32) {
    double* h = W.hP[i];
    const double n3[3] = {h[0], h[1], h[2]}, z3[3] = {0, 0, 1};
    bool skip = false;
    if (keep_end_faces && dot3(n3, path) / norm3(n3) / norm3(path) > 0.8) skip = true;
    if (keep_end_faces && !skip && fabs(dot3(n3, z3)) / norm3(n3) > 0.8) skip = true;
    if (!skip) h[3] += sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]) * shrink_size;
  }
  __syncwarp();
  for (int i = lane; i < nH * 4; i += 32) oh[i] = (&W.hP[0][0])[i];
  if (lane == 0) {
    unsigned long long rng = lp_seed(seed_base, si, 1);
    const bool ok = corridor_valid(&W.hP[0][0], nH, &rng, scr + kScrLp, scri, scr + kScrAlp);
    out_rows[seg] = nH;
    out_valid[seg] = ok ? 1 : 0;
  }
}

// =================================================================================================
// per-agent tail: cut at first invalid corridor, intersections, goal reachability
// =================================================================================================
__global__ void k_corridor_finish(const double* __restrict__ path, const int* __restrict__ len,
                                  const double* __restrict__ start, int id_base, int variant,
                                  const double* __restrict__ seg_h, const int* __restrict__ seg_rows,
                                  const int* __restrict__ seg_valid, double* __restrict__ scr_all,
                                  int* __restrict__ scri_all, int* __restrict__ out_npoly, double* __restrict__ out_goal,
                                  int a0, int n_agents) {
  const int a = a0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // n_agents = end of the range
  if (a >= n_agents || (threadIdx.x & 31) != 0) return;
  // reuse segment 0's scratch of this agent: holds up to kLpRows rows of LP workspace
  double* scr = scr_all + static_cast<size_t>(a) * POP_MAX_PIECES * kScrDoubles;
  int* scri = scri_all + static_cast<size_t>(a) * POP_MAX_PIECES * kScrInts;
  double* lpw = scr + kScrLp;
  double* ab = scr + static_cast<size_t>(kScrDoubles);  // segment 1's scratch: 4 * kLpRows doubles fit
  out_npoly[a] = corridor_finish_agent(variant, static_cast<unsigned long long>(id_base + a), len[a],
                                       path + static_cast<size_t>(a) * POP_MAX_PATH * 6,
                                       seg_h + static_cast<size_t>(a) * POP_MAX_PIECES * POP_MAX_FACES * 4,
                                       seg_rows + a * POP_MAX_PIECES, seg_valid + a * POP_MAX_PIECES, start + a * 9, ab,
                                       lpw, scri, out_goal + a * 6);
}

}  // namespace

// ---- host side -----------------------------------------------------------------------------------
struct CorridorDev {
  int n_agents = 0, cap = 0;
  SegInfo* segs = nullptr;
  int* pts = nullptr;
  double* distRs = nullptr;
  double* tang = nullptr;
  uint8_t* flags = nullptr;
  double* scr = nullptr;
  int* scri = nullptr;
  double* seg_h = nullptr;   // [agent][POP_MAX_PIECES][POP_MAX_FACES][4]
  int* seg_rows = nullptr;   // [agent][POP_MAX_PIECES]
  int* seg_valid = nullptr;
  int* npoly = nullptr;      // [agent]
  double* goal = nullptr;    // [agent][6]
  double* tst = nullptr;     // [agent] absolute traj_start_time
  int* err = nullptr;
};

int popk_corridor_alloc(pop_handle* h) {
  if (h->corridor) return POP_OK;
  CorridorDev* d = new CorridorDev();
  h->corridor = d;
  d->n_agents = h->cfg.n_agents_local;
  d->cap = h->cfg.max_obstacle_points;
  const size_t ns = static_cast<size_t>(d->n_agents) * POP_MAX_PIECES;
  bool ok = true;
  ok &= cudaMalloc(&d->segs, ns * sizeof(SegInfo)) == cudaSuccess;
  ok &= cudaMalloc(&d->pts, ns * d->cap * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->distRs, ns * d->cap * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->tang, ns * d->cap * 4 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->flags, ns * d->cap) == cudaSuccess;
  ok &= cudaMalloc(&d->scr, (ns + 2) * kScrDoubles * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->scri, (ns + 2) * kScrInts * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->seg_h, ns * POP_MAX_FACES * 4 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->seg_rows, ns * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->seg_valid, ns * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->npoly, d->n_agents * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->goal, d->n_agents * 6 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->tst, d->n_agents * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->err, sizeof(int)) == cudaSuccess;
  static_assert(4 * kLpRows <= kScrDoubles, "finish kernel borrows one segment scratch for A|b");
  auto finish = [&]() -> int {   // as popk_astar_alloc: a failed attempt leaves h->corridor null
    if (!ok) { cudaGetLastError(); return pop_fail(h, POP_E_CUDA, "corridor workspace allocation failed"); }
    POP_CUDA(h, cudaMemsetAsync(d->err, 0, sizeof(int), h->stream));
    POP_CUDA(h, cudaMemsetAsync(d->seg_h, 0, ns * POP_MAX_FACES * 4 * sizeof(double), h->stream));
    POP_CUDA(h, cudaMemsetAsync(d->goal, 0, d->n_agents * 6 * sizeof(double), h->stream));
    return POP_OK;
  };
  const int rc = finish();
  if (rc != POP_OK) popk_corridor_free(h);
  return rc;
}

void popk_corridor_free(pop_handle* h) {
  CorridorDev* d = h->corridor;
  if (!d) return;
  void* ptrs[] = {d->segs, d->pts, d->distRs, d->tang, d->flags, d->scr, d->scri, d->seg_h, d->seg_rows,
                  d->seg_valid, d->npoly, d->goal, d->tst, d->err};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete d;
  h->corridor = nullptr;
}

// path/len/start device-resident; traj_start_time in corridor->tst
int popk_corridor_run(pop_handle* h, const double* d_path, const int* d_len, const double* d_start, cudaStream_t st,
                      int a0, int na) {
  CorridorDev* d = h->corridor;
  const pop_config& c = h->cfg;
  const int ns = na * POP_MAX_PIECES, seg0 = a0 * POP_MAX_PIECES;
  pop_prof_begin(h, POP_STAGE_CORRIDOR);
  if (c.cell_dtype == POP_CELL_U8)
    k_corridor_points<uint8_t><<<ns, 256, 0, st>>>(h->geom, h->d_grid, h->d_poses, h->t_map, d_path, d_len,
                                                          d_start, d->tst, c.corridor_tau, c.init_range, d->segs,
                                                          d->pts, d->cap, seg0);
  else
    k_corridor_points<float><<<ns, 256, 0, st>>>(h->geom, h->d_grid, h->d_poses, h->t_map, d_path, d_len,
                                                        d_start, d->tst, c.corridor_tau, c.init_range, d->segs,
                                                        d->pts, d->cap, seg0);
  POP_LAUNCH_CHECK(h);
  k_corridor_firi<<<(ns + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, st>>>(
      h->geom, h->d_poses, d->segs, d->pts, d->cap, c.shrink_size, c.agent_id_base, d->distRs, d->tang, d->flags,
      d->scr, d->scri, d->seg_h, d->seg_rows, d->seg_valid, seg0, seg0 + ns, d->err);
  POP_LAUNCH_CHECK(h);
  k_corridor_finish<<<(na + 3) / 4, 128, 0, st>>>(d_path, d_len, d_start, c.agent_id_base, c.query_variant, d->seg_h,
                                                                 d->seg_rows, d->seg_valid, d->scr, d->scri, d->npoly,
                                                                 d->goal, a0, a0 + na);
  POP_LAUNCH_CHECK(h);
  pop_prof_end(h, POP_STAGE_CORRIDOR, 3);
  return POP_OK;
}

double* popk_corridor_hpoly_dev(pop_handle* h) { return h->corridor->seg_h; }
int* popk_corridor_rows_dev(pop_handle* h) { return h->corridor->seg_rows; }
int* popk_corridor_npoly_dev(pop_handle* h) { return h->corridor->npoly; }
double* popk_corridor_goal_dev(pop_handle* h) { return h->corridor->goal; }
double* popk_corridor_tst_dev(pop_handle* h) { return h->corridor->tst; }

int popk_corridor_check_err(pop_handle* h) {
  int e = 0;
  POP_CUDA(h, cudaMemcpyAsync(&e, h->corridor->err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  if (e) {
    cudaMemsetAsync(h->corridor->err, 0, sizeof(int), h->stream);
    return pop_fail(h, POP_E_CAPACITY, e == 1 ? "a corridor box holds more obstacle points than max_obstacle_points=%d"
                                             : "a corridor polytope needs more than POP_MAX_FACES faces (cap %d)",
                    e == 1 ? h->cfg.max_obstacle_points : POP_MAX_FACES);
  }
  return POP_OK;
}

extern "C" int pop_corridor_batch(pop_handle* h, const double* path, const int32_t* len, const double* start,
                                  const double* traj_start_time, double* out_hpoly, int32_t* out_rows,
                                  int32_t* out_npoly, double* out_goal) {
  if (!h || !path || !len || !start || !traj_start_time || !out_hpoly || !out_rows || !out_npoly || !out_goal)
    return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->map_ready) return pop_fail(h, POP_E_STATE, "corridor stage before the first map update");
  int rc = popk_corridor_alloc(h);
  if (rc) return rc;
  CorridorDev* d = h->corridor;
  const int n = d->n_agents;
  const size_t bpath = sizeof(double) * 6 * POP_MAX_PATH * n, blen = sizeof(int) * n, bstart = sizeof(double) * 9 * n;
  rc = pop_ensure_scratch(h, bpath + blen + bstart + 64, 0);
  if (rc) return rc;
  uint8_t* s = static_cast<uint8_t*>(h->d_scratch);
  double* d_path = reinterpret_cast<double*>(s);
  double* d_start = reinterpret_cast<double*>(s + bpath);
  int* d_len = reinterpret_cast<int*>(s + bpath + bstart);
  POP_CUDA(h, cudaMemcpyAsync(d_path, path, bpath, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d_start, start, bstart, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d_len, len, blen, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d->tst, traj_start_time, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  rc = popk_corridor_run(h, d_path, d_len, d_start, h->stream, 0, h->corridor->n_agents);
  if (rc) return rc;
  const size_t ns = static_cast<size_t>(n) * POP_MAX_PIECES;
  POP_CUDA(h, cudaMemcpyAsync(out_hpoly, d->seg_h, ns * POP_MAX_FACES * 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(out_rows, d->seg_rows, ns * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(out_npoly, d->npoly, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(out_goal, d->goal, n * 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  return popk_corridor_check_err(h);
}
