// k_qp.cu — K4: corridor-constrained piecewise Bezier min-jerk QP, assembled and solved on the
// device, one CTA per agent.
//
// Assembly restates BezierOpt::setup (traj_opt/src/bezier_optimizer.cpp:27-260): x = [piece]
// [ctrl pt 0..4][xyz], block-diagonal jerk cost, 9(M+1) continuity/boundary equalities, 12M
// velocity and 9M acceleration boxes, 5 rows per corridor face.  Rows are kept in ELL form
// (<= 6 non-zeros, ascending columns) plus a column index for A' products.
// The solver is the ADMM of OSQP 0.6 — what BezierOpt::optimize reaches through
// traj_opt/include/iosqp.hpp:40-115 with eps_abs = eps_rel = 1e-3 and default settings —
// Ruiz equilibration, per-row rho, alpha = 1.6 relaxation, termination test every 25
// iterations, residual-balancing rho updates at a fixed interval.  The KKT solve is done on the
// reduced SPD matrix P + sigma I + A' diag(rho) A, which for this problem is block-tridiagonal
// with 15x15 blocks (continuity rows couple neighbouring pieces only): banded Cholesky once per
// rho, then block forward/backward recurrences with pre-multiplied inverse blocks, all in
// shared memory.  Inner products and inf-norms use warp shuffles + one shared-memory stage.
#include <cfloat>

#include "pop_internal.h"

namespace {

constexpr int kT = 256;                       // threads per CTA
constexpr int kMaxN = 15 * POP_MAX_PIECES;    // 180 variables
constexpr int kMaxRows = 9 * (POP_MAX_PIECES + 1) + 21 * POP_MAX_PIECES + 5 * POP_MAX_PIECES * POP_MAX_FACES;
constexpr int kEll = 6;
constexpr int kBand = 16;                     // half bandwidth 15 + diagonal
constexpr double OSQP_INFTY = 1e30, MIN_SCALING = 1e-4, MAX_SCALING = 1e4;
constexpr double RHO_MIN = 1e-6, RHO_MAX = 1e6, RHO_EQ = 1e3, RHO_TOL = 1e-4;
constexpr double SIGMA = 1e-6, ALPHA = 1.6, RHO0 = 0.1, ADAPT_TOL = 5.0;

struct QpParams {
  int max_iter, adaptive_interval, check_interval, scaling_iters;
  double eps_abs, eps_rel, eps_prim_inf, vmax, amax;
};

}  // namespace

struct QpDev {
  int cap_agents = 0;
  // inputs (device resident)
  int* n_pieces = nullptr;    // [agent]
  double* init = nullptr;     // [agent][9]
  double* goal = nullptr;     // [agent][9]
  double* dur = nullptr;      // [agent][POP_MAX_PIECES]
  // workspaces per agent
  int* ell_col = nullptr;     // [agent][kMaxRows][6]
  double* ell_val = nullptr;  // [agent][kMaxRows][6]
  int* csc_ptr = nullptr;     // [agent][kMaxN+1]
  int* csc_idx = nullptr;     // [agent][kMaxRows*6]  → row*6+slot
  double* Pblk = nullptr;     // [agent][POP_MAX_PIECES][225]
  double* E = nullptr;        // [agent][kMaxRows]
  double* lo = nullptr;       // [agent][kMaxRows] scaled lower bounds
  double* up = nullptr;       // [agent][kMaxRows] scaled upper bounds
  double* dy = nullptr;       // [agent][kMaxRows]
  double* Kb = nullptr;       // [agent][kMaxN][kBand] banded lower factor
  double* QM = nullptr;       // [225] jerk cost block
  // outputs
  double* cpts = nullptr;     // [agent][POP_MAX_PIECES*5][3]
  int* status = nullptr;      // [agent]
  int* iters = nullptr;       // [agent]
};

namespace {

__device__ __forceinline__ double limit_scaling(double v) {
  v = v < MIN_SCALING ? 1.0 : v;
  v = v > MAX_SCALING ? MAX_SCALING : v;
  return v;
}

// block-wide max of a per-thread value (all threads receive the result)
__device__ double block_max(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = red[0];
  for (int w = 1; w < kT / 32; w++) r = fmax(r, red[w]);
  return r;
}
// deterministic block-wide sum (fixed tree)
__device__ double block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0;
  for (int w = 0; w < kT / 32; w++) r += red[w];
  return r;
}

// Shared-memory row capacity: rows beyond it make the solve report status -20 (the host turns
// that into POP_E_CAPACITY). 1536 rows = 233 corridor faces over 12 pieces.
constexpr int kRowCap = 1536;
// Packed lower-triangular inverse Cholesky factor; row i starts at a bank offset of i mod 16
// doubles so that both traversals (row-per-thread forward, column-per-thread backward) are
// bank-conflict free.
constexpr int kLinvCap = kMaxN * (kMaxN + 1) / 2 + 16 * kMaxN;

struct Smem {
  double x[kMaxN], xp[kMaxN], xt[kMaxN], rhs[kMaxN], tn[kMaxN], D[kMaxN];
  double Linv[kLinvCap];
  int row_off[kMaxN];
  double z[kRowCap], y[kRowCap], w[kRowCap];   // w = rho z - y
  signed char type[kRowCap];
  double red[kT / 32];
  int face_off[POP_MAX_PIECES + 1];
  double t[POP_MAX_PIECES];
  int M, n, m, status, fail;
  double rho, c_scale;
};

// ---- row construction (bezier_optimizer.cpp:133-260); entries ascending in column -----------------
__device__ void build_row(const Smem& S, int r, const double* init, const double* goal, const double* hpoly,
                          const int* rows, const QpParams& P, int* col, double* val, double* lo, double* up) {
  const int M = S.M;
  for (int k = 0; k < kEll; k++) { col[k] = -1; val[k] = 0.0; }
  const int nb = 3 * (M + 1);
  const double t0 = S.t[0], tM = S.t[M - 1];
  if (r < nb) {  // position
    const int b = r / 3, d = r % 3;
    if (b == 0) { col[0] = d; val[0] = 1; *lo = *up = init[d]; }
    else if (b == M) { col[0] = 15 * M - 3 + d; val[0] = 1; *lo = *up = goal[d]; }
    else { col[0] = 15 * b - 3 + d; val[0] = -1; col[1] = 15 * b + d; val[1] = 1; *lo = *up = 0.0; }
    return;
  }
  r -= nb;
  if (r < nb) {  // velocity
    const int b = r / 3, d = r % 3;
    if (b == 0) { col[0] = d; val[0] = -4; col[1] = 3 + d; val[1] = 4; *lo = *up = init[3 + d] * t0; }
    else if (b == M) { col[0] = 15 * M - 6 + d; val[0] = -4; col[1] = 15 * M - 3 + d; val[1] = 4; *lo = *up = goal[3 + d] * tM; }
    else {
      const double t1 = S.t[b], t1_ = S.t[b - 1];
      col[0] = 15 * b - 6 + d; val[0] = 4 * 1.0 / t1_;
      col[1] = 15 * b - 3 + d; val[1] = -4 * 1.0 / t1_;
      col[2] = 15 * b + d;     val[2] = -4 * 1.0 / t1;
      col[3] = 15 * b + 3 + d; val[3] = 4 * 1.0 / t1;
      *lo = *up = 0.0;
    }
    return;
  }
  r -= nb;
  if (r < nb) {  // acceleration: 12 c0 - 24 c1 + 12 c2
    const int b = r / 3, d = r % 3;
    if (b == 0) {
      col[0] = d; val[0] = 12; col[1] = 3 + d; val[1] = -24; col[2] = 6 + d; val[2] = 12;
      *lo = *up = init[6 + d] * t0 * t0;
    } else if (b == M) {
      col[0] = 15 * M - 9 + d; val[0] = 12; col[1] = 15 * M - 6 + d; val[1] = -24; col[2] = 15 * M - 3 + d; val[2] = 12;
      *lo = *up = goal[6 + d] * tM * tM;
    } else {
      const double t2 = S.t[b] * S.t[b], t2_ = S.t[b - 1] * S.t[b - 1];
      col[0] = 15 * b - 9 + d; val[0] = -12.0 / t2_;
      col[1] = 15 * b - 6 + d; val[1] = 24.0 / t2_;
      col[2] = 15 * b - 3 + d; val[2] = -12.0 / t2_;
      col[3] = 15 * b + d;     val[3] = 12.0 / t2;
      col[4] = 15 * b + 3 + d; val[4] = -24.0 / t2;
      col[5] = 15 * b + 6 + d; val[5] = 12.0 / t2;
      *lo = *up = 0.0;
    }
    return;
  }
  r -= nb;
  if (r < 12 * M) {  // velocity box
    const int i = r / 12, j = (r % 12) / 3, d = r % 3;
    col[0] = 15 * i + 3 * j + d; val[0] = -4; col[1] = 15 * i + 3 * j + 3 + d; val[1] = 4;
    *up = P.vmax * 1.0 * S.t[i]; *lo = -P.vmax * 1.0 * S.t[i];
    return;
  }
  r -= 12 * M;
  if (r < 9 * M) {  // acceleration box
    const int i = r / 9, j = (r % 9) / 3, d = r % 3;
    col[0] = 15 * i + 3 * j + d; val[0] = 12; col[1] = 15 * i + 3 * j + 3 + d; val[1] = -24;
    col[2] = 15 * i + 3 * j + 6 + d; val[2] = 12;
    *up = P.amax * 1.0 * S.t[i] * S.t[i]; *lo = -P.amax * 1.0 * S.t[i] * S.t[i];
    return;
  }
  r -= 9 * M;
  // safety: piece i, face j, ctrl point k
  int i = 0;
  while (i + 1 < M && r >= 5 * S.face_off[i + 1]) i++;
  const int rr = r - 5 * S.face_off[i];
  const int j = rr / 5, k = rr % 5;
  const double* h = hpoly + (static_cast<size_t>(i) * POP_MAX_FACES + j) * 4;
  for (int d = 0; d < 3; d++) { col[d] = 15 * i + 3 * k + d; val[d] = h[d]; }
  *up = -h[3];
  *lo = -OSQP_INFTY;
  (void)rows;
}

__device__ __forceinline__ double row_dot(const int* col, const double* val, const double* x) {
  double s = 0;
#pragma unroll
  for (int k = 0; k < kEll; k++) {
    const int c = col[k];
    if (c >= 0) s += val[k] * x[c];
  }
  return s;
}

// K = P + sigma I + A' diag(rho) A in banded-lower storage Kb[c][d] = K[c+d][c], then banded
// Cholesky in place and the block recurrence matrices Linv/G/H.
__device__ void factorize(Smem& S, const int* ell_col, const double* ell_val, const int* csc_ptr, const int* csc_idx,
                          const double* Pblk, double* Kb) {
  const int n = S.n, M = S.M;
  const int tid = threadIdx.x;
  for (int e = tid; e < n * kBand; e += kT) {
    const int c = e / kBand, d = e % kBand, r = c + d;
    double v = 0.0;
    if (r < n) {
      if (r / 15 == c / 15) v = Pblk[(c / 15) * 225 + (r % 15) * 15 + (c % 15)];
      if (d == 0) v += SIGMA;
      for (int q = csc_ptr[r]; q < csc_ptr[r + 1]; q++) {   // rows containing column r, ascending
        const int idx = csc_idx[q], row = idx / kEll;
        const int* rc = ell_col + row * kEll;
        for (int k = 0; k < kEll; k++)
          if (rc[k] == c) {
            const double rho = S.type[row] == 1 ? RHO_EQ * S.rho : (S.type[row] == 0 ? S.rho : RHO_MIN);
            v += (rho * ell_val[idx]) * ell_val[row * kEll + k];
            break;
          }
      }
    }
    Kb[e] = v;
  }
  __syncthreads();
  // banded Cholesky, column by column
  for (int j = 0; j < n; j++) {
    if (tid == 0) {
      double d = Kb[j * kBand];
      for (int k = max(0, j - 15); k < j; k++) { const double l = Kb[k * kBand + (j - k)]; d -= l * l; }
      if (!(d > 0.0)) S.fail = 1;
      Kb[j * kBand] = sqrt(d);
    }
    __syncthreads();
    if (tid >= 1 && tid <= 15 && j + tid < n) {
      const int i = j + tid;
      double s = Kb[j * kBand + tid];
      for (int k = max(0, i - 15); k < j; k++) s -= Kb[k * kBand + (i - k)] * Kb[k * kBand + (j - k)];
      Kb[j * kBand + tid] = s / Kb[j * kBand];
    }
    __syncthreads();
  }
  // Linv = L^-1 (dense lower triangular, packed rows): thread per column, forward substitution on
  // the banded factor; previously computed entries of the same column are read back from smem
  for (int c = tid; c < n; c += kT) {
    for (int r = c; r < n; r++) {
      double s = r == c ? 1.0 : 0.0;
      const int k0 = max(c, r - 15);
      for (int k = k0; k < r; k++) s -= Kb[k * kBand + (r - k)] * S.Linv[S.row_off[k] + c];
      S.Linv[S.row_off[r] + c] = s / Kb[r * kBand];
    }
  }
  __syncthreads();
}

// solve K xt = rhs:  y = L^-1 rhs,  xt = L^-T y  (two triangular mat-vecs from shared memory)
__device__ void kkt_solve(Smem& S) {
  const int n = S.n, tid = threadIdx.x;
  if (tid < n) {
    const double* row = S.Linv + S.row_off[tid];
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int j = 0;
    for (; j + 3 <= tid; j += 4) {
      a0 += row[j] * S.rhs[j];
      a1 += row[j + 1] * S.rhs[j + 1];
      a2 += row[j + 2] * S.rhs[j + 2];
      a3 += row[j + 3] * S.rhs[j + 3];
    }
    for (; j <= tid; j++) a0 += row[j] * S.rhs[j];
    S.tn[tid] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  if (tid < n) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int i = tid;
    for (; i + 3 < n; i += 4) {
      a0 += S.Linv[S.row_off[i] + tid] * S.tn[i];
      a1 += S.Linv[S.row_off[i + 1] + tid] * S.tn[i + 1];
      a2 += S.Linv[S.row_off[i + 2] + tid] * S.tn[i + 2];
      a3 += S.Linv[S.row_off[i + 3] + tid] * S.tn[i + 3];
    }
    for (; i < n; i++) a0 += S.Linv[S.row_off[i] + tid] * S.tn[i];
    S.xt[tid] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kT, 1) k_qp_admm(QpDev Dv, QpParams P, const double* __restrict__ hpoly_all,
                                                   const int* __restrict__ rows_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int a = blockIdx.x, tid = threadIdx.x;
  const int M = Dv.n_pieces[a];
  if (M <= 0 || M > POP_MAX_PIECES) {
    if (tid == 0) { Dv.status[a] = 0; Dv.iters[a] = 0; }
    return;
  }
  const double* init = Dv.init + a * 9;
  const double* goal = Dv.goal + a * 9;
  const double* hpoly = hpoly_all + static_cast<size_t>(a) * POP_MAX_PIECES * POP_MAX_FACES * 4;
  const int* rows = rows_all + a * POP_MAX_PIECES;
  int* ell_col = Dv.ell_col + static_cast<size_t>(a) * kMaxRows * kEll;
  double* ell_val = Dv.ell_val + static_cast<size_t>(a) * kMaxRows * kEll;
  int* csc_ptr = Dv.csc_ptr + static_cast<size_t>(a) * (kMaxN + 1);
  int* csc_idx = Dv.csc_idx + static_cast<size_t>(a) * kMaxRows * kEll;
  double* Pblk = Dv.Pblk + static_cast<size_t>(a) * POP_MAX_PIECES * 225;
  double* E = Dv.E + static_cast<size_t>(a) * kMaxRows;
  double* dyv = Dv.dy + static_cast<size_t>(a) * kMaxRows;
  double* Kb = Dv.Kb + static_cast<size_t>(a) * kMaxN * kBand;
  double* lov = Dv.lo + static_cast<size_t>(a) * kMaxRows;
  double* upv = Dv.up + static_cast<size_t>(a) * kMaxRows;

  if (tid == 0) {
    S.M = M; S.n = 15 * M;
    int off = 0;
    for (int i = 0; i < 15 * M; i++) {
      while ((off & 15) != (i & 15)) off++;
      S.row_off[i] = off;
      off += i + 1;
    }
    S.face_off[0] = 0;
    for (int i = 0; i < M; i++) {
      int r = rows[i];
      r = r < 0 ? 0 : (r > POP_MAX_FACES ? POP_MAX_FACES : r);
      S.face_off[i + 1] = S.face_off[i] + r;
      S.t[i] = Dv.dur[a * POP_MAX_PIECES + i];
    }
    S.m = 9 * (M + 1) + 21 * M + 5 * S.face_off[M];
    S.status = 0; S.fail = 0; S.rho = RHO0; S.c_scale = 1.0;
  }
  __syncthreads();
  const int n = S.n, m = S.m;
  if (m > kRowCap) {
    if (tid == 0) { Dv.status[a] = -20; Dv.iters[a] = 0; }
    return;
  }

  // ---- assembly ----
  for (int r = tid; r < m; r += kT) {
    int col[kEll];
    double val[kEll], lo, up;
    build_row(S, r, init, goal, hpoly, rows, P, col, val, &lo, &up);
    for (int k = 0; k < kEll; k++) { ell_col[r * kEll + k] = col[k]; ell_val[r * kEll + k] = val[k]; }
    lov[r] = fmax(lo, -OSQP_INFTY);
    upv[r] = fmin(up, OSQP_INFTY);
    S.z[r] = 0.0; S.y[r] = 0.0; S.w[r] = 0.0;
    E[r] = 1.0;
  }
  for (int e = tid; e < M * 225; e += kT) Pblk[e] = Dv.QM[e % 225];
  if (tid < n) { S.D[tid] = 1.0; S.x[tid] = 0.0; }
  __syncthreads();
  // column index: entries of column j in ascending row order
  if (tid == 0) csc_ptr[0] = 0;
  if (tid < n) {
    int cnt = 0;
    for (int r = 0; r < m; r++)
      for (int k = 0; k < kEll; k++) if (ell_col[r * kEll + k] == tid) cnt++;
    S.tn[tid] = cnt;
  }
  __syncthreads();
  if (tid == 0) { int run = 0; for (int j = 0; j < n; j++) { run += static_cast<int>(S.tn[j]); csc_ptr[j + 1] = run; } }
  __syncthreads();
  if (tid < n) {
    int w = csc_ptr[tid];
    for (int r = 0; r < m; r++)
      for (int k = 0; k < kEll; k++) if (ell_col[r * kEll + k] == tid) csc_idx[w++] = r * kEll + k;
  }
  __syncthreads();

  // ---- Ruiz equilibration (osqp scale_data, scaling = 10) ----
  for (int it = 0; it < P.scaling_iters; it++) {
    double dt = 1.0;
    if (tid < n) {
      double cp = 0, ca = 0;
      const int b = tid / 15, c = tid % 15;
      for (int r = 0; r < 15; r++) cp = fmax(cp, fabs(Pblk[b * 225 + r * 15 + c]));
      for (int q = csc_ptr[tid]; q < csc_ptr[tid + 1]; q++) ca = fmax(ca, fabs(ell_val[csc_idx[q]]));
      dt = 1.0 / sqrt(limit_scaling(fmax(cp, ca)));
      S.tn[tid] = dt;
    }
    __syncthreads();
    for (int r = tid; r < m; r += kT) {
      double rm = 0;
      for (int k = 0; k < kEll; k++) if (ell_col[r * kEll + k] >= 0) rm = fmax(rm, fabs(ell_val[r * kEll + k]));
      const double et = 1.0 / sqrt(limit_scaling(rm));
      for (int k = 0; k < kEll; k++) {
        const int c = ell_col[r * kEll + k];
        if (c >= 0) ell_val[r * kEll + k] = (et * ell_val[r * kEll + k]) * S.tn[c];
      }
      E[r] *= et;
    }
    for (int e = tid; e < M * 225; e += kT) {
      const int b = e / 225, r = (e % 225) / 15, c = e % 15;
      Pblk[e] = (S.tn[15 * b + r] * Pblk[e]) * S.tn[15 * b + c];
    }
    if (tid < n) S.D[tid] *= dt;
    __syncthreads();
    double cpm = 0.0;
    if (tid < n) {
      const int b = tid / 15, c = tid % 15;
      for (int r = 0; r < 15; r++) cpm = fmax(cpm, fabs(Pblk[b * 225 + r * 15 + c]));
      S.rhs[tid] = cpm;
    }
    __syncthreads();
    if (tid == 0) {
      double mean = 0;
      for (int j = 0; j < n; j++) mean += S.rhs[j];
      mean /= n;
      const double qn = limit_scaling(0.0);
      double ct = limit_scaling(fmax(mean, qn));
      ct = 1. / ct;
      S.rhs[0] = ct;
      S.c_scale *= ct;
    }
    __syncthreads();
    const double ct = S.rhs[0];
    for (int e = tid; e < M * 225; e += kT) Pblk[e] *= ct;
    __syncthreads();
  }
  // scaled bounds and constraint types (set_rho_vec)
  for (int r = tid; r < m; r += kT) {
    lov[r] *= E[r];
    upv[r] *= E[r];
    signed char ty;
    if (lov[r] < -OSQP_INFTY * MIN_SCALING && upv[r] > OSQP_INFTY * MIN_SCALING) ty = -1;
    else if (upv[r] - lov[r] < RHO_TOL) ty = 1;
    else ty = 0;
    S.type[r] = ty;
  }
  __syncthreads();
  factorize(S, ell_col, ell_val, csc_ptr, csc_idx, Pblk, Kb);
  if (S.fail) {
    if (tid == 0) { Dv.status[a] = -7; Dv.iters[a] = 0; }
    return;
  }
  const double cinv = 1.0 / S.c_scale;

  // ---- ADMM ----
  int iter, status = 0;
  bool can_check = false;
  double pri_res = 0, dua_res = 0;
  auto rho_of = [&](int r) { return S.type[r] == 1 ? RHO_EQ * S.rho : (S.type[r] == 0 ? S.rho : RHO_MIN); };

  // residuals + norms needed by check_termination / compute_rho_estimate.  Outputs (all threads):
  double nz = 0, nAx = 0, nAty = 0, nPx = 0, rp_scaled = 0, rd_scaled = 0, nz_un = 0, nd_un = 0;
  auto update_info = [&]() {
    double v_pri = 0, v_rp = 0, v_z = 0, v_ax = 0, v_ez = 0, v_eax = 0;
    for (int r = tid; r < m; r += kT) {
      const double ax = row_dot(ell_col + r * kEll, ell_val + r * kEll, S.x);
      const double res = ax - S.z[r];
      const double einv = 1.0 / E[r];
      v_pri = fmax(v_pri, fabs(einv * res));
      v_rp = fmax(v_rp, fabs(res));
      v_z = fmax(v_z, fabs(S.z[r]));
      v_ax = fmax(v_ax, fabs(ax));
      v_ez = fmax(v_ez, fabs(einv * S.z[r]));
      v_eax = fmax(v_eax, fabs(einv * ax));
    }
    double v_dua = 0, v_rd = 0, v_aty = 0, v_px = 0, v_daty = 0, v_dpx = 0;
    if (tid < n) {
      const int b = tid / 15, rr = tid % 15;
      double px = 0;
      for (int k = 0; k < 15; k++) px += Pblk[b * 225 + rr * 15 + k] * S.x[15 * b + k];
      double aty = 0;
      for (int q = csc_ptr[tid]; q < csc_ptr[tid + 1]; q++) { const int idx = csc_idx[q]; aty += ell_val[idx] * S.y[idx / kEll]; }
      const double res = (0.0 + px) + aty;
      const double dinv = 1.0 / S.D[tid];
      v_dua = fabs(dinv * res); v_rd = fabs(res); v_aty = fabs(aty); v_px = fabs(px);
      v_daty = fabs(dinv * aty); v_dpx = fabs(dinv * px);
    }
    pri_res = block_max(v_pri, S.red);
    rp_scaled = block_max(v_rp, S.red);
    nz = block_max(v_z, S.red);
    nAx = block_max(v_ax, S.red);
    const double ez = block_max(v_ez, S.red), eax = block_max(v_eax, S.red);
    dua_res = cinv * block_max(v_dua, S.red);
    rd_scaled = block_max(v_rd, S.red);
    nAty = block_max(v_aty, S.red);
    nPx = block_max(v_px, S.red);
    const double daty = block_max(v_daty, S.red), dpx = block_max(v_dpx, S.red);
    // unscaled norms for the tolerance computation (compute_pri_tol / compute_dua_tol)
    nz_un = fmax(ez, eax);
    nd_un = cinv * fmax(fmax(0.0, daty), dpx);
  };

  auto primal_infeasible = [&](double eps) -> bool {
    // is_primal_infeasible: delta_y projected on the polar of the recession cone, stored in dyv
    double v_nd = 0, v_lhs = 0;
    for (int r = tid; r < m; r += kT) {
      double d = dyv[r];
      if (upv[r] > OSQP_INFTY * MIN_SCALING) {
        if (lov[r] < -OSQP_INFTY * MIN_SCALING) d = 0.0;
        else d = fmin(d, 0.0);
      } else if (lov[r] < -OSQP_INFTY * MIN_SCALING) {
        d = fmax(d, 0.0);
      }
      dyv[r] = d;
      v_nd = fmax(v_nd, fabs(E[r] * d));
      v_lhs += upv[r] * fmax(d, 0.0) + lov[r] * fmin(d, 0.0);
    }
    const double nd = block_max(v_nd, S.red);
    const double lhs = block_sum(v_lhs, S.red);
    if (nd > eps && lhs < -eps * nd) {
      double v = 0;
      if (tid < n) {
        double s = 0;
        for (int q = csc_ptr[tid]; q < csc_ptr[tid + 1]; q++) { const int idx = csc_idx[q]; s += ell_val[idx] * dyv[idx / kEll]; }
        v = fabs((1.0 / S.D[tid]) * s);
      }
      const double nrm = block_max(v, S.red);
      return nrm < eps * nd;
    }
    return false;
  };
  auto check_termination = [&](bool approximate) -> bool {
    double eps_abs = P.eps_abs, eps_rel = P.eps_rel, eps_pi = P.eps_prim_inf;
    if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pi *= 10; }
    if (pri_res > OSQP_INFTY || dua_res > OSQP_INFTY) { status = -7; return true; }
    bool prim_ok = false, dual_ok = false, prim_inf = false;
    const double eps_prim = eps_abs + eps_rel * nz_un;
    if (pri_res < eps_prim) prim_ok = true;
    else prim_inf = primal_infeasible(eps_pi);
    const double eps_dual = eps_abs + eps_rel * nd_un;
    if (dua_res < eps_dual) dual_ok = true;
    if (prim_ok && dual_ok) { status = approximate ? 2 : 1; return true; }
    if (prim_inf) { status = approximate ? 3 : -3; return true; }
    return false;
  };

  for (iter = 1; iter <= P.max_iter; iter++) {
    // x_prev <- x ; rhs = sigma x_prev - q + A' w with w = rho z - y.  Half-warp per column, lanes
    // stride its entries, fixed shuffle tree (deterministic order).
    {
      const int grp = tid >> 4, ln = tid & 15;
      for (int j0 = 0; j0 < n; j0 += kT / 16) {
        const int j = j0 + grp;
        double s = 0.0;
        if (j < n) {
          const int q1 = csc_ptr[j + 1];
          for (int q = csc_ptr[j] + ln; q < q1; q += 16) {
            const int idx = csc_idx[q];
            s += ell_val[idx] * S.w[idx / kEll];
          }
        }
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (j < n && ln == 0) {
          const double xv = S.x[j];
          S.xp[j] = xv;
          S.rhs[j] = (SIGMA * xv - 0.0) + s;
        }
      }
    }
    __syncthreads();
    kkt_solve(S);
    __syncthreads();
    const bool store_dy = P.check_interval && (iter % P.check_interval == 0);
    for (int r = tid; r < m; r += kT) {
      const double zt = row_dot(ell_col + r * kEll, ell_val + r * kEll, S.xt);
      const double rho = rho_of(r), zp = S.z[r], yv = S.y[r];
      const double mix = ALPHA * zt + (1.0 - ALPHA) * zp;
      const double v = mix + (1. / rho) * yv;
      const double zn = fmin(fmax(v, lov[r]), upv[r]);
      const double d = rho * (mix - zn);
      S.z[r] = zn;
      S.y[r] = yv + d;
      S.w[r] = rho * zn - (yv + d);
      if (store_dy) dyv[r] = d;
    }
    if (tid < n) S.x[tid] = ALPHA * S.xt[tid] + (1.0 - ALPHA) * S.xp[tid];
    __syncthreads();
    can_check = P.check_interval && (iter % P.check_interval == 0);
    bool info = false;
    if (can_check) {
      update_info();
      info = true;
      if (check_termination(false)) break;
    }
    if (P.adaptive_interval && (iter % P.adaptive_interval == 0)) {
      if (!info) update_info();
      double pr = rp_scaled / (fmax(nz, nAx) + 1e-10);
      double dr = rd_scaled / (fmax(fmax(0.0, nAty), nPx) + 1e-10);
      double rho_new = S.rho * sqrt(pr / (dr + 1e-10));
      rho_new = fmin(fmax(rho_new, RHO_MIN), RHO_MAX);
      if (rho_new > S.rho * ADAPT_TOL || rho_new < S.rho / ADAPT_TOL) {
        __syncthreads();
        if (tid == 0) S.rho = rho_new;
        __syncthreads();
        for (int r = tid; r < m; r += kT) S.w[r] = rho_of(r) * S.z[r] - S.y[r];
        factorize(S, ell_col, ell_val, csc_ptr, csc_idx, Pblk, Kb);
      }
    }
  }
  if (status == 0 && !can_check) { update_info(); check_termination(false); }
  if (status == 0) { if (!check_termination(true)) status = -2; }
  __syncthreads();
  if (tid < n) Dv.cpts[static_cast<size_t>(a) * POP_MAX_PIECES * 15 + tid] = S.D[tid] * S.x[tid];
  if (tid == 0) {
    Dv.status[a] = status;
    Dv.iters[a] = iter < P.max_iter ? iter : P.max_iter;
  }
}

}  // namespace

// ---- host side -----------------------------------------------------------------------------------
static void host_jerk_block(double* QM) {
  // calcCtrlPtsCvtMat + calcMinJerkCost (bezier_optimizer.cpp:56-105), same evaluation order
  const int N_ = 4;
  double p2v[12][15] = {{0}}, v2a[9][12] = {{0}}, a2j[6][9] = {{0}};
  for (int i = 0; i < N_; i++) for (int d = 0; d < 3; d++) { p2v[i * 3 + d][i * 3 + d] = -N_; p2v[i * 3 + d][(i + 1) * 3 + d] = N_; }
  for (int i = 0; i < N_ - 1; i++) for (int d = 0; d < 3; d++) { v2a[i * 3 + d][i * 3 + d] = -(N_ - 1); v2a[i * 3 + d][(i + 1) * 3 + d] = N_ - 1; }
  for (int i = 0; i < N_ - 2; i++) for (int d = 0; d < 3; d++) { a2j[i * 3 + d][i * 3 + d] = -(N_ - 2); a2j[i * 3 + d][(i + 1) * 3 + d] = N_ - 2; }
  double av[6][12] = {{0}}, p2j[6][15] = {{0}};
  for (int i = 0; i < 6; i++) for (int j = 0; j < 12; j++) { double s = 0; for (int k = 0; k < 9; k++) s += a2j[i][k] * v2a[k][j]; av[i][j] = s; }
  for (int i = 0; i < 6; i++) for (int j = 0; j < 15; j++) { double s = 0; for (int k = 0; k < 12; k++) s += av[i][k] * p2v[k][j]; p2j[i][j] = s; }
  double Pm[6][6] = {{0}};
  for (int d = 0; d < 3; d++) { Pm[d][d] = 1.0 / 3; Pm[d][3 + d] = 1.0 / 6; Pm[3 + d][d] = 1.0 / 6; Pm[3 + d][3 + d] = 1.0 / 3; }
  double tP[15][6] = {{0}};
  for (int i = 0; i < 15; i++) for (int j = 0; j < 6; j++) { double s = 0; for (int k = 0; k < 6; k++) s += p2j[k][i] * Pm[k][j]; tP[i][j] = s; }
  for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { double s = 0; for (int k = 0; k < 6; k++) s += tP[i][k] * p2j[k][j]; QM[i * 15 + j] = s; }
}

int popk_qp_alloc(pop_handle* h, int n_agents) {
  if (h->qp && h->qp->cap_agents >= n_agents) return POP_OK;
  if (h->qp) popk_qp_free(h);
  QpDev* d = new QpDev();
  h->qp = d;
  d->cap_agents = n_agents;
  const size_t na = n_agents;
  bool ok = true;
  ok &= cudaMalloc(&d->n_pieces, na * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->init, na * 9 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->goal, na * 9 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->dur, na * POP_MAX_PIECES * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->ell_col, na * kMaxRows * kEll * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->ell_val, na * kMaxRows * kEll * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->csc_ptr, na * (kMaxN + 1) * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->csc_idx, na * kMaxRows * kEll * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->Pblk, na * POP_MAX_PIECES * 225 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->E, na * kMaxRows * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->dy, na * kMaxRows * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->lo, na * kMaxRows * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->up, na * kMaxRows * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->Kb, na * kMaxN * kBand * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->QM, 225 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->cpts, na * POP_MAX_PIECES * 15 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->status, na * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->iters, na * sizeof(int)) == cudaSuccess;
  if (!ok) { cudaGetLastError(); return pop_fail(h, POP_E_CUDA, "QP workspace allocation failed"); }
  double QM[225];
  host_jerk_block(QM);
  POP_CUDA(h, cudaMemcpyAsync(d->QM, QM, sizeof(QM), cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemsetAsync(d->cpts, 0, na * POP_MAX_PIECES * 15 * sizeof(double), h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  POP_CUDA(h, cudaFuncSetAttribute(k_qp_admm, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(Smem))));
  return POP_OK;
}

void popk_qp_free(pop_handle* h) {
  QpDev* d = h->qp;
  if (!d) return;
  void* ptrs[] = {d->n_pieces, d->init, d->goal, d->dur, d->ell_col, d->ell_val, d->csc_ptr, d->csc_idx, d->Pblk,
                  d->E, d->dy, d->lo, d->up, d->Kb, d->QM, d->cpts, d->status, d->iters};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete d;
  h->qp = nullptr;
}

// inputs resident in qp->{n_pieces, init, goal, dur}; corridors at d_hpoly / d_rows
int popk_qp_run(pop_handle* h, int n_agents, const double* d_hpoly, const int* d_rows, double vmax, double amax) {
  QpDev* d = h->qp;
  const pop_config& c = h->cfg;
  QpParams P;
  P.max_iter = c.osqp_max_iter; P.adaptive_interval = c.osqp_adaptive_rho_interval;
  P.check_interval = c.osqp_check_termination; P.scaling_iters = 10;
  P.eps_abs = c.osqp_eps_abs; P.eps_rel = c.osqp_eps_rel; P.eps_prim_inf = 1e-4;
  P.vmax = vmax; P.amax = amax;
  pop_prof_begin(h, POP_STAGE_QP);
  k_qp_admm<<<n_agents, kT, sizeof(Smem), h->stream>>>(*d, P, d_hpoly, d_rows);
  POP_LAUNCH_CHECK(h);
  pop_prof_end(h, POP_STAGE_QP, 1);
  return POP_OK;
}

int* popk_qp_npieces_dev(pop_handle* h) { return h->qp->n_pieces; }
double* popk_qp_init_dev(pop_handle* h) { return h->qp->init; }
double* popk_qp_goal_dev(pop_handle* h) { return h->qp->goal; }
double* popk_qp_dur_dev(pop_handle* h) { return h->qp->dur; }
double* popk_qp_cpts_dev(pop_handle* h) { return h->qp->cpts; }
int* popk_qp_status_dev(pop_handle* h) { return h->qp->status; }
int* popk_qp_iters_dev(pop_handle* h) { return h->qp->iters; }

extern "C" int pop_qp_solve_batch(pop_handle* h, int n, const int32_t* n_pieces, const double* init,
                                  const double* goal, const double* dur, const double* hpoly, const int32_t* rows,
                                  double vmax, double amax, double* out_cpts, int32_t* out_status,
                                  int32_t* out_iters) {
  if (!h || n <= 0 || !n_pieces || !init || !goal || !dur || !hpoly || !rows || !out_cpts || !out_status)
    return POP_E_INVALID;
  for (int i = 0; i < n; i++) {
    if (n_pieces[i] < 0 || n_pieces[i] > POP_MAX_PIECES) return pop_fail(h, POP_E_CAPACITY, "problem %d has %d pieces (max %d)", i, n_pieces[i], POP_MAX_PIECES);
    for (int k = 0; k < n_pieces[i]; k++)
      if (rows[i * POP_MAX_PIECES + k] < 0 || rows[i * POP_MAX_PIECES + k] > POP_MAX_FACES)
        return pop_fail(h, POP_E_CAPACITY, "problem %d corridor %d has %d faces (max %d)", i, k, rows[i * POP_MAX_PIECES + k], POP_MAX_FACES);
  }
  int rc = popk_qp_alloc(h, n > h->cfg.n_agents_local ? n : h->cfg.n_agents_local);
  if (rc) return rc;
  QpDev* d = h->qp;
  const size_t bh = sizeof(double) * POP_MAX_PIECES * POP_MAX_FACES * 4 * n, br = sizeof(int) * POP_MAX_PIECES * n;
  rc = pop_ensure_scratch(h, bh + br + 64, 0);
  if (rc) return rc;
  double* d_h = static_cast<double*>(h->d_scratch);
  int* d_r = reinterpret_cast<int*>(static_cast<uint8_t*>(h->d_scratch) + bh);
  POP_CUDA(h, cudaMemcpyAsync(d_h, hpoly, bh, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d_r, rows, br, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d->n_pieces, n_pieces, sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d->init, init, sizeof(double) * 9 * n, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d->goal, goal, sizeof(double) * 9 * n, cudaMemcpyHostToDevice, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(d->dur, dur, sizeof(double) * POP_MAX_PIECES * n, cudaMemcpyHostToDevice, h->stream));
  rc = popk_qp_run(h, n, d_h, d_r, vmax, amax);
  if (rc) return rc;
  POP_CUDA(h, cudaMemcpyAsync(out_cpts, d->cpts, sizeof(double) * POP_MAX_PIECES * 15 * n, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaMemcpyAsync(out_status, d->status, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
  if (out_iters) POP_CUDA(h, cudaMemcpyAsync(out_iters, d->iters, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int i = 0; i < n; i++)
    if (out_status[i] == -20)
      return pop_fail(h, POP_E_CAPACITY, "problem %d has more than %d constraint rows (shared-memory capacity)", i, kRowCap);
  return POP_OK;
}
