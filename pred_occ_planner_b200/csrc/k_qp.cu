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
// reduced SPD matrix K = P + sigma I + A' diag(rho) A, banded with half bandwidth 15
// (continuity rows couple neighbouring pieces only).
//
// How one ADMM iteration runs on the SM (all operands on chip, nothing touches L2/HBM):
//   * every thread owns up to kSlots constraint rows for the whole solve: coefficients, bounds,
//     z and y live in REGISTERS; after the z/y update the thread leaves coefficient * w
//     (w = rho z - y) at the row's slots of a column-ordered product array in shared memory;
//   * A'w is then a contiguous sum per column (kG lanes per column + shuffle);
//   * K^-1 is applied as L^-T (L^-1 rhs) with the dense inverse Cholesky factor held in
//     REGISTERS as 6x6 tiles, one tile per thread (n <= 150; larger problems stream the packed
//     factor from shared memory with 16-byte loads): tile x vector products, partial sums
//     exchanged through shared memory, four barriers;
//   * the factor itself (once per rho): banded K assembled in shared memory, warp-synchronous
//     banded Cholesky, then one thread per column of L^-1 by forward substitution with the last
//     15 entries in a register window.
#include <cfloat>

#include "pop_internal.h"

namespace {

constexpr int kT = 384;                       // threads per CTA, shared-memory factor variant (and upper bound)
constexpr int kTt = 384;                      // threads per CTA, register-tile variant (11 warps would be allocated like 12)
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
  // workspaces per agent (set-up only; the iteration never reads them)
  int* ell_col = nullptr;     // [agent][kMaxRows][6]
  double* ell_val = nullptr;  // [agent][kMaxRows][6]
  int* csc_ptr = nullptr;     // [agent][kMaxN+1]
  int* csc_idx = nullptr;     // [agent][kMaxRows*6]  → row*6+slot
  unsigned short* dst = nullptr;  // [agent][kMaxRows*6] position of (row,slot) in column order
  double* Pblk = nullptr;     // [agent][POP_MAX_PIECES][225]
  double* E = nullptr;        // [agent][kMaxRows]
  double* lo = nullptr;       // [agent][kMaxRows] scaled lower bounds
  double* up = nullptr;       // [agent][kMaxRows] scaled upper bounds
  double* QM = nullptr;       // [225] jerk cost block
  // outputs
  double* cpts = nullptr;     // [agent][POP_MAX_PIECES*5][3]
  int* status = nullptr;      // [agent]
  int* iters = nullptr;       // [agent]
  long long* prof = nullptr;  // [agent][8] SM cycles: setup, factorisations, ADMM loop, #factorisations,
                              //            and inside the loop: gather, KKT solve, row update, (unused)
};

namespace {

__device__ __forceinline__ double limit_scaling(double v) {
  v = v < MIN_SCALING ? 1.0 : v;
  v = v > MAX_SCALING ? MAX_SCALING : v;
  return v;
}

// Row capacity: every thread owns up to kSlots rows (row = slot*kT + tid); more rows (or a
// problem that does not fit the shared-memory plan below) make the solve report status -20
// (host: POP_E_CAPACITY).
constexpr int kSlotsMax = 6;            // shared-memory factor path (2 304 rows); the register-tile path owns 3
constexpr int kRowCap = kT * kSlotsMax;
constexpr int kMaxStruct = 9 * (POP_MAX_PIECES + 1) + 21 * POP_MAX_PIECES;  // 369 structured rows < kT
constexpr int kSmemBytes = 232448;   // 227 KB opt-in maximum per CTA on sm_100a
constexpr int kRedN = 12;
constexpr int kVec = kMaxN + 6;      // vectors padded so 6-wide tile loads stay in bounds
constexpr int kTileNB = 25;          // register-tile path: n <= 150 → <= 325 tiles of 6x6
// Shared memory of the register-tile variant: header + staging / partial sums (25 x 150) + Kb + products
// + types for n <= 150, m <= 3 kT.  Leaving the rest of the 256 KB to L1 keeps the set-up arrays
// and the few spilled row constants close.
constexpr int kSmemTiles = 136 * 1024;

// Shared-memory plan.  Fixed header, then per problem (n variables, m rows, nnz non-zeros of A):
//   Linv  ~n(n+1)/2 doubles packed lower-triangular inverse Cholesky factor, rows padded to an
//                           even length (row i at linv_off(i)); on the register-tile path it is
//                           only a staging area and the iteration reuses it for partial sums
//   Kb    16 n doubles      banded K / its Cholesky factor, Kb[c][d] = K[c+d][c]
//   pc    <= nnz+16n+2      coefficient * u products in column order; column c starts at cptr[c],
//                           padded so that start = 2c (mod 16): the two lanes per column of a
//                           warp's gather then hit all 16 eight-byte bank pairs evenly; the
//                           last slot absorbs the stores of unused row slots
//   type  m s8              constraint type per row (-1 loose, 0 inequality, 1 equality)
struct Smem {
  double x[kVec], xp[kVec], xt[kVec], rhs[kVec], tn[kVec], D[kVec], rd[kVec];
  double red[(kT / 32) * kRedN];
  double t[POP_MAX_PIECES];
  double rho, c_scale;
  double rho_v[3], rinv_v[3];   // indexed by type+1: loose, inequality, equality
  int face_off[POP_MAX_PIECES + 1];
  int M, n, m, ms, nnz, status, fail;
  unsigned short cptr[kMaxN + 2];   // start of column c in pc
  unsigned short clen[kMaxN + 2];   // entries of column c
  unsigned short cstr[kMaxN + 2];   // of which in structured rows (set-up only)
  unsigned short rounds[POP_MAX_PIECES];   // gather rounds of piece i: columns padded to 8 * rounds entries
  int pc_len, max_clen;
  long long dbg[4];   // POP_QP_PHASE_PROF: cycles of K assembly, Cholesky, substitution + tiles
};
constexpr int kHdrBytes = (sizeof(Smem) + 15) / 16 * 16;

// warp-wide max of a NON-NEGATIVE double: such values order like their 64-bit patterns, so two
// 32-bit REDUX instructions (high words, then low words among the lanes holding the top high word)
// replace five shuffle + max rounds
__device__ __forceinline__ double warp_max_nonneg(double v) {
  const unsigned hi = static_cast<unsigned>(__double2hiint(v)), lo = static_cast<unsigned>(__double2loint(v));
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return __hiloint2double(static_cast<int>(mh), static_cast<int>(ml));
}

// block-wide max of N per-thread NON-NEGATIVE values at once (all threads receive the results):
// REDUX inside the warps, one shared-memory stage reduced by the first N lanes of warp 0
template <int N>
__device__ void block_max_n(double (&v)[N], double* red) {
  const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < N; k++) v[k] = warp_max_nonneg(v[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; k++) red[w * N + k] = v[k];
  }
  __syncthreads();
  if (w == 0 && lane < N) {
    double r = red[lane];
    for (int q = 1; q < nw; q++) r = fmax(r, red[q * N + lane]);
    red[lane] = r;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; k++) v[k] = red[k];
}
__device__ double block_max(double v, double* red) {
  double a[1] = {v};
  block_max_n<1>(a, red);
  return a[0];
}
// deterministic block-wide sum (fixed tree)
__device__ double block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0;
  for (int w = 0; w < (blockDim.x >> 5); w++) r += red[w];
  return r;
}

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

// Packed storage of the lower-triangular L^-1: rows 2q and 2q+1 both take 2q+2 doubles (row 2q
// ends with one zero pad), so every row starts on a 16-byte boundary.
__device__ __forceinline__ int linv_off(int i) {
  const int q = i >> 1;
  return (i & 1) ? 2 * (q + 1) * (q + 1) : 2 * q * (q + 1);
}

struct TileId { int rb, cb, NB; bool act; };

// K = P + sigma I + A' diag(rho) A → banded Cholesky → L^-1, all in shared memory.
//   1. thread per band entry: P block + safety rows (same control point only) + structured rows
//      through the column lists; coefficients are read from the set-up arrays in global memory
//      (L2), a few dozen independent loads per entry;
//   2. warp 0: right-looking banded Cholesky, one pivot per step (sqrt, scale the column, rank-1
//      update of the 15x15 trailing window), __syncwarp between the three sub-steps;
//   3. thread per column c of L^-1: forward substitution x_r = (e_c - sum_k L[r][k] x_k) / L[r][r]
//      with the last 15 x in a rotating register window; all columns walk r together so the
//      factor rows are shared-memory broadcasts.  TILES: every 6 rows the block of L^-1 is
//      handed through a 6 x 6NB staging buffer to the threads whose register tile it fills;
//      otherwise the rows go to the packed factor in shared memory.
template <bool TILES, int NT>
__device__ void factorize(Smem& S, double* __restrict__ Linv, double (&T)[NT], const TileId& ti, double* __restrict__ Kb,
                          const signed char* __restrict__ type, const int* __restrict__ ell_col,
                          const double* __restrict__ ell_val, const int* __restrict__ csc_ptr,
                          const int* __restrict__ csc_idx, const double* __restrict__ Pblk) {
  const int n = S.n, ms = S.ms;
  const int tid = threadIdx.x;
#ifdef POP_QP_PHASE_PROF
  const long long tf0 = clock64();
#endif
  for (int e = tid; e < n * kBand; e += blockDim.x) {
    const int c = e / kBand, d = e % kBand, r = c + d;
    double v = 0.0;
    if (r < n) {
      if (r / 15 == c / 15) v = Pblk[(c / 15) * 225 + (r % 15) * 15 + (c % 15)];
      if (d == 0) v += SIGMA;
      // rows containing column r in ascending order: structured rows first, then (only when r
      // and c belong to the same control point) the safety rows of that piece
      const int q0 = csc_ptr[r], q1 = csc_ptr[r + 1];
      const bool same_pt = r / 3 == c / 3;
      for (int q = q0; q < q1; q++) {
        const int idx = csc_idx[q], row = idx / kEll;
        if (row >= ms) {
          if (!same_pt) break;
          const double rho = S.rho_v[type[row] + 1];
          v += (rho * ell_val[idx]) * ell_val[idx - (r - c)];   // slots of a safety row are x,y,z
          continue;
        }
        const int* rc = ell_col + row * kEll;
        for (int k = 0; k < kEll; k++)
          if (rc[k] == c) {
            const double rho = S.rho_v[type[row] + 1];
            v += (rho * ell_val[idx]) * ell_val[row * kEll + k];
            break;
          }
      }
    }
    Kb[e] = v;
  }
  __syncthreads();
#ifdef POP_QP_PHASE_PROF
  const long long tf1 = clock64();
#endif
  if (tid < 32) {
    const int lane = tid;
    int udi[4], udk[4];   // trailing-window entries of this lane: (di, dk), 1 <= dk <= di <= 15
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int e = lane + 32 * u;
      int di = 1, base = 0;
      while (base + di <= e) { base += di; di++; }
      udi[u] = e < 120 ? di : 0;
      udk[u] = e - base + 1;
    }
    for (int j = 0; j < n; j++) {
      // pivot
      const double dj = Kb[j * kBand];
      if (!(dj > 0.0)) { if (lane == 0) S.fail = 1; }
      const double lj = sqrt(dj > 0.0 ? dj : 1.0);
      const double inv = 1.0 / lj;
      __syncwarp();
      if (lane == 0) { Kb[j * kBand] = lj; S.rd[j] = inv; }
      if (lane >= 1 && lane < kBand && j + lane < n) Kb[j * kBand + lane] = Kb[j * kBand + lane] / lj;
      __syncwarp();
      // trailing update: a[i][k] -= l[i][j] l[k][j] for j < k <= i <= j+15 (120 entries, 4 per lane)
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int di = udi[u], dk = udk[u];
        if (di > 0 && j + di < n) Kb[(j + dk) * kBand + (di - dk)] -= Kb[j * kBand + di] * Kb[j * kBand + dk];
      }
      __syncwarp();
    }
  }
  __syncthreads();
#ifdef POP_QP_PHASE_PROF
  const long long tf2 = clock64();
#endif
  {
    const int c = tid;
    const int nst = 6 * ti.NB;   // staging row length (TILES)
    double w[15];   // w[k] = x_{r-1-k}
#pragma unroll
    for (int k = 0; k < 15; k++) w[k] = 0.0;
    if constexpr (TILES) {
#pragma unroll
      for (int k = 0; k < NT; k++) T[k] = 0.0;
    }
    const int r_end = TILES ? nst : n;
    for (int r = 0; r < r_end; r++) {
      double xr = 0.0;
      if (c < n && r < n && r >= c) {
        // L[r][r-1-k] = Kb[(r-1-k)*16 + (k+1)]; entries before the column start multiply w = 0
        double s0 = r == c ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < 15; k += 2) {
          s0 -= Kb[max(r - 1 - k, 0) * kBand + (k + 1)] * w[k];
          if (k + 1 < 15) s1 -= Kb[max(r - 2 - k, 0) * kBand + (k + 2)] * w[k + 1];
        }
        xr = (s0 + s1) * S.rd[r];
      }
#pragma unroll
      for (int k = 14; k > 0; k--) w[k] = w[k - 1];
      w[0] = xr;
      if constexpr (TILES) {
        // K^-1 = L^-T L^-1 = sum over rows r of (row r of L^-1)' (row r of L^-1): every staged block
        // of 6 rows adds its outer products to the register tile (rb, cb) it reaches (6b >= 6rb)
        if (c < nst) Linv[(r % 6) * nst + c] = xr;
        if (r % 6 == 5) {
          __syncthreads();
          if (ti.act && ti.rb <= r / 6) {
#pragma unroll
            for (int rr = 0; rr < 6; rr++) {
              const double2* pa = reinterpret_cast<const double2*>(Linv + rr * nst + 6 * ti.rb);
              const double2* pb = reinterpret_cast<const double2*>(Linv + rr * nst + 6 * ti.cb);
              const double2 a01 = pa[0], a23 = pa[1], a45 = pa[2], b01 = pb[0], b23 = pb[1], b45 = pb[2];
              const double av[6] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y};
              const double bv[6] = {b01.x, b01.y, b23.x, b23.y, b45.x, b45.y};
#pragma unroll
              for (int i = 0; i < 6; i++)
#pragma unroll
                for (int j = 0; j < 6; j++) T[i * 6 + j] = fma(av[i], bv[j], T[i * 6 + j]);
            }
          }
          __syncthreads();
        }
      } else {
        if (c < n && r < n && r >= c) Linv[linv_off(r) + c] = xr;
      }
    }
  }
  __syncthreads();
#ifdef POP_QP_PHASE_PROF
  if (tid == 0) { S.dbg[0] += tf1 - tf0; S.dbg[1] += tf2 - tf1; S.dbg[2] += clock64() - tf2; }
#endif
}

// ---- K^-1 application, register-tile path (n <= 150) -----------------------------------------
// Thread t < NB(NB+1)/2 holds the 6x6 tile (rb, cb), cb <= rb, of the SYMMETRIC K^-1 (NB = ceil(n/6))
// and uses it for both triangles in one pass.  Partial sums go to ONE dense array P[NB][Lp]:
//   P[cb][6 rb + i] = tile[i][:] . rhs[6 cb ..]        (row side)
//   P[rb][6 cb + j] = tile[:][j] . rhs[6 rb ..]        (column side, off-diagonal tiles only)
// Row k of P therefore holds, for every output i in block b, the contribution of column block k (k <= b: a row
// side, k > b: a column side): every entry is written exactly once per application — nothing to zero, nothing to
// mask — and xt[i] = sum_k P[k][i].  Each side is six contiguous doubles (three 16-byte stores); Lp = 2 (mod 4)
// makes the lanes of a quarter-warp (consecutive cb, stride Lp or 6 doubles) fall on distinct 16-byte bank groups,
// and the reduction reads 32 consecutive doubles per warp and k.  Then x = alpha xt + (1 - alpha) x_prev.
// FMA is fine here: bit parity of this solve with a CPU factorisation is not attainable anyway.
template <int NT>
__device__ __forceinline__ void kkt_tiles(Smem& S, const double (&T)[NT], const TileId& ti,
                                          double* __restrict__ P, int Lp) {
  const int n = S.n, tid = threadIdx.x;
  if (ti.act) {
    const double2* xc = reinterpret_cast<const double2*>(S.rhs + 6 * ti.cb);
    const double2 c01 = xc[0], c23 = xc[1], c45 = xc[2];
    double s[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
      double a = T[i * 6] * c01.x;
      a = fma(T[i * 6 + 1], c01.y, a);
      a = fma(T[i * 6 + 2], c23.x, a);
      a = fma(T[i * 6 + 3], c23.y, a);
      a = fma(T[i * 6 + 4], c45.x, a);
      a = fma(T[i * 6 + 5], c45.y, a);
      s[i] = a;
    }
    double2* outA = reinterpret_cast<double2*>(P + ti.cb * Lp + 6 * ti.rb);
    outA[0] = make_double2(s[0], s[1]); outA[1] = make_double2(s[2], s[3]); outA[2] = make_double2(s[4], s[5]);
    if (ti.rb != ti.cb) {
      const double2* xr = reinterpret_cast<const double2*>(S.rhs + 6 * ti.rb);
      const double2 r01 = xr[0], r23 = xr[1], r45 = xr[2];
      double o[6];
#pragma unroll
      for (int j = 0; j < 6; j++) {
        double a = T[j] * r01.x;
        a = fma(T[6 + j], r01.y, a);
        a = fma(T[12 + j], r23.x, a);
        a = fma(T[18 + j], r23.y, a);
        a = fma(T[24 + j], r45.x, a);
        a = fma(T[30 + j], r45.y, a);
        o[j] = a;
      }
      double2* outB = reinterpret_cast<double2*>(P + ti.rb * Lp + 6 * ti.cb);
      outB[0] = make_double2(o[0], o[1]); outB[1] = make_double2(o[2], o[3]); outB[2] = make_double2(o[4], o[5]);
    }
  }
  __syncthreads();
  if (tid < n) {
    // rows NB .. 4 ceil(NB / 4) of P are zero (never written by a tile): a uniform loop of four independent
    // loads and adds per round, no predicates
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const double* pr = P + tid;
    const int rounds = (ti.NB + 3) >> 2, step = 4 * Lp;
    for (int k = 0; k < rounds; k++) {
      const double v0 = pr[0], v1 = pr[Lp], v2 = pr[2 * Lp], v3 = pr[3 * Lp];
      pr += step;
      a0 += v0; a1 += v1; a2 += v2; a3 += v3;
    }
    const double xt = (a0 + a1) + (a2 + a3);
    S.xt[tid] = xt;
    S.x[tid] = ALPHA * xt + (1.0 - ALPHA) * S.xp[tid];
  }
  __syncthreads();
}

// ---- K^-1 application, shared-memory path (n > 150) -------------------------------------------
// Both triangular mat-vecs run on groups of 8 lanes with 16-byte shared-memory loads:
//   forward:  a group owns row blocks {2b, 2b+1} for b = g and b = NB-1-g (balanced), lanes stride
//             the 2-column chunks, each rhs chunk load feeds both rows;
//   backward: a group owns column pairs {2b, 2b+1} for the same two b, lanes stride the rows, one
//             load brings L^-1[i][2b..2b+1].
__device__ __forceinline__ void kkt_solve(Smem& S, const double* __restrict__ Linv) {
  const int n = S.n, tid = threadIdx.x;
  const int l = tid & 7, g = tid >> 3;
  const int NB = (n + 1) >> 1;
  const int b1 = g, b2 = NB - 1 - g;
  const bool act = b1 <= b2, two = b1 < b2;
  const double2* rhs2 = reinterpret_cast<const double2*>(S.rhs);
  double s00 = 0, s01 = 0, s10 = 0, s11 = 0;   // rows 2b1, 2b1+1, 2b2, 2b2+1
  if (act) {
    {
      const double2* r0 = reinterpret_cast<const double2*>(Linv + linv_off(2 * b1));
      const double2* r1 = reinterpret_cast<const double2*>(Linv + linv_off(2 * b1 + 1));
      double e0 = 0, e1 = 0;
      for (int k = l; k <= b1; k += 8) {
        const double2 x = rhs2[k], u = r0[k], v = r1[k];
        s00 = fma(u.x, x.x, s00); e0 = fma(u.y, x.y, e0);
        s01 = fma(v.x, x.x, s01); e1 = fma(v.y, x.y, e1);
      }
      s00 += e0; s01 += e1;
    }
    if (two) {
      const double2* r0 = reinterpret_cast<const double2*>(Linv + linv_off(2 * b2));
      const double2* r1 = reinterpret_cast<const double2*>(Linv + linv_off(2 * b2 + 1));
      double e0 = 0, e1 = 0;
      for (int k = l; k <= b2; k += 8) {
        const double2 x = rhs2[k], u = r0[k], v = r1[k];
        s10 = fma(u.x, x.x, s10); e0 = fma(u.y, x.y, e0);
        s11 = fma(v.x, x.x, s11); e1 = fma(v.y, x.y, e1);
      }
      s10 += e0; s11 += e1;
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    s00 += __shfl_xor_sync(0xffffffffu, s00, o);
    s01 += __shfl_xor_sync(0xffffffffu, s01, o);
    s10 += __shfl_xor_sync(0xffffffffu, s10, o);
    s11 += __shfl_xor_sync(0xffffffffu, s11, o);
  }
  if (act && l == 0) {
    S.tn[2 * b1] = s00;
    if (2 * b1 + 1 < n) S.tn[2 * b1 + 1] = s01;
    if (two) { S.tn[2 * b2] = s10; if (2 * b2 + 1 < n) S.tn[2 * b2 + 1] = s11; }
  }
  __syncthreads();
  s00 = s01 = s10 = s11 = 0;
  if (act) {
    {
      const int c0 = 2 * b1;
      int i = c0 + l, off = linv_off(i) + c0;
      double e0 = 0, e1 = 0;
      for (; i + 8 < n; i += 16) {
        const double2 u = *reinterpret_cast<const double2*>(Linv + off);
        const double t = S.tn[i];
        off += 8 * i + 40;
        const double2 v = *reinterpret_cast<const double2*>(Linv + off);
        const double t2 = S.tn[i + 8];
        off += 8 * (i + 8) + 40;
        s00 = fma(u.x, t, s00); s01 = fma(u.y, t, s01);
        e0 = fma(v.x, t2, e0); e1 = fma(v.y, t2, e1);
      }
      if (i < n) {
        const double2 u = *reinterpret_cast<const double2*>(Linv + off);
        const double t = S.tn[i];
        s00 = fma(u.x, t, s00); s01 = fma(u.y, t, s01);
      }
      s00 += e0; s01 += e1;
    }
    if (two) {
      const int c0 = 2 * b2;
      int i = c0 + l, off = linv_off(i) + c0;
      double e0 = 0, e1 = 0;
      for (; i + 8 < n; i += 16) {
        const double2 u = *reinterpret_cast<const double2*>(Linv + off);
        const double t = S.tn[i];
        off += 8 * i + 40;
        const double2 v = *reinterpret_cast<const double2*>(Linv + off);
        const double t2 = S.tn[i + 8];
        off += 8 * (i + 8) + 40;
        s10 = fma(u.x, t, s10); s11 = fma(u.y, t, s11);
        e0 = fma(v.x, t2, e0); e1 = fma(v.y, t2, e1);
      }
      if (i < n) {
        const double2 u = *reinterpret_cast<const double2*>(Linv + off);
        const double t = S.tn[i];
        s10 = fma(u.x, t, s10); s11 = fma(u.y, t, s11);
      }
      s10 += e0; s11 += e1;
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    s00 += __shfl_xor_sync(0xffffffffu, s00, o);
    s01 += __shfl_xor_sync(0xffffffffu, s01, o);
    s10 += __shfl_xor_sync(0xffffffffu, s10, o);
    s11 += __shfl_xor_sync(0xffffffffu, s11, o);
  }
  if (act && l == 0) {
    int c = 2 * b1;
    S.xt[c] = s00; S.x[c] = ALPHA * s00 + (1.0 - ALPHA) * S.xp[c];
    if (c + 1 < n) { S.xt[c + 1] = s01; S.x[c + 1] = ALPHA * s01 + (1.0 - ALPHA) * S.xp[c + 1]; }
    if (two) {
      c = 2 * b2;
      S.xt[c] = s10; S.x[c] = ALPHA * s10 + (1.0 - ALPHA) * S.xp[c];
      if (c + 1 < n) { S.xt[c + 1] = s11; S.x[c + 1] = ALPHA * s11 + (1.0 - ALPHA) * S.xp[c + 1]; }
    }
  }
  __syncthreads();
}

// (A' u)_c = sum of the column-c segment of pc.  Warp i gathers the 15 columns of piece i, two lanes per column
// (lane h takes entries h, h + 2, ... with four accumulators, then one shuffle).  All columns of a piece are padded
// with never-written zero slots to the same length 8 R_i (+2 slots of skew, see the set-up), so the loop is
// warp-uniform, unpredicated and unmasked: R_i rounds of four independent loads and adds per lane.  Fixed order.
// Every thread of the CTA calls this; the even lane of a column gets the sum.
__device__ __forceinline__ double gather_col(const double* __restrict__ pq, int rounds) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int r = 0; r < rounds; r++) {
    const double v0 = pq[0], v1 = pq[2], v2 = pq[4], v3 = pq[6];
    pq += 8;
    s0 += v0; s1 += v1; s2 += v2; s3 += v3;
  }
  double s = (s0 + s1) + (s2 + s3);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  return s;
}

// TILES = true: problems with n <= 150 and m <= 3 kT rows (L^-1 in registers, 3 row slots per
// thread); TILES = false: the rest (L^-1 streamed from shared memory, 6 row slots).  Both variants
// are launched over the whole batch; each CTA returns at once if the problem is not its kind.
template <bool TILES>
__global__ void __launch_bounds__(TILES ? kTt : kT, 1) k_qp_admm(QpDev Dv, QpParams P, const double* __restrict__ hpoly_all,
                                                                 const int* __restrict__ rows_all, int a0) {
  constexpr int kSlots = TILES ? 3 : kSlotsMax;
  constexpr int KT = TILES ? kTt : kT;   // threads of this variant
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int a = blockIdx.x + a0, tid = threadIdx.x;
  const long long clk0 = clock64();
  long long clk_fact = 0;
  int n_fact = 0;
  const int M = Dv.n_pieces[a];
  long long clk_g = 0, clk_k = 0, clk_r = 0, clk_e = 0, clk_e1 = 0, clk_e2 = 0;
  if (M <= 0 || M > POP_MAX_PIECES) {
    if (TILES && tid == 0) { Dv.status[a] = 0; Dv.iters[a] = 0; }
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
  unsigned short* dstv = Dv.dst + static_cast<size_t>(a) * kMaxRows * kEll;
  double* Pblk = Dv.Pblk + static_cast<size_t>(a) * POP_MAX_PIECES * 225;
  double* E = Dv.E + static_cast<size_t>(a) * kMaxRows;
  double* lov = Dv.lo + static_cast<size_t>(a) * kMaxRows;
  double* upv = Dv.up + static_cast<size_t>(a) * kMaxRows;

  if (tid == 0) {
    S.M = M; S.n = 15 * M;
    S.face_off[0] = 0;
    for (int i = 0; i < M; i++) {
      int r = rows[i];
      r = r < 0 ? 0 : (r > POP_MAX_FACES ? POP_MAX_FACES : r);
      S.face_off[i + 1] = S.face_off[i] + r;
      S.t[i] = Dv.dur[a * POP_MAX_PIECES + i];
    }
    S.ms = 9 * (M + 1) + 21 * M;
    S.m = S.ms + 5 * S.face_off[M];
    // non-zeros of A: pos 2 + vel 4 + acc 6 per join and dimension (1 + 2 + 3 at each end), velocity
    // boxes 2, acceleration boxes 3, safety rows 3
    S.nnz = 87 * M + 15 * S.face_off[M];   // 36 per join/end pair + 24 + 27 box entries per piece
    S.status = 0; S.fail = 0; S.rho = RHO0; S.c_scale = 1.0;
    S.dbg[0] = S.dbg[1] = S.dbg[2] = S.dbg[3] = 0;
    S.rho_v[0] = RHO_MIN; S.rho_v[1] = RHO0; S.rho_v[2] = RHO_EQ * RHO0;
    for (int k = 0; k < 3; k++) S.rinv_v[k] = 1. / S.rho_v[k];
  }
  __syncthreads();
  const int n = S.n, m = S.m, ms = S.ms, nnz = S.nnz;
  if (((n + 5) / 6 <= kTileNB && m <= 3 * kTt) != TILES) return;   // the other variant's problem
  const int NBt = (n + 5) / 6;
  // TILES: staging rows (6 x 6NB) / partial sums P[NB][Lp] instead of the packed factor
  int Lpt = 6 * NBt;
  if ((Lpt & 3) != 2) Lpt += 2;   // Lp = 2 (mod 4): see kkt_tiles
  // TILES: [staging 36 NB | Kb 16 n] during a factorisation, P during the iterations
  const int linv_len = TILES ? max(36 * NBt + n * kBand, ((NBt + 3) & ~3) * Lpt) : linv_off((n + 1) & ~1);
  // the banded K lives inside the staging area (TILES) or on top of the product array (it is dead
  // during a factorisation and rewritten right after), so it costs no shared memory of its own
  // product array: every column of piece i takes 8 R_i + 2 slots, R_i = ceil(longest column of the piece / 8); a column
  // holds one entry per face of its piece plus at most 11 structured entries
  const int pc_len_max = max(15 * (S.face_off[M] + 20 * M) + 8, TILES ? 0 : n * kBand);
  const size_t dyn_bytes = sizeof(double) * (static_cast<size_t>(linv_len) + pc_len_max) +
                           ((m + 15) & ~15) + 16 + (TILES ? sizeof(double) * 2 * static_cast<size_t>(m) : 0);
  if (m > kRowCap || kHdrBytes + dyn_bytes > (TILES ? kSmemTiles : kSmemBytes)) {
    if (tid == 0) { Dv.status[a] = -20; Dv.iters[a] = 0; }
    return;
  }
  double* Linv = reinterpret_cast<double*>(smem_raw + kHdrBytes);
  double* pc = Linv + linv_len;
  double* Kb = TILES ? Linv + 36 * NBt : pc;
  signed char* type = reinterpret_cast<signed char*>(pc + pc_len_max);
  // tile variant: row scalings and their reciprocals stay in shared memory for
  // the termination / rho-adaptation events (the other variant reads them from global memory)
  double* Es = reinterpret_cast<double*>(type + ((m + 15) & ~15));
  double* Eis = Es + m;

  // ---- assembly ----
  for (int r = tid; r < m; r += KT) {
    int col[kEll];
    double val[kEll], lo, up;
    build_row(S, r, init, goal, hpoly, rows, P, col, val, &lo, &up);
    for (int k = 0; k < kEll; k++) { ell_col[r * kEll + k] = col[k]; ell_val[r * kEll + k] = val[k]; }
    lov[r] = fmax(lo, -OSQP_INFTY);
    upv[r] = fmin(up, OSQP_INFTY);
    E[r] = 1.0;
  }
  for (int e = tid; e < M * 225; e += KT) Pblk[e] = Dv.QM[e % 225];
  for (int e = tid; e < linv_len; e += KT) Linv[e] = 0.0;   // row pads stay zero
  for (int e = tid; e < pc_len_max; e += KT) pc[e] = 0.0;   // the pad slots of the columns stay zero for the whole solve
  if (tid < kVec) { S.D[tid] = 1.0; S.x[tid] = 0.0; S.xp[tid] = 0.0; S.xt[tid] = 0.0; S.rhs[tid] = 0.0; S.tn[tid] = 0.0; S.rd[tid] = 0.0; }
  __syncthreads();
  // column index: entries of column j in ascending row order.  A column holds its structured entries (rows < ms:
  // boundary / continuity rows and the velocity and acceleration boxes) first, then one entry per face of its piece
  // (safety row of face j, control point k, dimension d sits in column 15 i + 3 k + d), so only the structured rows
  // need a scan: their column numbers are staged in shared memory (the product array is still free) and every
  // column thread walks that list with broadcast loads; the safety entries are placed in parallel by formula.
  {
    int* scol = reinterpret_cast<int*>(pc);   // [ms][6]
    for (int e = tid; e < ms * kEll; e += KT) scol[e] = ell_col[e];
    if (tid == 0) csc_ptr[0] = 0;
    __syncthreads();
    if (tid < n) {
      int cnt = 0;
      for (int e = 0; e < ms * kEll; e++) cnt += scol[e] == tid;
      S.cstr[tid] = static_cast<unsigned short>(cnt);
      const int i = tid / 15;
      S.clen[tid] = static_cast<unsigned short>(cnt + S.face_off[i + 1] - S.face_off[i]);
    }
    __syncthreads();
    if (tid == 0) {
      // Column c of piece i starts at base_i + (c - 15 i) S_i with S_i = 8 R_i + 2: S_i = 2 (mod 8) spreads the 30 lanes
      // of the gathering warp (entry h + 2k of 15 consecutive columns) evenly over the 16 eight-byte bank pairs.
      int run = 0, prun = 0, mx = 0;
      for (int i = 0; i < M; i++) {
        int L = 0;
        for (int j = 15 * i; j < 15 * i + 15; j++) L = max(L, static_cast<int>(S.clen[j]));
        const int R = (L + 7) >> 3, Si = 8 * R + 2;
        S.rounds[i] = static_cast<unsigned short>(R);
        for (int j = 0; j < 15; j++) S.cptr[15 * i + j] = static_cast<unsigned short>(prun + j * Si);
        prun += 15 * Si;
        mx = max(mx, L);
      }
      for (int j = 0; j < n; j++) { run += S.clen[j]; csc_ptr[j + 1] = run; }
      S.pc_len = prun;
      S.max_clen = mx;
      if (run != nnz || prun + 1 > pc_len_max) S.fail = 2;
    }
    __syncthreads();
    if (tid < n) {
      int w = csc_ptr[tid], pw = S.cptr[tid];
      for (int e = 0; e < ms * kEll; e++)
        if (scol[e] == tid) { dstv[e] = static_cast<unsigned short>(pw++); csc_idx[w++] = e; }
    }
    for (int r = ms + tid; r < m; r += KT) {
      const int q = (r - ms) / 5, k = (r - ms) % 5;
      int i = 0;
      while (i + 1 < M && q >= S.face_off[i + 1]) i++;
      const int j = q - S.face_off[i];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const int c = 15 * i + 3 * k + d;
        dstv[r * kEll + d] = static_cast<unsigned short>(S.cptr[c] + S.cstr[c] + j);
        csc_idx[csc_ptr[c] + S.cstr[c] + j] = r * kEll + d;
      }
    }
    __syncthreads();
    for (int e = tid; e < ms * kEll + 2; e += KT) pc[e >> 1] = 0.0;   // the staging area goes back to the product array
    __syncthreads();
  }
  if (S.fail) {
    if (tid == 0) { Dv.status[a] = -21; Dv.iters[a] = 0; }
    return;
  }

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
    for (int r = tid; r < m; r += KT) {
      double rm = 0;
      for (int k = 0; k < kEll; k++) if (ell_col[r * kEll + k] >= 0) rm = fmax(rm, fabs(ell_val[r * kEll + k]));
      const double et = 1.0 / sqrt(limit_scaling(rm));
      for (int k = 0; k < kEll; k++) {
        const int c = ell_col[r * kEll + k];
        if (c >= 0) ell_val[r * kEll + k] = (et * ell_val[r * kEll + k]) * S.tn[c];
      }
      E[r] *= et;
    }
    for (int e = tid; e < M * 225; e += KT) {
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
    for (int e = tid; e < M * 225; e += KT) Pblk[e] *= ct;
    __syncthreads();
  }
  if (tid < kVec) { S.rhs[tid] = 0.0; S.tn[tid] = 0.0; }   // scratch use above; pads must be zero
  // scaled bounds and constraint types (set_rho_vec)
  for (int r = tid; r < m; r += KT) {
    lov[r] *= E[r];
    upv[r] *= E[r];
    signed char ty;
    if (lov[r] < -OSQP_INFTY * MIN_SCALING && upv[r] > OSQP_INFTY * MIN_SCALING) ty = -1;
    else if (upv[r] - lov[r] < RHO_TOL) ty = 1;
    else ty = 0;
    type[r] = ty;
    if constexpr (TILES) { Es[r] = E[r]; Eis[r] = 1.0 / E[r]; }
  }
  __syncthreads();

  // ---- register tile of L^-1 (n <= 150) ----
  TileId ti;
  ti.NB = (n + 5) / 6;
  constexpr bool tiles = TILES;
  ti.rb = 0; ti.cb = 0;
  {
    int rb = static_cast<int>((sqrtf(8.0f * tid + 1.0f) - 1.0f) * 0.5f);
    while ((rb + 1) * (rb + 2) / 2 <= tid) rb++;
    while (rb * (rb + 1) / 2 > tid) rb--;
    ti.rb = rb; ti.cb = tid - rb * (rb + 1) / 2;
  }
  ti.act = tiles && ti.rb < ti.NB;
  double T[TILES ? 36 : 1];
#pragma unroll
  for (int k = 0; k < (TILES ? 36 : 1); k++) T[k] = 0.0;
  auto refactor = [&]() {
    const long long c0k = clock64();
    factorize<TILES>(S, Linv, T, ti, Kb, type, ell_col, ell_val, csc_ptr, csc_idx, Pblk);
    if constexpr (TILES) {   // the staging area becomes P again: its pad rows (NB .. 4 ceil(NB / 4)) must read as zero
      for (int e = NBt * Lpt + tid; e < ((NBt + 3) & ~3) * Lpt; e += KT) Linv[e] = 0.0;
    } else {                 // the banded K sat on the product array: its pad slots must read as zero again
      for (int e = tid; e < pc_len_max; e += KT) pc[e] = 0.0;
    }
    __syncthreads();
    clk_fact += clock64() - c0k; n_fact++;
  };
  refactor();
  if (S.fail) {
    if (tid == 0) { Dv.status[a] = -7; Dv.iters[a] = 0; }
    return;
  }
  const double cinv = 1.0 / S.c_scale;

  // ---- per-thread row state in registers ----
  // slot 0 (row tid): structured or safety, 6 coefficients with packed 8-bit columns and 16-bit
  // product slots; unused slots point at the scratch entry pc[nnz] with zero coefficients
  double v0[kEll], z0 = 0, y0 = 0, lo0 = -OSQP_INFTY, up0 = OSQP_INFTY;
  unsigned long long c0 = 0;
  unsigned d0[3];
  int ty0 = 0;
  const bool has0 = tid < m;
  {
    unsigned short ds[kEll];
#pragma unroll
    for (int k = 0; k < kEll; k++) { v0[k] = 0.0; ds[k] = static_cast<unsigned short>(S.pc_len); }
    if (has0) {
#pragma unroll
      for (int k = 0; k < kEll; k++) {
        const int c = ell_col[tid * kEll + k];
        if (c >= 0) {
          v0[k] = ell_val[tid * kEll + k];
          c0 |= static_cast<unsigned long long>(c) << (8 * k);
          ds[k] = dstv[tid * kEll + k];
        }
      }
      lo0 = lov[tid]; up0 = upv[tid]; ty0 = type[tid];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) d0[k] = ds[2 * k] | (static_cast<unsigned>(ds[2 * k + 1]) << 16);
  }
  // slots 1.. (rows s*KT + tid): safety rows only — 3 coefficients on consecutive columns, lower
  // bound -OSQP_INFTY*E (never active), type 0
  double vs[kSlots - 1][3], zs[kSlots - 1], ys[kSlots - 1], ups[kSlots - 1];
  int bs[kSlots - 1];
  unsigned long long dss[kSlots - 1];
#pragma unroll
  for (int s = 0; s < kSlots - 1; s++) {
    const int r = (s + 1) * KT + tid;
    zs[s] = ys[s] = 0.0; ups[s] = OSQP_INFTY; bs[s] = 0;
    vs[s][0] = vs[s][1] = vs[s][2] = 0.0;
    dss[s] = static_cast<unsigned long long>(S.pc_len) * 0x000100010001ull;
    if (r < m) {
      bs[s] = ell_col[r * kEll];
      vs[s][0] = ell_val[r * kEll]; vs[s][1] = ell_val[r * kEll + 1]; vs[s][2] = ell_val[r * kEll + 2];
      ups[s] = upv[r];
      dss[s] = static_cast<unsigned long long>(dstv[r * kEll]) | (static_cast<unsigned long long>(dstv[r * kEll + 1]) << 16) |
               (static_cast<unsigned long long>(dstv[r * kEll + 2]) << 32);
    }
  }
#define COL0(k) static_cast<int>((c0 >> (8 * (k))) & 0xffull)
#define DST0(k) static_cast<int>((d0[(k) >> 1] >> (16 * ((k) & 1))) & 0xffffu)
#define DSTS(s, k) static_cast<int>((dss[s] >> (16 * (k))) & 0xffffull)
  auto dot0 = [&](const double* x) {
    double s = 0;
#pragma unroll
    for (int k = 0; k < kEll; k++) s = fma(v0[k], x[COL0(k)], s);   // fused: see kkt_tiles on bit parity
    return s;
  };
  auto dots = [&](int s, const double* x) {
    return fma(vs[s][2], x[bs[s] + 2], fma(vs[s][1], x[bs[s] + 1], vs[s][0] * x[bs[s]]));
  };
  auto put0 = [&](double u) {
#pragma unroll
    for (int k = 0; k < kEll; k++) pc[DST0(k)] = v0[k] * u;
  };
  auto puts = [&](int s, double u) {
#pragma unroll
    for (int k = 0; k < 3; k++) pc[DSTS(s, k)] = vs[s][k] * u;
  };
  // leave u-weighted coefficients in pc for gather_col; sel: 0 → w = rho z - y, 1 → y, 2 → dy
  auto write_p = [&](int sel, const double (&dy)[kSlots]) {
    put0(sel == 0 ? S.rho_v[ty0 + 1] * z0 - y0 : (sel == 1 ? y0 : dy[0]));
#pragma unroll
    for (int s = 0; s < kSlots - 1; s++)
      if ((s + 1) * KT < m) puts(s, sel == 0 ? S.rho_v[1] * zs[s] - ys[s] : (sel == 1 ? ys[s] : dy[s + 1]));
  };

  // gather role: warp i < M takes piece i, lanes 2q and 2q + 1 its column 15 i + q (lanes 30, 31 and the warps beyond M
  // run the same loop on valid memory and drop the result)
  const int gw = tid >> 5, gln = tid & 31;
  const bool gact = gw < M && gln < 30;
  const int gc = gact ? 15 * gw + (gln >> 1) : 0;
  const bool gl0 = gact && (gln & 1) == 0;
  const double* gq = pc + (gact ? S.cptr[gc] + (gln & 1) : 0);
  const int grounds = gw < M ? S.rounds[gw] : 0;

  // ---- ADMM ----
  int iter, status = 0;
  bool can_check = false;
  double pri_res = 0, dua_res = 0;
  double nz = 0, nAx = 0, nAty = 0, nPx = 0, rp_scaled = 0, rd_scaled = 0, nz_un = 0, nd_un = 0;

  // residuals + norms needed by check_termination / compute_rho_estimate (all threads get them).
  // Leaves val*y in pc: the caller restores val*w before the next iteration.
  auto update_info = [&](const double (&dy)[kSlots]) {
    double mx[kRedN];
#pragma unroll
    for (int k = 0; k < kRedN; k++) mx[k] = 0.0;
    auto row_norms = [&](double ax, double z, int r) {
      const double res = ax - z;
      const double einv = TILES ? Eis[r] : 1.0 / E[r];
      mx[0] = fmax(mx[0], fabs(einv * res));
      mx[1] = fmax(mx[1], fabs(res));
      mx[2] = fmax(mx[2], fabs(z));
      mx[3] = fmax(mx[3], fabs(ax));
      mx[4] = fmax(mx[4], fabs(einv * z));
      mx[5] = fmax(mx[5], fabs(einv * ax));
    };
    if (has0) row_norms(dot0(S.x), z0, tid);
#pragma unroll
    for (int s = 0; s < kSlots - 1; s++)
      if ((s + 1) * KT + tid < m) row_norms(dots(s, S.x), zs[s], (s + 1) * KT + tid);
    write_p(1, dy);
    __syncthreads();
    const double aty = gather_col(gq, grounds);
    if (gl0) {
      const int b = gc / 15, rr = gc % 15;
      double px = 0;
      const double* Pb = Pblk + b * 225 + rr * 15;
#pragma unroll
      for (int k = 0; k < 15; k++) px += Pb[k] * S.x[15 * b + k];
      const double res = (0.0 + px) + aty;
      const double dinv = 1.0 / S.D[gc];
      mx[6] = fabs(dinv * res); mx[7] = fabs(res); mx[8] = fabs(aty); mx[9] = fabs(px);
      mx[10] = fabs(dinv * aty); mx[11] = fabs(dinv * px);
    }
    block_max_n<kRedN>(mx, S.red);
    pri_res = mx[0]; rp_scaled = mx[1]; nz = mx[2]; nAx = mx[3];
    dua_res = cinv * mx[6]; rd_scaled = mx[7]; nAty = mx[8]; nPx = mx[9];
    // unscaled norms for the tolerance computation (compute_pri_tol / compute_dua_tol)
    nz_un = fmax(mx[4], mx[5]);
    nd_un = cinv * fmax(fmax(0.0, mx[10]), mx[11]);
  };

  auto primal_infeasible = [&](double eps, double (&dy)[kSlots]) -> bool {
    // is_primal_infeasible: delta_y projected on the polar of the recession cone (kept in place)
    double v_nd = 0, v_lhs = 0;
    if (has0) {
      double d = dy[0];
      if (up0 > OSQP_INFTY * MIN_SCALING) {
        if (lo0 < -OSQP_INFTY * MIN_SCALING) d = 0.0;
        else d = fmin(d, 0.0);
      } else if (lo0 < -OSQP_INFTY * MIN_SCALING) {
        d = fmax(d, 0.0);
      }
      dy[0] = d;
      v_nd = fabs((TILES ? Es[tid] : E[tid]) * d);
      v_lhs = up0 * fmax(d, 0.0) + lo0 * fmin(d, 0.0);
    }
#pragma unroll
    for (int s = 0; s < kSlots - 1; s++)
      if ((s + 1) * KT + tid < m) {   // up finite, lo = -OSQP_INFTY*E: project on d >= 0
        const double d = fmax(dy[s + 1], 0.0);
        dy[s + 1] = d;
        v_nd = fmax(v_nd, fabs((TILES ? Es[(s + 1) * KT + tid] : E[(s + 1) * KT + tid]) * d));
        v_lhs += ups[s] * d;
      }
    const double nd = block_max(v_nd, S.red);
    const double lhs = block_sum(v_lhs, S.red);
    if (nd > eps && lhs < -eps * nd) {
      __syncthreads();
      write_p(2, dy);
      __syncthreads();
      const double s = gather_col(gq, grounds);
      const double v = gl0 ? fabs((1.0 / S.D[gc]) * s) : 0.0;
      const double nrm = block_max(v, S.red);
      return nrm < eps * nd;
    }
    return false;
  };
  auto check_termination = [&](bool approximate, double (&dy)[kSlots]) -> bool {
    double eps_abs = P.eps_abs, eps_rel = P.eps_rel, eps_pi = P.eps_prim_inf;
    if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pi *= 10; }
    if (pri_res > OSQP_INFTY || dua_res > OSQP_INFTY) { status = -7; return true; }
    bool prim_ok = false, dual_ok = false, prim_inf = false;
    const double eps_prim = eps_abs + eps_rel * nz_un;
    if (pri_res < eps_prim) prim_ok = true;
    else prim_inf = primal_infeasible(eps_pi, dy);
    const double eps_dual = eps_abs + eps_rel * nd_un;
    if (dua_res < eps_dual) dual_ok = true;
    if (prim_ok && dual_ok) { status = approximate ? 2 : 1; return true; }
    if (prim_inf) { status = approximate ? 3 : -3; return true; }
    return false;
  };

  {
    const double dy[kSlots] = {0};
    write_p(0, dy);
  }
  __syncthreads();
  auto admm_iter = [&](double (&dy)[kSlots]) {
#ifdef POP_QP_PHASE_PROF
    const long long t0 = clock64();
#endif
    // x_prev <- x ; rhs = sigma x_prev - q + A' w with w = rho z - y
    {
      const double s = gather_col(gq, grounds);
      if (gl0) {
        const double xv = S.x[gc];
        S.xp[gc] = xv;
        S.rhs[gc] = (SIGMA * xv - 0.0) + s;
      }
    }
    __syncthreads();
#ifdef POP_QP_PHASE_PROF
    const long long t1 = clock64();
#endif
    if constexpr (TILES) kkt_tiles(S, T, ti, Linv, Lpt);
    else kkt_solve(S, Linv);
#ifdef POP_QP_PHASE_PROF
    const long long t2 = clock64();
#endif
    // z, y updates of the rows this thread owns (branch-free: unused slots carry zero
    // coefficients and write to the scratch entry); leave val*w for the next gather
    {
      const double rho_in = S.rho_v[1], rinv_in = S.rinv_v[1];
      const double rho = S.rho_v[ty0 + 1], rinv = S.rinv_v[ty0 + 1];
      double zt[kSlots];
      zt[0] = dot0(S.xt);
#pragma unroll
      for (int s = 0; s < kSlots - 1; s++) zt[s + 1] = (s + 1) * KT < m ? dots(s, S.xt) : 0.0;
      {
        const double mix = ALPHA * zt[0] + (1.0 - ALPHA) * z0;
        const double v = mix + rinv * y0;
        const double zn = fmin(fmax(v, lo0), up0);
        const double d = rho * (mix - zn);
        z0 = zn; y0 = y0 + d; dy[0] = d;
        put0(rho * zn - y0);
      }
#pragma unroll
      for (int s = 0; s < kSlots - 1; s++) {
        if ((s + 1) * KT >= m) continue;   // uniform: no row of this problem lives in the slot
        const double mix = ALPHA * zt[s + 1] + (1.0 - ALPHA) * zs[s];
        const double v = mix + rinv_in * ys[s];
        const double zn = fmin(v, ups[s]);
        const double d = rho_in * (mix - zn);
        zs[s] = zn; ys[s] = ys[s] + d; dy[s + 1] = d;
        puts(s, rho_in * zn - ys[s]);
      }
    }
    __syncthreads();
#ifdef POP_QP_PHASE_PROF
    clk_g += t1 - t0; clk_k += t2 - t1; clk_r += clock64() - t2;
#endif
  };
  // Hot iterations run in a tight inner loop up to the next event (termination test, rho
  // adaptation, last iteration); only the event iteration keeps its delta_y.
  iter = 0;
  const long long clk_loop = clock64();
  while (true) {
    int next = P.max_iter;
    if (P.check_interval) next = min(next, (iter / P.check_interval + 1) * P.check_interval);
    if (P.adaptive_interval) next = min(next, (iter / P.adaptive_interval + 1) * P.adaptive_interval);
    for (; iter + 1 < next; iter++) {
      double scratch[kSlots];
      admm_iter(scratch);
    }
    double dy[kSlots] = {0};   // delta_y of this iteration (is_primal_infeasible)
    admm_iter(dy);
    iter++;
#ifdef POP_QP_PHASE_PROF
    const long long te0 = clock64();
    long long te1 = te0, te2 = te0;
#endif
    const bool last = iter >= P.max_iter;
    can_check = P.check_interval && (iter % P.check_interval == 0);
    bool info = false;
    if (can_check || last) {
      update_info(dy);
#ifdef POP_QP_PHASE_PROF
      te1 = clock64();
#endif
      info = true;
      if (check_termination(false, dy)) break;
#ifdef POP_QP_PHASE_PROF
      te2 = clock64();
#endif
      if (last) {   // osqp_solve's closing test with relaxed tolerances, else max-iter
        if (!check_termination(true, dy)) status = -2;
        break;
      }
    }
    bool new_rho = false;
    if (P.adaptive_interval && (iter % P.adaptive_interval == 0)) {
      if (!info) { update_info(dy); info = true; }
      double pr = rp_scaled / (fmax(nz, nAx) + 1e-10);
      double dr = rd_scaled / (fmax(fmax(0.0, nAty), nPx) + 1e-10);
      double rho_new = S.rho * sqrt(pr / (dr + 1e-10));
      rho_new = fmin(fmax(rho_new, RHO_MIN), RHO_MAX);
      if (rho_new > S.rho * ADAPT_TOL || rho_new < S.rho / ADAPT_TOL) {
        __syncthreads();
        if (tid == 0) {
          S.rho = rho_new;
          S.rho_v[1] = rho_new; S.rho_v[2] = RHO_EQ * rho_new;
          S.rinv_v[1] = 1. / S.rho_v[1]; S.rinv_v[2] = 1. / S.rho_v[2];
        }
        __syncthreads();
        new_rho = true;
      }
    }
    if (info) {   // pc holds val*y (or val*dy): restore val*w (with the new rho, if it changed)
      __syncthreads();
      if (new_rho) refactor();
      write_p(0, dy);
      __syncthreads();
    }
#ifdef POP_QP_PHASE_PROF
    clk_e += clock64() - te0; clk_e1 += te1 - te0; clk_e2 += te2 - te1;
#endif
  }
  __syncthreads();
  if (tid < n) Dv.cpts[static_cast<size_t>(a) * POP_MAX_PIECES * 15 + tid] = S.D[tid] * S.x[tid];
  if (tid == 0) {
    Dv.status[a] = status;
    Dv.iters[a] = iter < P.max_iter ? iter : P.max_iter;
    const long long now = clock64();
    Dv.prof[a * 8 + 0] = clk_loop - clk0;   // assembly, scaling, first factorisation, register load
    Dv.prof[a * 8 + 1] = clk_fact;
    Dv.prof[a * 8 + 2] = now - clk_loop;
    Dv.prof[a * 8 + 3] = n_fact;
    Dv.prof[a * 8 + 4] = clk_g; Dv.prof[a * 8 + 5] = clk_k; Dv.prof[a * 8 + 6] = clk_r;
    Dv.prof[a * 8 + 7] = clk_e;
#if defined(POP_QP_PHASE_PROF) && POP_QP_PHASE_PROF == 2
    Dv.prof[a * 8 + 4] = S.dbg[0]; Dv.prof[a * 8 + 5] = S.dbg[1]; Dv.prof[a * 8 + 6] = S.dbg[2];   // factorisation split
#endif
  }
#undef COL0
#undef DST0
#undef DSTS
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
  ok &= cudaMalloc(&d->lo, na * kMaxRows * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->up, na * kMaxRows * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->dst, na * kMaxRows * kEll * sizeof(unsigned short)) == cudaSuccess;
  ok &= cudaMalloc(&d->QM, 225 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->cpts, na * POP_MAX_PIECES * 15 * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&d->status, na * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->iters, na * sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d->prof, na * 8 * sizeof(long long)) == cudaSuccess;
  // as popk_astar_alloc: a failed attempt frees what it got and leaves h->qp null
  auto finish = [&]() -> int {
    if (!ok) { cudaGetLastError(); return pop_fail(h, POP_E_CUDA, "QP workspace allocation failed"); }
    double QM[225];
    host_jerk_block(QM);
    POP_CUDA(h, cudaMemcpyAsync(d->QM, QM, sizeof(QM), cudaMemcpyHostToDevice, h->stream));
    POP_CUDA(h, cudaMemsetAsync(d->cpts, 0, na * POP_MAX_PIECES * 15 * sizeof(double), h->stream));
    POP_CUDA(h, cudaStreamSynchronize(h->stream));
    POP_CUDA(h, cudaFuncSetAttribute(k_qp_admm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTiles));
    POP_CUDA(h, cudaFuncSetAttribute(k_qp_admm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    return POP_OK;
  };
  const int rc = finish();
  if (rc != POP_OK) { popk_qp_free(h); return rc; }
  return POP_OK;
}

void popk_qp_free(pop_handle* h) {
  QpDev* d = h->qp;
  if (!d) return;
  void* ptrs[] = {d->n_pieces, d->init, d->goal, d->dur, d->ell_col, d->ell_val, d->csc_ptr, d->csc_idx, d->Pblk,
                  d->E, d->lo, d->up, d->dst, d->QM, d->cpts, d->status, d->iters, d->prof};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete d;
  h->qp = nullptr;
}

// inputs resident in qp->{n_pieces, init, goal, dur}; corridors at d_hpoly / d_rows
int popk_qp_run(pop_handle* h, int n_agents, const double* d_hpoly, const int* d_rows, double vmax, double amax,
                cudaStream_t st, int a0) {
  if (!st) st = h->stream;
  QpDev* d = h->qp;
  const pop_config& c = h->cfg;
  QpParams P;
  P.max_iter = c.osqp_max_iter; P.adaptive_interval = c.osqp_adaptive_rho_interval;
  P.check_interval = c.osqp_check_termination; P.scaling_iters = 10;
  P.eps_abs = c.osqp_eps_abs; P.eps_rel = c.osqp_eps_rel; P.eps_prim_inf = 1e-4;
  P.vmax = vmax; P.amax = amax;
  pop_prof_begin(h, POP_STAGE_QP);
  k_qp_admm<true><<<n_agents, kTt, kSmemTiles, st>>>(*d, P, d_hpoly, d_rows, a0);
  POP_LAUNCH_CHECK(h);
  static const bool skip_big = getenv("POP_QP_SKIP_BIG") != nullptr;   // experiment switch
  if (!skip_big) {
    k_qp_admm<false><<<n_agents, kT, kSmemBytes, st>>>(*d, P, d_hpoly, d_rows, a0);
    POP_LAUNCH_CHECK(h);
  }
  pop_prof_end(h, POP_STAGE_QP, 2);
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
  PopDeviceGuard guard(h);
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
      return pop_fail(h, POP_E_CAPACITY, "problem %d exceeds the %d-row / shared-memory capacity of the QP kernel", i, kRowCap);
  return POP_OK;
}

// Measurement hook: SM cycles the last QP launch spent per problem in {setup, factorisations, ADMM
// loop} and the number of factorisations; out = n x 8 int64.
extern "C" int pop_qp_profile(pop_handle* h, int n, int64_t* out) {
  if (!h || !out || n <= 0) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  if (!h->qp || n > h->qp->cap_agents) return pop_fail(h, POP_E_STATE, "no QP launch to report on");
  POP_CUDA(h, cudaMemcpyAsync(out, h->qp->prof, sizeof(long long) * 8 * n, cudaMemcpyDeviceToHost, h->stream));
  POP_CUDA(h, cudaStreamSynchronize(h->stream));
  return POP_OK;
}
