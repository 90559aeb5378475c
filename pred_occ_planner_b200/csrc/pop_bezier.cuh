// pop_bezier.cuh — device Bernstein evaluation in the reference's operation order.
// Restates traj_utils/src/bernstein.cpp:25-59 (p = C^T * A * [1,s,..,s^4], left to right) and
// Bezier::locatePiece / calcPieces (traj_utils/include/traj_utils/bernstein.hpp:164-172,
// src/bernstein.cpp:184-196).
#pragma once
#include <cuda_runtime.h>

// s^n (n = 0..4) rounded once from a double-double product chain: agrees with a correctly
// rounded pow(s, n), which is what the CPU libm returns for these arguments.
__device__ __forceinline__ double pop_pow_int(double s, int n) {
  if (n == 0) return 1.0;
  if (n == 1) return s;
  double hi = s, lo = 0.0;
  for (int i = 1; i < n; i++) {
    const double ph = __dmul_rn(hi, s);
    double pl = __fma_rn(hi, s, -ph);          // exact error of hi*s
    pl = __dadd_rn(pl, __dmul_rn(lo, s));
    const double nh = __dadd_rn(ph, pl);
    lo = __dsub_rn(pl, __dsub_rn(nh, ph));
    hi = nh;
  }
  return hi;
}

static __constant__ double kPopBernsteinA[25] = {
    1, -4, 6, -4, 1, 0, 4, -12, 12, -4, 0, 0, 6, -12, 6, 0, 0, 0, 4, -4, 0, 0, 0, 0, 1};

// one piece on [t0, tf]; c = 5x3 row-major control points; deriv 0/1/2
__device__ inline void pop_piece_eval(const double* c, double t0, double tf, double tt, int deriv,
                                      double out[3]) {
  const double dur = __dsub_rn(tf, t0);
  const double s = __ddiv_rn(__dsub_rn(tt, t0), dur);
  double S[5];
  if (deriv == 0) {
    S[0] = 1;
    for (int i = 1; i <= 4; i++) S[i] = pop_pow_int(s, i);
  } else if (deriv == 1) {
    S[0] = 0; S[1] = 1;
    for (int i = 2; i <= 4; i++) S[i] = __dmul_rn(static_cast<double>(i), pop_pow_int(s, i - 1));
  } else {
    S[0] = 0; S[1] = 0; S[2] = 2;
    for (int i = 3; i <= 4; i++) S[i] = __dmul_rn(static_cast<double>(i * (i - 1)), pop_pow_int(s, i - 2));
  }
  for (int d = 0; d < 3; d++) {
    double CA[5];
    for (int j = 0; j < 5; j++) {
      double acc = 0;
      for (int i = 0; i < 5; i++) acc = __dadd_rn(acc, __dmul_rn(c[i * 3 + d], kPopBernsteinA[i * 5 + j]));
      CA[j] = acc;
    }
    double acc = 0;
    for (int j = 0; j < 5; j++) acc = __dadd_rn(acc, __dmul_rn(CA[j], S[j]));
    if (deriv == 1) acc = __ddiv_rn(acc, dur);
    if (deriv == 2) acc = __ddiv_rn(acc, __dmul_rn(dur, dur));
    out[d] = acc;
  }
}

// Bezier::getPos/getVel/getAcc with piece durations dur[M] (already widened to double)
template <typename DurT>
__device__ inline void pop_bezier_eval(const DurT* dur, int M, const double* cpts, double tt,
                                       int deriv, double out[3]) {
  double rem = tt;
  int piece = M - 1;
  for (int i = 0; i < M; i++) {
    rem = __dsub_rn(rem, static_cast<double>(dur[i]));
    if (rem < 0) { piece = i; break; }
  }
  double t0 = 0;
  for (int k = 0; k < piece; k++) t0 = __dadd_rn(t0, static_cast<double>(dur[k]));
  const double tf = __dadd_rn(t0, static_cast<double>(dur[piece]));
  pop_piece_eval(cpts + piece * 15, t0, tf, tt, deriv, out);
}
