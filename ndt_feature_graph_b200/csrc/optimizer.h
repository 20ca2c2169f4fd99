// optimizer.h — the Newton / More-Thuente controller of an NDT-D2D registration as a RESUMABLE state
// machine: it never evaluates the score itself; it asks for "derivatives at pose P (with/without
// Hessian)" and is resumed with the 28 reduced sums.  On the GPU one thread of the registration's
// thread-block (cluster) runs it between derivative passes, so a whole registration is one launch.
//
// Control flow restated from the reference (no code copied; the loop structure, constants and quirks
// are what defines the result):
//   match / matchFusion      ndt_feature/include/ndt_feature/ndt_matcher_d2d_fusion.h:797-1155
//                            (== upstream NDTMatcherD2D::match when no feature/soft terms are active;
//                             call site ndt_feature/src/ndt_feature_src/ndt_feature_graph.cpp:273)
//   lineSearchMT             ndt_matcher_d2d_fusion.h:390-793  (constants :400-408, info codes :655-678)
//   lineSearchMTFusionTcov   ndt_matcher_d2d_fusion.h:37-385   (soft prior on the accumulated local pose)
//   Mahalanobis prior        ndt_matcher_d2d_fusion.h:11-32
//   MoreThuente::cstep       [upstream] == MINPACK dcstep, SURVEY.md Appendix A4
//   increment convention     TR = Trans(p0..2)*Rx(p3)*Ry(p4)*Rz(p5), T <- TR*T  (:1035-1043)
#pragma once

#include <cstring>

#include "d2d_pair.h"

#ifdef __CUDACC__
#define NDTB_HDF __host__ __device__
#else
#define NDTB_HDF
#endif

namespace ndtb {

struct Pose {
  double R[9];  // row-major rotation
  double t[3];
};

struct OptParams {
  int itr_max, step_control, regularize;
  int planar;  // NDTMatcherD2D_2D: estimate (x, y, yaw) only
  int fusion, soft, tik;
  double delta_score;
  double Q[36];  // Tcov^-1 (fusion only), row-major
};

enum { PH_NEWTON = 0, PH_LS_INIT = 1, PH_LS_EVAL = 2, PH_FINAL = 3, PH_DONE = 4 };

struct OptState {
  int phase;
  int want_hess;  // what the pending evaluation must compute
  Pose Peval;     // pose of the pending evaluation
  Pose T, Tbest, Tinit;
  double score_best, score_here;
  int itr, ret, exit_code, n_hess, n_grad, nonfinite;
  double pose_local[6], x0[6], incr[6], scg[6];
  // line search
  int ls_soft;  // 1 while the (discarded) Tcov search of matchFusion :1008-1010 runs
  double X[6];
  double stp, dginit, dgtest, width, width1, finit, stx, fx, dgx, sty, fy, dgy, stmin, stmax;
  int infoc, nfev, brackt, stage1;
  // evaluations the reference repeats at a pose it has just evaluated are answered from these (bitwise the same
  // value a repeated pass would give): score + gradient of the Hessian pass at T (line-search start), and the pose /
  // score of the last evaluation (final score when the last step lands on an evaluated trial pose)
  double sg[7];
  Pose Plast;
  double score_last;
  int have_last;
  int n_exec;  // derivative passes actually executed on the device
  // stand-alone NDTMatcherD2D::lineSearchMT (host-driven, ndtb_d2d_line_search_cells): stop after the search
  int ls_only;
  double ls_result;
};

// ------------------------------------------------------------------ small dense algebra
NDTB_HDF inline void m3_mul(const double *A, const double *B, double *C) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
  for (int i = 0; i < 9; i++) C[i] = t[i];
}

NDTB_HDF inline Pose pose_from_cm(const double *T) {
  Pose P;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) P.R[i * 3 + j] = T[j * 4 + i];
    P.t[i] = T[12 + i];
  }
  return P;
}
NDTB_HDF inline void pose_to_cm(const Pose &P, double *T) {
  for (int i = 0; i < 16; i++) T[i] = 0.0;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[j * 4 + i] = P.R[i * 3 + j];
    T[12 + i] = P.t[i];
  }
  T[15] = 1.0;
}
NDTB_HDF inline Pose pose_from_vec(const double *p) {
  const double cx = cos(p[3]), sx = sin(p[3]), cy = cos(p[4]), sy = sin(p[4]), cz = cos(p[5]), sz = sin(p[5]);
  const double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
  const double Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
  const double Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  Pose P;
  double t[9];
  m3_mul(Rx, Ry, t);
  m3_mul(t, Rz, P.R);
  P.t[0] = p[0], P.t[1] = p[1], P.t[2] = p[2];
  return P;
}
NDTB_HDF inline Pose pose_mul(const Pose &A, const Pose &B) {
  Pose C;
  m3_mul(A.R, B.R, C.R);
  for (int i = 0; i < 3; i++)
    C.t[i] = (A.R[i * 3] * B.t[0] + A.R[i * 3 + 1] * B.t[1] + A.R[i * 3 + 2] * B.t[2]) + A.t[i];
  return C;
}
NDTB_HDF inline Pose pose_inverse(const Pose &A) {
  Pose I;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) I.R[i * 3 + j] = A.R[j * 3 + i];
  for (int i = 0; i < 3; i++) I.t[i] = -(I.R[i * 3] * A.t[0] + I.R[i * 3 + 1] * A.t[1] + I.R[i * 3 + 2] * A.t[2]);
  return I;
}
// ndt_feature::getRobustYawFromAffine3d (ndt_feature/include/ndt_feature/utils.h:30-40)
NDTB_HDF inline double robust_yaw(const Pose &P) {
  double d = P.R[0];
  d = d > 1.0 ? 1.0 : (d < -1.0 ? -1.0 : d);
  const double a = acos(d);
  return P.R[3] > 0 ? a : -a;
}

// cyclic Jacobi, symmetric n x n (n<=6), eigenvalues ascending, eigenvectors in the columns of V.
// Compile-time n: every loop unrolls and the 3x3 instance (one per NDT cell) lives in registers.
// max_sweeps < 64 is for callers that defer the rare non-converging matrices (exactly singular ones never meet the
// stopping test and run all 64 sweeps): returns false when the sweep cap was hit before the stopping test passed; the
// result is then NOT the 64-sweep result and must be recomputed with the full cap.
NDTB_HDF inline bool same_bits(double a, double b) {
#ifdef __CUDA_ARCH__
  return __double_as_longlong(a) == __double_as_longlong(b);
#else
  return std::memcmp(&a, &b, sizeof a) == 0;
#endif
}
// n == 3 only (one per NDT cell): a sweep that changes no bit of A or V is a fixed point of the deterministic iteration —
// every later sweep repeats it, so stopping there returns exactly what the remaining sweeps would.  The exactly singular
// covariances (collinear / 3-point cells, ~2 % of the cells) never meet the stopping test but all reach such a fixed point
// within 13 sweeps (200 000 random degenerate cells on the host), instead of running the 64.
// TRACK is a template flag because the bookkeeping costs the cells that converge in 2-4 sweeps more than it saves them (B200:
// k_eigen 1.01 -> 1.19 ms with it, k_eigen_hard 0.48 -> 0.11 ms): the map build turns it on for the deferred cells only.
template <int n, bool TRACK = false>
NDTB_HDF inline bool eig_sym_n(const double *Ain, double *evals, double *V, int max_sweeps = 64, int *sweeps_run = nullptr) {
  double A[n * n];
  for (int i = 0; i < n * n; i++) A[i] = Ain[i];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  bool converged = max_sweeps >= 64;
  int sweep = 0;
  for (; sweep < max_sweeps; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        const double a2 = A[i * n + j] * A[i * n + j];
        if (i == j) diag += a2; else off += a2;
      }
    if (off <= 1e-32 * diag || off == 0.0) {
      converged = true;
      break;
    }
    bool changed = !(TRACK && n == 3);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int p = 0; p < n - 1; p++)
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int q = p + 1; q < n; q++) {
        const double apq = A[p * n + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
          if (TRACK && n == 3) changed = changed || !same_bits(A[k * n + p], akp) || !same_bits(A[k * n + q], akq);
        }
        for (int k = 0; k < n; k++) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
          if (TRACK && n == 3) changed = changed || !same_bits(A[p * n + k], apk) || !same_bits(A[q * n + k], aqk);
        }
        for (int k = 0; k < n; k++) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq;
          V[k * n + q] = s * vkp + c * vkq;
          if (TRACK && n == 3) changed = changed || !same_bits(V[k * n + p], vkp) || !same_bits(V[k * n + q], vkq);
        }
      }
    if (!changed) {  // fixed point: the remaining sweeps would repeat this one
      converged = true;
      sweep++;
      break;
    }
  }
  if (sweeps_run) *sweeps_run = sweep;
  // stable ascending order by eigenvalue (selection network free of dynamic indexing for small n)
  double d[n];
  int order[n];
  for (int i = 0; i < n; i++) d[i] = A[i * n + i], order[i] = i;
  for (int i = 1; i < n; i++)
    for (int k = i; k > 0; k--) {
      if (d[k - 1] > d[k]) {
        const double td = d[k - 1]; d[k - 1] = d[k]; d[k] = td;
        const int to = order[k - 1]; order[k - 1] = order[k]; order[k] = to;
      }
    }
  double Vt[n * n];
  for (int j = 0; j < n; j++) {
    evals[j] = d[j];
    for (int i = 0; i < n; i++) {
      double v = 0.0;
      for (int o = 0; o < n; o++) v = (order[j] == o) ? V[i * n + o] : v;
      Vt[i * n + j] = v;
    }
  }
  for (int i = 0; i < n * n; i++) V[i] = Vt[i];
  return converged;
}
NDTB_HDF inline void eig_sym(int n, const double *Ain, double *evals, double *V) {
  if (n == 3) eig_sym_n<3>(Ain, evals, V);
  else eig_sym_n<6>(Ain, evals, V);
}

// x = A^-1 b, LDL^T with symmetric diagonal pivoting (Eigen::LDLT semantics), n = 6
NDTB_HDF inline void ldlt_solve6(const double *Ain, const double *b, double *x) {
  const int n = 6;
  double A[36];
  for (int i = 0; i < 36; i++) A[i] = Ain[i];
  int perm[6];
  for (int i = 0; i < n; i++) perm[i] = i;
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = fabs(A[k * n + k]);
    for (int i = k + 1; i < n; i++)
      if (fabs(A[i * n + i]) > best) best = fabs(A[i * n + i]), p = i;
    if (p != k) {
      for (int j = 0; j < n; j++) { const double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
      for (int i = 0; i < n; i++) { const double t = A[i * n + k]; A[i * n + k] = A[i * n + p]; A[i * n + p] = t; }
      const int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
    }
    const double d = A[k * n + k];
    if (d == 0.0) continue;
    for (int i = k + 1; i < n; i++) A[i * n + k] /= d;
    for (int i = k + 1; i < n; i++)
      for (int j = k + 1; j <= i; j++) {
        A[i * n + j] -= A[i * n + k] * d * A[j * n + k];
        A[j * n + i] = A[i * n + j];
      }
  }
  double y[6];
  for (int i = 0; i < n; i++) y[i] = b[perm[i]];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++) y[i] -= A[i * n + j] * y[j];
  for (int i = 0; i < n; i++) {
    const double d = A[i * n + i];
    y[i] = (fabs(d) > 2.2250738585072014e-308) ? y[i] / d : 0.0;
  }
  for (int i = n - 1; i >= 0; i--)
    for (int j = i + 1; j < n; j++) y[i] -= A[j * n + i] * y[j];
  for (int i = 0; i < n; i++) x[perm[i]] = y[i];
}

// general 6x6 inverse (Gauss-Jordan, partial pivoting); false if singular
NDTB_HDF inline bool inv6(const double *Ain, double *Ainv) {
  const int n = 6;
  double M[6][12];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) M[i][j] = Ain[i * n + j], M[i][n + j] = (i == j) ? 1.0 : 0.0;
  for (int c = 0; c < n; c++) {
    int p = c;
    for (int r = c + 1; r < n; r++)
      if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
    if (M[p][c] == 0.0 || M[p][c] != M[p][c]) return false;
    if (p != c)
      for (int j = 0; j < 2 * n; j++) { const double t = M[p][j]; M[p][j] = M[c][j]; M[c][j] = t; }
    const double id = 1.0 / M[c][c];
    for (int j = 0; j < 2 * n; j++) M[c][j] *= id;
    for (int r = 0; r < n; r++)
      if (r != c) {
        const double f = M[r][c];
        if (f != 0.0)
          for (int j = 0; j < 2 * n; j++) M[r][j] -= f * M[c][j];
      }
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) Ainv[i * n + j] = M[i][n + j];
  return true;
}

// ------------------------------------------------------------------ soft prior (ndt_matcher_d2d_fusion.h:11-32)
NDTB_HDF inline double maha_score(const double *x, const double *C) {
  double s = 0;
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) s += x[i] * C[i * 6 + j] * x[j];
  return s;
}
NDTB_HDF inline void maha_gradient(const double *x, const double *C, double *g) {
  for (int i = 0; i < 6; i++) {
    g[i] = 0;
    for (int j = 0; j < 6; j++) g[i] += (C[j * 6 + i] + C[i * 6 + j]) * x[j];
  }
}

// ------------------------------------------------------------------ MoreThuente::cstep
NDTB_HDF inline double mt_max3(double a, double b, double c) {
  a = fabs(a), b = fabs(b), c = fabs(c);
  return a > b ? (a > c ? a : c) : (b > c ? b : c);
}
NDTB_HDF inline int mt_cstep(double &stx, double &fx, double &dx, double &sty, double &fy, double &dy, double &stp,
                             double fp, double dp, int &brackt, double stmin, double stmax) {
  const double lo = stx < sty ? stx : sty, hi = stx > sty ? stx : sty;
  if ((brackt && (stp <= lo || stp >= hi)) || (dx * (stp - stx) >= 0.0) || (stmax < stmin)) return 0;
  const double sgnd = dp * (dx / fabs(dx));
  int info, bound;
  double theta, s, gamma, p, q, r, stpc, stpq, stpf;
  if (fp > fx) {  // higher value: minimum bracketed
    info = 1, bound = 1;
    theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
    s = mt_max3(theta, dx, dp);
    gamma = s * sqrt(((theta / s) * (theta / s)) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    p = (gamma - dx) + theta;
    q = ((gamma - dx) + gamma) + dp;
    r = p / q;
    stpc = stx + r * (stp - stx);
    stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2) * (stp - stx);
    stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2;
    brackt = 1;
  } else if (sgnd < 0.0) {  // lower value, derivatives of opposite sign
    info = 2, bound = 0;
    theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
    s = mt_max3(theta, dx, dp);
    gamma = s * sqrt(((theta / s) * (theta / s)) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = ((gamma - dp) + gamma) + dx;
    r = p / q;
    stpc = stp + r * (stx - stp);
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
    brackt = 1;
  } else if (fabs(dp) < fabs(dx)) {  // lower value, same sign, derivative magnitude decreases
    info = 3, bound = 1;
    theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
    s = mt_max3(theta, dx, dp);
    const double rad = (theta / s) * (theta / s) - (dx / s) * (dp / s);
    gamma = s * sqrt(rad > 0.0 ? rad : 0.0);
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = (gamma + (dx - dp)) + gamma;
    r = p / q;
    if (r < 0.0 && gamma != 0.0)
      stpc = stp + r * (stx - stp);
    else if (stp > stx)
      stpc = stmax;
    else
      stpc = stmin;
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt)
      stpf = (fabs(stp - stpc) < fabs(stp - stpq)) ? stpc : stpq;
    else
      stpf = (fabs(stp - stpc) > fabs(stp - stpq)) ? stpc : stpq;
  } else {  // lower value, same sign, derivative magnitude does not decrease
    info = 4, bound = 0;
    if (brackt) {
      theta = 3 * (fp - fy) / (sty - stp) + dy + dp;
      s = mt_max3(theta, dy, dp);
      gamma = s * sqrt(((theta / s) * (theta / s)) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      p = (gamma - dp) + theta;
      q = ((gamma - dp) + gamma) + dy;
      r = p / q;
      stpf = stp + r * (sty - stp);
    } else
      stpf = stp > stx ? stmax : stmin;
  }
  if (fp > fx) {
    sty = stp, fy = fp, dy = dp;
  } else {
    if (sgnd < 0.0) sty = stx, fy = fx, dy = dx;
    stx = stp, fx = fp, dx = dp;
  }
  stpf = stmax < stpf ? stmax : stpf;
  stpf = stmin > stpf ? stmin : stpf;
  stp = stpf;
  if (brackt && bound) {
    const double lim = stx + 0.66 * (sty - stx);
    if (sty > stx)
      stp = lim < stp ? lim : stp;
    else
      stp = lim > stp ? lim : stp;
  }
  return info;
}

// ------------------------------------------------------------------ the state machine
constexpr double LS_RECOVERY = 0.1, LS_FTOL = 0.11111, LS_GTOL = 0.99999, LS_STPMAX = 4.0, LS_STPMIN = 0.001,
                 LS_XTOL = 0.01;
constexpr int LS_MAXFEV = 40;

NDTB_HDF inline void opt_request(OptState &s, const Pose &P, int hess, int phase) {
  s.Peval = P;
  s.want_hess = hess;
  s.phase = phase;
}

NDTB_HDF inline void opt_begin(OptState &s, const OptParams &prm, const double *T0_colmajor) {
  s.T = pose_from_cm(T0_colmajor);
  s.Tbest = s.T;
  s.Tinit = s.T;
  s.score_best = prm.fusion ? 1.7976931348623157e308 : 2147483647.0;  // INT_MAX upstream
  s.score_here = 0;
  s.itr = 0, s.ret = 1, s.exit_code = 0, s.n_hess = 0, s.n_grad = 0, s.nonfinite = 0;
  for (int i = 0; i < 6; i++) s.pose_local[i] = s.x0[i] = s.incr[i] = s.scg[i] = s.X[i] = 0.0;
  s.ls_soft = 0;
  s.have_last = 0, s.n_exec = 0, s.score_last = 0;
  s.ls_only = 0, s.ls_result = 0.0;
  for (int i = 0; i < 7; i++) s.sg[i] = 0.0;
  opt_request(s, s.T, 1, PH_NEWTON);
}

NDTB_HDF inline void opt_finish(OptState &s) { s.phase = PH_DONE; }

// head of the More-Thuente loop: bracket bookkeeping, clamp, then ask for the trial evaluation
NDTB_HDF inline void ls_trial(OptState &s, const OptParams &prm) {
  if (s.brackt) {
    s.stmin = s.stx < s.sty ? s.stx : s.sty;
    s.stmax = s.stx > s.sty ? s.stx : s.sty;
  } else {
    s.stmin = s.stx;
    s.stmax = s.stp + 4 * (s.stp - s.stx);
  }
  s.stp = s.stp > LS_STPMIN ? s.stp : LS_STPMIN;
  s.stp = s.stp < LS_STPMAX ? s.stp : LS_STPMAX;
  if ((s.brackt && (s.stp <= s.stmin || s.stp >= s.stmax)) || (s.nfev >= LS_MAXFEV - 1) || (s.infoc == 0) ||
      (s.brackt && (s.stmax - s.stmin <= LS_XTOL * s.stmax)))
    s.stp = s.stx;
  double pincr[6];
  for (int i = 0; i < 6; i++) pincr[i] = s.stp * s.incr[i];
  if (s.ls_soft)
    for (int i = 0; i < 6; i++) s.X[i] += pincr[i];  // (sic) accumulates across evaluations, fusion.h:183
  (void)prm;
  opt_request(s, pose_mul(pose_from_vec(pincr), s.T), 0, PH_LS_EVAL);
}

NDTB_HDF inline void on_ls_init(OptState &s, const OptParams &prm, const double *sums7);
NDTB_HDF inline void on_final(OptState &s, const OptParams &prm, double score);
NDTB_HDF inline bool pose_same(const Pose &a, const Pose &b) {
  bool eq = true;
  for (int i = 0; i < 9; i++) eq = eq && (a.R[i] == b.R[i]);
  for (int i = 0; i < 3; i++) eq = eq && (a.t[i] == b.t[i]);
  return eq;
}

NDTB_HDF inline void opt_apply_step(OptState &s, const OptParams &prm, double step) {
  for (int i = 0; i < 6; i++) s.incr[i] *= step;
  s.T = pose_mul(pose_from_vec(s.incr), s.T);
  double nrm = 0;
  for (int i = 0; i < 6; i++) s.pose_local[i] += s.incr[i], nrm += s.incr[i] * s.incr[i];
  nrm = sqrt(nrm);
  bool convergence = false;
  if (s.itr > 0) convergence = nrm < prm.delta_score;
  if (s.itr > prm.itr_max) convergence = true, s.ret = 0, s.exit_code = 3;
  s.itr++;
  if (!convergence)
    opt_request(s, s.T, 1, PH_NEWTON);
  else if (s.have_last && pose_same(s.Plast, s.T))
    on_final(s, prm, s.score_last);  // the final pose is the last trial pose: its score is already known
  else
    opt_request(s, s.T, 0, PH_FINAL);
}

// lineSearchMT starts with a gradient-only derivativesNDT at the current pose (fusion.h:444 / :80): the Hessian pass of
// this Newton iteration was evaluated at exactly that pose, so its score and gradient are handed over directly.
NDTB_HDF inline void ls_start(OptState &s, const OptParams &prm, int soft) {
  s.ls_soft = soft;
  if (soft)
    for (int i = 0; i < 6; i++) s.X[i] = s.pose_local[i];
  on_ls_init(s, prm, s.sg);
}

NDTB_HDF inline void ls_finish_step(OptState &s, const OptParams &prm, double step) {
  if (s.ls_only) {  // lineSearchMT returns the step; the caller applies it
    s.ls_result = step;
    opt_finish(s);
    return;
  }
  if (prm.fusion) step = step > 0.0 ? step : 0.0;  // fusion.h:1018-1023 with step_size_feat == 0
  opt_apply_step(s, prm, step);
}

NDTB_HDF inline void ls_done(OptState &s, const OptParams &prm, double step) {
  if (s.ls_soft) {  // fusion.h:1008-1023: the Tcov search result is overwritten by the NDT search (same pose, same sums)
    s.ls_soft = 0;
    on_ls_init(s, prm, s.sg);
    return;
  }
  ls_finish_step(s, prm, step);
}

// Resume with the reduced sums of the evaluation that was requested: sums[0]=score, [1..6]=g,
// [7..27]=H upper triangle (only when want_hess).
NDTB_HDF inline void opt_advance(OptState &s, const OptParams &prm, const double *sums) {
  if (!(sums[0] * 0.0 == 0.0)) s.nonfinite = 1;
  switch (s.phase) {
    case PH_NEWTON: {
      s.n_hess++;
      s.n_exec++;
      for (int i = 0; i < 7; i++) s.sg[i] = sums[i];
      s.Plast = s.Peval, s.score_last = sums[0], s.have_last = 1;
      s.score_here = sums[0];
      double g[6], H[36];
      for (int i = 0; i < 6; i++) g[i] = sums[ACC_G + i];
      for (int p = 0; p < 6; p++)
        for (int q = p; q < 6; q++) H[p * 6 + q] = H[q * 6 + p] = sums[hidx(p, q)];
      if (prm.soft) {
        s.score_here += maha_score(s.pose_local, prm.Q);
        double gm[6];
        maha_gradient(s.pose_local, prm.Q, gm);
        for (int i = 0; i < 6; i++) {
          g[i] += gm[i];
          for (int j = 0; j < 6; j++) H[i * 6 + j] += prm.Q[j * 6 + i] + prm.Q[i * 6 + j];
        }
      }
      if (prm.tik) {  // fusion.h:894-911  g <- H^T g + Q x0 ; H <- H^T H + Q ; x0 = (x, y, 0, 0, 0, yaw) of T*Tinit^-1
        const Pose X0 = pose_mul(s.T, pose_inverse(s.Tinit));
        s.x0[0] = X0.t[0], s.x0[1] = X0.t[1], s.x0[2] = 0, s.x0[3] = 0, s.x0[4] = 0, s.x0[5] = robust_yaw(X0);
        double Hn[36], gn[6];
        for (int i = 0; i < 6; i++) {
          gn[i] = 0;
          for (int k = 0; k < 6; k++) gn[i] += H[k * 6 + i] * g[k] + prm.Q[i * 6 + k] * s.x0[k];
          for (int j = 0; j < 6; j++) {
            double a = 0;
            for (int k = 0; k < 6; k++) a += H[k * 6 + i] * H[k * 6 + j];
            Hn[i * 6 + j] = a + prm.Q[i * 6 + j];
          }
        }
        for (int i = 0; i < 36; i++) H[i] = Hn[i];
        for (int i = 0; i < 6; i++) g[i] = gn[i];
        s.score_here += maha_score(s.x0, prm.Q);
      }
      if (prm.planar) {
        // NDTMatcherD2D_2D [upstream] (matchFusion2d, fusion.h:1159-1176): z, roll and pitch are decoupled (zero gradient,
        // unit diagonal), so their increments are exactly 0 and the (x, y, yaw) block is solved as a 3x3 system
        const int drop[3] = {2, 3, 4};
        for (int d = 0; d < 3; d++) {
          g[drop[d]] = 0.0;
          for (int j = 0; j < 6; j++) H[drop[d] * 6 + j] = H[j * 6 + drop[d]] = 0.0;
          H[drop[d] * 6 + drop[d]] = 1.0;
        }
      }
      for (int i = 0; i < 6; i++) s.scg[i] = g[i];
      if (s.score_here < s.score_best) s.Tbest = s.T, s.score_best = s.score_here;
      double evals[6], evecs[36];
      eig_sym(6, H, evals, evecs);
      const double minC = evals[0], maxC = evals[5];
      double gnorm = 0;
      for (int i = 0; i < 6; i++) gnorm += g[i] * g[i];
      gnorm = sqrt(gnorm);
      if (minC < 0) {
        if (prm.regularize || prm.fusion) {
          double reg = gnorm;
          reg = reg + minC > 0 ? reg : 0.001 * maxC - minC;
          for (int i = 0; i < 6; i++) evals[i] += reg;
          for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) {
              double a = 0;
              for (int k = 0; k < 6; k++) a += evecs[i * 6 + k] * evals[k] * evecs[j * 6 + k];
              H[i * 6 + j] = a;
            }
        } else {
          if (s.score_here > s.score_best) s.T = s.Tbest;
          s.exit_code = 4;
          opt_finish(s);
          return;
        }
      }
      if (gnorm <= prm.delta_score) {
        if (s.score_here > s.score_best) s.T = s.Tbest;
        s.exit_code = 1;
        opt_finish(s);
        return;
      }
      double ng[6];
      ldlt_solve6(H, g, ng);
      double dginit = 0;
      for (int i = 0; i < 6; i++) s.incr[i] = -ng[i], dginit += s.incr[i] * s.scg[i];
      if (dginit > 0) {
        if (s.score_here > s.score_best) s.T = s.Tbest;
        s.exit_code = 2;
        opt_finish(s);
        return;
      }
      if (prm.step_control)
        ls_start(s, prm, prm.soft ? 1 : 0);
      else
        opt_apply_step(s, prm, 1.0);
      return;
    }
    case PH_LS_INIT: {  // (not requested any more: ls_start answers it from the Hessian pass)
      s.n_exec++;
      on_ls_init(s, prm, sums);
      return;
    }
    case PH_LS_EVAL: {
      s.n_grad++;
      s.n_exec++;
      s.Plast = s.Peval, s.score_last = sums[0], s.have_last = 1;
      double f = sums[0], gh[6];
      for (int i = 0; i < 6; i++) gh[i] = sums[ACC_G + i];
      if (s.ls_soft) {
        f += maha_score(s.X, prm.Q);
        double gm[6];
        maha_gradient(s.X, prm.Q, gm);
        for (int i = 0; i < 6; i++) gh[i] += gm[i];
      }
      double dg = 0;
      for (int i = 0; i < 6; i++) dg += s.incr[i] * gh[i];
      s.nfev++;
      const double ftest1 = s.finit + s.stp * s.dgtest;
      int info = 0;
      if ((s.brackt && (s.stp <= s.stmin || s.stp >= s.stmax)) || (s.infoc == 0)) info = 6;
      if ((s.stp == LS_STPMAX) && (f <= ftest1) && (dg <= s.dgtest)) info = 5;
      if ((s.stp == LS_STPMIN) && ((f > ftest1) || (dg >= s.dgtest))) info = 4;
      if (s.nfev >= LS_MAXFEV) info = 3;
      if (s.brackt && (s.stmax - s.stmin <= LS_XTOL * s.stmax)) info = 2;
      if ((f <= ftest1) && (fabs(dg) <= LS_GTOL * (-s.dginit))) info = 1;
      if (info != 0) {
        ls_done(s, prm, info != 1 ? LS_RECOVERY : s.stp);
        return;
      }
      const double mtol = LS_FTOL < LS_GTOL ? LS_FTOL : LS_GTOL;
      if (s.stage1 && (f <= ftest1) && (dg >= mtol * s.dginit)) s.stage1 = 0;
      if (s.stage1 && (f <= s.fx) && (f > ftest1)) {
        const double fm = f - s.stp * s.dgtest;
        double fxm = s.fx - s.stx * s.dgtest, fym = s.fy - s.sty * s.dgtest;
        const double dgm = dg - s.dgtest;
        double dgxm = s.dgx - s.dgtest, dgym = s.dgy - s.dgtest;
        s.infoc = mt_cstep(s.stx, fxm, dgxm, s.sty, fym, dgym, s.stp, fm, dgm, s.brackt, s.stmin, s.stmax);
        s.fx = fxm + s.stx * s.dgtest;
        s.fy = fym + s.sty * s.dgtest;
        s.dgx = dgxm + s.dgtest;
        s.dgy = dgym + s.dgtest;
      } else {
        s.infoc = mt_cstep(s.stx, s.fx, s.dgx, s.sty, s.fy, s.dgy, s.stp, f, dg, s.brackt, s.stmin, s.stmax);
      }
      if (s.brackt) {
        if (fabs(s.sty - s.stx) >= 0.66 * s.width1) s.stp = s.stx + 0.5 * (s.sty - s.stx);
        s.width1 = s.width;
        s.width = fabs(s.sty - s.stx);
      }
      ls_trial(s, prm);
      return;
    }
    case PH_FINAL: {
      s.n_exec++;
      on_final(s, prm, sums[0]);
      return;
    }
    default:
      return;
  }
}

NDTB_HDF inline void on_ls_init(OptState &s, const OptParams &prm, const double *sums7) {
  for (;;) {
    s.n_grad++;
    double score_init = sums7[0], gh[6];
    for (int i = 0; i < 6; i++) gh[i] = sums7[ACC_G + i];
    if (s.ls_soft) {
      score_init += maha_score(s.X, prm.Q);
      double gm[6];
      maha_gradient(s.X, prm.Q, gm);
      for (int i = 0; i < 6; i++) gh[i] += gm[i];
    }
    s.dginit = 0;
    for (int i = 0; i < 6; i++) s.dginit += s.incr[i] * gh[i];
    if (s.dginit >= 0.0) {
      for (int i = 0; i < 6; i++) s.incr[i] = -s.incr[i];
      s.dginit = -s.dginit;
      if (s.dginit >= 0.0) {  // no descent either way: the search returns the recovery step at once
        if (s.ls_soft) {
          s.ls_soft = 0;  // ... and the discarded Tcov search is followed by the NDT search
          continue;
        }
        ls_finish_step(s, prm, LS_RECOVERY);
        return;
      }
    }
    s.stp = 1.0;
    s.infoc = 1;
    s.brackt = 0, s.stage1 = 1, s.nfev = 0;
    s.dgtest = LS_FTOL * s.dginit;
    s.width = LS_STPMAX - LS_STPMIN;
    s.width1 = 2 * s.width;
    s.finit = score_init;
    s.stx = 0.0, s.fx = s.finit, s.dgx = s.dginit;
    s.sty = 0.0, s.fy = s.finit, s.dgy = s.dginit;
    ls_trial(s, prm);
    return;
  }
}

NDTB_HDF inline void on_final(OptState &s, const OptParams &prm, double score) {
  s.n_grad++;
  s.score_here = score;
  if (prm.soft) s.score_here += maha_score(s.pose_local, prm.Q);
  if (prm.tik) s.score_here += maha_score(s.x0, prm.Q);
  if (s.score_here > s.score_best) s.T = s.Tbest;
  opt_finish(s);
}

}  // namespace ndtb
