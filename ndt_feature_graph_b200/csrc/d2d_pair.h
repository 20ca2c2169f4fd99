// d2d_pair.h — per-(source cell, target cell) contribution to the NDT-D2D score, gradient and Hessian.
//
// Replaces the inner body of lslgeneric::NDTMatcherD2D::derivativesNDT [upstream perception_oru]
// (computeDerivativesLocal + update_gradient_hessian_local), reached from the reference at
// ndt_feature/include/ndt_feature/ndt_matcher_d2d_fusion.h:856 (Hessian) and :617,:444 (gradient only).
//
// Upstream materialises Jest (3x6), Hest (18x6), Zest (3x18), ZHest (18x18) per source cell and runs
// dense 6x6 loops per pair.  Here everything is contracted in closed form.  With the moved source cell
// (mu, C), target cell (m, S):  x = mu-m,  B = (C+S)^-1,  q = Bx,  w = Cq,  v = mu-w,
//   n_r = q x e_r,  J_r = e_r x mu,  z_r = Z_r q = e_r x w + C n_r,  d_r = J_r - z_r,  D_r = B d_r
//   Q   = 2 [ q ; v x q ]                              (gradient direction, g += factor*Q)
//   H/factor (upper triangle, mirrored like upstream):
//     tt: 2 B                      tr: 2 (D_b)_p        rr(a<=b): 2 d_a.D_b + 2 v_a q_b - 2 d_ab q.v - 2 n_a.C n_b
//     all: - (lfd2/2) Q_p Q_q
// which is term-by-term identical to 2 JtBJ + 2 xtBH - xtBZhBx - 2 xtBZBJ - 2 xtBZBJ^T + xtBZBZBx + its
// transpose - lfd2/2 QQ^T (SURVEY.md §8a "Specification of kernel (ii)").
//
// Accumulator layout (ACC_N doubles): [0]=score, [1..6]=g, [7..27]=H upper triangle row-major
// (00 01 02 03 04 05 11 12 ... 55).
#pragma once

#ifdef __CUDACC__
#define NDTB_HD __host__ __device__ __forceinline__
#else
#define NDTB_HD inline
#endif

#include <math.h>

namespace ndtb {

constexpr int ACC_G = 1;
constexpr int ACC_H = 7;
constexpr int ACC_N = 28;

// index of H(p,q), p<=q, in the packed upper triangle
NDTB_HD constexpr int hidx(int p, int q) { return ACC_H + p * 6 - (p * (p - 1)) / 2 + (q - p); }

// Adds the pair's contribution to acc.  g6 (optional, may be nullptr) receives this pair's own gradient
// contribution (used by covariance()).  Returns false if the pair is skipped (C+S not invertible by
// Eigen's computeInverseAndDetWithCheck threshold, or non-finite Mahalanobis distance).
// Where the 21 Hessian sums of a thread live: RegAcc = in the thread's accumulator array (acc[ACC_H + i]); a kernel
// that wants the registers for the pair arithmetic passes its own policy (d2d.cu: one shared-memory slot per lane).
struct RegAcc {
  static constexpr bool kPairHook = false;  // true: the policy wants every pair's own gradient (covariance pass)
  double *acc;
  NDTB_HD void add(int i, double v) const { acc[ACC_H + i] += v; }
};

template <bool HESS, class HA>
NDTB_HD bool pair_contrib_acc(const double mu0, const double mu1, const double mu2, const double *C /*6*/,
                              const double *m /*3*/, const double *S /*6*/, double lfd1, double lfd2, double *acc,
                              const HA &hacc, double *g6) {
  const double x0 = mu0 - m[0], x1 = mu1 - m[1], x2 = mu2 - m[2];
  const double a00 = C[0] + S[0], a01 = C[1] + S[1], a02 = C[2] + S[2];
  const double a11 = C[3] + S[3], a12 = C[4] + S[4], a22 = C[5] + S[5];
  const double c00 = a11 * a22 - a12 * a12;
  const double c10 = a02 * a12 - a01 * a22;
  const double c20 = a01 * a12 - a02 * a11;
  const double det = c00 * a00 + c10 * a01 + c20 * a02;
  if (!(fabs(det) > 1e-12)) return false;
#ifdef __CUDA_ARCH__
  const double id = __drcp_rn(det);  // the correctly rounded reciprocal == 1.0 / det bit for bit, a third of the instructions
#else
  const double id = 1.0 / det;
#endif
  const double b00 = c00 * id, b01 = c10 * id, b02 = c20 * id;
  const double b11 = (a00 * a22 - a02 * a02) * id;
  const double b12 = (a02 * a01 - a00 * a12) * id;
  const double b22 = (a00 * a11 - a01 * a01) * id;
  const double q0 = b00 * x0 + b01 * x1 + b02 * x2;
  const double q1 = b01 * x0 + b11 * x1 + b12 * x2;
  const double q2 = b02 * x0 + b12 * x1 + b22 * x2;
  const double l = x0 * q0 + x1 * q1 + x2 * q2;
  if (l * 0.0 != 0.0) return false;
  const double sh = -lfd1 * exp(-lfd2 * l * 0.5);
  const double factor = -(lfd2 * 0.5) * sh;
  acc[0] += sh;
  // w = C q, v = mu - w
  const double w0 = C[0] * q0 + C[1] * q1 + C[2] * q2;
  const double w1 = C[1] * q0 + C[3] * q1 + C[4] * q2;
  const double w2 = C[2] * q0 + C[4] * q1 + C[5] * q2;
  const double v0 = mu0 - w0, v1 = mu1 - w1, v2 = mu2 - w2;
  double Q[6];
  Q[0] = 2.0 * q0, Q[1] = 2.0 * q1, Q[2] = 2.0 * q2;
  Q[3] = 2.0 * (v1 * q2 - v2 * q1);
  Q[4] = 2.0 * (v2 * q0 - v0 * q2);
  Q[5] = 2.0 * (v0 * q1 - v1 * q0);
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int k = 0; k < 6; k++) {
    const double gk = factor * Q[k];
    acc[ACC_G + k] += gk;
    if (g6) g6[k] = gk;
  }
  if (!HESS) return true;

  // cn_r = C n_r with n_x = (0,q2,-q1), n_y = (-q2,0,q0), n_z = (q1,-q0,0)
  const double cnx0 = q2 * C[1] - q1 * C[2], cnx1 = q2 * C[3] - q1 * C[4], cnx2 = q2 * C[4] - q1 * C[5];
  const double cny0 = q0 * C[2] - q2 * C[0], cny1 = q0 * C[4] - q2 * C[1], cny2 = q0 * C[5] - q2 * C[2];
  const double cnz0 = q1 * C[0] - q0 * C[1], cnz1 = q1 * C[1] - q0 * C[3], cnz2 = q1 * C[2] - q0 * C[4];
  // d_r = J_r - z_r,  J_x=(0,-mu2,mu1) J_y=(mu2,0,-mu0) J_z=(-mu1,mu0,0),  z_r = e_r x w + cn_r
  // => d_x = (0,-v2,v1) - cn_x etc.
  const double dx0 = -cnx0, dx1 = -v2 - cnx1, dx2 = v1 - cnx2;
  const double dy0 = v2 - cny0, dy1 = -cny1, dy2 = -v0 - cny2;
  const double dz0 = -v1 - cnz0, dz1 = v0 - cnz1, dz2 = -cnz2;
  // D_r = B d_r
  const double Dx0 = b00 * dx0 + b01 * dx1 + b02 * dx2, Dx1 = b01 * dx0 + b11 * dx1 + b12 * dx2,
               Dx2 = b02 * dx0 + b12 * dx1 + b22 * dx2;
  const double Dy0 = b00 * dy0 + b01 * dy1 + b02 * dy2, Dy1 = b01 * dy0 + b11 * dy1 + b12 * dy2,
               Dy2 = b02 * dy0 + b12 * dy1 + b22 * dy2;
  const double Dz0 = b00 * dz0 + b01 * dz1 + b02 * dz2, Dz1 = b01 * dz0 + b11 * dz1 + b12 * dz2,
               Dz2 = b02 * dz0 + b12 * dz1 + b22 * dz2;
  const double qv = q0 * v0 + q1 * v1 + q2 * v2;
  const double hl = lfd2 * 0.5;
  // translation-translation
  hacc.add(hidx(0, 0) - ACC_H, factor * (2.0 * b00 - hl * Q[0] * Q[0]));
  hacc.add(hidx(0, 1) - ACC_H, factor * (2.0 * b01 - hl * Q[0] * Q[1]));
  hacc.add(hidx(0, 2) - ACC_H, factor * (2.0 * b02 - hl * Q[0] * Q[2]));
  hacc.add(hidx(1, 1) - ACC_H, factor * (2.0 * b11 - hl * Q[1] * Q[1]));
  hacc.add(hidx(1, 2) - ACC_H, factor * (2.0 * b12 - hl * Q[1] * Q[2]));
  hacc.add(hidx(2, 2) - ACC_H, factor * (2.0 * b22 - hl * Q[2] * Q[2]));
  // translation-rotation
  hacc.add(hidx(0, 3) - ACC_H, factor * (2.0 * Dx0 - hl * Q[0] * Q[3]));
  hacc.add(hidx(1, 3) - ACC_H, factor * (2.0 * Dx1 - hl * Q[1] * Q[3]));
  hacc.add(hidx(2, 3) - ACC_H, factor * (2.0 * Dx2 - hl * Q[2] * Q[3]));
  hacc.add(hidx(0, 4) - ACC_H, factor * (2.0 * Dy0 - hl * Q[0] * Q[4]));
  hacc.add(hidx(1, 4) - ACC_H, factor * (2.0 * Dy1 - hl * Q[1] * Q[4]));
  hacc.add(hidx(2, 4) - ACC_H, factor * (2.0 * Dy2 - hl * Q[2] * Q[4]));
  hacc.add(hidx(0, 5) - ACC_H, factor * (2.0 * Dz0 - hl * Q[0] * Q[5]));
  hacc.add(hidx(1, 5) - ACC_H, factor * (2.0 * Dz1 - hl * Q[1] * Q[5]));
  hacc.add(hidx(2, 5) - ACC_H, factor * (2.0 * Dz2 - hl * Q[2] * Q[5]));
  // rotation-rotation (a<=b): 2 d_a.D_b + 2 v_a q_b - 2 delta q.v - 2 n_a.cn_b
  // n_x.cn_b = q2*cn_b1 - q1*cn_b2 ; n_y.cn_b = -q2*cn_b0 + q0*cn_b2 ; n_z.cn_b = q1*cn_b0 - q0*cn_b1
  const double xx = dx0 * Dx0 + dx1 * Dx1 + dx2 * Dx2 + v0 * q0 - qv - (q2 * cnx1 - q1 * cnx2);
  const double xy = dx0 * Dy0 + dx1 * Dy1 + dx2 * Dy2 + v0 * q1 - (q2 * cny1 - q1 * cny2);
  const double xz = dx0 * Dz0 + dx1 * Dz1 + dx2 * Dz2 + v0 * q2 - (q2 * cnz1 - q1 * cnz2);
  const double yy = dy0 * Dy0 + dy1 * Dy1 + dy2 * Dy2 + v1 * q1 - qv - (q0 * cny2 - q2 * cny0);
  const double yz = dy0 * Dz0 + dy1 * Dz1 + dy2 * Dz2 + v1 * q2 - (q0 * cnz2 - q2 * cnz0);
  const double zz = dz0 * Dz0 + dz1 * Dz1 + dz2 * Dz2 + v2 * q2 - qv - (q1 * cnz0 - q0 * cnz1);
  hacc.add(hidx(3, 3) - ACC_H, factor * (2.0 * xx - hl * Q[3] * Q[3]));
  hacc.add(hidx(3, 4) - ACC_H, factor * (2.0 * xy - hl * Q[3] * Q[4]));
  hacc.add(hidx(3, 5) - ACC_H, factor * (2.0 * xz - hl * Q[3] * Q[5]));
  hacc.add(hidx(4, 4) - ACC_H, factor * (2.0 * yy - hl * Q[4] * Q[4]));
  hacc.add(hidx(4, 5) - ACC_H, factor * (2.0 * yz - hl * Q[4] * Q[5]));
  hacc.add(hidx(5, 5) - ACC_H, factor * (2.0 * zz - hl * Q[5] * Q[5]));
  return true;
}

template <bool HESS>
NDTB_HD bool pair_contrib(const double mu0, const double mu1, const double mu2, const double *C /*6*/, const double *m /*3*/,
                          const double *S /*6*/, double lfd1, double lfd2, double *acc, double *g6) {
  return pair_contrib_acc<HESS, RegAcc>(mu0, mu1, mu2, C, m, S, lfd1, lfd2, acc, RegAcc{acc}, g6);
}

// Branch-free gradient-only variant: a skipped pair contributes exact zeros through selects, so two calls on
// independent pairs form one basic block and the scheduler interleaves their (long, serial) dependency chains.
// Same operations in the same order as pair_contrib<false> for every pair that is not skipped.
NDTB_HD void pair_grad_nb(const double mu0, const double mu1, const double mu2, const double *C /*6*/, const double *m /*3*/,
                          const double *S /*6*/, double lfd1, double lfd2, bool live, double *acc, double *npairs) {
  const double x0 = mu0 - m[0], x1 = mu1 - m[1], x2 = mu2 - m[2];
  const double a00 = C[0] + S[0], a01 = C[1] + S[1], a02 = C[2] + S[2];
  const double a11 = C[3] + S[3], a12 = C[4] + S[4], a22 = C[5] + S[5];
  const double c00 = a11 * a22 - a12 * a12;
  const double c10 = a02 * a12 - a01 * a22;
  const double c20 = a01 * a12 - a02 * a11;
  const double det = c00 * a00 + c10 * a01 + c20 * a02;
  const bool ok1 = live && (fabs(det) > 1e-12);
  const double id = 1.0 / (ok1 ? det : 1.0);
  const double b00 = c00 * id, b01 = c10 * id, b02 = c20 * id;
  const double b11 = (a00 * a22 - a02 * a02) * id;
  const double b12 = (a02 * a01 - a00 * a12) * id;
  const double b22 = (a00 * a11 - a01 * a01) * id;
  const double q0 = b00 * x0 + b01 * x1 + b02 * x2;
  const double q1 = b01 * x0 + b11 * x1 + b12 * x2;
  const double q2 = b02 * x0 + b12 * x1 + b22 * x2;
  const double l = x0 * q0 + x1 * q1 + x2 * q2;
  const bool ok = ok1 && (l * 0.0 == 0.0);
  const double sh = -lfd1 * exp(-lfd2 * (ok ? l : 0.0) * 0.5);
  const double factor = -(lfd2 * 0.5) * sh;
  acc[0] += ok ? sh : 0.0;
  *npairs += ok ? 1.0 : 0.0;
  const double w0 = C[0] * q0 + C[1] * q1 + C[2] * q2;
  const double w1 = C[1] * q0 + C[3] * q1 + C[4] * q2;
  const double w2 = C[2] * q0 + C[4] * q1 + C[5] * q2;
  const double v0 = mu0 - w0, v1 = mu1 - w1, v2 = mu2 - w2;
  const double Q0 = 2.0 * q0, Q1 = 2.0 * q1, Q2 = 2.0 * q2;
  const double Q3 = 2.0 * (v1 * q2 - v2 * q1), Q4 = 2.0 * (v2 * q0 - v0 * q2), Q5 = 2.0 * (v0 * q1 - v1 * q0);
  acc[ACC_G + 0] += ok ? factor * Q0 : 0.0;
  acc[ACC_G + 1] += ok ? factor * Q1 : 0.0;
  acc[ACC_G + 2] += ok ? factor * Q2 : 0.0;
  acc[ACC_G + 3] += ok ? factor * Q3 : 0.0;
  acc[ACC_G + 4] += ok ? factor * Q4 : 0.0;
  acc[ACC_G + 5] += ok ? factor * Q5 : 0.0;
}

}  // namespace ndtb
