// slv_math.cuh -- FP64 device/host math of the SolEFP path: Kabsch superimposition, tensor
// rotation in reduced storage, pair tensors.  Everything is __host__ __device__ so the same
// code can be unit-checked on the build host (tests/host_math_check.cpp) before it runs on the GPU.
//
// Conventions (reference: solvshift/slvpar.py:682-687, solvshift/src/efprot.f:297-320):
//   row vectors, x' = x.rot + t;  mu'_i = mu_a R_ai;  Q'_ij = Q_ab R_ai R_bj;
//   O'_ijk = O_abc R_ai R_bj R_ck;  alpha' = R^T alpha R.
//   Q: XX YY ZZ XY XZ YZ ; O: XXX YYY ZZZ XXY XXZ XYY YYZ XZZ YZZ XYZ.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define SLV_HD __host__ __device__ __forceinline__
#else
#define SLV_HD inline
#endif

namespace slv {

SLV_HD constexpr int sym2(int a, int b) { return a == b ? a : a + b + 2; }
SLV_HD constexpr int sym3(int a, int b, int c) {
  // counts of x,y,z -> reduced octupole slot
  const int nx = (a == 0) + (b == 0) + (c == 0), ny = (a == 1) + (b == 1) + (c == 1);
  const int key = nx * 4 + ny;
  return key == 12 ? 0 : key == 3 ? 1 : key == 0 ? 2 : key == 9 ? 3 : key == 8 ? 4 : key == 6 ? 5
       : key == 2 ? 6 : key == 4 ? 7 : key == 1 ? 8 : 9;
}

// ---------------------------------------------------------------------------------------------
// Kabsch: optimal proper rotation (row-vector convention) taking centred parameter coordinates
// c_k onto centred target coordinates r_k, from the 3x3 correlation a[i][j] = sum_k c_k[i] r_k[j].
// Horn's quaternion form; the dominant eigenvector of the 4x4 symmetric matrix is found with
// cyclic Jacobi sweeps (robust for planar / near-degenerate atom sets such as water).  Gives the
// same rotation as the SVD route with the det<0 reflection fix (slvpar.py:646-652).
SLV_HD void kabsch_rot_jacobi(const double a[9], double rot[9]) {
  const double Sxx = a[0], Sxy = a[1], Sxz = a[2], Syx = a[3], Syy = a[4], Syz = a[5], Szx = a[6], Szy = a[7], Szz = a[8];
  double N[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                    {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                    {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                    {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0, dia = 0.0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      dia += N[p][p] * N[p][p];
#pragma unroll
      for (int q = p + 1; q < 4; ++q) off += N[p][q] * N[p][q];
    }
    if (off <= 1e-32 * dia || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        const double apq = N[p][q];
        if (apq == 0.0) continue;
        const double theta = (N[q][q] - N[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 4; ++k) {            // N <- N J  (columns p,q)
          const double nkp = N[k][p], nkq = N[k][q];
          N[k][p] = c * nkp - s * nkq; N[k][q] = s * nkp + c * nkq;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {            // N <- J^T N (rows p,q)
          const double npk = N[p][k], nqk = N[q][k];
          N[p][k] = c * npk - s * nqk; N[q][k] = s * npk + c * nqk;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
        }
      }
    }
  }
  int im = 0;
#pragma unroll
  for (int p = 1; p < 4; ++p) if (N[p][p] > N[im][im]) im = p;
  double q0 = V[0][0], q1 = V[1][0], q2 = V[2][0], q3 = V[3][0];
  if (im == 1) { q0 = V[0][1]; q1 = V[1][1]; q2 = V[2][1]; q3 = V[3][1]; }
  if (im == 2) { q0 = V[0][2]; q1 = V[1][2]; q2 = V[2][2]; q3 = V[3][2]; }
  if (im == 3) { q0 = V[0][3]; q1 = V[1][3]; q2 = V[2][3]; q3 = V[3][3]; }
  const double nrm = 1.0 / sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  q0 *= nrm; q1 *= nrm; q2 *= nrm; q3 *= nrm;
  // column-vector rotation M (r = M c); the row-vector matrix is its transpose: rot[a][i] = M[i][a]
  const double M00 = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, M01 = 2.0 * (q1 * q2 - q0 * q3), M02 = 2.0 * (q1 * q3 + q0 * q2);
  const double M10 = 2.0 * (q1 * q2 + q0 * q3), M11 = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, M12 = 2.0 * (q2 * q3 - q0 * q1);
  const double M20 = 2.0 * (q1 * q3 - q0 * q2), M21 = 2.0 * (q2 * q3 + q0 * q1), M22 = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
  rot[0] = M00; rot[1] = M10; rot[2] = M20;
  rot[3] = M01; rot[4] = M11; rot[5] = M21;
  rot[6] = M02; rot[7] = M12; rot[8] = M22;
}

// Fast path of the same problem: the dominant eigenvector of Horn's matrix N by shifted inverse iteration.
// With GA = sum|c_k|^2 and GB = sum|r_k|^2 (passed as gsum = GA + GB) the largest eigenvalue obeys
// lambda_1 <= gsum/2, with equality for a perfect fit, so mu = gsum/2 (1 + 1e-7) makes mu I - N symmetric
// positive definite: one 4x4 Cholesky factorisation, then a few triangular solves; the iteration contracts by
// (mu - lambda_1)/(mu - lambda_2) per step (<< 1 for rigid fragments).  Convergence is verified on the
// eigen-residual; anything that does not converge (poor fits, near-degenerate atom sets) goes to the
// Jacobi solver above, so the result is the same rotation either way.
SLV_HD void kabsch_rot(const double a[9], double gsum, double rot[9]) {
  const double Sxx = a[0], Sxy = a[1], Sxz = a[2], Syx = a[3], Syy = a[4], Syz = a[5], Szx = a[6], Szy = a[7], Szz = a[8];
  const double n00 = Sxx + Syy + Szz, n01 = Syz - Szy, n02 = Szx - Sxz, n03 = Sxy - Syx;
  const double n11 = Sxx - Syy - Szz, n12 = Sxy + Syx, n13 = Szx + Sxz;
  const double n22 = -Sxx + Syy - Szz, n23 = Syz + Szy, n33 = -Sxx - Syy + Szz;
  const double mu = 0.5 * gsum * (1.0 + 1e-7) + 1e-300;
  // Cholesky of M = mu I - N = L L^T
  const double m00 = mu - n00, m11 = mu - n11, m22 = mu - n22, m33 = mu - n33;
  bool ok = m00 > 0.0;
  const double l00 = sqrt(m00), i00 = 1.0 / l00;
  const double l10 = -n01 * i00, l20 = -n02 * i00, l30 = -n03 * i00;
  const double d1 = m11 - l10 * l10;
  ok = ok && d1 > 0.0;
  const double l11 = sqrt(d1), i11 = 1.0 / l11;
  const double l21 = (-n12 - l20 * l10) * i11, l31 = (-n13 - l30 * l10) * i11;
  const double d2 = m22 - l20 * l20 - l21 * l21;
  ok = ok && d2 > 0.0;
  const double l22 = sqrt(d2), i22 = 1.0 / l22;
  const double l32 = (-n23 - l30 * l20 - l31 * l21) * i22;
  const double d3 = m33 - l30 * l30 - l31 * l31 - l32 * l32;
  ok = ok && d3 > 0.0;
  const double i33 = 1.0 / sqrt(d3);
  double q0 = 1.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
  if (ok) {
    // start from the row of adj-like guess: any vector not orthogonal to v1; use (1,1,1,1) mixed with e0
    q0 = 1.0; q1 = 0.37; q2 = -0.23; q3 = 0.11;
    for (int it = 0; it < 4; ++it) {
      // L y = q
      const double y0 = q0 * i00, y1 = (q1 - l10 * y0) * i11, y2 = (q2 - l20 * y0 - l21 * y1) * i22;
      const double y3 = (q3 - l30 * y0 - l31 * y1 - l32 * y2) * i33;
      // L^T x = y
      const double x3 = y3 * i33, x2 = (y2 - l32 * x3) * i22, x1 = (y1 - l21 * x2 - l31 * x3) * i11;
      const double x0 = (y0 - l10 * x1 - l20 * x2 - l30 * x3) * i00;
      const double nr = 1.0 / sqrt(x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3);
      q0 = x0 * nr; q1 = x1 * nr; q2 = x2 * nr; q3 = x3 * nr;
    }
    // eigen-residual check
    const double w0 = n00 * q0 + n01 * q1 + n02 * q2 + n03 * q3, w1 = n01 * q0 + n11 * q1 + n12 * q2 + n13 * q3;
    const double w2 = n02 * q0 + n12 * q1 + n22 * q2 + n23 * q3, w3 = n03 * q0 + n13 * q1 + n23 * q2 + n33 * q3;
    const double lam = q0 * w0 + q1 * w1 + q2 * w2 + q3 * w3;
    const double r0 = w0 - lam * q0, r1 = w1 - lam * q1, r2 = w2 - lam * q2, r3 = w3 - lam * q3;
    const double res2 = r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
    ok = res2 <= 1e-28 * mu * mu && lam > 0.0;
  }
  if (!ok) { kabsch_rot_jacobi(a, rot); return; }
  const double M00 = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, M01 = 2.0 * (q1 * q2 - q0 * q3), M02 = 2.0 * (q1 * q3 + q0 * q2);
  const double M10 = 2.0 * (q1 * q2 + q0 * q3), M11 = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, M12 = 2.0 * (q2 * q3 - q0 * q1);
  const double M20 = 2.0 * (q1 * q3 - q0 * q2), M21 = 2.0 * (q2 * q3 + q0 * q1), M22 = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
  rot[0] = M00; rot[1] = M10; rot[2] = M20;
  rot[3] = M01; rot[4] = M11; rot[5] = M21;
  rot[6] = M02; rot[7] = M12; rot[8] = M22;
}

// x' = x.rot (+ t)
SLV_HD void rot_vec(const double R[9], const double x[3], double y[3]) {
  y[0] = x[0] * R[0] + x[1] * R[3] + x[2] * R[6];
  y[1] = x[0] * R[1] + x[1] * R[4] + x[2] * R[7];
  y[2] = x[0] * R[2] + x[1] * R[5] + x[2] * R[8];
}

// reduced symmetric rank-2: Q'_ij = Q_ab R_ai R_bj
SLV_HD void rot_quad(const double R[9], const double q[6], double o[6]) {
  double t[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      t[a][j] = q[sym2(a, 0)] * R[0 * 3 + j] + q[sym2(a, 1)] * R[1 * 3 + j] + q[sym2(a, 2)] * R[2 * 3 + j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j)
      o[sym2(i, j)] = R[0 * 3 + i] * t[0][j] + R[1 * 3 + i] * t[1][j] + R[2 * 3 + i] * t[2][j];
}

// reduced symmetric rank-3: O'_ijk = O_abc R_ai R_bj R_ck, staged contraction
SLV_HD void rot_oct(const double R[9], const double o[10], double out[10]) {
  double t1[6][3];                       // t1[(ab)][k] = sum_c O_abc R_ck
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = a; b < 3; ++b)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        t1[sym2(a, b)][k] = o[sym3(a, b, 0)] * R[0 * 3 + k] + o[sym3(a, b, 1)] * R[1 * 3 + k] + o[sym3(a, b, 2)] * R[2 * 3 + k];
  double t2[3][6];                       // t2[a][(jk)] = sum_b t1[(ab)][k] R_bj   (symmetric in j,k)
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = j; k < 3; ++k)
        t2[a][sym2(j, k)] = t1[sym2(a, 0)][k] * R[0 * 3 + j] + t1[sym2(a, 1)][k] * R[1 * 3 + j] + t1[sym2(a, 2)][k] * R[2 * 3 + j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j)
#pragma unroll
      for (int k = j; k < 3; ++k)
        out[sym3(i, j, k)] = R[0 * 3 + i] * t2[0][sym2(j, k)] + R[1 * 3 + i] * t2[1][sym2(j, k)] + R[2 * 3 + i] * t2[2][sym2(j, k)];
}

// general 3x3 (row-major): alpha' = R^T alpha R
SLV_HD void rot_mat(const double R[9], const double a[9], double o[9]) {
  double t[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      t[i * 3 + j] = a[i * 3 + 0] * R[0 * 3 + j] + a[i * 3 + 1] * R[1 * 3 + j] + a[i * 3 + 2] * R[2 * 3 + j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      o[i * 3 + j] = R[0 * 3 + i] * t[0 * 3 + j] + R[1 * 3 + i] * t[1 * 3 + j] + R[2 * 3 + i] * t[2 * 3 + j];
}

// rotate one 20-component moment row (q, mu3, Theta6, Omega10)
SLV_HD void rot_mom20(const double R[9], const double m[20], double o[20]) {
  o[0] = m[0];
  rot_vec(R, m + 1, o + 1);
  rot_quad(R, m + 4, o + 4);
  rot_oct(R, m + 10, o + 10);
}

// 1/sqrt(x) in full FP64 accuracy
SLV_HD double rsqrt64(double x) {
#ifdef __CUDA_ARCH__
  double y = rsqrt(x);
  return y;
#else
  return 1.0 / sqrt(x);
#endif
}

// 1/sqrt(x) for the hot pair loops: hardware seed (MUFU.RSQ64H, ~2^-20) refined by one third-order step
// y <- y + y e (1/2 + 3/8 e), e = 1 - x y^2, i.e. relative error ~e^3 ~ 1e-18: full FP64 accuracy, without
// the range / denormal guard of the library routine (arguments here are squared inter-site distances,
// always normal numbers).
SLV_HD double rsqrt_fast(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);
  const double pcoef = fma(e, 0.375, 0.5);
  return fma(pcoef * e, y, y);
#else
  return 1.0 / sqrt(x);
#endif
}

// ---------------------------------------------------------------------------------------------
// Electrostatic pair kernel.  For a solvent site B (lab-frame traceless moments mB[20] at rB) and a
// solute site at rA, builds the 23 "potential derivative" quantities P so that for any solute
// moment row sA (multiplicity-folded, see fold20) the interaction energy through R^-4
//   E = qA qB T + T_a(qA muB_a - muA_a qB) + T_ab((qA ThB_ab + ThA_ab qB)/3 - muA_a muB_b)
//       + T_abc((qA OmB_abc - OmA_abc qB)/15 - muA_a ThB_bc/3 + ThA_ab muB_c/3)
// (the truncation libbbg sdmtpm applies; reference call site solvshift/solefp.py:1175) is dot(sA, P[0..19]),
// and P[20..22] is the gradient at rA of the potential of the CHARGE of B alone (dmtcor rf2/rk2 terms).
SLV_HD void pair_potentials(const double rA[3], const double rB[3], const double mB[20], double P[23]) {
  const double Rx = rB[0] - rA[0], Ry = rB[1] - rA[1], Rz = rB[2] - rA[2];
  const double r2 = Rx * Rx + Ry * Ry + Rz * Rz;
  const double u = rsqrt_fast(r2), u2 = u * u, u3 = u2 * u, u5 = u3 * u2, u7 = u5 * u2;
  const double qB = mB[0], mx = mB[1], my = mB[2], mz = mB[3];
  const double *th = mB + 4, *om = mB + 10;
  const double RR[6] = {Rx * Rx, Ry * Ry, Rz * Rz, Rx * Ry, Rx * Rz, Ry * Rz};
  const double RRR[10] = {RR[0] * Rx, RR[1] * Ry, RR[2] * Rz, RR[0] * Ry, RR[0] * Rz, RR[1] * Rx, RR[1] * Rz, RR[2] * Rx, RR[2] * Ry, RR[3] * Rz};
  const double mR = mx * Rx + my * Ry + mz * Rz;
  const double tRx = th[0] * Rx + th[3] * Ry + th[4] * Rz;
  const double tRy = th[3] * Rx + th[1] * Ry + th[5] * Rz;
  const double tRz = th[4] * Rx + th[5] * Ry + th[2] * Rz;
  const double tRR = tRx * Rx + tRy * Ry + tRz * Rz;
  const double oRRR = om[0] * RRR[0] + om[1] * RRR[1] + om[2] * RRR[2]
                    + 3.0 * (om[3] * RRR[3] + om[4] * RRR[4] + om[5] * RRR[5] + om[6] * RRR[6] + om[7] * RRR[7] + om[8] * RRR[8])
                    + 6.0 * om[9] * RRR[9];
  P[0] = qB * u - mR * u3 + tRR * u5 - oRRR * u7;
  const double s1 = qB * u3 - 3.0 * mR * u5 + 5.0 * tRR * u7, t5 = 2.0 * u5;
  P[1] = Rx * s1 + mx * u3 - tRx * t5;
  P[2] = Ry * s1 + my * u3 - tRy * t5;
  P[3] = Rz * s1 + mz * u3 - tRz * t5;
  const double c2 = qB * u5 - 5.0 * mR * u7;
  P[4] = c2 * RR[0] + t5 * (Rx * mx);
  P[5] = c2 * RR[1] + t5 * (Ry * my);
  P[6] = c2 * RR[2] + t5 * (Rz * mz);
  P[7] = c2 * RR[3] + u5 * (Rx * my + Ry * mx);
  P[8] = c2 * RR[4] + u5 * (Rx * mz + Rz * mx);
  P[9] = c2 * RR[5] + u5 * (Ry * mz + Rz * my);
  const double c3 = qB * u7;
#pragma unroll
  for (int k = 0; k < 10; ++k) P[10 + k] = c3 * RRR[k];
  const double g = qB * u3;
  P[20] = g * Rx; P[21] = g * Ry; P[22] = g * Rz;
}

// fold the multiplicities of the reduced storage into a solute moment row so that the full
// contraction over symmetric tensors becomes a plain 20-term dot product with P
SLV_HD void fold20(double s[20]) {
  s[7] *= 2.0; s[8] *= 2.0; s[9] *= 2.0;
#pragma unroll
  for (int k = 13; k < 19; ++k) s[k] *= 3.0;
  s[19] *= 6.0;
}

// ---------------------------------------------------------------------------------------------
// Polarisation path (solvshift/src/solpol2.f).  m[20] = q, mu, Theta, Omega (traceless, reduced).
// out_a = sum_bc Omega_abc p_b q_c
SLV_HD void oct_bilin(const double om[10], const double p[3], const double q[3], double out[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
      for (int c = 0; c < 3; ++c) s += om[sym3(a, b, c)] * p[b] * q[c];
    out[a] = s;
  }
}
SLV_HD void quad_vec(const double th[6], const double p[3], double out[3]) {
  out[0] = th[0] * p[0] + th[3] * p[1] + th[4] * p[2];
  out[1] = th[3] * p[0] + th[1] * p[1] + th[5] * p[2];
  out[2] = th[4] * p[0] + th[5] * p[1] + th[2] * p[2];
}

// Field at displacement R = r_point - r_site of one site.  octfac = 5 reproduces FFFFFF
// (solpol2.f:1408-1429, the reference's own prefactor); octfac = 7 is the value AAAAAA uses for
// the derivative moments (solpol2.f:1096-1117).  Accumulates into F.
SLV_HD void site_field(const double R[3], const double m[20], double octfac, double F[3]) {
  const double r2 = R[0] * R[0] + R[1] * R[1] + R[2] * R[2];
  const double u = rsqrt64(r2), u2 = u * u, u3 = u2 * u, u5 = u3 * u2, u7 = u5 * u2, u9 = u7 * u2;
  const double mR = m[1] * R[0] + m[2] * R[1] + m[3] * R[2];
  double tR[3], oRR[3];
  quad_vec(m + 4, R, tR);
  oct_bilin(m + 10, R, R, oRR);
  const double tRR = tR[0] * R[0] + tR[1] * R[1] + tR[2] * R[2];
  const double oRRR = oRR[0] * R[0] + oRR[1] * R[1] + oRR[2] * R[2];
  const double sR = m[0] * u3 + 3.0 * mR * u5 + 5.0 * tRR * u7 + octfac * oRRR * u9;
#pragma unroll
  for (int a = 0; a < 3; ++a) F[a] += sR * R[a] - u3 * m[1 + a] - 2.0 * u5 * tR[a] - 3.0 * u7 * oRR[a];
}

// Directional derivative along v (of the field point) of the physical field (octupole factor 7),
// with the one slip of AAAAAA reproduced: in SUM2 = Omega_abc v_a R_b R_c the term O_xzz v_z R_x R_z
// is counted once, not twice (solpol2.f:466-483 and 959-976).  Accumulates sign*dF into F.
SLV_HD void site_dfield(const double R[3], const double v[3], const double m[20], double sign, double F[3]) {
  const double r2 = R[0] * R[0] + R[1] * R[1] + R[2] * R[2];
  const double u = rsqrt64(r2), u2 = u * u, u3 = u2 * u, u5 = u3 * u2, u7 = u5 * u2, u9 = u7 * u2;
  const double Rv = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  const double mR = m[1] * R[0] + m[2] * R[1] + m[3] * R[2];
  const double mv = m[1] * v[0] + m[2] * v[1] + m[3] * v[2];
  double tR[3], tv[3], oRR[3], ovR[3];
  quad_vec(m + 4, R, tR);
  quad_vec(m + 4, v, tv);
  oct_bilin(m + 10, R, R, oRR);
  oct_bilin(m + 10, v, R, ovR);
  const double RtR = tR[0] * R[0] + tR[1] * R[1] + tR[2] * R[2];
  const double vtR = tR[0] * v[0] + tR[1] * v[1] + tR[2] * v[2];
  const double oRRR = oRR[0] * R[0] + oRR[1] * R[1] + oRR[2] * R[2];
  const double sum2 = oRR[0] * v[0] + oRR[1] * v[1] + oRR[2] * v[2] - m[10 + 7] * v[2] * R[0] * R[2];
  // coefficients of R, v, mu, ThR, Thv, OmRR, OmvR
  const double cR = -3.0 * m[0] * u3 * Rv * u2 + 3.0 * u5 * (mv - 5.0 * u2 * mR * Rv)
                  + u5 * (10.0 * u2 * vtR - 35.0 * Rv * RtR * u2 * u2) + 7.0 * u9 * (3.0 * sum2 - 9.0 * u2 * oRRR * Rv);
  const double cv = m[0] * u3 + 3.0 * u5 * mR + 5.0 * u7 * RtR + 7.0 * u9 * oRRR;
  const double cm = 3.0 * u5 * Rv, ctR = 10.0 * u7 * Rv, ctv = -2.0 * u5, coRR = 21.0 * u9 * Rv, covR = -6.0 * u7;
#pragma unroll
  for (int a = 0; a < 3; ++a)
    F[a] += sign * (cR * R[a] + cv * v[a] + cm * m[1 + a] + ctR * tR[a] + ctv * tv[a] + coRR * oRR[a] + covR * ovR[a]);
}

// Lean form of site_field for the hot field loop of k_pol: same arithmetic, octupole contractions through
// the six R_b R_c monomials.  `w` masks the contribution (0 for sites of the centre's own molecule).
SLV_HD void site_field_acc(const double R[3], const double m[20], double octfac, double w, double F[3]) {
  const double r2 = R[0] * R[0] + R[1] * R[1] + R[2] * R[2];
  const double u = rsqrt_fast(r2), u2 = u * u, u3 = w * u2 * u, u5 = u3 * u2, u7 = u5 * u2, u9 = u7 * u2;
  const double xx = R[0] * R[0], yy = R[1] * R[1], zz = R[2] * R[2];
  const double xy = 2.0 * R[0] * R[1], xz = 2.0 * R[0] * R[2], yz = 2.0 * R[1] * R[2];
  const double *th = m + 4, *om = m + 10;
  const double mR = m[1] * R[0] + m[2] * R[1] + m[3] * R[2];
  const double tRx = th[0] * R[0] + th[3] * R[1] + th[4] * R[2];
  const double tRy = th[3] * R[0] + th[1] * R[1] + th[5] * R[2];
  const double tRz = th[4] * R[0] + th[5] * R[1] + th[2] * R[2];
  const double tRR = tRx * R[0] + tRy * R[1] + tRz * R[2];
  // O: XXX0 YYY1 ZZZ2 XXY3 XXZ4 XYY5 YYZ6 XZZ7 YZZ8 XYZ9
  const double oRRx = om[0] * xx + om[5] * yy + om[7] * zz + om[3] * xy + om[4] * xz + om[9] * yz;
  const double oRRy = om[3] * xx + om[1] * yy + om[8] * zz + om[5] * xy + om[9] * xz + om[6] * yz;
  const double oRRz = om[4] * xx + om[6] * yy + om[2] * zz + om[9] * xy + om[7] * xz + om[8] * yz;
  const double oRRR = oRRx * R[0] + oRRy * R[1] + oRRz * R[2];
  const double sR = m[0] * u3 + 3.0 * mR * u5 + 5.0 * tRR * u7 + octfac * oRRR * u9;
  const double c5 = 2.0 * u5, c7 = 3.0 * u7;
  F[0] += sR * R[0] - u3 * m[1] - c5 * tRx - c7 * oRRx;
  F[1] += sR * R[1] - u3 * m[2] - c5 * tRy - c7 * oRRy;
  F[2] += sR * R[2] - u3 * m[3] - c5 * tRz - c7 * oRRz;
}

}  // namespace slv
