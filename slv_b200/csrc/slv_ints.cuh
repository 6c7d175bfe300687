// slv_ints.cuh -- one-electron integrals over contracted Cartesian Gaussians (s, p, d) for the
// exchange-repulsion shift: overlap S, kinetic T = <k|-1/2 del^2|l>, and their derivatives with respect
// to the centre of the bra function contracted with a displacement vector L (the normal-mode direction
// of the bra's atom).  Replaces PyQuante.Ints.getSAB/getTAB/getSA1B/getTA1B at the reference call sites
// solvshift/solefp.py:1523-1526 (PyQuante "1.0.3-BBG" is un-vendored; convention: each primitive and each
// contracted Cartesian component normalised, see slv_b200/basis.py).  Obara-Saika 1-D recursions; the
// 4x5 table per axis is built unconditionally and the entries a given (l_bra, l_ket) needs are picked
// with selects, so the code is branch-free across mixed s/p/d pairs.  __host__ __device__ for host checks.
#pragma once
#include "slv_math.cuh"

namespace slv {

SLV_HD double sel3(int i, double v0, double v1, double v2) { return i == 0 ? v0 : (i == 1 ? v1 : v2); }

// For one axis: s(i,j) = int (x-A)^i (x-B)^j exp(-a(x-A)^2 - b(x-B)^2) dx / s(0,0).
// Returns overlap factors S0 = s(la,lb), Sp = s(la+1,lb), Sm = s(la-1,lb) and the matching kinetic
// factors t(i,lb) = b(2lb+1) s(i,lb) - 2b^2 s(i,lb+2) - lb(lb-1)/2 s(i,lb-2).
SLV_HD void axis_factors(double PA, double PB, double h, double b, int la, int lb,
                         double &S0, double &Sp, double &Sm, double &T0, double &Tp, double &Tm) {
  double s[4][5];
  s[0][0] = 1.0;
  s[0][1] = PB;
#pragma unroll
  for (int j = 1; j < 4; ++j) s[0][j + 1] = PB * s[0][j] + h * j * s[0][j - 1];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      double v = PA * s[i][j];
      if (i > 0) v += h * i * s[i - 1][j];
      if (j > 0) v += h * j * s[i][j - 1];
      s[i + 1][j] = v;
    }
  }
  // rows la-1, la, la+1 for the three columns lb-2, lb, lb+2
  double r0[3], rp[3], rm[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // column index: c=0 -> lb-2 (exists only for lb=2: column 0), c=1 -> lb, c=2 -> lb+2
    double col[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      col[i] = (c == 0) ? (lb == 2 ? s[i][0] : 0.0) : (c == 1) ? sel3(lb, s[i][0], s[i][1], s[i][2]) : sel3(lb, s[i][2], s[i][3], s[i][4]);
    r0[c] = sel3(la, col[0], col[1], col[2]);
    rp[c] = sel3(la, col[1], col[2], col[3]);
    rm[c] = sel3(la, 0.0, col[0], col[1]);
  }
  const double k0 = b * (2.0 * lb + 1.0), k2 = -2.0 * b * b, km = -0.5 * lb * (lb - 1.0);
  S0 = r0[1]; Sp = rp[1]; Sm = rm[1];
  T0 = k0 * r0[1] + k2 * r0[2] + km * r0[0];
  Tp = k0 * rp[1] + k2 * rp[2] + km * rp[0];
  Tm = k0 * rm[1] + k2 * rm[2] + km * rm[0];
}

// Integrals between contracted functions k (centre A, powers la[3], npa primitives ea/ca) and l (centre B, ...).
// out[0] = S, out[1] = T, out[2] = L . dS/dA, out[3] = L . dT/dA.
SLV_HD void ao_pair(const double A[3], const int la[3], int npa, const double *ea, const double *ca,
                    const double B[3], const int lb[3], int npb, const double *eb, const double *cb,
                    const double L[3], double out[4]) {
  const double ABx = A[0] - B[0], ABy = A[1] - B[1], ABz = A[2] - B[2];
  const double AB2 = ABx * ABx + ABy * ABy + ABz * ABz;
  double S = 0.0, T = 0.0, dS = 0.0, dT = 0.0;
  for (int p = 0; p < npa; ++p) {
    const double a = ea[p];
    for (int q = 0; q < npb; ++q) {
      const double b = eb[q];
      const double pp = a + b, ip = 1.0 / pp, h = 0.5 * ip;
      const double mu = a * b * ip;
      if (mu * AB2 > 46.0) continue;      // exp(-46) = 1e-20: below FP64 resolution of every contracted integral
      const double pref = ca[p] * cb[q] * exp(-mu * AB2) * (3.14159265358979323846 * ip) * sqrt(3.14159265358979323846 * ip);
      // P - A = -b/p (A-B) ; P - B = a/p (A-B)
      const double fa = -b * ip, fb = a * ip;
      double S0[3], Sp[3], Sm[3], T0[3], Tp[3], Tm[3];
      axis_factors(fa * ABx, fb * ABx, h, b, la[0], lb[0], S0[0], Sp[0], Sm[0], T0[0], Tp[0], Tm[0]);
      axis_factors(fa * ABy, fb * ABy, h, b, la[1], lb[1], S0[1], Sp[1], Sm[1], T0[1], Tp[1], Tm[1]);
      axis_factors(fa * ABz, fb * ABz, h, b, la[2], lb[2], S0[2], Sp[2], Sm[2], T0[2], Tp[2], Tm[2]);
      const double sxy = S0[0] * S0[1], sxz = S0[0] * S0[2], syz = S0[1] * S0[2];
      S += pref * sxy * S0[2];
      T += pref * (T0[0] * syz + T0[1] * sxz + T0[2] * sxy);
      const double a2 = 2.0 * a;
      // d/dA_x of the bra: 2a (l+1) - l (l-1)
      const double dsx = a2 * Sp[0] - la[0] * Sm[0], dsy = a2 * Sp[1] - la[1] * Sm[1], dsz = a2 * Sp[2] - la[2] * Sm[2];
      const double dtx = a2 * Tp[0] - la[0] * Tm[0], dty = a2 * Tp[1] - la[1] * Tm[1], dtz = a2 * Tp[2] - la[2] * Tm[2];
      const double gSx = dsx * syz, gSy = dsy * sxz, gSz = dsz * sxy;
      const double gTx = dtx * syz + dsx * (T0[1] * S0[2] + S0[1] * T0[2]);
      const double gTy = dty * sxz + dsy * (T0[0] * S0[2] + S0[0] * T0[2]);
      const double gTz = dtz * sxy + dsz * (T0[0] * S0[1] + S0[0] * T0[1]);
      dS += pref * (L[0] * gSx + L[1] * gSy + L[2] * gSz);
      dT += pref * (L[0] * gTx + L[1] * gTy + L[2] * gTz);
    }
  }
  out[0] = S; out[1] = T; out[2] = dS; out[3] = dT;
}

// Shell-pair form: all functions of a shell share exponents and centre, so the three 1-D tables of a primitive
// pair are built once and every (function of A, function of B) picks its entries from them.  nfa, nfb <= 6
// (S 1, P 3, D 6, L = SP 4).  la/lb: [nf][3] Cartesian powers; ca/cb: [nf][6] normalised coefficients;
// out[(fa*nfb + fb)*4 + q], q = S, T, L.dS, L.dT (overwritten).
SLV_HD void shell_pair(const double A[3], int nfa, const int *la, int npa, const double *ea, const double *ca,
                       const double B[3], int nfb, const int *lb, int npb, const double *eb, const double *cb,
                       const double L[3], double *out) {
  const double ABx = A[0] - B[0], ABy = A[1] - B[1], ABz = A[2] - B[2];
  const double AB2 = ABx * ABx + ABy * ABy + ABz * ABz;
  const int nout = nfa * nfb * 4;
  for (int w = 0; w < nout; ++w) out[w] = 0.0;
  double t[3][4][5];
  for (int p = 0; p < npa; ++p) {
    const double a = ea[p], a2 = 2.0 * a;
    for (int q = 0; q < npb; ++q) {
      const double b = eb[q];
      const double pp = a + b, ip = 1.0 / pp, h = 0.5 * ip;
      const double mu = a * b * ip;
      if (mu * AB2 > 46.0) continue;
      const double pref = exp(-mu * AB2) * (3.14159265358979323846 * ip) * sqrt(3.14159265358979323846 * ip);
      const double fa_ = -b * ip, fb_ = a * ip;
      const double AB[3] = {ABx, ABy, ABz};
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        const double PA = fa_ * AB[ax], PB = fb_ * AB[ax];
        double (*s)[5] = t[ax];
        s[0][0] = 1.0; s[0][1] = PB;
#pragma unroll
        for (int j = 1; j < 4; ++j) s[0][j + 1] = PB * s[0][j] + h * j * s[0][j - 1];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            double v = PA * s[i][j];
            if (i > 0) v += h * i * s[i - 1][j];
            if (j > 0) v += h * j * s[i][j - 1];
            s[i + 1][j] = v;
          }
      }
      const double k2 = -2.0 * b * b;
      for (int fa = 0; fa < nfa; ++fa) {
        const double cap = ca[fa * 6 + p] * pref;
        for (int fb = 0; fb < nfb; ++fb) {
          const double cc = cap * cb[fb * 6 + q];
          double S0[3], T0[3], dS[3], dT[3];
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) {
            const int i = la[fa * 3 + ax], j = lb[fb * 3 + ax];
            const double (*s)[5] = t[ax];
            const double k0 = b * (2.0 * j + 1.0), km = -0.5 * j * (j - 1.0);
            const double s0 = s[i][j], sp = s[i + 1][j], sm = i > 0 ? s[i - 1][j] : 0.0;
            const double s0p = s[i][j + 2], spp = s[i + 1][j + 2], smp = i > 0 ? s[i - 1][j + 2] : 0.0;
            const double s0m = j > 1 ? s[i][j - 2] : 0.0, spm = j > 1 ? s[i + 1][j - 2] : 0.0, smm = (i > 0 && j > 1) ? s[i - 1][j - 2] : 0.0;
            const double t0 = k0 * s0 + k2 * s0p + km * s0m;
            const double tp = k0 * sp + k2 * spp + km * spm;
            const double tm = k0 * sm + k2 * smp + km * smm;
            S0[ax] = s0; T0[ax] = t0;
            dS[ax] = a2 * sp - i * sm; dT[ax] = a2 * tp - i * tm;
          }
          const double sxy = S0[0] * S0[1], sxz = S0[0] * S0[2], syz = S0[1] * S0[2];
          const double gS = L[0] * dS[0] * syz + L[1] * dS[1] * sxz + L[2] * dS[2] * sxy;
          const double gT = L[0] * (dT[0] * syz + dS[0] * (T0[1] * S0[2] + S0[1] * T0[2]))
                          + L[1] * (dT[1] * sxz + dS[1] * (T0[0] * S0[2] + S0[0] * T0[2]))
                          + L[2] * (dT[2] * sxy + dS[2] * (T0[0] * S0[1] + S0[0] * T0[1]));
          double *o = out + (fa * nfb + fb) * 4;
          o[0] += cc * sxy * S0[2];
          o[1] += cc * (T0[0] * syz + T0[1] * sxz + T0[2] * sxy);
          o[2] += cc * gS;
          o[3] += cc * gT;
        }
      }
    }
  }
}

// ---- angular-momentum-specialised form: LA, LB = total angular momentum of the bra / ket function (0 s, 1 p, 2 d).
// The 1-D tables shrink to (LA+2) x (LB+3) entries and the entry selection to the powers that can occur; k_rep walks
// the AO pairs class by class, so a CTA executes one specialisation at a time.
template <int N>
SLV_HD double selN(int i, const double *v) {      // v[i] for i in 0..N without dynamic indexing
  if (N == 0) return v[0];
  if (N == 1) return i == 0 ? v[0] : v[1];
  return i == 0 ? v[0] : (i == 1 ? v[1] : v[2]);
}

template <int LA, int LB>
SLV_HD void axis_factors_t(double PA, double PB, double h, double b, int la, int lb,
                           double &S0, double &Sp, double &Sm, double &T0, double &Tp, double &Tm) {
  constexpr int NI = LA + 2, NJ = LB + 3;
  double s[NI][NJ];
  s[0][0] = 1.0;
  s[0][1] = PB;
#pragma unroll
  for (int j = 1; j < NJ - 1; ++j) s[0][j + 1] = PB * s[0][j] + h * j * s[0][j - 1];
#pragma unroll
  for (int i = 0; i < NI - 1; ++i) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      double v = PA * s[i][j];
      if (i > 0) v += h * i * s[i - 1][j];
      if (j > 0) v += h * j * s[i][j - 1];
      s[i + 1][j] = v;
    }
  }
  double cm[NI], c0[NI], cp[NI];                  // columns lb-2, lb, lb+2 of every row
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    double a0[3], a2[3];
#pragma unroll
    for (int j = 0; j <= LB; ++j) { a0[j] = s[i][j]; a2[j] = s[i][j + 2]; }
    c0[i] = selN<LB>(lb, a0);
    cp[i] = selN<LB>(lb, a2);
    cm[i] = (LB == 2 && lb == 2) ? s[i][0] : 0.0;
  }
  double lo0[3], mid0[3], hi0[3], lo1[3], mid1[3], hi1[3], lo2[3], mid2[3], hi2[3];
#pragma unroll
  for (int i = 0; i <= LA; ++i) {
    mid0[i] = cm[i]; hi0[i] = cm[i + 1]; lo0[i] = i > 0 ? cm[i - 1] : 0.0;
    mid1[i] = c0[i]; hi1[i] = c0[i + 1]; lo1[i] = i > 0 ? c0[i - 1] : 0.0;
    mid2[i] = cp[i]; hi2[i] = cp[i + 1]; lo2[i] = i > 0 ? cp[i - 1] : 0.0;
  }
  const double r0m = selN<LA>(la, mid0), rpm = selN<LA>(la, hi0), rmm = selN<LA>(la, lo0);
  const double r00 = selN<LA>(la, mid1), rp0 = selN<LA>(la, hi1), rm0 = selN<LA>(la, lo1);
  const double r0p = selN<LA>(la, mid2), rpp = selN<LA>(la, hi2), rmp = selN<LA>(la, lo2);
  const double k0 = b * (2.0 * lb + 1.0), k2 = -2.0 * b * b, km = -0.5 * lb * (lb - 1.0);
  S0 = r00; Sp = rp0; Sm = rm0;
  T0 = k0 * r00 + k2 * r0p + km * r0m;
  Tp = k0 * rp0 + k2 * rpp + km * rpm;
  Tm = k0 * rm0 + k2 * rmp + km * rmm;
}

template <int LA, int LB>
SLV_HD void ao_pair_t(const double A[3], const int la[3], int npa, const double *ea, const double *ca,
                      const double B[3], const int lb[3], int npb, const double *eb, const double *cb,
                      const double L[3], double out[4]) {
  const double ABx = A[0] - B[0], ABy = A[1] - B[1], ABz = A[2] - B[2];
  const double AB2 = ABx * ABx + ABy * ABy + ABz * ABz;
  double S = 0.0, T = 0.0, dS = 0.0, dT = 0.0;
  for (int p = 0; p < npa; ++p) {
    const double a = ea[p];
    for (int q = 0; q < npb; ++q) {
      const double b = eb[q];
      const double pp = a + b, ip = 1.0 / pp, h = 0.5 * ip;
      const double mu = a * b * ip;
      if (mu * AB2 > 46.0) continue;
      const double pref = ca[p] * cb[q] * exp(-mu * AB2) * (3.14159265358979323846 * ip) * sqrt(3.14159265358979323846 * ip);
      const double fa = -b * ip, fb = a * ip;
      double S0[3], Sp[3], Sm[3], T0[3], Tp[3], Tm[3];
      axis_factors_t<LA, LB>(fa * ABx, fb * ABx, h, b, la[0], lb[0], S0[0], Sp[0], Sm[0], T0[0], Tp[0], Tm[0]);
      axis_factors_t<LA, LB>(fa * ABy, fb * ABy, h, b, la[1], lb[1], S0[1], Sp[1], Sm[1], T0[1], Tp[1], Tm[1]);
      axis_factors_t<LA, LB>(fa * ABz, fb * ABz, h, b, la[2], lb[2], S0[2], Sp[2], Sm[2], T0[2], Tp[2], Tm[2]);
      const double sxy = S0[0] * S0[1], sxz = S0[0] * S0[2], syz = S0[1] * S0[2];
      S += pref * sxy * S0[2];
      T += pref * (T0[0] * syz + T0[1] * sxz + T0[2] * sxy);
      const double a2 = 2.0 * a;
      const double dsx = a2 * Sp[0] - la[0] * Sm[0], dsy = a2 * Sp[1] - la[1] * Sm[1], dsz = a2 * Sp[2] - la[2] * Sm[2];
      const double dtx = a2 * Tp[0] - la[0] * Tm[0], dty = a2 * Tp[1] - la[1] * Tm[1], dtz = a2 * Tp[2] - la[2] * Tm[2];
      const double gSx = dsx * syz, gSy = dsy * sxz, gSz = dsz * sxy;
      const double gTx = dtx * syz + dsx * (T0[1] * S0[2] + S0[1] * T0[2]);
      const double gTy = dty * sxz + dsy * (T0[0] * S0[2] + S0[0] * T0[2]);
      const double gTz = dtz * sxy + dsz * (T0[0] * S0[1] + S0[0] * T0[1]);
      dS += pref * (L[0] * gSx + L[1] * gSy + L[2] * gSz);
      dT += pref * (L[0] * gTx + L[1] * gTy + L[2] * gTz);
    }
  }
  out[0] = S; out[1] = T; out[2] = dS; out[3] = dT;
}


// ---- sub-shell-pair form (the one k_rep runs).  A sub-shell is the set of Cartesian components of one angular
// momentum on one atom sharing the primitive exponents (an SP "L" shell counts as an S and a P sub-shell).  One call
// handles all ket components and either all bra components (IA < 0; used when the ket is an s sub-shell) or the single
// bra component IA, so the exponential, the prefactor and the 1-D Obara-Saika tables of a primitive pair are computed
// once for up to six AO pairs, and every power is a compile-time constant: no table selects, no dynamic indexing.
// Contraction uses the coefficients of the FIRST component of each sub-shell; the caller rescales the other
// components (d_xy.. carry sqrt(3) relative to d_xx.., slv_b200/basis.py).
SLV_HD constexpr int cart_ncomp(int L) { return L == 0 ? 1 : (L == 1 ? 3 : 6); }
// power along axis ax of component c: s; x,y,z; xx,yy,zz,xy,xz,yz (efprot.f:17-20 order)
SLV_HD constexpr int cart_pow(int L, int c, int ax) {
  return L == 0 ? 0
       : L == 1 ? (c == ax ? 1 : 0)
       : c < 3  ? (c == ax ? 2 : 0)
       : c == 3 ? ((ax == 0 || ax == 1) ? 1 : 0)
       : c == 4 ? ((ax == 0 || ax == 2) ? 1 : 0)
                : ((ax == 1 || ax == 2) ? 1 : 0);
}

// One axis of one primitive pair.  IMAX: highest bra power handled on this axis; ALLI: values for every bra power
// 0..IMAX (else only IMAX).  Outputs [bra power slot][ket power 0..LB]: overlap s, kinetic t, and the centre
// derivatives ds = 2a s(i+1,j) - i s(i-1,j), dt likewise.
template <int IMAX, int LB, bool ALLI>
struct AxisVals {
  static constexpr int NI = ALLI ? IMAX + 1 : 1;
  double s[NI][LB + 1], t[NI][LB + 1], ds[NI][LB + 1], dt[NI][LB + 1];
  SLV_HD void build(double PA, double PB, double h, double a, double b) {
    constexpr int R = IMAX + 2, C = LB + 3;
    double e[R][C];
    e[0][0] = 1.0;
    e[0][1] = PB;
#pragma unroll
    for (int j = 1; j < C - 1; ++j) e[0][j + 1] = PB * e[0][j] + h * j * e[0][j - 1];
#pragma unroll
    for (int i = 0; i < R - 1; ++i) {
#pragma unroll
      for (int j = 0; j < C; ++j) {
        double v = PA * e[i][j];
        if (i > 0) v += h * i * e[i - 1][j];
        if (j > 0) v += h * j * e[i][j - 1];
        e[i + 1][j] = v;
      }
    }
    const double k2 = -2.0 * b * b;
    double k[R][LB + 1];                         // kinetic 1-D factor for every row
#pragma unroll
    for (int i = 0; i < R; ++i) {
#pragma unroll
      for (int j = 0; j <= LB; ++j) {
        double v = b * (2.0 * j + 1.0) * e[i][j] + k2 * e[i][j + 2];
        if (j >= 2) v += -0.5 * j * (j - 1.0) * e[i][j - 2];
        k[i][j] = v;
      }
    }
    const double a2 = 2.0 * a;
#pragma unroll
    for (int ii = 0; ii < NI; ++ii) {
      const int i = ALLI ? ii : IMAX;
#pragma unroll
      for (int j = 0; j <= LB; ++j) {
        s[ii][j] = e[i][j]; t[ii][j] = k[i][j];
        double d1 = a2 * e[i + 1][j], d2 = a2 * k[i + 1][j];
        if (i > 0) { d1 -= i * e[i - 1][j]; d2 -= i * k[i - 1][j]; }
        ds[ii][j] = d1; dt[ii][j] = d2;
      }
    }
  }
  SLV_HD void scale(double f) {
#pragma unroll
    for (int ii = 0; ii < NI; ++ii) {
#pragma unroll
      for (int j = 0; j <= LB; ++j) { s[ii][j] *= f; t[ii][j] *= f; ds[ii][j] *= f; dt[ii][j] *= f; }
    }
  }
};

// acc[(ca * NB + cb) * 4 + q] += contributions of all primitive pairs; q = S, T, L.dS/dA, L.dT/dA.
// ca = bra component slot (component index when IA < 0, else 0), cb = ket component.
template <int LA, int LB, int IA>
SLV_HD void subshell_pair_t(const double A[3], int npa, const double *ea, const double *ca,
                            const double B[3], int npb, const double *eb, const double *cb,
                            const double L[3], double *acc) {
  constexpr bool ALL = IA < 0;
  constexpr int NA = ALL ? cart_ncomp(LA) : 1, NB = cart_ncomp(LB);
  constexpr int MX = ALL ? LA : cart_pow(LA, IA, 0), MY = ALL ? LA : cart_pow(LA, IA, 1), MZ = ALL ? LA : cart_pow(LA, IA, 2);
  const double ABx = A[0] - B[0], ABy = A[1] - B[1], ABz = A[2] - B[2];
  const double AB2 = ABx * ABx + ABy * ABy + ABz * ABz;
#pragma unroll
  for (int w = 0; w < NA * NB * 4; ++w) acc[w] = 0.0;
  for (int p = 0; p < npa; ++p) {
    const double a = ea[p];
    for (int q = 0; q < npb; ++q) {
      const double b = eb[q];
      const double pp = a + b, ip = 1.0 / pp, h = 0.5 * ip;
      const double mu = a * b * ip;
      if (mu * AB2 > 46.0) continue;
      const double pref = ca[p] * cb[q] * exp(-mu * AB2) * (3.14159265358979323846 * ip) * sqrt(3.14159265358979323846 * ip);
      const double fa = -b * ip, fb = a * ip;
      AxisVals<MX, LB, ALL> X; AxisVals<MY, LB, ALL> Y; AxisVals<MZ, LB, ALL> Z;
      X.build(fa * ABx, fb * ABx, h, a, b); X.scale(pref);       // every term below carries exactly one x factor
      Y.build(fa * ABy, fb * ABy, h, a, b);
      Z.build(fa * ABz, fb * ABz, h, a, b);
#pragma unroll
      for (int c1 = 0; c1 < NA; ++c1) {
        const int cc = ALL ? c1 : IA;
        const int ix = ALL ? cart_pow(LA, cc, 0) : 0, iy = ALL ? cart_pow(LA, cc, 1) : 0, iz = ALL ? cart_pow(LA, cc, 2) : 0;
#pragma unroll
        for (int c2 = 0; c2 < NB; ++c2) {
          const int jx = cart_pow(LB, c2, 0), jy = cart_pow(LB, c2, 1), jz = cart_pow(LB, c2, 2);
          const double sx = X.s[ix][jx], sy = Y.s[iy][jy], sz = Z.s[iz][jz];
          const double tx = X.t[ix][jx], ty = Y.t[iy][jy], tz = Z.t[iz][jz];
          const double dsx = L[0] * X.ds[ix][jx], dsy = L[1] * Y.ds[iy][jy], dsz = L[2] * Z.ds[iz][jz];
          const double dtx = L[0] * X.dt[ix][jx], dty = L[1] * Y.dt[iy][jy], dtz = L[2] * Z.dt[iz][jz];
          const double sxy = sx * sy, syz = sy * sz;
          const double tyz = ty * sz + sy * tz;                 // (T_y S_z + S_y T_z)
          double *o = acc + (c1 * NB + c2) * 4;
          o[0] += sxy * sz;
          o[1] += tx * syz + sx * tyz;
          o[2] += dsx * syz + sx * (dsy * sz + sy * dsz);
          o[3] += dtx * syz + dsx * tyz + sx * (dty * sz + dsy * tz + sy * dtz + ty * dsz) + tx * (dsy * sz + sy * dsz);
        }
      }
    }
  }
}

// One work item of k_rep phase 1: sub-shells sa = {atom, first function, L, nprim} of the solute and sb of the
// solvent fragment; writes V[((f0a + ca) * nbsB + f0b + cb) * 4 + q].  eX/cX are the per-function [nbf][6] exponent and
// coefficient tables (coefficient x primitive norm x contracted norm), LcA the per-atom displacement vectors.
template <int LA, int LB, int IA>
SLV_HD void subshell_pair_store(const int *sa, const int *sb, const double *posA, const double *posB,
                                const double *eA, const double *cA, const double *eB, const double *cB,
                                const double *LcA, int nbsB, double *V) {
  constexpr int NA = IA < 0 ? cart_ncomp(LA) : 1, NB = cart_ncomp(LB);
  const int f0a = sa[1], f0b = sb[1];
  double acc[NA * NB * 4];
  subshell_pair_t<LA, LB, IA>(posA + sa[0] * 3, sa[3], eA + f0a * 6, cA + f0a * 6, posB + sb[0] * 3, sb[3], eB + f0b * 6, cB + f0b * 6,
                              LcA + sa[0] * 3, acc);
  const double ia0 = 1.0 / cA[f0a * 6], ib0 = 1.0 / cB[f0b * 6];
#pragma unroll
  for (int c1 = 0; c1 < NA; ++c1) {
    const int k = f0a + (IA < 0 ? c1 : IA);
    const double ra = (LA == 2) ? cA[k * 6] * ia0 : 1.0;          // only d components differ in norm within a sub-shell
#pragma unroll
    for (int c2 = 0; c2 < NB; ++c2) {
      const int l = f0b + c2;
      const double f = (LB == 2) ? ra * (cB[l * 6] * ib0) : ra;
      double *v = V + ((size_t)k * nbsB + l) * 4;
      v[0] = f * acc[(c1 * NB + c2) * 4]; v[1] = f * acc[(c1 * NB + c2) * 4 + 1];
      v[2] = f * acc[(c1 * NB + c2) * 4 + 2]; v[3] = f * acc[(c1 * NB + c2) * 4 + 3];
    }
  }
}

// The 23 (LA, LB, IA) specialisations in the order of the plan's work list (slv_b200/plan.py, IO_SHPAIR):
// ket s: all bra components at once; ket p and d: one bra component per item.
#define SLV_SUBSHELL_CLASSES(X) \
  X(0, 0, 0, -1) X(1, 1, 0, -1) X(2, 2, 0, -1) \
  X(3, 0, 1, 0) X(4, 1, 1, 0) X(5, 1, 1, 1) X(6, 1, 1, 2) \
  X(7, 2, 1, 0) X(8, 2, 1, 1) X(9, 2, 1, 2) X(10, 2, 1, 3) X(11, 2, 1, 4) X(12, 2, 1, 5) \
  X(13, 0, 2, 0) X(14, 1, 2, 0) X(15, 1, 2, 1) X(16, 1, 2, 2) \
  X(17, 2, 2, 0) X(18, 2, 2, 1) X(19, 2, 2, 2) X(20, 2, 2, 3) X(21, 2, 2, 4) X(22, 2, 2, 5)
constexpr int SUBSHELL_NCLASS = 23;

}  // namespace slv
