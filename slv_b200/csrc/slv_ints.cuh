// slv_ints.cuh -- one-electron integrals over contracted Cartesian Gaussians (s, p, d) for the
// exchange-repulsion shift: overlap S, kinetic T = <k|-1/2 del^2|l>, and their derivatives with respect
// to the centre of the bra function contracted with a displacement vector L (the normal-mode direction
// of the bra's atom).  Replaces PyQuante.Ints.getSAB/getTAB/getSA1B/getTA1B at the reference call sites
// solvshift/solefp.py:1523-1526 (PyQuante "1.0.3-BBG" is un-vendored; convention: each primitive and each
// contracted Cartesian component normalised, see slv_b200/basis.py).  Obara-Saika 1-D recursions, evaluated per
// pair of sub-shells with every angular power a template constant.  __host__ __device__: tests/host_math.cpp builds
// the same code for the CPU check against oracle/ints_oracle.py.
#pragma once
#include "slv_math.cuh"

namespace slv {

// A sub-shell is the set of Cartesian components of one angular
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

// Record of one primitive pair of a sub-shell pair, tabulated when the plan is built (slv_b200/plan.py
// primitive_pair_records): thr = 46 (a + b) / (a b), mu = a b / (a + b), pref0 = c_a c_b (pi / (a + b))^(3/2) with the
// coefficients of the first component of each sub-shell, ip = 1 / (a + b), a, b.  The records of a sub-shell pair are
// sorted by thr, largest first: the pairs with exp(-mu AB^2) >= exp(-46) at a distance AB are a prefix of the list, so a
// thread walks ITS survivors and stops at the first record whose threshold the distance exceeds -- no division, square
// root or screening loop per evaluation.
constexpr int PP_REC = 6;
constexpr int PP_SREC = 2;    // leading records of an item the integral kernel keeps in shared memory

// acc[(ca * NB + cb) * 4 + q] += contributions of all primitive pairs; q = S, T, L.dS/dA, L.dT/dA.
// ca = bra component slot (component index when IA < 0, else 0), cb = ket component.
template <int LA, int LB, int IA>
SLV_HD void subshell_pair_t(const double A[3], const double B[3], const double L[3], int npp, const double *__restrict__ rec,
                            const double *srec, int sstride, double *acc) {
  constexpr bool ALL = IA < 0;
  constexpr int NA = ALL ? cart_ncomp(LA) : 1, NB = cart_ncomp(LB);
  constexpr int MX = ALL ? LA : cart_pow(LA, IA, 0), MY = ALL ? LA : cart_pow(LA, IA, 1), MZ = ALL ? LA : cart_pow(LA, IA, 2);
  const double ABx = A[0] - B[0], ABy = A[1] - B[1], ABz = A[2] - B[2];
  const double AB2 = ABx * ABx + ABy * ABy + ABz * ABz;
#pragma unroll
  for (int w = 0; w < NA * NB * 4; ++w) acc[w] = 0.0;
  // records k < PP_SREC come from srec[(k * PP_REC + field) * sstride] (the kernel's shared-memory copy), the rest from rec
  for (int k = 0; k < npp; ++k) {
    double thr, mu, pref0, ip, a, b;
    if (srec != nullptr && k < PP_SREC) {                  // two separate paths: a pointer select would turn both into generic loads
      const double *r = srec + (size_t)k * PP_REC * sstride;
      thr = r[0]; mu = r[sstride]; pref0 = r[2 * sstride]; ip = r[3 * sstride]; a = r[4 * sstride]; b = r[5 * sstride];
    } else {
      const double *r = rec + (size_t)k * PP_REC;
      thr = r[0]; mu = r[1]; pref0 = r[2]; ip = r[3]; a = r[4]; b = r[5];
    }
    if (!(AB2 <= thr)) break;
    {
      const double pref = pref0 * exp(-mu * AB2);
      if constexpr (LA == 0 && LB == 0) {
        // s|s in closed form (a fifth of all items): S = pref, T = mu (3 - 2 mu AB^2) S, dS/dA = -2 mu (A - B) S,
        // dT/dA = -2 mu (A - B) (T + 2 mu S) -- a dozen live values instead of the three 1-D tables, so twice as many
        // warps fit an SM
        const double LAB = L[0] * ABx + L[1] * ABy + L[2] * ABz;
        const double T = mu * (3.0 - 2.0 * mu * AB2) * pref, g = -2.0 * mu * LAB;
        acc[0] += pref; acc[1] += T; acc[2] += g * pref; acc[3] += g * (T + 2.0 * mu * pref);
        (void)ip; (void)a; (void)b;
        continue;
      }
      const double h = 0.5 * ip;
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

// Norm of component c of a sub-shell relative to its first component: only d components differ, d_xy, d_xz, d_yz carry
// sqrt(3) relative to d_xx, d_yy, d_zz (primitive norms ~ 1/sqrt((2l-1)!!(2m-1)!!(2n-1)!!), equal contracted norms;
// slv_b200/basis.py) -- a compile-time constant, so the contraction can use the coefficients of the first component.
SLV_HD constexpr double cart_norm_ratio(int L, int c) { return (L == 2 && c >= 3) ? 1.7320508075688772 : 1.0; }

// Scale the contracted sums of one sub-shell pair and write them to V[((f0a + ca) * LP + f0b + cb) * 4 + q]
// (LP records of four kinds per solute basis function).
template <int LA, int LB, int IA>
SLV_HD void subshell_pair_write(const double *acc, int f0a, int f0b, int LP, double *V) {
  constexpr int NA = IA < 0 ? cart_ncomp(LA) : 1, NB = cart_ncomp(LB);
#pragma unroll
  for (int c1 = 0; c1 < NA; ++c1) {
    const int ka = IA < 0 ? c1 : IA, k = f0a + ka;
#pragma unroll
    for (int c2 = 0; c2 < NB; ++c2) {
      const int l = f0b + c2;
      const double f = cart_norm_ratio(LA, ka) * cart_norm_ratio(LB, c2);
      double *v = V + ((size_t)k * LP + l) * 4;                   // V[k][l][q], LP records per row
      double x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double y = (f == 1.0) ? acc[(c1 * NB + c2) * 4 + q] : f * acc[(c1 * NB + c2) * 4 + q];
        x[q] = (y - y == 0.0) ? y : 0.0;                            // non-finite -> 0: the pool is recycled and its cells are
      }                                                             // read (times zero) by later batches
#ifdef __CUDA_ARCH__
      // one 256-bit store per record (a whole 32-byte sector): the stores of a warp go to 32 different sectors, four
      // 64-bit stores per record quadrupled the requests the memory pipeline had to take
      asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(v), "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]) : "memory");
#else
      v[0] = x[0]; v[1] = x[1]; v[2] = x[2]; v[3] = x[3];
#endif
    }
  }
}

// Host-side restatement of the record builder (the plan builds the same table in NumPy): records of the primitive
// pairs of two sub-shells, thresholds descending.  rec must hold npa * npb * PP_REC doubles.
inline void primitive_pair_records(int npa, const double *ea, const double *ca, int npb, const double *eb, const double *cb, double *rec) {
  int n = 0;
  for (int p = 0; p < npa; ++p)
    for (int q = 0; q < npb; ++q, ++n) {
      const double a = ea[p], b = eb[q], ip = 1.0 / (a + b), pi = 3.14159265358979323846;
      double *r = rec + n * PP_REC;
      r[0] = 46.0 * (a + b) / (a * b); r[1] = a * b * ip; r[2] = ca[p] * cb[q] * (pi * ip) * sqrt(pi * ip); r[3] = ip; r[4] = a; r[5] = b;
    }
  for (int i = 1; i < n; ++i)                                       // insertion sort, thr descending (stable)
    for (int j = i; j > 0 && rec[(j - 1) * PP_REC] < rec[j * PP_REC]; --j)
      for (int d = 0; d < PP_REC; ++d) { const double t = rec[j * PP_REC + d]; rec[j * PP_REC + d] = rec[(j - 1) * PP_REC + d]; rec[(j - 1) * PP_REC + d] = t; }
}

// One work item of k_rep phase 1 in a single call (the integral kernel takes the records from the plan's table and
// hoists everything that does not depend on the geometry out of its loop over pairs; this is the same sequence, used
// by the host check): sub-shells sa = {atom, first function, L, nprim} of the solute and sb of the solvent fragment,
// eX/cX the per-function [nbf][6] exponent and coefficient tables (coefficient x primitive norm x contracted norm),
// LcA the per-atom displacement vectors.
template <int LA, int LB, int IA>
inline void subshell_pair_store(const int *sa, const int *sb, const double *posA, const double *posB,
                                const double *eA, const double *cA, const double *eB, const double *cB,
                                const double *LcA, int LP, double *V) {
  constexpr int NA = IA < 0 ? cart_ncomp(LA) : 1, NB = cart_ncomp(LB);
  const int f0a = sa[1], f0b = sb[1];
  double acc[NA * NB * 4], rec[36 * PP_REC];
  primitive_pair_records(sa[3], eA + f0a * 6, cA + f0a * 6, sb[3], eB + f0b * 6, cB + f0b * 6, rec);
  subshell_pair_t<LA, LB, IA>(posA + sa[0] * 3, posB + sb[0] * 3, LcA + sa[0] * 3, sa[3] * sb[3], rec, nullptr, 1, acc);
  subshell_pair_write<LA, LB, IA>(acc, f0a, f0b, LP, V);
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
