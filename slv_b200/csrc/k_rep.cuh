// k_rep.cuh -- kernel D: exchange-repulsion frequency shift.  Persistent CTAs with a dynamic frame queue;
// per frame the CTA walks the solvent fragments inside ecut.
//
// Reference: per solvent molecule B the Python layer builds both basis sets, calls PyQuante for the AO
// integrals S, T, dS/dA, dT/dA (solvshift/solefp.py:1497-1526) and then SHFTEX (solvshift/src/shftex.f):
//   CALCIJ (298-438): S_ij = C_A S C_B^T, T_ij, and per mode  S^m_ij = C_A (L^m_atom(k) . dS_kl/dA) C_B^T + C1^m_A S C_B^T
//   CALCFI (69-295) : fi_m = sum_ij TS1..4 + TA1..4  (EFP2 spherical-Gaussian-overlap form)
//   SHFTMA = -sum_m g_m fi_m .
// fi_m is linear in the mode quantities (S^m, T^m, fock1_m, lmoc1_m, lvec_m), so the plan's g-contracted tables
// turn the 30-mode loop into one evaluation.  Phases per (frame, B):
//   1  AO integrals for every (k,l) pair -> V[k][l][4] = S, T, L.dS, L.dT      (slv_ints.cuh)
//   2  half transform  X[k][q][j] = sum_l V[k][l][q] C_B[j][l]
//   3  LMO integrals   S_ij, T_ij, Sm_ij, Tm_ij
//   4  CALCFI per (i,j), block sum.
// AO->LMO coefficients are rotated into the lab frame per fragment (VECROT, efprot.f:22-151).
#pragma once
#include <string>
#include <vector>
#include "slv_common.cuh"
#include "slv_ints.cuh"

namespace slv {

constexpr int REP_THREADS = 256;

struct RepWork {            // per-CTA global workspace (L2-resident), strides in doubles
  double *CA;    // [2][nmosA][nbsA]  rotated vecl and g-contracted vecl1 of the solute
  double *CB;    // [max_nmos][max_nbas]
  double *V;     // [nbsA][max_nbas][4]
  double *X;     // [nbsA][4][max_nmos]
  double *IJ;    // [4][nmosA][max_nmos]
  size_t sCA, sCB, sV, sX, sIJ;
  int *queue;
};

inline int rep_alloc(RepWork &rw, std::vector<void *> &allocs, std::string &err, int ctas, int nmosc, int nbasc,
                     int max_nmos, int max_nbas) {
  rw.sCA = (size_t)2 * nmosc * nbasc; rw.sCB = (size_t)max_nmos * max_nbas; rw.sV = (size_t)nbasc * max_nbas * 4;
  rw.sX = (size_t)nbasc * 4 * max_nmos; rw.sIJ = (size_t)4 * nmosc * max_nmos;
  const size_t tot = rw.sCA + rw.sCB + rw.sV + rw.sX + rw.sIJ;
  double *base = nullptr;
  if (cudaMalloc(&base, tot * ctas * sizeof(double)) != cudaSuccess) { err = "cudaMalloc(rep workspace) failed"; return SLVGPU_ENOMEM; }
  allocs.push_back(base);
  rw.CA = base; rw.CB = rw.CA + rw.sCA * ctas; rw.V = rw.CB + rw.sCB * ctas; rw.X = rw.V + rw.sV * ctas; rw.IJ = rw.X + rw.sX * ctas;
  if (cudaMalloc(&rw.queue, sizeof(int)) != cudaSuccess) { err = "cudaMalloc failed"; return SLVGPU_ENOMEM; }
  allocs.push_back(rw.queue);
  return 0;
}

// shared memory: solute [posA natA*3][LcA natA*3][riA nmosA*3][ri1A nmosA*3] + solvent [posB max_nat*3][riB max_nmos*3]
inline size_t rep_smem_bytes(int nmosc, int max_nmos, int natc, int max_nat, int nbasc, int max_nbas) {
  // + basis tables of the solute (staged once) and of the current solvent fragment: per function 6 exponents,
  //   6 coefficients and 5 ints (atom, lx, ly, lz, nprim) -- the integral loop reads them from shared memory
  return sizeof(double) * (size_t)(6 * natc + 6 * nmosc + 3 * max_nat + 3 * max_nmos + 8 + 15 * (nbasc + max_nbas));
}

// phase 1 of k_rep for one angular-momentum class: pairs order[lo..hi) -> V[(k*nbsB + l)*4 + q]; basis tables in smem
struct RepPhase1 {
  const double *posA, *posB, *LcA; double *V; int nbsB; const int *order;
  const int *iA; const double *eA, *cA;
  const int *iB; const double *eB, *cB;
  template <int LA, int LB>
  __device__ __forceinline__ void run(int lo, int hi) const {
    for (int wi = lo + threadIdx.x; wi < hi; wi += blockDim.x) {
      const int w = order[wi];
      const int k = w / nbsB, l = w % nbsB;
      const int *ik = iA + k * 5, *il = iB + l * 5;
      const int ak = ik[0], al = il[0];
      const int la[3] = {ik[1], ik[2], ik[3]}, lb[3] = {il[1], il[2], il[3]};
      double o[4];
      ao_pair_t<LA, LB>(posA + ak * 3, la, ik[4], eA + k * 6, cA + k * 6, posB + al * 3, lb, il[4], eB + l * 6, cB + l * 6,
                        LcA + ak * 3, o);
      double *v = V + (size_t)w * 4;
      v[0] = o[0]; v[1] = o[1]; v[2] = o[2]; v[3] = o[3];
    }
  }
};

// rotate the AO->LMO rows of one fragment type: dst[i][k] for i < nrows, from src (body frame)
__device__ inline void rotate_vecl_rows(const DevPlan &p, const int *tm, const double *src, int nrows, const double R[9], double *dst) {
  const int nb = tm[SLVGPU_M_NBASIS];
  const int *bl = p.itab + tm[SLVGPU_M_IO_BFL];
  for (int w = threadIdx.x; w < nrows * nb; w += blockDim.x) {
    const int i = w / nb, k = w % nb;
    const int lx = bl[k * 3], ly = bl[k * 3 + 1], lz = bl[k * 3 + 2], Ls = lx + ly + lz;
    const double *c = src + (size_t)i * nb + k;
    double *d = dst + (size_t)i * nb + k;
    if (Ls == 0) d[0] = c[0];
    else if (Ls == 1 && lx == 1) { double v[3] = {c[0], c[1], c[2]}, o[3]; rot_vec(R, v, o); d[0] = o[0]; d[1] = o[1]; d[2] = o[2]; }
    else if (Ls == 2 && lx == 2) { double v[6] = {c[0], c[1], c[2], c[3], c[4], c[5]}, o[6]; rot_quad(R, v, o); for (int t = 0; t < 6; ++t) d[t] = o[t]; }
  }
}

__global__ void __launch_bounds__(REP_THREADS, 2)
k_rep(const DevPlan p, int frame0, int nframes, FrameLists fl, RepWork rw, double *__restrict__ out, int max_nbas) {
  extern __shared__ double sm[];
  __shared__ double s_red[32];
  __shared__ int s_fi;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int *tm0 = tmeta(p, p.inst_type[0]);
  const int natA = tm0[SLVGPU_M_NATOMS], nmosA = tm0[SLVGPU_M_NMOS], nbsA = tm0[SLVGPU_M_NBASIS];
  double *posA = sm, *LcA = posA + 3 * natA, *riA = LcA + 3 * natA, *ri1A = riA + 3 * nmosA;
  double *s_eA = ri1A + 3 * nmosA, *s_cA = s_eA + 6 * nbsA;                 // solute basis: exponents, coefficients
  int *s_iA = reinterpret_cast<int *>(s_cA + 6 * nbsA);                      // [nbsA][5] atom, lx, ly, lz, nprim
  double *s_eB = s_cA + 6 * nbsA + 3 * nbsA, *s_cB = s_eB + 6 * max_nbas;    // current solvent fragment's basis
  int *s_iB = reinterpret_cast<int *>(s_cB + 6 * max_nbas);
  double *solv = s_cB + 6 * max_nbas + 3 * max_nbas;
  double *CA = rw.CA + blockIdx.x * rw.sCA, *CA1 = CA + (size_t)nmosA * nbsA;
  double *CB = rw.CB + blockIdx.x * rw.sCB, *V = rw.V + blockIdx.x * rw.sV, *X = rw.X + blockIdx.x * rw.sX, *IJ = rw.IJ + blockIdx.x * rw.sIJ;
  const double *FA = p.dtab + tm0[SLVGPU_M_O_FOCK], *FA1 = p.dtab + p.sol_meta[SLVGPU_S_FOCK1], *ZA = p.dtab + tm0[SLVGPU_M_O_ZNUC];
  const int *blA = p.itab + tm0[SLVGPU_M_IO_BFL];
  const double *beA = p.dtab + tm0[SLVGPU_M_O_BFEXP], *bfA = p.dtab + tm0[SLVGPU_M_O_BFCOEF];

  {
    const int *bcA = p.itab + tm0[SLVGPU_M_IO_BFC], *bnA = p.itab + tm0[SLVGPU_M_IO_BFNP];
    for (int w = tid; w < nbsA * 6; w += NT) { s_eA[w] = beA[w]; s_cA[w] = bfA[w]; }
    for (int k = tid; k < nbsA; k += NT) {
      s_iA[k * 5] = bcA[k]; s_iA[k * 5 + 1] = blA[k * 3]; s_iA[k * 5 + 2] = blA[k * 3 + 1]; s_iA[k * 5 + 3] = blA[k * 3 + 2]; s_iA[k * 5 + 4] = bnA[k];
    }
  }
  for (;;) {
    __syncthreads();
    if (tid == 0) s_fi = atomicAdd(rw.queue, 1);
    __syncthreads();
    const int fi = s_fi;
    if (fi >= nframes) break;
    const int nl = fl.nlist[fi];
    const size_t lbase = (size_t)fi * p.list_cap;
    double R0[9], t0[3];
    {
      const double *sx = fl.sol_xf + (size_t)fi * 12;
      for (int d = 0; d < 9; ++d) R0[d] = sx[d];
      for (int d = 0; d < 3; ++d) t0[d] = sx[9 + d];
    }
    for (int a = tid; a < natA; a += NT) {
      double y[3];
      rot_vec(R0, p.dtab + tm0[SLVGPU_M_O_POS] + a * 3, y);
      posA[a * 3] = y[0] + t0[0]; posA[a * 3 + 1] = y[1] + t0[1]; posA[a * 3 + 2] = y[2] + t0[2];
      rot_vec(R0, p.dtab + p.sol_meta[SLVGPU_S_LVEC] + a * 3, LcA + a * 3);
    }
    for (int i = tid; i < nmosA; i += NT) {
      double y[3];
      rot_vec(R0, p.dtab + tm0[SLVGPU_M_O_LMOC] + i * 3, y);
      riA[i * 3] = y[0] + t0[0]; riA[i * 3 + 1] = y[1] + t0[1]; riA[i * 3 + 2] = y[2] + t0[2];
      rot_vec(R0, p.dtab + p.sol_meta[SLVGPU_S_LMOC1] + i * 3, ri1A + i * 3);
    }
    rotate_vecl_rows(p, tm0, p.dtab + tm0[SLVGPU_M_O_VECL], nmosA, R0, CA);
    rotate_vecl_rows(p, tm0, p.dtab + p.sol_meta[SLVGPU_S_VECL1], nmosA, R0, CA1);
    __syncthreads();

    double acc = 0.0;
    for (int e = 0; e < nl; ++e) {
      if (!(fl.list_flag[lbase + e] & 2)) continue;          // uniform across the CTA
      const int *tm = tmeta(p, p.inst_type[fl.list_inst[lbase + e]]);
      const int natB = tm[SLVGPU_M_NATOMS], nmosB = tm[SLVGPU_M_NMOS], nbsB = tm[SLVGPU_M_NBASIS];
      double *posB = solv, *riB = solv + 3 * natB;
      double R[9], t[3];
      const double *xf = fl.list_xf + (lbase + e) * 12;
      for (int d = 0; d < 9; ++d) R[d] = xf[d];
      for (int d = 0; d < 3; ++d) t[d] = xf[9 + d];
      __syncthreads();                                       // previous fragment's phase 4 is done with solv/IJ
      for (int a = tid; a < natB; a += NT) {
        double y[3];
        rot_vec(R, p.dtab + tm[SLVGPU_M_O_POS] + a * 3, y);
        posB[a * 3] = y[0] + t[0]; posB[a * 3 + 1] = y[1] + t[1]; posB[a * 3 + 2] = y[2] + t[2];
      }
      for (int j = tid; j < nmosB; j += NT) {
        double y[3];
        rot_vec(R, p.dtab + tm[SLVGPU_M_O_LMOC] + j * 3, y);
        riB[j * 3] = y[0] + t[0]; riB[j * 3 + 1] = y[1] + t[1]; riB[j * 3 + 2] = y[2] + t[2];
      }
      rotate_vecl_rows(p, tm, p.dtab + tm[SLVGPU_M_O_VECL], nmosB, R, CB);
      {
        const int *bcB = p.itab + tm[SLVGPU_M_IO_BFC], *blB = p.itab + tm[SLVGPU_M_IO_BFL], *bnB = p.itab + tm[SLVGPU_M_IO_BFNP];
        const double *beB = p.dtab + tm[SLVGPU_M_O_BFEXP], *bfB = p.dtab + tm[SLVGPU_M_O_BFCOEF];
        for (int w = tid; w < nbsB * 6; w += NT) { s_eB[w] = beB[w]; s_cB[w] = bfB[w]; }
        for (int l = tid; l < nbsB; l += NT) {
          s_iB[l * 5] = bcB[l]; s_iB[l * 5 + 1] = blB[l * 3]; s_iB[l * 5 + 2] = blB[l * 3 + 1]; s_iB[l * 5 + 3] = blB[l * 3 + 2]; s_iB[l * 5 + 4] = bnB[l];
        }
      }
      __syncthreads();
      // ---- phase 1: AO integrals, class by class (angular momentum of bra x ket) in primitive-count order
      {
        const int *po = p.itab + tm[SLVGPU_M_IO_PAIRORDER];     // [10] class offsets, then the pair list
        RepPhase1 ph{posA, posB, LcA, V, nbsB, po + 10, s_iA, s_eA, s_cA, s_iB, s_eB, s_cB};
        ph.run<0, 0>(po[0], po[1]); ph.run<0, 1>(po[1], po[2]); ph.run<0, 2>(po[2], po[3]);
        ph.run<1, 0>(po[3], po[4]); ph.run<1, 1>(po[4], po[5]); ph.run<1, 2>(po[5], po[6]);
        ph.run<2, 0>(po[6], po[7]); ph.run<2, 1>(po[7], po[8]); ph.run<2, 2>(po[8], po[9]);
      }
      __syncthreads();
      // ---- phase 2: X[k][q][j] = sum_l V[k][l][q] CB[j][l]
      for (int w = tid; w < nbsA * 4 * nmosB; w += NT) {
        const int j = w % nmosB, q = (w / nmosB) & 3, k = w / (4 * nmosB);
        const double *v = V + (size_t)k * nbsB * 4 + q, *cb = CB + (size_t)j * nbsB;
        double s = 0.0;
        for (int l = 0; l < nbsB; ++l) s += v[(size_t)l * 4] * cb[l];
        X[w] = s;
      }
      __syncthreads();
      // ---- phase 3: LMO-basis integrals
      const int nij = nmosA * nmosB;
      for (int w = tid; w < nij; w += NT) {
        const int i = w / nmosB, j = w % nmosB;
        const double *ca = CA + (size_t)i * nbsA, *ca1 = CA1 + (size_t)i * nbsA;
        double s = 0, tt = 0, smv = 0, tmv = 0;
        for (int k = 0; k < nbsA; ++k) {
          const double *x = X + (size_t)k * 4 * nmosB + j;
          const double xs = x[0], xt = x[nmosB], xds = x[2 * nmosB], xdt = x[3 * nmosB];
          s += ca[k] * xs; tt += ca[k] * xt;
          smv += ca[k] * xds + ca1[k] * xs; tmv += ca[k] * xdt + ca1[k] * xt;
        }
        IJ[w] = s; IJ[nij + w] = tt; IJ[2 * nij + w] = smv; IJ[3 * nij + w] = tmv;
      }
      __syncthreads();
      // ---- phase 4: CALCFI
      {
        const double *Sij = IJ, *Tij = IJ + nij, *Smij = IJ + 2 * nij, *Tmij = IJ + 3 * nij;
        const double *FB = p.dtab + tm[SLVGPU_M_O_FOCK], *ZB = p.dtab + tm[SLVGPU_M_O_ZNUC];
        const double PI = 3.14159265358979323846;
        for (int w = tid; w < nij; w += NT) {
          const int i = w / nmosB, j = w % nmosB;
          const double S = Sij[w], T = Tij[w], Sm = Smij[w], Tm = Tmij[w];
          const double sl = log(fabs(S)), S2 = S * S;
          const double ri[3] = {riA[i * 3], riA[i * 3 + 1], riA[i * 3 + 2]}, r1i[3] = {ri1A[i * 3], ri1A[i * 3 + 1], ri1A[i * 3 + 2]};
          const double rj[3] = {riB[j * 3], riB[j * 3 + 1], riB[j * 3 + 2]};
          double dx = ri[0] - rj[0], dy = ri[1] - rj[1], dz = ri[2] - rj[2];
          double rinv = rsqrt64(dx * dx + dy * dy + dz * dz);
          const double rmij = rinv * (dx * r1i[0] + dy * r1i[1] + dz * r1i[2]);
          const double sdr = S * rinv;
          const double aux = 2.0 * sqrt(-2.0 * sl / PI);
          double f = 4.0 * sdr * Sm * (sqrt(-1.0 / (2.0 * PI * sl)) - aux - 1.0) + 2.0 * sdr * sdr * rmij * (aux + 1.0)
                   + 4.0 * Sm * T + 4.0 * Tm * S;
          for (int k = 0; k < nmosA; ++k) {                 // TA1
            const double skj = Sij[k * nmosB + j], smkj = Smij[k * nmosB + j];
            dx = riA[k * 3] - rj[0]; dy = riA[k * 3 + 1] - rj[1]; dz = riA[k * 3 + 2] - rj[2];
            rinv = rsqrt64(dx * dx + dy * dy + dz * dz);
            const double rmkj = rinv * (dx * ri1A[k * 3] + dy * ri1A[k * 3 + 1] + dz * ri1A[k * 3 + 2]);
            const double fik = FA[i * nmosA + k], f1ik = FA1[i * nmosA + k];
            f += -2.0 * Sm * fik * skj + 8.0 * S * rinv * Sm - 2.0 * S * (f1ik * skj + fik * smkj) - 4.0 * S2 * rinv * rinv * rmkj;
          }
          for (int l = 0; l < nmosB; ++l) {                 // TA2
            const double sil = Sij[i * nmosB + l], smil = Smij[i * nmosB + l];
            dx = ri[0] - riB[l * 3]; dy = ri[1] - riB[l * 3 + 1]; dz = ri[2] - riB[l * 3 + 2];
            rinv = rsqrt64(dx * dx + dy * dy + dz * dz);
            const double rmil = rinv * (dx * r1i[0] + dy * r1i[1] + dz * r1i[2]);
            const double fjl = FB[j * nmosB + l];
            f += -2.0 * Sm * fjl * sil + 8.0 * S * rinv * Sm - 2.0 * S * fjl * smil - 4.0 * S2 * rinv * rinv * rmil;
          }
          for (int m = 0; m < natA; ++m) {                  // TA3
            dx = posA[m * 3] - rj[0]; dy = posA[m * 3 + 1] - rj[1]; dz = posA[m * 3 + 2] - rj[2];
            rinv = rsqrt64(dx * dx + dy * dy + dz * dz);
            const double rmm = rinv * (dx * LcA[m * 3] + dy * LcA[m * 3 + 1] + dz * LcA[m * 3 + 2]);
            f += 2.0 * S2 * ZA[m] * rinv * rinv * rmm - 4.0 * S * Sm * ZA[m] * rinv;
          }
          for (int n = 0; n < natB; ++n) {                  // TA4
            dx = ri[0] - posB[n * 3]; dy = ri[1] - posB[n * 3 + 1]; dz = ri[2] - posB[n * 3 + 2];
            rinv = rsqrt64(dx * dx + dy * dy + dz * dz);
            const double rmn = rinv * (dx * r1i[0] + dy * r1i[1] + dz * r1i[2]);
            f += 2.0 * S2 * ZB[n] * rinv * rinv * rmn - 4.0 * S * Sm * ZB[n] * rinv;
          }
          acc += f;
        }
      }
    }
    acc = block_sum(acc, s_red);
    if (tid == 0) out[(size_t)(frame0 + fi) * SLVGPU_NOUT + SLVGPU_REP_MEA] = -acc;
  }
}

}  // namespace slv
