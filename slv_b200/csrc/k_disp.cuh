// k_disp.cuh -- kernel B: dispersion shift, one CTA per MD frame, one thread per solvent LMO.
// LMO-LMO C6 dispersion between the solute and every solvent fragment inside pcut with the 12-point
// Gauss-Legendre imaginary-frequency quadrature (libbbg sdisp6 / tdisp6; call site
// solvshift/solefp.py:1427-1435; quadrature solvshift/slvpar.py:989-1019):
//   E_ani = -(1/2pi) sum_ij sum_k W_k T_ab aj(k)_cb T_cd ai(k)_da ,  T = (3RR - R^2)/R^5
//   E_iso = -(3/pi)  sum_ij sum_k W_k abar_i(k) abar_j(k) / R^6
//   shift = -sum_m g_m dE/dQ_m,  dE/dQ_m acting on ai (dpoli1_m) and on r_i (lmoc1_m).
// The mode sum is linear, so the plan carries A1 = sum_m g_m dpoli1_m and u = sum_m g_m lmoc1_m and one
// pass over the pairs replaces the reference's (nmodes+1) passes.
#pragma once
#include "slv_common.cuh"

namespace slv {

constexpr int DISP_THREADS = 256;
constexpr int DISP_TJ = 64;                 // solvent LMOs per tile; 4 threads (solute-LMO quarters) share one solvent LMO
constexpr int DISP_SOLROW = 6 + 216 + 24;   // pos3, u3, ai[12][9], A1[12][9], abar_i[12], A1bar_i[12]

__global__ void __launch_bounds__(DISP_THREADS, 2)
k_disp(const DevPlan p, int frame0, FrameLists fl, double *__restrict__ out) {
  extern __shared__ double sm[];
  const int fi = blockIdx.x;
  const int *tm0 = tmeta(p, p.inst_type[0]);
  const int npolc = tm0[SLVGPU_M_NPOL];
  double *s_sol = sm;                                   // [npolc][DISP_SOLROW]
  double *s_aj = sm + (size_t)npolc * DISP_SOLROW;      // [108][DISP_TJ]  lab-frame alpha_j(i w_k) of the tile
  double *s_rj = s_aj + 108 * DISP_TJ;                  // [4][DISP_TJ]    position + valid flag
  __shared__ double s_red[32];

  // solute LMO rows in the lab frame
  {
    double R[9], t[3];
    const double *sx = fl.sol_xf + (size_t)fi * 12;
#pragma unroll
    for (int d = 0; d < 9; ++d) R[d] = sx[d];
#pragma unroll
    for (int d = 0; d < 3; ++d) t[d] = sx[9 + d];
    for (int w = threadIdx.x; w < npolc * 25; w += blockDim.x) {
      const int i = w / 25, part = w % 25;
      double *row = s_sol + (size_t)i * DISP_SOLROW;
      if (part == 24) {
        double y[3];
        rot_vec(R, p.dtab + tm0[SLVGPU_M_O_RPOL] + i * 3, y);
        row[0] = y[0] + t[0]; row[1] = y[1] + t[1]; row[2] = y[2] + t[2];
        rot_vec(R, p.dtab + p.sol_meta[SLVGPU_S_LMOC1] + i * 3, y);
        row[3] = y[0]; row[4] = y[1]; row[5] = y[2];
      } else {
        const int k = part % 12, which = part / 12;
        const double *src = p.dtab + (which ? p.sol_meta[SLVGPU_S_DPOLI1] : tm0[SLVGPU_M_O_DPOLI]) + ((size_t)i * 12 + k) * 9;
        double a[9];
        rot_mat(R, src, a);
        double *dst = row + 6 + which * 108 + k * 9;
#pragma unroll
        for (int c = 0; c < 9; ++c) dst[c] = a[c];
        row[6 + 216 + which * 12 + k] = (a[0] + a[4] + a[8]) * (1.0 / 3.0);
      }
    }
  }

  // compact index of the solvent LMOs inside pcut: exclusive scan of npol over the list entries
  const int nl = fl.nlist[fi];
  int *s_off = reinterpret_cast<int *>(s_rj + 4 * DISP_TJ);         // [nl + 1]
  __shared__ int s_scan[DISP_THREADS / 32];
  __shared__ int s_run;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  for (int e0 = 0; e0 < nl; e0 += blockDim.x) {
    const int e = e0 + threadIdx.x;
    int n = 0;
    if (e < nl && (fl.list_flag[(size_t)fi * p.list_cap + e] & 1)) n = tmeta(p, p.inst_type[fl.list_inst[(size_t)fi * p.list_cap + e]])[SLVGPU_M_NPOL];
    int incl = n;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_scan[w] = incl;
    __syncthreads();
    int base = s_run, tot = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { if (i < w) base += s_scan[i]; tot += s_scan[i]; }
    if (e < nl) s_off[e] = base + incl - n;
    __syncthreads();
    if (threadIdx.x == 0) s_run += tot;
    __syncthreads();
  }
  const int items = s_run;
  if (threadIdx.x == 0) s_off[nl] = items;
  __syncthreads();
  const int jl = threadIdx.x & (DISP_TJ - 1), ig = threadIdx.x / DISP_TJ;    // ig in 0..3
  double acc_ani = 0.0, acc_iso = 0.0;
  for (int it0 = 0; it0 < items; it0 += DISP_TJ) {
    __syncthreads();
    // ---- phase 1: the four threads of a solvent LMO rotate three frequencies each into shared memory
    {
      const int it = it0 + jl;
      const bool valid = it < items;
      if (valid) {
        int lo = 0, hi = nl;                     // largest e with s_off[e] <= it (entries with npol = 0 are skipped by the upper bound)
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= it) lo = mid; else hi = mid; }
        const int e = lo, l = it - s_off[lo];
        const size_t le = (size_t)fi * p.list_cap + e;
        const int *tm = tmeta(p, p.inst_type[fl.list_inst[le]]);
        double R[9];
        const double *xf = fl.list_xf + le * 12;
#pragma unroll
        for (int d = 0; d < 9; ++d) R[d] = xf[d];
        for (int k = ig * 3; k < ig * 3 + 3; ++k) {
          double a[9];
          rot_mat(R, p.dtab + tm[SLVGPU_M_O_DPOLI] + ((size_t)l * 12 + k) * 9, a);
#pragma unroll
          for (int c = 0; c < 9; ++c) s_aj[(k * 9 + c) * DISP_TJ + jl] = a[c];
        }
        if (ig == 0) {
          double rj[3];
          rot_vec(R, p.dtab + tm[SLVGPU_M_O_RPOL] + l * 3, rj);
          s_rj[0 * DISP_TJ + jl] = rj[0] + xf[9]; s_rj[1 * DISP_TJ + jl] = rj[1] + xf[10]; s_rj[2 * DISP_TJ + jl] = rj[2] + xf[11];
        }
      }
      if (ig == 0) s_rj[3 * DISP_TJ + jl] = valid ? 1.0 : 0.0;
    }
    __syncthreads();
    // ---- phase 2: each thread takes a quarter of the solute LMOs against its solvent LMO
    if (s_rj[3 * DISP_TJ + jl] != 0.0) {
      const double rj[3] = {s_rj[jl], s_rj[DISP_TJ + jl], s_rj[2 * DISP_TJ + jl]};
      double abj[12];
#pragma unroll
      for (int k = 0; k < 12; ++k)
        abj[k] = (s_aj[(k * 9 + 0) * DISP_TJ + jl] + s_aj[(k * 9 + 4) * DISP_TJ + jl] + s_aj[(k * 9 + 8) * DISP_TJ + jl]) * (1.0 / 3.0);
      for (int i = ig; i < npolc; i += DISP_THREADS / DISP_TJ) {
        const double *row = s_sol + (size_t)i * DISP_SOLROW;
        const double Rx = row[0] - rj[0], Ry = row[1] - rj[1], Rz = row[2] - rj[2];
        const double ux = row[3], uy = row[4], uz = row[5];
        const double r2 = Rx * Rx + Ry * Ry + Rz * Rz;
        const double iv = rsqrt_fast(r2), iv2 = iv * iv, iv5 = iv2 * iv2 * iv;
        const double Ru = Rx * ux + Ry * uy + Rz * uz;
        // T = (3RR - r2)/r^5 and its directional derivative along u (u moves r_i)
        double T[9], dT[9];
        const double Rv[3] = {Rx, Ry, Rz}, uv[3] = {ux, uy, uz};
        const double f5 = 5.0 * Ru * iv2;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            const double dl = (a == b) ? 1.0 : 0.0;
            const double tt = (3.0 * Rv[a] * Rv[b] - r2 * dl) * iv5;
            T[a * 3 + b] = tt;
            dT[a * 3 + b] = (3.0 * (uv[a] * Rv[b] + Rv[a] * uv[b]) - 2.0 * Ru * dl) * iv5 - f5 * tt;
          }
        double s_ani = 0.0, s1 = 0.0, s2 = 0.0;
        for (int k = 0; k < 12; ++k) {
          double aj[9];
#pragma unroll
          for (int c = 0; c < 9; ++c) aj[c] = s_aj[(k * 9 + c) * DISP_TJ + jl];
          const double *ai = row + 6 + k * 9, *A1 = row + 6 + 108 + k * 9;
          double X[9], Xp[9];
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              X[a * 3 + c] = T[a * 3 + 0] * aj[c * 3 + 0] + T[a * 3 + 1] * aj[c * 3 + 1] + T[a * 3 + 2] * aj[c * 3 + 2];
              Xp[a * 3 + c] = dT[a * 3 + 0] * aj[c * 3 + 0] + dT[a * 3 + 1] * aj[c * 3 + 1] + dT[a * 3 + 2] * aj[c * 3 + 2];
            }
          double tr = 0.0;
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const double Y = X[a * 3 + 0] * T[0 * 3 + d] + X[a * 3 + 1] * T[1 * 3 + d] + X[a * 3 + 2] * T[2 * 3 + d];
              const double Yd = Xp[a * 3 + 0] * T[0 * 3 + d] + Xp[a * 3 + 1] * T[1 * 3 + d] + Xp[a * 3 + 2] * T[2 * 3 + d]
                              + X[a * 3 + 0] * dT[0 * 3 + d] + X[a * 3 + 1] * dT[1 * 3 + d] + X[a * 3 + 2] * dT[2 * 3 + d];
              tr += Y * A1[d * 3 + a] + Yd * ai[d * 3 + a];
            }
          s_ani += p.glw[k] * tr;
          s1 += p.glw[k] * row[6 + 216 + 12 + k] * abj[k];
          s2 += p.glw[k] * row[6 + 216 + k] * abj[k];
        }
        acc_ani += s_ani;
        const double iv6 = iv2 * iv2 * iv2;
        acc_iso += s1 * iv6 - 6.0 * Ru * s2 * iv6 * iv2;
      }
    }
  }
  acc_ani = block_sum(acc_ani, s_red);
  acc_iso = block_sum(acc_iso, s_red);
  if (threadIdx.x == 0) {
    double *o = out + (size_t)(frame0 + fi) * SLVGPU_NOUT;
    o[SLVGPU_DIS_MEA] = acc_ani * (0.5 / 3.14159265358979323846);
    o[SLVGPU_DIS_ISO] = acc_iso * (3.0 / 3.14159265358979323846);
  }
}

}  // namespace slv
