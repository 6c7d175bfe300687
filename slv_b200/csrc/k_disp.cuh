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

constexpr int DISP_THREADS = 64;
constexpr int DISP_SOLROW = 6 + 216 + 24;   // pos3, u3, ai[12][9], A1[12][9], abar_i[12], A1bar_i[12]

__global__ void __launch_bounds__(DISP_THREADS)
k_disp(const DevPlan p, int frame0, FrameLists fl, double *__restrict__ out) {
  extern __shared__ double sm[];
  const int fi = blockIdx.x;
  const int *tm0 = tmeta(p, p.inst_type[0]);
  const int npolc = tm0[SLVGPU_M_NPOL];
  double *s_sol = sm;                                   // [npolc][DISP_SOLROW]
  double *s_aj = sm + (size_t)npolc * DISP_SOLROW;      // [108][DISP_THREADS]
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
  __syncthreads();

  const int nl = fl.nlist[fi];
  const int items = nl * p.max_npol;
  double acc_ani = 0.0, acc_iso = 0.0;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int e = it / p.max_npol, l = it % p.max_npol;
    const size_t le = (size_t)fi * p.list_cap + e;
    if (!(fl.list_flag[le] & 1)) continue;
    const int *tm = tmeta(p, p.inst_type[fl.list_inst[le]]);
    if (l >= tm[SLVGPU_M_NPOL]) continue;
    double R[9], rj[3];
    const double *xf = fl.list_xf + le * 12;
#pragma unroll
    for (int d = 0; d < 9; ++d) R[d] = xf[d];
    rot_vec(R, p.dtab + tm[SLVGPU_M_O_RPOL] + l * 3, rj);
    rj[0] += xf[9]; rj[1] += xf[10]; rj[2] += xf[11];
    double abj[12];
    for (int k = 0; k < 12; ++k) {
      double a[9];
      rot_mat(R, p.dtab + tm[SLVGPU_M_O_DPOLI] + ((size_t)l * 12 + k) * 9, a);
#pragma unroll
      for (int c = 0; c < 9; ++c) s_aj[(k * 9 + c) * DISP_THREADS + threadIdx.x] = a[c];
      abj[k] = (a[0] + a[4] + a[8]) * (1.0 / 3.0);
    }
    for (int i = 0; i < npolc; ++i) {
      const double *row = s_sol + (size_t)i * DISP_SOLROW;
      const double Rx = row[0] - rj[0], Ry = row[1] - rj[1], Rz = row[2] - rj[2];
      const double ux = row[3], uy = row[4], uz = row[5];
      const double r2 = Rx * Rx + Ry * Ry + Rz * Rz;
      const double iv = rsqrt64(r2), iv2 = iv * iv, iv5 = iv2 * iv2 * iv;
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
        for (int c = 0; c < 9; ++c) aj[c] = s_aj[(k * 9 + c) * DISP_THREADS + threadIdx.x];
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
  acc_ani = block_sum(acc_ani, s_red);
  acc_iso = block_sum(acc_iso, s_red);
  if (threadIdx.x == 0) {
    double *o = out + (size_t)(frame0 + fi) * SLVGPU_NOUT;
    o[SLVGPU_DIS_MEA] = acc_ani * (0.5 / 3.14159265358979323846);
    o[SLVGPU_DIS_ISO] = acc_iso * (3.0 / 3.14159265358979323846);
  }
}

}  // namespace slv
