// k_elect.cuh -- kernel A: one CTA per MD frame.
//   * MOLLST cut-off test per solvent instance (solvshift/src/solpol2.f:3-115)
//   * Kabsch superimposition of the solute and of every solvent instance inside any cut-off
//     (Frag.sup, solvshift/slvpar.py:641-679) -- once per instance, not once per layer
//   * rotation of the instance's distributed multipoles into the lab frame (ROTDMA, efprot.f:286-741)
//   * solute-solvent multipole electrostatic shift through R^-4 with the mode-contracted SolCAMM
//     moment sets (libbbg sdmtpm/sdmtpe, call sites solvshift/solefp.py:1175-1182) and the
//     charge-only potential-gradient corrections (dmtcor rf2/rk2, solefp.py:1186-1190)
//   * compacted hand-over list (instance, flags, rot, transl) for the pol/disp/rep kernels
// Coordinates are read exactly once from HBM; rotated multipoles are never materialised.
#pragma once
#include "slv_common.cuh"

namespace slv {

constexpr int ELECT_THREADS = 256;
constexpr int SOLROW = 52;   // doubles per solute site in shared memory: pos3, smea20, sea20, cmea3, cea3 (+3 pad)
// dynamic shared memory: solute rows, then per pass the transforms, unit offsets and types of blockDim.x instances
inline size_t elect_smem(int ndmac) {
  return ((size_t)ndmac * SOLROW + (size_t)ELECT_THREADS * 12) * sizeof(double) + (2 * (size_t)ELECT_THREADS + 2) * sizeof(int);
}

__global__ void __launch_bounds__(ELECT_THREADS, 2)
k_frame_elect(const DevPlan p, const double *__restrict__ coords /* first frame of this launch */, int frame0 /* its row in out */,
              FrameLists fl, double *__restrict__ out) {
  extern __shared__ double sm[];
  double *s_sol = sm;                                   // [ndma_c][SOLROW]
  __shared__ double s_xf[16];                           // solute rot(9) transl(3) rms, centroid(3)
  __shared__ double s_red[32];
  __shared__ int s_cnt[ELECT_THREADS / 32 + 1], s_usum[ELECT_THREADS / 32 + 1];
  __shared__ int s_base;

  const int fi = blockIdx.x;                            // frame within chunk
  const double *x = coords + (size_t)fi * p.natoms_total * 3;
  const int *tm0 = tmeta(p, p.inst_type[0]);
  const int ndmac = tm0[SLVGPU_M_NDMA];
  const int a0c = p.inst_atom0[0], natc_md = p.inst_atom0[1] - a0c;

  if (threadIdx.x == 0) {
    double rot[9], tr[3], rms;
    fit_instance(p, tm0, x + a0c * 3, rot, tr, rms);
    double c[3] = {0, 0, 0};
    for (int k = 0; k < natc_md; ++k) { c[0] += x[(a0c + k) * 3]; c[1] += x[(a0c + k) * 3 + 1]; c[2] += x[(a0c + k) * 3 + 2]; }
    const double inv = 1.0 / (double)natc_md;
#pragma unroll
    for (int d = 0; d < 9; ++d) s_xf[d] = rot[d];
#pragma unroll
    for (int d = 0; d < 3; ++d) { s_xf[9 + d] = tr[d]; s_xf[13 + d] = c[d] * inv; }
    s_xf[12] = rms;
    double *sx = fl.sol_xf + (size_t)fi * 12;
#pragma unroll
    for (int d = 0; d < 9; ++d) sx[d] = rot[d];
#pragma unroll
    for (int d = 0; d < 3; ++d) sx[9 + d] = tr[d];
    s_base = 0;
  }
  __syncthreads();

  const bool global = (p.flags & SLVGPU_GLOBAL) != 0;     // energy mode: this kernel only superimposes and hands every instance over
  // solute site rows in the lab frame
  if ((p.flags & SLVGPU_ELECT) && !global && threadIdx.x < ndmac) {
    const int s = threadIdx.x;
    double R[9];
#pragma unroll
    for (int d = 0; d < 9; ++d) R[d] = s_xf[d];
    double *row = s_sol + s * SOLROW;
    double y[3];
    rot_vec(R, p.dtab + tm0[SLVGPU_M_O_RDMA] + s * 3, y);
    row[0] = y[0] + s_xf[9]; row[1] = y[1] + s_xf[10]; row[2] = y[2] + s_xf[11];
    double m[20];
    rot_mom20(R, p.dtab + p.sol_meta[SLVGPU_S_SMEA] + s * 20, m); fold20(m);
#pragma unroll
    for (int k = 0; k < 20; ++k) row[3 + k] = m[k];
    rot_mom20(R, p.dtab + p.sol_meta[SLVGPU_S_SEA] + s * 20, m); fold20(m);
#pragma unroll
    for (int k = 0; k < 20; ++k) row[23 + k] = m[k];
    rot_vec(R, p.dtab + p.sol_meta[SLVGPU_S_CMEA] + s * 3, y);
    row[43] = y[0]; row[44] = y[1]; row[45] = y[2];
    rot_vec(R, p.dtab + p.sol_meta[SLVGPU_S_CEA] + s * 3, y);
    row[46] = y[0]; row[47] = y[1]; row[48] = y[2];
  }
  __syncthreads();

  const double cx = s_xf[13], cy = s_xf[14], cz = s_xf[15];
  double e_mea = 0.0, e_ea = 0.0, c_mea = 0.0, c_ea = 0.0, rms_sum = 0.0, rms_max = 0.0;
  double v_mea = 0.0, v_ea = 0.0, w_mea = 0.0, w_ea = 0.0;          // the same four sums over the non-EFP environment
  int n_c = 0, n_p = 0, n_e = 0;
  const int nsolv = p.ninst - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // Two steps per pass of blockDim.x instances.  (1) One thread per instance: cut-off test, superimposition, hand-over
  // list.  (2) The multipole sites of the pass's instances inside ccut form one flat list of (instance, site) units
  // that all threads share evenly: fragment types with different site counts, instances outside the cut-off and a
  // partly filled last pass would otherwise leave most lanes of a warp idle in the pair loop.
  double *s_inst = s_sol + (size_t)ndmac * SOLROW;        // [blockDim.x][12]  rot, transl of the pass's instances
  int *s_uoff = reinterpret_cast<int *>(s_inst + (size_t)blockDim.x * 12);   // [blockDim.x + 1] first unit of each instance
  int *s_ityp = s_uoff + blockDim.x + 1;                  // [blockDim.x] fragment type
  const bool do_el = (p.flags & SLVGPU_ELECT) && !global;
  for (int base = 0; base < nsolv; base += blockDim.x) {
    const int i = base + threadIdx.x + 1;                // instance index
    int flag = 0, nd = 0;
    double rot[9], tr[3], rms = 0.0;
    if (i < p.ninst) {
      const int a0 = p.inst_atom0[i], nat = p.inst_atom0[i + 1] - a0;
      const double *xi = x + (size_t)a0 * 3;
      double mx = 0, my = 0, mz = 0;
      for (int k = 0; k < nat; ++k) { mx += xi[k * 3]; my += xi[k * 3 + 1]; mz += xi[k * 3 + 2]; }
      const double dn = (double)nat;
      mx /= dn; my /= dn; mz /= dn;
      const double dx = cx - mx, dy = cy - my, dz = cz - mz;
      const double rr = sqrt(dx * dx + dy * dy + dz * dz);
      flag = (rr < p.ccut ? 4 : 0) | (rr < p.pcut ? 1 : 0) | (rr < p.ecut ? 2 : 0);
      const int ty = p.inst_type[i];
      const int *tm = tmeta(p, ty);
      const bool env = tm[SLVGPU_M_REMOVABLE] == 2;       // environment of eval_dma: no cut-off, electrostatics only
      if (env) flag = 4;
      if (global) flag = 5;                               // no cut-offs: every fragment is in the electrostatic and polarisation layers
      if (flag) {
        fit_instance(p, tm, xi, rot, tr, rms);
        if (flag & 4) {
          if (!env) { ++n_c; rms_sum += rms; rms_max = fmax(rms_max, rms); }
          if (do_el) {
            nd = tm[SLVGPU_M_NDMA];
            double *si = s_inst + (size_t)threadIdx.x * 12;
#pragma unroll
            for (int d = 0; d < 9; ++d) si[d] = rot[d];
#pragma unroll
            for (int d = 0; d < 3; ++d) si[9 + d] = tr[d];
            s_ityp[threadIdx.x] = ty;
          }
        }
        if (flag & 1) ++n_p;
        if (flag & 2) ++n_e;
      }
    }
    // ordered compaction of the instances the later kernels need (inside pcut or ecut), and the exclusive scan of the
    // site counts (both through one pass of warp scans)
    const int keep = (flag & 3) != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    int incl = nd;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 0) s_cnt[warp] = __popc(bal);
    if (lane == 31) s_usum[warp] = incl;
    __syncthreads();
    int off = s_base, ubase = 0;
    for (int w = 0; w < warp; ++w) { off += s_cnt[w]; ubase += s_usum[w]; }
    off += __popc(bal & ((1u << lane) - 1u));
    s_uoff[threadIdx.x] = ubase + incl - nd;
    if (threadIdx.x == blockDim.x - 1) s_uoff[blockDim.x] = ubase + incl;
    if (keep && off < p.list_cap) {
      const size_t e = (size_t)fi * p.list_cap + off;
      fl.list_inst[e] = i; fl.list_flag[e] = flag & 3;
      double *xf = fl.list_xf + e * 12;
#pragma unroll
      for (int d = 0; d < 9; ++d) xf[d] = rot[d];
#pragma unroll
      for (int d = 0; d < 3; ++d) xf[9 + d] = tr[d];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_cnt[w];
      s_base += tot;
    }
    if (do_el) {
      const int nunit = s_uoff[blockDim.x];
      for (int u = threadIdx.x; u < nunit; u += blockDim.x) {
        int lo = 0, hi = blockDim.x;                     // owner: largest t with s_uoff[t] <= u (instances without sites share their successor's offset)
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_uoff[mid] <= u) lo = mid; else hi = mid; }
        const int sidx = u - s_uoff[lo];
        const int *tm = tmeta(p, s_ityp[lo]);
        const bool env = tm[SLVGPU_M_REMOVABLE] == 2;
        const double *si = s_inst + (size_t)lo * 12;
        double R[9];
#pragma unroll
        for (int d = 0; d < 9; ++d) R[d] = si[d];
        double rB[3], mB[20];
        rot_vec(R, p.dtab + tm[SLVGPU_M_O_RDMA] + sidx * 3, rB);
        rB[0] += si[9]; rB[1] += si[10]; rB[2] += si[11];
        rot_mom20(R, p.dtab + tm[SLVGPU_M_O_MOM] + sidx * 20, mB);
        double e1s = 0.0, e2s = 0.0, c1s = 0.0, c2s = 0.0;
        for (int a = 0; a < ndmac; ++a) {
          const double *row = s_sol + a * SOLROW;
          double P[23];
          pair_potentials(row, rB, mB, P);
          double e1 = 0.0, e2 = 0.0;
#pragma unroll
          for (int k = 0; k < 20; ++k) { e1 += row[3 + k] * P[k]; e2 += row[23 + k] * P[k]; }
          e1s += e1; e2s += e2;
          c1s += row[43] * P[20] + row[44] * P[21] + row[45] * P[22];
          c2s += row[46] * P[20] + row[47] * P[21] + row[48] * P[22];
        }
        if (env) { v_mea += e1s; v_ea += e2s; w_mea += c1s; w_ea += c2s; }
        else { e_mea += e1s; e_ea += e2s; c_mea += c1s; c_ea += c2s; }
      }
    }
    __syncthreads();
  }

  e_mea = block_sum(e_mea, s_red); e_ea = block_sum(e_ea, s_red);
  c_mea = block_sum(c_mea, s_red); c_ea = block_sum(c_ea, s_red);
  rms_sum = block_sum(rms_sum, s_red); rms_max = block_max(rms_max, s_red);
  const double dn_c = block_sum((double)n_c, s_red), dn_p = block_sum((double)n_p, s_red), dn_e = block_sum((double)n_e, s_red);
  v_mea = block_sum(v_mea, s_red); v_ea = block_sum(v_ea, s_red); w_mea = block_sum(w_mea, s_red); w_ea = block_sum(w_ea, s_red);
  if (threadIdx.x == 0) {
    double *o = out + (size_t)(frame0 + fi) * SLVGPU_NOUT;
    const bool el = (p.flags & SLVGPU_ELECT) != 0;
    o[SLVGPU_ELE_MEA] = el ? e_mea : 0.0;
    o[SLVGPU_ELE_EA] = (el && (p.flags & SLVGPU_EA)) ? e_ea : 0.0;
    o[SLVGPU_COR_MEA] = (el && (p.flags & SLVGPU_CORR)) ? c_mea : 0.0;
    o[SLVGPU_COR_EA] = (el && (p.flags & SLVGPU_CORR)) ? c_ea : 0.0;
    o[SLVGPU_POL_MEA] = 0.0; o[SLVGPU_REP_MEA] = 0.0; o[SLVGPU_DIS_MEA] = 0.0; o[SLVGPU_DIS_ISO] = 0.0;
    o[SLVGPU_RMS_C] = s_xf[12];
    o[SLVGPU_RMS_SMAX] = rms_max;
    o[SLVGPU_RMS_AVE] = (s_xf[12] + rms_sum) / (dn_c + 1.0);
    o[SLVGPU_N_ELECT] = dn_c; o[SLVGPU_N_POL] = dn_p; o[SLVGPU_N_REP] = dn_e;
    o[SLVGPU_POL_ITERS] = 0.0; o[SLVGPU_POL_RESID] = 0.0;
    for (int k = SLVGPU_POL_ADD; k < SLVGPU_NOUT; ++k) o[k] = 0.0;
    o[SLVGPU_ENV_ELE_MEA] = el ? v_mea : 0.0;
    o[SLVGPU_ENV_ELE_EA] = (el && (p.flags & SLVGPU_EA)) ? v_ea : 0.0;
    o[SLVGPU_ENV_COR_MEA] = (el && (p.flags & SLVGPU_CORR)) ? w_mea : 0.0;
    o[SLVGPU_ENV_COR_EA] = (el && (p.flags & SLVGPU_CORR)) ? w_ea : 0.0;
    fl.nlist[fi] = s_base < p.list_cap ? s_base : p.list_cap;
  }
}


// Energy mode (EFP(freq=False), solvshift/solefp.py:905-1009): electrostatic interaction energy of the whole cluster,
//   E = sum over site pairs (a, b) of DIFFERENT fragments of the multipole series through R^-4
// (libbbg.qm.clemtp.edmtpa is un-vendored: the same truncation as the shift path, slv::pair_potentials, is used -- DESIGN.md).
// Persistent CTAs, one frame at a time; the lab-frame site table [nsite][24] (pos3, mom20, fragment) lives in a per-CTA
// workspace; rows in registers, columns staged through shared memory, each unordered pair evaluated once.
constexpr int GLOB_THREADS = 256;
__global__ void __launch_bounds__(GLOB_THREADS, 2)
k_global_elect(const DevPlan p, int frame0, int nframes, FrameLists fl, double *__restrict__ site_ws, double *__restrict__ out) {
  constexpr int TJ = 128;                          // column sites per shared-memory tile
  __shared__ double s_tile[TJ * 24];
  __shared__ double s_red[32];
  __shared__ int s_n;
  double *ws = site_ws + (size_t)blockIdx.x * p.site_cap * 24;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int *tm0 = tmeta(p, p.inst_type[0]);
  for (int fi = blockIdx.x; fi < nframes; fi += gridDim.x) {
    const int nl = fl.nlist[fi];
    const size_t lbase = (size_t)fi * p.list_cap;
    __syncthreads();
    if (tid == 0) s_n = tm0[SLVGPU_M_NDMA];
    __syncthreads();
    // site table: fragment 0 first, then the list entries in order (offsets by a serial scan in chunks of NT entries)
    for (int e0 = -1; e0 < nl; e0 += NT) {
      const int e = e0 + tid;
      const bool on = e < nl;
      const int *tm = e < 0 ? tm0 : (on ? tmeta(p, p.inst_type[fl.list_inst[lbase + e]]) : tm0);
      const int nd = on ? tm[SLVGPU_M_NDMA] : 0;
      // exclusive scan of nd over the chunk
      __shared__ int s_cnt[GLOB_THREADS];
      s_cnt[tid] = (e < 0) ? 0 : nd;
      __syncthreads();
      int offs = (e < 0) ? 0 : s_n;
      if (e >= 0) for (int k = 0; k < tid; ++k) offs += s_cnt[k];
      if (on) {
        double R[9], t[3];
        const double *xf = e < 0 ? fl.sol_xf + (size_t)fi * 12 : fl.list_xf + (lbase + e) * 12;
        for (int d = 0; d < 9; ++d) R[d] = xf[d];
        for (int d = 0; d < 3; ++d) t[d] = xf[9 + d];
        for (int sidx = 0; sidx < nd; ++sidx) {
          if (offs + sidx >= p.site_cap) break;
          double y[3], m[20];
          rot_vec(R, p.dtab + tm[SLVGPU_M_O_RDMA] + sidx * 3, y);
          rot_mom20(R, p.dtab + tm[SLVGPU_M_O_MOM] + sidx * 20, m);
          double *row = ws + (size_t)(offs + sidx) * 24;
          for (int d = 0; d < 3; ++d) row[d] = y[d] + t[d];
          for (int d = 0; d < 20; ++d) row[3 + d] = m[d];
          row[23] = (double)(e + 1);
        }
      }
      __syncthreads();
      if (tid == 0 && e0 >= 0) { int tot = 0; for (int k = 0; k < NT; ++k) tot += s_cnt[k]; s_n += tot; }
      if (e0 < 0) e0 = -NT;              // the first pass handled the solute alone (e = -1); entries start at 0
      __syncthreads();
    }
    const int nsite = min(s_n, p.site_cap);
    double acc = 0.0;
    for (int i0 = 0; i0 < nsite; i0 += NT) {
      const int i = i0 + tid;
      const bool act = i < nsite;
      double ri[3] = {0, 0, 0}, mi[20]; int moli = -1;
      if (act) {
        const double *row = ws + (size_t)i * 24;
        for (int d = 0; d < 3; ++d) ri[d] = row[d];
        for (int d = 0; d < 20; ++d) mi[d] = row[3 + d];
        fold20(mi);
        moli = (int)row[23];
      }
      for (int j0 = i0; j0 < nsite; j0 += TJ) {           // columns j > i only
        const int jn = min(TJ, nsite - j0);
        __syncthreads();
        for (int w = tid; w < jn * 24; w += NT) s_tile[w] = ws[(size_t)j0 * 24 + w];
        __syncthreads();
        if (act) {
          for (int jj = 0; jj < jn; ++jj) {
            const double *rj = s_tile + jj * 24;
            if (j0 + jj <= i || (int)rj[23] == moli) continue;
            double P[23];
            pair_potentials(ri, rj, rj + 3, P);
            double e1 = 0.0;
#pragma unroll
            for (int k = 0; k < 20; ++k) e1 += mi[k] * P[k];
            acc += e1;
          }
        }
      }
    }
    acc = block_sum(acc, s_red);
    if (tid == 0) out[(size_t)(frame0 + fi) * SLVGPU_NOUT + SLVGPU_ELE_MEA] = acc;
  }
}

}  // namespace slv
