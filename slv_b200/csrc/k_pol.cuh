// k_pol.cuh -- kernel C: polarisation (induced-dipole) frequency shift.  Persistent CTAs, one MD frame at
// a time per CTA, scratch tables in an L2-resident per-CTA workspace.
//
// Reference: SFTPLI + FFFFFF + AAAAAA (solvshift/src/solpol2.f:118-239, 1270-1511, 242-1150) build the dense
// matrix D (alpha^-1 blocks, dipole tensor T between centres of different molecules), invert it by LU and do
// two NDIM^3 DGEMMs per normal mode.  Algebraically, with x = D^-1 F and y = D^-T F,
//      SHIFT = -1/2 [ y^T dD x - (x + y) . dF ],   dD = sum_m g_m dD_m,  dF = sum_m g_m dF_m,
// where dD has entries only in the solute rows/columns and dF is a sum over (solute, solvent) pairs, both
// linear in the mode-derivative parameters -- which the plan pre-contracts.  So the work is: the fields F,
// two linear solves, and O(N) pair contractions.  The two solves are one matrix-free BiCG run on the
// block-preconditioned operator A = I + alpha T (BiCG solves A x = b and A^T xt = bt together, which is
// exactly the D / D^T pair needed); every sweep over the centre pairs applies T to both search directions,
// and T is rebuilt from the centre coordinates each sweep (no N^2 storage).
// The reference's quirks are reproduced: octupole field prefactor 5 in F (solpol2.f:1425), 1/2 on the
// solute-solvent dT blocks (699-706), the SUM2 slip in the field derivatives (466-483, 959-976).
#pragma once
#include "slv_common.cuh"

namespace slv {

#ifndef SLV_POL_THREADS
#define SLV_POL_THREADS 256
#endif
#ifndef SLV_POL_CTAS_PER_SM
#define SLV_POL_CTAS_PER_SM 3
#endif
constexpr int POL_THREADS = SLV_POL_THREADS;
constexpr int POL_CTAS_PER_SM = SLV_POL_CTAS_PER_SM;
#ifndef SLV_POL_TILE
#define SLV_POL_TILE 768
#endif
constexpr int POL_TILE = SLV_POL_TILE;   // centres the resident sweep tile holds: {position, u, ut, molecule} = 80 bytes each;
                                         // 768 centres = 60 KB: three CTAs share an SM and the shared-memory carve-out leaves L1 for the workspace traffic
#ifndef SLV_POL_ROWS
#define SLV_POL_ROWS 2
#endif
constexpr int POL_ROWS = SLV_POL_ROWS;   // rows per lane in the resident sweep (2: the tile entry of a column is read once for two rows)
constexpr int POL_BSH = POL_ROWS == 2 ? 6 : 5;   // log2 of the rows of a sweep block
constexpr int SITE_ROW = 24;     // pos3, mom20, molecule id
constexpr int PST = 32;          // BiCG state per centre: x xt r rt p pt u ut q qt (3 each) + 2 pad: rows are 32-byte aligned, the
                                 // vector phases move them with 256-bit loads / stores (a third of the memory requests)
constexpr int AST = 12;          // row stride of the centre polarisabilities (9 used)

struct PolWork {                 // per-CTA workspace (device pointers already offset per CTA by the kernel)
  double *cent_pos;   // [cent_cap][3]
  double *cent_alpha; // [cent_cap][9]
  double *fld;        // [cent_cap][3]
  double *xy;         // [cent_cap][PST] BiCG state + [cent_cap][6] final x, y
  double *site;       // [site_cap][SITE_ROW]
  double *aux;        // [2][cent_cap][3] right-hand sides of the AVEC solve (detail mode)
  double *part;       // [POL_THREADS/32][2][64][6] partial sums of the balanced sweep
  int *cent_mol;      // [cent_cap]
  int *ent_coff;      // [list_cap] first centre of a list entry
  int *ent_soff;      // [list_cap] first site of a list entry
};

// 256-bit accesses to the L2-resident workspace (32-byte aligned rows); volatile + memory clobber: the rows are shared
// between the phases of the solver through barriers
__device__ __forceinline__ void ld4(const double *p, double *o) {
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st4(double *p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// One ordered centre pair of the BiCG sweep: dipole tensor of R = rc - r_j applied to the two search directions
// held in the tile entry tt = {r_j(3), p_j(3), ut_j(3), molecule_j}; same-molecule pairs (incl. self) are masked branch-free.
__device__ __forceinline__ void sweep_pair(const double rc[3], const bool same, const double tt[9], double ax[3], double ay[3]) {
  const double Rx = rc[0] - tt[0], Ry = rc[1] - tt[1], Rz = rc[2] - tt[2];
  const double r2 = same ? 1.0 : (Rx * Rx + Ry * Ry + Rz * Rz);
  const double u = rsqrt_fast(r2), u2 = u * u;
  const double u3 = same ? 0.0 : u2 * u, t5 = 3.0 * u3 * u2;
  const double Rxj = fma(Rz, tt[5], fma(Ry, tt[4], Rx * tt[3]));
  const double Ryj = fma(Rz, tt[8], fma(Ry, tt[7], Rx * tt[6]));
  const double fx = -t5 * Rxj, fy = -t5 * Ryj;
  ax[0] = fma(fx, Rx, fma(u3, tt[3], ax[0])); ax[1] = fma(fx, Ry, fma(u3, tt[4], ax[1])); ax[2] = fma(fx, Rz, fma(u3, tt[5], ax[2]));
  ay[0] = fma(fy, Rx, fma(u3, tt[6], ay[0])); ay[1] = fma(fy, Ry, fma(u3, tt[7], ay[1])); ay[2] = fma(fy, Rz, fma(u3, tt[8], ay[2]));
}

__global__ void __launch_bounds__(POL_THREADS, POL_CTAS_PER_SM)
k_pol(const DevPlan p, int frame0, int nframes, FrameLists fl, PolWork wk0, double *__restrict__ out, int *__restrict__ status,
      double *__restrict__ detail /* optional [nframes][cent_cap][9]: dipind, flds, rpol */, int *__restrict__ queue) {
  __shared__ double s_red[64], s_redA[32], s_redB[64];   // A / B: the two reductions of a BiCG iteration (one barrier each)
  __shared__ int s_fi;
  __shared__ unsigned long long s_scan[POL_THREADS / 32];
  __shared__ int s_n[4];
  extern __shared__ double sm[];    // [POL_TILE*10] sweep tile; the solute tables of the contraction phase reuse it
  double *s_tile = sm;

  PolWork wk = wk0;
  {
    const size_t cb = blockIdx.x;
    const size_t cc = (size_t)p.cent_cap;
    wk.cent_pos += cb * cc * 3; wk.cent_alpha += cb * cc * AST; wk.fld += cb * cc * 3; wk.xy += cb * cc * (PST + 8);
    wk.site += cb * (size_t)p.site_cap * SITE_ROW; wk.cent_mol += cb * cc; wk.aux += cb * cc * 6;
    wk.ent_coff += cb * (size_t)p.list_cap; wk.ent_soff += cb * (size_t)p.list_cap;
    wk.part += cb * (size_t)(POL_THREADS / 32) * 2 * 64 * 6;
  }
  const int *tm0 = tmeta(p, p.inst_type[0]);
  const int npolc = tm0[SLVGPU_M_NPOL], ndmac = tm0[SLVGPU_M_NDMA];
  const int tid = threadIdx.x, NT = blockDim.x;

  for (;;) {
    // frames are handed out dynamically: their cost varies with the number of centres inside pcut and the
    // BiCG iteration count, so a static round-robin leaves SMs idle at the tail
    __syncthreads();
    if (tid == 0) s_fi = atomicAdd(queue, 1);
    __syncthreads();
    const int fi = s_fi;
    if (fi >= nframes) break;
    const int nl = fl.nlist[fi];
    const size_t lbase = (size_t)fi * p.list_cap;
    // Pass -1 is the many-body solve over the pol-layer fragments whose type is not flagged removable; every
    // removable fragment (EFP.eval(remove_clashes=True), solvshift/solefp.py:1224-1236, 1333-1404, 1722-1773) then
    // gets its own two-body solve with the solute, and the shifts add.
    double shift_main = 0.0, shift_add = 0.0, resid_out = 0.0, pairs_out = 0.0, ncent_out = 0.0, nsite_out = 0.0;
    int iters_out = 0;
    for (int pass = -1;;) {
    auto included = [&](int e) -> bool {
      if (!(fl.list_flag[lbase + e] & 1)) return false;
      const bool rem = tmeta(p, p.inst_type[fl.list_inst[lbase + e]])[SLVGPU_M_REMOVABLE] == 1;
      return pass < 0 ? !rem : e == pass;
    };
    do {
    // ---- (a) offsets of every pol-layer entry, then the lab-frame centre and site tables
    // offsets by a block-wide exclusive scan over the list entries (ordered, deterministic)
    {
      int run_c = npolc, run_s = ndmac;
      double sq = 0.0;                       // sum of npol^2 over the included fragments (for the pair count)
      for (int e0 = 0; e0 < nl; e0 += NT) {
        const int e = e0 + tid;
        int nc = 0, ns = 0;
        if (e < nl && included(e)) {
          const int *tm = tmeta(p, p.inst_type[fl.list_inst[lbase + e]]);
          nc = tm[SLVGPU_M_NPOL]; ns = tm[SLVGPU_M_NDMA];
          sq += (double)nc * nc;
        }
        // pack both counts into one 64-bit value for a single scan
        unsigned long long v = ((unsigned long long)nc << 32) | (unsigned)ns, incl = v;
        const int lane = tid & 31, w = tid >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        __syncthreads();
        if (lane == 31) s_scan[w] = incl;
        __syncthreads();
        unsigned long long base = 0, tot = 0;
        for (int i = 0; i < (NT >> 5); ++i) { if (i < w) base += s_scan[i]; tot += s_scan[i]; }
        const unsigned long long excl = base + incl - v;
        if (e < nl) { wk.ent_coff[e] = run_c + (int)(excl >> 32); wk.ent_soff[e] = run_s + (int)(excl & 0xffffffffu); }
        run_c += (int)(tot >> 32); run_s += (int)(tot & 0xffffffffu);
      }
      sq = block_sum(sq, s_red);
      if (tid == 0) { s_n[0] = run_c; s_n[1] = run_s; }
      if (pass < 0) pairs_out = (double)run_c * run_c - sq - (double)npolc * npolc;
    }
    __syncthreads();
    const int ncent = s_n[0], nsite = s_n[1];
    if (ncent > p.cent_cap || nsite > p.site_cap || ncent == npolc) {
      if (tid == 0 && ncent != npolc) status[frame0 + fi] |= SLVGPU_ST_OVERFLOW;
      break;                                   // nothing to solve in this pass
    }
    for (int e = tid - 1; e < nl; e += NT) {          // e == -1 is the solute
      const int *tm; double R[9], t[3]; int c0, s0, mol;
      if (e < 0) {
        tm = tm0; const double *sx = fl.sol_xf + (size_t)fi * 12;
        for (int d = 0; d < 9; ++d) R[d] = sx[d];
        for (int d = 0; d < 3; ++d) t[d] = sx[9 + d];
        c0 = 0; s0 = 0; mol = 0;
      } else {
        if (!included(e)) continue;
        tm = tmeta(p, p.inst_type[fl.list_inst[lbase + e]]);
        const double *xf = fl.list_xf + (lbase + e) * 12;
        for (int d = 0; d < 9; ++d) R[d] = xf[d];
        for (int d = 0; d < 3; ++d) t[d] = xf[9 + d];
        c0 = wk.ent_coff[e]; s0 = wk.ent_soff[e]; mol = e + 1;
      }
      const int np_ = tm[SLVGPU_M_NPOL], nd = tm[SLVGPU_M_NDMA];
      for (int l = 0; l < np_; ++l) {
        double y[3], a[9];
        rot_vec(R, p.dtab + tm[SLVGPU_M_O_RPOL] + l * 3, y);
        rot_mat(R, p.dtab + tm[SLVGPU_M_O_POL] + l * 9, a);
        for (int d = 0; d < 3; ++d) wk.cent_pos[(size_t)(c0 + l) * 3 + d] = y[d] + t[d];
        for (int d = 0; d < 9; ++d) wk.cent_alpha[(size_t)(c0 + l) * AST + d] = a[d];
        wk.cent_mol[c0 + l] = mol;
      }
      for (int s = 0; s < nd; ++s) {
        double y[3], m[20];
        rot_vec(R, p.dtab + tm[SLVGPU_M_O_RDMA] + s * 3, y);
        rot_mom20(R, p.dtab + tm[SLVGPU_M_O_MOM] + s * 20, m);
        double *row = wk.site + (size_t)(s0 + s) * SITE_ROW;
        for (int d = 0; d < 3; ++d) row[d] = y[d] + t[d];
        for (int d = 0; d < 20; ++d) row[3 + d] = m[d];
        row[23] = (double)mol;
      }
    }
    __syncthreads();

    // ---- (b) fields at every centre from the multipoles of all OTHER molecules (FFFFFF); sites staged
    //      through shared memory in tiles, own-molecule sites masked branch-free
    {
      constexpr int FT = POL_TILE * 10 / SITE_ROW;          // sites per tile: the sweep tile's shared memory, 426 sites
      for (int c0 = 0; c0 < ncent; c0 += NT) {
        const int c = c0 + tid;
        const bool act = c < ncent;
        double rc[3] = {0, 0, 0}; int mc = -1;
        if (act) { rc[0] = wk.cent_pos[(size_t)c * 3]; rc[1] = wk.cent_pos[(size_t)c * 3 + 1]; rc[2] = wk.cent_pos[(size_t)c * 3 + 2]; mc = wk.cent_mol[c]; }
        double F[3] = {0, 0, 0};
        for (int s0 = 0; s0 < nsite; s0 += FT) {
          const int sn = min(FT, nsite - s0);
          __syncthreads();
          for (int w = tid; w < sn * SITE_ROW; w += NT) s_tile[w] = wk.site[(size_t)s0 * SITE_ROW + w];
          __syncthreads();
          if (act) {
#pragma unroll 2
            for (int s = 0; s < sn; ++s) {
              const double *row = s_tile + s * SITE_ROW;
              const bool same = ((int)row[23] == mc);
              const double R[3] = {same ? 1.0 : rc[0] - row[0], same ? 0.0 : rc[1] - row[1], same ? 0.0 : rc[2] - row[2]};
              site_field_acc(R, row + 3, 5.0, same ? 0.0 : 1.0, F);
            }
          }
        }
        if (act) {
          const double *a = wk.cent_alpha + (size_t)c * AST;
          wk.fld[(size_t)c * 3] = F[0]; wk.fld[(size_t)c * 3 + 1] = F[1]; wk.fld[(size_t)c * 3 + 2] = F[2];
          double *v = wk.xy + (size_t)c * PST;               // state slot 0..2 <- b = alpha F
          for (int i = 0; i < 3; ++i) v[i] = a[i * 3] * F[0] + a[i * 3 + 1] * F[1] + a[i * 3 + 2] * F[2];
        }
      }
    }
    __syncthreads();

    // ---- (c) BiCG on the block-preconditioned system  A x = alpha F,  A = I + alpha T,  together with the
    //      transposed system A^T xt = F (xt = alpha^-T y): one sweep over the centre pairs applies T to both
    //      search directions.  State per centre (PST doubles): x xt r rt p pt u ut q qt.
    int iters = 0;
    double resid = 1.0;
    double *stt = wk.xy;                                   // [ncent][PST]
    double *vfin = wk.xy + (size_t)p.cent_cap * PST;        // [ncent][6]  final x, y
    // b = state slot 0..2 (already alpha-preconditioned), bt = bt_src[c][3]; result x and y = alpha^T xt in vfin
    auto bicg = [&](const double *bt_src) {
    iters = 0; resid = 1.0;
    double rho = 0.0, bnorm = 0.0;
    {
      double prho = 0.0, pb = 0.0;
      for (int c = tid; c < ncent; c += NT) {
        double *s_ = stt + (size_t)c * PST;
        const double *F = bt_src + (size_t)c * 3;
        for (int i = 0; i < 3; ++i) {
          const double bi = s_[i];                         // preconditioned right-hand side
          s_[6 + i] = bi; s_[12 + i] = bi;                 // r = p = b
          s_[9 + i] = F[i]; s_[15 + i] = F[i];             // rt = pt = F
          s_[i] = 0.0; s_[3 + i] = 0.0;                    // x = xt = 0
          prho += F[i] * bi; pb = fmax(pb, fmax(fabs(bi), fabs(F[i])));
        }
      }
      rho = block_sum(prho, s_red);
      bnorm = block_max(pb, s_red);
    }
    bool broke = false;
    // When all centres fit one tile (the usual case: <= 1024 inside pcut) the tile stays resident in shared memory for
    // the whole solve: positions and molecule ids are staged once, and every iteration the owners write their search
    // directions straight into it -- no per-iteration round trip through the L2 workspace, one barrier per sweep.
    const bool resident = ncent <= POL_TILE;
    if (resident) {
      for (int j = tid; j < ncent; j += NT) {
        double *tt = s_tile + j * 10;
        tt[0] = wk.cent_pos[(size_t)j * 3]; tt[1] = wk.cent_pos[(size_t)j * 3 + 1]; tt[2] = wk.cent_pos[(size_t)j * 3 + 2];
        tt[9] = (double)wk.cent_mol[j];
      }
    }
    if (bnorm > 0.0) {
      // sweep operands of the first iteration: u = p, ut = alpha^T pt (later iterations write them where p, pt are updated)
      for (int c = tid; c < ncent; c += NT) {
        double *s_ = stt + (size_t)c * PST;
        const double *a = wk.cent_alpha + (size_t)c * AST;
        double *dst = resident ? s_tile + c * 10 + 3 : s_ + 18;
        for (int i = 0; i < 3; ++i) {
          dst[i] = s_[12 + i];
          dst[3 + i] = a[i] * s_[15] + a[3 + i] * s_[16] + a[6 + i] * s_[17];
        }
      }
      for (; iters < p.pol_maxit;) {
        if (resident) __syncthreads();
        double pq = 0.0;
        // q = p + alpha (T p), qt = pt + T ut for one row, from the accumulated tensor applications
        auto finish = [&](int c, const double *ax, const double *ay) {
          double *s_ = stt + (size_t)c * PST;
          double v[8], a[12], q[3], qt[3];                    // v = p(3) pt(3) u(2)
          ld4(s_ + 12, v); ld4(s_ + 16, v + 4);
          const double *ag = wk.cent_alpha + (size_t)c * AST;
          ld4(ag, a); ld4(ag + 4, a + 4); ld4(ag + 8, a + 8);
          for (int i = 0; i < 3; ++i) {
            q[i] = v[i] + a[i * 3] * ax[0] + a[i * 3 + 1] * ax[1] + a[i * 3 + 2] * ax[2];
            qt[i] = v[3 + i] + ay[i];
            pq += v[3 + i] * q[i];
          }
          st4(s_ + 24, q[0], q[1], q[2], qt[0]); st4(s_ + 28, qt[1], qt[2], 0.0, 0.0);
        };
        if (resident) {
          // Balanced sweep.  Rows are taken in blocks of 64 (two per lane: the tile entry of column j is read from
          // shared memory once for both rows and the twelve independent accumulation chains keep the FP64 pipe fed
          // with four warps per scheduler).  The (block, column) index space, nb2 x ncent, is cut into equal
          // contiguous ranges, one per warp, so the sweep costs ncent^2 / 256 per thread for every ncent instead of
          // whole passes of 512 rows (403 centres: 0.79 of two passes used; 520 centres: a third pass for 8 rows).
          // A range that covers a block only partly leaves partial sums in the workspace (head = slot 0, tail =
          // slot 1); after one barrier they are added in ascending warp order -- deterministic.
          const int warp = tid >> 5, lane = tid & 31, nwarp = NT >> 5;
          const int nb2 = (ncent + (1 << POL_BSH) - 1) >> POL_BSH;
          // cost-weighted ranges: a (block, column) unit costs 2 (two rows per lane) except in a last block whose second
          // row half is empty (one-row loop, cost 1); positions in cost space are even, so they map back to whole units
          const int nfull = nb2 - 1, ufull = nfull * ncent;
          const int clast = (POL_ROWS == 2 && ncent - (nfull << POL_BSH) <= 32) ? 1 : 2;
          const int F = 2 * ufull, W2 = F + clast * ncent;
          int Qc = (W2 + nwarp - 1) / nwarp; Qc += Qc & 1;
          auto unit_of = [&](int x) { x = min(x, W2); return x <= F ? x >> 1 : ufull + (x - F) / clast; };
          auto warp_of = [&](int u) { return (u < ufull ? 2 * u : F + (u - ufull) * clast) / Qc; };
          bool split = false;
          for (int w2 = 1; w2 < nwarp; ++w2) split |= (unit_of(w2 * Qc) % ncent) != 0;
          const int u1 = unit_of((warp + 1) * Qc);
          for (int u = unit_of(warp * Qc); u < u1;) {
            const int b = u / ncent, ja = u - b * ncent, jb = min(ncent, ja + (u1 - u));
            const int ca = (b << POL_BSH) + lane, cb = ca + 32;
            const bool acta = ca < ncent, actb = POL_ROWS == 2 && cb < ncent;
            const bool anyb = POL_ROWS == 2 && __any_sync(0xffffffffu, actb);
            double ra[3] = {0, 0, 0}, rb[3] = {0, 0, 0};
            int ma = -1, mb = -1;          // molecule ids: small integers held as doubles in the tile, compared through their high words
                                           // (an integer compare: a double compare would take a slot of the FP64 pipe per pair)
            if (acta) { ra[0] = s_tile[ca * 10]; ra[1] = s_tile[ca * 10 + 1]; ra[2] = s_tile[ca * 10 + 2]; ma = __double2hiint(s_tile[ca * 10 + 9]); }
            if (actb) { rb[0] = s_tile[cb * 10]; rb[1] = s_tile[cb * 10 + 1]; rb[2] = s_tile[cb * 10 + 2]; mb = __double2hiint(s_tile[cb * 10 + 9]); }
            double axa[3] = {0, 0, 0}, aya[3] = {0, 0, 0}, axb[3] = {0, 0, 0}, ayb[3] = {0, 0, 0};
            if (anyb) {
#pragma unroll 2
              for (int jj = ja; jj < jb; ++jj) {
                const double2 *t2 = reinterpret_cast<const double2 *>(s_tile + jj * 10);
                const double2 q0 = t2[0], q1 = t2[1], q2 = t2[2], q3 = t2[3], q4 = t2[4];
                const double tt[9] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y, q4.x};
                const int mj = __double2hiint(q4.y);
                sweep_pair(ra, mj == ma, tt, axa, aya);
                sweep_pair(rb, mj == mb, tt, axb, ayb);
              }
            } else if (acta) {
#pragma unroll 4
              for (int jj = ja; jj < jb; ++jj) {
                const double2 *t2 = reinterpret_cast<const double2 *>(s_tile + jj * 10);
                const double2 q0 = t2[0], q1 = t2[1], q2 = t2[2], q3 = t2[3], q4 = t2[4];
                const double tt[9] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y, q4.x};
                sweep_pair(ra, __double2hiint(q4.y) == ma, tt, axa, aya);
              }
            }
            if (ja == 0 && jb == ncent) {                       // the whole block: finish its rows here
              if (acta) finish(ca, axa, aya);
              if (actb) finish(cb, axb, ayb);
            } else {
              double *pp = wk.part + ((size_t)(warp * 2 + (ja > 0 ? 0 : 1)) * 64 + lane) * 6;
              for (int d = 0; d < 3; ++d) { pp[d] = axa[d]; pp[3 + d] = aya[d]; }
              if (POL_ROWS == 2) for (int d = 0; d < 3; ++d) { pp[192 + d] = axb[d]; pp[195 + d] = ayb[d]; }
            }
            u += jb - ja;
          }
          if (split) {                                            // CTA-uniform
            __syncthreads();
            for (int b = warp; b < nb2; b += nwarp) {
              const int wf = warp_of(b * ncent), wl = warp_of(b * ncent + ncent - 1);
              if (wf == wl) continue;                             // one warp had the whole block
              double axa[3] = {0, 0, 0}, aya[3] = {0, 0, 0}, axb[3] = {0, 0, 0}, ayb[3] = {0, 0, 0};
              for (int w2 = wf; w2 <= wl; ++w2) {
                const double *pp = wk.part + ((size_t)(w2 * 2 + (b * ncent < unit_of(w2 * Qc) ? 0 : 1)) * 64 + lane) * 6;
                for (int d = 0; d < 3; ++d) { axa[d] += pp[d]; aya[d] += pp[3 + d]; }
                if (POL_ROWS == 2) for (int d = 0; d < 3; ++d) { axb[d] += pp[192 + d]; ayb[d] += pp[195 + d]; }
              }
              const int ca = (b << POL_BSH) + lane, cb = ca + 32;
              if (ca < ncent) finish(ca, axa, aya);
              if (POL_ROWS == 2 && cb < ncent) finish(cb, axb, ayb);
            }
          }
        } else {
          // more centres than the resident tile holds: one row per thread, columns tile by tile
          for (int c0 = 0; c0 < ncent; c0 += NT) {
            const int ca = c0 + tid;
            const bool acta = ca < ncent;
            double ra[3] = {0, 0, 0}; int ma = -1;
            if (acta) { ra[0] = wk.cent_pos[(size_t)ca * 3]; ra[1] = wk.cent_pos[(size_t)ca * 3 + 1]; ra[2] = wk.cent_pos[(size_t)ca * 3 + 2]; ma = __double2hiint((double)wk.cent_mol[ca]); }
            double axa[3] = {0, 0, 0}, aya[3] = {0, 0, 0};
            for (int j0 = 0; j0 < ncent; j0 += POL_TILE) {
              __syncthreads();
              for (int tl = tid; tl < POL_TILE; tl += NT) {
                const int j = j0 + tl;
                if (j < ncent) {
                  double *tt = s_tile + tl * 10;
                  tt[0] = wk.cent_pos[(size_t)j * 3]; tt[1] = wk.cent_pos[(size_t)j * 3 + 1]; tt[2] = wk.cent_pos[(size_t)j * 3 + 2];
                  const double *vj = stt + (size_t)j * PST + 18;
                  for (int d = 0; d < 6; ++d) tt[3 + d] = vj[d];
                  tt[9] = (double)wk.cent_mol[j];
                }
              }
              __syncthreads();
              const int jn = min(POL_TILE, ncent - j0);
              if (acta) {
#pragma unroll 2
                for (int jj = 0; jj < jn; ++jj) {
                  const double2 *t2 = reinterpret_cast<const double2 *>(s_tile + jj * 10);
                  const double2 q0 = t2[0], q1 = t2[1], q2 = t2[2], q3 = t2[3], q4 = t2[4];
                  const double tt[9] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y, q4.x};
                  sweep_pair(ra, __double2hiint(q4.y) == ma, tt, axa, aya);
                }
              }
            }
            if (acta) finish(ca, axa, aya);
          }
        }
        pq = block_sum_1b(pq, s_redA);
        ++iters;
        if (!(fabs(pq) > 0.0) || !isfinite(pq)) { broke = true; break; }
        const double al = rho / pq;
        double prho = 0.0, pr = 0.0;
        for (int c = tid; c < ncent; c += NT) {
          double *s_ = stt + (size_t)c * PST;
          double v[20], w[8];                                  // v = x xt r rt p pt u(2), w = q qt pad(2)
          ld4(s_, v); ld4(s_ + 4, v + 4); ld4(s_ + 8, v + 8); ld4(s_ + 12, v + 12); ld4(s_ + 16, v + 16);
          ld4(s_ + 24, w); ld4(s_ + 28, w + 4);
          for (int i = 0; i < 3; ++i) {
            v[i] += al * v[12 + i]; v[3 + i] += al * v[15 + i];
            const double r = v[6 + i] - al * w[i], rt = v[9 + i] - al * w[3 + i];
            v[6 + i] = r; v[9 + i] = rt;
            prho += rt * r; pr = fmax(pr, fmax(fabs(r), fabs(rt)));
          }
          st4(s_, v[0], v[1], v[2], v[3]); st4(s_ + 4, v[4], v[5], v[6], v[7]); st4(s_ + 8, v[8], v[9], v[10], v[11]);
        }
        double rho1 = prho, rmax = pr;
        block_sum_max_1b(rho1, rmax, s_redB);
        resid = rmax / bnorm;
        if (!(resid > p.pol_tol)) break;
        if (!(fabs(rho) > 0.0) || !isfinite(rho1)) { broke = true; break; }
        const double be = rho1 / rho;
        rho = rho1;
        for (int c = tid; c < ncent; c += NT) {                 // new search directions, and the next sweep's operands
          double *s_ = stt + (size_t)c * PST;
          double *dst = resident ? s_tile + c * 10 + 3 : s_ + 18;
          double v[16], a[12], pn[3], ptn[3];                  // v = s_[4..19]: xt(2) r rt p pt u(2)
          ld4(s_ + 4, v); ld4(s_ + 8, v + 4); ld4(s_ + 12, v + 8); ld4(s_ + 16, v + 12);
          const double *ag = wk.cent_alpha + (size_t)c * AST;
          ld4(ag, a); ld4(ag + 4, a + 4); ld4(ag + 8, a + 8);
          for (int i = 0; i < 3; ++i) { pn[i] = v[2 + i] + be * v[8 + i]; ptn[i] = v[5 + i] + be * v[11 + i]; }
          st4(s_ + 12, pn[0], pn[1], pn[2], ptn[0]); st4(s_ + 16, ptn[1], ptn[2], v[14], v[15]);
          for (int i = 0; i < 3; ++i) {
            dst[i] = pn[i];
            dst[3 + i] = a[i] * ptn[0] + a[3 + i] * ptn[1] + a[6 + i] * ptn[2];
          }
        }
      }
    } else resid = 0.0;
    for (int c = tid; c < ncent; c += NT) {       // x, and y = alpha^T xt
      const double *s_ = stt + (size_t)c * PST;
      const double *a = wk.cent_alpha + (size_t)c * AST;
      double *v = vfin + (size_t)c * 6;
      for (int i = 0; i < 3; ++i) {
        v[i] = s_[i];
        v[3 + i] = a[i] * s_[3] + a[3 + i] * s_[4] + a[6 + i] * s_[5];
      }
    }
    __syncthreads();
    if (broke) resid = 1.0;
    };
    bicg(wk.fld);
    const double *vxy = vfin;

    // ---- (d) contractions with the solute rows: y^T dD x - (x+y).dF
    // solute tables in shared memory: centres [npolc][21]: pos3 r1(3) G(9) x3 y3 ; sites [ndmac][46]: pos3 mom20 dma1(20) L3
    double *s_c = sm, *s_s = s_c + (size_t)npolc * 21;      // the sweep tile is dead here (every pass restages it)
    double acc = 0.0;
    const bool global = (p.flags & SLVGPU_GLOBAL) != 0;
    if (global) {
      // EFP2 polarisation ENERGY of the whole cluster (SOLPOL, solvshift/src/solpol2.f:1153-1267): EPOL = -1/2 F . D^-1 F
      for (int c = tid; c < ncent; c += NT)
        acc += wk.fld[(size_t)c * 3] * vxy[(size_t)c * 6] + wk.fld[(size_t)c * 3 + 1] * vxy[(size_t)c * 6 + 1] + wk.fld[(size_t)c * 3 + 2] * vxy[(size_t)c * 6 + 2];
    } else {
    {
      const double *sx = fl.sol_xf + (size_t)fi * 12;
      double R[9];
      for (int d = 0; d < 9; ++d) R[d] = sx[d];
      for (int i = tid; i < npolc; i += NT) {
        double *row = s_c + (size_t)i * 21;
        for (int d = 0; d < 3; ++d) row[d] = wk.cent_pos[(size_t)i * 3 + d];
        rot_vec(R, p.dtab + p.sol_meta[SLVGPU_S_LMOC1] + i * 3, row + 3);
        rot_mat(R, p.dtab + p.sol_meta[SLVGPU_S_DPOL1] + i * 9, row + 6);
        for (int d = 0; d < 6; ++d) row[15 + d] = vxy[(size_t)i * 6 + d];
      }
      for (int a = tid; a < ndmac; a += NT) {
        double *row = s_s + (size_t)a * 46;
        const double *src = wk.site + (size_t)a * SITE_ROW;
        for (int d = 0; d < 23; ++d) row[d] = src[d];
        rot_mom20(R, p.dtab + p.sol_meta[SLVGPU_S_DMA1] + a * 20, row + 23);
        rot_vec(R, p.dtab + p.sol_meta[SLVGPU_S_LVEC] + a * 3, row + 43);
      }
    }
    __syncthreads();
    // term A: solute diagonal blocks  y_i^T G_i x_i,  G = -alpha^-1 (sum_m g_m alpha1_m) alpha^-1
    for (int i = tid; i < npolc; i += NT) {
      const double *row = s_c + (size_t)i * 21, *G = row + 6, *xi = row + 15, *yi = row + 18;
      for (int a = 0; a < 3; ++a) acc += yi[a] * (G[a * 3] * xi[0] + G[a * 3 + 1] * xi[1] + G[a * 3 + 2] * xi[2]);
    }
    // terms B and C2: one thread per solvent centre j
    for (int j = npolc + tid; j < ncent; j += NT) {
      const double rj[3] = {wk.cent_pos[(size_t)j * 3], wk.cent_pos[(size_t)j * 3 + 1], wk.cent_pos[(size_t)j * 3 + 2]};
      const double *vj = vxy + (size_t)j * 6;
      const double xj[3] = {vj[0], vj[1], vj[2]}, yj[3] = {vj[3], vj[4], vj[5]};
      for (int i = 0; i < npolc; ++i) {
        const double *row = s_c + (size_t)i * 21, *r1 = row + 3, *xi = row + 15, *yi = row + 18;
        const double R[3] = {row[0] - rj[0], row[1] - rj[1], row[2] - rj[2]};
        const double r2 = R[0] * R[0] + R[1] * R[1] + R[2] * R[2];
        const double u = rsqrt64(r2), u2 = u * u, u5 = u2 * u2 * u;
        const double R1R = R[0] * r1[0] + R[1] * r1[1] + R[2] * r1[2];
        const double Ryi = R[0] * yi[0] + R[1] * yi[1] + R[2] * yi[2], Rxj = R[0] * xj[0] + R[1] * xj[1] + R[2] * xj[2];
        const double Ryj = R[0] * yj[0] + R[1] * yj[1] + R[2] * yj[2], Rxi = R[0] * xi[0] + R[1] * xi[1] + R[2] * xi[2];
        const double yixj = yi[0] * xj[0] + yi[1] * xj[1] + yi[2] * xj[2], yjxi = yj[0] * xi[0] + yj[1] * xi[1] + yj[2] * xi[2];
        const double r1xj = r1[0] * xj[0] + r1[1] * xj[1] + r1[2] * xj[2], r1yi = r1[0] * yi[0] + r1[1] * yi[1] + r1[2] * yi[2];
        const double r1xi = r1[0] * xi[0] + r1[1] * xi[1] + r1[2] * xi[2], r1yj = r1[0] * yj[0] + r1[1] * yj[1] + r1[2] * yj[2];
        const double b1 = R1R * (yixj - 5.0 * u2 * Ryi * Rxj) + Ryi * r1xj + r1yi * Rxj;
        const double b2 = R1R * (yjxi - 5.0 * u2 * Ryj * Rxi) + Ryj * r1xi + r1yj * Rxi;
        acc += -1.5 * u5 * (b1 + b2);
      }
      double dF[3] = {0, 0, 0};
      for (int a = 0; a < ndmac; ++a) {
        const double *row = s_s + (size_t)a * 46;
        const double R[3] = {rj[0] - row[0], rj[1] - row[1], rj[2] - row[2]};
        site_field(R, row + 23, 7.0, dF);
        site_dfield(R, row + 43, row + 3, -1.0, dF);
      }
      acc -= (xj[0] + yj[0]) * dF[0] + (xj[1] + yj[1]) * dF[1] + (xj[2] + yj[2]) * dF[2];
    }
    // term C1: one thread per solvent site s, field derivative at the solute centres moving by r1
    for (int s = ndmac + tid; s < nsite; s += NT) {
      const double *srow = wk.site + (size_t)s * SITE_ROW;
      double m[20];
      for (int d = 0; d < 20; ++d) m[d] = srow[3 + d];
      const double rs[3] = {srow[0], srow[1], srow[2]};
      for (int i = 0; i < npolc; ++i) {
        const double *row = s_c + (size_t)i * 21;
        const double R[3] = {row[0] - rs[0], row[1] - rs[1], row[2] - rs[2]};
        double dF[3] = {0, 0, 0};
        site_dfield(R, row + 3, m, 1.0, dF);
        acc -= (row[15] + row[18]) * dF[0] + (row[16] + row[19]) * dF[1] + (row[17] + row[20]) * dF[2];
      }
    }
    }   // frequency-shift contractions
    acc = block_sum(acc, s_red);
    if (pass < 0) { shift_main = -0.5 * acc; iters_out = iters; resid_out = resid; ncent_out = ncent; nsite_out = nsite; }
    else { shift_add += -0.5 * acc; resid_out = fmax(resid_out, resid); }
    if (detail && pass < 0) {
      const size_t drow = (size_t)p.cent_cap * 12 + 1;
      double *dd = detail + (size_t)fi * drow;
      for (int c = tid; c < ncent; c += NT)
        for (int d = 0; d < 3; ++d) {
          dd[(size_t)c * 12 + d] = vxy[(size_t)c * 6 + d];
          dd[(size_t)c * 12 + 3 + d] = wk.fld[(size_t)c * 3 + d];
          dd[(size_t)c * 12 + 6 + d] = wk.cent_pos[(size_t)c * 3 + d];
        }
      if (tid == 0) dd[(size_t)p.cent_cap * 12] = (double)ncent;
      if (!global) {
      // ---- solvatochromic induced dipoles AVEC = sum_m g_m MIVEC_m = D^-T (dD^T y - dF) - D^-1 dF (solpol2.f:1122-1143):
      //      the vectors dF and w = dD^T y, then one more BiCG run with the right-hand sides (dF, w - dF)
      double *dFv = wk.aux, *btv = wk.aux + (size_t)p.cent_cap * 3;
      auto Bapply = [&](const double *R, const double *r1, const double *v, double *o) {   // o += B_ij v
        const double r2 = R[0] * R[0] + R[1] * R[1] + R[2] * R[2];
        const double u = rsqrt64(r2), u2 = u * u, u5 = u2 * u2 * u, f = -1.5 * u5;
        const double R1R = R[0] * r1[0] + R[1] * r1[1] + R[2] * r1[2];
        const double Rv = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], r1v = r1[0] * v[0] + r1[1] * v[1] + r1[2] * v[2];
        for (int a = 0; a < 3; ++a) o[a] += f * (R1R * (v[a] - 5.0 * u2 * R[a] * Rv) + R[a] * r1v + r1[a] * Rv);
      };
      for (int j = npolc + tid; j < ncent; j += NT) {                 // solvent centres
        const double rj[3] = {wk.cent_pos[(size_t)j * 3], wk.cent_pos[(size_t)j * 3 + 1], wk.cent_pos[(size_t)j * 3 + 2]};
        double dF[3] = {0, 0, 0}, w[3] = {0, 0, 0};
        for (int a = 0; a < ndmac; ++a) {
          const double *row = s_s + (size_t)a * 46;
          const double R[3] = {rj[0] - row[0], rj[1] - row[1], rj[2] - row[2]};
          site_field(R, row + 23, 7.0, dF);
          site_dfield(R, row + 43, row + 3, -1.0, dF);
        }
        for (int i = 0; i < npolc; ++i) {
          const double *row = s_c + (size_t)i * 21;
          const double R[3] = {row[0] - rj[0], row[1] - rj[1], row[2] - rj[2]};
          Bapply(R, row + 3, row + 18, w);
        }
        for (int d = 0; d < 3; ++d) { dFv[(size_t)j * 3 + d] = dF[d]; btv[(size_t)j * 3 + d] = w[d] - dF[d]; }
      }
      for (int i = tid; i < npolc; i += NT) {                          // solute centres
        const double *row = s_c + (size_t)i * 21, *G = row + 6, *yi = row + 18;
        double dF[3] = {0, 0, 0}, w[3];
        for (int a = 0; a < 3; ++a) w[a] = G[a] * yi[0] + G[3 + a] * yi[1] + G[6 + a] * yi[2];      // G^T y
        for (int sidx = ndmac; sidx < nsite; ++sidx) {
          const double *srow = wk.site + (size_t)sidx * SITE_ROW;
          const double R[3] = {row[0] - srow[0], row[1] - srow[1], row[2] - srow[2]};
          site_dfield(R, row + 3, srow + 3, 1.0, dF);
        }
        for (int j = npolc; j < ncent; ++j) {
          const double R[3] = {row[0] - wk.cent_pos[(size_t)j * 3], row[1] - wk.cent_pos[(size_t)j * 3 + 1], row[2] - wk.cent_pos[(size_t)j * 3 + 2]};
          Bapply(R, row + 3, vxy + (size_t)j * 6 + 3, w);
        }
        for (int d = 0; d < 3; ++d) { dFv[(size_t)i * 3 + d] = dF[d]; btv[(size_t)i * 3 + d] = w[d] - dF[d]; }
      }
      __syncthreads();
      for (int c = tid; c < ncent; c += NT) {                          // preconditioned rhs b = alpha dF
        const double *a = wk.cent_alpha + (size_t)c * AST, *F = dFv + (size_t)c * 3;
        double *v = wk.xy + (size_t)c * PST;
        for (int i = 0; i < 3; ++i) v[i] = a[i * 3] * F[0] + a[i * 3 + 1] * F[1] + a[i * 3 + 2] * F[2];
      }
      __syncthreads();
      const int it1 = iters; const double res1 = resid;
      bicg(btv);
      for (int c = tid; c < ncent; c += NT)
        for (int d = 0; d < 3; ++d) dd[(size_t)c * 12 + 9 + d] = vfin[(size_t)c * 6 + 3 + d] - vfin[(size_t)c * 6 + d];
      resid = fmax(res1, resid); iters = it1;
      iters_out = iters; resid_out = resid;
      }   // AVEC (frequency-shift mode only)
    }
    } while (0);
    // next removable fragment inside pcut, if any
    __syncthreads();
    if (tid == 0) {
      int nx = -1;
      for (int e = pass + 1; e < nl; ++e)
        if ((fl.list_flag[lbase + e] & 1) && tmeta(p, p.inst_type[fl.list_inst[lbase + e]])[SLVGPU_M_REMOVABLE] == 1) { nx = e; break; }
      s_n[2] = nx;
    }
    __syncthreads();
    pass = s_n[2];
    if (pass < 0) break;
    }   // passes
    if (tid == 0) {
      double *o = out + (size_t)(frame0 + fi) * SLVGPU_NOUT;
      o[SLVGPU_POL_MEA] = shift_main + shift_add;
      o[SLVGPU_POL_ADD] = shift_add;
      o[SLVGPU_POL_ITERS] = (double)iters_out;
      o[SLVGPU_POL_RESID] = resid_out;
      o[SLVGPU_POL_NCENT] = ncent_out; o[SLVGPU_POL_NSITE] = nsite_out; o[SLVGPU_POL_PAIRS] = pairs_out;
      if (resid_out > p.pol_tol) status[frame0 + fi] |= SLVGPU_ST_POL_NOCONV;
    }
    __syncthreads();
  }
}

}  // namespace slv
