// k_rep.cuh -- kernel group D: exchange-repulsion frequency shift, evaluated in batches of (frame, fragment) pairs.
//
// Reference: per solvent molecule B the Python layer builds both basis sets, calls PyQuante for the AO
// integrals S, T, dS/dA, dT/dA (solvshift/solefp.py:1497-1526) and then SHFTEX (solvshift/src/shftex.f):
//   CALCIJ (298-438): S_ij = C_A S C_B^T, T_ij, and per mode  S^m_ij = C_A (L^m_atom(k) . dS_kl/dA) C_B^T + C1^m_A S C_B^T
//   CALCFI (69-295) : fi_m = sum_ij TS1..4 + TA1..4  (EFP2 spherical-Gaussian-overlap form)
//   SHFTMA = -sum_m g_m fi_m .
// fi_m is linear in the mode quantities (S^m, T^m, fock1_m, lmoc1_m, lvec_m), so the plan's g-contracted tables
// turn the 30-mode loop into one evaluation.  Per batch of frames (as many as the integral pool holds):
//   k_rep_scan   exclusive scan of the per-frame counts of fragments inside ecut -> pair slots
//   k_rep_prep   per frame: lab-frame solute atoms and displacement vectors; per (frame, fragment) pair: the
//                fragment's atoms, and the pair -> (frame, list entry, type) map
//   k_rep_ints   AO integrals V[pair][k][l][4] = S, T, L.dS, L.dT: one launch per (fragment type, sub-shell class),
//                one thread per (pair, sub-shell-pair item), grid-wide -- every warp runs one specialisation of
//                slv_ints.cuh with near-equal loop lengths, at full occupancy
//   k_rep_pair   one CTA per (frame, fragment) pair:
//                2  half transform  X[k][q][j] = sum_l V[k][l][q] C_B[j][l]
//                3  LMO integrals   S_ij, T_ij, Sm_ij, Tm_ij
//                4  CALCFI per (i,j), block sum -> res[pair]
//   k_rep_sum    rep_mea of a frame = -(sum of its pair results, in list order)
// AO->LMO coefficients are rotated into the lab frame per fragment (VECROT, efprot.f:22-151): the solute's once per
// frame (k_rep_prep), the fragment's in k_rep_pair.
#pragma once
#include <cstdlib>
#include <string>
#include <vector>
#include "slv_common.cuh"
#include "slv_ints.cuh"

namespace slv {

constexpr int REP_THREADS = 128;
constexpr int REP_INT_THREADS = 128;

struct RepPool {            // one batch of (frame, fragment) pairs (V .. pair_type) + per-frame tables of the chunk
  double *V;       // [pair_cap][nbsA][max_nbas][4]  AO integrals
  double *posA;    // [max_chunk][natA][3]   solute atoms, lab frame
  double *LcA;     // [max_chunk][natA][3]   mode-contracted displacement vectors, lab frame
  double *riA;     // [max_chunk][2][nmosA][3]  LMO centroids and their mode-contracted derivatives
  double *CA;      // [max_chunk][2][nmosA][nbsA]  rotated vecl and g-contracted vecl1 of the solute
  double *posB;    // [pair_cap][max_nat][3]
  double *res;     // [pair_cap]  sum_ij fi of the pair
  int *pair_frame, *pair_entry, *pair_type;   // [pair_cap]
  int *pair0;      // [max_chunk + 1] exclusive scan of the per-frame pair counts
  size_t sV;
  int pair_cap, max_nat;
};

// Two pools (the batches alternate between two streams, so the integral kernels of one batch overlap the
// latency-bound transform kernel of the previous one); the per-frame tables are shared.  rep_alloc creates the
// per-frame tables and fixes the largest pool the plan may use; the pools themselves grow on demand (rep_reserve),
// so a plan that only ever sees a few fragments inside ecut does not hold gigabytes.
struct RepLimits { size_t max_pairs = 0; };

inline int rep_alloc(RepPool rp[2], RepLimits &lim, std::vector<void *> &allocs, std::string &err, int nmosc, int nbasc,
                     int natc, int max_nbas, int max_nat, int max_chunk, int list_cap) {
  // integral pools: together up to 8 GiB and at most a quarter of the free memory, never more than the worst case
  const size_t sV = (size_t)nbasc * max_nbas * 4;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  size_t budget = (size_t)8 << 30;
  if (budget > free_b / 4) budget = free_b / 4;
  size_t cap = budget / 2 / (sV * sizeof(double));
  const size_t want = (size_t)max_chunk * (size_t)list_cap;
  if (cap > want) cap = want;
  if (const char *ev = getenv("SLVGPU_REP_POOL_PAIRS")) {      // testing knob: force small batches
    const long v = atol(ev);
    if (v > 0 && (size_t)v < cap) cap = (size_t)v;
  }
  if (cap < 1) cap = 1;
  if (cap > 0x3fffffff) cap = 0x3fffffff;
  lim.max_pairs = cap;
  void *q = nullptr;
#define RA(ptr, n, T)                                                                                                   \
  if (cudaMalloc(&q, (size_t)(n) * sizeof(T)) != cudaSuccess) { err = "cudaMalloc(rep tables) failed"; return SLVGPU_ENOMEM; } \
  allocs.push_back(q); ptr = (T *)q;
  RA(rp[0].posA, (size_t)max_chunk * natc * 3, double)
  RA(rp[0].LcA, (size_t)max_chunk * natc * 3, double)
  RA(rp[0].riA, (size_t)max_chunk * nmosc * 6, double)
  RA(rp[0].CA, (size_t)max_chunk * 2 * nmosc * nbasc, double)
  RA(rp[0].pair0, (size_t)max_chunk + 1, int)
#undef RA
  rp[1] = rp[0];
  for (int h = 0; h < 2; ++h) { rp[h].sV = sV; rp[h].max_nat = max_nat; rp[h].pair_cap = 0; rp[h].V = nullptr; }
  return 0;
}

inline void rep_pool_free(RepPool &r) {
  if (r.pair_cap <= 0) return;
  cudaFree(r.V); cudaFree(r.posB); cudaFree(r.res); cudaFree(r.pair_frame); cudaFree(r.pair_entry); cudaFree(r.pair_type);
  r.V = nullptr; r.pair_cap = 0;
}

// make both pools hold at least `pairs` pairs (clamped to the plan's limit); the caller has synchronised the streams
inline int rep_reserve(RepPool rp[2], const RepLimits &lim, size_t pairs, std::string &err) {
  if (pairs > lim.max_pairs) pairs = lim.max_pairs;
  if (pairs < 64) pairs = lim.max_pairs < 64 ? lim.max_pairs : 64;
  if ((size_t)rp[0].pair_cap >= pairs) return 0;
  for (int h = 0; h < 2; ++h) {
    rep_pool_free(rp[h]);
    bool ok = cudaMalloc(&rp[h].V, pairs * rp[h].sV * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&rp[h].posB, pairs * rp[h].max_nat * 3 * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&rp[h].res, pairs * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&rp[h].pair_frame, pairs * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMalloc(&rp[h].pair_entry, pairs * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMalloc(&rp[h].pair_type, pairs * sizeof(int)) == cudaSuccess;
    if (!ok) { err = "cudaMalloc(rep integral pool) failed"; return SLVGPU_ENOMEM; }
    rp[h].pair_cap = (int)pairs;
  }
  return 0;
}

// shared memory of k_rep_pair: X tile [<= REP_THREADS rows x 5 LMOs][4] + CB[max_nmos][max_nbas] + IJ[4][nmosA][max_nmos]
// + geometry
inline size_t rep_smem_bytes(int nmosc, int nbasc, int max_nmos, int max_nbas, int natc, int max_nat) {
  (void)nbasc;
  return sizeof(double) * ((size_t)REP_THREADS * 4 * 5 + 32 + (size_t)max_nmos * max_nbas + (size_t)4 * nmosc * max_nmos
                           + 6 * natc + 6 * nmosc + 3 * max_nat + 3 * max_nmos + 8);
}

// exclusive scan of the number of fragments inside ecut (k_frame_elect wrote it to the result row) over the frames
// of a chunk: pair0[f] = first pair slot of frame f, pair0[nf] = total
__global__ void __launch_bounds__(1024) k_rep_scan(const double *__restrict__ out, int frame0, int nf, int *__restrict__ pair0) {
  __shared__ int s_part[1024];
  const int tid = threadIdx.x;
  const int per = (nf + 1023) / 1024;
  const int a = tid * per, b = min(nf, a + per);
  int s = 0;
  for (int f = a; f < b; ++f) s += (int)out[(size_t)(frame0 + f) * SLVGPU_NOUT + SLVGPU_N_REP];
  s_part[tid] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {                     // Hillis-Steele inclusive scan of the 1024 partial sums
    const int v = tid >= o ? s_part[tid - o] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int run = s_part[tid] - s;
  for (int f = a; f < b; ++f) { pair0[f] = run; run += (int)out[(size_t)(frame0 + f) * SLVGPU_NOUT + SLVGPU_N_REP]; }
  if (tid == 1023) pair0[nf] = s_part[1023];
}

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

// one CTA per frame of the batch [fs, fe): lab-frame solute tables and the pair map
__global__ void __launch_bounds__(128) k_rep_prep(const DevPlan p, int fs, FrameLists fl, RepPool rp) {
  __shared__ int s_cnt[4], s_base;
  const int fi = fs + blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
  const int *tm0 = tmeta(p, p.inst_type[0]);
  const int natA = tm0[SLVGPU_M_NATOMS];
  {
    const double *sx = fl.sol_xf + (size_t)fi * 12;
    double R0[9], t0[3];
    for (int d = 0; d < 9; ++d) R0[d] = sx[d];
    for (int d = 0; d < 3; ++d) t0[d] = sx[9 + d];
    for (int a = tid; a < natA; a += NT) {
      double y[3];
      rot_vec(R0, p.dtab + tm0[SLVGPU_M_O_POS] + a * 3, y);
      double *pa = rp.posA + ((size_t)fi * natA + a) * 3;
      pa[0] = y[0] + t0[0]; pa[1] = y[1] + t0[1]; pa[2] = y[2] + t0[2];
      rot_vec(R0, p.dtab + p.sol_meta[SLVGPU_S_LVEC] + a * 3, rp.LcA + ((size_t)fi * natA + a) * 3);
    }
    const int nmosA = tm0[SLVGPU_M_NMOS], nbsA = tm0[SLVGPU_M_NBASIS];
    double *ri = rp.riA + (size_t)fi * nmosA * 6;
    for (int i = tid; i < nmosA; i += NT) {
      double y[3];
      rot_vec(R0, p.dtab + tm0[SLVGPU_M_O_LMOC] + i * 3, y);
      ri[i * 3] = y[0] + t0[0]; ri[i * 3 + 1] = y[1] + t0[1]; ri[i * 3 + 2] = y[2] + t0[2];
      rot_vec(R0, p.dtab + p.sol_meta[SLVGPU_S_LMOC1] + i * 3, ri + (size_t)nmosA * 3 + i * 3);
    }
    double *ca = rp.CA + (size_t)fi * 2 * nmosA * nbsA;
    rotate_vecl_rows(p, tm0, p.dtab + tm0[SLVGPU_M_O_VECL], nmosA, R0, ca);
    rotate_vecl_rows(p, tm0, p.dtab + p.sol_meta[SLVGPU_S_VECL1], nmosA, R0, ca + (size_t)nmosA * nbsA);
  }
  const int nl = fl.nlist[fi];
  const size_t lbase = (size_t)fi * p.list_cap;
  if (tid == 0) s_base = rp.pair0[fi] - rp.pair0[fs];
  __syncthreads();
  for (int e0 = 0; e0 < nl; e0 += NT) {                    // ordered compaction of the entries inside ecut
    const int e = e0 + tid;
    const bool keep = e < nl && (fl.list_flag[lbase + e] & 2);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if ((tid & 31) == 0) s_cnt[tid >> 5] = __popc(bal);
    __syncthreads();
    int slot = s_base;
    for (int w = 0; w < (tid >> 5); ++w) slot += s_cnt[w];
    slot += __popc(bal & ((1u << (tid & 31)) - 1u));
    if (keep && slot < rp.pair_cap) {
      const int ty = p.inst_type[fl.list_inst[lbase + e]];
      const int *tm = tmeta(p, ty);
      rp.pair_frame[slot] = fi; rp.pair_entry[slot] = e; rp.pair_type[slot] = ty;
      const double *xf = fl.list_xf + (lbase + e) * 12;
      double R[9], t[3];
      for (int d = 0; d < 9; ++d) R[d] = xf[d];
      for (int d = 0; d < 3; ++d) t[d] = xf[9 + d];
      double *pb = rp.posB + (size_t)slot * rp.max_nat * 3;
      for (int a = 0; a < tm[SLVGPU_M_NATOMS]; ++a) {
        double y[3];
        rot_vec(R, p.dtab + tm[SLVGPU_M_O_POS] + a * 3, y);
        pb[a * 3] = y[0] + t[0]; pb[a * 3 + 1] = y[1] + t[1]; pb[a * 3 + 2] = y[2] + t[2];
      }
    }
    __syncthreads();
    if (tid == 0) { int tot = 0; for (int w = 0; w < (NT >> 5); ++w) tot += s_cnt[w]; s_base += tot; }
    __syncthreads();
  }
}

// AO integrals of one (L_bra, L_ket, bra component) class for the pairs of fragment type `type`: thread t handles
// pair t / n_items, item lo + t % n_items (items sorted by primitive-pair count, so neighbouring lanes run equal loops)
template <int LA, int LB, int IA>
__global__ void __launch_bounds__(REP_INT_THREADS)
k_rep_ints(const DevPlan p, int type, int lo, int n_items, int npairs, RepPool rp) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int pr = (int)(t / n_items);
  if (pr >= npairs || rp.pair_type[pr] != type) return;
  const int it = (int)(t - (long)pr * n_items);
  const int *tm0 = tmeta(p, p.inst_type[0]), *tm = tmeta(p, type);
  const int nsubB = tm[SLVGPU_M_NSUBSH], nbsB = tm[SLVGPU_M_NBASIS], natA = tm0[SLVGPU_M_NATOMS];
  const int w = p.itab[tm[SLVGPU_M_IO_SHPAIR] + SUBSHELL_NCLASS + 1 + lo + it];
  const int a = w / nsubB, b = w - a * nsubB;
  const size_t fa = (size_t)rp.pair_frame[pr] * natA * 3;
  subshell_pair_store<LA, LB, IA>(p.itab + tm0[SLVGPU_M_IO_SUBSH] + a * 4, p.itab + tm[SLVGPU_M_IO_SUBSH] + b * 4,
                                  rp.posA + fa, rp.posB + (size_t)pr * rp.max_nat * 3,
                                  p.dtab + tm0[SLVGPU_M_O_BFEXP], p.dtab + tm0[SLVGPU_M_O_BFCOEF],
                                  p.dtab + tm[SLVGPU_M_O_BFEXP], p.dtab + tm[SLVGPU_M_O_BFCOEF],
                                  rp.LcA + fa, nbsB, rp.V + (size_t)pr * rp.sV);
}

// One CTA per (frame, fragment) pair: half transform, LMO integrals, CALCFI -> res[pair]
#ifndef SLV_REP_PAIR_CTAS
#define SLV_REP_PAIR_CTAS 4
#endif
__global__ void __launch_bounds__(REP_THREADS, SLV_REP_PAIR_CTAS)
k_rep_pair(const DevPlan p, FrameLists fl, RepPool rp) {
  extern __shared__ double sm[];
  __shared__ double s_red[32];
  const int tid = threadIdx.x, NT = blockDim.x;
  const int pr = blockIdx.x;
  const int fi = rp.pair_frame[pr], e = rp.pair_entry[pr];
  const int *tm0 = tmeta(p, p.inst_type[0]), *tm = tmeta(p, rp.pair_type[pr]);
  const int natA = tm0[SLVGPU_M_NATOMS], nmosA = tm0[SLVGPU_M_NMOS], nbsA = tm0[SLVGPU_M_NBASIS];
  const int natB = tm[SLVGPU_M_NATOMS], nmosB = tm[SLVGPU_M_NMOS], nbsB = tm[SLVGPU_M_NBASIS];
  const int nij = nmosA * nmosB;
  double *X = sm, *CBs = X + (size_t)REP_THREADS * 4 * 5 + 32, *IJ = CBs + (size_t)nmosB * nbsB;
  double *posA = IJ + 4 * nij, *LcA = posA + 3 * natA, *riA = LcA + 3 * natA, *ri1A = riA + 3 * nmosA;
  double *posB = ri1A + 3 * nmosA, *riB = posB + 3 * natB;
  const double *V = rp.V + (size_t)pr * rp.sV;
  const double *CA = rp.CA + (size_t)fi * 2 * nmosA * nbsA, *CA1 = CA + (size_t)nmosA * nbsA;
  const double *FA = p.dtab + tm0[SLVGPU_M_O_FOCK], *FA1 = p.dtab + p.sol_meta[SLVGPU_S_FOCK1], *ZA = p.dtab + tm0[SLVGPU_M_O_ZNUC];
  // The pair's integrals were written by earlier kernels of the batch into a pool far larger than L2: rows are pulled
  // into L2 two tiles ahead of the half transform (all of them at once would be 0.55 MB per DMSO pair x 3 CTAs x 148
  // SMs -- more than L2 holds).
  auto prefetch_rows = [&](int r0, int r1) {
    r1 = min(r1, nbsA);
    if (r0 >= r1) return;
    const char *base = reinterpret_cast<const char *>(V) + (size_t)r0 * nbsB * 32;
    const size_t bytes = (size_t)(r1 - r0) * nbsB * 32;
    for (size_t b = (size_t)tid * 128; b < bytes; b += (size_t)NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + b));
  };
  const int nblk0 = (nmosB + 4) / 5, KT0 = max(1, NT / nblk0);
  prefetch_rows(0, 2 * KT0);
  double R[9], t[3];
  {
    const double *xf = fl.list_xf + ((size_t)fi * p.list_cap + e) * 12;
    for (int d = 0; d < 9; ++d) R[d] = xf[d];
    for (int d = 0; d < 3; ++d) t[d] = xf[9 + d];
  }
  for (int w = tid; w < natA * 3; w += NT) { posA[w] = rp.posA[(size_t)fi * natA * 3 + w]; LcA[w] = rp.LcA[(size_t)fi * natA * 3 + w]; }
  for (int w = tid; w < nmosA * 6; w += NT) riA[w] = rp.riA[(size_t)fi * nmosA * 6 + w];        // riA then ri1A, contiguous
  for (int w = tid; w < natB * 3; w += NT) posB[w] = rp.posB[(size_t)pr * rp.max_nat * 3 + w];
  for (int j = tid; j < nmosB; j += NT) {
    double y[3];
    rot_vec(R, p.dtab + tm[SLVGPU_M_O_LMOC] + j * 3, y);
    riB[j * 3] = y[0] + t[0]; riB[j * 3 + 1] = y[1] + t[1]; riB[j * 3 + 2] = y[2] + t[2];
  }
  const int *blB = p.itab + tm[SLVGPU_M_IO_BFL];
  const double *vecB = p.dtab + tm[SLVGPU_M_O_VECL];
  // rotate the fragment's AO->LMO rows into the lab frame (VECROT): s copy, p as vectors, d as symmetric tensors; one
  // thread per leading element of a block
  for (int w = tid; w < nmosB * nbsB; w += NT) {
    const int j = w / nbsB, k = w - j * nbsB;
    const int lx = blB[k * 3], ly = blB[k * 3 + 1], lz = blB[k * 3 + 2], Ls = lx + ly + lz;
    const double *c = vecB + (size_t)j * nbsB + k;
    double *d = CBs + (size_t)j * nbsB + k;
    if (Ls == 0) d[0] = c[0];
    else if (Ls == 1 && lx == 1) { double v[3] = {c[0], c[1], c[2]}, o[3]; rot_vec(R, v, o); d[0] = o[0]; d[1] = o[1]; d[2] = o[2]; }
    else if (Ls == 2 && lx == 2) { double v[6] = {c[0], c[1], c[2], c[3], c[4], c[5]}, o[6]; rot_quad(R, v, o); for (int q = 0; q < 6; ++q) d[q] = o[q]; }
  }
  for (int w = tid; w < 4 * nij; w += NT) IJ[w] = 0.0;
  // The solute's basis functions are taken in tiles of KT rows sized so that one tile's half transform is one round of
  // the CTA; a tile's X goes straight into the LMO integrals, so every integral record is read exactly once however
  // many LMOs the fragment has (three passes over V for DMSO's 21 LMOs cost 1.1 MB of DRAM reads per pair before).
  constexpr int JB = 5;
  const int nblk = (nmosB + JB - 1) / JB, JP = nblk * JB;
  const int KT = max(1, NT / nblk);
  for (int k0 = 0; k0 < nbsA; k0 += KT) {
    const int kn = min(KT, nbsA - k0);
    prefetch_rows(k0 + 2 * KT, k0 + 3 * KT);
    __syncthreads();                                         // CBs ready / the previous tile's X is consumed
    // ---- phase 2: X[kk][q][j] = sum_l V[k0+kk][l][q] CB[j][l]; one thread per (row, block of 5 LMOs): the row is
    //      streamed as 32-byte records, each feeding 20 FMAs (the nblk threads of a row sit in one warp)
    for (int w = tid; w < kn * nblk; w += NT) {
      const int kk = w / nblk, jb = (w - kk * nblk) * JB;
      const double2 *v2 = reinterpret_cast<const double2 *>(V + (size_t)(k0 + kk) * nbsB * 4);
      const double *cb[JB];
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) cb[jj] = CBs + (size_t)min(jb + jj, nmosB - 1) * nbsB;
      double acc[4][JB];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) acc[q][jj] = 0.0;
#pragma unroll 8
      for (int l = 0; l < nbsB; ++l) {
        const double2 va = v2[2 * l], vb = v2[2 * l + 1];
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
          const double c = cb[jj][l];
          acc[0][jj] = fma(va.x, c, acc[0][jj]); acc[1][jj] = fma(va.y, c, acc[1][jj]);
          acc[2][jj] = fma(vb.x, c, acc[2][jj]); acc[3][jj] = fma(vb.y, c, acc[3][jj]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) X[((size_t)kk * 4 + q) * JP + jb + jj] = acc[q][jj];
    }
    __syncthreads();
    // ---- phase 3: this tile's contribution to the LMO-basis integrals; (i, j) is owned by one thread
    for (int w = tid; w < nij; w += NT) {
      const int i = w / nmosB, j = w - i * nmosB;
      const double *ca = CA + (size_t)i * nbsA + k0, *ca1 = CA1 + (size_t)i * nbsA + k0;
      double sv = 0, tv = 0, smv = 0, tmv = 0;
#pragma unroll 8
      for (int kk = 0; kk < kn; ++kk) {
        const double *x = X + (size_t)kk * 4 * JP + j;
        const double xs = x[0], xt = x[JP], xds = x[2 * JP], xdt = x[3 * JP];
        const double c0 = ca[kk], c1 = ca1[kk];
        sv = fma(c0, xs, sv); tv = fma(c0, xt, tv);
        smv = fma(c0, xds, fma(c1, xs, smv)); tmv = fma(c0, xdt, fma(c1, xt, tmv));
      }
      IJ[w] += sv; IJ[nij + w] += tv; IJ[2 * nij + w] += smv; IJ[3 * nij + w] += tmv;
    }
  }
  __syncthreads();
  // ---- phase 4: CALCFI
  double acc = 0.0;
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
  acc = block_sum(acc, s_red);
  if (tid == 0) rp.res[pr] = acc;
}

// rep_mea of every frame of the batch: the pair results summed in list order (deterministic)
__global__ void k_rep_sum(int frame0, int fs, int nb, RepPool rp, double *__restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nb) return;
  const int a = rp.pair0[fs + f] - rp.pair0[fs], b = rp.pair0[fs + f + 1] - rp.pair0[fs];
  double s = 0.0;
  for (int q = a; q < b; ++q) s += rp.res[q];
  out[(size_t)(frame0 + fs + f) * SLVGPU_NOUT + SLVGPU_REP_MEA] = -s;
}

}  // namespace slv
