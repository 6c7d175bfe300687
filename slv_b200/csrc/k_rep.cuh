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
//   k_rep_prep   per frame: lab-frame solute atoms and displacement vectors, the solute's rotated AO->LMO coefficients
//                in k-octet tiles; per (frame, fragment) pair: the fragment's atoms and the pair -> (frame, entry, type) map
//   k_rep_ints   AO integrals V[pair][k][l][q] (q = S, T, L.dS, L.dT; l padded to a multiple of 4): one launch per (fragment
//                type, sub-shell class); a thread owns one (sub-shell pair) item of the class and walks 8 or 16 pairs of the
//                type's pair list (geometry staged per warp with cp.async, primitive pairs from the plan's sorted records,
//                one 256-bit store per record) -- every warp runs one specialisation of slv_ints.cuh with near-equal loops
//   k_rep_pair   one CTA per (frame, fragment) pair, one launch per fragment type.  The pair's integrals stream through
//                shared memory in tiles of KT solute basis functions (one cp.async.bulk per tile, completion on an
//                mbarrier, the next tile in flight while the current one is consumed); per tile
//                  A  X[(k,q)][j] = sum_l V[(k,q)][l] C_B[j][l]        FP64 tensor cores (mma.sync m8n8k4), operands in smem
//                  B  S_ij, T_ij, Sm_ij, Tm_ij += C_A / C1_A [i][k] X[k][q][j]   FP64 tensor cores, accumulators in registers
//                then CALCFI from distance tables (every centroid / atom distance evaluated once) -> res[pair]
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
constexpr int REP_INT_SREC = PP_SREC;   // primitive-pair records per item the integral kernel keeps in shared memory
// doubles of one geometry buffer of the integral kernel (solute atoms, displacement vectors, fragment atoms), even
__host__ __device__ inline int gws_of(int natA, int natB) { return (natA * 6 + natB * 3 + 1) & ~1; }
// dynamic shared memory of the integral kernel: two geometry buffers per warp, REP_INT_SREC records per thread
__host__ __device__ inline size_t rep_ints_smem(int threads, int natA, int max_natB) {
  return ((size_t)(threads / 32) * 2 * gws_of(natA, max_natB) + (size_t)threads * REP_INT_SREC * 6) * sizeof(double);
}
#ifndef SLV_REP_INT_PBLOCKS
#define SLV_REP_INT_PBLOCKS 4     // resident blocks per SM the p-shell classes are compiled for
#endif
#ifndef SLV_REP_INT_DBLOCKS
#define SLV_REP_INT_DBLOCKS 3     // resident blocks per SM the d-shell classes of the integral kernel are compiled for
#endif
constexpr int REP_CS = 12;        // row stride (doubles) of a solute coefficient octet tile: 8 k's + 4 pad, == 4 (mod 8)
constexpr int REP_MAX_NIT = 3;    // solute LMO tiles of 8 (nmos <= 24)
constexpr int REP_IP = 8 * REP_MAX_NIT;   // rows of a solute coefficient tile (fixed: strides become compile-time constants)
constexpr int REP_MAX_NNT = 5;    // fragment LMO tiles of 8 (nmos <= 40)

// Records per row of the integral block of a fragment with nbs basis functions: V[k][l][q], l padded to the k-step of the
// tensor-core MMA (4).  A fragment of eight (k, q) rows x four l reads 2 x 16 consecutive doubles: no bank conflicts.
__host__ __device__ inline int rep_lp(int nbs) { return (nbs + 3) & ~3; }

struct RepPool {            // one batch of (frame, fragment) pairs (V .. pair_type) + per-frame tables of the chunk
  double *V;       // [pair_cap][nbsA16][LP(type)][4]  AO integrals (S, T, L.dS, L.dT per record), block stride sV
  double *posA;    // [max_chunk][2][natA][3]  per frame: solute atoms in the lab frame, then the mode-contracted displacement
                   //                          vectors (one contiguous record: the integral kernels stage it in one piece)
  double *riA;     // [max_chunk][2][nmosA][3]  LMO centroids and their mode-contracted derivatives
  double *CA;      // [max_chunk][nbsA16/8][2][IP][REP_CS]  rotated vecl / g-contracted vecl1 of the solute in k-octet tiles
  double *posB;    // [pair_cap][max_nat][3]
  double *res;     // [pair_cap]  sum_ij fi of the pair
  int *pair_frame, *pair_entry, *pair_type;   // [pair_cap]
  int *pair0;      // [max_chunk + 1] exclusive scan of the per-frame pair counts
  int *tlist;      // [ntypes][pair_cap]  pair slots of every fragment type (filled by k_rep_prep, any order)
  int *tcount;     // [ntypes]            their number (zeroed before k_rep_prep)
  size_t sV, sCA;
  int pair_cap, max_nat, nbsA16, IP, ntypes;
};

// Two pools (the batches alternate between two streams, so the integral kernels of one batch overlap the
// transform kernel of the previous one); the per-frame tables are shared.  rep_alloc creates the
// per-frame tables and fixes the largest pool the plan may use; the pools themselves grow on demand (rep_reserve),
// so a plan that only ever sees a few fragments inside ecut does not hold gigabytes.
struct RepLimits { size_t max_pairs = 0; };

inline int rep_alloc(RepPool rp[2], RepLimits &lim, std::vector<void *> &allocs, std::string &err, int nmosc, int nbasc,
                     int natc, int max_nbas, int max_nat, int max_chunk, int list_cap, int ntypes) {
  // integral pools: together up to 24 GiB (SLVGPU_REP_POOL_GIB) and at most a quarter of the free memory, never more than the worst case
  const int nbsA16 = (nbasc + 15) & ~15, IP = REP_IP;
  (void)nmosc;
  const size_t sV = (size_t)nbsA16 * 4 * rep_lp(max_nbas);
  const size_t sCA = (size_t)(nbsA16 / 8) * 2 * IP * REP_CS;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  size_t budget = (size_t)24 << 30;
  if (const char *ev = getenv("SLVGPU_REP_POOL_GIB")) { const long v = atol(ev); if (v > 0 && v <= 128) budget = (size_t)v << 30; }
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
  RA(rp[0].posA, (size_t)max_chunk * natc * 6, double)
  RA(rp[0].riA, (size_t)max_chunk * nmosc * 6, double)
  RA(rp[0].CA, (size_t)max_chunk * sCA, double)
  RA(rp[0].pair0, (size_t)max_chunk + 1, int)
#undef RA
  // the pad rows / columns of the coefficient tiles are never written again: they must be zero
  if (cudaMemset(rp[0].CA, 0, (size_t)max_chunk * sCA * sizeof(double)) != cudaSuccess) { err = "cudaMemset(rep tables) failed"; return SLVGPU_ECUDA; }
  rp[1] = rp[0];
  for (int h = 0; h < 2; ++h) {
    rp[h].sV = sV; rp[h].sCA = sCA; rp[h].max_nat = max_nat; rp[h].pair_cap = 0; rp[h].V = nullptr; rp[h].nbsA16 = nbsA16; rp[h].IP = IP;
    rp[h].posB = nullptr; rp[h].res = nullptr; rp[h].pair_frame = rp[h].pair_entry = rp[h].pair_type = nullptr;
    rp[h].tlist = rp[h].tcount = nullptr; rp[h].ntypes = ntypes;
  }
  return 0;
}

inline void rep_pool_free(RepPool &r) {
  cudaFree(r.V); cudaFree(r.posB); cudaFree(r.res); cudaFree(r.pair_frame); cudaFree(r.pair_entry); cudaFree(r.pair_type);   // cudaFree(nullptr) is a no-op
  cudaFree(r.tlist); cudaFree(r.tcount);
  r.V = nullptr; r.posB = nullptr; r.res = nullptr; r.pair_frame = r.pair_entry = r.pair_type = nullptr;
  r.tlist = r.tcount = nullptr;
  r.pair_cap = 0;
}

// make both pools hold at least `pairs` pairs (clamped to the plan's limit); the caller has synchronised the streams
inline int rep_reserve(RepPool rp[2], const RepLimits &lim, size_t pairs, std::string &err) {
  if (pairs > lim.max_pairs) pairs = lim.max_pairs;
  if (pairs < 64) pairs = lim.max_pairs < 64 ? lim.max_pairs : 64;
  if ((size_t)rp[0].pair_cap >= pairs) return 0;
  for (int h = 0; h < 2; ++h) {
    rep_pool_free(rp[h]);
    bool ok = cudaMalloc(&rp[h].V, pairs * rp[h].sV * sizeof(double)) == cudaSuccess;
    // pad columns of the integral rows and the rows past the solute's last basis function are read by the tile loop
    // (times zero coefficients): stale finite values are harmless, uninitialised NaN patterns are not
    ok = ok && cudaMemset(rp[h].V, 0, pairs * rp[h].sV * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&rp[h].posB, pairs * rp[h].max_nat * 3 * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&rp[h].res, pairs * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&rp[h].pair_frame, pairs * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMalloc(&rp[h].pair_entry, pairs * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMalloc(&rp[h].pair_type, pairs * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMalloc(&rp[h].tlist, pairs * (size_t)rp[h].ntypes * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMalloc(&rp[h].tcount, (size_t)rp[h].ntypes * sizeof(int)) == cudaSuccess;
    // the memsets ran on the legacy default stream, which the plan's non-blocking streams do not wait for
    ok = ok && cudaDeviceSynchronize() == cudaSuccess;
    if (!ok) {                                     // leave no half-built pool behind (a retry starts from scratch)
      cudaGetLastError();
      for (int g = 0; g < 2; ++g) rep_pool_free(rp[g]);
      err = "cudaMalloc(rep integral pool) failed";
      return SLVGPU_ENOMEM;
    }
    rp[h].pair_cap = (int)pairs;
  }
  return 0;
}

// ---- shared-memory plan of k_rep_pair for one fragment type
struct RepSmem {
  int KT, NNT, nbuf, LP, LK, JS, IP, nit;
  int oV, oCA, oX, oCB, oIJ, oTab, oGeo, oBar;     // offsets in doubles
  size_t bytes;
};

inline RepSmem rep_smem_plan(int nmosA, int nbsA, int natA, int nmosB, int nbsB, int natB) {
  RepSmem s;
  (void)nbsA;
  s.LP = rep_lp(nbsB); s.LK = (nbsB + 3) & ~3;
  s.NNT = (nmosB + 7) / 8; s.JS = 8 * s.NNT + 4;
  s.IP = REP_IP; s.nit = (nmosA + 7) / 8;
  s.KT = (size_t)16 * 4 * s.LP * 8 <= 24 * 1024 ? 16 : 8;                     // tiles of 16 solute functions while they stay small
  const size_t tileV = (size_t)s.KT * 4 * s.LP, tileCA = (size_t)(s.KT / 8) * 2 * s.IP * REP_CS, tileX = (size_t)4 * s.KT * s.JS;
  const int nij = nmosA * nmosB;
  const size_t fixed = 2 * tileX + (size_t)s.LK * s.JS + (size_t)4 * nij + (size_t)2 * nij + 4 * nmosB + 4 * nmosA
                     + 6 * natA + 6 * nmosA + 3 * natB + 3 * nmosB + 8;
  s.nbuf = (fixed + 2 * tileV + 3 * tileCA) * 8 <= (size_t)110 * 1024 ? 2 : 1;  // two tiles in flight when two CTAs still fit an SM
  int o = 0;
  s.oV = o; o += (int)(s.nbuf * tileV);
  s.oCA = o; o += (int)((s.nbuf + 1) * tileCA);      // one coefficient buffer more than tiles in flight: phase B still reads the oldest
  s.oX = o; o += (int)(2 * tileX);
  s.oCB = o; o += s.LK * s.JS;
  s.oIJ = o; o += 4 * nij;
  s.oTab = o; o += 2 * nij + 4 * nmosB + 4 * nmosA;
  s.oGeo = o; o += 6 * natA + 6 * nmosA + 3 * natB + 3 * nmosB;
  o = (o + 1) & ~1;
  s.oBar = o; o += 4;
  s.bytes = (size_t)o * 8;
  return s;
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

// rotate the AO->LMO rows of one fragment type into the lab frame (VECROT): s copy, p triples as vectors, d sextets as
// symmetric tensors; one thread per leading element of a block; element (i, k) goes to dst[index(i, k)]
template <class Index>
__device__ inline void rotate_vecl_rows(const DevPlan &p, const int *tm, const double *src, int nrows, const double R[9], double *dst, Index index) {
  const int nb = tm[SLVGPU_M_NBASIS], nblk = tm[SLVGPU_M_NBFBLK];
  const int2 *blk = reinterpret_cast<const int2 *>(p.itab + tm[SLVGPU_M_IO_BFBLK]);     // (first function, L) of every block
  for (int w = threadIdx.x; w < nrows * nblk; w += blockDim.x) {
    const int i = w / nblk;
    const int2 bk = blk[w - i * nblk];
    const int k = bk.x;
    const double *c = src + (size_t)i * nb + k;
    if (bk.y == 0) dst[index(i, k)] = c[0];
    else if (bk.y == 1) { double v[3] = {c[0], c[1], c[2]}, o[3]; rot_vec(R, v, o); for (int t = 0; t < 3; ++t) dst[index(i, k + t)] = o[t]; }
    else { double v[6] = {c[0], c[1], c[2], c[3], c[4], c[5]}, o[6]; rot_quad(R, v, o); for (int t = 0; t < 6; ++t) dst[index(i, k + t)] = o[t]; }
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
      double *pa = rp.posA + ((size_t)fi * 2 * natA + a) * 3;
      pa[0] = y[0] + t0[0]; pa[1] = y[1] + t0[1]; pa[2] = y[2] + t0[2];
      rot_vec(R0, p.dtab + p.sol_meta[SLVGPU_S_LVEC] + a * 3, rp.posA + ((size_t)(fi * 2 + 1) * natA + a) * 3);
    }
    const int nmosA = tm0[SLVGPU_M_NMOS];
    double *ri = rp.riA + (size_t)fi * nmosA * 6;
    for (int i = tid; i < nmosA; i += NT) {
      double y[3];
      rot_vec(R0, p.dtab + tm0[SLVGPU_M_O_LMOC] + i * 3, y);
      ri[i * 3] = y[0] + t0[0]; ri[i * 3 + 1] = y[1] + t0[1]; ri[i * 3 + 2] = y[2] + t0[2];
      rot_vec(R0, p.dtab + p.sol_meta[SLVGPU_S_LMOC1] + i * 3, ri + (size_t)nmosA * 3 + i * 3);
    }
    // k-octet tiles [k / 8][vecl | vecl1][IP][REP_CS]: one contiguous block per tile of the transform kernel
    double *ca = rp.CA + (size_t)fi * rp.sCA;
    const int IP = rp.IP;
    rotate_vecl_rows(p, tm0, p.dtab + tm0[SLVGPU_M_O_VECL], nmosA, R0, ca,
                     [IP](int i, int k) { return ((size_t)(k >> 3) * 2 * IP + i) * REP_CS + (k & 7); });
    rotate_vecl_rows(p, tm0, p.dtab + p.sol_meta[SLVGPU_S_VECL1], nmosA, R0, ca,
                     [IP](int i, int k) { return (((size_t)(k >> 3) * 2 + 1) * IP + i) * REP_CS + (k & 7); });
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
      rp.tlist[(size_t)ty * rp.pair_cap + atomicAdd(rp.tcount + ty, 1)] = slot;   // order within a type is irrelevant: every pair is independent
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

// ---- small PTX wrappers: FP64 tensor-core MMA, mbarrier, 1-D bulk asynchronous copy (TMA)
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {   // 8 bytes global -> shared, no register staging
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// AO integrals of one (L_bra, L_ket, bra component) class for the pairs of fragment type `type`.  A thread owns ONE item
// of the class (a sub-shell pair; items sorted by primitive-pair count, so neighbouring lanes run equal loops) and walks
// the block's share of the type's pairs (blockIdx.y, `ppb` <= 32 pairs): everything that does not depend on the geometry
// -- the item record, its primitive-pair table, the V offsets -- is fetched once, and the geometry of
// the next pair (solute atoms and displacement vectors of its frame, the fragment's atoms) streams into the warp's
// shared-memory buffer with cp.async while the current pair is evaluated.  (One thread per (pair, item) paid the chain of
// five dependent global loads for ~500 instructions of arithmetic.)
template <int LA, int LB, int IA>
__global__ void __launch_bounds__(REP_INT_THREADS, (LB == 2 || (LA == 2 && LB == 0)) ? SLV_REP_INT_DBLOCKS : (LA + LB == 0 ? 10 : SLV_REP_INT_PBLOCKS))
k_rep_ints(const DevPlan p, int type, int lo, int n_items, int ppb, RepPool rp) {
  extern __shared__ double s_geo[];                        // [warps][2][gws]
  constexpr int NA = IA < 0 ? cart_ncomp(LA) : 1, NB = cart_ncomp(LB);
  const int cnt = rp.tcount[type];
  const int q0 = blockIdx.y * ppb;
  if (q0 >= cnt) return;
  const int nq = min(cnt - q0, ppb);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int *tm0 = tmeta(p, p.inst_type[0]), *tm = tmeta(p, type);
  const int natA = tm0[SLVGPU_M_NATOMS], natB = tm[SLVGPU_M_NATOMS];
  const int LP = rep_lp(tm[SLVGPU_M_NBASIS]);
  const int it = blockIdx.x * blockDim.x + tid;
  const bool act = it < n_items;
  // item record: 3 * atom_A, first function A, 3 * atom_B, first function B, first primitive-pair record, records
  const int4 ir = *reinterpret_cast<const int4 *>(p.itab + tm[SLVGPU_M_IO_SHITEM] + (size_t)(lo + (act ? it : n_items - 1)) * 8);
  const int2 ip = *reinterpret_cast<const int2 *>(p.itab + tm[SLVGPU_M_IO_SHITEM] + (size_t)(lo + (act ? it : n_items - 1)) * 8 + 4);
  const int atA = ir.x, f0a = ir.y, atB = ir.z, f0b = ir.w, npp = ip.y;
  const double *rec = p.dtab + tm[SLVGPU_M_O_PPTAB] + (size_t)ip.x * PP_REC;
  // the first REP_INT_SREC records of the item stay in shared memory for all its pairs ([record][field][thread]: lanes
  // read consecutive words); the deeper ones -- tight primitives, reached only at short distances -- come from the table
  double *srec = s_geo + (size_t)(blockDim.x >> 5) * 2 * gws_of(natA, natB) + tid;
  {
    const int n = min(npp, REP_INT_SREC);
    for (int k = 0; k < n; ++k)
#pragma unroll
      for (int d = 0; d < PP_REC; ++d) srec[(k * PP_REC + d) * blockDim.x] = rec[k * PP_REC + d];
  }
  const int nA6 = natA * 6, gw = nA6 + natB * 3, gws = gws_of(natA, natB);
  double *geo = s_geo + (size_t)warp * 2 * gws;
  // lane l keeps the slot and the frame of pair q0 + l
  const int my_slot = lane < nq ? rp.tlist[(size_t)type * rp.pair_cap + q0 + lane] : 0;
  const int my_frame = rp.pair_frame[my_slot];
  auto stage = [&](int i) {
    const int slot = __shfl_sync(0xffffffffu, my_slot, i), fr = __shfl_sync(0xffffffffu, my_frame, i);
    const double *gA = rp.posA + (size_t)fr * nA6, *gB = rp.posB + (size_t)slot * rp.max_nat * 3 - nA6;
    double *dst = geo + (i & 1) * gws;
    for (int x = lane; x < gw; x += 32) cp_async8(dst + x, (x < nA6 ? gA : gB) + x);
    cp_async_commit();
  };
  stage(0);
  for (int i = 0; i < nq; ++i) {
    cp_async_wait_all();
    __syncwarp();
    if (i + 1 < nq) stage(i + 1);                          // the other buffer: every lane finished pair i - 1 before the barrier above
    const int slot = __shfl_sync(0xffffffffu, my_slot, i);
    if (act) {
      const double *g = geo + (i & 1) * gws;
      const double A[3] = {g[atA], g[atA + 1], g[atA + 2]}, L[3] = {g[nA6 / 2 + atA], g[nA6 / 2 + atA + 1], g[nA6 / 2 + atA + 2]};
      const double B[3] = {g[nA6 + atB], g[nA6 + atB + 1], g[nA6 + atB + 2]};
      double acc[NA * NB * 4];
      subshell_pair_t<LA, LB, IA>(A, B, L, npp, rec, srec, (int)blockDim.x, acc);
      subshell_pair_write<LA, LB, IA>(acc, f0a, f0b, LP, rp.V + (size_t)slot * rp.sV);
    }
  }
}

// One CTA per (frame, fragment) pair of fragment type `type`: half transform and LMO integrals on the FP64 tensor cores,
// then CALCFI -> res[pair].  KT: solute basis functions per tile (8 or 16); NNT: fragment LMO tiles of 8.
template <int KT, int NNT>
__global__ void __launch_bounds__(REP_THREADS, NNT <= 2 ? 4 : 2)
k_rep_pair(const DevPlan p, int type, FrameLists fl, RepPool rp, RepSmem L) {
  extern __shared__ double sm[];
  __shared__ double s_red[32];
  const int tid = threadIdx.x, NT = REP_THREADS, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c4 = lane & 3;
  const int pr = blockIdx.x;
  if (rp.pair_type[pr] != type) return;
  const int fi = rp.pair_frame[pr], e = rp.pair_entry[pr];
  const int *tm0 = tmeta(p, p.inst_type[0]), *tm = tmeta(p, type);
  const int natA = tm0[SLVGPU_M_NATOMS], nmosA = tm0[SLVGPU_M_NMOS], nbsA = tm0[SLVGPU_M_NBASIS];
  const int natB = tm[SLVGPU_M_NATOMS], nmosB = tm[SLVGPU_M_NMOS], nbsB = tm[SLVGPU_M_NBASIS];
  const int nij = nmosA * nmosB;
  constexpr int JS = 8 * NNT + 4;                        // row stride of the CB / X tiles (== 4 mod 8: conflict-free B fragments)
  const int LP = L.LP, nit = L.nit;
  double *Vt = sm + L.oV, *CAs = sm + L.oCA, *Xs = sm + L.oX, *CBs = sm + L.oCB, *IJ = sm + L.oIJ;
  double *RI = sm + L.oTab, *RM = RI + nij, *col1 = RM + nij, *col2 = col1 + nmosB, *u3 = col2 + nmosB, *w3 = u3 + nmosB;
  double *row1 = w3 + nmosB, *row2 = row1 + nmosA, *u4 = row2 + nmosA, *w4 = u4 + nmosA;
  double *posA = sm + L.oGeo, *LcA = posA + 3 * natA, *riA = LcA + 3 * natA, *ri1A = riA + 3 * nmosA;
  double *posB = ri1A + 3 * nmosA, *riB = posB + 3 * natB;
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + L.oBar);
  const double *V = rp.V + (size_t)pr * rp.sV;
  const double *CAg = rp.CA + (size_t)fi * rp.sCA;
  const int ntile = (nbsA + KT - 1) / KT;
  const unsigned tileV_bytes = (unsigned)((size_t)KT * 4 * LP * sizeof(double));
  constexpr unsigned tileCA_doubles = (unsigned)((KT / 8) * 2 * REP_IP * REP_CS);
  const size_t tileV_doubles = (size_t)KT * 4 * LP, tileX = (size_t)4 * KT * JS;

  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  // tile t -> V buffer t % nbuf, coefficient buffer t % (nbuf + 1), both completing on barrier t & 1
  const int ncab = L.nbuf + 1;
  auto issue = [&](int t) {
    unsigned long long *b = bar + (t & 1);
    mbar_expect_tx(b, tileV_bytes + tileCA_doubles * (unsigned)sizeof(double));
    bulk_g2s(Vt + (size_t)(t % L.nbuf) * tileV_doubles, V + (size_t)t * tileV_doubles, tileV_bytes, b);
    bulk_g2s(CAs + (size_t)(t % ncab) * tileCA_doubles, CAg + (size_t)t * tileCA_doubles, tileCA_doubles * (unsigned)sizeof(double), b);
  };
  if (tid == 0) { issue(0); if (L.nbuf == 2 && ntile > 1) issue(1); }

  // ---- geometry, the fragment's rotated AO->LMO coefficients, and the distance tables of CALCFI (while tile 0 is in flight)
  double R[9], tr[3];
  {
    const double *xf = fl.list_xf + ((size_t)fi * p.list_cap + e) * 12;
    for (int d = 0; d < 9; ++d) R[d] = xf[d];
    for (int d = 0; d < 3; ++d) tr[d] = xf[9 + d];
  }
  for (int w = tid; w < natA * 6; w += NT) posA[w] = rp.posA[(size_t)fi * natA * 6 + w];                 // posA then LcA, contiguous
  for (int w = tid; w < nmosA * 6; w += NT) riA[w] = rp.riA[(size_t)fi * nmosA * 6 + w];        // riA then ri1A, contiguous
  for (int w = tid; w < natB * 3; w += NT) posB[w] = rp.posB[(size_t)pr * rp.max_nat * 3 + w];
  for (int j = tid; j < nmosB; j += NT) {
    double y[3];
    rot_vec(R, p.dtab + tm[SLVGPU_M_O_LMOC] + j * 3, y);
    riB[j * 3] = y[0] + tr[0]; riB[j * 3 + 1] = y[1] + tr[1]; riB[j * 3 + 2] = y[2] + tr[2];
  }
  // pad rows / columns of the coefficient tile are zero; the rotation writes every other element (disjoint: no barrier)
  for (int w = tid; w < L.LK * JS; w += NT) { const int l = w / JS, j = w - l * JS; if (l >= nbsB || j >= nmosB) CBs[w] = 0.0; }
  rotate_vecl_rows(p, tm, p.dtab + tm[SLVGPU_M_O_VECL], nmosB, R, CBs, [](int j, int l) { return l * JS + j; });   // CBs[l][j]
  __syncthreads();                                       // geometry tables complete
  for (int w = tid; w < nij; w += NT) {
    const int i = w / nmosB, j = w - i * nmosB;
    const double dx = riA[i * 3] - riB[j * 3], dy = riA[i * 3 + 1] - riB[j * 3 + 1], dz = riA[i * 3 + 2] - riB[j * 3 + 2];
    const double rinv = rsqrt64(dx * dx + dy * dy + dz * dz);
    RI[w] = rinv; RM[w] = rinv * (dx * ri1A[i * 3] + dy * ri1A[i * 3 + 1] + dz * ri1A[i * 3 + 2]);
  }
  __syncthreads();
  for (int w = tid; w < 2 * (nmosA + nmosB); w += NT) {
    if (w < nmosB) {                                     // column sums over the solute's LMOs (TA1)
      double s1 = 0, s2 = 0;
      for (int k = 0; k < nmosA; ++k) { const double r = RI[k * nmosB + w]; s1 += r; s2 += r * r * RM[k * nmosB + w]; }
      col1[w] = s1; col2[w] = s2;
    } else if (w < 2 * nmosB) {                          // solute atoms against fragment LMO j (TA3)
      const int j = w - nmosB;
      double su = 0, sw = 0;
      for (int m = 0; m < natA; ++m) {
        const double dx = posA[m * 3] - riB[j * 3], dy = posA[m * 3 + 1] - riB[j * 3 + 1], dz = posA[m * 3 + 2] - riB[j * 3 + 2];
        const double rinv = rsqrt64(dx * dx + dy * dy + dz * dz), z = p.dtab[tm0[SLVGPU_M_O_ZNUC] + m];
        su += z * rinv * rinv * rinv * (dx * LcA[m * 3] + dy * LcA[m * 3 + 1] + dz * LcA[m * 3 + 2]); sw += z * rinv;
      }
      u3[j] = su; w3[j] = sw;
    } else if (w < 2 * nmosB + nmosA) {                  // row sums over the fragment's LMOs (TA2)
      const int i = w - 2 * nmosB;
      double s1 = 0, s2 = 0;
      for (int l = 0; l < nmosB; ++l) { const double r = RI[i * nmosB + l]; s1 += r; s2 += r * r * RM[i * nmosB + l]; }
      row1[i] = s1; row2[i] = s2;
    } else {                                             // fragment atoms against solute LMO i (TA4)
      const int i = w - 2 * nmosB - nmosA;
      double su = 0, sw = 0;
      for (int n = 0; n < natB; ++n) {
        const double dx = riA[i * 3] - posB[n * 3], dy = riA[i * 3 + 1] - posB[n * 3 + 1], dz = riA[i * 3 + 2] - posB[n * 3 + 2];
        const double rinv = rsqrt64(dx * dx + dy * dy + dz * dz), z = p.dtab[tm[SLVGPU_M_O_ZNUC] + n];
        su += z * rinv * rinv * rinv * (dx * ri1A[i * 3] + dy * ri1A[i * 3 + 1] + dz * ri1A[i * 3 + 2]); sw += z * rinv;
      }
      u4[i] = su; w4[i] = sw;
    }
  }
  __syncthreads();

  // ---- tile loop.  Phase B units u = kind * nit + itile (kind: S, T, Sm, Tm), unit u belongs to warp u % 4.  Every
  //      offset below is a 32-bit count of doubles inside one shared-memory buffer, fixed before the loop.
  double accB[REP_MAX_NIT][NNT][2];
  int ukind[REP_MAX_NIT], uoffA[REP_MAX_NIT], uoffX0[REP_MAX_NIT], uoffX1[REP_MAX_NIT], urow[REP_MAX_NIT];
#pragma unroll
  for (int s = 0; s < REP_MAX_NIT; ++s) {
#pragma unroll
    for (int n = 0; n < NNT; ++n) accB[s][n][0] = accB[s][n][1] = 0.0;
    const int u = warp + 4 * s;
    const int kind = u < 4 * nit ? u / nit : -1, itile = u - kind * nit;
    ukind[s] = kind;
    urow[s] = itile * 8 + g;
    uoffA[s] = (itile * 8 + g) * REP_CS + c4;
    uoffX0[s] = (max(kind, 0) * KT + c4) * JS + g;
    uoffX1[s] = ((kind & 1) * KT + c4) * JS + g;           // Sm, Tm: the C1 product uses X_S, X_T
  }
  const int nks = L.LK >> 2;
  constexpr int MW = KT / 8;                              // M-tiles (8 integral rows = 2 basis functions x 4 kinds) per warp and tile
  constexpr int CA_P = REP_IP * REP_CS, CA_OCT = 2 * CA_P;
  const int offVa = ((warp * 2 + (g >> 2)) * LP + c4) * 4 + (g & 3), strideVm = 32 * LP;
  const int offCb = c4 * JS + g;
  int offXw[MW];                                          // where this lane's two X values of M-tile m go
#pragma unroll
  for (int m = 0; m < MW; ++m) {
    const int row = (m * 4 + warp) * 8 + g;               // tile row -> (basis function row >> 2, integral kind row & 3)
    offXw[m] = ((row & 3) * KT + (row >> 2)) * JS + 2 * c4;
  }
  int vbuf = 0, cabuf = 0;
  for (int t = 0; t < ntile; ++t) {
    mbar_wait(bar + (t & 1), (unsigned)((t >> 1) & 1));
    const double *Vb = Vt + (size_t)vbuf * tileV_doubles;
    double *Xb = Xs + (t & 1) * (int)tileX;
    // phase A: X[(k,q)][j] = sum_l V[(k,q)][l] CB[j][l]; warp w owns M-tiles w, w + 4
    {
      double acc[MW][NNT][2];
#pragma unroll
      for (int m = 0; m < MW; ++m)
#pragma unroll
        for (int n = 0; n < NNT; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
      // A fragment: row g of M-tile mt is (basis function 2 mt + g / 4, kind g % 4), column c4 is l = 4 ks + c4
      const double *va = Vb + offVa;
      const double *cb = CBs + offCb;
#pragma unroll 2
      for (int ks = 0; ks < nks; ++ks) {
        double a[MW], b[NNT];
#pragma unroll
        for (int m = 0; m < MW; ++m) a[m] = va[m * strideVm];
#pragma unroll
        for (int n = 0; n < NNT; ++n) b[n] = cb[n * 8];
#pragma unroll
        for (int m = 0; m < MW; ++m)
#pragma unroll
          for (int n = 0; n < NNT; ++n) dmma884(acc[m][n][0], acc[m][n][1], a[m], b[n]);
        va += 16; cb += 4 * JS;
      }
#pragma unroll
      for (int m = 0; m < MW; ++m)
#pragma unroll
        for (int n = 0; n < NNT; ++n) *reinterpret_cast<double2 *>(Xb + offXw[m] + n * 8) = make_double2(acc[m][n][0], acc[m][n][1]);
    }
    __syncthreads();                                       // X complete; every warp is done with this V buffer
    if (tid == 0 && t + L.nbuf < ntile) issue(t + L.nbuf);
    // phase B: IJ[kind][i][j] += sum_k C[i][k] X[k][q][j]
    {
      const double *cat = CAs + cabuf * (int)tileCA_doubles;
#pragma unroll
      for (int s = 0; s < REP_MAX_NIT; ++s) {
        if (ukind[s] >= 0) {
          const double *ca = cat + uoffA[s], *x0 = Xb + uoffX0[s], *x1 = Xb + uoffX1[s];
          const bool two = ukind[s] >= 2;
#pragma unroll
          for (int k4 = 0; k4 < KT / 4; ++k4) {
            const double a0 = ca[(k4 >> 1) * CA_OCT + (k4 & 1) * 4];
#pragma unroll
            for (int n = 0; n < NNT; ++n) dmma884(accB[s][n][0], accB[s][n][1], a0, x0[k4 * 4 * JS + n * 8]);
            if (two) {
              const double a1 = ca[(k4 >> 1) * CA_OCT + CA_P + (k4 & 1) * 4];
#pragma unroll
              for (int n = 0; n < NNT; ++n) dmma884(accB[s][n][0], accB[s][n][1], a1, x1[k4 * 4 * JS + n * 8]);
            }
          }
        }
      }
    }
    if (++vbuf == L.nbuf) vbuf = 0;
    if (++cabuf == ncab) cabuf = 0;
  }
#pragma unroll
  for (int s = 0; s < REP_MAX_NIT; ++s) {
    if (ukind[s] >= 0) {
      const int kind = ukind[s], i = urow[s];
#pragma unroll
      for (int n = 0; n < NNT; ++n) {
        const int j = n * 8 + 2 * c4;
        if (i < nmosA && j < nmosB) IJ[kind * nij + i * nmosB + j] = accB[s][n][0];
        if (i < nmosA && j + 1 < nmosB) IJ[kind * nij + i * nmosB + j + 1] = accB[s][n][1];
      }
    }
  }
  __syncthreads();
  // ---- CALCFI (shftex.f:69-295): every sum over centroid / atom distances comes from the tables above
  double acc = 0.0;
  {
    const double *Sij = IJ, *Tij = IJ + nij, *Smij = IJ + 2 * nij, *Tmij = IJ + 3 * nij;
    const double *FA = p.dtab + tm0[SLVGPU_M_O_FOCK], *FA1 = p.dtab + p.sol_meta[SLVGPU_S_FOCK1], *FB = p.dtab + tm[SLVGPU_M_O_FOCK];
    const double PI = 3.14159265358979323846;
    for (int w = tid; w < nij; w += NT) {
      const int i = w / nmosB, j = w - i * nmosB;
      const double S = Sij[w], T = Tij[w], Sm = Smij[w], Tm = Tmij[w];
      if (S == 0.0) continue;       // every primitive pair screened out (fragments far apart although their centres are inside
                                    // ecut): the term vanishes like S ln S; log(0) must not turn it into a NaN
      const double sl = log(fabs(S)), S2 = S * S;
      const double sdr = S * RI[w];
      const double aux = 2.0 * sqrt(-2.0 * sl / PI);
      double f = 4.0 * sdr * Sm * (sqrt(-1.0 / (2.0 * PI * sl)) - aux - 1.0) + 2.0 * sdr * sdr * RM[w] * (aux + 1.0)
               + 4.0 * Sm * T + 4.0 * Tm * S;
      double g1 = 0, g2 = 0, g3 = 0;                       // (F_A S)_ij, (F1_A S)_ij, (F_A Sm)_ij
      for (int k = 0; k < nmosA; ++k) {
        const double fik = FA[i * nmosA + k], skj = Sij[k * nmosB + j];
        g1 = fma(fik, skj, g1); g2 = fma(FA1[i * nmosA + k], skj, g2); g3 = fma(fik, Smij[k * nmosB + j], g3);
      }
      double h1 = 0, h2 = 0;                               // (S F_B^T)_ij, (Sm F_B^T)_ij
      for (int l = 0; l < nmosB; ++l) {
        const double fjl = FB[j * nmosB + l];
        h1 = fma(fjl, Sij[i * nmosB + l], h1); h2 = fma(fjl, Smij[i * nmosB + l], h2);
      }
      f += -2.0 * Sm * g1 - 2.0 * S * (g2 + g3) + 8.0 * S * Sm * col1[j] - 4.0 * S2 * col2[j];       // TA1
      f += -2.0 * Sm * h1 - 2.0 * S * h2 + 8.0 * S * Sm * row1[i] - 4.0 * S2 * row2[i];              // TA2
      f += 2.0 * S2 * u3[j] - 4.0 * S * Sm * w3[j];                                                   // TA3
      f += 2.0 * S2 * u4[i] - 4.0 * S * Sm * w4[i];                                                   // TA4
      acc += f;
    }
  }
  // the integral kernels store zeros for non-finite values (the pool is recycled and its pad cells are read by later
  // batches): a pair with non-finite coordinates must still come out non-finite
  for (int w = tid; w < 3 * (natA + natB); w += NT) acc += 0.0 * (w < 3 * natA ? posA[w] : posB[w - 3 * natA]);
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
