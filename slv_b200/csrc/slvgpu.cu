// slvgpu.cu -- C ABI (include/slvgpu.h) over the sm_100a kernels of the SolEFP frequency-shift path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC slvgpu.cu -o libslvgpu.so
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/slvgpu.h"
#include "k_disp.cuh"
#include "k_elect.cuh"
#include "k_pol.cuh"
#include "k_rep.cuh"
#include "k_spectra.cuh"
#include "k_util.cuh"
#include "xtc_reader.h"

using namespace slv;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(SLVGPU_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));               \
  } while (0)

struct slvgpu_plan {
  DevPlan dp;
  int max_chunk = 0, sm_count = 0, pol_ctas = 0, rep_ctas = 0;
  int ndmac = 0, npolc = 0, nmosc = 0, nbasc = 0, natc = 0;
  int max_nmos = 0, max_nbas = 0, max_nat = 0;
  std::vector<void *> allocs;
  FrameLists fl{};
  PolWork pw{};
  RepPool rp[2] = {};
  RepLimits rep_lim;
  cudaStream_t side = nullptr;                // plan-owned stream: every second exchange-repulsion batch runs on it
  cudaStream_t hp = nullptr, rep0 = nullptr;  // tail fill: k_pol on a high-priority stream, the other batches on a stream of their own
  cudaEvent_t ev_a = nullptr, ev_pol = nullptr, ev_rep0 = nullptr;
  int tailfill = 1;                           // SLVGPU_TAILFILL=0: k_pol and the first batch stream stay on the caller's stream
  cudaStream_t fan[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // per batch stream: extra streams the integral classes fan out over
  cudaEvent_t ev_side = nullptr, ev_fork[2] = {nullptr, nullptr}, ev_join[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  int rep_skip = 0;                           // experiments: SLVGPU_REP_SKIP=1 no integral kernels, 2 no transform kernels (results are then wrong)
  int rep_fan = -1;                           // extra streams per batch the integral kernels use (SLVGPU_REP_FAN, 0..2; -1: by type count)
  std::vector<int> rep_types;                 // fragment types that carry the wave-function tables
  std::vector<std::vector<int>> rep_cls;      // per such type: the 24 class offsets of its integral work list
  std::vector<RepSmem> rep_sm;                // per such type: shared-memory plan of k_rep_pair
  std::vector<int> h_pair0;
  double *d_glob_sites = nullptr; int glob_ctas = 0;   // energy mode: per-CTA site tables of k_global_elect
  int rep_ppb = 0, rep_bt = 128;              // integral kernels: pairs per block (0 = by batch size) and block size (testing knobs)
  double *d_detail = nullptr;
  int *d_queue = nullptr;
  int detail_frames = 0;
  int want_detail = 0;
  double *d_stage_coords2[2] = {nullptr, nullptr}, *d_stage_out = nullptr; int *d_stage_status = nullptr;
  void *d_stage_src[2] = {nullptr, nullptr}; size_t stage_src_bytes = 0; int32_t *d_idx = nullptr;   // float32 ingest
  int64_t stage_frames = 0, stage_slice = 0;
  int64_t last_launches = 0;
  cudaStream_t own_stream = nullptr, copy_stream = nullptr;
  std::vector<cudaEvent_t> copy_ev;
  int timing = 0;
  std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
  std::vector<int> ev_kind;            // kernel id of each (start, stop) pair in use
  double t_acc[4] = {0, 0, 0, 0}; int64_t t_cnt[4] = {0, 0, 0, 0};
};

template <class T>
static int dalloc(slvgpu_plan *pl, T **ptr, size_t n) {
  void *q = nullptr;
  if (n == 0) n = 1;
  cudaError_t e = cudaMalloc(&q, n * sizeof(T));
  if (e != cudaSuccess) return fail(SLVGPU_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  pl->allocs.push_back(q);
  *ptr = (T *)q;
  return 0;
}
template <class T>
static int upload(slvgpu_plan *pl, const T **dptr, const T *h, size_t n) {
  T *q = nullptr;
  int rc = dalloc(pl, &q, n);
  if (rc) return rc;
  if (n) CK(cudaMemcpy(q, h, n * sizeof(T), cudaMemcpyHostToDevice));
  *dptr = q;
  return 0;
}

extern "C" const char *slvgpu_last_error(void) { return g_err.c_str(); }
extern "C" const char *slvgpu_version(void) { return "slvgpu 0.1 (sm_100a)"; }

static void ev_begin(slvgpu_plan *pl, int kind, cudaStream_t st) {
  if (!pl->timing) return;
  if (pl->ev_used + 2 > pl->ev_pool.size()) {
    for (int i = 0; i < 64; ++i) { cudaEvent_t e; cudaEventCreate(&e); pl->ev_pool.push_back(e); }
  }
  pl->ev_kind.resize(pl->ev_used / 2);       // an error return between ev_begin and ev_end left an unrecorded entry behind
  cudaEventRecord(pl->ev_pool[pl->ev_used], st);
  pl->ev_kind.push_back(kind);
}
static void ev_end(slvgpu_plan *pl, cudaStream_t st) {
  if (!pl->timing) return;
  cudaEventRecord(pl->ev_pool[pl->ev_used + 1], st);
  pl->ev_used += 2;
}

extern "C" int slvgpu_set_timing(slvgpu_plan *pl, int on) {
  if (!pl) return fail(SLVGPU_EINVAL, "null plan");
  pl->timing = on; pl->ev_used = 0; pl->ev_kind.clear();
  for (int k = 0; k < 4; ++k) { pl->t_acc[k] = 0; pl->t_cnt[k] = 0; }
  return 0;
}

extern "C" int slvgpu_get_timing(slvgpu_plan *pl, double *ms4, int64_t *launches4) {
  if (!pl || !ms4 || !launches4) return fail(SLVGPU_EINVAL, "null argument");
  CK(cudaDeviceSynchronize());
  if (pl->ev_kind.size() > pl->ev_used / 2) pl->ev_kind.resize(pl->ev_used / 2);
  for (size_t i = 0; i < pl->ev_kind.size(); ++i) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, pl->ev_pool[2 * i], pl->ev_pool[2 * i + 1]));
    pl->t_acc[pl->ev_kind[i]] += ms; pl->t_cnt[pl->ev_kind[i]] += 1;
  }
  pl->ev_used = 0; pl->ev_kind.clear();
  for (int k = 0; k < 4; ++k) { ms4[k] = pl->t_acc[k]; launches4[k] = pl->t_cnt[k]; }
  return 0;
}

extern "C" void slvgpu_plan_destroy(slvgpu_plan *pl) {
  if (!pl) return;
  for (cudaEvent_t e : pl->ev_pool) cudaEventDestroy(e);
  for (void *q : pl->allocs) cudaFree(q);
  for (int b = 0; b < 2; ++b) rep_pool_free(pl->rp[b]);
  for (int b = 0; b < 2; ++b) if (pl->d_stage_coords2[b]) cudaFree(pl->d_stage_coords2[b]);
  for (int b = 0; b < 2; ++b) if (pl->d_stage_src[b]) cudaFree(pl->d_stage_src[b]);
  if (pl->d_idx) cudaFree(pl->d_idx);
  if (pl->d_stage_out) cudaFree(pl->d_stage_out);
  if (pl->d_stage_status) cudaFree(pl->d_stage_status);
  if (pl->own_stream) cudaStreamDestroy(pl->own_stream);
  if (pl->copy_stream) cudaStreamDestroy(pl->copy_stream);
  if (pl->side) cudaStreamDestroy(pl->side);
  if (pl->hp) cudaStreamDestroy(pl->hp);
  if (pl->rep0) cudaStreamDestroy(pl->rep0);
  if (pl->ev_side) cudaEventDestroy(pl->ev_side);
  if (pl->ev_a) cudaEventDestroy(pl->ev_a);
  if (pl->ev_pol) cudaEventDestroy(pl->ev_pol);
  if (pl->ev_rep0) cudaEventDestroy(pl->ev_rep0);
  for (int b = 0; b < 2; ++b) {
    if (pl->ev_fork[b]) cudaEventDestroy(pl->ev_fork[b]);
    for (int k = 0; k < 2; ++k) { if (pl->fan[b][k]) cudaStreamDestroy(pl->fan[b][k]); if (pl->ev_join[b][k]) cudaEventDestroy(pl->ev_join[b][k]); }
  }
  for (cudaEvent_t e : pl->copy_ev) cudaEventDestroy(e);
  delete pl;
}

extern "C" int slvgpu_plan_create(const slvgpu_plan_desc *d, slvgpu_plan **out) {
  if (!d || !out) return fail(SLVGPU_EINVAL, "null argument");
  if (d->ntypes < 1 || d->ninst < 1 || d->natoms_total < 1) return fail(SLVGPU_EINVAL, "empty plan");
  if (!d->type_meta || !d->sol_meta || !d->itab || !d->dtab || !d->inst_type || !d->inst_atom0)
    return fail(SLVGPU_EINVAL, "null table pointer");
  for (int i = 0; i < d->ninst; ++i)
    if (d->inst_type[i] < 0 || d->inst_type[i] >= d->ntypes) return fail(SLVGPU_EINVAL, "instance type out of range");
  // Every table a requested term reads must be present: a missing section of a .frg file leaves its offset at -1,
  // which the kernels would use as an address.  The reference raises KeyError on the missing parameter
  // (solvshift/solefp.py:1429 par['dpoli'], 1305 par['dpol'], 1531 par['fock']).
  {
    const int32_t *t0 = d->type_meta + (size_t)d->inst_type[0] * SLVGPU_NMETA;
    auto need_sol = [&](int k) { return d->sol_meta[k] >= 0; };
    auto miss = [&](const char *term, const char *what, int type) {
      return fail(SLVGPU_EINVAL, std::string(term) + " needs the '" + what + "' table of fragment type " + std::to_string(type) +
                                     ", which its parameter file does not provide");
    };
    const int ts = d->inst_type[0];
    const bool glob = (d->flags & SLVGPU_GLOBAL) != 0;       // energy mode: no mode-derivative tables are read
    if (glob && (d->flags & (SLVGPU_DISP | SLVGPU_REP)))
      return fail(SLVGPU_EINVAL, "the energy mode evaluates electrostatics and polarisation only (solvshift/solefp.py:906-908, 917-919)");
    if ((d->flags & SLVGPU_ELECT) && glob) {
      if (t0[SLVGPU_M_O_RDMA] < 0 || t0[SLVGPU_M_O_MOM] < 0) return miss("elect", "distributed multipoles", ts);
    } else if (d->flags & SLVGPU_ELECT) {
      if (t0[SLVGPU_M_O_RDMA] < 0 || t0[SLVGPU_M_NDMA] < 1) return miss("elect", "distributed multipoles", ts);
      if (!need_sol(SLVGPU_S_SMEA) || !need_sol(SLVGPU_S_SEA) || !need_sol(SLVGPU_S_CMEA) || !need_sol(SLVGPU_S_CEA))
        return miss("elect", "multipole derivatives (dma1/dma2, lvec, gijk)", ts);
    }
    if (d->flags & SLVGPU_DISP) {
      if (t0[SLVGPU_M_NPOL] < 1 || t0[SLVGPU_M_O_RPOL] < 0 || t0[SLVGPU_M_O_DPOLI] < 0) return miss("disp", "dpoli", ts);
      if (!need_sol(SLVGPU_S_DPOLI1)) return miss("disp", "dpoli1", ts);
      if (!need_sol(SLVGPU_S_LMOC1)) return miss("disp", "lmoc1", ts);
    }
    if ((d->flags & SLVGPU_POL) && glob) {
      if (t0[SLVGPU_M_NPOL] < 1 || t0[SLVGPU_M_O_RPOL] < 0 || t0[SLVGPU_M_O_POL] < 0) return miss("pol", "dpol", ts);
      if (t0[SLVGPU_M_O_RDMA] < 0 || t0[SLVGPU_M_O_MOM] < 0) return miss("pol", "distributed multipoles", ts);
    } else if (d->flags & SLVGPU_POL) {
      if (t0[SLVGPU_M_NPOL] < 1 || t0[SLVGPU_M_O_RPOL] < 0 || t0[SLVGPU_M_O_POL] < 0) return miss("pol", "dpol", ts);
      if (t0[SLVGPU_M_O_RDMA] < 0 || t0[SLVGPU_M_O_MOM] < 0) return miss("pol", "distributed multipoles", ts);
      if (!need_sol(SLVGPU_S_DPOL1)) return miss("pol", "dpol1", ts);
      if (!need_sol(SLVGPU_S_LMOC1)) return miss("pol", "lmoc1", ts);
      if (!need_sol(SLVGPU_S_DMA1) || !need_sol(SLVGPU_S_LVEC)) return miss("pol", "dma1 / lvec", ts);
    }
    for (int i = 1; i < d->ninst; ++i) {
      const int ty = d->inst_type[i];
      const int32_t *tm = d->type_meta + (size_t)ty * SLVGPU_NMETA;
      if ((d->flags & (SLVGPU_ELECT | SLVGPU_POL)) && tm[SLVGPU_M_NDMA] > 0 && (tm[SLVGPU_M_O_RDMA] < 0 || tm[SLVGPU_M_O_MOM] < 0))
        return miss("elect/pol", "distributed multipoles", ty);
      if (tm[SLVGPU_M_REMOVABLE] == 2) {                  // non-EFP environment: electrostatics only, nothing else is read
        if (tm[SLVGPU_M_NPOL] > 0) return fail(SLVGPU_EINVAL, "environment fragment types must not carry polarisabilities");
        continue;
      }
      if ((d->flags & (SLVGPU_POL | SLVGPU_DISP)) && tm[SLVGPU_M_NPOL] > 0 && (tm[SLVGPU_M_O_RPOL] < 0 || tm[SLVGPU_M_O_POL] < 0))
        return miss("pol/disp", "dpol", ty);
      if ((d->flags & SLVGPU_DISP) && tm[SLVGPU_M_NPOL] > 0 && tm[SLVGPU_M_O_DPOLI] < 0) return miss("disp", "dpoli", ty);
      if ((d->flags & SLVGPU_REP) && (tm[SLVGPU_M_O_VECL] < 0 || tm[SLVGPU_M_O_FOCK] < 0 || tm[SLVGPU_M_O_LMOC] < 0 ||
                                      tm[SLVGPU_M_IO_SHPAIR] < 0 || tm[SLVGPU_M_NMOS] < 1))
        return miss("rep", "wave function (vecl, fock, lmoc, basis)", ty);
    }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(SLVGPU_ECUDA, "no CUDA device: this library has no CPU path");
  slvgpu_plan *pl = new slvgpu_plan();
  DevPlan &p = pl->dp;
  memset(&p, 0, sizeof(p));
  p.ntypes = d->ntypes; p.ninst = d->ninst; p.natoms_total = d->natoms_total; p.flags = d->flags;
  p.ccut = d->ccut; p.pcut = d->pcut; p.ecut = d->ecut;
  p.pol_maxit = d->pol_maxit > 0 ? d->pol_maxit : 1000;
  p.pol_tol = d->pol_tol > 0 ? d->pol_tol : 1e-10;
  int rc = 0;
#define TRY(x) do { rc = (x); if (rc) { slvgpu_plan_destroy(pl); return rc; } } while (0)
  TRY(upload(pl, &p.type_meta, d->type_meta, (size_t)d->ntypes * SLVGPU_NMETA));
  TRY(upload(pl, &p.sol_meta, d->sol_meta, (size_t)SLVGPU_NSOL));
  TRY(upload(pl, &p.itab, d->itab, (size_t)d->n_itab));
  TRY(upload(pl, &p.dtab, d->dtab, (size_t)d->n_dtab));
  TRY(upload(pl, &p.inst_type, d->inst_type, (size_t)d->ninst));
  TRY(upload(pl, &p.inst_atom0, d->inst_atom0, (size_t)d->ninst + 1));
  // Gauss-Legendre weights of the imaginary-frequency quadrature (solvshift/slvpar.py:1000-1019), v = 0.3
  {
    static const double w12[12] = {0.0471753363865118, 0.1069393259953184, 0.1600783285433462, 0.2031674267230659,
                                   0.2334925365383548, 0.2491470458134028, 0.2491470458134028, 0.2334925365383548,
                                   0.2031674267230659, 0.1600783285433462, 0.1069393259953184, 0.0471753363865118};
    static const double t12[12] = {-0.9815606342467192, -0.9041172563704749, -0.7699026741943047, -0.5873179542866175,
                                   -0.3678314989981802, -0.1252334085114689, 0.1252334085114689, 0.3678314989981802,
                                   0.5873179542866175, 0.7699026741943047, 0.9041172563704749, 0.9815606342467192};
    for (int k = 0; k < 12; ++k) p.glw[k] = 2.0 * 0.3 * w12[k] / ((1.0 - t12[k]) * (1.0 - t12[k]));
  }
  // sizes
  long tot_npol = 0, tot_ndma = 0;
  p.max_npol = 0; p.max_ndma = 0;
  for (int t = 0; t < d->ntypes; ++t) {
    const int32_t *tm = d->type_meta + (size_t)t * SLVGPU_NMETA;
    if (tm[SLVGPU_M_NPOL] > p.max_npol) p.max_npol = tm[SLVGPU_M_NPOL];
    if (tm[SLVGPU_M_NDMA] > p.max_ndma) p.max_ndma = tm[SLVGPU_M_NDMA];
  }
  std::vector<char> solvent_type((size_t)d->ntypes, 0);        // types that occur among the solvent instances
  for (int i = 0; i < d->ninst; ++i) {
    const int32_t *tm = d->type_meta + (size_t)d->inst_type[i] * SLVGPU_NMETA;
    tot_npol += tm[SLVGPU_M_NPOL]; tot_ndma += tm[SLVGPU_M_NDMA];
    if (i > 0 && tm[SLVGPU_M_REMOVABLE] != 2) {                 // exchange-repulsion partner dimensions
      solvent_type[d->inst_type[i]] = 1;
      if (tm[SLVGPU_M_NMOS] > pl->max_nmos) pl->max_nmos = tm[SLVGPU_M_NMOS];
      if (tm[SLVGPU_M_NBASIS] > pl->max_nbas) pl->max_nbas = tm[SLVGPU_M_NBASIS];
      if (tm[SLVGPU_M_NATOMS] > pl->max_nat) pl->max_nat = tm[SLVGPU_M_NATOMS];
    }
  }
  if (pl->max_nmos < 1) pl->max_nmos = 1;
  if (pl->max_nbas < 1) pl->max_nbas = 1;
  if (pl->max_nat < 1) pl->max_nat = 1;
  const int32_t *tm0 = d->type_meta + (size_t)d->inst_type[0] * SLVGPU_NMETA;
  pl->ndmac = tm0[SLVGPU_M_NDMA]; pl->npolc = tm0[SLVGPU_M_NPOL]; pl->nmosc = tm0[SLVGPU_M_NMOS];
  pl->nbasc = tm0[SLVGPU_M_NBASIS]; pl->natc = tm0[SLVGPU_M_NATOMS];
  if ((d->flags & SLVGPU_POL) && !(d->flags & SLVGPU_GLOBAL) && pl->ndmac != pl->natc) {
    slvgpu_plan_destroy(pl);
    return fail(SLVGPU_EINVAL, "polarisation shift needs ndma == natoms on the solute (solpol2.f:738-741, 873-876)");
  }
  p.list_cap = d->ninst > 1 ? d->ninst - 1 : 1;
  p.cent_cap = (int)tot_npol; p.site_cap = (int)tot_ndma;
  if (d->pol_cent_cap > 0 && d->pol_cent_cap < p.cent_cap) {
    p.cent_cap = d->pol_cent_cap > pl->npolc ? d->pol_cent_cap : pl->npolc + 1;
    // sites scale with centres through the fragment types; bound by the worst ratio
    double ratio = 1.0;
    for (int t = 0; t < d->ntypes; ++t) {
      const int32_t *tm = d->type_meta + (size_t)t * SLVGPU_NMETA;
      if (tm[SLVGPU_M_NPOL] > 0) ratio = fmax(ratio, (double)tm[SLVGPU_M_NDMA] / tm[SLVGPU_M_NPOL]);
    }
    long sc = (long)(p.cent_cap * ratio) + pl->ndmac + 8;
    p.site_cap = (int)(sc < tot_ndma ? sc : tot_ndma);
  }
  cudaDeviceProp prop;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaGetDeviceProperties(&prop, dev);
  pl->sm_count = prop.multiProcessorCount;
  pl->max_chunk = d->max_chunk > 0 ? d->max_chunk : 4096;
  pl->want_detail = d->want_detail;
  const size_t ch = (size_t)pl->max_chunk;
  TRY(dalloc(pl, &pl->d_queue, 4));
  TRY(dalloc(pl, &pl->fl.sol_xf, ch * 12));
  TRY(dalloc(pl, &pl->fl.nlist, ch));
  TRY(dalloc(pl, &pl->fl.list_inst, ch * p.list_cap));
  TRY(dalloc(pl, &pl->fl.list_flag, ch * p.list_cap));
  TRY(dalloc(pl, &pl->fl.list_xf, ch * p.list_cap * 12));
  if (d->flags & SLVGPU_POL) {
    pl->pol_ctas = pl->sm_count * POL_CTAS_PER_SM;    // three persistent 256-thread CTAs per SM: one frame each, their barrier phases interleave (4 x 128 measured: better
    // for ~400 centres, worse at ~500 and a longer tail -- DESIGN.md)
    const size_t nc = (size_t)pl->pol_ctas, cc = (size_t)p.cent_cap;
    TRY(dalloc(pl, &pl->pw.cent_pos, nc * cc * 3));
    TRY(dalloc(pl, &pl->pw.cent_alpha, nc * cc * AST));
    TRY(dalloc(pl, &pl->pw.fld, nc * cc * 3));
    TRY(dalloc(pl, &pl->pw.xy, nc * cc * (PST + 8)));
    TRY(dalloc(pl, &pl->pw.site, nc * (size_t)p.site_cap * SITE_ROW));
    TRY(dalloc(pl, &pl->pw.aux, nc * cc * 6));
    TRY(dalloc(pl, &pl->pw.part, nc * (size_t)(POL_THREADS / 32) * 2 * 64 * 6));
    TRY(dalloc(pl, &pl->pw.cent_mol, nc * cc));
    TRY(dalloc(pl, &pl->pw.ent_coff, nc * (size_t)p.list_cap));
    TRY(dalloc(pl, &pl->pw.ent_soff, nc * (size_t)p.list_cap));
    if (pl->want_detail) {
      pl->detail_frames = pl->max_chunk;
      TRY(dalloc(pl, &pl->d_detail, (size_t)pl->detail_frames * (cc * 12 + 1)));
    }
  }
  if ((d->flags & SLVGPU_GLOBAL) && (d->flags & SLVGPU_ELECT)) {
    pl->glob_ctas = pl->sm_count * 2;
    TRY(dalloc(pl, &pl->d_glob_sites, (size_t)pl->glob_ctas * (size_t)p.site_cap * 24));
  }
  if (d->flags & SLVGPU_REP) {
    if (tm0[SLVGPU_M_O_VECL] < 0 || d->sol_meta[SLVGPU_S_VECL1] < 0) {
      slvgpu_plan_destroy(pl);
      return fail(SLVGPU_EINVAL, "exchange-repulsion needs the wave-function tables (vecl, fock, basis) in the plan");
    }
    if (pl->nmosc > 8 * REP_MAX_NIT) {
      slvgpu_plan_destroy(pl);
      return fail(SLVGPU_EINVAL, "exchange repulsion: the solute has more than 24 LMOs (k_rep_pair keeps its LMO integrals in registers)");
    }
    TRY(rep_alloc(pl->rp, pl->rep_lim, pl->allocs, g_err, pl->nmosc, pl->nbasc, pl->natc, pl->max_nbas, pl->max_nat, pl->max_chunk, p.list_cap, p.ntypes));
    for (int t = 0; t < d->ntypes; ++t) {
      const int32_t *tm = d->type_meta + (size_t)t * SLVGPU_NMETA;
      if (tm[SLVGPU_M_IO_SHPAIR] < 0 || !solvent_type[t]) continue;
      const RepSmem rs = rep_smem_plan(pl->nmosc, pl->nbasc, pl->natc, tm[SLVGPU_M_NMOS], tm[SLVGPU_M_NBASIS], tm[SLVGPU_M_NATOMS]);
      if (rs.NNT > REP_MAX_NNT || rs.bytes > (size_t)220 * 1024) {
        slvgpu_plan_destroy(pl);
        return fail(SLVGPU_EINVAL, "exchange repulsion: fragment type " + std::to_string(t) + " is too large for k_rep_pair "
                                   "(more than 40 LMOs, or its coefficient table does not fit shared memory)");
      }
      pl->rep_types.push_back(t);
      pl->rep_cls.emplace_back(d->itab + tm[SLVGPU_M_IO_SHPAIR], d->itab + tm[SLVGPU_M_IO_SHPAIR] + SUBSHELL_NCLASS + 1);
      pl->rep_sm.push_back(rs);
    }
    pl->h_pair0.resize((size_t)pl->max_chunk + 1);
  }
#undef TRY
  // opt in to large dynamic shared memory where needed
  cudaFuncSetAttribute(k_frame_elect, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_disp, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_pol, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
#define RPA(KT, NNT) cudaFuncSetAttribute(k_rep_pair<KT, NNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  RPA(8, 1) RPA(8, 2) RPA(8, 3) RPA(8, 4) RPA(8, 5) RPA(16, 1) RPA(16, 2) RPA(16, 3) RPA(16, 4) RPA(16, 5)
#undef RPA
#define X(ID, LA, LB, IA) cudaFuncSetAttribute(k_rep_ints<LA, LB, IA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  SLV_SUBSHELL_CLASSES(X)
#undef X
  if (const char *ev = getenv("SLVGPU_REP_SKIP")) pl->rep_skip = atoi(ev);
  if (const char *ev = getenv("SLVGPU_TAILFILL")) pl->tailfill = atoi(ev) != 0;
  if (const char *ev = getenv("SLVGPU_REP_FAN")) { const int v = atoi(ev); if (v >= 0 && v <= 2) pl->rep_fan = v; }
  if (const char *ev = getenv("SLVGPU_REP_PPB")) { const int v = atoi(ev); if (v >= 1 && v <= 32) pl->rep_ppb = v; }
  if (const char *ev = getenv("SLVGPU_REP_BT")) { const int v = atoi(ev); if (v == 32 || v == 64 || v == 128) pl->rep_bt = v; }
  *out = pl;
  return 0;
}

__global__ void k_status(const double *__restrict__ out, int *__restrict__ status, int64_t n) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const double *o = out + f * SLVGPU_NOUT;
  bool bad = false;
  for (int k = 0; k < 8; ++k) bad |= !isfinite(o[k]);
  if (bad) status[f] |= SLVGPU_ST_NONFINITE;
}

// Stage 1 of a chunk: superimposition + electrostatics of `n` frames whose coordinates start at frame `cframe0` of
// d_coords, then their dispersion (per frame, from the hand-over lists just written); results go to rows oframe0.. of
// d_out, the lists to slots lframe0.. of the chunk's lists.  The host entry feeds a chunk slice by slice (each slice's copy
// hides behind the previous slice's two kernels); coordinates are not read again after k_frame_elect.
static int launch_elect(slvgpu_plan *pl, const double *d_coords, int64_t cframe0, int64_t oframe0, int lframe0, int n, double *d_out,
                        cudaStream_t st) {
  const DevPlan &p = pl->dp;
  const size_t sm_el = elect_smem(pl->ndmac);
  FrameLists fl = pl->fl;
  fl.sol_xf += (size_t)lframe0 * 12; fl.nlist += lframe0;
  fl.list_inst += (size_t)lframe0 * p.list_cap; fl.list_flag += (size_t)lframe0 * p.list_cap; fl.list_xf += (size_t)lframe0 * p.list_cap * 12;
  ev_begin(pl, 0, st);
  k_frame_elect<<<n, ELECT_THREADS, sm_el, st>>>(p, d_coords + (size_t)cframe0 * p.natoms_total * 3, (int)oframe0, fl, d_out);
  ev_end(pl, st);
  ++pl->last_launches;
  if ((p.flags & SLVGPU_DISP) && p.ninst > 1) {
    const size_t sm_di = ((size_t)pl->npolc * DISP_SOLROW + (108 + 4) * DISP_TJ) * sizeof(double) + ((size_t)p.list_cap + 2) * sizeof(int);
    ev_begin(pl, 1, st);
    k_disp<<<n, DISP_THREADS, sm_di, st>>>(p, (int)oframe0, fl, d_out);
    ev_end(pl, st);
    ++pl->last_launches;
  }
  CK(cudaGetLastError());
  return 0;
}

// Stage 2 of a chunk: everything that works from the hand-over lists of its nf frames (rows f0.. of d_out).
static int launch_rest(slvgpu_plan *pl, int64_t f0, int nf, double *d_out, int32_t *d_status, cudaStream_t st) {
  const DevPlan &p = pl->dp;
  size_t sm_po = (size_t)POL_TILE * 10, sm_po_sol = (size_t)pl->npolc * 21 + (size_t)pl->ndmac * 46;   // tile; solute tables alias it
  sm_po = (sm_po > sm_po_sol ? sm_po : sm_po_sol) * sizeof(double);
  const bool do_pol = (p.flags & SLVGPU_POL) && p.ninst > 1;
  const bool do_rep = (p.flags & SLVGPU_REP) && p.ninst > 1;
  if (do_rep && !pl->side) {
    CK(cudaStreamCreateWithFlags(&pl->side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&pl->ev_side, cudaEventDisableTiming));
    int lo_pri = 0, hi_pri = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
    CK(cudaStreamCreateWithPriority(&pl->hp, cudaStreamNonBlocking, hi_pri));
    CK(cudaStreamCreateWithFlags(&pl->rep0, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&pl->ev_a, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&pl->ev_pol, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&pl->ev_rep0, cudaEventDisableTiming));
    for (int b = 0; b < 2; ++b) {
      CK(cudaEventCreateWithFlags(&pl->ev_fork[b], cudaEventDisableTiming));
      for (int k = 0; k < 2; ++k) {
        CK(cudaStreamCreateWithFlags(&pl->fan[b][k], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&pl->ev_join[b][k], cudaEventDisableTiming));
      }
    }
  }
  // Schedule of a chunk: k_disp and k_pol on the caller's stream, then the exchange-repulsion batches alternating between
  // it and a plan-owned stream (two integral pools: the integral kernels of one batch run beside the transform kernel of
  // the other), the integral classes of a batch fanned out over two more streams each (their grids are small: the tail of
  // one class fills with the next).  Running dispersion / exchange repulsion BESIDE k_pol was measured and dropped: k_pol
  // holds the whole register file of every SM, and with fewer CTAs it loses more than the side kernels gain
  // (profiles/overlap_pol_ctas_r02.jsonl); FP64 tensor-core and DFMA instructions share one pipe (profiles/dmma_mix_r02.json).
  {
    // The exchange-repulsion batch boundaries need the per-frame fragment counts on the host: one small copy and one
    // synchronisation of the caller's stream per chunk, taken here (k_frame_elect has written the counts) so that
    // everything after it is enqueued without another host stall.
    const bool tail = do_rep && do_pol && pl->tailfill;
    cudaStream_t s0 = tail ? pl->rep0 : st, s1 = pl->side;
    if (do_rep) {
      k_rep_scan<<<1, 1024, 0, st>>>(d_out, (int)f0, nf, pl->rp[0].pair0);
      ++pl->last_launches;
      CK(cudaMemcpyAsync(pl->h_pair0.data(), pl->rp[0].pair0, ((size_t)nf + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      const int *pair0 = pl->h_pair0.data();
      // pools sized for two concurrent batches covering the chunk, and for the largest single frame
      int fmax = 0;
      for (int f = 0; f < nf; ++f) fmax = pair0[f + 1] - pair0[f] > fmax ? pair0[f + 1] - pair0[f] : fmax;
      size_t need = ((size_t)pair0[nf] + 1) / 2;
      if (need < (size_t)fmax) need = (size_t)fmax;
      if (need > (size_t)pl->rp[0].pair_cap && (size_t)pl->rp[0].pair_cap < pl->rep_lim.max_pairs) {
        CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(pl->rep0));     // nothing of an earlier chunk may still use the pools
        const int rc = rep_reserve(pl->rp, pl->rep_lim, need, g_err);
        if (rc) return rc;
      }
    }
    if ((p.flags & SLVGPU_GLOBAL) && (p.flags & SLVGPU_ELECT) && p.ninst > 1) {
      const int g = nf < pl->glob_ctas ? nf : pl->glob_ctas;
      ev_begin(pl, 1, st);
      k_global_elect<<<g, GLOB_THREADS, 0, st>>>(p, (int)f0, nf, pl->fl, pl->d_glob_sites, d_out);
      ev_end(pl, st);
      ++pl->last_launches;
    }
    // Tail fill: k_pol's persistent CTAs hold every register of the SMs, so nothing runs beside them -- until the frame queue
    // is empty and they retire one by one (the last ~2 ms of the launch).  With k_pol on a high-priority stream of the plan
    // and the exchange-repulsion batches on streams that become runnable at the same moment, the integral kernels take the
    // SMs as they fall idle instead of waiting for the last CTA.
    cudaStream_t sp = tail ? pl->hp : st;
    if (tail) {
      CK(cudaEventRecord(pl->ev_a, st));
      CK(cudaStreamWaitEvent(pl->hp, pl->ev_a, 0));
      CK(cudaStreamWaitEvent(s0, pl->ev_a, 0));
      CK(cudaStreamWaitEvent(s1, pl->ev_a, 0));
    }
    if (do_pol) {
      const int g = nf < pl->pol_ctas ? nf : pl->pol_ctas;
      ev_begin(pl, 2, sp);
      CK(cudaMemsetAsync(pl->d_queue, 0, sizeof(int), sp));
      k_pol<<<g, POL_THREADS, sm_po, sp>>>(p, (int)f0, nf, pl->fl, pl->pw, d_out, d_status, pl->d_detail, pl->d_queue);
      ev_end(pl, sp);
      ++pl->last_launches;
    }
    if (do_rep) {
      ev_begin(pl, 3, sp);                                       // the bracket of the exchange repulsion opens where k_pol ends:
                                                                 // it times what the step pays for it
      if (!tail) {                                               // serial schedule: the second batch stream starts after k_pol too
        CK(cudaEventRecord(pl->ev_a, st));
        CK(cudaStreamWaitEvent(s1, pl->ev_a, 0));
      }
      // Exchange repulsion runs in batches of frames sized by the integral pool.
      const int *pair0 = pl->h_pair0.data();
      int batch = 0;
      for (int fs = 0; fs < nf; ++batch) {
        const RepPool &rp = pl->rp[batch & 1];
        cudaStream_t sb = (batch & 1) ? s1 : s0;
        int fe = fs;
        while (fe < nf && pair0[fe + 1] - pair0[fs] <= rp.pair_cap) ++fe;
        if (fe == fs) {
          cudaStreamSynchronize(s1); cudaStreamSynchronize(s0);
          ev_end(pl, s0);
          return fail(SLVGPU_ENOMEM, "exchange repulsion: one frame has more fragments inside ecut than the integral pool holds");
        }
        const int nb = fe - fs, npairs = pair0[fe] - pair0[fs];
        if (npairs > 0) {
          CK(cudaMemsetAsync(rp.tcount, 0, (size_t)p.ntypes * sizeof(int), sb));
          k_rep_prep<<<nb, 128, 0, sb>>>(p, fs, pl->fl, rp);
          ++pl->last_launches;
          // integral kernels: grid.x covers the items of a class, grid.y the pairs of the type in groups of ppb (the
          // type's pair count lives on the device; groups past it leave at once)
          const int ppb = pl->rep_ppb > 0 ? pl->rep_ppb : (npairs >= 8192 ? 16 : 8);
          const unsigned gy = (unsigned)((npairs + ppb - 1) / ppb);
          // many fragment types (protein side chains) mean many small launches per batch: those gain from the fan-out
          // (config 4: 7.1 -> 5.7 ms), two big types do not (config 3: 37.0 -> 37.6 ms)
          const int bi = batch & 1, nfan = pl->rep_fan >= 0 ? pl->rep_fan : (pl->rep_types.size() > 2 ? 2 : 0);
          cudaStream_t fs3[3] = {sb, pl->fan[bi][0], pl->fan[bi][1]};
          if (nfan > 0) {
            CK(cudaEventRecord(pl->ev_fork[bi], sb));
            for (int k = 0; k < nfan; ++k) CK(cudaStreamWaitEvent(pl->fan[bi][k], pl->ev_fork[bi], 0));
          }
          int turn = 0;
          const size_t sm_geo = rep_ints_smem(pl->rep_bt, pl->natc, pl->max_nat);
          for (size_t ti = 0; ti < pl->rep_types.size(); ++ti) {
            const int ty = pl->rep_types[ti];
            const std::vector<int> &co = pl->rep_cls[ti];
#define X(ID, LA, LB, IA)                                                                                         \
            if (co[ID + 1] > co[ID] && pl->rep_skip != 1) {                                                         \
              const int n_items = co[ID + 1] - co[ID];                                                              \
              k_rep_ints<LA, LB, IA><<<dim3((unsigned)((n_items + pl->rep_bt - 1) / pl->rep_bt), gy), pl->rep_bt, sm_geo,  \
                                       fs3[turn++ % (nfan + 1)]>>>(p, ty, co[ID], n_items, ppb, rp);                   \
              ++pl->last_launches;                                                                                  \
            }
            SLV_SUBSHELL_CLASSES(X)
#undef X
          }
          for (int k = 0; k < nfan; ++k) {
            CK(cudaEventRecord(pl->ev_join[bi][k], pl->fan[bi][k]));
            CK(cudaStreamWaitEvent(sb, pl->ev_join[bi][k], 0));
          }
          for (size_t ti = 0; ti < pl->rep_types.size(); ++ti) {
            const int ty = pl->rep_types[ti];
            const RepSmem &rs = pl->rep_sm[ti];
#define RP(KT_, NNT_) if (rs.KT == KT_ && rs.NNT == NNT_ && pl->rep_skip != 2) k_rep_pair<KT_, NNT_><<<npairs, REP_THREADS, rs.bytes, sb>>>(p, ty, pl->fl, rp, rs);
            RP(8, 1) RP(8, 2) RP(8, 3) RP(8, 4) RP(8, 5) RP(16, 1) RP(16, 2) RP(16, 3) RP(16, 4) RP(16, 5)
#undef RP
            ++pl->last_launches;
          }
        }
        k_rep_sum<<<(nb + 127) / 128, 128, 0, sb>>>((int)f0, fs, nb, rp, d_out);
        ++pl->last_launches;
        fs = fe;
      }
      if (batch > 1) {                                          // the first batch stream waits for the second
        CK(cudaEventRecord(pl->ev_side, s1));
        CK(cudaStreamWaitEvent(s0, pl->ev_side, 0));
      }
      if (tail) {                                               // the caller's stream joins k_pol and the batches
        CK(cudaEventRecord(pl->ev_rep0, s0));
        CK(cudaStreamWaitEvent(st, pl->ev_rep0, 0));
        CK(cudaEventRecord(pl->ev_pol, pl->hp));
        CK(cudaStreamWaitEvent(st, pl->ev_pol, 0));
        ev_end(pl, st);
      } else ev_end(pl, s0);
    }
  }
  CK(cudaGetLastError());
  return 0;
}

extern "C" int slvgpu_eval_frames(slvgpu_plan *pl, const double *d_coords, int64_t nframes, double *d_out,
                                  int32_t *d_status, void *stream_) {
  if (!pl || !d_coords || !d_out || !d_status) return fail(SLVGPU_EINVAL, "null argument");
  if (nframes < 0) return fail(SLVGPU_EINVAL, "negative frame count");
  cudaStream_t st = (cudaStream_t)stream_;
  pl->last_launches = 0;
  if (nframes == 0) return 0;
  CK(cudaMemsetAsync(d_status, 0, (size_t)nframes * sizeof(int32_t), st));
  for (int64_t f0 = 0; f0 < nframes; f0 += pl->max_chunk) {
    const int nf = (int)((nframes - f0) < pl->max_chunk ? (nframes - f0) : pl->max_chunk);
    int rc = launch_elect(pl, d_coords, f0, f0, 0, nf, d_out, st);
    if (rc) return rc;
    rc = launch_rest(pl, f0, nf, d_out, d_status, st);
    if (rc) return rc;
  }
  k_status<<<(unsigned)((nframes + 255) / 256), 256, 0, st>>>(d_out, d_status, nframes);
  ++pl->last_launches;
  CK(cudaGetLastError());
  return 0;
}

// float32 trajectory records -> the plan's FP64 frame layout: dst[f][a][d] = (double)src[f][idx ? idx[a] : a][d] * scale
// (the gather of frame[idx] and the unit conversion of util/slv_md-run_nopp:205, done at HBM speed instead of on the host)
__global__ void k_ingest_f32(const float *__restrict__ src, const int *__restrict__ idx, int natoms_src, int natoms, int64_t nframes,
                             double scale, double *__restrict__ dst) {
  const int64_t n = nframes * natoms * 3;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = w / ((int64_t)natoms * 3);
    const int r = (int)(w - f * natoms * 3), a = r / 3, d = r - a * 3;
    const int sa = idx ? idx[a] : a;
    dst[w] = (double)src[(f * natoms_src + sa) * 3 + d] * scale;
  }
}

// Common host entry: h_src holds nframes records of natoms_src atoms, FP64 (elem 8: natoms_src == natoms_total, no gather, no
// scaling) or float32 (elem 4: gathered through d_idx and scaled on the device).
static int eval_frames_host_any(slvgpu_plan *pl, const void *h_src, int elem, int natoms_src, const int32_t *h_idx, double scale,
                                int64_t nframes, double *h_out, int32_t *h_status) {
  const DevPlan &p = pl->dp;
  if (!pl->own_stream) CK(cudaStreamCreateWithFlags(&pl->own_stream, cudaStreamNonBlocking));
  // Device staging is bounded: two coordinate buffers of one SLICE each (double buffering), so trajectories far larger
  // than device memory stream through; the result rows of the whole call stay on the device until the end.  A chunk (up
  // to max_chunk frames: the unit the persistent kernels and the exchange-repulsion batches work on) is fed slice by slice:
  // copy -> (float32: gather + unit conversion) -> k_frame_elect, which is the only kernel that reads coordinates; the copy
  // of a slice hides behind the previous slice's kernel.  Then the rest of the chunk's kernels run once over all its frames
  // -- the step costs what the device-resident path costs plus the first slice's copy.
  const size_t src_frame_bytes = (size_t)natoms_src * 3 * elem;
  const size_t fstride = (size_t)p.natoms_total * 3;
  // slices are whole waves of the per-frame kernels (one CTA per frame, two CTAs per SM): a slice of 1.1 waves costs two
  const int64_t wave = 2 * (int64_t)pl->sm_count;
  int64_t slice = (int64_t)(48.0e6 / (double)src_frame_bytes) / wave * wave;
  slice = slice < wave ? wave : slice;
  if (const char *ev = getenv("SLVGPU_STAGE_SLICE")) { const long v = atol(ev); if (v > 0) slice = v; }   // testing knob: small slices
  if (slice > pl->max_chunk) slice = pl->max_chunk;
  const int64_t first = slice < wave / 2 ? slice : (slice < wave ? slice : wave / 2);     // a short first slice: its copy is the exposed one
  if (pl->stage_slice < slice) {
    for (int b = 0; b < 2; ++b) { if (pl->d_stage_coords2[b]) cudaFree(pl->d_stage_coords2[b]); pl->d_stage_coords2[b] = nullptr; }
    pl->stage_slice = 0;
    for (int b = 0; b < 2; ++b) CK(cudaMalloc(&pl->d_stage_coords2[b], (size_t)slice * fstride * sizeof(double)));
    pl->stage_slice = slice;
  }
  if (elem == 4 && pl->stage_src_bytes < (size_t)slice * src_frame_bytes) {
    for (int b = 0; b < 2; ++b) { if (pl->d_stage_src[b]) cudaFree(pl->d_stage_src[b]); pl->d_stage_src[b] = nullptr; }
    pl->stage_src_bytes = 0;
    for (int b = 0; b < 2; ++b) CK(cudaMalloc(&pl->d_stage_src[b], (size_t)slice * src_frame_bytes));
    pl->stage_src_bytes = (size_t)slice * src_frame_bytes;
  }
  if (pl->stage_frames < nframes) {
    if (pl->d_stage_out) { cudaFree(pl->d_stage_out); cudaFree(pl->d_stage_status); }
    pl->d_stage_out = nullptr; pl->d_stage_status = nullptr; pl->stage_frames = 0;
    CK(cudaMalloc(&pl->d_stage_out, (size_t)nframes * SLVGPU_NOUT * sizeof(double)));
    CK(cudaMalloc(&pl->d_stage_status, (size_t)nframes * sizeof(int32_t)));
    pl->stage_frames = nframes;
  }
  cudaStream_t st = pl->own_stream;
  if (!pl->copy_stream) CK(cudaStreamCreateWithFlags(&pl->copy_stream, cudaStreamNonBlocking));
  while (pl->copy_ev.size() < 4) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); pl->copy_ev.push_back(e); }
  if (elem == 4 && h_idx) {
    if (!pl->d_idx) CK(cudaMalloc(&pl->d_idx, (size_t)p.natoms_total * sizeof(int32_t)));
    CK(cudaMemcpyAsync(pl->d_idx, h_idx, (size_t)p.natoms_total * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemsetAsync(pl->d_stage_status, 0, (size_t)nframes * sizeof(int32_t), st));
  int64_t launches = 0, c = 0;
  pl->last_launches = 0;
  for (int64_t f0 = 0; f0 < nframes; f0 += pl->max_chunk) {
    const int nf = (int)((nframes - f0) < pl->max_chunk ? (nframes - f0) : pl->max_chunk);
    for (int64_t s0 = 0; s0 < nf; ++c) {
      const int b = (int)(c & 1);
      const int64_t want = (f0 == 0 && s0 == 0) ? first : slice;
      const int64_t n = (nf - s0) < want ? (nf - s0) : want;
      if (c >= 2) CK(cudaStreamWaitEvent(pl->copy_stream, pl->copy_ev[2 + b], 0));      // buffer b has been consumed
      void *dst = elem == 4 ? pl->d_stage_src[b] : (void *)pl->d_stage_coords2[b];
      CK(cudaMemcpyAsync(dst, (const char *)h_src + (size_t)(f0 + s0) * src_frame_bytes, (size_t)n * src_frame_bytes,
                         cudaMemcpyHostToDevice, pl->copy_stream));
      CK(cudaEventRecord(pl->copy_ev[b], pl->copy_stream));
      CK(cudaStreamWaitEvent(st, pl->copy_ev[b], 0));
      if (elem == 4) {
        const int64_t work = n * (int64_t)fstride;
        const unsigned grid = (unsigned)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16);
        k_ingest_f32<<<grid, 256, 0, st>>>((const float *)pl->d_stage_src[b], h_idx ? pl->d_idx : nullptr, natoms_src, p.natoms_total, n,
                                            scale, pl->d_stage_coords2[b]);
        ++launches;
      }
      int rc = launch_elect(pl, pl->d_stage_coords2[b], 0, f0 + s0, (int)s0, (int)n, pl->d_stage_out, st);
      if (rc) return rc;
      CK(cudaEventRecord(pl->copy_ev[2 + b], st));
      s0 += n;
    }
    int rc = launch_rest(pl, f0, nf, pl->d_stage_out, pl->d_stage_status, st);
    if (rc) return rc;
  }
  k_status<<<(unsigned)((nframes + 255) / 256), 256, 0, st>>>(pl->d_stage_out, pl->d_stage_status, nframes);
  launches += pl->last_launches + 1;
  pl->last_launches = launches;
  CK(cudaMemcpyAsync(h_out, pl->d_stage_out, (size_t)nframes * SLVGPU_NOUT * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_status, pl->d_stage_status, (size_t)nframes * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int slvgpu_eval_frames_host(slvgpu_plan *pl, const double *h_coords, int64_t nframes, double *h_out,
                                       int32_t *h_status) {
  if (!pl || !h_coords || !h_out || !h_status) return fail(SLVGPU_EINVAL, "null argument");
  if (nframes <= 0) return nframes == 0 ? 0 : fail(SLVGPU_EINVAL, "negative frame count");
  return eval_frames_host_any(pl, h_coords, 8, pl->dp.natoms_total, nullptr, 1.0, nframes, h_out, h_status);
}

extern "C" int slvgpu_eval_frames_host_f32(slvgpu_plan *pl, const float *h_coords, int64_t nframes, int32_t natoms_src,
                                           const int32_t *h_idx, double scale, double *h_out, int32_t *h_status) {
  if (!pl || !h_coords || !h_out || !h_status) return fail(SLVGPU_EINVAL, "null argument");
  if (nframes <= 0) return nframes == 0 ? 0 : fail(SLVGPU_EINVAL, "negative frame count");
  if (natoms_src < 1 || (!h_idx && natoms_src != pl->dp.natoms_total)) return fail(SLVGPU_EINVAL, "natoms_src does not match the plan (no index list given)");
  if (h_idx)
    for (int a = 0; a < pl->dp.natoms_total; ++a)
      if (h_idx[a] < 0 || h_idx[a] >= natoms_src) return fail(SLVGPU_EINVAL, "atom index outside the trajectory record");
  if (!(scale > 0.0)) return fail(SLVGPU_EINVAL, "scale must be positive");
  return eval_frames_host_any(pl, h_coords, 4, natoms_src, h_idx, scale, nframes, h_out, h_status);
}

extern "C" int64_t slvgpu_last_launches(const slvgpu_plan *pl) { return pl ? pl->last_launches : 0; }

extern "C" int slvgpu_get_pol_detail(slvgpu_plan *pl, int64_t frame, double *h_dipind, double *h_flds, double *h_rpol,
                                     double *h_avec, int32_t max_centres) {
  if (!pl || !pl->d_detail) return fail(SLVGPU_EINVAL, "plan was created without want_detail / pol");
  if (frame < 0 || frame >= pl->detail_frames) return fail(SLVGPU_EINVAL, "frame outside the last chunk");
  const int cc = pl->dp.cent_cap;
  std::vector<double> buf((size_t)cc * 12 + 1);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(buf.data(), pl->d_detail + (size_t)frame * ((size_t)cc * 12 + 1), buf.size() * sizeof(double), cudaMemcpyDeviceToHost));
  const int ncent = (int)buf[(size_t)cc * 12];
  const int n = max_centres < ncent ? max_centres : ncent;
  for (int c = 0; c < n; ++c)
    for (int d = 0; d < 3; ++d) {
      if (h_dipind) h_dipind[c * 3 + d] = buf[(size_t)c * 12 + d];
      if (h_flds) h_flds[c * 3 + d] = buf[(size_t)c * 12 + 3 + d];
      if (h_rpol) h_rpol[c * 3 + d] = buf[(size_t)c * 12 + 6 + d];
      if (h_avec) h_avec[c * 3 + d] = buf[(size_t)c * 12 + 9 + d];
    }
  return n;
}

// ------------------------------------------------------------------------------------------ Frag utilities
template <class F>
static int roundtrip(double *h, size_t n, F launch) {
  double *d = nullptr;
  CK(cudaMalloc(&d, n * sizeof(double)));
  cudaError_t e = cudaMemcpy(d, h, n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) { launch(d); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cudaMemcpy(h, d, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(SLVGPU_ECUDA, cudaGetErrorString(e));
  return 0;
}

extern "C" int slvgpu_kabsch(const double *h_ref, const double *h_target, int32_t n, double *h_rot, double *h_transl,
                             double *h_rms) {
  if (!h_ref || !h_target || n < 1 || !h_rot || !h_transl || !h_rms) return fail(SLVGPU_EINVAL, "bad argument");
  std::vector<double> buf((size_t)n * 6 + 13);
  memcpy(buf.data(), h_ref, (size_t)n * 3 * sizeof(double));
  memcpy(buf.data() + (size_t)n * 3, h_target, (size_t)n * 3 * sizeof(double));
  int rc = roundtrip(buf.data(), buf.size(), [&](double *d) { k_kabsch_one<<<1, 32>>>(d, d + (size_t)n * 3, n, d + (size_t)n * 6); });
  if (rc) return rc;
  memcpy(h_rot, buf.data() + (size_t)n * 6, 9 * sizeof(double));
  memcpy(h_transl, buf.data() + (size_t)n * 6 + 9, 3 * sizeof(double));
  *h_rms = buf[(size_t)n * 6 + 12];
  return 0;
}

static int with_rot(const double *h_rot, double **d_rot) {
  CK(cudaMalloc(d_rot, 12 * sizeof(double)));
  CK(cudaMemcpy(*d_rot, h_rot, 9 * sizeof(double), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int slvgpu_rotate_dma(double *h_dip, double *h_qad, double *h_oct, int64_t n, const double *h_rot) {
  if (n <= 0) return 0;
  if (!h_dip || !h_qad || !h_oct || !h_rot) return fail(SLVGPU_EINVAL, "null argument");
  std::vector<double> buf((size_t)n * 19);
  for (int64_t i = 0; i < n; ++i) {
    memcpy(&buf[i * 19], h_dip + i * 3, 3 * sizeof(double));
    memcpy(&buf[i * 19 + 3], h_qad + i * 6, 6 * sizeof(double));
    memcpy(&buf[i * 19 + 9], h_oct + i * 10, 10 * sizeof(double));
  }
  double *d_rot = nullptr;
  int rc = with_rot(h_rot, &d_rot);
  if (rc) return rc;
  rc = roundtrip(buf.data(), buf.size(), [&](double *d) { k_rot_dma<<<(unsigned)((n + 127) / 128), 128>>>(d, n, d_rot); });
  cudaFree(d_rot);
  if (rc) return rc;
  for (int64_t i = 0; i < n; ++i) {
    memcpy(h_dip + i * 3, &buf[i * 19], 3 * sizeof(double));
    memcpy(h_qad + i * 6, &buf[i * 19 + 3], 6 * sizeof(double));
    memcpy(h_oct + i * 10, &buf[i * 19 + 9], 10 * sizeof(double));
  }
  return 0;
}

extern "C" int slvgpu_rotate_pol(double *h_pol, int64_t n, const double *h_rot) {
  if (n <= 0) return 0;
  if (!h_pol || !h_rot) return fail(SLVGPU_EINVAL, "null argument");
  double *d_rot = nullptr;
  int rc = with_rot(h_rot, &d_rot);
  if (rc) return rc;
  rc = roundtrip(h_pol, (size_t)n * 9, [&](double *d) { k_rot_pol<<<(unsigned)((n + 127) / 128), 128>>>(d, n, d_rot); });
  cudaFree(d_rot);
  return rc;
}

extern "C" int slvgpu_rotate_vec(double *h_vec, int64_t n, const double *h_rot, const double *h_transl) {
  if (n <= 0) return 0;
  if (!h_vec || !h_rot) return fail(SLVGPU_EINVAL, "null argument");
  double *d_rot = nullptr;
  int rc = with_rot(h_rot, &d_rot);
  if (rc) return rc;
  double t[3] = {0, 0, 0};
  if (h_transl) memcpy(t, h_transl, sizeof(t));
  cudaMemcpy(d_rot + 9, t, sizeof(t), cudaMemcpyHostToDevice);
  rc = roundtrip(h_vec, (size_t)n * 3, [&](double *d) { k_rot_vec<<<(unsigned)((n + 127) / 128), 128>>>(d, n, d_rot); });
  cudaFree(d_rot);
  return rc;
}

extern "C" int slvgpu_rotate_vecl(double *h_vec, int64_t nrows, int32_t nbasis, const int32_t *h_typ, const double *h_rot) {
  if (nrows <= 0 || nbasis <= 0) return 0;
  if (!h_vec || !h_typ || !h_rot) return fail(SLVGPU_EINVAL, "null argument");
  // shell starts: the first function of each p triple / d sextet
  std::vector<int> start;
  for (int k = 0; k < nbasis;) {
    if (h_typ[k] == 1) { start.push_back(k); start.push_back(1); k += 3; }
    else if (h_typ[k] == 2) { start.push_back(k); start.push_back(2); k += 6; }
    else ++k;
  }
  if (start.empty()) return 0;
  double *d_rot = nullptr; int *d_start = nullptr;
  int rc = with_rot(h_rot, &d_rot);
  if (rc) return rc;
  CK(cudaMalloc(&d_start, start.size() * sizeof(int)));
  CK(cudaMemcpy(d_start, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice));
  const int nsh = (int)start.size() / 2;
  rc = roundtrip(h_vec, (size_t)nrows * nbasis, [&](double *d) {
    const int64_t work = nrows * nsh;
    k_rot_vecl<<<(unsigned)((work + 127) / 128), 128>>>(d, nrows, nbasis, d_start, nsh, d_rot);
  });
  cudaFree(d_rot); cudaFree(d_start);
  return rc;
}

// ------------------------------------------------------------------------------------------ spectra (row N2)
extern "C" int slvgpu_tcf(const double *h_a, const double *h_b, int64_t n, int64_t k, int32_t demean, double *h_tcf) {
  if (!h_a || !h_b || !h_tcf || n < 1 || k < 1 || k > n) return fail(SLVGPU_EINVAL, "bad argument");
  double *d = nullptr;
  CK(cudaMalloc(&d, (size_t)(2 * n + k + 2) * sizeof(double)));
  double *da = d, *db = d + n, *dt = d + 2 * n, *dm = dt + k;
  cudaError_t e = cudaMemcpy(da, h_a, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(db, h_b, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    if (demean) {
      k_mean<<<1, 1024>>>(da, n, dm); k_shift<<<(unsigned)((n + 255) / 256), 256>>>(da, n, dm);
      k_mean<<<1, 1024>>>(db, n, dm + 1); k_shift<<<(unsigned)((n + 255) / 256), 256>>>(db, n, dm + 1);
    }
    k_tcf<<<(unsigned)((k + TCF_LAGS - 1) / TCF_LAGS), SPEC_THREADS>>>(da, db, n, k, dt);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(h_tcf, dt, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(SLVGPU_ECUDA, cudaGetErrorString(e));
  return 0;
}

extern "C" int slvgpu_expres(const double *h_f, int64_t n, double dt, int64_t ndel, double *h_re, double *h_im) {
  if (!h_f || !h_re || !h_im || n < 2 || ndel < 0 || ndel >= n) return fail(SLVGPU_EINVAL, "bad argument");
  double *d = nullptr;
  CK(cudaMalloc(&d, (size_t)(2 * n + 1 + 2 * (ndel + 1)) * sizeof(double)));
  double *df = d, *dC = d + n, *dre = dC + n + 1, *dim = dre + ndel + 1;
  cudaError_t e = cudaMemcpy(df, h_f, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    k_trapz_prefix<<<1, 1024>>>(df, n, dt, dC);
    k_expres<<<(unsigned)(ndel + 1), SPEC_THREADS>>>(dC, n, ndel, dre, dim);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(h_re, dre, (size_t)(ndel + 1) * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(h_im, dim, (size_t)(ndel + 1) * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(SLVGPU_ECUDA, cudaGetErrorString(e));
  return 0;
}

extern "C" int slvgpu_trig_transform(const double *h_f, int64_t n, double y0, double dy, const double *h_x, int64_t m,
                                     int32_t kind, double *h_out) {
  if (!h_f || !h_x || !h_out || n < 2 || m < 1 || (kind != 0 && kind != 1)) return fail(SLVGPU_EINVAL, "bad argument");
  double *d = nullptr;
  CK(cudaMalloc(&d, (size_t)(n + 2 * m) * sizeof(double)));
  double *df = d, *dx = d + n, *dout = dx + m;
  cudaError_t e = cudaMemcpy(df, h_f, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dx, h_x, (size_t)m * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    k_trig_transform<<<(unsigned)m, SPEC_THREADS>>>(df, n, y0, dy, dx, kind, dout);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(h_out, dout, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(SLVGPU_ECUDA, cudaGetErrorString(e));
  return 0;
}

extern "C" int slvgpu_hist(const double *h_x, int64_t n, double lo, double hi, int32_t bins, int64_t *h_counts) {
  if (!h_x || !h_counts || n < 1 || bins < 1 || bins > 8192 || !(hi > lo)) return fail(SLVGPU_EINVAL, "bad argument");
  double *dx = nullptr; unsigned long long *dc = nullptr;
  CK(cudaMalloc(&dx, (size_t)n * sizeof(double)));
  cudaError_t e = cudaMalloc(&dc, (size_t)bins * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(dc, 0, (size_t)bins * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemcpy(dx, h_x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    const int64_t nb = (n + SPEC_THREADS - 1) / SPEC_THREADS;
    k_hist<<<(unsigned)(nb < 592 ? nb : 592), SPEC_THREADS, (size_t)bins * sizeof(unsigned int)>>>(dx, n, lo, hi, bins, dc);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(h_counts, dc, (size_t)bins * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree(dx); if (dc) cudaFree(dc);
  if (e != cudaSuccess) return fail(SLVGPU_ECUDA, cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------ XTC trajectories (row N1)
extern "C" int slvgpu_xtc_scan(const char *path, int64_t *nframes, int32_t *natoms) {
  if (!path || !nframes || !natoms) return fail(SLVGPU_EINVAL, "null argument");
  slvxtc::File f; f.fh = fopen(path, "rb");
  if (!f.fh) return fail(SLVGPU_EINVAL, std::string("cannot open ") + path);
  std::vector<unsigned char> buf; std::string err;
  int64_t n = 0; int32_t na = 0, na0 = -1, step; float time, box[9];
  for (;;) {
    const int rc = slvxtc::read_frame(f, na, step, time, box, nullptr, buf, err);
    if (rc == 0) break;
    if (rc < 0) return fail(SLVGPU_EINVAL, std::string(path) + ": " + err);
    if (na0 < 0) na0 = na;
    if (na != na0) return fail(SLVGPU_EINVAL, std::string(path) + ": atom count changes between frames");
    ++n;
  }
  *nframes = n; *natoms = na0 < 0 ? 0 : na0;
  return 0;
}

extern "C" int slvgpu_xtc_read(const char *path, int64_t first, int64_t count, float *h_xyz, int32_t *h_step, float *h_time,
                               float *h_box) {
  if (!path || !h_xyz || first < 0 || count < 0) return fail(SLVGPU_EINVAL, "bad argument");
  slvxtc::File f; f.fh = fopen(path, "rb");
  if (!f.fh) return fail(SLVGPU_EINVAL, std::string("cannot open ") + path);
  std::vector<unsigned char> buf; std::string err;
  int32_t na = 0, step; float time, box[9];
  for (int64_t k = 0; k < first + count; ++k) {
    // (k - first) * na: na is the atom count of the previous frame, and slvgpu_xtc_scan guarantees it does not change
    float *dst = k < first ? nullptr : (k == first ? h_xyz : h_xyz + (size_t)(k - first) * na * 3);
    const int rc = slvxtc::read_frame(f, na, step, time, box, dst, buf, err);
    if (rc == 0) return fail(SLVGPU_EINVAL, std::string(path) + ": fewer frames than requested");
    if (rc < 0) return fail(SLVGPU_EINVAL, std::string(path) + ": " + err);
    if (k >= first) {
      if (h_step) h_step[k - first] = step;
      if (h_time) h_time[k - first] = time;
      if (h_box) memcpy(h_box + (size_t)(k - first) * 9, box, sizeof(box));
    }
  }
  return 0;
}
