/* slvgpu.h -- C ABI of the B200-native SolEFP frequency-shift engine.
 *
 * The reference (globulion/slv) has no C ABI of its own: its native boundary is what f2py
 * generates for the Fortran kernels plus the Python classes above them.  Each entry point
 * below names the reference interface it replaces (file:line under /root/reference).
 *
 * Conventions: plain pointers and sizes only; FP64 arrays are C-contiguous; all lengths in
 * elements; every function returns 0 on success or a negative SLVGPU_E* code, and
 * slvgpu_last_error() returns a message for the calling thread.  Device buffers passed in
 * are owned by the caller (PyTorch in the shipped host code); the library allocates only its
 * own plan tables and scratch.  One plan per device / rank.  Atomic units throughout
 * (Bohr, Hartree); the cm^-1 conversion is the caller's (solvshift/solefp.py:1571-1579).
 */
#ifndef SLVGPU_H
#define SLVGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLVGPU_OK            0
#define SLVGPU_EINVAL       -1
#define SLVGPU_ECUDA        -2
#define SLVGPU_ENOMEM       -3

/* evaluation flags (EFP.__init__ keywords, solvshift/solefp.py:52-83) */
#define SLVGPU_ELECT   1   /* elect  */
#define SLVGPU_EA      2   /* ea     */
#define SLVGPU_CORR    4   /* corr   */
#define SLVGPU_POL     8   /* pol    */
#define SLVGPU_DISP   16   /* disp   */
#define SLVGPU_REP    32   /* rep    */
#define SLVGPU_GLOBAL 64   /* freq=False: EFP2 interaction ENERGIES of the whole cluster, every fragment pair, no cut-offs
                            * (EFP._eval_mode_global, solvshift/solefp.py:905-1009): the electrostatic energy comes back in
                            * the ELE_MEA column, the polarisation energy (SOLPOL) in POL_MEA, both in Hartree             */

/* columns of the per-frame output row (a.u.) */
enum {
  SLVGPU_ELE_MEA = 0, SLVGPU_ELE_EA = 1, SLVGPU_COR_MEA = 2, SLVGPU_COR_EA = 3,
  SLVGPU_POL_MEA = 4, SLVGPU_REP_MEA = 5, SLVGPU_DIS_MEA = 6, SLVGPU_DIS_ISO = 7,
  SLVGPU_RMS_C = 8, SLVGPU_RMS_SMAX = 9, SLVGPU_RMS_AVE = 10,
  SLVGPU_N_ELECT = 11, SLVGPU_N_POL = 12, SLVGPU_N_REP = 13,
  SLVGPU_POL_ITERS = 14, SLVGPU_POL_RESID = 15,
  SLVGPU_POL_ADD = 16,   /* part of POL_MEA from the additive two-body solves of removable fragments */
  SLVGPU_POL_NCENT = 17, /* polarisable centres of the many-body solve                                  */
  SLVGPU_POL_NSITE = 18, /* DMTP sites contributing fields                                              */
  SLVGPU_POL_PAIRS = 19, /* ordered centre pairs of different molecules = dipole-tensor applies per sweep */
  /* SolCAMM + corrections against the non-EFP environment (instances of types flagged SLVGPU_M_ENV): EFP.eval_dma,
   * solvshift/solefp.py:151-178, 798-903 -- no cut-off, never part of the polarisation / dispersion / repulsion layers */
  SLVGPU_ENV_ELE_MEA = 20, SLVGPU_ENV_ELE_EA = 21, SLVGPU_ENV_COR_MEA = 22, SLVGPU_ENV_COR_EA = 23,
  SLVGPU_NOUT = 24
};

/* per-frame status bits */
#define SLVGPU_ST_NONFINITE      1   /* a shift term is NaN/Inf                      */
#define SLVGPU_ST_POL_NOCONV     2   /* induced-dipole iteration did not converge    */
#define SLVGPU_ST_SINGULAR       4   /* degenerate superimposition                   */
#define SLVGPU_ST_OVERFLOW       8   /* more centres / sites in a layer than the plan's workspace caps */

/* integer columns of the per-fragment-type table `type_meta` [ntypes][SLVGPU_NMETA] */
enum {
  SLVGPU_M_NATOMS = 0, SLVGPU_M_NDMA, SLVGPU_M_NPOL, SLVGPU_M_NMOS, SLVGPU_M_NBASIS, SLVGPU_M_NSUPL,
  SLVGPU_M_IO_SUPL,   /* itab: [nsupl] parameter-order atom indices used for the Kabsch fit     */
  SLVGPU_M_IO_REORD,  /* itab: [natoms] MD atom (within the instance) of parameter atom i, -1 = absent */
  SLVGPU_M_O_POS,     /* dtab: [natoms][3] parameter geometry                                    */
  SLVGPU_M_O_RDMA,    /* dtab: [ndma][3]                                                         */
  SLVGPU_M_O_MOM,     /* dtab: [ndma][20] q, mu(3), Theta(6), Omega(10) -- Buckingham traceless  */
  SLVGPU_M_O_RPOL,    /* dtab: [npol][3]                                                         */
  SLVGPU_M_O_POL,     /* dtab: [npol][9] alpha                                                   */
  SLVGPU_M_O_POLINV,  /* dtab: [npol][9] alpha^-1                                                */
  SLVGPU_M_O_DPOLI,   /* dtab: [npol][12][9] alpha(i w_k)                                        */
  SLVGPU_M_O_LMOC,    /* dtab: [nmos][3]                                                         */
  SLVGPU_M_O_FOCK,    /* dtab: [nmos][nmos]                                                      */
  SLVGPU_M_O_VECL,    /* dtab: [nmos][nbasis]                                                    */
  SLVGPU_M_O_ZNUC,    /* dtab: [natoms] atomic numbers as doubles                                */
  SLVGPU_M_IO_BFC,    /* itab: [nbasis] atom of each basis function                              */
  SLVGPU_M_IO_BFL,    /* itab: [nbasis][3] Cartesian powers                                      */
  SLVGPU_M_IO_BFNP,   /* itab: [nbasis] number of primitives                                     */
  SLVGPU_M_O_BFEXP,   /* dtab: [nbasis][6] exponents                                             */
  SLVGPU_M_O_BFCOEF,  /* dtab: [nbasis][6] normalised contraction coefficients                   */
  SLVGPU_M_REMOVABLE, /* 1: polarisation of this type is treated pair-wise (remove_clashes); 2: the type is part of the
                       * non-EFP electrostatic environment of EFP.eval_dma (point multipoles riding on frame atoms)   */
  SLVGPU_M_NSUBSH,   /* number of basis sub-shells (an SP shell counts as an S and a P sub-shell)  */
  SLVGPU_M_IO_SUBSH, /* itab: [nsubsh][4] atom, first function, angular momentum, primitives      */
  SLVGPU_M_IO_SHPAIR, /* itab: exchange-repulsion integral work list against the solute: [24] class offsets, then the
                       * work items a * nsubsh + b (a: solute sub-shell, b: this type's), grouped by the 23
                       * (L_bra, L_ket, bra component) specialisations of slv_ints.cuh SLV_SUBSHELL_CLASSES,
                       * heavy primitive-pair counts first inside a class                                    */
  SLVGPU_M_IO_SHITEM, /* itab (16-byte aligned): [items][8] per work item of IO_SHPAIR, same order: 3 * solute atom, first
                       * solute function, 3 * fragment atom, first fragment function, first record in O_PPTAB, records   */
  SLVGPU_M_O_PPTAB,   /* dtab (16-byte aligned): [records][6] primitive pairs of every (solute sub-shell, sub-shell) pair,
                       * thr = 46 (a+b)/(ab), mu, c_a c_b (pi/(a+b))^1.5, 1/(a+b), a, b; thr descending within a pair    */
  SLVGPU_M_IO_BFBLK,  /* itab: [nbfblk][2] first function and angular momentum of every s / p / d block of the basis (the units
                       * VECROT rotates the AO->LMO coefficient rows in)                                                  */
  SLVGPU_M_NBFBLK,
  SLVGPU_NMETA = 32
};

/* offsets (into dtab) of the solute's mode-contracted tables `sol_meta` [SLVGPU_NSOL] */
enum {
  SLVGPU_S_SMEA = 0,  /* [ndma][20]  -sum_m g_m dma1_m            (SolCAMM, mechanical)          */
  SLVGPU_S_SEA,       /* [ndma][20]  dma2_j / den                  (SolCAMM, electronic)         */
  SLVGPU_S_CMEA,      /* [ndma][3]   -q_x sum_m g_m L^m_x          (correction, mechanical)      */
  SLVGPU_S_CEA,       /* [ndma][3]   2 q1^j_x L^j_x / den          (correction, electronic)      */
  SLVGPU_S_DPOLI1,    /* [npol][12][9] sum_m g_m dpoli1_m                                        */
  SLVGPU_S_LMOC1,     /* [npol][3]   sum_m g_m lmoc1_m                                           */
  SLVGPU_S_DPOL1,     /* [npol][9]   sum_m g_m dpol1_m                                           */
  SLVGPU_S_LVEC,      /* [natoms][3] sum_m g_m lvec_m                                            */
  SLVGPU_S_DMA1,      /* [ndma][20]  sum_m g_m dma1_m (traceless)                                */
  SLVGPU_S_FOCK1,     /* [nmos][nmos] sum_m g_m fock1_m                                          */
  SLVGPU_S_VECL1,     /* [nmos][nbasis] sum_m g_m vecl1_m                                        */
  SLVGPU_NSOL = 16
};

typedef struct slvgpu_plan_desc {
  int32_t ntypes;           /* fragment types (len(bsm))                                        */
  int32_t ninst;            /* fragment instances per frame, instance 0 = solute                */
  int32_t natoms_total;     /* atoms per frame (rows of frame[idx])                             */
  int32_t flags;            /* SLVGPU_ELECT | ...                                               */
  double  ccut, pcut, ecut; /* cut-offs in Bohr; strict '<' on centre distances (MOLLST)        */
  const int32_t *type_meta; /* [ntypes][SLVGPU_NMETA]                                           */
  const int32_t *sol_meta;  /* [SLVGPU_NSOL]                                                    */
  const int32_t *itab;  int64_t n_itab;
  const double  *dtab;  int64_t n_dtab;
  const int32_t *inst_type;   /* [ninst] index into type tables (EFP.set `ind`)                 */
  const int32_t *inst_atom0;  /* [ninst] first atom of the instance in the frame (cumsum nmol)  */
  int32_t max_chunk;        /* frames processed per internal pass (scratch is sized by it)      */
  int32_t pol_maxit;        /* cap on induced-dipole iterations                                 */
  double  pol_tol;          /* relative max-norm update at which the iteration stops            */
  int32_t pol_cent_cap;     /* workspace cap on polarisable centres inside pcut (0 = all)       */
  int32_t want_detail;      /* keep induced dipoles / fields of the last call for the getters   */
} slvgpu_plan_desc;

typedef struct slvgpu_plan slvgpu_plan;

/* Build the device-resident plan for one trajectory layout.  Replaces the per-frame Python work of
 * EFP.set / EFP._update / Frag.copy (solvshift/solefp.py:223-277, 738-787; slvpar.py:736-738):
 * parameters are uploaded once instead of deep-copied every frame. */
int slvgpu_plan_create(const slvgpu_plan_desc *desc, slvgpu_plan **out);
void slvgpu_plan_destroy(slvgpu_plan *plan);

/* Evaluate nframes frames held in DEVICE memory.  d_coords [nframes][natoms_total][3] FP64 Bohr,
 * d_out [nframes][SLVGPU_NOUT] FP64, d_status [nframes] int32; `stream` is a cudaStream_t (0 = legacy
 * default).  Replaces, per frame, EFP.set + EFP.eval + get_shifts/get_rms*
 * (solvshift/solefp.py:223, 93-149, 1011-1629; natives: solpol2.f MOLLST/SFTPLI, libbbg clemtp
 * sdmtpm/sdmtpe/dmtcor/sdisp6/tdisp6, shftex.f SHFTEX, PyQuante getSAB/getTAB/getSA1B/getTA1B).
 * Work is enqueued on `stream` and the call returns without waiting for it, except when the plan has
 * SLVGPU_REP: the exchange-repulsion batches are sized from per-frame fragment counts read back on the host, so
 * the call synchronises `stream` once per chunk of max_chunk frames (right after the per-frame kernels of the chunk; it is
 * therefore not graph-capturable); the polarisation kernel and the batches also use plan-owned streams (one of high
 * priority), all joined back into `stream` before the call returns. */
int slvgpu_eval_frames(slvgpu_plan *plan, const double *d_coords, int64_t nframes,
                       double *d_out, int32_t *d_status, void *stream);

/* Same through HOST buffers (pinned or pageable): copies coordinates in, results out; this is the call the
 * MD frame loop of util/slv_md-run(_nopp) (116-155, 202-209) maps onto. */
int slvgpu_eval_frames_host(slvgpu_plan *plan, const double *h_coords, int64_t nframes,
                            double *h_out, int32_t *h_status);

/* Same for float32 trajectory records as the MD engines write them (DCD, XTC, NetCDF hold float32): h_coords
 * [nframes][natoms_src][3] in the file's units; the gather frame[idx] (h_idx [natoms_total] record atom of every plan atom,
 * NULL = identity) and the conversion to Bohr (x scale; util/slv_md-run_nopp:205 `coords * AngstromToBohr`, md.py idx slice)
 * happen on the device, so half the bytes cross the host link and no host pass touches the coordinates.  The rows are
 * bit-identical to slvgpu_eval_frames_host on (double)h_coords[..] * scale. */
int slvgpu_eval_frames_host_f32(slvgpu_plan *plan, const float *h_coords, int64_t nframes, int32_t natoms_src,
                                const int32_t *h_idx, double scale, double *h_out, int32_t *h_status);

/* Number of kernel launches issued by the last slvgpu_eval_frames* call on this plan. */
int64_t slvgpu_last_launches(const slvgpu_plan *plan);

/* Per-kernel device timing (CUDA events recorded on the launch stream around every kernel of
 * slvgpu_eval_frames*): kernels 0 = superimposition+electrostatics, 1 = dispersion, 2 = polarisation,
 * 3 = exchange-repulsion.  get_timing synchronises and returns accumulated ms and launch counts since
 * set_timing(plan, 1). */
int slvgpu_set_timing(slvgpu_plan *plan, int on);
int slvgpu_get_timing(slvgpu_plan *plan, double *ms4, int64_t *launches4);

/* Optional per-frame detail of the LAST evaluated frame range: induced dipoles and fields at the
 * polarisable centres and the solvatochromic induced dipoles (EFP.get_dipind / get_fields / get_avec,
 * solvshift/solefp.py:303-313).  Buffers are
 * [ncent][3]; returns the number of centres written, or a negative error. */
int slvgpu_get_pol_detail(slvgpu_plan *plan, int64_t frame, double *h_dipind, double *h_flds,
                          double *h_rpol, double *h_avec, int32_t max_centres);

/* Kabsch superimposition of one fragment (Frag.sup, solvshift/slvpar.py:641-679, libbbg
 * SVDSuperimposer): host pointers, n atoms each; rot[9] row-major with x' = x.rot + transl. */
int slvgpu_kabsch(const double *h_ref, const double *h_target, int32_t n,
                  double *h_rot, double *h_transl, double *h_rms);

/* Rotate the distributed-multipole tensors of nsets sets (host pointers, in place):
 * dip [n][3], qad [n][6], oct [n][10] -- ROTDMA/ROTDM1 (solvshift/src/efprot.f:286-1211);
 * and 3x3 tensors alpha' = R^T alpha R (solvshift/slvpar.py:707-722) for pol [npol][9]. */
int slvgpu_rotate_dma(double *h_dip, double *h_qad, double *h_oct, int64_t n, const double *h_rot);
int slvgpu_rotate_pol(double *h_pol, int64_t n, const double *h_rot);
/* rows of vectors x' = x.rot (+ transl when non-NULL): positions, lvec, lmoc1 */
int slvgpu_rotate_vec(double *h_vec, int64_t n, const double *h_rot, const double *h_transl);
/* AO->LMO coefficient rotation, VECROT/VC1ROT (solvshift/src/efprot.f:22-283): vec [nrows][nbasis],
 * typ [nbasis] = 0 s, 1 p, 2 d (first component of each p triple / d sextet marks the shell). */
int slvgpu_rotate_vecl(double *h_vec, int64_t nrows, int32_t nbasis, const int32_t *h_typ, const double *h_rot);

/* Post-processing of the shift series, host pointers:
 * slvgpu_tcf: TCF(i) = 1/(n-i) sum_j a[j] b[j+i], i < k -- GTCF (util/lib/genres.f:1-23); with demean != 0 the means
 * are removed first, which makes it the `cf` of util/slv_calc-tcf:17-28 (cf(y1,y2)[i] = tcf(a=y2, b=y1)[i]).
 * slvgpu_expres: classical line-shape function of util/slv_calc-expres:73-101: re/im [ndel+1]. */
int slvgpu_tcf(const double *h_a, const double *h_b, int64_t n, int64_t k, int32_t demean, double *h_tcf);
int slvgpu_expres(const double *h_f, int64_t n, double dt, int64_t ndel, double *h_re, double *h_im);
/* out[i] = trapezoid integral over y of f(y) cos(x_i y) (kind 0) or f(y) sin(x_i y) (kind 1) on the uniform grid
 * y_j = y0 + j dy, j < n, for m abscissae x_i: the frequency and time integrals of util/slv_calc-ftir (spectral
 * density :223-226, line-broadening functions :252-259, classical line shape :399-407). */
int slvgpu_trig_transform(const double *h_f, int64_t n, double y0, double dy, const double *h_x, int64_t m,
                          int32_t kind, double *h_out);
/* counts of h_x in `bins` (<= 8192) equal-width bins over [lo, hi], last bin closed, as numpy / pylab.hist do for
 * util/slv_hist:60-61. */
int slvgpu_hist(const double *h_x, int64_t n, double lo, double hi, int32_t bins, int64_t *h_counts);

/* GROMACS XTC trajectories (the `gromacs` branch of util/slv_md-run:117-152, which goes through MDAnalysis): host-side
 * decoder of the xdr3dcoord compression.  slvgpu_xtc_scan counts the frames and returns the atom count; slvgpu_xtc_read
 * decodes `count` frames starting at frame `first` into h_xyz [count][natoms][3] float32 nm records (the precision the file
 * holds -- feed them to slvgpu_eval_frames_host_f32 with scale = 10 * AngstromToBohr); h_step / h_time [count] and h_box
 * [count][9] may be NULL. */
int slvgpu_xtc_scan(const char *path, int64_t *nframes, int32_t *natoms);
int slvgpu_xtc_read(const char *path, int64_t first, int64_t count, float *h_xyz, int32_t *h_step, float *h_time,
                    float *h_box);

const char *slvgpu_last_error(void);
const char *slvgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif
