// slv_common.cuh -- device plan view, per-frame scratch layout and small block-level helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/slvgpu.h"
#include "slv_math.cuh"

namespace slv {

// Device view of the plan (passed by value to kernels).
struct DevPlan {
  int ntypes, ninst, natoms_total, flags;
  double ccut, pcut, ecut;
  const int *type_meta;   // [ntypes][SLVGPU_NMETA]
  const int *sol_meta;    // [SLVGPU_NSOL]
  const int *itab;
  const double *dtab;
  const int *inst_type;   // [ninst]
  const int *inst_atom0;  // [ninst+1]
  int pol_maxit; double pol_tol;
  int list_cap;           // capacity of the per-frame pol/rep instance list
  int cent_cap;           // capacity (centres) of the per-CTA polarisation workspace
  int site_cap;           // capacity (DMTP sites) of the per-CTA polarisation workspace
  int max_npol, max_ndma; // largest npol / ndma over all fragment types
  double glw[12];         // Gauss-Legendre weights W_k of the imaginary-frequency quadrature (slvpar.py:989-1019)
};

// Per-frame hand-over from the superimposition/electrostatics kernel to the pol/disp/rep kernels:
// the solute transform and the compacted list of solvent instances inside pcut or ecut.
struct FrameLists {
  double *sol_xf;    // [chunk][12]       solute rot(9) + transl(3)
  int *nlist;        // [chunk]           entries in the list
  int *list_inst;    // [chunk][list_cap] instance index
  int *list_flag;    // [chunk][list_cap] bit0 = in pcut, bit1 = in ecut
  double *list_xf;   // [chunk][list_cap][12]
};

__device__ __forceinline__ const int *tmeta(const DevPlan &p, int t) { return p.type_meta + t * SLVGPU_NMETA; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Deterministic block reductions (fixed tree: lanes by xor-shuffle, then warps in index order).
// `scratch` must hold >= 32 doubles; result valid in every thread.
__device__ __forceinline__ double block_sum(double v, double *scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < nw; ++i) r += scratch[i];
  return r;
}
__device__ __forceinline__ double block_max(double v, double *scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double r = scratch[0];
  for (int i = 1; i < nw; ++i) r = fmax(r, scratch[i]);
  return r;
}

// Single-barrier variants for reductions that alternate between two scratch arrays inside a loop: the barrier that
// protects `scratch` from the readers of its previous use is whatever barrier the caller has between the two uses
// (k_pol's BiCG iteration: pq -> A, (rho, max) -> B, with the other reduction's barrier in between).
__device__ __forceinline__ double block_sum_1b(double v, double *scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < nw; ++i) r += scratch[i];
  return r;
}
__device__ __forceinline__ void block_sum_max_1b(double &s, double &m, double *scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  s = warp_sum(s); m = warp_max(m);
  if (lane == 0) { scratch[w] = s; scratch[32 + w] = m; }
  __syncthreads();
  double rs = 0.0, rm = scratch[32];
  for (int i = 0; i < nw; ++i) { rs += scratch[i]; rm = fmax(rm, scratch[32 + i]); }
  s = rs; m = rm;
}

// one sum and one max in a single pass (two barriers); `scratch` must hold >= 64 doubles
__device__ __forceinline__ void block_sum_max(double &s, double &m, double *scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  s = warp_sum(s); m = warp_max(m);
  __syncthreads();
  if (lane == 0) { scratch[w] = s; scratch[32 + w] = m; }
  __syncthreads();
  double rs = 0.0, rm = scratch[32];
  for (int i = 0; i < nw; ++i) { rs += scratch[i]; rm = fmax(rm, scratch[32 + i]); }
  s = rs; m = rm;
}

// Kabsch fit of one fragment instance: parameter atoms `supl` (parameter order) onto the frame atoms
// they map to through `reord`.  Follows Frag.sup / SVDSuperimposer (solvshift/slvpar.py:641-679):
// one-atom fragments get the identity and a pure shift (668-671).
__device__ inline void fit_instance(const DevPlan &p, const int *tm, const double *__restrict__ xyz /* instance atoms */,
                                    double rot[9], double tr[3], double &rms) {
  const int natoms = tm[SLVGPU_M_NATOMS], nsupl = tm[SLVGPU_M_NSUPL];
  const int *supl = p.itab + tm[SLVGPU_M_IO_SUPL];
  const int *reord = p.itab + tm[SLVGPU_M_IO_REORD];
  const double *pos = p.dtab + tm[SLVGPU_M_O_POS];
  if (natoms == 1) {
    rot[0] = rot[4] = rot[8] = 1.0; rot[1] = rot[2] = rot[3] = rot[5] = rot[6] = rot[7] = 0.0;
    const int md = reord[0];
#pragma unroll
    for (int d = 0; d < 3; ++d) tr[d] = (md >= 0 ? xyz[md * 3 + d] : 0.0) - pos[d];
    rms = 0.0;
    return;
  }
  double c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
  for (int k = 0; k < nsupl; ++k) {
    const int ip = supl[k], md = reord[ip];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      c1[d] += pos[ip * 3 + d];
      c2[d] += (md >= 0 ? xyz[md * 3 + d] : 0.0);      // absent atoms are zero rows (solefp.py:1633-1645)
    }
  }
  const double inv = 1.0 / (double)nsupl;
#pragma unroll
  for (int d = 0; d < 3; ++d) { c1[d] *= inv; c2[d] *= inv; }
  double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, gsum = 0.0;
  for (int k = 0; k < nsupl; ++k) {
    const int ip = supl[k], md = reord[ip];
    double c[3], r[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      c[d] = pos[ip * 3 + d] - c1[d];
      r[d] = (md >= 0 ? xyz[md * 3 + d] : 0.0) - c2[d];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) a[i * 3 + j] += c[i] * r[j];
    gsum += c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  }
  kabsch_rot(a, gsum, rot);
  double t1[3];
  rot_vec(rot, c1, t1);
#pragma unroll
  for (int d = 0; d < 3; ++d) tr[d] = c2[d] - t1[d];
  double ss = 0.0;
  for (int k = 0; k < nsupl; ++k) {
    const int ip = supl[k], md = reord[ip];
    double y[3];
    rot_vec(rot, pos + ip * 3, y);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double dd = y[d] + tr[d] - (md >= 0 ? xyz[md * 3 + d] : 0.0);
      ss += dd * dd;
    }
  }
  rms = sqrt(ss * inv);
}

}  // namespace slv
