// k_util.cuh -- single-fragment utilities behind Frag.sup / Frag.rotate (solvshift/slvpar.py:641-734):
// Kabsch fit of one fragment and rotation of parameter tensors (ROTDMA/ROTDM1, VECROT/VC1ROT in
// solvshift/src/efprot.f).  Bandwidth-trivial; they exist so the host API has no CPU arithmetic path.
#pragma once
#include "slv_common.cuh"

namespace slv {

// ref[n][3], tgt[n][3] -> res[13] = rot(9), transl(3), rms   (SVDSuperimposer semantics)
__global__ void k_kabsch_one(const double *ref, const double *tgt, int n, double *res) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tr[3], rms = 0.0;
  if (n == 1) {
    for (int d = 0; d < 3; ++d) tr[d] = tgt[d] - ref[d];
  } else {
    double c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
    for (int k = 0; k < n; ++k)
      for (int d = 0; d < 3; ++d) { c1[d] += ref[k * 3 + d]; c2[d] += tgt[k * 3 + d]; }
    for (int d = 0; d < 3; ++d) { c1[d] /= n; c2[d] /= n; }
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, gsum = 0.0;
    for (int k = 0; k < n; ++k)
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) a[i * 3 + j] += (ref[k * 3 + i] - c1[i]) * (tgt[k * 3 + j] - c2[j]);
        gsum += (ref[k * 3 + i] - c1[i]) * (ref[k * 3 + i] - c1[i]) + (tgt[k * 3 + i] - c2[i]) * (tgt[k * 3 + i] - c2[i]);
      }
    kabsch_rot(a, gsum, rot);
    double t1[3];
    rot_vec(rot, c1, t1);
    for (int d = 0; d < 3; ++d) tr[d] = c2[d] - t1[d];
    double ss = 0.0;
    for (int k = 0; k < n; ++k) {
      double y[3];
      rot_vec(rot, ref + k * 3, y);
      for (int d = 0; d < 3; ++d) { const double dd = y[d] + tr[d] - tgt[k * 3 + d]; ss += dd * dd; }
    }
    rms = sqrt(ss / n);
  }
  for (int d = 0; d < 9; ++d) res[d] = rot[d];
  for (int d = 0; d < 3; ++d) res[9 + d] = tr[d];
  res[12] = rms;
}

// rows of 19 doubles: dip3, qad6, oct10 (primitive or traceless -- rotation is linear)
__global__ void k_rot_dma(double *rows, int64_t n, const double *rot) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double R[9], in[19], o[19];
  for (int d = 0; d < 9; ++d) R[d] = rot[d];
  for (int d = 0; d < 19; ++d) in[d] = rows[i * 19 + d];
  rot_vec(R, in, o); rot_quad(R, in + 3, o + 3); rot_oct(R, in + 9, o + 9);
  for (int d = 0; d < 19; ++d) rows[i * 19 + d] = o[d];
}

__global__ void k_rot_pol(double *a, int64_t n, const double *rot) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double R[9], in[9], o[9];
  for (int d = 0; d < 9; ++d) { R[d] = rot[d]; in[d] = a[i * 9 + d]; }
  rot_mat(R, in, o);
  for (int d = 0; d < 9; ++d) a[i * 9 + d] = o[d];
}

__global__ void k_rot_vec(double *v, int64_t n, const double *rot /* 9 + transl 3 */) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double R[9], in[3], o[3];
  for (int d = 0; d < 9; ++d) R[d] = rot[d];
  for (int d = 0; d < 3; ++d) in[d] = v[i * 3 + d];
  rot_vec(R, in, o);
  for (int d = 0; d < 3; ++d) v[i * 3 + d] = o[d] + rot[9 + d];
}

// AO->LMO coefficient rows: p triples rotate as vectors; d sextets (xx yy zz xy xz yz) as the symmetric
// tensor whose off-diagonal elements are the xy/xz/yz coefficients themselves (efprot.f:78-136).
__global__ void k_rot_vecl(double *vec, int64_t nrows, int nbasis, const int *start, int nsh, const double *rot) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nrows * nsh) return;
  const int64_t row = w / nsh; const int sh = (int)(w % nsh);
  const int k = start[2 * sh], typ = start[2 * sh + 1];
  double R[9];
  for (int d = 0; d < 9; ++d) R[d] = rot[d];
  double *c = vec + row * nbasis + k;
  if (typ == 1) {
    double in[3] = {c[0], c[1], c[2]}, o[3];
    rot_vec(R, in, o);
    c[0] = o[0]; c[1] = o[1]; c[2] = o[2];
  } else {
    double in[6] = {c[0], c[1], c[2], c[3], c[4], c[5]}, o[6];
    rot_quad(R, in, o);
    for (int d = 0; d < 6; ++d) c[d] = o[d];
  }
}

}  // namespace slv
