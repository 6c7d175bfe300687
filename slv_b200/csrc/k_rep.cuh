// k_rep.cuh -- kernel D: exchange-repulsion shift (shftex.f SHFTEX = CALCIJ + CALCFI with one-electron
// integrals S, T, dS/dA, dT/dA evaluated on the fly).  Filled in below.
#pragma once
#include <string>
#include <vector>
#include "slv_common.cuh"

namespace slv {

constexpr int REP_THREADS = 256;
struct RepWork { double *buf; };

inline int rep_alloc(RepWork &rw, std::vector<void *> &allocs, std::string &err, int sm_count, int nmosc, int nbasc,
                     int max_nmos, int max_nbas, int *ctas) {
  (void)rw; (void)allocs; (void)nmosc; (void)nbasc; (void)max_nmos; (void)max_nbas;
  *ctas = sm_count;
  err = "exchange-repulsion kernel not built into this library";
  return SLVGPU_EINVAL;
}
inline size_t rep_smem_bytes(int, int, int, int) { return 0; }
__global__ void k_rep(const DevPlan, int, int, FrameLists, RepWork, double *) {}

}  // namespace slv
