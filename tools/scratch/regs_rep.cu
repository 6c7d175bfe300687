#include "../../slv_b200/csrc/slv_ints.cuh"
using namespace slv;
#define X(ID, LA, LB, IA) \
__global__ void __launch_bounds__(256, 2) k_##ID(const int* sa, const int* sb, const double* pA, const double* pB, const double* eA, const double* cA, const double* eB, const double* cB, const double* L, int n, double* V) { \
  subshell_pair_store<LA, LB, IA>(sa + threadIdx.x * 4, sb + threadIdx.x * 4, pA, pB, eA, cA, eB, cB, L, n, V); }
SLV_SUBSHELL_CLASSES(X)
