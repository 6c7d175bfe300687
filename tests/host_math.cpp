// Host build of slv_b200/csrc/slv_math.cuh for CPU unit checks (tests/test_host_math.py).
// TEST INFRASTRUCTURE: the product never loads this.
#include "../slv_b200/csrc/slv_math.cuh"
extern "C" {
void hm_kabsch_rot(const double* a, double* rot) { slv::kabsch_rot_jacobi(a, rot); }
void hm_kabsch_rot_fast(const double* a, double gsum, double* rot) { slv::kabsch_rot(a, gsum, rot); }
void hm_rot_mom20(const double* R, const double* m, double* o) { slv::rot_mom20(R, m, o); }
void hm_rot_mat(const double* R, const double* a, double* o) { slv::rot_mat(R, a, o); }
void hm_pair_potentials(const double* rA, const double* rB, const double* mB, double* P) { slv::pair_potentials(rA, rB, mB, P); }
void hm_fold20(double* s) { slv::fold20(s); }
}
extern "C" {
void hm_site_field(const double* R, const double* m, double f, double* F) { slv::site_field(R, m, f, F); }
void hm_site_dfield(const double* R, const double* v, const double* m, double sg, double* F) { slv::site_dfield(R, v, m, sg, F); }
}
#include "../slv_b200/csrc/slv_ints.cuh"
extern "C" void hm_site_field_acc(const double* R, const double* m, double f, double w, double* F) { slv::site_field_acc(R, m, f, w, F); }

// all AO integrals between two fragments through the sub-shell-pair specialisations, walking the classes the way
// k_rep does (tests/test_host_math.py compares with oracle/ints_oracle.py)
extern "C" void hm_subshell_V(int nsubA, const int* subA, const double* posA, const double* eA, const double* cA, const double* LcA,
                              int nsubB, const int* subB, const double* posB, const double* eB, const double* cB, int LP, double* V) {
  for (int a = 0; a < nsubA; ++a)
    for (int b = 0; b < nsubB; ++b) {
      const int* sa = subA + a * 4; const int* sb = subB + b * 4;
#define X(ID, LA, LB, IA) \
      if (sa[2] == LA && sb[2] == LB) slv::subshell_pair_store<LA, LB, IA>(sa, sb, posA, posB, eA, cA, eB, cB, LcA, LP, V);
      SLV_SUBSHELL_CLASSES(X)
#undef X
    }
}
