// TEST INFRASTRUCTURE ONLY -- not part of the shipped product.
//
// extern "C" handles onto the reference's own native kernels, compiled unmodified from
// /root/reference/src (see oracle/Makefile) against oracle/rcpp_compat/Rcpp.h.
// Used by tests/ (as the parity checker), by tests/golden/make_golden.py (to emit the
// committed fixtures) and by bench.py's cpu_baseline / --impl reference leg.
//
// Layout conventions are R's: column-major double / int32, 1-based where the reference is.
#include "Rcpp.h"

#include <cstdint>
#include <cstring>
#include <string>

// Declarations of the reference entry points (defined in /root/reference/src/*.cpp):
//   src/find_mutual_nns.cpp:8, src/smooth_gaussian_kernel.cpp:11, src/adjust_shift_variance.cpp:30
Rcpp::List find_mutual_nns(Rcpp::IntegerMatrix left, Rcpp::IntegerMatrix right);
SEXP smooth_gaussian_kernel(Rcpp::NumericMatrix averaged, Rcpp::IntegerVector index, Rcpp::NumericMatrix mat, double sigma2);
Rcpp::NumericVector adjust_shift_variance(Rcpp::NumericMatrix data1, Rcpp::NumericMatrix data2, Rcpp::NumericMatrix vect,
                                          double sigma2, Rcpp::IntegerVector restrict1, Rcpp::IntegerVector restrict2);

static thread_local std::string g_err;

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// left: [n1 x k2] int32 col-major (1-based ids into right's rows); right: [n2 x k1].
// first_out/second_out must hold n1*k2 entries; *np_out receives the pair count.
int ref_find_mutual_nns(const int32_t* left, int64_t n1, int k2, const int32_t* right, int64_t n2, int k1,
                        int32_t* first_out, int32_t* second_out, int64_t* np_out) {
    try {
        auto L = Rcpp::IntegerMatrix::view(const_cast<int*>(left), n1, k2);
        auto R = Rcpp::IntegerMatrix::view(const_cast<int*>(right), n2, k1);
        Rcpp::List out = find_mutual_nns(L, R);
        const auto& a = out[0];
        const auto& b = out[1];
        *np_out = static_cast<int64_t>(a.size());
        std::copy(a.begin(), a.end(), first_out);
        std::copy(b.begin(), b.end(), second_out);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// averaged [G x nmnn], index0 int32[nidx] (0-based), mat [Gdist x ncells] -> out [G x ncells]; all col-major.
int ref_smooth_gaussian_kernel(const double* averaged, int64_t G, int64_t nmnn, const int32_t* index0, int64_t nidx,
                               const double* mat, int64_t Gdist, int64_t ncells, double sigma2, double* out) {
    try {
        auto A = Rcpp::NumericMatrix::view(const_cast<double*>(averaged), G, nmnn);
        auto I = Rcpp::IntegerVector::view(const_cast<int*>(index0), nidx);
        auto M = Rcpp::NumericMatrix::view(const_cast<double*>(mat), Gdist, ncells);
        Rcpp::NumericMatrix o = smooth_gaussian_kernel(A, I, M, sigma2);
        std::copy(o.begin(), o.end(), out);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// data1 [G1 x n1], data2 [G2 x n2], vect [vr x vc] (the reference demands vr == n2, vc == G), restricts 0-based.
int ref_adjust_shift_variance(const double* data1, int64_t G1, int64_t n1, const double* data2, int64_t G2, int64_t n2,
                              const double* vect, int64_t vr, int64_t vc, double sigma2,
                              const int32_t* r1, int64_t nr1, const int32_t* r2, int64_t nr2, double* out) {
    try {
        auto D1 = Rcpp::NumericMatrix::view(const_cast<double*>(data1), G1, n1);
        auto D2 = Rcpp::NumericMatrix::view(const_cast<double*>(data2), G2, n2);
        auto V = Rcpp::NumericMatrix::view(const_cast<double*>(vect), vr, vc);
        auto R1 = Rcpp::IntegerVector::view(const_cast<int*>(r1), nr1);
        auto R2 = Rcpp::IntegerVector::view(const_cast<int*>(r2), nr2);
        Rcpp::NumericVector o = adjust_shift_variance(D1, D2, V, sigma2, R1, R2);
        std::copy(o.begin(), o.end(), out);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

}  // extern "C"
