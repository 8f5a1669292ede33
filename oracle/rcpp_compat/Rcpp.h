// TEST INFRASTRUCTURE ONLY -- not part of the shipped product.
//
// Minimal, header-only stand-in for <Rcpp.h>, just large enough that the reference's
// own native kernels
//     /root/reference/src/find_mutual_nns.cpp
//     /root/reference/src/smooth_gaussian_kernel.cpp
//     /root/reference/src/adjust_shift_variance.cpp
//     /root/reference/src/utils.cpp
// compile UNMODIFIED, from where they lie, into oracle/_ref/ (see oracle/Makefile).
// There is no R in this image, so this is how the reference's real code becomes the
// parity oracle for those three kernels.  Nothing here is copied from Rcpp; it only
// offers the handful of members those four files touch:
//   column-major Matrix<T> (nrow/ncol/size/begin/end/row()/column()), Vector<T>,
//   List::create, R_NegInf / R_NaReal / NA_INTEGER and R::logspace_add.
#ifndef B200MNN_ORACLE_RCPP_COMPAT_H
#define B200MNN_ORACLE_RCPP_COMPAT_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <deque>
#include <iterator>
#include <limits>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <vector>

static const double R_NegInf = -std::numeric_limits<double>::infinity();
static const double R_PosInf = std::numeric_limits<double>::infinity();
static const double R_NaReal = std::numeric_limits<double>::quiet_NaN();
static const double R_NaN = std::numeric_limits<double>::quiet_NaN();
static const int NA_INTEGER = std::numeric_limits<int>::min();

namespace R {
// Rmath's definition: log(exp(lx) + exp(ly)) without leaving log space.
inline double logspace_add(double lx, double ly) {
    return ((lx > ly) ? lx : ly) + std::log1p(std::exp(-std::fabs(lx - ly)));
}
}  // namespace R

namespace Rcpp {

// Random-access iterator with a runtime stride (matrix rows are strided in column-major storage).
template <typename T>
class StridedIter {
public:
    using iterator_category = std::random_access_iterator_tag;
    using value_type = typename std::remove_const<T>::type;
    using difference_type = std::ptrdiff_t;
    using pointer = T*;
    using reference = T&;

    StridedIter() : p_(nullptr), s_(1) {}
    StridedIter(T* p, std::ptrdiff_t s) : p_(p), s_(s) {}
    reference operator*() const { return *p_; }
    reference operator[](difference_type i) const { return p_[i * s_]; }
    StridedIter& operator++() { p_ += s_; return *this; }
    StridedIter operator++(int) { StridedIter t(*this); p_ += s_; return t; }
    StridedIter& operator--() { p_ -= s_; return *this; }
    StridedIter& operator+=(difference_type n) { p_ += n * s_; return *this; }
    StridedIter& operator-=(difference_type n) { p_ -= n * s_; return *this; }
    StridedIter operator+(difference_type n) const { return StridedIter(p_ + n * s_, s_); }
    StridedIter operator-(difference_type n) const { return StridedIter(p_ - n * s_, s_); }
    difference_type operator-(const StridedIter& o) const { return (p_ - o.p_) / s_; }
    bool operator==(const StridedIter& o) const { return p_ == o.p_; }
    bool operator!=(const StridedIter& o) const { return p_ != o.p_; }
    bool operator<(const StridedIter& o) const { return (s_ > 0) ? p_ < o.p_ : p_ > o.p_; }

private:
    T* p_;
    std::ptrdiff_t s_;
};

// A view of one row or one column.
template <typename T>
class Slice {
public:
    using iterator = StridedIter<T>;
    Slice(T* p, std::size_t n, std::ptrdiff_t stride) : p_(p), n_(n), s_(stride) {}
    iterator begin() const { return iterator(p_, s_); }
    iterator end() const { return iterator(p_ + static_cast<std::ptrdiff_t>(n_) * s_, s_); }
    std::size_t size() const { return n_; }
    T& operator[](std::size_t i) const { return p_[static_cast<std::ptrdiff_t>(i) * s_]; }

private:
    T* p_;
    std::size_t n_;
    std::ptrdiff_t s_;
};

// A contiguous view of one column: iterators are raw pointers (the reference hands
// column.begin() to functions taking `const double*`).
template <typename T>
class ColumnView {
public:
    using iterator = T*;
    ColumnView(T* p, std::size_t n) : p_(p), n_(n) {}
    iterator begin() const { return p_; }
    iterator end() const { return p_ + n_; }
    std::size_t size() const { return n_; }
    T& operator[](std::size_t i) const { return p_[i]; }

private:
    T* p_;
    std::size_t n_;
};

// R vectors are reference-like handles: copies share storage.  A Vector either owns its
// storage (shared) or views caller memory (the C wrappers in ref_capi.cpp use views).
template <typename T>
class Vector {
public:
    using iterator = T*;
    using const_iterator = const T*;

    Vector() : n_(0), p_(nullptr) {}
    explicit Vector(std::size_t n) : own_(std::make_shared<std::vector<T>>(n, T(0))), n_(n), p_(own_->data()) {}
    Vector(int n) : Vector(static_cast<std::size_t>(n)) {}
    template <typename It, typename = typename std::iterator_traits<It>::iterator_category>
    Vector(It first, It last) : own_(std::make_shared<std::vector<T>>(first, last)), n_(own_->size()), p_(own_->data()) {}
    static Vector view(T* p, std::size_t n) { Vector v; v.n_ = n; v.p_ = p; return v; }

    std::size_t size() const { return n_; }
    std::size_t length() const { return n_; }
    iterator begin() const { return p_; }
    iterator end() const { return p_ + n_; }
    T& operator[](std::size_t i) const { return p_[i]; }

private:
    std::shared_ptr<std::vector<T>> own_;
    std::size_t n_;
    T* p_;
};

template <typename T>
class Matrix {
public:
    using iterator = T*;
    using Row = Slice<T>;
    using Column = ColumnView<T>;

    Matrix() : nr_(0), nc_(0), p_(nullptr) {}
    Matrix(std::size_t nr, std::size_t nc)
        : own_(std::make_shared<std::vector<T>>(nr * nc, T(0))), nr_(nr), nc_(nc), p_(own_->data()) {}
    static Matrix view(T* p, std::size_t nr, std::size_t nc) { Matrix m; m.nr_ = nr; m.nc_ = nc; m.p_ = p; return m; }

    int nrow() const { return static_cast<int>(nr_); }
    int ncol() const { return static_cast<int>(nc_); }
    std::size_t size() const { return nr_ * nc_; }
    iterator begin() const { return p_; }
    iterator end() const { return p_ + nr_ * nc_; }
    T& operator()(std::size_t r, std::size_t c) const { return p_[r + c * nr_]; }
    Row row(std::size_t r) const { return Row(p_ + r, nc_, static_cast<std::ptrdiff_t>(nr_)); }
    Column column(std::size_t c) const { return Column(p_ + c * nr_, nr_); }

private:
    std::shared_ptr<std::vector<T>> own_;
    std::size_t nr_, nc_;
    T* p_;
};

using NumericVector = Vector<double>;
using IntegerVector = Vector<int>;
using NumericMatrix = Matrix<double>;
using IntegerMatrix = Matrix<int>;

// Only the two-integer-vector form is needed (find_mutual_nns' return value).
class List {
public:
    static List create(const IntegerVector& a, const IntegerVector& b) { List l; l.items_.push_back(a); l.items_.push_back(b); return l; }
    const IntegerVector& operator[](std::size_t i) const { return items_[i]; }
    std::size_t size() const { return items_.size(); }

private:
    std::vector<IntegerVector> items_;
};

}  // namespace Rcpp

// smooth_gaussian_kernel() is declared to return SEXP and returns a NumericMatrix.
using SEXP = Rcpp::NumericMatrix;

#endif
