// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Minimal stand-in for <Rcpp.h>, written from scratch, so that the reference's
// two translation units (src/minHash.cpp, src/pairwiseSeqAlign.cpp) compile
// UNMODIFIED, where they lie under /root/reference, into oracle/_ref/ without R.
// It provides only the Rcpp surface those two files touch:
//   Rcpp::stop / Rcpp::warning (printf-like, %s with std::string, %c, %d)
//   CharacterVector: (n) ctor, length(), operator[] (read as string, assign string)
//   NumericMatrix:   (r,c) ctor zero-filled column-major, operator()(i,j), attr("dimnames") = ..., attr(name) = double / std::vector<double>
//   List::create(a, b), as<std::string>(elem)
//   IntegerVector / NumericVector / IntegerMatrix (only what rpkg/src/dyna_shims.cpp needs, for the shim harness)
// Nothing here mirrors Rcpp's implementation; it is a behavioural stub.
#ifndef DYNA_ORACLE_RCPP_STUB_H
#define DYNA_ORACLE_RCPP_STUB_H

#include <cstddef>
#include <cstdint>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace Rcpp {

class exception : public std::runtime_error {
public:
  explicit exception(const std::string& w) : std::runtime_error(w) {}
};

namespace stubfmt {
inline void emit(std::ostringstream& os, const char*& f) {
  // copy literal text up to the next conversion (or end); "%%" -> '%'
  while (*f) {
    if (f[0] == '%' && f[1] == '%') { os << '%'; f += 2; continue; }
    if (f[0] == '%') return;
    os << *f++;
  }
}
inline void skip_spec(const char*& f) {
  if (*f != '%') return;
  ++f;
  while (*f && !((*f >= 'a' && *f <= 'z') || (*f >= 'A' && *f <= 'Z'))) ++f;
  if (*f) ++f;
}
inline void build(std::ostringstream& os, const char* f) { emit(os, f); }
template <class T, class... Rest>
void build(std::ostringstream& os, const char* f, const T& v, const Rest&... rest) {
  emit(os, f);
  if (*f == '%') { skip_spec(f); os << v; }
  build(os, f, rest...);
}
}  // namespace stubfmt

template <class... Args>
[[noreturn]] void stop(const char* fmt, const Args&... args) {
  std::ostringstream os;
  stubfmt::build(os, fmt, args...);
  throw exception(os.str());
}
[[noreturn]] inline void stop(const std::string& msg) { throw exception(msg); }

// warnings are recorded, not printed (the oracle wrapper can read the last one)
inline std::string& last_warning() { static thread_local std::string w; return w; }
template <class... Args>
void warning(const char* fmt, const Args&... args) {
  std::ostringstream os;
  stubfmt::build(os, fmt, args...);
  last_warning() = os.str();
}

// --- CharacterVector -------------------------------------------------------
class CharacterVector {
  std::vector<std::string> v_;
public:
  CharacterVector() {}
  explicit CharacterVector(std::size_t n) : v_(n) {}
  explicit CharacterVector(std::vector<std::string> v) : v_(std::move(v)) {}
  long length() const { return static_cast<long>(v_.size()); }
  long size() const { return length(); }
  std::string& operator[](std::size_t i) { return v_[i]; }
  const std::string& operator[](std::size_t i) const { return v_[i]; }
  const std::vector<std::string>& data() const { return v_; }
};

template <class T> T as(const std::string& s) { return T(s); }

// --- List / attributes -------------------------------------------------------
struct List {
  std::vector<CharacterVector> items;
  static List create(const CharacterVector& a, const CharacterVector& b) {
    List l; l.items.push_back(a); l.items.push_back(b); return l;
  }
};

// --- NumericMatrix -----------------------------------------------------------
class NumericMatrix {
  std::size_t nr_, nc_;
  std::vector<double> d_;
  List dimnames_;
  double scalar_attr_ = 0.0;
  std::vector<double> vector_attr_;
  struct AttrProxy {
    NumericMatrix* m;
    AttrProxy& operator=(const List& l) { m->dimnames_ = l; return *this; }
    AttrProxy& operator=(double v) { m->scalar_attr_ = v; return *this; }  // the one scalar attribute the shims set
    AttrProxy& operator=(const std::vector<double>& v) { m->vector_attr_ = v; return *this; }  // and the one vector attribute
  };
public:
  NumericMatrix() : nr_(0), nc_(0) {}
  NumericMatrix(std::size_t r, std::size_t c) : nr_(r), nc_(c), d_(r * c, 0.0) {}
  double& operator()(std::size_t i, std::size_t j) { return d_[i + j * nr_]; }          // column-major, as R
  const double& operator()(std::size_t i, std::size_t j) const { return d_[i + j * nr_]; }
  AttrProxy attr(const char*) { return AttrProxy{this}; }
  std::size_t nrow() const { return nr_; }
  std::size_t ncol() const { return nc_; }
  const double* begin() const { return d_.data(); }
  double* begin() { return d_.data(); }
  const List& dimnames() const { return dimnames_; }
  double scalar_attr() const { return scalar_attr_; }
  const std::vector<double>& vector_attr() const { return vector_attr_; }
};

// --- plain vectors / integer matrix (shim harness only) -------------------------
template <class T>
class StubVector {
  std::vector<T> v_;
public:
  StubVector() {}
  explicit StubVector(std::vector<T> v) : v_(std::move(v)) {}
  long length() const { return static_cast<long>(v_.size()); }
  typename std::vector<T>::const_iterator begin() const { return v_.begin(); }
  typename std::vector<T>::const_iterator end() const { return v_.end(); }
};
typedef StubVector<int> IntegerVector;
typedef StubVector<double> NumericVector;

class IntegerMatrix {
  int nr_, nc_;
  std::vector<int> d_;
public:
  IntegerMatrix(int r, int c, std::vector<int> d) : nr_(r), nc_(c), d_(std::move(d)) {}
  int nrow() const { return nr_; }
  int ncol() const { return nc_; }
  int operator()(int i, int j) const { return d_[static_cast<std::size_t>(i) + static_cast<std::size_t>(j) * nr_]; }
};

}  // namespace Rcpp

#endif
