// Rcpp.h -- STUB for a syntax / type check of the glue where no R toolchain exists (tests/test_glue_cpu.py):
//     g++ -std=c++17 -fsyntax-only -I r-glue/stub -I include r-glue/*.cpp
// Declares just the slice of the Rcpp API the glue uses, with the real signatures' shapes.  Not a replacement for Rcpp:
// nothing here is ever linked or run.
#ifndef BVG_STUB_RCPP_H
#define BVG_STUB_RCPP_H
#include <cstddef>
#include <string>
#include <vector>

typedef struct SEXPREC* SEXP;
typedef std::ptrdiff_t R_xlen_t;
extern SEXP R_NilValue;
bool Rf_isNull(SEXP);
double unif_rand(void);

namespace Rcpp {

class String {
 public:
  operator std::string() const;
};

class CharacterVector {
 public:
  CharacterVector();
  CharacterVector(SEXP);
  R_xlen_t size() const;
  String operator[](R_xlen_t) const;
};

class NumericVector {
 public:
  NumericVector();
  explicit NumericVector(R_xlen_t n);
  NumericVector(SEXP);
  R_xlen_t size() const;
  double& operator[](R_xlen_t);
  const double& operator[](R_xlen_t) const;
  const double* begin() const;
  const double* end() const;
  operator SEXP() const;
};

class NumericMatrix : public NumericVector {
 public:
  NumericMatrix();
  NumericMatrix(int nrow, int ncol);
  NumericMatrix(SEXP);
  int nrow() const;
  int ncol() const;
  double& operator()(int, int);
  const double& operator()(int, int) const;
};

template <class T> T as(SEXP);

namespace internal {
struct NamedValue { const char* name; SEXP value; };
}
class Named {
 public:
  explicit Named(const char*);
  Named(const Named&) = delete;
  Named& operator=(const Named&) = delete;
  template <class T> internal::NamedValue operator=(const T&) const;
};

class List {
 public:
  class Proxy {
   public:
    operator SEXP() const;
    operator NumericVector() const;
    operator NumericMatrix() const;
    template <class T> Proxy& operator=(const T&);
  };
  List();
  explicit List(int n);
  List(SEXP);
  List(const Proxy&);
  SEXP names() const;
  R_xlen_t size() const;
  Proxy operator[](const char*);
  const Proxy operator[](const char*) const;
  Proxy operator[](int);
  template <class... A> static List create(const A&...);
  operator SEXP() const;
};

void stop(const char*);
void stop(const std::string&);
template <class... A> void warning(const char* fmt, A...);
void checkUserInterrupt();

}  // namespace Rcpp
#endif
