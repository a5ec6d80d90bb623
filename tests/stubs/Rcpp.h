// Minimal stand-in for <Rcpp.h>: just enough declarations for a SYNTAX / TYPE check of shim/r_shim.cpp in an image
// without R (tests/test_r_shim.py compiles the shim with -fsyntax-only against this file).  Nothing here is ever
// linked or run; the real Rcpp provides the same names with the same shapes.
#pragma once
#include <cstddef>
#include <exception>
#include <string>

struct SEXPREC;
typedef SEXPREC* SEXP;
extern "C" SEXP R_NilValue;
extern "C" void Rf_error(const char*, ...) __attribute__((noreturn));
extern "C" double unif_rand(void);
#define RcppExport extern "C"

namespace Rcpp {

template <typename T>
class Vector {
 public:
  Vector();
  explicit Vector(int n);
  Vector(SEXP x);
  T* begin();
  const T* begin() const;
  long size() const;
  T& operator[](long i);
  operator SEXP() const;
};

template <typename T>
class Matrix {
 public:
  Matrix();
  Matrix(int nrow, int ncol);
  Matrix(SEXP x);
  T* begin();
  const T* begin() const;
  int nrow() const;
  int ncol() const;
  T& operator()(int i, int j);
  operator SEXP() const;
};

typedef Vector<double> NumericVector;
typedef Vector<int> IntegerVector;
typedef Vector<int> LogicalVector;  // R logicals are ints
typedef Matrix<double> NumericMatrix;
typedef Matrix<int> IntegerMatrix;
typedef Matrix<int> LogicalMatrix;

template <typename T>
T as(SEXP x);
template <typename T>
SEXP wrap(const T& x);
template <typename T>
T clone(const T& x);

struct NamedValue {
  const char* name;
  SEXP value;
};
class Named {
 public:
  explicit Named(const char* n);
  template <typename T>
  NamedValue operator=(const T& v) const;
};

class List {
 public:
  template <typename... Args>
  static SEXP create(const Args&... args);
};

class RNGScope {
 public:
  RNGScope();
  ~RNGScope();
};

}  // namespace Rcpp
