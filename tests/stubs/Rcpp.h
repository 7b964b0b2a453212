// Minimal stand-in for <Rcpp.h>, ONLY so that the Rcpp shim can be syntax-checked where R is not
// installed (tests/test_abi_cpu.py). It models the handful of Rcpp types the shim touches; nothing links.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>
typedef std::ptrdiff_t R_xlen_t;
struct SEXPREC;
typedef SEXPREC *SEXP;
inline bool Rf_isNull(SEXP) { return false; }
namespace Rcpp {
struct Proxy {
  std::string s;
  Proxy &operator=(const std::string &v) { s = v; return *this; }
  Proxy &operator=(const Proxy &v) { s = v.s; return *this; }
  Proxy() = default;
  Proxy(const Proxy &) = default;
  operator SEXP() const { return nullptr; }
};
template <typename T> T as(const Proxy &p) { return T(p.s); }
struct Names { template <typename V> Names &operator=(const V &) { return *this; } };
struct CharacterVector {
  std::vector<Proxy> v;
  CharacterVector() = default;
  explicit CharacterVector(std::size_t n) : v(n) {}
  CharacterVector(SEXP) {}
  R_xlen_t size() const { return (R_xlen_t)v.size(); }
  Proxy &operator[](R_xlen_t i) { return v[(std::size_t)i]; }
  void push_back(const std::string &s) { Proxy p; p.s = s; v.push_back(p); }
  void push_back(const Proxy &p) { v.push_back(p); }
};
struct ListProxy {
  template <typename V> ListProxy &operator=(const V &) { return *this; }
  operator SEXP() const { return nullptr; }
  operator CharacterVector() const { return CharacterVector(); }
};
struct List {
  List() = default;
  explicit List(std::size_t) {}
  R_xlen_t size() const { return 0; }
  ListProxy operator[](std::size_t) { return ListProxy(); }
  ListProxy operator[](const char *) { return ListProxy(); }
};
struct NumericVector {
  std::vector<double> v;
  explicit NumericVector(std::size_t n) : v(n) {}
  double &operator[](std::size_t i) { return v[i]; }
  Names names() { return Names(); }
};
[[noreturn]] inline void stop(const std::string &m) { throw std::runtime_error(m); }
inline void warning(const std::string &) {}
inline void checkUserInterrupt() {}
template <typename C> struct class_ {
  explicit class_(const char *) {}
  template <typename... A> class_ &constructor() { return *this; }
  template <typename G> class_ &property(const char *, G) { return *this; }
  template <typename G, typename S> class_ &property(const char *, G, S) { return *this; }
  template <typename M> class_ &method(const char *, M) { return *this; }
};
}  // namespace Rcpp
#define RCPP_MODULE(name) void rcpp_module_##name()
