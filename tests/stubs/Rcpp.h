// Stand-in for <Rcpp.h> with REAL storage, so that the Rcpp shim (elections.dtree_b200/host/r/R_tree_b200.cpp)
// can be compiled AND EXECUTED where R is not installed: tests/shim/shim_driver.cpp replays the reference's testthat
// scenarios through the shim's RDirichletTree class itself (tests/test_gpu_shim.py), and tests/test_abi_cpu.py
// compiles and links it. It models exactly the Rcpp surface the shim touches — CharacterVector, List (by index and
// by name), NumericVector with names, as<std::string>, stop / warning / checkUserInterrupt, class_ registration —
// with R's semantics where the shim relies on them (a List element can be NULL, a character vector, ...).
// Warnings are recorded in Rcpp::stub::warnings() so that a driver can assert them, as expect_warning does.
#pragma once
#include <cstddef>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

typedef std::ptrdiff_t R_xlen_t;

namespace Rcpp {
struct CharacterVector;
struct NumericVector;
namespace stub {
// An R value as far as the shim can see one: NULL, a character vector or a numeric vector.
struct Value {
  std::shared_ptr<std::vector<std::string>> chr;
  std::shared_ptr<std::vector<double>> num;
  std::shared_ptr<std::vector<std::string>> names;
  bool is_null() const { return !chr && !num; }
};
inline std::vector<std::string> &warnings() {
  static std::vector<std::string> w;
  return w;
}
inline bool &interrupt_pending() {
  static bool flag = false;
  return flag;
}
}  // namespace stub
}  // namespace Rcpp

typedef const Rcpp::stub::Value *SEXP;
inline bool Rf_isNull(SEXP s) { return s == nullptr || s->is_null(); }

namespace Rcpp {

struct exception : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct interrupted_exception : std::runtime_error {
  interrupted_exception() : std::runtime_error("interrupted") {}
};
[[noreturn]] inline void stop(const std::string &m) { throw exception(m); }
inline void warning(const std::string &m) { stub::warnings().push_back(m); }
inline void checkUserInterrupt() {
  if (stub::interrupt_pending()) throw interrupted_exception();
}

// Element of a character vector: assignable from strings and from other elements.
struct StringProxy {
  std::string *s;
  StringProxy &operator=(const std::string &v) { *s = v; return *this; }
  StringProxy &operator=(const char *v) { *s = v; return *this; }
  StringProxy &operator=(const StringProxy &v) { *s = *v.s; return *this; }
  operator std::string() const { return *s; }
};
template <typename T> T as(const StringProxy &p) { return T(*p.s); }

struct CharacterVector {
  stub::Value v;
  CharacterVector() { v.chr = std::make_shared<std::vector<std::string>>(); }
  explicit CharacterVector(std::size_t n) { v.chr = std::make_shared<std::vector<std::string>>(n); }
  CharacterVector(std::initializer_list<const char *> l) {
    v.chr = std::make_shared<std::vector<std::string>>();
    for (const char *s : l) v.chr->push_back(s);
  }
  CharacterVector(SEXP s) {  // what `Rcpp::CharacterVector x = list[i]` does; R would throw on a non-character value
    if (s == nullptr || !s->chr) {
      if (s != nullptr && s->num) throw exception("Expecting a string vector: not compatible with requested type");
      v.chr = std::make_shared<std::vector<std::string>>();
    } else {
      v = *s;
    }
  }
  R_xlen_t size() const { return (R_xlen_t)v.chr->size(); }
  StringProxy operator[](R_xlen_t i) { return StringProxy{&(*v.chr)[(std::size_t)i]}; }
  std::string at(std::size_t i) const { return (*v.chr)[i]; }
  void push_back(const std::string &s) {  // copy-on-append, like Rcpp (vectors are values in R)
    auto n = std::make_shared<std::vector<std::string>>(*v.chr);
    n->push_back(s);
    v.chr = n;
  }
  void push_back(const StringProxy &p) { push_back(*p.s); }
};

struct NumericVector {
  stub::Value v;
  explicit NumericVector(std::size_t n) { v.num = std::make_shared<std::vector<double>>(n, 0.0); }
  double &operator[](std::size_t i) { return (*v.num)[i]; }
  R_xlen_t size() const { return (R_xlen_t)v.num->size(); }
  struct NamesProxy {
    stub::Value *v;
    NamesProxy &operator=(const CharacterVector &c) {
      v->names = std::make_shared<std::vector<std::string>>(*c.v.chr);
      return *this;
    }
  };
  NamesProxy names() { return NamesProxy{&v}; }
  std::vector<std::string> names_vector() const { return v.names ? *v.names : std::vector<std::string>(); }
};

struct List {
  std::shared_ptr<std::vector<stub::Value>> items = std::make_shared<std::vector<stub::Value>>();
  std::shared_ptr<std::vector<std::string>> keys = std::make_shared<std::vector<std::string>>();
  List() = default;
  explicit List(std::size_t n) {
    items->resize(n);
    keys->resize(n);
  }
  R_xlen_t size() const { return (R_xlen_t)items->size(); }
  struct Proxy {
    stub::Value *slot;
    Proxy &operator=(const CharacterVector &c) { *slot = c.v; return *this; }
    Proxy &operator=(const NumericVector &n) { *slot = n.v; return *this; }
    operator SEXP() const { return slot; }
    operator CharacterVector() const { return CharacterVector((SEXP)slot); }
  };
  template <typename I, typename = typename std::enable_if<std::is_integral<I>::value>::type>
  Proxy operator[](I i) { return Proxy{&(*items)[(std::size_t)i]}; }
  Proxy operator[](const char *key) {
    for (std::size_t i = 0; i < keys->size(); ++i)
      if ((*keys)[i] == key) return Proxy{&(*items)[i]};
    items->emplace_back();
    keys->emplace_back(key);
    return Proxy{&items->back()};
  }
  void push_back(const CharacterVector &c) {
    items->push_back(c.v);
    keys->emplace_back();
  }
  void push_back_null() {
    items->emplace_back();
    keys->emplace_back();
  }
  void push_back(const NumericVector &n) {
    items->push_back(n.v);
    keys->emplace_back();
  }
};

// Module registration: nothing to register without R; the chain must merely compile.
template <typename C> struct class_ {
  explicit class_(const char *) {}
  template <typename... A> class_ &constructor() { return *this; }
  template <typename G> class_ &property(const char *, G) { return *this; }
  template <typename G, typename S> class_ &property(const char *, G, S) { return *this; }
  template <typename M> class_ &method(const char *, M) { return *this; }
};
}  // namespace Rcpp
#define RCPP_MODULE(name) void rcpp_module_##name()
