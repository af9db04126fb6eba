// segment_matcher.cpp — see segment_matcher.h. A port of the normal-form classifier of fuser_b200/matcher.py.
#include "segment_matcher.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <functional>
#include <set>
#include <stdexcept>

#include "../../../include/b200norm.h"

namespace b200norm {
namespace {

struct Unmatchable : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ---------------------------------------------------------------------------------------------
// rational exponents
// ---------------------------------------------------------------------------------------------
int64_t gcd64(int64_t a, int64_t b) {
  a = a < 0 ? -a : a;
  b = b < 0 ? -b : b;
  while (b != 0) {
    const int64_t t = a % b;
    a = b;
    b = t;
  }
  return a == 0 ? 1 : a;
}
struct Rat {
  int64_t n = 0, d = 1;
  Rat() = default;
  Rat(int64_t n_, int64_t d_ = 1) : n(n_), d(d_) {
    if (d < 0) {
      n = -n;
      d = -d;
    }
    const int64_t g = gcd64(n, d);
    n /= g;
    d /= g;
  }
  Rat operator+(const Rat& o) const { return Rat(n * o.d + o.n * d, d * o.d); }
  Rat operator*(const Rat& o) const { return Rat(n * o.n, d * o.d); }
  bool operator==(const Rat& o) const { return n == o.n && d == o.d; }
  bool operator!=(const Rat& o) const { return !(*this == o); }
  bool operator<(const Rat& o) const { return n * o.d < o.n * d; }
  bool is_int() const { return d == 1; }
  double val() const { return (double)n / (double)d; }
};
// Fraction(x).limit_denominator(64)
Rat rat_from_double(double x) {
  Rat best(std::llround(x), 1);
  double best_err = std::fabs(x - best.val());
  for (int64_t d = 2; d <= 64 && best_err > 1e-12; ++d) {
    const Rat c(std::llround(x * (double)d), d);
    const double e = std::fabs(x - c.val());
    if (e < best_err - 1e-15) {
      best = c;
      best_err = e;
    }
  }
  return best;
}

// ---------------------------------------------------------------------------------------------
// polynomial normal form:  sum_i coef_i * prod_j atom_ij ^ exp_ij
// ---------------------------------------------------------------------------------------------
using AtomId = int32_t;
using Monomial = std::vector<std::pair<AtomId, Rat>>;  // sorted by atom id, no zero exponents
struct MonoLess {
  bool operator()(const Monomial& a, const Monomial& b) const {
    if (a.size() != b.size()) return a.size() < b.size();
    for (size_t i = 0; i < a.size(); ++i) {
      if (a[i].first != b[i].first) return a[i].first < b[i].first;
      if (a[i].second != b[i].second) return a[i].second < b[i].second;
    }
    return false;
  }
};
// A term disappears only when its contributions CANCEL, never because it is small next to the others: an
// epsilon of 1e-12 next to a variance is a term like any other.
constexpr double kCancel = 1e-12;

struct Poly {
  std::map<Monomial, double, MonoLess> terms;

  static Poly constant(double c) {
    Poly p;
    if (c != 0.0) p.terms[Monomial()] = c;
    return p;
  }
  static Poly atom(AtomId a, Rat e = Rat(1)) {
    Poly p;
    p.terms[Monomial{{a, e}}] = 1.0;
    return p;
  }
  bool is_monomial() const { return terms.size() == 1; }
};

struct Acc {  // accumulate contributions, dropping keys whose contributions cancel
  std::map<Monomial, std::pair<double, double>, MonoLess> t;
  void add(const Monomial& k, double c) {
    auto& e = t[k];
    e.first += c;
    e.second += std::fabs(c);
  }
  Poly finish() const {
    Poly p;
    for (const auto& kv : t) {
      const double c = kv.second.first;
      if (c == c && std::fabs(c) > kCancel * kv.second.second) p.terms[kv.first] = c;
    }
    return p;
  }
};

Monomial mul_mono(const Monomial& a, const Monomial& b) {
  std::map<AtomId, Rat> d;
  for (const auto& f : a) d[f.first] = d.count(f.first) ? d[f.first] + f.second : f.second;
  for (const auto& f : b) d[f.first] = d.count(f.first) ? d[f.first] + f.second : f.second;
  Monomial r;
  for (const auto& kv : d)
    if (kv.second.n != 0) r.emplace_back(kv.first, kv.second);
  return r;
}
Poly add(const Poly& a, const Poly& b) {
  Acc acc;
  for (const auto& kv : a.terms) acc.add(kv.first, kv.second);
  for (const auto& kv : b.terms) acc.add(kv.first, kv.second);
  return acc.finish();
}
Poly scale(const Poly& a, double c) {
  Poly p;
  for (const auto& kv : a.terms)
    if (kv.second * c != 0.0) p.terms[kv.first] = kv.second * c;
  return p;
}
Poly sub(const Poly& a, const Poly& b) { return add(a, scale(b, -1.0)); }
Poly mul(const Poly& a, const Poly& b) {
  Acc acc;
  for (const auto& x : a.terms)
    for (const auto& y : b.terms) acc.add(mul_mono(x.first, y.first), x.second * y.second);
  return acc.finish();
}

// ---------------------------------------------------------------------------------------------
// canonicaliser
// ---------------------------------------------------------------------------------------------
enum AtomKind { A_IN, A_DIM, A_SUM, A_MAX, A_EXP, A_LOG, A_EXPR, A_KEEP };
struct AtomInfo {
  int kind;
  int32_t a;               // A_IN: node index; A_DIM: extent class; A_KEEP: node index of the uniform() tensor
  std::set<int32_t> red;   // A_SUM / A_MAX: reduced extent classes
  int32_t key;             // A_SUM / A_MAX / A_EXP / A_LOG / A_EXPR: id of the canonical inner expression
};

// One context per match: atoms and canonical keys are interned here, so equal expressions get equal ids no matter
// which side (the segment or a candidate definition) produced them first.
struct Context {
  std::map<std::string, AtomId> atom_ids;
  std::vector<AtomInfo> atoms;
  std::map<std::string, int32_t> key_ids;
  std::vector<std::set<int32_t>> axes_of_key;

  AtomId intern(const AtomInfo& a) {
    std::string r = std::to_string(a.kind) + ":" + std::to_string(a.a) + ":" + std::to_string(a.key) + ":";
    for (int32_t c : a.red) r += std::to_string(c) + ",";
    auto it = atom_ids.find(r);
    if (it != atom_ids.end()) return it->second;
    const AtomId id = (AtomId)atoms.size();
    atoms.push_back(a);
    atom_ids[r] = id;
    return id;
  }
};

std::string poly_key(const Poly& p) {
  std::string s;
  char buf[48];
  for (const auto& kv : p.terms) {
    for (const auto& f : kv.first) {
      s += std::to_string(f.first) + "^" + std::to_string(f.second.n) + "/" + std::to_string(f.second.d) + "*";
    }
    snprintf(buf, sizeof(buf), ":%.9e;", kv.second);
    s += buf;
  }
  return s;
}

struct Canon {
  const std::vector<ExprNode>& nodes;
  Context& ctx;
  std::map<int32_t, Poly> cache;

  Canon(const std::vector<ExprNode>& n, Context& c) : nodes(n), ctx(c) {}

  std::set<int32_t> atom_axes(AtomId id) const {
    const AtomInfo& a = ctx.atoms[id];
    std::set<int32_t> r;
    switch (a.kind) {
      case A_IN:
      case A_KEEP:  // the dropout mask uniform() < keep spans the axes of the uniform tensor
        for (int32_t d : nodes[a.a].dims)
          if (d >= 0) r.insert(d);
        return r;
      case A_DIM: return r;
      case A_SUM:
      case A_MAX:
        for (int32_t c : ctx.axes_of_key[a.key])
          if (!a.red.count(c)) r.insert(c);
        return r;
      default: return ctx.axes_of_key[a.key];
    }
  }
  std::set<int32_t> poly_axes(const Poly& p) const {
    std::set<int32_t> r;
    for (const auto& kv : p.terms)
      for (const auto& f : kv.first) {
        const auto ax = atom_axes(f.first);
        r.insert(ax.begin(), ax.end());
      }
    return r;
  }
  int32_t reg(const Poly& p) {
    const std::string k = poly_key(p);
    auto it = ctx.key_ids.find(k);
    if (it != ctx.key_ids.end()) return it->second;
    const int32_t id = (int32_t)ctx.axes_of_key.size();
    ctx.axes_of_key.push_back(poly_axes(p));
    ctx.key_ids[k] = id;
    return id;
  }

  Poly pow(const Poly& p, Rat e) {
    if (e == Rat(1)) return p;
    if (e.n == 0) return Poly::constant(1.0);
    if (p.terms.empty()) return Poly();
    if (p.is_monomial()) {
      const auto& kv = *p.terms.begin();
      if (kv.second > 0 || e.is_int()) {
        Monomial m;
        for (const auto& f : kv.first) {
          const Rat x = f.second * e;
          if (x.n != 0) m.emplace_back(f.first, x);
        }
        Poly r;
        r.terms[m] = std::pow(kv.second, e.val());
        return r;
      }
    }
    if (e.is_int() && e.n > 0 && e.n <= 4) {
      Poly r = p;
      for (int64_t i = 1; i < e.n; ++i) r = mul(r, p);
      return r;
    }
    AtomInfo a{A_EXPR, 0, {}, reg(p)};  // non-integer / negative power of a multi-term expression: opaque
    return Poly::atom(ctx.intern(a), e);
  }

  // sum over the logical axes red_idx of a tensor whose extent classes are tensor_dims
  Poly reduce_sum(const Poly& p, const std::vector<int32_t>& tensor_dims, const std::vector<int32_t>& red_idx) {
    std::set<int32_t> red;
    for (int32_t i : red_idx)
      if (tensor_dims.at(i) >= 0) red.insert(tensor_dims[i]);
    if (red.empty()) return p;  // reduction over broadcast axes only: identity
    Poly out;
    for (const auto& kv : p.terms) {
      Monomial dep, indep;
      std::set<int32_t> dep_axes;
      for (const auto& f : kv.first) {
        const auto ax = atom_axes(f.first);
        bool hits = false;
        for (int32_t c : ax) hits = hits || red.count(c) != 0;
        if (hits) {
          dep.push_back(f);
          dep_axes.insert(ax.begin(), ax.end());
        } else {
          indep.push_back(f);
        }
      }
      Poly term;
      term.terms[indep] = kv.second;
      std::set<int32_t> inner_red;
      for (int32_t c : red) {
        if (dep_axes.count(c)) {
          inner_red.insert(c);
        } else {  // a reduced axis nothing in the term depends on contributes its extent
          term = mul(term, Poly::atom(ctx.intern(AtomInfo{A_DIM, c, {}, -1})));
        }
      }
      if (!inner_red.empty()) {
        Poly inner;
        inner.terms[dep] = 1.0;
        term = mul(term, Poly::atom(ctx.intern(AtomInfo{A_SUM, 0, inner_red, reg(inner)})));
      }
      out = add(out, term);
    }
    return out;
  }

  Poly value(int32_t id) {
    auto it = cache.find(id);
    if (it != cache.end()) return it->second;
    Poly r = compute(id);
    cache[id] = r;
    return r;
  }

  Poly compute(int32_t id) {
    const ExprNode& n = nodes.at(id);
    auto arg = [&](int i) { return value(n.args.at(i)); };
    switch (n.kind) {
      case ExprKind::Input:
        if (n.is_scalar) throw Unmatchable("symbolic scalar inputs are not supported by the normalization kernels");
        return Poly::atom(ctx.intern(AtomInfo{A_IN, id, {}, -1}));
      case ExprKind::Constant: return Poly::constant(n.value);
      case ExprKind::Size: return Poly::atom(ctx.intern(AtomInfo{A_DIM, n.size_class, {}, -1}));
      case ExprKind::Cast:
      case ExprKind::Broadcast: return arg(0);
      case ExprKind::Neg: return scale(arg(0), -1.0);
      case ExprKind::Reciprocal: return pow(arg(0), Rat(-1));
      case ExprKind::Sqrt: return pow(arg(0), Rat(1, 2));
      case ExprKind::Rsqrt: return pow(arg(0), Rat(-1, 2));
      case ExprKind::Exp:
      case ExprKind::Log: {
        const int32_t k = reg(arg(0));
        return Poly::atom(ctx.intern(AtomInfo{n.kind == ExprKind::Exp ? A_EXP : A_LOG, 0, {}, k}));
      }
      case ExprKind::Uniform: throw Unmatchable("uniform() outside a dropout mask");
      case ExprKind::Lt: {
        // only as the dropout mask: uniform(0, 1) < constant (benchmarks/python/test_dropout_layernorm_fwd.py:38-41)
        const ExprNode& u = nodes.at(n.args.at(0));
        const ExprNode& c = nodes.at(n.args.at(1));
        if (u.kind == ExprKind::Uniform && !u.is_scalar && u.dtype == kF32 && c.kind == ExprKind::Constant &&
            u.args.size() == 2 && nodes.at(u.args[0]).kind == ExprKind::Constant && nodes.at(u.args[0]).value == 0.0 &&
            nodes.at(u.args[1]).kind == ExprKind::Constant && nodes.at(u.args[1]).value == 1.0)
          return Poly::atom(ctx.intern(AtomInfo{A_KEEP, n.args[0], {}, -1}));
        throw Unmatchable("lt outside a dropout mask");
      }
      case ExprKind::Add: return add(arg(0), arg(1));
      case ExprKind::Sub: return sub(arg(0), arg(1));
      case ExprKind::Mul: return mul(arg(0), arg(1));
      case ExprKind::Div: return mul(arg(0), pow(arg(1), Rat(-1)));
      case ExprKind::Pow: {
        const ExprNode& e = nodes.at(n.args.at(1));
        if (e.kind != ExprKind::Constant) throw Unmatchable("pow with a non-constant exponent");
        return pow(arg(0), rat_from_double(e.value));
      }
      case ExprKind::Sum:
      case ExprKind::Max:
      case ExprKind::Mean:
      case ExprKind::Var:
      case ExprKind::M2: {
        const ExprNode& src = nodes.at(n.args.at(0));
        const Poly x = arg(0);
        if (n.kind == ExprKind::Sum) return reduce_sum(x, src.dims, n.red_axes);
        std::set<int32_t> red;
        for (int32_t i : n.red_axes)
          if (src.dims.at(i) >= 0) red.insert(src.dims[i]);
        if (n.kind == ExprKind::Max) {
          if (red.empty()) return x;
          return Poly::atom(ctx.intern(AtomInfo{A_MAX, 0, red, reg(x)}));
        }
        Poly cnt = Poly::constant(1.0);
        for (int32_t c : red) cnt = mul(cnt, Poly::atom(ctx.intern(AtomInfo{A_DIM, c, {}, -1})));
        const Poly mean = mul(reduce_sum(x, src.dims, n.red_axes), pow(cnt, Rat(-1)));
        if (n.kind == ExprKind::Mean) return mean;
        const Poly d = sub(x, mean);
        const Poly m2 = reduce_sum(mul(d, d), src.dims, n.red_axes);
        if (n.kind == ExprKind::M2) return m2;
        const Poly denom = n.correction ? sub(cnt, Poly::constant((double)n.correction)) : cnt;
        return mul(m2, pow(denom, Rat(-1)));
      }
    }
    throw Unmatchable("op is not on the normalization hot path");
  }
};

// ---------------------------------------------------------------------------------------------
// candidate op definitions, written against a small graph builder (math of csrc/ops/normalization.cpp)
// ---------------------------------------------------------------------------------------------
struct Builder {
  std::vector<ExprNode>& nodes;
  explicit Builder(std::vector<ExprNode>& n) : nodes(n) {}
  int32_t push(ExprNode n) {
    nodes.push_back(std::move(n));
    return (int32_t)nodes.size() - 1;
  }
  int ndim(int32_t t) const { return (int)nodes[t].dims.size(); }
  int32_t constant(double v) {
    ExprNode n;
    n.kind = ExprKind::Constant;
    n.is_scalar = true;
    n.value = v;
    return push(n);
  }
  int32_t size(int32_t t, int axis) {
    const int a = axis < 0 ? axis + ndim(t) : axis;
    ExprNode n;
    n.kind = ExprKind::Size;
    n.is_scalar = true;
    n.size_class = nodes[t].dims.at(a);
    if (n.size_class < 0) return constant(1.0);
    return push(n);
  }
  int32_t unary(ExprKind k, int32_t a) {
    ExprNode n;
    n.kind = k;
    n.is_scalar = nodes[a].is_scalar;
    n.args = {a};
    n.dims = nodes[a].dims;
    return push(n);
  }
  int32_t binary(ExprKind k, int32_t a, int32_t b) {
    ExprNode n;
    n.kind = k;
    n.args = {a, b};
    const ExprNode &x = nodes[a], &y = nodes[b];
    n.is_scalar = x.is_scalar && y.is_scalar;
    if (x.is_scalar) {
      n.dims = y.dims;
    } else if (y.is_scalar) {
      n.dims = x.dims;
    } else {
      if (x.dims.size() != y.dims.size()) throw Unmatchable("candidate: rank mismatch");
      for (size_t i = 0; i < x.dims.size(); ++i) {
        if (x.dims[i] < 0) {
          n.dims.push_back(y.dims[i]);
        } else if (y.dims[i] < 0 || y.dims[i] == x.dims[i]) {
          n.dims.push_back(x.dims[i]);
        } else {
          throw Unmatchable("candidate: distinct extents meet (wrong role guess)");
        }
      }
    }
    return push(n);
  }
  int32_t add(int32_t a, int32_t b) { return binary(ExprKind::Add, a, b); }
  int32_t sub(int32_t a, int32_t b) { return binary(ExprKind::Sub, a, b); }
  int32_t mul(int32_t a, int32_t b) { return binary(ExprKind::Mul, a, b); }
  int32_t reduce(ExprKind k, int32_t a, const std::vector<int32_t>& axes, bool keepdim, int corr = 0) {
    ExprNode n;
    n.kind = k;
    n.args = {a};
    n.red_axes = axes;
    n.correction = corr;
    for (int i = 0; i < ndim(a); ++i) {
      const bool r = std::find(axes.begin(), axes.end(), i) != axes.end();
      if (r) {
        if (keepdim) n.dims.push_back(-1);
      } else {
        n.dims.push_back(nodes[a].dims[i]);
      }
    }
    return push(n);
  }
  // ops.broadcast(t, mask): a new size-1 axis wherever mask is true
  int32_t bcast_mask(int32_t t, const std::vector<bool>& mask) {
    ExprNode n;
    n.kind = ExprKind::Broadcast;
    n.args = {t};
    size_t k = 0;
    for (bool m : mask) n.dims.push_back(m ? -1 : nodes[t].dims.at(k++));
    if (k != nodes[t].dims.size()) throw Unmatchable("candidate: broadcast mask does not fit the operand");
    return push(n);
  }
  // [N] -> [.., N] aligned with the last axis of `like`
  int32_t bcast_last(int32_t t, int32_t like) {
    ExprNode n;
    n.kind = ExprKind::Broadcast;
    n.args = {t};
    n.dims.assign(ndim(like), -1);
    n.dims.back() = nodes[t].dims.at(0);
    return push(n);
  }
  // per-row statistic ([..] or [.., 1]) -> broadcastable against `like`
  int32_t bcast_rows(int32_t t, int32_t like) {
    if (ndim(t) == ndim(like)) return t;
    ExprNode n;
    n.kind = ExprKind::Broadcast;
    n.args = {t};
    n.dims = nodes[t].dims;
    n.dims.push_back(-1);
    return push(n);
  }
};

using Outs = std::map<std::string, int32_t>;
constexpr int32_t kNone = -1;

Outs def_layer_norm_fwd(Builder& b, int32_t x, int32_t w, int32_t bias, int32_t eps) {
  const std::vector<int32_t> last{b.ndim(x) - 1};
  const int32_t var = b.reduce(ExprKind::Var, x, last, true), mean = b.reduce(ExprKind::Mean, x, last, true);
  const int32_t invstd = b.unary(ExprKind::Rsqrt, b.add(var, eps));
  int32_t y = b.mul(b.sub(x, mean), invstd);
  if (w != kNone) y = b.mul(y, b.bcast_last(w, x));
  if (bias != kNone) y = b.add(y, b.bcast_last(bias, x));
  return {{"y", y}, {"mean", mean}, {"invstd", invstd}};
}
std::vector<int32_t> outer_axes(Builder& b, int32_t x) {
  std::vector<int32_t> o;
  for (int i = 0; i + 1 < b.ndim(x); ++i) o.push_back(i);
  return o;
}
Outs def_outer_norm_fwd(Builder& b, int32_t x, int32_t w, int32_t bias, int32_t eps) {
  const auto outer = outer_axes(b, x);
  const int32_t var = b.reduce(ExprKind::Var, x, outer, false), mean = b.reduce(ExprKind::Mean, x, outer, false);
  const int32_t invstd = b.unary(ExprKind::Rsqrt, b.add(var, eps));
  int32_t y = b.mul(b.sub(x, b.bcast_last(mean, x)), b.bcast_last(invstd, x));
  if (w != kNone) y = b.mul(y, b.bcast_last(w, x));
  if (bias != kNone) y = b.add(y, b.bcast_last(bias, x));
  return {{"y", y}, {"mean", mean}, {"invstd", invstd}};
}
Outs def_outer_norm_bwd(Builder& b, int32_t dy, int32_t x, int32_t mean, int32_t invstd, int32_t w) {
  const auto outer = outer_axes(b, x);
  int32_t num = b.size(x, 0);
  for (size_t i = 1; i < outer.size(); ++i) num = b.mul(num, b.size(x, outer[i]));
  const int32_t norm = b.unary(ExprKind::Reciprocal, num);
  const int32_t xc = b.sub(x, b.bcast_last(mean, x));
  const int32_t grad_sum = b.reduce(ExprKind::Sum, dy, outer, false);
  const int32_t dot_p = b.reduce(ExprKind::Sum, b.mul(dy, xc), outer, false);
  const int32_t grad_mean = b.bcast_last(b.mul(grad_sum, norm), x);
  const int32_t proj_scale = b.bcast_last(b.mul(b.mul(dot_p, norm), b.mul(invstd, invstd)), x);
  int32_t grad_scale = b.bcast_last(invstd, x);
  if (w != kNone) grad_scale = b.mul(b.bcast_last(w, x), grad_scale);
  const int32_t dx = b.mul(b.sub(b.sub(dy, b.mul(proj_scale, xc)), grad_mean), grad_scale);
  return {{"dx", dx}, {"dw", b.mul(dot_p, invstd)}, {"db", grad_sum}};
}
// instance norm (csrc/ops/normalization.cpp:790-950, :952-1060; benchmarks/python/normalization.py with
// norm = "instance_norm"): statistics over the spatial axes of every (sample, channel) plane, affine per channel
Outs def_instance_norm_fwd(Builder& b, int32_t x, int32_t w, int32_t bias, int32_t eps, int c_axis) {
  const int nd = b.ndim(x);
  std::vector<int32_t> red;
  std::vector<bool> smask(nd, false), cmask(nd, true);
  for (int a = 1; a < nd; ++a)
    if (a != c_axis) {
      red.push_back(a);
      smask[a] = true;
    }
  cmask[c_axis] = false;
  const int32_t var = b.reduce(ExprKind::Var, x, red, false), mean = b.reduce(ExprKind::Mean, x, red, false);
  const int32_t invstd = b.unary(ExprKind::Rsqrt, b.add(var, eps));
  int32_t y = b.mul(b.sub(x, b.bcast_mask(mean, smask)), b.bcast_mask(invstd, smask));
  if (w != kNone) y = b.mul(y, b.bcast_mask(w, cmask));
  if (bias != kNone) y = b.add(y, b.bcast_mask(bias, cmask));
  return {{"y", y}, {"mean", mean}, {"invstd", invstd}};
}
Outs def_instance_norm_bwd(Builder& b, int32_t dy, int32_t x, int32_t mean, int32_t invstd, int32_t w, int c_axis) {
  const int nd = b.ndim(x);
  std::vector<int32_t> red;
  std::vector<bool> smask(nd, false), cmask(nd, true);
  for (int a = 1; a < nd; ++a)
    if (a != c_axis) {
      red.push_back(a);
      smask[a] = true;
    }
  cmask[c_axis] = false;
  int32_t num = b.size(x, red.at(0));
  for (size_t i = 1; i < red.size(); ++i) num = b.mul(num, b.size(x, red[i]));
  const int32_t norm = b.unary(ExprKind::Reciprocal, num);
  const int32_t xc = b.sub(x, b.bcast_mask(mean, smask));
  const int32_t grad_sum = b.reduce(ExprKind::Sum, dy, red, false);
  const int32_t dot_p = b.reduce(ExprKind::Sum, b.mul(dy, xc), red, false);
  const int32_t grad_mean = b.bcast_mask(b.mul(grad_sum, norm), smask);
  const int32_t proj_scale = b.bcast_mask(b.mul(b.mul(dot_p, norm), b.mul(invstd, invstd)), smask);
  int32_t grad_scale = b.bcast_mask(invstd, smask);
  if (w != kNone) grad_scale = b.mul(b.bcast_mask(w, cmask), grad_scale);
  const int32_t dx = b.mul(b.sub(b.sub(dy, b.mul(proj_scale, xc)), grad_mean), grad_scale);
  return {{"dx", dx}, {"dw", b.reduce(ExprKind::Sum, b.mul(dot_p, invstd), {0}, false)},
          {"db", b.reduce(ExprKind::Sum, grad_sum, {0}, false)}};
}
Outs def_rms_norm_fwd(Builder& b, int32_t x, int32_t w, int32_t eps) {
  const std::vector<int32_t> last{b.ndim(x) - 1};
  const int32_t ms = b.mul(b.reduce(ExprKind::Sum, b.mul(x, x), last, true), b.unary(ExprKind::Reciprocal, b.size(x, -1)));
  const int32_t var_eps = b.add(ms, eps);
  const int32_t invstd = b.unary(ExprKind::Rsqrt, var_eps);
  int32_t y = b.mul(x, invstd);
  if (w != kNone) y = b.mul(y, b.bcast_last(w, x));
  return {{"y", y}, {"invstd", invstd}, {"rms", b.unary(ExprKind::Sqrt, var_eps)}};
}
Outs def_softmax_fwd(Builder& b, int32_t x, bool log, int axis) {
  const std::vector<int32_t> ax{axis < 0 ? axis + b.ndim(x) : axis};
  const int32_t d = b.sub(x, b.reduce(ExprKind::Max, x, ax, true));
  const int32_t e = b.unary(ExprKind::Exp, d);
  const int32_t s = b.reduce(ExprKind::Sum, e, ax, true);
  return {{"y", log ? b.sub(d, b.unary(ExprKind::Log, s)) : b.mul(e, b.unary(ExprKind::Reciprocal, s))}};
}
Outs def_softmax_bwd(Builder& b, int32_t dy, int32_t y, bool log, int axis) {
  const std::vector<int32_t> ax{axis < 0 ? axis + b.ndim(y) : axis};
  if (log) return {{"dx", b.sub(dy, b.mul(b.unary(ExprKind::Exp, y), b.reduce(ExprKind::Sum, dy, ax, true)))}};
  const int32_t g = b.mul(dy, y);
  return {{"dx", b.sub(g, b.mul(y, b.reduce(ExprKind::Sum, g, ax, true)))}};
}
Outs def_norm_bwd(Builder& b, int32_t dy, int32_t x, int32_t mean, int32_t stat, int32_t w, bool stat_is_rms,
                  bool centered, bool half_projection) {
  const std::vector<int32_t> last{b.ndim(x) - 1};
  const auto outer = outer_axes(b, x);
  const int32_t st = b.bcast_rows(stat, x);
  const int32_t invstd = stat_is_rms ? b.unary(ExprKind::Reciprocal, st) : st;
  const int32_t xc = mean != kNone ? b.sub(x, b.bcast_rows(mean, x)) : x;
  const int32_t xh = b.mul(xc, invstd);
  const int32_t g = w != kNone ? b.mul(dy, b.bcast_last(w, x)) : dy;
  const int32_t n = b.size(x, -1);
  int32_t inner = b.mul(n, g);
  if (centered) inner = b.sub(inner, b.reduce(ExprKind::Sum, g, last, true));
  int32_t proj = b.mul(xh, b.reduce(ExprKind::Sum, b.mul(g, xh), last, true));
  if (half_projection) proj = b.mul(proj, b.constant(0.5));
  inner = b.sub(inner, proj);
  const int32_t dx = b.mul(b.mul(b.unary(ExprKind::Reciprocal, n), invstd), inner);
  return {{"dx", dx}, {"db", b.reduce(ExprKind::Sum, dy, outer, false)},
          {"dw", b.reduce(ExprKind::Sum, b.mul(dy, xh), outer, false)}};
}
Outs def_reduce(Builder& b, int32_t x, bool is_sum, bool inner) {
  const std::vector<int32_t> axes = inner ? std::vector<int32_t>{b.ndim(x) - 1} : outer_axes(b, x);
  return {{"r", b.reduce(is_sum ? ExprKind::Sum : ExprKind::Max, x, axes, false)}};
}
Outs def_var_mean(Builder& b, int32_t x, int corr) {
  const std::vector<int32_t> last{b.ndim(x) - 1};
  return {{"var", b.reduce(ExprKind::Var, x, last, false, corr)}, {"mean", b.reduce(ExprKind::Mean, x, last, false)},
          {"m2", b.reduce(ExprKind::M2, x, last, false)}};
}

// ---------------------------------------------------------------------------------------------
// graph facts
// ---------------------------------------------------------------------------------------------
bool reduced(int32_t dt) { return dt == kBF16 || dt == kF16; }

struct Facts {
  std::vector<char> live;
  std::vector<char> consumed;
};
Facts facts_of(const ExprGraph& g) {
  Facts f;
  f.live.assign(g.nodes.size(), 0);
  f.consumed.assign(g.nodes.size(), 0);
  std::vector<int32_t> stack(g.outputs.begin(), g.outputs.end());
  while (!stack.empty()) {
    const int32_t v = stack.back();
    stack.pop_back();
    if (f.live[v]) continue;
    f.live[v] = 1;
    for (int32_t a : g.nodes[v].args) stack.push_back(a);
  }
  for (const auto& n : g.nodes)
    for (int32_t a : n.args) f.consumed[a] = 1;
  return f;
}
// scalar constants added to a tensor: the epsilon is one of them
std::vector<double> eps_candidates(const ExprGraph& g) {
  std::vector<double> out;
  for (const auto& n : g.nodes) {
    if (n.is_scalar || (n.kind != ExprKind::Add && n.kind != ExprKind::Sub)) continue;
    for (int32_t a : n.args) {
      const ExprNode& s = g.nodes[a];
      if (s.kind == ExprKind::Constant && std::fabs(s.value) > 0 && std::fabs(s.value) < 1.0 &&
          std::find(out.begin(), out.end(), s.value) == out.end())
        out.push_back(s.value);
    }
  }
  if (out.empty()) out.push_back(0.0);
  return out;
}
// a cast to a reduced-precision dtype whose result feeds more arithmetic (the row-sum round trip of
// benchmarks/python/test_softmax_bwd.py:49-60)
bool has_intermediate_round(const ExprGraph& g, const Facts& f) {
  std::set<int32_t> outs(g.outputs.begin(), g.outputs.end());
  for (size_t i = 0; i < g.nodes.size(); ++i) {
    const ExprNode& n = g.nodes[i];
    if (!n.is_scalar && n.kind == ExprKind::Cast && reduced(n.dtype) && f.consumed[i] && !outs.count((int32_t)i)) {
      const ExprNode& src = g.nodes[n.args.at(0)];
      if (!src.is_scalar && src.kind != ExprKind::Input) return true;
    }
  }
  return false;
}
// the residual sum is rounded to a reduced-precision dtype before it is normalised
bool sum_is_rounded(const ExprGraph& g, const Facts& f) {
  for (size_t i = 0; i < g.nodes.size(); ++i) {
    const ExprNode& n = g.nodes[i];
    if (!n.is_scalar && n.kind == ExprKind::Cast && reduced(n.dtype) && f.consumed[i]) {
      const ExprNode& src = g.nodes[n.args.at(0)];
      if (!src.is_scalar && src.kind == ExprKind::Add) return true;
    }
  }
  return false;
}

struct In {
  int32_t idx;   // position in graph.inputs
  int32_t node;  // node index
};

// Instance norm (training statistics) forward / backward on [B, C, spatial...] or [B, spatial..., C]: its parameters
// are indexed by an axis that is neither a row nor the column of the [rows, cols] view the other segments use, so it is
// recognised before that classification (mirrors matcher.py: _match_instance_norm).
bool match_instance_norm(const ExprGraph& g, const std::vector<In>& tin, std::vector<In> mats, Context& ctx,
                         const std::vector<std::string>& mine, SegmentMatch* out) {
  auto N = [&](const In& i) -> const ExprNode& { return g.nodes[i.node]; };
  size_t nd = 0;
  for (const In& i : mats) nd = std::max(nd, N(i).dims.size());
  {
    std::vector<In> keep;
    for (const In& i : mats)
      if (N(i).dims.size() == nd) keep.push_back(i);  // the saved [B, C] statistics of the backward are rank 2
    mats = keep;
  }
  if (nd < 3 || (mats.size() != 1 && mats.size() != 2)) return false;
  const ExprNode& x_like = N(mats[0]);
  for (const In& i : mats)
    if (N(i).dims != x_like.dims) return false;
  const int32_t io = x_like.dtype;
  if (io != kF32 && io != kF16 && io != kBF16) return false;
  for (const In& i : mats)
    if (N(i).dtype != io) return false;
  {
    std::set<int32_t> distinct(x_like.dims.begin(), x_like.dims.end());
    if (distinct.size() != nd) return false;  // two axes share an extent class: not decidable here
  }
  std::vector<In> others;
  for (const In& i : tin) {
    bool is_mat = false;
    for (const In& m : mats) is_mat = is_mat || m.node == i.node;
    if (!is_mat) others.push_back(i);
  }
  std::vector<std::string> names;
  auto attempt = [&](const std::function<Outs(Builder&)>& build) -> bool {
    std::vector<ExprNode> nodes = g.nodes;
    Builder b(nodes);
    std::map<std::string, std::string> keys;
    try {
      const Outs cand = build(b);
      Canon c(nodes, ctx);
      for (const auto& kv : cand)
        if (kv.second != kNone) keys[kv.first] = poly_key(c.value(kv.second));
    } catch (const Unmatchable&) {
      return false;
    } catch (const std::out_of_range&) {
      return false;
    }
    names.clear();
    for (const std::string& k : mine) {
      bool hit = false;
      for (const auto& kv : keys)
        if (kv.second == k) {
          names.push_back(kv.first);
          hit = true;
          break;
        }
      if (!hit) return false;
    }
    return true;
  };
  auto has = [&](const char* nm) { return std::find(names.begin(), names.end(), nm) != names.end(); };
  std::vector<double> eps_list = eps_candidates(g);
  for (int c_axis : {1, (int)nd - 1}) {
    const int32_t ccls = x_like.dims[c_axis], bcls = x_like.dims[0];
    std::vector<In> chan, plane;
    for (const In& i : others) {
      const ExprNode& n = N(i);
      if (n.dims.size() == 1 && n.dims[0] == ccls && n.dtype == io) chan.push_back(i);
      else if (n.dims.size() == 2 && n.dims[0] == bcls && n.dims[1] == ccls && n.dtype == kF32) plane.push_back(i);
    }
    if (chan.size() + plane.size() != others.size()) continue;
    if (mats.size() == 1 && plane.empty() && chan.size() <= 2) {
      const In& x = mats[0];
      std::vector<std::pair<const In*, const In*>> combos{{nullptr, nullptr}};
      if (chan.size() == 1) combos = {{&chan[0], nullptr}};
      if (chan.size() == 2) combos = {{&chan[0], &chan[1]}, {&chan[1], &chan[0]}};
      for (double eps : eps_list)
        for (const auto& wb : combos) {
          const In *w = wb.first, *bi = wb.second;
          if (attempt([&](Builder& b) {
                return def_instance_norm_fwd(b, x.node, w ? w->node : kNone, bi ? bi->node : kNone, b.constant(eps), c_axis);
              }) &&
              has("y")) {
            out->op = B200NORM_OP_LAYERNORM_FWD;
            out->flags = B200NORM_F_ROW_AFFINE | (w ? B200NORM_F_HAS_WEIGHT : 0) | (bi ? B200NORM_F_HAS_BIAS : 0);
            out->eps = eps;
            out->io_dtype = io;
            out->roles = {{"x", x.idx}};
            if (w) out->roles["w"] = w->idx;
            if (bi) out->roles["b"] = bi->idx;
            out->outputs = names;
            out->channels_last = c_axis != 1;
            out->description = "instance_norm forward";
            return true;
          }
        }
    }
    if (mats.size() == 2 && plane.size() == 2 && chan.size() <= 1) {
      const In* w = chan.empty() ? nullptr : &chan[0];
      for (int order = 0; order < 2; ++order) {
        const In &dy_ = mats[order], &x_ = mats[1 - order];
        for (int so = 0; so < 2; ++so) {
          const In &mean = plane[so], &st = plane[1 - so];
          if (attempt([&](Builder& b) {
                return def_instance_norm_bwd(b, dy_.node, x_.node, mean.node, st.node, w ? w->node : kNone, c_axis);
              }) &&
              has("dx")) {
            out->op = B200NORM_OP_LAYERNORM_BWD;
            out->flags = B200NORM_F_ROW_AFFINE | (w ? B200NORM_F_HAS_WEIGHT : 0);
            out->eps = 0.0;
            out->io_dtype = io;
            out->roles = {{"dy", dy_.idx}, {"x", x_.idx}, {"mean", mean.idx}, {"stat", st.idx}};
            if (w) out->roles["w"] = w->idx;
            out->outputs = names;
            out->channels_last = c_axis != 1;
            out->description = "instance_norm backward";
            return true;
          }
        }
      }
    }
  }
  return false;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// the classifier
// ---------------------------------------------------------------------------------------------
static void match_impl(const ExprGraph& g, SegmentMatch* out) {
  if (g.outputs.empty()) throw Unmatchable("fusion has no outputs");
  const Facts facts = facts_of(g);
  for (int32_t i : g.inputs)
    if (g.nodes[i].is_scalar) throw Unmatchable("symbolic scalar inputs are not supported by the normalization kernels");
  std::vector<In> tin;  // tensor inputs that reach an output (dead inputs are not part of the segment)
  for (size_t k = 0; k < g.inputs.size(); ++k)
    if (facts.live[g.inputs[k]]) tin.push_back(In{(int32_t)k, g.inputs[k]});

  Context ctx;
  std::vector<std::string> mine;
  {
    Canon c(g.nodes, ctx);
    for (int32_t o : g.outputs) mine.push_back(poly_key(c.value(o)));
  }
  auto N = [&](const In& i) -> const ExprNode& { return g.nodes[i.node]; };
  auto full = [&](const ExprNode& n) {
    return std::all_of(n.dims.begin(), n.dims.end(), [](int32_t d) { return d >= 0; });
  };
  std::vector<In> mats;
  for (const In& i : tin)
    if (N(i).dims.size() >= 2 && full(N(i))) mats.push_back(i);
  if (mats.empty()) throw Unmatchable("no [rows, cols] input");
  if (match_instance_norm(g, tin, mats, ctx, mine, out)) return;
  const ExprNode& x_like = N(mats[0]);
  const size_t nd = x_like.dims.size();
  const int32_t col_cls = x_like.dims.back();
  const std::vector<int32_t> row_cls(x_like.dims.begin(), x_like.dims.end() - 1);
  if (std::find(row_cls.begin(), row_cls.end(), col_cls) != row_cls.end())
    throw Unmatchable("square symbolic shapes: rows and cols were unified");
  std::vector<In> masks;  // a saved dropout mask (Bool matrix, backward)
  {
    std::vector<In> keep;
    for (const In& i : mats) {
      if (N(i).dims != x_like.dims) continue;
      (N(i).dtype == kBool ? masks : keep).push_back(i);
    }
    mats = keep;
  }
  if (mats.empty()) throw Unmatchable("no [rows, cols] input");
  const int32_t io = N(mats[0]).dtype;
  std::vector<In> vecs, cstats, stats;
  for (const In& i : tin) {
    const ExprNode& n = N(i);
    if (n.dims.size() == 1 && n.dims[0] == col_cls) {
      // fp32 per-COLUMN statistics next to reduced-precision activations (outer normalization backward)
      if (n.dtype == kF32 && io != kF32) {
        cstats.push_back(i);
      } else {
        vecs.push_back(i);
      }
    }
    const bool rows_only = n.dims.size() == nd - 1 && n.dims == row_cls;
    const bool rows_keep = n.dims.size() == nd && n.dims.back() < 0 &&
                           std::vector<int32_t>(n.dims.begin(), n.dims.end() - 1) == row_cls;
    if (rows_only || rows_keep) stats.push_back(i);
  }
  if (mats.size() + masks.size() + vecs.size() + cstats.size() + stats.size() != tin.size())
    throw Unmatchable("an input is neither [rows, cols], [cols] nor a per-row statistic");
  if (io != kF32 && io != kF16 && io != kBF16) throw Unmatchable("io dtype");
  for (const In& i : mats)
    if (N(i).dtype != io) throw Unmatchable("activations / weights of mixed dtypes");
  for (const In& i : vecs)
    if (N(i).dtype != io) throw Unmatchable("activations / weights of mixed dtypes");
  for (const In& i : stats)
    if (N(i).dtype != kF32) throw Unmatchable("per-row statistics must be Float");

  const std::vector<double> eps_list = eps_candidates(g);
  const size_t n_m = mats.size(), n_v = vecs.size(), n_s = stats.size();

  // Instantiate a candidate definition on the segment's inputs and compare normal forms; on success `names`
  // holds the role of every fusion output.
  std::vector<std::string> names;
  auto try_candidate = [&](const std::function<Outs(Builder&)>& build) -> bool {
    std::vector<ExprNode> nodes = g.nodes;
    Builder b(nodes);
    Outs cand;
    std::map<std::string, std::string> keys;
    try {
      cand = build(b);
      Canon c(nodes, ctx);
      for (const auto& kv : cand)
        if (kv.second != kNone) keys[kv.first] = poly_key(c.value(kv.second));
    } catch (const Unmatchable&) {
      return false;
    } catch (const std::out_of_range&) {
      return false;
    }
    names.clear();
    for (const std::string& k : mine) {
      bool hit = false;
      for (const auto& kv : keys) {
        if (kv.second == k) {
          names.push_back(kv.first);
          hit = true;
          break;
        }
      }
      if (!hit) return false;
    }
    return true;
  };
  auto has = [&](const char* nm) { return std::find(names.begin(), names.end(), nm) != names.end(); };
  auto finish = [&](int32_t op, uint32_t flags, double eps, std::map<std::string, int32_t> roles, const char* desc) {
    out->op = op;
    out->flags = flags;
    out->eps = eps;
    out->io_dtype = io;
    out->roles = std::move(roles);
    out->outputs = names;
    out->description = desc;
  };
  auto opt = [](const In* p) { return p ? p->node : kNone; };
  auto put = [](std::map<std::string, int32_t>& r, const char* k, const In* p) {
    if (p) r[k] = p->idx;
  };

  // ---- forward ops: one matrix input ----
  if (n_m == 1 && n_s == 0) {
    const In& x = mats[0];
    if (n_v == 0) {
      for (int axis : {-1, 0})
        for (bool log : {false, true})
          if (try_candidate([&](Builder& b) { return def_softmax_fwd(b, x.node, log, axis); })) {
            finish(log ? B200NORM_OP_LOGSOFTMAX_FWD : B200NORM_OP_SOFTMAX_FWD, axis == 0 ? B200NORM_F_OUTER : 0, 0.0,
                   {{"x", x.idx}}, log ? "log_softmax" : "softmax");
            return;
          }
      const struct { bool is_sum, inner; } reds[] = {{true, true}, {false, true}, {true, false}};
      for (const auto& rd : reds)
        if (try_candidate([&](Builder& b) { return def_reduce(b, x.node, rd.is_sum, rd.inner); })) {
          if (rd.inner) {
            finish(B200NORM_OP_REDUCE_INNER, rd.is_sum ? 0 : B200NORM_F_REDUCE_MAX, 0.0, {{"x", x.idx}},
                   rd.is_sum ? "sum over the last axis" : "max over the last axis");
          } else {
            finish(B200NORM_OP_REDUCE_OUTER, 0, 0.0, {{"x", x.idx}}, "sum over the outer axes");
          }
          return;
        }
      std::set<int> corrections{0};
      for (const auto& n : g.nodes)
        if (n.kind == ExprKind::Var) corrections.insert(n.correction);
      for (int corr : corrections)
        if (try_candidate([&](Builder& b) { return def_var_mean(b, x.node, corr); })) {
          if (has("var") && has("m2")) throw Unmatchable("both the variance and the sum of squared deviations requested");
          finish(B200NORM_OP_LAYERNORM_FWD, B200NORM_F_NO_Y | (has("m2") ? B200NORM_F_STAT_M2 : B200NORM_F_STAT_VAR), 0.0,
                 {{"x", x.idx}}, has("m2") ? "welford" : "var_mean");
          out->correction = corr;
          return;
        }
    }
    for (double eps : eps_list) {
      if (n_v <= 1) {
        const In* w = n_v == 1 ? &vecs[0] : nullptr;
        if (try_candidate([&](Builder& b) { return def_rms_norm_fwd(b, x.node, opt(w), b.constant(eps)); })) {
          if (has("rms") && has("invstd")) throw Unmatchable("both rms and invstd requested");
          std::map<std::string, int32_t> roles{{"x", x.idx}};
          put(roles, "w", w);
          finish(B200NORM_OP_RMSNORM_FWD, (w ? B200NORM_F_HAS_WEIGHT : 0) | (has("rms") ? B200NORM_F_STAT_IS_RMS : 0), eps,
                 roles, "rms_norm");
          return;
        }
      }
      std::vector<std::pair<const In*, const In*>> combos{{nullptr, nullptr}};
      if (n_v == 1) combos = {{&vecs[0], nullptr}};
      if (n_v == 2) combos = {{&vecs[0], &vecs[1]}, {&vecs[1], &vecs[0]}};
      for (const auto& wb : combos) {
        const In *w = wb.first, *bi = wb.second;
        std::map<std::string, int32_t> roles{{"x", x.idx}};
        put(roles, "w", w);
        put(roles, "b", bi);
        const uint32_t pf = (w ? B200NORM_F_HAS_WEIGHT : 0) | (bi ? B200NORM_F_HAS_BIAS : 0);
        if (try_candidate([&](Builder& b) { return def_layer_norm_fwd(b, x.node, opt(w), opt(bi), b.constant(eps)); })) {
          finish(B200NORM_OP_LAYERNORM_FWD, pf, eps, roles, "layer_norm");
          return;
        }
        if (try_candidate([&](Builder& b) { return def_outer_norm_fwd(b, x.node, opt(w), opt(bi), b.constant(eps)); }) &&
            has("y")) {
          finish(B200NORM_OP_LAYERNORM_FWD, B200NORM_F_OUTER | pf, eps, roles,
                 "batch_norm forward (statistics over the outer axes, channels last)");
          return;
        }
      }
    }
    throw Unmatchable("single-input fusion is not rms_norm / layer_norm / softmax / a reduction");
  }

  // ---- fused dropout (benchmarks/python/test_dropout_{layernorm,rmsnorm}_{fwd,bwd}.py) ----
  {
    std::vector<int32_t> keep_nodes;
    for (size_t i = 0; i < g.nodes.size(); ++i) {
      const ExprNode& n = g.nodes[i];
      if (n.kind == ExprKind::Lt && !n.is_scalar && !g.nodes[n.args.at(0)].is_scalar &&
          g.nodes[n.args[0]].kind == ExprKind::Uniform)
        keep_nodes.push_back((int32_t)i);
    }
    if (!keep_nodes.empty()) {
      if (keep_nodes.size() != 1 || !masks.empty() || n_m != 2 || n_s != 0 || n_v > 2)
        throw Unmatchable("dropout: one mask over norm(residual + dropout(x)) expected");
      const int32_t knode = keep_nodes[0];
      const ExprNode& kc = g.nodes[g.nodes[knode].args.at(1)];
      if (kc.kind != ExprKind::Constant || !(kc.value > 0.0) || kc.value > 1.0) throw Unmatchable("dropout keep probability");
      const double keep = kc.value;
      const uint32_t dflags = B200NORM_F_ADD_RESIDUAL | B200NORM_F_DROPOUT;
      for (int order = 0; order < 2; ++order) {
        const In &x = mats[order], &res = mats[1 - order];
        auto drop_sum = [&](Builder& b) { return b.add(res.node, b.mul(b.mul(x.node, knode), b.constant(1.0 / keep))); };
        for (double eps : eps_list) {
          if (n_v <= 1) {
            const In* w = n_v == 1 ? &vecs[0] : nullptr;
            if (try_candidate([&](Builder& b) {
                  const int32_t s = drop_sum(b);
                  Outs o = def_rms_norm_fwd(b, s, opt(w), b.constant(eps));
                  o["sum"] = s;
                  o["mask"] = knode;
                  return o;
                }) &&
                has("y")) {
              if (has("rms") && has("invstd")) throw Unmatchable("both rms and invstd requested");
              std::map<std::string, int32_t> roles{{"x", x.idx}, {"r", res.idx}};
              put(roles, "w", w);
              finish(B200NORM_OP_RMSNORM_FWD,
                     dflags | (w ? B200NORM_F_HAS_WEIGHT : 0) | (has("rms") ? B200NORM_F_STAT_IS_RMS : 0) |
                         (has("sum") ? B200NORM_F_OUT_SUM : 0),
                     eps, roles, "rms_norm(residual + dropout(x))");
              out->dropout_keep = keep;
              return;
            }
          }
          std::vector<std::pair<const In*, const In*>> combos{{nullptr, nullptr}};
          if (n_v == 1) combos = {{&vecs[0], nullptr}};
          if (n_v == 2) combos = {{&vecs[0], &vecs[1]}, {&vecs[1], &vecs[0]}};
          for (const auto& wb : combos) {
            const In *w = wb.first, *bi = wb.second;
            if (try_candidate([&](Builder& b) {
                  const int32_t s = drop_sum(b);
                  Outs o = def_layer_norm_fwd(b, s, opt(w), opt(bi), b.constant(eps));
                  o["sum"] = s;
                  o["mask"] = knode;
                  return o;
                }) &&
                has("y")) {
              std::map<std::string, int32_t> roles{{"x", x.idx}, {"r", res.idx}};
              put(roles, "w", w);
              put(roles, "b", bi);
              finish(B200NORM_OP_LAYERNORM_FWD,
                     dflags | (w ? B200NORM_F_HAS_WEIGHT : 0) | (bi ? B200NORM_F_HAS_BIAS : 0) |
                         (has("sum") ? B200NORM_F_OUT_SUM : 0),
                     eps, roles, "layer_norm(residual + dropout(x))");
              out->dropout_keep = keep;
              return;
            }
          }
        }
      }
      throw Unmatchable("a dropout mask outside norm(residual + dropout(x))");
    }
    if (!masks.empty()) {
      // backward: dy, x (the dropped branch), residual + the saved mask; the scale 1 / keep is one of the constants
      if (masks.size() != 1 || n_m != 3 || n_v > 1 || (n_s != 1 && n_s != 2))
        throw Unmatchable("dropout backward: dy, x, residual, one mask and the saved statistics expected");
      const In& mk = masks[0];
      const In* w = n_v == 1 ? &vecs[0] : nullptr;
      std::set<double> scales;
      for (const auto& n : g.nodes)
        if (n.kind == ExprKind::Constant && n.dtype != kInt && n.value >= 1.0 && n.value <= 1e6)
          scales.insert(n.value);
      for (double sc : scales) {
        std::vector<int> perm{0, 1, 2};
        do {
          const In &dy_ = mats[perm[0]], &x_ = mats[perm[1]], &r_ = mats[perm[2]];
          auto build = [&](Builder& b, int32_t mean, int32_t st, bool rms, bool centered) {
            const int32_t scn = b.constant(sc);
            const int32_t xe = b.add(r_.node, b.mul(b.mul(x_.node, mk.node), scn));
            Outs o = def_norm_bwd(b, dy_.node, xe, mean, st, opt(w), rms, centered, false);
            o["dres"] = o["dx"];
            o["dx"] = b.mul(b.mul(o["dres"], scn), mk.node);
            return o;
          };
          if (n_s == 1) {
            const struct { bool rms, centered; } forms[] = {{true, false}, {false, false}, {true, true}, {false, true}};
            for (const auto& f : forms)
              if (try_candidate([&](Builder& b) { return build(b, kNone, stats[0].node, f.rms, f.centered); }) && !has("db") &&
                  (has("dx") || has("dres"))) {
                std::map<std::string, int32_t> roles{{"dy", dy_.idx}, {"x", x_.idx}, {"r", r_.idx}, {"mask", mk.idx},
                                                     {"stat", stats[0].idx}};
                put(roles, "w", w);
                finish(B200NORM_OP_RMSNORM_BWD,
                       B200NORM_F_DROPOUT | (w ? B200NORM_F_HAS_WEIGHT : 0) | (f.rms ? B200NORM_F_STAT_IS_RMS : 0) |
                           (f.centered ? B200NORM_F_RMS_BWD_CENTERED : 0),
                       0.0, roles, "backward of rms_norm(residual + dropout(x))");
                out->dropout_keep = 1.0 / sc;
                return;
              }
          } else {
            for (int so = 0; so < 2; ++so) {
              const In &mean = stats[so], &st = stats[1 - so];
              if (try_candidate([&](Builder& b) { return build(b, mean.node, st.node, false, true); }) &&
                  (has("dx") || has("dres"))) {
                std::map<std::string, int32_t> roles{{"dy", dy_.idx}, {"x", x_.idx}, {"r", r_.idx}, {"mask", mk.idx},
                                                     {"mean", mean.idx}, {"stat", st.idx}};
                put(roles, "w", w);
                finish(B200NORM_OP_LAYERNORM_BWD, B200NORM_F_DROPOUT | (w ? B200NORM_F_HAS_WEIGHT : 0), 0.0, roles,
                       "backward of layer_norm(residual + dropout(x))");
                out->dropout_keep = 1.0 / sc;
                return;
              }
            }
          }
        } while (std::next_permutation(perm.begin(), perm.end()));
      }
      throw Unmatchable("a Bool matrix input outside a dropout backward");
    }
  }

  // ---- forward norms with a fused residual add: two matrix inputs, no statistics ----
  if (n_m == 2 && n_s == 0 && n_v <= 2) {
    const In &x = mats[0], &res = mats[1];
    const uint32_t rflags = B200NORM_F_ADD_RESIDUAL | (sum_is_rounded(g, facts) ? B200NORM_F_ROUND_SUM : 0);
    for (double eps : eps_list) {
      if (n_v <= 1) {
        const In* w = n_v == 1 ? &vecs[0] : nullptr;
        if (try_candidate([&](Builder& b) {
              const int32_t s = b.add(x.node, res.node);
              Outs o = def_rms_norm_fwd(b, s, opt(w), b.constant(eps));
              o["sum"] = s;
              return o;
            }) &&
            (has("y") || has("invstd") || has("rms"))) {
          if (has("rms") && has("invstd")) throw Unmatchable("both rms and invstd requested");
          std::map<std::string, int32_t> roles{{"x", x.idx}, {"r", res.idx}};
          put(roles, "w", w);
          finish(B200NORM_OP_RMSNORM_FWD,
                 rflags | (w ? B200NORM_F_HAS_WEIGHT : 0) | (has("rms") ? B200NORM_F_STAT_IS_RMS : 0) |
                     (has("sum") ? B200NORM_F_OUT_SUM : 0),
                 eps, roles, "rms_norm(x + residual)");
          return;
        }
      }
      std::vector<std::pair<const In*, const In*>> combos{{nullptr, nullptr}};
      if (n_v == 1) combos = {{&vecs[0], nullptr}};
      if (n_v == 2) combos = {{&vecs[0], &vecs[1]}, {&vecs[1], &vecs[0]}};
      for (const auto& wb : combos) {
        const In *w = wb.first, *bi = wb.second;
        if (try_candidate([&](Builder& b) {
              const int32_t s = b.add(x.node, res.node);
              Outs o = def_layer_norm_fwd(b, s, opt(w), opt(bi), b.constant(eps));
              o["sum"] = s;
              return o;
            }) &&
            (has("y") || has("mean") || has("invstd"))) {
          std::map<std::string, int32_t> roles{{"x", x.idx}, {"r", res.idx}};
          put(roles, "w", w);
          put(roles, "b", bi);
          finish(B200NORM_OP_LAYERNORM_FWD,
                 rflags | (w ? B200NORM_F_HAS_WEIGHT : 0) | (bi ? B200NORM_F_HAS_BIAS : 0) |
                     (has("sum") ? B200NORM_F_OUT_SUM : 0),
                 eps, roles, "layer_norm(x + residual)");
          return;
        }
      }
    }
  }

  // ---- outer normalization backward: two matrices + per-column mean, invstd (+ weight) ----
  {
    std::vector<In> pool = vecs;
    pool.insert(pool.end(), cstats.begin(), cstats.end());
    if (n_m == 2 && n_s == 0 && (pool.size() == 2 || pool.size() == 3)) {
      std::vector<int> perm(pool.size());
      for (size_t i = 0; i < perm.size(); ++i) perm[i] = (int)i;
      for (int order = 0; order < 2; ++order) {
        const In &a = mats[order], &b_ = mats[1 - order];
        std::sort(perm.begin(), perm.end());
        do {
          const In &mean = pool[perm[0]], &st = pool[perm[1]];
          const In* w = pool.size() == 3 ? &pool[perm[2]] : nullptr;
          if (N(mean).dtype != kF32 || N(st).dtype != kF32 || (w && N(*w).dtype != io)) continue;
          if (try_candidate([&](Builder& b) { return def_outer_norm_bwd(b, a.node, b_.node, mean.node, st.node, opt(w)); }) &&
              has("dx")) {
            std::map<std::string, int32_t> roles{{"dy", a.idx}, {"x", b_.idx}, {"mean", mean.idx}, {"stat", st.idx}};
            put(roles, "w", w);
            finish(B200NORM_OP_LAYERNORM_BWD, B200NORM_F_OUTER | (w ? B200NORM_F_HAS_WEIGHT : 0), 0.0, roles,
                   "batch_norm backward (statistics over the outer axes, channels last)");
            return;
          }
        } while (std::next_permutation(perm.begin(), perm.end()));
      }
    }
  }
  if (!cstats.empty()) throw Unmatchable("per-column fp32 statistics outside an outer normalization backward");

  // ---- backward ops: two matrix inputs ----
  if (n_m == 2) {
    const bool round_sum = has_intermediate_round(g, facts);
    for (int order = 0; order < 2; ++order) {
      const In &a = mats[order], &b_ = mats[1 - order];
      if (n_s == 0 && n_v == 0) {
        for (int axis : {-1, 0})
          for (bool log : {false, true})
            if (try_candidate([&](Builder& b) { return def_softmax_bwd(b, a.node, b_.node, log, axis); })) {
              finish(log ? B200NORM_OP_LOGSOFTMAX_BWD : B200NORM_OP_SOFTMAX_BWD,
                     ((round_sum && !log) ? B200NORM_F_ROUND_ROWSUM : 0) | (axis == 0 ? B200NORM_F_OUTER : 0), 0.0,
                     {{"dy", a.idx}, {"y", b_.idx}}, log ? "log_softmax_backward" : "softmax_backward");
              return;
            }
      }
      if (n_v > 1) continue;
      const In* w = n_v == 1 ? &vecs[0] : nullptr;
      if (n_s == 1) {  // RMSNorm backward
        const struct { bool rms, centered, half; } forms[] = {{false, false, false}, {false, true, false},
                                                              {true, false, false},  {true, true, false},
                                                              {true, false, true},   {false, false, true}};
        for (const auto& f : forms)
          if (try_candidate([&](Builder& b) {
                return def_norm_bwd(b, a.node, b_.node, kNone, stats[0].node, opt(w), f.rms, f.centered, f.half);
              }) &&
              !has("db")) {
            std::map<std::string, int32_t> roles{{"dy", a.idx}, {"x", b_.idx}, {"stat", stats[0].idx}};
            put(roles, "w", w);
            finish(B200NORM_OP_RMSNORM_BWD,
                   (w ? B200NORM_F_HAS_WEIGHT : 0) | (f.rms ? B200NORM_F_STAT_IS_RMS : 0) |
                       (f.centered ? B200NORM_F_RMS_BWD_CENTERED : 0) | (f.half ? B200NORM_F_RMS_BWD_HALF_PROJ : 0),
                   0.0, roles, f.half ? "thunder_rms_norm_backward (halved projection)" : "rms_norm_backward");
            return;
          }
      }
      if (n_s == 2) {  // LayerNorm backward
        for (int so = 0; so < 2; ++so) {
          const In &mean = stats[so], &st = stats[1 - so];
          if (try_candidate([&](Builder& b) {
                return def_norm_bwd(b, a.node, b_.node, mean.node, st.node, opt(w), false, true, false);
              })) {
            std::map<std::string, int32_t> roles{{"dy", a.idx}, {"x", b_.idx}, {"mean", mean.idx}, {"stat", st.idx}};
            put(roles, "w", w);
            finish(B200NORM_OP_LAYERNORM_BWD, w ? B200NORM_F_HAS_WEIGHT : 0, 0.0, roles, "layer_norm_backward");
            return;
          }
        }
      }
    }
    throw Unmatchable("two-input fusion is not a normalization / softmax backward");
  }
  throw Unmatchable("unsupported input signature");
}

bool match_segment(const ExprGraph& graph, SegmentMatch* out, std::string* why) {
  SegmentMatch m;
  try {
    for (size_t i = 0; i < graph.nodes.size(); ++i)
      for (int32_t a : graph.nodes[i].args)
        if (a < 0 || (size_t)a >= i) throw Unmatchable("malformed graph: arguments must precede their users");
    match_impl(graph, &m);
    // outputs must come in the dtypes the kernels produce: fp32 statistics, io dtype everywhere else
    for (size_t k = 0; k < graph.outputs.size(); ++k) {
      const std::string& role = m.outputs.at(k);
      const bool stat = role == "mean" || role == "invstd" || role == "rms" || role == "var" || role == "m2";
      const int32_t want = role == "mask" ? (int32_t)kBool : stat ? (int32_t)kF32 : m.io_dtype;
      if (graph.nodes[graph.outputs[k]].dtype != want)
        throw Unmatchable("output '" + role + "' has a dtype the kernels do not produce");
    }
  } catch (const Unmatchable& e) {
    if (why) *why = e.what();
    return false;
  } catch (const std::exception& e) {
    if (why) *why = std::string("malformed graph: ") + e.what();
    return false;
  }
  if (out) *out = m;
  return true;
}

}  // namespace b200norm
