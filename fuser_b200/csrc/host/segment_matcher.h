// segment_matcher.h — the C++ segment classifier behind NormExecutor::supported(Fusion*).
//
// In the reference, ExecutorDispatch::makeExecutor asks each executor type `supported(fusion)`
// (csrc/runtime/executor_dispatch.cpp:46-59; CutlassExecutor::supported, cutlass_executor.cpp:38-56) and the
// first that accepts compiles and runs the segment. This library's kernels are fixed, so the question is "is this
// segment, as written, one of the normalization / reduction ops the kernels implement — in ANY of the spellings
// the frontends produce (C++ op builders of csrc/ops/normalization.cpp, Thunder traces of benchmarks/python)?".
// It is answered on a small plain-data expression graph (ExprGraph) so that the classifier itself depends on
// neither nvFuser's IR headers nor Python:
//   * in an nvFuser tree a ~100-line adapter fills the graph by walking fusion->exprs() (UnaryOp, BinaryOp,
//     ReductionOp, WelfordOp, BroadcastOp, ...; extent classes from the exact IterDomain map) — INTEGRATION.md §2.2;
//   * stand-alone, fuser_b200/_export.py fills it from the recorded FusionDefinition; tests/test_cpp_matcher.py
//     checks that this classifier and the Python one (fuser_b200/matcher.py) agree on every spelling of the hot path.
// Method (same as matcher.py): every output is canonicalised to a polynomial normal form over atoms — inputs,
// extents, sums / maxima of canonical sub-expressions, exp / log, opaque non-integer powers — and compared, term by
// term, with the normal forms of candidate op definitions instantiated on the segment's own inputs.
#pragma once

#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace b200norm {

enum class ExprKind : int32_t {
  Input = 0,   // fusion input tensor
  Constant,    // scalar constant (value)
  Size,        // scalar: extent of an axis (size_class)
  Cast,        // value preserving in the algebra (decides io dtypes, see has_intermediate_round)
  Broadcast,   // broadcast / broadcast_in_dim / expand / squeeze: a view
  Neg, Reciprocal, Sqrt, Rsqrt, Exp, Log,
  Add, Sub, Mul, Div, Pow,
  Sum, Max,    // reductions over red_axes of args[0]
  Mean, Var, M2,  // WelfordOp outputs: avg, var (with correction), var_sum
  Uniform,        // RNGOp uniform(args[0], args[1]) of the tensor's shape (the dropout prologue)
  Lt              // BinaryOp lt -> Bool; only uniform(0, 1) < constant (the dropout mask) is understood
};

enum ExprDtype : int32_t { kF32 = 0, kBF16 = 1, kF16 = 2, kOtherFloat = 3, kInt = 4, kBool = 5 };

struct ExprNode {
  ExprKind kind = ExprKind::Input;
  int32_t dtype = kF32;
  bool is_scalar = false;
  std::vector<int32_t> args;      // node indices (always smaller than the node's own index)
  std::vector<int32_t> dims;      // tensors: per logical axis the extent class (>= 0) or -1 for a broadcast axis
  std::vector<int32_t> red_axes;  // reductions: logical axes of args[0] that are reduced
  int32_t correction = 0;         // Var
  double value = 0.0;             // Constant
  int32_t size_class = -1;        // Size
};

struct ExprGraph {
  std::vector<ExprNode> nodes;
  std::vector<int32_t> inputs;   // node indices, in fusion->inputs() order (scalar inputs included)
  std::vector<int32_t> outputs;  // node indices, in fusion->outputs() order
};

// What the classifier decided: the (op, flags, eps, correction) of a b200norm_desc, which fusion input plays
// which role, and the role of every fusion output (in order).
struct SegmentMatch {
  int32_t op = -1;
  uint32_t flags = 0;
  double eps = 0.0;
  int32_t io_dtype = kF32;
  int32_t correction = 0;
  double dropout_keep = 1.0;  // B200NORM_F_DROPOUT: the keep probability 1 - p
  bool channels_last = false;  // B200NORM_F_ROW_AFFINE (instance norm): [B, spatial..., C] instead of [B, C, spatial...]
  std::map<std::string, int32_t> roles;  // "x", "dy", "y", "w", "b", "r", "mean", "stat", "mask" -> index into graph.inputs
  std::vector<std::string> outputs;      // "y", "mean", "invstd", "rms", "sum", "dx", "dw", "db", "r", "var", "m2", "mask", "dres"
  std::string description;
};

// True and *out filled if the graph is a segment the kernel library implements; otherwise false and *why says
// what stopped it. Side-effect free and re-entrant (supported() runs on the compile thread pool).
bool match_segment(const ExprGraph& graph, SegmentMatch* out, std::string* why);

}  // namespace b200norm
