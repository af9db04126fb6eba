// norm_executor.h — C++ host executor for matched normalization / reduction segments.
//
// Mirror of the executor the reference slots into ExecutorDispatch for segments that
// bypass lowering / codegen / NVRTC (reference: CutlassExecutor,
// csrc/runtime/cutlass_executor.h:23-66, .cpp:58-115): `supported` (side-effect free),
// `compile` (classification result cached, no NVRTC), `isCompiled`, and `run`, which
// infers output shapes, allocates them with ATen (at::empty — the same allocator the
// reference uses, cutlass_executor.cpp:85-115 / executor.cpp:1094-1131), sizes a
// zero-initialised byte workspace (executor.cpp:1160-1170) and enqueues the kernels on
// at::cuda::getCurrentCUDAStream() under a device guard (cutlass_executor.cpp:79).
//
// at::Tensor appears only on this side of the boundary; everything below it is the C ABI
// of include/b200norm.h. In an nvFuser tree this class is wrapped 1:1 by a
// `NormExecutor : ExecutorAbstract` whose `run` converts KernelArgumentHolder <->
// std::vector<at::Tensor>.
#pragma once

#include <ATen/ATen.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/b200norm.h"
#include "segment_matcher.h"

namespace b200norm {

// What the matcher decided about a segment (fuser_b200/matcher.py produces the same
// triple; in-tree it comes from walking Fusion::exprs()).
struct SegmentSpec {
  int32_t op = 0;          // b200norm_op
  uint32_t flags = 0;      // B200NORM_F_*
  double eps = 0.0;
  int32_t variant = 0;     // b200norm_variant (0 = planner's choice)
  int32_t cluster_size = 0;
  int32_t correction = 0;  // var_mean's correction (B200NORM_F_STAT_VAR), travels in desc.reserved[0]
  double dropout_keep = 1.0;  // B200NORM_F_DROPOUT: keep probability 1 - p, travels in desc.reserved[1..2]
  int32_t ignore_index = -100;  // cross-entropy (REDUCE_NLL / NLL_SOFTMAX_BWD): at::nll_loss's ignore_index, desc.reserved[3]
};

class NormExecutor {
 public:
  NormExecutor(int64_t fusion_id = 0, int64_t concrete_id = 0, int64_t runtime_id = 0, int64_t group_id = 0)
      : fusion_id_(fusion_id), concrete_id_(concrete_id), runtime_id_(runtime_id), group_id_(group_id) {}

  // True if the kernel library has a kernel for this segment at this shape.
  static bool supported(const SegmentSpec& spec, at::ScalarType io_dtype, int64_t rows, int64_t cols);

  // The classifier of ExecutorDispatch: is this segment — given as the plain-data expression graph an adapter
  // fills from fusion->exprs() — one of the normalization / reduction ops of the
  // kernel library, in any spelling? Side-effect free (csrc/runtime/executor_dispatch.cpp:46-59 asks every
  // executor in turn). On success *spec is ready for compile() and *match tells which fusion input plays which
  // role and what every fusion output is.
  static bool supported(const ExprGraph& segment, SegmentSpec* spec, SegmentMatch* match = nullptr,
                        std::string* why = nullptr);

  void compile(const SegmentSpec& spec);
  bool isCompiled() const { return compiled_; }

  // Inputs in the order of include/b200norm.h's tensor table for the op; optional
  // tensors the flags declare absent are left out. Returns the outputs in the same
  // table's order (absent optional outputs are left out). `outputs`, when non-empty,
  // supplies preallocated output tensors.
  std::vector<at::Tensor> run(const std::vector<at::Tensor>& inputs, std::vector<at::Tensor> outputs = {});

  const SegmentSpec& spec() const { return spec_; }
  int64_t groupId() const { return group_id_; }

 private:
  std::vector<at::Tensor> allocateOutputs(const std::vector<at::Tensor>& inputs, int64_t rows, int64_t cols) const;

  SegmentSpec spec_;
  bool compiled_ = false;
  int64_t fusion_id_, concrete_id_, runtime_id_, group_id_;
};

// Thrown for non-zero library status (the NVF_THROW of this boundary).
struct NormError : public std::runtime_error {
  NormError(int code_, const std::string& msg) : std::runtime_error(msg), code(code_) {}
  int code;
};

}  // namespace b200norm
