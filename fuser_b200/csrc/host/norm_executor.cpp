// norm_executor.cpp — see norm_executor.h.
#include "norm_executor.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <cstring>
#include <map>
#include <mutex>
#include <sstream>

namespace b200norm {

namespace {

int dtype_id(at::ScalarType t) {
  switch (t) {
    case at::kFloat: return B200NORM_DT_F32;
    case at::kBFloat16: return B200NORM_DT_BF16;
    case at::kHalf: return B200NORM_DT_F16;
    default: return -1;
  }
}

[[noreturn]] void raise(int code) {
  std::ostringstream os;
  os << "b200norm status " << code << ": " << b200norm_last_error();
  throw NormError(code, os.str());
}
[[noreturn]] void raise_msg(int code, const std::string& msg) {
  std::ostringstream os;
  os << "b200norm status " << code << ": " << msg;
  throw NormError(code, os.str());
}

b200norm_desc make_desc(const SegmentSpec& s, int dt, int64_t rows, int64_t cols) {
  b200norm_desc d{};
  d.abi_version = B200NORM_ABI_VERSION;
  d.op = s.op;
  d.io_dtype = dt;
  d.variant = s.variant;
  d.rows = rows;
  d.cols = cols;
  d.eps = s.eps;
  d.flags = s.flags;
  d.cluster_size = s.cluster_size;
  d.reserved[0] = s.correction;
  d.reserved[3] = s.ignore_index;
  if (s.flags & B200NORM_F_DROPOUT) {
    uint64_t bits;
    static_assert(sizeof(bits) == sizeof(s.dropout_keep), "double is 64 bits");
    memcpy(&bits, &s.dropout_keep, sizeof(bits));
    d.reserved[1] = (int32_t)(uint32_t)bits;
    d.reserved[2] = (int32_t)(uint32_t)(bits >> 32);
  }
  return d;
}

b200norm_tensor as_arg(const at::Tensor& t) {
  b200norm_tensor a{};
  a.data = t.data_ptr();
  a.dtype = dtype_id(t.scalar_type());
  a.ndim = (int32_t)t.dim();
  if (a.ndim > B200NORM_MAX_DIMS) raise_msg(B200NORM_ERR_UNSUPPORTED, "tensor rank > 4 at the kernel boundary");
  for (int i = 0; i < a.ndim; ++i) {
    a.sizes[i] = t.size(i);
    a.strides[i] = t.stride(i);
  }
  return a;
}

// [.., N] -> [M, N] view without a copy (the kernels take any row stride).
at::Tensor as_rows(const at::Tensor& t, int64_t n) {
  if (t.dim() == 2 && t.size(1) == n) return t;
  if (t.dim() >= 2 && n > 0 && t.size(0) * n == t.numel() && t.size(-1) != n) {
    // outer ops: [M, d1, d2, ..] -> [M, d1*d2*..]
    if (!t.is_contiguous())
      raise_msg(B200NORM_ERR_UNSUPPORTED, "a non-contiguous tensor of rank > 2 is not viewable as [rows, cols]");
    return t.view({t.size(0), n});
  }
  if (t.numel() == 0) return t.reshape({-1, n});
  if (!t.is_contiguous())
    raise_msg(B200NORM_ERR_UNSUPPORTED, "a non-contiguous tensor of rank != 2 is not viewable as [rows, cols]");
  return t.view({-1, n});
}

// zero-initialised workspace per (device, stream); the library keeps the header zero
at::Tensor workspace(size_t nbytes, const at::Device& dev, cudaStream_t stream) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, at::Tensor> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair((int)dev.index(), stream);
  auto it = cache.find(key);
  if (it == cache.end() || (size_t)it->second.numel() < nbytes) {
    size_t n = nbytes < (1u << 20) ? (1u << 20) : nbytes;
    at::Tensor ws = at::zeros({(int64_t)n}, at::TensorOptions().dtype(at::kByte).device(dev));
    cache[key] = ws;
    return ws;
  }
  return it->second;
}

bool is_norm_bwd(int op) { return op == B200NORM_OP_LAYERNORM_BWD || op == B200NORM_OP_RMSNORM_BWD; }

}  // namespace

bool NormExecutor::supported(const SegmentSpec& spec, at::ScalarType io_dtype, int64_t rows, int64_t cols) {
  const int dt = dtype_id(io_dtype);
  if (dt < 0) return false;
  b200norm_desc d = make_desc(spec, dt, rows, cols);
  return b200norm_query(&d, nullptr) == B200NORM_OK;
}

bool NormExecutor::supported(const ExprGraph& segment, SegmentSpec* spec, SegmentMatch* match, std::string* why) {
  SegmentMatch m;
  if (!match_segment(segment, &m, why)) return false;
  if (spec != nullptr) {
    *spec = SegmentSpec();
    spec->op = m.op;
    spec->flags = m.flags;
    spec->eps = m.eps;
    spec->correction = m.correction;
    spec->dropout_keep = m.dropout_keep;
  }
  if (match != nullptr) *match = m;
  return true;
}

void NormExecutor::compile(const SegmentSpec& spec) {
  if (spec.op < 0 || spec.op >= B200NORM_OP_COUNT_) raise_msg(B200NORM_ERR_UNSUPPORTED, "unknown segment kind");
  spec_ = spec;
  compiled_ = true;
}

std::vector<at::Tensor> NormExecutor::allocateOutputs(const std::vector<at::Tensor>& in, int64_t rows,
                                                     int64_t cols) const {
  std::vector<at::Tensor> out;
  const uint32_t f = spec_.flags;
  const at::Tensor& first = in.at(0);
  const auto opts = first.options();
  const auto f32 = opts.dtype(at::kFloat);
  auto stat_shape = [&](const at::Tensor& x) {
    if (f & B200NORM_F_OUTER) return std::vector<int64_t>{cols};  // one statistic per column (keepdim = false)
    if (f & B200NORM_F_ROW_AFFINE) return std::vector<int64_t>{x.size(0), (int64_t)spec_.ignore_index};  // [B, C]
    std::vector<int64_t> s(x.sizes().begin(), x.sizes().end());
    s.back() = 1;
    return s;
  };
  switch (spec_.op) {
    case B200NORM_OP_LAYERNORM_FWD:
      if (!(f & B200NORM_F_NO_Y)) out.push_back(at::empty(first.sizes(), opts));
      if ((f & B200NORM_F_ADD_RESIDUAL) && (f & B200NORM_F_OUT_SUM)) out.push_back(at::empty(first.sizes(), opts));
      if (f & B200NORM_F_OUT_MEAN) out.push_back(at::empty(stat_shape(first), f32));
      if (f & B200NORM_F_OUT_INVSTD) out.push_back(at::empty(stat_shape(first), f32));
      if (f & B200NORM_F_DROPOUT) out.push_back(at::empty(first.sizes(), opts.dtype(at::kBool)));  // the mask
      break;
    case B200NORM_OP_RMSNORM_FWD:
      out.push_back(at::empty(first.sizes(), opts));
      if ((f & B200NORM_F_ADD_RESIDUAL) && (f & B200NORM_F_OUT_SUM)) out.push_back(at::empty(first.sizes(), opts));
      out.push_back(at::empty(stat_shape(first), f32));
      if (f & B200NORM_F_DROPOUT) out.push_back(at::empty(first.sizes(), opts.dtype(at::kBool)));  // the mask
      break;
    case B200NORM_OP_SOFTMAX_FWD:
    case B200NORM_OP_LOGSOFTMAX_FWD:
      out.push_back(at::empty(first.sizes(), opts));
      break;
    case B200NORM_OP_SOFTMAX_BWD:
    case B200NORM_OP_LOGSOFTMAX_BWD:
      out.push_back(at::empty(in.at(1).sizes(), opts));
      break;
    case B200NORM_OP_LAYERNORM_BWD:
    case B200NORM_OP_RMSNORM_BWD: {
      const bool ln = spec_.op == B200NORM_OP_LAYERNORM_BWD;
      if (f & B200NORM_F_OUT_DX) out.push_back(at::empty(in.at(1).sizes(), opts));
      if ((f & B200NORM_F_OUT_DX) && (f & B200NORM_F_DROPOUT)) out.push_back(at::empty(in.at(1).sizes(), opts));  // d residual
      const bool want_dw = (f & B200NORM_F_OUT_DW) != 0;
      const bool want_db = ln && (f & B200NORM_F_OUT_DB) != 0;
      if (f & B200NORM_F_ROW_AFFINE) {  // instance norm: per-row sums [B, C] fp32 (the caller adds them over the batch)
        if (want_dw) out.push_back(at::empty(stat_shape(in.at(1)), f32));
        if (want_db) out.push_back(at::empty(stat_shape(in.at(1)), f32));
      } else if ((want_dw || want_db) && !(f & B200NORM_F_PEER_PUBLISH)) {  // a publish leaves the sums in the exchange slot
        if (f & B200NORM_F_PARTIALS_F32) {
          out.push_back(at::empty({ln ? 2 : 1, cols}, f32));
        } else {
          if (want_dw) out.push_back(at::empty({cols}, opts));
          if (want_db) out.push_back(at::empty({cols}, opts));
        }
      }
      break;
    }
    case B200NORM_OP_REDUCE_INNER: {
      std::vector<int64_t> s(first.sizes().begin(), first.sizes().end() - 1);
      if (f & B200NORM_F_REDUCE_NLL) {  // cross-entropy forward: per-row loss and log-sum-exp, fp32
        out.push_back(at::empty(s, f32));
        out.push_back(at::empty(s, f32));
      } else {
        out.push_back(at::empty(s, opts));
      }
      break;
    }
    case B200NORM_OP_NLL_SOFTMAX_BWD:
      out.push_back(at::empty(first.sizes(), opts));
      break;
    case B200NORM_OP_REDUCE_OUTER:
      out.push_back(at::empty({cols}, (f & B200NORM_F_PARTIALS_F32) ? f32 : opts));
      break;
    default:
      raise_msg(B200NORM_ERR_UNSUPPORTED, "allocateOutputs: segment kind not handled");
  }
  (void)rows;
  return out;
}

std::vector<at::Tensor> NormExecutor::run(const std::vector<at::Tensor>& inputs, std::vector<at::Tensor> outputs) {
  if (!compiled_) raise_msg(B200NORM_ERR_BAD_ARG, "NormExecutor must be compiled before running");
  if (inputs.empty()) raise_msg(B200NORM_ERR_BAD_ARG, "no inputs");
  for (const auto& t : inputs) {
    if (!t.is_cuda())
      raise_msg(B200NORM_ERR_NO_DEVICE, "b200norm kernels run on CUDA tensors only (there is no CPU fallback)");
  }
  const at::Tensor& lead = inputs[0];
  const c10::cuda::CUDAGuard guard(lead.device());
  cudaStream_t stream = at::cuda::getCurrentCUDAStream(lead.device().index()).stream();

  // rows / cols of the segment
  int64_t cols = 0, rows = 0;
  int dt = -1;
  if (spec_.op == B200NORM_OP_FINALIZE_PARTIALS) {
    // inputs: sums [K, N] fp32; output dtype travels in the first preallocated output
    if (outputs.empty()) raise_msg(B200NORM_ERR_BAD_ARG, "finalize_partials needs preallocated outputs");
    rows = lead.size(0);
    cols = lead.size(1);
    dt = dtype_id(outputs[0].scalar_type());
  } else if (spec_.op == B200NORM_OP_PEER_COMBINE) {
    // inputs: any CUDA tensor of the device (the exchange buffer; not passed on); K preallocated outputs [N]
    if (outputs.empty()) raise_msg(B200NORM_ERR_BAD_ARG, "peer_combine needs preallocated outputs");
    rows = (int64_t)outputs.size();
    cols = outputs[0].numel();
    dt = dtype_id(outputs[0].scalar_type());
  } else if ((spec_.flags & B200NORM_F_OUTER) && spec_.op != B200NORM_OP_LAYERNORM_FWD &&
             spec_.op != B200NORM_OP_LAYERNORM_BWD) {
    // outer-persistent softmax: the reduction runs over dim 0; everything behind it is the iteration domain
    // (the outer normalization reduces every axis but the last: the [.., C] -> [rows, C] view below)
    if (lead.dim() < 2) raise_msg(B200NORM_ERR_UNSUPPORTED, "an outer reduction needs a tensor of rank >= 2");
    rows = lead.size(0);
    cols = rows > 0 ? lead.numel() / rows : 0;
    if (rows == 0)
      for (int64_t i = 1, c = 1; i < lead.dim(); ++i) cols = (c *= lead.size(i));
    dt = dtype_id(lead.scalar_type());
  } else if (spec_.flags & B200NORM_F_ROW_AFFINE) {
    // instance norm, channels first: [B, C, spatial...] -> rows = B * C (sample, channel) pairs of the spatial extent
    const int64_t ch = spec_.ignore_index;  // the channel count travels in desc.reserved[3]
    if (lead.dim() < 3 || ch < 1 || lead.size(1) != ch)
      raise_msg(B200NORM_ERR_BAD_ARG, "instance norm expects [B, C, spatial...] with C = the spec's channel count");
    rows = lead.size(0) * ch;
    cols = rows > 0 ? lead.numel() / rows : 0;
    dt = dtype_id(lead.scalar_type());
  } else {
    cols = lead.dim() > 0 ? lead.size(-1) : 1;
    rows = cols > 0 ? lead.numel() / cols : 0;
    dt = dtype_id(lead.scalar_type());
  }
  if (dt < 0) raise_msg(B200NORM_ERR_UNSUPPORTED, "dtype is not handled by b200norm");
  b200norm_desc d = make_desc(spec_, dt, rows, cols);

  if (outputs.empty()) outputs = allocateOutputs(inputs, rows, cols);

  // marshal: matrices as [rows, cols] views, vectors flattened
  const bool add_fwd = (spec_.op == B200NORM_OP_LAYERNORM_FWD || spec_.op == B200NORM_OP_RMSNORM_FWD) &&
                       (spec_.flags & B200NORM_F_ADD_RESIDUAL);
  const int n_mat_in = (spec_.op == B200NORM_OP_FINALIZE_PARTIALS) ? 0
                       : (spec_.op == B200NORM_OP_LAYERNORM_BWD || spec_.op == B200NORM_OP_RMSNORM_BWD ||
                          spec_.op == B200NORM_OP_SOFTMAX_BWD || spec_.op == B200NORM_OP_LOGSOFTMAX_BWD || add_fwd)
                           ? 2
                           : 1;
  std::vector<at::Tensor> keep;  // views must outlive the launch call
  keep.reserve(inputs.size() + outputs.size());
  std::vector<b200norm_tensor> in_args, out_args;
  in_args.reserve(inputs.size());
  out_args.reserve(outputs.size());
  // fused dropout, backward: the residual matrix (and the mask) follow the per-row / per-column vectors
  const bool drop_bwd = is_norm_bwd(spec_.op) && (spec_.flags & B200NORM_F_DROPOUT);
  const size_t residual_at = drop_bwd && inputs.size() >= 2 ? inputs.size() - 2 : (size_t)-1;
  for (size_t i = 0; i < inputs.size() && spec_.op != B200NORM_OP_PEER_COMBINE; ++i) {
    const bool mat = (int)i < n_mat_in || i == residual_at;
    keep.push_back(mat ? as_rows(inputs[i], cols) : inputs[i].reshape({-1}));
    in_args.push_back(as_arg(keep.back()));
  }
  int n_mat_out = 0;
  switch (spec_.op) {
    case B200NORM_OP_LAYERNORM_FWD:
    case B200NORM_OP_RMSNORM_FWD:
      n_mat_out = (add_fwd && (spec_.flags & B200NORM_F_OUT_SUM)) ? 2 : 1;
      if (spec_.op == B200NORM_OP_LAYERNORM_FWD && (spec_.flags & B200NORM_F_NO_Y)) n_mat_out = 0;
      break;
    case B200NORM_OP_SOFTMAX_FWD:
    case B200NORM_OP_LOGSOFTMAX_FWD:
    case B200NORM_OP_SOFTMAX_BWD:
    case B200NORM_OP_LOGSOFTMAX_BWD:
    case B200NORM_OP_NLL_SOFTMAX_BWD:
      n_mat_out = 1;
      break;
    case B200NORM_OP_LAYERNORM_BWD:
    case B200NORM_OP_RMSNORM_BWD:
      n_mat_out = (spec_.flags & B200NORM_F_OUT_DX) ? (drop_bwd ? 2 : 1) : 0;
      break;
    default:
      n_mat_out = 0;
  }
  for (size_t i = 0; i < outputs.size(); ++i) {
    keep.push_back((int)i < n_mat_out ? as_rows(outputs[i], cols) : outputs[i].view({-1}));
    out_args.push_back(as_arg(keep.back()));
  }

  const size_t ws_bytes = b200norm_workspace_bytes(&d);
  void* ws_ptr = nullptr;
  size_t ws_n = 0;
  at::Tensor ws;
  if (ws_bytes > 0) {
    ws = workspace(ws_bytes, lead.device(), stream);
    ws_ptr = ws.data_ptr();
    ws_n = (size_t)ws.numel();
  }
  const int rc = b200norm_launch(&d, in_args.data(), (int)in_args.size(), out_args.data(), (int)out_args.size(),
                                 ws_ptr, ws_n, (void*)stream);
  if (rc != B200NORM_OK) raise(rc);
  return outputs;
}

}  // namespace b200norm
