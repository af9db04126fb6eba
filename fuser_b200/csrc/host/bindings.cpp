// bindings.cpp — pybind11 surface of the C++ host executor (what python_direct/runtime.cpp
// does for FusionExecutorCache in the reference: python/python_direct/runtime.cpp:267-305).
#include <torch/extension.h>

#include "norm_executor.h"
#include "segment_matcher.h"

namespace py = pybind11;
using b200norm::NormExecutor;
using b200norm::SegmentSpec;

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "C++ host executor of the b200norm kernel library";
  // status codes travel as (code, message); fuser_b200/ops.py re-raises them as
  // B200NormError / UnsupportedSegment
  static py::exception<b200norm::NormError> norm_error(m, "NormError", PyExc_RuntimeError);
  py::register_exception_translator([](std::exception_ptr p) {
    try {
      if (p) std::rethrow_exception(p);
    } catch (const b200norm::NormError& e) {
      py::object exc = py::reinterpret_borrow<py::object>(norm_error.ptr())(e.code, std::string(e.what()));
      PyErr_SetObject(norm_error.ptr(), exc.ptr());
    }
  });

  py::class_<SegmentSpec>(m, "SegmentSpec")
      .def(py::init([](int op, uint32_t flags, double eps, int variant, int cluster_size, int correction, double dropout_keep, int ignore_index) {
             SegmentSpec s;
             s.op = op;
             s.flags = flags;
             s.eps = eps;
             s.variant = variant;
             s.cluster_size = cluster_size;
             s.correction = correction;
             s.dropout_keep = dropout_keep;
             s.ignore_index = ignore_index;
             return s;
           }),
           py::arg("op"), py::arg("flags") = 0, py::arg("eps") = 0.0, py::arg("variant") = 0,
           py::arg("cluster_size") = 0, py::arg("correction") = 0, py::arg("dropout_keep") = 1.0, py::arg("ignore_index") = -100)
      .def_readwrite("op", &SegmentSpec::op)
      .def_readwrite("flags", &SegmentSpec::flags)
      .def_readwrite("eps", &SegmentSpec::eps)
      .def_readwrite("variant", &SegmentSpec::variant)
      .def_readwrite("cluster_size", &SegmentSpec::cluster_size)
      .def_readwrite("correction", &SegmentSpec::correction)
      .def_readwrite("dropout_keep", &SegmentSpec::dropout_keep);

  // the C++ segment classifier on a plain-data expression graph (segment_matcher.h): nodes are tuples
  // (kind, dtype, is_scalar, args, dims, red_axes, correction, value, size_class)
  m.def(
      "match_segment",
      [](const std::vector<std::tuple<int, int, bool, std::vector<int32_t>, std::vector<int32_t>, std::vector<int32_t>, int,
                                      double, int>>& nodes,
         const std::vector<int32_t>& inputs, const std::vector<int32_t>& outputs) -> py::object {
        b200norm::ExprGraph g;
        for (const auto& t : nodes) {
          b200norm::ExprNode n;
          n.kind = static_cast<b200norm::ExprKind>(std::get<0>(t));
          n.dtype = std::get<1>(t);
          n.is_scalar = std::get<2>(t);
          n.args = std::get<3>(t);
          n.dims = std::get<4>(t);
          n.red_axes = std::get<5>(t);
          n.correction = std::get<6>(t);
          n.value = std::get<7>(t);
          n.size_class = std::get<8>(t);
          g.nodes.push_back(std::move(n));
        }
        g.inputs = inputs;
        g.outputs = outputs;
        b200norm::SegmentMatch sm;
        std::string why;
        py::dict d;
        if (!b200norm::match_segment(g, &sm, &why)) {
          d["matched"] = false;
          d["why"] = why;
          return d;
        }
        d["matched"] = true;
        d["op"] = sm.op;
        d["flags"] = sm.flags;
        d["eps"] = sm.eps;
        d["io_dtype"] = sm.io_dtype;
        d["correction"] = sm.correction;
        d["dropout_keep"] = sm.dropout_keep;
        d["channels_last"] = sm.channels_last;
        d["roles"] = sm.roles;
        d["outputs"] = sm.outputs;
        d["description"] = sm.description;
        return d;
      },
      py::arg("nodes"), py::arg("inputs"), py::arg("outputs"));

  py::class_<NormExecutor>(m, "NormExecutor")
      .def(py::init<int64_t, int64_t, int64_t, int64_t>(), py::arg("fusion_id") = 0, py::arg("concrete_id") = 0,
           py::arg("runtime_id") = 0, py::arg("group_id") = 0)
      .def_static("supported", static_cast<bool (*)(const SegmentSpec&, at::ScalarType, int64_t, int64_t)>(&NormExecutor::supported))
      .def("compile", &NormExecutor::compile)
      .def("is_compiled", &NormExecutor::isCompiled)
      .def("run", &NormExecutor::run, py::arg("inputs"), py::arg("outputs") = std::vector<at::Tensor>{});
}
