// bindings.cpp — pybind11 surface of the C++ host executor (what python_direct/runtime.cpp
// does for FusionExecutorCache in the reference: python/python_direct/runtime.cpp:267-305).
#include <torch/extension.h>

#include "norm_executor.h"

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
      .def(py::init([](int op, uint32_t flags, double eps, int variant, int cluster_size, int correction) {
             SegmentSpec s;
             s.op = op;
             s.flags = flags;
             s.eps = eps;
             s.variant = variant;
             s.cluster_size = cluster_size;
             s.correction = correction;
             return s;
           }),
           py::arg("op"), py::arg("flags") = 0, py::arg("eps") = 0.0, py::arg("variant") = 0,
           py::arg("cluster_size") = 0, py::arg("correction") = 0)
      .def_readwrite("op", &SegmentSpec::op)
      .def_readwrite("flags", &SegmentSpec::flags)
      .def_readwrite("eps", &SegmentSpec::eps)
      .def_readwrite("variant", &SegmentSpec::variant)
      .def_readwrite("cluster_size", &SegmentSpec::cluster_size)
      .def_readwrite("correction", &SegmentSpec::correction);

  py::class_<NormExecutor>(m, "NormExecutor")
      .def(py::init<int64_t, int64_t, int64_t, int64_t>(), py::arg("fusion_id") = 0, py::arg("concrete_id") = 0,
           py::arg("runtime_id") = 0, py::arg("group_id") = 0)
      .def_static("supported", &NormExecutor::supported)
      .def("compile", &NormExecutor::compile)
      .def("is_compiled", &NormExecutor::isCompiled)
      .def("run", &NormExecutor::run, py::arg("inputs"), py::arg("outputs") = std::vector<at::Tensor>{});
}
