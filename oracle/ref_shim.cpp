// TEST INFRASTRUCTURE ONLY -- never imported by the product path (occuseg_b200/).
//
// Shim that compiles the reference's own CPU arithmetic
//   sparseconvnet/SCN/CPU/{Convolution,Deconvolution,BatchNormalization,NetworkInNetwork}.cpp
// UNMODIFIED (they are #included from where they lie under /root/reference; nothing is copied
// into this repository) behind a stub Metadata<D> that serves rulebooks handed in from Python.
// The reference's real Metadata cannot be built here: Metadata.h:16,59 pull in google sparsehash,
// cudpp and the CUDA runtime, and GPU_GRID is hard-defined (Metadata.h:42).
//
// What the stub has to provide is read off the reference call sites:
//   m.getSubmanifoldRuleBook(inputSize, filterSize, true[, dilated_rate])   CPU/Convolution.cpp:121,163
//   m.getRuleBook(inputSize, outputSize, filterSize, filterStride, true)    CPU/Convolution.cpp:45,86; CPU/Deconvolution.cpp:15,56
//   m.getNActive(size)                                                      CPU/Convolution.cpp:47,88,122,164
//   Int, RuleBook                                                           Metadata/32bits.h:11, Metadata/Metadata.h:73
//   OptionalTensorData<T>(tensor)                                           Metadata/Metadata.h:366
#include <torch/extension.h>
#include <omp.h>
#include <cstdint>
#include <map>
#include <vector>
#include <stdexcept>

using Int = int32_t;
using RuleBook = std::vector<std::vector<Int>>;

template <typename T> T *OptionalTensorData(at::Tensor tensor) {
  return tensor.numel() ? tensor.data_ptr<T>() : nullptr;
}

static long size_key(const at::Tensor &sz) {
  // all test grids are cubes; the first entry identifies the scale
  return sz.numel() ? sz.data_ptr<long>()[0] : 0;
}

template <Int Dimension> struct Metadata {
  std::map<long, RuleBook> subm;                   // keyed by spatial size
  std::map<long, RuleBook> strided;                // keyed by the (fine) input spatial size
  std::map<long, Int> nActive;

  RuleBook &getSubmanifoldRuleBook(at::Tensor spatialSize, at::Tensor, bool, int = 1) {
    auto it = subm.find(size_key(spatialSize));
    if (it == subm.end()) throw std::runtime_error("ref_shim: no submanifold rulebook loaded for this size");
    return it->second;
  }
  RuleBook &getRuleBook(at::Tensor inputSpatialSize, at::Tensor, at::Tensor, at::Tensor, bool) {
    auto it = strided.find(size_key(inputSpatialSize));
    if (it == strided.end()) throw std::runtime_error("ref_shim: no strided rulebook loaded for this size");
    return it->second;
  }
  Int getNActive(at::Tensor spatialSize) {
    auto it = nActive.find(size_key(spatialSize));
    if (it == nActive.end()) throw std::runtime_error("ref_shim: nActive unknown for this size");
    return it->second;
  }
};

// ---- the reference's arithmetic, verbatim from its own tree (include path set by build_ref.py) ----
#include "CPU/Convolution.cpp"
#include "CPU/Deconvolution.cpp"
#include "CPU/BatchNormalization.cpp"
#include "CPU/NetworkInNetwork.cpp"

using M3 = Metadata<3>;

static RuleBook to_rulebook(const std::vector<at::Tensor> &lists) {
  RuleBook rb(lists.size());
  for (size_t k = 0; k < lists.size(); ++k) {
    auto t = lists[k].contiguous();
    TORCH_CHECK(t.scalar_type() == at::kInt, "rule lists must be int32 [n,2]");
    rb[k].assign(t.data_ptr<Int>(), t.data_ptr<Int>() + t.numel());
  }
  return rb;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  py::class_<M3>(m, "Metadata_3")
      .def(py::init<>())
      .def("load_submanifold", [](M3 &self, long size, std::vector<at::Tensor> lists) { self.subm[size] = to_rulebook(lists); })
      .def("load_strided", [](M3 &self, long in_size, std::vector<at::Tensor> lists) { self.strided[in_size] = to_rulebook(lists); })
      .def("set_nactive", [](M3 &self, long size, long n) { self.nActive[size] = (Int)n; });

  m.def("SubmanifoldConvolution_updateOutput",
        [](at::Tensor sz, at::Tensor fs, M3 &md, at::Tensor in, at::Tensor out, at::Tensor w, at::Tensor b) {
          return cpu_SubmanifoldConvolution_updateOutput<float, 3>(sz, fs, md, in, out, w, b);
        });
  m.def("SubmanifoldConvolution_backward",
        [](at::Tensor sz, at::Tensor fs, M3 &md, at::Tensor in, at::Tensor din, at::Tensor dout, at::Tensor w,
           at::Tensor dw, at::Tensor db) {
          cpu_SubmanifoldConvolution_backward<float, 3>(sz, fs, md, in, din, dout, w, dw, db, 1);
        });
  m.def("Convolution_updateOutput",
        [](at::Tensor isz, at::Tensor osz, at::Tensor fs, at::Tensor st, M3 &md, at::Tensor in, at::Tensor out,
           at::Tensor w, at::Tensor b) {
          return cpu_Convolution_updateOutput<float, 3>(isz, osz, fs, st, md, in, out, w, b);
        });
  m.def("Convolution_backward",
        [](at::Tensor isz, at::Tensor osz, at::Tensor fs, at::Tensor st, M3 &md, at::Tensor in, at::Tensor din,
           at::Tensor dout, at::Tensor w, at::Tensor dw, at::Tensor db) {
          cpu_Convolution_backward<float, 3>(isz, osz, fs, st, md, in, din, dout, w, dw, db);
        });
  m.def("Deconvolution_updateOutput",
        [](at::Tensor isz, at::Tensor osz, at::Tensor fs, at::Tensor st, M3 &md, at::Tensor in, at::Tensor out,
           at::Tensor w, at::Tensor b) {
          return cpu_Deconvolution_updateOutput<float, 3>(isz, osz, fs, st, md, in, out, w, b);
        });
  m.def("Deconvolution_backward",
        [](at::Tensor isz, at::Tensor osz, at::Tensor fs, at::Tensor st, M3 &md, at::Tensor in, at::Tensor din,
           at::Tensor dout, at::Tensor w, at::Tensor dw, at::Tensor db) {
          cpu_Deconvolution_backward<float, 3>(isz, osz, fs, st, md, in, din, dout, w, dw, db);
        });
  m.def("BatchNormalization_updateOutput",
        [](at::Tensor in, at::Tensor out, at::Tensor sm, at::Tensor si, at::Tensor rm, at::Tensor rv, at::Tensor w,
           at::Tensor b, double eps, double mom, bool train, double leak) {
          cpu_BatchNormalization_updateOutput<float>(in, out, sm, si, rm, rv, w, b, eps, mom, train, leak);
        });
  m.def("BatchNormalization_backward",
        [](at::Tensor in, at::Tensor din, at::Tensor out, at::Tensor dout, at::Tensor sm, at::Tensor si,
           at::Tensor rm, at::Tensor rv, at::Tensor w, at::Tensor b, at::Tensor dw, at::Tensor db, double leak) {
          cpu_BatchNormalization_backward<float>(in, din, out, dout, sm, si, rm, rv, w, b, dw, db, leak);
        });
  m.def("NetworkInNetwork_updateOutput", [](at::Tensor in, at::Tensor out, at::Tensor w, at::Tensor b) {
    return cpu_NetworkInNetwork_updateOutput<float>(in, out, w, b);
  });
  m.def("NetworkInNetwork_updateGradInput", [](at::Tensor din, at::Tensor dout, at::Tensor w) {
    cpu_NetworkInNetwork_updateGradInput<float>(din, dout, w);
  });
  m.def("NetworkInNetwork_accGradParameters", [](at::Tensor in, at::Tensor dout, at::Tensor dw, at::Tensor db) {
    cpu_NetworkInNetwork_accGradParameters<float>(in, dout, dw, db);
  });
  m.def("set_threads", [](int n) { omp_set_num_threads(n); at::set_num_threads(n); });
  m.def("max_threads", []() { return omp_get_max_threads(); });
}
