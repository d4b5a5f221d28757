// torch.ops.vp_b200.* — the C ABI of include/vp_b200.h registered with PyTorch's dispatcher (TORCH_LIBRARY), as SURVEY.md §8(b)
// specifies.  The wrappers are deliberately thin: a tensor becomes its raw device pointer (None -> null), the current CUDA
// stream becomes the trailing `stream` argument, a non-zero status becomes a c10::Error carrying vp_last_error().  Shapes,
// dtypes and contiguity are checked by the callers in videopainter_b200/ops.py; all arithmetic is in libvp_b200.so.
//
// One generic wrapper serves every entry point: the schema of an op is derived from the C prototype (pointer -> Tensor?,
// pointer array -> int[], integer -> int, float -> float), so the op takes exactly the C arguments in the C order, minus
// `stream`.
#include <ATen/core/Tensor.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <optional>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/vp_b200.h"

namespace {

using OptTensor = const std::optional<at::Tensor>&;

struct PtrArray {                        // int[] of device addresses -> `void* const*`
  std::vector<void*> v;
  operator void* const*() const { return v.data(); }
};

template <class T>
struct Arg;                              // C parameter type -> (schema type, conversion)
template <class P>
struct PtrArg {
  using type = OptTensor;
  static P get(OptTensor t) { return t.has_value() && t->defined() ? static_cast<P>(t->data_ptr()) : nullptr; }
};
template <> struct Arg<const void*> : PtrArg<const void*> {};
template <> struct Arg<void*> : PtrArg<void*> {};
template <> struct Arg<const float*> : PtrArg<const float*> {};
template <> struct Arg<float*> : PtrArg<float*> {};
template <> struct Arg<const int64_t*> : PtrArg<const int64_t*> {};
template <> struct Arg<const uint8_t*> : PtrArg<const uint8_t*> {};
template <> struct Arg<uint8_t*> : PtrArg<uint8_t*> {};
template <>
struct Arg<void* const*> {
  using type = at::IntArrayRef;
  static PtrArray get(at::IntArrayRef a) {
    PtrArray p;
    p.v.reserve(a.size());
    for (int64_t x : a) p.v.push_back(reinterpret_cast<void*>(static_cast<uintptr_t>(x)));
    return p;
  }
};
template <class I>
struct IntArg {
  using type = int64_t;
  static I get(int64_t v) { return static_cast<I>(v); }
};
template <> struct Arg<int> : IntArg<int> {};
template <> struct Arg<long long> : IntArg<long long> {};
template <> struct Arg<unsigned int> : IntArg<unsigned int> {};
template <>
struct Arg<float> {
  using type = double;
  static float get(double v) { return static_cast<float>(v); }
};

template <class F>
struct Traits;
template <class... A>
struct Traits<int (*)(A...)> {
  using args = std::tuple<A...>;
  static constexpr size_t n = sizeof...(A);
};

template <auto Fn, class Idx>
struct WrapImpl;
template <auto Fn, size_t... I>
struct WrapImpl<Fn, std::index_sequence<I...>> {
  using T = Traits<decltype(Fn)>;
  template <size_t K>
  using CArg = std::tuple_element_t<K, typename T::args>;
  static_assert(std::is_same_v<CArg<T::n - 1>, void*>, "the last C parameter must be the stream");
  static void call(typename Arg<CArg<I>>::type... a) {
    void* stream = static_cast<void*>(c10::cuda::getCurrentCUDAStream().stream());
    const int rc = Fn(Arg<CArg<I>>::get(a)..., stream);
    TORCH_CHECK(rc == 0, "vp_b200 kernel launch failed: rc=", rc, " (", vp_last_error(), "; cudaError=", vp_last_cuda_error(), ")");
  }
};
template <auto Fn>
using Wrap = WrapImpl<Fn, std::make_index_sequence<Traits<decltype(Fn)>::n - 1>>;

}  // namespace

#define VP_OP(name) m.def(#name, &Wrap<&vp_##name>::call)

TORCH_LIBRARY(vp_b200, m) {
  VP_OP(time_sinusoid);
  VP_OP(gemv);
  VP_OP(ln_modulate);
  VP_OP(ln_final);
  VP_OP(gemm_bias);
  VP_OP(gemm_gelu);
  VP_OP(gemm_gate_residual);
  VP_OP(gemm_qkv);
  VP_OP(gemm_qkv_peer);
  VP_OP(attention);
  VP_OP(attention_peer);
  VP_OP(peer_barrier);
  VP_OP(peer_scatter);
  VP_OP(a2a_unpack_heads);
  VP_OP(step_end);
  VP_OP(patchify);
  VP_OP(mask_pool);
  VP_OP(unpatchify);
  m.def("version", []() -> int64_t { return vp_version(); });
}
