// Minimal stand-in for jaxlib's xla/ffi/api/ffi.h, for the build container (no jaxlib): just enough of the
// binding DSL to TYPE-CHECK vajax_b200/csrc/vjx_ffi.cc - every handler must be invocable with exactly the
// context / attribute / argument / result types its Ffi::Bind() chain declares (tests/test_host_logic.py).
// It registers nothing and runs nothing.  Test infrastructure only.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>

namespace xla {
namespace ffi {

enum DataType { F64, S32 };
template <DataType dt> struct NativeOf;
template <> struct NativeOf<F64> { using type = double; };
template <> struct NativeOf<S32> { using type = int32_t; };

class Error {
 public:
  static Error Success() { return Error(false); }
  static Error InvalidArgument(const std::string&) { return Error(true); }
  static Error Internal(const std::string&) { return Error(true); }
  bool failure() const { return failed_; }
  bool success() const { return !failed_; }

 private:
  explicit Error(bool f) : failed_(f) {}
  bool failed_;
};

template <DataType dt>
class Buffer {
 public:
  size_t element_count() const { return n_; }
  typename NativeOf<dt>::type* typed_data() const { return p_; }

 private:
  typename NativeOf<dt>::type* p_ = nullptr;
  size_t n_ = 0;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

template <typename T> struct PlatformStream {};
template <typename T> struct CtxArg { using type = T; };
template <typename T> struct CtxArg<PlatformStream<T>> { using type = T; };

template <typename... Args>
struct Binding {
  template <typename T> Binding<Args..., typename CtxArg<T>::type> Ctx() const { return {}; }
  template <typename T> Binding<Args..., T> Attr(const char*) const { return {}; }
  template <typename T> Binding<Args..., T> Arg() const { return {}; }
  template <typename T> Binding<Args..., Result<T>> Ret() const { return {}; }
  template <typename F> static constexpr bool Matches = std::is_invocable_r<Error, F, Args...>::value;
  static constexpr size_t arity = sizeof...(Args);
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding)                                              \
  static_assert(decltype(binding)::template Matches<decltype(&impl)>,                                  \
                #impl ": handler signature does not match its Ffi::Bind() chain");                      \
  extern "C" int sym##_arity() { return (int)decltype(binding)::arity; }
