// MOCK of the parts of xla/ffi/api/ffi.h that cagpjax_b200/csrc/xla_ffi_shim.cc uses.  TEST INFRASTRUCTURE ONLY.
//
// Written from memory of the jax >= 0.4.31 FFI headers (this image has no JAX), so compiling the shim against it does
// NOT prove the shim builds against the real header.  What it does prove (tests/test_host_logic.py):
//   * the shim is syntactically valid C++17 and every identifier in the handler bodies exists;
//   * every call into include/cagp_b200.h matches that header's prototypes (argument count and types);
//   * every handler's parameter list matches its binding: XLA_FFI_DEFINE_HANDLER_SYMBOL static_asserts that the
//     implementation is invocable with (context, arguments, results, attributes) in binding order.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <vector>

namespace xla {
namespace ffi {

enum DataType { F32 = 11, F64 = 12, U8 = 6, S64 = 5 };

class Error {
 public:
  static Error Success() { return Error(); }
  static Error Internal(std::string message) { return Error(std::move(message)); }
  bool success() const { return message_.empty(); }

 private:
  Error() = default;
  explicit Error(std::string message) : message_(std::move(message)) {}
  std::string message_;
};

template <typename T>
class Span {
 public:
  Span() = default;
  Span(const T* data, size_t size) : data_(data), size_(size) {}
  const T* begin() const { return data_; }
  const T* end() const { return data_ + size_; }
  size_t size() const { return size_; }
  const T& operator[](size_t i) const { return data_[i]; }

 private:
  const T* data_ = nullptr;
  size_t size_ = 0;
};

class AnyBuffer {
 public:
  DataType element_type() const { return type_; }
  Span<int64_t> dimensions() const { return Span<int64_t>(dims_.data(), dims_.size()); }
  void* untyped_data() const { return data_; }
  size_t element_count() const {
    size_t n = 1;
    for (int64_t d : dims_) n *= (size_t)d;
    return n;
  }
  size_t size_bytes() const { return element_count() * (type_ == F64 || type_ == S64 ? 8 : type_ == F32 ? 4 : 1); }

 private:
  DataType type_ = F64;
  std::vector<int64_t> dims_;
  void* data_ = nullptr;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }

 private:
  T value_;
};

template <typename T>
struct PlatformStream {};

namespace internal {
template <typename... Ts>
struct Binding {
  template <typename T>
  static T stream_of(PlatformStream<T>*);
  template <typename C>
  constexpr Binding<Ts..., decltype(stream_of(static_cast<C*>(nullptr)))> Ctx() const { return {}; }
  template <typename T>
  constexpr Binding<Ts..., T> Arg() const { return {}; }
  template <typename T>
  constexpr Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename T>
  constexpr Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename F>
  static constexpr bool Matches() { return std::is_invocable_r<Error, F, Ts...>::value; }
};
}  // namespace internal

struct Ffi {
  static constexpr internal::Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)                                                      \
  static_assert(decltype(binding)::template Matches<decltype(&impl)>(), #impl " does not match its binding");    \
  extern "C" void* symbol(void* call_frame) { return call_frame; }
