// TEST INFRASTRUCTURE ONLY -- a stand-in for jaxlib's `xla/ffi/api/ffi.h` (absent from the build container), just deep
// enough to type-check galax_b200/csrc/gx_xla_ffi.cc: the same public names and shapes the handlers use
// (Buffer / ResultBuffer / Span / Error / PlatformStream / Ffi::Bind().Ctx().Arg().Attr().Ret() /
// XLA_FFI_DEFINE_HANDLER_SYMBOL), and -- like the real binder -- a compile-time check that the implementation's
// parameter list is exactly what the binding declares, in order.  Nothing here executes.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla::ffi {

enum DataType { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };
template <DataType dt> struct NativeOf;
template <> struct NativeOf<F64> { using type = double; };
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<S64> { using type = int64_t; };
template <> struct NativeOf<U8> { using type = uint8_t; };

template <class T>
class Span {
public:
    Span(const T *p = nullptr, size_t n = 0) : p_(p), n_(n) {}
    size_t size() const { return n_; }
    const T *begin() const { return p_; }
    const T *end() const { return p_ + n_; }
private:
    const T *p_;
    size_t n_;
};

template <DataType dt>
class Buffer {
public:
    using T = typename NativeOf<dt>::type;
    T *typed_data() const { return data_; }
    size_t element_count() const { return count_; }
    Span<int64_t> dimensions() const { return {}; }
private:
    T *data_ = nullptr;
    size_t count_ = 0;
};
template <class B>
class Result {
public:
    B *operator->() { return &b_; }
    B &operator*() { return b_; }
private:
    B b_;
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

enum class ErrorCode { kOk, kInvalidArgument, kInternal, kUnimplemented };
class Error {
public:
    Error() = default;
    Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
    static Error Success() { return Error(); }
    bool failure() const { return code_ != ErrorCode::kOk; }
    bool success() const { return code_ == ErrorCode::kOk; }
private:
    ErrorCode code_ = ErrorCode::kOk;
    std::string msg_;
};

template <class T> struct PlatformStream {};

template <class A> struct ArgOf { using type = A; };                       // Arg<Buffer<dt>> -> Buffer<dt>
template <class R> struct RetOf;                                           // Ret<Buffer<dt>> -> ResultBuffer<dt>
template <DataType dt> struct RetOf<Buffer<dt>> { using type = ResultBuffer<dt>; };
template <class C> struct CtxOf;                                           // Ctx<PlatformStream<T>> -> T
template <class T> struct CtxOf<PlatformStream<T>> { using type = T; };

template <class... Ts>
struct Binding {
    template <class C> Binding<Ts..., typename CtxOf<C>::type> Ctx() const { return {}; }
    template <class A> Binding<Ts..., typename ArgOf<A>::type> Arg() const { return {}; }
    template <class R> Binding<Ts..., typename RetOf<R>::type> Ret() const { return {}; }
    template <class T> Binding<Ts..., T> Attr(const char *) const { return {}; }
    template <class Fn>
    int To(Fn) const {
        static_assert(std::is_same_v<Fn, Error (*)(Ts...)>,
                      "handler implementation does not have the parameter list its binding declares");
        return 0;
    }
};
struct Ffi {
    static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                          \
    static const int name##_binding_check = (binding).To(&impl);                    \
    extern "C" XLA_FFI_Error *name(XLA_FFI_CallFrame *) { return nullptr; }
