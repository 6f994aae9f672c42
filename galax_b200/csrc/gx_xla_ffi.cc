// gx_xla_ffi.cc -- XLA FFI custom-call handlers wrapping the C ABI (include/galax_b200.h).
//
// NOT part of the default build: it needs the XLA FFI headers that ship inside jaxlib
// (`python -c "import jax.ffi; print(jax.ffi.include_dir())"`), which are absent from the build container.  What can be
// checked here is checked: tests/test_abi.py compiles this file against tests/fake_xla/xla/ffi/api/ffi.h, a stand-in with
// the same public names whose binder verifies -- as the real one does -- that every handler's parameter list is exactly
// what its Ffi::Bind() chain declares.  Where jaxlib is installed:
//
//   nvcc -std=c++17 -shared -Xcompiler -fPIC -I$(python -c "import jax.ffi as f; print(f.include_dir())") \
//        -I include -o galax_b200/libgalax_b200_ffi.so galax_b200/csrc/gx_xla_ffi.cc -Lgalax_b200 -lgalax_b200
//
// and on the Python side:
//
//   jax.ffi.register_ffi_target("gx_integrate_fixed", jax.ffi.pycapsule(lib.GxIntegrateFixed), platform="CUDA")
//   q, p, status = jax.ffi.ffi_call("gx_integrate_fixed", (ShapeDtypeStruct((N,T,3), f64),)*2 + (ShapeDtypeStruct((N,), i32),),
//                                   vmap_method="broadcast_all")(q0, p0, ts, pot=pot_bytes, t0=..., t1=..., dt0=..., ...)
//
// The handlers only enqueue on XLA's stream (ffi::PlatformStream<cudaStream_t>) and never synchronise.
#include <cuda_runtime.h>

#include <cstring>

#include "../../include/galax_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error FromRc(int rc) {
    if (rc == 0) return ffi::Error::Success();
    return ffi::Error(rc == GX_ERR_CUDA ? ffi::ErrorCode::kInternal : ffi::ErrorCode::kInvalidArgument, gx_strerror(rc));
}

static ffi::Error PotFromBytes(ffi::Span<const uint8_t> bytes, gx_potential *pot) {
    if (bytes.size() != sizeof(gx_potential))
        return ffi::Error(ffi::ErrorCode::kInvalidArgument, "pot attribute must be sizeof(gx_potential) bytes");
    std::memcpy(pot, bytes.begin(), sizeof(gx_potential));
    return ffi::Error::Success();
}

static ffi::Error PotentialEvalImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> xyz, ffi::Span<const uint8_t> pot_b,
                                    double t, ffi::ResultBuffer<ffi::F64> acc, ffi::ResultBuffer<ffi::F64> hess) {
    gx_potential pot;
    if (auto e = PotFromBytes(pot_b, &pot); e.failure()) return e;
    int64_t n = xyz.element_count() / 3;
    return FromRc(gx_potential_eval(&pot, xyz.typed_data(), t, n, GX_ACC | GX_HESS, nullptr, nullptr,
                                    acc->typed_data(), hess->typed_data(), stream));
}

static ffi::Error IntegrateFixedImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> q0, ffi::Buffer<ffi::F64> p0,
                                     ffi::Buffer<ffi::F64> ts, ffi::Span<const uint8_t> pot_b, double t0, double t1,
                                     double dt0, int32_t scheme, int64_t max_steps, ffi::ResultBuffer<ffi::F64> q,
                                     ffi::ResultBuffer<ffi::F64> p, ffi::ResultBuffer<ffi::S32> status) {
    gx_potential pot;
    if (auto e = PotFromBytes(pot_b, &pot); e.failure()) return e;
    int64_t n = q0.element_count() / 3;
    int32_t T = (int32_t)ts.element_count();
    return FromRc(gx_integrate_fixed(&pot, q0.typed_data(), p0.typed_data(), n, t0, t1, dt0, ts.typed_data(), T, scheme,
                                     max_steps, GX_LAYOUT_NT3, q->typed_data(), p->typed_data(), status->typed_data(),
                                     stream));
}

static ffi::Error IntegrateDopri8Impl(cudaStream_t stream, ffi::Buffer<ffi::F64> q0, ffi::Buffer<ffi::F64> p0,
                                      ffi::Buffer<ffi::F64> t0, ffi::Buffer<ffi::F64> ts,
                                      ffi::Span<const uint8_t> pot_b, ffi::Span<const uint8_t> pid_b, double t1,
                                      int64_t max_steps, ffi::ResultBuffer<ffi::F64> q, ffi::ResultBuffer<ffi::F64> p,
                                      ffi::ResultBuffer<ffi::S32> status, ffi::ResultBuffer<ffi::S32> n_acc,
                                      ffi::ResultBuffer<ffi::S32> n_att, ffi::ResultBuffer<ffi::S64> workspace) {
    gx_potential pot;
    if (auto e = PotFromBytes(pot_b, &pot); e.failure()) return e;
    if (pid_b.size() != sizeof(gx_pid)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "bad pid attribute");
    gx_pid pid;
    std::memcpy(&pid, pid_b.begin(), sizeof pid);
    int64_t n = q0.element_count() / 3;
    int32_t T = (int32_t)ts.element_count();
    return FromRc(gx_integrate_dopri8(&pot, &pid, q0.typed_data(), p0.typed_data(), n, t0.typed_data(), 0.0, t1,
                                      ts.typed_data(), T, max_steps, nullptr, GX_LAYOUT_NT3, q->typed_data(),
                                      p->typed_data(), status->typed_data(), n_acc->typed_data(), n_att->typed_data(),
                                      workspace->typed_data(), stream));
}

// The reference's joint batch semantics (one shared adaptive step for the whole (N, 3) batch): what galax's
// evaluate_orbit / OrbitSolver.solve compute for scalar times, hence the handler a jit-compiled galax would call.
static ffi::Error IntegrateJointImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> q0, ffi::Buffer<ffi::F64> p0,
                                     ffi::Buffer<ffi::F64> ts, ffi::Span<const uint8_t> pot_b,
                                     ffi::Span<const uint8_t> pid_b, double t0, double t1, int64_t max_steps,
                                     ffi::ResultBuffer<ffi::F64> q, ffi::ResultBuffer<ffi::F64> p,
                                     ffi::ResultBuffer<ffi::S32> status, ffi::ResultBuffer<ffi::S32> n_acc,
                                     ffi::ResultBuffer<ffi::S32> n_att, ffi::ResultBuffer<ffi::F64> workspace) {
    gx_potential pot;
    if (auto e = PotFromBytes(pot_b, &pot); e.failure()) return e;
    if (pid_b.size() != sizeof(gx_pid)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "bad pid attribute");
    gx_pid pid;
    std::memcpy(&pid, pid_b.begin(), sizeof pid);
    int64_t n = q0.element_count() / 3;
    if ((int64_t)workspace->element_count() * 8 < gx_joint_workspace_bytes(n))
        return ffi::Error(ffi::ErrorCode::kInvalidArgument, "workspace smaller than gx_joint_workspace_bytes(N)");
    return FromRc(gx_integrate_adaptive_joint(GX_SOLVER_DOPRI8, &pot, &pid, q0.typed_data(), p0.typed_data(), n, t0, t1,
                                              ts.typed_data(), (int32_t)ts.element_count(), max_steps, GX_LAYOUT_NT3,
                                              q->typed_data(), p->typed_data(), status->typed_data(), n_acc->typed_data(),
                                              n_att->typed_data(), workspace->typed_data(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(GxIntegrateJoint, IntegrateJointImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<ffi::Span<const uint8_t>>("pot")
                                  .Attr<ffi::Span<const uint8_t>>("pid")
                                  .Attr<double>("t0")
                                  .Attr<double>("t1")
                                  .Attr<int64_t>("max_steps")
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(GxPotentialEval, PotentialEvalImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<ffi::Span<const uint8_t>>("pot")
                                  .Attr<double>("t")
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(GxIntegrateFixed, IntegrateFixedImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<ffi::Span<const uint8_t>>("pot")
                                  .Attr<double>("t0")
                                  .Attr<double>("t1")
                                  .Attr<double>("dt0")
                                  .Attr<int32_t>("scheme")
                                  .Attr<int64_t>("max_steps")
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(GxIntegrateDopri8, IntegrateDopri8Impl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<ffi::Span<const uint8_t>>("pot")
                                  .Attr<ffi::Span<const uint8_t>>("pid")
                                  .Attr<double>("t1")
                                  .Attr<int64_t>("max_steps")
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S64>>());
