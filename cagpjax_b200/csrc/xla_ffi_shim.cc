// XLA FFI shim: exposes every compute entry point of include/cagp_b200.h as a JAX custom call
// (jax.ffi.register_ffi_target / jax.ffi.ffi_call), which is how the reference (pure Python/JAX)
// would bind libcagp_b200.so.  One handler per C symbol; both float types go through xla::ffi::AnyBuffer.
//
// This image has no JAX, hence no xla/ffi/api/ffi.h: the whole file is guarded and compiles to an EMPTY
// translation unit here (tests/test_host_logic.py::test_xla_ffi_shim_* checks the guard and that every
// compute symbol of the header has a handler).  Where JAX is installed, build with
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c 'import jax.ffi; print(jax.ffi.include_dir())')
//       -I/usr/local/cuda/include -Iinclude cagpjax_b200/csrc/xla_ffi_shim.cc
//       -Lcagpjax_b200/_lib -lcagp_b200 -o cagpjax_b200/_lib/libcagp_b200_xla.so       (one command line)
// The Python side is jax_binding/cagpjax_b200_jax/ (custom_vjp wrappers, compute engine, congruence overload).
// Never built against the real header (written against the jax >= 0.4.31 FFI API from memory).  tests/test_host_logic.py
// compiles it against a MOCK of that API (tests/mock_xla) and the real include/cagp_b200.h: syntax, every C-ABI call's
// argument list, and handler parameters vs bindings are checked that way; the real header's details are not.
//
// Conventions: scalar sizes travel as int64 attributes; workspaces are extra RESULT buffers sized by the Python
// side with the cagp_*_ws_bytes() helpers (XLA owns all memory); errors map to ffi::Error::Internal with
// cagp_last_error_string().
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define CAGP_HAVE_XLA_FFI 1
#endif
#endif

#ifdef CAGP_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include <cstdint>
#include <vector>

#include "../../include/cagp_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using Buf = ffi::AnyBuffer;
using Out = ffi::Result<ffi::AnyBuffer>;

namespace {

inline int32_t dtype_of(const Buf& b) { return b.element_type() == ffi::F64 ? CAGP_F64 : CAGP_F32; }
inline ffi::Error status(int rc) { return rc == CAGP_OK ? ffi::Error::Success() : ffi::Error::Internal(cagp_last_error_string()); }
inline cagp_kernel_desc desc(int64_t family, const Buf& ls, const void* bbox = nullptr) {
  cagp_kernel_desc kd;
  kd.family = (int32_t)family;
  kd.dtype = dtype_of(ls);
  kd.n_lengthscale = (int32_t)ls.element_count();
  kd.reserved = 0;
  kd.lengthscale = ls.untyped_data();
  kd.bbox = bbox;
  return kd;
}
// an empty (0-element) bbox operand means "no bounding box"
inline const void* opt(const Buf& b) { return b.element_count() ? b.untyped_data() : nullptr; }

// ---- cagp_scale_inputs: X (n, d), lengthscale, origin (d or 0 elements) -> xs (dpad, n) ----------------------------
ffi::Error ScaleInputs(cudaStream_t st, Buf X, Buf ls, Buf origin, Out xs, int64_t family) {
  const int64_t n = X.dimensions()[0], d = X.dimensions()[1];
  cagp_kernel_desc kd = desc(family, ls);
  return status(cagp_scale_inputs(&kd, X.untyped_data(), n, (int32_t)d, opt(origin), xs->untyped_data(), n, st));
}
// ---- cagp_bounding_box: xs (dpad, n) -> bbox (2 dpad) ---------------------------------------------------------------
ffi::Error BoundingBox(cudaStream_t st, Buf xs, Out bbox) {
  const int64_t dp = xs.dimensions()[0], n = xs.dimensions()[1];
  return status(cagp_bounding_box(dtype_of(xs), xs.untyped_data(), n, n, (int32_t)dp, bbox->untyped_data(), st));
}
// ---- cagp_ks_blocksparse_fwd ---------------------------------------------------------------------------------------
ffi::Error KsFwd(cudaStream_t st, Buf xs1, Buf xs2, Buf s, Buf ls, Buf bbox, Out Mt, int64_t family, int64_t d, int64_t k) {
  const int64_t m = xs1.dimensions()[1], n = xs2.dimensions()[1];
  cagp_kernel_desc kd = desc(family, ls, opt(bbox));
  return status(cagp_ks_blocksparse_fwd(&kd, xs1.untyped_data(), m, m, xs2.untyped_data(), n, n, (int32_t)d, s.untyped_data(),
                                        (int32_t)k, Mt->untyped_data(), m, st));
}
// ---- cagp_ks_blocksparse_bwd: results s_bar (n), ls_bar, ws (bytes) -------------------------------------------------
ffi::Error KsBwd(cudaStream_t st, Buf xs1, Buf xs2, Buf s, Buf ls, Buf Gt, Out s_bar, Out ls_bar, Out ws, int64_t family,
                 int64_t d, int64_t k) {
  const int64_t m = xs1.dimensions()[1], n = xs2.dimensions()[1];
  cagp_kernel_desc kd = desc(family, ls);
  return status(cagp_ks_blocksparse_bwd(&kd, xs1.untyped_data(), m, m, xs2.untyped_data(), n, n, (int32_t)d, s.untyped_data(),
                                        (int32_t)k, Gt.untyped_data(), Gt.dimensions()[1], s_bar->untyped_data(),
                                        ls_bar->untyped_data(), ws->untyped_data(), ws->size_bytes(), st));
}
// ---- cagp_ks_blocksparse_fwd_sym -----------------------------------------------------------------------------------
ffi::Error KsFwdSym(cudaStream_t st, Buf xs, Buf s, Buf ls, Buf bbox, Out Mt, int64_t family, int64_t d, int64_t k,
                    int64_t a_begin, int64_t a_end) {
  const int64_t n = xs.dimensions()[1];
  cagp_kernel_desc kd = desc(family, ls, opt(bbox));
  return status(cagp_ks_blocksparse_fwd_sym(&kd, xs.untyped_data(), n, n, (int32_t)d, s.untyped_data(), (int32_t)k,
                                            (int32_t)a_begin, (int32_t)a_end, Mt->untyped_data(), Mt->dimensions()[1], st));
}
// ---- cagp_ks_blocksparse_fwd_sym_peer: peer panel addresses and the row partition are int64 attribute spans ---------
ffi::Error KsFwdSymPeer(cudaStream_t st, Buf xs, Buf s, Buf ls, Buf bbox, Out own_panel, int64_t family, int64_t d, int64_t k,
                        int64_t a_begin, int64_t a_end, ffi::Span<const int64_t> panels, ffi::Span<const int64_t> row_begin,
                        int64_t me) {
  const int64_t n = xs.dimensions()[1];
  const int world = (int)panels.size();
  std::vector<void*> ptrs(world);
  for (int r = 0; r < world; ++r) ptrs[r] = r == me ? own_panel->untyped_data() : reinterpret_cast<void*>(panels[r]);
  cagp_kernel_desc kd = desc(family, ls, opt(bbox));
  return status(cagp_ks_blocksparse_fwd_sym_peer(&kd, xs.untyped_data(), n, n, (int32_t)d, s.untyped_data(), (int32_t)k,
                                                 (int32_t)a_begin, (int32_t)a_end, ptrs.data(), row_begin.begin(), world,
                                                 (int32_t)me, own_panel->dimensions()[1], st));
}
// ---- cagp_ks_blocksparse_bwd_sym -----------------------------------------------------------------------------------
ffi::Error KsBwdSym(cudaStream_t st, Buf xs, Buf s, Buf ls, Buf bbox, Buf Gt, Out s_bar, Out ls_bar, Out ws, int64_t family,
                    int64_t d, int64_t k, int64_t a_begin, int64_t a_end) {
  const int64_t n = xs.dimensions()[1];
  cagp_kernel_desc kd = desc(family, ls, opt(bbox));
  return status(cagp_ks_blocksparse_bwd_sym(&kd, xs.untyped_data(), n, n, (int32_t)d, s.untyped_data(), (int32_t)k,
                                            (int32_t)a_begin, (int32_t)a_end, Gt.untyped_data(), Gt.dimensions()[1],
                                            s_bar->untyped_data(), ls_bar->untyped_data(), ws->untyped_data(),
                                            ws->size_bytes(), st));
}
// ---- cagp_project_rows: Mt (k, m), s, yt (local rows) -> A (k, k), c (k) ---------------------------------------------
ffi::Error ProjectRows(cudaStream_t st, Buf Mt, Buf s, Buf yt, Out A, Out c, int64_t row_offset, int64_t n_total) {
  const int64_t k = Mt.dimensions()[0], m = Mt.dimensions()[1];
  return status(cagp_project_rows(dtype_of(Mt), Mt.untyped_data(), m, m, row_offset, n_total, (int32_t)k, s.untyped_data(),
                                  yt.untyped_data(), A->untyped_data(), c->untyped_data(), st));
}
// ---- cagp_gram_mtm: Mt -> B (k, k); ws result sized by cagp_gram_ws_bytes ---------------------------------------------
ffi::Error GramMtm(cudaStream_t st, Buf Mt, Out B, Out ws) {
  const int64_t k = Mt.dimensions()[0], m = Mt.dimensions()[1];
  return status(cagp_gram_mtm(dtype_of(Mt), Mt.untyped_data(), m, m, (int32_t)k, B->untyped_data(), ws->untyped_data(),
                              ws->size_bytes(), st));
}
// ---- cagp_assemble_cotangent -----------------------------------------------------------------------------------------
ffi::Error AssembleCotangent(cudaStream_t st, Buf Mt, Buf W, Buf Abar, Buf cbar, Buf s, Buf yt, Out Gt, int64_t row_offset,
                             int64_t n_total) {
  const int64_t k = Mt.dimensions()[0], m = Mt.dimensions()[1];
  return status(cagp_assemble_cotangent(dtype_of(Mt), Mt.untyped_data(), m, m, row_offset, n_total, (int32_t)k, W.untyped_data(),
                                        Abar.untyped_data(), cbar.untyped_data(), s.untyped_data(), yt.untyped_data(),
                                        Gt->untyped_data(), m, st));
}
// ---- cagp_assemble_cotangent_peer: peer buffer addresses / block partition as int64 attribute spans -------------------
ffi::Error AssembleCotangentPeer(cudaStream_t st, Buf Mt, Buf W, Buf Abar, Buf cbar, Buf s, Buf yt, Out own_buffer,
                                 int64_t row_offset, int64_t n_total, ffi::Span<const int64_t> buffers,
                                 ffi::Span<const int64_t> block_begin, int64_t me) {
  const int64_t k = Mt.dimensions()[0], m = Mt.dimensions()[1];
  const int world = (int)buffers.size();
  std::vector<void*> ptrs(world);
  std::vector<int32_t> bb(world + 1);
  for (int r = 0; r < world; ++r) ptrs[r] = r == me ? own_buffer->untyped_data() : reinterpret_cast<void*>(buffers[r]);
  for (int r = 0; r <= world; ++r) bb[r] = (int32_t)block_begin[r];
  return status(cagp_assemble_cotangent_peer(dtype_of(Mt), Mt.untyped_data(), m, m, row_offset, n_total, (int32_t)k,
                                             W.untyped_data(), Abar.untyped_data(), cbar.untyped_data(), s.untyped_data(),
                                             yt.untyped_data(), ptrs.data(), bb.data(), world, (int32_t)me,
                                             own_buffer->dimensions()[1], st));
}
// ---- cagp_project_rows_bwd -------------------------------------------------------------------------------------------
ffi::Error ProjectRowsBwd(cudaStream_t st, Buf Mt, Buf Abar, Buf cbar, Out s_dir, Out yt_bar, int64_t row_offset, int64_t n_total) {
  const int64_t k = Mt.dimensions()[0], m = Mt.dimensions()[1];
  return status(cagp_project_rows_bwd(dtype_of(Mt), Mt.untyped_data(), m, m, row_offset, n_total, (int32_t)k, Abar.untyped_data(),
                                      cbar.untyped_data(), s_dir->untyped_data(), yt_bar->untyped_data(), st));
}
// ---- cagp_predict_moments ----------------------------------------------------------------------------------------------
ffi::Error PredictMoments(cudaStream_t st, Buf Mt, Buf w, Buf Pinv, Buf scalars, Out mean, Out var, Out ws) {
  const int64_t k = Mt.dimensions()[0], m = Mt.dimensions()[1];
  return status(cagp_predict_moments(dtype_of(Mt), Mt.untyped_data(), m, m, (int32_t)k, w.untyped_data(), Pinv.untyped_data(),
                                     scalars.untyped_data(), mean->untyped_data(), var->untyped_data(), ws->untyped_data(),
                                     ws->size_bytes(), st));
}
// ---- cagp_ks_dense_fwd: V (n, p) -> Y (m, p) ---------------------------------------------------------------------------
ffi::Error KsDenseFwd(cudaStream_t st, Buf xs1, Buf xs2, Buf V, Buf ls, Out Y, int64_t family, int64_t d) {
  const int64_t m = xs1.dimensions()[1], n = xs2.dimensions()[1], p = V.dimensions()[1];
  cagp_kernel_desc kd = desc(family, ls);
  return status(cagp_ks_dense_fwd(&kd, xs1.untyped_data(), m, m, xs2.untyped_data(), n, n, (int32_t)d, V.untyped_data(), (int32_t)p,
                                  p, Y->untyped_data(), p, st));
}
// ---- cagp_ks_dense_bwd_ls: Vt (p, n), Ybt (p, m) -> ls_bar --------------------------------------------------------------
ffi::Error KsDenseBwdLs(cudaStream_t st, Buf xs1, Buf xs2, Buf Vt, Buf Ybt, Buf ls, Out ls_bar, Out ws, int64_t family, int64_t d) {
  const int64_t m = xs1.dimensions()[1], n = xs2.dimensions()[1], p = Vt.dimensions()[0];
  cagp_kernel_desc kd = desc(family, ls);
  return status(cagp_ks_dense_bwd_ls(&kd, xs1.untyped_data(), m, m, xs2.untyped_data(), n, n, (int32_t)d, Vt.untyped_data(), n,
                                     Ybt.untyped_data(), m, (int32_t)p, ls_bar->untyped_data(), ws->untyped_data(),
                                     ws->size_bytes(), st));
}
// ---- cagp_cross_cov_fwd / cagp_cross_cov_bwd ----------------------------------------------------------------------------
ffi::Error CrossCovFwd(cudaStream_t st, Buf xs1, Buf xs2, Buf ls, Out C, int64_t family, int64_t d) {
  const int64_t m = xs1.dimensions()[1], n = xs2.dimensions()[1];
  cagp_kernel_desc kd = desc(family, ls);
  return status(cagp_cross_cov_fwd(&kd, xs1.untyped_data(), m, m, xs2.untyped_data(), n, n, (int32_t)d, C->untyped_data(), n, st));
}
ffi::Error CrossCovBwd(cudaStream_t st, Buf xs1, Buf xs2, Buf Cbar, Buf ls, Out x2s_bar, Out ls_bar, Out ws, int64_t family,
                       int64_t d) {
  const int64_t m = xs1.dimensions()[1], n = xs2.dimensions()[1];
  cagp_kernel_desc kd = desc(family, ls);
  return status(cagp_cross_cov_bwd(&kd, xs1.untyped_data(), m, m, xs2.untyped_data(), n, n, (int32_t)d, Cbar.untyped_data(), n,
                                   x2s_bar->untyped_data(), ls_bar->untyped_data(), ws->untyped_data(), ws->size_bytes(), st));
}

}  // namespace

#define CAGP_BIND() ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
#define ARG .Arg<Buf>()
#define RET .Ret<Buf>()
#define I64(name) .Attr<int64_t>(name)

XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpScaleInputs, ScaleInputs, CAGP_BIND() ARG ARG ARG RET I64("family"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpBoundingBox, BoundingBox, CAGP_BIND() ARG RET);
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpKsBlocksparseFwd, KsFwd, CAGP_BIND() ARG ARG ARG ARG ARG RET I64("family") I64("d") I64("k"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpKsBlocksparseBwd, KsBwd,
                              CAGP_BIND() ARG ARG ARG ARG ARG RET RET RET I64("family") I64("d") I64("k"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpKsBlocksparseFwdSym, KsFwdSym,
                              CAGP_BIND() ARG ARG ARG ARG RET I64("family") I64("d") I64("k") I64("a_begin") I64("a_end"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpKsBlocksparseFwdSymPeer, KsFwdSymPeer,
                              CAGP_BIND() ARG ARG ARG ARG RET I64("family") I64("d") I64("k") I64("a_begin") I64("a_end")
                                  .Attr<ffi::Span<const int64_t>>("panels")
                                  .Attr<ffi::Span<const int64_t>>("row_begin") I64("me"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpKsBlocksparseBwdSym, KsBwdSym,
                              CAGP_BIND() ARG ARG ARG ARG ARG RET RET RET I64("family") I64("d") I64("k") I64("a_begin") I64("a_end"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpProjectRows, ProjectRows, CAGP_BIND() ARG ARG ARG RET RET I64("row_offset") I64("n_total"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpGramMtm, GramMtm, CAGP_BIND() ARG RET RET);
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpAssembleCotangent, AssembleCotangent,
                              CAGP_BIND() ARG ARG ARG ARG ARG ARG RET I64("row_offset") I64("n_total"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpAssembleCotangentPeer, AssembleCotangentPeer,
                              CAGP_BIND() ARG ARG ARG ARG ARG ARG RET I64("row_offset") I64("n_total")
                                  .Attr<ffi::Span<const int64_t>>("buffers")
                                  .Attr<ffi::Span<const int64_t>>("block_begin") I64("me"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpProjectRowsBwd, ProjectRowsBwd, CAGP_BIND() ARG ARG ARG RET RET I64("row_offset") I64("n_total"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpPredictMoments, PredictMoments, CAGP_BIND() ARG ARG ARG ARG RET RET RET);
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpKsDenseFwd, KsDenseFwd, CAGP_BIND() ARG ARG ARG ARG RET I64("family") I64("d"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpKsDenseBwdLs, KsDenseBwdLs, CAGP_BIND() ARG ARG ARG ARG ARG RET RET I64("family") I64("d"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpCrossCovFwd, CrossCovFwd, CAGP_BIND() ARG ARG ARG RET I64("family") I64("d"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpCrossCovBwd, CrossCovBwd, CAGP_BIND() ARG ARG ARG ARG RET RET RET I64("family") I64("d"));

#endif  // CAGP_HAVE_XLA_FFI
