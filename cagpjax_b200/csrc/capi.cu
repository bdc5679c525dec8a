// C ABI of libcagp_b200.so: argument checking, kernel selection and launches.  See
// include/cagp_b200.h for the contract of each entry point.
#include "../../include/cagp_b200.h"

#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "aux_kernels.cuh"
#include "host_common.cuh"
#include "sym_tiling.cuh"

namespace cagp_host {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CAGP_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return CAGP_OK;
}

int num_sms() {
    static int sms[64] = {0};  // per device (a process may drive several)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!sms[dev]) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = v > 0 ? v : 148;
    }
    return sms[dev];
}

}  // namespace cagp_host

namespace cagp_host {
// impl_gemm.cu: float64 contractions on the fp64 tensor core (gemm_dmma.cuh)
size_t gram_ws_bytes_f64(int64_t m, int k);
int gram_mtm_f64(const double* Mt, int64_t ldm, int64_t m, int k, double* B, void* ws, size_t ws_bytes, cudaStream_t stream);
int assemble_cotangent_f64(const double* Mt, int64_t ldm, int64_t m, int64_t row_offset, int64_t n_total, int k,
                           const double* W, const double* Abar, const double* cbar, const double* s, const double* yt,
                           double* Gt, int64_t ldg, cudaStream_t stream);
int assemble_cotangent_peer_f64(const double* Mt, int64_t ldm, int64_t m, int64_t row_offset, int64_t n_total, int k,
                                const double* W, const double* Abar, const double* cbar, const double* s, const double* yt,
                                void* const* buffers, const int32_t* block_begin, int world, int me, int64_t ld,
                                cudaStream_t stream);
size_t predict_ws_bytes_f64(int64_t m, int k);
int predict_moments_f64(const double* Mt, int64_t ldm, int64_t m, int k, const double* w, const double* Pinv,
                        const double* scalars, double* mean, double* var, void* ws, size_t ws_bytes, cudaStream_t stream);
}  // namespace cagp_host

using namespace cagp_host;

namespace {

// one cuBLAS handle per device, created lazily (the only global state of the library)
std::mutex g_handle_mu;
cublasHandle_t g_handles[64] = {nullptr};

int get_handle(cudaStream_t stream, cublasHandle_t* out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return fail(CAGP_ERR_CUDA, "cudaGetDevice failed");
    // the caller holds g_handle_mu until its gemm is enqueued (SetStream + gemm must not interleave between threads)
    if (!g_handles[dev]) {
        if (cublasCreate(&g_handles[dev]) != CUBLAS_STATUS_SUCCESS) return fail(CAGP_ERR_CUBLAS, "cublasCreate failed");
        cublasSetPointerMode(g_handles[dev], CUBLAS_POINTER_MODE_HOST);
        cublasSetMathMode(g_handles[dev], CUBLAS_DEFAULT_MATH);  // fp32 stays fp32 (TF32 needs an explicit opt-in)
    }
    if (cublasSetStream(g_handles[dev], stream) != CUBLAS_STATUS_SUCCESS) return fail(CAGP_ERR_CUBLAS, "cublasSetStream failed");
    *out = g_handles[dev];
    return CAGP_OK;
}

int check_desc(const cagp_kernel_desc* kd, int d) {
    if (!kd) return fail(CAGP_ERR_INVALID, "kernel descriptor is null");
    if (kd->family < 0 || kd->family > 3) return fail(CAGP_ERR_INVALID, "unknown kernel family %d", kd->family);
    if (kd->dtype != CAGP_F32 && kd->dtype != CAGP_F64) return fail(CAGP_ERR_INVALID, "unknown dtype %d", kd->dtype);
    if (d < 1) return fail(CAGP_ERR_INVALID, "input dimension must be >= 1");
    if (kd->n_lengthscale != 1 && kd->n_lengthscale != d)
        return fail(CAGP_ERR_INVALID, "n_lengthscale must be 1 or d (%d), got %d", d, kd->n_lengthscale);
    if (!kd->lengthscale) return fail(CAGP_ERR_INVALID, "lengthscale pointer is null");
    if (cagp_padded_dim(d) < 0) return fail(CAGP_ERR_UNSUPPORTED, "input dimension %d > 32 is not supported", d);
    return CAGP_OK;
}

}  // namespace

extern "C" {

const char* cagp_last_error_string(void) { return cagp_host::g_err; }
int cagp_version(void) { return 100; }

int cagp_padded_dim(int32_t d) {
    static const int dims[] = {1, 2, 3, 4, 8, 16, 32};
    for (int v : dims)
        if (d <= v) return v;
    return -1;
}

int cagp_scale_inputs(const cagp_kernel_desc* kd, const void* X, int64_t n, int32_t d, const void* origin, void* xs,
                      int64_t ld, cagp_stream_t stream) {
    if (int rc = check_desc(kd, d)) return rc;
    if (!X || !xs || n < 1 || ld < n) return fail(CAGP_ERR_INVALID, "scale_inputs: bad X/xs/n/ld");
    const int dp = cagp_padded_dim(d);
    const double c = cagp::family_scale(kd->family);
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)dp);
    if (kd->dtype == CAGP_F64)
        cagp::scale_inputs_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)X, n, d, (const double*)kd->lengthscale, kd->n_lengthscale, c, (const double*)origin, (double*)xs, ld);
    else
        cagp::scale_inputs_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)X, n, d, (const float*)kd->lengthscale, kd->n_lengthscale, (float)c, (const float*)origin, (float*)xs, ld);
    return check_launch("scale_inputs_kernel");
}

int cagp_bounding_box(int32_t dtype, const void* xs, int64_t n, int64_t ld, int32_t dpad, void* bbox, cagp_stream_t stream) {
    if (!xs || !bbox || n < 1 || ld < n || dpad < 1) return fail(CAGP_ERR_INVALID, "bounding_box: bad argument");
    if (dtype == CAGP_F64)
        cagp::bounding_box_kernel<double><<<dpad, 1024, 0, (cudaStream_t)stream>>>((const double*)xs, n, ld, dpad, (double*)bbox);
    else if (dtype == CAGP_F32)
        cagp::bounding_box_kernel<float><<<dpad, 1024, 0, (cudaStream_t)stream>>>((const float*)xs, n, ld, dpad, (float*)bbox);
    else
        return fail(CAGP_ERR_INVALID, "unknown dtype");
    return check_launch("bounding_box_kernel");
}

int cagp_ks_blocksparse_fwd(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2,
                            int64_t n, int64_t ld2, int32_t d, const void* s, int32_t k, void* Mt, int64_t ldm,
                            cagp_stream_t stream) {
    if (int rc = check_desc(kd, d)) return rc;
    if (!xs1 || !xs2 || !s || !Mt) return fail(CAGP_ERR_INVALID, "ks fwd: null pointer");
    if (m < 1 || n < 1 || k < 1 || ld1 < m || ld2 < n || ldm < m) return fail(CAGP_ERR_INVALID, "ks fwd: bad shape");
    if (kd->dtype == CAGP_F64) return fwd_impl<double>(kd, xs1, m, ld1, xs2, n, ld2, d, s, k, Mt, ldm, (cudaStream_t)stream);
    return fwd_impl<float>(kd, xs1, m, ld1, xs2, n, ld2, d, s, k, Mt, ldm, (cudaStream_t)stream);
}

int cagp_ks_blocksparse_fwd_sym(const cagp_kernel_desc* kd, const void* xs, int64_t n, int64_t ld, int32_t d,
                                const void* s, int32_t k, int32_t a_begin, int32_t a_end, void* Mt, int64_t ldm,
                                cagp_stream_t stream) {
    if (int rc = check_desc(kd, d)) return rc;
    if (!xs || !s || !Mt) return fail(CAGP_ERR_INVALID, "ks sym fwd: null pointer");
    if (n < 1 || k < 1 || ld < n || ldm < n || a_begin < 0 || a_end > k || a_begin > a_end) return fail(CAGP_ERR_INVALID, "ks sym fwd: bad shape");
    void* panels[1] = {Mt};
    const int64_t row_begin[2] = {0, n};
    if (kd->dtype == CAGP_F64) return sym_fwd_impl<double>(kd, xs, n, ld, d, s, k, a_begin, a_end, panels, row_begin, 1, 0, ldm, (cudaStream_t)stream);
    return sym_fwd_impl<float>(kd, xs, n, ld, d, s, k, a_begin, a_end, panels, row_begin, 1, 0, ldm, (cudaStream_t)stream);
}

int cagp_ks_blocksparse_fwd_sym_peer(const cagp_kernel_desc* kd, const void* xs, int64_t n, int64_t ld, int32_t d,
                                     const void* s, int32_t k, int32_t a_begin, int32_t a_end, void* const* panels,
                                     const int64_t* row_begin, int32_t world, int32_t me, int64_t ldm,
                                     cagp_stream_t stream) {
    if (int rc = check_desc(kd, d)) return rc;
    if (!xs || !s || !panels || !row_begin) return fail(CAGP_ERR_INVALID, "ks sym fwd peer: null pointer");
    if (world < 1 || world > 16 || me < 0 || me >= world) return fail(CAGP_ERR_INVALID, "ks sym fwd peer: bad world/rank");
    if (n < 1 || k < 1 || ld < n || a_begin < 0 || a_end > k || a_begin > a_end) return fail(CAGP_ERR_INVALID, "ks sym fwd peer: bad shape");
    if (row_begin[0] != 0 || row_begin[world] != n) return fail(CAGP_ERR_INVALID, "ks sym fwd peer: row_begin must span [0, n]");
    const int64_t bs = n / k;
    for (int r = 0; r < world; ++r) {
        if (!panels[r]) return fail(CAGP_ERR_INVALID, "ks sym fwd peer: null panel");
        if (row_begin[r + 1] < row_begin[r] || row_begin[r + 1] - row_begin[r] > ldm) return fail(CAGP_ERR_INVALID, "ks sym fwd peer: bad row partition");
        if (r + 1 < world && bs > 0 && row_begin[r + 1] % bs != 0) return fail(CAGP_ERR_INVALID, "ks sym fwd peer: rows must be split on action-block boundaries");
    }
    if (kd->dtype == CAGP_F64) return sym_fwd_impl<double>(kd, xs, n, ld, d, s, k, a_begin, a_end, panels, row_begin, world, me, ldm, (cudaStream_t)stream);
    return sym_fwd_impl<float>(kd, xs, n, ld, d, s, k, a_begin, a_end, panels, row_begin, world, me, ldm, (cudaStream_t)stream);
}

int cagp_ks_sym_fast_path(const cagp_kernel_desc* kd, int64_t n, int32_t d, int32_t k, int32_t backward) {
    if (check_desc(kd, d) || n < 1 || k < 1) return 0;
    return symx_applicable(kd, n, k, cagp_padded_dim(d), backward != 0) ? 1 : 0;
}

int cagp_ks_fwd_sym_zero_rows(const cagp_kernel_desc* kd, int64_t n, int32_t d, int32_t k) {
    if (check_desc(kd, d) || n < 1 || k < 1) return 2;
    const int dp = cagp_padded_dim(d);
    const SymShape sh = sym_shape(n, k, dp, false);
    cagp::SymTiling t = make_sym_tiling(n, k, 0, k, sh);
    int tiles_main = t.tiles_main, tiles_last = t.tiles_last;
    if (symx_applicable(kd, n, k, dp, false)) {
        const cagp::SymTiling tf = make_sym_tiling(n, k, 0, k, symx_shape(false));
        tiles_main = tf.tiles_main > tiles_main ? tf.tiles_main : tiles_main;
        tiles_last = tf.tiles_last > tiles_last ? tf.tiles_last : tiles_last;
    }
    return tiles_main > 1 ? 2 : (tiles_last > 1 ? 1 : 0);
}

size_t cagp_ks_bwd_sym_ws_bytes(const cagp_kernel_desc* kd, int64_t n, int32_t d, int32_t k, int32_t a_begin, int32_t a_end) {
    if (check_desc(kd, d) || n < 1 || k < 1) return 0;
    const bool ard = kd->n_lengthscale > 1;
    return kd->dtype == CAGP_F64 ? sym_bwd_ws_bytes<double>(n, d, k, ard, a_begin, a_end) : sym_bwd_ws_bytes<float>(n, d, k, ard, a_begin, a_end);
}

int cagp_ks_blocksparse_bwd_sym(const cagp_kernel_desc* kd, const void* xs, int64_t n, int64_t ld, int32_t d,
                                const void* s, int32_t k, int32_t a_begin, int32_t a_end, const void* Gt, int64_t ldg,
                                void* s_bar, void* ls_bar, void* ws, size_t ws_bytes, cagp_stream_t stream) {
    if (int rc = check_desc(kd, d)) return rc;
    if (!xs || !s || !Gt || !s_bar || !ls_bar || !ws) return fail(CAGP_ERR_INVALID, "ks sym bwd: null pointer");
    if (n < 1 || k < 1 || ld < n || ldg < n || a_begin < 0 || a_end > k || a_begin > a_end) return fail(CAGP_ERR_INVALID, "ks sym bwd: bad shape");
    if (kd->dtype == CAGP_F64)
        return sym_bwd_impl<double>(kd, xs, n, ld, d, s, k, a_begin, a_end, Gt, ldg, s_bar, ls_bar, ws, ws_bytes, (cudaStream_t)stream);
    return sym_bwd_impl<float>(kd, xs, n, ld, d, s, k, a_begin, a_end, Gt, ldg, s_bar, ls_bar, ws, ws_bytes, (cudaStream_t)stream);
}

size_t cagp_ks_bwd_ws_bytes(const cagp_kernel_desc* kd, int64_t m, int64_t n, int32_t d, int32_t k) {
    if (check_desc(kd, d) || m < 1 || n < 1 || k < 1) return 0;
    const bool ard = kd->n_lengthscale > 1;
    return kd->dtype == CAGP_F64 ? bwd_ws_bytes<double>(m, n, d, k, ard) : bwd_ws_bytes<float>(m, n, d, k, ard);
}

int cagp_ks_blocksparse_bwd(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2,
                            int64_t n, int64_t ld2, int32_t d, const void* s, int32_t k, const void* Gt, int64_t ldg,
                            void* s_bar, void* ls_bar, void* ws, size_t ws_bytes, cagp_stream_t stream) {
    if (int rc = check_desc(kd, d)) return rc;
    if (!xs1 || !xs2 || !s || !Gt || !s_bar || !ls_bar || !ws) return fail(CAGP_ERR_INVALID, "ks bwd: null pointer");
    if (m < 1 || n < 1 || k < 1 || ld1 < m || ld2 < n || ldg < m) return fail(CAGP_ERR_INVALID, "ks bwd: bad shape");
    if (kd->dtype == CAGP_F64)
        return bwd_impl<double>(kd, xs1, m, ld1, xs2, n, ld2, d, s, k, Gt, ldg, s_bar, ls_bar, ws, ws_bytes, (cudaStream_t)stream);
    return bwd_impl<float>(kd, xs1, m, ld1, xs2, n, ld2, d, s, k, Gt, ldg, s_bar, ls_bar, ws, ws_bytes, (cudaStream_t)stream);
}

int cagp_project_rows(int32_t dtype, const void* Mt, int64_t ldm, int64_t m, int64_t row_offset, int64_t n_total,
                      int32_t k, const void* s, const void* yt, void* A, void* c, cagp_stream_t stream) {
    if (!Mt || !s || !yt || !A || !c) return fail(CAGP_ERR_INVALID, "project_rows: null pointer");
    if (m < 1 || k < 1 || ldm < m || row_offset < 0 || row_offset + m > n_total) return fail(CAGP_ERR_INVALID, "project_rows: bad shape");
    cagp::BlockLayout lay{n_total, k, n_total / k};
    if (dtype == CAGP_F64)
        cagp::project_rows_kernel<double><<<k, 256, 0, (cudaStream_t)stream>>>((const double*)Mt, ldm, m, row_offset, lay, (const double*)s, (const double*)yt, (double*)A, (double*)c);
    else if (dtype == CAGP_F32)
        cagp::project_rows_kernel<float><<<k, 256, 0, (cudaStream_t)stream>>>((const float*)Mt, ldm, m, row_offset, lay, (const float*)s, (const float*)yt, (float*)A, (float*)c);
    else
        return fail(CAGP_ERR_INVALID, "unknown dtype");
    return check_launch("project_rows_kernel");
}

size_t cagp_gram_ws_bytes(int32_t dtype, int64_t m, int32_t k) {
    if (dtype != CAGP_F64 || m < 1 || k < 1) return 0;  // float32 goes through cuBLAS Sgemm, no scratch
    return gram_ws_bytes_f64(m, k);
}

int cagp_gram_mtm(int32_t dtype, const void* Mt, int64_t ldm, int64_t m, int32_t k, void* B, void* ws, size_t ws_bytes,
                  cagp_stream_t stream) {
    if (!Mt || !B || m < 1 || k < 1 || ldm < m) return fail(CAGP_ERR_INVALID, "gram_mtm: bad argument");
    if (dtype == CAGP_F64)  // hand-written DMMA kernel: lower-triangular tiles only, deterministic split-m reduction
        return gram_mtm_f64((const double*)Mt, ldm, m, k, (double*)B, ws, ws_bytes, (cudaStream_t)stream);
    if (dtype != CAGP_F32) return fail(CAGP_ERR_INVALID, "unknown dtype");
    if (m > 0x7fffffffLL || ldm > 0x7fffffffLL) return fail(CAGP_ERR_UNSUPPORTED, "gram_mtm: m exceeds int32 (shard rows)");
    cublasHandle_t h;
    std::unique_lock<std::mutex> lock(g_handle_mu);  // held across SetStream + gemm: one handle per device
    if (int rc = get_handle((cudaStream_t)stream, &h)) return rc;
    // Mt row-major (k, ldm) == column-major (ldm, k) matrix M with m valid rows: B = M^T M (float32 only)
    const float one = 1.f, zero = 0.f;
    cublasStatus_t st = cublasSgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, k, k, (int)m, &one, (const float*)Mt, (int)ldm, (const float*)Mt, (int)ldm, &zero, (float*)B, k);
    lock.unlock();
    if (st != CUBLAS_STATUS_SUCCESS) return fail(CAGP_ERR_CUBLAS, "cublas gemm (gram) failed with status %d", (int)st);
    cagp::mirror_lower_kernel<float><<<dim3((k + 15) / 16, (k + 15) / 16), dim3(16, 16), 0, (cudaStream_t)stream>>>((float*)B, k);
    return check_launch("mirror_lower_kernel");
}

int cagp_assemble_cotangent(int32_t dtype, const void* Mt, int64_t ldm, int64_t m, int64_t row_offset, int64_t n_total,
                            int32_t k, const void* W, const void* Abar, const void* cbar, const void* s, const void* yt,
                            void* Gt, int64_t ldg, cagp_stream_t stream) {
    if (!Mt || !W || !Abar || !cbar || !s || !yt || !Gt) return fail(CAGP_ERR_INVALID, "assemble_cotangent: null pointer");
    if (m < 1 || k < 1 || ldm < m || ldg < m || row_offset < 0 || row_offset + m > n_total) return fail(CAGP_ERR_INVALID, "assemble_cotangent: bad shape");
    if (dtype == CAGP_F64)  // one DMMA kernel: W Mt with the outer terms as its epilogue
        return assemble_cotangent_f64((const double*)Mt, ldm, m, row_offset, n_total, k, (const double*)W, (const double*)Abar,
                                      (const double*)cbar, (const double*)s, (const double*)yt, (double*)Gt, ldg, (cudaStream_t)stream);
    if (dtype != CAGP_F32) return fail(CAGP_ERR_INVALID, "unknown dtype");
    if (m > 0x7fffffffLL || ldm > 0x7fffffffLL || ldg > 0x7fffffffLL) return fail(CAGP_ERR_UNSUPPORTED, "assemble_cotangent: m exceeds int32");
    cagp::BlockLayout lay{n_total, k, n_total / k};
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)k);
    // float32: column-major G (m x k, ld ldg) = M (m x k, ld ldm) * W^T (k x k) + prefilled outer terms (cuBLAS Sgemm)
    cagp::cotangent_prefill_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(m, row_offset, lay, (const float*)Abar, (const float*)cbar, (const float*)s, (const float*)yt, (float*)Gt, ldg);
    if (int rc = check_launch("cotangent_prefill_kernel")) return rc;
    cublasHandle_t h;
    std::unique_lock<std::mutex> lock(g_handle_mu);
    if (int rc = get_handle((cudaStream_t)stream, &h)) return rc;
    const float one = 1.f;
    cublasStatus_t st = cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, (int)m, k, k, &one, (const float*)Mt, (int)ldm, (const float*)W, k, &one, (float*)Gt, (int)ldg);
    lock.unlock();
    if (st != CUBLAS_STATUS_SUCCESS) return fail(CAGP_ERR_CUBLAS, "cublas gemm failed with status %d", (int)st);
    return CAGP_OK;
}

int cagp_assemble_cotangent_peer(int32_t dtype, const void* Mt, int64_t ldm, int64_t m, int64_t row_offset,
                                 int64_t n_total, int32_t k, const void* W, const void* Abar, const void* cbar,
                                 const void* s, const void* yt, void* const* buffers, const int32_t* block_begin,
                                 int32_t world, int32_t me, int64_t ld, cagp_stream_t stream) {
    if (!Mt || !W || !Abar || !cbar || !s || !yt || !buffers || !block_begin) return fail(CAGP_ERR_INVALID, "assemble_cotangent_peer: null pointer");
    if (dtype != CAGP_F64) return fail(CAGP_ERR_UNSUPPORTED, "assemble_cotangent_peer: float64 only");
    if (world < 1 || world > 16 || me < 0 || me >= world) return fail(CAGP_ERR_INVALID, "assemble_cotangent_peer: bad world/rank");
    if (m < 1 || k < 1 || ldm < m || ld < n_total || row_offset < 0 || row_offset + m > n_total) return fail(CAGP_ERR_INVALID, "assemble_cotangent_peer: bad shape");
    if (block_begin[0] != 0 || block_begin[world] != k) return fail(CAGP_ERR_INVALID, "assemble_cotangent_peer: block_begin must span [0, k]");
    for (int h = 0; h < world; ++h) {
        if (!buffers[h]) return fail(CAGP_ERR_INVALID, "assemble_cotangent_peer: null buffer");
        if (block_begin[h + 1] < block_begin[h]) return fail(CAGP_ERR_INVALID, "assemble_cotangent_peer: bad block partition");
    }
    return assemble_cotangent_peer_f64((const double*)Mt, ldm, m, row_offset, n_total, k, (const double*)W, (const double*)Abar,
                                       (const double*)cbar, (const double*)s, (const double*)yt, buffers, block_begin, world, me,
                                       ld, (cudaStream_t)stream);
}

int cagp_project_rows_bwd(int32_t dtype, const void* Mt, int64_t ldm, int64_t m, int64_t row_offset, int64_t n_total,
                          int32_t k, const void* Abar, const void* cbar, void* s_dir, void* yt_bar, cagp_stream_t stream) {
    if (!Mt || !Abar || !cbar || !s_dir || !yt_bar) return fail(CAGP_ERR_INVALID, "project_rows_bwd: null pointer");
    if (m < 1 || k < 1 || ldm < m || row_offset < 0 || row_offset + m > n_total) return fail(CAGP_ERR_INVALID, "project_rows_bwd: bad shape");
    cagp::BlockLayout lay{n_total, k, n_total / k};
    const unsigned grid = (unsigned)((m + 255) / 256);
    if (dtype == CAGP_F64)
        cagp::project_rows_bwd_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)Mt, ldm, m, row_offset, lay, (const double*)Abar, (const double*)cbar, (double*)s_dir, (double*)yt_bar);
    else if (dtype == CAGP_F32)
        cagp::project_rows_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)Mt, ldm, m, row_offset, lay, (const float*)Abar, (const float*)cbar, (float*)s_dir, (float*)yt_bar);
    else
        return fail(CAGP_ERR_INVALID, "unknown dtype");
    return check_launch("project_rows_bwd_kernel");
}

int cagp_ks_dense_fwd(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
                      int64_t ld2, int32_t d, const void* V, int32_t p, int64_t ldv, void* Y, int64_t ldy,
                      cagp_stream_t stream) {
    if (int rc = check_desc(kd, d)) return rc;
    if (!xs1 || !xs2 || !V || !Y) return fail(CAGP_ERR_INVALID, "ks dense fwd: null pointer");
    if (m < 1 || n < 1 || p < 1 || ld1 < m || ld2 < n || ldv < p || ldy < p) return fail(CAGP_ERR_INVALID, "ks dense fwd: bad shape");
    if (kd->dtype == CAGP_F64) return dense_fwd_impl<double>(kd, xs1, m, ld1, xs2, n, ld2, d, V, p, ldv, Y, ldy, (cudaStream_t)stream);
    return dense_fwd_impl<float>(kd, xs1, m, ld1, xs2, n, ld2, d, V, p, ldv, Y, ldy, (cudaStream_t)stream);
}

size_t cagp_ks_dense_bwd_ls_ws_bytes(const cagp_kernel_desc* kd, int64_t m, int64_t n, int32_t d) {
    if (check_desc(kd, d) || m < 1 || n < 1) return 0;
    const bool ard = kd->n_lengthscale > 1;
    return kd->dtype == CAGP_F64 ? dense_ls_ws_bytes<double>(m, n, d, ard) : dense_ls_ws_bytes<float>(m, n, d, ard);
}

int cagp_ks_dense_bwd_ls(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
                         int64_t ld2, int32_t d, const void* V, int64_t ldv, const void* Ybar, int64_t ldyb, int32_t p,
                         void* ls_bar, void* ws, size_t ws_bytes, cagp_stream_t stream) {
    if (int rc = check_desc(kd, d)) return rc;
    if (!xs1 || !xs2 || !V || !Ybar || !ls_bar || !ws) return fail(CAGP_ERR_INVALID, "ks dense bwd: null pointer");
    if (m < 1 || n < 1 || p < 1 || ld1 < m || ld2 < n || ldv < n || ldyb < m) return fail(CAGP_ERR_INVALID, "ks dense bwd: bad shape");
    if (kd->dtype == CAGP_F64)
        return dense_ls_impl<double>(kd, xs1, m, ld1, xs2, n, ld2, d, V, ldv, Ybar, ldyb, p, ls_bar, ws, ws_bytes, (cudaStream_t)stream);
    return dense_ls_impl<float>(kd, xs1, m, ld1, xs2, n, ld2, d, V, ldv, Ybar, ldyb, p, ls_bar, ws, ws_bytes, (cudaStream_t)stream);
}

int cagp_cross_cov_fwd(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
                       int64_t ld2, int32_t d, void* C, int64_t ldc, cagp_stream_t stream) {
    if (int rc = check_desc(kd, d)) return rc;
    if (!xs1 || !xs2 || !C) return fail(CAGP_ERR_INVALID, "cross_cov fwd: null pointer");
    if (m < 1 || n < 1 || ld1 < m || ld2 < n || ldc < n) return fail(CAGP_ERR_INVALID, "cross_cov fwd: bad shape");
    if (kd->dtype == CAGP_F64) return cross_fwd_impl<double>(kd, xs1, m, ld1, xs2, n, ld2, d, C, ldc, (cudaStream_t)stream);
    return cross_fwd_impl<float>(kd, xs1, m, ld1, xs2, n, ld2, d, C, ldc, (cudaStream_t)stream);
}

size_t cagp_cross_cov_bwd_ws_bytes(const cagp_kernel_desc* kd, int64_t m, int64_t n, int32_t d) {
    if (check_desc(kd, d) || m < 1 || n < 1) return 0;
    const bool ard = kd->n_lengthscale > 1;
    return kd->dtype == CAGP_F64 ? cross_bwd_ws_bytes<double>(m, n, d, ard) : cross_bwd_ws_bytes<float>(m, n, d, ard);
}

int cagp_cross_cov_bwd(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
                       int64_t ld2, int32_t d, const void* Cbar, int64_t ldc, void* x2s_bar, void* ls_bar, void* ws,
                       size_t ws_bytes, cagp_stream_t stream) {
    if (int rc = check_desc(kd, d)) return rc;
    if (!xs1 || !xs2 || !Cbar || !x2s_bar || !ls_bar || !ws) return fail(CAGP_ERR_INVALID, "cross_cov bwd: null pointer");
    if (m < 1 || n < 1 || ld1 < m || ld2 < n || ldc < n) return fail(CAGP_ERR_INVALID, "cross_cov bwd: bad shape");
    if (kd->dtype == CAGP_F64)
        return cross_bwd_impl<double>(kd, xs1, m, ld1, xs2, n, ld2, d, Cbar, ldc, x2s_bar, ls_bar, ws, ws_bytes, (cudaStream_t)stream);
    return cross_bwd_impl<float>(kd, xs1, m, ld1, xs2, n, ld2, d, Cbar, ldc, x2s_bar, ls_bar, ws, ws_bytes, (cudaStream_t)stream);
}

size_t cagp_predict_ws_bytes(int32_t dtype, int64_t m, int32_t k) {
    if (dtype == CAGP_F64) return predict_ws_bytes_f64(m, k);
    return (size_t)k * (size_t)m * 4;
}

int cagp_predict_moments(int32_t dtype, const void* Mt, int64_t ldm, int64_t m, int32_t k, const void* w,
                         const void* Pinv, const void* scalars, void* mean, void* var, void* ws, size_t ws_bytes,
                         cagp_stream_t stream) {
    if (!Mt || !w || !Pinv || !scalars || !mean || !var || !ws) return fail(CAGP_ERR_INVALID, "predict_moments: null pointer");
    if (m < 1 || k < 1 || ldm < m) return fail(CAGP_ERR_INVALID, "predict_moments: bad shape");
    if (ws_bytes < cagp_predict_ws_bytes(dtype, m, k)) return fail(CAGP_ERR_WORKSPACE, "predict_moments: workspace too small");
    if (dtype == CAGP_F64)  // DMMA kernel; Z = Pinv Mt is contracted in its epilogue and never written
        return predict_moments_f64((const double*)Mt, ldm, m, k, (const double*)w, (const double*)Pinv, (const double*)scalars,
                                   (double*)mean, (double*)var, ws, ws_bytes, (cudaStream_t)stream);
    if (dtype != CAGP_F32) return fail(CAGP_ERR_INVALID, "unknown dtype");
    if (m > 0x7fffffffLL || ldm > 0x7fffffffLL) return fail(CAGP_ERR_UNSUPPORTED, "predict_moments: m exceeds int32");
    cublasHandle_t h;
    std::unique_lock<std::mutex> lock(g_handle_mu);
    if (int rc = get_handle((cudaStream_t)stream, &h)) return rc;
    // float32: column-major Z (m x k, ld m) = M (m x k) * Pinv (k x k, symmetric) via cuBLAS Sgemm
    const float one = 1.f, zero = 0.f;
    cublasStatus_t st = cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, (int)m, k, k, &one, (const float*)Mt, (int)ldm, (const float*)Pinv, k, &zero, (float*)ws, (int)m);
    lock.unlock();
    if (st != CUBLAS_STATUS_SUCCESS) return fail(CAGP_ERR_CUBLAS, "cublas gemm failed with status %d", (int)st);
    const unsigned grid = (unsigned)((m + 255) / 256);
    cagp::predict_moments_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)Mt, ldm, (const float*)ws, m, m, k, (const float*)w, (const float*)scalars, (float*)mean, (float*)var);
    return check_launch("predict_moments_kernel");
}

}  // extern "C"
