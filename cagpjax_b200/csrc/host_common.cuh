// Host-side helpers shared by the translation units of libcagp_b200.so.
#pragma once
#include "../../include/cagp_b200.h"

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <type_traits>

namespace cagp_host {

int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
int num_sms();

template <typename F>
int dispatch_dim(int dp, F&& f) {
    switch (dp) {
        case 1: return f(std::integral_constant<int, 1>{});
        case 2: return f(std::integral_constant<int, 2>{});
        case 3: return f(std::integral_constant<int, 3>{});
        case 4: return f(std::integral_constant<int, 4>{});
        case 8: return f(std::integral_constant<int, 8>{});
        case 16: return f(std::integral_constant<int, 16>{});
        case 32: return f(std::integral_constant<int, 32>{});
    }
    return fail(CAGP_ERR_UNSUPPORTED, "padded dimension %d not instantiated", dp);
}
template <typename F>
int dispatch_family(int fam, F&& f) {
    switch (fam) {
        case CAGP_RBF: return f(std::integral_constant<int, 0>{});
        case CAGP_MATERN12: return f(std::integral_constant<int, 1>{});
        case CAGP_MATERN32: return f(std::integral_constant<int, 2>{});
        case CAGP_MATERN52: return f(std::integral_constant<int, 3>{});
    }
    return fail(CAGP_ERR_INVALID, "unknown family");
}

// rows per thread: coordinates of R points live in registers, so R shrinks with the dimension
template <int D>
struct MaxRows {
    static constexpr int value = D <= 4 ? 4 : (D <= 8 ? 2 : 1);
};
template <int D>
struct ColsPerStage {
    static constexpr int value = D <= 8 ? 256 : 128;
};
inline int max_rows(int dp) { return dp <= 4 ? 4 : (dp <= 8 ? 2 : 1); }
// smallest R in {1, 2, 4} (<= rmax) such that nt * R covers a block of `bs` rows, else rmax
inline int choose_rows(int64_t bs, int nt, int rmax) {
    for (int r = 1; r <= rmax; r *= 2)
        if ((int64_t)nt * r >= bs) return r;
    return rmax;
}
template <int D, typename F>
int dispatch_rows(int r, F&& f) {
    if (r == 1) return f(std::integral_constant<int, 1>{});
    if constexpr (MaxRows<D>::value >= 2) {
        if (r == 2) return f(std::integral_constant<int, 2>{});
    }
    if constexpr (MaxRows<D>::value >= 4) {
        if (r == 4) return f(std::integral_constant<int, 4>{});
    }
    return fail(CAGP_ERR_UNSUPPORTED, "rows per thread %d not instantiated", r);
}

// ---- implemented in impl_*.cu, one object per floating-point type ---------------------------
template <typename T>
int fwd_impl(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
             int64_t ld2, int d, const void* s, int k, void* Mt, int64_t ldm, cudaStream_t stream);
template <typename T>
size_t bwd_ws_bytes(int64_t m, int64_t n, int d, int k, bool ard);
template <typename T>
int bwd_impl(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
             int64_t ld2, int d, const void* s, int k, const void* Gt, int64_t ldg, void* s_bar, void* ls_bar,
             void* ws, size_t ws_bytes, cudaStream_t stream);
template <typename T>
int sym_fwd_impl(const cagp_kernel_desc* kd, const void* xs, int64_t n, int64_t ld, int d, const void* s, int k,
                 int a_begin, int a_end, void* const* panels, const int64_t* row_begin, int world, int me, int64_t ldm,
                 cudaStream_t stream);
template <typename T>
size_t sym_bwd_ws_bytes(int64_t n, int d, int k, bool ard, int a_begin, int a_end);
template <typename T>
int sym_bwd_impl(const cagp_kernel_desc* kd, const void* xs, int64_t n, int64_t ld, int d, const void* s, int k,
                 int a_begin, int a_end, const void* Gt, int64_t ldg, void* s_bar, void* ls_bar, void* ws,
                 size_t ws_bytes, cudaStream_t stream);

template <typename T>
int dense_fwd_impl(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
                   int64_t ld2, int d, const void* V, int p, int64_t ldv, void* Y, int64_t ldy, cudaStream_t stream);
template <typename T>
size_t dense_ls_ws_bytes(int64_t m, int64_t n, int d, bool ard);
template <typename T>
int dense_ls_impl(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
                  int64_t ld2, int d, const void* V, int64_t ldv, const void* Ybar, int64_t ldyb, int p, void* ls_bar,
                  void* ws, size_t ws_bytes, cudaStream_t stream);

template <typename T>
int cross_fwd_impl(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
                   int64_t ld2, int d, void* C, int64_t ldc, cudaStream_t stream);
template <typename T>
size_t cross_bwd_ws_bytes(int64_t m, int64_t n, int d, bool ard);
template <typename T>
int cross_bwd_impl(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
                   int64_t ld2, int d, const void* Cbar, int64_t ldc, void* x2s_bar, void* ls_bar, void* ws,
                   size_t ws_bytes, cudaStream_t stream);

}  // namespace cagp_host
