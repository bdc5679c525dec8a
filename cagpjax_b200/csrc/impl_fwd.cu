// Launcher of the rectangular forward kernel (one object per floating-point type: -DCAGP_REAL=...).
#include "host_common.cuh"
#include "ks_blocksparse.cuh"

namespace cagp_host {

template <typename T>
int fwd_impl(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
             int64_t ld2, int d, const void* s, int k, void* Mt, int64_t ldm, cudaStream_t stream) {
    const int dp = cagp_padded_dim(d);
    return dispatch_dim(dp, [&](auto dc) {
        return dispatch_family(kd->family, [&](auto fc) {
            constexpr int D = decltype(dc)::value;
            constexpr int FAM = decltype(fc)::value;
            // float64, D <= 4: 8 rows per thread (the per-column loads and loop control are shared by 8 pairs)
            constexpr int R = (sizeof(T) == 8 && D <= 4) ? 8 : MaxRows<D>::value;
            constexpr int NT = 128;
            constexpr int JC = ColsPerStage<D>::value;
            cagp::FwdArgs<T> a;
            a.xs1 = (const T*)xs1; a.ld1 = ld1; a.m = m;
            a.xs2 = (const T*)xs2; a.ld2 = ld2;
            a.s = (const T*)s;
            a.lay = cagp::BlockLayout{n, k, n / k};
            a.Mt = (T*)Mt; a.ldm = ldm;
            a.bbox = (const T*)kd->bbox;
            const int64_t gx = (m + NT * R - 1) / (NT * R);
            const int64_t target = (int64_t)num_sms() * 4 * 32;
            int64_t gy = (target + gx - 1) / gx;
            if (gy > k) gy = k;
            if (gy < 1) gy = 1;
            if (gy > 65535) gy = 65535;
            a.blocks_per_cta = (int)((k + gy - 1) / gy);
            gy = (k + a.blocks_per_cta - 1) / a.blocks_per_cta;
            const size_t smem = (cagp::FwdCanExpand<T, FAM>::kTabElems + (size_t)JC * cagp::EntrySize<D>::value) * sizeof(T);
            auto kern = cagp::ks_fwd_kernel<T, D, FAM, R, NT, JC>;
            if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<dim3((unsigned)gx, (unsigned)gy), NT, smem, stream>>>(a);
            return check_launch("ks_fwd_kernel");
        });
    });
}

template int fwd_impl<CAGP_REAL>(const cagp_kernel_desc*, const void*, int64_t, int64_t, const void*, int64_t, int64_t,
                                 int, const void*, int, void*, int64_t, cudaStream_t);

}  // namespace cagp_host
