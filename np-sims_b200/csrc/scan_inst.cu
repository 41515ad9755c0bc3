// scan_inst.cu — explicit instantiations of the width-specialised scan kernel.
// Compiled four times with -DNPS_WGROUP=0..3 (8 widths each) so the builds run in parallel,
// plus once with -DNPS_WGROUP=4 for the generic kernel and the dispatch table.
#if NPS_WGROUP == 4
#define NPS_SCAN_GENERIC 1
#endif
#include "scan.cuh"
#define NPS_REF_DEVICE_CODE 1
#include "ref_scan.cuh"
#define NPS_HIST_DEVICE_CODE 1
#include "hist_select.cuh"

namespace nps {

template <int W>
static cudaError_t launch_scan_w(const ScanParams& p, int grid, size_t smem, cudaStream_t stream) {
    constexpr int R = rows_per_thread(W);
    auto kern = scan_topk_kernel<W, R>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kScanThreads, smem, stream>>>(p);
    count_launch();
    return cudaGetLastError();
}

template <int W>
static cudaError_t launch_ref_scan_w(const RefParams& p, dim3 grid, size_t smem, cudaStream_t stream) {
    constexpr int R = rows_per_thread(W);
    auto kern = ref_scan_kernel<W, R>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kScanThreads, smem, stream>>>(p);
    count_launch();
    return cudaGetLastError();
}

template <int W>
static cudaError_t launch_hist_scan_w(const HistParams& p, int grid, size_t smem, cudaStream_t stream) {
    constexpr int R = rows_per_thread(W);
    auto kern = hist_scan_kernel<W, R>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kScanThreads, smem, stream>>>(p);
    count_launch();
    return cudaGetLastError();
}

#define NPS_DECL_GROUP(g) ScanLaunchFn scan_launcher_group##g(int words); RefLaunchFn ref_launcher_group##g(int words); HistLaunchFn hist_launcher_group##g(int words);
NPS_DECL_GROUP(0) NPS_DECL_GROUP(1) NPS_DECL_GROUP(2) NPS_DECL_GROUP(3)

#if NPS_WGROUP < 4
#define NPS_CASE(w) case (NPS_WGROUP * 8 + w): return launch_scan_w<NPS_WGROUP * 8 + w>;
#define NPS_DEFINE_GROUP_(g)                                               \
    ScanLaunchFn scan_launcher_group##g(int words) {                       \
        switch (words) {                                                   \
            NPS_CASE(1) NPS_CASE(2) NPS_CASE(3) NPS_CASE(4)                \
            NPS_CASE(5) NPS_CASE(6) NPS_CASE(7) NPS_CASE(8)                \
            default: return nullptr;                                       \
        }                                                                  \
    }
#define NPS_RCASE(w) case (NPS_WGROUP * 8 + w): return launch_ref_scan_w<NPS_WGROUP * 8 + w>;
#define NPS_DEFINE_RGROUP_(g)                                              \
    RefLaunchFn ref_launcher_group##g(int words) {                         \
        switch (words) {                                                   \
            NPS_RCASE(1) NPS_RCASE(2) NPS_RCASE(3) NPS_RCASE(4)            \
            NPS_RCASE(5) NPS_RCASE(6) NPS_RCASE(7) NPS_RCASE(8)            \
            default: return nullptr;                                       \
        }                                                                  \
    }
#define NPS_HCASE(w) case (NPS_WGROUP * 8 + w): return launch_hist_scan_w<NPS_WGROUP * 8 + w>;
#define NPS_DEFINE_HGROUP_(g)                                              \
    HistLaunchFn hist_launcher_group##g(int words) {                       \
        switch (words) {                                                   \
            NPS_HCASE(1) NPS_HCASE(2) NPS_HCASE(3) NPS_HCASE(4)            \
            NPS_HCASE(5) NPS_HCASE(6) NPS_HCASE(7) NPS_HCASE(8)            \
            default: return nullptr;                                       \
        }                                                                  \
    }
#define NPS_DEFINE_GROUP(g) NPS_DEFINE_GROUP_(g) NPS_DEFINE_RGROUP_(g) NPS_DEFINE_HGROUP_(g)
NPS_DEFINE_GROUP(NPS_WGROUP)
#else
ScanLaunchFn scan_launcher_for_width(int words) {
    if (words < 1 || words > 32) return nullptr;
    switch ((words - 1) / 8) {
        case 0: return scan_launcher_group0(words);
        case 1: return scan_launcher_group1(words);
        case 2: return scan_launcher_group2(words);
        default: return scan_launcher_group3(words);
    }
}
HistLaunchFn hist_launcher_for_width(int words) {
    if (words < 1 || words > 32) return nullptr;
    switch ((words - 1) / 8) {
        case 0: return hist_launcher_group0(words);
        case 1: return hist_launcher_group1(words);
        case 2: return hist_launcher_group2(words);
        default: return hist_launcher_group3(words);
    }
}
cudaError_t launch_hist_scan_generic(const HistParams& p, int grid, size_t smem, cudaStream_t stream) {
    cudaError_t e =
        cudaFuncSetAttribute(hist_scan_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hist_scan_generic_kernel<<<grid, kScanThreads, smem, stream>>>(p);
    count_launch();
    return cudaGetLastError();
}
RefLaunchFn ref_launcher_for_width(int words) {
    if (words < 1 || words > 32) return nullptr;
    switch ((words - 1) / 8) {
        case 0: return ref_launcher_group0(words);
        case 1: return ref_launcher_group1(words);
        case 2: return ref_launcher_group2(words);
        default: return ref_launcher_group3(words);
    }
}
cudaError_t launch_ref_scan_generic(const RefParams& p, dim3 grid, size_t smem, cudaStream_t stream) {
    cudaError_t e =
        cudaFuncSetAttribute(ref_scan_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ref_scan_generic_kernel<<<grid, kScanThreads, smem, stream>>>(p);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_scan_generic(const ScanParams& p, int grid, size_t smem, cudaStream_t stream) {
    cudaError_t e =
        cudaFuncSetAttribute(scan_topk_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    scan_topk_generic_kernel<<<grid, kScanThreads, smem, stream>>>(p);
    count_launch();
    return cudaGetLastError();
}
#endif

}  // namespace nps
