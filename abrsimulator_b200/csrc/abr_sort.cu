// Session order by trace: sessions are independent (Simulator.py:93-131 holds exactly one), so the
// order in which an environment keeps them is free.  Kept sorted by trace, every thread block of the step kernels
// follows one trace and takes the shared-memory path whatever order the caller's sessions come in.
//
// abr_sort_by_trace builds that order: perm[p] = caller's index of the session at environment position p, a stable
// (hence deterministic) radix sort of (trace id, session index) pairs.  Set-up code, run once per session->trace
// assignment; the radix sort itself is CUB's (header-only, part of the CUDA toolkit).
#include <cub/device/device_radix_sort.cuh>

#include "abr_common.cuh"

namespace abr {

namespace {

__global__ void __launch_bounds__(256) abr_iota_kernel(int32_t* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

}  // namespace

// d_tmp: caller-provided scratch of *tmp_bytes bytes; with d_tmp == nullptr only *tmp_bytes is set (layout:
// [sorted keys n][iota n][CUB temp storage]).
cudaError_t launch_sort_by_trace(const int32_t* d_trace_id, int n, int n_traces, int32_t* d_perm, void* d_tmp,
                                 size_t* tmp_bytes, cudaStream_t st) {
    int end_bit = 1;
    while (end_bit < 31 && (1ll << end_bit) < (long long)n_traces) ++end_bit;
    size_t cub_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                                    (const int32_t*)nullptr, (int32_t*)nullptr, n, 0, end_bit, st);
    if (e != cudaSuccess) return e;
    const size_t arr = ((size_t)n * sizeof(int32_t) + 255) & ~(size_t)255;
    if (!d_tmp) { *tmp_bytes = 2 * arr + cub_bytes; return cudaSuccess; }
    int32_t* keys_out = reinterpret_cast<int32_t*>(d_tmp);
    int32_t* iota = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(d_tmp) + arr);
    void* cub_tmp = reinterpret_cast<char*>(d_tmp) + 2 * arr;
    abr_iota_kernel<<<(n + 255) / 256, 256, 0, st>>>(iota, n);
    count_launch();
    e = cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, d_trace_id, keys_out, iota, d_perm, n, 0, end_bit, st);
    count_launch(2);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace abr
