// Single-warp issue-rate probe (sm_100a): cycles per instruction of ONE warp executing K independent dependency
// chains, for integer (IMAD), FP64 (DADD) and shared-memory loads.  Tells how much instruction-level parallelism an
// in-order warp can actually use — the ceiling for software pipelining inside one thread.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_issue warp_issue.cu ; run: ./warp_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int KIND>
__global__ void probe(int iters, long long* cyc, double* sink, int warps) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
    __syncthreads();
    unsigned x[K]; double d[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { x[k] = threadIdx.x + k; d[k] = 1.0 + k + threadIdx.x * 1e-9; }
    const double c = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (KIND == 0) x[k] = x[k] * 3u + 7u;                       // IMAD chain
                else if (KIND == 1) d[k] = __dadd_rn(d[k], c);              // DADD chain
                else if (KIND == 2) x[k] = ((unsigned*)sm)[(x[k] & 1023u)];  // dependent LDS chain
                else { x[k] = x[k] * 3u + 7u; d[k] = __dadd_rn(d[k], c); }  // one IMAD + one DADD per chain
            }
        }
    }
    long long t1 = clock64();
    double s = 0; unsigned u = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) { s += d[k]; u += x[k]; }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s + u;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K, int KIND>
void run(const char* name, int warps) {
    long long* cyc; double* sink;
    cudaMalloc(&cyc, 8 * 1024); cudaMalloc(&sink, 8 * 1024 * 1024);
    const int iters = 2000;
    probe<K, KIND><<<1, 32 * warps>>>(iters, cyc, sink, warps);
    probe<K, KIND><<<1, 32 * warps>>>(iters, cyc, sink, warps);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (KIND == 3 ? 2.0 : 1.0) * K * 16.0 * iters;
    printf("%-28s K=%d warps/SM=%2d : %.2f cycles per warp-instruction (per warp), %.2f per scheduler slot\n", name, K, warps,
           (double)h / per, (double)h / per / ((warps + 3) / 4));
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    run<1, 0>("IMAD, 1 chain", 1); run<2, 0>("IMAD, 2 chains", 1); run<4, 0>("IMAD, 4 chains", 1); run<8, 0>("IMAD, 8 chains", 1);
    run<1, 1>("DADD, 1 chain", 1); run<2, 1>("DADD, 2 chains", 1); run<4, 1>("DADD, 4 chains", 1); run<8, 1>("DADD, 8 chains", 1);
    run<1, 2>("LDS dependent, 1 chain", 1); run<4, 2>("LDS dependent, 4 chains", 1);
    run<1, 3>("IMAD+DADD, 1 chain", 1); run<4, 3>("IMAD+DADD, 4 chains", 1);
    run<4, 0>("IMAD, 4 chains", 4); run<4, 0>("IMAD, 4 chains", 16);
    run<4, 1>("DADD, 4 chains", 4); run<4, 1>("DADD, 4 chains", 16);
    run<1, 3>("IMAD+DADD, 1 chain", 16); run<4, 3>("IMAD+DADD, 4 chains", 16);
    return 0;
}
