// fp64_latency.cu -- single-thread latency of the fp64 operations k_peac_frame's merge step is made of (measurement aid):
// dependent chains of DADD / DMUL / division / sqrt, and peac::eig33 (the cyclic Jacobi eigen-solve of a candidate fit) on
// scatter matrices like the ones a 10x10 block of a planar depth patch produces.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I manhattanslam_b200/csrc -o /tmp/fp64_latency tools/fp64_latency.cu && /tmp/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
#include "peac_frame.cuh"

__global__ void k_lat(double *out, long long *cyc, double seed) {
    double x = seed, y = seed * 0.37 + 1.0;
    long long t0, t1;
    const int N = 256;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; i++) x = x + y;
    t1 = clock64(); cyc[0] = (t1 - t0) / N;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; i++) x = x * 1.0000001;
    t1 = clock64(); cyc[1] = (t1 - t0) / N;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; i++) x = y / (x + 2.0);
    t1 = clock64(); cyc[2] = (t1 - t0) / N;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; i++) x = sqrt(x + 2.0);
    t1 = clock64(); cyc[3] = (t1 - t0) / N;
    // eig33 on a planar-patch scatter matrix, perturbed every iteration
    double K[9] = {3.1e4, 1.2e3, -4.0e2, 1.2e3, 2.7e4, 9.0e2, -4.0e2, 9.0e2, 1.5e1};
    double s[3], V[9], acc = 0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
        K[0] += 1.0 + acc * 1e-30, K[4] += 0.5, K[8] += 0.01;
        peac::eig33(K, s, V);
        acc += s[0] + V[0];
    }
    t1 = clock64(); cyc[4] = (t1 - t0) / 64;
    // a whole candidate fit: merged sums -> Stats::compute
    peac::Node a, b, m;
    for (int k = 0; k < 9; k++) a.s[k] = 1.0 + k * seed, b.s[k] = 2.0 + k;
    a.s[3] = 5e4, a.s[4] = 4e4, a.s[5] = 9e4, b.s[3] = 6e4, b.s[4] = 3e4, b.s[5] = 8e4;
    a.N = 100, b.N = 100, a.rid = 0, b.rid = 1;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
        a.s[3] += 1.0 + acc * 1e-30;
        peac::merged(a, b, m);
        acc += m.mse;
    }
    t1 = clock64(); cyc[5] = (t1 - t0) / 64;
    out[0] = x + acc;
}

int main() {
    double *d_out;
    long long *d_cyc, h[6];
    cudaMalloc(&d_out, 8);
    cudaMalloc(&d_cyc, 48);
    k_lat<<<1, 1>>>(d_out, d_cyc, 1.25);
    k_lat<<<1, 1>>>(d_out, d_cyc, 1.25);
    cudaMemcpy(h, d_cyc, 48, cudaMemcpyDeviceToHost);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("fp64 dependent-chain latency in cycles (SM clock %d kHz): dadd %lld  dmul %lld  ddiv(+dadd) %lld  dsqrt(+dadd) %lld  eig33 %lld  merged() %lld\n",
           clk, h[0], h[1], h[2], h[3], h[4], h[5]);
    return 0;
}
