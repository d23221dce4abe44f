// Repro of a ptxas 12.9 / driver-JIT miscompile on sm_100: integer max(best, max(mn, -mx)) with mn a min-chain and
// mx a max-chain returns max(d) instead (variants tA/tC/tD/tE print 43, expected 3; tB/tF/tG are correct; -Xptxas -O0 is correct).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o repro tools/ptxas_vimnmx3_repro.cu
#include <cstdio>
#include <cstdint>
__global__ void tA(const int *d, int *out) {
    int best = 0;
    for (int s0 = 0; s0 < 16; s0++) {
        int mn = d[s0], mx = d[s0];
        for (int k = 1; k < 9; k++) { int v = d[(s0 + k) & 15]; mn = min(mn, v); mx = max(mx, v); }
        best = max(best, max(mn, -mx));
    }
    out[0] = best;
}
__global__ void tB(const int *d, int *out) {
    int best = 0;
    for (int s0 = 0; s0 < 16; s0++) {
        int mn = d[s0];
        for (int k = 1; k < 9; k++) { int v = d[(s0 + k) & 15]; mn = min(mn, v); }
        best = max(best, mn);
    }
    out[0] = best;
}
__global__ void tC(const int *d, int *out) {
    int best = 0;
    for (int s0 = 0; s0 < 16; s0++) {
        int mn = d[s0], mq = -d[s0];
        for (int k = 1; k < 9; k++) { int v = d[(s0 + k) & 15]; mn = min(mn, v); mq = min(mq, -v); }
        best = max(best, max(mn, mq));
    }
    out[0] = best;
}
__device__ __forceinline__ int pmin(int a, int b) { int r; asm volatile("min.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ int pmax(int a, int b) { int r; asm volatile("max.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__global__ void tD(const int *d, int *out) {
    int best = 0;
    for (int s0 = 0; s0 < 16; s0++) {
        int mn = d[s0], mx = d[s0];
        for (int k = 1; k < 9; k++) { int v = d[(s0 + k) & 15]; mn = pmin(mn, v); mx = pmax(mx, v); }
        best = pmax(best, pmax(mn, -mx));
    }
    out[0] = best;
}
__global__ void tE(const int *dd, int *out) {
    int d[16];
    for (int i = 0; i < 16; i++) d[i] = dd[i];
    int lo2[16], hi2[16], lo4[16], hi4[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { lo2[i] = min(d[i], d[(i + 1) & 15]); hi2[i] = max(d[i], d[(i + 1) & 15]); }
#pragma unroll
    for (int i = 0; i < 16; i++) { lo4[i] = min(lo2[i], lo2[(i + 2) & 15]); hi4[i] = max(hi2[i], hi2[(i + 2) & 15]); }
    int best = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        int lo9 = min(min(lo4[i], lo4[(i + 4) & 15]), d[(i + 8) & 15]);
        int hi9 = max(max(hi4[i], hi4[(i + 4) & 15]), d[(i + 8) & 15]);
        best = max(best, max(lo9, -hi9));
    }
    out[0] = best;
}
__global__ void tF(const int *d, int *out) {
    float best = 0;
    for (int s0 = 0; s0 < 16; s0++) {
        float mn = d[s0], mx = d[s0];
        for (int k = 1; k < 9; k++) { float v = d[(s0 + k) & 15]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
        best = fmaxf(best, fmaxf(mn, -mx));
    }
    out[0] = (int)best;
}
// G: min path only but with separate hi path stored to out[1] (no negation)
__global__ void tG(const int *d, int *out) {
    int best = 0, bestn = 1000;
    for (int s0 = 0; s0 < 16; s0++) {
        int mn = d[s0], mx = d[s0];
        for (int k = 1; k < 9; k++) { int v = d[(s0 + k) & 15]; mn = min(mn, v); mx = max(mx, v); }
        best = max(best, mn); bestn = min(bestn, mx);
    }
    out[0] = max(best, -bestn);
}
int main() {
    int h[16] = {3, 6, 5, 43, 39, 40, 7, 1, 7, 4, 4, 6, 5, 3, 6, 4};
    int *d, *o; cudaMalloc(&d, 64); cudaMalloc(&o, 64); cudaMemcpy(d, h, 64, cudaMemcpyHostToDevice);
    int r;
#define RUN(K) K<<<1, 1>>>(d, o); cudaMemcpy(&r, o, 4, cudaMemcpyDeviceToHost); printf(#K " %d\n", r);
    RUN(tA) RUN(tB) RUN(tC) RUN(tD) RUN(tE) RUN(tF) RUN(tG)
    return 0;
}
