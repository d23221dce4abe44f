// fp64_throughput.cu -- per-SM issue rate of the operations the reference's float/double cost expressions are made of
// (calculateCost, src/SurfelFusion.cpp:333-355: every term is computed in float, widened, added in double, narrowed):
// F2F.F64.F32, F2F.F32.F64, DADD, DFMA next to FFMA and integer IMAD/LOP3 as yardsticks.  One CTA of 1024 threads per SM,
// eight independent chains per thread, so that latency is hidden and the pipe's rate shows (measurement aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/fp64_throughput tools/fp64_throughput.cu && /tmp/fp64_throughput
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 512, CH = 8;

template <int OP>
__global__ void __launch_bounds__(1024) k_thr(float *out, long long *cyc, float seed) {
    float f[CH];
    double d[CH];
    int n[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) f[c] = seed + c + threadIdx.x * 1e-3f, d[c] = (double)f[c] * 1.0000001, n[c] = (int)threadIdx.x + c;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (OP == 0) {  // float -> double -> (DADD) is avoided: widen, xor a bit in the integer domain, narrow = 1 F2F.F64.F32 + 1 F2F.F32.F64
                double w;
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(w) : "f"(f[c]));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[c]) : "d"(w));
            } else if (OP == 1) {
                asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(1.0e-9));
            } else if (OP == 2) {
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(1.0000001), "d"(1.0e-9));
            } else if (OP == 3) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(1.0000001f), "f"(1.0e-9f));
            } else if (OP == 4) {
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(n[c]) : "r"(3), "r"(7));
            } else if (OP == 5) {  // widen only (the double is consumed by a cheap integer op on its high word)
                double w;
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(w) : "f"(f[c]));
                int lo, hi;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "d"(w));
                f[c] = __int_as_float((hi ^ lo) | 0x3f000000);
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) acc += f[c] + (float)d[c] + (float)n[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
static double run(float *d_out, long long *d_cyc, int sms, int opsPerIter) {
    k_thr<OP><<<sms, 1024>>>(d_out, d_cyc, 1.25f);
    k_thr<OP><<<sms, 1024>>>(d_out, d_cyc, 1.25f);
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; i++) avg += (double)h[i];
    avg /= sms;
    return (double)ITER * CH * opsPerIter * 1024 / avg;  // thread-operations per cycle per SM
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *d_out;
    long long *d_cyc;
    cudaMalloc(&d_out, sizeof(float) * 1024 * sms);
    cudaMalloc(&d_cyc, sizeof(long long) * sms);
    printf("thread-operations per cycle per SM (%d SMs, 32 warps per SM, %d chains per thread):\n", sms, CH);
    printf("  cvt.f64.f32 + cvt.rn.f32.f64 pair   %.1f conversions\n", run<0>(d_out, d_cyc, sms, 2));
    printf("  cvt.f64.f32 (+ mov/xor/or)          %.1f\n", run<5>(d_out, d_cyc, sms, 1));
    printf("  add.rn.f64                          %.1f\n", run<1>(d_out, d_cyc, sms, 1));
    printf("  fma.rn.f64                          %.1f\n", run<2>(d_out, d_cyc, sms, 1));
    printf("  fma.rn.f32                          %.1f\n", run<3>(d_out, d_cyc, sms, 1));
    printf("  mad.lo.s32                          %.1f\n", run<4>(d_out, d_cyc, sms, 1));
    return 0;
}
