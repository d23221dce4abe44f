// msl_common.cuh -- shared helpers for the sm_100a front-end kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/msl_frontend.h"

namespace msl {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

#define MSL_CUDA(expr)                                                                           \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return msl::fail(MSL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));  \
    } while (0)

#define MSL_LAUNCH_CHECK()                                                                        \
    do {                                                                                          \
        msl::g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess)                                                                    \
            return msl::fail(MSL_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
    } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

#ifdef __CUDACC__

__device__ __forceinline__ int warp_incl_scan(int v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (unsigned)o) v += t;
    }
    return v;
}

// In-place exclusive scan of a[0..n) (shared or global memory) by the whole CTA (1-D block, <= 1024
// threads).  ws = shared scratch of >= 34 ints.  Returns the total.  Contains __syncthreads().
__device__ inline int block_excl_scan(int *a, int n, int *ws) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5, nw = (nt + 31) >> 5;
    if (tid == 0) ws[33] = 0;
    __syncthreads();
    for (int base = 0; base < n; base += nt) {
        int i = base + tid;
        int v = (i < n) ? a[i] : 0;
        int inc = warp_incl_scan(v);
        if (lane == 31) ws[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int s = (lane < nw) ? ws[lane] : 0;
            int si = warp_incl_scan(s);
            ws[lane] = si - s;
            if (lane == 31) ws[32] = si;
        }
        __syncthreads();
        int carry = ws[33];
        if (i < n) a[i] = carry + ws[wid] + inc - v;
        __syncthreads();
        if (tid == 0) ws[33] = carry + ws[32];
        __syncthreads();
    }
    const int total = ws[33];
    __syncthreads();  // ws may be reused by the next call
    return total;
}

#endif  // __CUDACC__

}  // namespace msl
