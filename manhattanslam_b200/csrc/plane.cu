// plane.cu -- B200-native plane pre-stage (replaces PlaneDetection::readDepthImage,
// src/PlaneExtractor.cpp:44-76, and the peac pre-stage: PlaneSeg ctor + Stats::compute per 10x10 block,
// include/peac/AHCPlaneSeg.hpp:148-181,235-312, and the node/edge initialisation of
// PlaneFitter::initGraph, include/peac/AHCPlaneFitter.hpp:756-928).
//
//   k_plane_cloud   P1     one thread per half-resolution pixel: u16 depth -> double XYZ (optional output)
//   k_plane_blocks  P2-P4  one thread per 10x10 block: strict validity / depth-discontinuity scan in the
//                          reference's row-major order (bit-exact double sums), 3x3 scatter matrix,
//                          cyclic-Jacobi eigen-solve, normal orientation, mse, curvature, seed test
//   k_plane_edges   P5     one CTA per frame: the row pass then the column pass of initGraph with their
//                          --j/++j stepping, one thread per row / column
//   k_peac_frame    f2     one CTA per frame: ahCluster + refineDetails (peac_frame.cuh) -> membershipImg, planes
// All arithmetic is fp64 in the reference's operation order (file built with -fmad=false).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "msl_common.cuh"
#include "peac_frame.cuh"

using namespace msl;

namespace {

constexpr int WIN = 10;  // windowWidth/windowHeight, AHCPlaneFitter.hpp:156-160
// ParamSet defaults (AHCParamSet.hpp:68-75); never overridden by ManhattanSLAM
#define P_DEPTH_SIGMA 1.6e-6
#define P_STDTOL_INIT 5.0
#define P_Z_NEAR 500.0
#define P_Z_FAR 4000.0
#define P_DEPTH_ALPHA 0.04
#define P_DEPTH_CHANGE_TOL 0.02

struct PlaneParams {
    int w, h, W2, H2, Nw, Nh;
    int dstride;            // depth row stride in pixels
    size_t frameStride;     // depth frame stride in pixels
    float fx, fy, cx, cy, factor;
    double thNear;          // T_ang(P_INIT, z <= z_near), evaluated on the host with std::cos
    double angFactor, angNear;
};

__device__ __forceinline__ double depth_z(const uint16_t *depth, const PlaneParams &P, int row, int col) {
    return (double)depth[(size_t)(2 * row) * P.dstride + 2 * col] * (double)P.factor;  // :64
}

__global__ void __launch_bounds__(256) k_plane_cloud(PlaneParams P, const uint16_t *__restrict__ depth, double *__restrict__ cloud) {
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), row = blockIdx.y * 8 + (threadIdx.x >> 5), b = blockIdx.z;
    if (col >= P.W2 || row >= P.H2) return;
    const uint16_t *d = depth + b * P.frameStride;
    const double z = depth_z(d, P, row, col);
    const double x = ((double)(2 * col) - (double)P.cx) * z / (double)P.fx;  // :69-70
    const double y = ((double)(2 * row) - (double)P.cy) * z / (double)P.fy;
    double *o = cloud + ((size_t)b * P.W2 * P.H2 + (size_t)row * P.W2 + col) * 3;
    o[0] = x, o[1] = y, o[2] = z;
}

// cyclic Jacobi, identical operation order to oracle/plane_oracle.cpp (stand-in for Eigen's solver)
__device__ void eig33sym_dev(const double K[9], double s[3], double V[9]) {
    double a[3][3] = {{K[0], K[1], K[2]}, {K[3], K[4], K[5]}, {K[6], K[7], K[8]}};
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 32; sweep++) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-60 || off <= 1e-34 * diag) break;
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - sn * akq;
                    a[k][q] = sn * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - sn * aqk;
                    a[q][k] = sn * apk + c * aqk;
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - sn * vkq;
                    v[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    int o0 = 0, o1 = 1, o2 = 2;
    const double e[3] = {a[0][0], a[1][1], a[2][2]};
    // same selection order as the oracle's exchange sort
    if (e[o1] < e[o0]) { int t = o0; o0 = o1; o1 = t; }
    if (e[o2] < e[o0]) { int t = o0; o0 = o2; o2 = t; }
    if (e[o2] < e[o1]) { int t = o1; o1 = o2; o2 = t; }
    const int ord[3] = {o0, o1, o2};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        s[i] = e[ord[i]];
#pragma unroll
        for (int k = 0; k < 3; k++) V[k * 3 + i] = v[k][ord[i]];
    }
}

__global__ void __launch_bounds__(128)
    k_plane_blocks(PlaneParams P, const uint16_t *__restrict__ depth, msl_block_stat *__restrict__ blocks, uint8_t *__restrict__ seed) {
    const int blk = blockIdx.x * 128 + threadIdx.x, b = blockIdx.y;
    const int nb = P.Nw * P.Nh;
    if (blk >= nb) return;
    const uint16_t *d = depth + b * P.frameStride;
    const int bi = blk / P.Nw, bj = blk % P.Nw;
    double sx = 0, sy = 0, sz = 0, sxx = 0, syy = 0, szz = 0, sxy = 0, syz = 0, sxz = 0;
    int N = 0;
    bool valid = true;
    const int r0 = bi * WIN, c0 = bj * WIN;
    for (int i = r0, ic = 0; ic < WIN && i < P.H2 && valid; ++i, ++ic) {
        for (int j = c0, jc = 0; jc < WIN && j < P.W2; ++j, ++jc) {
            const double z = depth_z(d, P, i, j);
            if (z == 0) {  // ImagePointCloud::get, include/PlaneExtractor.h:48-56 (u16 input is never NaN)
                valid = false;
                break;
            }
            const double tdz = P_DEPTH_ALPHA * fabs(z) + P_DEPTH_CHANGE_TOL;  // T_dz, AHCParamSet.hpp:140-142
            if (j + 1 < P.W2) {
                const double zn = depth_z(d, P, i, j + 1);
                if (zn != 0 && fabs(z - zn) > tdz) {
                    valid = false;
                    break;
                }
            }
            if (i + 1 < P.H2) {
                const double zn = depth_z(d, P, i + 1, j);
                if (zn != 0 && fabs(z - zn) > tdz) {
                    valid = false;
                    break;
                }
            }
            const double x = ((double)(2 * j) - (double)P.cx) * z / (double)P.fx;
            const double y = ((double)(2 * i) - (double)P.cy) * z / (double)P.fy;
            sx += x, sy += y, sz += z;
            sxx += x * x, syy += y * y, szz += z * z;
            sxy += x * y, syz += y * z, sxz += x * z;
            ++N;
        }
    }
    msl_block_stat B;
    memset(&B, 0, sizeof(B));
    B.N = valid ? N : 0;
    B.nouse = valid ? 0 : 1;
    if (B.N < 4) {
        B.mse = B.curvature = __longlong_as_double(0x7ff8000000000000LL);  // quiet NaN, :296
    } else {
        const double sc = 1.0 / N;
        B.center[0] = sx * sc, B.center[1] = sy * sc, B.center[2] = sz * sc;
        double K[9] = {sxx - sx * sx * sc, sxy - sx * sy * sc, sxz - sx * sz * sc, 0, syy - sy * sy * sc,
                       syz - sy * sz * sc, 0, 0, szz - sz * sz * sc};
        K[3] = K[1], K[6] = K[2], K[7] = K[5];
        double sv[3], V[9];
        eig33sym_dev(K, sv, V);
        if ((V[0] * B.center[0] + V[3] * B.center[1]) + V[6] * B.center[2] <= 0) {
            B.normal[0] = V[0], B.normal[1] = V[3], B.normal[2] = V[6];
        } else {
            B.normal[0] = -V[0], B.normal[1] = -V[3], B.normal[2] = -V[6];
        }
        B.mse = sv[0] * sc;
        B.curvature = sv[0] / ((sv[0] + sv[1]) + sv[2]);
    }
    const double tm = P_DEPTH_SIGMA * B.center[2] * B.center[2] + P_STDTOL_INIT;  // T_mse(P_INIT), :92
    blocks[(size_t)b * nb + blk] = B;
    seed[(size_t)b * nb + blk] = (B.mse < tm * tm && !B.nouse) ? 1 : 0;
}

__device__ __forceinline__ double t_ang(const PlaneParams &P, double z) {  // T_ang(P_INIT), :113-118
    if (z <= P_Z_NEAR) return P.thNear;
    double cz = z < P_Z_FAR ? z : P_Z_FAR;
    return cos(P.angFactor * cz + P.angNear - P.angFactor * P_Z_NEAR);
}
__device__ __forceinline__ double nsim(const msl_block_stat *B, int a, int b) {
    return fabs((B[a].normal[0] * B[b].normal[0] + B[a].normal[1] * B[b].normal[1]) + B[a].normal[2] * B[b].normal[2]);
}

__global__ void __launch_bounds__(128)
    k_plane_edges(PlaneParams P, const msl_block_stat *__restrict__ blocks, const uint8_t *__restrict__ seed, uint8_t *__restrict__ edges) {
    const int b = blockIdx.x, t = threadIdx.x, Nw = P.Nw, Nh = P.Nh, nb = Nw * Nh;
    const msl_block_stat *B = blocks + (size_t)b * nb;
    const uint8_t *G = seed + (size_t)b * nb;
    uint8_t *E = edges + (size_t)b * nb;
    for (int k = t; k < nb; k += blockDim.x) E[k] = 0;
    __syncthreads();
    for (int i = t; i < Nh; i += blockDim.x) {  // row pass, AHCPlaneFitter.hpp:840-874
        for (int j = 1; j < Nw; j += 2) {
            const int c = i * Nw + j;
            if (G[c - 1] == 0) { --j; continue; }
            if (G[c] == 0) continue;
            if (j < Nw - 1 && G[c + 1] == 0) { ++j; continue; }
            const double th = t_ang(P, B[c].center[2]);
            if ((j < Nw - 1 && nsim(B, c - 1, c + 1) >= th) || (j == Nw - 1 && nsim(B, c, c - 1) >= th)) {
                E[c] |= 1, E[c - 1] |= 2;
                if (j < Nw - 1) E[c] |= 2, E[c + 1] |= 1;
            } else
                --j;
        }
    }
    __syncthreads();
    for (int j = t; j < Nw; j += blockDim.x) {  // column pass, :876-910
        for (int i = 1; i < Nh; i += 2) {
            const int c = i * Nw + j;
            if (G[c - Nw] == 0) { --i; continue; }
            if (G[c] == 0) continue;
            if (i < Nh - 1 && G[c + Nw] == 0) { ++i; continue; }
            const double th = t_ang(P, B[c].center[2]);
            if ((i < Nh - 1 && nsim(B, c - Nw, c + Nw) >= th) || (i == Nh - 1 && nsim(B, c, c - Nw) >= th)) {
                E[c] |= 4, E[c - Nw] |= 8;
                if (i < Nh - 1) E[c] |= 8, E[c + Nw] |= 4;
            } else
                --i;
        }
    }
}

static_assert(sizeof(msl_plane_rec) == sizeof(peac::PlaneOut) && sizeof(msl_plane_rec) == 64, "plane record layout");

// ahCluster + refineDetails of one frame per CTA (peac_frame.cuh); 196 KB of dynamic shared memory
struct PeacScratch {  // per-frame strides are implied: npix, rfqCap, visCap
    float *dist;
    uint32_t *rfq;
    int *own, *visC;
    float *visDist;
    uint8_t *visFlag;
    int rfqCap, visCap;
    long long *prof;  // 8 per frame, or null
};

// SH = peac::Shared: the working set in the CTA's dynamic shared memory (frames of <= 768 blocks); SH = peac::SharedBig: in
// global memory, one record per frame (`big`), for frames of up to 3072 blocks (1280x960)
template <class SH, int MAXT>
__global__ void __launch_bounds__(MAXT)
    k_peac_frame(peac::Geo g, size_t frameStride, const uint16_t *__restrict__ depth, const msl_block_stat *__restrict__ blocks,
                 const uint8_t *__restrict__ seed, const uint8_t *__restrict__ edges, int32_t *__restrict__ membership, PeacScratch sc,
                 SH *big, peac::PlaneOut *__restrict__ planes, int planeCap, int32_t *__restrict__ planeCount,
                 int32_t *__restrict__ frameError) {
    extern __shared__ __align__(16) unsigned char peac_smem[];
    const int b = blockIdx.x, nb = g.Nw * g.Nh;
    SH &S = big ? big[b] : *reinterpret_cast<SH *>(peac_smem);
    const size_t npix = (size_t)g.W2 * g.H2;
    peac::Flood F;
    F.distMap = sc.dist + b * npix, F.own = sc.own + b * npix;
    F.rfq = sc.rfq + (size_t)b * sc.rfqCap, F.rfqCap = sc.rfqCap;
    F.visC = sc.visC + (size_t)b * sc.visCap, F.visDist = sc.visDist + (size_t)b * sc.visCap, F.visFlag = sc.visFlag + (size_t)b * sc.visCap;
    F.visCap = sc.visCap;
    F.prof = sc.prof ? sc.prof + 16 * (size_t)b : nullptr;
    peac::frame(S, g, depth + b * frameStride, blocks + (size_t)b * nb, seed + (size_t)b * nb, edges + (size_t)b * nb,
                membership + b * npix, F, planes + (size_t)b * planeCap, planeCap, planeCount + b, frameError + b, (int)threadIdx.x,
                (int)blockDim.x);
}

}  // namespace

struct msl_plane {
    int w, h, maxBatch, device;
    int W2, H2, Nw, Nh;
    cudaStream_t stream = nullptr;
    uint16_t *d_depth = nullptr;
    double *d_cloud = nullptr;
    msl_block_stat *d_blocks = nullptr;
    uint8_t *d_seed = nullptr, *d_edges = nullptr;
    // msl_plane_detect*: allocated on first use
    int32_t *d_mem = nullptr, *d_count = nullptr, *d_ferr = nullptr;
    float *d_dist = nullptr, *d_visDist = nullptr;
    uint32_t *d_rfq = nullptr;
    int *d_own = nullptr, *d_visC = nullptr;
    uint8_t *d_visFlag = nullptr;
    peac::SharedBig *d_big = nullptr;  // frames of more than 768 blocks: the working set in global memory
    msl_plane_rec *d_planes = nullptr;
    int planeCap = 0;
    long long *d_prof = nullptr;  // msl_plane_debug_profile: 8 stamps per frame of the last detect call
    int profFrames = 0;
    int pendingCheck = 0;  // frames of an enqueued detect whose per-frame error words have not been read yet
};

static void plane_free(msl_plane *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    void *ptrs[] = {p->d_depth, p->d_cloud, p->d_blocks, p->d_seed, p->d_edges, p->d_mem, p->d_count, p->d_ferr,
                    p->d_dist, p->d_rfq, p->d_planes, p->d_own, p->d_visC, p->d_visDist, p->d_visFlag, p->d_big, p->d_prof};
    for (void *q : ptrs)
        if (q) cudaFree(q);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

extern "C" {

int msl_plane_create(int w, int h, int max_batch, int device, msl_plane **out) {
    if (!out) return fail(MSL_ERR_INVALID, "msl_plane_create: null out");
    *out = nullptr;
    if (w < 2 * WIN || h < 2 * WIN || max_batch < 1) return fail(MSL_ERR_INVALID, "msl_plane_create: parameter out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device || device < 0)
        return fail(MSL_ERR_CUDA, "msl_plane_create: no usable CUDA device (there is no CPU fallback)");
    MSL_CUDA(cudaSetDevice(device));
    msl_plane *p = new msl_plane();
    p->w = w, p->h = h, p->maxBatch = max_batch, p->device = device;
    p->W2 = (int)ceil(w / 2.0), p->H2 = (int)ceil(h / 2.0);  // src/PlaneExtractor.cpp:51-52
    p->Nw = p->W2 / WIN, p->Nh = p->H2 / WIN;
    const size_t B = max_batch, nb = (size_t)p->Nw * p->Nh;
    cudaError_t e = cudaMalloc((void **)&p->d_depth, B * w * h * 2);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_cloud, B * p->W2 * p->H2 * 3 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_blocks, B * nb * sizeof(msl_block_stat));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_seed, B * nb);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_edges, B * nb);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        plane_free(p);
        return fail(MSL_ERR_CUDA, std::string("msl_plane_create: ") + cudaGetErrorString(e));
    }
    *out = p;
    return MSL_OK;
}

void msl_plane_destroy(msl_plane *p) { plane_free(p); }
void *msl_plane_stream(msl_plane *p) { return p ? (void *)p->stream : nullptr; }
static int plane_check_frames(msl_plane *p) {  // after a synchronize: did any frame of the last detect overflow?
    if (!p->pendingCheck) return MSL_OK;
    const int n = p->pendingCheck;
    p->pendingCheck = 0;
    std::vector<int32_t> e(n);
    MSL_CUDA(cudaMemcpy(e.data(), p->d_ferr, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
    for (int b = 0; b < n; b++)
        if (e[b] != 0)
            return fail(MSL_ERR_CAPACITY, e[b] == peac::PEAC_ERR_QUEUE ? "msl_plane_detect: region-grow queue exceeded"
                                                                      : "msl_plane_detect: more than 128 planes in a frame");
    return MSL_OK;
}

int msl_plane_sync(msl_plane *p) {
    if (!p) return fail(MSL_ERR_INVALID, "null handle");
    MSL_CUDA(cudaSetDevice(p->device));
    MSL_CUDA(cudaStreamSynchronize(p->stream));
    return plane_check_frames(p);
}

int msl_plane_prestage_dev(msl_plane *p, const uint16_t *d_depth, int dstride_px, size_t frame_stride_px, int batch,
                           const float K[4], float depth_map_factor, double *d_cloud_xyz, msl_block_stat *d_blocks,
                           uint8_t *d_seed, uint8_t *d_edges) {
    if (!p || !d_depth || !K) return fail(MSL_ERR_INVALID, "msl_plane_prestage_dev: null argument");
    if (batch < 1 || batch > p->maxBatch || dstride_px < p->w) return fail(MSL_ERR_INVALID, "msl_plane_prestage_dev: bad batch/stride");
    MSL_CUDA(cudaSetDevice(p->device));
    PlaneParams P;
    P.w = p->w, P.h = p->h, P.W2 = p->W2, P.H2 = p->H2, P.Nw = p->Nw, P.Nh = p->Nh;
    P.dstride = dstride_px, P.frameStride = frame_stride_px;
    P.fx = K[0], P.fy = K[1], P.cx = K[2], P.cy = K[3], P.factor = depth_map_factor;
    const double angle_near = 15.0 * M_PI / 180.0, angle_far = 90.0 * M_PI / 180.0;
    P.angFactor = (angle_far - angle_near) / (P_Z_FAR - P_Z_NEAR);
    P.angNear = angle_near;
    P.thNear = std::cos(P.angFactor * P_Z_NEAR + angle_near - P.angFactor * P_Z_NEAR);
    msl_block_stat *blocks = d_blocks ? d_blocks : p->d_blocks;
    uint8_t *seed = d_seed ? d_seed : p->d_seed;
    uint8_t *edges = d_edges ? d_edges : p->d_edges;
    if (d_cloud_xyz) {
        k_plane_cloud<<<dim3(cdiv(P.W2, 32), cdiv(P.H2, 8), batch), 256, 0, p->stream>>>(P, d_depth, d_cloud_xyz);
        MSL_LAUNCH_CHECK();
    }
    k_plane_blocks<<<dim3(cdiv(P.Nw * P.Nh, 128), batch), 128, 0, p->stream>>>(P, d_depth, blocks, seed);
    MSL_LAUNCH_CHECK();
    k_plane_edges<<<batch, 128, 0, p->stream>>>(P, blocks, seed, edges);
    MSL_LAUNCH_CHECK();
    return MSL_OK;
}

constexpr int PLANE_CAP_INTERNAL = 128;  // records per frame of the handle's own buffer (host entry point) = peac::MAXPL

static int plane_detect_alloc(msl_plane *p) {
    if (p->d_mem) return MSL_OK;
    const size_t B = p->maxBatch, npix = (size_t)p->W2 * p->H2;
    cudaError_t e = cudaMalloc((void **)&p->d_mem, B * npix * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_dist, B * npix * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_rfq, B * npix * 4 * sizeof(uint32_t));
    // region grow by levels: one level holds at most npix entries = 4 * npix visits
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_own, B * npix * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_visC, B * npix * 4 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_visDist, B * npix * 4 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_visFlag, B * npix * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_count, B * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_ferr, B * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_planes, B * PLANE_CAP_INTERNAL * sizeof(msl_plane_rec));
    if (e == cudaSuccess) e = cudaMemset(p->d_ferr, 0, B * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_prof, B * 16 * sizeof(long long));
    if (e == cudaSuccess && p->Nw * p->Nh > peac::MAXB) e = cudaMalloc((void **)&p->d_big, B * sizeof(peac::SharedBig));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_peac_frame<peac::Shared, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(peac::Shared));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_peac_frame<peac::Shared, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(peac::Shared));
    if (e != cudaSuccess) return fail(MSL_ERR_CUDA, std::string("msl_plane_detect: ") + cudaGetErrorString(e));
    p->planeCap = PLANE_CAP_INTERNAL;
    return MSL_OK;
}

int msl_plane_detect_dev(msl_plane *p, const uint16_t *d_depth, int dstride_px, size_t frame_stride_px, int batch,
                         const float K[4], float depth_map_factor, int32_t *d_membership, int32_t *d_plane_count,
                         msl_plane_rec *d_planes, int plane_cap) {
    if (!p || !d_depth || !K || !d_membership || !d_plane_count || (plane_cap > 0 && !d_planes) || plane_cap < 0)
        return fail(MSL_ERR_INVALID, "msl_plane_detect_dev: bad argument");
    if (p->Nw * p->Nh > peac::MAXB_BIG) return fail(MSL_ERR_INVALID, "msl_plane_detect: frames of more than 3072 blocks (1280x960) are not supported");
    if ((size_t)p->W2 * p->H2 >= (1u << 20)) return fail(MSL_ERR_INVALID, "msl_plane_detect: frame too large");
    int rc = plane_detect_alloc(p);
    if (rc) return rc;
    // the pre-stage into the handle's own buffers (P1-P5), then one CTA per frame
    rc = msl_plane_prestage_dev(p, d_depth, dstride_px, frame_stride_px, batch, K, depth_map_factor, nullptr, p->d_blocks, p->d_seed,
                                p->d_edges);
    if (rc) return rc;
    peac::Geo g;
    g.W2 = p->W2, g.H2 = p->H2, g.Nw = p->Nw, g.Nh = p->Nh, g.dstride = dstride_px;
    g.fx = K[0], g.fy = K[1], g.cx = K[2], g.cy = K[3], g.factor = depth_map_factor;
    g.thMerge = std::cos(60.0 * M_PI / 180.0), g.thRefine = std::cos(30.0 * M_PI / 180.0);  // AHCParamSet.hpp:72-73
    const char *fs = getenv("MSL_PEAC_FLOOD_SERIAL");  // 1: the region grow as a FIFO on thread 0 (A/B, cross-check)
    g.floodSerial = (fs && atoi(fs) != 0) ? 1 : 0;
    PeacScratch sc;
    sc.dist = p->d_dist, sc.rfq = p->d_rfq, sc.own = p->d_own, sc.visC = p->d_visC, sc.visDist = p->d_visDist, sc.visFlag = p->d_visFlag;
    sc.rfqCap = 4 * p->W2 * p->H2, sc.visCap = 4 * p->W2 * p->H2;
    sc.prof = p->d_prof, p->profFrames = batch;
    int threads = 512;  // MSL_PEAC_THREADS = 64 | 128 | 256 | 512 | 1024: CTA size of k_peac_frame (measured r02e: 512 is the fastest -- the region grow scales with the threads, the eigen-solver spills at the 64 registers 1024 threads leave)
    if (const char *e = getenv("MSL_PEAC_THREADS")) {
        const int v = atoi(e);
        if (v == 64 || v == 128 || v == 256 || v == 512 || v == 1024) threads = v;
    }
#define PEAC_ARGS g, frame_stride_px, d_depth, p->d_blocks, p->d_seed, p->d_edges, d_membership, sc
#define PEAC_TAIL reinterpret_cast<peac::PlaneOut *>(d_planes), plane_cap, d_plane_count, p->d_ferr
    // two register budgets: up to 512 threads per CTA at 110 registers (no spills in the fp64 eigen-solver), 1024 at 64
    if (p->Nw * p->Nh <= peac::MAXB) {
        if (threads <= 512)
            k_peac_frame<peac::Shared, 512><<<batch, threads, sizeof(peac::Shared), p->stream>>>(PEAC_ARGS, (peac::Shared *)nullptr, PEAC_TAIL);
        else
            k_peac_frame<peac::Shared, 1024><<<batch, threads, sizeof(peac::Shared), p->stream>>>(PEAC_ARGS, (peac::Shared *)nullptr, PEAC_TAIL);
    } else {
        if (threads <= 512)
            k_peac_frame<peac::SharedBig, 512><<<batch, threads, 0, p->stream>>>(PEAC_ARGS, p->d_big, PEAC_TAIL);
        else
            k_peac_frame<peac::SharedBig, 1024><<<batch, threads, 0, p->stream>>>(PEAC_ARGS, p->d_big, PEAC_TAIL);
    }
#undef PEAC_ARGS
#undef PEAC_TAIL
    MSL_LAUNCH_CHECK();
    p->pendingCheck = batch > p->pendingCheck ? batch : p->pendingCheck;
    return MSL_OK;
}

// Measurement aid: per frame of the last detect call 16 values -- out[16 f + k], k = 0..6 globaltimer stamps (ns): start, graph
// built, ahCluster done, block membership + region-grow seeds done, region grow done, final merge done, end; k = 7 merge
// steps taken; k = 8..13 SM cycles of ahCluster's sub-phases summed over its steps: queue pop, candidate fits, selection,
// publish + decision, adjacency update, node copy + queue push.
int msl_plane_debug_profile(msl_plane *p, int64_t *out, int frames) {
    if (!p || !out || frames < 1 || !p->d_prof || frames > p->profFrames) return fail(MSL_ERR_INVALID, "msl_plane_debug_profile: bad argument");
    MSL_CUDA(cudaSetDevice(p->device));
    MSL_CUDA(cudaStreamSynchronize(p->stream));
    MSL_CUDA(cudaMemcpy(out, p->d_prof, sizeof(long long) * 16 * (size_t)frames, cudaMemcpyDeviceToHost));
    return MSL_OK;
}

int msl_plane_detect(msl_plane *p, const uint16_t *depth, int dstride_px, size_t frame_stride_px, int batch, const float K[4],
                     float depth_map_factor, int32_t *membership, int32_t *plane_count, msl_plane_rec *planes, int plane_cap) {
    if (!p || !depth || !K || !membership || !plane_count || (plane_cap > 0 && !planes) || plane_cap < 0)
        return fail(MSL_ERR_INVALID, "msl_plane_detect: bad argument");
    if (batch < 1 || batch > p->maxBatch || dstride_px < p->w) return fail(MSL_ERR_INVALID, "msl_plane_detect: bad batch/stride");
    MSL_CUDA(cudaSetDevice(p->device));
    int rc = plane_detect_alloc(p);
    if (rc) return rc;
    const size_t fr = (size_t)p->w * p->h, npix = (size_t)p->W2 * p->H2;
    for (int b = 0; b < batch; b++)
        MSL_CUDA(cudaMemcpy2DAsync(p->d_depth + b * fr, (size_t)p->w * 2, depth + b * frame_stride_px, (size_t)dstride_px * 2,
                                   (size_t)p->w * 2, p->h, cudaMemcpyHostToDevice, p->stream));
    rc = msl_plane_detect_dev(p, p->d_depth, p->w, fr, batch, K, depth_map_factor, p->d_mem, p->d_count, p->d_planes, p->planeCap);
    if (rc) return rc;
    MSL_CUDA(cudaMemcpyAsync(membership, p->d_mem, sizeof(int32_t) * npix * batch, cudaMemcpyDeviceToHost, p->stream));
    MSL_CUDA(cudaMemcpyAsync(plane_count, p->d_count, sizeof(int32_t) * batch, cudaMemcpyDeviceToHost, p->stream));
    std::vector<msl_plane_rec> rec((size_t)batch * p->planeCap);
    MSL_CUDA(cudaMemcpyAsync(rec.data(), p->d_planes, sizeof(msl_plane_rec) * rec.size(), cudaMemcpyDeviceToHost, p->stream));
    MSL_CUDA(cudaStreamSynchronize(p->stream));
    rc = plane_check_frames(p);
    if (rc) return rc;
    for (int b = 0; b < batch; b++) {
        if (plane_count[b] > p->planeCap) return fail(MSL_ERR_CAPACITY, "msl_plane_detect: more planes than the handle's record buffer holds");
        for (int i = 0; i < plane_count[b] && i < plane_cap; i++) planes[(size_t)b * plane_cap + i] = rec[(size_t)b * p->planeCap + i];
    }
    return MSL_OK;
}

int msl_plane_prestage(msl_plane *p, const uint16_t *depth, int dstride_px, size_t frame_stride_px, int batch,
                       const float K[4], float depth_map_factor, double *cloud_xyz, msl_block_stat *blocks,
                       uint8_t *seed, uint8_t *edges) {
    if (!p || !depth || !K) return fail(MSL_ERR_INVALID, "msl_plane_prestage: null argument");
    if (batch < 1 || batch > p->maxBatch || dstride_px < p->w) return fail(MSL_ERR_INVALID, "msl_plane_prestage: bad batch/stride");
    MSL_CUDA(cudaSetDevice(p->device));
    const size_t fr = (size_t)p->w * p->h;
    for (int b = 0; b < batch; b++)
        MSL_CUDA(cudaMemcpy2DAsync(p->d_depth + b * fr, (size_t)p->w * 2, depth + b * frame_stride_px, (size_t)dstride_px * 2,
                                   (size_t)p->w * 2, p->h, cudaMemcpyHostToDevice, p->stream));
    int rc = msl_plane_prestage_dev(p, p->d_depth, p->w, fr, batch, K, depth_map_factor, cloud_xyz ? p->d_cloud : nullptr,
                                    p->d_blocks, p->d_seed, p->d_edges);
    if (rc) return rc;
    const size_t nb = (size_t)p->Nw * p->Nh;
    if (cloud_xyz) MSL_CUDA(cudaMemcpyAsync(cloud_xyz, p->d_cloud, sizeof(double) * 3 * p->W2 * p->H2 * batch, cudaMemcpyDeviceToHost, p->stream));
    if (blocks) MSL_CUDA(cudaMemcpyAsync(blocks, p->d_blocks, sizeof(msl_block_stat) * nb * batch, cudaMemcpyDeviceToHost, p->stream));
    if (seed) MSL_CUDA(cudaMemcpyAsync(seed, p->d_seed, nb * batch, cudaMemcpyDeviceToHost, p->stream));
    if (edges) MSL_CUDA(cudaMemcpyAsync(edges, p->d_edges, nb * batch, cudaMemcpyDeviceToHost, p->stream));
    MSL_CUDA(cudaStreamSynchronize(p->stream));
    return MSL_OK;
}

}  // extern "C"
