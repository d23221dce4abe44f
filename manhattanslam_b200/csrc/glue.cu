// glue.cu -- the frame glue either side of the ORB extractor, so a frame can stay on the device from decode to
// the feature grid (razayunus/ManhattanSLAM):
//   Tracking::GrabImage          src/Tracking.cc:184-211  cvtColor RGB/BGR(A) -> GRAY (15-bit fixed point of OpenCV 4),
//                                                          depth convertTo(CV_32F, mDepthMapFactor)
//   Frame::UndistortKeyPoints    src/Frame.cc:437-463     cv::undistortPoints (5 fixed-point iterations in fp64)
//   Frame::ComputeStereoFromRGBD src/Frame.cc:495-513     depth at the truncated keypoint, uRight = x_un - bf / d
// All kernels are element-parallel and HBM-bound: 128-bit loads where the layout allows, one pass, no reuse.
#include <vector>

#include "msl_common.cuh"

using namespace msl;

namespace {

// ---- cvtColor: 4 pixels per thread when the rows are 4-pixel aligned (12 or 16 source bytes -> one 32-bit store)
__device__ __forceinline__ unsigned gray15(unsigned r, unsigned g, unsigned b) {
    return (r * 9798u + g * 19235u + b * 3735u + (1u << 14)) >> 15;  // RGB2Gray<uchar>, OpenCV >= 4.0
}

__global__ void __launch_bounds__(256)
    k_cvt_gray(const uint8_t *__restrict__ src, int w, int h, int stride, size_t frameStride, int channels, int rgbOrder,
               uint8_t *__restrict__ dst, int dstride, size_t dframe) {
    const int x4 = (blockIdx.x * 256 + threadIdx.x) * 4, y = blockIdx.y, b = blockIdx.z;
    if (x4 >= w) return;
    const uint8_t *s = src + b * frameStride + (size_t)y * stride + (size_t)x4 * channels;
    uint8_t *d = dst + b * dframe + (size_t)y * dstride + x4;
    const int ri = rgbOrder ? 0 : 2, bi = rgbOrder ? 2 : 0;
    const int n = min(4, w - x4);
    unsigned out = 0;
    if (n == 4 && channels == 4 && (((size_t)s) & 15) == 0) {
        const uint4 v = *(const uint4 *)s;
        const unsigned px[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const unsigned c0 = px[q] & 255u, c1 = (px[q] >> 8) & 255u, c2 = (px[q] >> 16) & 255u;
            out |= gray15(rgbOrder ? c0 : c2, c1, rgbOrder ? c2 : c0) << (8 * q);
        }
    } else if (n == 4 && channels == 3 && (((size_t)s) & 3) == 0) {
        const unsigned a = ((const unsigned *)s)[0], bb = ((const unsigned *)s)[1], c = ((const unsigned *)s)[2];
        const unsigned char by[12] = {(unsigned char)a, (unsigned char)(a >> 8), (unsigned char)(a >> 16), (unsigned char)(a >> 24),
                                      (unsigned char)bb, (unsigned char)(bb >> 8), (unsigned char)(bb >> 16), (unsigned char)(bb >> 24),
                                      (unsigned char)c, (unsigned char)(c >> 8), (unsigned char)(c >> 16), (unsigned char)(c >> 24)};
#pragma unroll
        for (int q = 0; q < 4; q++) out |= gray15(by[3 * q + ri], by[3 * q + 1], by[3 * q + bi]) << (8 * q);
    } else {
        for (int q = 0; q < n; q++) out |= gray15(s[q * channels + ri], s[q * channels + 1], s[q * channels + bi]) << (8 * q);
    }
    if (n == 4 && (((size_t)d) & 3) == 0)
        *(unsigned *)d = out;
    else
        for (int q = 0; q < n; q++) d[q] = (uint8_t)(out >> (8 * q));
}

// ---- depth: 8 pixels per thread (one 128-bit load, two 128-bit stores)
__global__ void __launch_bounds__(256) k_depth_to_float(const uint16_t *__restrict__ src, long long n, float factor, float *__restrict__ dst) {
    const long long i8 = ((long long)blockIdx.x * 256 + threadIdx.x) * 8;
    if (i8 >= n) return;
    if (i8 + 8 <= n && (((size_t)(src + i8)) & 15) == 0 && (((size_t)(dst + i8)) & 15) == 0) {
        const uint4 v = __ldcs((const uint4 *)(src + i8));
        const unsigned wv[4] = {v.x, v.y, v.z, v.w};
        float o[8];
#pragma unroll
        for (int q = 0; q < 4; q++) o[2 * q] = (float)(wv[q] & 0xffffu) * factor, o[2 * q + 1] = (float)(wv[q] >> 16) * factor;
        __stcs((float4 *)(dst + i8), make_float4(o[0], o[1], o[2], o[3]));
        __stcs((float4 *)(dst + i8 + 4), make_float4(o[4], o[5], o[6], o[7]));
    } else {
        for (long long i = i8; i < n && i < i8 + 8; i++) dst[i] = (float)src[i] * factor;
    }
}

struct Undist {
    double fx, fy, cx, cy, ifx, ify, k[5];
    int identity;  // mDistCoef.at<float>(0) == 0.0 (src/Frame.cc:438-441)
};

// cv::undistortPoints(src, dst, K, D, noArray(), K): TermCriteria(MAX_ITER, 5, 0.01), all in fp64, same operation order
// as cvUndistortPointsInternal (the k4..k6, s1..s4 terms are present as exact zeros there; they do not change a bit)
__device__ __forceinline__ void undistort_one(const Undist &U, float xin, float yin, float &xo, float &yo) {
    double x = xin, y = yin;
    const double u = x, v = y;
    x = (x - U.cx) * U.ifx;
    y = (y - U.cy) * U.ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = 1.0 / (1 + ((U.k[4] * r2 + U.k[1]) * r2 + U.k[0]) * r2);
        if (icdist < 0) {
            x = (u - U.cx) * U.ifx;
            y = (v - U.cy) * U.ify;
            break;
        }
        const double deltaX = 2 * U.k[2] * x * y + U.k[3] * (r2 + 2 * x * x);
        const double deltaY = U.k[2] * (r2 + 2 * y * y) + 2 * U.k[3] * x * y;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    xo = (float)(U.fx * x + U.cx);
    yo = (float)(U.fy * y + U.cy);
}

// UndistortKeyPoints + ComputeStereoFromRGBD for `batch` frames of ragged keypoint lists (rows per frame reserved,
// counts filled): reads msl_keypoint {x, y, ...} (28 B), writes undistorted xy, uRight, depth.
__global__ void __launch_bounds__(256)
    k_keypoint_glue(const msl_keypoint *__restrict__ kps, int rows, const int32_t *__restrict__ counts, int nFixed, Undist U,
                    const float *__restrict__ depth, int w, int h, float mbf, float *__restrict__ xyUn, float *__restrict__ uRight,
                    float *__restrict__ kDepth) {
    const int i = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
    const int n = counts ? min(counts[b], rows) : nFixed;
    if (i >= n) return;
    const size_t o = (size_t)b * rows + i;
    const float kx = kps[o].x, ky = kps[o].y;
    float ux = kx, uy = ky;
    if (!U.identity) undistort_one(U, kx, ky, ux, uy);
    if (xyUn) xyUn[2 * o] = ux, xyUn[2 * o + 1] = uy;
    if (depth) {
        float ur = -1.f, kd = -1.f;
        const int px = (int)kx, py = (int)ky;  // imDepth.at<float>(v, u): float -> int truncation (:505)
        // keypoints lie inside the image (ORB border 16 px); the guard only keeps a malformed input from faulting
        const float d = (px >= 0 && px < w && py >= 0 && py < h) ? __ldg(depth + (size_t)b * w * h + (size_t)py * w + px) : 0.f;
        if (d > 0) {
            kd = d;
            ur = ux - mbf / d;
        }
        uRight[o] = ur, kDepth[o] = kd;
    }
}

}  // namespace

struct msl_glue {
    int w, h, maxBatch, device;
    cudaStream_t stream = nullptr;
    uint8_t *d_in = nullptr, *d_out = nullptr;  // staging for the host entry points
    size_t inCap = 0, outCap = 0;
    // frame sets (msl_glue_upload_frames): the sensor frames of a batch, uploaded once for every stage
    cudaStream_t upStream = nullptr;
    struct FrameSet {
        uint8_t *gray = nullptr;
        uint16_t *d16 = nullptr;
        float *depth = nullptr;
        int32_t *aux = nullptr;
        size_t auxCap = 0;
        cudaEvent_t ready = nullptr;
        bool used = false;
    } fs[MSL_GLUE_FRAME_SETS];
};

static int glue_reserve(msl_glue *g, size_t in, size_t out) {
    if (in > g->inCap) {
        if (g->d_in) cudaFree(g->d_in);
        g->d_in = nullptr, g->inCap = 0;
        MSL_CUDA(cudaMalloc((void **)&g->d_in, in));
        g->inCap = in;
    }
    if (out > g->outCap) {
        if (g->d_out) cudaFree(g->d_out);
        g->d_out = nullptr, g->outCap = 0;
        MSL_CUDA(cudaMalloc((void **)&g->d_out, out));
        g->outCap = out;
    }
    return MSL_OK;
}

static Undist make_undist(const float K4[4], const float D5[5]) {
    Undist U;
    U.fx = K4[0], U.fy = K4[1], U.cx = K4[2], U.cy = K4[3];
    U.ifx = 1. / U.fx, U.ify = 1. / U.fy;
    for (int q = 0; q < 5; q++) U.k[q] = D5 ? (double)D5[q] : 0.0;
    U.identity = !D5 || D5[0] == 0.0f;
    return U;
}

extern "C" {

int msl_glue_create(int w, int h, int max_batch, int device, msl_glue **out) {
    if (!out) return fail(MSL_ERR_INVALID, "msl_glue_create: null out");
    *out = nullptr;
    if (w < 1 || h < 1 || max_batch < 1) return fail(MSL_ERR_INVALID, "msl_glue_create: parameter out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device || device < 0)
        return fail(MSL_ERR_CUDA, "msl_glue_create: no usable CUDA device (there is no CPU fallback)");
    MSL_CUDA(cudaSetDevice(device));
    msl_glue *g = new msl_glue();
    g->w = w, g->h = h, g->maxBatch = max_batch, g->device = device;
    if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete g;
        return fail(MSL_ERR_CUDA, "msl_glue_create: stream creation failed");
    }
    *out = g;
    return MSL_OK;
}

void msl_glue_destroy(msl_glue *g) {
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->d_in) cudaFree(g->d_in);
    if (g->d_out) cudaFree(g->d_out);
    for (auto &f : g->fs) {
        if (f.gray) cudaFree(f.gray);
        if (f.d16) cudaFree(f.d16);
        if (f.depth) cudaFree(f.depth);
        if (f.aux) cudaFree(f.aux);
        if (f.ready) cudaEventDestroy(f.ready);
    }
    if (g->upStream) cudaStreamDestroy(g->upStream);
    if (g->stream) cudaStreamDestroy(g->stream);
    delete g;
}
void *msl_glue_stream(msl_glue *g) { return g ? (void *)g->stream : nullptr; }
int msl_glue_sync(msl_glue *g) {
    if (!g) return fail(MSL_ERR_INVALID, "null handle");
    MSL_CUDA(cudaSetDevice(g->device));
    MSL_CUDA(cudaStreamSynchronize(g->stream));
    return MSL_OK;
}

int msl_glue_cvt_gray_dev(msl_glue *g, const uint8_t *d_src, int stride, size_t frame_stride, int channels, int rgb_order,
                          int batch, uint8_t *d_gray, int gray_stride, size_t gray_frame_stride) {
    if (!g || !d_src || !d_gray) return fail(MSL_ERR_INVALID, "msl_glue_cvt_gray_dev: null argument");
    if ((channels != 3 && channels != 4) || batch < 1 || stride < g->w * channels || gray_stride < g->w)
        return fail(MSL_ERR_INVALID, "msl_glue_cvt_gray_dev: bad layout");
    MSL_CUDA(cudaSetDevice(g->device));
    k_cvt_gray<<<dim3(cdiv(cdiv(g->w, 4), 256), g->h, batch), 256, 0, g->stream>>>(d_src, g->w, g->h, stride, frame_stride, channels,
                                                                                  rgb_order, d_gray, gray_stride, gray_frame_stride);
    MSL_LAUNCH_CHECK();
    return MSL_OK;
}

int msl_glue_cvt_gray(msl_glue *g, const uint8_t *src, int stride, int channels, int rgb_order, int batch, uint8_t *gray) {
    if (!g || !src || !gray) return fail(MSL_ERR_INVALID, "msl_glue_cvt_gray: null argument");
    if (batch < 1 || batch > g->maxBatch) return fail(MSL_ERR_INVALID, "msl_glue_cvt_gray: bad batch");
    MSL_CUDA(cudaSetDevice(g->device));
    const size_t in = (size_t)stride * g->h * batch, out = (size_t)g->w * g->h * batch;
    int rc = glue_reserve(g, in, out);
    if (rc) return rc;
    MSL_CUDA(cudaMemcpyAsync(g->d_in, src, in, cudaMemcpyHostToDevice, g->stream));
    rc = msl_glue_cvt_gray_dev(g, g->d_in, stride, (size_t)stride * g->h, channels, rgb_order, batch, g->d_out, g->w, (size_t)g->w * g->h);
    if (rc) return rc;
    MSL_CUDA(cudaMemcpyAsync(gray, g->d_out, out, cudaMemcpyDeviceToHost, g->stream));
    MSL_CUDA(cudaStreamSynchronize(g->stream));
    return MSL_OK;
}

int msl_glue_depth_to_float_dev(msl_glue *g, const uint16_t *d_depth16, int64_t n, float factor, float *d_depth) {
    if (!g || !d_depth16 || !d_depth || n < 1) return fail(MSL_ERR_INVALID, "msl_glue_depth_to_float_dev: bad argument");
    MSL_CUDA(cudaSetDevice(g->device));
    k_depth_to_float<<<(unsigned)((n + 2047) / 2048), 256, 0, g->stream>>>(d_depth16, n, factor, d_depth);
    MSL_LAUNCH_CHECK();
    return MSL_OK;
}

int msl_glue_depth_to_float(msl_glue *g, const uint16_t *depth16, int batch, float factor, float *depth) {
    if (!g || !depth16 || !depth) return fail(MSL_ERR_INVALID, "msl_glue_depth_to_float: null argument");
    if (batch < 1 || batch > g->maxBatch) return fail(MSL_ERR_INVALID, "msl_glue_depth_to_float: bad batch");
    MSL_CUDA(cudaSetDevice(g->device));
    const size_t n = (size_t)g->w * g->h * batch;
    int rc = glue_reserve(g, n * 2, n * 4);
    if (rc) return rc;
    MSL_CUDA(cudaMemcpyAsync(g->d_in, depth16, n * 2, cudaMemcpyHostToDevice, g->stream));
    rc = msl_glue_depth_to_float_dev(g, (const uint16_t *)g->d_in, (int64_t)n, factor, (float *)g->d_out);
    if (rc) return rc;
    MSL_CUDA(cudaMemcpyAsync(depth, g->d_out, n * 4, cudaMemcpyDeviceToHost, g->stream));
    MSL_CUDA(cudaStreamSynchronize(g->stream));
    return MSL_OK;
}

// The sensor frames of a batch uploaded ONCE for every stage of the front-end (the reference's Frame constructor hands the same
// mImGray / imDepth to ExtractORB, ComputeStereoFromRGBD, ExtractPlanes and, through the KeyFrame, to SurfelFusion --
// src/Frame.cc:90-110, src/Tracking.cc:184-211): gray and the sensor's CV_16U depth go up on the handle's copy stream, the
// CV_32F depth of Tracking::GrabImageRGBD (imDepth.convertTo(CV_32F, mDepthMapFactor)) is produced on the device.
int msl_glue_upload_frames(msl_glue *g, int slot, const uint8_t *gray, int gray_stride, const uint16_t *depth16, int depth_stride_px,
                           int batch, float factor, const int32_t *aux, size_t aux_ints, const uint8_t **d_gray,
                           const uint16_t **d_depth16, const float **d_depth, const int32_t **d_aux) {
    if (!g || !gray || !depth16 || slot < 0 || slot >= MSL_GLUE_FRAME_SETS) return fail(MSL_ERR_INVALID, "msl_glue_upload_frames: bad argument");
    if (batch < 1 || batch > g->maxBatch) return fail(MSL_ERR_INVALID, "msl_glue_upload_frames: bad batch");
    if (gray_stride < g->w || depth_stride_px < g->w) return fail(MSL_ERR_INVALID, "msl_glue_upload_frames: bad stride");
    MSL_CUDA(cudaSetDevice(g->device));
    msl_glue::FrameSet &f = g->fs[slot];
    const size_t npx = (size_t)g->w * g->h, cap = npx * g->maxBatch;
    if (!g->upStream) MSL_CUDA(cudaStreamCreateWithFlags(&g->upStream, cudaStreamNonBlocking));
    if (!f.gray) {
        MSL_CUDA(cudaMalloc((void **)&f.gray, cap));
        MSL_CUDA(cudaMalloc((void **)&f.d16, cap * 2));
        MSL_CUDA(cudaMalloc((void **)&f.depth, cap * 4));
        MSL_CUDA(cudaEventCreateWithFlags(&f.ready, cudaEventDisableTiming));
    }
    if (aux && aux_ints > f.auxCap) {
        if (f.aux) cudaFree(f.aux);
        f.aux = nullptr, f.auxCap = 0;
        MSL_CUDA(cudaMalloc((void **)&f.aux, aux_ints * 4));
        f.auxCap = aux_ints;
    }
    cudaStream_t st = g->upStream;
    if (gray_stride == g->w) MSL_CUDA(cudaMemcpyAsync(f.gray, gray, npx * batch, cudaMemcpyHostToDevice, st));
    else MSL_CUDA(cudaMemcpy2DAsync(f.gray, g->w, gray, gray_stride, g->w, (size_t)g->h * batch, cudaMemcpyHostToDevice, st));
    if (depth_stride_px == g->w) MSL_CUDA(cudaMemcpyAsync(f.d16, depth16, npx * batch * 2, cudaMemcpyHostToDevice, st));
    else MSL_CUDA(cudaMemcpy2DAsync(f.d16, (size_t)g->w * 2, depth16, (size_t)depth_stride_px * 2, (size_t)g->w * 2, (size_t)g->h * batch, cudaMemcpyHostToDevice, st));
    if (aux && aux_ints) MSL_CUDA(cudaMemcpyAsync(f.aux, aux, aux_ints * 4, cudaMemcpyHostToDevice, st));
    const int64_t n = (int64_t)(npx * batch);
    k_depth_to_float<<<(unsigned)((n + 2047) / 2048), 256, 0, st>>>(f.d16, n, factor, f.depth);
    MSL_LAUNCH_CHECK();
    MSL_CUDA(cudaEventRecord(f.ready, st));
    f.used = true;
    if (d_gray) *d_gray = f.gray;
    if (d_depth16) *d_depth16 = f.d16;
    if (d_depth) *d_depth = f.depth;
    if (d_aux) *d_aux = aux ? f.aux : nullptr;
    return MSL_OK;
}

int msl_glue_frames_wait(msl_glue *g, int slot, void *stream) {
    if (!g || slot < 0 || slot >= MSL_GLUE_FRAME_SETS || !g->fs[slot].used) return fail(MSL_ERR_INVALID, "msl_glue_frames_wait: no upload in this slot");
    MSL_CUDA(cudaSetDevice(g->device));
    if (stream) MSL_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, g->fs[slot].ready, 0));
    else MSL_CUDA(cudaEventSynchronize(g->fs[slot].ready));
    return MSL_OK;
}

int msl_glue_keypoints_dev(msl_glue *g, const msl_keypoint *d_kps, int rows, const int32_t *d_counts, int batch, const float K4[4],
                           const float D5[5], const float *d_depth, float mbf, float *d_xy_un, float *d_uright, float *d_kdepth,
                           void *stream) {
    if (!g || !d_kps || !K4 || rows < 1 || batch < 1) return fail(MSL_ERR_INVALID, "msl_glue_keypoints_dev: bad argument");
    if (d_depth && (!d_uright || !d_kdepth)) return fail(MSL_ERR_INVALID, "msl_glue_keypoints_dev: depth given without outputs");
    MSL_CUDA(cudaSetDevice(g->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : g->stream;
    k_keypoint_glue<<<dim3(cdiv(rows, 256), batch), 256, 0, st>>>(d_kps, rows, d_counts, rows, make_undist(K4, D5), d_depth, g->w, g->h,
                                                                 mbf, d_xy_un, d_uright, d_kdepth);
    MSL_LAUNCH_CHECK();
    return MSL_OK;
}

int msl_glue_keypoints(msl_glue *g, const msl_keypoint *kps, int n, const float K4[4], const float D5[5], const float *depth,
                       float mbf, float *xy_un, float *uright, float *kdepth) {
    if (!g || !K4 || n < 0 || (n && !kps)) return fail(MSL_ERR_INVALID, "msl_glue_keypoints: bad argument");
    if (depth && (!uright || !kdepth)) return fail(MSL_ERR_INVALID, "msl_glue_keypoints: depth given without outputs");
    if (n == 0) return MSL_OK;
    MSL_CUDA(cudaSetDevice(g->device));
    const size_t npx = (size_t)g->w * g->h;
    const size_t inKp = align_up(sizeof(msl_keypoint) * (size_t)n, 256), in = inKp + (depth ? npx * 4 : 0);
    int rc = glue_reserve(g, in, (size_t)n * 16);
    if (rc) return rc;
    MSL_CUDA(cudaMemcpyAsync(g->d_in, kps, sizeof(msl_keypoint) * (size_t)n, cudaMemcpyHostToDevice, g->stream));
    const float *d_depth = nullptr;
    if (depth) {
        MSL_CUDA(cudaMemcpyAsync(g->d_in + inKp, depth, npx * 4, cudaMemcpyHostToDevice, g->stream));
        d_depth = (const float *)(g->d_in + inKp);
    }
    float *o = (float *)g->d_out;
    rc = msl_glue_keypoints_dev(g, (const msl_keypoint *)g->d_in, n, nullptr, 1, K4, D5, d_depth, mbf, xy_un ? o : nullptr, o + 2 * (size_t)n,
                                o + 3 * (size_t)n, nullptr);
    if (rc) return rc;
    if (xy_un) MSL_CUDA(cudaMemcpyAsync(xy_un, o, sizeof(float) * 2 * (size_t)n, cudaMemcpyDeviceToHost, g->stream));
    if (depth) {
        MSL_CUDA(cudaMemcpyAsync(uright, o + 2 * (size_t)n, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, g->stream));
        MSL_CUDA(cudaMemcpyAsync(kdepth, o + 3 * (size_t)n, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, g->stream));
    }
    MSL_CUDA(cudaStreamSynchronize(g->stream));
    return MSL_OK;
}

}  // extern "C"
