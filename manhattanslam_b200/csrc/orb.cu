// orb.cu -- B200-native ORB extraction (replaces src/ORBextractor.cc of razayunus/ManhattanSLAM).
//
// Batched over independent frames; every stage is a hand-written sm_100a kernel:
//   k_resize      O1  8U bilinear pyramid level l from level l-1 (OpenCV fixed-point arithmetic)
//   k_fast_cells  O2  one CTA per 30-px FAST cell: ROI tile -> shared memory, threshold-free FAST-9-16
//                     score, in-cell 3x3 NMS, ini/min threshold fallback, ordered warp-ballot compaction
//   k_octree      O3  one CTA per (frame, level): DistributeOctTree reformulated as rounds of parallel
//                     quad splits with prefix sums (list order, largest-first expansion and the
//                     creation-order tie-break reproduced exactly) + best-response pick
//   k_blur        O6  7x7 sigma-2 fixed-point Gaussian (8.8 taps), separable, shared-memory tiles
//   k_describe    O5+O7+O8  one warp per keypoint: IC-angle (integer moments + fastAtan2 polynomial),
//                     steered 256-bit BRIEF on the blurred level, output scaling
// All integer stages are bit-exact with the CPU oracle; float math avoids FMA contraction (-fmad=false).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include <cuda.h>  // CUtensorMap + cuTensorMapEncodeTiled prototype (resolved through cudaGetDriverEntryPoint, no -lcuda)

#include "msl_common.cuh"

namespace msl {
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};
}  // namespace msl

using namespace msl;

namespace {

constexpr int EDGE_THRESHOLD = 19;  // src/ORBextractor.cc:72
constexpr int PATCH_SIZE = 31;
constexpr int HALF_PATCH = 15;
constexpr int MIN_BORDER = EDGE_THRESHOLD - 3;  // 16
constexpr int MAX_LEVELS = 16;
constexpr int CELL_TILE = 68;  // max ROI edge (wCell < 60, +6 ring, rounded)

struct LevelInfo {
    int w, h, pitch;     // level image size and row pitch (bytes)
    int offset;          // byte offset inside the per-frame pyramid block
    int cellBase, nCells;
    int nFeatures;       // mnFeaturesPerLevel
    int nIni;            // DistributeOctTree root count
    float hX;
    int width, height;   // maxBorder - minBorder
    int tabOff;          // offset (in shorts) of the resize tables of this level
    int xtOff;           // offset (in int2 entries, even) of this level's packed column table (k_resize4)
    int xSpan;           // max over groups of four output columns of (last source column + 1) - (first source column & ~3)
    float scale;         // mvScaleFactor[level]
    int kpBase, kpCap;   // slot range of this level in the per-frame level-keypoint array
    int candBase, candCap;  // slot range of this level in the per-frame candidate arrays (nCells*capCell: cannot overflow)
    int scaledPatch;     // (int)(31 * mvScaleFactor)
};

struct Cell {
    short x0, y0, x1, y1;  // ROI [x0,x1) x [y0,y1) in level coordinates (includes the 3-px FAST ring)
};

struct LevelKp {
    short x, y;      // level coordinates
    short score;     // FAST response
    short level;
};

struct DevPattern {
    int8_t v[1024];
};
__constant__ DevPattern c_pattern = {{
#include "rbrief_pattern_31.inc"
}};
__constant__ int c_umax[16];

inline int cv_round_f(float v) { return (int)nearbyintf(v); }
inline int cv_floor_d(double v) {
    int i = (int)v;
    return i - (i > v);
}
inline int cv_ceil_d(double v) {
    int i = (int)v;
    return i + (i < v);
}
inline short sat_short(float v) {
    int iv = cv_round_f(v);
    return (short)(iv < -32768 ? -32768 : iv > 32767 ? 32767 : iv);
}

// Level 0 = the input frame (src/ORBextractor.cc:886-891), re-pitched into the pyramid block.
__global__ void __launch_bounds__(256) k_load_level0(const uint8_t *__restrict__ src, int stride, size_t frameStride,
                                                     uint8_t *__restrict__ pyr, int pitch, size_t pyrStride, int w,
                                                     int h, int vec) {
    const int y = blockIdx.y, b = blockIdx.z;
    const uint8_t *s = src + b * frameStride + (size_t)y * stride;
    uint8_t *d = pyr + b * pyrStride + (size_t)y * pitch;
    if (vec) {
        const int x = (blockIdx.x * 256 + threadIdx.x) * 16;
        if (x < w) *(uint4 *)(d + x) = *(const uint4 *)(s + x);
    } else {
        for (int x = blockIdx.x * 256 * 16 + threadIdx.x; x < min(w, (int)(blockIdx.x + 1) * 256 * 16); x += 256) d[x] = s[x];
    }
}

// ------------------------------------------------------------------------------------------ O1
// cv::resize INTER_LINEAR 8UC1: D = ((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2 with
// H = S[sx]*a0 + S[sx+1]*a1 (11-bit coefficients).  Tables are built on the host exactly as OpenCV does.
__global__ void __launch_bounds__(256) k_resize(const LevelInfo *__restrict__ lv, int level, uint8_t *pyr,
                                                size_t frameStride, const short *__restrict__ tab) {
    const LevelInfo L = lv[level], P = lv[level - 1];
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= L.w || y >= L.h) return;
    const short *t = tab + L.tabOff;
    const int sx = t[x], a0 = t[L.w + x], a1 = t[2 * L.w + x];
    const short *ty = t + 3 * L.w;
    const int sy = ty[y], b0 = ty[L.h + y], b1 = ty[2 * L.h + y];
    const int sx1 = min(sx + 1, P.w - 1);
    const int sy0 = min(max(sy, 0), P.h - 1), sy1 = min(max(sy + 1, 0), P.h - 1);
    const uint8_t *src = pyr + blockIdx.z * frameStride + P.offset;
    const uint8_t *r0 = src + (size_t)sy0 * P.pitch, *r1 = src + (size_t)sy1 * P.pitch;
    const int h0 = r0[sx] * a0 + r0[sx1] * a1;
    const int h1 = r1[sx] * a0 + r1[sx1] * a1;
    const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    pyr[blockIdx.z * frameStride + L.offset + (size_t)y * L.pitch + x] = (uint8_t)v;
}

// The same, four output pixels of a row per thread (default; MSL_ORB_RESIZE4=0 for the one-pixel form).  k_resize is a
// few loads per thread behind two dependent round trips, in ~50 k tiny CTAs per level and batch: it runs at the latency of
// a CTA times the number of waves, not at any throughput.  Here a thread's four columns come from two 16-byte loads of a
// packed column table {sx | sx1 << 16, a0 | a1 << 16}, and the source bytes of both rows from three aligned words each
// (four output columns span at most 12 source bytes from the word the first one lies in -- xSpan, checked on the host);
// the arithmetic per pixel is k_resize's.
__device__ __forceinline__ unsigned resize_pick(unsigned w0, unsigned w1, unsigned w2, int o) {  // byte o (0..11) of {w0, w1, w2}
    const unsigned w = o < 8 ? (o < 4 ? w0 : w1) : w2;
    return (w >> (8 * (o & 3))) & 255u;
}
__global__ void __launch_bounds__(256) k_resize4(const LevelInfo *__restrict__ lv, int level, uint8_t *pyr, size_t frameStride,
                                                 const short *__restrict__ tab, const int4 *__restrict__ xt4) {
    const LevelInfo L = lv[level], P = lv[level - 1];
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4, y = blockIdx.y * 8 + threadIdx.y;
    if (x0 >= L.w || y >= L.h) return;
    const int4 e01 = __ldg(xt4 + (L.xtOff + x0) / 2), e23 = __ldg(xt4 + (L.xtOff + x0) / 2 + 1);  // (the table is padded to four columns)
    const int ex[4] = {e01.x, e01.z, e23.x, e23.z}, ea[4] = {e01.y, e01.w, e23.y, e23.w};
    const short *ty = tab + L.tabOff + 3 * L.w;
    const int sy = ty[y], b0 = ty[L.h + y], b1 = ty[2 * L.h + y];
    const int sy0 = min(max(sy, 0), P.h - 1), sy1 = min(max(sy + 1, 0), P.h - 1);
    const uint8_t *src = pyr + blockIdx.z * frameStride + P.offset;
    const int base = (ex[0] & 0xffff) & ~3;
    const unsigned *r0 = reinterpret_cast<const unsigned *>(src + (size_t)sy0 * P.pitch + base);
    const unsigned *r1 = reinterpret_cast<const unsigned *>(src + (size_t)sy1 * P.pitch + base);
    const unsigned u0 = r0[0], u1 = r0[1], u2 = r0[2], v0 = r1[0], v1 = r1[1], v2 = r1[2];
    unsigned out = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int o0 = (ex[q] & 0xffff) - base, o1 = (int)((unsigned)ex[q] >> 16) - base;
        const int a0 = ea[q] & 0xffff, a1 = (int)((unsigned)ea[q] >> 16);
        const int h0 = (int)resize_pick(u0, u1, u2, o0) * a0 + (int)resize_pick(u0, u1, u2, o1) * a1;
        const int h1 = (int)resize_pick(v0, v1, v2, o0) * a0 + (int)resize_pick(v0, v1, v2, o1) * a1;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        out |= (unsigned)(v & 255) << (8 * q);
    }
    uint8_t *dst = pyr + blockIdx.z * frameStride + L.offset + (size_t)y * L.pitch + x0;
    if (x0 + 3 < L.w) {
        *reinterpret_cast<unsigned *>(dst) = out;  // rows are 16-byte aligned and padded: an aligned word inside the row
    } else {
        for (int q = 0; x0 + q < L.w; q++) dst[q] = (uint8_t)(out >> (8 * q));
    }
}

// ------------------------------------------------------------------------------------------ O2
// Threshold-free FAST-9-16 score: S_max = max over the 16 nine-pixel arcs, both polarities, of the
// minimum signed difference; cv::FAST's response is S_max-1 and "corner at t" <=> S_max > t.
// minTh: scores <= minTh are never used (a candidate needs S_max > minTh and NMS only ever compares a candidate
// against neighbours whose true score is below its own when they are <= minTh), so they are returned as 0 after a
// cheap test: any 9-arc of the 16-ring contains two of the four compass pixels (0, 4, 8, 12) that are 4 apart or
// three of them, so S_max > minTh needs >= 2 compass pixels beyond the threshold on the same side.
__device__ __forceinline__ bool fast_quick(const uint8_t *c, int pitch, int minTh) {
    const int v = c[0];
    const int c0 = v - c[3 * pitch], c4 = v - c[3], c8 = v - c[-3 * pitch], c12 = v - c[-3];
    const int pos = (c0 > minTh) + (c4 > minTh) + (c8 > minTh) + (c12 > minTh);
    const int neg = (c0 < -minTh) + (c4 < -minTh) + (c8 < -minTh) + (c12 < -minTh);
    return pos >= 2 || neg >= 2;
}
// the full score of a pixel that passed fast_quick (the two are separate passes in k_fast_cells: about one pixel in ten
// passes the compass test on natural texture, but six warps in ten hold at least one that does -- the survivors are
// compacted into a list first so that the 200-instruction arc reduction runs on full warps)
__device__ __forceinline__ int fast_smax(const uint8_t *c, int pitch) {
    const int v = c[0];
    int d[16];
    d[0] = v - c[3 * pitch];
    d[1] = v - c[3 * pitch + 1];
    d[2] = v - c[2 * pitch + 2];
    d[3] = v - c[pitch + 3];
    d[4] = v - c[3];
    d[5] = v - c[-pitch + 3];
    d[6] = v - c[-2 * pitch + 2];
    d[7] = v - c[-3 * pitch + 1];
    d[8] = v - c[-3 * pitch];
    d[9] = v - c[-3 * pitch - 1];
    d[10] = v - c[-2 * pitch - 2];
    d[11] = v - c[-pitch - 3];
    d[12] = v - c[-3];
    d[13] = v - c[pitch - 3];
    d[14] = v - c[2 * pitch - 2];
    d[15] = v - c[3 * pitch - 1];
    int lo2[16], hi2[16], lo4[16], hi4[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        lo2[i] = min(d[i], d[(i + 1) & 15]);
        hi2[i] = max(d[i], d[(i + 1) & 15]);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) {
        lo4[i] = min(lo2[i], lo2[(i + 2) & 15]);
        hi4[i] = max(hi2[i], hi2[(i + 2) & 15]);
    }
    // NOTE: the two polarities are reduced separately and combined once at the end.  Folding them as
    // max(best, max(lo9, -hi9)) inside the loop is miscompiled by ptxas 12.9 for sm_100a (a negated
    // operand feeding a fused 3-input VIMNMX3; repro in tools/ptxas_vimnmx3_repro.cu).
    int a = lo4[0], b = hi4[0];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int lo9 = min(min(lo4[i], lo4[(i + 4) & 15]), d[(i + 8) & 15]);
        const int hi9 = max(max(hi4[i], hi4[(i + 4) & 15]), d[(i + 8) & 15]);
        a = (i == 0) ? lo9 : max(a, lo9);
        b = (i == 0) ? hi9 : min(b, hi9);
    }
    const int dark = 0 - b;
    int best = a > dark ? a : dark;
    best = best > 0 ? best : 0;
    return best;
}

// One CTA per FAST cell (src/ORBextractor.cc:745-780).  Output: ordered survivor records of the cell
// (x | y<<12 | smax<<24, x/y relative to minBorder) into its staging slot + the cell count.
// ---- TMA (tensor-map) helpers for the FAST cell tiles: SASS UTMALDG + SYNCS
__device__ __forceinline__ uint32_t orb_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool orb_mbar_wait_bounded(uint64_t *b, uint32_t parity) {
    // bounded spin: a malformed descriptor must surface as an error code, never as a hung GPU
    for (int it = 0; it < (1 << 20); it++) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(orb_smem_u32(b)), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

// TMA tiled loads trap ("illegal instruction") unless innermost-coordinate * elemSize is 16-byte aligned (measured on
// B200, driver 580): the box starts at x0 & ~15 and is wide enough for ROI width (<= 66) + 15, rounded up to 16.
constexpr int TILE_W_MAX = 96;

// One CTA per FAST cell (src/ORBextractor.cc:745-780).  The cell's ROI (cell + 3-px FAST ring) is staged into
// shared memory by ONE 3-D TMA tensor load (x, y, frame) of the level image; out-of-image columns/rows are
// zero-filled by the TMA unit and never read.  Output: ordered survivor records of the cell
// (x | y<<12 | smax<<24, x/y relative to minBorder) into its staging slot + the cell count.
__global__ void __launch_bounds__(128)
    k_fast_cells(const LevelInfo *__restrict__ lv, const Cell *__restrict__ cells, const short *__restrict__ cellLevel,
                 const uint8_t *__restrict__ pyr, size_t frameStride, const CUtensorMap *__restrict__ tmaps, int tileW,
                 int tileBytes, int totalCells, int capCell, int iniTh, int minTh, uint32_t *__restrict__ staging,
                 int *__restrict__ cellCount, int *__restrict__ err) {
    __shared__ __align__(128) uint8_t tile[TILE_W_MAX * CELL_TILE];
    __shared__ uint8_t sc[CELL_TILE * CELL_TILE];
    __shared__ uint8_t sv[CELL_TILE * CELL_TILE];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint16_t cand[CELL_TILE * CELL_TILE];  // pixels that passed the compass test (any order)
    __shared__ int wsum[4];
    __shared__ int s_any, s_ncand;
    const int cell = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x;
    const Cell C = cells[cell];
    const int level = cellLevel[cell];
    const LevelInfo L = lv[level];
    const int rw = C.x1 - C.x0, rh = C.y1 - C.y0;
    const int cw = rw - 6, ch = rh - 6;  // detectable area
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(orb_smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(orb_smem_u32(&mbar)), "r"(tileBytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                orb_smem_u32(tile)),
            "l"((unsigned long long)(tmaps + level)), "r"((int)C.x0 & ~15), "r"((int)C.y0), "r"(frame), "r"(orb_smem_u32(&mbar))
            : "memory");
    }
    __syncthreads();
    if (!orb_mbar_wait_bounded(&mbar, 0)) {  // should never happen; keep the result correct and report it
        if (tid == 0) atomicExch(err, 4);
        const uint8_t *img = pyr + frame * frameStride + L.offset;
        for (int p = tid; p < rw * rh; p += 128) {
            int ry = p / rw, rx = p - ry * rw;
            tile[ry * tileW + (C.x0 & 15) + rx] = img[(size_t)(C.y0 + ry) * L.pitch + C.x0 + rx];
        }
    }
    const uint8_t *roi = tile + (C.x0 & 15);  // ROI origin inside the 16-byte aligned box
    if (tid == 0) s_any = 0, s_ncand = 0;
    __syncthreads();
    const int np = cw * ch;
    const int lane = tid & 31, wid = tid >> 5;
    // pixel p = tid, tid + 128, ... of the cell in row-major order: (cy, cx) advanced by a fixed step instead of divided out
    // of p in each of the three passes (the divisions were a fifth of the kernel's instructions)
    const int cwd = max(cw, 1);  // (an empty cell has np <= 0: no pass runs)
    const int stepY = 128 / cwd, stepX = 128 - stepY * cwd;
    const int cy0 = tid / cwd, cx0 = tid - cy0 * cwd;
#define FAST_ADVANCE(cy, cx) \
    do {                     \
        cx += stepX;         \
        cy += stepY;         \
        if (cx >= cw) {      \
            cx -= cw;        \
            cy++;            \
        }                    \
    } while (0)
    // scores with a one-pixel border of zeros (a neighbour outside the cell counts as 0 in the NMS): no bounds tests
    uint8_t *const scI = sc + CELL_TILE + 1;  // score of (cy, cx) at scI[cy * CELL_TILE + cx]
    for (int i = tid; i < 2 * (cw + 2) + 2 * ch; i += 128) {
        int by, bx;
        if (i < cw + 2) by = -1, bx = i - 1;
        else if (i < 2 * (cw + 2)) by = ch, bx = i - (cw + 2) - 1;
        else if (i < 2 * (cw + 2) + ch) by = i - 2 * (cw + 2), bx = -1;
        else by = i - 2 * (cw + 2) - ch, bx = cw;
        scI[by * CELL_TILE + bx] = 0;
    }
    {
        const int thq = min(minTh, iniTh);
        int cy = cy0, cx = cx0;
        for (int p0 = 0; p0 < np; p0 += 128) {
            const int p = p0 + tid;
            bool pass = false;
            int q = 0;
            if (p < np) {
                q = cy * CELL_TILE + cx;
                pass = fast_quick(roi + (cy + 3) * tileW + cx + 3, tileW, thq);
                if (!pass) scI[q] = 0;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            int base = 0;
            if (lane == 0 && bal) base = atomicAdd(&s_ncand, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) cand[base + __popc(bal & ((1u << lane) - 1))] = (uint16_t)q;
            FAST_ADVANCE(cy, cx);
        }
        __syncthreads();
        const int nc = s_ncand;
        for (int i = tid; i < nc; i += 128) {
            const int q = cand[i];
            const int cy = q / CELL_TILE, cx = q - cy * CELL_TILE;
            scI[q] = (uint8_t)fast_smax(roi + (cy + 3) * tileW + cx + 3, tileW);
        }
    }
    __syncthreads();
    // in-cell NMS: strict maximum over the 8 neighbours; neighbours outside the cell's ring count as 0
    int any = 0;
    {
        int cy = cy0, cx = cx0;
        for (int p = tid; p < np; p += 128) {
            const uint8_t *c = scI + cy * CELL_TILE + cx;
            const int s = c[0];
            int keep = 0;
            if (s > minTh) {
                const int m0 = max(max((int)c[-CELL_TILE - 1], (int)c[-CELL_TILE]), max((int)c[-CELL_TILE + 1], (int)c[-1]));
                const int m1 = max(max((int)c[1], (int)c[CELL_TILE - 1]), max((int)c[CELL_TILE], (int)c[CELL_TILE + 1]));
                keep = s > max(m0, m1);
            }
            sv[cy * CELL_TILE + cx] = keep ? (uint8_t)s : 0;
            any |= (keep && s > iniTh);
            FAST_ADVANCE(cy, cx);
        }
    }
    if (any) s_any = 1;
    __syncthreads();
    const int thr = s_any ? iniTh : minTh;  // FAST(iniThFAST) non-empty, else FAST(minThFAST)
    uint32_t *out = staging + ((size_t)frame * totalCells + cell) * capCell;
    const int offX = C.x0 + 3 - MIN_BORDER, offY = C.y0 + 3 - MIN_BORDER;
    int base = 0;
    {
        int cy = cy0, cx = cx0;
        for (int p0 = 0; p0 < np; p0 += 128) {
            const int p = p0 + tid;
            const int s = p < np ? (int)sv[cy * CELL_TILE + cx] : 0;
            const bool f = s > thr;
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (lane == 0) wsum[wid] = __popc(bal);
            __syncthreads();
            int pre = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < 4; w++) {
                int c = wsum[w];
                if (w < wid) pre += c;
                tot += c;
            }
            if (f) out[base + pre + __popc(bal & ((1u << lane) - 1))] =
                       (uint32_t)(cx + offX) | ((uint32_t)(cy + offY) << 12) | ((uint32_t)s << 24);
            base += tot;
            FAST_ADVANCE(cy, cx);
            __syncthreads();
        }
    }
#undef FAST_ADVANCE
    if (tid == 0) cellCount[(size_t)frame * totalCells + cell] = base;
}

// ------------------------------------------------------------------------------------------ O3
// DistributeOctTree (src/ORBextractor.cc:531-721) as rounds of parallel quad splits.  The node list is
// kept in list order in shared memory and rebuilt every round:
//   new list = reverse(children in creation order) ++ (undivided nodes in old order)
// which is exactly what push_front/erase produce.  Phase 1 divides every node with >1 keys; once
// size + 3*nToExpand > N the largest-first phase divides the previous round's multi-key children
// in (size desc, creation desc) order -- creation desc == list position asc -- up to the node that
// makes size >= N.  Finally each node keeps its max-response key (first in candidate order on ties).
struct OctSmem {
    short4 *rect[2];
    int *cnt[2];
    int *pidx;     // eligible enumeration (exclusive scan)
    int *rank;     // creation order index of eligible node (by pidx)
    int *ordPos;   // list position by creation order index
    int *cc;       // 4 child counts per eligible node (by pidx); later reused as child new-position
    int *flag;     // scan scratch (4 per order index)
    int *surv;     // survivor scan
    int *ws;       // block scan scratch (34 ints)
};

__device__ __forceinline__ int quadrant_of(const short4 r, int x, int y) {
    // ExtractorNode::DivideNode, src/ORBextractor.cc:477-529
    const int halfX = (int)ceilf((float)(r.z - r.x) / 2.f);
    const int halfY = (int)ceilf((float)(r.w - r.y) / 2.f);
    const int q = (x < r.x + halfX) ? 0 : 1;
    return (y < r.y + halfY) ? q : q + 2;
}

__device__ __forceinline__ short4 child_rect(const short4 r, int q) {
    const int halfX = (int)ceilf((float)(r.z - r.x) / 2.f);
    const int halfY = (int)ceilf((float)(r.w - r.y) / 2.f);
    short4 c;
    c.x = (q & 1) ? r.x + halfX : r.x;
    c.z = (q & 1) ? r.z : r.x + halfX;
    c.y = (q & 2) ? r.y + halfY : r.y;
    c.w = (q & 2) ? r.w : r.y + halfY;
    return c;
}

__global__ void __launch_bounds__(256)
    k_octree(const LevelInfo *__restrict__ lv, int nlevels, int totalCells, int capCell,
             const uint32_t *__restrict__ staging, const int *__restrict__ cellCount, uint32_t *__restrict__ candRec,
             unsigned short *__restrict__ candNode, int *__restrict__ candCount, LevelKp *__restrict__ lvlKps,
             int *__restrict__ lvlCount, int kpCapTotal, int candCapTotal, int maxNodes, int *__restrict__ err) {
    extern __shared__ __align__(16) int smem_raw[];
    const int level = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    const LevelInfo L = lv[level];
    OctSmem S;
    {
        int *p = smem_raw;
        S.rect[0] = (short4 *)p; p += 2 * maxNodes;
        S.rect[1] = (short4 *)p; p += 2 * maxNodes;
        S.cnt[0] = p; p += maxNodes;
        S.cnt[1] = p; p += maxNodes;
        S.pidx = p; p += maxNodes;
        S.rank = p; p += maxNodes;
        S.ordPos = p; p += maxNodes;
        S.surv = p; p += maxNodes;
        S.cc = p; p += 4 * maxNodes;
        S.flag = p; p += 4 * maxNodes;
        S.ws = p; p += 40;
    }
    __shared__ int s_size, s_phase, s_clast, s_finish, s_J, s_m;
    uint32_t *rec = candRec + (size_t)frame * candCapTotal + L.candBase;
    unsigned short *node = candNode + (size_t)frame * candCapTotal + L.candBase;

    // ---- gather the ordered candidate list: cells in (row, col) order, row-major inside a cell
    const int *cc = cellCount + (size_t)frame * totalCells + L.cellBase;
    int *cellOff = S.flag;  // nCells <= 4*maxNodes is checked on the host
    for (int c = tid; c < L.nCells; c += nt) cellOff[c] = cc[c];
    __syncthreads();
    const int n = block_excl_scan(cellOff, L.nCells, S.ws);
    if (tid == 0) candCount[frame * nlevels + level] = n;
    __syncthreads();
    if (n == 0) {
        if (tid == 0) lvlCount[frame * nlevels + level] = 0;
        return;
    }
    for (int c = tid >> 5; c < L.nCells; c += nt >> 5) {  // one warp per cell
        const uint32_t *src = staging + ((size_t)frame * totalCells + L.cellBase + c) * capCell;
        const int k = cc[c], o = cellOff[c];
        for (int i = tid & 31; i < k; i += 32) rec[o + i] = src[i];
    }
    // ---- roots (src/ORBextractor.cc:536-571)
    const int N = L.nFeatures;
    for (int i = tid; i < L.nIni; i += nt) {
        short4 r;
        r.x = (short)(int)(L.hX * (float)i);
        r.z = (short)(int)(L.hX * (float)(i + 1));
        r.y = 0;
        r.w = (short)L.height;
        S.rect[0][i] = r;
        S.cnt[0][i] = 0;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        const int x = rec[i] & 0xfff;
        const int r = (int)((float)x / L.hX);
        node[i] = (unsigned short)r;
        atomicAdd(&S.cnt[0][r], 1);
    }
    __syncthreads();
    int cur = 0;
    {   // erase empty roots, keep order
        for (int i = tid; i < L.nIni; i += nt) S.surv[i] = S.cnt[0][i] > 0;
        __syncthreads();
        const int live = block_excl_scan(S.surv, L.nIni, S.ws);
        for (int i = tid; i < L.nIni; i += nt)
            if (S.cnt[0][i] > 0) {
                S.rect[1][S.surv[i]] = S.rect[0][i];
                S.cnt[1][S.surv[i]] = S.cnt[0][i];
                S.pidx[i] = S.surv[i];
            } else
                S.pidx[i] = -1;
        __syncthreads();
        if (live != L.nIni)
            for (int i = tid; i < n; i += nt) node[i] = (unsigned short)S.pidx[node[i]];
        cur = 1;
        if (tid == 0) {
            s_size = live;
            s_phase = 1;
            s_clast = 0;
            s_finish = 0;
        }
        __syncthreads();
    }

    // ---- split rounds
    for (int round = 0; round < 64; round++) {
        const short4 *rect = S.rect[cur];
        const int *cnt = S.cnt[cur];
        short4 *rect2 = S.rect[cur ^ 1];
        int *cnt2 = S.cnt[cur ^ 1];
        const int size = s_size, phase = s_phase, clast = s_clast;
        // 1. eligible nodes, enumerated by list position
        for (int p = tid; p < size; p += nt) S.pidx[p] = (cnt[p] > 1) && (phase == 1 || p < clast);
        __syncthreads();
        const int m = block_excl_scan(S.pidx, size, S.ws);
        if (m == 0) break;  // size == prevSize -> bFinish
        for (int k = tid; k < 4 * m; k += nt) S.cc[k] = 0;
        __syncthreads();
        auto eligible = [&](int p) { return (cnt[p] > 1) && (phase == 1 || p < clast); };
        // 2. child occupancy of every eligible node
        for (int i = tid; i < n; i += nt) {
            const int p = node[i];
            if (eligible(p)) {
                const uint32_t r = rec[i];
                atomicAdd(&S.cc[4 * S.pidx[p] + quadrant_of(rect[p], r & 0xfff, (r >> 12) & 0xfff)], 1);
            }
        }
        __syncthreads();
        // 3. creation order: list order (phase 1) or size desc / position asc (phase 2)
        for (int p = tid; p < size; p += nt) {
            if (!eligible(p)) continue;
            int rk;
            if (phase == 1)
                rk = S.pidx[p];
            else {
                rk = 0;
                const int c = cnt[p];
                for (int p2 = 0; p2 < clast; p2++) {
                    const int c2 = cnt[p2];
                    rk += (c2 > 1) && (c2 > c || (c2 == c && p2 < p));
                }
            }
            S.rank[S.pidx[p]] = rk;
            S.ordPos[rk] = p;
        }
        __syncthreads();
        // 4. cut-off: divide in creation order until size >= N (phase 2 only)
        int J = m - 1;
        if (phase == 2) {
            for (int e = tid; e < m; e += nt) {
                const int k = S.pidx[S.ordPos[e]];
                S.flag[e] = (S.cc[4 * k] > 0) + (S.cc[4 * k + 1] > 0) + (S.cc[4 * k + 2] > 0) + (S.cc[4 * k + 3] > 0) - 1;
            }
            __syncthreads();
            block_excl_scan(S.flag, m, S.ws);  // flag[e] = growth before e
            if (tid == 0) s_J = m - 1;
            __syncthreads();
            for (int e = tid; e < m; e += nt) {
                const int k = S.pidx[S.ordPos[e]];
                const int nz = (S.cc[4 * k] > 0) + (S.cc[4 * k + 1] > 0) + (S.cc[4 * k + 2] > 0) + (S.cc[4 * k + 3] > 0);
                const int before = size + S.flag[e], after = before + nz - 1;
                if (before < N && after >= N) s_J = e;  // unique: first e reaching N
            }
            __syncthreads();
            J = s_J;
            __syncthreads();
        }
        // 5. creation index of every child of the divided nodes
        for (int k = tid; k < 4 * (J + 1); k += nt) {
            const int e = k >> 2, q = k & 3;
            S.flag[k] = S.cc[4 * S.pidx[S.ordPos[e]] + q] > 0;
        }
        __syncthreads();
        const int created = block_excl_scan(S.flag, 4 * (J + 1), S.ws);
        // 6. survivors keep their relative order behind the new children
        for (int p = tid; p < size; p += nt) S.surv[p] = !(eligible(p) && S.rank[S.pidx[p]] <= J);
        __syncthreads();
        const int nsurv = block_excl_scan(S.surv, size, S.ws);
        const int newSize = created + nsurv;
        // 7. build the new list (position = created-1-creationIdx for children)
        for (int k = tid; k < 4 * (J + 1); k += nt) {
            const int e = k >> 2, q = k & 3;
            const int p = S.ordPos[e];
            const int c = S.cc[4 * S.pidx[p] + q];
            if (c > 0) {
                const int np = created - 1 - S.flag[k];
                rect2[np] = child_rect(rect[p], q);
                cnt2[np] = c;
            }
        }
        for (int p = tid; p < size; p += nt)
            if (!(eligible(p) && S.rank[S.pidx[p]] <= J)) {
                const int np = created + S.surv[p];
                rect2[np] = rect[p];
                cnt2[np] = cnt[p];
            }
        // 8. move the keys
        for (int i = tid; i < n; i += nt) {
            const int p = node[i];
            int np;
            if (eligible(p) && S.rank[S.pidx[p]] <= J) {
                const uint32_t r = rec[i];
                const int q = quadrant_of(rect[p], r & 0xfff, (r >> 12) & 0xfff);
                np = created - 1 - S.flag[4 * S.rank[S.pidx[p]] + q];
            } else
                np = created + S.surv[p];
            node[i] = (unsigned short)np;
        }
        __syncthreads();
        // 9. termination / phase switch (src/ORBextractor.cc:641-700)
        if (tid == 0) s_m = 0;
        __syncthreads();
        int nexp = 0;
        for (int p = tid; p < created; p += nt) nexp += cnt2[p] > 1;
        if (nexp) atomicAdd(&s_m, nexp);
        __syncthreads();
        if (tid == 0) {
            const int nToExpand = s_m;
            if (newSize >= N || newSize == size)
                s_finish = 1;
            else if (phase == 1 && newSize + nToExpand * 3 > N)
                s_phase = 2;
            s_size = newSize;
            s_clast = created;
        }
        cur ^= 1;
        __syncthreads();
        if (s_finish) break;
    }
    __syncthreads();
    // ---- best key per node: max response, first in candidate order on ties (src/ORBextractor.cc:702-718)
    const int size = s_size;
    unsigned long long *best = (unsigned long long *)S.cc;  // 8-byte aligned: offset is a multiple of 2 ints
    for (int p = tid; p < size; p += nt) best[p] = 0ull;
    __syncthreads();
    for (int i = tid; i < n; i += nt)
        atomicMax(&best[node[i]], ((unsigned long long)(rec[i] >> 24) << 32) | (0xffffffffu - (unsigned)i));
    __syncthreads();
    if (size > L.kpCap) {
        if (tid == 0) {
            atomicExch(err, 2);
            lvlCount[frame * nlevels + level] = 0;
        }
        return;
    }
    LevelKp *out = lvlKps + (size_t)frame * kpCapTotal + L.kpBase;
    for (int p = tid; p < size; p += nt) {
        const uint32_t r = rec[0xffffffffu - (unsigned)(best[p] & 0xffffffffull)];
        LevelKp k;
        k.x = (short)((r & 0xfff) + MIN_BORDER);
        k.y = (short)(((r >> 12) & 0xfff) + MIN_BORDER);
        k.score = (short)((r >> 24) - 1);
        k.level = (short)level;
        out[p] = k;
    }
    if (tid == 0) lvlCount[frame * nlevels + level] = size;
}

// ------------------------------------------------------------------------------------------ O6
// cv::GaussianBlur 7x7 sigma 2, BORDER_REFLECT_101, 8.8 fixed-point taps {18,34,48,56,48,34,18}.
constexpr int BLUR_TW = 64, BLUR_TH = 16;
__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = (p < 0) ? -p : 2 * (len - 1) - p;
    return p;
}
// Each thread produces FOUR horizontally adjacent outputs per pass (byte loads, one division per element and 16-bit
// shared-memory traffic per pixel made the per-pixel form issue-bound at ~170 instructions per pixel): interior tiles are
// staged with aligned 32-bit loads (rows are padded to 16 bytes and tiles start at multiples of 64), the horizontal pass
// reads three words and writes four 16-bit sums with one 64-bit store, the vertical pass reads seven 64-bit words and
// writes four bytes with one store.  Border tiles take the per-byte path with BORDER_REFLECT_101.
__global__ void __launch_bounds__(256)
    k_blur(const LevelInfo *__restrict__ lv, const int *__restrict__ tileBase, int nlevels,
           const uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur, size_t frameStride) {
    constexpr int INW = (BLUR_TW + 8) / 4;                 // words per staged row: columns tx - 4 .. tx + 67
    __shared__ __align__(16) uint32_t in32[(BLUR_TH + 6) * INW];
    __shared__ __align__(16) uint16_t hz[(BLUR_TH + 6)][BLUR_TW];
    int level = 0;
    while (level + 1 < nlevels && (int)blockIdx.x >= tileBase[level + 1]) level++;
    const LevelInfo L = lv[level];
    const int t = blockIdx.x - tileBase[level];
    const int tilesX = (L.w + BLUR_TW - 1) / BLUR_TW;
    const int tx = (t % tilesX) * BLUR_TW, ty = (t / tilesX) * BLUR_TH;
    const uint8_t *src = pyr + blockIdx.y * frameStride + L.offset;
    const int tid = threadIdx.x;
    const bool interior = tx >= 4 && tx + BLUR_TW + 4 <= L.w && ty >= 3 && ty + BLUR_TH + 3 <= L.h;
    if (interior) {
        for (int i = tid; i < (BLUR_TH + 6) * INW; i += 256) {
            const int ry = i / INW, w = i - ry * INW;
            in32[i] = *reinterpret_cast<const uint32_t *>(src + (size_t)(ty - 3 + ry) * L.pitch + tx - 4 + 4 * w);
        }
    } else {
        uint8_t *in8 = reinterpret_cast<uint8_t *>(in32);
        for (int p = tid; p < (BLUR_TH + 6) * (BLUR_TW + 6); p += 256) {
            const int ry = p / (BLUR_TW + 6), rx = p - ry * (BLUR_TW + 6);
            const int gy = reflect101(ty + ry - 3, L.h), gx = reflect101(tx + rx - 3, L.w);
            in8[ry * (INW * 4) + rx + 1] = src[(size_t)gy * L.pitch + gx];
        }
    }
    __syncthreads();
    for (int i = tid; i < (BLUR_TH + 6) * (BLUR_TW / 4); i += 256) {
        const int ry = i / (BLUR_TW / 4), g = i - ry * (BLUR_TW / 4);
        const uint32_t *row = in32 + ry * INW + g;
        const uint32_t w0 = row[0], w1 = row[1], w2 = row[2];
        // bytes 1 .. 10 of the three words: columns x - 3 .. x + 6 for the outputs x .. x + 3
        const uint32_t b1 = (w0 >> 8) & 0xff, b2 = (w0 >> 16) & 0xff, b3 = w0 >> 24;
        const uint32_t b4 = w1 & 0xff, b5 = (w1 >> 8) & 0xff, b6 = (w1 >> 16) & 0xff, b7 = w1 >> 24;
        const uint32_t b8 = w2 & 0xff, b9 = (w2 >> 8) & 0xff, b10 = (w2 >> 16) & 0xff;
        const uint32_t h0 = 18 * (b1 + b7) + 34 * (b2 + b6) + 48 * (b3 + b5) + 56 * b4;
        const uint32_t h1 = 18 * (b2 + b8) + 34 * (b3 + b7) + 48 * (b4 + b6) + 56 * b5;
        const uint32_t h2 = 18 * (b3 + b9) + 34 * (b4 + b8) + 48 * (b5 + b7) + 56 * b6;
        const uint32_t h3 = 18 * (b4 + b10) + 34 * (b5 + b9) + 48 * (b6 + b8) + 56 * b7;
        *reinterpret_cast<uint2 *>(&hz[ry][4 * g]) = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));  // each <= 65280
    }
    __syncthreads();
    uint8_t *dst = blur + blockIdx.y * frameStride + L.offset;
    {
        const int ry = tid >> 4, g = tid & 15;  // 16 rows x 16 groups of four columns
        const int gx = tx + 4 * g, gy = ty + ry;
        if (gx < L.w && gy < L.h) {
            uint2 r[7];
#pragma unroll
            for (int k = 0; k < 7; k++) r[k] = *reinterpret_cast<const uint2 *>(&hz[ry + k][4 * g]);
            uint32_t o[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t v[7];
#pragma unroll
                for (int k = 0; k < 7; k++) {
                    const uint32_t w = (c < 2) ? r[k].x : r[k].y;
                    v[k] = (c & 1) ? (w >> 16) : (w & 0xffff);
                }
                const uint32_t acc = 18u * (v[0] + v[6]) + 34u * (v[1] + v[5]) + 48u * (v[2] + v[4]) + 56u * v[3];
                o[c] = (acc + (1u << 15)) >> 16;
            }
            uint8_t *out = dst + (size_t)gy * L.pitch + gx;
            if (gx + 3 < L.w) {
                *reinterpret_cast<uint32_t *>(out) = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
            } else {
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (gx + c < L.w) out[c] = (uint8_t)o[c];
            }
        }
    }
}

// ------------------------------------------------------------------------------ O5 + O7 + O8
// cv::fastAtan2 (scalar atan_f32): 7th-order odd polynomial in degrees; no FMA (file built -fmad=false).
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float p1 = 0.9997878412794807f * (float)(180 / 3.14159265358979323846);
    const float p3 = -0.3258083974640975f * (float)(180 / 3.14159265358979323846);
    const float p5 = 0.1555786518463281f * (float)(180 / 3.14159265358979323846);
    const float p7 = -0.04432655554792128f * (float)(180 / 3.14159265358979323846);
    const float eps = (float)2.2204460492503131e-16;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// One warp per keypoint.  Level offsets come from an in-kernel prefix over the per-level counts.
__global__ void __launch_bounds__(256)
    k_describe(const LevelInfo *__restrict__ lv, int nlevels, const uint8_t *__restrict__ pyr,
               const uint8_t *__restrict__ blur, size_t frameStride, const LevelKp *__restrict__ lvlKps,
               const int *__restrict__ lvlCount, int kpCapTotal, int capOut, msl_keypoint *__restrict__ kps,
               uint8_t *__restrict__ desc, int *__restrict__ counts, int *__restrict__ err) {
    __shared__ int8_t pat[1024];  // transposed: pat[(k*4 + c)*32 + byte] -> conflict-free per-lane reads
    __shared__ int lvlOff[MAX_LEVELS + 1];
    const int frame = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < 1024; i += 256) {
        // source layout: byte b (0..31), test k (0..7), component c (x0,y0,x1,y1) at b*32 + k*4 + c
        int b = i >> 5, kc = i & 31;
        pat[kc * 32 + b] = c_pattern.v[i];
    }
    if (tid == 0) {
        int o = 0;
        for (int l = 0; l < nlevels; l++) {
            lvlOff[l] = o;
            o += lvlCount[frame * nlevels + l];
        }
        lvlOff[nlevels] = o;
        if (blockIdx.x == 0) {
            counts[frame] = min(o, capOut);
            if (o > capOut) atomicExch(err, 3);
        }
    }
    __syncthreads();
    const int total = min(lvlOff[nlevels], capOut);
    for (int k = blockIdx.x * 8 + wid; k < total; k += gridDim.x * 8) {
        int level = 0;
        while (level + 1 < nlevels && k >= lvlOff[level + 1]) level++;
        const LevelInfo L = lv[level];
        const LevelKp kp = lvlKps[(size_t)frame * kpCapTotal + L.kpBase + (k - lvlOff[level])];
        // IC_Angle (src/ORBextractor.cc:75-99): lanes = columns u in [-15,15], loop rows v
        const uint8_t *img = pyr + frame * frameStride + L.offset + (size_t)kp.y * L.pitch + kp.x;
        const int u = lane - HALF_PATCH;
        int m10 = 0, m01 = 0;
        if (lane < 31) {
#pragma unroll 1
            for (int v = -HALF_PATCH; v <= HALF_PATCH; v++) {
                if (abs(u) <= c_umax[abs(v)]) {
                    const int val = img[v * L.pitch + u];
                    m10 += u * val;
                    m01 += v * val;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        }
        const float angle = fast_atan2_deg((float)m01, (float)m10);
        // computeOrbDescriptor (src/ORBextractor.cc:104-149): a=cosf(angle*pi/180), b=sinf(..)
        const float factorPI = (float)(3.14159265358979323846 / 180.f);
        const float ang = __fmul_rn(angle, factorPI);
        const float a = (float)cos((double)ang), b = (float)sin((double)ang);
        const uint8_t *ctr = blur + frame * frameStride + L.offset + (size_t)kp.y * L.pitch + kp.x;
        int val = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const float x0 = (float)pat[(t * 4 + 0) * 32 + lane], y0 = (float)pat[(t * 4 + 1) * 32 + lane];
            const float x1 = (float)pat[(t * 4 + 2) * 32 + lane], y1 = (float)pat[(t * 4 + 3) * 32 + lane];
            const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
            const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
            const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
            const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
            const int t0 = ctr[r0 * L.pitch + c0], t1 = ctr[r1 * L.pitch + c1];
            val |= (t0 < t1) << t;
        }
        desc[((size_t)frame * capOut + k) * 32 + lane] = (uint8_t)val;
        if (lane == 0) {
            msl_keypoint o;
            float fx = (float)kp.x, fy = (float)kp.y;
            if (level != 0) {  // src/ORBextractor.cc:861-866
                fx = __fmul_rn(fx, L.scale);
                fy = __fmul_rn(fy, L.scale);
            }
            o.x = fx;
            o.y = fy;
            o.size = (float)L.scaledPatch;
            o.angle = angle;
            o.response = (float)kp.score;
            o.octave = level;
            o.class_id = -1;
            kps[(size_t)frame * capOut + k] = o;
        }
    }
}

}  // namespace

// =============================================================================================
struct msl_orb {
    msl_orb_params prm;
    int w, h, maxBatch, device;
    cudaStream_t stream = nullptr;
    int nlevels;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> featPerLevel;
    std::vector<LevelInfo> lv;
    std::vector<int> blurTileBase;
    int totalCells = 0, capCell = 0, kpCapTotal = 0, capOut = 0, maxNodes = 0, blurTiles = 0, candCapTotal = 0;
    int tileW = 0, tileH = 0;   // TMA box of the FAST cell tiles
    CUtensorMap *d_tmaps = nullptr;
    size_t pyrBytes = 0;
    size_t octSmem = 0;
    // device
    LevelInfo *d_lv = nullptr;
    Cell *d_cells = nullptr;
    short *d_cellLevel = nullptr, *d_tab = nullptr;
    int4 *d_xt4 = nullptr;   // packed column tables of k_resize4 (two int2 entries per int4)
    bool resize4 = true;     // MSL_ORB_RESIZE4
    int *d_blurTileBase = nullptr;
    uint8_t *d_pyr = nullptr, *d_blur = nullptr;
    uint32_t *d_staging = nullptr, *d_candRec = nullptr;
    unsigned short *d_candNode = nullptr;
    int *d_cellCount = nullptr, *d_candCount = nullptr, *d_lvlCount = nullptr, *d_err = nullptr;
    LevelKp *d_lvlKps = nullptr;
    msl_keypoint *d_kps = nullptr;
    uint8_t *d_desc = nullptr;
    int *d_counts = nullptr;
    int lastBatch = 0;
};

static void orb_free(msl_orb *o) {
    if (!o) return;
    cudaSetDevice(o->device);
    void *ptrs[] = {o->d_lv, o->d_cells, o->d_cellLevel, o->d_tab, o->d_xt4, o->d_blurTileBase, o->d_pyr, o->d_blur,
                    o->d_staging, o->d_candRec, o->d_candNode, o->d_cellCount, o->d_candCount, o->d_lvlCount,
                    o->d_err, o->d_lvlKps, o->d_kps, o->d_desc, o->d_counts, o->d_tmaps};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (o->stream) cudaStreamDestroy(o->stream);
    delete o;
}

extern "C" {

const char *msl_last_error(void) { return g_last_error.c_str(); }
const char *msl_version(void) { return "manhattanslam_b200 0.1 (sm_100a)"; }
uint64_t msl_kernel_launch_count(void) { return g_launches.load(); }

int msl_orb_create(const msl_orb_params *prm, int w, int h, int max_batch, int device, msl_orb **out) {
    if (!prm || !out) return fail(MSL_ERR_INVALID, "msl_orb_create: null argument");
    *out = nullptr;
    if (prm->nlevels < 1 || prm->nlevels > MAX_LEVELS || prm->nfeatures < 1 || !(prm->scale_factor > 1.0f) ||
        w < 64 || h < 64 || w > 4095 || h > 4095 || max_batch < 1)
        return fail(MSL_ERR_INVALID, "msl_orb_create: parameter out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device || device < 0)
        return fail(MSL_ERR_CUDA, "msl_orb_create: no usable CUDA device (there is no CPU fallback)");
    MSL_CUDA(cudaSetDevice(device));
    msl_orb *o = new msl_orb();
    o->prm = *prm;
    o->w = w, o->h = h, o->maxBatch = max_batch, o->device = device;
    const int nl = o->nlevels = prm->nlevels;
    // ---- ORBextractor::ORBextractor, src/ORBextractor.cc:412-445 (float/double mix kept literally)
    const double scaleFactor = prm->scale_factor;
    o->scale.resize(nl), o->invScale.resize(nl), o->sigma2.resize(nl), o->invSigma2.resize(nl);
    o->scale[0] = 1.0f, o->sigma2[0] = 1.0f;
    for (int i = 1; i < nl; i++) {
        o->scale[i] = (float)(o->scale[i - 1] * scaleFactor);
        o->sigma2[i] = o->scale[i] * o->scale[i];
    }
    for (int i = 0; i < nl; i++) {
        o->invScale[i] = 1.0f / o->scale[i];
        o->invSigma2[i] = 1.0f / o->sigma2[i];
    }
    o->featPerLevel.resize(nl);
    {
        float factor = (float)(1.0f / scaleFactor);
        float nDesired = prm->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
        int sum = 0;
        for (int l = 0; l < nl - 1; l++) {
            o->featPerLevel[l] = cv_round_f(nDesired);
            sum += o->featPerLevel[l];
            nDesired *= factor;
        }
        o->featPerLevel[nl - 1] = std::max(prm->nfeatures - sum, 0);
    }
    int umax[16];
    {   // src/ORBextractor.cc:453-467
        int v, v0, vmax = cv_floor_d(HALF_PATCH * sqrtf(2.f) / 2 + 1);
        int vmin = cv_ceil_d(HALF_PATCH * sqrtf(2.f) / 2);
        const double hp2 = HALF_PATCH * HALF_PATCH;
        for (v = 0; v <= vmax; ++v) umax[v] = (int)nearbyint(sqrt(hp2 - v * v));
        for (v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }
    // ---- level geometry, resize tables, FAST cells
    o->lv.resize(nl);
    std::vector<short> tab;
    std::vector<int> xtab;  // k_resize4's packed column tables, two ints per column
    std::vector<Cell> cells;
    std::vector<short> cellLevel;
    size_t off = 0;
    int kpBase = 0, maxNodes = 0, capCell = 0, blurTiles = 0;
    o->blurTileBase.resize(nl + 1);
    for (int l = 0; l < nl; l++) {
        LevelInfo &L = o->lv[l];
        L.w = cv_round_f((float)w * o->invScale[l]);  // src/ORBextractor.cc:875
        L.h = cv_round_f((float)h * o->invScale[l]);
        L.pitch = (int)align_up(L.w, 16);
        L.offset = (int)off;
        off = align_up(off + (size_t)L.pitch * L.h, 256);
        L.scale = o->scale[l];
        L.scaledPatch = (int)(PATCH_SIZE * o->scale[l]);
        L.nFeatures = o->featPerLevel[l];
        const int maxBorderX = L.w - EDGE_THRESHOLD + 3, maxBorderY = L.h - EDGE_THRESHOLD + 3;
        L.width = maxBorderX - MIN_BORDER, L.height = maxBorderY - MIN_BORDER;
        const float W = 30;
        const float width = (float)L.width, height = (float)L.height;
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols < 1 || nRows < 1 || L.height <= 0) {
            orb_free(o);
            return fail(MSL_ERR_INVALID, "msl_orb_create: pyramid level too small for a 30-px FAST cell");
        }
        const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
        L.nIni = (int)roundf((float)L.width / (float)L.height);  // src/ORBextractor.cc:535
        if (L.nIni < 1) {
            orb_free(o);
            return fail(MSL_ERR_INVALID, "msl_orb_create: aspect ratio gives zero octree roots");
        }
        L.hX = (float)L.width / L.nIni;
        L.cellBase = (int)cells.size();
        for (int i = 0; i < nRows; i++) {  // src/ORBextractor.cc:745-761
            const float iniY = (float)(MIN_BORDER + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) continue;
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(MIN_BORDER + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) continue;
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                Cell c = {(short)iniX, (short)iniY, (short)maxX, (short)maxY};
                if (c.x1 - c.x0 < 7 || c.y1 - c.y0 < 7) continue;  // cv::FAST on <7 px finds nothing
                if (c.x1 - c.x0 > CELL_TILE || c.y1 - c.y0 > CELL_TILE) {
                    orb_free(o);
                    return fail(MSL_ERR_INVALID, "msl_orb_create: FAST cell exceeds tile");
                }
                cells.push_back(c);
                cellLevel.push_back((short)l);
                o->tileW = std::max(o->tileW, (int)align_up(c.x1 - c.x0 + 15, 16));
                o->tileH = std::max(o->tileH, c.y1 - c.y0);
            }
        }
        L.nCells = (int)cells.size() - L.cellBase;
        capCell = std::max(capCell, ((wCell + 1) / 2) * ((hCell + 1) / 2));
        L.kpCap = std::max(L.nFeatures + 3, 4 * L.nIni) + 1;
        L.kpBase = kpBase;
        kpBase += L.kpCap;
        maxNodes = std::max(maxNodes, L.kpCap + 1);
        o->blurTileBase[l] = blurTiles;
        blurTiles += cdiv(L.w, BLUR_TW) * cdiv(L.h, BLUR_TH);
        // resize tables for level l from level l-1 (cv::resize, imgproc/resize.cpp)
        L.tabOff = (int)tab.size();
        L.xtOff = 0, L.xSpan = 0;
        if (l > 0) {
            const LevelInfo &P = o->lv[l - 1];
            const double sx_ = 1. / ((double)L.w / P.w), sy_ = 1. / ((double)L.h / P.h);
            std::vector<short> xo(L.w), a0(L.w), a1(L.w), yo(L.h), b0(L.h), b1(L.h);
            for (int dx = 0; dx < L.w; dx++) {
                float fx = (float)((dx + 0.5) * sx_ - 0.5);
                int sx = cv_floor_d(fx);
                fx -= sx;
                if (sx < 0) fx = 0, sx = 0;
                if (sx >= P.w - 1) fx = 0, sx = P.w - 1;
                xo[dx] = (short)sx;
                a0[dx] = sat_short((1.f - fx) * 2048);
                a1[dx] = sat_short(fx * 2048);
            }
            for (int dy = 0; dy < L.h; dy++) {
                float fy = (float)((dy + 0.5) * sy_ - 0.5);
                int sy = cv_floor_d(fy);
                fy -= sy;
                yo[dy] = (short)sy;
                b0[dy] = sat_short((1.f - fy) * 2048);
                b1[dy] = sat_short(fy * 2048);
            }
            for (auto *v : {&xo, &a0, &a1, &yo, &b0, &b1}) tab.insert(tab.end(), v->begin(), v->end());
            // packed column table of k_resize4: {sx | sx1 << 16, a0 | a1 << 16}, padded to four columns with the last one
            L.xtOff = (int)(xtab.size() / 2);
            L.xSpan = 0;
            const int wPad = (L.w + 3) & ~3;
            for (int dx = 0; dx < wPad; dx++) {
                const int c = std::min(dx, L.w - 1);
                const int sx0_ = xo[c], sx1_ = std::min(sx0_ + 1, P.w - 1);
                xtab.push_back((sx0_ & 0xffff) | (sx1_ << 16));
                xtab.push_back(((int)a0[c] & 0xffff) | ((int)a1[c] << 16));
                if (a0[c] < 0 || a1[c] < 0) L.xSpan = 1 << 20;  // (coefficients are 0..2048: never)
            }
            for (int g = 0; g < wPad; g += 4) {
                const int first = xo[std::min(g, L.w - 1)] & ~3;
                for (int q = 0; q < 4; q++) {
                    const int c = std::min(g + q, L.w - 1);
                    L.xSpan = std::max(L.xSpan, std::min((int)xo[c] + 1, P.w - 1) + 1 - first);
                    if (xo[c] < xo[std::min(g, L.w - 1)]) L.xSpan = 1 << 20;  // (source columns ascend: never)
                }
            }
        }
    }
    o->blurTileBase[nl] = blurTiles;
    o->blurTiles = blurTiles;
    o->pyrBytes = align_up(off, 256);
    o->totalCells = (int)cells.size();
    o->capCell = capCell;
    o->kpCapTotal = kpBase;
    o->capOut = kpBase;
    o->maxNodes = maxNodes;
    for (int l = 0; l < nl; l++)
        if (o->lv[l].nCells > 4 * maxNodes) o->maxNodes = maxNodes = cdiv(o->lv[l].nCells, 4) + 1;
    {   // candidate slots per level: every cell can hold at most capCell NMS survivors
        int cb = 0;
        for (int l = 0; l < nl; l++) {
            o->lv[l].candBase = cb;
            o->lv[l].candCap = o->lv[l].nCells * capCell;
            cb += (o->lv[l].candCap + 3) & ~3;
        }
        o->candCapTotal = cb;
    }
    o->octSmem = (size_t)(2 * 2 + 2 + 4 + 8) * maxNodes * sizeof(int) + 64 * sizeof(int);
    if (o->octSmem > 200 * 1024) {
        orb_free(o);
        return fail(MSL_ERR_INVALID, "msl_orb_create: nfeatures per level too large for the octree kernel");
    }
    if (tab.empty()) tab.push_back(0);
    // ---- device allocations
    const size_t B = max_batch;
#define ALLOC(ptr, bytes)                                   \
    do {                                                    \
        cudaError_t e_ = cudaMalloc((void **)&(ptr), (bytes)); \
        if (e_ != cudaSuccess) {                            \
            orb_free(o);                                    \
            return fail(MSL_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e_)); \
        }                                                   \
    } while (0)
    ALLOC(o->d_lv, sizeof(LevelInfo) * nl);
    ALLOC(o->d_cells, sizeof(Cell) * cells.size());
    ALLOC(o->d_cellLevel, sizeof(short) * cells.size());
    ALLOC(o->d_tab, sizeof(short) * tab.size());
    ALLOC(o->d_xt4, sizeof(int) * std::max<size_t>(xtab.size(), 4));
    ALLOC(o->d_blurTileBase, sizeof(int) * (nl + 1));
    ALLOC(o->d_pyr, B * o->pyrBytes);
    ALLOC(o->d_blur, B * o->pyrBytes);
    ALLOC(o->d_staging, B * o->totalCells * (size_t)capCell * sizeof(uint32_t));
    ALLOC(o->d_candRec, B * (size_t)o->candCapTotal * sizeof(uint32_t));
    ALLOC(o->d_candNode, B * (size_t)o->candCapTotal * sizeof(unsigned short));
    ALLOC(o->d_cellCount, B * o->totalCells * sizeof(int));
    ALLOC(o->d_candCount, B * nl * sizeof(int));
    ALLOC(o->d_lvlCount, B * nl * sizeof(int));
    ALLOC(o->d_err, sizeof(int));
    ALLOC(o->d_lvlKps, B * o->kpCapTotal * sizeof(LevelKp));
    ALLOC(o->d_kps, B * o->capOut * sizeof(msl_keypoint));
    ALLOC(o->d_desc, B * o->capOut * 32);
    ALLOC(o->d_counts, B * sizeof(int));
#undef ALLOC
    MSL_CUDA(cudaStreamCreateWithFlags(&o->stream, cudaStreamNonBlocking));
    MSL_CUDA(cudaMemcpy(o->d_lv, o->lv.data(), sizeof(LevelInfo) * nl, cudaMemcpyHostToDevice));
    MSL_CUDA(cudaMemcpy(o->d_cells, cells.data(), sizeof(Cell) * cells.size(), cudaMemcpyHostToDevice));
    MSL_CUDA(cudaMemcpy(o->d_cellLevel, cellLevel.data(), sizeof(short) * cells.size(), cudaMemcpyHostToDevice));
    MSL_CUDA(cudaMemcpy(o->d_tab, tab.data(), sizeof(short) * tab.size(), cudaMemcpyHostToDevice));
    if (!xtab.empty()) MSL_CUDA(cudaMemcpy(o->d_xt4, xtab.data(), sizeof(int) * xtab.size(), cudaMemcpyHostToDevice));
    if (const char *e = getenv("MSL_ORB_RESIZE4")) o->resize4 = atoi(e) != 0;
    MSL_CUDA(cudaMemcpy(o->d_blurTileBase, o->blurTileBase.data(), sizeof(int) * (nl + 1), cudaMemcpyHostToDevice));
    {   // one 3-D tensor map (x, y, frame) per pyramid level for the TMA tile loads of k_fast_cells
        typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
            qres != cudaDriverEntryPointSuccess || o->tileW > TILE_W_MAX || o->tileH > CELL_TILE) {
            orb_free(o);
            return fail(MSL_ERR_CUDA, "msl_orb_create: cuTensorMapEncodeTiled unavailable (TMA is required)");
        }
        std::vector<CUtensorMap> maps(nl);
        for (int l = 0; l < nl; l++) {
            const LevelInfo &L = o->lv[l];
            const cuuint64_t gdim[3] = {(cuuint64_t)L.w, (cuuint64_t)L.h, (cuuint64_t)max_batch};
            const cuuint64_t gstr[2] = {(cuuint64_t)L.pitch, (cuuint64_t)o->pyrBytes};
            const cuuint32_t box[3] = {(cuuint32_t)o->tileW, (cuuint32_t)o->tileH, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = ((EncodeTiled)fn)(&maps[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, o->d_pyr + L.offset, gdim, gstr, box, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                orb_free(o);
                return fail(MSL_ERR_CUDA, "msl_orb_create: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
            }
        }
        cudaError_t e_ = cudaMalloc((void **)&o->d_tmaps, sizeof(CUtensorMap) * nl);
        if (e_ == cudaSuccess) e_ = cudaMemcpy(o->d_tmaps, maps.data(), sizeof(CUtensorMap) * nl, cudaMemcpyHostToDevice);
        if (e_ != cudaSuccess) {
            orb_free(o);
            return fail(MSL_ERR_CUDA, std::string("tensor maps: ") + cudaGetErrorString(e_));
        }
    }
    MSL_CUDA(cudaMemcpyToSymbol(c_umax, umax, sizeof(umax)));
    MSL_CUDA(cudaMemset(o->d_err, 0, sizeof(int)));
    MSL_CUDA(cudaMemset(o->d_pyr, 0, B * o->pyrBytes));
    MSL_CUDA(cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)o->octSmem));
    *out = o;
    return MSL_OK;
}

void msl_orb_destroy(msl_orb *o) { orb_free(o); }
int msl_orb_levels(const msl_orb *o) { return o ? o->nlevels : 0; }
int msl_orb_capacity(const msl_orb *o) { return o ? o->capOut : 0; }
void *msl_orb_stream(msl_orb *o) { return o ? (void *)o->stream : nullptr; }

int msl_orb_scale_factors(const msl_orb *o, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2) {
    if (!o) return fail(MSL_ERR_INVALID, "null handle");
    for (int i = 0; i < o->nlevels; i++) {
        if (scale) scale[i] = o->scale[i];
        if (inv_scale) inv_scale[i] = o->invScale[i];
        if (sigma2) sigma2[i] = o->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = o->invSigma2[i];
    }
    return MSL_OK;
}

// Enqueue the whole pipeline; level 0 of every frame must already be in d_pyr.
static int orb_run(msl_orb *o, int batch, msl_keypoint *d_kps, uint8_t *d_desc, int32_t *d_counts) {
    cudaStream_t st = o->stream;
    const int nl = o->nlevels;
    for (int l = 1; l < nl; l++) {
        if (o->resize4 && o->lv[l].xSpan <= 12) {  // four columns per thread: their source bytes lie in three aligned words
            dim3 g(cdiv(cdiv(o->lv[l].w, 4), 32), cdiv(o->lv[l].h, 8), batch);
            k_resize4<<<g, dim3(32, 8), 0, st>>>(o->d_lv, l, o->d_pyr, o->pyrBytes, o->d_tab, o->d_xt4);
        } else {
            dim3 g(cdiv(o->lv[l].w, 32), cdiv(o->lv[l].h, 8), batch);
            k_resize<<<g, dim3(32, 8), 0, st>>>(o->d_lv, l, o->d_pyr, o->pyrBytes, o->d_tab);
        }
        MSL_LAUNCH_CHECK();
    }
    k_fast_cells<<<dim3(o->totalCells, batch), 128, 0, st>>>(o->d_lv, o->d_cells, o->d_cellLevel, o->d_pyr,
                                                             o->pyrBytes, o->d_tmaps, o->tileW, o->tileW * o->tileH,
                                                             o->totalCells, o->capCell, o->prm.ini_th_fast,
                                                             o->prm.min_th_fast, o->d_staging, o->d_cellCount, o->d_err);
    MSL_LAUNCH_CHECK();
    k_octree<<<dim3(nl, batch), 256, o->octSmem, st>>>(o->d_lv, nl, o->totalCells, o->capCell, o->d_staging,
                                                       o->d_cellCount, o->d_candRec, o->d_candNode, o->d_candCount,
                                                       o->d_lvlKps, o->d_lvlCount, o->kpCapTotal, o->candCapTotal, o->maxNodes,
                                                       o->d_err);
    MSL_LAUNCH_CHECK();
    k_blur<<<dim3(o->blurTiles, batch), 256, 0, st>>>(o->d_lv, o->d_blurTileBase, nl, o->d_pyr, o->d_blur,
                                                      o->pyrBytes);
    MSL_LAUNCH_CHECK();
    k_describe<<<dim3(cdiv(o->capOut, 8), batch), 256, 0, st>>>(o->d_lv, nl, o->d_pyr, o->d_blur, o->pyrBytes,
                                                                o->d_lvlKps, o->d_lvlCount, o->kpCapTotal, o->capOut,
                                                                d_kps, d_desc, d_counts, o->d_err);
    MSL_LAUNCH_CHECK();
    o->lastBatch = batch;
    return MSL_OK;
}

static int orb_check_err(msl_orb *o) {
    int e = 0;
    MSL_CUDA(cudaMemcpyAsync(&e, o->d_err, sizeof(int), cudaMemcpyDeviceToHost, o->stream));
    MSL_CUDA(cudaStreamSynchronize(o->stream));
    if (e) {
        cudaMemsetAsync(o->d_err, 0, sizeof(int), o->stream);
        if (e == 4) return fail(MSL_ERR_CUDA, "ORB: TMA tile load did not complete (fallback loads were used)");
        return fail(MSL_ERR_CAPACITY, e == 2 ? "ORB: octree produced more nodes than the level capacity"
                                               : "ORB: keypoint output capacity exceeded");
    }
    return MSL_OK;
}

int msl_orb_extract_dev(msl_orb *o, const uint8_t *d_gray, int stride, size_t frame_stride, int batch,
                        msl_keypoint *d_kps, uint8_t *d_desc, int32_t *d_counts) {
    if (!o || !d_gray || !d_kps || !d_desc || !d_counts) return fail(MSL_ERR_INVALID, "msl_orb_extract_dev: null argument");
    if (batch < 1 || batch > o->maxBatch || stride < o->w) return fail(MSL_ERR_INVALID, "msl_orb_extract_dev: bad batch/stride");
    MSL_CUDA(cudaSetDevice(o->device));
    const int vec = (o->w % 16 == 0) && (stride % 16 == 0) && (frame_stride % 16 == 0) && (((uintptr_t)d_gray) % 16 == 0);
    k_load_level0<<<dim3(cdiv(o->w, 256 * 16), o->h, batch), 256, 0, o->stream>>>(d_gray, stride, frame_stride, o->d_pyr,
                                                                               o->lv[0].pitch, o->pyrBytes, o->w, o->h, vec);
    MSL_LAUNCH_CHECK();
    return orb_run(o, batch, d_kps, d_desc, d_counts);
}

int msl_orb_sync(msl_orb *o) {
    if (!o) return fail(MSL_ERR_INVALID, "null handle");
    MSL_CUDA(cudaSetDevice(o->device));
    return orb_check_err(o);
}

int msl_orb_extract(msl_orb *o, const uint8_t *gray, int stride, size_t frame_stride, int batch, msl_keypoint *kps,
                    uint8_t *desc, int32_t *counts) {
    if (!o || !kps || !desc || !counts) return fail(MSL_ERR_INVALID, "msl_orb_extract: null argument");
    if (batch < 1 || batch > o->maxBatch) return fail(MSL_ERR_INVALID, "msl_orb_extract: bad batch");
    if (!gray) {  // _image.empty() => silent return, src/ORBextractor.cc:815-816
        for (int b = 0; b < batch; b++) counts[b] = 0;
        return MSL_OK;
    }
    if (stride < o->w) return fail(MSL_ERR_INVALID, "msl_orb_extract: stride < width");
    MSL_CUDA(cudaSetDevice(o->device));
    if (stride == o->w && o->lv[0].pitch == o->w) {  // dense frames: one strided copy, one "row" per frame
        MSL_CUDA(cudaMemcpy2DAsync(o->d_pyr, o->pyrBytes, gray, frame_stride, (size_t)o->w * o->h, batch,
                                   cudaMemcpyHostToDevice, o->stream));
    } else {
        for (int b = 0; b < batch; b++)
            MSL_CUDA(cudaMemcpy2DAsync(o->d_pyr + b * o->pyrBytes, o->lv[0].pitch, gray + b * frame_stride, stride,
                                       o->w, o->h, cudaMemcpyHostToDevice, o->stream));
    }
    int rc = orb_run(o, batch, o->d_kps, o->d_desc, o->d_counts);
    if (rc) return rc;
    MSL_CUDA(cudaMemcpyAsync(counts, o->d_counts, sizeof(int) * batch, cudaMemcpyDeviceToHost, o->stream));
    MSL_CUDA(cudaMemcpyAsync(kps, o->d_kps, sizeof(msl_keypoint) * (size_t)batch * o->capOut, cudaMemcpyDeviceToHost, o->stream));
    MSL_CUDA(cudaMemcpyAsync(desc, o->d_desc, (size_t)batch * o->capOut * 32, cudaMemcpyDeviceToHost, o->stream));
    return orb_check_err(o);
}

int msl_orb_debug_level_size(const msl_orb *o, int level, int *w, int *h) {
    if (!o || level < 0 || level >= o->nlevels) return fail(MSL_ERR_INVALID, "bad level");
    *w = o->lv[level].w, *h = o->lv[level].h;
    return MSL_OK;
}

int msl_orb_debug_level(msl_orb *o, int frame, int level, int blurred, uint8_t *out) {
    if (!o || level < 0 || level >= o->nlevels || frame < 0 || frame >= o->maxBatch) return fail(MSL_ERR_INVALID, "bad level/frame");
    MSL_CUDA(cudaSetDevice(o->device));
    const LevelInfo &L = o->lv[level];
    const uint8_t *src = (blurred ? o->d_blur : o->d_pyr) + frame * o->pyrBytes + L.offset;
    MSL_CUDA(cudaStreamSynchronize(o->stream));
    MSL_CUDA(cudaMemcpy2D(out, L.w, src, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    return MSL_OK;
}

int msl_orb_debug_candidates(msl_orb *o, int frame, int level, int32_t *xyr, int cap, int *n) {
    if (!o || level < 0 || level >= o->nlevels || frame < 0 || frame >= o->maxBatch) return fail(MSL_ERR_INVALID, "bad level/frame");
    MSL_CUDA(cudaSetDevice(o->device));
    MSL_CUDA(cudaStreamSynchronize(o->stream));
    int cnt = 0;
    MSL_CUDA(cudaMemcpy(&cnt, o->d_candCount + frame * o->nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    *n = cnt;
    int m = std::min(std::min(cnt, cap), o->lv[level].candCap);
    std::vector<uint32_t> rec(m);
    if (m) MSL_CUDA(cudaMemcpy(rec.data(), o->d_candRec + (size_t)frame * o->candCapTotal + o->lv[level].candBase, m * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < m; i++) {
        xyr[3 * i] = rec[i] & 0xfff;
        xyr[3 * i + 1] = (rec[i] >> 12) & 0xfff;
        xyr[3 * i + 2] = (int)(rec[i] >> 24) - 1;
    }
    return MSL_OK;
}

}  // extern "C"
