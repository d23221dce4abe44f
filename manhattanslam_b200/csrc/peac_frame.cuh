// peac_frame.cuh -- per-frame agglomerative plane clustering + refinement: the part of ahc::PlaneFitter::run that follows
// the pre-stage.  Replaces ahCluster (include/peac/AHCPlaneFitter.hpp:939-1143) and
// refineDetails (:294-374) with findBlockMembership (:480-582) and floodFill (:422-471), on top of PlaneSeg's merge
// constructor / connect / disconnectAllNbs / mergeNbsFrom (include/peac/AHCPlaneSeg.hpp:321-437), Stats::compute
// (:148-181) and DisjointSet (include/peac/DisjointSet.hpp).  Output: PlaneFitter::membershipImg and extractedPlanes.
//
// One CTA per frame, everything but the region grow in shared memory (196 KB; frames of more than 768 blocks: the same
// working set, SharedT<3072>, in global memory):
//   * node table: one 144-byte record per 10x10 block slot (the nine running sums, centre, normal, mse, N, rid, creation
//     sequence).  A merged node REUSES the slot of the node that was popped from the queue (that slot is referenced
//     nowhere else any more), so 768 slots serve the <= 1535 nodes a frame can create;
//   * adjacency: one 768-bit mask per slot instead of std::set<PlaneSeg*>.  The reference visits neighbours in
//     heap-address order; that order only decides EXACT mse ties between merge candidates, which are resolved here by
//     creation sequence (what the reference does inside a bump arena, oracle/ref_arena.hpp);
//   * min-MSE queue: a binary heap of slot ids maintained by thread 0 with libstdc++'s push_heap / pop_heap sift
//     order, so that equal keys leave the queue in the same order as std::priority_queue;
//   * per merge step the candidate fits (merged sums -> 3x3 scatter -> Jacobi eigen-solve) of ALL neighbours run in
//     parallel, one thread per neighbour slot; the adjacency update runs one thread per mask row;
//   * findBlockMembership: one thread per block; the seeds of the region grow are written in the reference's order
//     through a per-block count + exclusive scan;
//   * floodFill is an order-dependent FIFO (a pixel keeps the first plane that reaches it with a smaller distance, and
//     counts failed visits in negative "trail" values).  It is run LEVEL BY LEVEL with the same result: the entries a
//     level pushes form the next level; a visit (entry, neighbour) only reads and writes the state of its TARGET pixel
//     (trail value, best distance), so the visits of a level are resolved in parallel across target pixels and in visit
//     order per pixel (rounds of atomicMin(own[pixel], visit): the earliest pending visit of every pixel resolves);
//     the geometric part of a visit (point-plane distance, 3-sigma test) is state-independent and precomputed for all
//     visits at once; the next level is written in visit order through a chunked scan of the push flags.  The plain
//     FIFO on thread 0 is kept behind Geo::floodSerial (MSL_PEAC_FLOOD_SERIAL=1) as a cross-check;
//   * the final merge reuses the cluster loop on the extracted planes; the plane-id remap and the per-plane pixel counts
//     run one thread per pixel.
// All arithmetic is fp64 in the reference's operation order (built with -fmad=false).
//
// The same source compiles for the host so that it can be checked on CPU-only machines by
// tests/test_peac_host_emulation.py: PEAC_HOST_EMULATION (one "thread", barriers are no-ops) checks the algorithm against
// the oracle; PEAC_HOST_EMULATION_MT (several real threads, PEAC_SYNC = a pthread barrier, built with ThreadSanitizer)
// checks that every shared-memory hand-over between the parallel phases is separated by a barrier.  Test harnesses, not a
// fallback: the product entry points (msl_plane_detect*) only ever launch the kernel.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__) && !defined(PEAC_HOST_EMULATION)
#define PEAC_HD __device__ __forceinline__
#define PEAC_D __device__
#define PEAC_SYNC() __syncthreads()
#define PEAC_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define PEAC_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define PEAC_ATOMIC_OR(p, v) atomicOr((p), (v))
#define PEAC_LOAD(p) (*(volatile const int *)(p))
#define PEAC_STORE(p, v) (*(volatile int *)(p) = (v))
#define PEAC_UNROLL _Pragma("unroll")
#define PEAC_STORE_FLAG(p) (*(p) = 1)  // several threads may store the same 1
#define PEAC_STAMP(F, i, tid)                         \
    do {                                              \
        if ((F).prof && (tid) == 0) {                 \
            unsigned long long t_;                    \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); \
            (F).prof[i] = (long long)t_;              \
        }                                             \
    } while (0)
#define PEAC_CLOCK() clock64()
#elif defined(PEAC_HOST_EMULATION_MT)  // tests/host_emul/peac_host_mt.cpp: real threads + a barrier, under ThreadSanitizer
#define PEAC_CLOCK() 0LL
void peac_emu_sync();
#define PEAC_HD inline
#define PEAC_D inline
#define PEAC_SYNC() peac_emu_sync()
#define PEAC_ATOMIC_ADD(p, v) __atomic_fetch_add((p), (v), __ATOMIC_RELAXED)
#define PEAC_ATOMIC_MIN(p, v) peac_emu_atomic_min((p), (v))
#define PEAC_ATOMIC_OR(p, v) __atomic_fetch_or((p), (v), __ATOMIC_RELAXED)
#define PEAC_LOAD(p) __atomic_load_n((p), __ATOMIC_RELAXED)
#define PEAC_STORE(p, v) __atomic_store_n((p), (v), __ATOMIC_RELAXED)
inline void peac_emu_atomic_min(int *p, int v) {
    int cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
}
#define PEAC_STORE_FLAG(p) __atomic_store_n((p), (unsigned char)1, __ATOMIC_RELAXED)
#define PEAC_UNROLL
#define PEAC_STAMP(F, i, tid) ((void)0)
#else
#define PEAC_HD inline
#define PEAC_D inline
#define PEAC_SYNC() ((void)0)
#define PEAC_ATOMIC_ADD(p, v) ((*(p) += (v)) - (v))  // returns the old value, like atomicAdd
#define PEAC_ATOMIC_MIN(p, v) (*(p) = *(p) < (v) ? *(p) : (v))
#define PEAC_ATOMIC_OR(p, v) (*(p) |= (v))
#define PEAC_LOAD(p) (*(p))
#define PEAC_STORE(p, v) (*(p) = (v))
#define PEAC_STORE_FLAG(p) (*(p) = 1)
#define PEAC_UNROLL
#define PEAC_STAMP(F, i, tid) ((void)0)
#define PEAC_CLOCK() 0LL
#endif

namespace peac {

constexpr int WIN = 10;           // windowWidth / windowHeight, AHCPlaneFitter.hpp:156-160
constexpr int MAXB = 768;         // block slots of the shared-memory instance (640x480 input: 32 x 24 blocks)
constexpr int MAXB_BIG = 3072;    // block slots of the global-memory instance (1280x960 input: 64 x 48 blocks)
constexpr int MAXPL = 128;        // extracted planes (>= 3000 of <= 76,800 points each: at most 25)
constexpr int MIN_SUPPORT = 3000, MAX_STEP = 100000;  // :155-156
#define PEAC_DEPTH_SIGMA 1.6e-6
#define PEAC_STDTOL_MERGE 8.0
#define PEAC_DEPTH_ALPHA 0.04
#define PEAC_DEPTH_CHANGE_TOL 0.02

struct Geo {
    int W2, H2, Nw, Nh;
    int dstride;  // depth row stride in pixels (full resolution)
    float fx, fy, cx, cy, factor;
    double thMerge, thRefine;  // cos(60 deg), cos(30 deg) evaluated on the host with std::cos (AHCParamSet.hpp:72-73)
    int floodSerial;           // 1: the region grow as the reference's FIFO on thread 0 (A/B and cross-check); 0: by levels
};

struct Node {        // PlaneSeg
    double s[9];     // Stats: sx sy sz sxx syy szz sxy syz sxz
    double center[3], normal[3], mse;
    int N, rid, seq, nouse;
};

struct PlaneOut {  // = msl_plane_rec
    double normal[3], center[3];
    int32_t N, rid, vertices, pad;
};

// The working set of one frame.  SharedT<768> (196 KB) lives in the CTA's shared memory; frames with more blocks use
// SharedT<3072> in global memory (1.7 MB per frame, L2-resident) -- same code, the reference S is all it sees.
template <int NB> struct SharedT {
    static constexpr int SLOTS = NB;       // block slots
    static constexpr int WORDS = NB / 32;  // adjacency mask words per slot
    Node node[NB];
    uint32_t nbs[NB][WORDS];
    double candMse[NB];  // reused as int scratch by the membership pass and the region grow
    int16_t parent[NB], dsize[NB], heap[NB], blkMap[NB];
    double heapKey[NB];  // mse of heap[i]: a queued node's mse never changes, and the sift loops compare keys without the
                         // dependent load heap[i] -> node[heap[i]].mse
    int16_t red[NB];    // the candidates (neighbour slots that pass the normal test) of the current merge step
    uint32_t tmpMask[WORDS];
    int16_t extracted[MAXPL], oldPl[MAXPL], plidmap[MAXPL];
    uint8_t valid[MAXPL];
    int32_t count[MAXPL];
    Node tmp;
    int heapN, nExtracted, nOld, seqNext, step;
    int curP, curNb, decision;  // broadcast from thread 0
    int nCand;                  // threads that hold a merge candidate of the current step (entries of red[])
    int lvlBegin, lvlEnd, pending;  // region grow by levels
    int error;
};
typedef SharedT<MAXB> Shared;
typedef SharedT<MAXB_BIG> SharedBig;

// per-frame scratch of the region grow in global memory
struct Flood {
    float *distMap;   // H2*W2
    uint32_t *rfq;    // the queue: rfqCap entries
    int rfqCap;
    int *own;         // H2*W2: earliest pending visit of a pixel in the current round
    int *visC;        // visCap: target pixel of a visit (-1: nothing to do)
    float *visDist;   // visCap
    uint8_t *visFlag; // visCap: bit0 near, bit1 pending, bit2 pushes
    int visCap;
    long long *prof;  // optional: 8 globaltimer stamps of the frame's phases + 8 cycle counters of ahCluster's sub-phases (msl_plane_debug_profile), else null
};

enum { PEAC_OK = 0, PEAC_ERR_QUEUE = 1, PEAC_ERR_PLANES = 2 };

// ---- symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations (the repository's stand-in for Eigen's solver,
// identical operation order to oracle/plane_oracle.cpp); eigenvalues ascending, V[3 * k + i] = component k of vector i
PEAC_HD void eig33(const double K[9], double s[3], double V[9]) {
    double a[3][3] = {{K[0], K[1], K[2]}, {K[3], K[4], K[5]}, {K[6], K[7], K[8]}};
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 32; sweep++) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-60 || off <= 1e-34 * diag) break;
        PEAC_UNROLL
        for (int p = 0; p < 2; p++) {
            PEAC_UNROLL
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                PEAC_UNROLL
                for (int k = 0; k < 3; k++) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - sn * akq;
                    a[k][q] = sn * akp + c * akq;
                }
                PEAC_UNROLL
                for (int k = 0; k < 3; k++) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - sn * aqk;
                    a[q][k] = sn * apk + c * aqk;
                }
                PEAC_UNROLL
                for (int k = 0; k < 3; k++) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - sn * vkq;
                    v[k][q] = sn * vkp + c * vkq;
                }
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
    PEAC_UNROLL
    for (int i = 0; i < 3; i++) {
        s[i] = e[ord[i]];
        PEAC_UNROLL
        for (int k = 0; k < 3; k++) V[k * 3 + i] = v[k][ord[i]];
    }
}

// Stats::compute, AHCPlaneSeg.hpp:148-181 (curvature is not consumed downstream)
PEAC_HD void fit(Node &n) {
    const double *s = n.s;
    const double sc = 1.0 / n.N;
    n.center[0] = s[0] * sc, n.center[1] = s[1] * sc, n.center[2] = s[2] * sc;
    double K[9] = {s[3] - s[0] * s[0] * sc, s[6] - s[0] * s[1] * sc, s[8] - s[0] * s[2] * sc, 0, s[4] - s[1] * s[1] * sc,
                   s[7] - s[1] * s[2] * sc, 0, 0, s[5] - s[2] * s[2] * sc};
    K[3] = K[1], K[6] = K[2], K[7] = K[5];
    double sv[3], V[9];
    eig33(K, sv, V);
    if ((V[0] * n.center[0] + V[3] * n.center[1]) + V[6] * n.center[2] <= 0) {
        n.normal[0] = V[0], n.normal[1] = V[3], n.normal[2] = V[6];
    } else {
        n.normal[0] = -V[0], n.normal[1] = -V[3], n.normal[2] = -V[6];
    }
    n.mse = sv[0] * sc;
}

PEAC_HD double nsim(const Node &a, const Node &b) {  // normalSimilarity :349-353
    return fabs((a.normal[0] * b.normal[0] + a.normal[1] * b.normal[1]) + a.normal[2] * b.normal[2]);
}
PEAC_HD double t_mse_merge(double z) {  // T_mse(P_MERGING | P_REFINE, z), AHCParamSet.hpp:94-97
    const double v = PEAC_DEPTH_SIGMA * z * z + PEAC_STDTOL_MERGE;
    return v * v;
}
// PlaneSeg(pa, pb), AHCPlaneSeg.hpp:321-343
PEAC_HD void merged(const Node &a, const Node &b, Node &m) {
    PEAC_UNROLL
    for (int k = 0; k < 9; k++) m.s[k] = a.s[k] + b.s[k];
    m.N = a.N + b.N;
    m.rid = a.N >= b.N ? a.rid : b.rid;
    m.nouse = 0;
    fit(m);
}

// ImagePointCloud::get (include/PlaneExtractor.h:48-56) on the cloud of readDepthImage (src/PlaneExtractor.cpp:60-74),
// recomputed from the depth image: row / col are half-resolution coordinates
PEAC_HD bool point(const Geo &g, const uint16_t *depth, int row, int col, double pt[3]) {
    const double z = (double)depth[(size_t)(2 * row) * g.dstride + 2 * col] * (double)g.factor;
    if (z == 0) return false;
    pt[0] = ((double)(2 * col) - (double)g.cx) * z / (double)g.fx;
    pt[1] = ((double)(2 * row) - (double)g.cy) * z / (double)g.fy;
    pt[2] = z;
    return true;
}

// the nine running sums of a block in the reference's row-major order (PlaneSeg ctor, AHCPlaneSeg.hpp:235-312); only
// called for blocks the pre-stage accepted, so the validity tests are not repeated
PEAC_HD void block_sums(const Geo &g, const uint16_t *depth, int blk, double s[9]) {
    const int r0 = (blk / g.Nw) * WIN, c0 = (blk % g.Nw) * WIN;
    double sx = 0, sy = 0, sz = 0, sxx = 0, syy = 0, szz = 0, sxy = 0, syz = 0, sxz = 0;
    for (int i = r0; i < r0 + WIN; ++i)
        for (int j = c0; j < c0 + WIN; ++j) {
            double p[3] = {0, 0, 0};
            point(g, depth, i, j, p);
            const double x = p[0], y = p[1], z = p[2];
            sx += x, sy += y, sz += z;
            sxx += x * x, syy += y * y, szz += z * z;
            sxy += x * y, syz += y * z, sxz += x * z;
        }
    s[0] = sx, s[1] = sy, s[2] = sz, s[3] = sxx, s[4] = syy, s[5] = szz, s[6] = sxy, s[7] = syz, s[8] = sxz;
}

// ---- adjacency masks
PEAC_HD bool bit(const uint32_t *m, int i) { return (m[i >> 5] >> (i & 31)) & 1u; }
PEAC_HD void setbit(uint32_t *m, int i) { m[i >> 5] |= 1u << (i & 31); }
PEAC_HD void clrbit(uint32_t *m, int i) { m[i >> 5] &= ~(1u << (i & 31)); }

// ---- std::priority_queue<.., PlaneSegMinMSECmp> on slot ids: comp(a, b) = mse[b] < mse[a]; libstdc++'s sift order.
// The key (mse) of every entry is kept beside its slot id.
template <class SH>
PEAC_HD void heap_sift_up(SH &S, int hole, int top, int value, double key) {  // std::__push_heap
    int parent = (hole - 1) / 2;
    while (hole > top && key < S.heapKey[parent]) {  // comp(heap[parent], value) = mse[value] < mse[heap[parent]]
        S.heap[hole] = S.heap[parent], S.heapKey[hole] = S.heapKey[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    S.heap[hole] = (int16_t)value, S.heapKey[hole] = key;
}
template <class SH>
PEAC_HD void heap_push(SH &S, int slot) {
    const int at = S.heapN;
    S.heapN = at + 1;
    heap_sift_up(S, at, 0, slot, S.node[slot].mse);
}
template <class SH>
PEAC_HD int heap_pop(SH &S) {  // top(), then std::pop_heap + pop_back
    const int top = S.heap[0];
    const int last = S.heapN - 1;
    if (last > 0) {
        const int value = S.heap[last];
        const double key = S.heapKey[last];
        const int len = last;  // std::__adjust_heap(first, 0, len, value)
        int hole = 0, second = 0;
        while (second < (len - 1) / 2) {
            second = 2 * (second + 1);
            if (S.heapKey[second - 1] < S.heapKey[second]) second--;  // comp(heap[second], heap[second - 1])
            S.heap[hole] = S.heap[second], S.heapKey[hole] = S.heapKey[second];
            hole = second;
        }
        if ((len & 1) == 0 && second == (len - 2) / 2) {
            second = 2 * (second + 1);
            S.heap[hole] = S.heap[second - 1], S.heapKey[hole] = S.heapKey[second - 1];
            hole = second - 1;
        }
        heap_sift_up(S, hole, 0, value, key);
    }
    S.heapN = last;
    return top;
}

// ---- DisjointSet.hpp (thread 0 mutates; the read-only find is for the parallel passes)
template <class SH>
PEAC_HD int ds_find(SH &S, int x) {
    int r = x;
    while (S.parent[r] != r) r = S.parent[r];
    while (S.parent[x] != r) {  // path compression (the recursion of the reference leaves the same parents)
        const int nx = S.parent[x];
        S.parent[x] = (int16_t)r;
        x = nx;
    }
    return r;
}
template <class SH>
PEAC_HD int ds_find_ro(const SH &S, int x) {
    while (S.parent[x] != x) x = S.parent[x];
    return x;
}
template <class SH>
PEAC_HD void ds_union(SH &S, int x, int y) {
    const int xr = ds_find(S, x), yr = ds_find(S, y);
    if (xr == yr) return;
    if (S.dsize[xr] < S.dsize[yr])
        S.parent[xr] = (int16_t)yr, S.dsize[yr] = (int16_t)(S.dsize[yr] + S.dsize[xr]);
    else
        S.parent[yr] = (int16_t)xr, S.dsize[xr] = (int16_t)(S.dsize[xr] + S.dsize[yr]);
}

// ---- std::sort(extractedPlanes.begin(), extractedPlanes.end(), PlaneSegSizeCmp()) exactly as libstdc++ performs it
// (introsort: median-of-three pivot + unguarded partition down to 16 elements, then one insertion sort).  std::sort is not
// stable and planes of equal size are common (sizes are multiples of 100 points), so the order of equal keys -- and with
// it the plane ids written into membershipImg -- is whatever this algorithm leaves.  comp(a, b) = N[b] < N[a].
template <class SH>
PEAC_HD bool plane_comp(const SH &S, int a, int b) { return S.node[b].N < S.node[a].N; }
template <class SH>
PEAC_HD void plane_linear_insert(SH &S, int16_t *v, int last) {  // std::__unguarded_linear_insert
    const int16_t val = v[last];
    int next = last - 1;
    while (plane_comp(S, val, v[next])) {
        v[last] = v[next];
        last = next;
        --next;
    }
    v[last] = val;
}
template <class SH>
PEAC_HD void plane_insertion_sort(SH &S, int16_t *v, int first, int last) {  // std::__insertion_sort
    for (int i = first + 1; i < last; ++i) {
        if (plane_comp(S, v[i], v[first])) {
            const int16_t val = v[i];
            for (int k = i; k > first; --k) v[k] = v[k - 1];
            v[first] = val;
        } else {
            plane_linear_insert(S, v, i);
        }
    }
}
template <class SH>
PEAC_HD bool sort_planes(SH &S) {
    int16_t *v = S.extracted;
    const int n = S.nExtracted;
    if (n < 2) return true;
    // std::__introsort_loop with its recursion on the right part turned into a stack of (first, last, depth) ranges
    int lg = 0;
    while ((n >> (lg + 1)) > 0) lg++;
    int stF[32], stL[32], stD[32], sp = 0;
    stF[0] = 0, stL[0] = n, stD[0] = 2 * lg, sp = 1;
    while (sp > 0) {
        --sp;
        int first = stF[sp], last = stL[sp], depth = stD[sp];
        while (last - first > 16) {
            if (depth == 0) return false;  // libstdc++ would heap-sort here; not reachable with <= 128 planes in practice
            --depth;
            // std::__move_median_to_first(first, first + 1, mid, last - 1)
            const int a = first + 1, b = first + (last - first) / 2, c = last - 1;
            int pick;
            if (plane_comp(S, v[a], v[b])) {
                if (plane_comp(S, v[b], v[c])) pick = b;
                else if (plane_comp(S, v[a], v[c])) pick = c;
                else pick = a;
            } else if (plane_comp(S, v[a], v[c])) pick = a;
            else if (plane_comp(S, v[b], v[c])) pick = c;
            else pick = b;
            { const int16_t t = v[first]; v[first] = v[pick]; v[pick] = t; }
            // std::__unguarded_partition(first + 1, last, pivot = first)
            int lo = first + 1, hi = last;
            for (;;) {
                while (plane_comp(S, v[lo], v[first])) ++lo;
                --hi;
                while (plane_comp(S, v[first], v[hi])) --hi;
                if (!(lo < hi)) break;
                { const int16_t t = v[lo]; v[lo] = v[hi]; v[hi] = t; }
                ++lo;
            }
            if (sp >= 32) return false;
            stF[sp] = lo, stL[sp] = last, stD[sp] = depth, ++sp;  // __introsort_loop(cut, last, depth_limit)
            last = lo;
        }
    }
    // std::__final_insertion_sort
    if (n > 16) {
        plane_insertion_sort(S, v, 0, 16);
        for (int i = 16; i < n; ++i) plane_linear_insert(S, v, i);
    } else {
        plane_insertion_sort(S, v, 0, n);
    }
    return true;
}

// ---- ahCluster, AHCPlaneFitter.hpp:939-1143.  nslots = number of block slots of the frame.
// prof (optional, thread 0): cycles accumulated per sub-phase of a merge step -- [0] queue pop, [1] candidate fits,
// [2] selection, [3] publish + decision, [4] adjacency update, [5] node copy + queue push
template <class SH>
PEAC_D void cluster(SH &S, const Geo &g, int nslots, int tid, int nt, long long *prof = nullptr) {
    long long tc = PEAC_CLOCK(), acc[6] = {0, 0, 0, 0, 0, 0};
#define PEAC_LAP(i)                          \
    do {                                     \
        if (prof && tid == 0) {              \
            const long long t_ = PEAC_CLOCK(); \
            acc[i] += t_ - tc;               \
            tc = t_;                         \
        }                                    \
    } while (0)
    for (;;) {
        if (tid == 0) {
            int p = -1;
            while (S.heapN > 0 && S.step <= MAX_STEP) {
                const int x = heap_pop(S);
                if (S.node[x].nouse) continue;
                p = x;
                break;
            }
            S.curP = p;
        }
        PEAC_SYNC();
        PEAC_LAP(0);
        const int p = S.curP;
        if (p < 0) break;
        // candidate merges with every neighbour, in parallel; every thread keeps the best of its own slots
        // (least mse, then earliest creation: the first minimum in the reference's neighbour order) together with the
        // merged node itself, and every candidate goes to a compact list (a node has a handful of neighbours: thread 0's
        // selection and tie walk look at those entries only -- scanning all 768 slots for ties cost more than the fits --
        // and the winner's fit is not computed a second time)
        int mine = -1;
        Node mineNode;
        for (int k = tid; k < nslots; k += nt) {
            if (!bit(S.nbs[p], k)) continue;
            if (nsim(S.node[p], S.node[k]) < g.thMerge) continue;
            Node m;
            merged(S.node[p], S.node[k], m);
            S.candMse[k] = m.mse;
            S.red[PEAC_ATOMIC_ADD(&S.nCand, 1)] = (int16_t)k;
            if (mine < 0 || m.mse < S.candMse[mine] || (m.mse == S.candMse[mine] && S.node[k].seq < S.node[mine].seq)) mine = k, mineNode = m;
        }
        PEAC_SYNC();
        PEAC_LAP(1);
#if defined(__CUDA_ARCH__) && !defined(PEAC_HOST_EMULATION) && defined(PEAC_WARP_SELECT)
        // merged planes collect dozens of neighbours: the first minimum (least mse, then earliest creation -- a total order,
        // creation sequences are unique) is found by the first warp with shuffles; thread 0 then only handles exact ties
        int warpBest = -1, warpTies = 0;
        if (tid < 32) {
            const int nc = S.nCand;
            int best = -1;
            double bm = 0;
            int bs = 0;
            for (int t = tid; t < nc; t += 32) {
                const int k = S.red[t];
                const double m = S.candMse[k];
                const int sq = S.node[k].seq;
                if (best < 0 || m < bm || (m == bm && sq < bs)) best = k, bm = m, bs = sq;
            }
            for (int o = 16; o > 0; o >>= 1) {
                const int ob = __shfl_xor_sync(0xffffffffu, best, o);
                const double om = __shfl_xor_sync(0xffffffffu, bm, o);
                const int os = __shfl_xor_sync(0xffffffffu, bs, o);
                if (ob >= 0 && (best < 0 || om < bm || (om == bm && os < bs))) best = ob, bm = om, bs = os;
            }
            int ties = 0;  // candidates sharing the minimal mse (the winner included)
            for (int t = tid; t < nc; t += 32) ties += S.candMse[S.red[t]] == bm;
            for (int o = 16; o > 0; o >>= 1) ties += __shfl_xor_sync(0xffffffffu, ties, o);
            warpBest = best, warpTies = ties;
        }
#endif
        if (tid == 0) {
            // :1064-1072 in neighbour order (creation sequence): the first minimum wins; on an exact tie the reference
            // replaces the candidate iff cand->N < merge->mse
            int best = -1;
            const int nc = S.nCand;
#if defined(__CUDA_ARCH__) && !defined(PEAC_HOST_EMULATION) && defined(PEAC_WARP_SELECT)
            best = warpBest;
            const bool mayTie = warpTies > 1;
#else
            const bool mayTie = true;
            for (int t = 0; t < nc; t++) {
                const int k = S.red[t];
                if (best < 0 || S.candMse[k] < S.candMse[best] || (S.candMse[k] == S.candMse[best] && S.node[k].seq < S.node[best].seq))
                    best = k;
            }
#endif
            S.nCand = 0;
            if (best >= 0 && mayTie) {
                const double mn = S.candMse[best];
                // cand->N = N(p) + N(neighbour) > N(p): the tie rule can only fire when the tied mse exceeds N(p)
                if (mn > (double)S.node[p].N) {
                    int lastSeq = S.node[best].seq;
                    for (;;) {  // walk the tie group in creation order (over the step's candidates: a handful)
                        int nx = -1;
                        for (int t = 0; t < nc; t++) {
                            const int k = S.red[t];
                            if (S.candMse[k] == mn && S.node[k].seq > lastSeq && (nx < 0 || S.node[k].seq < S.node[nx].seq)) nx = k;
                        }
                        if (nx < 0) break;
                        lastSeq = S.node[nx].seq;
                        if ((double)(S.node[p].N + S.node[best].N) < mn) best = nx;
                    }
                }
            }
            S.curNb = best;
        }
        PEAC_SYNC();
        PEAC_LAP(2);
        {   // the thread that evaluated the winning neighbour already holds the merged node (a second fit would repeat the
            // same arithmetic): it publishes it and takes the merge decision; without a candidate thread 0 decides
            const int best = S.curNb;
            if ((best >= 0 && (best % nt) == tid) || (best < 0 && tid == 0)) {
                if (best >= 0) {
                    if (mine == best)
                        S.tmp = mineNode;
                    else
                        merged(S.node[p], S.node[best], S.tmp);  // the tie walk picked a slot that was not this thread's best
                }
                if (best >= 0 && S.tmp.mse < t_mse_merge(S.tmp.center[2])) {
                    S.decision = 1;
                    ds_union(S, S.node[p].rid, S.node[best].rid);  // mergeNbsFrom :383
                } else {
                    S.decision = 0;
                    if (S.node[p].N >= MIN_SUPPORT) {
                        if (S.nExtracted < MAXPL)
                            S.extracted[S.nExtracted++] = (int16_t)p;
                        else
                            S.error = PEAC_ERR_PLANES;
                    }
                }
            }
        }
        PEAC_SYNC();
        PEAC_LAP(3);
        if (S.decision) {
            const int nb = S.curNb;
            for (int w = tid; w < SH::WORDS; w += nt) {  // union of the two neighbour sets without the two nodes themselves
                uint32_t m = S.nbs[p][w] | S.nbs[nb][w];
                if ((p >> 5) == w) m &= ~(1u << (p & 31));
                if ((nb >> 5) == w) m &= ~(1u << (nb & 31));
                S.tmpMask[w] = m;
            }
            PEAC_SYNC();
            // disconnectAllNbs of both, then the merged node (in p's slot) becomes a neighbour of the union; the two rows and
            // the node record are written word by word by as many threads
            for (int k = tid; k < nslots; k += nt) {
                if (k != p && k != nb) {
                    clrbit(S.nbs[k], nb);
                    if (bit(S.tmpMask, k))
                        setbit(S.nbs[k], p);
                    else
                        clrbit(S.nbs[k], p);
                }
            }
            for (int w = tid; w < SH::WORDS; w += nt) S.nbs[p][w] = S.tmpMask[w], S.nbs[nb][w] = 0;
            if (tid == nt - 1) S.tmp.seq = S.seqNext++;
            PEAC_SYNC();
            PEAC_LAP(4);
            {
                const double *src = reinterpret_cast<const double *>(&S.tmp);
                double *dst = reinterpret_cast<double *>(&S.node[p]);
                for (int w = tid; w < (int)(sizeof(Node) / sizeof(double)); w += nt) dst[w] = src[w];
            }
            if (tid == 0) S.node[nb].nouse = 1, S.step++;
            PEAC_SYNC();
            if (tid == 0) heap_push(S, p);
        } else {
            for (int k = tid; k < nslots; k += nt) {  // p->disconnectAllNbs()
                if (k == p) {
                    for (int w = 0; w < SH::WORDS; w++) S.nbs[k][w] = 0;
                } else {
                    clrbit(S.nbs[k], p);
                }
            }
            if (tid == 0) S.step++;
        }
        PEAC_SYNC();
        PEAC_LAP(5);
    }
    if (tid == 0 && !sort_planes(S)) S.error = PEAC_ERR_PLANES;  // std::sort(extractedPlanes, PlaneSegSizeCmp) :1139-1141
    PEAC_SYNC();
    if (prof && tid == 0)
        for (int i = 0; i < 6; i++) prof[i] += acc[i];
#undef PEAC_LAP
}

PEAC_HD int nbs4(int i, int j, int Hh, int Ww, int out[4]) {  // getValid4Neighbor :388-400
    const int id = i * Ww + j;
    int c = 0;
    if (j > 0) out[c++] = id - 1;
    if (j < Ww - 1) out[c++] = id + 1;
    if (i > 0) out[c++] = id - Ww;
    if (i < Hh - 1) out[c++] = id + Ww;
    return c;
}
PEAC_HD int block_of(const Geo &g, int px, int py) {  // getBlockIdx :408-415
    const int by = py / WIN, bx = px / WIN;
    return (by < g.Nh && bx < g.Nw) ? by * g.Nw + bx : -1;
}

// region-grow queue entry: pixel | plane id << 20
PEAC_HD uint32_t rf_pack(int pix, int plid) { return (uint32_t)pix | ((uint32_t)plid << 20); }

// floodFill :422-471 as the reference's FIFO, on thread 0 (Geo::floodSerial)
template <class SH>
PEAC_D void flood_serial(SH &S, const Geo &g, const uint16_t *depth, int32_t *membership, const Flood &F, int tid) {
    if (tid != 0 || S.error != PEAC_OK) return;
    int qn = S.curP;
    for (int k = 0; k < qn; ++k) {
        const uint32_t ent = F.rfq[k];
        const int sIdx = (int)(ent & 0xfffffu), plid = (int)(ent >> 20);
        const int seedy = sIdx / g.W2, seedx = sIdx - seedy * g.W2;
        const Node &pl = S.node[S.extracted[plid]];
        int q[4];
        const int nn = nbs4(seedy, seedx, g.H2, g.W2, q);
        for (int it = 0; it < nn; ++it) {
            const int cIdx = q[it];
            int32_t trail = membership[cIdx];
            if (trail <= -6) continue;
            if (trail >= 0 && trail == plid) continue;
            const int cy = cIdx / g.W2, cx = cIdx - cy * g.W2;
            const int blkid = block_of(g, cx, cy);
            if (blkid >= 0 && S.blkMap[blkid] >= 0) continue;
            double pt[3];
            float cdist = -1;
            bool near = false;
            if (point(g, depth, cy, cx, pt)) {
                const double sd = (pl.normal[0] * (pt[0] - pl.center[0]) + pl.normal[1] * (pt[1] - pl.center[1])) +
                                  pl.normal[2] * (pt[2] - pl.center[2]);  // signedDist :355-359
                cdist = (float)fabs(sd);
                near = (double)cdist * (double)cdist < 9 * pl.mse + 1e-5;  // std::pow(float, 2): evaluated in double
            }
            if (near) {
                if (trail >= 0) {
                    const int a = S.extracted[trail], b = S.extracted[plid];
                    if (nsim(pl, S.node[a]) >= g.thRefine) setbit(S.nbs[a], b), setbit(S.nbs[b], a);  // n_pl.connect(pl)
                }
                if (cdist < F.distMap[cIdx]) {
                    membership[cIdx] = plid;
                    F.distMap[cIdx] = cdist;
                    if (qn < F.rfqCap)
                        F.rfq[qn++] = rf_pack(cIdx, plid);
                    else
                        S.error = PEAC_ERR_QUEUE;
                } else if (trail < 0) {
                    membership[cIdx] = trail - 1;
                }
            } else if (trail < 0) {
                membership[cIdx] = trail - 1;
            }
        }
        if (S.error != PEAC_OK) break;
    }
}

// floodFill :422-471 level by level (see the header comment); exactly the FIFO's result
template <class SH>
PEAC_D void flood_levels(SH &S, const Geo &g, const uint16_t *depth, int32_t *membership, const Flood &F, int tid, int nt) {
    const int npix = g.W2 * g.H2;
    for (int px = tid; px < npix; px += nt) F.own[px] = 0x7fffffff;
    if (tid == 0) S.lvlBegin = 0, S.lvlEnd = S.curP;
    PEAC_SYNC();
    int *scan = (int *)S.candMse;  // nt + 1 ints
    for (;;) {
        const int begin = S.lvlBegin, L = S.lvlEnd - begin, nv = 4 * L;
        if (L <= 0 || S.error != PEAC_OK) break;  // uniform: written by thread 0 before the last barrier
        if (nv > F.visCap) {
            if (tid == 0) S.error = PEAC_ERR_QUEUE;
            break;
        }
        // (a) the state-independent part of every visit (entry k, neighbour it); four visits per thread at a time so that
        // their queue and depth loads are in flight together (one CTA per frame: nothing else hides the L2 round trips)
        for (int v0 = tid; v0 < nv; v0 += 4 * nt) {
            int cc[4], plid[4];
            uint16_t raw[4];
            PEAC_UNROLL
            for (int u = 0; u < 4; u++) {
                const int v = v0 + u * nt;
                cc[u] = -1, plid[u] = 0, raw[u] = 0;
                if (v < nv) {
                    const uint32_t ent = F.rfq[begin + (v >> 2)];
                    const int sIdx = (int)(ent & 0xfffffu), it = v & 3;
                    plid[u] = (int)(ent >> 20);
                    const int seedy = sIdx / g.W2, seedx = sIdx - seedy * g.W2;
                    int q[4];
                    const int nn = nbs4(seedy, seedx, g.H2, g.W2, q);
                    if (it < nn) {
                        const int c = q[it];
                        const int cy = c / g.W2, cx = c - cy * g.W2;
                        const int blkid = block_of(g, cx, cy);
                        if (!(blkid >= 0 && S.blkMap[blkid] >= 0)) {  // inside a kept block: never touched
                            cc[u] = c;
                            raw[u] = depth[(size_t)(2 * cy) * g.dstride + 2 * cx];
                        }
                    }
                }
            }
            PEAC_UNROLL
            for (int u = 0; u < 4; u++) {
                const int v = v0 + u * nt;
                if (v >= nv) continue;
                float cdist = -1;
                uint8_t fl = 0;
                if (cc[u] >= 0) {
                    fl = 2;
                    const double z = (double)raw[u] * (double)g.factor;  // point(): ImagePointCloud::get on readDepthImage's cloud
                    if (z != 0) {
                        const int cy = cc[u] / g.W2, cx = cc[u] - cy * g.W2;
                        const double pt[3] = {((double)(2 * cx) - (double)g.cx) * z / (double)g.fx,
                                              ((double)(2 * cy) - (double)g.cy) * z / (double)g.fy, z};
                        const Node &pl = S.node[S.extracted[plid[u]]];
                        const double sd = (pl.normal[0] * (pt[0] - pl.center[0]) + pl.normal[1] * (pt[1] - pl.center[1])) +
                                          pl.normal[2] * (pt[2] - pl.center[2]);
                        cdist = (float)fabs(sd);
                        if ((double)cdist * (double)cdist < 9 * pl.mse + 1e-5) fl = 3;
                    }
                }
                F.visC[v] = cc[u], F.visDist[v] = cdist, F.visFlag[v] = fl;
            }
        }
        PEAC_SYNC();
        // (b) resolve in visit order per target pixel: every round the earliest pending visit of each pixel
        for (;;) {
            if (tid == 0) S.pending = 0;
            for (int v0 = tid; v0 < nv; v0 += 4 * nt) {  // four at a time: flag and target loads in flight together
                int tc[4];
                PEAC_UNROLL
                for (int u = 0; u < 4; u++) {
                    const int v = v0 + u * nt;
                    tc[u] = (v < nv && (F.visFlag[v] & 2)) ? F.visC[v] : -1;
                }
                PEAC_UNROLL
                for (int u = 0; u < 4; u++)
                    if (tc[u] >= 0) PEAC_ATOMIC_MIN(&F.own[tc[u]], v0 + u * nt);
            }
            PEAC_SYNC();
            for (int v0 = tid; v0 < nv; v0 += 4 * nt) {
              // four visits at a time: flags and targets, then the owners, then the state of the resolved targets are loaded
              // together (the visits a round resolves have distinct target pixels, so they do not see each other's writes)
              uint8_t flb[4];
              int cb[4], ownb[4], plidb[4];
              int32_t trailb[4];
              float distb[4];
              PEAC_UNROLL
              for (int u = 0; u < 4; u++) {
                  const int v = v0 + u * nt;
                  flb[u] = v < nv ? F.visFlag[v] : 0;
                  cb[u] = (flb[u] & 2) ? F.visC[v] : -1;
              }
              PEAC_UNROLL
              for (int u = 0; u < 4; u++) ownb[u] = cb[u] >= 0 ? PEAC_LOAD(&F.own[cb[u]]) : -1;
              PEAC_UNROLL
              for (int u = 0; u < 4; u++) {
                  const int v = v0 + u * nt;
                  const bool mineNow = cb[u] >= 0 && ownb[u] == v;
                  plidb[u] = mineNow ? (int)(F.rfq[begin + (v >> 2)] >> 20) : 0;
                  trailb[u] = mineNow ? membership[cb[u]] : 0;
                  distb[u] = mineNow ? F.distMap[cb[u]] : 0.f;
              }
              PEAC_UNROLL
              for (int u = 0; u < 4; u++) {
                const int v = v0 + u * nt;
                const uint8_t fl = flb[u];
                if (!(fl & 2)) continue;
                const int c = cb[u];
                if (ownb[u] != v) {
                    PEAC_STORE(&S.pending, 1);
                    continue;
                }
                const int plid = plidb[u];
                const int32_t trail = trailb[u];
                uint8_t out = fl & 1;  // pending cleared
                if (trail <= -6 || (trail >= 0 && trail == plid)) {
                } else if (fl & 1) {
                    if (trail >= 0) {
                        const int a = S.extracted[trail], b = S.extracted[plid];
                        if (nsim(S.node[b], S.node[a]) >= g.thRefine) {
                            PEAC_ATOMIC_OR(&S.nbs[a][b >> 5], 1u << (b & 31));
                            PEAC_ATOMIC_OR(&S.nbs[b][a >> 5], 1u << (a & 31));
                        }
                    }
                    const float cdist = F.visDist[v];
                    if (cdist < distb[u]) {
                        membership[c] = plid;
                        F.distMap[c] = cdist;
                        out |= 4;
                    } else if (trail < 0) {
                        membership[c] = trail - 1;
                    }
                } else if (trail < 0) {
                    membership[c] = trail - 1;
                }
                F.visFlag[v] = out;
                PEAC_STORE(&F.own[c], 0x7fffffff);
              }
            }
            PEAC_SYNC();
            if (!S.pending) break;
            PEAC_SYNC();  // everyone has read `pending` before thread 0 clears it
        }
        // (c) the next level, in visit order: chunked scan of the push flags
        const int chunk = (nv + nt - 1) / nt, lo = tid * chunk < nv ? tid * chunk : nv, hi = lo + chunk < nv ? lo + chunk : nv;
        int cnt = 0;
        for (int v = lo; v < hi; v++) cnt += (F.visFlag[v] >> 2) & 1;
        scan[tid] = cnt;
        PEAC_SYNC();
        if (tid == 0) {
            int acc = 0;
            for (int t = 0; t < nt; t++) {
                const int c = scan[t];
                scan[t] = acc;
                acc += c;
            }
            if (S.lvlEnd + acc > F.rfqCap) S.error = PEAC_ERR_QUEUE;
            scan[nt] = acc;
        }
        PEAC_SYNC();
        if (S.error == PEAC_OK) {
            uint32_t *o = F.rfq + S.lvlEnd + scan[tid];
            for (int v = lo; v < hi; v++)
                if (F.visFlag[v] & 4) *o++ = rf_pack(F.visC[v], (int)(F.rfq[begin + (v >> 2)] >> 20));
        }
        PEAC_SYNC();
        if (tid == 0) S.lvlBegin = S.lvlEnd, S.lvlEnd += scan[nt];
        PEAC_SYNC();
    }
}

// One frame.  blocks / seed / edges: the pre-stage outputs of the frame (centre, normal, mse, N per block; node mask;
// edge mask bit0=left,1=right,2=up,3=down).  membership: H2*W2 int32 out.  F: per-frame scratch of the region grow in
// global memory.  planes: <= planeCap records out; *planeCount out.
template <class SH, typename BlockStat>
PEAC_D void frame(SH &S, const Geo &g, const uint16_t *depth, const BlockStat *blocks, const uint8_t *seed, const uint8_t *edges,
                  int32_t *membership, const Flood &F, PlaneOut *planes, int planeCap, int32_t *planeCount, int32_t *errorOut, int tid,
                  int nt) {
    float *const distMap = F.distMap;
    uint32_t *const rfq = F.rfq;
    const int rfqCap = F.rfqCap;
    const int nb = g.Nw * g.Nh, npix = g.W2 * g.H2;
    PEAC_STAMP(F, 0, tid);
    // ---- initGraph's nodes and edges from the pre-stage (AHCPlaneFitter.hpp:756-928)
    for (int b = tid; b < nb; b += nt) {
        Node &n = S.node[b];
        n.N = blocks[b].N, n.rid = b, n.seq = b, n.nouse = blocks[b].nouse ? 1 : 0, n.mse = blocks[b].mse;
        for (int k = 0; k < 3; k++) n.center[k] = blocks[b].center[k], n.normal[k] = blocks[b].normal[k];
        if (seed[b])
            block_sums(g, depth, b, n.s);
        else
            for (int k = 0; k < 9; k++) n.s[k] = 0;
        S.parent[b] = (int16_t)b, S.dsize[b] = 1;
        uint32_t *m = S.nbs[b];
        for (int w = 0; w < SH::WORDS; w++) m[w] = 0;
        const int e = edges[b];
        if (e & 1) setbit(m, b - 1);
        if (e & 2) setbit(m, b + 1);
        if (e & 4) setbit(m, b - g.Nw);
        if (e & 8) setbit(m, b + g.Nw);
    }
    if (tid == 0) S.heapN = 0, S.nExtracted = 0, S.seqNext = nb, S.step = 0, S.error = PEAC_OK, S.nCand = 0;
    PEAC_SYNC();
    if (tid == 0)
        for (int b = 0; b < nb; b++)
            if (seed[b]) heap_push(S, b);  // minQ.push in block order (:779)
    PEAC_SYNC();
    PEAC_STAMP(F, 1, tid);
    if (F.prof && tid == 0)
        for (int i = 8; i < 16; i++) F.prof[i] = 0;
    cluster(S, g, nb, tid, nt, F.prof ? F.prof + 8 : nullptr);
    PEAC_STAMP(F, 2, tid);

    // ---- refineDetails :294-374.  findBlockMembership(isValidExtractedPlane) :480-582, ERODE_ALL_BORDER
    for (int i = tid; i < MAXPL; i += nt) S.valid[i] = 0, S.count[i] = 0, S.plidmap[i] = -1;
    PEAC_SYNC();
    for (int b = tid; b < nb; b += nt) {
        const int i = b / g.Nw, j = b - i * g.Nw;
        const int setid = ds_find_ro(S, b);
        int bm = -1;
        if ((int)S.dsize[setid] * (WIN * WIN) >= MIN_SUPPORT) {
            int q[4];
            const int nn = nbs4(i, j, g.Nh, g.Nw, q);
            bool same = true;
            for (int k = 0; k < nn; k++)
                if (ds_find_ro(S, q[k]) != setid) {
                    same = false;
                    break;
                }
            int plid = 0;  // rid2plid[setid]: std::map::operator[] yields 0 for a root that is no plane's rid
            for (int e = 0; e < S.nExtracted; e++)
                if (S.node[S.extracted[e]].rid == setid) {
                    plid = e;
                    break;
                }
            if (same) {
                bm = plid;
                if (plid < S.nExtracted) PEAC_STORE_FLAG(&S.valid[plid]);  // every writer stores the same 1
            }
        }
        S.blkMap[b] = (int16_t)bm;
    }
    PEAC_SYNC();
    int *cnt = (int *)S.candMse;  // per-block number of region-grow seeds, then their exclusive scan
    for (int b = tid; b < nb; b += nt) {
        const int i = b / g.Nw, j = b - i * g.Nw, bm = S.blkMap[b];
        int c = 0;
        if (bm < 0) {
            if (i > 0 && S.blkMap[b - g.Nw] >= 0) c += WIN - 1;
            if (j > 0 && S.blkMap[b - 1] >= 0) c += WIN - 1;
        } else {
            if (i > 0 && S.blkMap[b - g.Nw] != bm) c += WIN - 1;
            if (j > 0 && S.blkMap[b - 1] != bm) c += WIN - 1;
        }
        cnt[b] = c;
    }
    for (int px = tid; px < npix; px += nt) {  // membershipImg.setTo(-1) + the block fills; distMap = FLT_MAX
        const int y = px / g.W2, x = px - y * g.W2, b = block_of(g, x, y);
        membership[px] = (b >= 0 && S.blkMap[b] >= 0) ? S.blkMap[b] : -1;
        distMap[px] = FLT_MAX;
    }
    PEAC_SYNC();
    if (tid == 0) {
        int acc = 0;
        for (int b = 0; b < nb; b++) {
            const int c = cnt[b];
            cnt[b] = acc;
            acc += c;
        }
        S.curP = acc;  // queue length so far
        if (acc > rfqCap) S.error = PEAC_ERR_QUEUE;
    }
    PEAC_SYNC();
    if (S.error == PEAC_OK)
        for (int b = tid; b < nb; b += nt) {  // the seeds of the region grow, in the reference's order (:537-580)
            const int i = b / g.Nw, j = b - i * g.Nw, bm = S.blkMap[b];
            uint32_t *o = rfq + cnt[b];
            if (bm < 0) {
                if (i > 0 && S.blkMap[b - g.Nw] >= 0) {
                    const int s = (i * WIN - 1) * g.W2 + j * WIN, pl = S.blkMap[b - g.Nw];
                    for (int k = 1; k < WIN; ++k) *o++ = rf_pack(s + k, pl);
                }
                if (j > 0 && S.blkMap[b - 1] >= 0) {
                    const int s = (i * WIN) * g.W2 + j * WIN - 1, pl = S.blkMap[b - 1];
                    for (int k = 0; k < WIN - 1; ++k) *o++ = rf_pack(s + k * g.W2, pl);
                }
            } else {
                if (i > 0 && S.blkMap[b - g.Nw] != bm) {
                    const int s = (i * WIN) * g.W2 + j * WIN;
                    for (int k = 0; k < WIN - 1; ++k) *o++ = rf_pack(s + k, bm);
                }
                if (j > 0 && S.blkMap[b - 1] != bm) {
                    const int s = (i * WIN) * g.W2 + j * WIN;
                    for (int k = 1; k < WIN; ++k) *o++ = rf_pack(s + k * g.W2, bm);
                }
            }
        }
    PEAC_SYNC();

    // ---- floodFill :422-471
    PEAC_STAMP(F, 3, tid);
    if (g.floodSerial)
        flood_serial(S, g, depth, membership, F, tid);
    else
        flood_levels(S, g, depth, membership, F, tid, nt);
    PEAC_SYNC();
    PEAC_STAMP(F, 4, tid);

    // ---- "try to merge one last time" :312-320: the valid planes re-enter the queue in plane order
    if (tid == 0) {
        S.nOld = S.nExtracted;
        for (int i = 0; i < S.nOld; i++) S.oldPl[i] = S.extracted[i];
        S.nExtracted = 0, S.heapN = 0;
        for (int i = 0; i < S.nOld; i++)
            if (S.valid[i]) heap_push(S, S.oldPl[i]);
    }
    PEAC_SYNC();
    cluster(S, g, nb, tid, nt);
    PEAC_STAMP(F, 5, tid);
    if (tid == 0) {  // plidmap :322-337 (a merged plane lives in the slot of one of its parts; ds roots are unaffected)
        for (int i = 0; i < S.nOld; i++) {
            if (!S.valid[i]) continue;
            const int root = ds_find(S, S.node[S.oldPl[i]].rid);
            for (int j = 0; j < S.nExtracted; j++)
                if (root == S.node[S.extracted[j]].rid) {
                    S.plidmap[i] = (int16_t)j;
                    break;
                }
        }
    }
    PEAC_SYNC();
    for (int px = tid; px < npix; px += nt) {  // :352-366
        const int plid = membership[px];
        if (plid >= 0 && plid < MAXPL && S.plidmap[plid] >= 0) {
            const int np = S.plidmap[plid];
            membership[px] = np;
            PEAC_ATOMIC_ADD(&S.count[np], 1);
        }
    }
    PEAC_SYNC();
    for (int i = tid; i < S.nExtracted && i < planeCap; i += nt) {
        const Node &n = S.node[S.extracted[i]];
        PlaneOut &o = planes[i];
        for (int k = 0; k < 3; k++) o.normal[k] = n.normal[k], o.center[k] = n.center[k];
        o.N = n.N, o.rid = n.rid, o.vertices = S.count[i], o.pad = 0;
    }
    if (tid == 0) *planeCount = S.nExtracted, *errorOut = S.error;
    PEAC_STAMP(F, 6, tid);
    if (F.prof && tid == 0) F.prof[7] = S.step;
}

}  // namespace peac
