// peac_host.cpp -- TEST HARNESS (tests/test_peac_host_emulation.py): manhattanslam_b200/csrc/peac_frame.cuh -- the code the
// k_peac_frame kernel runs -- compiled for the host with one "thread" and no-op barriers, so that the f2 algorithm
// (slot reuse, mask adjacency, heap order, region-grow seeds, flood fill, final merge, plane-id remap) can be checked
// against the oracle on machines without a GPU.  Never linked into the product library.
#define PEAC_HOST_EMULATION
#include "../../manhattanslam_b200/csrc/peac_frame.cuh"

#include <cstring>
#include <vector>

struct BlockStat {  // = msl_block_stat / orc_block_stat
    double center[3], normal[3], mse, curvature;
    int32_t N, nouse;
};

extern "C" int peac_host_frame(const uint16_t *depth, int w, int h, int dstride_px, float fx, float fy, float cx, float cy,
                               float factor, const BlockStat *blocks, const uint8_t *seed, const uint8_t *edges,
                               int32_t *membership, peac::PlaneOut *planes, int cap, int32_t *error, int rfq_cap, int flood_serial) {
    peac::Geo g;
    g.W2 = (int)std::ceil(w / 2.0), g.H2 = (int)std::ceil(h / 2.0);
    g.Nw = g.W2 / peac::WIN, g.Nh = g.H2 / peac::WIN;
    if (g.Nw * g.Nh > peac::MAXB_BIG) return -1;
    g.dstride = dstride_px;
    g.fx = fx, g.fy = fy, g.cx = cx, g.cy = cy, g.factor = factor;
    g.thMerge = std::cos(60.0 * M_PI / 180.0), g.thRefine = std::cos(30.0 * M_PI / 180.0);
    g.floodSerial = flood_serial;
    const size_t npix = (size_t)g.W2 * g.H2;
    std::vector<float> dist(npix), visDist(4 * npix);
    std::vector<uint32_t> rfq((size_t)rfq_cap);
    std::vector<int> own(npix), visC(4 * npix);
    std::vector<uint8_t> visFlag(4 * npix);
    peac::Flood F{dist.data(), rfq.data(), rfq_cap, own.data(), visC.data(), visDist.data(), visFlag.data(), (int)(4 * npix)};
    int32_t count = 0;
    if (g.Nw * g.Nh <= peac::MAXB) {  // the shared-memory instance of the kernel
        std::vector<peac::Shared> S(1);
        peac::frame(S[0], g, depth, blocks, seed, edges, membership, F, planes, cap, &count, error, 0, 1);
    } else {  // the global-memory instance
        std::vector<peac::SharedBig> S(1);
        peac::frame(S[0], g, depth, blocks, seed, edges, membership, F, planes, cap, &count, error, 0, 1);
    }
    return count;
}

extern "C" int peac_host_shared_bytes() { return (int)sizeof(peac::Shared); }
extern "C" int peac_host_shared_big_bytes() { return (int)sizeof(peac::SharedBig); }
