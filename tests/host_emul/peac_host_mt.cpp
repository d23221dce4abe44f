// peac_host_mt.cpp -- TEST HARNESS (tests/test_peac_host_emulation.py): manhattanslam_b200/csrc/peac_frame.cuh run by several
// REAL threads that play the threads of one CTA, PEAC_SYNC() = a pthread barrier, built with -fsanitize=thread.  A missing
// barrier between two phases of the kernel shows up as a ThreadSanitizer data-race report (and usually as a wrong result).
// usage: peac_host_mt <threads> <in.bin> <out.bin> [flood_serial]; in.bin = header {w, h, cap} as int32, {fx, fy, cx, cy, factor} as float,
// depth u16[w*h], blocks 72 B x nb, seed u8[nb], edges u8[nb]; out.bin = count i32, error i32, membership i32[h2*w2],
// planes 64 B x cap.  Never linked into the product library.
#define PEAC_HOST_EMULATION_MT
#include <pthread.h>

#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../manhattanslam_b200/csrc/peac_frame.cuh"

static pthread_barrier_t g_bar;
void peac_emu_sync() { pthread_barrier_wait(&g_bar); }

struct BlockStat {
    double center[3], normal[3], mse, curvature;
    int32_t N, nouse;
};

int main(int argc, char **argv) {
    if (argc != 4 && argc != 5) return 2;
    const int nt = atoi(argv[1]);
    FILE *f = fopen(argv[2], "rb");
    if (!f) return 2;
    int32_t hd[3];
    float kf[5];
    if (fread(hd, 4, 3, f) != 3 || fread(kf, 4, 5, f) != 5) return 2;
    const int w = hd[0], h = hd[1], cap = hd[2];
    peac::Geo g;
    g.W2 = (int)std::ceil(w / 2.0), g.H2 = (int)std::ceil(h / 2.0);
    g.Nw = g.W2 / peac::WIN, g.Nh = g.H2 / peac::WIN;
    g.dstride = w, g.fx = kf[0], g.fy = kf[1], g.cx = kf[2], g.cy = kf[3], g.factor = kf[4];
    g.thMerge = std::cos(60.0 * M_PI / 180.0), g.thRefine = std::cos(30.0 * M_PI / 180.0);
    g.floodSerial = argc == 5 ? atoi(argv[4]) : 0;
    const int nb = g.Nw * g.Nh, npix = g.W2 * g.H2;
    if (nb > peac::MAXB_BIG) return 3;
    std::vector<uint16_t> depth((size_t)w * h);
    std::vector<BlockStat> blocks(nb);
    std::vector<uint8_t> seed(nb), edges(nb);
    if (fread(depth.data(), 2, depth.size(), f) != depth.size() || fread(blocks.data(), sizeof(BlockStat), nb, f) != (size_t)nb ||
        fread(seed.data(), 1, nb, f) != (size_t)nb || fread(edges.data(), 1, nb, f) != (size_t)nb)
        return 2;
    fclose(f);
    std::vector<peac::Shared> S(nb <= peac::MAXB ? 1 : 0);
    std::vector<peac::SharedBig> SB(nb <= peac::MAXB ? 0 : 1);
    std::vector<int32_t> mem(npix);
    std::vector<float> dist(npix);
    std::vector<uint32_t> rfq((size_t)4 * npix);
    std::vector<int> own(npix), visC((size_t)4 * npix);
    std::vector<float> visDist((size_t)4 * npix);
    std::vector<uint8_t> visFlag((size_t)4 * npix);
    peac::Flood F{dist.data(), rfq.data(), (int)rfq.size(), own.data(), visC.data(), visDist.data(), visFlag.data(), 4 * npix};
    std::vector<peac::PlaneOut> planes(cap);
    memset(planes.data(), 0, sizeof(peac::PlaneOut) * cap);
    int32_t count = 0, error = 0;
    pthread_barrier_init(&g_bar, nullptr, nt);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&, t] {
            if (nb <= peac::MAXB)
                peac::frame(S[0], g, depth.data(), blocks.data(), seed.data(), edges.data(), mem.data(), F, planes.data(), cap, &count,
                            &error, t, nt);
            else
                peac::frame(SB[0], g, depth.data(), blocks.data(), seed.data(), edges.data(), mem.data(), F, planes.data(), cap, &count,
                            &error, t, nt);
        });
    for (auto &t : th) t.join();
    f = fopen(argv[3], "wb");
    fwrite(&count, 4, 1, f), fwrite(&error, 4, 1, f);
    fwrite(mem.data(), 4, npix, f), fwrite(planes.data(), sizeof(peac::PlaneOut), cap, f);
    fclose(f);
    return 0;
}
