// PlaneExtractor_msl.cpp -- PlaneDetection::readDepthImage on the B200 front-end (drop-in for the definition
// in src/PlaneExtractor.cpp:44-76).  Frame::ExtractPlanes (src/Frame.cc:605-609) is unchanged.  The call also
// returns peac's per-block statistics / seeds / edges; msl_prestage_of() hands them to a PlaneFitter whose
// initGraph consumes them instead of re-scanning the cloud (optional step).
#include <mutex>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "PlaneExtractor.h"
#include "msl_frontend.h"

namespace {
struct Pre {
    msl_plane *h = nullptr;
    int w = 0, hgt = 0;
    std::vector<msl_block_stat> blocks;
    std::vector<uint8_t> seed, edges;
};
std::mutex g_mu;
std::unordered_map<const PlaneDetection *, Pre> g_pre;  // one PlaneDetection per Frame thread
}  // namespace

const std::vector<msl_block_stat> *msl_prestage_of(const PlaneDetection *pd, const uint8_t **seed, const uint8_t **edges) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_pre.find(pd);
    if (it == g_pre.end()) return nullptr;
    *seed = it->second.seed.data(), *edges = it->second.edges.data();
    return &it->second.blocks;
}

bool PlaneDetection::readDepthImage(const cv::Mat depthImg, const cv::Mat &K, const float &depthMapFactor) {
    cv::Mat depth_img = depthImg;
    if (depth_img.empty() || depth_img.depth() != CV_16U) {
        std::cout << "WARNING: cannot read depth image. No such a file, or the image format is not 16UC1" << std::endl;
        return false;
    }
    const int W2 = (int)ceil(depthImg.cols / 2.0), H2 = (int)ceil(depthImg.rows / 2.0);
    cloud.vertices.resize((size_t)H2 * W2);
    cloud.verticesColour.resize((size_t)H2 * W2);
    cloud.w = W2, cloud.h = H2;
    seg_img_ = cv::Mat(H2, W2, CV_8UC3);
    Pre *p;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        p = &g_pre[this];
    }
    if (!p->h || p->w != depth_img.cols || p->hgt != depth_img.rows) {
        if (p->h) msl_plane_destroy(p->h);
        if (msl_plane_create(depth_img.cols, depth_img.rows, 1, 0, &p->h) != MSL_OK) throw std::runtime_error(msl_last_error());
        p->w = depth_img.cols, p->hgt = depth_img.rows;
        const size_t nb = (size_t)(W2 / 10) * (H2 / 10);
        p->blocks.resize(nb), p->seed.resize(nb), p->edges.resize(nb);
    }
    const float Kf[4] = {K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2)};
    static_assert(sizeof(VertexType) == 3 * sizeof(double), "Eigen::Vector3d is three packed doubles");
    if (msl_plane_prestage(p->h, depth_img.ptr<uint16_t>(), (int)(depth_img.step / 2), (size_t)(depth_img.step / 2) * depth_img.rows,
                           1, Kf, depthMapFactor, reinterpret_cast<double *>(cloud.vertices.data()), p->blocks.data(),
                           p->seed.data(), p->edges.data()) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    for (int i = 0, v = 0; i < depth_img.rows; i += 2)  // colours are only read back for plane members (src/Frame.cc:616-621)
        for (int j = 0; j < depth_img.cols; j += 2, v++) cloud.verticesColour[v] = color_img_.at<cv::Vec3b>(i, j);
    return true;
}
