// PlaneExtractor_msl.cpp -- PlaneDetection::readDepthImage and ::runPlaneDetection on the B200 front-end (drop-in for the
// definitions in src/PlaneExtractor.cpp:44-82).  Frame::ExtractPlanes (src/Frame.cc:605-652) and Tracking's use of
// plane_filter.membershipImg (src/Tracking.cc:228,497) are unchanged.  readDepthImage also keeps peac's per-block
// statistics / seeds / edges (msl_prestage_of()); runPlaneDetection runs the whole fitter on the device
// (msl_plane_detect: ahCluster + refineDetails) and fills what the callers read: plane_filter.membershipImg,
// plane_filter.extractedPlanes[i]->normal / center / N / rid, plane_vertices_, plane_num_.  seg_img_ (the debug colouring)
// is not produced.  Frames of more than 3072 blocks (beyond 1280x960) fall outside msl_plane_detect: keep the reference's
// definition there.
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
    cv::Mat depth;  // the image of the last readDepthImage (a header sharing the caller's buffer, as in the reference)
    float K[4] = {0, 0, 0, 0}, factor = 0;
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
    p->depth = depth_img, p->factor = depthMapFactor;
    for (int k = 0; k < 4; k++) p->K[k] = Kf[k];
    static_assert(sizeof(VertexType) == 3 * sizeof(double), "Eigen::Vector3d is three packed doubles");
    if (msl_plane_prestage(p->h, depth_img.ptr<uint16_t>(), (int)(depth_img.step / 2), (size_t)(depth_img.step / 2) * depth_img.rows,
                           1, Kf, depthMapFactor, reinterpret_cast<double *>(cloud.vertices.data()), p->blocks.data(),
                           p->seed.data(), p->edges.data()) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    for (int i = 0, v = 0; i < depth_img.rows; i += 2)  // colours are only read back for plane members (src/Frame.cc:616-621)
        for (int j = 0; j < depth_img.cols; j += 2, v++) cloud.verticesColour[v] = color_img_.at<cv::Vec3b>(i, j);
    return true;
}

void PlaneDetection::runPlaneDetection() {
    Pre *p;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_pre.find(this);
        if (it == g_pre.end() || it->second.depth.empty()) throw std::runtime_error("runPlaneDetection before readDepthImage");
        p = &it->second;
    }
    const int W2 = cloud.w, H2 = cloud.h, cap = 128;  // peac::MAXPL, the library's own limit (msl_plane_detect)
    cv::Mat &mem = plane_filter.membershipImg;
    mem.create(H2, W2, CV_32SC1);
    if (!mem.isContinuous()) throw std::runtime_error("membershipImg must be continuous");
    int32_t n = 0;
    std::vector<msl_plane_rec> rec(cap);
    if (msl_plane_detect(p->h, p->depth.ptr<uint16_t>(), (int)(p->depth.step / 2), (size_t)(p->depth.step / 2) * p->depth.rows, 1, p->K,
                         p->factor, mem.ptr<int32_t>(), &n, rec.data(), cap) != MSL_OK)
        throw std::runtime_error(msl_last_error());
    if (n > cap) {
        // more planes than the record array holds (never seen below 1280x960): the reference has no limit, so fail loudly rather
        // than leave labels >= cap in membershipImg without an extractedPlanes / plane_vertices_ entry
        throw std::runtime_error("runPlaneDetection: more than 128 planes extracted");
    }
    // extractedPlanes: PlaneSeg has no default constructor; build each on an empty cloud (rejected: N = 0, nouse) and
    // set the fields Frame::ExtractPlanes reads (src/Frame.cc:626-632)
    plane_filter.extractedPlanes.clear();
    const ahc::NullImage3D none;
    for (int i = 0; i < n; i++) {
        ahc::PlaneSeg::shared_ptr s(new ahc::PlaneSeg(none, rec[i].rid, 0, 0, 0, 0, plane_filter.windowWidth, plane_filter.windowHeight,
                                                      plane_filter.params));
        for (int k = 0; k < 3; k++) s->normal[k] = rec[i].normal[k], s->center[k] = rec[i].center[k];
        s->N = rec[i].N, s->rid = rec[i].rid, s->nouse = false;
        plane_filter.extractedPlanes.push_back(s);
    }
    // plane_vertices_[i] = pixels labelled i in row-major order (AHCPlaneFitter.hpp:352-363)
    plane_vertices_.assign((size_t)n, std::vector<int>());
    for (int i = 0; i < n; i++) plane_vertices_[i].reserve((size_t)rec[i].vertices);
    const int32_t *m = mem.ptr<int32_t>();
    for (int i = 0; i < W2 * H2; i++)
        if (m[i] >= 0 && m[i] < n) plane_vertices_[m[i]].push_back(i);
    plane_num_ = (int)plane_vertices_.size();
}
