// mock_abi.cpp -- TEST HARNESS (tests/test_adapters_on_mock_abi.py): the subset of include/msl_frontend.h that
// adapters/ORBextractor_msl.cc, PlaneExtractor_msl.cpp and SurfelFusion_msl.cpp call, implemented on the CPU ORACLE, so
// that the bindings' marshalling code -- which needs OpenCV / Eigen and a GPU to run for real -- can be EXECUTED on a
// CPU-only machine and compared, at the level of the reference's class surfaces, with the reference's own classes
// (oracle/_ref).  It says nothing about the kernels; it checks the glue between the reference's types and the C ABI.
// Lives under tests/ and is never linked into the product library.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "msl_frontend.h"
#include "msl_oracle.h"

static thread_local std::string g_err;
static int fail(int code, const char *what) {
    g_err = what;
    return code;
}

struct msl_orb {
    orc_orb *o;
    int w, h, cap;
};
struct msl_plane {
    int w, h;
};
struct msl_surfel_fusion {
    orc_surfel_fusion *o;
    int w, h;
    std::vector<orc_surfel> map;
    orc_surfel_mapping *mp;  // the inactive store behind msl_surfel_move_add
};

extern "C" {

const char *msl_last_error(void) { return g_err.c_str(); }

int msl_orb_create(const msl_orb_params *p, int w, int h, int max_batch, int, msl_orb **out) {
    if (!p || !out || max_batch < 1) return fail(MSL_ERR_INVALID, "msl_orb_create: bad argument");
    orc_orb *o = orc_orb_create(p->nfeatures, p->scale_factor, p->nlevels, p->ini_th_fast, p->min_th_fast);
    if (!o) return fail(MSL_ERR_INVALID, "msl_orb_create: bad parameters");
    *out = new msl_orb{o, w, h, p->nfeatures + 8 * p->nlevels + 64};
    return MSL_OK;
}
void msl_orb_destroy(msl_orb *h) {
    if (h) orc_orb_destroy(h->o), delete h;
}
int msl_orb_capacity(const msl_orb *h) { return h->cap; }
int msl_orb_extract(msl_orb *h, const uint8_t *gray, int stride, size_t frame_stride, int batch, msl_keypoint *kps, uint8_t *desc,
                    int32_t *counts) {
    static_assert(sizeof(msl_keypoint) == sizeof(orc_keypoint), "keypoint layout");
    for (int b = 0; b < batch; b++) {
        const int n = orc_orb_extract(h->o, gray + b * frame_stride, h->w, h->h, stride, (orc_keypoint *)(kps + (size_t)b * h->cap),
                                      desc + (size_t)b * h->cap * 32, h->cap);
        if (n < 0) return fail(MSL_ERR_CAPACITY, "msl_orb_extract: capacity");
        counts[b] = n;
    }
    return MSL_OK;
}

int msl_plane_create(int w, int h, int max_batch, int, msl_plane **out) {
    if (!out || max_batch < 1) return fail(MSL_ERR_INVALID, "msl_plane_create: bad argument");
    *out = new msl_plane{w, h};
    return MSL_OK;
}
void msl_plane_destroy(msl_plane *p) { delete p; }
int msl_plane_prestage(msl_plane *p, const uint16_t *depth, int dstride_px, size_t frame_stride_px, int batch, const float K[4],
                       float factor, double *cloud, msl_block_stat *blocks, uint8_t *seed, uint8_t *edges) {
    static_assert(sizeof(msl_block_stat) == sizeof(orc_block_stat), "block layout");
    const int W2 = (int)std::ceil(p->w / 2.0), H2 = (int)std::ceil(p->h / 2.0), nb = (W2 / 10) * (H2 / 10);
    for (int b = 0; b < batch; b++)
        orc_plane_prestage(depth + b * frame_stride_px, p->w, p->h, dstride_px, K[0], K[1], K[2], K[3], factor,
                           cloud ? cloud + (size_t)b * W2 * H2 * 3 : nullptr, (orc_block_stat *)(blocks ? blocks + (size_t)b * nb : nullptr),
                           seed ? seed + (size_t)b * nb : nullptr, edges ? edges + (size_t)b * nb : nullptr);
    return MSL_OK;
}
int msl_plane_detect(msl_plane *p, const uint16_t *depth, int dstride_px, size_t frame_stride_px, int batch, const float K[4],
                     float factor, int32_t *membership, int32_t *plane_count, msl_plane_rec *planes, int plane_cap) {
    const int W2 = (int)std::ceil(p->w / 2.0), H2 = (int)std::ceil(p->h / 2.0);
    for (int b = 0; b < batch; b++) {
        std::vector<double> nrm(3 * 128), cen(3 * 128);
        std::vector<int32_t> N(128), rid(128), nv(128);
        const int n = orc_plane_detect(depth + b * frame_stride_px, p->w, p->h, dstride_px, K[0], K[1], K[2], K[3], factor,
                                       membership + (size_t)b * W2 * H2, nrm.data(), cen.data(), N.data(), rid.data(), nv.data(), 128);
        plane_count[b] = n;
        for (int i = 0; i < n && i < plane_cap; i++) {
            msl_plane_rec &r = planes[(size_t)b * plane_cap + i];
            for (int k = 0; k < 3; k++) r.normal[k] = nrm[3 * i + k], r.center[k] = cen[3 * i + k];
            r.N = N[i], r.rid = rid[i], r.vertices = nv[i], r.pad = 0;
        }
    }
    return MSL_OK;
}

int msl_surfel_create(int w, int h, float fx, float fy, float cx, float cy, float fuse_far, float fuse_near, int64_t, int,
                      msl_surfel_fusion **out) {
    if (!out) return fail(MSL_ERR_INVALID, "msl_surfel_create: null out");
    *out = new msl_surfel_fusion{orc_surfel_create(w, h, fx, fy, cx, cy, fuse_far, fuse_near), w, h, {}, orc_mapping_create()};
    return MSL_OK;
}
void msl_surfel_destroy(msl_surfel_fusion *h) {
    if (h) orc_surfel_destroy(h->o), orc_mapping_destroy(h->mp), delete h;
}
int msl_surfel_move_add(msl_surfel_fusion *h, const int32_t *rem, int n_rem, const int32_t *add, int n_add, int64_t stats[3]) {
    const int64_t n0 = (int64_t)h->map.size(), cap = n0 + orc_mapping_inactive(h->mp, nullptr, 0);
    h->map.resize((size_t)cap);
    const int64_t n = orc_move_add_surfels(h->mp, h->map.data(), n0, cap, rem, n_rem, add, n_add);
    if (n < 0) return fail(MSL_ERR_STATE, "msl_surfel_move_add: pose state");
    h->map.resize((size_t)n);
    if (stats) stats[0] = stats[1] = 0, stats[2] = n;
    return MSL_OK;
}
int64_t msl_surfel_inactive_size(const msl_surfel_fusion *h) { return orc_mapping_inactive(h->mp, nullptr, 0); }
int msl_surfel_download_inactive(msl_surfel_fusion *h, msl_surfel *out, int64_t cap, int64_t *n) {
    const int64_t sz = orc_mapping_inactive(h->mp, nullptr, 0);
    if (n) *n = sz;
    if (!out) return MSL_OK;
    if (sz > cap) return fail(MSL_ERR_CAPACITY, "msl_surfel_download_inactive: capacity");
    orc_mapping_inactive(h->mp, (orc_surfel *)out, cap);
    return MSL_OK;
}
int msl_surfel_upload_map(msl_surfel_fusion *h, const msl_surfel *local, int64_t n) {
    static_assert(sizeof(msl_surfel) == sizeof(orc_surfel), "surfel layout");
    h->map.assign((const orc_surfel *)local, (const orc_surfel *)local + n);
    return MSL_OK;
}
int msl_surfel_download_map(msl_surfel_fusion *h, msl_surfel *local, int64_t cap, int64_t *n) {
    if (n) *n = (int64_t)h->map.size();
    if (!local) return MSL_OK;  // size query
    if ((int64_t)h->map.size() > cap) return fail(MSL_ERR_CAPACITY, "msl_surfel_download_map: capacity");
    memcpy(local, h->map.data(), sizeof(orc_surfel) * h->map.size());
    return MSL_OK;
}
int msl_surfel_download_changed(msl_surfel_fusion *h, int ref, int32_t *idx, msl_surfel *rec, int64_t cap, int64_t *n) {
    int64_t c = 0;
    for (size_t i = 0; i < h->map.size(); i++) {
        const orc_surfel &e = h->map[i];
        if (e.updateTimes != 0 && e.lastUpdate != ref) continue;
        if (idx && rec) {
            if (c >= cap) return fail(MSL_ERR_CAPACITY, "msl_surfel_download_changed: capacity");
            idx[c] = (int32_t)i;
            memcpy(rec + c, &e, sizeof(e));
        }
        c++;
    }
    if (n) *n = c;
    return MSL_OK;
}
int msl_surfel_fuse(msl_surfel_fusion *h, int ref, const uint8_t *gray, int gray_stride, const float *depth,
                    const int32_t *membership, const float Twc[16], msl_surfel *new_surfels, int cap_new, int compact, int64_t stats[4]) {
    std::vector<orc_surfel> own;
    if (!new_surfels) {  // the caller does not want the new surfels back (device-resident mode)
        own.resize((size_t)(h->w / 8) * (h->h / 8));
        new_surfels = (msl_surfel *)own.data(), cap_new = (int)own.size();
    }
    const int n = orc_surfel_fuse(h->o, ref, gray, gray_stride, depth, membership, Twc, h->map.data(), (int64_t)h->map.size(),
                                  (orc_surfel *)new_surfels, cap_new, 1);
    if (n < 0 || n > cap_new) return fail(MSL_ERR_CAPACITY, "msl_surfel_fuse: capacity");
    if (compact) {  // the tail of SurfelMapping::fuseMap (src/SurfelMapping.cpp:366-391)
        const int64_t n0 = (int64_t)h->map.size();
        h->map.resize((size_t)(n0 + n));
        h->map.resize((size_t)orc_surfel_compact(h->map.data(), n0, (const orc_surfel *)new_surfels, n));
    }
    if (stats) stats[0] = n, stats[1] = stats[2] = 0, stats[3] = (int64_t)h->map.size();
    return MSL_OK;
}

}  // extern "C"
