// mock_abi_matcher.cpp -- TEST HARNESS (tests/test_adapters_on_mock_abi.py): the matcher entry points of
// include/msl_frontend.h that adapters/ORBmatcher_msl.cc calls, implemented on the CPU ORACLE (see mock_abi.cpp).
#include <string>
#include <vector>

#include "msl_frontend.h"
#include "msl_oracle.h"

static_assert(sizeof(msl_frame_geom) == sizeof(orc_frame_geom), "frame geometry layout");
struct msl_matcher {
    int dummy;
};
static thread_local std::string g_merr;
#define G(g) ((const orc_frame_geom *)(g))

extern "C" {

const char *msl_last_error(void) __attribute__((weak));
const char *msl_last_error(void) { return g_merr.c_str(); }

int msl_matcher_create(int, int, int, int, msl_matcher **out) {
    *out = new msl_matcher{0};
    return MSL_OK;
}
void msl_matcher_destroy(msl_matcher *m) { delete m; }

int msl_search_by_projection_frame(msl_matcher *, const msl_frame_geom *g, const float Tc[16], const float Tl[16], float th, int chk,
                                   int n_last, const uint8_t *has, const uint8_t *outl, const uint8_t *obs, const float *world,
                                   const uint8_t *desc, const int32_t *oct, const float *ang, int n_cur, const float *xy,
                                   const int32_t *coct, const float *cang, const float *ur, const uint8_t *cdesc, const uint8_t *occ,
                                   int32_t *match, int32_t *nmatches) {
    *nmatches = orc_search_by_projection_frame(G(g), Tc, Tl, th, chk, n_last, has, outl, obs, world, desc, oct, ang, n_cur, xy, coct, cang,
                                               ur, cdesc, occ, match);
    return MSL_OK;
}
int msl_search_by_projection_points(msl_matcher *, const msl_frame_geom *g, float th, float nnratio, int n_mp, const uint8_t *valid,
                                    const uint8_t *obs, const float *proj, const int32_t *lvl, const float *vcos, const uint8_t *desc,
                                    int n_cur, const float *xy, const int32_t *coct, const float *ur, const uint8_t *cdesc,
                                    const uint8_t *occ, int32_t *match, int32_t *nmatches) {
    *nmatches = orc_search_by_projection_points(G(g), th, nnratio, n_mp, valid, obs, proj, lvl, vcos, desc, n_cur, xy, coct, ur, cdesc, occ, match);
    return MSL_OK;
}
int msl_search_by_projection_keyframe(msl_matcher *, const msl_frame_geom *g, const float Tc[16], float th, int orb_dist, int chk,
                                      float lsf, int n_kf, const uint8_t *valid, const float *world, const uint8_t *desc,
                                      const float *dist, const float *ang, int n_cur, const float *xy, const int32_t *coct,
                                      const float *cang, const uint8_t *cdesc, const uint8_t *occ, int32_t *match, int32_t *nmatches) {
    *nmatches = orc_search_by_projection_keyframe(G(g), Tc, th, orb_dist, chk, lsf, n_kf, valid, world, desc, dist, ang, n_cur, xy, coct, cang,
                                                  cdesc, occ, match);
    return MSL_OK;
}
int msl_search_by_bow(msl_matcher *, float nnratio, int chk, int nk_nodes, const uint32_t *kid, const int32_t *koff, const int32_t *kfeat,
                      int nf_nodes, const uint32_t *fid, const int32_t *foff, const int32_t *ffeat, int n_kf, const uint8_t *valid,
                      const uint8_t *kdesc, const float *kang, int n_f, const uint8_t *fdesc, const float *fang, int32_t *match,
                      int32_t *nmatches) {
    *nmatches = orc_search_by_bow(nnratio, chk, nk_nodes, kid, koff, kfeat, nf_nodes, fid, foff, ffeat, n_kf, valid, kdesc, kang, n_f, fdesc,
                                  fang, match);
    return MSL_OK;
}
int msl_search_for_triangulation(msl_matcher *, const float F12[9], const float Cw1[3], const float Tcw2[16], const float K2[4],
                                 int only_stereo, int chk, int nlevels, const float *sf2, const float *ls2, int nn1, const uint32_t *id1,
                                 const int32_t *off1, const int32_t *ft1, int nn2, const uint32_t *id2, const int32_t *off2,
                                 const int32_t *ft2, int n1, const uint8_t *has1, const float *ur1, const float *xy1, const float *ang1,
                                 const uint8_t *desc1, int n2, const uint8_t *has2, const float *ur2, const float *xy2,
                                 const int32_t *oct2, const float *ang2, const uint8_t *desc2, int32_t *m12, int32_t *nmatches) {
    *nmatches = orc_search_for_triangulation(F12, Cw1, Tcw2, K2, only_stereo, chk, nlevels, sf2, ls2, nn1, id1, off1, ft1, nn2, id2, off2, ft2,
                                             n1, has1, ur1, xy1, ang1, desc1, n2, has2, ur2, xy2, oct2, ang2, desc2, m12);
    return MSL_OK;
}
int msl_fuse_search(msl_matcher *, const msl_frame_geom *g, const float Tcw[16], float th, float lsf, const float *ils, int n_mp,
                    const uint8_t *valid, const float *world, const float *normal, const float *dist, const uint8_t *desc, int n_kf,
                    const float *xy, const int32_t *oct, const float *ur, const uint8_t *kdesc, int32_t *best_idx, int32_t *best_dist,
                    int32_t *nfused) {
    *nfused = orc_fuse_search(G(g), Tcw, th, lsf, ils, n_mp, valid, world, normal, dist, desc, n_kf, xy, oct, ur, kdesc, best_idx, best_dist);
    return MSL_OK;
}

}  // extern "C"

extern "C" int msl_distinctive_descriptors(msl_matcher *, int n_points, const int32_t *offsets, const uint8_t *desc, int32_t *best_idx,
                                           int32_t *best_median) {
    std::vector<int32_t> bm(n_points > 0 ? n_points : 1);
    orc_distinctive_descriptors(n_points, offsets, desc, best_idx, best_median ? best_median : bm.data());
    return MSL_OK;
}
