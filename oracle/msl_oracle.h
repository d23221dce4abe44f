/*
 * msl_oracle.h -- CPU oracle for the ManhattanSLAM RGB-D front-end hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a dependency-free restatement of the
 * reference's CPU algorithm (razayunus/ManhattanSLAM) used as the parity
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing in the product path (manhattanslam_b200/,
 * include/) may include, link or call it.
 *
 * Parity pinning: the reference ships no tests or golden vectors (SURVEY.md
 * section 4) and cannot be built as a whole here (no OpenCV / Eigen / PCL).  Two
 * anchors replace them:
 *  (1) the third-party primitives restated here (cv::FAST, cv::resize INTER_LINEAR
 *      8U, cv::GaussianBlur 7x7 8U, cv::fastAtan2, cvRound, cv::gemm / cv::norm on
 *      3x3 / 3x1 floats, cv::cvtColor, cv::undistortPoints) are pinned bit-exactly
 *      against Python cv2 4.13.0 in tests/test_oracle_primitives.py and through
 *      committed fixtures in tests/golden/;
 *  (2) everything above the primitives is pinned against the REFERENCE'S OWN
 *      SOURCE: oracle/Makefile compiles src/ORBextractor.cc, src/ORBmatcher.cc,
 *      src/PlaneExtractor.cpp (+ include/peac/), src/SurfelFusion.cpp,
 *      src/SurfelMapping.cpp and src/MapPoint.cc where they lie, unmodified, against stand-in headers (oracle/ref_shim_cv, ref_shim_match, ref_shim_map, ref_shim_mp) into
 *      oracle/_ref/, and tests/test_oracle_ref.py requires the restatements here to
 *      equal them bit for bit.  What the stand-ins decide -- and what therefore
 *      stays "parity unpinned" -- is listed in DESIGN.md section 2: Eigen's
 *      arithmetic (3x3 eigen-solver, 4x4 inverse, fixed-size products), cv::Mat::dot,
 *      Mat::convertTo, the thread schedule of the racy `stable` flag, and heap-address
 *      order (creation order here) where the reference sorts or iterates by pointer.
 *
 * All functions are plain C ABI so that Python ctypes can drive them.
 */
#ifndef MSL_ORACLE_H
#define MSL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ ORB --- */

typedef struct {
    float x, y;      /* cv::KeyPoint::pt (level-0 coordinates after O8 scaling) */
    float size;      /* (int)(31 * scaleFactor[level])                          */
    float angle;     /* degrees, cv::fastAtan2                                  */
    float response;  /* FAST corner score                                       */
    int32_t octave;  /* pyramid level                                           */
    int32_t class_id;/* always -1                                               */
} orc_keypoint;      /* 28 bytes */

typedef struct orc_orb orc_orb;

/* ORBextractor::ORBextractor, src/ORBextractor.cc:412-468 */
orc_orb *orc_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
void orc_orb_destroy(orc_orb *);
/* getters: vectors of length nlevels (ORBextractor.h:58-82) */
int orc_orb_levels(const orc_orb *);
void orc_orb_scale_factors(const orc_orb *, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2);
void orc_orb_features_per_level(const orc_orb *, int32_t *n);
void orc_orb_umax(const orc_orb *, int32_t *umax16);

/* ORBextractor::operator(), src/ORBextractor.cc:813-870.  Returns the number of
 * keypoints (<= cap), -1 if cap is too small, -2 for a geometry on which the reference itself is undefined (a level more
 * than twice as tall as wide: zero octree roots, division by zero at :535).  desc is n x 32 bytes. */
int orc_orb_extract(orc_orb *, const uint8_t *gray, int w, int h, int stride,
                    orc_keypoint *kps, uint8_t *desc, int cap);

/* Stage dumps of the LAST orc_orb_extract call (for stage-level parity). */
int orc_orb_level_size(const orc_orb *, int level, int *w, int *h);
const uint8_t *orc_orb_level_image(const orc_orb *, int level);          /* w*h, dense */
const uint8_t *orc_orb_level_blurred(const orc_orb *, int level);        /* w*h, dense */
/* candidates fed to DistributeOctTree (x,y relative to minBorder, response) */
int orc_orb_level_candidates(const orc_orb *, int level, int32_t *xyr, int cap);
/* keypoints per level after the octree (level coordinates, before O8 scaling) */
int orc_orb_level_keypoints(const orc_orb *, int level, orc_keypoint *kps, int cap);

/* Primitives (pinned against cv2 4.13). */
void orc_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride,
                          uint8_t *dst, int dw, int dh, int dstride);
void orc_gaussian_blur_7x7_s2_u8(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride);
/* cv::FAST(img, kps, threshold, nonmaxSuppression=true, TYPE_9_16): out = (x,y,score) triples */
int orc_fast_9_16(const uint8_t *img, int w, int h, int stride, int threshold, int nms,
                  int32_t *xyr, int cap);
/* threshold-free score map: S_max(p) (0..255); FAST score = S_max-1; corner at t <=> S_max > t */
void orc_fast_score_map(const uint8_t *img, int w, int h, int stride, uint8_t *smax, int ostride);
float orc_fast_atan2(float y, float x);
int orc_cv_round_f(float v);

/* -------------------------------------------------------------- matcher --- */

/* ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:835-849 */
int orc_descriptor_distance(const uint8_t *a, const uint8_t *b);
/* cv::Mat arithmetic as the matcher restatements evaluate it (pinned against cv2.gemm / cv2.norm):
 * R * x + t (flags 0: float row sums), -R.t() * t (transposed operand: double accumulation),
 * -Rwc * t with Rwc = Rcw.t() materialised (flags 0, alpha -1), cv::norm of a 3-vector. */
void orc_cv_rx_plus_t(const float R[9], const float x[3], const float t[3], float out[3]);
void orc_cv_neg_rt_times_t(const float R[9], const float t[3], float out[3]);
void orc_cv_neg_rwc_times_t(const float Rcw[9], const float t[3], float out[3]);
double orc_cv_norm3(const float v[3]);
void orc_hamming_best2(const uint8_t *q, int nq, const uint8_t *t, int nt, int32_t *best_idx, int32_t *best_dist,
                       int32_t *second_dist);

typedef struct {
    float fx, fy, cx, cy;
    float mnMinX, mnMinY, mnMaxX, mnMaxY;
    float gridWInv, gridHInv; /* mfGridElementWidthInv / HeightInv */
    float mb, mbf;
    int32_t nlevels;
    float scaleFactors[16];
} orc_frame_geom;

/* Frame::AssignFeaturesToGrid + GetFeaturesInArea, src/Frame.cc:155-168,332-381,418-427.
 * Returns the number of indices written (candidate order of the reference). */
int orc_features_in_area(const orc_frame_geom *g, const float *kp_xy, const int32_t *kp_octave, int n,
                         float x, float y, float r, int minLevel, int maxLevel, int32_t *out, int cap);

/* ORBmatcher::SearchByProjection(Frame&Cur, const Frame&Last, th), src/ORBmatcher.cc:548-678.
 * Last-frame side: per keypoint i: has_mp[i], outlier[i], world position, MapPoint descriptor,
 * octave and (undistorted) angle.  Cur side: undistorted keypoints, uRight, descriptors and
 * cur_occupied[i] = (mvpMapPoints[i] && Observations()>0) on entry.  last_mp_obs[i] =
 * (pMP->Observations() > 0) decides whether a slot assigned during this call blocks later queries.
 * Output: cur_match[i2] = index of the Last keypoint whose MapPoint was assigned to slot i2
 * (-1 none; slots that were occupied on entry keep -2).  Returns nmatches. */
int orc_search_by_projection_frame(const orc_frame_geom *g, const float Tcw_cur[16], const float Tcw_last[16],
                                   float th, int check_orientation,
                                   int n_last, const uint8_t *last_has_mp, const uint8_t *last_outlier,
                                   const uint8_t *last_mp_obs, const float *last_mp_world,
                                   const uint8_t *last_mp_desc, const int32_t *last_octave, const float *last_angle,
                                   int n_cur, const float *cur_xy, const int32_t *cur_octave, const float *cur_angle,
                                   const float *cur_uright, const uint8_t *cur_desc, const uint8_t *cur_occupied,
                                   int32_t *cur_match);

/* ORBmatcher::SearchByProjection(Frame&F, const vector<MapPoint*>&, th), src/ORBmatcher.cc:40-117.
 * Per map point: track_in_view && !bad flag, mTrackProjX/Y/XR, mnTrackScaleLevel, mTrackViewCos,
 * descriptor, mp_obs = Observations()>0.  Output cur_match as above.  Returns nmatches. */
int orc_search_by_projection_points(const orc_frame_geom *g, float th, float nnratio,
                                    int n_mp, const uint8_t *mp_valid, const uint8_t *mp_obs,
                                    const float *mp_proj_xyr, const int32_t *mp_level, const float *mp_viewcos,
                                    const uint8_t *mp_desc,
                                    int n_cur, const float *cur_xy, const int32_t *cur_octave,
                                    const float *cur_uright, const uint8_t *cur_desc, const uint8_t *cur_occupied,
                                    int32_t *cur_match);

/* ORBmatcher::SearchByProjection(Frame&Cur, KeyFrame*, const set<MapPoint*>&sAlreadyFound, th, ORBdist),
 * src/ORBmatcher.cc:680-797 (relocalisation).  Per KeyFrame keypoint i: kf_valid[i] = (pMP && !isBad() &&
 * !sAlreadyFound.count(pMP)), world position, descriptor, kf_mp_dist[2i..] = (mfMinDistance, mfMaxDistance),
 * kf_angle = pKF->mvKeysUn[i].angle.  cur_occupied[j] = (mvpMapPoints[j] != NULL).  log_scale_factor =
 * Frame::mfLogScaleFactor, g->nlevels = Frame::mnScaleLevels.  Output cur_match as above.  Returns nmatches. */
int orc_search_by_projection_keyframe(const orc_frame_geom *g, const float Tcw_cur[16], float th, int orb_dist,
                                      int check_orientation, float log_scale_factor,
                                      int n_kf, const uint8_t *kf_valid, const float *kf_mp_world,
                                      const uint8_t *kf_mp_desc, const float *kf_mp_dist, const float *kf_angle,
                                      int n_cur, const float *cur_xy, const int32_t *cur_octave, const float *cur_angle,
                                      const uint8_t *cur_desc, const uint8_t *cur_occupied, int32_t *cur_match);

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), src/ORBmatcher.cc:146-255.  The two
 * DBoW2::FeatureVector maps in CSR form (node ids ascending; node k owns features feat[off[k] .. off[k+1]) in their
 * stored order).  kf_valid[i] = (pMP && !pMP->isBad()); kf_angle = pKF->mvKeysUn[i].angle; f_angle = F.mvKeys[j].angle.
 * Output f_match[j] = index of the KeyFrame keypoint whose MapPoint was assigned to F's slot j, -1 none, -3 assigned
 * and then reset by the rotation check.  Returns nmatches. */
int orc_search_by_bow(float nnratio, int check_orientation, int n_nodes_kf, const uint32_t *kf_node_id,
                      const int32_t *kf_node_off, const int32_t *kf_node_feat, int n_nodes_f, const uint32_t *f_node_id,
                      const int32_t *f_node_off, const int32_t *f_node_feat, int n_kf, const uint8_t *kf_valid,
                      const uint8_t *kf_desc, const float *kf_angle, int n_f, const uint8_t *f_desc, const float *f_angle,
                      int32_t *f_match);

/* ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:257-406 (+ CheckDistEpipolarLine :127-144).  F12 row-major 3x3,
 * Cw1 = pKF1->GetCameraCenter(), Tcw2 = pKF2 pose (row-major 4x4), K2 = pKF2 fx, fy, cx, cy; scale_factors2 /
 * level_sigma2_2 = pKF2->mvScaleFactors / mvLevelSigma2; has_mp = (GetMapPoint(idx) != NULL); xy / angle = mvKeysUn.
 * Output matches12[idx1] = idx2, -1 none, -3 removed by the rotation check.  Returns nmatches. */
int orc_search_for_triangulation(const float F12[9], const float Cw1[3], const float Tcw2[16], const float K2[4],
                                 int only_stereo, int check_orientation, int nlevels, const float *scale_factors2,
                                 const float *level_sigma2_2, int n_nodes1, const uint32_t *node_id1,
                                 const int32_t *node_off1, const int32_t *node_feat1, int n_nodes2,
                                 const uint32_t *node_id2, const int32_t *node_off2, const int32_t *node_feat2, int n1,
                                 const uint8_t *has_mp1, const float *uright1, const float *xy1, const float *angle1,
                                 const uint8_t *desc1, int n2, const uint8_t *has_mp2, const float *uright2,
                                 const float *xy2, const int32_t *octave2, const float *angle2, const uint8_t *desc2,
                                 int32_t *matches12);

/* The search part of ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th), src/ORBmatcher.cc:408-519: per map
 * point the best KeyFrame keypoint (best_idx -1 / best_dist 256 if none).  mp_valid[i] = (pMP && !pMP->isBad() &&
 * !pMP->IsInKeyFrame(pKF)); mp_dist = (mfMinDistance, mfMaxDistance); g = the KeyFrame's intrinsics, image bounds, grid
 * and scale factors; Tcw = its pose.  Returns the number of map points with best_dist <= TH_LOW. */
int orc_fuse_search(const orc_frame_geom *g, const float Tcw[16], float th, float log_scale_factor,
                    const float *inv_level_sigma2, int n_mp, const uint8_t *mp_valid, const float *mp_world,
                    const float *mp_normal, const float *mp_dist, const uint8_t *mp_desc, int n_kf, const float *kf_xy,
                    const int32_t *kf_octave, const float *kf_uright, const uint8_t *kf_desc, int32_t *best_idx,
                    int32_t *best_dist);

/* MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:210-263, for a batch of map points: point k owns
 * desc[off[k] .. off[k+1]).  best_idx[k] = BestIdx inside the point's list (-1 if it has no descriptor),
 * best_median[k] = BestMedian. */
void orc_distinctive_descriptors(int n_points, const int32_t *off, const uint8_t *desc, int32_t *best_idx,
                                 int32_t *best_median);

/* ---------------------------------------------------- plane pre-stage --- */

typedef struct {
    double center[3];
    double normal[3];
    double mse;
    double curvature;
    int32_t N;
    int32_t nouse;
} orc_block_stat;

/* PlaneDetection::readDepthImage (src/PlaneExtractor.cpp:44-76) -> cloud (h2*w2*3 doubles),
 * PlaneSeg ctor + Stats::compute per 10x10 block (AHCPlaneSeg.hpp:235-312,148-181),
 * initGraph seed test and edges (AHCPlaneFitter.hpp:756-928).
 * seed[b] = 1 if block b becomes a graph node; edges[b] bit0=left,1=right,2=up,3=down. */
void orc_plane_prestage(const uint16_t *depth, int w, int h, int dstride_px,
                        float fx, float fy, float cx, float cy, float depthMapFactor,
                        double *cloud_xyz, orc_block_stat *blocks, uint8_t *seed, uint8_t *edges);
/* Frame::ExtractPlanes' readDepthImage + runPlaneDetection (src/Frame.cc:607-609): the whole of ahc::PlaneFitter::run --
 * the pre-stage above, ahCluster (include/peac/AHCPlaneFitter.hpp:939-1143) and refineDetails (:294-374); restated in
 * peac_oracle.inc.  membership: h2*w2 int32 (plane id, -1, or floodFill's trail counters <= -2); per extracted plane
 * (<= cap): normal[3], center[3], N, rid, number of member pixels.  Returns the number of planes. */
int orc_plane_detect(const uint16_t *depth, int w, int h, int dstride_px, float fx, float fy, float cx, float cy,
                     float depthMapFactor, int32_t *membership, double *plane_normal, double *plane_center, int32_t *plane_N,
                     int32_t *plane_rid, int32_t *plane_vertices, int cap);
/* symmetric 3x3 eigen-decomposition used above (ascending eigenvalues, columns of V) */
void orc_eig33sym(const double K[9], double s[3], double V[9]);

/* -------------------------------------------------------------- surfels --- */

typedef struct {
    float px, py, pz;
    float nx, ny, nz;
    float size;
    float color;
    int32_t r, g, b;
    float weight;
    int32_t updateTimes;
    int32_t lastUpdate;
} orc_surfel; /* include/Surfel.h:28-37, 56 bytes */

typedef struct {
    float x, y;
    float size;
    float normX, normY, normZ;
    float posX, posY, posZ;
    float viewCos;
    float meanDepth;
    float meanIntensity;
    int32_t r, g, b;
    int32_t fused, stable, use;
} orc_seed; /* SuperpixelSeed, include/SurfelFusion.h:46-58 (bools widened to int32) */

typedef struct orc_surfel_fusion orc_surfel_fusion;

/* SurfelFusion::SurfelFusion, src/SurfelFusion.cpp:29-38 */
orc_surfel_fusion *orc_surfel_create(int w, int h, float fx, float fy, float cx, float cy,
                                     float fuseFar, float fuseNear);
void orc_surfel_destroy(orc_surfel_fusion *);

/* SurfelFusion::fuseInitializeMap, src/SurfelFusion.cpp:40-73.  local is updated in place,
 * new_surfels receives up to cap_new entries; returns the number of new surfels.
 * threads: 1 = sequential slices 0..9 (the deterministic oracle order); >1 uses std::thread
 * over the same slices for the fuse scan only (result is slice-independent, see S8). */
int orc_surfel_fuse(orc_surfel_fusion *, int referenceFrameIndex,
                    const uint8_t *gray, int gray_stride, const float *depth, const int32_t *membership,
                    const float Twc[16], orc_surfel *local, int64_t n_local,
                    orc_surfel *new_surfels, int cap_new, int threads);

/* Stage dumps of the last call. */
const int32_t *orc_surfel_index(const orc_surfel_fusion *);         /* w*h */
const orc_seed *orc_surfel_seeds(const orc_surfel_fusion *);        /* (w/8)*(h/8) */
const float *orc_surfel_normmap(const orc_surfel_fusion *);         /* w*h*3 */
/* snapshot of seeds/index after updateSeeds of iteration it (0..2) */
const orc_seed *orc_surfel_seeds_iter(const orc_surfel_fusion *, int it);
const int32_t *orc_surfel_index_iter(const orc_surfel_fusion *, int it);

/* SurfelMapping::fuseMap tail, src/SurfelMapping.cpp:366-391: refill deleted slots from the back of
 * the deleted list with new surfels, append the rest, swap-remove remaining deleted slots.
 * local must have room for n_local + n_new.  Returns the new local size. */
int64_t orc_surfel_compact(orc_surfel *local, int64_t n_local, const orc_surfel *new_surfels, int n_new);

/* SurfelMapping::moveAddSurfels, src/SurfelMapping.cpp:194-304, on the state it touches (posesDatabase's
 * attachedSurfels / pointsBeginIndex / pointsPoseIndex, pointcloudPoseIndex, Map::mvInactiveSurfels).  The two pose
 * lists are what getAddRemovePoses (:306-326) returned.  local has room for cap_local surfels.  Returns the new
 * local size (surfels moved out stay as updateTimes == 0 slots, as in the reference), -1 if a pose to add was never
 * moved out, -2 if cap_local is too small. */
typedef struct orc_surfel_mapping orc_surfel_mapping;
orc_surfel_mapping *orc_mapping_create(void);
void orc_mapping_destroy(orc_surfel_mapping *);
int64_t orc_move_add_surfels(orc_surfel_mapping *, orc_surfel *local, int64_t n_local, int64_t cap_local,
                             const int32_t *posesToRemove, int n_remove, const int32_t *posesToAdd, int n_add);
/* Map::mvInactiveSurfels (copied to out, at most cap entries); returns its size */
int64_t orc_mapping_inactive(const orc_surfel_mapping *, orc_surfel *out, int64_t cap);

/* ------------------------------------------------------------ frame glue --- */
/* cv::cvtColor RGB/BGR(A) -> GRAY on CV_8U (src/Tracking.cc:189-200); channels 3|4, rgb_order 1 = R first */
void orc_cvt_gray(const uint8_t *src, int w, int h, int stride, int channels, int rgb_order, uint8_t *dst, int dstride);
/* Mat::convertTo(CV_32F, factor) on CV_16U depth (src/Tracking.cc:205-207) */
void orc_depth_to_float(const uint16_t *src, int64_t n, float factor, float *dst);
/* cv::undistortPoints(src, dst, K, D, noArray(), K) with K4 = fx,fy,cx,cy and D5 = k1,k2,p1,p2,k3 */
void orc_undistort_points(int n, const float *xy, const float K4[4], const float D5[5], float *out);
/* Frame::UndistortKeyPoints (src/Frame.cc:437-463): copy when D5[0] == 0 */
void orc_undistort_keypoints(int n, const float *xy, const float K4[4], const float D5[5], float *out);
/* Frame::ComputeStereoFromRGBD (src/Frame.cc:495-513): kp_xy = mvKeys, kpun_xy = mvKeysUn, dense w-wide depth */
void orc_stereo_from_rgbd(int n, const float *kp_xy, const float *kpun_xy, const float *depth, int w, float mbf,
                          float *uright, float *kdepth);

#ifdef __cplusplus
}
#endif
#endif
