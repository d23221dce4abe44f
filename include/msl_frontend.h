/*
 * msl_frontend.h -- C ABI of the B200-native RGB-D front-end for ManhattanSLAM.
 *
 * This is the drop-in boundary: a thin extern "C" shim (plain pointers and sizes, no C++/torch
 * types) over hand-written sm_100a CUDA kernels.  The reference-side adapters in adapters/ implement
 * the reference's C++ class surfaces on top of it so Tracking.cc / Frame.cc / SurfelMapping.cpp
 * compile unchanged.  Citations are into razayunus/ManhattanSLAM.
 *
 * Conventions
 *   - every function returns MSL_OK (0) or a negative msl_status; msl_last_error() gives the text of
 *     the last failure on the calling thread;
 *   - handles are opaque, own one CUDA stream plus all device scratch, and -- like the reference
 *     objects they replace (one ORBextractor per Tracking thread, one SurfelFusion per SurfelMapping
 *     thread) -- are NOT thread-safe;
 *   - "host" entry points take caller-allocated host buffers and include the H2D/D2H copies;
 *     "_dev" entry points take device pointers, enqueue on the handle's stream and do not
 *     synchronise (call msl_*_sync);
 *   - there is no CPU fallback: without a CUDA device every create call fails with MSL_ERR_CUDA.
 */
#ifndef MSL_FRONTEND_H
#define MSL_FRONTEND_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    MSL_OK = 0,
    MSL_ERR_INVALID = -1,   /* bad argument */
    MSL_ERR_CUDA = -2,      /* CUDA runtime failure (no device, OOM, launch error) */
    MSL_ERR_CAPACITY = -3,  /* a device-side capacity was exceeded (candidates, nodes, new surfels) */
    MSL_ERR_STATE = -4      /* call order violated (e.g. fuse before upload_map) */
} msl_status;

const char *msl_last_error(void);
const char *msl_version(void);
/* number of CUDA kernels this library has launched since load (all handles, all threads) */
uint64_t msl_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------ ORB
 * Replaces ORB_SLAM2::ORBextractor (include/ORBextractor.h:42-104, src/ORBextractor.cc:412-893). */

typedef struct {
    float x, y;       /* cv::KeyPoint::pt   (src/ORBextractor.cc:788-797, scaled at :861-866) */
    float size;       /* cv::KeyPoint::size = (int)(31*scaleFactor[octave])                    */
    float angle;      /* IC_Angle, degrees  (:75-99)                                           */
    float response;   /* FAST score                                                            */
    int32_t octave;
    int32_t class_id; /* -1 */
} msl_keypoint;       /* 28 bytes, field order of cv::KeyPoint */

typedef struct {
    int32_t nfeatures;   /* ORBextractor.nFeatures  (Example/TUM1.yaml:42) */
    float scale_factor;  /* ORBextractor.scaleFactor */
    int32_t nlevels;     /* ORBextractor.nLevels (<= 16) */
    int32_t ini_th_fast; /* ORBextractor.iniThFAST */
    int32_t min_th_fast; /* ORBextractor.minThFAST */
} msl_orb_params;

typedef struct msl_orb msl_orb;

/* ORBextractor::ORBextractor (src/ORBextractor.cc:412-468).  The handle is sized for frames of
 * exactly w x h and up to max_batch frames per call. */
int msl_orb_create(const msl_orb_params *params, int w, int h, int max_batch, int device, msl_orb **out);
void msl_orb_destroy(msl_orb *);

/* Getters of include/ORBextractor.h:58-82; arrays of nlevels floats (any may be NULL). */
int msl_orb_levels(const msl_orb *);
int msl_orb_scale_factors(const msl_orb *, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2);
/* rows of kps/desc reserved per frame by msl_orb_extract (>= nfeatures + 3*nlevels) */
int msl_orb_capacity(const msl_orb *);

/* ORBextractor::operator() (src/ORBextractor.cc:813-870) on `batch` independent gray frames.
 *   gray   : batch images, image b at gray + b*frame_stride, rows `stride` bytes apart (CV_8UC1)
 *   kps    : batch x capacity msl_keypoint          desc : batch x capacity x 32 bytes
 *   counts : batch int32 (keypoints found per frame)
 * Keypoint order = reference order (levels ascending, octree list order inside a level). */
int msl_orb_extract(msl_orb *, const uint8_t *gray, int stride, size_t frame_stride, int batch,
                    msl_keypoint *kps, uint8_t *desc, int32_t *counts);
/* Same with device pointers; asynchronous on the handle's stream. */
int msl_orb_extract_dev(msl_orb *, const uint8_t *d_gray, int stride, size_t frame_stride, int batch,
                        msl_keypoint *d_kps, uint8_t *d_desc, int32_t *d_counts);
int msl_orb_sync(msl_orb *);
void *msl_orb_stream(msl_orb *); /* cudaStream_t */

/* Stage read-back of the last extract call (parity tests): level image / blurred level image of
 * frame `frame` (dense w*h bytes), FAST candidates fed to the octree (x,y relative to minBorder, score). */
int msl_orb_debug_level_size(const msl_orb *, int level, int *w, int *h);
int msl_orb_debug_level(msl_orb *, int frame, int level, int blurred, uint8_t *out);
int msl_orb_debug_candidates(msl_orb *, int frame, int level, int32_t *xyr, int cap, int *n);

/* -------------------------------------------------------------------------------------- matcher
 * Replaces ORBmatcher::DescriptorDistance and the window searches of ORBmatcher::SearchByProjection
 * (src/ORBmatcher.cc:40-117, 548-678, 680-797, 835-849) including Frame::GetFeaturesInArea semantics
 * (src/Frame.cc:155-168, 332-381, 418-427). */

typedef struct {
    float fx, fy, cx, cy;                 /* Frame::fx.. */
    float mnMinX, mnMinY, mnMaxX, mnMaxY; /* Frame::mnMinX.. (src/Frame.cc:466-493) */
    float gridWInv, gridHInv;             /* Frame::mfGridElementWidthInv / HeightInv (:137-138) */
    float mb, mbf;                        /* Frame::mb, Frame::mbf */
    int32_t nlevels;
    float scaleFactors[16];               /* Frame::mvScaleFactors */
} msl_frame_geom;

typedef struct msl_matcher msl_matcher;
typedef struct msl_glue msl_glue; /* frame glue handle, declared below */

/* max_queries / max_train / max_batch size the Hamming batch buffers and the INITIAL scratch of the searches.  The scratch
 * grows with the call: the reference hands SearchByProjection the whole of mvpLocalMapPoints (src/Tracking.cc:1693) and Fuse
 * the map points of every neighbour keyframe (src/LocalMapping.cc:569), so the number of map points / queries of a search
 * is unbounded; only the keypoints of ONE frame or keyframe are limited (4096: the kernels' shared-memory grid). */
int msl_matcher_create(int max_queries, int max_train, int max_batch, int device, msl_matcher **out);
void msl_matcher_destroy(msl_matcher *);
int msl_matcher_sync(msl_matcher *);
void *msl_matcher_stream(msl_matcher *); /* cudaStream_t */
/* Deferred batches.  The reference calls the searches in loops over keyframes -- Tracking::Relocalization runs SearchByBoW
 * against every candidate keyframe (src/Tracking.cc:1930-1950), LocalMapping::CreateNewMapPoints runs SearchForTriangulation
 * against every neighbour (src/LocalMapping.cc:330-351), SearchInNeighbors runs Fuse against every target keyframe
 * (src/LocalMapping.cc:540-570) -- and one call is a few tens of microseconds of kernel behind a PCIe round trip.  Between
 * msl_matcher_batch_begin and msl_matcher_batch_end the six window / vocabulary searches (msl_search_by_projection_frame /
 * _points / _keyframe, msl_search_by_bow, msl_search_for_triangulation, msl_fuse_search) only pack their inputs (the input
 * arrays may be reused as soon as the call returns); msl_matcher_batch_end uploads everything in one copy, runs one CTA per
 * recorded call and then fills every call's output arrays, which must stay valid until then.  Calls with an empty side
 * return their (empty) result at once.  If the recorded calls outgrow the handle's scratch arena the recorded part is
 * executed early -- results are only promised at batch_end.  A call that fails inside a batch discards everything recorded
 * so far (the handle stays usable; the batch stays open).  Outside a batch every search is a batch of one. */
int msl_matcher_batch_begin(msl_matcher *);
int msl_matcher_batch_end(msl_matcher *);
/* measurement aid: with timing on, every execution (a batch, or a single search) is bracketed by CUDA events on the handle's
 * stream; msl_matcher_last_execution returns the device time of the last one -- upload + kernels + download -- and the
 * number of calls it ran */
int msl_matcher_set_timing(msl_matcher *, int on);
int msl_matcher_last_execution(msl_matcher *, double *device_ms, int *calls);

/* All-pairs Hamming distance of 256-bit descriptors: dist[b][i][j] = popcount(q[b][i] ^ t[b][j]),
 * batch pairs; q: batch x nq x 32, t: batch x nt x 32, dist: batch x nq x nt uint16. */
int msl_hamming_all_pairs(msl_matcher *, const uint8_t *q, int nq, const uint8_t *t, int nt, int batch,
                          uint16_t *dist);
/* Brute-force best/second-best per query row (ties: lowest train index). */
int msl_hamming_best2(msl_matcher *, const uint8_t *q, int nq, const uint8_t *t, int nt, int batch,
                      int32_t *best_idx, int32_t *best_dist, int32_t *second_dist);
int msl_hamming_best2_dev(msl_matcher *, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt, int batch,
                          int32_t *d_best_idx, int32_t *d_best_dist, int32_t *d_second_dist);
/* Same on the ragged output of msl_orb_extract_dev: every batch entry reserves `rows` descriptor rows (and
 * `rows` output slots), of which d_qcounts[b] / d_tcounts[b] are filled.  stream = cudaStream_t to enqueue on
 * (NULL = the handle's stream), e.g. the ORB handle's stream to chain extraction -> matching without a sync. */
int msl_hamming_best2_counts_dev(msl_matcher *, const uint8_t *d_q, const uint8_t *d_t, int rows,
                                 const int32_t *d_qcounts, const int32_t *d_tcounts, int batch, int32_t *d_best_idx,
                                 int32_t *d_best_dist, int32_t *d_second_dist, void *stream);

/* ORBmatcher::SearchByProjection(Frame &Cur, const Frame &Last, th) (src/ORBmatcher.cc:548-678).
 * The adapter flattens the Frame/MapPoint graph into arrays:
 *   Last side, per keypoint i < n_last: last_has_mp (mvpMapPoints[i]!=NULL), last_outlier (mvbOutlier),
 *     last_mp_obs (pMP->Observations()>0), last_mp_world (GetWorldPos, xyz), last_mp_desc (GetDescriptor,
 *     32 B), last_octave (mvKeys[i].octave), last_angle (mvKeysUn[i].angle);
 *   Cur side, per keypoint j < n_cur: cur_xy (mvKeysUn pt), cur_octave, cur_angle, cur_uright (mvuRight),
 *     cur_desc (mDescriptors), cur_occupied (mvpMapPoints[j] && Observations()>0 on entry).
 * Output: cur_match[j] = index i of the Last keypoint whose MapPoint now sits in slot j, -1 if the slot
 * was not touched, -2 if it was occupied on entry, -3 if it was matched and then reset to NULL by the
 * rotation-consistency check; *nmatches = return value of the reference method. */
int msl_search_by_projection_frame(msl_matcher *, const msl_frame_geom *geom, const float Tcw_cur[16],
                                   const float Tcw_last[16], float th, int check_orientation, int n_last,
                                   const uint8_t *last_has_mp, const uint8_t *last_outlier,
                                   const uint8_t *last_mp_obs, const float *last_mp_world,
                                   const uint8_t *last_mp_desc, const int32_t *last_octave,
                                   const float *last_angle, int n_cur, const float *cur_xy,
                                   const int32_t *cur_octave, const float *cur_angle, const float *cur_uright,
                                   const uint8_t *cur_desc, const uint8_t *cur_occupied, int32_t *cur_match,
                                   int32_t *nmatches);

/* The same search for a BATCH of consecutive frames straight from device-resident extractor / glue output (no host round
 * trip): pair p has Last = frame p and Current = frame p + 1, p < n_frames - 1.  The Last side is built on the device as
 * Tracking::TrackWithMotionModel finds it after Tracking::UpdateLastFrame (src/Tracking.cc:1052-1104): every keypoint with
 * depth > 0 in (depth, index) order up to th_depth (Tracking::mThDepth), at least the 100 closest, carries a MapPoint at
 * Frame::UnprojectStereo (src/Frame.cc:515-526) with no observations; no slot of the Current frame is occupied.
 * d_kps / d_desc / d_counts: msl_orb_extract_dev's output (`rows` rows per frame, rows <= 4096); d_xy_un / d_uright /
 * d_kdepth: msl_glue_keypoints_dev's output; Tcw: n_frames x 16 floats on the host (pose of every frame).
 * d_cur_match: (n_frames - 1) x rows int32, row p = cur_match of pair p as above; d_nmatches: n_frames - 1 int32.
 * Asynchronous on `stream` (NULL = the handle's stream). */
int msl_search_by_projection_frames_dev(msl_matcher *, const msl_frame_geom *geom, float th, int check_orientation, float th_depth,
                                        const msl_keypoint *d_kps, const uint8_t *d_desc, int rows, const int32_t *d_counts,
                                        int n_frames, const float *d_xy_un, const float *d_uright, const float *d_kdepth,
                                        const float *Tcw, int32_t *d_cur_match, int32_t *d_nmatches, void *stream);

/* Host form: extractor output (kps / desc / counts, `rows` rows per frame), the CV_32F depth frames (dense w x h) and the poses
 * in host memory; uploads, runs msl_glue_keypoints_dev (undistortion-free camera) + the search, downloads cur_match
 * ((n_frames - 1) x rows) and nmatches.  `glue`: any glue handle of this frame size (supplies the kernels). */
int msl_search_by_projection_frames(msl_matcher *, msl_glue *glue, const msl_frame_geom *geom, float th, int check_orientation,
                                    float th_depth, const msl_keypoint *kps, const uint8_t *desc, int rows, const int32_t *counts,
                                    int n_frames, const float *depth, int w, int h, const float *Tcw, int32_t *cur_match,
                                    int32_t *nmatches);

/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &, th) (src/ORBmatcher.cc:40-117).
 * Per map point k: mp_valid (mbTrackInView && !isBad()), mp_obs (Observations()>0), mp_proj_xyr
 * (mTrackProjX, mTrackProjY, mTrackProjXR), mp_level (mnTrackScaleLevel), mp_viewcos (mTrackViewCos),
 * mp_desc.  nnratio = ORBmatcher::mfNNratio.  cur_match[j] = k as above. */
int msl_search_by_projection_points(msl_matcher *, const msl_frame_geom *geom, float th, float nnratio,
                                    int n_mp, const uint8_t *mp_valid, const uint8_t *mp_obs,
                                    const float *mp_proj_xyr, const int32_t *mp_level, const float *mp_viewcos,
                                    const uint8_t *mp_desc, int n_cur, const float *cur_xy,
                                    const int32_t *cur_octave, const float *cur_uright, const uint8_t *cur_desc,
                                    const uint8_t *cur_occupied, int32_t *cur_match, int32_t *nmatches);

/* ORBmatcher::SearchByProjection(Frame &Cur, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist)
 * (src/ORBmatcher.cc:680-797, relocalisation; called at src/Tracking.cc:2006,2019) including
 * MapPoint::PredictScale (src/MapPoint.cc:350-364) and Get{Min,Max}DistanceInvariance (:324-332).
 *   KeyFrame side, per keypoint i < n_kf (pKF->GetMapPointMatches()): kf_valid (pMP && !pMP->isBad() &&
 *     !sAlreadyFound.count(pMP)), kf_mp_world (GetWorldPos), kf_mp_desc (GetDescriptor), kf_mp_dist
 *     (2 floats: mfMinDistance, mfMaxDistance), kf_angle (pKF->mvKeysUn[i].angle);
 *   Cur side: cur_occupied[j] = (mvpMapPoints[j] != NULL) -- this overload has no Observations() test.
 *   log_scale_factor = Frame::mfLogScaleFactor; geom->nlevels = Frame::mnScaleLevels.
 * cur_match / nmatches as for msl_search_by_projection_frame. */
int msl_search_by_projection_keyframe(msl_matcher *, const msl_frame_geom *geom, const float Tcw_cur[16], float th,
                                      int orb_dist, int check_orientation, float log_scale_factor, int n_kf,
                                      const uint8_t *kf_valid, const float *kf_mp_world, const uint8_t *kf_mp_desc,
                                      const float *kf_mp_dist, const float *kf_angle, int n_cur, const float *cur_xy,
                                      const int32_t *cur_octave, const float *cur_angle, const uint8_t *cur_desc,
                                      const uint8_t *cur_occupied, int32_t *cur_match, int32_t *nmatches);

/* ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches) (src/ORBmatcher.cc:146-255;
 * called at src/Tracking.cc:859,1158,1942).  The two DBoW2::FeatureVector maps (pKF->mFeatVec, F.mFeatVec) in CSR form:
 * node ids in ascending (std::map) order, node k owns the feature indices feat[off[k] .. off[k+1]) in their stored
 * order.  kf_valid[i] = (vpMapPointsKF[i] && !isBad()); kf_angle = pKF->mvKeysUn[i].angle; f_angle = F.mvKeys[j].angle.
 * f_match[j] = index i of the KeyFrame keypoint whose MapPoint now sits in vpMapPointMatches[j], -1 none, -3 assigned
 * and then reset to NULL by the rotation-consistency check; *nmatches = return value of the reference method.
 * MSL_ERR_INVALID for a malformed feature vector (ids not ascending, offsets not monotone, index out of range). */
int msl_search_by_bow(msl_matcher *, float nnratio, int check_orientation, int n_nodes_kf, const uint32_t *kf_node_id,
                      const int32_t *kf_node_off, const int32_t *kf_node_feat, int n_nodes_f, const uint32_t *f_node_id,
                      const int32_t *f_node_off, const int32_t *f_node_feat, int n_kf, const uint8_t *kf_valid,
                      const uint8_t *kf_desc, const float *kf_angle, int n_f, const uint8_t *f_desc, const float *f_angle,
                      int32_t *f_match, int32_t *nmatches);

/* ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12, vMatchedPairs, bOnlyStereo)
 * (src/ORBmatcher.cc:257-406, with CheckDistEpipolarLine :127-144; called at src/LocalMapping.cc:351).
 * F12 row-major 3x3; Cw1 = pKF1->GetCameraCenter(); Tcw2 = pKF2's pose (row-major 4x4); K2 = pKF2 fx, fy, cx, cy;
 * scale_factors2 / level_sigma2_2 = pKF2->mvScaleFactors / mvLevelSigma2 (nlevels entries); feature vectors as above;
 * has_mp = (GetMapPoint(idx) != NULL); uright = mvuRight; xy / angle / octave = mvKeysUn.
 * matches12[idx1] = idx2 (vMatchedPairs = the pairs with matches12 >= 0 in idx1 order), -1 none, -3 removed by the
 * rotation check; *nmatches = return value. */
int msl_search_for_triangulation(msl_matcher *, const float F12[9], const float Cw1[3], const float Tcw2[16],
                                 const float K2[4], int only_stereo, int check_orientation, int nlevels,
                                 const float *scale_factors2, const float *level_sigma2_2, int n_nodes1,
                                 const uint32_t *node_id1, const int32_t *node_off1, const int32_t *node_feat1,
                                 int n_nodes2, const uint32_t *node_id2, const int32_t *node_off2,
                                 const int32_t *node_feat2, int n1, const uint8_t *has_mp1, const float *uright1,
                                 const float *xy1, const float *angle1, const uint8_t *desc1, int n2,
                                 const uint8_t *has_mp2, const float *uright2, const float *xy2, const int32_t *octave2,
                                 const float *angle2, const uint8_t *desc2, int32_t *matches12, int32_t *nmatches);

/* The search part of ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint*> &vpMapPoints, th)
 * (src/ORBmatcher.cc:408-519; called at src/LocalMapping.cc:549,569): per map point the best keypoint of the KeyFrame.
 * geom = the KeyFrame's intrinsics, image bounds, grid and scale factors; Tcw = its pose; inv_level_sigma2 =
 * pKF->mvInvLevelSigma2; log_scale_factor = pKF->mfLogScaleFactor.  Per map point: mp_valid (pMP && !isBad() &&
 * !IsInKeyFrame(pKF)), mp_world (GetWorldPos), mp_normal (GetNormal), mp_dist (mfMinDistance, mfMaxDistance),
 * mp_desc (GetDescriptor).  KeyFrame keypoints: kf_xy / kf_octave = mvKeysUn, kf_uright = mvuRight, kf_desc.
 * best_idx[i] = bestIdx (-1 none), best_dist[i] = bestDist (256 none).  The caller applies :521-541 (Replace /
 * AddObservation / AddMapPoint) to every map point with best_dist <= TH_LOW (50), in order; *nfused = their number. */
int msl_fuse_search(msl_matcher *, const msl_frame_geom *geom, const float Tcw[16], float th, float log_scale_factor,
                    const float *inv_level_sigma2, int n_mp, const uint8_t *mp_valid, const float *mp_world,
                    const float *mp_normal, const float *mp_dist, const uint8_t *mp_desc, int n_kf, const float *kf_xy,
                    const int32_t *kf_octave, const float *kf_uright, const uint8_t *kf_desc, int32_t *best_idx,
                    int32_t *best_dist, int32_t *nfused);

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:210-263; called per map point at src/LocalMapping.cc and
 * src/Tracking.cc after observations change) for a BATCH of map points: point k owns the descriptors
 * desc[offsets[k] .. offsets[k+1]) x 32 bytes -- its observations in non-bad KeyFrames, in std::map order (:228-234).
 * best_idx[k] = BestIdx (the row with the least median distance to the rest, first such row; -1 for a point without
 * descriptors, which the reference leaves untouched); best_median (optional) = BestMedian.  The caller copies
 * descriptor best_idx[k] into mDescriptor (:255-258). */
int msl_distinctive_descriptors(msl_matcher *, int n_points, const int32_t *offsets, const uint8_t *desc,
                                int32_t *best_idx, int32_t *best_median);

/* ---------------------------------------------------------------------------------- frame glue
 * The per-frame steps either side of the ORB extractor, so that a frame can stay on the device from decode to the
 * feature grid: Tracking::GrabImage's cvtColor and depth conversion (src/Tracking.cc:184-211),
 * Frame::UndistortKeyPoints (src/Frame.cc:437-463) and Frame::ComputeStereoFromRGBD (src/Frame.cc:495-513). */

int msl_glue_create(int w, int h, int max_batch, int device, msl_glue **out);
void msl_glue_destroy(msl_glue *);
int msl_glue_sync(msl_glue *);
void *msl_glue_stream(msl_glue *); /* cudaStream_t */

/* cvtColor(mImGray, mImGray, CV_RGB2GRAY | CV_BGR2GRAY | CV_RGBA2GRAY | CV_BGRA2GRAY) (src/Tracking.cc:189-200):
 * channels = 3 or 4, rgb_order = Tracking::mbRGB.  src rows `stride` bytes apart; gray dense w*h per frame. */
int msl_glue_cvt_gray(msl_glue *, const uint8_t *src, int stride, int channels, int rgb_order, int batch, uint8_t *gray);
int msl_glue_cvt_gray_dev(msl_glue *, const uint8_t *d_src, int stride, size_t frame_stride, int channels, int rgb_order,
                          int batch, uint8_t *d_gray, int gray_stride, size_t gray_frame_stride);
/* mImDepth.convertTo(mImDepth, CV_32F, mDepthMapFactor) (src/Tracking.cc:205-207) on CV_16U depth */
int msl_glue_depth_to_float(msl_glue *, const uint16_t *depth16, int batch, float factor, float *depth);
int msl_glue_depth_to_float_dev(msl_glue *, const uint16_t *d_depth16, int64_t n, float factor, float *d_depth);
/* The sensor frames of a batch uploaded ONCE for every stage.  The reference's Frame constructor hands the same host images to
 * every consumer -- ExtractORB(imGray), ComputeStereoFromRGBD(imDepth), ExtractPlanes(imDepth as CV_16U) (src/Frame.cc:90-110,
 * :604-610) and later SurfelFusion through the KeyFrame (src/SurfelMapping.cpp:353-364) -- so a binding that goes through the
 * host entry points uploads gray twice and depth three times.  Here gray (CV_8U) and the sensor depth (CV_16U) go up once, on
 * the handle's own copy stream, into frame set `slot` (0 .. MSL_GLUE_FRAME_SETS - 1: with three sets the upload of batch k+2 can
 * be issued while batch k is being worked on, so that batch k+1's frames are on the device before batch k's fuse chain starts
 * and its superpixel stage can run beside that chain), and the CV_32F depth of Tracking::GrabImageRGBD (convertTo(CV_32F, mDepthMapFactor),
 * src/Tracking.cc:205-207) is produced on the device.  `aux` (optional, aux_ints int32 values) rides along, e.g. a
 * membership image computed on the host.  Returns at once; the device pointers stay valid until the slot's next upload and are
 * meant for the *_dev entry points after msl_glue_frames_wait(glue, slot, <that handle's stream>).  The caller must not upload
 * into a slot whose consumers of the previous upload have not been synchronised. */
#define MSL_GLUE_FRAME_SETS 3
int msl_glue_upload_frames(msl_glue *, int slot, const uint8_t *gray, int gray_stride, const uint16_t *depth16,
                           int depth_stride_px, int batch, float factor, const int32_t *aux, size_t aux_ints,
                           const uint8_t **d_gray, const uint16_t **d_depth16, const float **d_depth, const int32_t **d_aux);
/* makes `stream` (cudaStream_t; NULL = block the host) wait for the upload + conversion of frame set `slot` */
int msl_glue_frames_wait(msl_glue *, int slot, void *stream);
/* Frame::UndistortKeyPoints + Frame::ComputeStereoFromRGBD for one frame's keypoints (mvKeys as returned by
 * msl_orb_extract): K4 = fx, fy, cx, cy; D5 = k1, k2, p1, p2, k3 (NULL or D5[0] == 0: mvKeysUn = mvKeys);
 * depth = CV_32F metres, dense w*h, or NULL to skip the stereo part; mbf = Frame::mbf.
 * Outputs: xy_un (2 floats per keypoint, may be NULL), uright = mvuRight, kdepth = mvDepth (-1 where depth <= 0). */
int msl_glue_keypoints(msl_glue *, const msl_keypoint *kps, int n, const float K4[4], const float D5[5],
                       const float *depth, float mbf, float *xy_un, float *uright, float *kdepth);
/* Batched device form on the ragged output of msl_orb_extract_dev (`rows` keypoint rows reserved per frame, d_counts
 * filled); d_depth: batch dense frames; stream = cudaStream_t to enqueue on (NULL = the handle's), e.g. the ORB
 * handle's stream to chain extraction -> glue without a sync. */
int msl_glue_keypoints_dev(msl_glue *, const msl_keypoint *d_kps, int rows, const int32_t *d_counts, int batch,
                           const float K4[4], const float D5[5], const float *d_depth, float mbf, float *d_xy_un,
                           float *d_uright, float *d_kdepth, void *stream);

/* ----------------------------------------------------------------------------- plane pre-stage
 * Replaces PlaneDetection::readDepthImage (src/PlaneExtractor.cpp:44-76) and the peac pre-stage:
 * PlaneSeg ctor + Stats::compute per 10x10 block (include/peac/AHCPlaneSeg.hpp:148-181, 235-312)
 * and the node/edge initialisation of PlaneFitter::initGraph (include/peac/AHCPlaneFitter.hpp:756-928). */

typedef struct {
    double center[3];
    double normal[3];
    double mse;
    double curvature;
    int32_t N;     /* member points (0 if rejected) */
    int32_t nouse; /* 1 = rejected by missing data / depth discontinuity */
} msl_block_stat;  /* 72 bytes */

typedef struct msl_plane msl_plane;

int msl_plane_create(int w, int h, int max_batch, int device, msl_plane **out);
void msl_plane_destroy(msl_plane *);
/* depth: batch CV_16U images (row stride in pixels = dstride_px, image b at depth + b*frame_stride_px).
 * K = {fx, fy, cx, cy}.  Outputs per frame (any may be NULL): cloud_xyz (h2*w2*3 doubles, h2=ceil(h/2)),
 * blocks ((h2/10)*(w2/10) msl_block_stat), seed (1 = graph node), edges (bit0 left,1 right,2 up,3 down). */
int msl_plane_prestage(msl_plane *, const uint16_t *depth, int dstride_px, size_t frame_stride_px, int batch,
                       const float K[4], float depth_map_factor, double *cloud_xyz, msl_block_stat *blocks,
                       uint8_t *seed, uint8_t *edges);
int msl_plane_prestage_dev(msl_plane *, const uint16_t *d_depth, int dstride_px, size_t frame_stride_px,
                           int batch, const float K[4], float depth_map_factor, double *d_cloud_xyz,
                           msl_block_stat *d_blocks, uint8_t *d_seed, uint8_t *d_edges);
int msl_plane_sync(msl_plane *);
void *msl_plane_stream(msl_plane *);

/* Plane detection proper: replaces PlaneDetection::runPlaneDetection (src/PlaneExtractor.cpp:78-82), i.e. the whole of
 * ahc::PlaneFitter<ImagePointCloud>::run (include/peac/AHCPlaneFitter.hpp:207-258): the pre-stage above, ahCluster
 * (:939-1143) and refineDetails (:294-374) with findBlockMembership (:480-582) and floodFill (:422-471), with the
 * reference's default parameters (minSupport 3000, 10x10 windows, ERODE_ALL_BORDER, doRefine).
 * Per frame: membership = PlaneFitter::membershipImg (h2*w2 int32: plane id, -1 = none, <= -2 = floodFill's trail
 * counters -- what Tracking hands to SurfelFusion, src/Tracking.cc:228,497); plane_count = plane_num_; planes =
 * plane_cap records per frame, the first min(plane_count, plane_cap) filled in extractedPlanes order (descending N):
 * normal / center of extractedPlanes[i] (src/Frame.cc:626-632), N, rid, and vertices = plane_vertices_[i].size(); the
 * members of plane i are the pixels whose membership equals i, in row-major order (src/Frame.cc:612-621).
 * Frames of up to 768 blocks (640x480) are processed in shared memory, up to 3072 blocks (1280x960) in global memory;
 * larger ones: MSL_ERR_INVALID.  MSL_ERR_CAPACITY if a frame exceeds the region-grow queue (4 entries per half-resolution
 * pixel) or 128 planes. */
typedef struct {
    double normal[3];
    double center[3];
    int32_t N;        /* PlaneSeg::N (points of the merged blocks; not updated by the region grow, as in the reference) */
    int32_t rid;      /* PlaneSeg::rid (root block id) */
    int32_t vertices; /* number of member pixels after refinement */
    int32_t pad;
} msl_plane_rec;      /* 64 bytes */

int msl_plane_detect(msl_plane *, const uint16_t *depth, int dstride_px, size_t frame_stride_px, int batch, const float K[4],
                     float depth_map_factor, int32_t *membership, int32_t *plane_count, msl_plane_rec *planes, int plane_cap);
/* device buffers in, device buffers out (enqueued on the handle's stream; errors of the frames surface at the next
 * msl_plane_sync) */
int msl_plane_detect_dev(msl_plane *, const uint16_t *d_depth, int dstride_px, size_t frame_stride_px, int batch,
                         const float K[4], float depth_map_factor, int32_t *d_membership, int32_t *d_plane_count,
                         msl_plane_rec *d_planes, int plane_cap);

/* Measurement aid: 16 values per frame of the last msl_plane_detect* call: out[16 f + k], k = 0..6 globaltimer stamps (ns) --
 * start, graph built, ahCluster done, block membership + region-grow seeds, region grow done, final merge done, end;
 * k = 7 merge steps taken; k = 8..13 SM cycles of ahCluster's sub-phases summed over its steps (queue pop, candidate fits,
 * selection, publish + decision, adjacency update, node copy + queue push). */
int msl_plane_debug_profile(msl_plane *, int64_t *out, int frames);

/* ------------------------------------------------------------------------------------- surfels
 * Replaces SurfelFusion (include/SurfelFusion.h:44-139, src/SurfelFusion.cpp), the compaction tail of
 * SurfelMapping::fuseMap (src/SurfelMapping.cpp:366-391) and SurfelMapping::moveAddSurfels (:194-304), so that
 * Map::mvLocalSurfels and Map::mvInactiveSurfels can stay on the device between keyframes. */

typedef struct {
    float px, py, pz;
    float nx, ny, nz;
    float size;
    float color;
    int32_t r, g, b;
    float weight;
    int32_t updateTimes;
    int32_t lastUpdate;
} msl_surfel; /* include/Surfel.h:28-37, 56 bytes */

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
} msl_seed; /* SurfelFusion::SuperpixelSeed (include/SurfelFusion.h:46-58), bools widened */

typedef struct msl_surfel_fusion msl_surfel_fusion;

/* SurfelFusion::SurfelFusion (src/SurfelFusion.cpp:29-38); max_surfels = capacity of the
 * device-resident local map (Map::mvLocalSurfels). */
int msl_surfel_create(int w, int h, float fx, float fy, float cx, float cy, float fuse_far, float fuse_near,
                      int64_t max_surfels, int device, msl_surfel_fusion **out);
void msl_surfel_destroy(msl_surfel_fusion *);

/* Explicit sync points for the device-resident map (the host mutates mvLocalSurfels between
 * keyframes in SurfelMapping::moveAddSurfels, src/SurfelMapping.cpp:194-304). */
int msl_surfel_upload_map(msl_surfel_fusion *, const msl_surfel *local, int64_t n);
int msl_surfel_download_map(msl_surfel_fusion *, msl_surfel *local, int64_t cap, int64_t *n);
int64_t msl_surfel_map_size(const msl_surfel_fusion *);
/* Dirty download for the exact drop-in (adapters/SurfelFusion_msl.cpp, host vector authoritative): the surfels the last
 * NON-compacting msl_surfel_fuse call with reference index `ref` changed -- updated (lastUpdate == ref,
 * src/SurfelFusion.cpp:275) or deleted (updateTimes == 0, :182 / :210 / :237) -- as (index, record) pairs in ascending index
 * order.  idx / rec NULL: only *n.  About 30 % of the map instead of all of it over PCIe per keyframe. */
int msl_surfel_download_changed(msl_surfel_fusion *, int ref, int32_t *idx, msl_surfel *rec, int64_t cap, int64_t *n);

/* SurfelMapping::moveAddSurfels (src/SurfelMapping.cpp:194-304) on the device-resident maps.  poses_to_remove /
 * poses_to_add are what SurfelMapping::getAddRemovePoses (:306-326) returned (that walk over the pose graph stays on
 * the host).  Moving out: local surfels with updateTimes > 0 && lastUpdate == pose are appended -- pose after pose, map
 * order inside a pose -- to the device-side Map::mvInactiveSurfels (= PoseElement::attachedSurfels) and their local slot
 * keeps updateTimes = 0 until the next fuse compacts it, as in the reference.  Moving in: the attached surfels of the
 * poses to add are appended to the local map in list order and leave the inactive store.
 *   stats (may be NULL): {moved_out, moved_in, local size afterwards (dead slots included)}.
 * MSL_ERR_STATE if a pose to add was never moved out or a pose to remove is already inactive. */
int msl_surfel_move_add(msl_surfel_fusion *, const int32_t *poses_to_remove, int n_remove, const int32_t *poses_to_add,
                        int n_add, int64_t stats[3]);
/* Map::mvInactiveSurfels in the reference's order (poses in the order they were moved out, minus those moved back) */
int64_t msl_surfel_inactive_size(const msl_surfel_fusion *);
int msl_surfel_download_inactive(msl_surfel_fusion *, msl_surfel *out, int64_t cap, int64_t *n);

/* SurfelFusion::fuseInitializeMap (src/SurfelFusion.cpp:40-73) on the device-resident map.
 *   gray (CV_8UC1, row stride gray_stride), depth (CV_32F metres, dense w*h), membership (CV_32SC1,
 *   ceil(h/2) x ceil(w/2), -1 = no plane), Twc (row-major 4x4 camera->world).
 *   new_surfels (cap_new entries, may be NULL) receives SurfelFusion's newSurfels in seed order.
 *   compact != 0 additionally applies the fuseMap tail (src/SurfelMapping.cpp:366-391) on the device.
 *   stats (may be NULL): {n_new, n_updated, n_deleted, map_size_after}. */
int msl_surfel_fuse(msl_surfel_fusion *, int reference_frame_index, const uint8_t *gray, int gray_stride,
                    const float *depth, const int32_t *membership, const float Twc[16], msl_surfel *new_surfels,
                    int cap_new, int compact, int64_t stats[4]);
/* Device-pointer variant (inputs already resident); asynchronous, stats are device-resident until
 * msl_surfel_read_stats. */
int msl_surfel_fuse_dev(msl_surfel_fusion *, int reference_frame_index, const uint8_t *d_gray, int gray_stride,
                        const float *d_depth, const int32_t *d_membership, const float Twc[16], int compact);
/* Stream of `batch` consecutive keyframes (reference indices ref0, ref0+1, ...): the map-independent
 * superpixel stage runs batched over all frames, then fuse/initialise/compact runs frame by frame in
 * order on the device-resident map -- identical to `batch` consecutive fuseInitializeMap calls.
 * Twc: batch x 16 floats (host).  stats accumulate over the batch (n_new, n_updated, n_deleted) and
 * report the final map size. */
int msl_surfel_fuse_batch(msl_surfel_fusion *, int ref0, const uint8_t *gray, int gray_stride, const float *depth,
                          const int32_t *membership, const float *Twc, int batch, int compact, int64_t stats[4]);
int msl_surfel_fuse_batch_dev(msl_surfel_fusion *, int ref0, const uint8_t *d_gray, int gray_stride,
                              size_t gray_frame_stride, const float *d_depth, const int32_t *d_membership,
                              const float *Twc, int batch, int compact);
int msl_surfel_read_stats(msl_surfel_fusion *, int64_t stats[4]);
int msl_surfel_read_new(msl_surfel_fusion *, msl_surfel *new_surfels, int cap_new, int *n_new);
int msl_surfel_sync(msl_surfel_fusion *);
/* Measurement aid: CUDA events on the handle's stream around the kernels of the per-frame chain.  One event record
 * costs about 2.7 us of stream time, so mode 1 -- meant for use inside a timed region -- marks only k_fuse_scan and
 * k_fuse_apply (3 events) on every 8th frame, mode 2 marks all six points of every frame, mode 0 is off.
 * msl_surfel_fuse_kernel_time synchronises and returns the summed k_fuse_scan time and the number of timed launches;
 * msl_surfel_chain_times returns out = {scan, apply, post, list, cmp_apply} in ms summed over the timed frames
 * (mode 1: scan and apply only) and resets the tally. */
int msl_surfel_set_timing(msl_surfel_fusion *, int mode);
int msl_surfel_fuse_kernel_time(msl_surfel_fusion *, double *total_ms, int *launches);
int msl_surfel_chain_times(msl_surfel_fusion *, double out[5], int *frames);
/* Number of kernels fuseSurfelsKernel (src/SurfelFusion.cpp:167-283) runs as: 1 = k_fuse_one (scan and fuse in one
 * kernel; the "scan" interval of the timing aid is that kernel and "apply" is empty), 2 = k_fuse_scan + k_fuse_apply
 * (environment MSL_FUSE_ONE=0). */
int msl_surfel_fuse_kernels(const msl_surfel_fusion *);
/* Per-frame count table of the batched stream API (what the ranks all-gather): after every following
 * msl_surfel_fuse_batch_dev call d_table[2 b] = new surfels of frame b (initializeSurfels, src/SurfelFusion.cpp:285-331),
 * d_table[2 b + 1] = surfels frame b updated (fuseSurfelsKernel, :240-279); batch x 2 int32 in device memory, written on the
 * handle's stream.  NULL switches it off. */
int msl_surfel_set_count_table(msl_surfel_fusion *, int32_t *d_table);
/* launch geometry of the last fuseSurfelsKernel launch (test / bench evidence that the persistent multi-draw path ran):
 * out = {kernels, form (MSL_FUSE_ONE), persistent, grid (CTAs), warps per CTA, 128-surfel segments} */
int msl_surfel_launch_info(const msl_surfel_fusion *, int32_t out[6]);
/* CTAs per SM of the persistent fuse kernel (k_fuse_pipe; 3 fill an SM's registers).  batch_ctas applies to the per-frame
 * chain of msl_surfel_fuse_batch(_dev) calls of >= 8 frames, where the NEXT batch's superpixel stage runs beside the chain:
 * the default 2 leaves a third of every SM to it (measured: 8.50 -> 8.15 ms per 64-frame step; the kernel itself is 6 %
 * slower alone); single_ctas (default 3) applies to single frames and short batches.  0 keeps a value. */
int msl_surfel_set_fuse_ctas_per_sm(msl_surfel_fusion *, int batch_ctas, int single_ctas);
void *msl_surfel_stream(msl_surfel_fusion *);
/* cudaStream_t on which msl_surfel_fuse_batch_dev / msl_surfel_fuse_dev READ their frame inputs (gray, depth, membership): a
 * producer of device-resident inputs (msl_plane_detect_dev's membership image, src/Tracking.cc:227-229) makes it wait for
 * its event before the call, and records an event on it after the call to learn when the buffer may be overwritten. */
void *msl_surfel_input_stream(msl_surfel_fusion *);

/* Validation aid: k_fuse_scan divides by the camera-frame depth with a hand-scheduled IEEE sequence that shares one
 * reciprocal between the two image coordinates; this compares it bit for bit with the compiler's division on n random
 * operand triples (divisor in [c_lo, c_hi], numerators in [-a_max, a_max]) and returns the number of mismatches. */
int msl_surfel_selftest_div(msl_surfel_fusion *, int64_t n, uint64_t seed, float c_lo, float c_hi, float a_max,
                            int64_t *mismatches);

/* Batched superpixel generation only (generateSuperPixels, src/SurfelFusion.cpp:805-816) for `batch`
 * independent frames: seeds (batch x (w/8)*(h/8) msl_seed), index (batch x w*h int32, may be NULL). */
int msl_surfel_superpixels(msl_surfel_fusion *, const uint8_t *gray, int gray_stride, const float *depth,
                           const int32_t *membership, int batch, msl_seed *seeds, int32_t *index);

/* Stage read-back of the last fuse call. */
int msl_surfel_debug_seeds(msl_surfel_fusion *, msl_seed *seeds);
int msl_surfel_debug_index(msl_surfel_fusion *, int32_t *index);

#ifdef __cplusplus
}
#endif
#endif /* MSL_FRONTEND_H */
