/*
 * plslam_c.h — C ABI of the B200-native point+line front-end.
 *
 * This is the drop-in boundary for the data-parallel hot path of
 * wolfcanli/ORB_SLAM2_Modification_with-point-and-line-feature.  The reference has no FFI layer; its boundary
 * is four C++ classes (ORBextractor, LineExtractor, ORBmatcher, LineMatcher).  The host-side mirrors of those
 * classes (package host/ directory) flatten Frame / MapPoint / MapLine state into the POD arrays declared here
 * and call these entry points.  Every entry point cites the reference interface it replaces (file:line, relative
 * to the reference root).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++ / torch / OpenCV types.
 *   - every function returns PL_OK (0) or a negative pl_status; pl_last_error() gives a thread-local message.
 *     (The reference itself has no error convention: empty image -> silent return, ORBextractor.cc:1046; wrong
 *     type -> assert, :1050.  The host mirrors translate PL_ERR_EMPTY into the silent return.)
 *   - "host" entry points take host pointers and perform the H2D / D2H copies themselves on the handle's stream;
 *     "_dev" entry points take device pointers (inputs already resident in HBM) and are asynchronous on the
 *     handle's stream until pl_*_sync().
 *   - there is NO CPU fallback: if no sm_100-class device is usable, creation fails with PL_ERR_CUDA.
 */
#ifndef PLSLAM_C_H
#define PLSLAM_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PL_API __attribute__((visibility("default")))

typedef enum {
    PL_OK = 0,
    PL_ERR_ARG = -1,      /* bad argument (null pointer, non-positive size, capacity too small) */
    PL_ERR_EMPTY = -2,    /* empty image: mirrors the reference's silent return                  */
    PL_ERR_CUDA = -3,     /* CUDA runtime failure or no usable device                            */
    PL_ERR_CAPACITY = -4, /* an internal or caller capacity was exceeded; output not valid       */
    PL_ERR_STATE = -5     /* call sequence error (e.g. reading a pyramid before an extract)      */
} pl_status;

/* == cv::KeyPoint (28 bytes).  reference: every std::vector<cv::KeyPoint> crossing ORBextractor.h:59-61 */
typedef struct {
    float x, y;      /* pt */
    float size;
    float angle;     /* degrees, [0,360) */
    float response;
    int octave;
    int class_id;
} pl_keypoint;

/* == cv::line_descriptor::KeyLine (68 bytes, 17 four-byte fields in declaration order).
 * reference: std::vector<KeyLine> crossing LineExtractor.h:25-30 */
typedef struct {
    float angle;     /* radians, atan2(dy,dx) */
    int class_id;
    int octave;
    float pt_x, pt_y;
    float response;
    float size;
    float sx, sy, ex, ey;                       /* start/end in the original image      */
    float sx_oct, sy_oct, ex_oct, ey_oct;       /* start/end in the octave image        */
    float length;
    int num_pixels;
} pl_keyline;

PL_API const char* pl_last_error(void);
PL_API int pl_device_count(void);
/* version / build tag of the native library ("sm_100a") */
PL_API const char* pl_build_info(void);

/* ------------------------------------------------------------------------------------------------------------
 * A. ORB extraction — replaces ORBextractor::ORBextractor (ORBextractor.cc:410-470) and
 *    ORBextractor::operator() (ORBextractor.cc:1043-1105; header ORBextractor.h:51-61).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct pl_orb pl_orb;

/* ctor: same five parameters as the reference constructor, plus device ordinal and the maximum image size /
 * batch this handle will see (device buffers are sized once; nothing is allocated per call). */
PL_API int pl_orb_create(pl_orb** out, int nfeatures, float scale_factor, int nlevels, int ini_th_fast,
                         int min_th_fast, int device, int max_cols, int max_rows, int max_batch);
PL_API void pl_orb_destroy(pl_orb* h);

/* getters — ORBextractor.h:63-84.  Each writes nlevels floats (or ints). */
PL_API int pl_orb_levels(const pl_orb* h);
PL_API float pl_orb_scale_factor(const pl_orb* h);
PL_API int pl_orb_scale_factors(const pl_orb* h, float* out);
PL_API int pl_orb_inv_scale_factors(const pl_orb* h, float* out);
PL_API int pl_orb_level_sigma2(const pl_orb* h, float* out);
PL_API int pl_orb_inv_level_sigma2(const pl_orb* h, float* out);
PL_API int pl_orb_features_per_level(const pl_orb* h, int* out);
/* capacity (in keypoints) that is always sufficient for one frame: sum over levels of quota+3 */
PL_API int pl_orb_max_keypoints(const pl_orb* h);

/* operator(): one 8UC1 image (host pointer, row stride `step` bytes) -> keypoints + 32-byte descriptors.
 * `cap` = capacity of kps / desc in keypoints.  Synchronous.  */
PL_API int pl_orb_extract(pl_orb* h, const uint8_t* gray, int rows, int cols, size_t step, pl_keypoint* kps,
                          uint8_t* desc, int cap, int* n_out);

/* Batched operator() over n_frames independent images of identical size (config 5 of BASELINE.json; frames are
 * independent because operator() keeps no state across calls, ORBextractor.cc:1053-1073).
 * frame f starts at gray + f*frame_stride.  Outputs for frame f are written at kps + f*cap, desc + f*cap*32;
 * n_out[f] receives its count.  Host pointers; synchronous. */
PL_API int pl_orb_extract_batch(pl_orb* h, const uint8_t* gray, int n_frames, int rows, int cols, size_t step,
                                size_t frame_stride, pl_keypoint* kps, uint8_t* desc, int cap, int* n_out);

/* Same, but every pointer is a DEVICE pointer on the handle's device; asynchronous on the handle's stream.
 * Used for HBM-resident measurement and for chaining extract -> match without a host round trip. */
PL_API int pl_orb_extract_batch_dev(pl_orb* h, const uint8_t* d_gray, int n_frames, int rows, int cols,
                                    size_t step, size_t frame_stride, pl_keypoint* d_kps, uint8_t* d_desc, int cap,
                                    int* d_n_out);
/* waits for the handle's stream; returns PL_ERR_CAPACITY when a frame extracted through the device-pointer API since the
 * previous call exceeded a capacity (the host-pointer calls report it themselves). */
PL_API int pl_orb_sync(pl_orb* h);
/* the CUDA stream of the handle as an opaque pointer (cudaStream_t) — for event timing by the caller */
PL_API void* pl_orb_stream(pl_orb* h);

/* Backs the PUBLIC member ORBextractor::mvImagePyramid (ORBextractor.h:85; read by Frame.cc:895,985,997,1002).
 * Copies pyramid level `level` of frame `frame` of the LAST extract call into `out` (host) as the BORDERED plane
 * (EDGE_THRESHOLD=19 px of BORDER_REFLECT_101 on each side, ORBextractor.cc:1107-1130).  The un-bordered image
 * is the ROI at (19,19).  out_step >= cols+38. */
PL_API int pl_orb_pyramid_dims(const pl_orb* h, int level, int* rows, int* cols);
PL_API int pl_orb_pyramid_read(pl_orb* h, int frame, int level, uint8_t* out, size_t out_step);

/* The same level as a DEVICE pointer to the un-bordered image (pixel (0,0) of mvImagePyramid[level]; the 19-pixel border lies
 * around it in the same allocation, row pitch `pitch` bytes) — for consumers that stay on the device (pl_frame_compute_stereo_matches).
 * Valid until the next extract call on the handle; after the device-pointer extract call pl_orb_sync first. */
PL_API int pl_orb_pyramid_dev(pl_orb* h, int frame, int level, const uint8_t** d_image, size_t* pitch, int* rows, int* cols);

/* Debug / test hook: the ordered FAST candidate list handed to DistributeOctTree for (frame, level) of the last
 * extract call (ORBextractor.cc:779-829: vToDistributeKeys), coordinates relative to minBorder. */
PL_API int pl_orb_candidates_read(pl_orb* h, int frame, int level, float* xs, float* ys, float* responses, int cap,
                                  int* n_out);
/* Test hook: the 7x7 sigma=2 blurred plane (ORBextractor.cc:1085-1086) of (frame, level): cols x rows bytes. */
PL_API int pl_orb_blurred_read(pl_orb* h, int frame, int level, uint8_t* out, size_t out_step);
/* number of kernel launches issued by the last extract call (bench.py's gpu_launches claim) */
PL_API int pl_orb_last_launches(const pl_orb* h);
/* Measurement hooks.  With profiling on, every chunk records CUDA events on the launching stream between the
 * stages {0 pyramid, 1 FAST cells, 2 quadtree, 3 blur, 4 orientation+BRIEF} and pl_orb_stage_ms returns the
 * accumulated milliseconds per stage since profiling was switched on (and the number of chunks). */
PL_API int pl_orb_set_profiling(pl_orb* h, int on);
PL_API int pl_orb_stage_ms(pl_orb* h, float* out5, int* chunks);
/* sizes of the current geometry: allocated pyramid / blurred bytes per frame, sum of level pixels (P) and of
 * bordered level pixels (Pb) — the quantities SURVEY.md §8(d)'s algorithmic-bytes formulas use */
PL_API int pl_orb_bytes_per_frame(const pl_orb* h, long long* pyr, long long* blur, long long* level_px, long long* bordered_px);

/* ------------------------------------------------------------------------------------------------------------
 * B. Line extraction — replaces LineExtractor::ExtractLineSegment (LineExtractor.cpp:12-70; LineExtractor.h:25-30):
 *    LSD (cv::line_descriptor::LSDDetector::detect, 1 octave, LSD_REFINE_ADV) -> keep `max_lines` longest ->
 *    LBD 256-bit descriptors (BinaryDescriptor::compute) -> normalised homogeneous line coefficients.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct pl_line pl_line;

PL_API int pl_line_create(pl_line** out, int device, int max_cols, int max_rows, int max_batch);
PL_API void pl_line_destroy(pl_line* h);

/* one image -> <= max_lines keylines (reference constant 80, LineExtractor.cpp:23), n x 32 descriptor bytes,
 * n x 3 doubles of unit-norm line coefficients (LineExtractor.cpp:60-69).  Host pointers, synchronous. */
PL_API int pl_line_extract(pl_line* h, const uint8_t* gray, int rows, int cols, size_t step, int max_lines,
                           pl_keyline* kls, uint8_t* desc, double* coeffs, int* n_out);
PL_API int pl_line_extract_batch(pl_line* h, const uint8_t* gray, int n_frames, int rows, int cols, size_t step,
                                 size_t frame_stride, int max_lines, pl_keyline* kls, uint8_t* desc, double* coeffs,
                                 int* n_out);
/* The Frame constructor hands BOTH extractors the same image (Frame.cc:152-155): after a host-pointer call of the ORB extractor
 * the frames of its last chunk are still staged in HBM.  pl_orb_staged_images_dev returns them (valid until the handle's next
 * extract call; n_frames = the whole batch when it fitted one chunk, max_batch >= n), pl_line_extract_batch_from_dev is
 * pl_line_extract_batch reading its images from device memory (results to host memory, synchronous) — so a frame travels to the
 * device once. */
PL_API int pl_orb_staged_images_dev(pl_orb* h, const uint8_t** d_images, int* n_frames, int* rows, int* cols, size_t* step, size_t* frame_stride);
/* pl_orb_extract_batch in two halves, for ONE chunk (n_frames <= max_batch): pl_orb_stage_batch only enqueues the copy of the frames
 * into the staging buffer and returns; pl_orb_extract_staged is the rest of the call (kernels, read-back, capacity check; synchronous,
 * same results and error behaviour as pl_orb_extract_batch).  In between, pl_orb_stream_wait_staged makes another stream — the line
 * extractor's, pl_line_stream() — wait for the staging copy, so that pl_line_extract_batch_from_dev on the staged frames
 * (pl_orb_staged_images_dev) can be issued from the second host thread at once: the two extractors of the Frame constructor
 * (Frame.cc:152-155) start together on one upload. */
PL_API int pl_orb_stage_batch(pl_orb* h, const uint8_t* gray, int n_frames, int rows, int cols, size_t step, size_t frame_stride);
PL_API int pl_orb_stream_wait_staged(pl_orb* h, void* stream);
PL_API int pl_orb_extract_staged(pl_orb* h, pl_keypoint* kps, uint8_t* desc, int cap, int* n_out);
PL_API int pl_line_extract_batch_from_dev(pl_line* h, const uint8_t* d_gray, int n_frames, int rows, int cols, size_t step,
                                          size_t frame_stride, int max_lines, pl_keyline* kls, uint8_t* desc, double* coeffs,
                                          int* n_out);
/* Makes `stream` (a cudaStream_t, e.g. pl_orb_stream(orb)) wait for the point of the line extraction most recently enqueued on h at
 * which its streaming stages (scale, gradient, seed sort) are done and its region grower is launched.  The grower holds its SMs for
 * tens of milliseconds without filling them: work enqueued behind this point runs NEXT to it, whereas work started together with the
 * line extractor first competes with the streaming stages for the whole GPU. */
PL_API int pl_line_stream_wait_grow_start(pl_line* h, void* stream);
/* Tracking mode: with on != 0 the ~16 stream operations of a chunk of at most four frames are captured into a CUDA graph on first use
 * and replayed as one launch while the call's pointers and sizes stay the same (a host-pointer pl_line_extract always qualifies: it
 * stages into the handle's own buffers); other parameters re-capture.  Off by default (PLSLAM_LINE_GRAPH=1 turns it on at creation). */
PL_API int pl_line_set_graph(pl_line* h, int on);
PL_API int pl_line_extract_batch_dev(pl_line* h, const uint8_t* d_gray, int n_frames, int rows, int cols,
                                     size_t step, size_t frame_stride, int max_lines, pl_keyline* d_kls,
                                     uint8_t* d_desc, double* d_coeffs, int* d_n_out);
/* waits for the handle's stream; returns PL_ERR_CAPACITY when a frame extracted through the device-pointer API since the
 * previous call exceeded a capacity (the host-pointer calls report it themselves). */
PL_API int pl_line_sync(pl_line* h);
PL_API void* pl_line_stream(pl_line* h);
PL_API int pl_line_last_launches(const pl_line* h);
/* The ordered region grower runs as persistent CTAs that own their SM, so kernels of other streams (the matchers, the Frame
 * glue) wait for it unless some SMs are left free: n SMs are kept out of the grower's grid whenever a batch has more frames
 * than the remaining SMs (default 0).  Results do not depend on it. */
PL_API int pl_line_set_reserved_sms(pl_line* h, int n);

/* Measurement hooks: stages {0 blur+scale+gradient, 1 seed sort, 2 region growing/NFA, 3 KeyLines+blur5+Sobel, 4 LBD} */
PL_API int pl_line_set_profiling(pl_line* h, int on);
PL_API int pl_line_stage_ms(pl_line* h, float* out5, int* chunks);
/* of stage 2 (grow + NFA) the part k_lsd_nfa takes after the grower has ended (what its tail helpers did not validate), accumulated like
 * pl_line_stage_ms */
PL_API int pl_line_nfa_ms(pl_line* h, float* ms);
/* with profiling on, 16 values of frame `frame` of the last chunk from k_lsd_grow2's own clock64 accounting:
 * {0 sequencer cycles in commit-time re-growth, 1 cycles the frame was active, 2 sequencer cycles in commits (without the re-growth),
 *  3 tickets issued, 4 regions re-grown at commit, 5 regions committed, 6 tickets deferred, 7 tickets void, 8 sequencer cycles
 *  issuing tickets, 9 sequencer cycles idle, 10 / 11 / 12 grower cycles (summed over the grower warps) growing / waiting for a
 *  ticket / parking results, 13 grower warps, 14 grower cycles in growth that was given up, 15 reserved} */
PL_API int pl_line_grow_phases(pl_line* h, int frame, long long* out16);
/* Test hook: the double-precision sin / cos the line path uses on the device where one ulp decides a discrete outcome
 * (region2rect of lsd.cpp: rec.dx / rec.dy; the seeds' initial sums) — glibc's own arithmetic restated
 * (csrc/pl_glibc_sincos.cuh), so that the results equal std::sin / std::cos of the reference's host. */
PL_API int pl_test_sincos(const double* x, int n, double* s, double* c);
/* Test hooks: the 0.8-scaled 8-bit image LSD works on, its level-line angle map (float degrees, -1024 = undefined,
 * rows x cols of the scaled image) and the float LBD descriptors (n x 72) of frame `frame` of the last call. */
PL_API int pl_line_scaled_dims(const pl_line* h, int* rows, int* cols);
PL_API int pl_line_scaled_read(pl_line* h, int frame, uint8_t* out, size_t out_step);
PL_API int pl_line_angles_read(pl_line* h, int frame, float* out);
PL_API int pl_line_fdesc_read(pl_line* h, int frame, float* out, int n);

/* Test hook: ALL LSD segments of frame `frame` of the last call before the top-`max_lines` filter, in detection
 * order: x1,y1,x2,y2 (float, image coordinates) + width, prec(p), nfa as doubles
 * (cv::LineSegmentDetector::detect outputs). */
PL_API int pl_line_lsd_read(pl_line* h, int frame, float* xyxy, double* width, double* prec, double* nfa, int cap,
                            int* n_out);

/* ------------------------------------------------------------------------------------------------------------
 * C/D. Descriptor search.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct pl_match pl_match;
PL_API int pl_match_create(pl_match** out, int device);
PL_API void pl_match_destroy(pl_match* h);
PL_API int pl_match_sync(pl_match* h);
PL_API void* pl_match_stream(pl_match* h);
PL_API int pl_match_last_launches(const pl_match* h);

/* C1/D1: ORBmatcher::DescriptorDistance (ORBmatcher.cc:2083-2103) == LineMatcher::DescriptorDistance
 * (LineMatcher.cpp:20-39): 256-bit Hamming distance of row pairs (a[i], b[i]), i < n. */
PL_API int pl_hamming_pairs(pl_match* h, const uint8_t* a, const uint8_t* b, int n, int* dist);

/* D6: brute-force 2-nearest-neighbour Hamming search == cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, 2)
 * (LineMatcher.cpp:496-503, :1179-1185).  idx/dist are nq x 2; ties resolve to the lowest train index; when
 * nt < 2 the missing entries are idx=-1, dist=-1 (the reference reads them out of bounds, :507). */
PL_API int pl_hamming_knn2(pl_match* h, const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx, int* dist);
PL_API int pl_hamming_knn2_dev(pl_match* h, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int* d_idx,
                               int* d_dist);

/* Candidate scoring shared by ORBmatcher::SearchByProjection / SearchByBoW (ORBmatcher.cc:123-164, :1802-1829,
 * :306-336): for query row i, distances to train rows cand_idx[cand_off[i] .. cand_off[i+1]) in that order. */
PL_API int pl_hamming_candidates(pl_match* h, const uint8_t* q, int nq, const uint8_t* t, int nt,
                                 const int* cand_off, const int* cand_idx, int* dist_out);

/* Frame-side view used by the projection searches: the subset of ORB_SLAM2::Frame the matcher reads. */
typedef struct {
    int n;                        /* Frame::N                                                    */
    const pl_keypoint* keys_un;   /* Frame::mvKeysUn (x,y,octave,angle read)                     */
    const uint8_t* desc;          /* Frame::mDescriptors, n x 32                                 */
    const float* u_right;         /* Frame::mvuRight, n (<=0: no stereo)                         */
    const int* claimed;           /* n; 1 when mvpMapPoints[i] && Observations()>0 (ORBmatcher.cc:128-130) */
    float min_x, min_y, max_x, max_y;   /* Frame::mnMinX.. (Frame.cc:184-185)                    */
    float fx, fy, cx, cy, bf, b;  /* intrinsics, mbf, mb                                         */
    float tcw[12];                /* Frame::mTcw rows 0..2 (row-major 3x4)                       */
    int n_levels;
    const float* scale_factors;   /* Frame::mvScaleFactors                                       */
} pl_frame_view;

/* C2: ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) — ORBmatcher.cc:72-183.
 * Map points are given as the fields IsInFrustum (Frame.cc:345-401) leaves on them.  match_of_feature[n] receives,
 * per frame feature, the index of the map point assigned to it or -1 (== F.mvpMapPoints writes).  Returns the
 * count through n_matches. */
typedef struct {
    int n;
    const uint8_t* desc;          /* MapPoint::GetDescriptor(), n x 32            */
    const uint8_t* track_in_view; /* mbTrackInView && !isBad()                    */
    const float* proj_x;          /* mTrackProjX                                  */
    const float* proj_y;          /* mTrackProjY                                  */
    const float* proj_xr;         /* mTrackProjXR                                 */
    const int* scale_level;       /* mnTrackScaleLevel                            */
    const float* view_cos;        /* mTrackViewCos                                */
    const uint8_t* has_observations; /* Observations()>0: a feature assigned to such a point becomes "claimed" (:128-130) */
} pl_mappoint_view;
PL_API int pl_orb_search_local_points(pl_match* h, const pl_frame_view* F, const pl_mappoint_view* mps, float th,
                                      float nn_ratio, int* match_of_feature, int* n_matches);

/* C3: ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) — ORBmatcher.cc:1710-1879.
 * last_* describe LastFrame.mvpMapPoints: per Last feature i, valid[i] = (pMP && !mvbOutlier[i]), world position,
 * the map point's descriptor, mvKeys[i].octave, mvKeysUn[i].angle and Last's Tcw.  match_of_feature as above
 * (index i of the Last feature whose map point was assigned, or -1). */
typedef struct {
    int n;
    const uint8_t* valid;
    const float* world_pos;       /* n x 3 */
    const uint8_t* desc;          /* n x 32 */
    const int* octave;
    const float* angle;
    const uint8_t* has_observations; /* pMP->Observations()>0 (temporal VO points have 0, ORBmatcher.cc:1807-1809) */
    float tcw[12];
} pl_lastframe_view;
PL_API int pl_orb_search_last_frame(pl_match* h, const pl_frame_view* Cur, const pl_lastframe_view* Last, float th,
                                    int mono, int check_orientation, int* match_of_feature, int* n_matches);

/* Batched forms: n independent reference calls (e.g. the n frame pairs of an offline sequence whose pose priors are
 * already known) evaluated in one pass — every input array travels in one packed H2D copy, the kernels run one CTA /
 * one grid row per call, results come back in one D2H copy.  match_of_feature[i] has cur[i].n entries. */
PL_API int pl_orb_search_last_frame_batch(pl_match* h, int n, const pl_frame_view* cur, const pl_lastframe_view* last, float th,
                                          int mono, int check_orientation, int* const* match_of_feature, int* n_matches);
PL_API int pl_orb_search_local_points_batch(pl_match* h, int n, const pl_frame_view* F, const pl_mappoint_view* mps, float th,
                                            float nn_ratio, int* const* match_of_feature, int* n_matches);

/* Tracking::SearchLocalPoints (Tracking.cc:1746-1813) as one call per batch of frames: Frame::isInFrustum(pMP, 0.5) of every
 * local map point (Frame.cc:345-401, with MapPoint::PredictScale, MapPoint.cc:416-431 — the F4 arithmetic of
 * pl_frame_is_in_frustum_batch) followed by ORBmatcher(nn_ratio).SearchByProjection(mCurrentFrame, mvpLocalMapPoints, th)
 * (ORBmatcher.cc:44-124 — C2).  What IsInFrustum leaves on the map points (mbTrackInView, mTrackProjX / Y / XR,
 * mnTrackScaleLevel, mTrackViewCos) stays in HBM between the two: nothing per (frame, map point) crosses the host.
 * maps = the distinct snapshots of the local map (consecutive frames between two key frames see the same one and share its
 * upload), map_of_frame[i] = the snapshot of frame i, ow = n x 3 camera centres (Frame::mOw), F[i].claimed = features already
 * matched (mvpMapPoints[idx] != NULL, e.g. by C3).  n_in_view[i] (nullable) = map points that passed IsInFrustum (what
 * Tracking.cc:1792 counts with IncreaseVisible).  Results equal pl_frame_is_in_frustum_batch + pl_orb_search_local_points_batch. */
typedef struct pl_localmap_view {
    int n;
    const float* world_pos;          /* n x 3  MapPoint::GetWorldPos */
    const float* normal;             /* n x 3  MapPoint::GetNormal */
    const uint8_t* desc;             /* n x 32 MapPoint::GetDescriptor */
    const float* min_dist_inv;       /* MapPoint::GetMinDistanceInvariance */
    const float* max_dist_inv;       /* MapPoint::GetMaxDistanceInvariance */
    const float* max_dist;           /* mfMaxDistance, read by PredictScale */
    const uint8_t* has_observations; /* nullable: all 1 */
} pl_localmap_view;
PL_API int pl_orb_search_local_map_batch(pl_match* h, int n, const pl_frame_view* F, const float* ow, int n_maps, const pl_localmap_view* maps,
                                         const int* map_of_frame, float viewing_cos_limit, float log_scale_factor, float th, float nn_ratio,
                                         int* const* match_of_feature, int* n_matches, int* n_in_view);

/* D2: LineMatcher::LineMatching predicate (LineMatcher.cpp:1463-1504) evaluated for all pairs, followed by the
 * "last matching i wins, every hit counts" rule of LineMatcher::SearchByProjection (LineMatcher.cpp:215-233) and
 * its relaxed retry (:235-261).  proj = already projected/clipped keylines (new_KeyLines) with descriptors;
 * cur = current frame's keylines.  cur_claimed[j]=1 skips j (:217-219).  match_of_line[j] = index into proj or -1.
 * (D3/D4/D5 differ only in how proj is produced; see pl_line_project.) */
PL_API int pl_line_match_pairs(pl_match* h, const pl_keyline* proj, const uint8_t* proj_desc, int n_proj,
                               const pl_keyline* cur, const uint8_t* cur_desc, const uint8_t* cur_claimed, int n_cur,
                               int* match_of_line, int* n_matches, int* used_relaxed);

/* D3/D4/D5 front half: project 3-D map lines with Tcw, handle the behind-camera endpoint, LiangBarsky clip and
 * UpdateKeyLineData (LineMatcher.cpp:96-212, :1389-1460, :1601-1624).  valid[i] gates line i.  Outputs the
 * compacted new_KeyLines, new_kl_index. */
PL_API int pl_line_project(pl_match* h, const double* start3d, const double* end3d, const pl_keyline* src_kl,
                           const uint8_t* valid, int n, const float tcw[12], float fx, float fy, float cx, float cy,
                           float min_x, float min_y, float max_x, float max_y, int img_cols, int img_rows,
                           pl_keyline* out_kl, int* out_index, int* n_out);

/* D3/D4/D5 complete: LineMatcher::SearchByProjection(Frame& Cur, const Frame& Last) (LineMatcher.cpp:72-269),
 * (Frame&, KeyFrame*) (:527-721) and (Frame&, const vector<MapLine*>&) (:755-952) — they differ only in which map lines
 * are offered and how `valid` is computed by the caller.  For each of the n calls: project + clip the 3-D lines with
 * the frame's Tcw, then all-pairs LineMatching with the relaxed retry.  match_of_line[i][j] = index of the ORIGINAL map
 * line (into lines[i], i.e. new_kl_index applied) assigned to current line j, or -1.  The optional outputs return the
 * reference's new_KeyLines / new_kl_index (what its test-only overloads expose). */
typedef struct {
    int n;
    const double* start3d;   /* MapLine::mStart3d, n x 3 */
    const double* end3d;     /* MapLine::mEnd3d,   n x 3 */
    const pl_keyline* kl;    /* the KeyLine the map line was created from (mvKeyLinesUn[i]) */
    const uint8_t* desc;     /* MapLine::mLineDescriptor, n x 32 */
    const uint8_t* valid;    /* pML && !outlier && !isBad() (:101-107), or mbTrackInView (:790-800) */
} pl_mapline_view;
typedef struct {
    int n;                   /* Frame::NL */
    const pl_keyline* kl;    /* Frame::mvKeyLinesUn */
    const uint8_t* desc;     /* Frame::mLineDescriptors, n x 32 */
    const uint8_t* claimed;  /* mvpMapLines[j] && Observations()>0 (:217-219); may be NULL */
    float tcw[12];
    float fx, fy, cx, cy, min_x, min_y, max_x, max_y;
    int cols, rows;          /* im_gray_ size used by UpdateKeyLineData (:1614-1623) */
} pl_lineframe_view;
PL_API int pl_line_search_by_projection_batch(pl_match* h, int n, const pl_lineframe_view* cur, const pl_mapline_view* lines,
                                              int* const* match_of_line, int* n_matches, int* used_relaxed,
                                              pl_keyline* const* new_keylines /* may be NULL */, int* const* new_kl_index /* may be NULL */,
                                              int* n_projected /* may be NULL */);

/* ------------------------------------------------------------------------------------------------------------
 * C4 / C5: pose-based projection searches.  The map points are described by what the reference reads from them.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    int n;
    const uint8_t* valid;        /* C4: pMP && !isBad() && !sAlreadyFound.count(pMP) (ORBmatcher.cc:1913-1918);
                                    C5: !isBad() && !spAlreadyFound.count(pMP) (:455-459)                     */
    const float* world_pos;      /* MapPoint::GetWorldPos(), n x 3                                            */
    const uint8_t* desc;         /* MapPoint::GetDescriptor(), n x 32                                         */
    const float* min_dist_inv;   /* GetMinDistanceInvariance() (0.8f * mfMinDistance, MapPoint.cc:395-399)    */
    const float* max_dist_inv;   /* GetMaxDistanceInvariance() (1.2f * mfMaxDistance, :401-405)               */
    const float* max_dist;       /* mfMaxDistance, what PredictScale divides (MapPoint.cc:416-431)            */
    const float* angle;          /* C4: pKF->mvKeysUn[i].angle (rotation histogram, :1976); may be NULL for C5 */
    const float* normal;         /* C5: GetNormal(), n x 3 (:494-497); may be NULL for C4                     */
} pl_posepoint_view;

/* C4: ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th,
 * ORBdist) — ORBmatcher.cc:1891-2024 (relocalisation refinement).  pts[i] = pKF->GetMapPointMatches() of call i;
 * cur[i].claimed[i2] = (CurrentFrame.mvpMapPoints[i2] != NULL) (:1965); ow = Frame::mOw, log_scale_factor =
 * Frame::mfLogScaleFactor.  match_of_feature[i][i2] = index of the KeyFrame feature whose map point was assigned to
 * feature i2 of the frame, or -1. */
PL_API int pl_orb_search_keyframe_points_batch(pl_match* h, int n, const pl_frame_view* cur, const pl_posepoint_view* pts,
                                               const float* ow /* n x 3 */, const float* log_scale_factor /* n */, float th,
                                               int orb_dist, int check_orientation, int* const* match_of_feature, int* n_matches);

/* C5: ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, vpPoints, vpMatched, th) — ORBmatcher.cc:423-554
 * (loop closing).  kf[i] is the KeyFrame as a frame view whose tcw holds Rcw | tcw AFTER the scale was divided out
 * (:435-439, done by the caller with the reference's own cv::Mat expressions) and whose claimed[idx] =
 * (vpMatched[idx] != NULL) (:521); ow = -Rcw^T tcw (:439).  match_of_feature[i][idx] = index into pts[i] of the map
 * point written to vpMatched[idx], or -1. */
PL_API int pl_orb_search_sim3_points_batch(pl_match* h, int n, const pl_frame_view* kf, const pl_posepoint_view* pts,
                                           const float* ow /* n x 3 */, const float* log_scale_factor /* n */, int th,
                                           int* const* match_of_feature, int* n_matches);

/* ------------------------------------------------------------------------------------------------------------
 * C6 / C7: ORBmatcher::SearchByBoW.  A DBoW2::FeatureVector (map<NodeId, vector<unsigned>>, FeatureVector.h:21-22)
 * travels flattened in key order: node_id[k] ascending, the features of node k = feat_idx[node_off[k]..node_off[k+1]).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    int n;                      /* features (KeyFrame::N / Frame::N)                                           */
    const float* angle;         /* mvKeysUn[i].angle (KeyFrame) / mvKeys[i].angle (Frame), n                    */
    const uint8_t* desc;        /* mDescriptors, n x 32                                                         */
    const uint8_t* valid;       /* pMP && !isBad() per feature (:301-305, :775-780, :797-802); NULL = all valid */
    int n_nodes;
    const unsigned int* node_id;
    const int* node_off;        /* n_nodes + 1                                                                  */
    const unsigned int* feat_idx;
} pl_bow_view;
/* mode 0 = C6 SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) — ORBmatcher.cc:247-410: a = key frame, b = frame
 *          (b.valid ignored); accept best <= TH_LOW && best < nn_ratio * second (:338-341);
 *          match_out[i] has b.n entries: index of the key-frame feature whose map point went to F feature j, or -1.
 * mode 1 = C7 SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) — :729-872: accept best < TH_LOW (strict, :814);
 *          match_out[i] has a.n entries: index of the KF2 feature whose map point went to vpMatches12[idx1], or -1. */
PL_API int pl_orb_search_bow_batch(pl_match* h, int n, const pl_bow_view* a, const pl_bow_view* b, int mode, float nn_ratio,
                                   int check_orientation, int* const* match_out, int* n_matches);

/* ------------------------------------------------------------------------------------------------------------
 * D6: the brute-force line matchers (cv::BFMatcher(NORM_HAMMING) + a rule).
 * ---------------------------------------------------------------------------------------------------------- */
/* LineMatcher::SearchByProjection(Frame&, KeyFrame*, vector<MapLine*>&) — LineMatcher.cpp:489-525: knnMatch(ref, cur, 2),
 * accept best/second < 0.75; match_of_line[j] (n_cur entries) = the LAST reference line i whose best match is j
 * (:511), n_matches counts every hit (:513).  n_cur < 2 (the reference reads out of bounds) gives no match. */
PL_API int pl_line_match_knn_ratio(pl_match* h, const uint8_t* ref_desc, int n_ref, const uint8_t* cur_desc, int n_cur,
                                   int* match_of_line, int* n_matches);
/* LineMatcher::SearchForTriangulation — LineMatcher.cpp:1174-1204 with KeyFrame::lineDescriptorMAD
 * (KeyFrame.cc:773-797): knnMatch(desc1, desc2, 2); accept query i when d2 - d1 > 0.1 * nn12_mad.
 * pairs = n1 x 2 ints, the accepted (queryIdx, trainIdx) in query order.  n1 == 0 or n2 < 2 gives no match. */
PL_API int pl_line_search_for_triangulation(pl_match* h, const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, int* pairs,
                                            int* n_matches, double* nn_mad, double* nn12_mad);
/* LineMatcher::Fuse, the descriptor half of the active branch — LineMatcher.cpp:1207-1379: for every valid map line
 * the nearest key-frame line (BruteForce-Hamming match of one row) and the 1.5 * min(100, dist) rule (:1300-1312).
 * tdx[i] = matched key-frame line or -1; n_fused = number of hits.  The MapLine::Replace bookkeeping stays with
 * the caller. */
PL_API int pl_line_fuse_candidates(pl_match* h, const uint8_t* ml_desc, const uint8_t* valid, int n, const uint8_t* kf_desc,
                                   int n_kf, int* tdx, int* n_fused);

/* ------------------------------------------------------------------------------------------------------------
 * E: the remaining ORBmatcher entry points — the LocalMapping / LoopClosing / Initializer callers.
 * ---------------------------------------------------------------------------------------------------------- */
/* E1: ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) — ORBmatcher.cc:1107-1277 (variant 0) and
 * ORBmatcher::Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint) — :1290-1427 (variant 1): the projection search
 * of every map point into the key frame.  The points do not interact (no claims), so every one is searched at once.
 * kf[i] = the key frame (tcw = Rcw | tcw; for variant 1 after the scale was divided out by the caller, :1299-1302);
 * pts[i].valid = pMP && !isBad() && !IsInKeyFrame(pKF) (:1128-1133) / !isBad() && !spAlreadyFound.count(pMP) (:1317-1320);
 * ow = camera centre; inv_level_sigma2 = pKF->mvInvLevelSigma2 (kf[i].n_levels floats per call, used by the chi-square
 * gates :1217-1235 of variant 0; may be NULL for variant 1).
 * best_idx[i][k] = key-frame feature the point k fuses with (bestDist <= TH_LOW, :1249 / :1404) or -1;
 * best_dist[i][k] = its distance (256 when no candidate survived).  n_fused[i] = number of hits == the reference's return
 * value when vpMapPoints holds no duplicates.  The Replace / AddObservation bookkeeping (:1251-1272, :1406-1418) stays
 * with the caller, which walks best_idx in order. */
PL_API int pl_orb_fuse_candidates_batch(pl_match* h, int n, const pl_frame_view* kf, const pl_posepoint_view* pts, const float* ow /* n x 3 */,
                                        const float* log_scale_factor /* n */, const float* const* inv_level_sigma2 /* n pointers or NULL */,
                                        float th, int variant, int* const* best_idx, int* const* best_dist /* may be NULL */, int* n_fused);

/* E2: ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) — ORBmatcher.cc:1441-1692.
 * kf1 / kf2: the key frames (tcw = R1w | t1w and R2w | t2w; intrinsics of kf1 are used for both projections as in :1445-1448).
 * pts1 / pts2 = GetMapPointMatches() of each; valid = pMP && !vbAlreadyMatched && !isBad() (:1477-1484, :1551-1557).
 * t21 = sR21 | t21 and t12 = sR12 | t12 as 3x4 row-major, evaluated by the caller with the reference's cv::Mat
 * expressions (:1457-1460).  match12[i1] = feature of kf2 whose map point is written to vpMatches12[i1] (:1632-1647), or -1. */
PL_API int pl_orb_search_by_sim3(pl_match* h, const pl_frame_view* kf1, const pl_frame_view* kf2, const pl_posepoint_view* pts1,
                                 const pl_posepoint_view* pts2, const float t21[12], const float t12[12], float log_scale_factor1,
                                 float log_scale_factor2, float th, int* match12, int* n_found);

/* E3: ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) — ORBmatcher.cc:573-717.
 * Only claimed / tcw-independent fields of the views are read.  prev_matched (F1.n x 2 floats, in/out) = vbPrevMatched;
 * matches12 (F1.n) = vnMatches12.  F2.n <= 20000 (the replay keeps vMatchedDistance / vnMatches21 in shared memory). */
PL_API int pl_orb_search_for_initialization(pl_match* h, const pl_frame_view* F1, const pl_frame_view* F2, float* prev_matched,
                                            int window_size, float nn_ratio, int check_orientation, int* matches12, int* n_matches);

/* E4: ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) — ORBmatcher.cc:884-1095 with
 * CheckDistEpipolarLine (:205-232).  a / b: bow.valid[i] = (GetMapPoint(i) == NULL) (:938-940, :964-966), bow.angle =
 * mvKeysUn[i].angle; keys_un / u_right = mvKeysUn / mvuRight.  cw1 = pKF1->GetCameraCenter(); kf2_tcw = R2w | t2w; the
 * epipole (:897-903) is evaluated here like the reference's cv::Mat expression.  pairs receives n_matches (idx1, idx2)
 * pairs in ascending idx1 (:1085-1091); capacity a.bow.n pairs. */
typedef struct {
    pl_bow_view bow;
    const pl_keypoint* keys_un;
    const float* u_right;
} pl_triang_view;
PL_API int pl_orb_search_for_triangulation(pl_match* h, const pl_triang_view* a, const pl_triang_view* b, const float f12[9],
                                           const float cw1[3], const float kf2_tcw[12], float fx2, float fy2, float cx2, float cy2,
                                           const float* scale_factors2, const float* level_sigma2_2, int n_levels2, int only_stereo,
                                           int check_orientation, int* pairs, int* n_matches);

/* E5: MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:256-321) == MapLine::ComputeDistinctiveDescriptors
 * (MapLine.cpp:269-330) for n_groups map points / lines at once: group g owns the descriptor rows
 * [group_off[g], group_off[g+1]) (the observations whose key frame is not bad, in map order).  best_row[g] = row (relative to
 * the group) with the least median distance to the others (median = sorted[int(0.5*(N-1))], first minimum wins), or -1 for
 * an empty group. */
PL_API int pl_distinctive_descriptors(pl_match* h, const uint8_t* desc, const int* group_off, int n_groups, int* best_row);

/* ------------------------------------------------------------------------------------------------------------
 * F: per-feature maps of Frame that sit between the extractors and the matchers.  Pure
 *    element-wise work; batched over the frames of a sequence so that one call covers what the reference does in
 *    N Frame constructors / N SearchLocalPoints passes.  Frame f owns the features [off[f], off[f+1]).
 * ---------------------------------------------------------------------------------------------------------- */
/* F1: Frame::UndistortKeyPoints (Frame.cc:737-765) / UndistortKeyLines endpoints (:767-800) ==
 * cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) for n (x, y) pairs; dist_coef = k1 k2 p1 p2 k3
 * (Tracking.cc:79-91).  k1 == 0 copies the input, as the reference does (:739-743). */
PL_API int pl_frame_undistort_points(pl_match* h, const float* xy, int n, float fx, float fy, float cx, float cy, const float dist_coef[5],
                                     float* xy_out);
/* Frame::UndistortKeyLines (Frame.cc:767-845): both end points through cv::undistortPoints, then the derived KeyLine fields
 * (pt, lineLength, numOfPixels = cv::LineIterator(...).count on the img_cols x img_rows image, angle, size, response) are
 * recomputed from them; every other field is copied.  k1 == 0 copies the key lines (:768-771). */
PL_API int pl_frame_undistort_keylines(pl_match* h, const pl_keyline* kls, int n, float fx, float fy, float cx, float cy, const float dist_coef[5],
                                       int img_cols, int img_rows, pl_keyline* out);
/* Frame::AssignFeaturesToGrid (Frame.cc:265-287) with PosInGrid (:527-538): mGrid flattened — cell c = x * 48 + y (64 x 48
 * cells over bounds = mnMinX, mnMinY, mnMaxX, mnMaxY) owns sorted_idx[cell_start[c] .. cell_start[c + 1]) in ascending
 * feature index; features outside the grid are dropped.  cell_start has 64 * 48 + 1 entries, sorted_idx n. */
PL_API int pl_frame_assign_features_to_grid(pl_match* h, const pl_keypoint* keys_un, int n, const float bounds[4], int* cell_start, int* sorted_idx);
/* F2: Frame::ComputeStereoFromRGBD (Frame.cc:1065-1117): d = imDepth.at<float>(v, u) at the truncated DISTORTED position
 * xy (mvKeys[i].pt, or a KeyLine end point), depth_out = d and u_right_out = x_un - bf / d when d > 0, else -1 / -1.
 * depth = n_frames float images (rows x cols, row stride step_bytes, frame f at depth + f * frame_stride_bytes).
 * depth_is_device != 0: `depth` is a device pointer (images already resident in HBM), everything else stays host. */
PL_API int pl_frame_stereo_from_rgbd_batch(pl_match* h, int n_frames, const float* depth, int depth_is_device, int rows, int cols, size_t step_bytes,
                                           size_t frame_stride_bytes, const int* off, const float* xy, const float* x_un, float bf, float* depth_out,
                                           float* u_right_out);
/* F3: Frame::UnprojectStereo (Frame.cc:1120-1134) == UnprojectStereoLine{,Start,End} (:1140-1205) per end point:
 * x = (u-cx)*z*invfx, y = (v-cy)*z*invfy, world = mRwc * (x, y, z) + mOw for z > 0 (world is left 0 and valid = 0 otherwise).
 * rwc = n_frames x 9 (row-major mRwc), ow = n_frames x 3; invfx = 1.0f / fx as Frame.cc computes it. */
PL_API int pl_frame_unproject_batch(pl_match* h, int n_frames, const int* off, const float* xy_un, const float* z, const float* rwc, const float* ow,
                                    float fx, float fy, float cx, float cy, float* world /* total x 3 */, uint8_t* valid /* total */);
/* F4: Frame::IsInFrustum(MapPoint*, viewingCosLimit) (Frame.cc:345-401) with MapPoint::PredictScale (MapPoint.cc:416-431) for
 * n_frames frames against one snapshot of m map points: the fields it leaves on the map point, as [n_frames x m] planes —
 * exactly the inputs of pl_orb_search_local_points (C2).  tcw = n_frames x 12 (mRcw | mtcw), ow = n_frames x 3,
 * bounds = mnMinX, mnMinY, mnMaxX, mnMaxY.  Planes other than in_view are only defined where in_view = 1. */
PL_API int pl_frame_is_in_frustum_batch(pl_match* h, int n_frames, const float* tcw, const float* ow, float fx, float fy, float cx, float cy, float bf,
                                        const float bounds[4], int n_levels, float log_scale_factor, int m, const float* world_pos,
                                        const float* normal, const float* min_dist_inv, const float* max_dist_inv, const float* max_dist,
                                        float viewing_cos_limit, uint8_t* in_view, float* proj_x, float* proj_y, float* proj_xr, int* scale_level,
                                        float* view_cos);
/* Frame::IsInFrustum(MapLine*, viewingCosLimit) (Frame.cc:403-430): a map line is in view unless both end points are behind
 * the camera.  start3d / end3d = m x 3 doubles (MapLine::mStart3d / mEnd3d), converted to float as the reference does. */
PL_API int pl_frame_lines_in_frustum_batch(pl_match* h, int n_frames, const float* tcw, int m, const double* start3d, const double* end3d,
                                           uint8_t* in_view /* n_frames x m */);

/* F5: Frame::ComputeStereoMatches (Frame.cc:888-1062): per left key point the best right key point in
 * its row band (Hamming, level +-1, disparity range), the 11x11 SAD sub-pixel refinement over 11 shifts read from BOTH
 * extractors' mvImagePyramid (which never leaves the device), and the final 1.5*1.4*median outlier rule.  left / right = the two
 * extractors after operator() on the rectified pair (frame `frame` of their last call); keys = mvKeys / mvKeysRight (image
 * coordinates), bf = mbf, b = mb.  u_right / depth = mvuRight / mvDepth (-1 = no match).  n_left, n_right <= 16384.
 * Returns PL_ERR_CAPACITY (outputs filled) when one of the extractors had a frame over its key point capacity pending from the
 * device-pointer API: this call waits for both extractors and therefore consumes that report. */
PL_API int pl_frame_compute_stereo_matches(pl_match* h, pl_orb* left, pl_orb* right, int frame, const pl_keypoint* keys_left,
                                           const uint8_t* desc_left, int n_left, const pl_keypoint* keys_right, const uint8_t* desc_right,
                                           int n_right, float bf, float b, float* u_right, float* depth);

/* ------------------------------------------------------------------------------------------------------------
 * G: Frame::ComputeBoW (Frame.cc:721-735) — DBoW2 TemplatedVocabulary<FORB>::transform(features, mBowVec, mFeatVec, 4)
 *    (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1194, :1218-1259; FORB::distance FORB.cpp:85-110).  SURVEY.md §8(f)
 *    rank 1: the same Hamming kernel walks the vocabulary tree (k children per level, first minimum wins) and produces the
 *    FeatureVector that SearchByBoW consumes, flattened exactly like pl_bow_view.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct pl_voc pl_voc;
/* The tree as loadFromTextFile builds it (TemplatedVocabulary.h:1338-1422): node ids 1..n_nodes in file order (0 = root),
 * parent[i] < i+1, is_leaf / 32 descriptor bytes / weight per node; word ids are assigned to leaves in file order.
 * scoring / weighting = the DBoW2 enums (BowVector.h:36-53; ORBvoc.txt: 0 0 = L1_NORM, TF_IDF).  A node has <= 32 children. */
PL_API int pl_voc_create(pl_voc** out, int device, int k, int L, int scoring, int weighting, int n_nodes, const int* parent,
                         const uint8_t* is_leaf, const uint8_t* desc, const double* weight);
/* ORBVocabulary::loadFromTextFile(strVocFile) (System.cc:67).  Blank lines are skipped (the reference turns a trailing blank
 * line into a node with an uninitialised descriptor). */
PL_API int pl_voc_load_text(pl_voc** out, int device, const char* filename);
PL_API void pl_voc_destroy(pl_voc* v);
PL_API int pl_voc_info(const pl_voc* v, int* k, int* L, int* n_nodes /* incl. root */, int* n_words);
/* transform() for n_frames frames: frame f owns the descriptor rows [off[f], off[f+1]) (<= 8192 per frame).  All outputs are
 * segmented like the input: the BowVector of frame f is word_id / word_value[off[f] .. off[f] + n_words[f]) in ascending word
 * id; its FeatureVector is node_id[off[f] .. off[f] + n_fv_nodes[f]) ascending, node k owning
 * feat_idx[off[f] + node_off[off[f] + f + k] .. off[f] + node_off[off[f] + f + k + 1]) (frame-relative feature indices,
 * ascending) — node_off has total + n_frames entries.  A leaf shallower than L - levelsup leaves the reference's nid
 * uninitialised; here such a feature is filed under node 0. */
PL_API int pl_voc_transform_batch(pl_voc* v, int n_frames, const int* off, const uint8_t* desc, int levelsup, int* n_words, unsigned int* word_id,
                                  double* word_value, int* n_fv_nodes, unsigned int* node_id, int* node_off, unsigned int* feat_idx);

#ifdef __cplusplus
}
#endif
#endif /* PLSLAM_C_H */
