/* oracle.h — C interface of the CPU oracle (TEST INFRASTRUCTURE ONLY; see the header of each .cpp).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it. */
#ifndef PL_ORACLE_H
#define PL_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/plslam_c.h"
#ifdef __cplusplus
extern "C" {
#endif
#define ORC_API __attribute__((visibility("default")))

/* ---- OpenCV primitives restated (pinned to cv2 4.13 by tests/test_oracle_orb.py) ---- */
ORC_API float orc_fast_atan2(float y, float x);
ORC_API void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep);
ORC_API void orc_border_reflect101_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep, int border);
ORC_API void orc_gaussian_blur_fixed_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep, const int* kernel, int ksize);
ORC_API void orc_gaussian_blur7_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep);
ORC_API int orc_fast_score_9_16(const uint8_t* p, size_t step);
ORC_API int orc_fast_detect(const uint8_t* img, int w, int h, size_t step, int threshold, int nonmax, float* xs, float* ys, float* resp, int cap);

/* ---- ORBextractor ---- */
typedef struct orc_orb orc_orb;
ORC_API orc_orb* orc_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh);
ORC_API void orc_orb_destroy(orc_orb* o);
ORC_API void orc_orb_tables(const orc_orb* o, float* sf, float* invsf, float* sigma2, float* invsigma2, int* perLevel, int* umax);
ORC_API int orc_distribute_octtree(const float* xs, const float* ys, const float* resp, int n, int minX, int maxX, int minY, int maxY, int N, float* oxs, float* oys, float* oresp, int cap);
ORC_API int orc_orb_extract(orc_orb* o, const uint8_t* img, int rows, int cols, size_t step, pl_keypoint* kps, uint8_t* desc, int cap, int* n_out);
ORC_API int orc_orb_level_dims(const orc_orb* o, int level, int* w, int* h);
ORC_API int orc_orb_level_bordered(const orc_orb* o, int level, uint8_t* out);
ORC_API int orc_orb_level_blurred(const orc_orb* o, int level, uint8_t* out);
ORC_API int orc_orb_level_candidates(const orc_orb* o, int level, float* xs, float* ys, float* resp, int cap);

/* ---- descriptor search (match_oracle.cpp) ---- */
ORC_API int orc_descriptor_distance(const uint8_t* a, const uint8_t* b);
ORC_API void orc_compute_three_maxima(const int* counts, int L, int* ind);
ORC_API void orc_hamming_pairs(const uint8_t* a, const uint8_t* b, int n, int* dist);
ORC_API void orc_hamming_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx, int* dist, int threads);
ORC_API void orc_hamming_candidates(const uint8_t* q, int nq, const uint8_t* t, const int* off, const int* cidx, int* dist);

/* ---- projection searches / line matching over the POD views of plslam_c.h (match_oracle.cpp) ---- */
ORC_API int orc_orb_search_local_points(const pl_frame_view* F, const pl_mappoint_view* M, float th, float nn_ratio, int* match_of_feature, int* n_matches);
ORC_API int orc_orb_search_last_frame(const pl_frame_view* C, const pl_lastframe_view* L, float th, int mono, int check_orientation, int* match_of_feature, int* n_matches);
ORC_API int orc_line_iterator_count(float x0, float y0, float x1, float y1, int cols, int rows);
ORC_API int orc_line_project(const double* start3d, const double* end3d, const pl_keyline* src_kl, const uint8_t* valid, int n, const float tcw[12], float fx, float fy, float cx, float cy, float min_x, float min_y, float max_x, float max_y, int img_cols, int img_rows, pl_keyline* out_kl, int* out_index, int* n_out);
ORC_API int orc_line_match_pairs(const pl_keyline* proj, const uint8_t* proj_desc, int n_proj, const pl_keyline* cur, const uint8_t* cur_desc, const uint8_t* cur_claimed, int n_cur, int* match_of_line, int* n_matches, int* used_relaxed);
ORC_API int orc_orb_search_keyframe_points(const pl_frame_view* C, const pl_posepoint_view* P, const float* ow, float log_scale_factor, float th, int orb_dist, int check_orientation, int* match_of_feature, int* n_matches);
ORC_API int orc_orb_search_sim3_points(const pl_frame_view* K, const pl_posepoint_view* P, const float* ow, float log_scale_factor, int th, int* match_of_feature, int* n_matches);
ORC_API int orc_orb_search_bow(const pl_bow_view* A, const pl_bow_view* B, int mode, float nn_ratio, int check_orientation, int* match_out, int* n_matches);
ORC_API int orc_line_match_knn_ratio(const uint8_t* ref_desc, int n_ref, const uint8_t* cur_desc, int n_cur, int* match_of_line, int* n_matches);
ORC_API int orc_line_search_for_triangulation(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, int* pairs, int* n_matches, double* nn_mad, double* nn12_mad);
ORC_API int orc_line_fuse_candidates(const uint8_t* ml_desc, const uint8_t* valid, int n, const uint8_t* kf_desc, int n_kf, int* tdx, int* n_fused);
ORC_API int orc_orb_fuse_candidates(const pl_frame_view* K, const pl_posepoint_view* P, const float* ow, float log_scale_factor, const float* inv_level_sigma2, float th, int variant, int* best_idx, int* best_dist, int* n_fused);
ORC_API int orc_orb_search_by_sim3(const pl_frame_view* K1, const pl_frame_view* K2, const pl_posepoint_view* P1, const pl_posepoint_view* P2, const float* t21, const float* t12, float log_sf1, float log_sf2, float th, int* match12, int* n_found);
ORC_API int orc_orb_search_for_initialization(const pl_frame_view* F1, const pl_frame_view* F2, float* prev_matched, int window_size, float nn_ratio, int check_orientation, int* matches12, int* n_matches);
ORC_API int orc_orb_search_for_triangulation(const pl_triang_view* A, const pl_triang_view* B, const float* F12, const float* cw1, const float* kf2_tcw, float fx2, float fy2, float cx2, float cy2, const float* scale_factors2, const float* level_sigma2_2, int n_levels2, int only_stereo, int check_orientation, int* pairs, int* n_matches);
ORC_API int orc_distinctive_descriptors(const uint8_t* desc, const int* group_off, int n_groups, int* best_row);

/* ---- Frame glue (frame_oracle.cpp) ---- */
ORC_API int orc_frame_undistort_points(const float* xy, int n, float fx, float fy, float cx, float cy, const float* dist_coef, float* xy_out);
ORC_API int orc_frame_stereo_from_rgbd_batch(int n_frames, const float* depth, int rows, int cols, size_t step_bytes, size_t frame_stride_bytes, const int* off, const float* xy, const float* x_un, float bf, float* depth_out, float* u_right_out);
ORC_API int orc_frame_unproject_batch(int n_frames, const int* off, const float* xy_un, const float* z, const float* rwc, const float* ow, float fx, float fy, float cx, float cy, float* world, uint8_t* valid);
ORC_API int orc_frame_is_in_frustum_batch(int n_frames, const float* tcw, const float* ow, float fx, float fy, float cx, float cy, float bf, const float* bounds, int n_levels, float log_scale_factor, int m, const float* world_pos, const float* normal, const float* min_dist_inv, const float* max_dist_inv, const float* max_dist, float viewing_cos_limit, uint8_t* in_view, float* proj_x, float* proj_y, float* proj_xr, int* scale_level, float* view_cos);
ORC_API int orc_frame_compute_stereo_matches(const orc_orb* left, const orc_orb* right, const pl_keypoint* keysL, const uint8_t* descL, int N, const pl_keypoint* keysR, const uint8_t* descR, int Nr, float mbf, float mb, float* mvuRight, float* mvDepth);
ORC_API int orc_frame_undistort_keylines(const pl_keyline* kls, int n, float fx, float fy, float cx, float cy, const float* dist_coef, int img_cols, int img_rows, pl_keyline* out);
ORC_API int orc_frame_assign_features_to_grid(const pl_keypoint* keys_un, int n, const float* bounds, int* cell_start, int* sorted_idx);
ORC_API int orc_frame_lines_in_frustum_batch(int n_frames, const float* tcw, int m, const double* start3d, const double* end3d, uint8_t* in_view);

/* ---- DBoW2 vocabulary transform (bow_oracle.cpp) ---- */
typedef struct orc_voc orc_voc;
ORC_API orc_voc* orc_voc_create(int k, int L, int scoring, int weighting, int n_nodes, const int* parent, const uint8_t* is_leaf, const uint8_t* desc, const double* weight);
ORC_API orc_voc* orc_voc_load_text(const char* filename);
ORC_API void orc_voc_destroy(orc_voc* v);
ORC_API int orc_voc_info(const orc_voc* v, int* k, int* L, int* n_nodes, int* n_words);
ORC_API int orc_voc_transform(const orc_voc* v, const uint8_t* desc, int n, int levelsup, int* n_words, unsigned* word_id, double* word_value, int* n_fv_nodes, unsigned* node_id, int* node_off, unsigned* feat_idx);

/* ---- line extraction (line_oracle.cpp) ---- */
ORC_API int orc_lsd_trace(const uint8_t* img, int rows, int cols, size_t step, int* out4, int cap);
ORC_API int orc_lsd_detect(const uint8_t* img, int rows, int cols, size_t step, int order_mode, float* xyxy, double* width, double* prec, double* nfa, int cap);
ORC_API int orc_lsd_angles(const uint8_t* img, int rows, int cols, size_t step, double* out, int* ow, int* oh);
ORC_API int orc_lsd_scaled(const uint8_t* img, int rows, int cols, size_t step, uint8_t* out, int* ow, int* oh);
ORC_API int orc_lbd_compute(const uint8_t* img, int rows, int cols, size_t step, const pl_keyline* kls, int n, uint8_t* desc, float* fdesc);
ORC_API int orc_line_extract(const uint8_t* img, int rows, int cols, size_t step, int max_lines, int order_mode, pl_keyline* kls, uint8_t* desc, double* coeffs, int* n_out);

#ifdef __cplusplus
}
#endif
#endif
