// frame_kernels.cu — the per-feature maps of ORB_SLAM2::Frame between the extractors and the matchers on sm_100a
// (SURVEY.md §8(f) rank 4): the CUDA path behind pl_frame_* (include/plslam_c.h).
//
// Reference functions replaced:
//   Frame::UndistortKeyPoints / UndistortKeyLines end points   src/Frame.cc:737-765, 767-800  (cv::undistortPoints)
//   Frame::ComputeStereoFromRGBD                               src/Frame.cc:1065-1117
//   Frame::UnprojectStereo / UnprojectStereoLine{,Start,End}   src/Frame.cc:1120-1205
//   Frame::IsInFrustum(MapPoint*) + MapPoint::PredictScale     src/Frame.cc:345-401, src/MapPoint.cc:416-431
//   Frame::IsInFrustum(MapLine*)                               src/Frame.cc:403-430
//
// All of it is element-wise: one thread per feature (or per frame x map point), coalesced SoA planes, batched over the
// frames of a sequence.  The library is compiled with -fmad=false, so the double / float expressions below round exactly
// like the reference's scalar C++ (cv::Mat products: double accumulation, rounded once).
#include "match_common.cuh"

namespace pl {

__global__ void __launch_bounds__(256) k_undistort(const float2* __restrict__ xy, int n, double fx, double fy, double cx, double cy, double k0,
                                                   double k1, double k2, double k3, double k4, float2* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 p = xy[i];
    out[i] = undistort_point(p.x, p.y, fx, fy, cx, cy, k0, k1, k2, k3, k4);
}

// frame of feature i: binary search in the offsets (n_frames + 1 entries)
__device__ __forceinline__ int frame_of(const int* __restrict__ off, int n_frames, int i) {
    int lo = 0, hi = n_frames;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_stereo_from_rgbd(const uint8_t* __restrict__ depth, size_t step, size_t frame_stride, const int* __restrict__ off,
                                                          int n_frames, int total, const float2* __restrict__ xy, const float* __restrict__ x_un,
                                                          float bf, float* __restrict__ depth_out, float* __restrict__ u_right_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int f = frame_of(off, n_frames, i);
    const float2 p = xy[i];
    const float d = *(const float*)(depth + (size_t)f * frame_stride + (size_t)(int)p.y * step + (size_t)(int)p.x * 4);
    const bool ok = d > 0;
    depth_out[i] = ok ? d : -1.f;
    u_right_out[i] = ok ? __fsub_rn(x_un[i], __fdiv_rn(bf, d)) : -1.f;
}

// cv::Mat (CV_32F) R*x + t, one output row: double accumulation, rounded once
__device__ __forceinline__ float gemm_row(float r0, float r1, float r2, float t, float x, float y, float z) {
    double s = 0;
    s = __dadd_rn(s, __dmul_rn((double)r0, (double)x));
    s = __dadd_rn(s, __dmul_rn((double)r1, (double)y));
    s = __dadd_rn(s, __dmul_rn((double)r2, (double)z));
    return (float)__dadd_rn(s, (double)t);
}

__global__ void __launch_bounds__(256) k_unproject(const int* __restrict__ off, int n_frames, int total, const float2* __restrict__ xy_un,
                                                   const float* __restrict__ z, const float* __restrict__ rwc, const float* __restrict__ ow, float cx,
                                                   float cy, float invfx, float invfy, float* __restrict__ world, uint8_t* __restrict__ valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float zz = z[i];
    float wx = 0, wy = 0, wz = 0;
    const bool ok = zz > 0;
    if (ok) {
        const int f = frame_of(off, n_frames, i);
        const float* R = rwc + 9 * (size_t)f;
        const float* O = ow + 3 * (size_t)f;
        const float2 p = xy_un[i];
        const float x = __fmul_rn(__fmul_rn(__fsub_rn(p.x, cx), zz), invfx);
        const float y = __fmul_rn(__fmul_rn(__fsub_rn(p.y, cy), zz), invfy);
        wx = gemm_row(R[0], R[1], R[2], O[0], x, y, zz);
        wy = gemm_row(R[3], R[4], R[5], O[1], x, y, zz);
        wz = gemm_row(R[6], R[7], R[8], O[2], x, y, zz);
    }
    world[3 * (size_t)i] = wx;
    world[3 * (size_t)i + 1] = wy;
    world[3 * (size_t)i + 2] = wz;
    valid[i] = ok ? 1 : 0;
}

struct FrustumArgs {
    float fx, fy, cx, cy, bf, min_x, min_y, max_x, max_y, log_sf, cos_limit;
    int n_levels, m;
};

// grid.y = frame, one thread per map point; the map-point planes are read once per frame row (they stay in L2)
__global__ void __launch_bounds__(256) k_is_in_frustum(const float* __restrict__ tcw, const float* __restrict__ ow, FrustumArgs A,
                                                       const float* __restrict__ world_pos, const float* __restrict__ normal,
                                                       const float* __restrict__ min_inv, const float* __restrict__ max_inv,
                                                       const float* __restrict__ max_raw, uint8_t* __restrict__ in_view, float* __restrict__ proj_x,
                                                       float* __restrict__ proj_y, float* __restrict__ proj_xr, int* __restrict__ scale_level,
                                                       float* __restrict__ view_cos) {
    __shared__ float sT[15];
    const int f = blockIdx.y;
    if (threadIdx.x < 12) sT[threadIdx.x] = tcw[12 * (size_t)f + threadIdx.x];
    else if (threadIdx.x < 15) sT[threadIdx.x] = ow[3 * (size_t)f + threadIdx.x - 12];
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.m) return;
    const size_t o = (size_t)f * A.m + j;
    const float X = world_pos[3 * (size_t)j], Y = world_pos[3 * (size_t)j + 1], Z = world_pos[3 * (size_t)j + 2];
    bool ok = false;
    float u = 0, v = 0, xr = 0, vc = 0;
    int lvl = 0;
    const float pcx = gemm_row(sT[0], sT[1], sT[2], sT[3], X, Y, Z);
    const float pcy = gemm_row(sT[4], sT[5], sT[6], sT[7], X, Y, Z);
    const float pcz = gemm_row(sT[8], sT[9], sT[10], sT[11], X, Y, Z);
    if (!(pcz < 0.0f)) {
        const float invz = __fdiv_rn(1.0f, pcz);
        u = __fadd_rn(__fmul_rn(__fmul_rn(A.fx, pcx), invz), A.cx);
        v = __fadd_rn(__fmul_rn(__fmul_rn(A.fy, pcy), invz), A.cy);
        if (!(u < A.min_x || u > A.max_x) && !(v < A.min_y || v > A.max_y)) {
            const float px = __fsub_rn(X, sT[12]), py = __fsub_rn(Y, sT[13]), pz = __fsub_rn(Z, sT[14]);
            double ss = __dmul_rn((double)px, (double)px);
            ss = __dadd_rn(ss, __dmul_rn((double)py, (double)py));
            ss = __dadd_rn(ss, __dmul_rn((double)pz, (double)pz));
            const float dist = (float)sqrt(ss);
            if (!(dist < min_inv[j] || dist > max_inv[j])) {
                double dot = __dmul_rn((double)px, (double)normal[3 * (size_t)j]);
                dot = __dadd_rn(dot, __dmul_rn((double)py, (double)normal[3 * (size_t)j + 1]));
                dot = __dadd_rn(dot, __dmul_rn((double)pz, (double)normal[3 * (size_t)j + 2]));
                vc = (float)(dot / (double)dist);
                if (!(vc < A.cos_limit)) {
                    const float ratio = __fdiv_rn(max_raw[j], dist);
                    lvl = (int)ceilf(__fdiv_rn(glibc_logf(ratio), A.log_sf));
                    if (lvl < 0) lvl = 0;
                    else if (lvl >= A.n_levels) lvl = A.n_levels - 1;
                    xr = __fsub_rn(u, __fmul_rn(A.bf, invz));
                    ok = true;
                }
            }
        }
    }
    in_view[o] = ok ? 1 : 0;
    proj_x[o] = ok ? u : 0.f;
    proj_y[o] = ok ? v : 0.f;
    proj_xr[o] = ok ? xr : 0.f;
    scale_level[o] = ok ? lvl : 0;
    view_cos[o] = ok ? vc : 0.f;
}

__global__ void __launch_bounds__(256) k_lines_in_frustum(const float* __restrict__ tcw, int m, const double* __restrict__ s3, const double* __restrict__ e3,
                                                          uint8_t* __restrict__ in_view) {
    const int f = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const float* T = tcw + 12 * (size_t)f;
    const float zs = gemm_row(T[8], T[9], T[10], T[11], (float)s3[3 * (size_t)j], (float)s3[3 * (size_t)j + 1], (float)s3[3 * (size_t)j + 2]);
    const float ze = gemm_row(T[8], T[9], T[10], T[11], (float)e3[3 * (size_t)j], (float)e3[3 * (size_t)j + 1], (float)e3[3 * (size_t)j + 2]);
    in_view[(size_t)f * m + j] = (zs < 0.0f && ze < 0.0f) ? 0 : 1;
}

}  // namespace pl

using namespace pl;

namespace {
inline size_t padb(size_t b) { return PlStage::pad(b); }
int check_offsets(const int* off, int n_frames) {
    PL_CHECK_ARG(off && off[0] == 0);
    for (int f = 0; f < n_frames; f++) PL_CHECK_ARG(off[f + 1] >= off[f]);
    return PL_OK;
}
}  // namespace

extern "C" {

PL_API int pl_frame_undistort_points(pl_match* h, const float* xy, int n, float fx, float fy, float cx, float cy, const float dist_coef[5],
                                     float* xy_out) {
    PL_CHECK_ARG(h && n >= 0 && dist_coef && (n == 0 || (xy && xy_out)) && fx != 0.f && fy != 0.f);
    if (n == 0) return PL_OK;
    if (dist_coef[0] == 0.0f) {  // Frame.cc:739-743: mvKeysUn = mvKeys
        memmove(xy_out, xy, (size_t)n * 8);
        return PL_OK;
    }
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    int rc = h->in.reserve(padb((size_t)n * 8) * 2);
    if (rc != PL_OK) return rc;
    const float2* d_xy = (const float2*)h->in.put(xy, (size_t)n * 2);
    float* h_out;
    float2* d_out = (float2*)h->in.out<float>((size_t)n * 2, &h_out);
    cudaStream_t st = h->stream;
    if ((rc = h->in.upload(st)) != PL_OK) return rc;
    k_undistort<<<(n + 255) / 256, 256, 0, st>>>(d_xy, n, (double)fx, (double)fy, (double)cx, (double)cy, (double)dist_coef[0], (double)dist_coef[1],
                                                 (double)dist_coef[2], (double)dist_coef[3], (double)dist_coef[4], d_out);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(h_out, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    memcpy(xy_out, h_out, (size_t)n * 8);
    return PL_OK;
}

PL_API int pl_frame_stereo_from_rgbd_batch(pl_match* h, int n_frames, const float* depth, int depth_is_device, int rows, int cols, size_t step_bytes,
                                           size_t frame_stride_bytes, const int* off, const float* xy, const float* x_un, float bf, float* depth_out,
                                           float* u_right_out) {
    PL_CHECK_ARG(h && n_frames >= 0 && rows > 0 && cols > 0 && step_bytes >= (size_t)cols * 4 && frame_stride_bytes >= step_bytes * (size_t)rows);
    if (n_frames == 0) return PL_OK;
    int rc = check_offsets(off, n_frames);
    if (rc != PL_OK) return rc;
    const int total = off[n_frames];
    if (total == 0) return PL_OK;
    PL_CHECK_ARG(depth && xy && x_un && depth_out && u_right_out);
    // imDepth.at<float>(v, u) has no bounds check in the reference; here a position outside the image is an argument error
    for (int i = 0; i < total; i++)
        PL_CHECK_ARG((int)xy[2 * i] >= 0 && (int)xy[2 * i] < cols && (int)xy[2 * i + 1] >= 0 && (int)xy[2 * i + 1] < rows && xy[2 * i] > -1.f && xy[2 * i + 1] > -1.f);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    const size_t img_bytes = depth_is_device ? 0 : frame_stride_bytes * (size_t)n_frames;
    if ((rc = h->in.reserve(padb(img_bytes) + padb((size_t)(n_frames + 1) * 4) + padb((size_t)total * 8) + padb((size_t)total * 4) * 3)) != PL_OK) return rc;
    const uint8_t* d_depth = depth_is_device ? (const uint8_t*)depth : h->in.put((const uint8_t*)depth, img_bytes);
    const int* d_off = h->in.put(off, (size_t)n_frames + 1);
    const float2* d_xy = (const float2*)h->in.put(xy, (size_t)total * 2);
    const float* d_xun = h->in.put(x_un, (size_t)total);
    float *h_d, *h_ur;
    float* d_d = h->in.out<float>((size_t)total, &h_d);
    float* d_ur = h->in.out<float>((size_t)total, &h_ur);
    cudaStream_t st = h->stream;
    if ((rc = h->in.upload(st)) != PL_OK) return rc;
    k_stereo_from_rgbd<<<(total + 255) / 256, 256, 0, st>>>(d_depth, step_bytes, frame_stride_bytes, d_off, n_frames, total, d_xy, d_xun, bf, d_d, d_ur);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(h_d, d_d, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(cudaMemcpyAsync(h_ur, d_ur, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    memcpy(depth_out, h_d, (size_t)total * 4);
    memcpy(u_right_out, h_ur, (size_t)total * 4);
    return PL_OK;
}

PL_API int pl_frame_unproject_batch(pl_match* h, int n_frames, const int* off, const float* xy_un, const float* z, const float* rwc, const float* ow,
                                    float fx, float fy, float cx, float cy, float* world, uint8_t* valid) {
    PL_CHECK_ARG(h && n_frames >= 0 && fx != 0.f && fy != 0.f);
    if (n_frames == 0) return PL_OK;
    int rc = check_offsets(off, n_frames);
    if (rc != PL_OK) return rc;
    const int total = off[n_frames];
    if (total == 0) return PL_OK;
    PL_CHECK_ARG(xy_un && z && rwc && ow && world && valid);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    if ((rc = h->in.reserve(padb((size_t)(n_frames + 1) * 4) + padb((size_t)total * 8) + padb((size_t)total * 4) + padb((size_t)n_frames * 36) +
                            padb((size_t)n_frames * 12) + padb((size_t)total * 12) + padb((size_t)total))) != PL_OK)
        return rc;
    const int* d_off = h->in.put(off, (size_t)n_frames + 1);
    const float2* d_xy = (const float2*)h->in.put(xy_un, (size_t)total * 2);
    const float* d_z = h->in.put(z, (size_t)total);
    const float* d_r = h->in.put(rwc, (size_t)n_frames * 9);
    const float* d_o = h->in.put(ow, (size_t)n_frames * 3);
    float* h_w;
    uint8_t* h_v;
    float* d_w = h->in.out<float>((size_t)total * 3, &h_w);
    uint8_t* d_v = h->in.out<uint8_t>((size_t)total, &h_v);
    cudaStream_t st = h->stream;
    if ((rc = h->in.upload(st)) != PL_OK) return rc;
    volatile float invfx = 1.0f / fx, invfy = 1.0f / fy;  // Frame.cc:180-181
    k_unproject<<<(total + 255) / 256, 256, 0, st>>>(d_off, n_frames, total, d_xy, d_z, d_r, d_o, cx, cy, invfx, invfy, d_w, d_v);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(h_w, d_w, (size_t)total * 12, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(cudaMemcpyAsync(h_v, d_v, (size_t)total, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    memcpy(world, h_w, (size_t)total * 12);
    memcpy(valid, h_v, (size_t)total);
    return PL_OK;
}

PL_API int pl_frame_is_in_frustum_batch(pl_match* h, int n_frames, const float* tcw, const float* ow, float fx, float fy, float cx, float cy, float bf,
                                        const float bounds[4], int n_levels, float log_scale_factor, int m, const float* world_pos,
                                        const float* normal, const float* min_dist_inv, const float* max_dist_inv, const float* max_dist,
                                        float viewing_cos_limit, uint8_t* in_view, float* proj_x, float* proj_y, float* proj_xr, int* scale_level,
                                        float* view_cos) {
    PL_CHECK_ARG(h && n_frames >= 0 && m >= 0 && n_frames <= 65535 && bounds && n_levels >= 1 && log_scale_factor > 0.f);
    if (n_frames == 0 || m == 0) return PL_OK;
    PL_CHECK_ARG(tcw && ow && world_pos && normal && min_dist_inv && max_dist_inv && max_dist && in_view && proj_x && proj_y && proj_xr && scale_level &&
                 view_cos);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    const size_t nm = (size_t)n_frames * m;
    int rc = h->in.reserve(padb((size_t)n_frames * 48) + padb((size_t)n_frames * 12) + padb((size_t)m * 12) * 2 + padb((size_t)m * 4) * 3);
    if (rc != PL_OK) return rc;
    if ((rc = h->res.reserve(padb(nm) + padb(nm * 4) * 5)) != PL_OK) return rc;
    const float* d_t = h->in.put(tcw, (size_t)n_frames * 12);
    const float* d_o = h->in.put(ow, (size_t)n_frames * 3);
    const float* d_p = h->in.put(world_pos, (size_t)m * 3);
    const float* d_n = h->in.put(normal, (size_t)m * 3);
    const float* d_mi = h->in.put(min_dist_inv, (size_t)m);
    const float* d_ma = h->in.put(max_dist_inv, (size_t)m);
    const float* d_mr = h->in.put(max_dist, (size_t)m);
    uint8_t* h_iv;
    float *h_x, *h_y, *h_xr, *h_vc;
    int* h_l;
    uint8_t* d_iv = h->res.out<uint8_t>(nm, &h_iv);
    float* d_x = h->res.out<float>(nm, &h_x);
    float* d_y = h->res.out<float>(nm, &h_y);
    float* d_xr = h->res.out<float>(nm, &h_xr);
    int* d_l = h->res.out<int>(nm, &h_l);
    float* d_vc = h->res.out<float>(nm, &h_vc);
    cudaStream_t st = h->stream;
    if ((rc = h->in.upload(st)) != PL_OK) return rc;
    FrustumArgs A{fx, fy, cx, cy, bf, bounds[0], bounds[1], bounds[2], bounds[3], log_scale_factor, viewing_cos_limit, n_levels, m};
    k_is_in_frustum<<<dim3((m + 255) / 256, n_frames), 256, 0, st>>>(d_t, d_o, A, d_p, d_n, d_mi, d_ma, d_mr, d_iv, d_x, d_y, d_xr, d_l, d_vc);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(h->res.h, h->res.d, h->res.cur, cudaMemcpyDeviceToHost, st));  // the six planes: one copy
    PL_CUDA_TRY(pl::stream_sync(st));
    memcpy(in_view, h_iv, nm);
    memcpy(proj_x, h_x, nm * 4);
    memcpy(proj_y, h_y, nm * 4);
    memcpy(proj_xr, h_xr, nm * 4);
    memcpy(scale_level, h_l, nm * 4);
    memcpy(view_cos, h_vc, nm * 4);
    return PL_OK;
}

}  // extern "C"

namespace pl {
// k_is_in_frustum for callers in other translation units (pl_orb_search_local_map_batch keeps the planes on the device)
int launch_is_in_frustum(cudaStream_t st, int n_frames, const float* d_tcw, const float* d_ow, float fx, float fy, float cx, float cy, float bf,
                         const float bounds[4], float log_sf, float cos_limit, int n_levels, int m, const float* d_pos, const float* d_normal,
                         const float* d_min_inv, const float* d_max_inv, const float* d_max_raw, uint8_t* d_in_view, float* d_x, float* d_y,
                         float* d_xr, int* d_lvl, float* d_vc) {
    if (n_frames <= 0 || m <= 0) return PL_OK;
    FrustumArgs A{fx, fy, cx, cy, bf, bounds[0], bounds[1], bounds[2], bounds[3], log_sf, cos_limit, n_levels, m};
    k_is_in_frustum<<<dim3((m + 255) / 256, n_frames), 256, 0, st>>>(d_tcw, d_ow, A, d_pos, d_normal, d_min_inv, d_max_inv, d_max_raw, d_in_view, d_x,
                                                                      d_y, d_xr, d_lvl, d_vc);
    PL_CUDA_TRY(cudaGetLastError());
    return PL_OK;
}
}  // namespace pl

extern "C" {

PL_API int pl_frame_lines_in_frustum_batch(pl_match* h, int n_frames, const float* tcw, int m, const double* start3d, const double* end3d,
                                           uint8_t* in_view) {
    PL_CHECK_ARG(h && n_frames >= 0 && m >= 0 && n_frames <= 65535);
    if (n_frames == 0 || m == 0) return PL_OK;
    PL_CHECK_ARG(tcw && start3d && end3d && in_view);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    const size_t nm = (size_t)n_frames * m;
    int rc = h->in.reserve(padb((size_t)n_frames * 48) + padb((size_t)m * 24) * 2 + padb(nm));
    if (rc != PL_OK) return rc;
    const float* d_t = h->in.put(tcw, (size_t)n_frames * 12);
    const double* d_s = h->in.put(start3d, (size_t)m * 3);
    const double* d_e = h->in.put(end3d, (size_t)m * 3);
    uint8_t* h_iv;
    uint8_t* d_iv = h->in.out<uint8_t>(nm, &h_iv);
    cudaStream_t st = h->stream;
    if ((rc = h->in.upload(st)) != PL_OK) return rc;
    k_lines_in_frustum<<<dim3((m + 255) / 256, n_frames), 256, 0, st>>>(d_t, m, d_s, d_e, d_iv);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(h_iv, d_iv, nm, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    memcpy(in_view, h_iv, nm);
    return PL_OK;
}

}  // extern "C"
