// match_common.cuh — pieces shared by match_kernels.cu (Hamming searches) and search_kernels.cu (projection searches,
// line matching): the 256-bit Hamming distance, the pl_match handle and its scratch / staging buffers.
#pragma once
#include <algorithm>
#include <thread>
#include <unordered_map>
#include <vector>

#include "pl_common.cuh"

namespace pl {
// 256-bit Hamming distance of two 32-byte rows held as 2 x uint4
// (ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:2083-2103 == LineMatcher::DescriptorDistance, src/LineMatcher.cpp:20-39)
__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
           __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

#ifdef __CUDACC__
// cv::undistortPoints(src, dst, K, D = (k1 k2 p1 p2 k3), noArray(), P = K) for one point: five fixed-point iterations of the inverse
// Brown model in double (TermCriteria(MAX_ITER, 5, 0.01)), re-projection with P.  Pinned bit-exactly against cv2 4.13.
__device__ __forceinline__ float2 undistort_point(float px, float py, double fx, double fy, double cx, double cy, double k0, double k1, double k2,
                                                  double k3, double k4) {
    const double ifx = 1. / fx, ify = 1. / fy;
    const double u = px, v = py;
    double x = (u - cx) * ifx, y = (v - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((0. * r2 + 0.) * r2 + 0.) * r2) / (1 + ((k4 * r2 + k1) * r2 + k0) * r2);
        if (icdist < 0) {
            x = (u - cx) * ifx;
            y = (v - cy) * ify;
            break;
        }
        const double deltaX = 2 * k2 * x * y + k3 * (r2 + 2 * x * x) + 0. * r2 + 0. * r2 * r2;
        const double deltaY = k2 * (r2 + 2 * y * y) + 2 * k3 * x * y + 0. * r2 + 0. * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    const double xx = fx * x + 0. * y + cx, yy = 0. * x + fy * y + cy, ww = 1. / (0. * x + 0. * y + 1.);
    return make_float2((float)(xx * ww), (float)(yy * ww));
}
#endif
}  // namespace pl

// one pinned host buffer + one device buffer: every input of a call is packed into it and uploaded with ONE copy
struct PlStage {
    uint8_t* h = nullptr;
    uint8_t* d = nullptr;
    size_t cap = 0, cur = 0;
    // An API call that returns early with an error after upload() leaves the copy out of the pinned buffer (and the kernels reading
    // the device side) in flight: the next call must not refill or free the buffers under them.
    cudaStream_t busy_stream = nullptr;
    bool in_flight = false;
    void quiesce() {
        if (in_flight) {
            (void)cudaStreamSynchronize(busy_stream);
            in_flight = false;
        }
    }
    int reserve(size_t bytes) {
        quiesce();
        cur = 0;
        overflow = false;
        jobs.clear();
        seen.clear();
        if (bytes <= cap) return PL_OK;
        if (h) cudaFreeHost(h);
        if (d) cudaFree(d);
        h = d = nullptr;
        cap = 0;
        size_t want = std::max(bytes + bytes / 4, (size_t)1 << 20);
        PL_CUDA_TRY(cudaMallocHost((void**)&h, want));
        PL_CUDA_TRY(cudaMalloc((void**)&d, want));
        cap = want;
        return PL_OK;
    }
    static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
    // copies n elements into the pinned buffer; returns the DEVICE address they will have after upload()
    // Large arrays are not copied at once: the copy is recorded and executed by a few host threads in upload(), so the
    // source must stay alive until then (it always is: every put() happens inside the API call that uploads).
    bool overflow = false;  // a put() / out() did not fit what reserve() was asked for (checked by upload())
    struct Job { size_t off; const void* src; size_t bytes; };
    std::vector<Job> jobs;
    // a large array that several instances of a batch share (the same local-map snapshot for consecutive frames, a frame that
    // is "current" in one instance and "last" in the next) travels once: same host pointer and size -> same device copy
    std::unordered_map<const void*, std::pair<size_t, size_t>> seen;  // src -> (bytes, offset)
    template <typename T>
    const T* put(const T* src, size_t n) {
        const size_t off = cur, bytes = n * sizeof(T);
        if (off + pad(bytes) > cap) {  // the caller's reserve() sum and its put() sequence disagree: never write past the buffers
            overflow = true;
            return (const T*)d;
        }
        if (n && src) {
            if (bytes >= (size_t)8 << 10) {
                auto it = seen.find((const void*)src);
                if (it != seen.end() && it->second.first == bytes) return (const T*)(d + it->second.second);
                seen[(const void*)src] = std::make_pair(bytes, off);
                jobs.push_back(Job{off, src, bytes});
            } else {
                memcpy(h + off, src, bytes);
            }
        }
        cur += pad(bytes);
        return (const T*)(d + off);
    }
    // executes the recorded copies (up to 4 threads when there is enough to copy) and sends the packed buffer to the device
    int upload(cudaStream_t st) {
        if (overflow) {
            pl::set_error("internal: a packed call needed more staging memory than it reserved");
            return PL_ERR_CAPACITY;
        }
        size_t total = 0;
        for (const Job& j : jobs) total += j.bytes;
        const int nt = total >= ((size_t)4 << 20) ? 4 : 1;
        auto work = [&](int t) {
            // thread t takes the jobs whose cumulative start falls into its share of the bytes
            const size_t lo = total * t / nt, hi = total * (t + 1) / nt;
            size_t acc = 0;
            for (const Job& j : jobs) {
                if (acc >= lo && acc < hi) memcpy(h + j.off, j.src, j.bytes);
                acc += j.bytes;
            }
        };
        if (nt == 1) {
            work(0);
        } else {
            std::thread th[3];
            for (int t = 1; t < nt; t++) th[t - 1] = std::thread(work, t);
            work(0);
            for (int t = 1; t < nt; t++) th[t - 1].join();
        }
        jobs.clear();
        busy_stream = st;
        in_flight = true;
        PL_CUDA_TRY(cudaMemcpyAsync(d, h, cur, cudaMemcpyHostToDevice, st));
        return PL_OK;
    }
    // space that only exists on the device side (outputs / scratch): returns device address, and host mirror address
    template <typename T>
    T* out(size_t n, T** host_mirror = nullptr) {
        const size_t off = cur;
        if (off + pad(n * sizeof(T)) > cap) {
            overflow = true;
            if (host_mirror) *host_mirror = (T*)h;
            return (T*)d;
        }
        cur += pad(n * sizeof(T));
        if (host_mirror) *host_mirror = (T*)(h + off);
        return (T*)(d + off);
    }
    void release() {
        quiesce();
        if (h) cudaFreeHost(h);
        if (d) cudaFree(d);
        h = d = nullptr;
        cap = cur = 0;
    }
};

struct pl_match {
    int device = 0;
    cudaStream_t stream = nullptr;
    int last_launches = 0;
    // growable device scratch
    uint8_t* d_buf[24] = {nullptr};
    size_t d_cap[24] = {0};
    int sm_count = 148;
    PlStage in, res;   // packed inputs / packed results of the batched searches
};

namespace pl {
// frame_kernels.cu: Frame::IsInFrustum of n_frames frames against one snapshot of m map points, device pointers throughout
int launch_is_in_frustum(cudaStream_t st, int n_frames, const float* d_tcw, const float* d_ow, float fx, float fy, float cx, float cy, float bf,
                         const float bounds[4], float log_sf, float cos_limit, int n_levels, int m, const float* d_pos, const float* d_normal,
                         const float* d_min_inv, const float* d_max_inv, const float* d_max_raw, uint8_t* d_in_view, float* d_x, float* d_y,
                         float* d_xr, int* d_lvl, float* d_vc);
inline int match_scratch(pl_match* h, int slot, size_t bytes, void** out) {
    if (h->d_cap[slot] < bytes) {
        if (h->d_buf[slot]) cudaFree(h->d_buf[slot]);
        h->d_buf[slot] = nullptr;
        h->d_cap[slot] = 0;
        size_t want = std::max(bytes + bytes / 4, (size_t)1 << 16);
        PL_CUDA_TRY(cudaMalloc((void**)&h->d_buf[slot], want));
        h->d_cap[slot] = want;
    }
    *out = h->d_buf[slot];
    return PL_OK;
}
}  // namespace pl
