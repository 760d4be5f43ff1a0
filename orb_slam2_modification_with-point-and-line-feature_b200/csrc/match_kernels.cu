// match_kernels.cu — descriptor search on sm_100a: the CUDA path behind pl_hamming_* / pl_orb_search_* /
// pl_line_match_* (include/plslam_c.h).
//
// Reference functions replaced:
//   ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2083-2103) == LineMatcher::DescriptorDistance
//     (src/LineMatcher.cpp:20-39)                                          -> hamming256()
//   cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) (LineMatcher.cpp:496-503,1179-1185)  -> k_knn2_partial + k_knn2_merge
//   candidate loops of SearchByProjection / SearchByBoW (ORBmatcher.cc:123-164, 306-336, 1802-1829)
//                                                                          -> k_hamming_candidates
#include <algorithm>
#include <vector>

#include "match_common.cuh"

namespace pl {

__global__ void __launch_bounds__(256) k_hamming_pairs(const uint4* __restrict__ a, const uint4* __restrict__ b, int n,
                                                       int* __restrict__ dist) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dist[i] = hamming256(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}

// ---------------------------------------------------------------------------------------------------------------
// brute-force kNN-2.  Packed key = dist << 20 | train index: the two smallest keys are the two nearest
// neighbours with ties resolved to the lowest train index (cv::BFMatcher behaviour).
// grid = (ceil(nq/128), n_chunks): every CTA scans one chunk of the train set through shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kKnnThreads = 128;
constexpr uint32_t kKeyNone = 0xFFFFFFFFu;

__global__ void __launch_bounds__(kKnnThreads) k_knn2_partial(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t,
                                                              int nt, int chunk, uint32_t* __restrict__ part, int n_chunks) {
    __shared__ uint4 s_t[kKnnThreads * 2];
    const int tid = threadIdx.x;
    const int qi = blockIdx.x * kKnnThreads + tid;
    const int c0 = blockIdx.y * chunk, c1 = min(nt, c0 + chunk);
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    if (qi < nq) { q0 = q[2 * qi]; q1 = q[2 * qi + 1]; }
    uint32_t b1 = kKeyNone, b2 = kKeyNone;
    for (int base = c0; base < c1; base += kKnnThreads) {
        const int n = min(kKnnThreads, c1 - base);
        __syncthreads();
        if (tid < n) {
            s_t[2 * tid] = t[2 * (size_t)(base + tid)];
            s_t[2 * tid + 1] = t[2 * (size_t)(base + tid) + 1];
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < n; j++) {
            const uint32_t key = ((uint32_t)hamming256(q0, q1, s_t[2 * j], s_t[2 * j + 1]) << 20) | (uint32_t)(base + j);
            const uint32_t lo = min(key, b1);
            b2 = min(b2, max(key, b1));
            b1 = lo;
        }
    }
    if (qi < nq) {
        part[((size_t)qi * n_chunks + blockIdx.y) * 2] = b1;
        part[((size_t)qi * n_chunks + blockIdx.y) * 2 + 1] = b2;
    }
}

__global__ void __launch_bounds__(256) k_knn2_merge(const uint32_t* __restrict__ part, int nq, int n_chunks, int* __restrict__ idx,
                                                    int* __restrict__ dist) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    uint32_t b1 = kKeyNone, b2 = kKeyNone;
    for (int c = 0; c < 2 * n_chunks; c++) {
        const uint32_t key = part[(size_t)qi * n_chunks * 2 + c];
        const uint32_t lo = min(key, b1);
        b2 = min(b2, max(key, b1));
        b1 = lo;
    }
    idx[2 * qi] = b1 == kKeyNone ? -1 : (int)(b1 & 0xFFFFFu);
    dist[2 * qi] = b1 == kKeyNone ? -1 : (int)(b1 >> 20);
    idx[2 * qi + 1] = b2 == kKeyNone ? -1 : (int)(b2 & 0xFFFFFu);
    dist[2 * qi + 1] = b2 == kKeyNone ? -1 : (int)(b2 >> 20);
}

// one warp per query row: distances to its CSR candidate list, in list order
__global__ void __launch_bounds__(256) k_hamming_candidates(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t,
                                                            const int* __restrict__ off, const int* __restrict__ cidx,
                                                            int* __restrict__ dist) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nq) return;
    const uint4 q0 = q[2 * w], q1 = q[2 * w + 1];
    const int b = off[w], e = off[w + 1];
    for (int k = b + lane; k < e; k += 32) {
        const int j = cidx[k];
        dist[k] = hamming256(q0, q1, t[2 * (size_t)j], t[2 * (size_t)j + 1]);
    }
}


}  // namespace pl

using namespace pl;

namespace {
inline int scratch(pl_match* h, int slot, size_t bytes, void** out) { return pl::match_scratch(h, slot, bytes, out); }

int knn2_launch(pl_match* h, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int* d_idx, int* d_dist) {
    // enough CTAs for >= 2 waves when the problem allows it; chunks are multiples of the 128-row smem tile
    const int qblocks = (nq + kKnnThreads - 1) / kKnnThreads;
    int want_chunks = std::max(1, (2 * h->sm_count + qblocks - 1) / qblocks);
    int max_chunks = std::max(1, (nt + kKnnThreads - 1) / kKnnThreads);
    int n_chunks = std::min(want_chunks, max_chunks);
    int chunk = ((nt + n_chunks - 1) / n_chunks + kKnnThreads - 1) / kKnnThreads * kKnnThreads;
    if (chunk < kKnnThreads) chunk = kKnnThreads;
    n_chunks = std::max(1, (nt + chunk - 1) / chunk);
    void* part = nullptr;
    int rc = scratch(h, 0, (size_t)nq * n_chunks * 2 * sizeof(uint32_t), &part);
    if (rc != PL_OK) return rc;
    k_knn2_partial<<<dim3(qblocks, n_chunks), kKnnThreads, 0, h->stream>>>((const uint4*)d_q, nq, (const uint4*)d_t, nt, chunk,
                                                                            (uint32_t*)part, n_chunks);
    k_knn2_merge<<<(nq + 255) / 256, 256, 0, h->stream>>>((const uint32_t*)part, nq, n_chunks, d_idx, d_dist);
    h->last_launches += 2;
    PL_CUDA_TRY(cudaGetLastError());
    return PL_OK;
}
}  // namespace

// used by bow_kernels.cu (the D6 line matchers)
int pl_knn2_launch_dev(pl_match* h, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int* d_idx, int* d_dist) {
    return knn2_launch(h, d_q, nq, d_t, nt, d_idx, d_dist);
}

extern "C" {

PL_API int pl_match_create(pl_match** out, int device) {
    PL_CHECK_ARG(out != nullptr);
    *out = nullptr;
    int ndev = 0;
    PL_CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) {
        set_error("device %d not available (%d CUDA devices); this library has no CPU fallback", device, ndev);
        return PL_ERR_CUDA;
    }
    PL_CUDA_TRY(cudaSetDevice(device));
    pl_match* h = new pl_match();
    h->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_error("pl_match_create: %s", cudaGetErrorString(e));
        delete h;
        return PL_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = h;
    return PL_OK;
}

PL_API void pl_match_destroy(pl_match* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) pl::stream_sync(h->stream);
    for (int i = 0; i < 24; i++)
        if (h->d_buf[i]) cudaFree(h->d_buf[i]);
    h->in.release();
    h->res.release();
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

PL_API int pl_match_sync(pl_match* h) {
    PL_CHECK_ARG(h);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    return PL_OK;
}
PL_API void* pl_match_stream(pl_match* h) { return h ? (void*)h->stream : nullptr; }
PL_API int pl_match_last_launches(const pl_match* h) { return h ? h->last_launches : 0; }

PL_API int pl_hamming_pairs(pl_match* h, const uint8_t* a, const uint8_t* b, int n, int* dist) {
    PL_CHECK_ARG(h && a && b && dist && n >= 0);
    if (n == 0) return PL_OK;
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    void *da, *db, *dd;
    int rc;
    if ((rc = scratch(h, 1, (size_t)n * 32, &da)) != PL_OK) return rc;
    if ((rc = scratch(h, 2, (size_t)n * 32, &db)) != PL_OK) return rc;
    if ((rc = scratch(h, 3, (size_t)n * 4, &dd)) != PL_OK) return rc;
    PL_CUDA_TRY(cudaMemcpyAsync(da, a, (size_t)n * 32, cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(db, b, (size_t)n * 32, cudaMemcpyHostToDevice, h->stream));
    k_hamming_pairs<<<(n + 255) / 256, 256, 0, h->stream>>>((const uint4*)da, (const uint4*)db, n, (int*)dd);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(dist, dd, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    return PL_OK;
}

PL_API int pl_hamming_knn2_dev(pl_match* h, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int* d_idx, int* d_dist) {
    PL_CHECK_ARG(h && d_q && d_idx && d_dist && nq >= 0 && nt >= 0 && nt < (1 << 20));
    PL_CHECK_ARG(((uintptr_t)d_q & 15) == 0 && ((uintptr_t)d_t & 15) == 0);
    if (nq == 0) return PL_OK;
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    return knn2_launch(h, d_q, nq, d_t, nt, d_idx, d_dist);
}

PL_API int pl_hamming_knn2(pl_match* h, const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx, int* dist) {
    PL_CHECK_ARG(h && q && idx && dist && nq >= 0 && nt >= 0 && nt < (1 << 20) && (t || nt == 0));
    if (nq == 0) return PL_OK;
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    void *dq, *dt, *di, *dd;
    int rc;
    if ((rc = scratch(h, 1, (size_t)nq * 32, &dq)) != PL_OK) return rc;
    if ((rc = scratch(h, 2, (size_t)std::max(nt, 1) * 32, &dt)) != PL_OK) return rc;
    if ((rc = scratch(h, 3, (size_t)nq * 8, &di)) != PL_OK) return rc;
    if ((rc = scratch(h, 4, (size_t)nq * 8, &dd)) != PL_OK) return rc;
    PL_CUDA_TRY(cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, h->stream));
    if (nt) PL_CUDA_TRY(cudaMemcpyAsync(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice, h->stream));
    rc = knn2_launch(h, (const uint8_t*)dq, nq, (const uint8_t*)dt, nt, (int*)di, (int*)dd);
    if (rc != PL_OK) return rc;
    PL_CUDA_TRY(cudaMemcpyAsync(idx, di, (size_t)nq * 8, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(dist, dd, (size_t)nq * 8, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    return PL_OK;
}

PL_API int pl_hamming_candidates(pl_match* h, const uint8_t* q, int nq, const uint8_t* t, int nt, const int* cand_off,
                                 const int* cand_idx, int* dist_out) {
    PL_CHECK_ARG(h && q && t && cand_off && nq >= 0 && nt > 0);
    if (nq == 0) return PL_OK;
    PL_CHECK_ARG(cand_off[0] == 0);
    for (int i = 0; i < nq; i++) PL_CHECK_ARG(cand_off[i + 1] >= cand_off[i]);  // a CSR: offsets start at 0 and never decrease
    const int total = cand_off[nq];
    PL_CHECK_ARG(total >= 0 && (total == 0 || (cand_idx && dist_out)));
    if (total == 0) return PL_OK;
    for (int i = 0; i < total; i++) PL_CHECK_ARG(cand_idx[i] >= 0 && cand_idx[i] < nt);
    PL_CUDA_TRY(cudaSetDevice(h->device));
    h->last_launches = 0;
    void *dq, *dt, *doff, *dci, *dd;
    int rc;
    if ((rc = scratch(h, 1, (size_t)nq * 32, &dq)) != PL_OK) return rc;
    if ((rc = scratch(h, 2, (size_t)nt * 32, &dt)) != PL_OK) return rc;
    if ((rc = scratch(h, 3, (size_t)(nq + 1) * 4, &doff)) != PL_OK) return rc;
    if ((rc = scratch(h, 4, (size_t)total * 4, &dci)) != PL_OK) return rc;
    if ((rc = scratch(h, 5, (size_t)total * 4, &dd)) != PL_OK) return rc;
    PL_CUDA_TRY(cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(doff, cand_off, (size_t)(nq + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    PL_CUDA_TRY(cudaMemcpyAsync(dci, cand_idx, (size_t)total * 4, cudaMemcpyHostToDevice, h->stream));
    k_hamming_candidates<<<(nq * 32 + 255) / 256, 256, 0, h->stream>>>((const uint4*)dq, nq, (const uint4*)dt, (const int*)doff,
                                                                       (const int*)dci, (int*)dd);
    h->last_launches++;
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(dist_out, dd, (size_t)total * 4, cudaMemcpyDeviceToHost, h->stream));
    PL_CUDA_TRY(pl::stream_sync(h->stream));
    return PL_OK;
}

}  // extern "C"
