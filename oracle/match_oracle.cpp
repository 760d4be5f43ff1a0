// match_oracle.cpp — CPU ORACLE for the descriptor searches.  TEST INFRASTRUCTURE ONLY (see orb_oracle.cpp header).
//
// Restates, sequentially and literally:
//   ORBmatcher::DescriptorDistance (reference src/ORBmatcher.cc:2083-2103) == LineMatcher::DescriptorDistance
//     (src/LineMatcher.cpp:20-39): SWAR popcount over eight 32-bit words;
//   cv::BFMatcher(NORM_HAMMING, crossCheck=false).knnMatch(q, t, 2) as called at src/LineMatcher.cpp:496-503 and
//     :1179-1185 — pinned to python cv2 4.13 (ties resolve to the lowest train index) by tests/test_oracle_match.py
//     and tests/golden/knn_*.npz.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "oracle.h"

static inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    const int32_t* pa = reinterpret_cast<const int32_t*>(a);
    const int32_t* pb = reinterpret_cast<const int32_t*>(b);
    int dist = 0;
    for (int i = 0; i < 8; i++, pa++, pb++) {
        unsigned int v = *pa ^ *pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

extern "C" int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

extern "C" void orc_hamming_pairs(const uint8_t* a, const uint8_t* b, int n, int* dist) {
    for (int i = 0; i < n; i++) dist[i] = descriptor_distance(a + 32 * (size_t)i, b + 32 * (size_t)i);
}

static void knn2_range(const uint8_t* q, int q0, int q1, const uint8_t* t, int nt, int* idx, int* dist) {
    for (int i = q0; i < q1; i++) {
        int b1 = -1, d1 = 1 << 30, b2 = -1, d2 = 1 << 30;
        for (int j = 0; j < nt; j++) {
            int d = descriptor_distance(q + 32 * (size_t)i, t + 32 * (size_t)j);
            if (d < d1) { b2 = b1; d2 = d1; b1 = j; d1 = d; }
            else if (d < d2) { b2 = j; d2 = d; }
        }
        idx[2 * i] = b1; dist[2 * i] = b1 < 0 ? -1 : d1;
        idx[2 * i + 1] = b2; dist[2 * i + 1] = b2 < 0 ? -1 : d2;
    }
}

// threads <= 1: scalar single thread.  The reference's BFMatcher call is single-threaded per call site; the
// multi-thread mode only exists for the "all host cores" CPU baseline of bench.py.
extern "C" void orc_hamming_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx, int* dist, int threads) {
    if (threads <= 1) { knn2_range(q, 0, nq, t, nt, idx, dist); return; }
    std::vector<std::thread> th;
    int per = (nq + threads - 1) / threads;
    for (int k = 0; k < threads; k++) {
        int a = k * per, b = std::min(nq, a + per);
        if (a >= b) break;
        th.emplace_back(knn2_range, q, a, b, t, nt, idx, dist);
    }
    for (auto& x : th) x.join();
}

extern "C" void orc_hamming_candidates(const uint8_t* q, int nq, const uint8_t* t, const int* off, const int* cidx, int* dist) {
    for (int i = 0; i < nq; i++)
        for (int k = off[i]; k < off[i + 1]; k++) dist[k] = descriptor_distance(q + 32 * (size_t)i, t + 32 * (size_t)cidx[k]);
}
