// Exercises the host-side mirrors (ORBextractor / LineExtractor / matchers) the way Frame.cc uses them and prints
// checksums that tests/test_host_mirror_gpu.py compares with the oracle.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../orb_slam2_modification_with-point-and-line-feature_b200/host/LineExtractor.h"
#include "../../orb_slam2_modification_with-point-and-line-feature_b200/host/Matchers.h"
#include "../../orb_slam2_modification_with-point-and-line-feature_b200/host/ORBextractor.h"

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s raw.gray cols rows\n", argv[0]); return 2; }
    const int cols = atoi(argv[2]), rows = atoi(argv[3]);
    cv::Mat im(rows, cols, cv::CV_8UC1);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(im.data, 1, (size_t)rows * cols, f) != (size_t)rows * cols) { fprintf(stderr, "cannot read image\n"); return 2; }
    fclose(f);
    ORB_SLAM2::ORBextractor* mpORBextractorLeft = new ORB_SLAM2::ORBextractor(1000, 1.2f, 8, 20, 7, cols, rows);
    std::vector<cv::KeyPoint> mvKeys;
    cv::Mat mDescriptors;
    (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors);  // Frame.cc:321
    unsigned long long h = 1469598103934665603ull;
    for (size_t i = 0; i < mvKeys.size(); i++) {
        const unsigned char* p = (const unsigned char*)&mvKeys[i];
        for (size_t k = 0; k < sizeof(cv::KeyPoint); k++) h = (h ^ p[k]) * 1099511628211ull;
    }
    for (int i = 0; i < mDescriptors.rows * 32; i++) h = (h ^ mDescriptors.data[i]) * 1099511628211ull;
    printf("orb n=%zu levels=%d sf=%.6f hash=%llu pyr7=%dx%d px=%d\n", mvKeys.size(), mpORBextractorLeft->GetLevels(),
           mpORBextractorLeft->GetScaleFactors()[7], h, mpORBextractorLeft->mvImagePyramid[7].cols, mpORBextractorLeft->mvImagePyramid[7].rows,
           (int)mpORBextractorLeft->mvImagePyramid[3].ptr(5)[7]);
    ORB_SLAM2::LineExtractor le(cols, rows);
    std::vector<cv::line_descriptor::KeyLine> mvKeyLines;
    cv::Mat mLineDescriptors;
    std::vector<Eigen::Vector3d> mvKeyLineCoefficient;
    le.ExtractLineSegment(im, mvKeyLines, mLineDescriptors, mvKeyLineCoefficient);  // Frame.cc:327
    unsigned long long hl = 1469598103934665603ull;
    for (int i = 0; i < mLineDescriptors.rows * 32; i++) hl = (hl ^ mLineDescriptors.data[i]) * 1099511628211ull;
    printf("lines n=%zu coeffs=%zu class0=%d deschash=%llu\n", mvKeyLines.size(), mvKeyLineCoefficient.size(), mvKeyLines.empty() ? -1 : mvKeyLines[0].class_id, hl);
    cv::Mat empty;
    std::vector<cv::KeyPoint> k2;
    cv::Mat d2;
    (*mpORBextractorLeft)(empty, cv::Mat(), k2, d2);  // empty image: silent return
    printf("empty n=%zu\n", k2.size());
    ORB_SLAM2::ORBmatcher matcher(0.9f, true);
    std::vector<int> dist(mvKeys.size());
    if (!mvKeys.empty()) matcher.DescriptorDistance(mDescriptors.data, mDescriptors.data, (int)mvKeys.size(), dist.data());
    long s = 0;
    for (int d : dist) s += d;
    printf("selfdist=%ld\n", s);
    delete mpORBextractorLeft;
    return 0;
}
