// line_oracle.cpp — CPU ORACLE for the line-extraction path.  TEST INFRASTRUCTURE ONLY (see orb_oracle.cpp header).
//
// Restates LineExtractor::ExtractLineSegment (reference src/LineExtractor.cpp:12-70) and the un-vendored OpenCV code
// it calls:
//   * cv::line_descriptor::LSDDetector::detect (opencv_contrib line_descriptor/src/LSDDetector.cpp) — octave 0 is the
//     image itself; cv::createLineSegmentDetector(LSD_REFINE_ADV) (opencv imgproc/src/lsd.cpp) does the work.
//     PINNED: tests/test_oracle_line.py compares orc_lsd_detect with cv2.createLineSegmentDetector(LSD_REFINE_ADV)
//     of python cv2 4.13 (segments, width, prec, nfa) and with tests/golden/lsd_*.npz.
//   * cv::line_descriptor::BinaryDescriptor::compute (line_descriptor/src/binary_descriptor.cpp: computeSobel,
//     computeLBD, binaryConversion).  PARITY UNPINNED: opencv_contrib is not available in this image (cv2 has no
//     line_descriptor module) and the reference ships no golden vector, so this is a restatement of the published
//     algorithm (Zhang & Koch 2013) with OpenCV's parameters; its primitives (GaussianBlur 5x5 sigma 1, Sobel 3x3
//     CV_16S) are pinned to cv2.  The bit order inside a descriptor byte (1<<i) is the one detail that cannot be
//     checked; it does not affect any Hamming distance.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

const double kPi = 3.14159265358979323846;
const double M_3_2_PI_ = (3 * kPi) / 2;
const double M_2__PI_ = 2 * kPi;
const double NOTDEF = -1024.0;
const double DEG_TO_RADS = kPi / 180;
const double RELATIVE_ERROR_FACTOR = 100.0;

inline int cvRoundF(float v) { return (int)lrintf(v); }
inline int cvRoundD(double v) { return (int)lrint(v); }

inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// cv::resize(src, dst, Size(), fx, fy, INTER_LINEAR_EXACT) for CV_8UC1 (resize.cpp: resize_bitExact, ufixedpoint16)
void linear_exact_axis(int dn, int sn, double inv_scale, std::vector<int>& ofs, std::vector<int>& c0, std::vector<int>& c1) {
    ofs.assign(dn, 0); c0.assign(dn, 256); c1.assign(dn, 0);
    const double scale = 1.0 / inv_scale;
    for (int v = 0; v < dn; v++) {
        double f = scale * (v + 0.5) - 0.5;
        int i = (int)std::floor(f);
        if (i >= 0 && sn > 1) {
            if (i < sn - 1) {
                ofs[v] = i;
                c1[v] = cvRoundD((f - i) * 256);
                c0[v] = 256 - c1[v];
            } else {
                ofs[v] = sn - 1;
            }
        }
    }
}
void resize_linear_exact(const std::vector<uint8_t>& src, int sw, int sh, double fx, double fy, std::vector<uint8_t>& dst, int& dw, int& dh) {
    dw = cvRoundD(sw * fx);
    dh = cvRoundD(sh * fy);
    std::vector<int> xo, xc0, xc1, yo, yc0, yc1;
    linear_exact_axis(dw, sw, fx, xo, xc0, xc1);
    linear_exact_axis(dh, sh, fy, yo, yc0, yc1);
    dst.assign((size_t)dw * dh, 0);
    std::vector<uint32_t> r0(dw), r1(dw);
    for (int y = 0; y < dh; y++) {
        const uint8_t* s0 = &src[(size_t)yo[y] * sw];
        const uint8_t* s1 = &src[(size_t)std::min(yo[y] + 1, sh - 1) * sw];
        for (int x = 0; x < dw; x++) {
            int a = xo[x], b = std::min(xo[x] + 1, sw - 1);
            r0[x] = s0[a] * xc0[x] + s0[b] * xc1[x];
            r1[x] = s1[a] * xc0[x] + s1[b] * xc1[x];
        }
        for (int x = 0; x < dw; x++) dst[(size_t)y * dw + x] = (uint8_t)((r0[x] * yc0[y] + r1[x] * yc1[y] + 32768u) >> 16);
    }
}

struct RegionPoint { int x, y; double angle, modgrad; };
struct Rect {
    double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p;
};
struct NormPoint { int x, y, norm; };

struct Lsd {
    // parameters of createLineSegmentDetector(LSD_REFINE_ADV) defaults
    const double SCALE = 0.8, SIGMA_SCALE = 0.6, QUANT = 2.0, ANG_TH = 22.5, LOG_EPS = 0, DENSITY_TH = 0.7;
    const int N_BINS = 1024;
    int W = 0, H = 0;
    double LOG_NT = 0;
    std::vector<uint8_t> scaled;
    std::vector<double> angles, modgrad;
    std::vector<uint8_t> used;
    std::vector<NormPoint> ordered;
    // 0: stable order inside a gradient bin (raster order) — what cv2 4.13 does (verified: exact equality of all
    //    segments on every test frame); 1: std::sort (unstable) — kept only to document that it does NOT match
    int order_mode = 0;

    inline double ang(int x, int y) const { return angles[(size_t)y * W + x]; }

    static double angle_diff_signed(double a, double b) {
        double diff = a - b;
        while (diff <= -kPi) diff += M_2__PI_;
        while (diff > kPi) diff -= M_2__PI_;
        return diff;
    }
    static double angle_diff(double a, double b) { return std::fabs(angle_diff_signed(a, b)); }
    static bool double_equal(double a, double b) {
        if (a == b) return true;
        double abs_diff = std::fabs(a - b), aa = std::fabs(a), bb = std::fabs(b);
        double abs_max = aa > bb ? aa : bb;
        if (abs_max < DBL_MIN) abs_max = DBL_MIN;
        return (abs_diff / abs_max) <= (RELATIVE_ERROR_FACTOR * DBL_EPSILON);
    }
    static double distSq(double x1, double y1, double x2, double y2) { return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1); }
    static double dist(double x1, double y1, double x2, double y2) { return std::sqrt(distSq(x1, y1, x2, y2)); }
    static double log_gamma_windschitl(double x) {
        return 0.918938533204673 + (x - 0.5) * std::log(x) - x + 0.5 * x * std::log(x * std::sinh(1 / x) + 1 / (810.0 * std::pow(x, 6.0)));
    }
    static double log_gamma_lanczos(double x) {
        static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
        double a = (x + 0.5) * std::log(x + 5.5) - (x + 5.5);
        double b = 0;
        for (int n = 0; n < 7; ++n) {
            a -= std::log(x + double(n));
            b += q[n] * std::pow(x, double(n));
        }
        return a + std::log(b);
    }
    static double log_gamma(double x) { return x > 15.0 ? log_gamma_windschitl(x) : log_gamma_lanczos(x); }

    void ll_angle(double threshold) {
        angles.assign((size_t)W * H, NOTDEF);
        modgrad.assign((size_t)W * H, 0.0);
        double max_grad = -1;
        for (int y = 0; y < H - 1; ++y) {
            const uint8_t* r0 = &scaled[(size_t)y * W];
            const uint8_t* r1 = &scaled[(size_t)(y + 1) * W];
            for (int x = 0; x < W - 1; ++x) {
                int DA = r1[x + 1] - r0[x];
                int BC = r0[x + 1] - r1[x];
                int gx = DA + BC, gy = DA - BC;
                double norm = std::sqrt((gx * gx + gy * gy) / 4.0);
                modgrad[(size_t)y * W + x] = norm;
                if (norm <= threshold) {
                    angles[(size_t)y * W + x] = NOTDEF;
                } else {
                    angles[(size_t)y * W + x] = orc_fast_atan2(float(gx), float(-gy)) * DEG_TO_RADS;
                    if (norm > max_grad) max_grad = norm;
                }
            }
        }
        double bin_coef = (max_grad > 0) ? double(N_BINS - 1) / max_grad : 0;
        ordered.clear();
        ordered.reserve((size_t)(W - 1) * (H - 1));
        for (int y = 0; y < H - 1; ++y)
            for (int x = 0; x < W - 1; ++x) {
                NormPoint p;
                p.x = x; p.y = y;
                p.norm = int(modgrad[(size_t)y * W + x] * bin_coef);
                ordered.push_back(p);
            }
        if (order_mode == 0)
            std::stable_sort(ordered.begin(), ordered.end(), [](const NormPoint& a, const NormPoint& b) { return a.norm > b.norm; });
        else
            std::sort(ordered.begin(), ordered.end(), [](const NormPoint& a, const NormPoint& b) { return a.norm > b.norm; });
    }

    bool isAligned(int x, int y, double theta, double prec) const {
        if (x < 0 || y < 0 || x >= W || y >= H) return false;
        const double a = ang(x, y);
        if (a == NOTDEF) return false;
        double n_theta = theta - a;
        if (n_theta < 0) n_theta = -n_theta;
        if (n_theta > M_3_2_PI_) {
            n_theta -= M_2__PI_;
            if (n_theta < 0) n_theta = -n_theta;
        }
        return n_theta <= prec;
    }

    void region_grow(int sx, int sy, std::vector<RegionPoint>& reg, double& reg_angle, double prec) {
        reg.clear();
        RegionPoint seed;
        seed.x = sx; seed.y = sy;
        reg_angle = ang(sx, sy);
        seed.angle = reg_angle;
        seed.modgrad = modgrad[(size_t)sy * W + sx];
        reg.push_back(seed);
        float sumdx = float(std::cos(reg_angle));
        float sumdy = float(std::sin(reg_angle));
        used[(size_t)sy * W + sx] = 1;
        for (size_t i = 0; i < reg.size(); i++) {
            const RegionPoint rp = reg[i];
            int xx_min = std::max(rp.x - 1, 0), xx_max = std::min(rp.x + 1, W - 1);
            int yy_min = std::max(rp.y - 1, 0), yy_max = std::min(rp.y + 1, H - 1);
            for (int yy = yy_min; yy <= yy_max; ++yy)
                for (int xx = xx_min; xx <= xx_max; ++xx) {
                    uint8_t& is_used = used[(size_t)yy * W + xx];
                    if (is_used != 1 && isAligned(xx, yy, reg_angle, prec)) {
                        const double angle = ang(xx, yy);
                        is_used = 1;
                        RegionPoint np;
                        np.x = xx; np.y = yy;
                        np.modgrad = modgrad[(size_t)yy * W + xx];
                        np.angle = angle;
                        reg.push_back(np);
                        sumdx += cosf(float(angle));
                        sumdy += sinf(float(angle));
                        reg_angle = orc_fast_atan2(sumdy, sumdx) * DEG_TO_RADS;
                    }
                }
        }
    }

    double get_theta(const std::vector<RegionPoint>& reg, double x, double y, double reg_angle, double prec) const {
        double Ixx = 0.0, Iyy = 0.0, Ixy = 0.0;
        for (size_t i = 0; i < reg.size(); ++i) {
            const double regx = reg[i].x, regy = reg[i].y, weight = reg[i].modgrad;
            double dx = regx - x, dy = regy - y;
            Ixx += dy * dy * weight;
            Iyy += dx * dx * weight;
            Ixy -= dx * dy * weight;
        }
        double lambda = 0.5 * (Ixx + Iyy - std::sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
        double theta = (std::fabs(Ixx) > std::fabs(Iyy)) ? double(orc_fast_atan2(float(lambda - Ixx), float(Ixy)))
                                                         : double(orc_fast_atan2(float(Ixy), float(lambda - Iyy)));
        theta *= DEG_TO_RADS;
        if (angle_diff(theta, reg_angle) > prec) theta += kPi;
        return theta;
    }

    void region2rect(const std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec) const {
        double x = 0, y = 0, sum = 0;
        for (size_t i = 0; i < reg.size(); ++i) {
            const double weight = reg[i].modgrad;
            x += double(reg[i].x) * weight;
            y += double(reg[i].y) * weight;
            sum += weight;
        }
        x /= sum;
        y /= sum;
        double theta = get_theta(reg, x, y, reg_angle, prec);
        double dx = std::cos(theta), dy = std::sin(theta);
        double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
        for (size_t i = 0; i < reg.size(); ++i) {
            double regdx = double(reg[i].x) - x, regdy = double(reg[i].y) - y;
            double l = regdx * dx + regdy * dy;
            double w = -regdx * dy + regdy * dx;
            if (l > l_max) l_max = l;
            else if (l < l_min) l_min = l;
            if (w > w_max) w_max = w;
            else if (w < w_min) w_min = w;
        }
        rec.x1 = x + l_min * dx; rec.y1 = y + l_min * dy;
        rec.x2 = x + l_max * dx; rec.y2 = y + l_max * dy;
        rec.width = w_max - w_min;
        rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy; rec.prec = prec; rec.p = p;
        if (rec.width < 1.0) rec.width = 1.0;
    }

    bool reduce_region_radius(std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec, double density,
                              double density_th) {
        double xc = double(reg[0].x), yc = double(reg[0].y);
        double radSq1 = distSq(xc, yc, rec.x1, rec.y1), radSq2 = distSq(xc, yc, rec.x2, rec.y2);
        double radSq = radSq1 > radSq2 ? radSq1 : radSq2;
        while (density < density_th) {
            radSq *= 0.75 * 0.75;
            for (size_t i = 0; i < reg.size(); ++i) {
                if (distSq(xc, yc, double(reg[i].x), double(reg[i].y)) > radSq) {
                    used[(size_t)reg[i].y * W + reg[i].x] = 0;
                    std::swap(reg[i], reg[reg.size() - 1]);
                    reg.pop_back();
                    --i;
                }
            }
            if (reg.size() < 2) return false;
            region2rect(reg, reg_angle, prec, p, rec);
            density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        }
        return true;
    }

    bool refine(std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec, double density_th) {
        double density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density >= density_th) return true;
        double xc = double(reg[0].x), yc = double(reg[0].y);
        const double ang_c = reg[0].angle;
        double sum = 0, s_sum = 0;
        int n = 0;
        for (size_t i = 0; i < reg.size(); ++i) {
            used[(size_t)reg[i].y * W + reg[i].x] = 0;
            if (dist(xc, yc, reg[i].x, reg[i].y) < rec.width) {
                const double angle = reg[i].angle;
                double ang_d = angle_diff_signed(angle, ang_c);
                sum += ang_d;
                s_sum += ang_d * ang_d;
                ++n;
            }
        }
        double mean_angle = sum / double(n);
        double tau = 2.0 * std::sqrt((s_sum - 2.0 * mean_angle * sum) / double(n) + mean_angle * mean_angle);
        region_grow(reg[0].x, reg[0].y, reg, reg_angle, tau);
        if (reg.size() < 2) return false;
        region2rect(reg, reg_angle, prec, p, rec);
        density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density < density_th) return reduce_region_radius(reg, reg_angle, prec, p, rec, density, density_th);
        return true;
    }

    double nfa(int n, int k, double p) const {
        if (n == 0 || k == 0) return -LOG_NT;
        if (n == k) return -LOG_NT - double(n) * std::log10(p);
        double p_term = p / (1 - p);
        double log1term = log_gamma(double(n) + 1) - log_gamma(double(k) + 1) - log_gamma(double(n - k) + 1) + double(k) * std::log(p) +
                          double(n - k) * std::log(1.0 - p);
        double term = std::exp(log1term);
        if (double_equal(term, 0)) {
            if (k > n * p) return -log1term / M_LN10 - LOG_NT;
            else return -LOG_NT;
        }
        double bin_tail = term;
        double tolerance = 0.1;
        for (int i = k + 1; i <= n; ++i) {
            double bin_term = double(n - i + 1) / double(i);
            double mult_term = bin_term * p_term;
            term *= mult_term;
            bin_tail += term;
            if (bin_term < 1) {
                double err = term * ((1 - std::pow(mult_term, double(n - i + 1))) / (1 - mult_term) - 1);
                if (err < tolerance * std::fabs(-std::log10(bin_tail) - LOG_NT) * bin_tail) break;
            }
        }
        return -std::log10(bin_tail) - LOG_NT;
    }

    // rect_nfa of OpenCV >= 4.5.x (imgproc/src/lsd.cpp): real-valued corners, rotated so that the first one has
    // the smallest y (ties: smallest x); per row y in [ceil(top), ceil(bottom)] the x range is
    // [ceil(left_limit), int(right_limit)] with limits interpolated along the two edge chains.
    // (Recovered from the behaviour of cv2 4.13 — the pre-4.5 integer scan gives different counts — and pinned by
    // tests/test_oracle_line.py against cv2's nfa output.)
    static double get_slope(double px, double py, double qx, double qy) {
        return (int(std::ceil(py)) == int(std::ceil(qy))) ? 0.0 : (qx - px) / (qy - py);
    }
    double rect_nfa(const Rect& rec) const {
        int total_pts = 0, alg_pts = 0;
        double half_width = rec.width / 2.0;
        double dyhw = rec.dy * half_width, dxhw = rec.dx * half_width;
        double vx[4] = {rec.x1 - dyhw, rec.x2 - dyhw, rec.x2 + dyhw, rec.x1 + dyhw};
        double vy[4] = {rec.y1 + dxhw, rec.y2 + dxhw, rec.y2 - dxhw, rec.y1 - dxhw};
        int offset = 0;
        for (int i = 1; i < 4; ++i)
            if (vy[i] < vy[offset] || (vy[i] == vy[offset] && vx[i] < vx[offset])) offset = i;
        double ox[4], oy[4];
        for (int i = 0; i < 4; ++i) { ox[i] = vx[(i + offset) % 4]; oy[i] = vy[(i + offset) % 4]; }
        const double flstep = get_slope(ox[0], oy[0], ox[1], oy[1]);
        const double slstep = get_slope(ox[1], oy[1], ox[2], oy[2]);
        const double frstep = get_slope(ox[0], oy[0], ox[3], oy[3]);
        const double srstep = get_slope(ox[3], oy[3], ox[2], oy[2]);
        const int y_begin = int(std::ceil(oy[0])), y_end = int(std::ceil(oy[2]));
        const int c1 = int(std::ceil(oy[1])), c3 = int(std::ceil(oy[3]));
        for (int y = y_begin; y <= y_end; ++y) {
            if (y < 0 || y >= H) continue;
            double left_limit = (y <= c1) ? ox[0] + (double(y) - oy[0]) * flstep : ox[1] + (double(y) - oy[1]) * slstep;
            double right_limit = (y < c3) ? ox[0] + (double(y) - oy[0]) * frstep : ox[3] + (double(y) - oy[3]) * srstep;
            for (int x = int(std::ceil(left_limit)); x <= int(right_limit); ++x) {
                if (x < 0 || x >= W) continue;
                ++total_pts;
                if (isAligned(x, y, rec.theta, rec.prec)) ++alg_pts;
            }
        }
        return nfa(total_pts, alg_pts, rec.p);
    }

    double rect_improve(Rect& rec) const {
        double delta = 0.5, delta_2 = delta / 2.0;
        double log_nfa = rect_nfa(rec);
        if (log_nfa > LOG_EPS) return log_nfa;
        Rect r = rec;
        for (int n = 0; n < 5; ++n) {
            r.p /= 2;
            r.prec = r.p * kPi;
            double v = rect_nfa(r);
            if (v > log_nfa) { log_nfa = v; rec = r; }
        }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (unsigned n = 0; n < 5; ++n) {
            if ((r.width - delta) >= 0.5) {
                r.width -= delta;
                double v = rect_nfa(r);
                if (v > log_nfa) { rec = r; log_nfa = v; }
            }
        }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (unsigned n = 0; n < 5; ++n) {
            if ((r.width - delta) >= 0.5) {
                r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2;
                r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2;
                r.width -= delta;
                double v = rect_nfa(r);
                if (v > log_nfa) { rec = r; log_nfa = v; }
            }
        }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (unsigned n = 0; n < 5; ++n) {
            if ((r.width - delta) >= 0.5) {
                r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2;
                r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2;
                r.width -= delta;
                double v = rect_nfa(r);
                if (v > log_nfa) { rec = r; log_nfa = v; }
            }
        }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (unsigned n = 0; n < 5; ++n) {
            if ((r.width - delta) >= 0.5) {
                r.p /= 2;
                r.prec = r.p * kPi;
                double v = rect_nfa(r);
                if (v > log_nfa) { rec = r; log_nfa = v; }
            }
        }
        return log_nfa;
    }

    struct Seg { float x1, y1, x2, y2; double width, p, nfa; };

    std::vector<int>* trace = nullptr;  // lab hook: per grown region {seed pixel, size of the first growth, final size, outcome}

    void detect(const uint8_t* img, int w, int h, size_t step, std::vector<Seg>& out) {
        out.clear();
        const double prec = kPi * ANG_TH / 180;
        const double p = ANG_TH / 180;
        const double rho = QUANT / std::sin(prec);
        // GaussianBlur(image, 7x7, sigma = 0.6/0.8) in 8-bit fixed point, then resize(0.8, INTER_LINEAR_EXACT)
        const double sigma = SIGMA_SCALE / SCALE;
        const double sprec = 3;
        const unsigned hk = (unsigned)(std::ceil(sigma * std::sqrt(2 * sprec * std::log(10.0))));
        (void)hk;  // == 3 -> 7x7; the 8.8 fixed-point kernel of (7, 0.75) is {0,4,56,136,56,4,0} (pinned to cv2 by test)
        static const int k7[7] = {0, 4, 56, 136, 56, 4, 0};
        std::vector<uint8_t> blurred((size_t)w * h);
        orc_gaussian_blur_fixed_u8(img, w, h, step, blurred.data(), w, k7, 7);
        resize_linear_exact(blurred, w, h, SCALE, SCALE, scaled, W, H);
        ll_angle(rho);
        LOG_NT = 5 * (std::log10(double(W)) + std::log10(double(H))) / 2 + std::log10(11.0);
        const size_t min_reg_size = size_t(-LOG_NT / std::log10(p));
        used.assign((size_t)W * H, 0);
        std::vector<RegionPoint> reg;
        for (size_t i = 0; i < ordered.size(); ++i) {
            const int px = ordered[i].x, py = ordered[i].y;
            if (used[(size_t)py * W + px] == 0 && ang(px, py) != NOTDEF) {
                double reg_angle;
                region_grow(px, py, reg, reg_angle, prec);
                const int n_first = (int)reg.size();
                if (reg.size() < min_reg_size) {
                    if (trace) trace->insert(trace->end(), {py * W + px, n_first, n_first, 0});
                    continue;
                }
                Rect rec;
                region2rect(reg, reg_angle, prec, p, rec);
                const bool refined = refine(reg, reg_angle, prec, p, rec, DENSITY_TH);
                if (trace) trace->insert(trace->end(), {py * W + px, n_first, (int)reg.size(), refined ? 1 : 0});
                if (!refined) continue;
                double log_nfa = rect_improve(rec);
                if (log_nfa <= LOG_EPS) continue;
                rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
                rec.x1 /= SCALE; rec.y1 /= SCALE; rec.x2 /= SCALE; rec.y2 /= SCALE;
                rec.width /= SCALE;
                out.push_back({float(rec.x1), float(rec.y1), float(rec.x2), float(rec.y2), rec.width, rec.p, log_nfa});
            }
        }
    }
};

// ---- LSDDetector::detectImpl keyline fill (opencv_contrib LSDDetector.cpp) ----
void fill_keyline(const Lsd::Seg& s, int cols, int rows, int class_id, pl_keyline& kl) {
    float e[4] = {s.x1, s.y1, s.x2, s.y2};
    // checkLineExtremes
    if (e[0] < 0) e[0] = 0;
    if (e[0] >= cols) e[0] = (float)cols - 1.0f;
    if (e[2] < 0) e[2] = 0;
    if (e[2] >= cols) e[2] = (float)cols - 1.0f;
    if (e[1] < 0) e[1] = 0;
    if (e[1] >= rows) e[1] = (float)rows - 1.0f;
    if (e[3] < 0) e[3] = 0;
    if (e[3] >= rows) e[3] = (float)rows - 1.0f;
    const float octaveScale = 1.0f;  // pow((float)scale, 0)
    kl.sx = e[0] * octaveScale; kl.sy = e[1] * octaveScale; kl.ex = e[2] * octaveScale; kl.ey = e[3] * octaveScale;
    kl.sx_oct = e[0]; kl.sy_oct = e[1]; kl.ex_oct = e[2]; kl.ey_oct = e[3];
    kl.length = (float)std::sqrt(std::pow(e[0] - e[2], 2) + std::pow(e[1] - e[3], 2));
    // cv::LineIterator(img, Point2f, Point2f).count: endpoints cvRound'ed, inside the image -> max(|dx|,|dy|)+1
    int x0 = cvRoundF(e[0]), y0 = cvRoundF(e[1]), x1 = cvRoundF(e[2]), y1 = cvRoundF(e[3]);
    kl.num_pixels = std::max(std::abs(x1 - x0), std::abs(y1 - y0)) + 1;
    kl.angle = (float)std::atan2((double)(kl.ey - kl.sy), (double)(kl.ex - kl.sx));
    kl.class_id = class_id;
    kl.octave = 0;
    kl.size = (kl.ex - kl.sx) * (kl.ey - kl.sy);
    kl.response = kl.length / std::max(cols, rows);
    kl.pt_x = (kl.ex + kl.sx) / 2;
    kl.pt_y = (kl.ey + kl.sy) / 2;
}

// ---- BinaryDescriptor (LBD) ----
const int kBands = 9, kBandW = 7;
const int kComb[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7},
                          {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8}, {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

struct Lbd {
    int W = 0, H = 0;
    std::vector<int16_t> dx, dy;
    double gaussL[kBandW * 3], gaussG[kBands * kBandW];
    Lbd() {
        double u = (kBandW * 3 - 1) / 2;
        double sigma = (kBandW * 2 + 1) / 2;
        double invsigma2 = -1 / (2 * sigma * sigma);
        for (int i = 0; i < kBandW * 3; i++) {
            double dis = i - u;
            gaussL[i] = std::exp(dis * dis * invsigma2);
        }
        u = (kBands * kBandW - 1) / 2;
        sigma = kBands * kBandW / 2;
        invsigma2 = -1 / (2 * sigma * sigma);
        for (int i = 0; i < kBands * kBandW; i++) {
            double dis = i - u;
            gaussG[i] = std::exp(dis * dis * invsigma2);
        }
    }
    // computeSobel: GaussianBlur(5x5, sigma 1) then Sobel 3x3 CV_16S dx / dy (BORDER_REFLECT_101)
    void prepare(const uint8_t* img, int w, int h, size_t step) {
        W = w; H = h;
        static const int k5[5] = {14, 62, 104, 62, 14};
        std::vector<uint8_t> b((size_t)w * h);
        orc_gaussian_blur_fixed_u8(img, w, h, step, b.data(), w, k5, 5);
        dx.assign((size_t)w * h, 0);
        dy.assign((size_t)w * h, 0);
        for (int y = 0; y < h; y++) {
            const uint8_t* r0 = &b[(size_t)reflect101(y - 1, h) * w];
            const uint8_t* r1 = &b[(size_t)y * w];
            const uint8_t* r2 = &b[(size_t)reflect101(y + 1, h) * w];
            for (int x = 0; x < w; x++) {
                int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
                dx[(size_t)y * w + x] = (int16_t)((r0[xp] + 2 * r1[xp] + r2[xp]) - (r0[xm] + 2 * r1[xm] + r2[xm]));
                dy[(size_t)y * w + x] = (int16_t)((r2[xm] + 2 * r2[x] + r2[xp]) - (r0[xm] + 2 * r0[x] + r0[xp]));
            }
        }
    }
    void describe(const pl_keyline& kl, float* desVec /*72*/, uint8_t* bits /*32*/) const {
        const short heightOfLSP = (short)(kBandW * kBands);
        const short descriptor_size = kBands * 8;
        float pgdLBandSum[kBands] = {0}, ngdLBandSum[kBands] = {0}, pgdL2BandSum[kBands] = {0}, ngdL2BandSum[kBands] = {0};
        float pgdOBandSum[kBands] = {0}, ngdOBandSum[kBands] = {0}, pgdO2BandSum[kBands] = {0}, ngdO2BandSum[kBands] = {0};
        const short halfHeight = (heightOfLSP - 1) / 2;
        const short realWidth = (short)W;
        const short imageWidth = realWidth - 1, imageHeight = (short)(H - 1);
        const short lengthOfLSP = (short)kl.num_pixels;
        const short halfWidth = (lengthOfLSP - 1) / 2;
        const float lineMiddlePointX = (kl.sx_oct + kl.ex_oct) / 2;
        const float lineMiddlePointY = (kl.sy_oct + kl.ey_oct) / 2;
        float dL[2], dO[2];
        dL[0] = (float)std::cos((double)kl.angle);
        dL[1] = (float)std::sin((double)kl.angle);
        dO[0] = -dL[1];
        dO[1] = dL[0];
        float sCorX0 = -dL[0] * halfWidth + dL[1] * halfHeight + lineMiddlePointX;
        float sCorY0 = -dL[1] * halfWidth - dL[0] * halfHeight + lineMiddlePointY;
        for (short hID = 0; hID < heightOfLSP; hID++) {
            float sCorX = sCorX0, sCorY = sCorY0;
            float pgdLRowSum = 0, ngdLRowSum = 0, pgdORowSum = 0, ngdORowSum = 0;
            for (short wID = 0; wID < lengthOfLSP; wID++) {
                short tempCor = (short)std::round(sCorX);
                short xCor = (tempCor < 0) ? 0 : (tempCor > imageWidth) ? imageWidth : tempCor;
                tempCor = (short)std::round(sCorY);
                short yCor = (tempCor < 0) ? 0 : (tempCor > imageHeight) ? imageHeight : tempCor;
                short ddx = dx[(size_t)yCor * realWidth + xCor], ddy = dy[(size_t)yCor * realWidth + xCor];
                float gDL = ddx * dL[0] + ddy * dL[1];
                float gDO = ddx * dO[0] + ddy * dO[1];
                if (gDL > 0) pgdLRowSum += gDL; else ngdLRowSum -= gDL;
                if (gDO > 0) pgdORowSum += gDO; else ngdORowSum -= gDO;
                sCorX += dL[0];
                sCorY += dL[1];
            }
            sCorX0 -= dL[1];
            sCorY0 += dL[0];
            float coef = (float)gaussG[hID];
            pgdLRowSum = coef * pgdLRowSum;
            ngdLRowSum = coef * ngdLRowSum;
            float pgdL2RowSum = pgdLRowSum * pgdLRowSum, ngdL2RowSum = ngdLRowSum * ngdLRowSum;
            pgdORowSum = coef * pgdORowSum;
            ngdORowSum = coef * ngdORowSum;
            float pgdO2RowSum = pgdORowSum * pgdORowSum, ngdO2RowSum = ngdORowSum * ngdORowSum;
            auto add = [&](short band, float c) {
                pgdLBandSum[band] += c * pgdLRowSum;
                ngdLBandSum[band] += c * ngdLRowSum;
                pgdL2BandSum[band] += c * c * pgdL2RowSum;
                ngdL2BandSum[band] += c * c * ngdL2RowSum;
                pgdOBandSum[band] += c * pgdORowSum;
                ngdOBandSum[band] += c * ngdORowSum;
                pgdO2BandSum[band] += c * c * pgdO2RowSum;
                ngdO2BandSum[band] += c * c * ngdO2RowSum;
            };
            short bandID = (short)(hID / kBandW);
            add(bandID, (float)gaussL[hID % kBandW + kBandW]);
            bandID--;
            if (bandID >= 0) add(bandID, (float)gaussL[hID % kBandW + 2 * kBandW]);
            bandID = bandID + 2;
            if (bandID < kBands) add(bandID, (float)gaussL[hID % kBandW]);
        }
        const float invN2 = (float)(1.0 / (kBandW * 2.0)), invN3 = (float)(1.0 / (kBandW * 3.0));
        for (short bandID = 0; bandID < kBands; bandID++) {
            const float invN = (bandID == 0 || bandID == kBands - 1) ? invN2 : invN3;
            const short desID = bandID * 8;
            float temp = pgdLBandSum[bandID] * invN;
            desVec[desID] = temp;
            desVec[desID + 4] = std::sqrt(pgdL2BandSum[bandID] * invN - temp * temp);
            temp = ngdLBandSum[bandID] * invN;
            desVec[desID + 1] = temp;
            desVec[desID + 5] = std::sqrt(ngdL2BandSum[bandID] * invN - temp * temp);
            temp = pgdOBandSum[bandID] * invN;
            desVec[desID + 2] = temp;
            desVec[desID + 6] = std::sqrt(pgdO2BandSum[bandID] * invN - temp * temp);
            temp = ngdOBandSum[bandID] * invN;
            desVec[desID + 3] = temp;
            desVec[desID + 7] = std::sqrt(ngdO2BandSum[bandID] * invN - temp * temp);
        }
        float tempM = 0, tempS = 0;
        for (short b = 0; b < kBands; b++) {
            const float* d = desVec + b * 8;
            tempM += d[0] * d[0]; tempM += d[1] * d[1]; tempM += d[2] * d[2]; tempM += d[3] * d[3];
            tempS += d[4] * d[4]; tempS += d[5] * d[5]; tempS += d[6] * d[6]; tempS += d[7] * d[7];
        }
        tempM = 1 / std::sqrt(tempM);
        tempS = 1 / std::sqrt(tempS);
        for (short b = 0; b < kBands; b++) {
            float* d = desVec + b * 8;
            d[0] *= tempM; d[1] *= tempM; d[2] *= tempM; d[3] *= tempM;
            d[4] *= tempS; d[5] *= tempS; d[6] *= tempS; d[7] *= tempS;
        }
        for (short i = 0; i < descriptor_size; i++)
            if (desVec[i] > 0.4) desVec[i] = (float)0.4;
        float temp = 0;
        for (short i = 0; i < descriptor_size; i++) temp += desVec[i] * desVec[i];
        temp = 1 / std::sqrt(temp);
        for (short i = 0; i < descriptor_size; i++) desVec[i] = desVec[i] * temp;
        for (int c = 0; c < 32; c++) {
            const float* f1 = desVec + 8 * kComb[c][0];
            const float* f2 = desVec + 8 * kComb[c][1];
            uint8_t r = 0;
            for (int i = 0; i < 8; i++)
                if (f1[i] > f2[i]) r += (uint8_t)(1 << i);
            bits[c] = r;
        }
    }
};

}  // namespace

// cv::LineSegmentDetector(LSD_REFINE_ADV)::detect — returns the number of segments (all are written up to cap)
extern "C" int orc_lsd_detect(const uint8_t* img, int rows, int cols, size_t step, int order_mode, float* xyxy, double* width,
                              double* prec, double* nfa, int cap) {
    Lsd lsd;
    lsd.order_mode = order_mode;
    std::vector<Lsd::Seg> segs;
    lsd.detect(img, cols, rows, step, segs);
    for (size_t i = 0; i < segs.size() && (int)i < cap; i++) {
        xyxy[4 * i] = segs[i].x1; xyxy[4 * i + 1] = segs[i].y1; xyxy[4 * i + 2] = segs[i].x2; xyxy[4 * i + 3] = segs[i].y2;
        if (width) width[i] = segs[i].width;
        if (prec) prec[i] = segs[i].p;
        if (nfa) nfa[i] = segs[i].nfa;
    }
    return (int)segs.size();
}

// lab hook: the regions flsd() grows, in order: {seed pixel (y * W + x of the scaled image), size of the first growth, final size
// (after refine), 1 = got a rectangle}; returns the number of regions
extern "C" int orc_lsd_trace(const uint8_t* img, int rows, int cols, size_t step, int* out4, int cap) {
    Lsd lsd;
    std::vector<int> tr;
    lsd.trace = &tr;
    std::vector<Lsd::Seg> segs;
    lsd.detect(img, cols, rows, step, segs);
    const int n = (int)(tr.size() / 4);
    for (int i = 0; i < n && i < cap; i++)
        for (int k = 0; k < 4; k++) out4[4 * i + k] = tr[4 * i + k];
    return n;
}

// lab hook: level-line angle map (radians, NOTDEF=-1024) of the scaled image
extern "C" int orc_lsd_angles(const uint8_t* img, int rows, int cols, size_t step, double* out, int* ow, int* oh) {
    Lsd lsd;
    static const int k7[7] = {0, 4, 56, 136, 56, 4, 0};
    std::vector<uint8_t> blurred((size_t)cols * rows);
    orc_gaussian_blur_fixed_u8(img, cols, rows, step, blurred.data(), cols, k7, 7);
    resize_linear_exact(blurred, cols, rows, 0.8, 0.8, lsd.scaled, lsd.W, lsd.H);
    lsd.ll_angle(2.0 / std::sin(kPi * 22.5 / 180));
    *ow = lsd.W; *oh = lsd.H;
    if (out) std::memcpy(out, lsd.angles.data(), sizeof(double) * lsd.angles.size());
    return 0;
}

// test hook: the scaled 8-bit image LSD works on
extern "C" int orc_lsd_scaled(const uint8_t* img, int rows, int cols, size_t step, uint8_t* out, int* ow, int* oh) {
    static const int k7[7] = {0, 4, 56, 136, 56, 4, 0};
    std::vector<uint8_t> blurred((size_t)cols * rows), sc;
    orc_gaussian_blur_fixed_u8(img, cols, rows, step, blurred.data(), cols, k7, 7);
    int w, h;
    resize_linear_exact(blurred, cols, rows, 0.8, 0.8, sc, w, h);
    *ow = w; *oh = h;
    if (out) std::memcpy(out, sc.data(), sc.size());
    return 0;
}

// LBD of given keylines (BinaryDescriptor::compute): desc n x 32 bytes, optional float descriptors n x 72
extern "C" int orc_lbd_compute(const uint8_t* img, int rows, int cols, size_t step, const pl_keyline* kls, int n, uint8_t* desc,
                               float* fdesc) {
    Lbd lbd;
    lbd.prepare(img, cols, rows, step);
    float tmp[72];
    for (int i = 0; i < n; i++) lbd.describe(kls[i], fdesc ? fdesc + 72 * (size_t)i : tmp, desc + 32 * (size_t)i);
    return 0;
}

// LineExtractor::ExtractLineSegment — src/LineExtractor.cpp:12-70
extern "C" int orc_line_extract(const uint8_t* img, int rows, int cols, size_t step, int max_lines, int order_mode, pl_keyline* kls,
                                uint8_t* desc, double* coeffs, int* n_out) {
    *n_out = 0;
    if (!img || rows <= 0 || cols <= 0) return PL_ERR_EMPTY;
    Lsd lsd;
    lsd.order_mode = order_mode;
    std::vector<Lsd::Seg> segs;
    lsd.detect(img, cols, rows, step, segs);
    std::vector<pl_keyline> key_lines(segs.size());
    for (size_t i = 0; i < segs.size(); i++) fill_keyline(segs[i], cols, rows, (int)i, key_lines[i]);
    if ((int)key_lines.size() > max_lines) {
        // :27-30 — sort by response, descending.  The reference uses std::sort, whose order among EXACT float ties is
        // an artefact of libstdc++'s introsort; the contract here is: ties keep detection order (stable).
        std::stable_sort(key_lines.begin(), key_lines.end(), [](const pl_keyline& a, const pl_keyline& b) { return a.response > b.response; });
        key_lines.resize(max_lines);
    }
    const int n = (int)key_lines.size();
    if (n) {
        Lbd lbd;
        lbd.prepare(img, cols, rows, step);
        float tmp[72];
        for (int i = 0; i < n; i++) lbd.describe(key_lines[i], tmp, desc + 32 * (size_t)i);
    }
    for (int i = 0; i < n; i++) {
        kls[i] = key_lines[i];
        // :60-69 — l = (sx,sy,1) x (ex,ey,1), normalised (Eigen::Vector3d)
        const double sx = key_lines[i].sx, sy = key_lines[i].sy, ex = key_lines[i].ex, ey = key_lines[i].ey;
        double l0 = sy * 1.0 - 1.0 * ey, l1 = 1.0 * ex - sx * 1.0, l2 = sx * ey - sy * ex;
        double nrm = std::sqrt(l0 * l0 + l1 * l1 + l2 * l2);
        if (nrm > 0) { l0 /= nrm; l1 /= nrm; l2 /= nrm; }
        coeffs[3 * i] = l0; coeffs[3 * i + 1] = l1; coeffs[3 * i + 2] = l2;
    }
    *n_out = n;
    return PL_OK;
}
