// host_cpu_check.cc -- the parts of the host shim that need no GPU, checked on their own:
//   * minimal::drawSamples: every RANSAC trial gets its own 9 distinct point indices (the reference's
//     draw, minimal.cc:226-244, reseeds per trial; drawn back to back that would repeat the sample)
//   * SubsetDrawer reproduces the reference's persistent-permutation draw for a given rand() stream
//   * cv::imwrite / cv::imread (stand-in) round-trip 8-bit grey and BGR images
//   * the VelocityErrors / TrueValues surface of errorMeasure.h:18-44
// Exit code 0 = all good; prints the first failure otherwise.
#include <algorithm>
#include <cstdio>
#include <set>

#include "errorMeasure.h"
#include "minimal.h"

#define CHECK(cond, msg) do { if (!(cond)) { std::printf("FAILED: %s\n", msg); return 1; } } while (0)

int main(int argc, char **argv)
{
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    // ---- distinct samples per trial
    const int n = 2073600, H = 50;
    std::vector<int32_t> s = minimal::drawSamples(n, H);
    CHECK((int)s.size() == 9 * H, "drawSamples size");
    std::set<std::vector<int32_t>> distinct;
    for (int t = 0; t < H; ++t) {
        std::vector<int32_t> one(s.begin() + 9 * t, s.begin() + 9 * t + 9);
        std::set<int32_t> u(one.begin(), one.end());
        CHECK(u.size() == 9, "a trial repeats a point");
        for (int32_t i : one) CHECK(i >= 0 && i < n, "index out of range");
        std::sort(one.begin(), one.end());
        distinct.insert(one);
    }
    CHECK((int)distinct.size() == H, "two trials drew the same sample");
    // ---- the draw itself against a literal transcription of the rule "swap the pick to the end of the live range"
    {
        const int m = 40;
        unsigned state = 12345u;
        auto lcg = [&state] { state = state * 1103515245u + 12345u; return (int)((state >> 16) & 0x7fff); };
        std::vector<int32_t> got;
        minimal::SubsetDrawer drawer(m);
        for (int t = 0; t < 3; ++t) drawer.draw(lcg, got);
        state = 12345u;
        std::vector<int> idx(m);
        for (int i = 0; i < m; ++i) idx[i] = i;
        std::vector<int32_t> want;
        for (int t = 0; t < 3; ++t)
            for (int j = 0, live = m; j < 9; ++j, --live) { std::swap(idx[live - 1], idx[lcg() % live]); want.push_back(idx[live - 1]); }
        CHECK(got == want, "SubsetDrawer differs from the reference rule");
    }
    // ---- PNG round trip
    {
        cv::Mat g(37, 53, CV_8UC1), c(300, 411, CV_8UC3);           // the colour image spans several 64 KB stored blocks
        for (int y = 0; y < g.rows; ++y) for (int x = 0; x < g.cols; ++x) g.at<unsigned char>(y, x) = (unsigned char)(x * 7 + y * 13);
        for (int y = 0; y < c.rows; ++y) for (int x = 0; x < c.cols; ++x) c.at<cv::Vec3b>(y, x) = cv::Vec3b((unsigned char)x, (unsigned char)(y + x), (unsigned char)(3 * y));
        CHECK(cv::imwrite(dir + "/rsdsfm_g.png", g) && cv::imwrite(dir + "/rsdsfm_c.png", c), "imwrite");
        cv::Mat g2 = cv::imread(dir + "/rsdsfm_g.png", 0), c2 = cv::imread(dir + "/rsdsfm_c.png");
        CHECK(g2.rows == g.rows && g2.cols == g.cols && g2.type() == CV_8UC1 && !std::memcmp(g.data, g2.data, g.total_bytes()), "grey PNG round trip");
        CHECK(c2.rows == c.rows && c2.cols == c.cols && c2.type() == CV_8UC3 && !std::memcmp(c.data, c2.data, c.total_bytes()), "BGR PNG round trip");
    }
    // ---- errorMeasure.h surface
    {
        error_measure::TrueValues tv(Eigen::Vector3d(1, 2, 3), Eigen::Vector3d(4, 5, 6));
        CHECK(tv.w(0) == 1 && tv.v(2) == 6, "TrueValues(w, v)");
        Eigen::Array3Xd a = Eigen::Array3Xd::Zero(3, 2);
        Eigen::ArrayXd b = Eigen::ArrayXd::Zero(2);
        a(0, 1) = 7.0;
        error_measure::VelocityErrors e(a, a, b, b, a, a, 0.1, 0.2, 0.3);
        CHECK(e.error_w == 0.1 && e.error_v == 0.2 && e.error_reproject == 0.3 && e.error_v_vec.col(1)(0) == 7.0 && e.k(1) == 0.0, "VelocityErrors");
    }
    std::printf("host shim CPU checks: ok\n");
    return 0;
}
