// host_cpu_check.cc -- the parts of the host shim that need no GPU, checked on their own:
//   * minimal::drawSamples: every RANSAC trial gets its own 9 distinct point indices (the reference's
//     draw, minimal.cc:226-244, reseeds per trial; drawn back to back that would repeat the sample)
//   * SubsetDrawer reproduces the reference's persistent-permutation draw for a given rand() stream
//   * cv::imwrite / cv::imread (stand-in) round-trip 8-bit grey and BGR images
//   * the VelocityErrors / TrueValues surface of errorMeasure.h:18-44
//   * the fixture loaders (intrinsics, scanline poses, RS / GS unprojection maps) against files written here,
//     setSyntheticDepthMapRs / Gs, and the flow visualisations (getImageOpticalFlow, flowArrows)
// Exit code 0 = all good; prints the first failure otherwise.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <functional>
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
    // ---- fixture files: write a synthetic frame's CSVs, load them through the reference-named loaders, compare
    {
        const int rows = 9, cols = 13;
        auto wr = [&](const std::string &name, int nl, int nf, const std::function<double(int, int)> &f) {
            std::ofstream o(dir + "/" + name);
            o.precision(17);
            for (int i = 0; i < nl; ++i) { for (int j = 0; j < nf; ++j) o << (j ? "," : "") << f(i, j); o << "\n"; }
        };
        auto ux = [](int y, int x) { return (y == 2 && x == 3) ? 0.0 : 0.1 * x - 0.2 * y + 0.01; };
        auto uy = [](int y, int x) { return (y == 2 && x == 3) ? 0.0 : 0.3 * y - 0.05 * x; };
        auto uz = [](int y, int x) { return (y == 2 && x == 3) ? 0.0 : 4.0 + 0.02 * x * y; };
        wr("1_rs_unproject_x.csv", rows, cols, ux); wr("1_rs_unproject_y.csv", rows, cols, uy); wr("1_rs_unproject_z.csv", rows, cols, uz);
        wr("1_gs_unproject_x.csv", rows, cols, [&](int y, int x) { return ux(y, x) + 0.5; });
        wr("1_gs_unproject_y.csv", rows, cols, [&](int y, int x) { return uy(y, x) - 0.25; });
        wr("1_gs_unproject_z.csv", rows, cols, [&](int y, int x) { return uz(y, x) + 1.0; });
        wr("1_rs_t.csv", rows, 3, [](int i, int j) { return 0.01 * i * (j + 1); });
        wr("1_rs_r.csv", rows, 9, [](int i, int j) { return (j % 4 == 0 ? 1.0 : 0.0) + (j == 1 ? 1e-3 * i : 0.0) - (j == 3 ? 1e-3 * i : 0.0); });
        wr("A.csv", 3, 3, [](int i, int j) { const double K[9] = {500, 0, 6, 0, 510, 4, 0, 0, 1}; return K[3 * i + j]; });
        Camera cam;
        CHECK(cam.loadIntrinsicsFromFile(dir + "/A.csv", false), "loadIntrinsicsFromFile");
        CHECK(cam.getIntrinsics()(0, 0) == 500 && cam.getIntrinsics()(1, 2) == 4, "intrinsics values");
        RsFrame fr;
        fr.setIntrinsics(cam.getIntrinsics());
        fr.setImage(cv::Mat(rows, cols, CV_8UC3));
        CHECK(fr.setPoses(dir + "/1_rs_t.csv", dir + "/1_rs_r.csv"), "setPoses");
        CHECK(fr.setUnprojectionMapRs(dir + "/1_rs_unproject_x.csv", dir + "/1_rs_unproject_y.csv", dir + "/1_rs_unproject_z.csv"), "setUnprojectionMapRs");
        CHECK(fr.setUnprojectionMapGs(dir + "/1_gs_unproject_x.csv", dir + "/1_gs_unproject_y.csv", dir + "/1_gs_unproject_z.csv"), "setUnprojectionMapGs");
        CHECK(!fr.setUnprojectionMapGs(dir + "/1_rs_t.csv", dir + "/1_rs_t.csv", dir + "/1_rs_t.csv"), "a file of the wrong shape must be refused");
        const Eigen::Vector3d P = fr.getUnprojectedWorldCoordinates(Eigen::Vector2d(5, 7));
        CHECK(P.x() == ux(7, 5) && P.y() == uy(7, 5) && P.z() == uz(7, 5), "unprojection map values survive the round trip");
        fr.setSyntheticDepthMapRs();
        fr.setSyntheticDepthMapGs();
        // row 7: R = I + small skew, t = (0.07, 0.14, 0.21): z_cam = R.row(2) . P + t.z
        const double z_rs = uz(7, 5) + 0.01 * 7 * 3;
        CHECK(std::fabs(fr.getDepthMap()(7, 5) - z_rs) < 1e-12, "setSyntheticDepthMapRs uses the pixel's own scanline");
        CHECK(std::fabs(fr.getGsDepthMap()(7, 5) - (uz(7, 5) + 1.0)) < 1e-12, "setSyntheticDepthMapGs uses scanline 0");
        CHECK(fr.getDepthMap()(2, 3) == 0.0 && fr.getGsDepthMap()(2, 3) != 0.0, "a pixel without a world point has depth 0");
    }
    // ---- flow visualisations: colours of the four axis directions, arrows where the flow is
    {
        cv::Mat_<cv::Point_<double>> flow(20, 30);
        for (int y = 0; y < 20; ++y) for (int x = 0; x < 30; ++x) flow(y, x) = cv::Point_<double>(0, 0);
        flow(0, 0) = cv::Point_<double>(10, 0); flow(0, 10) = cv::Point_<double>(0, 10); flow(10, 0) = cv::Point_<double>(-5, 0);
        Camera cam;
        cv::Mat hsv = cam.getImageOpticalFlow(flow);
        const cv::Vec3f right = hsv.at<cv::Vec3f>(0, 0), down = hsv.at<cv::Vec3f>(0, 10), left = hsv.at<cv::Vec3f>(10, 0), none = hsv.at<cv::Vec3f>(5, 5);
        CHECK(right[2] == 1.f && right[1] == 0.f && right[0] == 0.f, "flow to the right: hue 0 = red at full value");
        CHECK(std::fabs(down[1] - 1.f) < 1e-5f && std::fabs(down[2] - 0.5f) < 1e-5f && down[0] == 0.f, "flow downwards: hue 90");
        CHECK(std::fabs(left[1] - 0.5f) < 1e-5f && std::fabs(left[0] - 0.5f) < 1e-5f && left[2] == 0.f, "flow to the left: hue 180 at half value");
        CHECK(none[0] == 0.f && none[1] == 0.f && none[2] == 0.f, "no flow: black");
        cv::Mat img(20, 30, CV_8UC3);
        std::memset(img.data, 0, img.total_bytes());
        cv::Mat arrows = cam.flowArrows(img, flow, 10, 10);
        CHECK(arrows.at<cv::Vec3b>(0, 5)[2] == 255 && arrows.at<cv::Vec3b>(5, 10)[2] == 255 && arrows.at<cv::Vec3b>(10, 0)[2] == 255, "arrow shafts");
        CHECK(arrows.at<cv::Vec3b>(15, 25)[2] == 0 && arrows.at<cv::Vec3b>(10, 20)[2] == 0, "no arrow without flow");
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
