// example_single_run.cc -- the glue of the reference's evaluateSingleRun() (main.cc:398-523) written
// against the host shim: flatten -> alpha -> RANSAC -> nonLinearRefinement -> sign fix -> depth
// raster -> setPose -> backProject -> interpolateCrackyImage, on a small analytic RS pair.
//
//   g++ -std=c++17 -O2 -I include -I rs-aware-differential-sfm_b200/host
//       rs-aware-differential-sfm_b200/host/example_single_run.cc
//       -L rs-aware-differential-sfm_b200 -lrsdsfm -Wl,-rpath,'$ORIGIN/..' -o rs-aware-differential-sfm_b200/host/example_single_run
//
// Prints the recovered motion; exits non-zero if w is not recovered (exact constant-velocity data).
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>

#include "camera.h"
#include "errorMeasure.h"
#include "minimal.h"
#include "nonlinearRefinement.h"

using Eigen::ArrayXd;
using nonlinear_refinement::nonLinearRefinement;

int main()
{
    const int rows = 240, cols = 320;
    const double gamma = 0.95;
    const Eigen::Vector3d v_true(0.30, 0.05, 0.02), w_true(0.002, -0.004, 0.0087);

    Camera camera;
    Eigen::Matrix3d Kin;
    Kin(0, 0) = 400.0; Kin(1, 1) = 398.0; Kin(0, 2) = 160.0; Kin(1, 2) = 120.0; Kin(2, 2) = 1.0;
    camera.setIntrinsics(Kin);
    cv::Mat rs1(rows, cols, CV_8UC3), rs2(rows, cols, CV_8UC3);
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x)
            rs1.at<cv::Vec3b>(y, x) = cv::Vec3b((unsigned char)(40 + (x * 3 + y) % 200), (unsigned char)(60 + (x + 2 * y) % 180), (unsigned char)(90 + (x * y) % 150));
    camera.addFrameReal(rs1);
    camera.addFrameReal(rs2);
    camera.setGamma(gamma);

    // exact constant-velocity RS flow: u = alpha (A v d + B w), alpha = 1 + gamma flow_y / h  (closed form)
    Eigen::Matrix3d K = camera.getIntrinsics();
    const double f_x = K(0, 0), f_y = K(1, 1), c_x = K(0, 2), c_y = K(1, 2);
    cv::Mat_<cv::Point_<double>> flow_image(rows, cols);
    for (int j = 0; j < rows; ++j)
        for (int i = 0; i < cols; ++i) {
            const double x = (i - c_x) / f_x, y = (j - c_y) / f_y;
            const double d = 0.08 + 0.05 * std::sin(0.03 * i) * std::cos(0.02 * j) + 0.03 * x;
            const double gx = (v_true(0) - x * v_true(2)) * d + (-x * y * w_true(0) + (1 + x * x) * w_true(1) - y * w_true(2));
            const double gy = (v_true(1) - y * v_true(2)) * d + (-(1 + y * y) * w_true(0) + x * y * w_true(1) + x * w_true(2));
            const double uy = gy / (1.0 - gy * f_y / rows), a = 1.0 + uy * f_y / rows;
            flow_image(j, i) = cv::Point_<double>(a * gx * f_x / gamma, uy * f_y / gamma);
        }
    camera.setCachedFlow(flow_image);

    // ---- main.cc:398-432
    error_measure::Flattened F = error_measure::flattenFlow(camera.calculateDeepFlow(1, 2), K, gamma, 1e-10, false);
    Eigen::Matrix2Xd &coord = F.coord, &flow = F.flow, &coord_pixel = F.coord_pixel, &flow_pixel = F.flow_pixel;
    // ---- main.cc:437-438
    ArrayXd alpha = minimal::getAlpha(flow_pixel, rows, gamma);
    ArrayXd alphaK = minimal::getAlphaK(coord_pixel, flow_pixel, rows, gamma);
    // ---- main.cc:447
    RansacValues ransac_results = minimal::ransac(coord, flow, alpha, alphaK, false, 5, 0.05, false);
    std::cout << "ransac numInliers: " << ransac_results.num_inliers << std::endl;
    std::cout << "ransac w: " << ransac_results.w.transpose() << std::endl;
    std::cout << "ransac v: " << ransac_results.v.transpose() << std::endl;
    // ---- main.cc:455-458
    RansacValues results = nonLinearRefinement(flow, ransac_results, false, true);
    // ---- main.cc:466-478
    double count_z = 0;
    for (int i = 0; i < results.num_inliers; ++i) count_z += results.inliers(2, i);
    if (count_z * 1.0 / results.num_inliers < 0) { results.inliers.row(2) *= -1.0; results.v *= -1.0; }
    std::cout << "final w: " << results.w.transpose() << std::endl;
    std::cout << "final v: " << results.v.transpose() << std::endl;
    // ---- main.cc:496-509
    Eigen::MatrixXd depth_map = Eigen::MatrixXd::Zero(rows, cols);
    for (int i = 0; i < results.num_inliers; i++) {
        const int x = int(f_x * results.inliers(0, i) + c_x + 0.5), y = int(f_y * results.inliers(1, i) + c_y + 0.5);
        if (x >= 0 && x < cols && y >= 0 && y < rows) depth_map(y, x) = results.inliers(2, i);
    }
    // ---- main.cc:516-523
    camera.setPose(1, results.k, results.v, results.w);
    camera.setDepthMap(1, depth_map);
    camera.backProject(1);
    cv::Mat backprojection = camera.interpolateCrackyImage(camera.getFrame(1).getGsImage(), 1);
    long filled = 0;
    for (int y = 0; y < rows; ++y) for (int x = 0; x < cols; ++x) if (backprojection.at<cv::Vec3b>(y, x) != cv::Vec3b(0, 0, 0)) filled++;
    std::printf("rectified image: %ld of %d pixels filled\n", filled, rows * cols);

    // ---- accuracy metric of the sweep driver (main.cc:262-266): ground truth attached instead of loaded.
    // Scanline poses follow the same small-motion model with the true motion; world = scanline-0 frame.
    Eigen::MatrixXd ux = Eigen::MatrixXd::Zero(rows, cols), uy = ux, uz = ux;
    for (int j = 0; j < rows; ++j) {
        const double beta = gamma * j / rows;                         // constant velocity: k = 0
        Eigen::Matrix3d Rj;
        Rj(0, 0) = 1; Rj(0, 1) = -beta * w_true(2); Rj(0, 2) = beta * w_true(1);
        Rj(1, 0) = beta * w_true(2); Rj(1, 1) = 1; Rj(1, 2) = -beta * w_true(0);
        Rj(2, 0) = -beta * w_true(1); Rj(2, 1) = beta * w_true(0); Rj(2, 2) = 1;
        const Eigen::Vector3d tj = v_true * beta;
        camera.setScanlinePose(1, j, Rj, tj);
        for (int i = 0; i < cols; ++i) {
            const double x = (i - c_x) / f_x, y = (j - c_y) / f_y;
            const double d = 0.08 + 0.05 * std::sin(0.03 * i) * std::cos(0.02 * j) + 0.03 * x;
            const Eigen::Vector3d Xc(x / d, y / d, 1.0 / d);
            const Eigen::Vector3d Xw = Rj.transpose() * (Xc - tj);   // first-order inverse, like cameraToWorldFrame
            ux(j, i) = Xw(0); uy(j, i) = Xw(1); uz(j, i) = Xw(2);
        }
    }
    camera.setUnprojectionMaps(1, ux, uy, uz);
    // ground-truth flow towards frame 2 (camera.cc:209-249): frame 2 is read out one frame period later
    for (int j = 0; j < rows; ++j) {
        const double beta = 1.0 + gamma * j / rows;
        Eigen::Matrix3d Rj;
        Rj(0, 0) = 1; Rj(0, 1) = -beta * w_true(2); Rj(0, 2) = beta * w_true(1);
        Rj(1, 0) = beta * w_true(2); Rj(1, 1) = 1; Rj(1, 2) = -beta * w_true(0);
        Rj(2, 0) = -beta * w_true(1); Rj(2, 1) = beta * w_true(0); Rj(2, 2) = 1;
        camera.setScanlinePose(2, j, Rj, v_true * beta);
    }
    cv::Mat_<cv::Point_<double>> true_flow = camera.calculateTrueFlow(1, 2);
    double flow_dev = 0, flow_mag = 0;
    for (int j = 0; j < rows; ++j)
        for (int i = 0; i < cols; ++i) {
            flow_dev += std::fabs(true_flow(j, i).x - flow_image(j, i).x) + std::fabs(true_flow(j, i).y - flow_image(j, i).y);
            flow_mag += std::fabs(flow_image(j, i).x) + std::fabs(flow_image(j, i).y);
        }
    std::printf("ground-truth flow vs differential model: mean |difference| / mean |flow| = %.3e\n", flow_dev / flow_mag);
    const double mean_error = camera.meanReprojectionError(1);
    cv::Mat error_image = camera.createErrorImage(1, 1.0);
    Eigen::MatrixXd gt_depth = camera.getFrame(1).getGroundtruthDepthMap();
    std::printf("mean reprojection error %.4e (ground-truth depth at the centre %.3f, error image %dx%d)\n", mean_error,
                gt_depth(rows / 2, cols / 2), error_image.cols, error_image.rows);

    // ---- the same ground truth through the reference's fixture files (A.csv, N_rs_t.csv, N_rs_r.csv,
    // N_rs_unproject_{x,y,z}.csv; camera.cc:49-176, rsframe.cc:58-218, :444-553): written here, loaded back
    double mean_error_csv = -1.0;
    {
        const std::string dir = "/tmp/rsdsfm_example_";
        auto open = [&](const char *name) { std::ofstream f(dir + name); f << std::setprecision(17); return f; };
        {
            std::ofstream A = open("A.csv");
            for (int r = 0; r < 3; ++r) A << K(r, 0) << "," << K(r, 1) << "," << K(r, 2) << "\n";
            std::ofstream T = open("1_rs_t.csv"), Rf = open("1_rs_r.csv");
            for (int j = 0; j < rows; ++j) {
                const double beta = gamma * j / rows;
                T << v_true(0) * beta << "," << v_true(1) * beta << "," << v_true(2) * beta << "\n";
                const double Rj[9] = {1, -beta * w_true(2), beta * w_true(1), beta * w_true(2), 1, -beta * w_true(0),
                                      -beta * w_true(1), beta * w_true(0), 1};
                for (int a = 0; a < 9; ++a) Rf << Rj[a] << (a == 8 ? "\n" : ",");
            }
            std::ofstream X = open("1_rs_unproject_x.csv"), Y = open("1_rs_unproject_y.csv"), Z = open("1_rs_unproject_z.csv");
            for (int j = 0; j < rows; ++j)
                for (int i = 0; i < cols; ++i) {
                    const char *sep = (i == cols - 1) ? "\n" : ",";
                    X << ux(j, i) << sep; Y << uy(j, i) << sep; Z << uz(j, i) << sep;
                }
        }
        Camera camera2;
        const bool ok = camera2.loadIntrinsicsFromFile(dir + "A.csv", true);
        camera2.addFrameSynthetic(rs1, rs1, depth_map, dir + "1_rs_t.csv", dir + "1_rs_r.csv", dir + "1_rs_unproject_x.csv",
                                  dir + "1_rs_unproject_y.csv", dir + "1_rs_unproject_z.csv");
        camera2.setGamma(gamma);
        camera2.setPose(1, results.k, results.v, results.w);
        camera2.backProject(1);
        mean_error_csv = ok ? camera2.meanReprojectionError(1) : -1.0;
        camera2.createPointCloud(1, dir + "cloud.ply");                      // main.cc:552 (ASCII PLY, world frame)
        std::printf("mean reprojection error through the CSV fixtures %.4e\n", mean_error_csv);
    }

    double werr = 0;
    for (int a = 0; a < 3; ++a) werr = std::fmax(werr, std::fabs(results.w(a) - w_true(a)));
    const double cosang = results.v.dot(v_true) / (results.v.norm() * v_true.norm());
    std::printf("max |w - w_true| = %.3e, angle(v, v_true) = %.3e rad\n", werr, std::acos(std::fmin(1.0, cosang)));
    return (werr < 1e-6 && cosang > 1.0 - 1e-9 && filled > rows * cols * 9 / 10 && mean_error < 0.5 && mean_error_csv == mean_error) ? 0 : 1;
}
