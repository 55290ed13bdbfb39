// rsdsfm_host.h -- host-side mirror of the reference's C++ surface for the dense optimisation
// core (Camera, RsFrame, Scanline, minimal::*, nonlinear_refinement::*, error_measure::*).
// Same names, argument meaning and (absence of) error behaviour as the reference headers
//   minimal.h:41-160, nonlinearRefinement.h:42-111, camera.h:33-340, rsframe.h:33-315,
//   scanline.h:30-101, errorMeasure.h:18-64
// so that drivers written like main.cc:398-523 / errorMeasure.cpp:66-226 compile against it
// unchanged; every computation forwards to the sm_100a library through include/rsdsfm.h.
// Header-only; link with librsdsfm.so.  There is no CPU fallback: without a B200 the first call
// throws std::runtime_error carrying rsdsfm_last_error().
#pragma once

#include <cmath>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "rs_cv.h"
#include "rs_eigen.h"
#include "rsdsfm.h"

namespace rsdsfm_host {

// One context per thread (contexts are independent; a single one is not re-entrant).
inline rsdsfm_ctx *context()
{
    thread_local struct Holder {
        rsdsfm_ctx *ctx = nullptr;
        ~Holder() { if (ctx) rsdsfm_destroy(ctx); }
    } h;
    if (!h.ctx) {
        const char *dev = std::getenv("RSDSFM_DEVICE");
        if (rsdsfm_create(dev ? std::atoi(dev) : 0, nullptr, &h.ctx) != RSDSFM_OK)
            throw std::runtime_error(std::string("rsdsfm: ") + rsdsfm_last_error(nullptr));
    }
    return h.ctx;
}
inline void check(int rc, const char *what)
{
    if (rc != RSDSFM_OK) throw std::runtime_error(std::string(what) + ": " + rsdsfm_last_error(context()));
}

// The reference's fixture files are plain CSV: one matrix row per line, comma separated,
// fields converted with ::atof (camera.cc:137-160, rsframe.cc:444-553, :58-218).  Reads a whole file;
// false when it cannot be opened or does not have `lines` lines of `fields` fields (0: not checked).
inline bool readCsv(const std::string &path, long lines, int fields, std::vector<double> &values, const char *what, bool verbose = true)
{
    std::ifstream in(path);
    if (!in) {
        if (verbose) std::cout << path << " Is not a valid file path for the " << what << " file!" << std::endl;
        return false;
    }
    values.clear();
    std::string line, field;
    long n_lines = 0;
    bool fields_ok = true;
    while (std::getline(in, line)) {
        if (line.empty() && in.eof()) break;
        std::stringstream ss(line);
        int n = 0;
        while (std::getline(ss, field, ',')) { values.push_back(::atof(field.c_str())); ++n; }
        if (fields > 0 && n != fields) fields_ok = false;
        ++n_lines;
    }
    if ((lines > 0 && n_lines != lines) || !fields_ok) {
        if (verbose)
            std::cout << "The number of lines: " << n_lines << " in the file: " << path << " does not conform with the expected " << lines
                      << " lines of " << fields << " values!" << std::endl;
        return false;
    }
    return true;
}

}  // namespace rsdsfm_host

// ---------------------------------------------------------------------------- minimal.h:41-76
struct Velocities {
    Eigen::Vector3d w;
    Eigen::Vector3d v;
    double k;
    Velocities(Eigen::Vector3d w_init, Eigen::Vector3d v_init) : w(w_init), v(v_init), k(0) {}
    Velocities(Eigen::Vector3d w_init, Eigen::Vector3d v_init, double k_init) : w(w_init), v(v_init), k(k_init) {}
};

struct RansacValues {
    int num_inliers;
    Eigen::Array3Xd inliers;      // 3 x m: x, y, depth z (= 1 / inverse depth)
    Eigen::VectorXd beta;
    Eigen::VectorXd alpha;
    Eigen::VectorXd alpha_k;
    Eigen::Vector3d w;
    Eigen::Vector3d v;
    double k;
    RansacValues(int num_inliers_init, Eigen::Array3Xd inliers_init, Eigen::VectorXd beta_init, Eigen::Vector3d w_init,
                 Eigen::Vector3d v_init)
        : num_inliers(num_inliers_init), inliers(inliers_init), beta(beta_init), alpha(beta_init), alpha_k(beta_init),
          w(w_init), v(v_init), k(0) {}
    RansacValues(int num_inliers_init, Eigen::Array3Xd inliers_init, Eigen::VectorXd alpha_init, Eigen::VectorXd alpha_k_init,
                 Eigen::Vector3d w_init, Eigen::Vector3d v_init, double k_init)
        : num_inliers(num_inliers_init), inliers(inliers_init), beta(alpha_init), alpha(alpha_init), alpha_k(alpha_k_init),
          w(w_init), v(v_init), k(k_init) {}
};

// ---------------------------------------------------------------------------- nonlinearRefinement.h:77-111
namespace nonlinear_refinement {

inline Eigen::ArrayXd estimateInverseDepths(const Eigen::Array2Xd &normalized_coordinates, const Eigen::Array2Xd &flow,
                                            const Eigen::Vector3d &linear_velocity, const Eigen::Vector3d &angular_velocity,
                                            const double &k, const Eigen::ArrayXd &alpha, const Eigen::ArrayXd &alphaK,
                                            bool show_messages)
{
    const int n = (int)normalized_coordinates.cols();
    Eigen::ArrayXd out(n);
    rsdsfm_lm_summary s;
    rsdsfm_host::check(rsdsfm_estimate_inverse_depths(rsdsfm_host::context(), RSDSFM_HOST, normalized_coordinates.data(), flow.data(),
                                                      n, linear_velocity.data(), angular_velocity.data(), k, alpha.data(),
                                                      alphaK.data(), out.data(), &s),
                       "estimateInverseDepths");
    if (show_messages)
        std::cout << std::endl << "LM: iterations " << s.iterations << ", cost " << s.initial_cost << " -> " << s.final_cost
                  << ", " << s.device_ms * 1e-3 << " s on the GPU" << std::endl;
    return out;
}

inline double estimateInverseDepth(const Eigen::Vector2d &normalized_coordinates, const Eigen::Vector3d &linear_velocity,
                                   const Eigen::Vector3d &angular_velocity, const Eigen::Vector2d &flow, const double &k,
                                   const double &alpha, const double &alphaK, bool show_messages)
{
    Eigen::Array2Xd q(2, 1), u(2, 1);
    q(0, 0) = normalized_coordinates(0); q(1, 0) = normalized_coordinates(1);
    u(0, 0) = flow(0); u(1, 0) = flow(1);
    Eigen::ArrayXd a(1), ak(1);
    a(0) = alpha; ak(0) = alphaK;
    return estimateInverseDepths(q, u, linear_velocity, angular_velocity, k, a, ak, show_messages)(0);
}

// flow: the array the caller has (residual i reads flow(:, i), exactly like nonlinearRefinement.cc:209-212).
inline RansacValues nonLinearRefinement(const Eigen::Array2Xd &flow, const RansacValues &inliers, bool const_acceleration,
                                        bool show_messages)
{
    const int m = inliers.num_inliers;
    Eigen::Vector3d v = inliers.v, w = inliers.w;
    double k = inliers.k;
    std::vector<double> z((size_t)(m > 0 ? m : 1));
    rsdsfm_lm_summary s;
    rsdsfm_host::check(rsdsfm_refine(rsdsfm_host::context(), RSDSFM_HOST, flow.data(), inliers.inliers.data(), inliers.alpha.data(),
                                     inliers.alpha_k.data(), m, v.data(), w.data(), &k, const_acceleration ? 1 : 0, nullptr, nullptr,
                                     z.data(), &s),
                       "nonLinearRefinement");
    if (show_messages)
        std::cout << "LM: iterations " << s.iterations << ", cost " << s.initial_cost << " -> " << s.final_cost << ", "
                  << s.device_ms * 1e-3 << " s on the GPU" << std::endl << std::endl;
    Eigen::Array3Xd new_inliers = Eigen::Array3Xd::Zero(3, m);
    for (int i = 0; i < m; ++i) {
        new_inliers(0, i) = inliers.inliers(0, i);
        new_inliers(1, i) = inliers.inliers(1, i);
        new_inliers(2, i) = z[(size_t)i];
    }
    return RansacValues(m, new_inliers, inliers.alpha, inliers.alpha_k, w, v, k);
}

}  // namespace nonlinear_refinement

// ---------------------------------------------------------------------------- minimal.h:91-160
namespace minimal {

inline Velocities calculateVelocities(const Eigen::Array2Xd &q, const Eigen::Array2Xd &u, const Eigen::ArrayXd &alpha,
                                      const Eigen::ArrayXd &alpha_k, bool use_alpha_k)
{
    double out[7];
    rsdsfm_solve9(q.data(), u.data(), alpha.data(), alpha_k.data(), use_alpha_k ? 1 : 0, out);
    return Velocities(Eigen::Vector3d(out[0], out[1], out[2]), Eigen::Vector3d(out[3], out[4], out[5]), out[6]);
}

inline Eigen::ArrayXd getAlpha(const Eigen::Array2Xd &flow, double h, double gamma)
{
    const int n = (int)flow.cols();
    Eigen::ArrayXd alpha(n);
    rsdsfm_host::check(rsdsfm_alpha(rsdsfm_host::context(), RSDSFM_HOST, flow.data(), nullptr, n, h, gamma, alpha.data(), nullptr), "getAlpha");
    return alpha;
}

inline Eigen::ArrayXd getAlphaK(const Eigen::Array2Xd &q, const Eigen::Array2Xd &flow, double h, double gamma)
{
    const int n = (int)q.cols();
    Eigen::ArrayXd alpha_k(n);
    rsdsfm_host::check(rsdsfm_alpha(rsdsfm_host::context(), RSDSFM_HOST, flow.data(), q.data(), n, h, gamma, nullptr, alpha_k.data()), "getAlphaK");
    return alpha_k;
}

// minimal::ransac with the sample list made explicit (the reference draws it with srand(time)/rand).
inline RansacValues ransacWithSamples(const Eigen::Array2Xd &q, const Eigen::Array2Xd &u, const Eigen::ArrayXd &alpha,
                                      const Eigen::ArrayXd &alpha_k, bool use_alpha_k, const std::vector<int32_t> &samples,
                                      double tolerance, bool show_messages)
{
    const int n = (int)q.cols(), H = (int)(samples.size() / 9);
    std::vector<int> counts((size_t)H);
    std::vector<double> sumerr((size_t)H), invd((size_t)n);
    std::vector<uint8_t> mask((size_t)n);
    int best = -1;
    double best7[7];
    rsdsfm_ctx *ctx = rsdsfm_host::context();
    rsdsfm_host::check(rsdsfm_ransac(ctx, RSDSFM_HOST, q.data(), u.data(), alpha.data(), alpha_k.data(), n, use_alpha_k ? 1 : 0,
                                     samples.data(), H, tolerance, counts.data(), sumerr.data(), &best, best7, mask.data(),
                                     invd.data(), nullptr),
                       "ransac");
    if (show_messages)
        for (int i = 0, mx = -1; i < H; ++i) {
            if (counts[(size_t)i] > mx) mx = counts[(size_t)i];
            std::cout << "Finished " << i + 1 << " RANSAC trials. The current maximum number of inliers is " << mx << "." << std::endl;
        }
    std::vector<double> inl((size_t)3 * n + 3), a((size_t)n + 1), ak((size_t)n + 1);
    int m = 0;
    rsdsfm_host::check(rsdsfm_gather_inliers(ctx, RSDSFM_HOST, q.data(), alpha.data(), alpha_k.data(), n, mask.data(), invd.data(),
                                             inl.data(), a.data(), ak.data(), nullptr, &m),
                       "ransac (gather)");
    Eigen::Array3Xd pts = Eigen::Array3Xd::Zero(3, m);
    Eigen::VectorXd al(m), alk(m);
    for (int j = 0; j < m; ++j) {
        pts(0, j) = inl[(size_t)3 * j]; pts(1, j) = inl[(size_t)3 * j + 1]; pts(2, j) = inl[(size_t)3 * j + 2];
        al(j) = a[(size_t)j]; alk(j) = ak[(size_t)j];
    }
    return RansacValues(m, pts, al, alk, Eigen::Vector3d(best7[0], best7[1], best7[2]), Eigen::Vector3d(best7[3], best7[4], best7[5]),
                        best7[6]);
}

// The 9-subsets of minimal.cc:226-244.  The reference keeps ONE permutation of 0..n-1 alive across all
// trials and, per trial, moves 9 entries picked with rand() % (slots left) to the shrinking tail (Q5).
// Given the same rand() stream the picks below are the same indices in the same order.  The stream
// is seeded ONCE per call: the reference reseeds with time(NULL) before every trial, which only works
// there because a trial takes seconds -- here all trials are drawn within microseconds and a per-trial
// reseed would hand every trial the same nine values.
class SubsetDrawer {
public:
    explicit SubsetDrawer(int n) : perm_((size_t)n) { for (int i = 0; i < n; ++i) perm_[(size_t)i] = i; }
    // appends one trial (9 distinct point indices) to `out`; next_random() plays the role of rand()
    template <class Rng>
    void draw(Rng &&next_random, std::vector<int32_t> &out)
    {
        for (size_t left = perm_.size(), taken = 0; taken < 9; ++taken, --left) {
            const size_t slot = (size_t)next_random() % left;
            const int picked = perm_[slot];
            perm_[slot] = perm_[left - 1];
            perm_[left - 1] = picked;
            out.push_back(picked);
        }
    }

private:
    std::vector<int> perm_;
};

// iterations x 9 point indices for minimal::ransac, drawn from the C library generator
inline std::vector<int32_t> drawSamples(int n, int iterations)
{
    std::vector<int32_t> samples;
    samples.reserve((size_t)9 * (iterations > 0 ? iterations : 0));
    SubsetDrawer drawer(n);
    // RSDSFM_RANSAC_SEED (environment): a fixed seed instead of the reference's time(NULL), for reproducible runs and tests
    const char *fixed = getenv("RSDSFM_RANSAC_SEED");
    srand(fixed ? (unsigned)strtoul(fixed, nullptr, 10) : (unsigned)time(NULL));
    for (int trial = 0; trial < iterations; ++trial) drawer.draw([] { return rand(); }, samples);
    return samples;
}

inline RansacValues ransac(const Eigen::Array2Xd &q, const Eigen::Array2Xd &u, const Eigen::ArrayXd &alpha,
                           const Eigen::ArrayXd &alpha_k, bool use_alpha_k, int iterations, double tolerance, bool show_messages)
{
    return ransacWithSamples(q, u, alpha, alpha_k, use_alpha_k, drawSamples((int)q.cols(), iterations), tolerance, show_messages);
}
inline RansacValues ransac(const Eigen::Array2Xd &q, const Eigen::Array2Xd &u, const Eigen::ArrayXd &alpha, int iterations,
                           double tolerance, bool show_messages)
{
    return ransac(q, u, alpha, alpha, false, iterations, tolerance, show_messages);
}
inline RansacValues ransac(const Eigen::Array2Xd &q, const Eigen::Array2Xd &u, const Eigen::ArrayXd &alpha,
                           const Eigen::ArrayXd &alpha_k, int iterations, double tolerance, bool show_messages)
{
    return ransac(q, u, alpha, alpha_k, true, iterations, tolerance, show_messages);
}

}  // namespace minimal

// ---------------------------------------------------------------------------- scanline.h:30-101
class Scanline {
public:
    Scanline() {}
    Scanline(const Eigen::Matrix3d &rotation, const Eigen::Vector3d &translation) : rotation_(rotation), translation_(translation) {}
    const Eigen::Matrix3d &getRotation() const { return rotation_; }
    const Eigen::Matrix3d &getRelativeRotation() const { return relative_rotation_; }
    const Eigen::Vector3d &getTranslation() const { return translation_; }
    const Eigen::Vector3d &getRelativeTranslation() const { return relative_translation_; }
    void setRotation(const Eigen::Matrix3d &rotation) { rotation_ = rotation; }
    void setTranslation(const Eigen::Vector3d &translation) { translation_ = translation; }
    void setRelativeRotation(const Eigen::Matrix3d &rotation) { relative_rotation_ = rotation; }
    void setRelativeTranslation(const Eigen::Vector3d &translation) { relative_translation_ = translation; }

private:
    Eigen::Matrix3d rotation_;
    Eigen::Vector3d translation_;
    Eigen::Matrix3d relative_rotation_;
    Eigen::Vector3d relative_translation_;
};

// ---------------------------------------------------------------------------- rsframe.h:33-315 (hot-path slice)
class RsFrame {
public:
    void setImage(cv::Mat image)
    {   // rsframe.cc:32-40
        rows_ = image.rows; cols_ = image.cols; image_ = image;
        for (int i = 0; i < rows_; ++i) scanlines_.push_back(Scanline());
    }
    void setDepthMap(const Eigen::MatrixXd &depth_map) { depth_map_ = depth_map; }
    void setGsImage(cv::Mat image) { gs_image_ = image; }
    void setGamma(const double gamma) { gamma_ = gamma; }
    void setIntrinsics(const Eigen::Matrix3d &intrinsics)
    {   // rsframe.cc:556-561
        f_x_ = intrinsics(0, 0); f_y_ = intrinsics(1, 1); c_x_ = intrinsics(0, 2); c_y_ = intrinsics(1, 2);
    }
    int getRows() const { return rows_; }
    int getCols() const { return cols_; }
    cv::Mat getRsImage() { return image_; }
    cv::Mat getGsImage() { return gs_image_; }
    Eigen::MatrixXd getDepthMap() { return depth_map_; }
    cv::Mat get3dCoordinates() { return coordinates_3d_; }
    unsigned long getNrScannlines() const { return scanlines_.size(); }

    // rsframe.cc:771-800
    void setRelativePose(const Eigen::Vector3d &linear_velocity, const Eigen::Vector3d &angular_velocity, const double k)
    {
        std::vector<double> R((size_t)9 * rows_), t((size_t)3 * rows_);
        rsdsfm_set_relative_pose(linear_velocity.data(), angular_velocity.data(), k, gamma_, rows_, R.data(), t.data());
        for (int i = 0; i < rows_; ++i) {
            Eigen::Matrix3d Ri;
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Ri(r, c) = R[(size_t)9 * i + 3 * r + c];
            scanlines_[(size_t)i].setRelativeRotation(Ri);
            scanlines_[(size_t)i].setRelativeTranslation(Eigen::Vector3d(t[(size_t)3 * i], t[(size_t)3 * i + 1], t[(size_t)3 * i + 2]));
        }
    }

    // rsframe.cc:629-736 -- scalar helpers kept on the host (they are not on the per-pixel path here)
    Eigen::Vector2d spaceToPlane(const Eigen::Vector3d &Point)
    {
        return Eigen::Vector2d(Point.x() / Point.z() * f_x_ + c_x_, Point.y() / Point.z() * f_x_ + c_y_);   // y uses f_x (Q12)
    }
    Eigen::Vector3d planeToSpace(const Eigen::Vector2d &point, double z_value = 0)
    {
        if (z_value == 0) z_value = depth_map_.coeff(int(point.y()), int(point.x()));
        return Eigen::Vector3d((point.x() - c_x_) * 1.0 / f_x_, (point.y() - c_y_) * 1.0 / f_y_, 1.0) * z_value;
    }
    Eigen::Vector3d worldToCameraFrame(const Eigen::Vector3d &Point, const int scanlineNr, bool useRelative = true)
    {
        const Scanline &s = scanlines_[(size_t)scanlineNr];
        const Eigen::Matrix3d &R = useRelative ? s.getRelativeRotation() : s.getRotation();
        const Eigen::Vector3d &t = useRelative ? s.getRelativeTranslation() : s.getTranslation();
        return R * Point + t;
    }
    Eigen::Vector3d cameraToWorldFrame(const Eigen::Vector3d &Point, const int scanlineNr, bool useRelative = true)
    {
        const Scanline &s = scanlines_[(size_t)scanlineNr];
        const Eigen::Matrix3d Rt = (useRelative ? s.getRelativeRotation() : s.getRotation()).transpose();
        const Eigen::Vector3d &t = useRelative ? s.getRelativeTranslation() : s.getTranslation();
        return Rt * Point - Rt * t;
    }

    // rsframe.cc:803-839 / :842-878
    void backProject() { backProjectImpl(0); }
    void backProjectGs() { backProjectImpl(1); }

    // ---- ground truth of a synthetic frame.  The reference fills these members from
    // the fixture CSVs (rsframe.cc:222-378, :444-553); the loaders are upstream of this path, so the
    // values are attached instead.
    void setUnprojectionMaps(const Eigen::MatrixXd &x, const Eigen::MatrixXd &y, const Eigen::MatrixXd &z)
    {
        unprojection_map_x_ = x; unprojection_map_y_ = y; unprojection_map_z_ = z;
    }
    void setScanlinePose(const int scanlineNr, const Eigen::Matrix3d &rotation, const Eigen::Vector3d &translation)
    {
        // the fixture loaders initialise the absolute and the relative pose alike (rsframe.cc:505-540)
        scanlines_[(size_t)scanlineNr].setRotation(rotation);
        scanlines_[(size_t)scanlineNr].setTranslation(translation);
        scanlines_[(size_t)scanlineNr].setRelativeRotation(rotation);
        scanlines_[(size_t)scanlineNr].setRelativeTranslation(translation);
    }
    // rsframe.cc:444-553: N_rs_t.csv (rows x 3) and N_rs_r.csv (rows x 9, row-major 3x3); both the absolute
    // and the relative pose of every scanline are initialised.  Nothing is set unless both files fit.
    bool setPoses(std::string csv_poses, std::string csv_orientation)
    {
        std::vector<double> t, R;
        if (!rsdsfm_host::readCsv(csv_poses, rows_, 3, t, "poses") || !rsdsfm_host::readCsv(csv_orientation, rows_, 9, R, "orientation")) {
            std::cout << "NO poses and NO orientations were set" << std::endl;
            return false;
        }
        for (int i = 0; i < rows_; ++i) {
            Eigen::Matrix3d Ri;
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Ri(r, c) = R[(size_t)9 * i + 3 * r + c];
            setScanlinePose(i, Ri, Eigen::Vector3d(t[(size_t)3 * i], t[(size_t)3 * i + 1], t[(size_t)3 * i + 2]));
        }
        return true;
    }
    // rsframe.cc:58-218: N_rs_unproject_{x,y,z}.csv, rows lines of cols values each
    bool setUnprojectionMapRs(const std::string csv_unprojection_x, const std::string csv_unprojection_y,
                              const std::string csv_unprojection_z)
    {
        std::vector<double> v[3];
        const std::string *paths[3] = {&csv_unprojection_x, &csv_unprojection_y, &csv_unprojection_z};
        for (int a = 0; a < 3; ++a)
            if (!rsdsfm_host::readCsv(*paths[a], rows_, cols_, v[a], "unprojection map")) {
                std::cout << "NO unprojection map was set" << std::endl;
                return false;
            }
        Eigen::MatrixXd m[3] = {Eigen::MatrixXd::Zero(rows_, cols_), Eigen::MatrixXd::Zero(rows_, cols_), Eigen::MatrixXd::Zero(rows_, cols_)};
        for (int a = 0; a < 3; ++a)
            for (int y = 0; y < rows_; ++y) for (int x = 0; x < cols_; ++x) m[a](y, x) = v[a][(size_t)y * cols_ + x];
        setUnprojectionMaps(m[0], m[1], m[2]);
        return true;
    }
    // rsframe.cc:222-378: N_gs_unproject_{x,y,z}.csv -- the world points seen by the global-shutter frame
    bool setUnprojectionMapGs(const std::string csv_unprojection_x, const std::string csv_unprojection_y,
                              const std::string csv_unprojection_z)
    {
        std::vector<double> v[3];
        const std::string *paths[3] = {&csv_unprojection_x, &csv_unprojection_y, &csv_unprojection_z};
        for (int a = 0; a < 3; ++a)
            if (!rsdsfm_host::readCsv(*paths[a], rows_, cols_, v[a], "GS unprojection map")) {
                std::cout << "Unprojection maps for GS image not set" << std::endl;
                return false;
            }
        Eigen::MatrixXd *dst[3] = {&gs_unprojection_map_x_, &gs_unprojection_map_y_, &gs_unprojection_map_z_};
        for (int a = 0; a < 3; ++a) {
            *dst[a] = Eigen::MatrixXd::Zero(rows_, cols_);
            for (int y = 0; y < rows_; ++y) for (int x = 0; x < cols_; ++x) (*dst[a])(y, x) = v[a][(size_t)y * cols_ + x];
        }
        return true;
    }
    // rsframe.cc:565-586 / :589-614: depth of every pixel's world point in the camera frame -- of scanline 0 for the GS
    // frame, of the pixel's own scanline for the RS frame; pixels without a world point (all-zero entry) get depth 0.
    void setSyntheticDepthMapGs()
    {
        gs_depth_map_ = Eigen::MatrixXd::Zero(rows_, cols_);
        for (int y = 0; y < (int)scanlines_.size(); ++y)
            for (int x = 0; x < cols_; ++x) {
                const Eigen::Vector3d Pw(gs_unprojection_map_x_(y, x), gs_unprojection_map_y_(y, x), gs_unprojection_map_z_(y, x));
                gs_depth_map_(y, x) = (Pw.norm() > 0) ? worldToCameraFrame(Pw, 0).z() : 0.0;
            }
    }
    void setSyntheticDepthMapRs()
    {
        depth_map_ = Eigen::MatrixXd::Zero(rows_, cols_);
        double z_sum = 0;
        long z_count = 0;
        for (int y = 0; y < (int)scanlines_.size(); ++y)
            for (int x = 0; x < cols_; ++x) {
                const Eigen::Vector3d Pw(unprojection_map_x_(y, x), unprojection_map_y_(y, x), unprojection_map_z_(y, x));
                if (Pw.norm() > 0) { const double z = worldToCameraFrame(Pw, y).z(); depth_map_(y, x) = z; z_sum += z; ++z_count; }
            }
        std::cout << "z mean: " << z_sum * 1.0 / z_count << std::endl;
    }
    Eigen::MatrixXd getGsDepthMap() { return gs_depth_map_; }
    // rsframe.cc:617-625
    Eigen::Vector3d getUnprojectedWorldCoordinates(const Eigen::Vector2d &point)
    {
        const int r = (int)point.y(), c = (int)point.x();
        return Eigen::Vector3d(unprojection_map_x_(r, c), unprojection_map_y_(r, c), unprojection_map_z_(r, c));
    }
    bool hasUnprojectionMaps() const { return unprojection_map_x_.rows() == rows_ && unprojection_map_x_.cols() == cols_ && rows_ > 0; }
    // camera.cc:209-249 seen from frame 1: flow towards `frame2` (its relative scanline poses, rsframe.h:239)
    cv::Mat_<cv::Point_<double>> trueFlowTo(const RsFrame &frame2)
    {
        std::vector<double> R((size_t)9 * rows_), t((size_t)3 * rows_);
        for (int i = 0; i < rows_; ++i) {
            const Eigen::Matrix3d &Ri = frame2.scanlines_[(size_t)i].getRelativeRotation();
            const Eigen::Vector3d &ti = frame2.scanlines_[(size_t)i].getRelativeTranslation();
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[(size_t)9 * i + 3 * r + c] = Ri(r, c);
            for (int a = 0; a < 3; ++a) t[(size_t)3 * i + a] = ti(a);
        }
        const double K4[4] = {f_x_, f_y_, c_x_, c_y_};
        cv::Mat_<cv::Point_<double>> flow(rows_, cols_);
        rsdsfm_host::check(rsdsfm_true_flow(rsdsfm_host::context(), RSDSFM_HOST, unprojection_map_x_.data(), unprojection_map_y_.data(),
                                            unprojection_map_z_.data(), R.data(), t.data(), RSDSFM_DEPTH_COLMAJOR, rows_, cols_, K4,
                                            reinterpret_cast<double *>(flow.data)),
                           "calculateTrueFlow");
        return flow;
    }
    // rsframe.cc:416-436
    Eigen::MatrixXd getGroundtruthDepthMap()
    {
        Eigen::MatrixXd out = Eigen::MatrixXd::Zero(rows_, cols_);
        double mean_error = 0;
        std::vector<float> no_estimate((size_t)3 * rows_ * cols_, 0.f);
        reprojection(no_estimate.data(), 1.0, &mean_error, nullptr, out.data());
        return out;
    }
    // rsframe.cc:953-967
    void relocatePose()
    {
        std::vector<double> R, t;
        packPoses(R, t);
        std::vector<double> Ro(R.size()), to(t.size());
        rsdsfm_host::check(rsdsfm_relocate_pose(R.data(), t.data(), rows_, Ro.data(), to.data()), "relocatePose");
        for (int i = 0; i < rows_; ++i) {
            Eigen::Matrix3d Ri; Eigen::Vector3d ti;
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Ri(r, c) = Ro[(size_t)9 * i + 3 * r + c];
            for (int a = 0; a < 3; ++a) ti(a) = to[(size_t)3 * i + a];
            scanlines_[(size_t)i].setRotation(Ri);
            scanlines_[(size_t)i].setTranslation(ti);
        }
    }
    // the per-pixel map-reduce behind Camera::meanReprojectionError / createErrorImage (camera.cc:503-691)
    void reprojection(const float *coords3d, double max_norm, double *mean_error, unsigned char *error_image, double *gt_depth)
    {
        std::vector<double> R, t;
        packPoses(R, t);
        const double K4[4] = {f_x_, f_y_, c_x_, c_y_};
        Eigen::MatrixXd depth = depth_map_;
        if (depth.rows() != rows_ || depth.cols() != cols_) depth = Eigen::MatrixXd::Zero(rows_, cols_);
        rsdsfm_host::check(rsdsfm_reprojection_error(rsdsfm_host::context(), RSDSFM_HOST, coords3d, unprojection_map_x_.data(),
                                                     unprojection_map_y_.data(), unprojection_map_z_.data(), R.data(), t.data(),
                                                     depth.data(), RSDSFM_DEPTH_COLMAJOR, rows_, cols_, K4, max_norm, mean_error,
                                                     nullptr, nullptr, nullptr, error_image, gt_depth),
                           "meanReprojectionError");
    }

private:
    void packPoses(std::vector<double> &R, std::vector<double> &t) const
    {
        R.resize((size_t)9 * rows_); t.resize((size_t)3 * rows_);
        for (int i = 0; i < rows_; ++i) {
            const Eigen::Matrix3d &Ri = scanlines_[(size_t)i].getRotation();
            const Eigen::Vector3d &ti = scanlines_[(size_t)i].getTranslation();
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[(size_t)9 * i + 3 * r + c] = Ri(r, c);
            for (int a = 0; a < 3; ++a) t[(size_t)3 * i + a] = ti(a);
        }
    }
    void backProjectImpl(int gs_mode)
    {
        std::vector<double> R((size_t)9 * rows_), t((size_t)3 * rows_);
        for (int i = 0; i < rows_; ++i) {
            const Eigen::Matrix3d &Ri = scanlines_[(size_t)i].getRelativeRotation();
            const Eigen::Vector3d &ti = scanlines_[(size_t)i].getRelativeTranslation();
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[(size_t)9 * i + 3 * r + c] = Ri(r, c);
            for (int a = 0; a < 3; ++a) t[(size_t)3 * i + a] = ti(a);
        }
        const double K4[4] = {f_x_, f_y_, c_x_, c_y_};
        cv::Mat gs_image = image_.clone();
        cv::Mat coordinates_3d(rows_, cols_, CV_32FC3);
        rsdsfm_host::check(rsdsfm_backproject(rsdsfm_host::context(), RSDSFM_HOST, image_.data, depth_map_.data(), RSDSFM_DEPTH_COLMAJOR,
                                              rows_, cols_, K4, R.data(), t.data(), gs_mode, gs_image.data,
                                              reinterpret_cast<float *>(coordinates_3d.data)),
                           "backProject");
        coordinates_3d_ = coordinates_3d;       // rsframe.cc:837-838: both members are replaced (Q20)
        gs_image_ = gs_image;
    }

    int rows_ = 0, cols_ = 0;
    double f_x_ = 0, f_y_ = 0, c_x_ = 0, c_y_ = 0, gamma_ = 0;
    cv::Mat image_, gs_image_, coordinates_3d_;
    Eigen::MatrixXd depth_map_;
    Eigen::MatrixXd unprojection_map_x_, unprojection_map_y_, unprojection_map_z_;
    Eigen::MatrixXd gs_unprojection_map_x_, gs_unprojection_map_y_, gs_unprojection_map_z_, gs_depth_map_;
    std::vector<Scanline> scanlines_;
};

// ---------------------------------------------------------------------------- camera.h:33-340 (hot-path slice)
class Camera {
public:
    Eigen::Matrix3d getIntrinsics() { return K_; }
    void addFrameReal(cv::Mat rs_image)
    {   // camera.cc:39-46
        RsFrame frame;
        frame.setIntrinsics(K_);
        frame.setImage(rs_image);
        frames_.push_back(frame);
    }
    void setIntrinsics(const Eigen::Matrix3d &intrinsics) { K_ = intrinsics; }
    // camera.cc:179-206: the five hard-coded phone calibrations
    void setIntrinsics(const std::string source_camera)
    {
        double fx = 0, fy = 0, cx = 0, cy = 0;
        if (source_camera == "iphone") { fx = 1505.1283359786307; fy = 1513.7789208311444; cx = 657.81734686405991; cy = 349.91807538147589; }
        else if (source_camera == "galaxy_stabil") { fx = 1803.29785922382; fy = 1799.35406531529; cx = 945.304708272490; cy = 544.684292978344; }
        else if (source_camera == "galaxy") { fx = 1492.41306997746; fy = 1491.09286590722; cx = 949.571146410704; cy = 554.675409391795; }
        else if (source_camera == "galaxy_old") { fx = 3154.53208221173; fy = 3152.28696217577; cx = 1969.87107268891; cy = 1521.27056048818; }
        else if (source_camera == "galaxy_vga") { fx = 484.450845764569; fy = 485.345469134313; cx = 313.442094604855; cy = 241.383116350144; }
        else std::cerr << "No valid source camera specified";
        Eigen::Matrix3d K;
        K(0, 0) = fx; K(1, 1) = fy; K(0, 2) = cx; K(1, 2) = cy; K(2, 2) = 1.0;
        K_ = K;
    }
    RsFrame getFrame(const int frameNr) { return frames_[(size_t)frameNr - 1]; }          // 1-based, deep copy
    void setPose(const int frameNr, const double k, const Eigen::Vector3d &linear_velocity, const Eigen::Vector3d &angular_velocity)
    {
        frames_[(size_t)frameNr - 1].setRelativePose(linear_velocity, angular_velocity, k);
    }
    void setGamma(const double gamma) { for (auto &f : frames_) f.setGamma(gamma); }
    void backProject(const int frameNr) { frames_[(size_t)frameNr - 1].backProject(); }
    void backProjectGs(const int frameNr) { frames_[(size_t)frameNr - 1].backProjectGs(); }
    void setDepthMap(const int frameNr, Eigen::MatrixXd depth_map) { frames_[(size_t)frameNr - 1].setDepthMap(depth_map); }

    // camera.cc:253-277 computes DeepFlow with OpenCV's optflow module; flow production is upstream
    // of this path (both sides consume a cached flow field), so the flow is attached instead.
    void setCachedFlow(const cv::Mat_<cv::Point_<double>> &flow) { cached_flow_ = flow; }
    cv::Mat_<cv::Point_<double>> calculateDeepFlow(const int, const int) { return cached_flow_; }
    // camera.cc:209-249: from the attached ground truth when there is one, else the cached field
    cv::Mat_<cv::Point_<double>> calculateTrueFlow(const int frameNr1, const int frameNr2)
    {
        RsFrame &f1 = frames_[(size_t)frameNr1 - 1];
        if (!f1.hasUnprojectionMaps()) return cached_flow_;
        return f1.trueFlowTo(frames_[(size_t)frameNr2 - 1]);
    }

    // camera.cc:423-491: ASCII PLY of the back-projected 3D points (world frame) coloured by the RS image.
    // File format (fixed by the reference's output): the 11 header lines below, then one vertex per pixel
    // in raster order as "x y z r g b": coordinates with 9 significant digits, each followed by a blank,
    // colour as decimal RGB (the image is BGR) separated by single blanks.
    void createPointCloud(const int frameNr, const std::string fileName)
    {
        RsFrame frame = frames_[(size_t)frameNr - 1];
        const cv::Mat xyz = frame.get3dCoordinates(), bgr = frame.getRsImage();
        const size_t vertices = (size_t)frame.getRows() * (size_t)frame.getCols();
        if ((size_t)xyz.rows * (size_t)xyz.cols != vertices) throw std::runtime_error("createPointCloud: backProject has not run");
        static const char *const properties[] = {"float x", "float y", "float z", "uchar red", "uchar green", "uchar blue"};
        std::ofstream ply(fileName);
        ply << "ply\nformat ascii 1.0\ncomment PLY File created by RS aware SfM wrapper\nelement vertex " << vertices << "\n";
        for (const char *p : properties) ply << "property " << p << "\n";
        ply << "end_header" << std::endl;
        ply << std::setprecision(9);
        const float *point = reinterpret_cast<const float *>(xyz.data);
        const unsigned char *colour = bgr.data;
        for (size_t vtx = 0; vtx < vertices; ++vtx, point += 3, colour += 3)
            ply << point[0] << ' ' << point[1] << ' ' << point[2] << ' ' << (unsigned)colour[2] << ' ' << (unsigned)colour[1] << ' '
                << (unsigned)colour[0] << '\n';
        ply.close();
        std::cout << "Point cloud file " << fileName << " created." << std::endl;
    }
    // camera.cc:280-308: the flow field as colours -- hue = direction (degrees), value = magnitude / largest magnitude,
    // full saturation -- returned as a float BGR image in [0, 1] (main.cc:390-392 scales it by 255 and stores it).
    // The reference goes through cv::cartToPolar / cv::cvtColor; this is the same colour model in plain arithmetic
    // (OpenCV's fast arctangent is accurate to 0.3 degrees, so the last grey level may differ).
    cv::Mat getImageOpticalFlow(cv::Mat_<cv::Point_<double>> flow)
    {
        const int rows = flow.rows, cols = flow.cols;
        cv::Mat bgr(rows, cols, CV_32FC3);
        float mag_max = 0.f;
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) {
                const float fx = (float)flow(y, x).x, fy = (float)flow(y, x).y;
                mag_max = std::max(mag_max, std::sqrt(fx * fx + fy * fy));
            }
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) {
                const float fx = (float)flow(y, x).x, fy = (float)flow(y, x).y;
                float hue = std::atan2(fy, fx) * 57.29577951308232f;
                if (hue < 0.f) hue += 360.f;
                const float val = mag_max > 0.f ? std::sqrt(fx * fx + fy * fy) / mag_max : 0.f;
                // HSV -> BGR with S = 1 (cv::COLOR_HSV2BGR on float images: H in [0, 360), S and V in [0, 1])
                const float h6 = hue / 60.f;
                const int sector = ((int)std::floor(h6)) % 6;
                const float f = h6 - std::floor(h6), p = 0.f, q = val * (1.f - f), t = val * f;
                float r, g, b;
                switch (sector) {
                    case 0: r = val; g = t; b = p; break;
                    case 1: r = q; g = val; b = p; break;
                    case 2: r = p; g = val; b = t; break;
                    case 3: r = p; g = q; b = val; break;
                    case 4: r = t; g = p; b = val; break;
                    default: r = val; g = p; b = q; break;
                }
                bgr.at<cv::Vec3f>(y, x) = cv::Vec3f(b, g, r);
            }
        return bgr;
    }
    // camera.cc:311-332: every delta-th pixel with a non-zero flow gets a red arrow from the pixel to where the flow
    // (truncated to whole pixels) takes it.  Drawn with a plain integer line and a two-stroke head; the reference
    // draws with cv::arrowedLine (anti-aliased), so pixel values along the strokes differ, their geometry does not.
    cv::Mat flowArrows(const cv::Mat image, const cv::Mat_<cv::Point_<double>> flow, const int delta_x, const int delta_y)
    {
        cv::Mat out = image;
        auto put = [&](int x, int y) { if (x >= 0 && y >= 0 && x < out.cols && y < out.rows) out.at<cv::Vec3b>(y, x) = cv::Vec3b(0, 0, 255); };
        auto line = [&](int x0, int y0, int x1, int y1) {
            const int dx = std::abs(x1 - x0), dy = -std::abs(y1 - y0), sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
            for (int e = dx + dy;;) {
                put(x0, y0);
                if (x0 == x1 && y0 == y1) break;
                const int e2 = 2 * e;
                if (e2 >= dy) { e += dy; x0 += sx; }
                if (e2 <= dx) { e += dx; y0 += sy; }
            }
        };
        for (int y = 0; y < image.rows; y += delta_y)
            for (int x = 0; x < image.cols; x += delta_x) {
                const double dx = flow(y, x).x, dy = flow(y, x).y;
                const double l = std::sqrt(dx * dx + dy * dy);
                if (!(l > 0)) continue;
                const int x2 = x + (int)dx, y2 = y + (int)dy;
                line(x, y, x2, y2);
                // head: two strokes of a tenth of the arrow's length at +-45 degrees from the reversed direction (cv::arrowedLine, tipLength 0.1)
                const double tip = 0.1 * std::sqrt((double)((x2 - x) * (x2 - x) + (y2 - y) * (y2 - y)));
                const double ang = std::atan2((double)(y - y2), (double)(x - x2));
                for (int sgn = -1; sgn <= 1; sgn += 2)
                    line(x2, y2, (int)std::lround(x2 + tip * std::cos(ang + sgn * 0.7853981633974483)),
                         (int)std::lround(y2 + tip * std::sin(ang + sgn * 0.7853981633974483)));
            }
        return out;
    }
    // camera.cc:99-176: A.csv, the 3 x 3 intrinsic matrix
    bool loadIntrinsicsFromFile(const std::string csv_intrinsic_matrix, bool show_messages)
    {
        std::vector<double> a;
        if (!rsdsfm_host::readCsv(csv_intrinsic_matrix, 3, 3, a, "intrinsic", show_messages)) {
            if (show_messages) std::cout << "No itrinsics were set" << std::endl;
            return false;
        }
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) K_(r, c) = a[(size_t)3 * r + c];
        return true;
    }
    // camera.cc:49-70: a synthetic frame with its ground truth (the images are decoded by the caller)
    void addFrameSynthetic(cv::Mat rs_image, cv::Mat gs_image, const Eigen::MatrixXd &depth_map, const std::string poses_csv,
                           const std::string orientation_csv, const std::string csv_unproject_x, const std::string csv_unproject_y,
                           const std::string csv_unproject_z)
    {
        RsFrame frame;
        frame.setIntrinsics(K_);
        frame.setImage(rs_image);
        frame.setGsImage(gs_image);
        frame.setDepthMap(depth_map);
        frame.setPoses(poses_csv, orientation_csv);
        frame.setUnprojectionMapRs(csv_unproject_x, csv_unproject_y, csv_unproject_z);
        frames_.push_back(frame);
    }
    // ground-truth attachments without files
    void setUnprojectionMaps(const int frameNr, const Eigen::MatrixXd &x, const Eigen::MatrixXd &y, const Eigen::MatrixXd &z)
    {
        frames_[(size_t)frameNr - 1].setUnprojectionMaps(x, y, z);
    }
    void setScanlinePose(const int frameNr, const int scanlineNr, const Eigen::Matrix3d &rotation, const Eigen::Vector3d &translation)
    {
        frames_[(size_t)frameNr - 1].setScanlinePose(scanlineNr, rotation, translation);
    }
    // camera.cc:593-691 -- works on a copy of the frame, like the reference
    double meanReprojectionError(const int frameNr)
    {
        RsFrame frame = frames_[(size_t)frameNr - 1];
        cv::Mat coords = frame.get3dCoordinates();
        double mean_error = 0;
        frame.reprojection(reinterpret_cast<const float *>(coords.data), 1.0, &mean_error, nullptr, nullptr);
        return mean_error;
    }
    // camera.cc:503-590
    cv::Mat createErrorImage(const int frameNr, const double max_norm)
    {
        RsFrame frame = frames_[(size_t)frameNr - 1];
        cv::Mat coords = frame.get3dCoordinates();
        cv::Mat error_image(frame.getRows(), frame.getCols(), CV_8UC1, cv::Scalar(0));
        double mean_error = 0;
        frame.reprojection(reinterpret_cast<const float *>(coords.data), max_norm, &mean_error, error_image.data, nullptr);
        return error_image;
    }

    // camera.cc:753-774
    cv::Mat interpolateCrackyImage(cv::Mat image_in, const unsigned offset)
    {
        cv::Mat image_out = image_in.clone();
        rsdsfm_host::check(rsdsfm_fill_cracks(rsdsfm_host::context(), RSDSFM_HOST, image_in.data, image_in.rows, image_in.cols, offset,
                                              image_out.data),
                           "interpolateCrackyImage");
        return image_out;
    }

private:
    std::vector<RsFrame> frames_;
    Eigen::Matrix3d K_;
    cv::Mat_<cv::Point_<double>> cached_flow_;
};

// ---------------------------------------------------------------------------- errorMeasure.h:18-64
namespace error_measure {

// errorMeasure.h:18-24 (constructed as TrueValues(w, v), main.cc:245)
struct TrueValues {
    Eigen::Vector3d w;
    Eigen::Vector3d v;
    TrueValues(Eigen::Vector3d w_init, Eigen::Vector3d v_init) : w(w_init), v(v_init) {}
};

// errorMeasure.h:29-44: same members, same constructor (read by main.cc:275-283 as .col(j) / (j)).
// The reference hands its per-evaluation scalar errors (ArrayXd) to the Array3Xd members (Q23, a shape
// bug: only column 0 is defined there).  evaluateVelocities below fills well-formed 3 x num_evaluations
// arrays instead: column j = (error of evaluation j, 0, 0).
struct VelocityErrors {
    Eigen::Array3Xd w;
    Eigen::Array3Xd v;
    Eigen::ArrayXd k;
    Eigen::ArrayXd error_reproject_vec;
    Eigen::Array3Xd error_v_vec;
    Eigen::Array3Xd error_w_vec;
    double error_w;
    double error_v;
    double error_reproject;
    VelocityErrors(Eigen::Array3Xd w_init, Eigen::Array3Xd v_init, Eigen::ArrayXd k_init, Eigen::ArrayXd error_reproject_vec_init,
                   Eigen::Array3Xd error_v_vec_init, Eigen::Array3Xd error_w_vec_init, double error_w_init, double error_v_init,
                   double error_reproject_init)
        : w(w_init), v(v_init), k(k_init), error_reproject_vec(error_reproject_vec_init), error_v_vec(error_v_vec_init),
          error_w_vec(error_w_vec_init), error_w(error_w_init), error_v(error_v_init), error_reproject(error_reproject_init) {}
};

// The flatten + normalise glue both reference drivers share (main.cc:398-432, errorMeasure.cpp:66-97).
struct Flattened {
    Eigen::Matrix2Xd coord, flow, coord_pixel, flow_pixel;
    int n = 0;
};
inline Flattened flattenFlow(const cv::Mat_<cv::Point_<double>> &flow_image, const Eigen::Matrix3d &K, double gamma,
                             double flow_threshold, bool truncate)
{
    const int rows = flow_image.rows, cols = flow_image.cols;
    Flattened F;
    F.coord = Eigen::Matrix2Xd(2, (std::ptrdiff_t)rows * cols); F.flow = F.coord; F.coord_pixel = F.coord; F.flow_pixel = F.coord;
    const double K4[4] = {K(0, 0), K(1, 1), K(0, 2), K(1, 2)};
    rsdsfm_host::check(rsdsfm_flatten(rsdsfm_host::context(), RSDSFM_HOST, reinterpret_cast<const double *>(flow_image.data), rows, cols,
                                      K4, gamma, flow_threshold, F.coord.data(), F.flow.data(), F.coord_pixel.data(),
                                      F.flow_pixel.data(), nullptr, &F.n),
                       "flatten");
    if (truncate) { F.coord.conservativeResize(2, F.n); F.flow.conservativeResize(2, F.n); }   // errorMeasure.cpp:96-97
    return F;
}

// errorMeasure.cpp:41-254.  image_path: prefix of the per-evaluation artefacts "<prefix><eval>.png" (8-bit depth
// image, :191-206) and "<prefix><eval>.ply" (:229-230); an EMPTY prefix writes no files (the reference has no
// such switch: it always writes).  The reprojection error (:229) needs the frame's ground truth (unprojection
// maps and scanline poses, attached or loaded from the fixture CSVs); without it the entry is NaN.
inline VelocityErrors evaluateVelocities(Camera camera, TrueValues true_values, double gamma, int ransac_trials, int num_evaluations,
                                         bool use_deep_flow, bool constant_acceleration, bool global_shutter,
                                         bool optimize_results, bool show_messages, std::string image_path)
{
    const double THRESHOLD_FLOW = 0.0000000001, TOL_RANSAC = 0.05;
    cv::Mat_<cv::Point_<double>> flow_image = use_deep_flow ? camera.calculateDeepFlow(1, 2) : camera.calculateTrueFlow(1, 2);
    const int rows = flow_image.rows, cols = flow_image.cols;
    Eigen::Matrix3d K = camera.getIntrinsics();
    const double f_x = K(0, 0), f_y = K(1, 1), c_x = K(0, 2), c_y = K(1, 2);
    camera.setGamma(gamma);
    Flattened F = flattenFlow(flow_image, K, gamma, THRESHOLD_FLOW, true);
    // alpha factors from the un-truncated pixel-unit arrays (errorMeasure.cpp:104-105); entries >= n are never read
    Eigen::ArrayXd alpha = minimal::getAlpha(F.flow_pixel, rows, gamma);
    Eigen::ArrayXd alphaK = minimal::getAlphaK(F.coord_pixel, F.flow_pixel, rows, gamma);
    if (global_shutter) { alpha *= 0; alpha += 1; constant_acceleration = false; }
    Eigen::ArrayXd v_errors = Eigen::ArrayXd::Zero(num_evaluations), w_errors = Eigen::ArrayXd::Zero(num_evaluations),
                   mean_errors = Eigen::ArrayXd::Zero(num_evaluations), k = Eigen::ArrayXd::Zero(num_evaluations);
    Eigen::Array3Xd w = Eigen::Array3Xd::Zero(3, num_evaluations), v = Eigen::Array3Xd::Zero(3, num_evaluations);
    const double true_v_norm = true_values.v.norm();
    // first-order "rotations" I + [w]x of errorMeasure.cpp:127-129, :180-183
    auto rot = [](const Eigen::Vector3d &r) {
        Eigen::Matrix3d R = Eigen::Matrix3d::Identity();
        R(0, 1) = -r(2); R(0, 2) = r(1); R(1, 0) = r(2); R(1, 2) = -r(0); R(2, 0) = -r(1); R(2, 1) = r(0);
        return R;
    };
    const Eigen::Matrix3d true_rot_t = rot(true_values.w).transpose();
    const bool have_ground_truth = camera.getFrame(1).hasUnprojectionMaps();
    for (int eval_num = 0; eval_num < num_evaluations; eval_num++) {
        RansacValues ransac_results = minimal::ransac(F.coord, F.flow, alpha, alphaK, constant_acceleration, ransac_trials, TOL_RANSAC, show_messages);
        RansacValues results = ransac_results;
        if (optimize_results) results = nonlinear_refinement::nonLinearRefinement(F.flow, ransac_results, constant_acceleration, show_messages);
        double z_count = 0;
        for (int i = 0; i < results.num_inliers; i++) z_count += results.inliers(2, i);
        if (z_count * 1.0 / results.num_inliers < 0) { results.inliers.row(2) *= -1.0; results.v *= -1.0; }
        for (int a = 0; a < 3; ++a) { w(a, eval_num) = results.w(a); v(a, eval_num) = results.v(a); }
        k(eval_num) = results.k;
        const Eigen::Matrix3d Re = rot(results.w);
        auto err_rot = [&](int r, int c) { return Re(r, 0) * true_rot_t(0, c) + Re(r, 1) * true_rot_t(1, c) + Re(r, 2) * true_rot_t(2, c); };
        w_errors(eval_num) = Eigen::Vector3d(err_rot(2, 1), err_rot(0, 2), err_rot(1, 0)).norm();
        v_errors(eval_num) = std::acos(results.v.dot(true_values.v) / (results.v.norm() * true_v_norm));
        // depth range with the reference's start values (errorMeasure.cpp:187-197), 8-bit depth image, depth map
        double z_min = 100000, z_max = 0;
        for (int i = 0; i < results.num_inliers; ++i) {
            if (results.inliers(2, i) < z_min) z_min = results.inliers(2, i);
            if (results.inliers(2, i) > z_max) z_max = results.inliers(2, i);
        }
        const double multiplier = 244.0 / (z_max - z_min);
        cv::Mat depth_est(rows, cols, CV_8UC1, cv::Scalar(0));
        Eigen::MatrixXd depth_map = Eigen::MatrixXd::Zero(rows, cols);
        for (int i = 0; i < results.num_inliers; i++) {
            const int x = int(f_x * results.inliers(0, i) + c_x + 0.5), y = int(f_y * results.inliers(1, i) + c_y + 0.5);
            if (x < 0 || x >= cols || y < 0 || y >= rows) continue;       // the reference writes out of bounds here
            depth_est.at<unsigned char>(y, x) = (unsigned char)(10 + int((results.inliers(2, i) - z_min) * multiplier));
            depth_map(y, x) = results.inliers(2, i);
        }
        if (!image_path.empty()) cv::imwrite(image_path + std::to_string(eval_num) + ".png", depth_est);
        camera.setPose(1, results.k, results.v, results.w);
        camera.setDepthMap(1, depth_map);
        if (global_shutter) camera.backProjectGs(1); else camera.backProject(1);
        mean_errors(eval_num) = have_ground_truth ? camera.meanReprojectionError(1) : std::nan("");
        if (!image_path.empty()) camera.createPointCloud(1, image_path + std::to_string(eval_num) + ".ply");
        if (show_messages)
            std::cout << std::endl << "The rotation error for the current evaluation is " << w_errors(eval_num) << std::endl
                      << "The translation error for the current evaluation is " << v_errors(eval_num) << std::endl
                      << "The reprojection error for the current evaluation is " << mean_errors(eval_num) << std::endl;
        else
            std::cout << "Finished evaluation " << eval_num + 1 << "/" << num_evaluations << ". error_w = " << w_errors(eval_num)
                      << ". error_v = " << v_errors(eval_num) << ". reprojection error = " << mean_errors(eval_num) << std::endl;
    }
    const double ave_v_error = v_errors.mean(), ave_w_error = w_errors.mean(), ave_mean_error = mean_errors.mean();
    std::cout << std::endl << "The average rotation error is " << ave_w_error << std::endl
              << "The average translation error is " << ave_v_error << std::endl
              << "The average reprojection error is " << ave_mean_error << std::endl;
    // Q23: the Array3Xd error members get one well-formed column per evaluation
    Eigen::Array3Xd v_err3 = Eigen::Array3Xd::Zero(3, num_evaluations), w_err3 = Eigen::Array3Xd::Zero(3, num_evaluations);
    for (int j = 0; j < num_evaluations; ++j) { v_err3(0, j) = v_errors(j); w_err3(0, j) = w_errors(j); }
    return VelocityErrors(w, v, k, mean_errors, v_err3, w_err3, ave_w_error, ave_v_error, ave_mean_error);
}

}  // namespace error_measure
