// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing in the product path may include, link or call this.
//
// CPU restatement of pcl::NormalDistributionsTransform<PointXYZ, PointXYZ> as the multi-lidar calibrator drives it:
//   Calibration_Tookit/multi_lidar/src/multi_lidar_calibration/src/multi_lidar_calibrator.cpp
//     :35-43   ndt.setTransformationEpsilon / setStepSize / setResolution / setMaximumIterations / setInputSource / setInputTarget
//     :62      ndt.align(*output_cloud, current_guess_)
//     :64-65   ndt.hasConverged(), ndt.getFitnessScore(), ndt.getTransformationProbability()
//     :69,71   ndt.getFinalTransformation()
// PCL (1.10, the Noetic generation the toolkit's Docker images install) is not vendored in the reference and not
// installable here, so this follows its published algorithm (pcl/registration/impl/ndt.hpp,
// pcl/filters/impl/voxel_grid_covariance.hpp; Magnusson 2009 eq. 6.8-6.21; More & Thuente 1994):
//   target      VoxelGridCovariance, leaf = resolution: per voxel the double sums of p and p p^T, mean, the single-pass
//               covariance scaled by (n-1)/n, >= 6 points, eigenvalues below 0.01 * largest inflated, inverse covariance;
//               a float centroid per voxel feeds the radius search (radius = resolution, strict <, float L2)
//   derivatives score, 6-gradient, 6x6 Hessian over all (source point, neighbour voxel) pairs, eq. 6.9-6.13, with the
//               precomputed angular terms of eq. 6.19 / 6.21 (and PCL's small-angle shortcut |angle| < 10e-5)
//   iteration   delta = JacobiSVD(H).solve(-g); More-Thuente step length in [epsilon/2, step_size] (at most 10 trial
//               steps; the Hessian is recomputed afterwards with the angular second-derivative terms of the FIRST trial
//               step, as PCL does); pose vector p += delta; stop when iterations > max or |step| < epsilon
//   transforms  float 4x4 = Translation * Rx * Ry * Rz from the double pose vector, points transformed in float
// parity unpinned: no PCL binary or golden vector exists in the reference or in this image. Choices this oracle pins:
// the symmetric 3x3 / 6x6 eigen-decompositions are cyclic Jacobi in double (PCL: Eigen's tridiagonal QL and two-sided
// Jacobi SVD), eulerAngles(0,1,2) follows Eigen 3.3, neighbour voxels are visited in ascending (distance, voxel index),
// the float point transform sums left to right.
#include <vector>
#include <map>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cfloat>
#include <algorithm>
#include <numeric>
#include "o_small_f64.h"
#include "o_kdtree.h"

namespace {

struct Leaf {
    int nr_points = 0;
    double mean[3] = {0, 0, 0};
    double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double icov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    float centroid[3] = {0, 0, 0};
    int rank = -1;          // position in the centroid cloud (>= min points), -1 otherwise
};

// symmetric NxN Jacobi in double, eigenvalues unsorted in w, eigenvectors in the columns of V
template <int N>
void jacobi_sym(const double* A, double* w, double* V) {
    double a[N * N];
    std::memcpy(a, A, sizeof(a));
    for (int i = 0; i < N * N; i++) V[i] = (i % (N + 1) == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0;
        for (int p = 0; p < N; p++) for (int q = p + 1; q < N; q++) off += a[p * N + q] * a[p * N + q];
        if (off < 1e-300) break;
        for (int p = 0; p < N - 1; p++) for (int q = p + 1; q < N; q++) {
            const double apq = a[p * N + q];
            if (apq == 0.0) continue;
            const double theta = (a[q * N + q] - a[p * N + p]) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < N; k++) { const double akp = a[k * N + p], akq = a[k * N + q]; a[k * N + p] = c * akp - s * akq; a[k * N + q] = s * akp + c * akq; }
            for (int k = 0; k < N; k++) { const double apk = a[p * N + k], aqk = a[q * N + k]; a[p * N + k] = c * apk - s * aqk; a[q * N + k] = s * apk + c * aqk; }
            for (int k = 0; k < N; k++) { const double vkp = V[k * N + p], vkq = V[k * N + q]; V[k * N + p] = c * vkp - s * vkq; V[k * N + q] = s * vkp + c * vkq; }
        }
    }
    for (int i = 0; i < N; i++) w[i] = a[i * N + i];
}

// Eigen::JacobiSVD(H).solve(b) for a symmetric H: x = sum over singular values above the rank threshold
void svd_solve6(const double H[36], const double b[6], double x[6]) {
    double w[6], V[36];
    jacobi_sym<6>(H, w, V);
    double smax = 0;
    for (int i = 0; i < 6; i++) smax = std::max(smax, std::fabs(w[i]));
    const double thr = std::max(smax * 6.0 * DBL_EPSILON, DBL_MIN);
    for (int i = 0; i < 6; i++) x[i] = 0;
    for (int k = 0; k < 6; k++) {
        if (!(std::fabs(w[k]) > thr)) continue;
        double dot = 0;
        for (int i = 0; i < 6; i++) dot += V[i * 6 + k] * b[i];
        const double f = dot / w[k];
        for (int i = 0; i < 6; i++) x[i] += f * V[i * 6 + k];
    }
}

struct Ndt {
    float resolution = 1.0f;
    double step_size = 0.1, outlier_ratio = 0.55, trans_eps = 0.1;
    int max_iterations = 35;
    int min_points_per_voxel = 6;
    double min_covar_eigvalue_mult = 0.01;
    // target voxels
    float inv_leaf = 1.0f;
    int min_b[3] = {0, 0, 0}, div_b[3] = {1, 1, 1};
    std::map<int, Leaf> leaves;
    std::vector<int> centroid_leaf;          // leaf index of every centroid, ascending
    std::vector<float> target;               // xyz
    std::vector<float> source;               // xyz
    // state of the optimisation
    double gauss_d1 = 0, gauss_d2 = 0;
    double j_ang[8][3];
    double h_ang[15][3];
    float final_T[16];
    int nr_iterations = 0;
    bool converged = false;
    double trans_probability = 0;
    int evaluations = 0;
    long long pairs_last = 0;

    void set_target(const float* xyz, int n) {
        target.assign(xyz, xyz + (size_t)n * 3);
        leaves.clear(); centroid_leaf.clear();
        inv_leaf = 1.0f / resolution;
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        bool any = false;
        for (int i = 0; i < n; i++) {
            const float* p = &xyz[(size_t)i * 3];
            if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
            any = true;
            for (int d = 0; d < 3; d++) { mn[d] = std::min(mn[d], p[d]); mx[d] = std::max(mx[d], p[d]); }
        }
        if (!any) return;
        int64_t dx[3];
        for (int d = 0; d < 3; d++) dx[d] = (int64_t)((mx[d] - mn[d]) * inv_leaf) + 1;
        if (dx[0] * dx[1] * dx[2] > (int64_t)INT32_MAX) return;           // PCL warns and leaves the grid empty
        int max_b[3];
        for (int d = 0; d < 3; d++) { min_b[d] = (int)std::floor(mn[d] * inv_leaf); max_b[d] = (int)std::floor(mx[d] * inv_leaf); div_b[d] = max_b[d] - min_b[d] + 1; }
        const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
        std::map<int, std::vector<float>> csum;     // float centroid sums (x, y, z)
        for (int i = 0; i < n; i++) {
            const float* p = &xyz[(size_t)i * 3];
            if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
            int ijk[3];
            for (int d = 0; d < 3; d++) ijk[d] = (int)(std::floor(p[d] * inv_leaf) - (float)min_b[d]);
            const int idx = ijk[0] * mul[0] + ijk[1] * mul[1] + ijk[2] * mul[2];
            Leaf& lf = leaves[idx];
            const double q[3] = {p[0], p[1], p[2]};
            for (int a = 0; a < 3; a++) { lf.mean[a] += q[a]; for (int b = 0; b < 3; b++) lf.cov[a * 3 + b] += q[a] * q[b]; }
            for (int d = 0; d < 3; d++) lf.centroid[d] += p[d];
            lf.nr_points++;
        }
        for (auto& kv : leaves) {
            Leaf& lf = kv.second;
            const double npts = (double)lf.nr_points;
            for (int d = 0; d < 3; d++) lf.centroid[d] /= (float)lf.nr_points;
            double pt_sum[3] = {lf.mean[0], lf.mean[1], lf.mean[2]};
            for (int d = 0; d < 3; d++) lf.mean[d] /= npts;
            if (lf.nr_points < min_points_per_voxel) continue;
            lf.rank = (int)centroid_leaf.size();
            centroid_leaf.push_back(kv.first);
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++)
                lf.cov[a * 3 + b] = (lf.cov[a * 3 + b] - 2 * (pt_sum[a] * lf.mean[b])) / npts + lf.mean[a] * lf.mean[b];
            for (int q = 0; q < 9; q++) lf.cov[q] *= (npts - 1.0) / npts;
            double w[3], V[9];
            jacobi3d(lf.cov, w, V);                     // ascending, columns
            if (w[0] < 0 || w[1] < 0 || w[2] <= 0) { lf.nr_points = -1; continue; }
            const double min_ev = min_covar_eigvalue_mult * w[2];
            if (w[0] < min_ev) {
                w[0] = min_ev;
                if (w[1] < min_ev) w[1] = min_ev;
                double Vi[9];
                inv3d(V, Vi);
                for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
                    double s = 0;
                    for (int k = 0; k < 3; k++) s += V[a * 3 + k] * w[k] * Vi[k * 3 + b];
                    lf.cov[a * 3 + b] = s;
                }
            }
            inv3d(lf.cov, lf.icov);
            double mxc = -INFINITY, mnc = INFINITY;
            for (int q = 0; q < 9; q++) { mxc = std::max(mxc, lf.icov[q]); mnc = std::min(mnc, lf.icov[q]); }
            if (mxc == (double)std::numeric_limits<float>::infinity() || mnc == -(double)std::numeric_limits<float>::infinity()) lf.nr_points = -1;
        }
    }

    // neighbour voxels of a (float) point: centroids with float squared distance < (float)(resolution^2), ascending
    void neighbours(const float q[3], std::vector<std::pair<float, int>>& out) const {
        out.clear();
        if (centroid_leaf.empty()) return;
        const float r2 = (float)((double)resolution * (double)resolution);
        int c[3];
        for (int d = 0; d < 3; d++) c[d] = (int)(std::floor(q[d] * inv_leaf) - (float)min_b[d]);
        for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
            const int x = c[0] + dx, y = c[1] + dy, z = c[2] + dz;
            if (x < 0 || y < 0 || z < 0 || x >= div_b[0] || y >= div_b[1] || z >= div_b[2]) continue;
            const int idx = x + y * div_b[0] + z * div_b[0] * div_b[1];
            auto it = leaves.find(idx);
            if (it == leaves.end() || it->second.rank < 0) continue;
            const float* ce = it->second.centroid;
            float d2 = 0;
            for (int d = 0; d < 3; d++) { const float diff = q[d] - ce[d]; d2 += diff * diff; }
            if (d2 < r2) out.emplace_back(d2, idx);
        }
        std::sort(out.begin(), out.end());
    }

    static void pose_to_matrix(const double p[6], float T[16]) {
        // Translation3f * AngleAxisf(rx, X) * AngleAxisf(ry, Y) * AngleAxisf(rz, Z), all in float
        const float rx = (float)p[3], ry = (float)p[4], rz = (float)p[5];
        const float cx = std::cos(rx), sx = std::sin(rx), cy = std::cos(ry), sy = std::sin(ry), cz = std::cos(rz), sz = std::sin(rz);
        const float Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
        const float Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
        const float Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
        float A[9], R[9];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { float s = 0; for (int k = 0; k < 3; k++) s += Rx[i * 3 + k] * Ry[k * 3 + j]; A[i * 3 + j] = s; }
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { float s = 0; for (int k = 0; k < 3; k++) s += A[i * 3 + k] * Rz[k * 3 + j]; R[i * 3 + j] = s; }
        for (int i = 0; i < 16; i++) T[i] = 0;
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) T[i * 4 + j] = R[i * 3 + j]; T[i * 4 + 3] = (float)p[i]; }
        T[15] = 1;
    }

    static void transform_cloud(const std::vector<float>& in, const float T[16], std::vector<float>& out) {
        out.resize(in.size());
        for (size_t i = 0; i < in.size() / 3; i++) {
            const float x = in[i * 3], y = in[i * 3 + 1], z = in[i * 3 + 2];
            out[i * 3] = T[0] * x + T[1] * y + T[2] * z + T[3];
            out[i * 3 + 1] = T[4] * x + T[5] * y + T[6] * z + T[7];
            out[i * 3 + 2] = T[8] * x + T[9] * y + T[10] * z + T[11];
        }
    }

    void angle_derivatives(const double p[6], bool compute_hessian) {
        double cx, cy, cz, sx, sy, sz;
        if (std::fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p[3]); sx = std::sin(p[3]); }
        if (std::fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p[4]); sy = std::sin(p[4]); }
        if (std::fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p[5]); sz = std::sin(p[5]); }
        const double ja[8][3] = {{-sx * sz + cx * sy * cz, -sx * cz - cx * sy * sz, -cx * cy},
                                 {cx * sz + sx * sy * cz, cx * cz - sx * sy * sz, -sx * cy},
                                 {-sy * cz, sy * sz, cy},
                                 {sx * cy * cz, -sx * cy * sz, sx * sy},
                                 {-cx * cy * cz, cx * cy * sz, -cx * sy},
                                 {-cy * sz, -cy * cz, 0},
                                 {cx * cz - sx * sy * sz, -cx * sz - sx * sy * cz, 0},
                                 {sx * cz + cx * sy * sz, cx * sy * cz - sx * sz, 0}};
        std::memcpy(j_ang, ja, sizeof(ja));
        if (compute_hessian) {
            const double ha[15][3] = {{-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, sx * cy},      // a2
                                      {-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, -cx * cy},     // a3
                                      {cx * cy * cz, -cx * cy * sz, cx * sy},                           // b2
                                      {sx * cy * cz, -sx * cy * sz, sx * sy},                           // b3
                                      {-sx * cz - cx * sy * sz, sx * sz - cx * sy * cz, 0},             // c2
                                      {cx * cz - sx * sy * sz, -sx * sy * cz - cx * sz, 0},             // c3
                                      {-cy * cz, cy * sz, sy},                                          // d1
                                      {-sx * sy * cz, sx * sy * sz, sx * cy},                           // d2
                                      {cx * sy * cz, -cx * sy * sz, -cx * cy},                          // d3
                                      {sy * sz, sy * cz, 0},                                            // e1
                                      {-sx * cy * sz, -sx * cy * cz, 0},                                // e2
                                      {cx * cy * sz, cx * cy * cz, 0},                                  // e3
                                      {-cy * cz, cy * sz, 0},                                           // f1
                                      {-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, 0},            // f2
                                      {-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, 0}};           // f3
            std::memcpy(h_ang, ha, sizeof(ha));
        }
    }

    // point_gradient (3x6) and point_hessian (18x6) of eq. 6.18-6.21 for the untransformed point x
    void point_derivatives(const double x[3], double pg[18], double ph[108], bool compute_hessian) const {
        auto dot = [&](const double v[3]) { return x[0] * v[0] + x[1] * v[1] + x[2] * v[2]; };
        for (int i = 0; i < 18; i++) pg[i] = 0;
        pg[0 * 6 + 0] = 1; pg[1 * 6 + 1] = 1; pg[2 * 6 + 2] = 1;
        pg[1 * 6 + 3] = dot(j_ang[0]); pg[2 * 6 + 3] = dot(j_ang[1]);
        pg[0 * 6 + 4] = dot(j_ang[2]); pg[1 * 6 + 4] = dot(j_ang[3]); pg[2 * 6 + 4] = dot(j_ang[4]);
        pg[0 * 6 + 5] = dot(j_ang[5]); pg[1 * 6 + 5] = dot(j_ang[6]); pg[2 * 6 + 5] = dot(j_ang[7]);
        if (!compute_hessian) return;
        for (int i = 0; i < 108; i++) ph[i] = 0;
        const double a[3] = {0, dot(h_ang[0]), dot(h_ang[1])}, b[3] = {0, dot(h_ang[2]), dot(h_ang[3])}, c[3] = {0, dot(h_ang[4]), dot(h_ang[5])};
        const double d[3] = {dot(h_ang[6]), dot(h_ang[7]), dot(h_ang[8])}, e[3] = {dot(h_ang[9]), dot(h_ang[10]), dot(h_ang[11])};
        const double f[3] = {dot(h_ang[12]), dot(h_ang[13]), dot(h_ang[14])};
        auto put = [&](int row, int col, const double v[3]) { for (int k = 0; k < 3; k++) ph[(row + k) * 6 + col] = v[k]; };
        put(9, 3, a); put(12, 3, b); put(15, 3, c);
        put(9, 4, b); put(12, 4, d); put(15, 4, e);
        put(9, 5, c); put(12, 5, e); put(15, 5, f);
    }

    // mode 0: score + gradient; 1: + hessian; 2: hessian only (computeHessian)
    double derivatives(const std::vector<float>& trans, const double p[6], int mode, double grad[6], double hess[36], bool refresh_angles) {
        if (refresh_angles) angle_derivatives(p, mode == 1);
        if (mode != 2) for (int i = 0; i < 6; i++) grad[i] = 0;
        if (mode != 0) for (int i = 0; i < 36; i++) hess[i] = 0;
        double score = 0;
        long long pairs = 0;
        std::vector<std::pair<float, int>> nb;
        const size_t n = source.size() / 3;
        double pg[18], ph[108];
        for (size_t idx = 0; idx < n; idx++) {
            const float* xt = &trans[idx * 3];
            neighbours(xt, nb);
            for (auto& pr : nb) {
                const Leaf& cell = leaves.at(pr.second);
                const double x[3] = {source[idx * 3], source[idx * 3 + 1], source[idx * 3 + 2]};
                double xtr[3] = {(double)xt[0] - cell.mean[0], (double)xt[1] - cell.mean[1], (double)xt[2] - cell.mean[2]};
                const double* ci = cell.icov;
                point_derivatives(x, pg, ph, mode != 0);
                pairs++;
                double cx[3];
                for (int a = 0; a < 3; a++) cx[a] = ci[a * 3] * xtr[0] + ci[a * 3 + 1] * xtr[1] + ci[a * 3 + 2] * xtr[2];
                const double xcx = xtr[0] * cx[0] + xtr[1] * cx[1] + xtr[2] * cx[2];
                double e = std::exp(-gauss_d2 * xcx / 2);
                const double score_inc = -gauss_d1 * e;
                e = gauss_d2 * e;
                if (e > 1 || e < 0 || e != e) continue;
                e *= gauss_d1;
                if (mode != 2) score += score_inc;
                double cdx[6][3], xdot[6];
                for (int i = 0; i < 6; i++) {
                    for (int a = 0; a < 3; a++) cdx[i][a] = ci[a * 3] * pg[0 * 6 + i] + ci[a * 3 + 1] * pg[1 * 6 + i] + ci[a * 3 + 2] * pg[2 * 6 + i];
                    xdot[i] = xtr[0] * cdx[i][0] + xtr[1] * cdx[i][1] + xtr[2] * cdx[i][2];
                }
                for (int i = 0; i < 6; i++) {
                    if (mode != 2) grad[i] += xdot[i] * e;
                    if (mode == 0) continue;
                    for (int j = 0; j < 6; j++) {
                        double chh[3];
                        const double hx = ph[(3 * i) * 6 + j], hy = ph[(3 * i + 1) * 6 + j], hz = ph[(3 * i + 2) * 6 + j];
                        for (int a = 0; a < 3; a++) chh[a] = ci[a * 3] * hx + ci[a * 3 + 1] * hy + ci[a * 3 + 2] * hz;
                        const double t2 = xtr[0] * chh[0] + xtr[1] * chh[1] + xtr[2] * chh[2];
                        const double t3 = pg[0 * 6 + j] * cdx[i][0] + pg[1 * 6 + j] * cdx[i][1] + pg[2 * 6 + j] * cdx[i][2];
                        hess[i * 6 + j] += e * (-gauss_d2 * xdot[i] * xdot[j] + t2 + t3);
                    }
                }
            }
        }
        pairs_last = pairs;
        evaluations++;
        return score;
    }

    static double psi(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
    static double dpsi(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

    static bool update_interval(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t, double g_t) {
        if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
        if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
        if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
        return true;
    }

    static double trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
        if (f_t > f_l) {
            const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
            const double w = std::sqrt(z * z - g_t * g_l);
            const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
            const double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
            if (std::fabs(a_c - a_l) < std::fabs(a_q - a_l)) return a_c;
            return 0.5 * (a_q + a_c);
        }
        if (g_t * g_l < 0) {
            const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
            const double w = std::sqrt(z * z - g_t * g_l);
            const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
            const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
            if (std::fabs(a_c - a_t) >= std::fabs(a_s - a_t)) return a_c;
            return a_s;
        }
        if (std::fabs(g_t) <= std::fabs(g_l)) {
            const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
            const double w = std::sqrt(z * z - g_t * g_l);
            const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
            const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
            const double a_t_next = (std::fabs(a_c - a_t) < std::fabs(a_s - a_t)) ? a_c : a_s;
            if (a_t > a_l) return std::min(a_t + 0.66 * (a_u - a_t), a_t_next);
            return std::max(a_t + 0.66 * (a_u - a_t), a_t_next);
        }
        const double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
        const double w = std::sqrt(z * z - g_t * g_u);
        return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
    }

    double step_length_mt(const double x[6], double step_dir[6], double step_init, double step_max, double step_min, double& score,
                          double grad[6], double hess[36], std::vector<float>& trans) {
        const double phi_0 = -score;
        double d_phi_0 = 0;
        for (int i = 0; i < 6; i++) d_phi_0 += grad[i] * step_dir[i];
        d_phi_0 = -d_phi_0;
        double x_t[6];
        if (d_phi_0 >= 0) {
            if (d_phi_0 == 0) return 0;
            d_phi_0 *= -1;
            for (int i = 0; i < 6; i++) step_dir[i] *= -1;
        }
        const int max_step_iterations = 10;
        int step_iterations = 0;
        const double mu = 1.e-4, nu = 0.9;
        double a_l = 0, a_u = 0;
        double f_l = psi(a_l, phi_0, phi_0, d_phi_0, mu), g_l = dpsi(d_phi_0, d_phi_0, mu);
        double f_u = psi(a_u, phi_0, phi_0, d_phi_0, mu), g_u = dpsi(d_phi_0, d_phi_0, mu);
        bool interval_converged = (step_max - step_min) < 0, open_interval = true;
        double a_t = step_init;
        a_t = std::min(a_t, step_max);
        a_t = std::max(a_t, step_min);
        for (int i = 0; i < 6; i++) x_t[i] = x[i] + step_dir[i] * a_t;
        pose_to_matrix(x_t, final_T);
        transform_cloud(source, final_T, trans);
        score = derivatives(trans, x_t, 1, grad, hess, true);
        double phi_t = -score, d_phi_t = 0;
        for (int i = 0; i < 6; i++) d_phi_t += grad[i] * step_dir[i];
        d_phi_t = -d_phi_t;
        double psi_t = psi(a_t, phi_t, phi_0, d_phi_0, mu), d_psi_t = dpsi(d_phi_t, d_phi_0, mu);
        while (!interval_converged && step_iterations < max_step_iterations && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
            if (open_interval) a_t = trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
            else a_t = trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
            a_t = std::min(a_t, step_max);
            a_t = std::max(a_t, step_min);
            for (int i = 0; i < 6; i++) x_t[i] = x[i] + step_dir[i] * a_t;
            pose_to_matrix(x_t, final_T);
            transform_cloud(source, final_T, trans);
            score = derivatives(trans, x_t, 0, grad, hess, true);
            phi_t = -score; d_phi_t = 0;
            for (int i = 0; i < 6; i++) d_phi_t += grad[i] * step_dir[i];
            d_phi_t = -d_phi_t;
            psi_t = psi(a_t, phi_t, phi_0, d_phi_0, mu); d_psi_t = dpsi(d_phi_t, d_phi_0, mu);
            if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
                open_interval = false;
                f_l = f_l + phi_0 - mu * d_phi_0 * a_l; g_l = g_l + mu * d_phi_0;
                f_u = f_u + phi_0 - mu * d_phi_0 * a_u; g_u = g_u + mu * d_phi_0;
            }
            if (open_interval) interval_converged = update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
            else interval_converged = update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
            step_iterations++;
        }
        // PCL: computeHessian reuses the angular second derivatives left by the last compute_hessian = true call
        if (step_iterations) derivatives(trans, x_t, 2, grad, hess, false);
        return a_t;
    }

    static void euler_012(const float T[16], float out[3]) {
        // Eigen 3.3 Matrix3f::eulerAngles(0, 1, 2)
        auto c = [&](int r, int col) { return T[r * 4 + col]; };
        float res[3];
        res[0] = std::atan2(c(1, 2), c(2, 2));
        const float c2 = std::sqrt(c(0, 0) * c(0, 0) + c(0, 1) * c(0, 1));
        if (res[0] > 0.f) {
            if (res[0] > 0.f) res[0] -= (float)M_PI; else res[0] += (float)M_PI;
            res[1] = std::atan2(-c(0, 2), -c2);
        } else res[1] = std::atan2(-c(0, 2), c2);
        const float s1 = std::sin(res[0]), c1 = std::cos(res[0]);
        res[2] = std::atan2(s1 * c(2, 0) - c1 * c(1, 0), c1 * c(1, 1) - s1 * c(2, 1));
        for (int i = 0; i < 3; i++) out[i] = -res[i];
    }

    void align(const float guess[16]) {
        nr_iterations = 0; converged = false; evaluations = 0;
        const double gauss_c1 = 10 * (1 - outlier_ratio);
        const double gauss_c2 = outlier_ratio / std::pow((double)resolution, 3);
        const double gauss_d3 = -std::log(gauss_c2);
        gauss_d1 = -std::log(gauss_c1 + gauss_c2) - gauss_d3;
        gauss_d2 = -2 * std::log((-std::log(gauss_c1 * std::exp(-0.5) + gauss_c2) - gauss_d3) / gauss_d1);
        static const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        std::memcpy(final_T, I, sizeof(I));
        std::vector<float> output = source;
        if (std::memcmp(guess, I, sizeof(I)) != 0) {
            std::memcpy(final_T, guess, sizeof(I));
            transform_cloud(source, guess, output);
        }
        double p[6], delta_p[6], grad[6], hess[36];
        float rot[3];
        euler_012(final_T, rot);
        p[0] = final_T[3]; p[1] = final_T[7]; p[2] = final_T[11];
        p[3] = rot[0]; p[4] = rot[1]; p[5] = rot[2];
        double score = derivatives(output, p, 1, grad, hess, true);
        const size_t npts = source.size() / 3;
        while (!converged) {
            double neg[6];
            for (int i = 0; i < 6; i++) neg[i] = -grad[i];
            svd_solve6(hess, neg, delta_p);
            double nrm = 0;
            for (int i = 0; i < 6; i++) nrm += delta_p[i] * delta_p[i];
            double delta_p_norm = std::sqrt(nrm);
            if (delta_p_norm == 0 || delta_p_norm != delta_p_norm) {
                trans_probability = score / (double)npts;
                converged = delta_p_norm == delta_p_norm;
                return;
            }
            for (int i = 0; i < 6; i++) delta_p[i] /= delta_p_norm;
            delta_p_norm = step_length_mt(p, delta_p, delta_p_norm, step_size, trans_eps / 2, score, grad, hess, output);
            for (int i = 0; i < 6; i++) delta_p[i] *= delta_p_norm;
            for (int i = 0; i < 6; i++) p[i] = p[i] + delta_p[i];
            if (nr_iterations > max_iterations || (nr_iterations && (std::fabs(delta_p_norm) < trans_eps))) converged = true;
            nr_iterations++;
        }
        trans_probability = score / (double)npts;
    }

    // getFitnessScore(): mean squared distance from the transformed source to its nearest target point (float kd-tree)
    double fitness(const float T[16]) const {
        std::vector<float> tr;
        transform_cloud(source, T, tr);
        const int nt = (int)(target.size() / 3);
        orc::KdTree kd;
        kd.build(target.data(), nt, 3);
        double sum = 0; int nr = 0;
        for (size_t i = 0; i < tr.size() / 3; i++) {
            int idx; float d2;
            const float q[3] = {tr[i * 3], tr[i * 3 + 1], tr[i * 3 + 2]};
            if (kd.knn(q, 1, &idx, &d2) < 1) continue;
            sum += d2; nr++;
        }
        return nr > 0 ? sum / nr : DBL_MAX;
    }
};

}  // namespace

extern "C" {

void* o_ndt_create(float resolution, double step_size, double trans_eps, int max_iterations) {
    Ndt* n = new Ndt();
    n->resolution = resolution; n->step_size = step_size; n->trans_eps = trans_eps; n->max_iterations = max_iterations;
    return n;
}
void o_ndt_destroy(void* h) { delete (Ndt*)h; }
int o_ndt_set_target(void* h, const float* xyz, int n) { Ndt* d = (Ndt*)h; d->set_target(xyz, n); return (int)d->centroid_leaf.size(); }
void o_ndt_set_source(void* h, const float* xyz, int n) { ((Ndt*)h)->source.assign(xyz, xyz + (size_t)n * 3); }
// per valid voxel (ascending voxel index): index, point count (-1 = rejected covariance), float centroid, mean, inverse covariance
void o_ndt_get_voxels(void* h, int* leaf_idx, int* npts, float* centroid, double* mean, double* icov) {
    Ndt* d = (Ndt*)h;
    int k = 0;
    for (int idx : d->centroid_leaf) {
        const Leaf& lf = d->leaves.at(idx);
        leaf_idx[k] = idx; npts[k] = lf.nr_points;
        std::memcpy(&centroid[k * 3], lf.centroid, 12); std::memcpy(&mean[k * 3], lf.mean, 24); std::memcpy(&icov[k * 9], lf.icov, 72);
        k++;
    }
}
void o_ndt_grid_geometry(void* h, int min_b[3], int div_b[3]) { Ndt* d = (Ndt*)h; std::memcpy(min_b, d->min_b, 12); std::memcpy(div_b, d->div_b, 12); }
// one derivative pass at pose vector p (x, y, z, rx, ry, rz); returns the score; hess may be NULL
double o_ndt_derivatives(void* h, const double p[6], double grad[6], double hess[36], long long* pairs) {
    Ndt* d = (Ndt*)h;
    const double c1 = 10 * (1 - d->outlier_ratio), c2 = d->outlier_ratio / std::pow((double)d->resolution, 3), d3 = -std::log(c2);
    d->gauss_d1 = -std::log(c1 + c2) - d3;
    d->gauss_d2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / d->gauss_d1);
    float T[16];
    Ndt::pose_to_matrix(p, T);
    std::vector<float> tr;
    Ndt::transform_cloud(d->source, T, tr);
    double g[6], H[36];
    const double s = d->derivatives(tr, p, 1, g, H, true);
    std::memcpy(grad, g, sizeof(g));
    if (hess) std::memcpy(hess, H, sizeof(H));
    if (pairs) *pairs = d->pairs_last;
    return s;
}
void o_ndt_pose_to_matrix(const double p[6], float T[16]) { Ndt::pose_to_matrix(p, T); }
void o_ndt_align(void* h, const float guess[16], float final_T[16], int* iterations, int* converged, double* trans_probability, int* evaluations) {
    Ndt* d = (Ndt*)h;
    d->align(guess);
    std::memcpy(final_T, d->final_T, 64);
    *iterations = d->nr_iterations; *converged = d->converged ? 1 : 0; *trans_probability = d->trans_probability; *evaluations = d->evaluations;
}
double o_ndt_fitness(void* h, const float T[16]) { return ((Ndt*)h)->fitness(T); }
void o_svd_solve6(const double H[36], const double b[6], double x[6]) { svd_solve6(H, b, x); }

}  // extern "C"
