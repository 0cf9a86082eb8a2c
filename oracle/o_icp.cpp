// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing in the product path may include, link or call this.
//
// CPU restatement of the yaw grid search of SensorsCalibration's lidar2lidar auto-calibration:
//   Calibration_Tookit/SensorsCalibration/lidar2lidar/auto_calib/src/registration_icp.cpp
//     :38-47   GetDeltaT(const float yaw): Rz(yaw * M_PI / 180) — the caller passes radians, the helper converts again;
//              reproduced as written
//     :49-76   RegistrationByICP: 5 rounds, step 5 deg halved per round, half-range 10, 5, 2, 1, 0 (integer division): 37 evaluations
//     :78-100  CalculateICPError: pcl::transformPointCloud(src, T = GetDeltaT(yaw) * init_guess) with a double matrix
//              (result rounded to float), 1-NN in a pcl::KdTreeFLANN over the target, dist_sum += squared distance
// and of pcl::IterativeClosestPoint<PointType, PointType> as the loop-closure thread of LIO-SAM configures it:
//   liosam_ws/src/LIO-SAM/src/mapOptmization.cpp
//     :559-565  setMaxCorrespondenceDistance(historyKeyframeSearchRadius * 2), setMaximumIterations(100),
//               setTransformationEpsilon(1e-6), setEuclideanFitnessEpsilon(1e-6), setRANSACIterations(0)
//     :568-571  setInputSource / setInputTarget / align(unused_result)          (identity guess)
//     :573      hasConverged(), getFitnessScore()     :580,:586  getFinalTransformation()
//   following PCL 1.10 (registration/impl/icp.hpp, default_convergence_criteria.hpp, transformation_estimation_svd.hpp,
//   Eigen::umeyama without scaling): 1-NN correspondences within the distance threshold, rigid transform from the
//   cross-covariance SVD, the working cloud transformed incrementally in float, final = T * final, and the three-stage
//   convergence test (iterations; cos(angle) >= 1 - eps and |t|^2 <= eps; absolute 1e-12 / relative eps change of the
//   mean squared correspondence distance). The means, the covariance and its SVD are evaluated in double here (Eigen's
//   float vectorised reductions cannot be restated bit for bit): tolerance parity with the real library, unpinned.
// PCL / FLANN are not vendored: the kd-tree is the exact float search of o_kdtree.h (FLANN L2_Simple arithmetic).
// parity unpinned against the real libraries; the definition of every step is the reference's own source above.
#include <vector>
#include <cmath>
#include <cstring>
#include "o_kdtree.h"
#include "o_small_f64.h"
#include <cfloat>

namespace {

// rigid transform (R, t) minimising sum |R p + t - q|^2 from n, sum p, sum q, sum q p^T (Eigen::umeyama, no scaling)
void umeyama_from_sums(double n, const double sp[3], const double sq[3], const double sqp[9], float T[16]) {
    double mp[3], mq[3], S[9];
    for (int d = 0; d < 3; d++) { mp[d] = sp[d] / n; mq[d] = sq[d] / n; }
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) S[a * 3 + b] = sqp[a * 3 + b] / n - mq[a] * mp[b];      // dst_demean * src_demean^T / n
    // SVD of S through the eigen-decomposition of S^T S = V diag(s^2) V^T, U = S V diag(1/s)
    double StS[9], w[3], V[9];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { double acc = 0; for (int k = 0; k < 3; k++) acc += S[k * 3 + a] * S[k * 3 + b]; StS[a * 3 + b] = acc; }
    jacobi3d(StS, w, V);                                   // ascending; columns of V
    // descending order like a singular value decomposition
    double Vd[9], sv[3];
    for (int c = 0; c < 3; c++) { sv[c] = std::sqrt(std::max(w[2 - c], 0.0)); for (int r = 0; r < 3; r++) Vd[r * 3 + c] = V[r * 3 + (2 - c)]; }
    double U[9];
    for (int c = 0; c < 3; c++) {
        double col[3];
        for (int r = 0; r < 3; r++) col[r] = S[r * 3] * Vd[c] + S[r * 3 + 1] * Vd[3 + c] + S[r * 3 + 2] * Vd[6 + c];
        const double nrm = std::sqrt(col[0] * col[0] + col[1] * col[1] + col[2] * col[2]);
        if (c < 2 || nrm > 1e-12 * (sv[0] + 1e-300)) for (int r = 0; r < 3; r++) U[r * 3 + c] = nrm > 0 ? col[r] / nrm : (r == c ? 1.0 : 0.0);
        else {                                             // rank-deficient: complete the basis
            U[0 * 3 + 2] = U[1 * 3 + 0] * U[2 * 3 + 1] - U[2 * 3 + 0] * U[1 * 3 + 1];
            U[1 * 3 + 2] = U[2 * 3 + 0] * U[0 * 3 + 1] - U[0 * 3 + 0] * U[2 * 3 + 1];
            U[2 * 3 + 2] = U[0 * 3 + 0] * U[1 * 3 + 1] - U[1 * 3 + 0] * U[0 * 3 + 1];
        }
    }
    auto det3 = [](const double* M) { return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]); };
    const double sgn = det3(U) * det3(Vd) < 0 ? -1.0 : 1.0;
    double R[9];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) R[a * 3 + b] = U[a * 3] * Vd[b * 3] + U[a * 3 + 1] * Vd[b * 3 + 1] + sgn * U[a * 3 + 2] * Vd[b * 3 + 2];
    for (int i = 0; i < 16; i++) T[i] = 0.f;
    for (int a = 0; a < 3; a++) {
        for (int b = 0; b < 3; b++) T[a * 4 + b] = (float)R[a * 3 + b];
        T[a * 4 + 3] = (float)(mq[a] - (R[a * 3] * mp[0] + R[a * 3 + 1] * mp[1] + R[a * 3 + 2] * mp[2]));
    }
    T[15] = 1.f;
}

void mul4f(const float* A, const float* B, float* C) {
    float t[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { float s = 0; for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j]; t[i * 4 + j] = s; }
    std::memcpy(C, t, sizeof(t));
}

void delta_t(float yaw, double T[16]) {
    const double a = (double)yaw * M_PI / 180.0;
    const double c = std::cos(a), s = std::sin(a);
    const double R[16] = {c, -s, 0, 0, s, c, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    std::memcpy(T, R, sizeof(R));
}
void mul4(const double* A, const double* B, double* C) {
    double t[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j]; t[i * 4 + j] = s; }
    std::memcpy(C, t, sizeof(t));
}

struct IcpErr {
    std::vector<float> tgt, src;
    orc::KdTree kd;
    int evaluations = 0;
    double error(const double T[16]) {
        double sum = 0;
        const size_t n = src.size() / 3;
        for (size_t j = 0; j < n; j++) {
            const double x = src[j * 3], y = src[j * 3 + 1], z = src[j * 3 + 2];
            const float q[3] = {(float)(T[0] * x + T[1] * y + T[2] * z + T[3]), (float)(T[4] * x + T[5] * y + T[6] * z + T[7]),
                                (float)(T[8] * x + T[9] * y + T[10] * z + T[11])};
            int idx; float d2;
            if (kd.knn(q, 1, &idx, &d2) > 0) sum += d2;
        }
        evaluations++;
        return sum;
    }
    double error_yaw(const double init[16], float yaw) {
        double D[16], T[16];
        delta_t(yaw, D); mul4(D, init, T);
        return error(T);
    }
};

}  // namespace

extern "C" {

void* o_icperr_create(const float* tgt, int nt, const float* src, int ns) {
    IcpErr* h = new IcpErr();
    h->tgt.assign(tgt, tgt + (size_t)nt * 3); h->src.assign(src, src + (size_t)ns * 3);
    h->kd.build(h->tgt.data(), nt, 3);
    return h;
}
void o_icperr_destroy(void* h) { delete (IcpErr*)h; }
double o_icperr_evaluate(void* h, const double T[16]) { return ((IcpErr*)h)->error(T); }
// RegistrationByICP (:49-76). Returns the number of error evaluations.
int o_icperr_yaw_search(void* hh, const double init[16], double T_out[16], double* best_yaw_out, double* min_error_out) {
    IcpErr* h = (IcpErr*)hh;
    h->evaluations = 0;
    double cur_yaw = 0;
    double min_error = h->error_yaw(init, (float)cur_yaw);
    double best_yaw = cur_yaw;
    const float degree_2_radian = 0.017453293f;
    int iter_cnt = 0;
    double step = 5;
    int search_range = 10;
    while (iter_cnt < 5) {
        for (int delta = -search_range; delta < search_range; delta++) {
            const double yaw = cur_yaw + delta * step * degree_2_radian;
            const double error = h->error_yaw(init, (float)yaw);
            if (error < min_error) { min_error = error; best_yaw = yaw; }
        }
        search_range = (int)(search_range / 2 + 0.5);
        step /= 2;
        cur_yaw = best_yaw;
        iter_cnt++;
    }
    double D[16];
    delta_t((float)best_yaw, D);
    mul4(D, init, T_out);
    *best_yaw_out = best_yaw; *min_error_out = min_error;
    return h->evaluations;
}

// pcl::IterativeClosestPoint::align with an identity guess. Returns the number of iterations. history (optional): per
// iteration 16 floats of the incremental transformation.
int o_icp_align(const float* src, int ns, const float* tgt, int nt, double max_corr_dist, int max_iterations, double transformation_epsilon,
                double euclidean_fitness_epsilon, float final_T[16], int* converged, double* fitness_score, float* history) {
    orc::KdTree kd;
    kd.build(tgt, nt, 3);
    std::vector<float> cur(src, src + (size_t)ns * 3);
    static const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    float fin[16], tr[16];
    std::memcpy(fin, I, sizeof(I)); std::memcpy(tr, I, sizeof(I));
    int nr_iterations = 0;
    bool conv = false;
    double prev_mse = DBL_MAX;
    const double rotation_threshold = 1.0 - transformation_epsilon, translation_threshold = transformation_epsilon;
    const double mse_abs = 1e-12, mse_rel = euclidean_fitness_epsilon;
    const double max_d2 = max_corr_dist * max_corr_dist;
    do {
        double n = 0, sp[3] = {0, 0, 0}, sq[3] = {0, 0, 0}, sqp[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, sd = 0;
        for (int i = 0; i < ns; i++) {
            int j; float d2;
            if (kd.knn(&cur[(size_t)i * 3], 1, &j, &d2) < 1) continue;
            if ((double)d2 > max_d2) continue;
            const double p[3] = {cur[(size_t)i * 3], cur[(size_t)i * 3 + 1], cur[(size_t)i * 3 + 2]};
            const double q[3] = {tgt[(size_t)j * 3], tgt[(size_t)j * 3 + 1], tgt[(size_t)j * 3 + 2]};
            n += 1; sd += d2;
            for (int a = 0; a < 3; a++) { sp[a] += p[a]; sq[a] += q[a]; for (int b = 0; b < 3; b++) sqp[a * 3 + b] += q[a] * p[b]; }
        }
        if (n < 3) { conv = false; break; }
        umeyama_from_sums(n, sp, sq, sqp, tr);
        for (int i = 0; i < ns; i++) {
            const float x = cur[(size_t)i * 3], y = cur[(size_t)i * 3 + 1], z = cur[(size_t)i * 3 + 2];
            cur[(size_t)i * 3] = tr[0] * x + tr[1] * y + tr[2] * z + tr[3];
            cur[(size_t)i * 3 + 1] = tr[4] * x + tr[5] * y + tr[6] * z + tr[7];
            cur[(size_t)i * 3 + 2] = tr[8] * x + tr[9] * y + tr[10] * z + tr[11];
        }
        mul4f(tr, fin, fin);
        if (history) std::memcpy(history + (size_t)nr_iterations * 16, tr, 64);
        ++nr_iterations;
        // DefaultConvergenceCriteria::hasConverged
        if (nr_iterations >= max_iterations) { conv = true; break; }
        const double cos_angle = 0.5 * ((double)tr[0] + (double)tr[5] + (double)tr[10] - 1);
        const double translation_sqr = (double)tr[3] * tr[3] + (double)tr[7] * tr[7] + (double)tr[11] * tr[11];
        if (cos_angle >= rotation_threshold && translation_sqr <= translation_threshold) { conv = true; break; }
        const double cur_mse = sd / n;
        if (std::fabs(cur_mse - prev_mse) < mse_abs) { conv = true; break; }
        if (std::fabs(cur_mse - prev_mse) / prev_mse < mse_rel) { conv = true; break; }
        prev_mse = cur_mse;
    } while (!conv);
    std::memcpy(final_T, fin, sizeof(fin));
    *converged = conv ? 1 : 0;
    // getFitnessScore(): the input transformed by the final transformation, mean squared 1-NN distance
    double sum = 0; int nr = 0;
    for (int i = 0; i < ns; i++) {
        const float x = src[(size_t)i * 3], y = src[(size_t)i * 3 + 1], z = src[(size_t)i * 3 + 2];
        const float q[3] = {fin[0] * x + fin[1] * y + fin[2] * z + fin[3], fin[4] * x + fin[5] * y + fin[6] * z + fin[7], fin[8] * x + fin[9] * y + fin[10] * z + fin[11]};
        int j; float d2;
        if (kd.knn(q, 1, &j, &d2) < 1) continue;
        sum += d2; nr++;
    }
    *fitness_score = nr ? sum / nr : DBL_MAX;
    return nr_iterations;
}

}  // extern "C"
