// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing in the product path may include, link or call this.
//
// CPU restatement of the yaw grid search of SensorsCalibration's lidar2lidar auto-calibration:
//   Calibration_Tookit/SensorsCalibration/lidar2lidar/auto_calib/src/registration_icp.cpp
//     :38-47   GetDeltaT(const float yaw): Rz(yaw * M_PI / 180) — the caller passes radians, the helper converts again;
//              reproduced as written
//     :49-76   RegistrationByICP: 5 rounds, step 5 deg halved per round, half-range 10, 5, 2, 1, 0 (integer division): 37 evaluations
//     :78-100  CalculateICPError: pcl::transformPointCloud(src, T = GetDeltaT(yaw) * init_guess) with a double matrix
//              (result rounded to float), 1-NN in a pcl::KdTreeFLANN over the target, dist_sum += squared distance
// PCL / FLANN are not vendored: the kd-tree is the exact float search of o_kdtree.h (FLANN L2_Simple arithmetic).
// parity unpinned against the real libraries; the definition of every step is the reference's own source above.
#include <vector>
#include <cmath>
#include <cstring>
#include "o_kdtree.h"

namespace {

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

}  // extern "C"
