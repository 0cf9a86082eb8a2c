// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing in the product path may include, link or call this.
//
// CPU restatement of LIO-SAM's scan-to-map Levenberg-Marquardt step, following
// liosam_ws/src/LIO-SAM/src/mapOptmization.cpp line by line:
//   pointAssociateToMap        :278-284     transformPointCloud :286-305
//   updatePointAssociateToMap  :969-972     cornerOptimization  :974-1064
//   surfOptimization           :1066-1135   combineOptimizationCoeffs :1137-1156
//   LMOptimization             :1158-1280   scan2MapOptimization      :1282-1310
// Types and comparison precisions follow SURVEY.md Appendix A (double literals promote the
// comparisons; float overloads of sqrt/fabs/sin/cos). Library calls (kd-tree, cv::eigen, cv::solve,
// Eigen QR) are restated in o_kdtree.h / o_math.h.
// parity: unpinned by reference tests (the reference ships none for this path); the math kernels
// are pinned against cv2 4.13, the kNN against brute force/scipy (tests/test_oracle_*.py).
#include "o_math.h"
#include "o_kdtree.h"
#include <vector>
#include <cstdint>
#include <cstring>
#include <omp.h>

namespace {

struct S2M {
    int threads = 1;
    std::vector<float> mapC, mapS, scanC, scanS;   // xyzi, 16 B stride
    orc::KdTree kdC, kdS;
    bool isDegenerate = false;                     // member, persists across scans (:136)
    float matP[36] = {0};                          // member, persists (:234 zero-initialised)
    // per-iteration scratch (laserCloudOri*Vec / coeffSel*Vec / *Flag)
    std::vector<float> oriC, coeffC, oriS, coeffS;
    std::vector<uint8_t> flagC, flagS;
    std::vector<int> knnC, knnS;
    std::vector<float> d2C, d2S;
    std::vector<float> selOri, selCoeff;           // laserCloudOri / coeffSel
    float lastAtA[36] = {0}, lastAtB[6] = {0}, lastX[6] = {0};
};

inline void associate(const float t[12], const float* pi, float* po) {
    po[0] = t[0] * pi[0] + t[1] * pi[1] + t[2] * pi[2] + t[3];
    po[1] = t[4] * pi[0] + t[5] * pi[1] + t[6] * pi[2] + t[7];
    po[2] = t[8] * pi[0] + t[9] * pi[1] + t[10] * pi[2] + t[11];
    po[3] = pi[3];
}

// trans2Affine3f(transformTobeMapped): (roll, pitch, yaw, x, y, z) -> 3x4  (mapOptmization.cpp:340-343)
inline void pose_to_affine(const float T[6], float t[12]) {
    orc::pcl_get_transformation(T[3], T[4], T[5], T[0], T[1], T[2], t);
}

void corner_pass(S2M& s, const float T[6]) {
    float t[12]; pose_to_affine(T, t);
    const int n = (int)s.scanC.size() / 4;
    const float* map = s.mapC.data();
#pragma omp parallel for num_threads(s.threads) schedule(static)
    for (int i = 0; i < n; i++) {
        float pointOri[4], pointSel[4], coeff[4];
        std::memcpy(pointOri, &s.scanC[(size_t)i * 4], 16);
        associate(t, pointOri, pointSel);
        int* ind = &s.knnC[(size_t)i * 5]; float* sq = &s.d2C[(size_t)i * 5];
        for (int j = 0; j < 5; j++) { ind[j] = -1; sq[j] = INFINITY; }
        int found = s.kdC.knn(pointSel, 5, ind, sq);
        s.flagC[i] = 0;
        if (found < 5) continue;          // reference indexes [4] unconditionally; <5 map points never passes its guards
        if (sq[4] < 1.0) {
            float cx = 0, cy = 0, cz = 0;
            for (int j = 0; j < 5; j++) {
                cx += map[(size_t)ind[j] * 4 + 0];
                cy += map[(size_t)ind[j] * 4 + 1];
                cz += map[(size_t)ind[j] * 4 + 2];
            }
            cx /= 5; cy /= 5; cz /= 5;
            float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
            for (int j = 0; j < 5; j++) {
                float ax = map[(size_t)ind[j] * 4 + 0] - cx;
                float ay = map[(size_t)ind[j] * 4 + 1] - cy;
                float az = map[(size_t)ind[j] * 4 + 2] - cz;
                a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
                a22 += ay * ay; a23 += ay * az;
                a33 += az * az;
            }
            a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
            float A1[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33};
            float D1[3], V1[9];
            orc::jacobi_eigen_f32<3>(A1, D1, V1);
            if (D1[0] > 3 * D1[1]) {
                float x0 = pointSel[0], y0 = pointSel[1], z0 = pointSel[2];
                float x1 = cx + 0.1 * V1[0];
                float y1 = cy + 0.1 * V1[1];
                float z1 = cz + 0.1 * V1[2];
                float x2 = cx - 0.1 * V1[0];
                float y2 = cy - 0.1 * V1[1];
                float z2 = cz - 0.1 * V1[2];

                float a012 = std::sqrt(((x0 - x1)*(y0 - y2) - (x0 - x2)*(y0 - y1)) * ((x0 - x1)*(y0 - y2) - (x0 - x2)*(y0 - y1))
                                     + ((x0 - x1)*(z0 - z2) - (x0 - x2)*(z0 - z1)) * ((x0 - x1)*(z0 - z2) - (x0 - x2)*(z0 - z1))
                                     + ((y0 - y1)*(z0 - z2) - (y0 - y2)*(z0 - z1)) * ((y0 - y1)*(z0 - z2) - (y0 - y2)*(z0 - z1)));
                float l12 = std::sqrt((x1 - x2)*(x1 - x2) + (y1 - y2)*(y1 - y2) + (z1 - z2)*(z1 - z2));
                float la = ((y1 - y2)*((x0 - x1)*(y0 - y2) - (x0 - x2)*(y0 - y1))
                          + (z1 - z2)*((x0 - x1)*(z0 - z2) - (x0 - x2)*(z0 - z1))) / a012 / l12;
                float lb = -((x1 - x2)*((x0 - x1)*(y0 - y2) - (x0 - x2)*(y0 - y1))
                           - (z1 - z2)*((y0 - y1)*(z0 - z2) - (y0 - y2)*(z0 - z1))) / a012 / l12;
                float lc = -((x1 - x2)*((x0 - x1)*(z0 - z2) - (x0 - x2)*(z0 - z1))
                           + (y1 - y2)*((y0 - y1)*(z0 - z2) - (y0 - y2)*(z0 - z1))) / a012 / l12;
                float ld2 = a012 / l12;
                float sw = 1 - 0.9 * std::fabs(ld2);
                coeff[0] = sw * la; coeff[1] = sw * lb; coeff[2] = sw * lc; coeff[3] = sw * ld2;
                if (sw > 0.1) {
                    std::memcpy(&s.oriC[(size_t)i * 4], pointOri, 16);
                    std::memcpy(&s.coeffC[(size_t)i * 4], coeff, 16);
                    s.flagC[i] = 1;
                }
            }
        }
    }
}

void surf_pass(S2M& s, const float T[6]) {
    float t[12]; pose_to_affine(T, t);
    const int n = (int)s.scanS.size() / 4;
    const float* map = s.mapS.data();
#pragma omp parallel for num_threads(s.threads) schedule(static)
    for (int i = 0; i < n; i++) {
        float pointOri[4], pointSel[4], coeff[4];
        std::memcpy(pointOri, &s.scanS[(size_t)i * 4], 16);
        associate(t, pointOri, pointSel);
        int* ind = &s.knnS[(size_t)i * 5]; float* sq = &s.d2S[(size_t)i * 5];
        for (int j = 0; j < 5; j++) { ind[j] = -1; sq[j] = INFINITY; }
        int found = s.kdS.knn(pointSel, 5, ind, sq);
        s.flagS[i] = 0;
        if (found < 5) continue;
        if (sq[4] < 1.0) {
            float A0[15], B0[5] = {-1, -1, -1, -1, -1}, X0[3];
            for (int j = 0; j < 5; j++) {
                A0[j * 3 + 0] = map[(size_t)ind[j] * 4 + 0];
                A0[j * 3 + 1] = map[(size_t)ind[j] * 4 + 1];
                A0[j * 3 + 2] = map[(size_t)ind[j] * 4 + 2];
            }
            orc::colpiv_qr_solve_5x3(A0, B0, X0);
            float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
            float ps = std::sqrt(pa * pa + pb * pb + pc * pc);
            pa /= ps; pb /= ps; pc /= ps; pd /= ps;
            bool planeValid = true;
            for (int j = 0; j < 5; j++) {
                if (std::fabs(pa * map[(size_t)ind[j] * 4 + 0] +
                              pb * map[(size_t)ind[j] * 4 + 1] +
                              pc * map[(size_t)ind[j] * 4 + 2] + pd) > 0.2) {
                    planeValid = false;
                    break;
                }
            }
            if (planeValid) {
                float pd2 = pa * pointSel[0] + pb * pointSel[1] + pc * pointSel[2] + pd;
                float sw = 1 - 0.9 * std::fabs(pd2) / std::sqrt(std::sqrt(pointOri[0] * pointOri[0]
                         + pointOri[1] * pointOri[1] + pointOri[2] * pointOri[2]));
                coeff[0] = sw * pa; coeff[1] = sw * pb; coeff[2] = sw * pc; coeff[3] = sw * pd2;
                if (sw > 0.1) {
                    std::memcpy(&s.oriS[(size_t)i * 4], pointOri, 16);
                    std::memcpy(&s.coeffS[(size_t)i * 4], coeff, 16);
                    s.flagS[i] = 1;
                }
            }
        }
    }
}

void combine(S2M& s) {
    s.selOri.clear(); s.selCoeff.clear();
    const int nc = (int)s.flagC.size(), ns = (int)s.flagS.size();
    for (int i = 0; i < nc; i++) if (s.flagC[i]) {
        s.selOri.insert(s.selOri.end(), &s.oriC[(size_t)i * 4], &s.oriC[(size_t)i * 4] + 4);
        s.selCoeff.insert(s.selCoeff.end(), &s.coeffC[(size_t)i * 4], &s.coeffC[(size_t)i * 4] + 4);
    }
    for (int i = 0; i < ns; i++) if (s.flagS[i]) {
        s.selOri.insert(s.selOri.end(), &s.oriS[(size_t)i * 4], &s.oriS[(size_t)i * 4] + 4);
        s.selCoeff.insert(s.selCoeff.end(), &s.coeffS[(size_t)i * 4], &s.coeffS[(size_t)i * 4] + 4);
    }
    // flags are kept (not reset) so the debug getters can read them; every pass rewrites all of them
}

// returns: 1 converged, 0 keep optimising; *ran = 0 when K < 50 (LMOptimization returned false early)
int lm_optimization(S2M& s, float T[6], int iterCount, int* ran) {
    float srx = std::sin(T[1]), crx = std::cos(T[1]);
    float sry = std::sin(T[2]), cry = std::cos(T[2]);
    float srz = std::sin(T[0]), crz = std::cos(T[0]);
    const int K = (int)s.selOri.size() / 4;
    *ran = 0;
    if (K < 50) return 0;
    *ran = 1;
    std::vector<float> matA((size_t)K * 6), matB(K);
    for (int i = 0; i < K; i++) {
        float pox = s.selOri[(size_t)i * 4 + 1], poy = s.selOri[(size_t)i * 4 + 2], poz = s.selOri[(size_t)i * 4 + 0];
        float cfx = s.selCoeff[(size_t)i * 4 + 1], cfy = s.selCoeff[(size_t)i * 4 + 2], cfz = s.selCoeff[(size_t)i * 4 + 0];
        float cfi = s.selCoeff[(size_t)i * 4 + 3];
        float arx = (crx*sry*srz*pox + crx*crz*sry*poy - srx*sry*poz) * cfx
                  + (-srx*srz*pox - crz*srx*poy - crx*poz) * cfy
                  + (crx*cry*srz*pox + crx*cry*crz*poy - cry*srx*poz) * cfz;
        float ary = ((cry*srx*srz - crz*sry)*pox
                  + (sry*srz + cry*crz*srx)*poy + crx*cry*poz) * cfx
                  + ((-cry*crz - srx*sry*srz)*pox
                  + (cry*srz - crz*srx*sry)*poy - crx*sry*poz) * cfz;
        float arz = ((crz*srx*sry - cry*srz)*pox + (-cry*crz-srx*sry*srz)*poy)*cfx
                  + (crx*crz*pox - crx*srz*poy) * cfy
                  + ((sry*srz + cry*crz*srx)*pox + (crz*sry-cry*srx*srz)*poy)*cfz;
        matA[(size_t)i * 6 + 0] = arz; matA[(size_t)i * 6 + 1] = arx; matA[(size_t)i * 6 + 2] = ary;
        matA[(size_t)i * 6 + 3] = cfz; matA[(size_t)i * 6 + 4] = cfx; matA[(size_t)i * 6 + 5] = cfy;
        matB[i] = -cfi;
    }
    float AtA[36], AtB[6], X[6];
    for (int r = 0; r < 6; r++) {
        for (int c = 0; c < 6; c++) {
            double acc = 0.0;
            for (int k = 0; k < K; k++) acc += (double)matA[(size_t)k * 6 + r] * (double)matA[(size_t)k * 6 + c];
            AtA[r * 6 + c] = (float)acc;
        }
        double acc = 0.0;
        for (int k = 0; k < K; k++) acc += (double)matA[(size_t)k * 6 + r] * (double)matB[k];
        AtB[r] = (float)acc;
    }
    std::memcpy(s.lastAtA, AtA, sizeof(AtA)); std::memcpy(s.lastAtB, AtB, sizeof(AtB));
    {
        float Aw[36]; std::memcpy(Aw, AtA, sizeof(Aw)); std::memcpy(X, AtB, sizeof(X));
        if (!orc::qr_solve_f32<6>(Aw, X)) std::memset(X, 0, sizeof(X));   // cv::solve zeroes dst on failure
    }
    if (iterCount == 0) {
        float Aw[36], E[6], V[36], V2[36], Vinv[36];
        std::memcpy(Aw, AtA, sizeof(Aw));
        orc::jacobi_eigen_f32<6>(Aw, E, V);
        std::memcpy(V2, V, sizeof(V2));
        s.isDegenerate = false;
        const float eignThre[6] = {100, 100, 100, 100, 100, 100};
        for (int i = 5; i >= 0; i--) {
            if (E[i] < eignThre[i]) {
                for (int j = 0; j < 6; j++) V2[i * 6 + j] = 0;
                s.isDegenerate = true;
            } else break;
        }
        orc::lu_invert_f32<6>(V, Vinv);
        orc::gemm_f32_dacc(Vinv, V2, s.matP, 6, 6, 6);
    }
    if (s.isDegenerate) {
        float X2[6]; std::memcpy(X2, X, sizeof(X2));
        orc::gemm_f32_dacc(s.matP, X2, X, 6, 6, 1);
    }
    std::memcpy(s.lastX, X, sizeof(X));
    for (int k = 0; k < 6; k++) T[k] += X[k];
    // pcl::rad2deg(float) = alpha * 57.29578f ; pow(float, int) evaluates in double
    auto r2d = [](float a) { return a * 57.29578f; };
    float deltaR = std::sqrt(std::pow((double)r2d(X[0]), 2) + std::pow((double)r2d(X[1]), 2) + std::pow((double)r2d(X[2]), 2));
    float deltaT = std::sqrt(std::pow((double)(X[3] * 100), 2) + std::pow((double)(X[4] * 100), 2) + std::pow((double)(X[5] * 100), 2));
    if (deltaR < 0.05 && deltaT < 0.05) return 1;
    return 0;
}

}  // namespace

extern "C" {

// the team size a num_threads(requested) region gets here (bench.py reports it as cpu_baseline.cores)
int o_omp_threads(int requested) {
    int got = 1;
#pragma omp parallel num_threads(requested > 0 ? requested : 1)
    {
#pragma omp single
        got = omp_get_num_threads();
    }
    return got;
}

void* o_s2m_create(int threads) { S2M* s = new S2M(); s->threads = threads > 0 ? threads : 1; return s; }
void o_s2m_destroy(void* h) { delete (S2M*)h; }

// replaces kdtreeCornerFromMap->setInputCloud / kdtreeSurfFromMap->setInputCloud (:1289-1290)
void o_s2m_set_map(void* h, const float* corner, int nc, const float* surf, int ns) {
    S2M& s = *(S2M*)h;
    s.mapC.assign(corner, corner + (size_t)nc * 4);
    s.mapS.assign(surf, surf + (size_t)ns * 4);
    s.kdC.build(s.mapC.data(), nc, 4);
    s.kdS.build(s.mapS.data(), ns, 4);
}

void o_s2m_set_scan(void* h, const float* corner, int nc, const float* surf, int ns) {
    S2M& s = *(S2M*)h;
    s.scanC.assign(corner, corner + (size_t)nc * 4);
    s.scanS.assign(surf, surf + (size_t)ns * 4);
    s.oriC.assign((size_t)nc * 4, 0.f); s.coeffC.assign((size_t)nc * 4, 0.f); s.flagC.assign(nc, 0);
    s.oriS.assign((size_t)ns * 4, 0.f); s.coeffS.assign((size_t)ns * 4, 0.f); s.flagS.assign(ns, 0);
    s.knnC.assign((size_t)nc * 5, -1); s.d2C.assign((size_t)nc * 5, 0.f);
    s.knnS.assign((size_t)ns * 5, -1); s.d2S.assign((size_t)ns * 5, 0.f);
}

void o_s2m_set_state(void* h, int degenerate, const float* matP) {
    S2M& s = *(S2M*)h; s.isDegenerate = degenerate != 0; if (matP) std::memcpy(s.matP, matP, sizeof(s.matP));
}
void o_s2m_get_state(void* h, int* degenerate, float* matP) {
    S2M& s = *(S2M*)h; if (degenerate) *degenerate = s.isDegenerate; if (matP) std::memcpy(matP, s.matP, sizeof(s.matP));
}

// one pass of cornerOptimization + surfOptimization + combineOptimizationCoeffs + LMOptimization(iter)
// pose: in/out transformTobeMapped (roll, pitch, yaw, x, y, z). returns 1 when LMOptimization reports convergence.
int o_s2m_iterate(void* h, float* pose, int iter, int* n_sel, int* ran, float* AtA, float* AtB, float* X) {
    S2M& s = *(S2M*)h;
    corner_pass(s, pose);
    surf_pass(s, pose);
    combine(s);
    if (n_sel) *n_sel = (int)s.selOri.size() / 4;
    int r = 0;
    int conv = lm_optimization(s, pose, iter, &r);
    if (ran) *ran = r;
    if (AtA) std::memcpy(AtA, s.lastAtA, sizeof(s.lastAtA));
    if (AtB) std::memcpy(AtB, s.lastAtB, sizeof(s.lastAtB));
    if (X) std::memcpy(X, s.lastX, sizeof(s.lastX));
    return conv;
}

// per-feature results of the last pass. which: 0 corner, 1 surf
void o_s2m_get_pass(void* h, int which, int* knn_idx, float* knn_d2, float* coeff, unsigned char* flag) {
    S2M& s = *(S2M*)h;
    auto& k = which ? s.knnS : s.knnC; auto& d = which ? s.d2S : s.d2C;
    auto& c = which ? s.coeffS : s.coeffC; auto& f = which ? s.flagS : s.flagC;
    if (knn_idx) std::memcpy(knn_idx, k.data(), k.size() * sizeof(int));
    if (knn_d2) std::memcpy(knn_d2, d.data(), d.size() * sizeof(float));
    if (coeff) std::memcpy(coeff, c.data(), c.size() * sizeof(float));
    if (flag) std::memcpy(flag, f.data(), f.size());
}

// scan2MapOptimization (:1282-1310) without transformUpdate. Guards as the reference:
// runs only when n_corner > edge_min and n_surf > surf_min; else returns -1 and leaves pose.
// pose_hist: optional max_iters*6 floats (pose after each iteration); nsel_hist optional max_iters ints.
int o_s2m_solve(void* h, float* pose, int max_iters, int edge_min, int surf_min, int* iters_done,
                int* converged, float* pose_hist, int* nsel_hist) {
    S2M& s = *(S2M*)h;
    if (iters_done) *iters_done = 0;
    if (converged) *converged = 0;
    const int nc = (int)s.scanC.size() / 4, ns = (int)s.scanS.size() / 4;
    if (!(nc > edge_min && ns > surf_min)) return -1;
    int it = 0;
    for (; it < max_iters; it++) {
        int nsel = 0, ran = 0;
        int conv = o_s2m_iterate(h, pose, it, &nsel, &ran, nullptr, nullptr, nullptr);
        if (pose_hist) std::memcpy(&pose_hist[(size_t)it * 6], pose, 24);
        if (nsel_hist) nsel_hist[it] = nsel;
        if (conv) { if (converged) *converged = 1; ++it; break; }
    }
    if (iters_done) *iters_done = it;
    return 0;
}

// stand-alone batched kNN (kd-tree) for pinning against brute force / scipy
void o_knn(const float* pts, int n, const float* queries, int m, int k, int* idx, float* d2, int threads) {
    orc::KdTree kd; kd.build(pts, n, 4);
#pragma omp parallel for num_threads(threads > 0 ? threads : 1)
    for (int i = 0; i < m; i++) {
        for (int j = 0; j < k; j++) { idx[(size_t)i * k + j] = -1; d2[(size_t)i * k + j] = INFINITY; }
        kd.knn(&queries[(size_t)i * 4], k, &idx[(size_t)i * k], &d2[(size_t)i * k]);
    }
}

// brute force with the same metric and tie rule: the kd-tree's own checker
void o_knn_brute(const float* pts, int n, const float* queries, int m, int k, int* idx, float* d2, int threads) {
#pragma omp parallel for num_threads(threads > 0 ? threads : 1)
    for (int i = 0; i < m; i++) {
        for (int j = 0; j < k; j++) { idx[(size_t)i * k + j] = -1; d2[(size_t)i * k + j] = INFINITY; }
        orc::KdTree::Result res{k, 0, &idx[(size_t)i * k], &d2[(size_t)i * k]};
        const float* q = &queries[(size_t)i * 4];
        for (int p = 0; p < n; p++) {
            float dx = q[0] - pts[(size_t)p * 4], dy = q[1] - pts[(size_t)p * 4 + 1], dz = q[2] - pts[(size_t)p * 4 + 2];
            float d = 0.f; d += dx * dx; d += dy * dy; d += dz * dz;
            res.add(d, p);
        }
    }
}

// transformPointCloud (:286-305): p' = T(pose6) * p, pose = (roll, pitch, yaw, x, y, z)
void o_transform_cloud(const float* in, int n, const float* pose, float* out, int threads) {
    float t[12]; pose_to_affine(pose, t);
#pragma omp parallel for num_threads(threads > 0 ? threads : 1)
    for (int i = 0; i < n; i++) associate(t, &in[(size_t)i * 4], &out[(size_t)i * 4]);
}

// exposed math kernels (pinned against cv2 in tests)
void o_eigen3(const float* A, float* W, float* V) { float a[9]; std::memcpy(a, A, sizeof(a)); orc::jacobi_eigen_f32<3>(a, W, V); }
void o_eigen6(const float* A, float* W, float* V) { float a[36]; std::memcpy(a, A, sizeof(a)); orc::jacobi_eigen_f32<6>(a, W, V); }
int o_qr_solve6(const float* A, const float* b, float* x) { float a[36]; std::memcpy(a, A, sizeof(a)); std::memcpy(x, b, 24); return orc::qr_solve_f32<6>(a, x); }
int o_lu_invert6(const float* A, float* Ainv) { return orc::lu_invert_f32<6>(A, Ainv); }
void o_plane5(const float* A, float* x) { float b[5] = {-1, -1, -1, -1, -1}; orc::colpiv_qr_solve_5x3(A, b, x); }
void o_pose_affine(const float* pose, float* t12) { pose_to_affine(pose, t12); }

}  // extern "C"
