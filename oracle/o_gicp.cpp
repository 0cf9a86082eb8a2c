// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing in the product path may include, link or call this.
//
// CPU restatement (fp64) of the Open3D calls Multi_LiCa makes for its GICP calibration:
//   Calibration_Tookit/Multi_LiCa/multi_lidar_calibrator/calibration/Calibration.py
//     :314-315  pcd.voxel_down_sample(voxel_size)
//     :327-328  pcd.estimate_normals()                       (default KDTreeSearchParamKNN(30))
//     :331-340  registration_generalized_icp(source, target, max_corr, init,
//                   TransformationEstimationForGeneralizedICP(epsilon), ICPConvergenceCriteria(rel_fit, rel_rmse, max_it))
// Open3D is an unpinned, un-vendored pip dependency (Multi_LiCa/requirements.txt:1; 0.17-0.19 by the Dockerfile's
// Python 3.10 / Ubuntu 22.04) and is not installed here: this follows its published algorithm —
//   voxel_down_sample : voxel = floor((p - (min - v/2)) / v), one mean (double) per voxel
//   estimate_normals  : per point, covariance of its 30 nearest neighbours (itself included), eigenvector of the smallest
//                       eigenvalue
//   GICP              : per-point covariance R diag(eps,1,1) R^T from the normal (Rodrigues rotation of e1 onto n);
//                       ICP loop: 1-NN within max_corr -> fitness = n_corr/n_src, inlier_rmse = sqrt(sum d^2 / n_corr);
//                       per pair M = Ct + Cs, W = (M^-1)^(1/2), r = W(vs - vt), J = W[-[vs]x | I]; x = LDLT(JtJ)^-1(-Jtr);
//                       T <- [Rz(x2)Ry(x1)Rx(x0) | x3..5] T; stop on |dfitness| < rel_fit && |drmse| < rel_rmse or max_it.
// parity unpinned: no Open3D binary or golden vector exists in the reference or this image. Definitions this oracle
// pins (and the CUDA path shares): output order of voxel_down_sample = ascending (z,y,x) voxel key (Open3D's is hash-map
// order); the 3x3 eigenvector comes from a cyclic Jacobi in fp64 (Open3D uses a closed-form solver); ties in
// nearest-neighbour distance fall to the smaller index; sums run in source index order; every iteration transforms the
// ORIGINAL source by the accumulated T (Open3D transforms its working copy incrementally: equal up to rounding).
#include <vector>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <numeric>
#include <map>
#include <omp.h>
#include "o_small_f64.h"

namespace {

// ---- fp64 kd-tree (exact kNN, ties to the smaller index)
struct KdTreeD {
    struct Node { int left = -1, right = -1, lo = 0, hi = 0, dim = 0; double cut = 0; };
    const double* pts = nullptr; int n = 0;
    std::vector<int> perm; std::vector<Node> nodes;
    void build(const double* p, int n_) {
        pts = p; n = n_; perm.resize(n); std::iota(perm.begin(), perm.end(), 0); nodes.clear();
        if (n) divide(0, n);
    }
    int divide(int lo, int hi) {
        int id = (int)nodes.size(); nodes.emplace_back();
        if (hi - lo <= 16) { nodes[id].lo = lo; nodes[id].hi = hi; return id; }
        double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
        for (int i = lo; i < hi; i++) for (int d = 0; d < 3; d++) { double v = pts[(size_t)perm[i] * 3 + d]; mn[d] = std::min(mn[d], v); mx[d] = std::max(mx[d], v); }
        int dim = 0; for (int d = 1; d < 3; d++) if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
        int mid = (lo + hi) / 2;
        std::nth_element(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi, [&](int a, int b) { return pts[(size_t)a * 3 + dim] < pts[(size_t)b * 3 + dim]; });
        double cut = pts[(size_t)perm[mid] * 3 + dim];
        int L = divide(lo, mid), R = divide(mid, hi);
        nodes[id].left = L; nodes[id].right = R; nodes[id].dim = dim; nodes[id].cut = cut;
        return id;
    }
    struct Res { int k, cnt; int* idx; double* d2; };
    static void add(Res& r, double d, int i) {
        if (r.cnt == r.k && (d > r.d2[r.k - 1] || (d == r.d2[r.k - 1] && i > r.idx[r.k - 1]))) return;
        int pos = r.cnt < r.k ? r.cnt : r.k - 1;
        while (pos > 0 && (r.d2[pos - 1] > d || (r.d2[pos - 1] == d && r.idx[pos - 1] > i))) { r.d2[pos] = r.d2[pos - 1]; r.idx[pos] = r.idx[pos - 1]; --pos; }
        r.d2[pos] = d; r.idx[pos] = i; if (r.cnt < r.k) ++r.cnt;
    }
    void search(Res& r, const double* q, int node) const {
        const Node& nd = nodes[node];
        if (nd.left < 0) {
            for (int i = nd.lo; i < nd.hi; i++) {
                const double* p = &pts[(size_t)perm[i] * 3];
                double dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
                add(r, dx * dx + dy * dy + dz * dz, perm[i]);
            }
            return;
        }
        double diff = q[nd.dim] - nd.cut;
        int first = diff < 0 ? nd.left : nd.right, second = diff < 0 ? nd.right : nd.left;
        search(r, q, first);
        if (r.cnt < r.k || diff * diff <= r.d2[r.k - 1]) search(r, q, second);
    }
    int knn(const double* q, int k, int* idx, double* d2) const { Res r{k, 0, idx, d2}; if (n) search(r, q, 0); return r.cnt; }
};

void cov_from_normal(const double n[3], double eps, double C[9]) {
    // Rx = I + [v]x + [v]x^2 / (1 + c), v = e1 x n, c = e1 . n ; identity when c < -0.99 (Open3D GetRotationFromE1ToX)
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double c = n[0];
    if (!(c < -0.99)) {
        const double v[3] = {0.0, -n[2], n[1]};
        const double sv[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
        double sv2[9];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += sv[i * 3 + k] * sv[k * 3 + j]; sv2[i * 3 + j] = s; }
        const double f = 1 / (1 + c);
        for (int i = 0; i < 9; i++) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + sv[i] + sv2[i] * f;
    }
    const double D[3] = {eps, 1, 1};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += R[i * 3 + k] * D[k] * R[j * 3 + k]; C[i * 3 + j] = s; }
}

void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += A[i * 3 + k] * B[k * 3 + j]; C[i * 3 + j] = s; }
}

// W = (M^-1)^(1/2) for symmetric positive definite M, through its eigen-decomposition
void inv_sqrt_spd(const double M[9], double W[9]) {
    double w[3], V[9];
    jacobi3d(M, w, V);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += V[i * 3 + k] * (1.0 / std::sqrt(w[k])) * V[j * 3 + k];
        W[i * 3 + j] = s;
    }
}

bool ldlt_solve6(const double A[36], const double b[6], double x[6]) {
    double L[36] = {0}, D[6];
    for (int j = 0; j < 6; j++) {
        double d = A[j * 6 + j];
        for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
        D[j] = d;
        if (d == 0.0) return false;
        for (int i = j + 1; i < 6; i++) {
            double v = A[i * 6 + j];
            for (int k = 0; k < j; k++) v -= L[i * 6 + k] * L[j * 6 + k] * D[k];
            L[i * 6 + j] = v / d;
        }
    }
    double y[6];
    for (int i = 0; i < 6; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= L[i * 6 + k] * y[k]; y[i] = v; }
    for (int i = 0; i < 6; i++) y[i] /= D[i];
    for (int i = 5; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < 6; k++) v -= L[k * 6 + i] * x[k]; x[i] = v; }
    return true;
}

void vec6_to_mat4(const double x[6], double T[16]) {
    const double ca = std::cos(x[0]), sa = std::sin(x[0]), cb = std::cos(x[1]), sb = std::sin(x[1]), cg = std::cos(x[2]), sg = std::sin(x[2]);
    // Rz(x2) * Ry(x1) * Rx(x0)
    const double R[9] = {cg * cb, cg * sb * sa - sg * ca, cg * sb * ca + sg * sa,
                         sg * cb, sg * sb * sa + cg * ca, sg * sb * ca - cg * sa,
                         -sb, cb * sa, cb * ca};
    for (int i = 0; i < 16; i++) T[i] = 0;
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) T[i * 4 + j] = R[i * 3 + j]; T[i * 4 + 3] = x[3 + i]; }
    T[15] = 1;
}

void mat4_mul(const double* A, const double* B, double* C) {
    double t[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j]; t[i * 4 + j] = s; }
    std::memcpy(C, t, sizeof(t));
}

}  // namespace

extern "C" {

// Open3D voxel_down_sample. out capacity n. returns number of voxels; out_key (optional) = packed (z,y,x) voxel key.
int o_o3d_voxel_down_sample(const double* pts, int n, double voxel, double* out, int* voxel_of_point_rank) {
    if (n <= 0) return 0;
    double mn[3] = {pts[0], pts[1], pts[2]};
    for (int i = 1; i < n; i++) for (int d = 0; d < 3; d++) mn[d] = std::min(mn[d], pts[(size_t)i * 3 + d]);
    for (int d = 0; d < 3; d++) mn[d] -= voxel * 0.5;
    struct Acc { double s[3] = {0, 0, 0}; int cnt = 0; };
    std::map<uint64_t, Acc> vox;
    std::vector<uint64_t> keys(n);
    for (int i = 0; i < n; i++) {
        uint64_t k = 0;
        uint64_t c[3];
        for (int d = 0; d < 3; d++) c[d] = (uint64_t)(int64_t)std::floor((pts[(size_t)i * 3 + d] - mn[d]) / voxel);
        k = (c[2] << 42) | (c[1] << 21) | c[0];
        keys[i] = k;
        Acc& a = vox[k];
        for (int d = 0; d < 3; d++) a.s[d] += pts[(size_t)i * 3 + d];
        a.cnt++;
    }
    int m = 0;
    std::map<uint64_t, int> rank;
    for (auto& kv : vox) {
        for (int d = 0; d < 3; d++) out[(size_t)m * 3 + d] = kv.second.s[d] / (double)kv.second.cnt;
        rank[kv.first] = m++;
    }
    if (voxel_of_point_rank) for (int i = 0; i < n; i++) voxel_of_point_rank[i] = rank[keys[i]];
    return m;
}

// estimate_normals(KNN k) + GICP covariances. normals/covs optional outputs (n*3 / n*9).
void o_gicp_normals_covs(const double* pts, int n, int k, double eps, double* normals, double* covs, int threads) {
    KdTreeD kd; kd.build(pts, n);
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        std::vector<int> idx(k); std::vector<double> d2(k);
        int found = kd.knn(&pts[(size_t)i * 3], k, idx.data(), d2.data());
        double nrm[3] = {0, 0, 1};
        if (found >= 3) {
            // Open3D ComputeCovariance: cumulants over the neighbours, cov = E[xx^T] - mean mean^T
            double c[9] = {0};
            for (int j = 0; j < found; j++) {
                const double* p = &pts[(size_t)idx[j] * 3];
                c[0] += p[0]; c[1] += p[1]; c[2] += p[2];
                c[3] += p[0] * p[0]; c[4] += p[0] * p[1]; c[5] += p[0] * p[2];
                c[6] += p[1] * p[1]; c[7] += p[1] * p[2]; c[8] += p[2] * p[2];
            }
            for (int q = 0; q < 9; q++) c[q] /= (double)found;
            double C[9];
            C[0] = c[3] - c[0] * c[0]; C[4] = c[6] - c[1] * c[1]; C[8] = c[8] - c[2] * c[2];
            C[1] = C[3] = c[4] - c[0] * c[1]; C[2] = C[6] = c[5] - c[0] * c[2]; C[5] = C[7] = c[7] - c[1] * c[2];
            double w[3], V[9];
            jacobi3d(C, w, V);
            nrm[0] = V[0]; nrm[1] = V[3]; nrm[2] = V[6];
            double nn = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
            if (nn == 0.0) { nrm[0] = 0; nrm[1] = 0; nrm[2] = 1; }
            // sign convention pinned by the oracle: first non-zero of (z, y, x) is positive
            if (nrm[2] < 0 || (nrm[2] == 0 && (nrm[1] < 0 || (nrm[1] == 0 && nrm[0] < 0)))) { nrm[0] = -nrm[0]; nrm[1] = -nrm[1]; nrm[2] = -nrm[2]; }
        }
        if (normals) std::memcpy(&normals[(size_t)i * 3], nrm, 24);
        if (covs) cov_from_normal(nrm, eps, &covs[(size_t)i * 9]);
    }
}

// One linearisation at transform T (4x4 row-major): correspondences + sums.
// sums[30] = JtJ upper (21, row-major r<=c), Jtr (6), n_corr, sum d^2 (point distance), sum r^2.
// corr (optional, n_src ints): target index or -1.
void o_gicp_linearize(const double* src, const double* src_cov, int ns, const double* tgt, const double* tgt_cov, int nt,
                      const void* tgt_tree, const double* T, double max_corr, double* sums, int* corr, int threads) {
    const KdTreeD& kd = *(const KdTreeD*)tgt_tree;
    const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    const double t[3] = {T[3], T[7], T[11]};
    std::vector<double> rows((size_t)ns * 30, 0.0);
    std::vector<int> cidx(ns, -1);
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 1024)
    for (int i = 0; i < ns; i++) {
        const double* p = &src[(size_t)i * 3];
        double vs[3];
        for (int r = 0; r < 3; r++) vs[r] = R[r * 3] * p[0] + R[r * 3 + 1] * p[1] + R[r * 3 + 2] * p[2] + t[r];
        int j; double d2;
        if (kd.knn(vs, 1, &j, &d2) < 1 || !(d2 < max_corr * max_corr)) continue;      // radius search is strict
        cidx[i] = j;
        double RC[9], Cs[9], Rt[9];
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Rt[a * 3 + b] = R[b * 3 + a];
        mat3_mul(R, &src_cov[(size_t)i * 9], RC);
        mat3_mul(RC, Rt, Cs);
        double M[9], W[9];
        for (int q = 0; q < 9; q++) M[q] = tgt_cov[(size_t)j * 9 + q] + Cs[q];
        inv_sqrt_spd(M, W);
        const double* vt = &tgt[(size_t)j * 3];
        const double d[3] = {vs[0] - vt[0], vs[1] - vt[1], vs[2] - vt[2]};
        // J = W [ -[vs]x | I ]
        const double S[9] = {0, vs[2], -vs[1], -vs[2], 0, vs[0], vs[1], -vs[0], 0};     // -skew(vs)
        double WS[9]; mat3_mul(W, S, WS);
        double* out = &rows[(size_t)i * 30];
        for (int r = 0; r < 3; r++) {
            const double Jr[6] = {WS[r * 3], WS[r * 3 + 1], WS[r * 3 + 2], W[r * 3], W[r * 3 + 1], W[r * 3 + 2]};
            const double res = W[r * 3] * d[0] + W[r * 3 + 1] * d[1] + W[r * 3 + 2] * d[2];
            int q = 0;
            for (int a = 0; a < 6; a++) for (int b = a; b < 6; b++) out[q++] += Jr[a] * Jr[b];
            for (int a = 0; a < 6; a++) out[21 + a] += Jr[a] * res;
            out[29] += res * res;
        }
        out[27] = 1.0; out[28] = d2;
    }
    for (int q = 0; q < 30; q++) sums[q] = 0.0;
    for (int i = 0; i < ns; i++) if (cidx[i] >= 0) for (int q = 0; q < 30; q++) sums[q] += rows[(size_t)i * 30 + q];
    if (corr) std::memcpy(corr, cidx.data(), (size_t)ns * sizeof(int));
    (void)nt;
}

void* o_gicp_tree_create(const double* tgt, int nt) { KdTreeD* k = new KdTreeD(); k->build(tgt, nt); return k; }
void o_gicp_tree_destroy(void* t) { delete (KdTreeD*)t; }

// x = LDLT(JtJ)^-1 (-Jtr); update = [Rz Ry Rx | t]. returns 0 when the solve fails (identity update)
int o_gicp_solve_update(const double* sums, double* update) {
    double A[36], b[6], x[6];
    int q = 0;
    for (int r = 0; r < 6; r++) for (int c = r; c < 6; c++) { A[r * 6 + c] = sums[q]; A[c * 6 + r] = sums[q]; q++; }
    for (int r = 0; r < 6; r++) b[r] = -sums[21 + r];
    if (sums[27] < 1.0 || !ldlt_solve6(A, b, x)) { for (int i = 0; i < 16; i++) update[i] = (i % 5 == 0) ? 1.0 : 0.0; return 0; }
    vec6_to_mat4(x, update);
    return 1;
}

// The whole registration_generalized_icp loop. covariances are given in the clouds' own frames.
int o_gicp_register(const double* src, const double* src_cov, int ns, const double* tgt, const double* tgt_cov, int nt,
                    const double* init, double max_corr, double rel_fit, double rel_rmse, int max_it,
                    double* T_out, double* fitness, double* rmse, int threads) {
    KdTreeD kd; kd.build(tgt, nt);
    double T[16]; std::memcpy(T, init, sizeof(T));
    double sums[30];
    o_gicp_linearize(src, src_cov, ns, tgt, tgt_cov, nt, &kd, T, max_corr, sums, nullptr, threads);
    double fit = ns ? sums[27] / ns : 0.0, rm = sums[27] > 0 ? std::sqrt(sums[28] / sums[27]) : 0.0;
    int it = 0;
    for (; it < max_it; it++) {
        double U[16];
        o_gicp_solve_update(sums, U);
        mat4_mul(U, T, T);
        const double bf = fit, br = rm;
        o_gicp_linearize(src, src_cov, ns, tgt, tgt_cov, nt, &kd, T, max_corr, sums, nullptr, threads);
        fit = ns ? sums[27] / ns : 0.0; rm = sums[27] > 0 ? std::sqrt(sums[28] / sums[27]) : 0.0;
        if (std::fabs(bf - fit) < rel_fit && std::fabs(br - rm) < rel_rmse) { ++it; break; }
    }
    std::memcpy(T_out, T, sizeof(T));
    *fitness = fit; *rmse = rm;
    return it;
}

}  // extern "C"
