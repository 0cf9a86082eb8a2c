// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing in the product path may include, link or call this.
//
// CPU restatement of the small dense-math library routines the reference's hot path calls.
// None of these libraries is vendored under /root/reference, so each routine restates the
// *published algorithm* of the pinned generation and is pinned by tests against
// cv2 4.13 (cv2.eigen / cv2.solve(DECOMP_QR) / cv2.invert) in tests/test_oracle_math.py.
//
//   jacobi_eigen_f32   <- cv::eigen on CV_32F symmetric (OpenCV "JacobiImpl_"), used at
//                         liosam_ws/src/LIO-SAM/src/mapOptmization.cpp:1018 (3x3) and :1235 (6x6)
//   qr_solve_f32       <- cv::solve(..., DECOMP_QR) (OpenCV hal "QRImpl"), mapOptmization.cpp:1227
//   lu_invert_f32      <- cv::Mat::inv() default DECOMP_LU (OpenCV hal "LUImpl"), mapOptmization.cpp:1250
//   gemm_f32_dacc      <- cv::gemm on CV_32F: double accumulation, float store, mapOptmization.cpp:1225-1226,1250,1257
//   colpiv_qr_solve_5x3<- Eigen::ColPivHouseholderQR<Matrix<float,5,3>>::solve, mapOptmization.cpp:1096
//   pcl_get_transformation <- pcl::getTransformation(x,y,z,roll,pitch,yaw) (PCL common/eigen.hpp)
//
// Build: -O3, no -ffast-math, -ffp-contract=off (the reference builds for baseline x86-64:
// liosam_ws/src/LIO-SAM/CMakeLists.txt:4-6, so no FMA contraction anywhere).
#pragma once
#include <cmath>
#include <cfloat>
#include <cstring>
#include <algorithm>
#include <utility>

namespace orc {

// OpenCV's own scaled hypot (lapack.cpp), float: max*sqrt(1 + (min/max)^2).
static inline float hypot_f32(float a, float b) {
    a = std::fabs(a);
    b = std::fabs(b);
    if (a > b) {
        b /= a;
        return a * std::sqrt(1 + b * b);
    }
    if (b > 0) {
        a /= b;
        return b * std::sqrt(1 + a * a);
    }
    return 0;
}

// Symmetric eigen-decomposition, cyclic-by-pivot Jacobi, float. A is n x n row-major (destroyed).
// W: eigenvalues descending; V: eigenvectors as ROWS (V[k*n + i] is component i of eigenvector k).
template <int N>
static inline void jacobi_eigen_f32(float* A, float* W, float* V) {
    const int n = N;
    const float eps = FLT_EPSILON;
    int indR[N], indC[N];
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) V[i * n + j] = 0.f;
        V[i * n + i] = 1.f;
    }
    float mv = 0.f;
    int m;
    for (int k = 0; k < n; k++) {
        W[k] = A[(n + 1) * k];
        if (k < n - 1) {
            m = k + 1; mv = std::fabs(A[n * k + m]);
            for (int i = k + 2; i < n; i++) {
                float val = std::fabs(A[n * k + i]);
                if (mv < val) { mv = val; m = i; }
            }
            indR[k] = m;
        }
        if (k > 0) {
            m = 0; mv = std::fabs(A[k]);
            for (int i = 1; i < k; i++) {
                float val = std::fabs(A[n * i + k]);
                if (mv < val) { mv = val; m = i; }
            }
            indC[k] = m;
        }
    }
    if (n > 1) {
        const int maxIters = n * n * 30;
        for (int iters = 0; iters < maxIters; iters++) {
            int k = 0; mv = std::fabs(A[indR[0]]);
            for (int i = 1; i < n - 1; i++) {
                float val = std::fabs(A[n * i + indR[i]]);
                if (mv < val) { mv = val; k = i; }
            }
            int l = indR[k];
            for (int i = 1; i < n; i++) {
                float val = std::fabs(A[n * indC[i] + i]);
                if (mv < val) { mv = val; k = indC[i]; l = i; }
            }
            float p = A[n * k + l];
            if (std::fabs(p) <= eps) break;
            float y = (float)((W[l] - W[k]) * 0.5);
            float t = std::fabs(y) + hypot_f32(p, y);
            float s = hypot_f32(p, t);
            float c = t / s;
            s = p / s; t = (p / t) * p;
            if (y < 0) { s = -s; t = -t; }
            A[n * k + l] = 0;
            W[k] -= t; W[l] += t;
            float a0, b0;
#define ORC_ROT(v0, v1) do { a0 = (v0); b0 = (v1); (v0) = a0 * c - b0 * s; (v1) = a0 * s + b0 * c; } while (0)
            for (int i = 0; i < k; i++) ORC_ROT(A[n * i + k], A[n * i + l]);
            for (int i = k + 1; i < l; i++) ORC_ROT(A[n * k + i], A[n * i + l]);
            for (int i = l + 1; i < n; i++) ORC_ROT(A[n * k + i], A[n * l + i]);
            for (int i = 0; i < n; i++) ORC_ROT(V[n * k + i], V[n * l + i]);
#undef ORC_ROT
            for (int j = 0; j < 2; j++) {
                int idx = j == 0 ? k : l;
                if (idx < n - 1) {
                    m = idx + 1; mv = std::fabs(A[n * idx + m]);
                    for (int i = idx + 2; i < n; i++) {
                        float val = std::fabs(A[n * idx + i]);
                        if (mv < val) { mv = val; m = i; }
                    }
                    indR[idx] = m;
                }
                if (idx > 0) {
                    m = 0; mv = std::fabs(A[idx]);
                    for (int i = 1; i < idx; i++) {
                        float val = std::fabs(A[n * i + idx]);
                        if (mv < val) { mv = val; m = i; }
                    }
                    indC[idx] = m;
                }
            }
        }
    }
    for (int k = 0; k < n - 1; k++) {
        m = k;
        for (int i = k + 1; i < n; i++) if (W[m] < W[i]) m = i;
        if (k != m) {
            std::swap(W[m], W[k]);
            for (int i = 0; i < n; i++) std::swap(V[n * m + i], V[n * k + i]);
        }
    }
}

// Householder-QR solve of a square n x n system (single right-hand side), float, in place.
// Returns 0 when a diagonal of R is below eps (singular), else 1. A and b are destroyed; x in b.
template <int N>
static inline int qr_solve_f32(float* A, float* b) {
    const int m = N, n = N;
    const float eps = FLT_EPSILON * 10;   // OpenCV passes FLT_EPSILON*10 to its QR kernel
    float vl[N], hF[N];
    for (int l = 0; l < n; l++) {
        int vlSize = m - l;
        float vlNorm = 0.f;
        for (int i = 0; i < vlSize; i++) { vl[i] = A[(l + i) * n + l]; vlNorm += vl[i] * vl[i]; }
        float tmpV = vl[0];
        vl[0] = vl[0] + ((vl[0] >= 0.0f) ? 1 : -1) * std::sqrt(vlNorm);
        vlNorm = std::sqrt(vlNorm + vl[0] * vl[0] - tmpV * tmpV);
        for (int i = 0; i < vlSize; i++) vl[i] /= vlNorm;
        for (int j = l; j < n; j++) {
            float v_lA = 0.f;
            for (int i = l; i < m; i++) v_lA += vl[i - l] * A[i * n + j];
            for (int i = l; i < m; i++) A[i * n + j] -= 2 * vl[i - l] * v_lA;
        }
        hF[l] = vl[0] * vl[0];
        for (int i = 1; i < vlSize; i++) A[(l + i) * n + l] = vl[i] / vl[0];
    }
    for (int l = 0; l < n; l++) {
        vl[0] = 1.f;
        for (int j = 1; j < m - l; j++) vl[j] = A[(j + l) * n + l];
        float v_lB = 0.f;
        for (int i = l; i < m; i++) v_lB += vl[i - l] * b[i];
        for (int i = l; i < m; i++) b[i] -= 2 * vl[i - l] * v_lB * hF[l];
    }
    for (int i = n - 1; i >= 0; i--) {
        for (int j = n - 1; j > i; j--) b[i] -= b[j] * A[i * n + j];
        if (std::fabs(A[i * n + i]) < eps) return 0;
        b[i] /= A[i * n + i];
    }
    return 1;
}

// Inverse by LU with partial pivoting against an identity right-hand side, float.
// Returns 0 (and leaves Ainv zeroed, like cv::invert on failure) when a pivot is below eps.
template <int N>
static inline int lu_invert_f32(const float* Ain, float* Ainv) {
    const int m = N;
    const float eps = FLT_EPSILON * 10;
    float A[N * N];
    std::memcpy(A, Ain, sizeof(A));
    float* b = Ainv;
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) b[i * m + j] = (i == j) ? 1.f : 0.f;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (std::fabs(A[j * m + i]) > std::fabs(A[k * m + i])) k = j;
        if (std::fabs(A[k * m + i]) < eps) { std::memset(Ainv, 0, sizeof(float) * N * N); return 0; }
        if (k != i) {
            for (int j = i; j < m; j++) std::swap(A[i * m + j], A[k * m + j]);
            for (int j = 0; j < m; j++) std::swap(b[i * m + j], b[k * m + j]);
        }
        float d = -1 / A[i * m + i];
        for (int j = i + 1; j < m; j++) {
            float alpha = A[j * m + i] * d;
            for (int kk = i + 1; kk < m; kk++) A[j * m + kk] += alpha * A[i * m + kk];
            for (int kk = 0; kk < m; kk++) b[j * m + kk] += alpha * b[i * m + kk];
        }
    }
    for (int i = m - 1; i >= 0; i--)
        for (int j = 0; j < m; j++) {
            float s = b[i * m + j];
            for (int k = i + 1; k < m; k++) s -= A[i * m + k] * b[k * m + j];
            b[i * m + j] = s / A[i * m + i];
        }
    return 1;
}

// C(MxN) = A(MxK) * B(KxN), float operands, products and sums in double, one rounding on store.
static inline void gemm_f32_dacc(const float* A, const float* B, float* C, int M, int K, int N) {
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            double s = 0.0;
            for (int k = 0; k < K; k++) s += (double)A[i * K + k] * (double)B[k * N + j];
            C[i * N + j] = (float)s;
        }
}

// Least squares min |A x - b| for A 5x3 by Householder QR with column pivoting (float).
// A is row-major 5x3. Restates the Eigen 3.3 algorithm (largest-updated-norm pivot, LAPACK WN176
// norm downdate, nonzero-pivot threshold). Sums are plain left-to-right; Eigen's packet order
// differs, so agreement with Eigen itself is tolerance-level (:1081-1096).
static inline void colpiv_qr_solve_5x3(const float* Ain, const float* bin, float* x) {
    const int rows = 5, cols = 3, size = 3;
    float qr[5][3];
    for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) qr[i][j] = Ain[i * 3 + j];
    float hCoeffs[3], normsUpdated[3], normsDirect[3];
    int transp[3];
    for (int k = 0; k < cols; k++) {
        float s = 0.f;
        for (int i = 0; i < rows; i++) s += qr[i][k] * qr[i][k];
        normsDirect[k] = std::sqrt(s);
        normsUpdated[k] = normsDirect[k];
    }
    float maxn = std::max(normsUpdated[0], std::max(normsUpdated[1], normsUpdated[2]));
    float th = maxn * FLT_EPSILON;
    const float threshold_helper = (th * th) / (float)rows;
    const float norm_downdate_threshold = std::sqrt(FLT_EPSILON);
    int nonzero_pivots = size;
    for (int k = 0; k < size; k++) {
        int big = k; float bigv = normsUpdated[k];
        for (int j = k + 1; j < cols; j++) if (normsUpdated[j] > bigv) { bigv = normsUpdated[j]; big = j; }
        float big_sq = bigv * bigv;
        if (nonzero_pivots == size && big_sq < threshold_helper * (float)(rows - k)) nonzero_pivots = k;
        transp[k] = big;
        if (k != big) {
            for (int i = 0; i < rows; i++) std::swap(qr[i][k], qr[i][big]);
            std::swap(normsUpdated[k], normsUpdated[big]);
            std::swap(normsDirect[k], normsDirect[big]);
        }
        // Householder vector of column k, rows k..4
        float tailSq = 0.f;
        for (int i = k + 1; i < rows; i++) tailSq += qr[i][k] * qr[i][k];
        float c0 = qr[k][k];
        float tau, beta;
        if (tailSq <= FLT_MIN) {
            tau = 0.f; beta = c0;
            for (int i = k + 1; i < rows; i++) qr[i][k] = 0.f;
        } else {
            beta = std::sqrt(c0 * c0 + tailSq);
            if (c0 >= 0.f) beta = -beta;
            float den = c0 - beta;
            for (int i = k + 1; i < rows; i++) qr[i][k] = qr[i][k] / den;
            tau = (beta - c0) / beta;
        }
        hCoeffs[k] = tau;
        qr[k][k] = beta;
        // apply H = I - tau v v^T (v = [1, essential]) to the trailing columns
        if (tau != 0.f) {
            for (int j = k + 1; j < cols; j++) {
                float tmp = 0.f;
                for (int i = k + 1; i < rows; i++) tmp += qr[i][k] * qr[i][j];
                tmp += qr[k][j];
                qr[k][j] -= tau * tmp;
                for (int i = k + 1; i < rows; i++) qr[i][j] -= tau * qr[i][k] * tmp;
            }
        }
        for (int j = k + 1; j < cols; j++) {
            if (normsUpdated[j] != 0.f) {
                float temp = std::fabs(qr[k][j]) / normsUpdated[j];
                temp = (1.f + temp) * (1.f - temp);
                temp = temp < 0.f ? 0.f : temp;
                float r = normsUpdated[j] / normsDirect[j];
                float temp2 = temp * (r * r);
                if (temp2 <= norm_downdate_threshold) {
                    float s = 0.f;
                    for (int i = k + 1; i < rows; i++) s += qr[i][j] * qr[i][j];
                    normsDirect[j] = std::sqrt(s);
                    normsUpdated[j] = normsDirect[j];
                } else {
                    normsUpdated[j] *= std::sqrt(temp);
                }
            }
        }
    }
    // column permutation indices: apply transpositions in order to the identity
    int perm[3] = {0, 1, 2};
    for (int k = 0; k < size; k++) std::swap(perm[k], perm[transp[k]]);
    float c[5];
    for (int i = 0; i < rows; i++) c[i] = bin[i];
    x[0] = x[1] = x[2] = 0.f;
    if (nonzero_pivots == 0) return;
    // c = Q^T b : apply H_0, H_1, ... in order
    for (int k = 0; k < nonzero_pivots; k++) {
        float tau = hCoeffs[k];
        if (tau == 0.f) continue;
        float tmp = 0.f;
        for (int i = k + 1; i < rows; i++) tmp += qr[i][k] * c[i];
        tmp += c[k];
        c[k] -= tau * tmp;
        for (int i = k + 1; i < rows; i++) c[i] -= tau * qr[i][k] * tmp;
    }
    // back substitution on the leading nonzero_pivots block
    for (int i = nonzero_pivots - 1; i >= 0; i--) {
        float s = c[i];
        for (int j = i + 1; j < nonzero_pivots; j++) s -= qr[i][j] * c[j];
        c[i] = s / qr[i][i];
    }
    for (int i = 0; i < nonzero_pivots; i++) x[perm[i]] = c[i];
}

// pcl::getTransformation(x, y, z, roll, pitch, yaw): rows of the 3x4 affine, float throughout.
// Follows PCL common/impl/eigen.hpp (1.8-1.10): DE = D*E and DF = D*F are formed first.
static inline void pcl_get_transformation(float x, float y, float z, float roll, float pitch, float yaw,
                                          float t[12]) {
    float A = std::cos(yaw), B = std::sin(yaw), C = std::cos(pitch), D = std::sin(pitch),
          E = std::cos(roll), F = std::sin(roll), DE = D * E, DF = D * F;
    t[0] = A * C;  t[1] = A * DF - B * E;  t[2]  = B * F + A * DE;  t[3]  = x;
    t[4] = B * C;  t[5] = A * E + B * DF;  t[6]  = B * DE - A * F;  t[7]  = y;
    t[8] = -D;     t[9] = C * F;           t[10] = C * E;           t[11] = z;
}

}  // namespace orc
