// Small dense solvers used inside the scan-to-map kernels, in the operation order of the library routines the
// reference calls (so that accept/reject thresholds see the same bits as the CPU path):
//   sym_eigen_jacobi<N>   cv::eigen on CV_32F        mapOptmization.cpp:1018 (3x3), :1235 (6x6)
//   solve_householder<N>  cv::solve(DECOMP_QR)       mapOptmization.cpp:1227
//   invert_lu<N>          cv::Mat::inv() (LU)        mapOptmization.cpp:1250
//   plane_lsq_5x3         Eigen colPivHouseholderQr  mapOptmization.cpp:1096
// All float, IEEE div/sqrt, no FMA (-fmad=false). Host+device so the host epilogue tests can reuse them.
#pragma once
#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b2 {

B2_HD float scaled_hypot(float a, float b) {
    a = fabsf(a); b = fabsf(b);
    if (a > b) { b /= a; return a * sqrtf(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrtf(1 + a * a); }
    return 0.f;
}

// largest |a(row, j)| for j in (row, N): index of first maximum
template <int N>
B2_HD int argmax_row(const float* a, int row) {
    int m = row + 1; float mv = fabsf(a[N * row + m]);
    for (int i = row + 2; i < N; i++) { float v = fabsf(a[N * row + i]); if (mv < v) { mv = v; m = i; } }
    return m;
}
// largest |a(i, col)| for i in [0, col): index of first maximum
template <int N>
B2_HD int argmax_col(const float* a, int col) {
    int m = 0; float mv = fabsf(a[col]);
    for (int i = 1; i < col; i++) { float v = fabsf(a[N * i + col]); if (mv < v) { mv = v; m = i; } }
    return m;
}

// Jacobi eigen-decomposition of a symmetric NxN (row-major, destroyed). w descending, v rows = eigenvectors.
template <int N>
B2_HD void sym_eigen_jacobi(float* a, float* w, float* v) {
    int rmax[N], cmax[N];
    for (int i = 0; i < N; i++) { for (int j = 0; j < N; j++) v[i * N + j] = (i == j) ? 1.f : 0.f; }
    for (int k = 0; k < N; k++) {
        w[k] = a[(N + 1) * k];
        if (k < N - 1) rmax[k] = argmax_row<N>(a, k);
        if (k > 0) cmax[k] = argmax_col<N>(a, k);
    }
    const int max_rot = N * N * 30;
    for (int it = 0; it < max_rot; it++) {
        int k = 0; float mv = fabsf(a[rmax[0]]);
        for (int i = 1; i < N - 1; i++) { float val = fabsf(a[N * i + rmax[i]]); if (mv < val) { mv = val; k = i; } }
        int l = rmax[k];
        for (int i = 1; i < N; i++) { float val = fabsf(a[N * cmax[i] + i]); if (mv < val) { mv = val; k = cmax[i]; l = i; } }
        float p = a[N * k + l];
        if (fabsf(p) <= FLT_EPSILON) break;
        float y = (float)((w[l] - w[k]) * 0.5);
        float t = fabsf(y) + scaled_hypot(p, y);
        float s = scaled_hypot(p, t);
        float c = t / s;
        s = p / s; t = (p / t) * p;
        if (y < 0) { s = -s; t = -t; }
        a[N * k + l] = 0;
        w[k] -= t; w[l] += t;
        for (int i = 0; i < k; i++)     { float a0 = a[N * i + k], b0 = a[N * i + l]; a[N * i + k] = a0 * c - b0 * s; a[N * i + l] = a0 * s + b0 * c; }
        for (int i = k + 1; i < l; i++) { float a0 = a[N * k + i], b0 = a[N * i + l]; a[N * k + i] = a0 * c - b0 * s; a[N * i + l] = a0 * s + b0 * c; }
        for (int i = l + 1; i < N; i++) { float a0 = a[N * k + i], b0 = a[N * l + i]; a[N * k + i] = a0 * c - b0 * s; a[N * l + i] = a0 * s + b0 * c; }
        for (int i = 0; i < N; i++)     { float a0 = v[N * k + i], b0 = v[N * l + i]; v[N * k + i] = a0 * c - b0 * s; v[N * l + i] = a0 * s + b0 * c; }
        for (int j = 0; j < 2; j++) {
            int idx = j == 0 ? k : l;
            if (idx < N - 1) rmax[idx] = argmax_row<N>(a, idx);
            if (idx > 0) cmax[idx] = argmax_col<N>(a, idx);
        }
    }
    for (int k = 0; k < N - 1; k++) {
        int m = k;
        for (int i = k + 1; i < N; i++) if (w[m] < w[i]) m = i;
        if (k != m) {
            float tw = w[m]; w[m] = w[k]; w[k] = tw;
            for (int i = 0; i < N; i++) { float tv = v[N * m + i]; v[N * m + i] = v[N * k + i]; v[N * k + i] = tv; }
        }
    }
}

// 3x3 specialisation of sym_eigen_jacobi with every element in a named register. The pivot bookkeeping of the
// generic routine (row maxima rmax[], column maxima cmax[], refreshed only for the two rotated indices) is
// reproduced literally: for N = 3 it reduces to rmax[0] in {1,2} and cmax[2] in {0,1} (rmax[1] = 2, cmax[1] = 0).
// Same operations in the same order as sym_eigen_jacobi<3>, so the results are bit-identical.
B2_HD void sym_eigen_jacobi3(float a00, float a01, float a02, float a11, float a12, float a22, float (&w)[3], float (&v)[9]) {
    float w0 = a00, w1 = a11, w2 = a22;
    float v00 = 1.f, v01 = 0.f, v02 = 0.f, v10 = 0.f, v11 = 1.f, v12 = 0.f, v20 = 0.f, v21 = 0.f, v22 = 1.f;
    int r0 = (fabsf(a01) < fabsf(a02)) ? 2 : 1;
    int c2 = (fabsf(a02) < fabsf(a12)) ? 1 : 0;
    for (int it = 0; it < 3 * 3 * 30; it++) {
        // pivot search
        int k = 0;
        float mv = fabsf(r0 == 1 ? a01 : a02);
        { float val = fabsf(a12); if (mv < val) { mv = val; k = 1; } }
        int l = (k == 0) ? r0 : 2;
        { float val = fabsf(a01); if (mv < val) { mv = val; k = 0; l = 1; } }
        { float val = fabsf(c2 == 0 ? a02 : a12); if (mv < val) { mv = val; k = c2; l = 2; } }
        const int pair = (k == 0) ? (l == 1 ? 0 : 1) : 2;            // (0,1) (0,2) (1,2)
        const float p = pair == 0 ? a01 : (pair == 1 ? a02 : a12);
        if (fabsf(p) <= FLT_EPSILON) break;
        const float wk = (k == 0) ? w0 : w1, wl = (l == 1) ? w1 : w2;
        float y = (float)((wl - wk) * 0.5);
        float t = fabsf(y) + scaled_hypot(p, y);
        float s = scaled_hypot(p, t);
        float c = t / s;
        s = p / s; t = (p / t) * p;
        if (y < 0) { s = -s; t = -t; }
#define B2_ROT(x, z) do { float a0_ = (x), b0_ = (z); (x) = a0_ * c - b0_ * s; (z) = a0_ * s + b0_ * c; } while (0)
        if (pair == 0) {
            a01 = 0; w0 -= t; w1 += t;
            B2_ROT(a02, a12);
            B2_ROT(v00, v10); B2_ROT(v01, v11); B2_ROT(v02, v12);
            r0 = (fabsf(a01) < fabsf(a02)) ? 2 : 1;      // idx 0: row max ; idx 1: rmax[1] = 2, cmax[1] = 0 fixed
        } else if (pair == 1) {
            a02 = 0; w0 -= t; w2 += t;
            B2_ROT(a01, a12);
            B2_ROT(v00, v20); B2_ROT(v01, v21); B2_ROT(v02, v22);
            r0 = (fabsf(a01) < fabsf(a02)) ? 2 : 1;      // idx 0: row max
            c2 = (fabsf(a02) < fabsf(a12)) ? 1 : 0;      // idx 2: column max
        } else {
            a12 = 0; w1 -= t; w2 += t;
            B2_ROT(a01, a02);
            B2_ROT(v10, v20); B2_ROT(v11, v21); B2_ROT(v12, v22);
            c2 = (fabsf(a02) < fabsf(a12)) ? 1 : 0;      // idx 2: column max (idx 1 entries are fixed)
        }
#undef B2_ROT
    }
    // selection sort, descending, rows of v follow
#define B2_SWAPROW(wa, wb, ra0, ra1, ra2, rb0, rb1, rb2) do { float t_ = wa; wa = wb; wb = t_; t_ = ra0; ra0 = rb0; rb0 = t_; \
        t_ = ra1; ra1 = rb1; rb1 = t_; t_ = ra2; ra2 = rb2; rb2 = t_; } while (0)
    {
        int m = 0;
        if (w0 < w1) m = 1;
        if ((m == 0 ? w0 : w1) < w2) m = 2;
        if (m == 1) B2_SWAPROW(w1, w0, v10, v11, v12, v00, v01, v02);
        else if (m == 2) B2_SWAPROW(w2, w0, v20, v21, v22, v00, v01, v02);
        if (w1 < w2) B2_SWAPROW(w2, w1, v20, v21, v22, v10, v11, v12);
    }
#undef B2_SWAPROW
    w[0] = w0; w[1] = w1; w[2] = w2;
    v[0] = v00; v[1] = v01; v[2] = v02; v[3] = v10; v[4] = v11; v[5] = v12; v[6] = v20; v[7] = v21; v[8] = v22;
}

// Householder QR solve of a square system, one right-hand side; a, b destroyed, x left in b. 0 = singular.
// No pivoting, so every index is a compile-time constant once the loops are unrolled: the whole factorisation
// stays in registers (a serial single-thread epilogue is latency bound, local memory would triple its time).
template <int N>
B2_HD int solve_householder(float* a, float* b) {
    const float eps = FLT_EPSILON * 10;
    float u[N], hf[N];
#pragma unroll
    for (int l = 0; l < N; l++) {
        const int len = N - l;
        float nrm = 0.f;
#pragma unroll
        for (int i = 0; i < len; i++) { u[i] = a[(l + i) * N + l]; nrm += u[i] * u[i]; }
        float head = u[0];
        u[0] = u[0] + ((u[0] >= 0.0f) ? 1 : -1) * sqrtf(nrm);
        nrm = sqrtf(nrm + u[0] * u[0] - head * head);
#pragma unroll
        for (int i = 0; i < len; i++) u[i] /= nrm;
#pragma unroll
        for (int j = l; j < N; j++) {
            float dot = 0.f;
#pragma unroll
            for (int i = l; i < N; i++) dot += u[i - l] * a[i * N + j];
#pragma unroll
            for (int i = l; i < N; i++) a[i * N + j] -= 2 * u[i - l] * dot;
        }
        hf[l] = u[0] * u[0];
#pragma unroll
        for (int i = 1; i < len; i++) a[(l + i) * N + l] = u[i] / u[0];
    }
#pragma unroll
    for (int l = 0; l < N; l++) {
        u[0] = 1.f;
#pragma unroll
        for (int j = 1; j < N - l; j++) u[j] = a[(j + l) * N + l];
        float dot = 0.f;
#pragma unroll
        for (int i = l; i < N; i++) dot += u[i - l] * b[i];
#pragma unroll
        for (int i = l; i < N; i++) b[i] -= 2 * u[i - l] * dot * hf[l];
    }
    int ok = 1;
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
#pragma unroll
        for (int j = N - 1; j > i; j--) b[i] -= b[j] * a[i * N + j];
        if (ok && fabsf(a[i * N + i]) < eps) ok = 0;
        if (ok) b[i] /= a[i * N + i];
    }
    return ok;
}

// Inverse via LU with partial pivoting applied to an identity right-hand side; a destroyed. 0 = singular (inv zeroed).
template <int N>
B2_HD int invert_lu(float* a, float* inv) {
    const float eps = FLT_EPSILON * 10;
    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) inv[i * N + j] = (i == j) ? 1.f : 0.f;
    for (int i = 0; i < N; i++) {
        int piv = i;
        for (int j = i + 1; j < N; j++) if (fabsf(a[j * N + i]) > fabsf(a[piv * N + i])) piv = j;
        if (fabsf(a[piv * N + i]) < eps) { for (int q = 0; q < N * N; q++) inv[q] = 0.f; return 0; }
        if (piv != i) {
            for (int j = i; j < N; j++) { float t = a[i * N + j]; a[i * N + j] = a[piv * N + j]; a[piv * N + j] = t; }
            for (int j = 0; j < N; j++) { float t = inv[i * N + j]; inv[i * N + j] = inv[piv * N + j]; inv[piv * N + j] = t; }
        }
        float d = -1 / a[i * N + i];
        for (int j = i + 1; j < N; j++) {
            float alpha = a[j * N + i] * d;
            for (int q = i + 1; q < N; q++) a[j * N + q] += alpha * a[i * N + q];
            for (int q = 0; q < N; q++) inv[j * N + q] += alpha * inv[i * N + q];
        }
    }
    for (int i = N - 1; i >= 0; i--)
        for (int j = 0; j < N; j++) {
            float s = inv[i * N + j];
            for (int q = i + 1; q < N; q++) s -= a[i * N + q] * inv[q * N + j];
            inv[i * N + j] = s / a[i * N + i];
        }
    return 1;
}

// C = A*B for NxN float with double accumulation and one rounding per element (cv::gemm on CV_32F)
template <int N, int NC>
B2_HD void matmul_dacc(const float* A, const float* B, float* C) {
    for (int i = 0; i < N; i++)
        for (int j = 0; j < NC; j++) {
            double s = 0.0;
            for (int q = 0; q < N; q++) s += (double)A[i * N + q] * (double)B[q * NC + j];
            C[i * NC + j] = (float)s;
        }
}

// min |M x + 1| over x for the 5x3 matrix of neighbour coordinates (rows = points): column-pivoted Householder QR
// (pivot = largest running column norm, LAPACK WN176 downdate, rank threshold eps^2*maxnorm^2/rows*(rows-k)).
B2_HD void plane_lsq_5x3(const float (&px)[5], const float (&py)[5], const float (&pz)[5], float& xa, float& xb, float& xc) {
    // Every loop below has a compile-time trip count and every index is static after unrolling: the pivot column is
    // brought in with predicated swaps instead of a data-dependent subscript, so q/tau/norms never leave registers.
    float q[5][3];
#pragma unroll
    for (int i = 0; i < 5; i++) { q[i][0] = px[i]; q[i][1] = py[i]; q[i][2] = pz[i]; }
    float tau[3], nu[3], nd[3];
    int perm[3] = {0, 1, 2};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 5; i++) s += q[i][k] * q[i][k];
        nd[k] = sqrtf(s); nu[k] = nd[k];
    }
    float mx = fmaxf(nu[0], fmaxf(nu[1], nu[2]));
    float th = mx * FLT_EPSILON;
    const float thr_helper = (th * th) / 5.0f;
    const float downdate_thr = sqrtf(FLT_EPSILON);
    int rank = 3;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int big = k; float bv = nu[k];
#pragma unroll
        for (int j = k + 1; j < 3; j++) if (nu[j] > bv) { bv = nu[j]; big = j; }
        if (rank == 3 && bv * bv < thr_helper * (float)(5 - k)) rank = k;
#pragma unroll
        for (int j = k + 1; j < 3; j++) {
            if (big == j) {
#pragma unroll
                for (int i = 0; i < 5; i++) { float t = q[i][k]; q[i][k] = q[i][j]; q[i][j] = t; }
                float t = nu[k]; nu[k] = nu[j]; nu[j] = t;
                t = nd[k]; nd[k] = nd[j]; nd[j] = t;
                int tp = perm[k]; perm[k] = perm[j]; perm[j] = tp;
            }
        }
        float tail = 0.f;
#pragma unroll
        for (int i = k + 1; i < 5; i++) tail += q[i][k] * q[i][k];
        float c0 = q[k][k], beta, tk;
        if (tail <= FLT_MIN) {
            tk = 0.f; beta = c0;
#pragma unroll
            for (int i = k + 1; i < 5; i++) q[i][k] = 0.f;
        } else {
            beta = sqrtf(c0 * c0 + tail);
            if (c0 >= 0.f) beta = -beta;
            float den = c0 - beta;
#pragma unroll
            for (int i = k + 1; i < 5; i++) q[i][k] = q[i][k] / den;
            tk = (beta - c0) / beta;
        }
        tau[k] = tk; q[k][k] = beta;
        if (tk != 0.f) {
#pragma unroll
            for (int j = k + 1; j < 3; j++) {
                float t = 0.f;
#pragma unroll
                for (int i = k + 1; i < 5; i++) t += q[i][k] * q[i][j];
                t += q[k][j];
                q[k][j] -= tk * t;
#pragma unroll
                for (int i = k + 1; i < 5; i++) q[i][j] -= tk * q[i][k] * t;
            }
        }
#pragma unroll
        for (int j = k + 1; j < 3; j++) {
            if (nu[j] != 0.f) {
                float t = fabsf(q[k][j]) / nu[j];
                t = (1.f + t) * (1.f - t);
                t = t < 0.f ? 0.f : t;
                float r = nu[j] / nd[j];
                float t2 = t * (r * r);
                if (t2 <= downdate_thr) {
                    float s = 0.f;
#pragma unroll
                    for (int i = k + 1; i < 5; i++) s += q[i][j] * q[i][j];
                    nd[j] = sqrtf(s); nu[j] = nd[j];
                } else {
                    nu[j] *= sqrtf(t);
                }
            }
        }
    }
    float c[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (k < rank && tau[k] != 0.f) {
            float t = 0.f;
#pragma unroll
            for (int i = k + 1; i < 5; i++) t += q[i][k] * c[i];
            t += c[k];
            c[k] -= tau[k] * t;
#pragma unroll
            for (int i = k + 1; i < 5; i++) c[i] -= tau[k] * q[i][k] * t;
        }
    }
#pragma unroll
    for (int i = 2; i >= 0; i--) {
        if (i < rank) {
            float s = c[i];
#pragma unroll
            for (int j = i + 1; j < 3; j++) if (j < rank) s -= q[i][j] * c[j];
            c[i] = s / q[i][i];
        }
    }
    // x[perm[i]] = c[i] for i < rank, zero elsewhere — as selects
    float x[3];
#pragma unroll
    for (int o = 0; o < 3; o++) {
        float val = 0.f;
#pragma unroll
        for (int i = 0; i < 3; i++) if (i < rank && perm[i] == o) val = c[i];
        x[o] = val;
    }
    xa = x[0]; xb = x[1]; xc = x[2];
}


// ---------------------------------------------------------------------------------------------------------------------
// Compact variants: same operations in the same order as the routines above, written as rolled loops over small
// per-thread arrays and never inlined. On the latency-critical path of one LM iteration a single warp walks this code
// once per launch; measured on B200 that path is bound by instruction fetch (stall_no_inst, ~37 cycles per SASS
// instruction for cold straight-line code), so code bytes matter more than register residency there.
#if defined(__CUDACC__)
#define B2_NOINLINE __device__ __noinline__
#define B2_ROLL _Pragma("unroll 1")

template <int N>
B2_NOINLINE void sym_eigen_jacobi_c(float* a, float* w, float* v) {
    int rmax[N], cmax[N];
    B2_ROLL for (int i = 0; i < N * N; i++) v[i] = 0.f;
    B2_ROLL for (int i = 0; i < N; i++) v[i * N + i] = 1.f;
    B2_ROLL for (int k = 0; k < N; k++) {
        w[k] = a[(N + 1) * k];
        if (k < N - 1) {
            int m = k + 1; float mv = fabsf(a[N * k + m]);
            B2_ROLL for (int i = k + 2; i < N; i++) { float val = fabsf(a[N * k + i]); if (mv < val) { mv = val; m = i; } }
            rmax[k] = m;
        }
        if (k > 0) {
            int m = 0; float mv = fabsf(a[k]);
            B2_ROLL for (int i = 1; i < k; i++) { float val = fabsf(a[N * i + k]); if (mv < val) { mv = val; m = i; } }
            cmax[k] = m;
        }
    }
    B2_ROLL for (int it = 0; it < N * N * 30; it++) {
        int k = 0; float mv = fabsf(a[rmax[0]]);
        B2_ROLL for (int i = 1; i < N - 1; i++) { float val = fabsf(a[N * i + rmax[i]]); if (mv < val) { mv = val; k = i; } }
        int l = rmax[k];
        B2_ROLL for (int i = 1; i < N; i++) { float val = fabsf(a[N * cmax[i] + i]); if (mv < val) { mv = val; k = cmax[i]; l = i; } }
        float p = a[N * k + l];
        if (fabsf(p) <= FLT_EPSILON) break;
        float y = (float)((w[l] - w[k]) * 0.5);
        float t = fabsf(y) + scaled_hypot(p, y);
        float s = scaled_hypot(p, t);
        float c = t / s;
        s = p / s; t = (p / t) * p;
        if (y < 0) { s = -s; t = -t; }
        a[N * k + l] = 0;
        w[k] -= t; w[l] += t;
        // the three rotation ranges of the upper triangle, as one loop over i != k, l
        B2_ROLL for (int i = 0; i < N; i++) {
            if (i == k || i == l) continue;
            float* x = (i < k) ? &a[N * i + k] : &a[N * k + i];
            float* z = (i < l) ? &a[N * i + l] : &a[N * l + i];
            float a0 = *x, b0 = *z;
            *x = a0 * c - b0 * s; *z = a0 * s + b0 * c;
        }
        B2_ROLL for (int i = 0; i < N; i++) { float a0 = v[N * k + i], b0 = v[N * l + i]; v[N * k + i] = a0 * c - b0 * s; v[N * l + i] = a0 * s + b0 * c; }
        B2_ROLL for (int j = 0; j < 2; j++) {
            int idx = j == 0 ? k : l;
            if (idx < N - 1) {
                int m = idx + 1; float mm = fabsf(a[N * idx + m]);
                B2_ROLL for (int i = idx + 2; i < N; i++) { float val = fabsf(a[N * idx + i]); if (mm < val) { mm = val; m = i; } }
                rmax[idx] = m;
            }
            if (idx > 0) {
                int m = 0; float mm = fabsf(a[idx]);
                B2_ROLL for (int i = 1; i < idx; i++) { float val = fabsf(a[N * i + idx]); if (mm < val) { mm = val; m = i; } }
                cmax[idx] = m;
            }
        }
    }
    B2_ROLL for (int k = 0; k < N - 1; k++) {
        int m = k;
        B2_ROLL for (int i = k + 1; i < N; i++) if (w[m] < w[i]) m = i;
        if (k != m) {
            float tw = w[m]; w[m] = w[k]; w[k] = tw;
            B2_ROLL for (int i = 0; i < N; i++) { float tv = v[N * m + i]; v[N * m + i] = v[N * k + i]; v[N * k + i] = tv; }
        }
    }
}

template <int N>
B2_NOINLINE int solve_householder_c(float* a, float* b) {
    const float eps = FLT_EPSILON * 10;
    float u[N], hf[N];
    B2_ROLL for (int l = 0; l < N; l++) {
        const int len = N - l;
        float nrm = 0.f;
        B2_ROLL for (int i = 0; i < len; i++) { u[i] = a[(l + i) * N + l]; nrm += u[i] * u[i]; }
        float head = u[0];
        u[0] = u[0] + ((u[0] >= 0.0f) ? 1 : -1) * sqrtf(nrm);
        nrm = sqrtf(nrm + u[0] * u[0] - head * head);
        B2_ROLL for (int i = 0; i < len; i++) u[i] /= nrm;
        B2_ROLL for (int j = l; j < N; j++) {
            float dot = 0.f;
            B2_ROLL for (int i = l; i < N; i++) dot += u[i - l] * a[i * N + j];
            B2_ROLL for (int i = l; i < N; i++) a[i * N + j] -= 2 * u[i - l] * dot;
        }
        hf[l] = u[0] * u[0];
        B2_ROLL for (int i = 1; i < len; i++) a[(l + i) * N + l] = u[i] / u[0];
    }
    B2_ROLL for (int l = 0; l < N; l++) {
        u[0] = 1.f;
        B2_ROLL for (int j = 1; j < N - l; j++) u[j] = a[(j + l) * N + l];
        float dot = 0.f;
        B2_ROLL for (int i = l; i < N; i++) dot += u[i - l] * b[i];
        B2_ROLL for (int i = l; i < N; i++) b[i] -= 2 * u[i - l] * dot * hf[l];
    }
    B2_ROLL for (int i = N - 1; i >= 0; i--) {
        B2_ROLL for (int j = N - 1; j > i; j--) b[i] -= b[j] * a[i * N + j];
        if (fabsf(a[i * N + i]) < eps) return 0;
        b[i] /= a[i * N + i];
    }
    return 1;
}

// q: 5x3 row-major neighbour coordinates (destroyed); x: 3 outputs
B2_NOINLINE void plane_lsq_5x3_c(float* q, float* x) {
    float tau[3], nu[3], nd[3];
    int perm[3];
    B2_ROLL for (int k = 0; k < 3; k++) {
        float s = 0.f;
        B2_ROLL for (int i = 0; i < 5; i++) s += q[i * 3 + k] * q[i * 3 + k];
        nd[k] = sqrtf(s); nu[k] = nd[k]; perm[k] = k;
    }
    float mx = fmaxf(nu[0], fmaxf(nu[1], nu[2]));
    float th = mx * FLT_EPSILON;
    const float thr_helper = (th * th) / 5.0f;
    const float downdate_thr = sqrtf(FLT_EPSILON);
    int rank = 3;
    B2_ROLL for (int k = 0; k < 3; k++) {
        int big = k; float bv = nu[k];
        B2_ROLL for (int j = k + 1; j < 3; j++) if (nu[j] > bv) { bv = nu[j]; big = j; }
        if (rank == 3 && bv * bv < thr_helper * (float)(5 - k)) rank = k;
        if (big != k) {
            B2_ROLL for (int i = 0; i < 5; i++) { float t = q[i * 3 + k]; q[i * 3 + k] = q[i * 3 + big]; q[i * 3 + big] = t; }
            float t = nu[k]; nu[k] = nu[big]; nu[big] = t;
            t = nd[k]; nd[k] = nd[big]; nd[big] = t;
            int tp = perm[k]; perm[k] = perm[big]; perm[big] = tp;
        }
        float tail = 0.f;
        B2_ROLL for (int i = k + 1; i < 5; i++) tail += q[i * 3 + k] * q[i * 3 + k];
        float c0 = q[k * 3 + k], beta, tk;
        if (tail <= FLT_MIN) {
            tk = 0.f; beta = c0;
            B2_ROLL for (int i = k + 1; i < 5; i++) q[i * 3 + k] = 0.f;
        } else {
            beta = sqrtf(c0 * c0 + tail);
            if (c0 >= 0.f) beta = -beta;
            float den = c0 - beta;
            B2_ROLL for (int i = k + 1; i < 5; i++) q[i * 3 + k] = q[i * 3 + k] / den;
            tk = (beta - c0) / beta;
        }
        tau[k] = tk; q[k * 3 + k] = beta;
        B2_ROLL for (int j = k + 1; j < 3; j++) {
            if (tk != 0.f) {
                float t = 0.f;
                B2_ROLL for (int i = k + 1; i < 5; i++) t += q[i * 3 + k] * q[i * 3 + j];
                t += q[k * 3 + j];
                q[k * 3 + j] -= tk * t;
                B2_ROLL for (int i = k + 1; i < 5; i++) q[i * 3 + j] -= tk * q[i * 3 + k] * t;
            }
            if (nu[j] != 0.f) {
                float t = fabsf(q[k * 3 + j]) / nu[j];
                t = (1.f + t) * (1.f - t);
                t = t < 0.f ? 0.f : t;
                float r = nu[j] / nd[j];
                float t2 = t * (r * r);
                if (t2 <= downdate_thr) {
                    float s = 0.f;
                    B2_ROLL for (int i = k + 1; i < 5; i++) s += q[i * 3 + j] * q[i * 3 + j];
                    nd[j] = sqrtf(s); nu[j] = nd[j];
                } else {
                    nu[j] *= sqrtf(t);
                }
            }
        }
    }
    float c[5];
    B2_ROLL for (int i = 0; i < 5; i++) c[i] = -1.f;
    x[0] = 0.f; x[1] = 0.f; x[2] = 0.f;
    B2_ROLL for (int k = 0; k < rank; k++) {
        if (tau[k] == 0.f) continue;
        float t = 0.f;
        B2_ROLL for (int i = k + 1; i < 5; i++) t += q[i * 3 + k] * c[i];
        t += c[k];
        c[k] -= tau[k] * t;
        B2_ROLL for (int i = k + 1; i < 5; i++) c[i] -= tau[k] * q[i * 3 + k] * t;
    }
    B2_ROLL for (int i = rank - 1; i >= 0; i--) {
        float s = c[i];
        B2_ROLL for (int j = i + 1; j < rank; j++) s -= q[i * 3 + j] * c[j];
        c[i] = s / q[i * 3 + i];
    }
    B2_ROLL for (int i = 0; i < rank; i++) x[perm[i]] = c[i];
}
#endif  // __CUDACC__

}  // namespace b2
