// ORACLE — TEST INFRASTRUCTURE ONLY. Small fp64 dense helpers shared by the GICP and NDT restatements.
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>

namespace {

// cyclic Jacobi, symmetric 3x3, fp64: eigenvalues ascending in w, eigenvectors as columns of V
void jacobi3d(const double A[9], double w[3], double V[9]) {
    double a[9]; std::memcpy(a, A, sizeof(a));
    for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 50; sweep++) {
        double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
        if (off < 1e-300) break;
        for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
            double apq = a[p * 3 + q];
            if (apq == 0.0) continue;
            double theta = (a[q * 3 + q] - a[p * 3 + p]) / (2.0 * apq);
            double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 3; k++) { double akp = a[k * 3 + p], akq = a[k * 3 + q]; a[k * 3 + p] = c * akp - s * akq; a[k * 3 + q] = s * akp + c * akq; }
            for (int k = 0; k < 3; k++) { double apk = a[p * 3 + k], aqk = a[q * 3 + k]; a[p * 3 + k] = c * apk - s * aqk; a[q * 3 + k] = s * apk + c * aqk; }
            for (int k = 0; k < 3; k++) { double vkp = V[k * 3 + p], vkq = V[k * 3 + q]; V[k * 3 + p] = c * vkp - s * vkq; V[k * 3 + q] = s * vkp + c * vkq; }
        }
    }
    w[0] = a[0]; w[1] = a[4]; w[2] = a[8];
    for (int i = 0; i < 2; i++) for (int j = i + 1; j < 3; j++) if (w[j] < w[i]) {
        std::swap(w[i], w[j]);
        for (int k = 0; k < 3; k++) std::swap(V[k * 3 + i], V[k * 3 + j]);
    }
}


// inverse of a general 3x3 through its cofactors (what Eigen does for fixed-size 3x3 inverse())
inline bool inv3d(const double A[9], double B[9]) {
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    const double id = 1.0 / det;
    B[0] = c00 * id; B[1] = (A[2] * A[7] - A[1] * A[8]) * id; B[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    B[3] = c01 * id; B[4] = (A[0] * A[8] - A[2] * A[6]) * id; B[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    B[6] = c02 * id; B[7] = (A[1] * A[6] - A[0] * A[7]) * id; B[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    return det != 0.0;
}

}  // namespace
