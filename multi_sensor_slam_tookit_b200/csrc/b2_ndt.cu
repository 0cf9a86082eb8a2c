// Normal Distributions Transform on the device. Replaces pcl::NormalDistributionsTransform<PointXYZ, PointXYZ> as
//   Calibration_Tookit/multi_lidar/src/multi_lidar_calibration/src/multi_lidar_calibrator.cpp:35-72
// drives it (setTransformationEpsilon / setStepSize / setResolution / setMaximumIterations / setInputSource /
// setInputTarget / align / hasConverged / getFitnessScore / getTransformationProbability / getFinalTransformation).
//
// Target (setInputTarget): pcl::VoxelGridCovariance with leaf = resolution. Points are keyed by PCL's float voxel
// arithmetic, sorted (stable) by voxel, and one thread per voxel accumulates sum p, sum p p^T (double) and the float
// centroid in ascending input index, i.e. in PCL's own order; then mean, single-pass covariance * (n-1)/n, Jacobi
// eigen-decomposition, eigenvalue inflation (0.01 * largest), cofactor inverse. HBM layout: a dense int table
// cell -> voxel rank (-1 = fewer than 6 points), and per voxel a float4 centroid, 3 doubles mean, 9 doubles inverse
// covariance (96 + 16 B per voxel; ~10^4 voxels, L2-resident).
//
// Derivative pass (one kernel per evaluation): one thread per source point — float transform, the 27 voxels around it,
// float centroid distance < resolution (what PCL's radius search over voxel centroids returns), eq. 6.9-6.13 of
// Magnusson 2009 in double with the angular terms precomputed on the host exactly as PCL does; 1 + 6 + 21 terms are
// reduced over the warp as they are produced, CTA partials are added in CTA order by the last CTA.
// The Newton step (SVD solve of the 6x6), the More-Thuente line search and the convergence rule run on the host
// between evaluations, in PCL's operation order; each evaluation returns 28 doubles.
#include "b2_cloud.cuh"
#include "b2_bvh.cuh"
#include <atomic>
#include <cmath>
#include <cfloat>
#include <climits>
#include <algorithm>
#include <vector>

namespace b2 {

constexpr int NDT_THREADS = 256;
constexpr int NDT_WARPS = NDT_THREADS / 32;
constexpr int NDT_NSUM = 29;                 // score, 6 gradient, 21 Hessian upper, pair count
constexpr size_t NDT_MAX_CELLS = (size_t)1 << 28;

struct NdtGrid {
    float inv_leaf, r2;
    int min_b[3], div_b[3];
    const int32_t* rank_of_cell;     // dense, div_b[0]*div_b[1]*div_b[2]
    const float4* centroid;          // per voxel
    const double* mean;              // 3 per voxel
    const double* icov;              // 9 per voxel
};

struct NdtAngular { double j[8][3]; double h[15][3]; };

struct NdtArgs {
    NdtGrid g;
    const float* src;                // 3 floats per source point
    uint32_t n_src;
    float T[12];                     // float 3x4, row-major
    NdtAngular ang;
    double d1, d2;
    int mode;                        // 0: score + gradient, 1: + Hessian, 2: Hessian only
    double* partials;
    double* sums;                    // NDT_NSUM
    unsigned* ticket;
    double* host_sums;               // the same sums in mapped host memory, host_sums[32] = seq once they are complete
    double seq;
};

// ---- target voxel build ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ndt_bbox(const float* __restrict__ xyz, uint32_t n, uint32_t* __restrict__ bb) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
        if (isfinite(x) && isfinite(y) && isfinite(z)) {
            mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
            mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int d = 0; d < 3; d++)
            if (mn[d] <= mx[d]) { atomicMin(&bb[d], float_flip(mn[d])); atomicMax(&bb[3 + d], float_flip(mx[d])); }
}
__global__ void k_ndt_bbox_init(uint32_t* bb) {
    if (threadIdx.x < 3) bb[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) bb[threadIdx.x] = 0u;
}

struct NdtKeyGeom { float inv_leaf; int min_b[3]; int mul[3]; uint32_t invalid; };

__global__ void __launch_bounds__(256) k_ndt_key(const float* __restrict__ xyz, uint32_t n, NdtKeyGeom g, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
    uint32_t key = g.invalid;
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
        const int i0 = (int)(floorf(x * g.inv_leaf) - (float)g.min_b[0]);
        const int i1 = (int)(floorf(y * g.inv_leaf) - (float)g.min_b[1]);
        const int i2 = (int)(floorf(z * g.inv_leaf) - (float)g.min_b[2]);
        key = (uint32_t)(i0 * g.mul[0] + i1 * g.mul[1] + i2 * g.mul[2]);
    }
    keys[i] = key; vals[i] = i;
}
// flags[i] = 1 where a voxel run starts in the sorted order (never for the non-finite bucket); flags[n] = 0
__global__ void __launch_bounds__(256) k_ndt_heads(const uint32_t* __restrict__ keys, uint32_t n, uint32_t invalid, uint32_t* __restrict__ flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint32_t f = 0;
    if (i < n) { const uint32_t k = keys[i]; f = (k != invalid) && (i == 0 || keys[i - 1] != k); }
    flags[i] = f;
}
__global__ void __launch_bounds__(256) k_ndt_seg_start(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ segid, uint32_t n, uint32_t invalid,
                                                       uint32_t* __restrict__ seg_start) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { if (keys[n - 1] != invalid) seg_start[segid[n]] = n; return; }
    const uint32_t k = keys[i];
    const bool head = (i == 0) || (keys[i - 1] != k);
    if (!head) return;
    if (k == invalid) seg_start[segid[n]] = i;
    else seg_start[segid[i]] = i;
}
// keep[s] = 1 when run s holds at least min_pts points
__global__ void __launch_bounds__(256) k_ndt_keep(const uint32_t* __restrict__ seg_start, uint32_t nseg, uint32_t min_pts, uint32_t* __restrict__ keep) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > nseg) return;
    keep[s] = (s < nseg && (seg_start[s + 1] - seg_start[s]) >= min_pts) ? 1u : 0u;
}

__device__ __forceinline__ void jacobi3_ndt(const double A[9], double w[3], double V[9]) {
    double a[9];
#pragma unroll
    for (int i = 0; i < 9; i++) { a[i] = A[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 50; sweep++) {
        const double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
        if (off < 1e-300) break;
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int q = p + 1; q < 3; q++) {
                const double apq = a[p * 3 + q];
                if (apq == 0.0) continue;
                const double theta = (a[q * 3 + q] - a[p * 3 + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
                for (int k = 0; k < 3; k++) { const double akp = a[k * 3 + p], akq = a[k * 3 + q]; a[k * 3 + p] = c * akp - s * akq; a[k * 3 + q] = s * akp + c * akq; }
#pragma unroll
                for (int k = 0; k < 3; k++) { const double apk = a[p * 3 + k], aqk = a[q * 3 + k]; a[p * 3 + k] = c * apk - s * aqk; a[q * 3 + k] = s * apk + c * aqk; }
#pragma unroll
                for (int k = 0; k < 3; k++) { const double vkp = V[k * 3 + p], vkq = V[k * 3 + q]; V[k * 3 + p] = c * vkp - s * vkq; V[k * 3 + q] = s * vkp + c * vkq; }
            }
    }
    w[0] = a[0]; w[1] = a[4]; w[2] = a[8];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = i + 1; j < 3; j++)
            if (w[j] < w[i]) {
                double t = w[i]; w[i] = w[j]; w[j] = t;
#pragma unroll
                for (int k = 0; k < 3; k++) { t = V[k * 3 + i]; V[k * 3 + i] = V[k * 3 + j]; V[k * 3 + j] = t; }
            }
}
__device__ __forceinline__ void inv3_cofactor(const double A[9], double B[9]) {
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    const double id = 1.0 / det;
    B[0] = c00 * id; B[1] = (A[2] * A[7] - A[1] * A[8]) * id; B[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    B[3] = c01 * id; B[4] = (A[0] * A[8] - A[2] * A[6]) * id; B[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    B[6] = c02 * id; B[7] = (A[1] * A[6] - A[0] * A[7]) * id; B[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

// one thread per voxel with >= min points: sums in ascending input index, then the statistics of VoxelGridCovariance
__global__ void __launch_bounds__(128) k_ndt_voxel_stats(const float* __restrict__ xyz, const uint32_t* __restrict__ order, const uint32_t* __restrict__ keys,
                                                         const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ keep_scan, uint32_t nseg,
                                                         double eig_mult, int32_t* __restrict__ rank_of_cell, float4* __restrict__ centroid,
                                                         double* __restrict__ mean, double* __restrict__ icov, int32_t* __restrict__ vox_index,
                                                         int32_t* __restrict__ vox_npts) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const uint32_t rank = keep_scan[s];
    if (keep_scan[s + 1] == rank) return;
    const uint32_t b = seg_start[s], e = seg_start[s + 1];
    double m[3] = {0, 0, 0}, C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    float cs[3] = {0.f, 0.f, 0.f};
    for (uint32_t i = b; i < e; i++) {
        const size_t src = order[i];
        const float fx = xyz[3 * src], fy = xyz[3 * src + 1], fz = xyz[3 * src + 2];
        const double q[3] = {(double)fx, (double)fy, (double)fz};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            m[a] += q[a];
#pragma unroll
            for (int c = 0; c < 3; c++) C[a * 3 + c] += q[a] * q[c];
        }
        cs[0] += fx; cs[1] += fy; cs[2] += fz;
    }
    int npts = (int)(e - b);
    const double nd = (double)npts;
    const float nf = (float)npts;
    const double ps[3] = {m[0], m[1], m[2]};
#pragma unroll
    for (int a = 0; a < 3; a++) m[a] /= nd;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int c = 0; c < 3; c++) C[a * 3 + c] = (C[a * 3 + c] - 2 * (ps[a] * m[c])) / nd + m[a] * m[c];
#pragma unroll
    for (int q = 0; q < 9; q++) C[q] *= (nd - 1.0) / nd;
    double w[3], V[9], ic[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    jacobi3_ndt(C, w, V);
    if (w[0] < 0 || w[1] < 0 || w[2] <= 0) npts = -1;
    else {
        const double min_ev = eig_mult * w[2];
        if (w[0] < min_ev) {
            w[0] = min_ev;
            if (w[1] < min_ev) w[1] = min_ev;
            double Vi[9];
            inv3_cofactor(V, Vi);
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    double acc = 0;
#pragma unroll
                    for (int k = 0; k < 3; k++) acc += V[a * 3 + k] * w[k] * Vi[k * 3 + c];
                    C[a * 3 + c] = acc;
                }
        }
        inv3_cofactor(C, ic);
        double mxc = -INFINITY, mnc = INFINITY;
#pragma unroll
        for (int q = 0; q < 9; q++) { mxc = fmax(mxc, ic[q]); mnc = fmin(mnc, ic[q]); }
        if (mxc == (double)INFINITY || mnc == -(double)INFINITY) npts = -1;
    }
    const uint32_t cell = keys[b];
    rank_of_cell[cell] = (int32_t)rank;
    centroid[rank] = make_float4(cs[0] / nf, cs[1] / nf, cs[2] / nf, 0.f);
#pragma unroll
    for (int a = 0; a < 3; a++) mean[3 * (size_t)rank + a] = m[a];
#pragma unroll
    for (int q = 0; q < 9; q++) icov[9 * (size_t)rank + q] = ic[q];
    vox_index[rank] = (int32_t)cell;
    vox_npts[rank] = npts;
}

// ---- derivative pass -------------------------------------------------------------------------------------------
__device__ __forceinline__ double dot3(const double v[3], double x, double y, double z) { return x * v[0] + y * v[1] + z * v[2]; }

#ifndef NDT_MIN_BLOCKS
#define NDT_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(NDT_THREADS, NDT_MIN_BLOCKS) k_ndt_derivatives(NdtArgs A) {
    __shared__ double s_red[NDT_WARPS][32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double tot = 0.0;
    const uint32_t n_chunks = (A.n_src + 31) >> 5;
    const NdtGrid& g = A.g;
    for (uint32_t chunk = blockIdx.x * NDT_WARPS + warp; chunk < n_chunks; chunk += gridDim.x * NDT_WARPS) {
        const uint32_t i = (chunk << 5) + lane;
        const bool valid = i < A.n_src;
        double term[32];
#pragma unroll
        for (int q = 0; q < 32; q++) term[q] = 0.0;
        if (valid) {
            const float sx = A.src[3 * (size_t)i], sy = A.src[3 * (size_t)i + 1], sz = A.src[3 * (size_t)i + 2];
            const float tx = A.T[0] * sx + A.T[1] * sy + A.T[2] * sz + A.T[3];
            const float ty = A.T[4] * sx + A.T[5] * sy + A.T[6] * sz + A.T[7];
            const float tz = A.T[8] * sx + A.T[9] * sy + A.T[10] * sz + A.T[11];
            const int c0 = (int)(floorf(tx * g.inv_leaf) - (float)g.min_b[0]);
            const int c1 = (int)(floorf(ty * g.inv_leaf) - (float)g.min_b[1]);
            const int c2 = (int)(floorf(tz * g.inv_leaf) - (float)g.min_b[2]);
            const bool fin = isfinite(tx) && isfinite(ty) && isfinite(tz);
            // point gradient / Hessian of eq. 6.18-6.21: depend on the untransformed point only
            const double x = (double)sx, y = (double)sy, z = (double)sz;
            double pg[3][6];
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 6; c++) pg[r][c] = (r == c) ? 1.0 : 0.0;
            pg[1][3] = dot3(A.ang.j[0], x, y, z); pg[2][3] = dot3(A.ang.j[1], x, y, z);
            pg[0][4] = dot3(A.ang.j[2], x, y, z); pg[1][4] = dot3(A.ang.j[3], x, y, z); pg[2][4] = dot3(A.ang.j[4], x, y, z);
            pg[0][5] = dot3(A.ang.j[5], x, y, z); pg[1][5] = dot3(A.ang.j[6], x, y, z); pg[2][5] = dot3(A.ang.j[7], x, y, z);
            // second derivatives: blocks (3+i, 3+j) of the 18x6 point Hessian, vectors a..f of eq. 6.21
            double hv[6][3];
            if (A.mode != 0) {
                hv[0][0] = 0; hv[0][1] = dot3(A.ang.h[0], x, y, z); hv[0][2] = dot3(A.ang.h[1], x, y, z);       // a
                hv[1][0] = 0; hv[1][1] = dot3(A.ang.h[2], x, y, z); hv[1][2] = dot3(A.ang.h[3], x, y, z);       // b
                hv[2][0] = 0; hv[2][1] = dot3(A.ang.h[4], x, y, z); hv[2][2] = dot3(A.ang.h[5], x, y, z);       // c
                hv[3][0] = dot3(A.ang.h[6], x, y, z); hv[3][1] = dot3(A.ang.h[7], x, y, z); hv[3][2] = dot3(A.ang.h[8], x, y, z);      // d
                hv[4][0] = dot3(A.ang.h[9], x, y, z); hv[4][1] = dot3(A.ang.h[10], x, y, z); hv[4][2] = dot3(A.ang.h[11], x, y, z);    // e
                hv[5][0] = dot3(A.ang.h[12], x, y, z); hv[5][1] = dot3(A.ang.h[13], x, y, z); hv[5][2] = dot3(A.ang.h[14], x, y, z);   // f
            }
            if (fin) {
                // The voxels in range first, as a per-lane list (cheap tests, in cell order), the derivative terms afterwards:
                // a point has about three voxels in range, its warp about twenty of the 27 slots with a hit in some lane, and
                // in one loop the warp pays the ~400-flop block below once per such slot instead of once per list entry of
                // its busiest lane. The per-point order of the voxels is unchanged, so every sum is bit-identical.
                int32_t hit[27];
                int nh = 0;
#pragma unroll 1
                for (int cell = 0; cell < 27; cell++) {
                    const int vx = c0 + (cell % 3) - 1, vy = c1 + ((cell / 3) % 3) - 1, vz = c2 + (cell / 9) - 1;
                    if (vx < 0 || vy < 0 || vz < 0 || vx >= g.div_b[0] || vy >= g.div_b[1] || vz >= g.div_b[2]) continue;
                    const int32_t rk = __ldg(&g.rank_of_cell[((size_t)vz * g.div_b[1] + vy) * g.div_b[0] + vx]);
                    if (rk < 0) continue;
                    const float4 ce = __ldg(&g.centroid[rk]);
                    const float ddx = tx - ce.x, ddy = ty - ce.y, ddz = tz - ce.z;
                    float dd = ddx * ddx;
                    dd = dd + ddy * ddy;
                    dd = dd + ddz * ddz;
                    if (!(dd < g.r2)) continue;
                    hit[nh++] = rk;
                }
#pragma unroll 1
                for (int hi = 0; hi < nh; hi++) {
                    const int32_t rk = hit[hi];
                    const double* mu = &g.mean[3 * (size_t)rk];
                    const double* ci = &g.icov[9 * (size_t)rk];
                    const double xt0 = (double)tx - mu[0], xt1 = (double)ty - mu[1], xt2 = (double)tz - mu[2];
                    double c[9];
#pragma unroll
                    for (int q = 0; q < 9; q++) c[q] = ci[q];
                    const double cx0 = c[0] * xt0 + c[1] * xt1 + c[2] * xt2, cx1 = c[3] * xt0 + c[4] * xt1 + c[5] * xt2, cx2 = c[6] * xt0 + c[7] * xt1 + c[8] * xt2;
                    const double xcx = xt0 * cx0 + xt1 * cx1 + xt2 * cx2;
                    double e = exp(-A.d2 * xcx / 2);
                    const double score_inc = -A.d1 * e;
                    e = A.d2 * e;
                    term[28] += 1.0;
                    if (e > 1 || e < 0 || e != e) continue;
                    e *= A.d1;
                    if (A.mode != 2) term[0] += score_inc;
                    double cdx[6][3], xdot[6];
#pragma unroll
                    for (int k = 0; k < 6; k++) {
#pragma unroll
                        for (int a = 0; a < 3; a++) cdx[k][a] = c[a * 3] * pg[0][k] + c[a * 3 + 1] * pg[1][k] + c[a * 3 + 2] * pg[2][k];
                        xdot[k] = xt0 * cdx[k][0] + xt1 * cdx[k][1] + xt2 * cdx[k][2];
                        if (A.mode != 2) term[1 + k] += xdot[k] * e;
                    }
                    if (A.mode != 0) {
                        int q = 7;
#pragma unroll
                        for (int a = 0; a < 6; a++)
#pragma unroll
                            for (int bcol = a; bcol < 6; bcol++) {
                                double t2 = 0.0;
                                if (a >= 3) {
                                    // block (a, bcol) of the point Hessian: a..f by (a-3, bcol-3) = (0,0)a (0,1)b (0,2)c (1,1)d (1,2)e (2,2)f
                                    const int ia = a - 3, ib = bcol - 3;
                                    const int hvi = (ia == 0) ? ib : (ia == 1 ? 2 + ib : 5);
                                    const double hx = hv[hvi][0], hy = hv[hvi][1], hz = hv[hvi][2];
                                    const double h0 = c[0] * hx + c[1] * hy + c[2] * hz, h1 = c[3] * hx + c[4] * hy + c[5] * hz, h2 = c[6] * hx + c[7] * hy + c[8] * hz;
                                    t2 = xt0 * h0 + xt1 * h1 + xt2 * h2;
                                }
                                const double t3 = pg[0][bcol] * cdx[a][0] + pg[1][bcol] * cdx[a][1] + pg[2][bcol] * cdx[a][2];
                                term[q] += e * (-A.d2 * xdot[a] * xdot[bcol] + t2 + t3);
                                q++;
                            }
                    }
                }
            }
        }
        __syncwarp();
        tot += warp_reduce_scatter32(term);
    }
    s_red[warp][lane] = tot;
    __syncthreads();
    if (threadIdx.x < NDT_NSUM) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < NDT_WARPS; w++) v += s_red[w][threadIdx.x];
        A.partials[(size_t)blockIdx.x * NDT_NSUM + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(A.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < NDT_NSUM) {
        double v = 0.0;
        for (unsigned b = 0; b < gridDim.x; b++) v += __ldcg(&A.partials[(size_t)b * NDT_NSUM + threadIdx.x]);
        A.sums[threadIdx.x] = v;
        if (A.host_sums) A.host_sums[threadIdx.x] = v;
    }
    if (A.host_sums) {
        // the host spins on the sequence number instead of paying a copy and a stream synchronisation per evaluation
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) { *reinterpret_cast<volatile double*>(A.host_sums + 32) = A.seq; __threadfence_system(); }
    }
    if (threadIdx.x == 0) *A.ticket = 0;
}

// ---- fitness score: mean squared distance to the nearest target point -------------------------------------------
// A warp answers NDT_FIT_QPW queries one after the other (each a warp-cooperative descent of the BVH: a chain of dependent loads).
// With 32 queries per warp a 85 k-point source was 2 700 warps — 18 per SM, each a 32-long serial chain: 1 ms. Eight per warp
// gives the SMs four times the chains to overlap. The sums are accumulated with atomics (order-free, as before).
constexpr uint32_t NDT_FIT_QPW = 8;
__global__ void __launch_bounds__(128) k_ndt_fitness(const __grid_constant__ BvhDev T, const float* __restrict__ src, uint32_t n_src, NdtArgs X, double* __restrict__ out) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t base = warp * NDT_FIT_QPW;
    if (base >= n_src) return;
    float tx = 0, ty = 0, tz = 0;
    if ((uint32_t)lane < NDT_FIT_QPW && base + lane < n_src) {
        const size_t i = base + lane;
        const float sx = src[3 * i], sy = src[3 * i + 1], sz = src[3 * i + 2];
        tx = X.T[0] * sx + X.T[1] * sy + X.T[2] * sz + X.T[3];
        ty = X.T[4] * sx + X.T[5] * sy + X.T[6] * sz + X.T[7];
        tz = X.T[8] * sx + X.T[9] * sy + X.T[10] * sz + X.T[11];
    }
    const int nq = (int)min(NDT_FIT_QPW, n_src - base);
    double sum = 0.0, cnt = 0.0;
    for (int j = 0; j < nq; j++) {
        const double qx = (double)__shfl_sync(full, tx, j), qy = (double)__shfl_sync(full, ty, j), qz = (double)__shfl_sync(full, tz, j);
        double d2 = INFINITY;
        if (isfinite(qx) && isfinite(qy) && isfinite(qz)) d2 = bvh_nn1_warp(T, qx, qy, qz);
        if (d2 < INFINITY) { sum += d2; cnt += 1.0; }
    }
    if (lane == 0) { atomicAdd(&out[0], sum); atomicAdd(&out[1], cnt); }
}

__global__ void __launch_bounds__(256) k_ndt_transform_out(const float* __restrict__ src, uint32_t n, NdtArgs X, unsigned char* __restrict__ out, size_t stride) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sx = src[3 * (size_t)i], sy = src[3 * (size_t)i + 1], sz = src[3 * (size_t)i + 2];
    float* o = reinterpret_cast<float*>(out + (size_t)i * stride);
    o[0] = X.T[0] * sx + X.T[1] * sy + X.T[2] * sz + X.T[3];
    o[1] = X.T[4] * sx + X.T[5] * sy + X.T[6] * sz + X.T[7];
    o[2] = X.T[8] * sx + X.T[9] * sy + X.T[10] * sz + X.T[11];
    if (stride >= 16) o[3] = 1.0f;
}

__global__ void __launch_bounds__(256) k_ndt_pack_xyz(const unsigned char* __restrict__ raw, size_t stride, uint32_t n, float* __restrict__ xyz) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    xyz[3 * (size_t)i] = p[0]; xyz[3 * (size_t)i + 1] = p[1]; xyz[3 * (size_t)i + 2] = p[2];
}
__global__ void __launch_bounds__(256) k_ndt_widen(const float* __restrict__ xyz, uint32_t n, double* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * n) out[i] = (double)xyz[i];
}
__global__ void k_fill_i32_ndt(int32_t* p, size_t n, int32_t v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---- host numerics (PCL's operation order) ------------------------------------------------------------------------
template <int N>
static void jacobi_sym_host(const double* A, double* w, double* V) {
    double a[N * N];
    memcpy(a, A, sizeof(a));
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

// Eigen::JacobiSVD(H).solve(b) for the symmetric 6x6 Hessian (ndt.hpp: sv.solve(-score_gradient))
static void svd_solve6_host(const double H[36], const double b[6], double x[6]) {
    double w[6], V[36];
    jacobi_sym_host<6>(H, w, V);
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

static void pose_to_matrix_f(const double p[6], float T[16]) {
    // Translation3f * AngleAxisf(rx, X) * AngleAxisf(ry, Y) * AngleAxisf(rz, Z) (ndt.hpp computeStepLengthMT), in float
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

static void euler_012_f(const float T[16], float out[3]) {
    // Eigen 3.3 Matrix3f::eulerAngles(0, 1, 2) (ndt.hpp: eig_transformation.rotation().eulerAngles(0, 1, 2))
    auto c = [&](int r, int col) { return T[r * 4 + col]; };
    float res[3];
    res[0] = std::atan2(c(1, 2), c(2, 2));
    const float c2 = std::sqrt(c(0, 0) * c(0, 0) + c(0, 1) * c(0, 1));
    if (res[0] > 0.f) { res[0] -= (float)M_PI; res[1] = std::atan2(-c(0, 2), -c2); }
    else res[1] = std::atan2(-c(0, 2), c2);
    const float s1 = std::sin(res[0]), c1 = std::cos(res[0]);
    res[2] = std::atan2(s1 * c(2, 0) - c1 * c(1, 0), c1 * c(1, 1) - s1 * c(2, 1));
    for (int i = 0; i < 3; i++) out[i] = -res[i];
}

}  // namespace b2

using namespace b2;

struct b2_ndt_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    double* h_res = nullptr; double* d_res = nullptr; double res_seq = 0.0;   // mapped result block (64 doubles)
    // parameters (PCL defaults: ndt.hpp constructor)
    float resolution = 1.0f;
    double step_size = 0.1, outlier_ratio = 0.55, trans_eps = 0.1;
    int max_iterations = 35;
    int min_points_per_voxel = 6;
    double min_covar_eigvalue_mult = 0.01;
    // target
    DevBuf tgt_xyz, rank_of_cell, centroid, mean, icov, vox_index, vox_npts, work;
    size_t n_tgt = 0; uint32_t n_vox = 0;
    int min_b[3] = {0, 0, 0}, div_b[3] = {1, 1, 1};
    float inv_leaf = 1.0f;
    bool have_tgt = false, grid_stale = true;
    BvhIndex bvh; bool bvh_valid = false;
    // source
    DevBuf src_xyz; size_t n_src = 0; bool have_src = false;
    // evaluation plumbing
    DevBuf partials, sums;
    PinBuf pin, stage;
    int grid_blocks = 1;
    NdtAngular ang{};
    double gauss_d1 = 0, gauss_d2 = 0;
    // results
    float final_T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    int nr_iterations = 0, evaluations = 0, launches = 0;
    bool converged = false;
    double trans_probability = 0;
    long long pairs_last = 0;
    float last_ms = 0.f;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

static int ndt_build_target(b2_ndt_s* h) {
    cudaStream_t s = h->stream;
    h->n_vox = 0; h->grid_stale = false;
    h->inv_leaf = 1.0f / h->resolution;
    h->min_b[0] = h->min_b[1] = h->min_b[2] = 0; h->div_b[0] = h->div_b[1] = h->div_b[2] = 1;
    const size_t n = h->n_tgt;
    B2_CHECK(h->rank_of_cell.reserve(4));
    B2_CUDA(cudaMemsetAsync(h->rank_of_cell.p, 0xff, 4, s));
    if (n == 0) return B2_OK;
    const float* xyz = h->tgt_xyz.as<float>();
    B2_CHECK(h->work.reserve(64));
    uint32_t* bb = h->work.as<uint32_t>();
    k_ndt_bbox_init<<<1, 32, 0, s>>>(bb); count_launch();
    k_ndt_bbox<<<(int)std::min<size_t>((n + 255) / 256, (size_t)device_sm_count() * 4), 256, 0, s>>>(xyz, (uint32_t)n, bb); count_launch();
    uint32_t hbb[6];
    B2_CUDA(cudaMemcpyAsync(hbb, bb, sizeof(hbb), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    if (hbb[0] == 0xffffffffu) return B2_OK;
    auto unflip = [](uint32_t u) { uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; float f; memcpy(&f, &v, 4); return f; };
    float mn[3], mx[3];
    for (int d = 0; d < 3; d++) { mn[d] = unflip(hbb[d]); mx[d] = unflip(hbb[3 + d]); }
    int64_t dx[3];
    for (int d = 0; d < 3; d++) dx[d] = (int64_t)((mx[d] - mn[d]) * h->inv_leaf) + 1;
    if (dx[0] * dx[1] * dx[2] > (int64_t)INT32_MAX) return B2_OK;        // PCL: "leaf size is too small", empty grid
    for (int d = 0; d < 3; d++) {
        h->min_b[d] = (int)std::floor(mn[d] * h->inv_leaf);
        h->div_b[d] = (int)std::floor(mx[d] * h->inv_leaf) - h->min_b[d] + 1;
    }
    const size_t ncell = (size_t)h->div_b[0] * h->div_b[1] * h->div_b[2];
    if (ncell > NDT_MAX_CELLS) { set_error("ndt: %zu voxels of %.3f m exceed the dense-table budget", ncell, h->resolution); return B2_ERR_TOO_LARGE; }
    NdtKeyGeom kg;
    kg.inv_leaf = h->inv_leaf;
    for (int d = 0; d < 3; d++) kg.min_b[d] = h->min_b[d];
    kg.mul[0] = 1; kg.mul[1] = h->div_b[0]; kg.mul[2] = h->div_b[0] * h->div_b[1];
    kg.invalid = (uint32_t)ncell;
    int bits = 1;
    while (((size_t)1 << bits) <= ncell) bits++;
    const size_t nal = (n + 63) & ~(size_t)63, np1 = n + 1, np1al = (np1 + 63) & ~(size_t)63;
    const size_t need = 4 * nal * 4 + 3 * np1al * 4 + sort_tmp_bytes(n) + scan_tmp_bytes(np1) + 1024;
    B2_CHECK(h->work.reserve(need));
    uint32_t* ka = h->work.as<uint32_t>();
    uint32_t* va = ka + nal; uint32_t* kb = va + nal; uint32_t* vb = kb + nal;
    uint32_t* flags = vb + nal; uint32_t* seg_start = flags + np1al; uint32_t* keep = seg_start + np1al;
    char* scratch = reinterpret_cast<char*>(keep + np1al);
    char* scan_scratch = scratch + sort_tmp_bytes(n);
    const unsigned nblk = (unsigned)((n + 255) / 256), nblk1 = (unsigned)((np1 + 255) / 256);
    k_ndt_key<<<nblk, 256, 0, s>>>(xyz, (uint32_t)n, kg, ka, va); count_launch();
    uint32_t *ks, *vs;
    B2_CHECK(radix_sort_pairs(ka, va, kb, vb, n, bits, scratch, s, &ks, &vs));
    k_ndt_heads<<<nblk1, 256, 0, s>>>(ks, (uint32_t)n, kg.invalid, flags); count_launch();
    B2_CHECK(exclusive_scan_u32(flags, np1, scan_scratch, s));
    k_ndt_seg_start<<<nblk1, 256, 0, s>>>(ks, flags, (uint32_t)n, kg.invalid, seg_start); count_launch();
    uint32_t nseg = 0;
    B2_CUDA(cudaMemcpyAsync(&nseg, flags + n, 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    if (nseg == 0) return B2_OK;
    k_ndt_keep<<<(nseg + 1 + 255) / 256, 256, 0, s>>>(seg_start, nseg, (uint32_t)h->min_points_per_voxel, keep); count_launch();
    B2_CHECK(exclusive_scan_u32(keep, (size_t)nseg + 1, scan_scratch, s));
    uint32_t nvox = 0;
    B2_CUDA(cudaMemcpyAsync(&nvox, keep + nseg, 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    B2_CHECK(h->rank_of_cell.reserve(ncell * 4));
    k_fill_i32_ndt<<<(int)std::min<size_t>((ncell + 255) / 256, (size_t)device_sm_count() * 8), 256, 0, s>>>(h->rank_of_cell.as<int32_t>(), ncell, -1); count_launch();
    const size_t nv1 = std::max<uint32_t>(nvox, 1);
    B2_CHECK(h->centroid.reserve(nv1 * 16)); B2_CHECK(h->mean.reserve(nv1 * 24)); B2_CHECK(h->icov.reserve(nv1 * 72));
    B2_CHECK(h->vox_index.reserve(nv1 * 4)); B2_CHECK(h->vox_npts.reserve(nv1 * 4));
    if (nvox) {
        k_ndt_voxel_stats<<<(nseg + 127) / 128, 128, 0, s>>>(xyz, vs, ks, seg_start, keep, nseg, h->min_covar_eigvalue_mult, h->rank_of_cell.as<int32_t>(),
                                                           h->centroid.as<float4>(), h->mean.as<double>(), h->icov.as<double>(),
                                                           h->vox_index.as<int32_t>(), h->vox_npts.as<int32_t>()); count_launch();
    }
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaStreamSynchronize(s));
    h->n_vox = nvox;
    return B2_OK;
}

static void ndt_angle_derivatives(b2_ndt_s* h, const double p[6], bool compute_hessian) {
    double cx, cy, cz, sx, sy, sz;
    if (std::fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p[3]); sx = std::sin(p[3]); }
    if (std::fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p[4]); sy = std::sin(p[4]); }
    if (std::fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p[5]); sz = std::sin(p[5]); }
    const double ja[8][3] = {{-sx * sz + cx * sy * cz, -sx * cz - cx * sy * sz, -cx * cy}, {cx * sz + sx * sy * cz, cx * cz - sx * sy * sz, -sx * cy},
                             {-sy * cz, sy * sz, cy}, {sx * cy * cz, -sx * cy * sz, sx * sy}, {-cx * cy * cz, cx * cy * sz, -cx * sy},
                             {-cy * sz, -cy * cz, 0}, {cx * cz - sx * sy * sz, -cx * sz - sx * sy * cz, 0}, {sx * cz + cx * sy * sz, cx * sy * cz - sx * sz, 0}};
    memcpy(h->ang.j, ja, sizeof(ja));
    if (compute_hessian) {
        const double ha[15][3] = {{-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, sx * cy}, {-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, -cx * cy},
                                  {cx * cy * cz, -cx * cy * sz, cx * sy}, {sx * cy * cz, -sx * cy * sz, sx * sy},
                                  {-sx * cz - cx * sy * sz, sx * sz - cx * sy * cz, 0}, {cx * cz - sx * sy * sz, -sx * sy * cz - cx * sz, 0},
                                  {-cy * cz, cy * sz, sy}, {-sx * sy * cz, sx * sy * sz, sx * cy}, {cx * sy * cz, -cx * sy * sz, -cx * cy},
                                  {sy * sz, sy * cz, 0}, {-sx * cy * sz, -sx * cy * cz, 0}, {cx * cy * sz, cx * cy * cz, 0},
                                  {-cy * cz, cy * sz, 0}, {-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, 0}, {-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, 0}};
        memcpy(h->ang.h, ha, sizeof(ha));
    }
}

static void ndt_gauss(b2_ndt_s* h) {
    const double c1 = 10 * (1 - h->outlier_ratio), c2 = h->outlier_ratio / std::pow((double)h->resolution, 3), d3 = -std::log(c2);
    h->gauss_d1 = -std::log(c1 + c2) - d3;
    h->gauss_d2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / h->gauss_d1);
}

static void ndt_fill_args(b2_ndt_s* h, NdtArgs& a, const float T[16], int mode) {
    a.g.inv_leaf = h->inv_leaf;
    a.g.r2 = (float)((double)h->resolution * (double)h->resolution);
    for (int d = 0; d < 3; d++) { a.g.min_b[d] = h->min_b[d]; a.g.div_b[d] = h->div_b[d]; }
    a.g.rank_of_cell = h->rank_of_cell.as<int32_t>();
    a.g.centroid = h->centroid.as<float4>(); a.g.mean = h->mean.as<double>(); a.g.icov = h->icov.as<double>();
    a.src = h->src_xyz.as<float>(); a.n_src = (uint32_t)h->n_src;
    for (int i = 0; i < 12; i++) a.T[i] = T[i];
    a.ang = h->ang; a.d1 = h->gauss_d1; a.d2 = h->gauss_d2; a.mode = mode;
    a.partials = h->partials.as<double>(); a.sums = h->sums.as<double>();
    a.ticket = reinterpret_cast<unsigned*>(h->sums.as<double>() + 32);
    a.host_sums = nullptr; a.seq = 0.0;
}

static int ndt_prepare(b2_ndt_s* h) {
    if (!h->have_tgt || !h->have_src) { set_error("ndt: setInputTarget and setInputSource first"); return B2_ERR_STATE; }
    if (h->grid_stale) B2_CHECK(ndt_build_target(h));
    const uint32_t chunks = (uint32_t)((h->n_src + 31) / 32);
    h->grid_blocks = (int)std::max<uint32_t>(1u, std::min<uint32_t>((chunks + NDT_WARPS - 1) / NDT_WARPS, (uint32_t)device_sm_count() * 2u));
    B2_CHECK(h->partials.reserve((size_t)device_sm_count() * 2 * NDT_NSUM * 8 + 256));
    if (!h->sums.p) { B2_CHECK(h->sums.reserve(64 * 8)); B2_CUDA(cudaMemsetAsync(h->sums.p, 0, 64 * 8, h->stream)); }
    B2_CHECK(h->pin.reserve(64 * 8));
    return B2_OK;
}

// one derivative pass at transform T (float 4x4) with the angular constants currently in h->ang
static int ndt_evaluate(b2_ndt_s* h, const float T[16], int mode, double* score, double grad[6], double hess[36]) {
    NdtArgs a;
    ndt_fill_args(h, a, T, mode);
    if (!h->h_res && !getenv("B2_NDT_NO_MAPPED")) {
        if (cudaHostAlloc(reinterpret_cast<void**>(&h->h_res), 64 * sizeof(double), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->d_res), h->h_res, 0) != cudaSuccess) { cudaGetLastError(); h->h_res = nullptr; h->d_res = nullptr; }
        else memset(h->h_res, 0, 64 * sizeof(double));
    }
    if (h->h_res) { h->res_seq += 1.0; a.host_sums = h->d_res; a.seq = h->res_seq; }
    k_ndt_derivatives<<<h->grid_blocks, NDT_THREADS, 0, h->stream>>>(a); count_launch(); h->launches++;
    B2_CUDA(cudaGetLastError());
    double* hs = h->pin.as<double>();
    bool got = false;
    if (h->h_res) {
        // one evaluation is ~50 us of kernel: spin on the sequence number the last CTA writes after the sums
        volatile double* flag = h->h_res + 32;
        for (long spin = 0; spin < 40000000L; spin++) {
            if (*flag == h->res_seq) { got = true; break; }
            if ((spin & 0xfff) == 0xfff && cudaStreamQuery(h->stream) != cudaErrorNotReady && *flag != h->res_seq) break;   // finished without the flag: an error
        }
        if (got) { std::atomic_thread_fence(std::memory_order_acquire); for (int i = 0; i < NDT_NSUM; i++) hs[i] = h->h_res[i]; }
    }
    if (!got) {
        B2_CUDA(cudaMemcpyAsync(hs, h->sums.p, NDT_NSUM * 8, cudaMemcpyDeviceToHost, h->stream));
        B2_CUDA(cudaStreamSynchronize(h->stream));
    }
    if (mode != 2) { if (score) *score = hs[0]; for (int i = 0; i < 6; i++) grad[i] = hs[1 + i]; }
    if (mode != 0) {
        int q = 7;
        for (int r = 0; r < 6; r++) for (int c = r; c < 6; c++) { hess[r * 6 + c] = hs[q]; hess[c * 6 + r] = hs[q]; q++; }
    }
    h->pairs_last = (long long)hs[28];
    h->evaluations++;
    return B2_OK;
}

static double ndt_psi(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
static double ndt_dpsi(double g_a, double g_0, double mu) { return g_a - mu * g_0; }
static bool ndt_update_interval(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t, double g_t) {
    if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
    if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
    if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
    return true;
}
static double ndt_trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
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

// ndt.hpp computeStepLengthMT; every trial step is one derivative pass on the device
static int ndt_step_length(b2_ndt_s* h, const double x[6], double step_dir[6], double step_init, double step_max, double step_min,
                           double& score, double grad[6], double hess[36], double* step_out) {
    const double phi_0 = -score;
    double d_phi_0 = 0;
    for (int i = 0; i < 6; i++) d_phi_0 += grad[i] * step_dir[i];
    d_phi_0 = -d_phi_0;
    double x_t[6];
    if (d_phi_0 >= 0) {
        if (d_phi_0 == 0) { *step_out = 0; return B2_OK; }
        d_phi_0 *= -1;
        for (int i = 0; i < 6; i++) step_dir[i] *= -1;
    }
    const int max_step_iterations = 10;
    int step_iterations = 0;
    const double mu = 1.e-4, nu = 0.9;
    double a_l = 0, a_u = 0;
    double f_l = ndt_psi(a_l, phi_0, phi_0, d_phi_0, mu), g_l = ndt_dpsi(d_phi_0, d_phi_0, mu);
    double f_u = ndt_psi(a_u, phi_0, phi_0, d_phi_0, mu), g_u = ndt_dpsi(d_phi_0, d_phi_0, mu);
    bool interval_converged = (step_max - step_min) < 0, open_interval = true;
    double a_t = step_init;
    a_t = std::min(a_t, step_max);
    a_t = std::max(a_t, step_min);
    for (int i = 0; i < 6; i++) x_t[i] = x[i] + step_dir[i] * a_t;
    pose_to_matrix_f(x_t, h->final_T);
    ndt_angle_derivatives(h, x_t, true);
    B2_CHECK(ndt_evaluate(h, h->final_T, 1, &score, grad, hess));
    double phi_t = -score, d_phi_t = 0;
    for (int i = 0; i < 6; i++) d_phi_t += grad[i] * step_dir[i];
    d_phi_t = -d_phi_t;
    double psi_t = ndt_psi(a_t, phi_t, phi_0, d_phi_0, mu), d_psi_t = ndt_dpsi(d_phi_t, d_phi_0, mu);
    while (!interval_converged && step_iterations < max_step_iterations && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
        if (open_interval) a_t = ndt_trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
        else a_t = ndt_trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
        a_t = std::min(a_t, step_max);
        a_t = std::max(a_t, step_min);
        for (int i = 0; i < 6; i++) x_t[i] = x[i] + step_dir[i] * a_t;
        pose_to_matrix_f(x_t, h->final_T);
        ndt_angle_derivatives(h, x_t, false);
        B2_CHECK(ndt_evaluate(h, h->final_T, 0, &score, grad, hess));
        phi_t = -score; d_phi_t = 0;
        for (int i = 0; i < 6; i++) d_phi_t += grad[i] * step_dir[i];
        d_phi_t = -d_phi_t;
        psi_t = ndt_psi(a_t, phi_t, phi_0, d_phi_0, mu); d_psi_t = ndt_dpsi(d_phi_t, d_phi_0, mu);
        if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
            open_interval = false;
            f_l = f_l + phi_0 - mu * d_phi_0 * a_l; g_l = g_l + mu * d_phi_0;
            f_u = f_u + phi_0 - mu * d_phi_0 * a_u; g_u = g_u + mu * d_phi_0;
        }
        if (open_interval) interval_converged = ndt_update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
        else interval_converged = ndt_update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
        step_iterations++;
    }
    // PCL's computeHessian: the angular second-derivative terms are the ones left by the last compute_hessian = true call
    if (step_iterations) B2_CHECK(ndt_evaluate(h, h->final_T, 2, nullptr, grad, hess));
    *step_out = a_t;
    return B2_OK;
}

static int ndt_set_cloud(b2_ndt_s* h, const void* xyz, size_t stride, size_t n, DevBuf& dst) {
    if ((n && !xyz) || stride < 12 || (stride & 3)) return B2_ERR_ARG;
    if (n > 0x7fffffffull) return B2_ERR_ARG;
    B2_CUDA(cudaSetDevice(h->device));
    B2_CHECK(dst.reserve(std::max<size_t>(n, 1) * 12));
    if (!n) return B2_OK;
    if (stride == 12) {
        B2_CUDA(cudaMemcpyAsync(dst.p, xyz, n * 12, cudaMemcpyHostToDevice, h->stream));
    } else {
        B2_CHECK(h->work.reserve(n * stride));
        B2_CUDA(cudaMemcpyAsync(h->work.p, xyz, n * stride, cudaMemcpyHostToDevice, h->stream));
        k_ndt_pack_xyz<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->work.as<unsigned char>(), stride, (uint32_t)n, dst.as<float>()); count_launch();
        B2_CUDA(cudaGetLastError());
    }
    B2_CUDA(cudaStreamSynchronize(h->stream));
    return B2_OK;
}

extern "C" {

int b2_ndt_create(b2_ndt_t* out) {
    if (!out) return B2_ERR_ARG;
    *out = nullptr;
    b2_ndt_s* h = new b2_ndt_s();
    if (cudaGetDevice(&h->device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->e0) != cudaSuccess || cudaEventCreate(&h->e1) != cudaSuccess) {
        set_error("b2_ndt_create: %s", cudaGetErrorString(cudaGetLastError()));
        delete h; return B2_ERR_CUDA;
    }
    *out = h;
    return B2_OK;
}

int b2_ndt_destroy(b2_ndt_t h) {
    if (!h) return B2_OK;
    h->tgt_xyz.release(); h->rank_of_cell.release(); h->centroid.release(); h->mean.release(); h->icov.release();
    h->vox_index.release(); h->vox_npts.release(); h->work.release(); h->src_xyz.release(); h->partials.release(); h->sums.release();
    h->bvh.release(); h->pin.release(); h->stage.release();
    if (h->h_res) cudaFreeHost(h->h_res);
    if (h->e0) cudaEventDestroy(h->e0);
    if (h->e1) cudaEventDestroy(h->e1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return B2_OK;
}

int b2_ndt_set_transformation_epsilon(b2_ndt_t h, double eps) { if (!h) return B2_ERR_ARG; h->trans_eps = eps; return B2_OK; }
int b2_ndt_set_step_size(b2_ndt_t h, double step) { if (!h) return B2_ERR_ARG; h->step_size = step; return B2_OK; }
int b2_ndt_set_resolution(b2_ndt_t h, float resolution) {
    if (!h || !(resolution > 0.f)) return B2_ERR_ARG;
    if (resolution != h->resolution) h->grid_stale = true;     // PCL re-initialises the voxel grid when the resolution changes
    h->resolution = resolution;
    return B2_OK;
}
int b2_ndt_set_maximum_iterations(b2_ndt_t h, int n) { if (!h) return B2_ERR_ARG; h->max_iterations = n; return B2_OK; }

int b2_ndt_set_input_target(b2_ndt_t h, const void* xyz, size_t stride, size_t n) {
    B2_NVTX("b2_ndt_set_input_target");
    if (!h) return B2_ERR_ARG;
    h->have_tgt = false; h->bvh_valid = false;
    B2_CHECK(ndt_set_cloud(h, xyz, stride, n, h->tgt_xyz));
    h->n_tgt = n;
    cudaEventRecord(h->e0, h->stream);
    B2_CHECK(ndt_build_target(h));
    cudaEventRecord(h->e1, h->stream); cudaEventSynchronize(h->e1);
    cudaEventElapsedTime(&h->last_ms, h->e0, h->e1);
    h->have_tgt = true;
    return B2_OK;
}

int b2_ndt_set_input_source(b2_ndt_t h, const void* xyz, size_t stride, size_t n) {
    B2_NVTX("b2_ndt_set_input_source");
    if (!h) return B2_ERR_ARG;
    h->have_src = false;
    B2_CHECK(ndt_set_cloud(h, xyz, stride, n, h->src_xyz));
    h->n_src = n;
    h->have_src = true;
    return B2_OK;
}

int b2_ndt_get_voxels(b2_ndt_t h, size_t capacity, size_t* n_voxels, int32_t* voxel_index, int32_t* n_points, float* centroid_xyz,
                      double* mean, double* inverse_covariance, int32_t min_b[3], int32_t div_b[3]) {
    if (!h || !h->have_tgt) return B2_ERR_STATE;
    B2_CUDA(cudaSetDevice(h->device));
    if (h->grid_stale) B2_CHECK(ndt_build_target(h));
    if (n_voxels) *n_voxels = h->n_vox;
    if (min_b) for (int d = 0; d < 3; d++) min_b[d] = h->min_b[d];
    if (div_b) for (int d = 0; d < 3; d++) div_b[d] = h->div_b[d];
    const size_t m = std::min<size_t>(capacity, h->n_vox);
    if (!m) return B2_OK;
    if (voxel_index) B2_CUDA(cudaMemcpyAsync(voxel_index, h->vox_index.p, m * 4, cudaMemcpyDeviceToHost, h->stream));
    if (n_points) B2_CUDA(cudaMemcpyAsync(n_points, h->vox_npts.p, m * 4, cudaMemcpyDeviceToHost, h->stream));
    if (mean) B2_CUDA(cudaMemcpyAsync(mean, h->mean.p, m * 24, cudaMemcpyDeviceToHost, h->stream));
    if (inverse_covariance) B2_CUDA(cudaMemcpyAsync(inverse_covariance, h->icov.p, m * 72, cudaMemcpyDeviceToHost, h->stream));
    if (centroid_xyz) {
        std::vector<float> tmp(m * 4);
        B2_CUDA(cudaMemcpyAsync(tmp.data(), h->centroid.p, m * 16, cudaMemcpyDeviceToHost, h->stream));
        B2_CUDA(cudaStreamSynchronize(h->stream));
        for (size_t i = 0; i < m; i++) for (int d = 0; d < 3; d++) centroid_xyz[i * 3 + d] = tmp[i * 4 + d];
    }
    B2_CUDA(cudaStreamSynchronize(h->stream));
    return B2_OK;
}

int b2_ndt_derivatives(b2_ndt_t h, const double p[6], double* score, double gradient[6], double hessian[36], long long* n_pairs) {
    if (!h || !p || !gradient) return B2_ERR_ARG;
    B2_CUDA(cudaSetDevice(h->device));
    B2_CHECK(ndt_prepare(h));
    ndt_gauss(h);
    float T[16];
    pose_to_matrix_f(p, T);
    ndt_angle_derivatives(h, p, true);
    double H[36], s = 0;
    B2_CHECK(ndt_evaluate(h, T, 1, &s, gradient, H));
    if (score) *score = s;
    if (hessian) memcpy(hessian, H, sizeof(H));
    if (n_pairs) *n_pairs = h->pairs_last;
    return B2_OK;
}

/* ndt.hpp computeTransformation */
int b2_ndt_align(b2_ndt_t h, const float guess[16], void* out_cloud, size_t out_stride) {
    B2_NVTX("b2_ndt_align");
    if (!h || !guess) return B2_ERR_ARG;
    B2_CUDA(cudaSetDevice(h->device));
    B2_CHECK(ndt_prepare(h));
    h->nr_iterations = 0; h->converged = false; h->evaluations = 0; h->launches = 0;
    ndt_gauss(h);
    cudaEventRecord(h->e0, h->stream);
    static const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    memcpy(h->final_T, I, sizeof(I));
    if (memcmp(guess, I, sizeof(I)) != 0) memcpy(h->final_T, guess, sizeof(I));
    double p[6], delta_p[6], grad[6], hess[36];
    float rot[3];
    euler_012_f(h->final_T, rot);
    p[0] = h->final_T[3]; p[1] = h->final_T[7]; p[2] = h->final_T[11];
    p[3] = rot[0]; p[4] = rot[1]; p[5] = rot[2];
    double score = 0;
    ndt_angle_derivatives(h, p, true);
    B2_CHECK(ndt_evaluate(h, h->final_T, 1, &score, grad, hess));       // the cloud transformed by the guess itself
    const double npts = (double)h->n_src;
    bool early = false;
    while (!h->converged) {
        double neg[6];
        for (int i = 0; i < 6; i++) neg[i] = -grad[i];
        svd_solve6_host(hess, neg, delta_p);
        double nrm = 0;
        for (int i = 0; i < 6; i++) nrm += delta_p[i] * delta_p[i];
        double delta_p_norm = std::sqrt(nrm);
        if (delta_p_norm == 0 || delta_p_norm != delta_p_norm) {
            h->trans_probability = score / npts;
            h->converged = delta_p_norm == delta_p_norm;
            early = true;
            break;
        }
        for (int i = 0; i < 6; i++) delta_p[i] /= delta_p_norm;
        double step = 0;
        B2_CHECK(ndt_step_length(h, p, delta_p, delta_p_norm, h->step_size, h->trans_eps / 2, score, grad, hess, &step));
        delta_p_norm = step;
        for (int i = 0; i < 6; i++) delta_p[i] *= delta_p_norm;
        for (int i = 0; i < 6; i++) p[i] = p[i] + delta_p[i];
        if (h->nr_iterations > h->max_iterations || (h->nr_iterations && (std::fabs(delta_p_norm) < h->trans_eps))) h->converged = true;
        h->nr_iterations++;
    }
    if (!early) h->trans_probability = score / npts;
    if (out_cloud && h->n_src) {
        if (out_stride < 12 || (out_stride & 3)) return B2_ERR_ARG;
        NdtArgs a;
        ndt_fill_args(h, a, h->final_T, 0);
        B2_CHECK(h->work.reserve(h->n_src * out_stride));
        B2_CUDA(cudaMemsetAsync(h->work.p, 0, h->n_src * out_stride, h->stream));
        k_ndt_transform_out<<<(unsigned)((h->n_src + 255) / 256), 256, 0, h->stream>>>(a.src, a.n_src, a, h->work.as<unsigned char>(), out_stride); count_launch();
        B2_CUDA(cudaMemcpyAsync(out_cloud, h->work.p, h->n_src * out_stride, cudaMemcpyDeviceToHost, h->stream));
    }
    cudaEventRecord(h->e1, h->stream);
    B2_CUDA(cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&h->last_ms, h->e0, h->e1);
    return B2_OK;
}

int b2_ndt_has_converged(b2_ndt_t h, int* converged) { if (!h || !converged) return B2_ERR_ARG; *converged = h->converged ? 1 : 0; return B2_OK; }
int b2_ndt_get_final_transformation(b2_ndt_t h, float T[16]) { if (!h || !T) return B2_ERR_ARG; memcpy(T, h->final_T, 64); return B2_OK; }
int b2_ndt_get_transformation_probability(b2_ndt_t h, double* prob) { if (!h || !prob) return B2_ERR_ARG; *prob = h->trans_probability; return B2_OK; }
int b2_ndt_get_final_num_iteration(b2_ndt_t h, int* n) { if (!h || !n) return B2_ERR_ARG; *n = h->nr_iterations; return B2_OK; }

/* pcl::Registration::getFitnessScore(): mean squared distance from every transformed source point to its nearest target point */
int b2_ndt_get_fitness_score(b2_ndt_t h, double* score) {
    B2_NVTX("b2_ndt_get_fitness_score");
    if (!h || !score) return B2_ERR_ARG;
    if (!h->have_tgt || !h->have_src) return B2_ERR_STATE;
    B2_CUDA(cudaSetDevice(h->device));
    *score = DBL_MAX;
    if (!h->n_tgt || !h->n_src) return B2_OK;
    cudaStream_t s = h->stream;
    if (!h->bvh_valid) {
        DevBuf wide;
        B2_CHECK(wide.reserve(h->n_tgt * 24));
        k_ndt_widen<<<(unsigned)((3 * h->n_tgt + 255) / 256), 256, 0, s>>>(h->tgt_xyz.as<float>(), (uint32_t)h->n_tgt, wide.as<double>()); count_launch();
        const int st = h->bvh.build(wide.as<double>(), h->n_tgt, h->work, s);
        cudaStreamSynchronize(s);
        wide.release();
        if (st != B2_OK) return st;
        h->bvh_valid = true;
    }
    if (!h->bvh.dev.n) return B2_OK;
    B2_CHECK(ndt_prepare(h));
    NdtArgs a;
    ndt_fill_args(h, a, h->final_T, 0);
    double* acc = h->sums.as<double>() + 40;
    B2_CUDA(cudaMemsetAsync(acc, 0, 16, s));
    const uint32_t warps = (uint32_t)((h->n_src + NDT_FIT_QPW - 1) / NDT_FIT_QPW);
    k_ndt_fitness<<<(warps + 3) / 4, 128, 0, s>>>(h->bvh.dev, a.src, a.n_src, a, acc); count_launch();
    B2_CUDA(cudaGetLastError());
    double hacc[2];
    B2_CUDA(cudaMemcpyAsync(hacc, acc, 16, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    if (hacc[1] > 0) *score = hacc[0] / hacc[1];
    return B2_OK;
}

int b2_ndt_last_gpu_ms(b2_ndt_t h, float* ms, int* launches, int* evaluations, long long* pairs_last) {
    if (!h) return B2_ERR_ARG;
    if (ms) *ms = h->last_ms;
    if (launches) *launches = h->launches;
    if (evaluations) *evaluations = h->evaluations;
    if (pairs_last) *pairs_last = h->pairs_last;
    return B2_OK;
}

}  // extern "C"
