// LIO-SAM per-scan front end on the device: range-image projection with first-hit occupancy, IMU-rotation deskew,
// per-ring compaction, curvature / occlusion masks, per-ring x 6-sector feature selection and the per-ring VoxelGrid of
// the surf candidates.
//
// Replaces (liosam_ws/src/LIO-SAM/src):
//   imageProjection.cpp   findRotation :446-471, deskewPoint :489-519, projectPointCloud :521-572, cloudExtraction :574-598
//   featureExtraction.cpp calculateSmoothness :81-101, markOccludedPoints :103-139, extractFeatures :141-238
//
// Sequential semantics of the reference, made parallel:
//   first hit wins a range-image cell      -> atomicMin of the point index per cell, then one thread per cell
//   transStartInverse = first accepted pt  -> atomicMin over accepted indices, one thread builds the inverse
//   ring compaction                        -> per-ring counts, then per-ring block scan at the ring's base offset
//   sort + greedy suppression per sector   -> one CTA per ring (rings are independent, sectors of a ring are not):
//                                             bitonic sort of (curvature, index) in shared memory, one lane walks the
//                                             greedy passes with early exit, the CTA compacts the surf candidates
//   per-ring VoxelGrid (featureExtraction.cpp:233-234) -> same CTA: bbox, PCL's index arithmetic, bitonic sort of
//                                             (voxel, position), one thread per voxel sums in position order
// Expression types follow SURVEY.md Appendix A; -fmad=false keeps every threshold operand at the reference's rounding.
#include "b2_common.cuh"
#include "b2_atan2f.cuh"
#include <cfloat>
#include <climits>
#include <cmath>
#include <algorithm>
#include <vector>

namespace b2 {

struct ScanDev {
    int n_scan, H, downsample;
    float rmin, rmax, ang_res_x;
    float edge_th, surf_th, surf_leaf;
};

__device__ __forceinline__ float sin_rnf(float a) { return (float)sin((double)a); }
__device__ __forceinline__ float cos_rnf(float a) { return (float)cos((double)a); }

// pcl::getTransformation(0,0,0,roll,pitch,yaw) rows (translation zero)
__device__ __forceinline__ void rot_affine(float roll, float pitch, float yaw, float* t) {
    float A = cos_rnf(yaw), B = sin_rnf(yaw), C = cos_rnf(pitch), D = sin_rnf(pitch), E = cos_rnf(roll), F = sin_rnf(roll);
    float DE = D * E, DF = D * F;
    t[0] = A * C;  t[1] = A * DF - B * E;  t[2]  = B * F + A * DE;  t[3]  = 0.f;
    t[4] = B * C;  t[5] = A * E + B * DF;  t[6]  = B * DE - A * F;  t[7]  = 0.f;
    t[8] = -D;     t[9] = C * F;           t[10] = C * E;           t[11] = 0.f;
}

// findRotation (imageProjection.cpp:446-471); cur = imuPointerCur (last valid index)
__device__ __forceinline__ void find_rotation(const double* __restrict__ t, const double* __restrict__ rx, const double* __restrict__ ry,
                                              const double* __restrict__ rz, int cur, double pointTime, float& ox, float& oy, float& oz) {
    int front = 0;
    while (front < cur) {
        if (pointTime < t[front]) break;
        ++front;
    }
    if (pointTime > t[front] || front == 0) {
        ox = (float)rx[front]; oy = (float)ry[front]; oz = (float)rz[front];
    } else {
        const int back = front - 1;
        const double ratioFront = (pointTime - t[back]) / (t[front] - t[back]);
        const double ratioBack = (t[front] - pointTime) / (t[front] - t[back]);
        ox = (float)(rx[front] * ratioFront + rx[back] * ratioBack);
        oy = (float)(ry[front] * ratioFront + ry[back] * ratioBack);
        oz = (float)(rz[front] * ratioFront + rz[back] * ratioBack);
    }
}

struct RawPoint { float x, y, z, intensity, time; int ring; };
__device__ __forceinline__ RawPoint load_raw(const unsigned char* __restrict__ raw, size_t i) {
    const uint4* p = reinterpret_cast<const uint4*>(raw + i * 32);     // PointXYZIRT is 32 B, 16 B aligned
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    RawPoint r;
    r.x = __uint_as_float(a.x); r.y = __uint_as_float(a.y); r.z = __uint_as_float(a.z);
    r.intensity = __uint_as_float(b.x); r.ring = (int)(b.y & 0xffffu); r.time = __uint_as_float(b.z);
    return r;
}

__global__ void __launch_bounds__(256) k_scan_init(int* __restrict__ winner, int cells, int* __restrict__ first_idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cells) winner[i] = INT_MAX;
    if (i == 0) *first_idx = INT_MAX;
}

// projectPointCloud, everything before the first-hit test: one thread per raw point
__global__ void __launch_bounds__(256) k_scan_project(const unsigned char* __restrict__ raw, int n, ScanDev s, int* __restrict__ winner,
                                                      int* __restrict__ first_idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int cell = -1;
    if (i < n) {
        const RawPoint p = load_raw(raw, i);
        const float range = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
        const int row = p.ring;
        // the reference requires a dense cloud (:264-268); non-finite points are dropped here
        bool ok = isfinite(p.x) && isfinite(p.y) && isfinite(p.z) && !(range < s.rmin || range > s.rmax) && row >= 0 && row < s.n_scan;
        ok = ok && (row % s.downsample == 0);
        if (ok) {
            // atan2 on float arguments (glibc's atan2f, restated in b2_atan2f.cuh), then *180 in float, /M_PI in double, narrowed to float (:547)
            const float at = port_atan2f(p.x, p.y);
            const float horizonAngle = (float)((double)(at * 180) / M_PI);
            int col = (int)(-round(((double)horizonAngle - 90.0) / (double)s.ang_res_x) + (double)(s.H / 2));
            if (col >= s.H) col -= s.H;
            if (col >= 0 && col < s.H) cell = row * s.H + col;
        }
    }
    if (cell >= 0) atomicMin(&winner[cell], i);
    // the first accepted point in input order seeds transStartInverse: one atomic per warp
    const int m = __reduce_min_sync(0xffffffffu, cell >= 0 ? i : INT_MAX);
    if ((threadIdx.x & 31) == 0 && m != INT_MAX) atomicMin(first_idx, m);
}

__device__ __forceinline__ void affine_inverse(const float* a, float* o) {
    const float m00 = a[0], m01 = a[1], m02 = a[2], m10 = a[4], m11 = a[5], m12 = a[6], m20 = a[8], m21 = a[9], m22 = a[10];
    const float c00 = m11 * m22 - m12 * m21, c10 = m12 * m20 - m10 * m22, c20 = m10 * m21 - m11 * m20;
    const float det = c00 * m00 + c10 * m01 + c20 * m02;
    const float invdet = 1.0f / det;
    o[0] = c00 * invdet; o[1] = (m02 * m21 - m01 * m22) * invdet; o[2]  = (m01 * m12 - m02 * m11) * invdet;
    o[4] = c10 * invdet; o[5] = (m00 * m22 - m02 * m20) * invdet; o[6]  = (m02 * m10 - m00 * m12) * invdet;
    o[8] = c20 * invdet; o[9] = (m01 * m20 - m00 * m21) * invdet; o[10] = (m00 * m11 - m01 * m10) * invdet;
    o[3]  = -(o[0] * a[3] + o[1] * a[7] + o[2] * a[11]);
    o[7]  = -(o[4] * a[3] + o[5] * a[7] + o[6] * a[11]);
    o[11] = -(o[8] * a[3] + o[9] * a[7] + o[10] * a[11]);
}

// transStartInverse from the first accepted point (deskewPoint :502-506)
__global__ void k_scan_start_inverse(const unsigned char* __restrict__ raw, const int* __restrict__ first_idx,
                                     const double* __restrict__ it, const double* __restrict__ irx, const double* __restrict__ iry,
                                     const double* __restrict__ irz, int cur, double t_scan, float* __restrict__ start_inv) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int i = *first_idx;
    if (i == INT_MAX) return;
    const RawPoint p = load_raw(raw, i);
    float rx, ry, rz, tf[12];
    find_rotation(it, irx, iry, irz, cur, t_scan + (double)p.time, rx, ry, rz);
    rot_affine(rx, ry, rz, tf);
    float inv[12];
    affine_inverse(tf, inv);
    for (int k = 0; k < 12; k++) start_inv[k] = inv[k];
}

// one thread per range-image cell: the cell's winner is deskewed and written (rest of projectPointCloud)
__global__ void __launch_bounds__(256) k_scan_cells(const unsigned char* __restrict__ raw, const int* __restrict__ winner, int cells,
                                                    const double* __restrict__ it, const double* __restrict__ irx, const double* __restrict__ iry,
                                                    const double* __restrict__ irz, int cur, double t_scan, int deskew,
                                                    const float* __restrict__ start_inv, float* __restrict__ range_mat, float4* __restrict__ full_cloud) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const int w = winner[c];
    if (w == INT_MAX) {
        range_mat[c] = FLT_MAX;
        const float qn = __int_as_float(0x7fc00000);
        full_cloud[c] = make_float4(qn, qn, qn, qn);
        return;
    }
    const RawPoint p = load_raw(raw, w);
    const float range = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
    float nx = p.x, ny = p.y, nz = p.z;
    if (deskew) {
        float rx, ry, rz, tf[12], bt[12];
        find_rotation(it, irx, iry, irz, cur, t_scan + (double)p.time, rx, ry, rz);
        rot_affine(rx, ry, rz, tf);
#pragma unroll
        for (int r = 0; r < 3; r++) {
#pragma unroll
            for (int q = 0; q < 3; q++)
                bt[r * 4 + q] = start_inv[r * 4 + 0] * tf[0 * 4 + q] + start_inv[r * 4 + 1] * tf[1 * 4 + q] + start_inv[r * 4 + 2] * tf[2 * 4 + q];
            bt[r * 4 + 3] = start_inv[r * 4 + 0] * tf[3] + start_inv[r * 4 + 1] * tf[7] + start_inv[r * 4 + 2] * tf[11] + start_inv[r * 4 + 3];
        }
        nx = bt[0] * p.x + bt[1] * p.y + bt[2] * p.z + bt[3];
        ny = bt[4] * p.x + bt[5] * p.y + bt[6] * p.z + bt[7];
        nz = bt[8] * p.x + bt[9] * p.y + bt[10] * p.z + bt[11];
    }
    range_mat[c] = range;
    full_cloud[c] = make_float4(nx, ny, nz, p.intensity);
}

// cloudExtraction, pass 1: occupied cells per ring
__global__ void __launch_bounds__(256) k_scan_ring_count(const int* __restrict__ winner, int H, int* __restrict__ ring_count) {
    __shared__ int ws[8];
    const int ring = blockIdx.x;
    int c = 0;
    for (int j = threadIdx.x; j < H; j += blockDim.x) c += winner[ring * H + j] != INT_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) t += ws[w]; ring_count[ring] = t; }
}

// cloudExtraction, pass 2: ordered compaction of each ring at its base offset
__global__ void __launch_bounds__(256) k_scan_ring_compact(const int* __restrict__ winner, const float* __restrict__ range_mat,
                                                           const float4* __restrict__ full_cloud, int H, int n_scan, const int* __restrict__ ring_count,
                                                           float4* __restrict__ extracted, int* __restrict__ col_ind, float* __restrict__ pt_range,
                                                           int* __restrict__ start_ring, int* __restrict__ end_ring, int* __restrict__ total) {
    __shared__ int ws[9];
    __shared__ int s_base;
    const int ring = blockIdx.x;
    if (threadIdx.x == 0) {
        int b = 0;
        for (int r = 0; r < ring; r++) b += ring_count[r];
        s_base = b;
        start_ring[ring] = b - 1 + 5;
        end_ring[ring] = b + ring_count[ring] - 1 - 5;
        if (ring == n_scan - 1) *total = b + ring_count[ring];
    }
    __syncthreads();
    int run = s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j0 = 0; j0 < H; j0 += 256) {
        const int j = j0 + threadIdx.x;
        const int occ = (j < H) && winner[ring * H + j] != INT_MAX;
        const unsigned bal = __ballot_sync(0xffffffffu, occ);
        if (lane == 0) ws[warp] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) { int c = ws[w]; ws[w] = t; t += c; } ws[8] = t; }
        __syncthreads();
        if (occ) {
            const int pos = run + ws[warp] + __popc(bal & ((1u << lane) - 1u));
            const int cell = ring * H + j;
            col_ind[pos] = j;
            pt_range[pos] = range_mat[cell];
            extracted[pos] = full_cloud[cell];
        }
        run += ws[8];
        __syncthreads();
    }
}

// calculateSmoothness + the zeroing the oracle defines for the never-written slots
__global__ void __launch_bounds__(256) k_scan_curvature(const float* __restrict__ r, const int* __restrict__ total, float* __restrict__ curv,
                                                        int* __restrict__ picked, int* __restrict__ label) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int M = *total;
    if (i >= M) return;
    float cv = 0.f;
    if (i >= 5 && i < M - 5) {
        float d = r[i - 5] + r[i - 4] + r[i - 3] + r[i - 2] + r[i - 1] - r[i] * 10 + r[i + 1] + r[i + 2] + r[i + 3] + r[i + 4] + r[i + 5];
        cv = d * d;
    }
    curv[i] = cv; picked[i] = 0; label[i] = 0;
}

// markOccludedPoints: every mark is a store of 1, so the parallel order does not matter
__global__ void __launch_bounds__(256) k_scan_masks(const float* __restrict__ r, const int* __restrict__ col, const int* __restrict__ total,
                                                    int* __restrict__ picked) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 5;
    const int M = *total;
    if (i >= M - 6) return;
    const float depth1 = r[i], depth2 = r[i + 1];
    const int columnDiff = abs(col[i + 1] - col[i]);
    if (columnDiff < 10) {
        if ((double)(depth1 - depth2) > 0.3) {
#pragma unroll
            for (int k = 0; k <= 5; k++) picked[i - k] = 1;
        } else if ((double)(depth2 - depth1) > 0.3) {
#pragma unroll
            for (int k = 1; k <= 6; k++) picked[i + k] = 1;
        }
    }
    const float diff1 = fabsf(r[i - 1] - r[i]);
    const float diff2 = fabsf(r[i + 1] - r[i]);
    if ((double)diff1 > 0.02 * (double)r[i] && (double)diff2 > 0.02 * (double)r[i]) picked[i] = 1;
}

// ---- shared-memory bitonic sort of 64-bit keys, ascending; P is a power of two, all threads of the CTA call it
__device__ void bitonic_sort_u64(unsigned long long* k, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = k[lo], b = k[hi];
                if ((a > b) == up) { k[lo] = b; k[hi] = a; }
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void suppress(int ind, int M, const int* __restrict__ col, int* __restrict__ picked) {
    for (int l = 1; l <= 5; l++) {
        if (ind + l >= M) break;
        if (abs(col[ind + l] - col[ind + l - 1]) > 10) break;
        picked[ind + l] = 1;
    }
    for (int l = -1; l >= -5; l--) {
        if (ind + l < 0) break;
        if (abs(col[ind + l] - col[ind + l + 1]) > 10) break;
        picked[ind + l] = 1;
    }
}

// The same for a whole warp walking one candidate: lanes 0-4 look at ind+1..ind+5, lanes 5-9 at ind-1..ind-5; the "stop at the first
// column gap" of the two loops is the lowest set bit of a ballot. All lanes pass the same ind. Ends with a warp barrier so the
// next candidate's picked test sees the marks.
__device__ __forceinline__ void suppress_warp(int ind, int M, const int* col, int* picked, int lane) {
    const bool up = lane < 5, down = lane >= 5 && lane < 10;
    const int l = up ? lane + 1 : -(lane - 4);
    bool stop = false;
    if (up) stop = (ind + l >= M) || abs(col[min(ind + l, M - 1)] - col[min(ind + l, M - 1) - 1]) > 10;
    if (down) stop = (ind + l < 0) || abs(col[max(ind + l, 0)] - col[max(ind + l, 0) + 1]) > 10;
    const unsigned b = __ballot_sync(0xffffffffu, stop);
    const unsigned bu = b & 0x1fu, bd = (b >> 5) & 0x1fu;
    const int first_up = bu ? __ffs(bu) - 1 : 5, first_down = bd ? __ffs(bd) - 1 : 5;      // offsets 1..first are marked
    if (up && lane < first_up) picked[ind + l] = 1;
    if (down && (lane - 5) < first_down) picked[ind + l] = 1;
    __syncwarp();
}

// extractFeatures: one CTA per ring. Dynamic shared memory: keys[P] (u64), pts[P] (float4) with P = pow2 >= H.
// Per-ring outputs go to fixed-capacity slots (corner: 120 per ring, surf: H per ring) and are concatenated afterwards.
__global__ void __launch_bounds__(256) k_scan_features(const float4* __restrict__ extracted, const int* __restrict__ col, const int* __restrict__ total,
                                                       const int* __restrict__ start_ring, const int* __restrict__ end_ring, ScanDev s, int P,
                                                       const float* __restrict__ curv, int* __restrict__ picked, int* __restrict__ label,
                                                       int* __restrict__ corner_idx, int* __restrict__ corner_cnt,
                                                       int* __restrict__ surf_idx, int* __restrict__ surf_cnt,
                                                       float4* __restrict__ surf_ds, int* __restrict__ surf_ds_cnt) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);
    float4* spts = reinterpret_cast<float4*>(smem + (size_t)P * 8);
    __shared__ int s_ncorner, s_nsurf, s_scan[9];
    __shared__ float s_mn[3], s_mx[3];
    __shared__ int s_geom[8];
    __shared__ float s_inv;
    const int ring = blockIdx.x;
    const int M = *total;
    const int sri = start_ring[ring], eri = end_ring[ring];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_ncorner = 0; s_nsurf = 0; }
    // The greedy walks below are one thread stepping through a sector with dependent reads of curvature / picked / column and the
    // +-5 suppression writes: staged in shared memory they run at ~30 cycles per access instead of a global round trip each
    // (this kernel was 82 % of the C2 front end). Window = the ring's own points and one neighbour on the low side, which is
    // all the walks can touch (candidates lie in [sri, eri] = [first + 4, last - 5], suppression reaches 5 further).
    const int Wcap = s.H + 16;
    float* g_curv_s = reinterpret_cast<float*>(smem + (size_t)P * 24);
    int* g_pick_s = reinterpret_cast<int*>(g_curv_s + Wcap);
    int* g_col_s = g_pick_s + Wcap;
    int* g_lab_s = g_col_s + Wcap;
    const int w0 = max(sri - 5, 0), w1 = min(eri + 5, M - 1);
    for (int k = w0 + threadIdx.x; k <= w1 && k - w0 < Wcap; k += blockDim.x) {
        g_curv_s[k - w0] = curv[k]; g_pick_s[k - w0] = picked[k]; g_col_s[k - w0] = col[k]; g_lab_s[k - w0] = label[k];
    }
    const float* curv_l = g_curv_s - w0;
    int* picked_l = g_pick_s - w0;
    const int* col_l = g_col_s - w0;
    int* label_l = g_lab_s - w0;
    __syncthreads();
    for (int j = 0; j < 6; j++) {
        const int sp = (sri * (6 - j) + eri * j) / 6;
        const int ep = (sri * (5 - j) + eri * (j + 1)) / 6 - 1;
        if (sp >= ep) continue;                                   // uniform for the CTA
        const int n = ep - sp;                                    // [sp, ep) is sorted, ep itself is not (:162)
        int Q = 1; while (Q < n) Q <<= 1;
        for (int t = threadIdx.x; t < Q; t += blockDim.x) {
            unsigned long long key = ~0ull;
            if (t < n) {
                const int k = sp + t;
                const float v = (k >= 5 && k < M - 5) ? curv_l[k] : 0.f;      // oracle definition outside [5, M-5)
                key = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)k;
            }
            keys[t] = key;
        }
        bitonic_sort_u64(keys, Q);
        if (warp == 0) {
            // One warp walks the sector (every lane runs the same control flow on the same shared-memory words; the +-5 suppression is
            // spread over ten lanes). descending: at most 20 corners with curvature > edgeThreshold (:165-195). The sorted part is
            // descending in curvature, so the first candidate at or below the threshold ends the walk (nothing later can pass).
            int largestPickedNum = 0;
            int nc = s_ncorner;
            for (int k = ep; k >= sp; k--) {
                const int ind = (k == ep) ? ep : (int)(unsigned)keys[k - sp];
                const float cv = curv_l[ind];
                if (k != ep && !(cv > s.edge_th)) break;
                if (picked_l[ind] == 0 && cv > s.edge_th) {
                    largestPickedNum++;
                    if (largestPickedNum <= 20) { if (lane == 0) { label_l[ind] = 1; corner_idx[ring * 120 + nc] = ind; } nc++; }
                    else break;
                    if (lane == 0) picked_l[ind] = 1;
                    suppress_warp(ind, M, col_l, picked_l, lane);
                }
            }
            __syncwarp();
            if (lane == 0) s_ncorner = nc;
            // ascending: surf labels (:197-225); ep is visited last
            for (int k = sp; k <= ep; k++) {
                const int ind = (k == ep) ? ep : (int)(unsigned)keys[k - sp];
                const float cv = curv_l[ind];
                if (k != ep && !(cv < s.surf_th)) { k = ep - 1; continue; }     // jump to the unsorted tail element
                if (picked_l[ind] == 0 && cv < s.surf_th) {
                    if (lane == 0) { label_l[ind] = -1; picked_l[ind] = 1; }
                    suppress_warp(ind, M, col_l, picked_l, lane);
                }
            }
            __syncwarp();
        }
        __syncthreads();
        // surf candidates of the sector, in index order (:227-232)
        int run = s_nsurf;
        for (int k0 = sp; k0 <= ep; k0 += 256) {
            const int k = k0 + threadIdx.x;
            const int f = (k <= ep) && label_l[k] <= 0;
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (lane == 0) s_scan[warp] = __popc(bal);
            __syncthreads();
            if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) { int c = s_scan[w]; s_scan[w] = t; t += c; } s_scan[8] = t; }
            __syncthreads();
            if (f) surf_idx[ring * s.H + run + s_scan[warp] + __popc(bal & ((1u << lane) - 1u))] = k;
            run += s_scan[8];
            __syncthreads();
        }
        if (threadIdx.x == 0) s_nsurf = run;
        __syncthreads();
    }
    const int ns = s_nsurf;
    if (threadIdx.x == 0) { corner_cnt[ring] = s_ncorner; surf_cnt[ring] = ns; }
    // labels and picked flags of the ring's own points go back to memory (cloudLabel is an output of the stage)
    for (int k = max(sri - 4, w0) + threadIdx.x; k <= w1; k += blockDim.x) { label[k] = label_l[k]; picked[k] = picked_l[k]; }
    // ---------------- per-ring VoxelGrid of the surf candidates (downSizeFilter, :233-236)
    if (ns == 0) { if (threadIdx.x == 0) surf_ds_cnt[ring] = 0; return; }
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int t = threadIdx.x; t < ns; t += blockDim.x) {
        const float4 p = extracted[surf_idx[ring * s.H + t]];
        spts[t] = p;
        mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
        mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    }
    __shared__ float s_red[8][6];
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    if (lane == 0) { for (int d = 0; d < 3; d++) { s_red[warp][d] = mn[d]; s_red[warp][3 + d] = mx[d]; } }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int d = 0; d < 3; d++) {
            float a = s_red[0][d], b = s_red[0][3 + d];
            for (int w = 1; w < 8; w++) { a = fminf(a, s_red[w][d]); b = fmaxf(b, s_red[w][3 + d]); }
            s_mn[d] = a; s_mx[d] = b;
        }
        const float inv = 1.0f / s.surf_leaf;
        s_inv = inv;
        long long dx = (long long)((s_mx[0] - s_mn[0]) * inv) + 1, dy = (long long)((s_mx[1] - s_mn[1]) * inv) + 1, dz = (long long)((s_mx[2] - s_mn[2]) * inv) + 1;
        s_geom[6] = (dx * dy * dz > (long long)INT_MAX) ? 1 : 0;             // PCL refuses: output = input
        int minb[3], maxb[3];
        for (int d = 0; d < 3; d++) { minb[d] = (int)floorf(s_mn[d] * inv); maxb[d] = (int)floorf(s_mx[d] * inv); s_geom[d] = minb[d]; }
        s_geom[3] = 1; s_geom[4] = maxb[0] - minb[0] + 1; s_geom[5] = (maxb[0] - minb[0] + 1) * (maxb[1] - minb[1] + 1);
    }
    __syncthreads();
    if (s_geom[6]) {
        for (int t = threadIdx.x; t < ns; t += blockDim.x) surf_ds[ring * s.H + t] = spts[t];
        if (threadIdx.x == 0) surf_ds_cnt[ring] = ns;
        return;
    }
    int Q = 1; while (Q < ns) Q <<= 1;
    for (int t = threadIdx.x; t < Q; t += blockDim.x) {
        unsigned long long key = ~0ull;
        if (t < ns) {
            const float4 p = spts[t];
            const int i0 = (int)(floorf(p.x * s_inv) - (float)s_geom[0]);
            const int i1 = (int)(floorf(p.y * s_inv) - (float)s_geom[1]);
            const int i2 = (int)(floorf(p.z * s_inv) - (float)s_geom[2]);
            const int idx = i0 * s_geom[3] + i1 * s_geom[4] + i2 * s_geom[5];
            key = ((unsigned long long)(unsigned)idx << 32) | (unsigned)t;
        }
        keys[t] = key;
    }
    bitonic_sort_u64(keys, Q);
    // one thread per run head: centroid in position order, rank = number of heads before it
    int run = 0;
    for (int t0 = 0; t0 < ns; t0 += 256) {
        const int t = t0 + threadIdx.x;
        const int head = (t < ns) && (t == 0 || (unsigned)(keys[t] >> 32) != (unsigned)(keys[t - 1] >> 32));
        const unsigned bal = __ballot_sync(0xffffffffu, head);
        if (lane == 0) s_scan[warp] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) { int q = 0; for (int w = 0; w < 8; w++) { int c = s_scan[w]; s_scan[w] = q; q += c; } s_scan[8] = q; }
        __syncthreads();
        if (head) {
            const unsigned vox = (unsigned)(keys[t] >> 32);
            float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
            int e = t;
            for (; e < ns && (unsigned)(keys[e] >> 32) == vox; e++) {
                const float4 p = spts[(unsigned)keys[e]];
                sx += p.x; sy += p.y; sz += p.z; si += p.w;
            }
            const float cnt = (float)(e - t);
            surf_ds[ring * s.H + run + s_scan[warp] + __popc(bal & ((1u << lane) - 1u))] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
        }
        run += s_scan[8];
        __syncthreads();
    }
    if (threadIdx.x == 0) surf_ds_cnt[ring] = run;
}

// concatenate the per-ring slots in ring order
__global__ void __launch_bounds__(256) k_scan_gather_out(const float4* __restrict__ extracted, int n_scan, int H,
                                                         const int* __restrict__ corner_idx, const int* __restrict__ corner_cnt,
                                                         const float4* __restrict__ surf_ds, const int* __restrict__ surf_ds_cnt,
                                                         float4* __restrict__ corner_out, int* __restrict__ corner_idx_out, float4* __restrict__ surf_out,
                                                         int* __restrict__ totals) {
    const int ring = blockIdx.x;
    __shared__ int s_cb, s_sb;
    if (threadIdx.x == 0) {
        int cb = 0, sb = 0;
        for (int r = 0; r < ring; r++) { cb += corner_cnt[r]; sb += surf_ds_cnt[r]; }
        s_cb = cb; s_sb = sb;
        if (ring == n_scan - 1) { totals[0] = cb + corner_cnt[ring]; totals[1] = sb + surf_ds_cnt[ring]; }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < corner_cnt[ring]; t += blockDim.x) {
        const int ind = corner_idx[ring * 120 + t];
        corner_out[s_cb + t] = extracted[ind];
        corner_idx_out[s_cb + t] = ind;
    }
    for (int t = threadIdx.x; t < surf_ds_cnt[ring]; t += blockDim.x) surf_out[s_sb + t] = surf_ds[ring * H + t];
}

}  // namespace b2

using namespace b2;

struct b2_scan_s {
    b2_scan_params prm;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int device = b2::current_device();      // the device the handle was created on
    DevBuf raw, imu, winner, small, range_mat, full_cloud, extracted, col_ind, pt_range, rings;
    DevBuf curv, picked, label, corner_idx, surf_idx, surf_ds, corner_out, corner_idx_out, surf_out;
    DevBuf wire;                          // cloud_info message staging (device)
    int n_raw = 0, M = 0, nc = 0, ns = 0;
    bool projected = false, featured = false;
    float last_ms = 0.f;
};

extern "C" {

void b2_scan_default_params(b2_scan_params* p) {
    if (!p) return;
    p->n_scan = 16; p->horizon_scan = 1800; p->downsample_rate = 1;          // utility.h:197-199
    p->lidar_min_range = 1.0f; p->lidar_max_range = 1000.0f;                 // utility.h:200-201
    p->edge_threshold = 1.0f; p->surf_threshold = 0.1f;                      // config/params.yaml:57-58
    p->odometry_surf_leaf_size = 0.4f;                                       // config/params.yaml:63
}

int b2_scan_create(b2_scan_t* out, const b2_scan_params* params) {
    if (!out) { set_error("b2_scan_create: null out"); return B2_ERR_ARG; }
    b2_scan_s* h = new b2_scan_s();
    if (params) h->prm = *params; else b2_scan_default_params(&h->prm);
    const b2_scan_params& p = h->prm;
    if (p.n_scan < 1 || p.n_scan > 1024 || p.horizon_scan < 16 || p.horizon_scan > 8192 || p.downsample_rate < 1 || !(p.odometry_surf_leaf_size > 0.f)) {
        delete h; set_error("b2_scan_create: bad params"); return B2_ERR_ARG;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e != cudaSuccess) { set_error("b2_scan_create: %s", cudaGetErrorString(e)); delete h; return B2_ERR_CUDA; }
    *out = h;
    return B2_OK;
}

int b2_scan_destroy(b2_scan_t h) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h) return B2_ERR_ARG;
    DevBuf* bufs[] = {&h->raw, &h->imu, &h->winner, &h->small, &h->range_mat, &h->full_cloud, &h->extracted, &h->col_ind, &h->pt_range, &h->rings,
                      &h->curv, &h->picked, &h->label, &h->corner_idx, &h->surf_idx, &h->surf_ds, &h->corner_out, &h->corner_idx_out, &h->surf_out, &h->wire};
    for (DevBuf* b : bufs) b->release();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return B2_OK;
}

static ScanDev scan_dev(const b2_scan_params& p);
// cloudInfo.{startRingIndex,endRingIndex,pointColInd,pointRange}.assign(.., 0) of allocateMemory (imageProjection.cpp:116-120):
// entries past the current scan's count keep what an earlier scan left there, and start at zero
static int scan_reserve_info(b2_scan_s* h, const ScanDev& s, cudaStream_t st) {
    const size_t cells = (size_t)s.n_scan * s.H;
    B2_CHECK(h->small.reserve(256));
    if (h->col_ind.cap < cells * 4) { B2_CHECK(h->col_ind.reserve(cells * 4)); B2_CUDA(cudaMemsetAsync(h->col_ind.p, 0, cells * 4, st)); }
    if (h->pt_range.cap < cells * 4) { B2_CHECK(h->pt_range.reserve(cells * 4)); B2_CUDA(cudaMemsetAsync(h->pt_range.p, 0, cells * 4, st)); }
    if (h->rings.cap < (size_t)s.n_scan * 3 * 4 + 64) { B2_CHECK(h->rings.reserve((size_t)s.n_scan * 3 * 4 + 64)); B2_CUDA(cudaMemsetAsync(h->rings.p, 0, (size_t)s.n_scan * 3 * 4, st)); }
    B2_CHECK(h->extracted.reserve(cells * 16));
    return B2_OK;
}

static ScanDev scan_dev(const b2_scan_params& p) {
    ScanDev s;
    s.n_scan = p.n_scan; s.H = p.horizon_scan; s.downsample = p.downsample_rate;
    s.rmin = p.lidar_min_range; s.rmax = p.lidar_max_range;
    s.ang_res_x = (float)(360.0 / (double)(float)p.horizon_scan);           // static float ang_res_x = 360.0/float(Horizon_SCAN)
    s.edge_th = p.edge_threshold; s.surf_th = p.surf_threshold; s.surf_leaf = p.odometry_surf_leaf_size;
    return s;
}

int b2_scan_project(b2_scan_t h, const void* xyzirt, size_t n, const double* imu_time, const double* imu_rot_x, const double* imu_rot_y,
                    const double* imu_rot_z, int n_imu, double time_scan_cur, int deskew_flag,
                    size_t* n_extracted, float* extracted_xyzi, int32_t* point_col_ind, float* point_range,
                    int32_t* start_ring_index, int32_t* end_ring_index, float* range_mat, float* full_cloud) {
    B2_NVTX("b2_scan_project");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || (n && !xyzirt) || !n_extracted || n > 0x7ffffff0ull || (n_imu > 0 && (!imu_time || !imu_rot_x || !imu_rot_y || !imu_rot_z))) {
        set_error("b2_scan_project: bad argument"); return B2_ERR_ARG;
    }
    const ScanDev s = scan_dev(h->prm);
    const int cells = s.n_scan * s.H;
    cudaStream_t st = h->stream;
    B2_CHECK(h->raw.reserve(std::max<size_t>(n, 1) * 32));
    B2_CHECK(h->winner.reserve((size_t)cells * 4));
    B2_CHECK(h->small.reserve(256));
    B2_CHECK(h->range_mat.reserve((size_t)cells * 4));
    B2_CHECK(h->full_cloud.reserve((size_t)cells * 16));
    B2_CHECK(h->extracted.reserve((size_t)cells * 16));
    B2_CHECK(scan_reserve_info(h, s, st));
    const int ni = std::max(n_imu, 1);
    B2_CHECK(h->imu.reserve((size_t)ni * 4 * 8));
    double* d_imu = h->imu.as<double>();
    int* first_idx = h->small.as<int>();
    int* total = first_idx + 1;
    float* start_inv = reinterpret_cast<float*>(first_idx + 4);
    int* ring_count = h->rings.as<int>();
    int* d_start = ring_count + s.n_scan;
    int* d_end = d_start + s.n_scan;
    if (n) B2_CUDA(cudaMemcpyAsync(h->raw.p, xyzirt, n * 32, cudaMemcpyHostToDevice, st));
    if (n_imu > 0) {
        B2_CUDA(cudaMemcpyAsync(d_imu, imu_time, (size_t)n_imu * 8, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(d_imu + ni, imu_rot_x, (size_t)n_imu * 8, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(d_imu + 2 * ni, imu_rot_y, (size_t)n_imu * 8, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(d_imu + 3 * ni, imu_rot_z, (size_t)n_imu * 8, cudaMemcpyHostToDevice, st));
    }
    B2_CUDA(cudaEventRecord(h->ev0, st));                        // b2_scan_last_gpu_ms: kernels only, the raw sweep already in HBM
    const int cur = n_imu - 1;                                   // imuPointerCur after the decrement of :355
    const int deskew = (deskew_flag != -1 && cur > 0) ? 1 : 0;  // deskewFlag == -1 || imuAvailable == false -> untouched point
    k_scan_init<<<(cells + 255) / 256, 256, 0, st>>>(h->winner.as<int>(), cells, first_idx); count_launch();
    if (n) { k_scan_project<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->raw.as<unsigned char>(), (int)n, s, h->winner.as<int>(), first_idx); count_launch(); }
    if (deskew) {
        k_scan_start_inverse<<<1, 32, 0, st>>>(h->raw.as<unsigned char>(), first_idx, d_imu, d_imu + ni, d_imu + 2 * ni, d_imu + 3 * ni, cur, time_scan_cur, start_inv);
        count_launch();
    }
    k_scan_cells<<<(cells + 255) / 256, 256, 0, st>>>(h->raw.as<unsigned char>(), h->winner.as<int>(), cells, d_imu, d_imu + ni, d_imu + 2 * ni, d_imu + 3 * ni,
                                                      cur, time_scan_cur, deskew, start_inv, h->range_mat.as<float>(), h->full_cloud.as<float4>()); count_launch();
    k_scan_ring_count<<<s.n_scan, 256, 0, st>>>(h->winner.as<int>(), s.H, ring_count); count_launch();
    k_scan_ring_compact<<<s.n_scan, 256, 0, st>>>(h->winner.as<int>(), h->range_mat.as<float>(), h->full_cloud.as<float4>(), s.H, s.n_scan, ring_count,
                                                  h->extracted.as<float4>(), h->col_ind.as<int>(), h->pt_range.as<float>(), d_start, d_end, total); count_launch();
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaEventRecord(h->ev1, st));
    int M = 0;
    B2_CUDA(cudaMemcpyAsync(&M, total, 4, cudaMemcpyDeviceToHost, st));
    if (start_ring_index) B2_CUDA(cudaMemcpyAsync(start_ring_index, d_start, (size_t)s.n_scan * 4, cudaMemcpyDeviceToHost, st));
    if (end_ring_index) B2_CUDA(cudaMemcpyAsync(end_ring_index, d_end, (size_t)s.n_scan * 4, cudaMemcpyDeviceToHost, st));
    if (range_mat) B2_CUDA(cudaMemcpyAsync(range_mat, h->range_mat.p, (size_t)cells * 4, cudaMemcpyDeviceToHost, st));
    if (full_cloud) B2_CUDA(cudaMemcpyAsync(full_cloud, h->full_cloud.p, (size_t)cells * 16, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    if (M) {
        if (extracted_xyzi) B2_CUDA(cudaMemcpyAsync(extracted_xyzi, h->extracted.p, (size_t)M * 16, cudaMemcpyDeviceToHost, st));
        if (point_col_ind) B2_CUDA(cudaMemcpyAsync(point_col_ind, h->col_ind.p, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
        if (point_range) B2_CUDA(cudaMemcpyAsync(point_range, h->pt_range.p, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
    }
    B2_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    h->n_raw = (int)n; h->M = M; h->projected = true; h->featured = false;
    *n_extracted = (size_t)M;
    return B2_OK;
}

int b2_scan_extract_features(b2_scan_t h, size_t* n_corner, float* corner_xyzi, int32_t* corner_index, size_t* n_surf, float* surf_xyzi,
                             float* curvature, int32_t* picked_after_mask, int32_t* label) {
    B2_NVTX("b2_scan_extract_features");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !n_corner || !n_surf) { set_error("b2_scan_extract_features: bad argument"); return B2_ERR_ARG; }
    if (!h->projected) { set_error("b2_scan_extract_features: call b2_scan_project first"); return B2_ERR_STATE; }
    const ScanDev s = scan_dev(h->prm);
    const int cells = s.n_scan * s.H;
    const int M = h->M;
    cudaStream_t st = h->stream;
    *n_corner = 0; *n_surf = 0;
    B2_CHECK(h->curv.reserve((size_t)cells * 4)); B2_CHECK(h->picked.reserve((size_t)cells * 4 * 2)); B2_CHECK(h->label.reserve((size_t)cells * 4));
    B2_CHECK(h->corner_idx.reserve((size_t)s.n_scan * 120 * 4 + (size_t)s.n_scan * 3 * 4 + 64));
    B2_CHECK(h->surf_idx.reserve((size_t)cells * 4)); B2_CHECK(h->surf_ds.reserve((size_t)cells * 16));
    B2_CHECK(h->corner_out.reserve((size_t)s.n_scan * 120 * 16)); B2_CHECK(h->corner_idx_out.reserve((size_t)s.n_scan * 120 * 4));
    B2_CHECK(h->surf_out.reserve((size_t)cells * 16));
    int* first_idx = h->small.as<int>();
    int* total = first_idx + 1;
    int* totals_out = first_idx + 32;
    int* ring_count = h->rings.as<int>();
    int* d_start = ring_count + s.n_scan;
    int* d_end = d_start + s.n_scan;
    int* corner_cnt = h->corner_idx.as<int>() + (size_t)s.n_scan * 120;
    int* surf_cnt = corner_cnt + s.n_scan;
    int* surf_ds_cnt = surf_cnt + s.n_scan;
    int* picked = h->picked.as<int>();
    int* picked_mask_copy = picked + cells;
    B2_CUDA(cudaEventRecord(h->ev0, st));
    if (M > 0) {
        k_scan_curvature<<<(M + 255) / 256, 256, 0, st>>>(h->pt_range.as<float>(), total, h->curv.as<float>(), picked, h->label.as<int>()); count_launch();
        if (M > 11) { k_scan_masks<<<(M - 11 + 255) / 256, 256, 0, st>>>(h->pt_range.as<float>(), h->col_ind.as<int>(), total, picked); count_launch(); }
        if (picked_after_mask) B2_CUDA(cudaMemcpyAsync(picked_mask_copy, picked, (size_t)M * 4, cudaMemcpyDeviceToDevice, st));
    }
    int P = 1; while (P < s.H) P <<= 1;
    const size_t smem = (size_t)P * (8 + 16) + 4 * (size_t)(s.H + 16) * 4;
    static bool attr_set = false;
    if (!attr_set) { B2_CUDA(cudaFuncSetAttribute(k_scan_features, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr_set = true; }
    if (smem > 200 * 1024) { set_error("b2_scan_extract_features: Horizon_SCAN too large for the per-ring kernel"); return B2_ERR_TOO_LARGE; }
    k_scan_features<<<s.n_scan, 256, smem, st>>>(h->extracted.as<float4>(), h->col_ind.as<int>(), total, d_start, d_end, s, P, h->curv.as<float>(), picked,
                                                 h->label.as<int>(), h->corner_idx.as<int>(), corner_cnt, h->surf_idx.as<int>(), surf_cnt,
                                                 h->surf_ds.as<float4>(), surf_ds_cnt); count_launch();
    k_scan_gather_out<<<s.n_scan, 256, 0, st>>>(h->extracted.as<float4>(), s.n_scan, s.H, h->corner_idx.as<int>(), corner_cnt, h->surf_ds.as<float4>(), surf_ds_cnt,
                                                h->corner_out.as<float4>(), h->corner_idx_out.as<int>(), h->surf_out.as<float4>(), totals_out); count_launch();
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaEventRecord(h->ev1, st));
    int tot[2] = {0, 0};
    B2_CUDA(cudaMemcpyAsync(tot, totals_out, 8, cudaMemcpyDeviceToHost, st));
    if (M > 0) {
        if (curvature) B2_CUDA(cudaMemcpyAsync(curvature, h->curv.p, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
        if (picked_after_mask) B2_CUDA(cudaMemcpyAsync(picked_after_mask, picked_mask_copy, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
        if (label) B2_CUDA(cudaMemcpyAsync(label, h->label.p, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
    }
    B2_CUDA(cudaStreamSynchronize(st));
    if (tot[0]) {
        if (corner_xyzi) B2_CUDA(cudaMemcpyAsync(corner_xyzi, h->corner_out.p, (size_t)tot[0] * 16, cudaMemcpyDeviceToHost, st));
        if (corner_index) B2_CUDA(cudaMemcpyAsync(corner_index, h->corner_idx_out.p, (size_t)tot[0] * 4, cudaMemcpyDeviceToHost, st));
    }
    if (tot[1] && surf_xyzi) B2_CUDA(cudaMemcpyAsync(surf_xyzi, h->surf_out.p, (size_t)tot[1] * 16, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    B2_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    h->nc = tot[0]; h->ns = tot[1]; h->featured = true;
    *n_corner = (size_t)tot[0]; *n_surf = (size_t)tot[1];
    return B2_OK;
}

int b2_scan_last_gpu_ms(b2_scan_t h, float* ms) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !ms) return B2_ERR_ARG;
    *ms = h->last_ms;
    return B2_OK;
}

// imuDeskewInfo (imageProjection.cpp:305-362): host only. The table is at most ~60 entries per scan and every entry
// depends on the one before it (rot[k] = rot[k-1] + w[k] * dt in double), so SURVEY.md 8a row a3 keeps it on the host; the
// device consumes the finished table in b2_scan_project. Three steps instead of the reference's single loop:
// the window [first, last] of the queue, the roll/pitch/yaw of the newest message at or before the scan start, and the
// running sum over the window. Same order of additions, so the table is bit-identical to the reference's.
static void quat_to_rpy_tf(const double* q, double* roll, double* pitch, double* yaw) {
    // tf::Matrix3x3::setRotation(q) followed by getEulerYPR(yaw, pitch, roll) (what getRPY calls), in double
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double s = 2.0 / (x * x + y * y + z * z + w * w);
    const double xs = x * s, ys = y * s, zs = z * s;
    const double r00 = 1.0 - (y * ys + z * zs), r10 = x * ys + w * zs;
    const double r20 = x * zs - w * ys, r21 = y * zs + w * xs, r22 = 1.0 - (x * xs + y * ys);
    if (fabs(r20) >= 1) {                       // gimbal lock branch of getEulerYPR
        *yaw = 0;
        *roll = atan2(r21, r22);
        *pitch = r20 < 0 ? M_PI / 2.0 : -M_PI / 2.0;
        return;
    }
    *pitch = -asin(r20);
    const double c = cos(*pitch);
    *roll = atan2(r21 / c, r22 / c);
    *yaw = atan2(r10 / c, r00 / c);
}

int b2_imu_deskew_info(const double* stamp, const double* orientation_xyzw, const double* angular_velocity, int n,
                       double time_scan_cur, double time_scan_end,
                       double* imu_time, double* imu_rot_x, double* imu_rot_y, double* imu_rot_z, int capacity,
                       int* n_table, int* n_popped, int* imu_available, float rpy_init[3]) {
    B2_NVTX("b2_imu_deskew_info");
    if (n < 0 || (n > 0 && (!stamp || !angular_velocity)) || !imu_time || !imu_rot_x || !imu_rot_y || !imu_rot_z || capacity < 1 ||
        !n_table || !n_popped || !imu_available) {
        b2::set_error("b2_imu_deskew_info: bad argument");
        return B2_ERR_ARG;
    }
    *imu_available = 0; *n_table = 0;
    // messages older than 10 ms before the scan start leave the queue (:309-315)
    int first = 0;
    while (first < n && stamp[first] < time_scan_cur - 0.01) ++first;
    *n_popped = first;
    if (first == n) return B2_OK;
    // the loop of :321-354 stops at the first message later than 10 ms after the scan end; that message is still
    // looked at by the roll/pitch/yaw test of :329 before the break
    int stop = first;
    while (stop < n && !(stamp[stop] > time_scan_end + 0.01)) ++stop;
    const int seen = stop < n ? stop + 1 : n;
    if (orientation_xyzw && rpy_init) {
        int newest = -1;
        for (int i = first; i < seen; ++i)
            if (stamp[i] <= time_scan_cur) newest = i;
        if (newest >= 0) {
            double r, p, y;
            quat_to_rpy_tf(orientation_xyzw + 4 * (size_t)newest, &r, &p, &y);
            rpy_init[0] = (float)r; rpy_init[1] = (float)p; rpy_init[2] = (float)y;
        }
    }
    const int count = stop - first;
    if (count > capacity) {
        b2::set_error("b2_imu_deskew_info: %d gyro samples in the scan window, table holds %d (queueLength, imageProjection.cpp:45)", count, capacity);
        return B2_ERR_CAPACITY;
    }
    double ax = 0, ay = 0, az = 0;
    for (int k = 0; k < count; ++k) {
        const int i = first + k;
        if (k > 0) {
            const double dt = stamp[i] - imu_time[k - 1];
            ax = ax + angular_velocity[3 * (size_t)i + 0] * dt;
            ay = ay + angular_velocity[3 * (size_t)i + 1] * dt;
            az = az + angular_velocity[3 * (size_t)i + 2] * dt;
        }
        imu_time[k] = stamp[i]; imu_rot_x[k] = ax; imu_rot_y[k] = ay; imu_rot_z[k] = az;
    }
    *n_table = count;                           // imuPointerCur == count - 1 after the decrement of :356
    *imu_available = count - 1 > 0 ? 1 : 0;
    return B2_OK;
}


// ================================================================================================================
// cloud_info wire format: msg/cloud_info.msg:1-35, ROS1 serialisation — little endian, arrays and
// strings prefixed by a u32 count, sensor_msgs/PointCloud2 = Header, height, width, PointField[] {name, offset, datatype,
// count}, is_bigendian, point_step, row_step, u8[] data, is_dense.
// ================================================================================================================
}  // extern "C"

namespace b2 {

int scan_features_dev(b2_scan_s* h, const void** d_corner, size_t* n_corner, const void** d_surf, size_t* n_surf) {
    if (!h || !h->featured) { set_error("front end: b2_scan_extract_features has not run"); return B2_ERR_STATE; }
    *d_corner = h->corner_out.p; *n_corner = (size_t)h->nc; *d_surf = h->surf_out.p; *n_surf = (size_t)h->ns;
    return B2_OK;
}

struct WireSect { const void* src; unsigned long long off; unsigned n; int kind; };   // kind 0: 4-byte words, 1: float4 -> 32-byte PointXYZI
struct WireArgs { WireSect s[7]; int n_sect; unsigned char* dst; };

__device__ __forceinline__ void wire_put(unsigned char* p, uint32_t w) {
    if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) *reinterpret_cast<uint32_t*>(p) = w;
    else { p[0] = (unsigned char)w; p[1] = (unsigned char)(w >> 8); p[2] = (unsigned char)(w >> 16); p[3] = (unsigned char)(w >> 24); }
}
__device__ __forceinline__ uint32_t wire_get(const unsigned char* p) {
    if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) return *reinterpret_cast<const uint32_t*>(p);
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// One launch packs every payload section of the message at its wire offset (blockIdx.y = section).
__global__ void __launch_bounds__(256) k_wire_pack(WireArgs a) {
    const WireSect sc = a.s[blockIdx.y];
    unsigned char* base = a.dst + sc.off;
    if (sc.kind == 0) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(sc.src);
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < sc.n; i += gridDim.x * blockDim.x) wire_put(base + (size_t)i * 4, src[i]);
    } else {
        // eight threads per point, one word each: x y z 1.0f | intensity 0 0 0   (pcl::PointXYZI, 32 bytes)
        const float4* src = reinterpret_cast<const float4*>(sc.src);
        const unsigned long long words = (unsigned long long)sc.n * 8;
        for (unsigned long long t = blockIdx.x * blockDim.x + threadIdx.x; t < words; t += (unsigned long long)gridDim.x * blockDim.x) {
            const unsigned i = (unsigned)(t >> 3), w = (unsigned)(t & 7);
            const float4 p = src[i];
            const uint32_t v = w == 0 ? __float_as_uint(p.x) : w == 1 ? __float_as_uint(p.y) : w == 2 ? __float_as_uint(p.z)
                             : w == 3 ? 0x3f800000u : w == 4 ? __float_as_uint(p.w) : 0u;
            wire_put(base + t * 4, v);
        }
    }
}

struct UnwireArgs {
    const unsigned char* msg;
    unsigned long long off_start, off_end, off_col, off_range, off_data;
    unsigned n_ring, n_col, n_range, n_pts, point_step;
    int ox, oy, oz, oi;
    int *d_start, *d_end, *col_ind, *total;
    float* pt_range;
    float4* extracted;
};

// blockIdx.y: 0 ring tables + count, 1 pointColInd, 2 pointRange, 3 the deskewed cloud (pcl::fromROSMsg by field name)
__global__ void __launch_bounds__(256) k_wire_unpack(UnwireArgs a) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, step = gridDim.x * blockDim.x;
    if (blockIdx.y == 0) {
        for (unsigned i = tid; i < a.n_ring; i += step) {
            a.d_start[i] = (int)wire_get(a.msg + a.off_start + (size_t)i * 4);
            a.d_end[i] = (int)wire_get(a.msg + a.off_end + (size_t)i * 4);
        }
        if (tid == 0) *a.total = (int)a.n_pts;
    } else if (blockIdx.y == 1) {
        for (unsigned i = tid; i < a.n_col; i += step) a.col_ind[i] = (int)wire_get(a.msg + a.off_col + (size_t)i * 4);
    } else if (blockIdx.y == 2) {
        for (unsigned i = tid; i < a.n_range; i += step) a.pt_range[i] = __uint_as_float(wire_get(a.msg + a.off_range + (size_t)i * 4));
    } else {
        for (unsigned i = tid; i < a.n_pts; i += step) {
            const unsigned char* r = a.msg + a.off_data + (size_t)i * a.point_step;
            float4 p;
            p.x = a.ox >= 0 ? __uint_as_float(wire_get(r + a.ox)) : 0.f;
            p.y = a.oy >= 0 ? __uint_as_float(wire_get(r + a.oy)) : 0.f;
            p.z = a.oz >= 0 ? __uint_as_float(wire_get(r + a.oz)) : 0.f;
            p.w = a.oi >= 0 ? __uint_as_float(wire_get(r + a.oi)) : 0.f;
            a.extracted[i] = p;
        }
    }
}

namespace {

// Serialiser over an optional output buffer: always advances, writes only when there is room (so the same walk sizes
// the message, places the headers, and records where the device-packed payloads go).
struct Wr {
    unsigned char* p; size_t cap; size_t off = 0;
    Wr(void* out, size_t c) : p(reinterpret_cast<unsigned char*>(out)), cap(c) {}
    void raw(const void* s, size_t n) { if (p && off + n <= cap && n) memcpy(p + off, s, n); off += n; }
    void u8(uint8_t v) { raw(&v, 1); }
    void u32(uint32_t v) { raw(&v, 4); }
    void i64(int64_t v) { raw(&v, 8); }
    void f32(float v) { raw(&v, 4); }
    void str(const char* s) { const uint32_t n = s ? (uint32_t)strlen(s) : 0; u32(n); raw(s, n); }
    size_t skip(size_t n) { const size_t o = off; off += n; return o; }
};

void wr_header(Wr& w, uint32_t seq, uint32_t sec, uint32_t nsec, const char* frame) { w.u32(seq); w.u32(sec); w.u32(nsec); w.str(frame); }

// pcl::toROSMsg(pcl::PointCloud<pcl::PointXYZI>) + publishCloud's stamp / frame_id (utility.h:286-295); returns the data offset
size_t wr_cloud_xyzi(Wr& w, const b2_cloud_info_meta& m, uint32_t n) {
    wr_header(w, 0, m.stamp_sec, m.stamp_nsec, m.cloud_frame_id);
    w.u32(1); w.u32(n);                                    // height, width
    w.u32(4);
    static const struct { const char* name; uint32_t off; } F[4] = {{"x", 0}, {"y", 4}, {"z", 8}, {"intensity", 16}};
    for (const auto& f : F) { w.str(f.name); w.u32(f.off); w.u8(7 /* FLOAT32 */); w.u32(1); }
    w.u8(0);                                               // is_bigendian
    w.u32(32); w.u32(32u * n);                             // point_step, row_step
    w.u32(32u * n);
    const size_t at = w.skip((size_t)32 * n);
    w.u8(1);                                               // is_dense
    return at;
}

void wr_cloud_empty(Wr& w) {                               // default-constructed sensor_msgs::PointCloud2
    wr_header(w, 0, 0, 0, nullptr);
    w.u32(0); w.u32(0); w.u32(0); w.u8(0); w.u32(0); w.u32(0); w.u32(0); w.u8(0);
}

struct Rd {
    const unsigned char* p; size_t n; size_t off = 0; bool ok = true;
    bool need(size_t k) { if (!ok || k > n - off) { ok = false; return false; } return true; }
    uint8_t u8() { if (!need(1)) return 0; return p[off++]; }
    uint32_t u32() { uint32_t v = 0; if (need(4)) { memcpy(&v, p + off, 4); off += 4; } return v; }
    int64_t i64() { int64_t v = 0; if (need(8)) { memcpy(&v, p + off, 8); off += 8; } return v; }
    float f32() { float v = 0; if (need(4)) { memcpy(&v, p + off, 4); off += 4; } return v; }
    const unsigned char* bytes(size_t k) { if (!need(k)) return nullptr; const unsigned char* r = p + off; off += k; return r; }
    const unsigned char* arr(uint32_t* count, size_t elem) { *count = u32(); if (!ok) return nullptr; if ((uint64_t)*count * elem > n - off) { ok = false; return nullptr; } return bytes((size_t)*count * elem); }
};

void rd_cloud(Rd& r, b2_cloud2_view* c) {
    memset(c, 0, sizeof(*c));
    c->off_x = c->off_y = c->off_z = c->off_intensity = -1;
    r.u32(); r.u32(); r.u32();
    uint32_t fl = 0; r.arr(&fl, 1);
    c->height = r.u32(); c->width = r.u32();
    c->n_fields = r.u32();
    for (uint32_t f = 0; r.ok && f < c->n_fields; f++) {
        uint32_t nl = 0; const unsigned char* name = r.arr(&nl, 1);
        const uint32_t off = r.u32(); const uint8_t dt = r.u8(); r.u32();
        if (!r.ok) break;
        auto is = [&](const char* s) { return nl == strlen(s) && memcmp(name, s, nl) == 0; };
        if (dt == 7) {
            if (is("x")) c->off_x = (int32_t)off; else if (is("y")) c->off_y = (int32_t)off;
            else if (is("z")) c->off_z = (int32_t)off; else if (is("intensity")) c->off_intensity = (int32_t)off;
        }
    }
    c->is_bigendian = r.u8(); c->point_step = r.u32(); c->row_step = r.u32();
    uint32_t nd = 0; c->data = r.arr(&nd, 1);
    c->is_dense = r.u8();
    if (r.ok && (uint64_t)c->width * c->height * c->point_step > nd) r.ok = false;
}

}  // namespace
}  // namespace b2

extern "C" {

int b2_cloud_info_parse(const void* msg, size_t n_bytes, b2_cloud_info_view* v) {
    if (!msg || !v) { set_error("b2_cloud_info_parse: null argument"); return B2_ERR_ARG; }
    memset(v, 0, sizeof(*v));
    Rd r{reinterpret_cast<const unsigned char*>(msg), n_bytes};
    v->seq = r.u32(); v->stamp_sec = r.u32(); v->stamp_nsec = r.u32();
    v->frame_id = reinterpret_cast<const char*>(r.arr(&v->frame_id_len, 1));
    v->start_ring_index = r.arr(&v->n_start_ring_index, 4);
    v->end_ring_index = r.arr(&v->n_end_ring_index, 4);
    v->point_col_ind = r.arr(&v->n_point_col_ind, 4);
    v->point_range = r.arr(&v->n_point_range, 4);
    v->imu_available = r.i64(); v->odom_available = r.i64();
    v->imu_roll_init = r.f32(); v->imu_pitch_init = r.f32(); v->imu_yaw_init = r.f32();
    v->initial_guess_x = r.f32(); v->initial_guess_y = r.f32(); v->initial_guess_z = r.f32();
    v->initial_guess_roll = r.f32(); v->initial_guess_pitch = r.f32(); v->initial_guess_yaw = r.f32();
    b2_cloud2_view* clouds[7] = {&v->cloud_deskewed, &v->cloud_corner, &v->cloud_surface, &v->key_frame_cloud, &v->key_frame_color,
                                 &v->key_frame_poses, &v->key_frame_map};
    for (b2_cloud2_view* c : clouds) rd_cloud(r, c);
    if (!r.ok) { set_error("b2_cloud_info_parse: truncated or malformed message (%zu bytes, stopped at %zu)", n_bytes, r.off); return B2_ERR_ARG; }
    if (r.off != n_bytes) { set_error("b2_cloud_info_parse: %zu trailing bytes", n_bytes - r.off); return B2_ERR_ARG; }
    return B2_OK;
}

int b2_scan_write_cloud_info(b2_scan_t h, const b2_cloud_info_meta* meta, int stage, void* out, size_t capacity, size_t* n_bytes) {
    B2_NVTX("b2_scan_write_cloud_info");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !meta || !n_bytes || (stage != 0 && stage != 1)) { set_error("b2_scan_write_cloud_info: bad argument"); return B2_ERR_ARG; }
    if (!h->projected || (stage == 1 && !h->featured)) { set_error("b2_scan_write_cloud_info: stage %d needs b2_scan_%s first", stage, stage ? "extract_features" : "project"); return B2_ERR_STATE; }
    const ScanDev s = scan_dev(h->prm);
    const uint32_t cells = (uint32_t)(s.n_scan * s.H);
    const int* ring_count = h->rings.as<int>();
    WireArgs a{}; a.n_sect = 0;
    size_t total = 0;
    // one walk with out == nullptr sizes the message and finds the payload offsets; the second one writes the small fields
    for (int pass = 0; pass < 2; pass++) {
        Wr w(pass ? out : nullptr, capacity);
        a.n_sect = 0;
        auto sect = [&](const void* src, size_t off, unsigned n, int kind) { if (n) { a.s[a.n_sect].src = src; a.s[a.n_sect].off = off; a.s[a.n_sect].n = n; a.s[a.n_sect].kind = kind; a.n_sect++; } };
        wr_header(w, meta->seq, meta->stamp_sec, meta->stamp_nsec, meta->frame_id);
        if (stage == 0) {
            w.u32((uint32_t)s.n_scan); sect(ring_count + s.n_scan, w.skip((size_t)s.n_scan * 4), (unsigned)s.n_scan, 0);
            w.u32((uint32_t)s.n_scan); sect(ring_count + 2 * s.n_scan, w.skip((size_t)s.n_scan * 4), (unsigned)s.n_scan, 0);
            w.u32(cells); sect(h->col_ind.p, w.skip((size_t)cells * 4), cells, 0);
            w.u32(cells); sect(h->pt_range.p, w.skip((size_t)cells * 4), cells, 0);
        } else {
            w.u32(0); w.u32(0); w.u32(0); w.u32(0);                    // freeCloudInfoMemory (featureExtraction.cpp:240-246)
        }
        w.i64(meta->imu_available); w.i64(meta->odom_available);
        w.f32(meta->imu_roll_init); w.f32(meta->imu_pitch_init); w.f32(meta->imu_yaw_init);
        w.f32(meta->initial_guess_x); w.f32(meta->initial_guess_y); w.f32(meta->initial_guess_z);
        w.f32(meta->initial_guess_roll); w.f32(meta->initial_guess_pitch); w.f32(meta->initial_guess_yaw);
        sect(h->extracted.p, wr_cloud_xyzi(w, *meta, (uint32_t)h->M), (unsigned)h->M, 1);
        if (stage == 1) {
            sect(h->corner_out.p, wr_cloud_xyzi(w, *meta, (uint32_t)h->nc), (unsigned)h->nc, 1);
            sect(h->surf_out.p, wr_cloud_xyzi(w, *meta, (uint32_t)h->ns), (unsigned)h->ns, 1);
        } else { wr_cloud_empty(w); wr_cloud_empty(w); }
        for (int k = 0; k < 4; k++) wr_cloud_empty(w);                 // key_frame_cloud / color / poses / map
        total = w.off;
        if (pass == 0) {
            *n_bytes = total;
            if (!out) return B2_OK;
            if (capacity < total) { set_error("b2_scan_write_cloud_info: message is %zu bytes, capacity %zu", total, capacity); return B2_ERR_CAPACITY; }
            // payloads: packed on the device at their wire offsets, one copy back; the header walk then fills the gaps
            cudaStream_t st = h->stream;
            B2_CHECK(h->wire.reserve(total + 16));
            a.dst = h->wire.as<unsigned char>();
            B2_CUDA(cudaEventRecord(h->ev0, st));
            if (a.n_sect) {
                unsigned most = 0;
                for (int k = 0; k < a.n_sect; k++) most = std::max(most, a.s[k].kind ? a.s[k].n * 8u : a.s[k].n);
                dim3 grid(std::max(1u, std::min((most + 255u) / 256u, (unsigned)device_sm_count() * 8u)), (unsigned)a.n_sect);
                k_wire_pack<<<grid, 256, 0, st>>>(a); count_launch();
                B2_CUDA(cudaGetLastError());
            }
            B2_CUDA(cudaEventRecord(h->ev1, st));
            // from the first payload byte to the last one: the gaps in between are rewritten by the second walk
            if (a.n_sect) {
                const size_t lo = a.s[0].off, hi = a.s[a.n_sect - 1].off + (size_t)a.s[a.n_sect - 1].n * (a.s[a.n_sect - 1].kind ? 32 : 4);
                B2_CUDA(cudaMemcpyAsync(reinterpret_cast<unsigned char*>(out) + lo, a.dst + lo, hi - lo, cudaMemcpyDeviceToHost, st));
            }
            B2_CUDA(cudaStreamSynchronize(st));
            B2_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
        }
    }
    return B2_OK;
}

int b2_scan_set_from_cloud_info(b2_scan_t h, const void* msg, size_t n_bytes, size_t* n_extracted) {
    B2_NVTX("b2_scan_set_from_cloud_info");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !msg) { set_error("b2_scan_set_from_cloud_info: null argument"); return B2_ERR_ARG; }
    b2_cloud_info_view v;
    B2_CHECK(b2_cloud_info_parse(msg, n_bytes, &v));
    const ScanDev s = scan_dev(h->prm);
    const size_t cells = (size_t)s.n_scan * s.H;
    const b2_cloud2_view& c = v.cloud_deskewed;
    const size_t M = (size_t)c.width * c.height;
    if (v.n_start_ring_index != (uint32_t)s.n_scan || v.n_end_ring_index != (uint32_t)s.n_scan) {
        set_error("b2_scan_set_from_cloud_info: %u / %u ring entries, N_SCAN is %d", v.n_start_ring_index, v.n_end_ring_index, s.n_scan); return B2_ERR_ARG;
    }
    if (M > cells || v.n_point_col_ind < M || v.n_point_range < M || v.n_point_col_ind > cells || v.n_point_range > cells) {
        set_error("b2_scan_set_from_cloud_info: %zu points, %u / %u per-point entries, range image has %zu cells", M, v.n_point_col_ind, v.n_point_range, cells); return B2_ERR_ARG;
    }
    if (M && (c.is_bigendian || c.off_x < 0 || c.off_y < 0 || c.off_z < 0 || (c.point_step < 12))) {
        set_error("b2_scan_set_from_cloud_info: cloud_deskewed needs little-endian FLOAT32 x, y, z fields"); return B2_ERR_ARG;
    }
    const int offs[4] = {c.off_x, c.off_y, c.off_z, c.off_intensity};
    for (int o : offs) if (M && o >= 0 && (uint32_t)o + 4 > c.point_step) { set_error("b2_scan_set_from_cloud_info: field offset %d outside point_step %u", o, c.point_step); return B2_ERR_ARG; }
    cudaStream_t st = h->stream;
    B2_CHECK(scan_reserve_info(h, s, st));
    B2_CHECK(h->wire.reserve(n_bytes + 16));
    B2_CUDA(cudaEventRecord(h->ev0, st));
    B2_CUDA(cudaMemcpyAsync(h->wire.p, msg, n_bytes, cudaMemcpyHostToDevice, st));
    const unsigned char* base = reinterpret_cast<const unsigned char*>(msg);
    auto at = [&](const void* p) { return p ? (unsigned long long)(reinterpret_cast<const unsigned char*>(p) - base) : 0ull; };
    UnwireArgs a{};
    a.msg = h->wire.as<unsigned char>();
    a.off_start = at(v.start_ring_index); a.off_end = at(v.end_ring_index); a.off_col = at(v.point_col_ind); a.off_range = at(v.point_range);
    a.off_data = at(c.data);
    a.n_ring = (unsigned)s.n_scan; a.n_col = v.n_point_col_ind; a.n_range = v.n_point_range; a.n_pts = (unsigned)M; a.point_step = c.point_step;
    a.ox = c.off_x; a.oy = c.off_y; a.oz = c.off_z; a.oi = c.off_intensity;
    int* first_idx = h->small.as<int>();
    int* ring_count = h->rings.as<int>();
    a.total = first_idx + 1; a.d_start = ring_count + s.n_scan; a.d_end = ring_count + 2 * s.n_scan;
    a.col_ind = h->col_ind.as<int>(); a.pt_range = h->pt_range.as<float>(); a.extracted = h->extracted.as<float4>();
    const unsigned most = (unsigned)std::max<size_t>(std::max<size_t>(v.n_point_col_ind, v.n_point_range), std::max<size_t>(M, (size_t)s.n_scan));
    dim3 grid(std::max(1u, std::min((most + 255u) / 256u, (unsigned)device_sm_count() * 8u)), 4);
    k_wire_unpack<<<grid, 256, 0, st>>>(a); count_launch();
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaEventRecord(h->ev1, st));
    B2_CUDA(cudaStreamSynchronize(st));
    B2_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    h->M = (int)M; h->n_raw = 0; h->projected = true; h->featured = false;
    if (n_extracted) *n_extracted = M;
    return B2_OK;
}

}  // extern "C"
