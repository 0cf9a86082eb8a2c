// 32-ary bounding-volume hierarchy over Morton-sorted fp64 points, and the warp-wide exact k-nearest-neighbour search on it.
// Used by estimate_normals (30-NN, b2_cloud.cu) and by the NDT fitness score (1-NN over the full target, b2_ndt.cu).
#pragma once
#include "b2_gridd.cuh"

namespace b2 {

// Lidar clouds vary in density by orders of magnitude, so neighbourhoods are not found with a uniform grid here: a
// leaf is 32 consecutive points of the Morton order, a node of level l holds 32 nodes of level l-1, and a warp tests
// the 32 children of a node in one step (one lane per child box). Children are visited nearest first and skipped
// once their box is farther than the current k-th neighbour (strictly: equal distances are still visited, a tied
// point with a smaller index may hide there). Box distances use the same (dx*dx + dy*dy) + dz*dz association as
// point distances, so rounding is monotone and the pruning is exact.
constexpr int BVH_MAXL = 7;
struct BvhDev {
    const P4d* pts;
    const double* box[BVH_MAXL];     // level l: count[l] boxes of 6 doubles (min xyz, max xyz)
    uint32_t count[BVH_MAXL];
    int levels;
    uint32_t n;                      // finite points
};


// host side (b2_cloud.cu): Morton sort + bottom-up boxes. `work` is scratch; non-finite points are left out.
struct BvhIndex {
    DevBuf pts, boxes;
    BvhDev dev{};
    int build(const double* d_xyz, size_t n, DevBuf& work, cudaStream_t s);
    void release() { pts.release(); boxes.release(); dev = BvhDev{}; }
};

#ifdef __CUDACC__

// A warp keeps the 32 best (distance, index) pairs seen so far, one per lane, ascending; slot K-1 is the k-th best.
struct WarpList {
    double sd; int si; uint32_t sp;       // this lane's slot
    double kd; int ki;                    // k-th best (uniform)
};
__device__ __forceinline__ void warp_offer(WarpList& L, int K, double d, int idx, uint32_t pos, bool valid) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned mask = __ballot_sync(full, valid && (d < L.kd || (d == L.kd && idx < L.ki)));
    while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const double cd = shfl_d(full, d, src);
        const int ci = __shfl_sync(full, idx, src);
        const uint32_t cp = __shfl_sync(full, pos, src);
        if (!(cd < L.kd || (cd == L.kd && ci < L.ki))) continue;
        const bool gt = (L.sd > cd) || (L.sd == cd && L.si > ci);
        const unsigned gm = __ballot_sync(full, gt);
        const int ins = __ffs(gm) - 1;
        const double ud = shfl_up_d(full, L.sd, 1);
        const int ui = __shfl_up_sync(full, L.si, 1);
        const uint32_t up = __shfl_up_sync(full, L.sp, 1);
        if (lane > ins) { L.sd = ud; L.si = ui; L.sp = up; }
        else if (lane == ins) { L.sd = cd; L.si = ci; L.sp = cp; }
        L.kd = shfl_d(full, L.sd, K - 1);
        L.ki = __shfl_sync(full, L.si, K - 1);
    }
}

__device__ __forceinline__ double box_dist2(const double* __restrict__ b, double qx, double qy, double qz) {
    const double2 a0 = __ldg(reinterpret_cast<const double2*>(b)), a1 = __ldg(reinterpret_cast<const double2*>(b) + 1),
                  a2 = __ldg(reinterpret_cast<const double2*>(b) + 2);
    // a0 = (lo.x, lo.y), a1 = (lo.z, hi.x), a2 = (hi.y, hi.z)
    const double dx = fmax(0.0, fmax(a0.x - qx, qx - a1.y));
    const double dy = fmax(0.0, fmax(a0.y - qy, qy - a2.x));
    const double dz = fmax(0.0, fmax(a1.x - qz, qz - a2.y));
    return dx * dx + dy * dy + dz * dz;
}

__device__ __forceinline__ void bvh_scan_leaf(const BvhDev& T, WarpList& L, int K, double qx, double qy, double qz, uint32_t leaf) {
    const int lane = threadIdx.x & 31;
    const uint32_t p = leaf * 32u + lane;
    const bool valid = p < T.n;
    double d = INFINITY; int idx = 0x7fffffff;
    if (valid) {
        double x, y, z; long long id;
        load_p4d(&T.pts[p], x, y, z, id);
        const double dx = qx - x, dy = qy - y, dz = qz - z;
        d = dx * dx + dy * dy + dz * dz;
        idx = (int)id;
    }
    warp_offer(L, K, d, idx, p, valid);
}

// warp-wide exact K nearest neighbours of (qx, qy, qz); on return lane r < K holds the r-th neighbour in L
__device__ __forceinline__ void bvh_knn_warp(const BvhDev& T, WarpList& L, int K, double qx, double qy, double qz) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int top = T.levels - 1;
    if (top == 0) { bvh_scan_leaf(T, L, K, qx, qy, qz, 0); return; }
    double dch[BVH_MAXL];
    uint32_t node[BVH_MAXL];
    int lv = top;
    node[lv] = 0;
    auto expand = [&](int l, uint32_t nd) {
        const uint32_t c = nd * 32u + lane;
        dch[l] = (c < T.count[l - 1]) ? box_dist2(T.box[l - 1] + 6 * (size_t)c, qx, qy, qz) : INFINITY;
    };
    expand(lv, 0);
    for (;;) {
        double dmin = dch[lv]; int jmin = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = shfl_xor_d(full, dmin, o);
            const int oj = __shfl_xor_sync(full, jmin, o);
            if (od < dmin || (od == dmin && oj < jmin)) { dmin = od; jmin = oj; }
        }
        if (dmin == INFINITY || dmin > L.kd) {          // nothing left worth visiting below this node
            if (++lv > top) break;
            continue;
        }
        if (lane == jmin) dch[lv] = INFINITY;
        const uint32_t child = node[lv] * 32u + (uint32_t)jmin;
        if (lv == 1) bvh_scan_leaf(T, L, K, qx, qy, qz, child);
        else { lv--; node[lv] = child; expand(lv, child); }
    }
}


// warp-wide exact nearest neighbour (squared distance only): the same walk with a scalar bound, one lane per leaf point
__device__ __forceinline__ double bvh_nn1_warp(const BvhDev& T, double qx, double qy, double qz) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    double best = INFINITY;
    auto scan_leaf = [&](uint32_t leaf) {
        const uint32_t p = leaf * 32u + lane;
        double d = INFINITY;
        if (p < T.n) {
            double x, y, z; long long id;
            load_p4d(&T.pts[p], x, y, z, id);
            const double dx = qx - x, dy = qy - y, dz = qz - z;
            d = dx * dx + dy * dy + dz * dz;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d = fmin(d, shfl_xor_d(full, d, o));
        best = fmin(best, d);
    };
    const int top = T.levels - 1;
    if (top == 0) { scan_leaf(0); return best; }
    double dch[BVH_MAXL];
    uint32_t node[BVH_MAXL];
    int lv = top;
    node[lv] = 0;
    auto expand = [&](int l, uint32_t nd) {
        const uint32_t c = nd * 32u + lane;
        dch[l] = (c < T.count[l - 1]) ? box_dist2(T.box[l - 1] + 6 * (size_t)c, qx, qy, qz) : INFINITY;
    };
    expand(lv, 0);
    for (;;) {
        double dmin = dch[lv]; int jmin = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = shfl_xor_d(full, dmin, o);
            const int oj = __shfl_xor_sync(full, jmin, o);
            if (od < dmin || (od == dmin && oj < jmin)) { dmin = od; jmin = oj; }
        }
        if (dmin == INFINITY || dmin >= best) {          // distance only: a tie cannot improve the result
            if (++lv > top) break;
            continue;
        }
        if (lane == jmin) dch[lv] = INFINITY;
        const uint32_t child = node[lv] * 32u + (uint32_t)jmin;
        if (lv == 1) scan_leaf(child);
        else { lv--; node[lv] = child; expand(lv, child); }
    }
    return best;
}

#endif  // __CUDACC__

}  // namespace b2
