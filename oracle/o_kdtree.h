// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing in the product path may include, link or call this.
//
// Exact k-nearest-neighbour search over xyz, standing in for pcl::KdTreeFLANN<PointType>
// (setInputCloud / nearestKSearch) as used at
//   liosam_ws/src/LIO-SAM/src/mapOptmization.cpp:1289-1290 (build), :987 and :1079 (5-NN).
// PCL/FLANN are not vendored under /root/reference; this restates the published behaviour of
// FLANN's KDTreeSingleIndex (leaf size 15, as PCL configures it) with the L2_Simple metric:
//   * squared distance = ((dx*dx) + dy*dy) + dz*dz accumulated left to right in float, no FMA;
//   * exact search (eps = 0), results ascending by distance, indices into the input cloud.
// One deliberate pin (SURVEY.md §7 "Exact kNN", north_star "ties broken by index"): equal
// distances are ordered by the smaller point index, where FLANN's order is traversal-dependent.
#pragma once
#include <vector>
#include <cstdint>
#include <cmath>
#include <algorithm>
#include <numeric>

namespace orc {

struct KdTree {
    struct Node {
        int left = -1, right = -1;   // children (internal) ; left<0 => leaf
        int lo = 0, hi = 0;          // leaf: range in perm
        int dim = 0;
        float divlow = 0.f, divhigh = 0.f;
    };
    const float* pts = nullptr;  // stride floats between points
    int stride = 4;
    int n = 0;
    std::vector<int> perm;
    std::vector<float> packed;   // xyz reordered by perm for locality (FLANN "reorder" = true)
    std::vector<Node> nodes;
    float bbmin[3], bbmax[3];
    static constexpr int LEAF = 15;

    void build(const float* p, int n_, int stride_) {
        pts = p; n = n_; stride = stride_;
        perm.resize(n); std::iota(perm.begin(), perm.end(), 0);
        nodes.clear(); nodes.reserve(n / 4 + 8);
        if (n == 0) return;
        for (int d = 0; d < 3; d++) { bbmin[d] = bbmax[d] = p[d]; }
        for (int i = 1; i < n; i++) for (int d = 0; d < 3; d++) {
            float v = p[(size_t)i * stride + d];
            bbmin[d] = std::min(bbmin[d], v); bbmax[d] = std::max(bbmax[d], v);
        }
        float bmin[3] = {bbmin[0], bbmin[1], bbmin[2]}, bmax[3] = {bbmax[0], bbmax[1], bbmax[2]};
        divide(0, n, bmin, bmax);
        packed.resize((size_t)n * 3);
        for (int i = 0; i < n; i++) for (int d = 0; d < 3; d++) packed[(size_t)i * 3 + d] = p[(size_t)perm[i] * stride + d];
    }

    int divide(int lo, int hi, float* bmin, float* bmax) {
        int id = (int)nodes.size();
        nodes.emplace_back();
        if (hi - lo <= LEAF) {
            nodes[id].lo = lo; nodes[id].hi = hi;
            for (int d = 0; d < 3; d++) { bmin[d] = bmax[d] = pts[(size_t)perm[lo] * stride + d]; }
            for (int i = lo + 1; i < hi; i++) for (int d = 0; d < 3; d++) {
                float v = pts[(size_t)perm[i] * stride + d];
                bmin[d] = std::min(bmin[d], v); bmax[d] = std::max(bmax[d], v);
            }
            return id;
        }
        // split the widest dimension of the bounding box at its middle, clamped into the data span
        int dim = 0; float span = bmax[0] - bmin[0];
        for (int d = 1; d < 3; d++) if (bmax[d] - bmin[d] > span) { span = bmax[d] - bmin[d]; dim = d; }
        float vmin = pts[(size_t)perm[lo] * stride + dim], vmax = vmin;
        for (int i = lo + 1; i < hi; i++) { float v = pts[(size_t)perm[i] * stride + dim]; vmin = std::min(vmin, v); vmax = std::max(vmax, v); }
        float cut = (bmin[dim] + bmax[dim]) / 2;
        cut = std::min(std::max(cut, vmin), vmax);
        // three-way partition: < cut | == cut | > cut ; put the boundary as close to the middle as possible
        int* P = perm.data();
        int l = lo, r = hi - 1;
        for (;;) {
            while (l <= r && pts[(size_t)P[l] * stride + dim] < cut) ++l;
            while (l <= r && pts[(size_t)P[r] * stride + dim] >= cut) --r;
            if (l > r) break;
            std::swap(P[l], P[r]); ++l; --r;
        }
        int lim1 = l;
        r = hi - 1;
        for (;;) {
            while (l <= r && pts[(size_t)P[l] * stride + dim] <= cut) ++l;
            while (l <= r && pts[(size_t)P[r] * stride + dim] > cut) --r;
            if (l > r) break;
            std::swap(P[l], P[r]); ++l; --r;
        }
        int lim2 = l;
        int half = (hi - lo) / 2, mid;
        if (lim1 - lo > half) mid = lim1;
        else if (lim2 - lo < half) mid = lim2;
        else mid = lo + half;
        if (mid == lo || mid == hi) mid = lo + half;   // all equal along dim
        float lmin[3] = {bmin[0], bmin[1], bmin[2]}, lmax[3] = {bmax[0], bmax[1], bmax[2]};
        float rmin[3] = {bmin[0], bmin[1], bmin[2]}, rmax[3] = {bmax[0], bmax[1], bmax[2]};
        lmax[dim] = cut; rmin[dim] = cut;
        int L = divide(lo, mid, lmin, lmax);
        int R = divide(mid, hi, rmin, rmax);
        nodes[id].left = L; nodes[id].right = R; nodes[id].dim = dim;
        nodes[id].divlow = lmax[dim]; nodes[id].divhigh = rmin[dim];
        for (int d = 0; d < 3; d++) { bmin[d] = std::min(lmin[d], rmin[d]); bmax[d] = std::max(lmax[d], rmax[d]); }
        return id;
    }

    struct Result {
        int k, count = 0;
        int* idx; float* d2;
        float worst() const { return count < k ? INFINITY : d2[k - 1]; }
        int worst_idx() const { return count < k ? INT32_MAX : idx[k - 1]; }
        void add(float d, int i) {
            if (count == k) {
                if (d > d2[k - 1] || (d == d2[k - 1] && i > idx[k - 1])) return;
            }
            int pos = count < k ? count : k - 1;
            while (pos > 0 && (d2[pos - 1] > d || (d2[pos - 1] == d && idx[pos - 1] > i))) {
                d2[pos] = d2[pos - 1]; idx[pos] = idx[pos - 1]; --pos;
            }
            d2[pos] = d; idx[pos] = i;
            if (count < k) ++count;
        }
    };

    void search_level(Result& res, const float* q, int node, float mindist, float* dists) const {
        const Node& nd = nodes[node];
        if (nd.left < 0) {
            for (int i = nd.lo; i < nd.hi; i++) {
                const float* p = &packed[(size_t)i * 3];
                float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
                float d = 0.f; d += dx * dx; d += dy * dy; d += dz * dz;
                res.add(d, perm[i]);
            }
            return;
        }
        int dim = nd.dim;
        float val = q[dim];
        float diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
        int best, other; float cut;
        if (diff1 + diff2 < 0) { best = nd.left; other = nd.right; cut = diff2 * diff2; }
        else { best = nd.right; other = nd.left; cut = diff1 * diff1; }
        search_level(res, q, best, mindist, dists);
        float dst = dists[dim];
        mindist = mindist + cut - dst;
        dists[dim] = cut;
        // "<=" keeps exact ties reachable so the (d2, index) order is well defined; the bound is a
        // float estimate, so it is relaxed by a few ulp before it is allowed to prune
        if (mindist * 0.9999f <= res.worst()) search_level(res, q, other, mindist, dists);
        dists[dim] = dst;
    }

    // returns number found (min(k, n)); ascending (d2, index)
    int knn(const float* q, int k, int* idx, float* d2) const {
        Result res{k, 0, idx, d2};
        if (n == 0) return 0;
        float dists[3] = {0, 0, 0};
        float distsq = 0.f;
        for (int d = 0; d < 3; d++) {
            if (q[d] < bbmin[d]) { dists[d] = (q[d] - bbmin[d]) * (q[d] - bbmin[d]); distsq += dists[d]; }
            if (q[d] > bbmax[d]) { dists[d] = (q[d] - bbmax[d]) * (q[d] - bbmax[d]); distsq += dists[d]; }
        }
        search_level(res, q, 0, distsq, dists);
        return res.count;
    }
};

}  // namespace orc
