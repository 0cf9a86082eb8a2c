// Uniform-grid exact k-nearest-neighbour search, device side.
// Replaces pcl::KdTreeFLANN::nearestKSearch (liosam_ws/src/LIO-SAM/src/mapOptmization.cpp:987,1079).
// One query is served by a group of LPF consecutive lanes: the 27 cells around the query are 9 x-rows that are
// contiguous runs of the cell-sorted float4 array, each lane strides a row with 128-bit loads and keeps a
// register top-K, then the group merges with shuffles. Ordering key = (d2 bits << 32) | original index, so ties
// fall to the smaller index. d2 = ((dx*dx)+dy*dy)+dz*dz in float, no FMA (FLANN L2_Simple).
#pragma once
#include "b2_common.cuh"

namespace b2 {

constexpr unsigned long long KNN_EMPTY = 0xffffffffffffffffull;

template <int K>
struct TopK {
    unsigned long long key[K];
    uint32_t pos[K];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < K; i++) { key[i] = KNN_EMPTY; pos[i] = 0; }
    }
    __device__ __forceinline__ void push(unsigned long long k, uint32_t p) {
        if (k < key[K - 1]) {
            key[K - 1] = k; pos[K - 1] = p;
#pragma unroll
            for (int j = K - 1; j > 0; j--) {
                if (key[j] < key[j - 1]) {
                    unsigned long long tk = key[j]; key[j] = key[j - 1]; key[j - 1] = tk;
                    uint32_t tp = pos[j]; pos[j] = pos[j - 1]; pos[j - 1] = tp;
                }
            }
        }
    }
    __device__ __forceinline__ void pop() {
#pragma unroll
        for (int j = 0; j < K - 1; j++) { key[j] = key[j + 1]; pos[j] = pos[j + 1]; }
        key[K - 1] = KNN_EMPTY;
    }
};

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned mask, unsigned long long v, int o) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_xor_sync(mask, lo, o); hi = __shfl_xor_sync(mask, hi, o);
    return ((unsigned long long)hi << 32) | lo;
}

// All LPF lanes of a group call this with the same query. On return every lane holds the K winners:
// out_key[r] (KNN_EMPTY when fewer than r+1 candidates exist in the 27 cells) and out_pos[r] (position in g.pts).
// bound2: the caller knows K map points whose squared distance to the query is <= bound2 (INFINITY: knows nothing). Nothing
// beyond it can be one of the K nearest, so such candidates are not inserted and cells the ball cannot reach are not read.
// The cell test is conservative: a point binned into cell i by floorf((p - o) * inv_h) lies within `slack` of the cell's
// faces (a few float ulps of the cell coordinate; slack is ~80x that), and the faces' distances are shortened by it.
template <int K, int LPF>
__device__ __forceinline__ void knn_group(const GridDev& g, float qx, float qy, float qz, bool active, float bound2,
                                          unsigned long long (&out_key)[K], uint32_t (&out_pos)[K], uint32_t* visited = nullptr) {
    uint32_t n_visited = 0;                       // candidates this lane loaded (statistics for the roofline numerator; dead code when unused)
    const int lane = threadIdx.x & 31;
    const int sub = lane & (LPF - 1);
    TopK<K> top; top.clear();
    if (active) {
        float fx = (qx - g.ox) * g.inv_h, fy = (qy - g.oy) * g.inv_h, fz = (qz - g.oz) * g.inv_h;
        // the comparisons are false for NaN, so a non-finite query scans nothing
        if (fx > -2.f && fx < (float)(g.nx + 1) && fy > -2.f && fy < (float)(g.ny + 1) && fz > -2.f && fz < (float)(g.nz + 1)) {
            const int cx = (int)floorf(fx), cy = (int)floorf(fy), cz = (int)floorf(fz);
            // accepted candidates: d < max_d2 (the reference gates on sqDis[K-1] < max_dist^2, so a neighbour at or beyond it
            // can never be one of K accepted neighbours) and d <= bound2
            // (wide groups, LPF >= 8, are the latency-bound shapes: they take no bound and skip the face arithmetic)
            constexpr bool PRUNE = LPF < 8;
            const float lim = PRUNE ? fminf(bound2, __uint_as_float(__float_as_uint(g.max_d2) - 1u)) : __uint_as_float(__float_as_uint(g.max_d2) - 1u);
            const float h = g.h;
            const float slack = 1e-5f * (fabsf(fx) + fabsf(fy) + fabsf(fz) + 4.f) * h;
            // squared distance from the query to the neighbouring slab on each side (0 when the query's own cell is not a
            // regular cell of the grid, i.e. no pruning there)
            const bool inx = PRUNE && cx >= 0 && cx < g.nx, iny = PRUNE && cy >= 0 && cy < g.ny, inz = PRUNE && cz >= 0 && cz < g.nz;
            const float tx = fx - (float)cx, ty = fy - (float)cy, tz = fz - (float)cz;
            float gxm = inx ? fmaxf(tx * h - slack, 0.f) : 0.f, gxp = inx ? fmaxf((1.f - tx) * h - slack, 0.f) : 0.f;
            float gym = iny ? fmaxf(ty * h - slack, 0.f) : 0.f, gyp = iny ? fmaxf((1.f - ty) * h - slack, 0.f) : 0.f;
            float gzm = inz ? fmaxf(tz * h - slack, 0.f) : 0.f, gzp = inz ? fmaxf((1.f - tz) * h - slack, 0.f) : 0.f;
            gxm *= gxm; gxp *= gxp; gym *= gym; gyp *= gyp; gzm *= gzm; gzp *= gzp;
            float lim_now = lim;
            auto row_range = [&](int r, uint32_t& s0, uint32_t& e0) {
                const int y = cy + (r % 3) - 1, z = cz + (r / 3) - 1;
                const float rowgap = ((r % 3) == 0 ? gym : (r % 3) == 2 ? gyp : 0.f) + ((r / 3) == 0 ? gzm : (r / 3) == 2 ? gzp : 0.f);
                const int x0 = max(rowgap + gxm <= lim_now ? cx - 1 : cx, 0), x1 = min(rowgap + gxp <= lim_now ? cx + 1 : cx, g.nx - 1);
                const bool ok = (x0 <= x1) && y >= 0 && y < g.ny && z >= 0 && z < g.nz && rowgap <= lim_now;
                const size_t row = ((size_t)(ok ? z : 0) * g.ny + (ok ? y : 0)) * g.nx;
                s0 = ok ? __ldg(&g.cell_start[row + x0]) : 0u;
                e0 = ok ? __ldg(&g.cell_start[row + x1 + 1]) : 0u;
            };
            auto scan_row = [&](uint32_t p, uint32_t e, float4 c) {
                for (;;) {
                    n_visited++;
                    const float dx = qx - c.x, dy = qy - c.y, dz = qz - c.z;
                    float d = dx * dx;
                    d = d + dy * dy;
                    d = d + dz * dz;
                    if (d <= lim_now) top.push(((unsigned long long)__float_as_uint(d) << 32) | (uint32_t)__float_as_int(c.w), p);
                    p += LPF;
                    if (p >= e) break;
                    c = ldg4(&g.pts[p]);
                }
            };
            // Lanes own the positions congruent to `sub` modulo LPF (absolute positions: the lanes of a group may prune a row to
            // different ranges, the partition must not depend on where a lane's range starts).
            // The query's own row first: once this lane holds K candidates, the K-th one bounds everything that follows
            // (K points exist within that distance), which usually shuts most of the other eight rows
            // (not for wide groups: a lane of 8 or 16 rarely collects K candidates from one row, and the extra dependent
            // load stage costs the latency-bound single-scan shape 8 %)
            constexpr bool CENTER_FIRST = PRUNE;
            if constexpr (CENTER_FIRST) {
                uint32_t s4, e4;
                row_range(4, s4, e4);
                const uint32_t p = s4 + ((sub - s4) & (LPF - 1));
                if (p < e4) scan_row(p, e4, ldg4(&g.pts[p]));
                if (top.key[K - 1] != KNN_EMPTY) lim_now = fminf(lim_now, __uint_as_float((uint32_t)(top.key[K - 1] >> 32)));
            }
            if constexpr (PRUNE) {
                // The eight outer rows as ONE candidate stream per lane. Walking them row by row makes the warp pay, for every
                // row, the longest row among its 32 lanes (measured: 14-15 of 32 threads active per instruction); as a stream
                // the warp runs as many steps as its busiest lane has candidates. The non-empty ranges are compacted into a
                // small per-lane table. (Requesting the next candidate before testing the current one measured no gain: 32 warps per
                // SM hide the load, and the extra float4 spills at 64 registers.)
                uint32_t lb[8], le[8];
                int nr = 0;
#pragma unroll
                for (int r = 0; r < 9; r++) {
                    if (r == 4) continue;
                    uint32_t s0, e0;
                    row_range(r, s0, e0);
                    const uint32_t p0 = s0 + ((sub - s0) & (LPF - 1));
                    if (p0 < e0) { lb[nr] = p0; le[nr] = e0; nr++; }
                }
                int k = 0;
                uint32_t p = 0, e = 0;
                bool have = nr > 0;
                if (have) { p = lb[0]; e = le[0]; k = 1; }
                while (have) {
                    const float4 c = ldg4(&g.pts[p]);
                    n_visited++;
                    const float dx = qx - c.x, dy = qy - c.y, dz = qz - c.z;
                    float d = dx * dx;
                    d = d + dy * dy;
                    d = d + dz * dz;
                    if (d <= lim_now) top.push(((unsigned long long)__float_as_uint(d) << 32) | (uint32_t)__float_as_int(c.w), p);
                    p += LPF;
                    if (p >= e) { if (k < nr) { p = lb[k]; e = le[k]; k++; } else have = false; }
                }
            } else {
                uint32_t rs[9], re[9];
#pragma unroll
                for (int r = 0; r < 9; r++) row_range(r, rs[r], re[r]);
                // first chunk of all nine rows is requested before any of it is consumed: independent 128-bit loads
                // in flight per lane instead of a load->compare->load chain (rows rarely exceed LPF points)
                float4 first[9];
#pragma unroll
                for (int r = 0; r < 9; r++) {
                    const uint32_t p = rs[r] + ((sub - rs[r]) & (LPF - 1));
                    first[r] = (p < re[r]) ? ldg4(&g.pts[p]) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int r = 0; r < 9; r++) {
                    const uint32_t p = rs[r] + ((sub - rs[r]) & (LPF - 1));
                    if (p < re[r]) scan_row(p, re[r], first[r]);
                }
            }
        }
    }
    if (visited) *visited += n_visited;
    // merge the LPF sorted lists: K rounds of group-min. The scan above diverges per lane (row lengths differ); without an
    // explicit reconvergence point the full-mask shuffles below run on the divergent slow path (measured 5x slower).
    __syncwarp();
    const unsigned full = 0xffffffffu;
    const unsigned gmask = (LPF == 32) ? full : (((1u << LPF) - 1u) << (lane & ~(LPF - 1)));
#pragma unroll
    for (int r = 0; r < K; r++) {
        unsigned long long m = top.key[0];
#pragma unroll
        for (int o = LPF >> 1; o > 0; o >>= 1) { unsigned long long t = shfl_xor_u64(full, m, o); m = t < m ? t : m; }
        const bool own = (top.key[0] == m) && (m != KNN_EMPTY);
        const unsigned owners = __ballot_sync(full, own) & gmask;
        const int src = owners ? (__ffs(owners) - 1) : lane;
        out_key[r] = m;
        out_pos[r] = __shfl_sync(full, top.pos[0], src);
        if (own) top.pop();
        __syncwarp();
    }
}

}  // namespace b2
