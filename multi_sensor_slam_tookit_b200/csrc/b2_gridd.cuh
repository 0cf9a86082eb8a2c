// fp64 uniform-grid index over a device-resident cloud, with exact ring-expanding nearest-neighbour search.
// Replaces the kd-trees Open3D builds inside estimate_normals() and registration_generalized_icp()
// (Calibration_Tookit/Multi_LiCa/multi_lidar_calibrator/calibration/Calibration.py:327-340) and the FLANN tree PCL's NDT
// uses for getFitnessScore (Calibration_Tookit/multi_lidar/.../multi_lidar_calibrator.cpp:66).
//
// HBM layout:  pts        P4d[n]        cell-sorted, one 32-byte sector per point: x, y, z, original index
//              cell_start u32[ncell+2]  exclusive prefix of per-cell counts, x fastest
// A row of consecutive x cells is one contiguous run of pts, so a (2r+1)^3 block costs (2r+1)^2 pairs of cell_start
// loads and then streams candidates with aligned 128-bit loads.
//
// Exactness: after every cell within Chebyshev ring r of the query's cell has been scanned, an unseen point is at
// least (r + m) * h away, m = distance (in cells) from the query to the nearest face of its own cell. The search
// stops when the k-th best squared distance is below that bound (shrunk by 1e-9 for the rounding of the cell
// assignment), when the bound passes the caller's radius, or when the ring covers the grid.
// Distances are (dx*dx + dy*dy) + dz*dz in double without FMA; ordering is (distance, original index).
#pragma once
#include "b2_common.cuh"

namespace b2 {

struct alignas(32) P4d { double x, y, z; long long idx; };

struct GridDDev {
    const P4d* pts;
    const uint32_t* cell_start;
    const uint4* cell_rec;     // optional, with pts_rel (coarse grids): per cell {begin, end, box lo, box hi}: the run of its points and their
                               // bounding box relative to the cell's corner in 1/256 cell, 8 bits per axis (lo rounded down, hi = last step touched)
    const float4* pts_rel;     // fp32 screening copy, same order as pts: xyz relative to the corner of the point's own cell
    double ox, oy, oz, h, inv_h;
    int nx, ny, nz;
    uint32_t n;
};

struct GridD {
    DevBuf pts, cell_start, work, cell_rec, pts_rel;
    GridDDev dev{};
    size_t n = 0;
    double ppc = 0.0;          // points per occupied cell of the built grid
    // d_xyz: device, n*3 doubles. h_request > 0 fixes the cell edge (raised only if the cell budget requires it);
    // otherwise the edge is chosen so that an occupied cell holds about target_ppc points.
    int build(const double* d_xyz, size_t n, double h_request, double target_ppc, cudaStream_t s);
    // cell records + fp32 screening copy of the points (for grids whose cells hold many points: nn1_scan_block_pruned skips a
    // cell on its box without touching its points, and tests the fp32 copy before the fp64 point)
    int build_cell_records(cudaStream_t s);
    void release() { pts.release(); cell_start.release(); work.release(); cell_rec.release(); pts_rel.release(); dev = GridDDev{}; n = 0; }
};

#ifdef __CUDACC__

__device__ __forceinline__ void load_p4d(const P4d* p, double& x, double& y, double& z, long long& idx) {
    const double2* q = reinterpret_cast<const double2*>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1);
    x = a.x; y = a.y; z = b.x; idx = __double_as_longlong(b.y);
}

__device__ __forceinline__ double shfl_d(unsigned mask, double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(mask, lo, src); hi = __shfl_sync(mask, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_d(unsigned mask, double v, int o) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(mask, lo, o); hi = __shfl_xor_sync(mask, hi, o);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_up_d(unsigned mask, double v, int o) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(mask, lo, o); hi = __shfl_up_sync(mask, hi, o);
    return __hiloint2double(hi, lo);
}

// Transposing warp reduction: every lane brings 32 values, lane i leaves with the sum over lanes of value i.
// 31 shuffle-adds instead of 32 x 5 (halves are exchanged, so the live set shrinks 32 -> 16 -> ... -> 1). Fixed tree order.
__device__ __forceinline__ double warp_reduce_scatter32(double (&v)[32]) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < o; k++) {
            const double keep = up ? v[k + o] : v[k];
            const double send = up ? v[k] : v[k + o];
            v[k] = keep + shfl_xor_d(full, send, o);
        }
    }
    return v[0];
}

// The same for 16 values: every lane leaves with the 32-lane sum of value (lane & 15). Half the live registers of the
// 32-value form (a kernel capped at 80 registers keeps 16 doubles in registers, 32 go to local memory).
__device__ __forceinline__ double warp_reduce_scatter16(double (&v)[16]) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < o; k++) {
            const double keep = up ? v[k + o] : v[k];
            const double send = up ? v[k] : v[k + o];
            v[k] = keep + shfl_xor_d(full, send, o);
        }
    }
    return v[0] + shfl_xor_d(full, v[0], 16);
}

// Geometry of one query against a grid: cell, and the per-ring exactness bound.
struct QueryCell {
    int cx, cy, cz;
    double m;          // distance to the nearest own-cell face, in cells (0..0.5)
    bool finite;
};
__device__ __forceinline__ QueryCell query_cell(const GridDDev& g, double qx, double qy, double qz) {
    QueryCell q;
    const double fx = (qx - g.ox) * g.inv_h, fy = (qy - g.oy) * g.inv_h, fz = (qz - g.oz) * g.inv_h;
    q.finite = isfinite(fx) && isfinite(fy) && isfinite(fz) && fabs(fx) < 1e9 && fabs(fy) < 1e9 && fabs(fz) < 1e9;
    const double flx = floor(fx), fly = floor(fy), flz = floor(fz);
    q.cx = q.finite ? (int)flx : 0; q.cy = q.finite ? (int)fly : 0; q.cz = q.finite ? (int)flz : 0;
    double m = fmin(fx - flx, 1.0 - (fx - flx));
    m = fmin(m, fmin(fy - fly, 1.0 - (fy - fly)));
    m = fmin(m, fmin(fz - flz, 1.0 - (fz - flz)));
    q.m = m;
    return q;
}
// squared lower bound on the distance of any point outside ring r (see header comment)
__device__ __forceinline__ double ring_bound2(const GridDDev& g, const QueryCell& q, int r) {
    const double b = ((double)r + q.m) * g.h;
    return b * b * (1.0 - 1e-9);
}
// rings needed to cover the whole grid from this query's cell
__device__ __forceinline__ int rings_to_cover(const GridDDev& g, const QueryCell& q) {
    int r = max(max(q.cx, g.nx - 1 - q.cx), max(max(q.cy, g.ny - 1 - q.cy), max(q.cz, g.nz - 1 - q.cz)));
    return max(r, 0);
}

// [begin, end) of the run of pts covering cells x0..x1 (clamped) of row (y, z); empty when the row is outside
__device__ __forceinline__ void row_range(const GridDDev& g, int x0, int x1, int y, int z, uint32_t& b, uint32_t& e) {
    x0 = max(x0, 0); x1 = min(x1, g.nx - 1);
    const bool ok = (x0 <= x1) && y >= 0 && y < g.ny && z >= 0 && z < g.nz;
    const size_t row = ((size_t)(ok ? z : 0) * g.ny + (ok ? y : 0)) * g.nx;
    b = ok ? __ldg(&g.cell_start[row + x0]) : 0u;
    e = ok ? __ldg(&g.cell_start[row + x1 + 1]) : 0u;
}

// Exact 1-NN for one query per thread.
//   fine   : the 3x3x3 block of the fine grid (nine contiguous runs); certified when the best squared distance is below
//            the ring-1 bound or the bound already exceeds radius2;
//   coarse : otherwise the 3x3x3 block of a second grid over the same points whose cell edge is >= the search radius,
//            which contains every point closer than the radius: one pass, no shell expansion.
// radius2: candidates at or beyond it are ignored (GICP max_correspondence_distance^2). Ties fall to the smaller index.
// Result: d2 (INFINITY = none), idx (original index), x/y/z the neighbour's coordinates, pos its position in the FINE
// grid's order (through `fine_pos_of`, original index -> fine position, when the coarse pass found it).
struct NN1 { double d2; long long idx; double x, y, z; uint32_t pos; };
#ifdef B2_NN1_STATS
// developer statistics (variant build only): per-thread counters the kernel adds up
struct NN1Stats { unsigned long long fine_cycles, coarse_cycles, coarse_queries, coarse_cands, coarse_exact, coarse_cells, unseeded; };
#define B2_STAT(...) __VA_ARGS__
#else
#define B2_STAT(...)
#endif

__device__ __forceinline__ void nn1_consider(NN1& best, double radius2, double qx, double qy, double qz, double x, double y, double z,
                                             long long idx, uint32_t p) {
    const double ddx = qx - x, ddy = qy - y, ddz = qz - z;
    const double d = ddx * ddx + ddy * ddy + ddz * ddz;
    if (d < radius2 && (d < best.d2 || (d == best.d2 && idx < best.idx))) { best.d2 = d; best.idx = idx; best.x = x; best.y = y; best.z = z; best.pos = p; }
}

// candidates [b, e): four loads are issued before the first is consumed (the loop is bound by load latency otherwise)
__device__ __forceinline__ void nn1_scan_run(const GridDDev& g, uint32_t b, uint32_t e, double qx, double qy, double qz, double radius2, NN1& best) {
    for (uint32_t p = b; p < e; p += 4) {
        double x[4], y[4], z[4]; long long id[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t pk = min(p + k, e - 1);
            load_p4d(&g.pts[pk], x[k], y[k], z[k], id[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (p + k < e) nn1_consider(best, radius2, qx, qy, qz, x[k], y[k], z[k], id[k], p + k);
    }
}

// the 3x3x3 block as nine contiguous runs, all eighteen bounds requested up front
// tab: this thread's column of an 18-row table in shared memory (row i at tab[i * TS]); a rolled loop over per-thread arrays
// would index them dynamically, which puts them in local memory (measured: half of the kernel's L2 traffic)
template <int TS>
__device__ __forceinline__ void nn1_scan_block(const GridDDev& g, const QueryCell& qc, double qx, double qy, double qz, double radius2, NN1& best, uint32_t* tab) {
#pragma unroll
    for (int i = 0; i < 9; i++) {
        uint32_t b, e;
        row_range(g, qc.cx - 1, qc.cx + 1, qc.cy + (i % 3) - 1, qc.cz + (i / 3) - 1, b, e);
        tab[i * TS] = b; tab[(9 + i) * TS] = e;
    }
#pragma unroll 1
    for (int i = 0; i < 9; i++) nn1_scan_run(g, tab[i * TS], tab[(9 + i) * TS], qx, qy, qz, radius2, best);
}

// The 3x3x3 block of a grid with cell records (the coarse pass), for a query the fine block could not certify — in the first
// evaluations of a registration that is almost every query, and most of them have NO target point within the radius, so the
// search has to prove the ball empty. That is 27 cells x (record, maybe candidates): a chain of dependent loads when walked
// cell by cell (measured: 120 k cycles per query, the launch latency-bound). Here instead
//   1. the 27 records (9 rows x 48 contiguous bytes) are prefetched into L1 — no registers held, one round trip;
//   2. sweep 1 tests every cell's box against the ball (fp32, conservative) and prefetches the points of the survivors;
//   3. sweep 2 walks the survivors nearest first, re-tests the box against the best so far and screens the fp32 copy of the
//      points; only a candidate that could win or tie loads its fp64 point and is decided exactly as before.
// Roundings: a box step is h/256 and the box is widened by one part in 1e6 of h; the fp32 coordinates (relative to the cell
// corner, below 2h) are within 2^-24 * 2h; a squared distance below h^2 moves by less than 1.3e-6 h^2 and the margin is eight
// times that, so nothing that could win is skipped: the result is the exact search's.
__device__ __forceinline__ void nn1_scan_block_pruned(const GridDDev& g, const QueryCell& qc, double qx, double qy, double qz, double radius2, NN1& best
                                                      B2_STAT(, NN1Stats* stats = nullptr)) {
    const float hf = (float)g.h;
    const float margin = 1e-5f * hf * hf;
    float limf = __double2float_ru(fmin(best.d2, radius2)) * (1.f + 1e-5f) + margin;
    // query relative to the corner of its own cell (fp64 difference, then narrowed); a neighbour's corner is +-h away
    const float q0x = (float)(qx - (g.ox + (double)qc.cx * g.h)), q0y = (float)(qy - (g.oy + (double)qc.cy * g.h)), q0z = (float)(qz - (g.oz + (double)qc.cz * g.h));
    const bool xin = qc.cx >= 1 && qc.cx + 1 < g.nx;
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const int y = qc.cy + (r % 3) - 1, z = qc.cz + (r / 3) - 1;
        if (y >= 0 && y < g.ny && z >= 0 && z < g.nz && qc.cx + 1 >= 0 && qc.cx - 1 < g.nx) {
            const uint4* row = g.cell_rec + ((size_t)z * g.ny + y) * g.nx;
            asm volatile("prefetch.global.L1 [%0];" :: "l"(row + min(max(qc.cx - 1, 0), g.nx - 1)));
            if (xin) asm volatile("prefetch.global.L1 [%0];" :: "l"(row + qc.cx + 1));
        }
    }
    const float step = hf * (1.f / 256.f), slack = 1e-6f * hf;
    // box distance of cell (ox, oy, oz) from the query, squared (fp32, never above the true distance to any of its points)
    auto box_d2 = [&](const uint4 rec, int ox, int oy, int oz) {
        const float qfx = q0x - (float)ox * hf, qfy = q0y - (float)oy * hf, qfz = q0z - (float)oz * hf;     // relative to that cell's corner
        const float lx = (float)(rec.z & 255u) * step - slack, ly = (float)((rec.z >> 8) & 255u) * step - slack, lz = (float)((rec.z >> 16) & 255u) * step - slack;
        const float ux = (float)((rec.w & 255u) + 1u) * step + slack, uy = (float)(((rec.w >> 8) & 255u) + 1u) * step + slack, uz = (float)(((rec.w >> 16) & 255u) + 1u) * step + slack;
        const float ex = fmaxf(0.f, fmaxf(lx - qfx, qfx - ux)), ey = fmaxf(0.f, fmaxf(ly - qfy, qfy - uy)), ez = fmaxf(0.f, fmaxf(lz - qfz, qfz - uz));
        return (ex * ex + ey * ey + ez * ez) * (1.f - 1e-5f);
    };
    // nearest-first rank of cell k = (oz + 1) * 9 + (oy + 1) * 3 + (ox + 1): centre, 6 faces, 12 edges, 8 corners
    constexpr unsigned char rank_of[27] = {19, 11, 20, 12, 5, 13, 21, 14, 22,   7, 3, 8, 1, 0, 2, 9, 4, 10,   23, 15, 24, 16, 6, 17, 25, 18, 26};
    // the inverse, 5 bits per entry: ranks 0-11, 12-23, 24-26
    const unsigned long long ord0 = 13ull | (12ull << 5) | (14ull << 10) | (10ull << 15) | (16ull << 20) | (4ull << 25) | (22ull << 30)
                                  | (9ull << 35) | (11ull << 40) | (15ull << 45) | (17ull << 50) | (1ull << 55);
    const unsigned long long ord1 = 3ull | (5ull << 5) | (7ull << 10) | (19ull << 15) | (21ull << 20) | (23ull << 25) | (25ull << 30)
                                  | (0ull << 35) | (2ull << 40) | (6ull << 45) | (8ull << 50) | (18ull << 55);
    const unsigned ord2 = 20u | (24u << 5) | (26u << 10);
    // sweep 1 (rolled: the kernel is at the size of the instruction cache as it is — fully unrolled this loop made the whole
    // launch 15 % slower): bit `rank` of `live` marks a non-empty cell whose box the ball reaches
    uint32_t live = 0;
#pragma unroll 1
    for (int r = 0; r < 9; r++) {
        const int oy = r % 3 - 1, oz = r / 3 - 1;
        const int y = qc.cy + oy, z = qc.cz + oz;
        if (!(y >= 0 && y < g.ny && z >= 0 && z < g.nz)) continue;
        const uint4* row = g.cell_rec + (((size_t)z * g.ny + y) * g.nx + qc.cx);
#pragma unroll 1
        for (int a = 0; a < 3; a++) {
            const int x = qc.cx + a - 1;
            if (x < 0 || x >= g.nx) continue;
            const uint4 rec = __ldg(row + (a - 1));
            if (rec.x >= rec.y || box_d2(rec, a - 1, oy, oz) > limf) continue;
            live |= 1u << rank_of[r * 3 + a];
            // its points, 128 bytes (8 points) per request, up to 2 KB: sweep 2 then reads them from L1
            const uint32_t pe = min(rec.y, rec.x + 128u);
            for (uint32_t p = rec.x; p < pe; p += 8u) asm volatile("prefetch.global.L1 [%0];" :: "l"(g.pts_rel + p));
        }
    }
    // sweep 2: the marked cells, nearest first
    while (live) {
        const int i = __ffs(live) - 1;
        live &= live - 1u;
        const int k = (int)((i < 12 ? ord0 >> (5 * i) : (i < 24 ? ord1 >> (5 * (i - 12)) : (unsigned long long)(ord2 >> (5 * (i - 24))))) & 31u);
        const int ox = k % 3 - 1, oy = (k / 3) % 3 - 1, oz = k / 9 - 1;
        const uint4 rec = __ldg(&g.cell_rec[((size_t)(qc.cz + oz) * g.ny + (qc.cy + oy)) * g.nx + (qc.cx + ox)]);
        if (box_d2(rec, ox, oy, oz) > limf) continue;
        B2_STAT(stats->coarse_cells++; stats->coarse_cands += rec.y - rec.x;)
        const float qfx = q0x - (float)ox * hf, qfy = q0y - (float)oy * hf, qfz = q0z - (float)oz * hf;
        const uint32_t b = rec.x, e = rec.y;
        for (uint32_t p = b; p < e; p += 4) {
            float4 c[4];
#pragma unroll
            for (int q = 0; q < 4; q++) c[q] = __ldg(&g.pts_rel[min(p + q, e - 1)]);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float dx = qfx - c[q].x, dy = qfy - c[q].y, dz = qfz - c[q].z;
                const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if (p + q < e && d <= limf) {
                    B2_STAT(stats->coarse_exact++;)
                    double x, y, z; long long id;
                    load_p4d(&g.pts[p + q], x, y, z, id);
                    nn1_consider(best, radius2, qx, qy, qz, x, y, z, id, p + q);
                    limf = __double2float_ru(fmin(best.d2, radius2)) * (1.f + 1e-5f) + margin;
                }
            }
        }
    }
}

#ifndef B2_NN1_STREAM
#define B2_NN1_STREAM 1
#endif
// the 3x3x3 block again, for a query that already holds a candidate (best.d2 finite): rows and end cells the ball of
// radius sqrt(best.d2) cannot reach are not read. The faces' distances are shortened by an absolute slack (1e-7 cells,
// nine orders above the rounding of the cell coordinate) and the ball is inflated by 1e-9, so the test is conservative;
// equal distances stay reachable (the test is a strict >), which keeps ties falling to the smaller index.
template <int TS>
__device__ __forceinline__ void nn1_scan_block_bounded(const GridDDev& g, const QueryCell& qc, double qx, double qy, double qz, double radius2, NN1& best, uint32_t* tab) {
    const double fx = (qx - g.ox) * g.inv_h - (double)qc.cx, fy = (qy - g.oy) * g.inv_h - (double)qc.cy, fz = (qz - g.oz) * g.inv_h - (double)qc.cz;
    const double lim = best.d2 * (1.0 + 1e-9);
    const double sl = 1e-7;
    const bool inx = qc.cx >= 0 && qc.cx < g.nx, iny = qc.cy >= 0 && qc.cy < g.ny, inz = qc.cz >= 0 && qc.cz < g.nz;
    double gxm = inx ? fmax(fx - sl, 0.0) * g.h : 0.0, gxp = inx ? fmax(1.0 - fx - sl, 0.0) * g.h : 0.0;
    double gym = iny ? fmax(fy - sl, 0.0) * g.h : 0.0, gyp = iny ? fmax(1.0 - fy - sl, 0.0) * g.h : 0.0;
    double gzm = inz ? fmax(fz - sl, 0.0) * g.h : 0.0, gzp = inz ? fmax(1.0 - fz - sl, 0.0) * g.h : 0.0;
    gxm *= gxm; gxp *= gxp; gym *= gym; gyp *= gyp; gzm *= gzm; gzp *= gzp;
    uint32_t rb[9], re[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const double rowgap = ((i % 3) == 0 ? gym : (i % 3) == 2 ? gyp : 0.0) + ((i / 3) == 0 ? gzm : (i / 3) == 2 ? gzp : 0.0);
        rb[i] = re[i] = 0u;
        if (!(rowgap > lim))
            row_range(g, (rowgap + gxm > lim) ? qc.cx : qc.cx - 1, (rowgap + gxp > lim) ? qc.cx : qc.cx + 1, qc.cy + (i % 3) - 1, qc.cz + (i / 3) - 1, rb[i], re[i]);
    }
#if B2_NN1_STREAM
    // the surviving rows as one candidate stream (lanes of a warp keep different rows of different lengths: walked row by
    // row the warp pays the longest row of every row slot)
    // (the compacted table lives in the thread's shared-memory column: nr indexes it dynamically)
    int nr = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) if (rb[i] < re[i]) { tab[nr * TS] = rb[i]; tab[(9 + nr) * TS] = re[i]; nr++; }
    int k = 0;
    uint32_t p = 0, e = 0;
    bool have = nr > 0;
    if (have) { p = tab[0]; e = tab[9 * TS]; k = 1; }
    while (have) {
        double x0, y0, z0, x1, y1, z1; long long i0, i1;
        const bool two = p + 1 < e;
        load_p4d(&g.pts[p], x0, y0, z0, i0);
        load_p4d(&g.pts[two ? p + 1 : p], x1, y1, z1, i1);
        nn1_consider(best, radius2, qx, qy, qz, x0, y0, z0, i0, p);
        if (two) nn1_consider(best, radius2, qx, qy, qz, x1, y1, z1, i1, p + 1);
        p += 2;
        if (p >= e) { if (k < nr) { p = tab[k * TS]; e = tab[(9 + k) * TS]; k++; } else have = false; }
    }
#else
#pragma unroll 1
    for (int i = 0; i < 9; i++) nn1_scan_run(g, rb[i], re[i], qx, qy, qz, radius2, best);
#endif
}

// seed: position (fine order) of a target point worth trying first — the previous iteration's correspondence — or
// 0xffffffff. It only narrows the search; the result is the exact nearest neighbour either way.
// tab: the calling thread's column of an 18-row table in shared memory (row i at tab[i * TS]).
// (Parking the queries that need the coarse pass in a per-warp queue and running it for 32 of them at a time measured no
// gain: they come in runs of consecutive source points anyway, so the warps that take the coarse pass are already full.)
template <int TS>
__device__ __forceinline__ NN1 nn1_thread(const GridDDev& fine, const GridDDev& coarse, const uint32_t* __restrict__ fine_pos_of,
                                          bool have_coarse, double qx, double qy, double qz, double radius2, uint32_t* tab, uint32_t seed = 0xffffffffu
                                          B2_STAT(, NN1Stats* stats = nullptr)) {
    B2_STAT(const long long t0 = clock64();)
    NN1 best; best.d2 = INFINITY; best.idx = 0x7fffffffffffffffLL; best.x = best.y = best.z = 0.0; best.pos = 0xffffffffu;
    const QueryCell qc = query_cell(fine, qx, qy, qz);
    if (!qc.finite) return best;
    if (seed != 0xffffffffu) {
        double x, y, z; long long id;
        load_p4d(&fine.pts[seed], x, y, z, id);
        nn1_consider(best, radius2, qx, qy, qz, x, y, z, id, seed);
    }
    B2_STAT(if (best.pos == 0xffffffffu) stats->unseeded++;)
    // (Sending a query without a usable seed straight to the coarse block — exact on its own — measured slower: from the
    // second evaluation on, many of them have a neighbour the fine block certifies for a fraction of the coarse pass.)
    if (best.pos != 0xffffffffu) nn1_scan_block_bounded<TS>(fine, qc, qx, qy, qz, radius2, best, tab);
    else nn1_scan_block<TS>(fine, qc, qx, qy, qz, radius2, best, tab);
    const double bound2 = ring_bound2(fine, qc, 1);
    B2_STAT(const long long t1 = clock64(); stats->fine_cycles += t1 - t0;)
    if (best.d2 < bound2 || bound2 >= radius2 || !have_coarse) return best;
    const QueryCell qcc = query_cell(coarse, qx, qy, qz);
    const long long before = best.idx;
    nn1_scan_block_pruned(coarse, qcc, qx, qy, qz, radius2, best B2_STAT(, stats));
    if (best.idx != before) best.pos = __ldg(&fine_pos_of[best.idx]);
    B2_STAT(stats->coarse_cycles += clock64() - t1; stats->coarse_queries++;)
    return best;
}

#endif  // __CUDACC__

}  // namespace b2
