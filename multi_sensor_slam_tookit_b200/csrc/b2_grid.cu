// Uniform-grid index build (replaces pcl::KdTreeFLANN::setInputCloud, mapOptmization.cpp:1289-1290) and the
// batched k-NN entry points of the C ABI.
//
// HBM layout after build():  pts   float4[n]      cell-sorted (x, y, z, original index as int bits)
//                            cell_start u32[ncell+2]  exclusive prefix of per-cell counts (x fastest)
// Build = bbox -> (cell, rank-in-cell) per point with one counting atomic -> prefix sum -> scatter: a counting sort, two
// launches (bounding box; one cooperative kernel for the rest) and one 24-byte readback for a 100 k-point map.
// Cell edge h = max_dist * (1 + 2^-7): any point closer than max_dist to a query lies in the 3x3x3 block
// around the query's cell even after float rounding of (p - origin) * (1/h).
#include "b2_grid.cuh"
#include <cooperative_groups.h>
#include <cmath>
#include <vector>

namespace b2 {

constexpr size_t GRID_MAX_CELLS = (size_t)1 << 27;   // 512 MiB of cell_start at most

__global__ void k_bbox_init(uint32_t* bb) {
    if (threadIdx.x < 3) bb[threadIdx.x] = 0xffffffffu;        // min (flipped)
    else if (threadIdx.x < 6) bb[threadIdx.x] = 0u;            // max (flipped)
}

__global__ void __launch_bounds__(256) k_bbox(const unsigned char* __restrict__ raw, size_t stride, size_t n, uint32_t* __restrict__ bb) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float* p = reinterpret_cast<const float*>(raw + i * stride);
        float x = p[0], y = p[1], z = p[2];
        if (isfinite(x) && isfinite(y) && isfinite(z)) {
            mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
            mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    }
    // six atomics per CTA, not per warp (2 600 warps queueing on six words were most of this kernel's 8.7 us)
    __shared__ float s_mn[8][3], s_mx[8][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 3; d++) { s_mn[warp][d] = mn[d]; s_mx[warp][d] = mx[d]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int d = threadIdx.x;
        float a = s_mn[0][d], b = s_mx[0][d];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) { a = fminf(a, s_mn[w][d]); b = fmaxf(b, s_mx[w][d]); }
        if (a <= b) { atomicMin(&bb[d], float_flip(a)); atomicMax(&bb[3 + d], float_flip(b)); }
    }
}

struct GridGeom { float ox, oy, oz, inv_h; int nx, ny, nz; uint32_t ncell; };

__device__ __forceinline__ uint32_t cell_of_point(const GridGeom& g, float x, float y, float z) {
    if (!(isfinite(x) && isfinite(y) && isfinite(z))) return g.ncell;
    int cx = (int)floorf((x - g.ox) * g.inv_h), cy = (int)floorf((y - g.oy) * g.inv_h), cz = (int)floorf((z - g.oz) * g.inv_h);
    cx = min(max(cx, 0), g.nx - 1); cy = min(max(cy, 0), g.ny - 1); cz = min(max(cz, 0), g.nz - 1);
    return (uint32_t)(((size_t)cz * g.ny + cy) * g.nx + cx);
}

// cell of every point, and its rank inside the cell (the value the counting atomic returns): after the prefix sum of the
// counts the point's slot is cell_start[cell] + rank. The order of points inside a cell is whatever the atomics gave; every
// consumer orders candidates by (distance, original index), so results do not depend on it.
__global__ void __launch_bounds__(256) k_cell_key(const unsigned char* __restrict__ raw, size_t stride, uint32_t n, GridGeom g,
                                                  uint32_t* __restrict__ cell_of, uint32_t* __restrict__ rank_in_cell, uint32_t* __restrict__ count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    uint32_t c = cell_of_point(g, p[0], p[1], p[2]);
    cell_of[i] = c;
    rank_in_cell[i] = atomicAdd(&count[c], 1u);
}

__global__ void __launch_bounds__(256) k_cell_scatter(const unsigned char* __restrict__ raw, size_t stride, uint32_t n,
                                                      const uint32_t* __restrict__ cell_of, const uint32_t* __restrict__ rank_in_cell,
                                                      const uint32_t* __restrict__ cell_start, float4* __restrict__ out, uint32_t* __restrict__ bb) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    // the bounding box has been read back by now: leave it reset for the next build (saves that build a launch)
    if (i < 6) bb[i] = i < 3 ? 0xffffffffu : 0u;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    out[cell_start[cell_of[i]] + rank_in_cell[i]] = make_float4(p[0], p[1], p[2], __int_as_float((int)i));
}

// The whole counting sort in ONE cooperative launch (the build of a 100 k-point map is bound by the host's launch calls,
// not by the device): zero the counts | cell + rank per point | exclusive prefix sum of the counts | scatter, separated by
// grid-wide barriers. The prefix sum gives every CTA one contiguous chunk of the table: chunk totals, barrier, every CTA
// adds up the totals of the chunks before its own (at most a few hundred), then scans its chunk in place.
constexpr int GB_THREADS = 256;
__global__ void __launch_bounds__(GB_THREADS) k_grid_build(const unsigned char* __restrict__ raw, size_t stride, uint32_t n, GridGeom g,
                                                           uint32_t* __restrict__ cell_of, uint32_t* __restrict__ rank_in_cell,
                                                           uint32_t* __restrict__ cell_start, uint32_t ncount, uint32_t* __restrict__ chunk_sum,
                                                           float4* __restrict__ out, uint32_t* __restrict__ bb) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const uint32_t tid = blockIdx.x * GB_THREADS + threadIdx.x, nthr = gridDim.x * GB_THREADS;
    __shared__ uint32_t s_w[GB_THREADS / 32];
    __shared__ uint32_t s_base;
    for (uint32_t i = tid; i < ncount; i += nthr) cell_start[i] = 0u;
    grid.sync();
    for (uint32_t i = tid; i < n; i += nthr) {
        const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
        const uint32_t c = cell_of_point(g, p[0], p[1], p[2]);
        cell_of[i] = c;
        rank_in_cell[i] = atomicAdd(&cell_start[c], 1u);
    }
    grid.sync();
    // ---- exclusive prefix sum of cell_start[0, ncount)
    const uint32_t chunk = (ncount + gridDim.x - 1) / gridDim.x;
    const uint32_t c0 = min(blockIdx.x * chunk, ncount), c1 = min(c0 + chunk, ncount);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto block_sum = [&](uint32_t v) {                    // every thread gets the CTA total
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (lane == 0) s_w[warp] = v;
        __syncthreads();
        uint32_t t = 0;
        for (int w = 0; w < GB_THREADS / 32; w++) t += s_w[w];
        return t;
    };
    {
        uint32_t v = 0;
        for (uint32_t i = c0 + threadIdx.x; i < c1; i += GB_THREADS) v += cell_start[i];
        v = block_sum(v);
        if (threadIdx.x == 0) chunk_sum[blockIdx.x] = v;
    }
    grid.sync();
    {
        uint32_t v = 0;
        for (uint32_t b = threadIdx.x; b < blockIdx.x; b += GB_THREADS) v += __ldcg(&chunk_sum[b]);
        v = block_sum(v);
        if (threadIdx.x == 0) s_base = v;
        __syncthreads();
    }
    uint32_t running = s_base;
    for (uint32_t t0 = c0; t0 < c1; t0 += GB_THREADS) {   // tile of 256 entries: inclusive warp scans + warp totals
        const uint32_t i = t0 + threadIdx.x;
        const uint32_t v = i < c1 ? cell_start[i] : 0u;
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        __syncthreads();
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int w = 0; w < GB_THREADS / 32; w++) { const uint32_t t = s_w[w]; if (w < warp) before += t; total += t; }
        if (i < c1) cell_start[i] = running + before + inc - v;
        running += total;
    }
    grid.sync();
    if (tid < 6) bb[tid] = tid < 3 ? 0xffffffffu : 0u;   // the bounding box has been read back: leave it reset for the next build
    for (uint32_t i = tid; i < n; i += nthr) {
        const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
        out[cell_start[cell_of[i]] + rank_in_cell[i]] = make_float4(p[0], p[1], p[2], __int_as_float((int)i));
    }
}

// CTAs that can be co-resident for the cooperative launch (0: cooperative launch unavailable)
static int grid_build_max_ctas() {
    static int cached = -1;
    if (cached >= 0) return cached;
    int dev = 0, coop = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_grid_build, GB_THREADS, 0) != cudaSuccess) { cudaGetLastError(); cached = 0; return 0; }
    cached = std::min(per_sm, 2) * device_sm_count();
    return cached;
}

int GridIndex::begin(const void* host_pts, size_t stride, size_t n_, float max_dist, cudaStream_t s) {
    n = n_; stride_ = stride; max_dist_ = max_dist;
    dev = GridDev{};
    h = max_dist * 1.0078125f;
    if (n == 0) {
        B2_CHECK(cell_start.reserve(3 * sizeof(uint32_t)));
        B2_CUDA(cudaMemsetAsync(cell_start.p, 0, 3 * sizeof(uint32_t), s));
        dev.pts = nullptr; dev.cell_start = cell_start.as<uint32_t>();
        dev.ox = dev.oy = dev.oz = 0.f; dev.inv_h = 1.0f / h; dev.h = h; dev.nx = dev.ny = dev.nz = 1; dev.n = 0; dev.max_d2 = max_dist * max_dist;
        return B2_OK;
    }
    if (n > 0x7fffffffull) { set_error("grid index: too many points (%zu)", n); return B2_ERR_ARG; }
    tl_rec(0, s);
    if (host_pts) {
        B2_CHECK(raw.reserve(n * stride));
        B2_CUDA(cudaMemcpyAsync(raw.p, host_pts, n * stride, cudaMemcpyHostToDevice, s));
        src_ = raw.as<unsigned char>();
    }
    tl_rec(1, s);
    // bbox on device, one small readback to size the cell table
    B2_CHECK(tmp.reserve(64));
    uint32_t* bb = tmp.as<uint32_t>();
    if (!bb_ready_) { k_bbox_init<<<1, 32, 0, s>>>(bb); count_launch(); }
    bb_ready_ = false;                                           // true again once a scatter that resets it is enqueued
    int nb = (int)std::min<size_t>((n + 255) / 256, (size_t)device_sm_count() * 4);
    k_bbox<<<nb, 256, 0, s>>>(src_, stride, n, bb); count_launch();
    B2_CUDA(cudaGetLastError());
    B2_CHECK(stage.reserve(64));
    B2_CUDA(cudaMemcpyAsync(stage.p, bb, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    tl_rec(2, s);
    return B2_OK;
}

int GridIndex::begin_device(const void* d_pts, size_t stride, size_t n_, float max_dist, cudaStream_t s) {
    src_ = static_cast<const unsigned char*>(d_pts);
    return begin(nullptr, stride, n_, max_dist, s);
}

int GridIndex::finish(cudaStream_t s) {
    if (n == 0) return B2_OK;
    const size_t stride = stride_;
    const float max_dist = max_dist_;
    B2_CUDA(cudaStreamSynchronize(s));
    const uint32_t* hbb = stage.as<uint32_t>();
    auto unflip = [](uint32_t u) { uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; float f; memcpy(&f, &v, 4); return f; };
    float mn[3], mx[3];
    for (int d = 0; d < 3; d++) { mn[d] = unflip(hbb[d]); mx[d] = unflip(hbb[3 + d]); }
    GridGeom g;
    g.inv_h = 1.0f / h;
    if (!(mn[0] <= mx[0])) { mn[0] = mn[1] = mn[2] = 0.f; mx[0] = mx[1] = mx[2] = 0.f; }   // no finite point
    g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2];
    double ex[3];
    for (int d = 0; d < 3; d++) ex[d] = std::floor(((double)mx[d] - (double)mn[d]) * (double)g.inv_h) + 2.0;
    if (ex[0] * ex[1] * ex[2] > (double)GRID_MAX_CELLS) {
        set_error("grid index: %.0f x %.0f x %.0f cells of %.3f m exceed the %zu-cell budget", ex[0], ex[1], ex[2], h, GRID_MAX_CELLS);
        return B2_ERR_TOO_LARGE;
    }
    g.nx = (int)ex[0]; g.ny = (int)ex[1]; g.nz = (int)ex[2];
    g.ncell = (uint32_t)((size_t)g.nx * g.ny * g.nz);
    const size_t ncount = (size_t)g.ncell + 2;
    B2_CHECK(cell_start.reserve(ncount * sizeof(uint32_t)));
    // cell + rank per point, prefix sum of the counts, one scatter (a counting sort: no radix passes)
    const size_t nal = (n + 63) & ~(size_t)63;
    const size_t need = 2 * nal * sizeof(uint32_t) + scan_tmp_bytes(ncount) + 4096;
    B2_CHECK(cell_of.reserve(need));
    B2_CHECK(pts.reserve(n * sizeof(float4)));
    uint32_t* d_cell = cell_of.as<uint32_t>();
    uint32_t* d_rank = d_cell + nal;
    char* scratch = reinterpret_cast<char*>(d_rank + nal);
    const int coop_max = ncount <= 0xffffffffull ? grid_build_max_ctas() : 0;
    if (coop_max > 0) {
        const size_t work = std::max<size_t>(n, ncount);
        int ctas = (int)std::min<size_t>((size_t)coop_max, (work + GB_THREADS * 4 - 1) / (GB_THREADS * 4));
        ctas = std::max(1, std::min(ctas, 1024));                 // chunk totals live in the first 4 KiB of the scratch area
        const unsigned char* a_raw = src_; size_t a_stride = stride; uint32_t a_n = (uint32_t)n, a_ncount = (uint32_t)ncount;
        uint32_t* a_cs = cell_start.as<uint32_t>(); uint32_t* a_chunk = reinterpret_cast<uint32_t*>(scratch);
        float4* a_out = pts.as<float4>(); uint32_t* a_bb = tmp.as<uint32_t>();
        void* args[] = {&a_raw, &a_stride, &a_n, &g, &d_cell, &d_rank, &a_cs, &a_ncount, &a_chunk, &a_out, &a_bb};
        B2_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_grid_build), dim3((unsigned)ctas), dim3(GB_THREADS), args, 0, s)); count_launch();
    } else {
        B2_CUDA(cudaMemsetAsync(cell_start.p, 0, ncount * sizeof(uint32_t), s));
        const unsigned nblk = (unsigned)((n + 255) / 256);
        k_cell_key<<<nblk, 256, 0, s>>>(src_, stride, (uint32_t)n, g, d_cell, d_rank, cell_start.as<uint32_t>()); count_launch();
        B2_CUDA(cudaGetLastError());
        B2_CHECK(exclusive_scan_u32(cell_start.as<uint32_t>(), ncount, scratch, s));
        k_cell_scatter<<<nblk, 256, 0, s>>>(src_, stride, (uint32_t)n, d_cell, d_rank, cell_start.as<uint32_t>(), pts.as<float4>(), tmp.as<uint32_t>()); count_launch();
        B2_CUDA(cudaGetLastError());
    }
    tl_rec(3, s);
    bb_ready_ = true;
    dev.pts = pts.as<float4>(); dev.cell_start = cell_start.as<uint32_t>();
    dev.ox = g.ox; dev.oy = g.oy; dev.oz = g.oz; dev.inv_h = g.inv_h; dev.h = h;
    dev.nx = g.nx; dev.ny = g.ny; dev.nz = g.nz; dev.n = (int)n; dev.max_d2 = max_dist * max_dist;
    return B2_OK;
}

#ifndef B2_GRID_ZERO_COPY_DEFAULT
#define B2_GRID_ZERO_COPY_DEFAULT false
#endif
// ------------------------------------------------------------------------------------------------ device-sized build
// The same counting sort with the geometry decided on the device: bounding box | geometry + zeroed counts | cell + rank per
// point | chunk totals | prefix sum | scatter, six phases of one cooperative launch. The host never waits for the bounding
// box (that wait, two per scan, was a third of the end-to-end scan-to-map step).
__global__ void __launch_bounds__(GB_THREADS) k_grid_build_dev(const unsigned char* __restrict__ raw, size_t stride, uint32_t n, float h, float max_d2,
                                                               uint32_t cell_budget, uint32_t* __restrict__ cell_of, uint32_t* __restrict__ rank_in_cell,
                                                               uint32_t* __restrict__ cell_start, uint32_t* __restrict__ other_table, int active_half,
                                                               uint32_t* __restrict__ chunk_sum, float4* __restrict__ out, GridDevMem* __restrict__ m,
                                                               uint32_t bar_base, const float4* __restrict__ pinned_src, volatile uint32_t* ingest_flag, uint32_t ingest_seq) {
    // Grid barrier on a counter in device memory (arrivals accumulate over the launches: barrier k of this launch is complete at
    // bar_base + (k + 1) * gridDim.x). A plain launch: measured on B200, a cooperative launch of this kernel starts ~12 us later
    // than a plain one, more than the kernel runs. All CTAs are resident by construction (one per SM, the host checks it).
    uint32_t bar_target = bar_base;
    auto grid_sync = [&]() {
        bar_target += gridDim.x;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(&m->bar, 1u);
            while ((int)(*reinterpret_cast<volatile uint32_t*>(&m->bar) - bar_target) < 0) { }
            __threadfence();
        }
        __syncthreads();
    };
    const uint32_t tid = blockIdx.x * GB_THREADS + threadIdx.x, nthr = gridDim.x * GB_THREADS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ uint32_t s_w[GB_THREADS / 32];
    __shared__ uint32_t s_base;
    __shared__ float s_mn[GB_THREADS / 32][3], s_mx[GB_THREADS / 32][3];
    const uint32_t other_dirty = __ldcg(&m->dirty[active_half ^ 1]);   // read before the first barrier, rewritten after the last one
    auto stamp = [&](int k) { if (tid == 0) { long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); m->tl[k] = gt; } };
    stamp(0);
    // ---- bounding box of the finite points
    {
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        auto take = [&](float x, float y, float z) {
            if (isfinite(x) && isfinite(y) && isfinite(z)) {
                mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
                mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
            }
        };
        if (pinned_src) {
            // the upload itself: 16-byte records from pinned host memory, four loads in flight per thread, into the device copy
            float4* dst = reinterpret_cast<float4*>(const_cast<unsigned char*>(raw));
            uint32_t i = tid;
            for (; i + 3 * nthr < n; i += 4 * nthr) {
                const float4 a = __ldcs(pinned_src + i), b = __ldcs(pinned_src + i + nthr), c = __ldcs(pinned_src + i + 2 * nthr), d = __ldcs(pinned_src + i + 3 * nthr);
                dst[i] = a; dst[i + nthr] = b; dst[i + 2 * nthr] = c; dst[i + 3 * nthr] = d;
                take(a.x, a.y, a.z); take(b.x, b.y, b.z); take(c.x, c.y, c.z); take(d.x, d.y, d.z);
            }
            for (; i < n; i += nthr) { const float4 a = __ldcs(pinned_src + i); dst[i] = a; take(a.x, a.y, a.z); }
        } else {
            for (uint32_t i = tid; i < n; i += nthr) {
                const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
                take(p[0], p[1], p[2]);
            }
        }
#pragma unroll
        for (int d = 0; d < 3; d++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
                mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < 3; d++) { s_mn[warp][d] = mn[d]; s_mx[warp][d] = mx[d]; }
        }
        __syncthreads();
        if (threadIdx.x < 3) {
            const int d = threadIdx.x;
            float a = s_mn[0][d], b = s_mx[0][d];
            for (int w = 1; w < GB_THREADS / 32; w++) { a = fminf(a, s_mn[w][d]); b = fmaxf(b, s_mx[w][d]); }
            if (a <= b) { atomicMin(&m->bb[d], float_flip(a)); atomicMax(&m->bb[3 + d], float_flip(b)); }
        }
    }
    grid_sync();
    if (pinned_src && tid == 0) { *ingest_flag = ingest_seq; __threadfence_system(); }      // every CTA is past its last read of the caller's buffer
    stamp(1);
    // ---- geometry: every thread derives the same numbers (the arithmetic of GridIndex::finish)
    GridGeom g;
    {
        float mn[3], mx[3];
#pragma unroll
        for (int d = 0; d < 3; d++) { mn[d] = float_unflip(__ldcg(&m->bb[d])); mx[d] = float_unflip(__ldcg(&m->bb[3 + d])); }
        if (!(mn[0] <= mx[0])) { mn[0] = mn[1] = mn[2] = 0.f; mx[0] = mx[1] = mx[2] = 0.f; }     // no finite point
        g.inv_h = 1.0f / h;
        g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2];
        double ex[3];
#pragma unroll
        for (int d = 0; d < 3; d++) ex[d] = floor(((double)mx[d] - (double)mn[d]) * (double)g.inv_h) + 2.0;
        if (ex[0] * ex[1] * ex[2] > (double)cell_budget) {
            if (tid == 0) m->status = B2_ERR_TOO_LARGE;      // bb is left as it is: rebuild_exact() resets it
            if (threadIdx.x == 0) atomicAdd(&m->bar, 3u);    // the three barriers nobody will reach (the host counts four per launch)
            return;                                          // uniform over the whole grid
        }
        g.nx = (int)ex[0]; g.ny = (int)ex[1]; g.nz = (int)ex[2];
        g.ncell = (uint32_t)((size_t)g.nx * g.ny * g.nz);
    }
    const uint32_t ncount = g.ncell + 2;
    // the table arrives zeroed (the previous build cleared it after its scatter): counting starts right away. Four points per
    // thread are in flight (loads first, then the counting atomics): a 100 k-point map is 2-3 points per thread on 148 CTAs.
    for (uint32_t i0 = tid; i0 < n; i0 += 4 * nthr) {
        uint32_t c[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t i = i0 + k * nthr;
            if (i < n) { const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride); c[k] = cell_of_point(g, p[0], p[1], p[2]); }
        }
        uint32_t r[4];
#pragma unroll
        for (int k = 0; k < 4; k++) if (i0 + k * nthr < n) r[k] = atomicAdd(&cell_start[c[k]], 1u);
#pragma unroll
        for (int k = 0; k < 4; k++) if (i0 + k * nthr < n) { cell_of[i0 + k * nthr] = c[k]; rank_in_cell[i0 + k * nthr] = r[k]; }
    }
    grid_sync();
    stamp(2);
    // ---- exclusive prefix sum of cell_start[0, ncount): one contiguous chunk per CTA, a multiple of 8 entries. The usual map
    // (a few hundred thousand cells) gives every thread at most 8 consecutive entries: they are loaded once (two 128-bit loads),
    // stay in registers across the barrier and are written back as offsets — one pass over the table instead of two.
    const uint32_t chunk = (((ncount + gridDim.x - 1) / gridDim.x) + 7u) & ~7u;
    const uint32_t c0 = min(blockIdx.x * chunk, ncount), c1 = min(c0 + chunk, ncount);
    const bool in_regs = chunk <= GB_THREADS * 8;
    auto block_sum = [&](uint32_t v) {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (lane == 0) s_w[warp] = v;
        __syncthreads();
        uint32_t t = 0;
        for (int w = 0; w < GB_THREADS / 32; w++) t += s_w[w];
        return t;
    };
    uint32_t e[8];
    uint32_t mine = 0;
    {
        uint32_t v = 0;
        if (in_regs) {
            const uint32_t b = c0 + threadIdx.x * 8;          // c0 and the table are 32-byte aligned
            if (b + 8 <= c1) {
                const uint4 lo = *reinterpret_cast<const uint4*>(cell_start + b), hi = *reinterpret_cast<const uint4*>(cell_start + b + 4);
                e[0] = lo.x; e[1] = lo.y; e[2] = lo.z; e[3] = lo.w; e[4] = hi.x; e[5] = hi.y; e[6] = hi.z; e[7] = hi.w;
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) e[k] = (b + k < c1) ? cell_start[b + k] : 0u;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) v += e[k];
            mine = v;
        } else {
            for (uint32_t i = c0 + threadIdx.x; i < c1; i += GB_THREADS) v += cell_start[i];
        }
        v = block_sum(v);
        if (threadIdx.x == 0) chunk_sum[blockIdx.x] = v;
    }
    grid_sync();
    stamp(3);
    {
        uint32_t v = 0;
        for (uint32_t b = threadIdx.x; b < blockIdx.x; b += GB_THREADS) v += __ldcg(&chunk_sum[b]);
        v = block_sum(v);
        if (threadIdx.x == 0) s_base = v;
        __syncthreads();
    }
    if (in_regs) {
        uint32_t inc = mine;                                   // exclusive scan of the per-thread totals over the CTA
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        uint32_t before = 0;
        for (int w = 0; w < warp; w++) before += s_w[w];
        uint32_t run = s_base + before + inc - mine;
        const uint32_t b = c0 + threadIdx.x * 8;
        uint32_t o8[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { o8[k] = run; run += e[k]; }
        if (b + 8 <= c1) {
            *reinterpret_cast<uint4*>(cell_start + b) = make_uint4(o8[0], o8[1], o8[2], o8[3]);
            *reinterpret_cast<uint4*>(cell_start + b + 4) = make_uint4(o8[4], o8[5], o8[6], o8[7]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) if (b + k < c1) cell_start[b + k] = o8[k];
        }
    } else {
        uint32_t running = s_base;
        for (uint32_t t0 = c0; t0 < c1; t0 += GB_THREADS) {
            const uint32_t i = t0 + threadIdx.x;
            const uint32_t v = i < c1 ? cell_start[i] : 0u;
            uint32_t inc = v;
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
            __syncthreads();
            if (lane == 31) s_w[warp] = inc;
            __syncthreads();
            uint32_t before = 0, total = 0;
            for (int w = 0; w < GB_THREADS / 32; w++) { const uint32_t t = s_w[w]; if (w < warp) before += t; total += t; }
            if (i < c1) cell_start[i] = running + before + inc - v;
            running += total;
        }
    }
    grid_sync();
    stamp(4);
    for (uint32_t i = tid; i < n; i += nthr) {
        const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
        out[cell_start[cell_of[i]] + rank_in_cell[i]] = make_float4(p[0], p[1], p[2], __int_as_float((int)i));
    }
    stamp(5);
    // the other half of the table (the previous build's offsets, nobody reads them any more) is cleared for the next build
    for (uint32_t i = tid; i < other_dirty; i += nthr) other_table[i] = 0u;
    if (tid == 0) {
        m->dirty[active_half] = ncount; m->dirty[active_half ^ 1] = 0u;
        stamp(6);
        GridDev d;
        d.pts = out; d.cell_start = cell_start;
        d.ox = g.ox; d.oy = g.oy; d.oz = g.oz; d.inv_h = g.inv_h; d.h = h;
        d.nx = g.nx; d.ny = g.ny; d.nz = g.nz; d.n = (int)n; d.max_d2 = max_d2;
        m->g = d;
        m->status = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) m->bb[k] = k < 3 ? 0xffffffffu : 0u;      // ready for the next build
    }
}

static int grid_build_dev_max_ctas() {
    static int cached = -1;
    if (cached >= 0) return cached;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_grid_build_dev, GB_THREADS, 0) != cudaSuccess) { cudaGetLastError(); cached = 0; return 0; }
    // four builds may be in flight on one device (two handles, two maps each): each gets at most a quarter of the resident slots
    cached = per_sm >= 4 ? device_sm_count() : (per_sm * device_sm_count()) / 4;
    return cached;
}

constexpr size_t GRID_DEFAULT_BUDGET = (size_t)4 << 20;       // cells of a device-sized table before anything is known: 16 MiB

static int store_devmem(GridIndex& gi, const GridDev& d, int status, cudaStream_t s) {
    GridDevMem hm{};
    hm.g = d; hm.status = status;
    for (int k = 0; k < 6; k++) hm.bb[k] = k < 3 ? 0xffffffffu : 0u;
    B2_CUDA(cudaMemcpyAsync(gi.devmem.p, &hm, sizeof(hm), cudaMemcpyHostToDevice, s));   // pageable source: staged before the call returns
    gi.bar_total_ = 0;
    return B2_OK;
}

int GridIndex::upload_async(const void* host_pts, const void* dev_pts, size_t stride, size_t n_, float max_dist, cudaStream_t s, bool copy_dev) {
    n = n_; stride_ = stride; max_dist_ = max_dist;
    h = max_dist * 1.0078125f;
    if (n > 0x7fffffffull) { set_error("grid index: too many points (%zu)", n); return B2_ERR_ARG; }
    if (!devmem.p) {
        B2_CHECK(devmem.reserve(sizeof(GridDevMem)));
        GridDev none{};
        B2_CHECK(store_devmem(*this, none, 0, s));
    }
    tl_rec(0, s);
    pinned_src_ = nullptr;
    if (n && host_pts) {
        B2_CHECK(raw.reserve(n * stride));
        // Pinned 16-byte records: the build kernel's first phase reads them over PCIe itself (128-bit loads from mapped host
        // memory), writes the device copy and reduces the bounding box on the way — no DMA descriptor, no copy -> kernel
        // hand-off. Measured (B2_GRID_ZERO_COPY=1): the PCIe read runs at the copy engine's speed (1.3 MB in 43 us), and as long as
        // b2_s2m_set_map keeps its promise that the caller's buffers are free when it returns (it waits for the kernel's last
        // read, as it waits for the DMA) the end-to-end step is the same 159-160 us either way; returning early instead would save
        // 5-10 us and break that promise. Off by default; anything else goes through the copy engine.
        const bool zero_copy = getenv("B2_GRID_ZERO_COPY") ? atoi(getenv("B2_GRID_ZERO_COPY")) != 0 : B2_GRID_ZERO_COPY_DEFAULT;
        if (zero_copy && stride == 16 && (reinterpret_cast<uintptr_t>(host_pts) & 15u) == 0 && grid_build_dev_max_ctas() > 0) {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, host_pts) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
                pinned_src_ = static_cast<const unsigned char*>(at.devicePointer);
            else cudaGetLastError();
        }
        if (!pinned_src_) B2_CUDA(cudaMemcpyAsync(raw.p, host_pts, n * stride, cudaMemcpyHostToDevice, s));
        src_ = raw.as<unsigned char>();
    } else if (n && dev_pts && copy_dev && dev_pts != raw.p) {
        // device points owned by somebody else (the local map's VoxelGrid output): keep a private copy so that the build,
        // which is still queued when this returns, does not depend on the owner leaving them alone
        B2_CHECK(raw.reserve(n * stride));
        B2_CUDA(cudaMemcpyAsync(raw.p, dev_pts, n * stride, cudaMemcpyDeviceToDevice, s));
        src_ = raw.as<unsigned char>();
    } else {
        src_ = static_cast<const unsigned char*>(dev_pts);
    }
    tl_rec(1, s);
    return B2_OK;
}

int GridIndex::build_async(cudaStream_t s) {
    const float max_dist = max_dist_;
    dev = GridDev{};                                  // the host copy is not valid after a device-sized build
    if (n == 0) {
        B2_CHECK(cell_start.reserve(3 * sizeof(uint32_t)));
        B2_CUDA(cudaMemsetAsync(cell_start.p, 0, 3 * sizeof(uint32_t), s));
        GridDev d{};
        d.pts = nullptr; d.cell_start = cell_start.as<uint32_t>();
        d.ox = d.oy = d.oz = 0.f; d.inv_h = 1.0f / h; d.h = h; d.nx = d.ny = d.nz = 1; d.n = 0; d.max_d2 = max_dist * max_dist;
        dev = d;
        tables_clean_ = false;
        return store_devmem(*this, d, 0, s);
    }
    const int coop_max = grid_build_dev_max_ctas();
    if (coop_max <= 0) {
        if (pinned_src_) { B2_CUDA(cudaMemcpyAsync(raw.p, pinned_src_, n * stride_, cudaMemcpyDefault, s)); pinned_src_ = nullptr; }
        return rebuild_exact(s);
    }
    if (cell_budget == 0) cell_budget = GRID_DEFAULT_BUDGET;
    // two tables of cell_budget + 2 entries: the build counts into the one the previous build left zeroed, and zeroes the other
    // one (its predecessor's offsets) when it is done — no clearing pass and no barrier for it on the critical path
    const size_t half = (cell_budget + 2 + 63) & ~(size_t)63;
    const void* before = cell_start.p;
    B2_CHECK(cell_start.reserve(2 * half * sizeof(uint32_t)));
    if (cell_start.p != before || !tables_clean_) {
        B2_CUDA(cudaMemsetAsync(cell_start.p, 0, 2 * half * sizeof(uint32_t), s));
        const uint32_t zero2[2] = {0u, 0u};
        B2_CUDA(cudaMemcpyAsync(&devmem.as<GridDevMem>()->dirty[0], zero2, sizeof(zero2), cudaMemcpyHostToDevice, s));
        tables_clean_ = true; table_active_ = 0;
    }
    table_active_ ^= 1;
    uint32_t* a_cs = cell_start.as<uint32_t>() + (size_t)table_active_ * half;
    uint32_t* a_other = cell_start.as<uint32_t>() + (size_t)(table_active_ ^ 1) * half;
    int a_half = table_active_;
    const size_t nal = (n + 63) & ~(size_t)63;
    B2_CHECK(cell_of.reserve(2 * nal * sizeof(uint32_t) + 4096));
    B2_CHECK(pts.reserve(n * sizeof(float4)));
    uint32_t* d_cell = cell_of.as<uint32_t>();
    uint32_t* d_rank = d_cell + nal;
    uint32_t* d_chunk = d_rank + nal;                 // chunk totals: at most 1024 CTAs
    // one CTA per SM: the phases are short (a 100 k-point map), what counts is the cost of the four grid barriers
    int ctas = std::min(coop_max, device_sm_count());
    ctas = std::max(1, std::min(ctas, 1024));
    static const int env_ctas = getenv("B2_GRID_CTAS") ? atoi(getenv("B2_GRID_CTAS")) : 0;          // kernel experiments
    if (env_ctas > 0) ctas = std::min(env_ctas, coop_max);
    const unsigned char* a_raw = src_; size_t a_stride = stride_; uint32_t a_n = (uint32_t)n; float a_h = h, a_md2 = max_dist * max_dist;
    uint32_t a_budget = (uint32_t)std::min<size_t>(cell_budget, 0xfffffff0u);
    float4* a_out = pts.as<float4>(); GridDevMem* a_m = devmem.as<GridDevMem>();
    volatile uint32_t* a_flag = nullptr; uint32_t a_seq = 0;
    if (pinned_src_) {
        if (!ingest_flag_.p) { B2_CHECK(ingest_flag_.reserve(64)); *ingest_flag_.as<uint32_t>() = 0; }
        a_flag = ingest_flag_.as<volatile uint32_t>(); a_seq = ++ingest_seq_;
    }
    k_grid_build_dev<<<(unsigned)ctas, GB_THREADS, 0, s>>>(a_raw, a_stride, a_n, a_h, a_md2, a_budget, d_cell, d_rank, a_cs, a_other, a_half, d_chunk, a_out, a_m, bar_total_, reinterpret_cast<const float4*>(pinned_src_), a_flag, a_seq);
    count_launch();
    B2_CUDA(cudaGetLastError());
    ingest_wait_ = pinned_src_ != nullptr;
    pinned_src_ = nullptr;                            // the device copy exists once this launch has run: later rebuilds read it
    bar_total_ += 4u * (uint32_t)ctas;
    tl_rec(3, s);
    return B2_OK;
}

int GridIndex::rebuild_exact(cudaStream_t s) {
    tables_clean_ = false;                            // the host-sized path uses the front of cell_start as one table
    const unsigned char* keep = src_;
    B2_CHECK(begin_device(keep, stride_, n, max_dist_, s));
    B2_CHECK(finish(s));
    if (!devmem.p) B2_CHECK(devmem.reserve(sizeof(GridDevMem)));
    B2_CHECK(store_devmem(*this, dev, 0, s));
    const size_t cells = (size_t)dev.nx * dev.ny * dev.nz;
    cell_budget = std::min<size_t>(GRID_MAX_CELLS, std::max(cell_budget, cells + cells / 2));
    return B2_OK;
}

// ------------------------------------------------------------------------------------------------ batched kNN
template <int K>
__global__ void __launch_bounds__(256) k_knn(GridDev g, const unsigned char* __restrict__ q, size_t stride, uint32_t m,
                                             float max_d2, int32_t* __restrict__ idx, float* __restrict__ d2) {
    constexpr int LPF = 8;
    const uint32_t qi = (blockIdx.x * blockDim.x + threadIdx.x) / LPF;
    const bool active = qi < m;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) { const float* p = reinterpret_cast<const float*>(q + (size_t)qi * stride); qx = p[0]; qy = p[1]; qz = p[2]; }
    unsigned long long key[K]; uint32_t pos[K];
    knn_group<K, LPF>(g, qx, qy, qz, active, INFINITY, key, pos);
    const int sub = threadIdx.x & (LPF - 1);
    if (active && sub == 0) {
#pragma unroll
        for (int r = 0; r < K; r++) {
            float d = __uint_as_float((uint32_t)(key[r] >> 32));
            bool ok = key[r] != KNN_EMPTY && d < max_d2;
            idx[(size_t)qi * K + r] = ok ? (int32_t)(uint32_t)key[r] : -1;
            d2[(size_t)qi * K + r] = ok ? d : INFINITY;
        }
    }
}

}  // namespace b2

struct b2_knn_s {
    b2::GridIndex grid;
    b2::DevBuf q, oidx, od2;
    cudaStream_t stream = nullptr;
    float max_dist = 1.0f;
    int device = b2::current_device();      // the device the handle was created on
    bool built = false;
};

extern "C" {

int b2_knn_create(b2_knn_t* out, float max_dist) {
    if (!out || !(max_dist > 0.f)) { b2::set_error("b2_knn_create: bad argument"); return B2_ERR_ARG; }
    b2_knn_s* h = new b2_knn_s();
    h->max_dist = max_dist;
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { b2::set_error("cudaStreamCreate -> %s", cudaGetErrorString(e)); delete h; return B2_ERR_CUDA; }
    *out = h;
    return B2_OK;
}

int b2_knn_destroy(b2_knn_t h) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h) return B2_ERR_ARG;
    h->grid.release(); h->q.release(); h->oidx.release(); h->od2.release();
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return B2_OK;
}

int b2_knn_set_input_cloud(b2_knn_t h, const void* pts, size_t stride, size_t n) {
    B2_NVTX("b2_knn_set_input_cloud");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || (n && !pts) || stride < 12 || (stride & 3)) { b2::set_error("b2_knn_set_input_cloud: bad argument"); return B2_ERR_ARG; }
    B2_CHECK(h->grid.build(pts, stride, n, h->max_dist, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    h->built = true;
    return B2_OK;
}

int b2_knn_nearest_k_search(b2_knn_t h, const void* queries, size_t stride, size_t m, int k, int32_t* indices, float* sq_dists) {
    B2_NVTX("b2_knn_nearest_k_search");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || (m && (!queries || !indices || !sq_dists)) || stride < 12 || (stride & 3) || k < 1 || k > 8) {
        b2::set_error("b2_knn_nearest_k_search: bad argument"); return B2_ERR_ARG;
    }
    if (!h->built) { b2::set_error("b2_knn_nearest_k_search: no input cloud"); return B2_ERR_STATE; }
    if (m == 0) return B2_OK;
    B2_CHECK(h->q.reserve(m * stride));
    B2_CHECK(h->oidx.reserve(m * k * sizeof(int32_t)));
    B2_CHECK(h->od2.reserve(m * k * sizeof(float)));
    B2_CUDA(cudaMemcpyAsync(h->q.p, queries, m * stride, cudaMemcpyHostToDevice, h->stream));
    const unsigned nblk = (unsigned)((m * 8 + 255) / 256);
    const float md2 = h->max_dist * h->max_dist;
#define B2_KNN_CASE(KK) case KK: b2::k_knn<KK><<<nblk, 256, 0, h->stream>>>(h->grid.dev, h->q.as<unsigned char>(), stride, (uint32_t)m, md2, \
                                                                          h->oidx.as<int32_t>(), h->od2.as<float>()); b2::count_launch(); break;
    switch (k) {
        B2_KNN_CASE(1) B2_KNN_CASE(2) B2_KNN_CASE(3) B2_KNN_CASE(4) B2_KNN_CASE(5) B2_KNN_CASE(6) B2_KNN_CASE(7) B2_KNN_CASE(8)
    }
#undef B2_KNN_CASE
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaMemcpyAsync(indices, h->oidx.p, m * k * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaMemcpyAsync(sq_dists, h->od2.p, m * k * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    return B2_OK;
}

}  // extern "C"
