// Device-resident fp64 cloud: grid index build, Open3D-style voxel_down_sample and estimate_normals (32-ary BVH over
// the Morton order), rigid transform.
// Replaces, for Multi_LiCa's GICP calibration (Calibration_Tookit/Multi_LiCa/multi_lidar_calibrator/calibration/Calibration.py):
//   :314-315  pcd.voxel_down_sample(voxel_size)   voxel = floor((p - (min - v/2)) / v), one double mean per voxel
//   :327-328  pcd.estimate_normals()              KDTreeSearchParamKNN(30): covariance of the 30 nearest neighbours
//                                                 (the point itself included), eigenvector of the smallest eigenvalue
//   Calibration.py:347-358 / Lidar.py             pcd.transform(T)
// Definitions shared with the oracle (Open3D leaves them unspecified): downsampled points come out in ascending
// (z, y, x) voxel order, sums inside a voxel run in ascending input index, neighbour ties fall to the smaller index,
// the normal's sign makes the first non-zero of (z, y, x) positive.
#include "b2_cloud.cuh"
#include "b2_comm.cuh"
#include "b2_bvh.cuh"
#include <cmath>
#include <algorithm>

namespace b2 {

// ------------------------------------------------------------------------------------------------ bbox (fp64)
__device__ __forceinline__ unsigned long long dflip(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
static inline double dunflip(unsigned long long u) {
    unsigned long long v = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    double d; memcpy(&d, &v, 8); return d;
}

__global__ void k_bboxd_init(unsigned long long* bb) {
    if (threadIdx.x < 3) bb[threadIdx.x] = ~0ull;
    else if (threadIdx.x < 6) bb[threadIdx.x] = 0ull;
}

__global__ void __launch_bounds__(256) k_bboxd(const double* __restrict__ xyz, size_t n, unsigned long long* __restrict__ bb) {
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        if (isfinite(x) && isfinite(y) && isfinite(z)) {
            mn[0] = fmin(mn[0], x); mn[1] = fmin(mn[1], y); mn[2] = fmin(mn[2], z);
            mx[0] = fmax(mx[0], x); mx[1] = fmax(mx[1], y); mx[2] = fmax(mx[2], z);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fmin(mn[d], shfl_xor_d(0xffffffffu, mn[d], o));
            mx[d] = fmax(mx[d], shfl_xor_d(0xffffffffu, mx[d], o));
        }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int d = 0; d < 3; d++)
            if (mn[d] <= mx[d]) { atomicMin(&bb[d], dflip(mn[d])); atomicMax(&bb[3 + d], dflip(mx[d])); }
}

int bbox_f64(const double* d_xyz, size_t n, DevBuf& scratch, cudaStream_t s, double mn[3], double mx[3]) {
    for (int d = 0; d < 3; d++) { mn[d] = INFINITY; mx[d] = -INFINITY; }
    if (n == 0) return B2_OK;
    B2_CHECK(scratch.reserve(64));
    unsigned long long* bb = scratch.as<unsigned long long>();
    k_bboxd_init<<<1, 32, 0, s>>>(bb); count_launch();
    const int nb = (int)std::min<size_t>((n + 255) / 256, (size_t)device_sm_count() * 8);
    k_bboxd<<<nb, 256, 0, s>>>(d_xyz, n, bb); count_launch();
    B2_CUDA(cudaGetLastError());
    unsigned long long h[6];
    B2_CUDA(cudaMemcpyAsync(h, bb, sizeof(h), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    if (h[0] == ~0ull) return B2_OK;           // no finite point
    for (int d = 0; d < 3; d++) { mn[d] = dunflip(h[d]); mx[d] = dunflip(h[3 + d]); }
    return B2_OK;
}

// ------------------------------------------------------------------------------------------------ grid build
struct GridGeomD { double ox, oy, oz, inv_h; int nx, ny, nz; uint32_t ncell; };

__device__ __forceinline__ uint32_t cell_of_point_d(const GridGeomD& g, double x, double y, double z) {
    if (!(isfinite(x) && isfinite(y) && isfinite(z))) return g.ncell;
    int cx = (int)floor((x - g.ox) * g.inv_h), cy = (int)floor((y - g.oy) * g.inv_h), cz = (int)floor((z - g.oz) * g.inv_h);
    cx = min(max(cx, 0), g.nx - 1); cy = min(max(cy, 0), g.ny - 1); cz = min(max(cz, 0), g.nz - 1);
    return (uint32_t)(((size_t)cz * g.ny + cy) * g.nx + cx);
}

// count[c]++ ; optionally record (key, input index) for the sort
__global__ void __launch_bounds__(256) k_celld_count(const double* __restrict__ xyz, uint32_t n, GridGeomD g, uint32_t* __restrict__ count,
                                                     uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = cell_of_point_d(g, xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]);
    atomicAdd(&count[c], 1u);
    if (keys) { keys[i] = c; vals[i] = i; }
}

__global__ void __launch_bounds__(256) k_count_nonzero(const uint32_t* __restrict__ count, size_t ncell, unsigned long long* __restrict__ out) {
    unsigned local = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (size_t)gridDim.x * blockDim.x) local += count[i] != 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, (unsigned long long)local);
}

__global__ void __launch_bounds__(256) k_celld_gather(const double* __restrict__ xyz, uint32_t n, const uint32_t* __restrict__ order, P4d* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t src = order[i];
    P4d p; p.x = xyz[3 * (size_t)src]; p.y = xyz[3 * (size_t)src + 1]; p.z = xyz[3 * (size_t)src + 2]; p.idx = (long long)src;
    double2* o = reinterpret_cast<double2*>(&out[i]);
    o[0] = make_double2(p.x, p.y); o[1] = make_double2(p.z, __longlong_as_double(p.idx));
}

static double cells_at(const double ext[3], double h) {
    double c = 1.0;
    for (int d = 0; d < 3; d++) c *= std::floor(ext[d] / h) + 1.0;
    return c;
}

int GridD::build(const double* d_xyz, size_t n_, double h_request, double target_ppc, cudaStream_t s) {
    n = n_; dev = GridDDev{}; ppc = 0.0;
    if (n > 0x7fffffffull) { set_error("grid index: too many points (%zu)", n); return B2_ERR_ARG; }
    double mn[3], mx[3];
    B2_CHECK(bbox_f64(d_xyz, n, work, s, mn, mx));
    if (n == 0 || !(mn[0] <= mx[0])) {
        B2_CHECK(cell_start.reserve(3 * sizeof(uint32_t)));
        B2_CUDA(cudaMemsetAsync(cell_start.p, 0, 3 * sizeof(uint32_t), s));
        dev.pts = nullptr; dev.cell_start = cell_start.as<uint32_t>();
        dev.h = h_request > 0 ? h_request : 1.0; dev.inv_h = 1.0 / dev.h; dev.nx = dev.ny = dev.nz = 1; dev.n = 0;
        return B2_OK;
    }
    double ext[3];
    for (int d = 0; d < 3; d++) ext[d] = mx[d] - mn[d];
    const double emax = std::max(ext[0], std::max(ext[1], ext[2]));
    const double max_cells = (double)std::min<size_t>(std::max<size_t>(16 * n, (size_t)1 << 16), (size_t)1 << 30);
    // smallest admissible edge under the cell budget (cells_at is non-increasing in h)
    double h_floor = 0.0;
    if (emax > 0) {
        double lo = emax * 1e-9, hi = emax + 1.0;
        if (cells_at(ext, lo) <= max_cells) h_floor = lo;
        else {
            for (int it = 0; it < 100; it++) { const double mid = 0.5 * (lo + hi); if (cells_at(ext, mid) <= max_cells) hi = mid; else lo = mid; }
            h_floor = hi;
        }
    } else h_floor = 1.0;
    double h;
    const bool adapt = !(h_request > 0);
    if (!adapt) h = std::max(h_request, h_floor);
    else {
        double vol = 1.0;
        for (int d = 0; d < 3; d++) vol *= std::max(ext[d], emax * 1e-3 + 1e-12);
        h = std::max(h_floor, std::cbrt(vol * target_ppc / (double)n));
    }
    GridGeomD g;
    const unsigned nblk = (unsigned)((n + 255) / 256);
    const size_t nal = (n + 63) & ~(size_t)63;
    unsigned long long* d_occ = nullptr;
    for (int attempt = 0;; attempt++) {
        g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2]; g.inv_h = 1.0 / h;
        g.nx = (int)(std::floor(ext[0] / h) + 1.0); g.ny = (int)(std::floor(ext[1] / h) + 1.0); g.nz = (int)(std::floor(ext[2] / h) + 1.0);
        g.ncell = (uint32_t)((size_t)g.nx * g.ny * g.nz);
        const size_t ncount = (size_t)g.ncell + 2;
        B2_CHECK(cell_start.reserve(ncount * sizeof(uint32_t)));
        B2_CUDA(cudaMemsetAsync(cell_start.p, 0, ncount * sizeof(uint32_t), s));
        const size_t need = 4 * nal * sizeof(uint32_t) + sort_tmp_bytes(n) + scan_tmp_bytes(ncount) + 1024;
        B2_CHECK(work.reserve(need));
        uint32_t* ka = work.as<uint32_t>();
        uint32_t* va = ka + nal;
        k_celld_count<<<nblk, 256, 0, s>>>(d_xyz, (uint32_t)n, g, cell_start.as<uint32_t>(), ka, va); count_launch();
        B2_CUDA(cudaGetLastError());
        // occupancy of this edge
        d_occ = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(work.p) + need - 512);
        B2_CUDA(cudaMemsetAsync(d_occ, 0, 8, s));
        const int nb = (int)std::min<size_t>(((size_t)g.ncell + 255) / 256, (size_t)device_sm_count() * 8);
        k_count_nonzero<<<nb, 256, 0, s>>>(cell_start.as<uint32_t>(), (size_t)g.ncell, d_occ); count_launch();
        unsigned long long occ = 0;
        B2_CUDA(cudaMemcpyAsync(&occ, d_occ, 8, cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaStreamSynchronize(s));
        ppc = occ ? (double)n / (double)occ : 0.0;
        if (!adapt || attempt >= 5 || occ == 0) break;
        const double ratio = ppc / target_ppc;
        if (ratio >= 0.7 && ratio <= 1.5) break;
        if (ratio > 1.5 && h <= h_floor) break;
        // points lie on surfaces: occupancy grows with h^2
        h = std::max(h_floor, h * std::min(4.0, std::max(0.25, std::sqrt(1.0 / ratio))));
    }
    const size_t ncount = (size_t)g.ncell + 2;
    uint32_t* ka = work.as<uint32_t>();
    uint32_t* va = ka + nal; uint32_t* kb = va + nal; uint32_t* vb = kb + nal;
    char* scratch = reinterpret_cast<char*>(vb + nal);
    B2_CHECK(exclusive_scan_u32(cell_start.as<uint32_t>(), ncount, scratch + sort_tmp_bytes(n), s));
    int bits = 1;
    while (((size_t)1 << bits) <= (size_t)g.ncell) bits++;
    uint32_t *ks, *vs;
    B2_CHECK(radix_sort_pairs(ka, va, kb, vb, n, bits, scratch, s, &ks, &vs));
    B2_CHECK(pts.reserve(n * sizeof(P4d)));
    k_celld_gather<<<nblk, 256, 0, s>>>(d_xyz, (uint32_t)n, vs, pts.as<P4d>()); count_launch();
    B2_CUDA(cudaGetLastError());
    dev.pts = pts.as<P4d>(); dev.cell_start = cell_start.as<uint32_t>();
    dev.ox = g.ox; dev.oy = g.oy; dev.oz = g.oz; dev.h = h; dev.inv_h = g.inv_h;
    dev.nx = g.nx; dev.ny = g.ny; dev.nz = g.nz;
    // valid points = everything before the non-finite bucket; read lazily by kernels through cell_start[ncell]
    dev.n = (uint32_t)n;
    return B2_OK;
}

// one thread per cell: the cell's record (run of its points + their box in 1/256 cell steps relative to the cell's corner)
// and the fp32 copy of its points relative to the same corner — the expression nn1_scan_block_pruned evaluates on the query's side
__global__ void __launch_bounds__(256) k_grid_cell_records(GridDDev g, size_t ncell, uint4* __restrict__ rec, float4* __restrict__ rel) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const uint32_t b = g.cell_start[c], e = g.cell_start[c + 1];
    if (b >= e) { rec[c] = make_uint4(b, b, 0u, 0u); return; }
    const int cx = (int)(c % (size_t)g.nx), cy = (int)((c / (size_t)g.nx) % (size_t)g.ny), cz = (int)(c / ((size_t)g.nx * g.ny));
    const double cox = g.ox + (double)cx * g.h, coy = g.oy + (double)cy * g.h, coz = g.oz + (double)cz * g.h;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t p = b; p < e; p++) {
        double x, y, z; long long id;
        load_p4d(&g.pts[p], x, y, z, id);
        x -= cox; y -= coy; z -= coz;
        rel[p] = make_float4((float)x, (float)y, (float)z, 0.f);
        lo[0] = fmin(lo[0], x); lo[1] = fmin(lo[1], y); lo[2] = fmin(lo[2], z);
        hi[0] = fmax(hi[0], x); hi[1] = fmax(hi[1], y); hi[2] = fmax(hi[2], z);
    }
    // step index of a coordinate: floor(v / h * 256) clamped to the cell (a point a rounding error outside its cell lands on
    // the first / last step; the search widens the box by 1e-6 h, far above that)
    uint32_t ql = 0, qh = 0;
    for (int a = 0; a < 3; a++) {
        const double sl = floor(lo[a] * g.inv_h * 256.0), sh = floor(hi[a] * g.inv_h * 256.0);
        ql |= (uint32_t)fmin(fmax(sl, 0.0), 255.0) << (8 * a);
        qh |= (uint32_t)fmin(fmax(sh, 0.0), 255.0) << (8 * a);
    }
    rec[c] = make_uint4(b, e, ql, qh);
}

int GridD::build_cell_records(cudaStream_t s) {
    dev.cell_rec = nullptr; dev.pts_rel = nullptr;
    if (!dev.pts || !n) return B2_OK;
    const size_t ncell = (size_t)dev.nx * dev.ny * dev.nz;
    B2_CHECK(cell_rec.reserve(ncell * sizeof(uint4)));
    B2_CHECK(pts_rel.reserve(n * sizeof(float4)));
    k_grid_cell_records<<<(unsigned)((ncell + 255) / 256), 256, 0, s>>>(dev, ncell, cell_rec.as<uint4>(), pts_rel.as<float4>()); count_launch();
    B2_CUDA(cudaGetLastError());
    dev.cell_rec = cell_rec.as<uint4>(); dev.pts_rel = pts_rel.as<float4>();
    return B2_OK;
}

// ------------------------------------------------------------------------------------------------ voxel_down_sample
struct VdsGeom { double vmin[3]; double voxel; unsigned long long nx, ny; unsigned long long invalid; };

__global__ void __launch_bounds__(256) k_vds_key(const double* __restrict__ xyz, uint32_t n, VdsGeom g, unsigned long long* __restrict__ lin,
                                                 uint32_t* __restrict__ key_lo, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
    unsigned long long k = g.invalid;
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
        const unsigned long long cx = (unsigned long long)(long long)floor((x - g.vmin[0]) / g.voxel);
        const unsigned long long cy = (unsigned long long)(long long)floor((y - g.vmin[1]) / g.voxel);
        const unsigned long long cz = (unsigned long long)(long long)floor((z - g.vmin[2]) / g.voxel);
        k = (cz * g.ny + cy) * g.nx + cx;
    }
    lin[i] = k; key_lo[i] = (uint32_t)k; vals[i] = i;
}
__global__ void __launch_bounds__(256) k_vds_key_hi(const unsigned long long* __restrict__ lin, const uint32_t* __restrict__ order, uint32_t n,
                                                    uint32_t* __restrict__ key_hi) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) key_hi[i] = (uint32_t)(lin[order[i]] >> 32);
}
// flags[i] = 1 where a voxel run starts in sorted order (the non-finite bucket never starts one); flags[n] = 0
__global__ void __launch_bounds__(256) k_vds_heads(const unsigned long long* __restrict__ lin, const uint32_t* __restrict__ order, uint32_t n,
                                                   unsigned long long invalid, uint32_t* __restrict__ flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint32_t f = 0;
    if (i < n) { const unsigned long long k = lin[order[i]]; f = (k != invalid) && (i == 0 || lin[order[i - 1]] != k); }
    flags[i] = f;
}
// seg_start[run] = first sorted position of the run; seg_start[nruns] = end of the finite points
__global__ void __launch_bounds__(256) k_vds_seg_start(const unsigned long long* __restrict__ lin, const uint32_t* __restrict__ order,
                                                       const uint32_t* __restrict__ segid, uint32_t n, unsigned long long invalid,
                                                       uint32_t* __restrict__ seg_start, int32_t* __restrict__ rank_of_point) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { if (lin[order[n - 1]] != invalid) seg_start[segid[n]] = n; return; }
    const unsigned long long k = lin[order[i]];
    const bool head = (i == 0) || (lin[order[i - 1]] != k);
    if (rank_of_point) rank_of_point[order[i]] = (k == invalid) ? -1 : (int32_t)(segid[i] + (head ? 1u : 0u)) - 1;
    if (!head) return;
    if (k == invalid) seg_start[segid[n]] = i;
    else seg_start[segid[i]] = i;
}
// one thread per voxel: double sums in ascending input index, then the mean
__global__ void __launch_bounds__(128) k_vds_mean(const double* __restrict__ xyz, const uint32_t* __restrict__ order,
                                                  const uint32_t* __restrict__ seg_start, uint32_t nseg, double* __restrict__ out) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const uint32_t b = seg_start[s], e = seg_start[s + 1];
    double sx = 0, sy = 0, sz = 0;
    for (uint32_t i = b; i < e; i++) {
        const size_t src = order[i];
        sx += xyz[3 * src]; sy += xyz[3 * src + 1]; sz += xyz[3 * src + 2];
    }
    const double cnt = (double)(e - b);
    out[3 * (size_t)s] = sx / cnt; out[3 * (size_t)s + 1] = sy / cnt; out[3 * (size_t)s + 2] = sz / cnt;
}

static int sort_by_u64(const unsigned long long* lin, size_t n, int bits, uint32_t* ka, uint32_t* va, uint32_t* kb, uint32_t* vb,
                       void* scratch, cudaStream_t s, uint32_t** order);

static int voxel_down_sample(b2_cloud_s* c, double voxel, b2_cloud_s* out, int32_t* h_rank) {
    cudaStream_t s = c->stream;
    const size_t n = c->n;
    out->n = 0; out->has_normals = false;
    if (n == 0) return B2_OK;
    if (n > 0x7fffffffull) { set_error("voxel_down_sample: too many points"); return B2_ERR_ARG; }
    const double* xyz = c->xyz.as<double>();
    double mn[3], mx[3];
    B2_CHECK(bbox_f64(xyz, n, c->work, s, mn, mx));
    if (!(mn[0] <= mx[0])) return B2_OK;
    VdsGeom g;
    g.voxel = voxel;
    double cnt[3];
    for (int d = 0; d < 3; d++) {
        g.vmin[d] = mn[d] - voxel * 0.5;
        cnt[d] = std::floor((mx[d] - g.vmin[d]) / voxel) + 1.0;
        if (!(cnt[d] < 2147483647.0)) { set_error("voxel_down_sample: voxel_size is too small"); return B2_ERR_TOO_LARGE; }
    }
    if (cnt[0] * cnt[1] * cnt[2] >= 4.0e18) { set_error("voxel_down_sample: voxel_size is too small"); return B2_ERR_TOO_LARGE; }
    g.nx = (unsigned long long)cnt[0]; g.ny = (unsigned long long)cnt[1];
    const unsigned long long total = g.nx * g.ny * (unsigned long long)cnt[2];
    g.invalid = total;
    int bits = 1;
    while (bits < 64 && (1ull << bits) <= total) bits++;
    const size_t nal = (n + 63) & ~(size_t)63;
    const size_t np1 = n + 1;
    const size_t need = nal * 8 + 4 * nal * 4 + (np1 + 64) * 4 * 2 + sort_tmp_bytes(n) + scan_tmp_bytes(np1) + 2048;
    B2_CHECK(c->work.reserve(need));
    unsigned long long* lin = c->work.as<unsigned long long>();
    uint32_t* ka = reinterpret_cast<uint32_t*>(lin + nal);
    uint32_t* va = ka + nal; uint32_t* kb = va + nal; uint32_t* vb = kb + nal;
    uint32_t* flags = vb + nal;
    uint32_t* seg_start = flags + ((np1 + 63) & ~(size_t)63);
    char* scratch = reinterpret_cast<char*>(seg_start + ((np1 + 63) & ~(size_t)63));
    const unsigned nblk = (unsigned)((n + 255) / 256), nblk1 = (unsigned)((np1 + 255) / 256);
    k_vds_key<<<nblk, 256, 0, s>>>(xyz, (uint32_t)n, g, lin, ka, va); count_launch();
    B2_CUDA(cudaGetLastError());
    uint32_t* vs = nullptr;
    B2_CHECK(sort_by_u64(lin, n, bits, ka, va, kb, vb, scratch, s, &vs));
    k_vds_heads<<<nblk1, 256, 0, s>>>(lin, vs, (uint32_t)n, g.invalid, flags); count_launch();
    B2_CUDA(cudaGetLastError());
    B2_CHECK(exclusive_scan_u32(flags, np1, scratch + sort_tmp_bytes(n), s));
    int32_t* d_rank = nullptr;
    DevBuf rankbuf;
    if (h_rank) { B2_CHECK(rankbuf.reserve(n * 4)); d_rank = rankbuf.as<int32_t>(); }
    k_vds_seg_start<<<nblk1, 256, 0, s>>>(lin, vs, flags, (uint32_t)n, g.invalid, seg_start, d_rank); count_launch();
    B2_CUDA(cudaGetLastError());
    uint32_t nseg = 0;
    B2_CUDA(cudaMemcpyAsync(&nseg, flags + n, 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    B2_CHECK(out->xyz.reserve((size_t)std::max<uint32_t>(nseg, 1) * 24));
    if (nseg) {
        k_vds_mean<<<(nseg + 127) / 128, 128, 0, s>>>(xyz, vs, seg_start, nseg, out->xyz.as<double>()); count_launch();
        B2_CUDA(cudaGetLastError());
    }
    if (h_rank) B2_CUDA(cudaMemcpyAsync(h_rank, d_rank, n * 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    rankbuf.release();
    out->n = nseg;
    return B2_OK;
}

// ------------------------------------------------------------------------------------------------ estimate_normals

// cyclic Jacobi of a symmetric 3x3 in double: w ascending, eigenvectors in the columns of V. Same sweep order, rotation
// formulas and stopping rule as the oracle, IEEE div/sqrt, no FMA.
__device__ __forceinline__ void jacobi3_f64(const double A[9], double w[3], double V[9]) {
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

__device__ __forceinline__ unsigned long long morton_expand21(unsigned long long v) {
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

__global__ void __launch_bounds__(256) k_morton_key(const double* __restrict__ xyz, uint32_t n, double ox, double oy, double oz, double scale,
                                                    unsigned long long* __restrict__ lin, uint32_t* __restrict__ key_lo, uint32_t* __restrict__ vals,
                                                    unsigned int* __restrict__ n_valid) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (i < n) {
        const double x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
        unsigned long long k = ~0ull;
        if (isfinite(x) && isfinite(y) && isfinite(z)) {
            const unsigned long long qx = (unsigned long long)fmin(2097151.0, fmax(0.0, (x - ox) * scale));
            const unsigned long long qy = (unsigned long long)fmin(2097151.0, fmax(0.0, (y - oy) * scale));
            const unsigned long long qz = (unsigned long long)fmin(2097151.0, fmax(0.0, (z - oz) * scale));
            k = morton_expand21(qx) | (morton_expand21(qy) << 1) | (morton_expand21(qz) << 2);
            ok = true;
        }
        lin[i] = k; key_lo[i] = (uint32_t)k; vals[i] = i;
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_valid, (unsigned)__popc(m));
}

// level 0: one warp per leaf of 32 points
__global__ void __launch_bounds__(256) k_bvh_leaf_boxes(const P4d* __restrict__ pts, uint32_t n, double* __restrict__ box, uint32_t n_leaf) {
    const uint32_t leaf = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (leaf >= n_leaf) return;
    const uint32_t p = leaf * 32u + lane;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (p < n) { double x, y, z; long long id; load_p4d(&pts[p], x, y, z, id); lo[0] = hi[0] = x; lo[1] = hi[1] = y; lo[2] = hi[2] = z; }
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo[d] = fmin(lo[d], shfl_xor_d(0xffffffffu, lo[d], o)); hi[d] = fmax(hi[d], shfl_xor_d(0xffffffffu, hi[d], o)); }
    if (lane < 3) box[6 * (size_t)leaf + lane] = lo[lane];
    else if (lane < 6) box[6 * (size_t)leaf + lane] = hi[lane - 3];
}
// level l: one warp per node, lane j reads child j
__global__ void __launch_bounds__(256) k_bvh_node_boxes(const double* __restrict__ child, uint32_t n_child, double* __restrict__ box, uint32_t n_node) {
    const uint32_t node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (node >= n_node) return;
    const uint32_t c = node * 32u + lane;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (c < n_child) {
#pragma unroll
        for (int d = 0; d < 3; d++) { lo[d] = child[6 * (size_t)c + d]; hi[d] = child[6 * (size_t)c + 3 + d]; }
    }
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo[d] = fmin(lo[d], shfl_xor_d(0xffffffffu, lo[d], o)); hi[d] = fmax(hi[d], shfl_xor_d(0xffffffffu, hi[d], o)); }
    if (lane < 3) box[6 * (size_t)node + lane] = lo[lane];
    else if (lane < 6) box[6 * (size_t)node + lane] = hi[lane - 3];
}

// One warp serves one leaf = 32 consecutive Morton-sorted points, all 32 queries at once (one lane per query).
//   search   the tree is walked ONCE per leaf with a group bound: a node is visited when its box is no farther from the
//            leaf's box than the largest current K-th distance among the 32 queries; children nearest first. Each leaf
//            reached is staged in shared memory and every lane scans its 32 points, inserting into its own sorted
//            K-list (shared memory, [slot][lane] so lanes never conflict). Box-to-box and point distances use the same
//            summation order, so rounding is monotone and the group pruning is exact for every query of the leaf.
//   normal   each lane: cumulants over its neighbours in ascending (distance, index), covariance, Jacobi, normal.
// Normals are written at the points' original indices.
constexpr int NRM_WARPS = 2;
constexpr int NRM_MAXK = 32;

__device__ __forceinline__ double box_box_dist2(const double* __restrict__ b, const double (&glo)[3], const double (&ghi)[3]) {
    const double2 a0 = __ldg(reinterpret_cast<const double2*>(b)), a1 = __ldg(reinterpret_cast<const double2*>(b) + 1),
                  a2 = __ldg(reinterpret_cast<const double2*>(b) + 2);
    const double dx = fmax(0.0, fmax(a0.x - ghi[0], glo[0] - a1.y));
    const double dy = fmax(0.0, fmax(a0.y - ghi[1], glo[1] - a2.x));
    const double dz = fmax(0.0, fmax(a1.x - ghi[2], glo[2] - a2.y));
    return dx * dx + dy * dy + dz * dz;
}

// leaf_begin / leaf_end: the leaves (32 consecutive points of the Morton order) this launch serves; morton_out: write normal i of
// the Morton order to nrm[3 i] instead of nrm[3 * original index] (the sharded set-up all-gathers contiguous leaf ranges).
__global__ void __launch_bounds__(NRM_WARPS * 32) k_normals(const __grid_constant__ BvhDev T, int K, double* __restrict__ nrm, uint32_t leaf_begin, uint32_t leaf_end, int morton_out) {
    __shared__ uint2 s_h[NRM_WARPS][NRM_MAXK][32];      // heap entry: (fp32 image of the squared distance, position in T.pts)
    __shared__ double s_cx[NRM_WARPS][32], s_cy[NRM_WARPS][32], s_cz[NRM_WARPS][32];
    __shared__ int s_ci[NRM_WARPS][32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t leaf0 = leaf_begin + blockIdx.x * NRM_WARPS + warp;
    if (leaf0 >= leaf_end) return;
    const uint32_t me = leaf0 * 32u + lane;
    const bool valid = me < T.n;
    double qx = 0, qy = 0, qz = 0; long long mid = -1;
    if (valid) load_p4d(&T.pts[me], qx, qy, qz, mid);
    double glo[3], ghi[3];
#pragma unroll
    for (int d = 0; d < 3; d++) { glo[d] = T.box[0][6 * (size_t)leaf0 + d]; ghi[d] = T.box[0][6 * (size_t)leaf0 + 3 + d]; }
    int cnt = 0;
    uint32_t kkey = 0x7f800000u;                        // key of this lane's K-th best so far (+inf until the heap is full)
    double kd = INFINITY;                               // an upper bound of its squared distance (the next float up), for the box tests

    // The K best of a lane live in a max-heap in shared memory ([slot][lane]: a lane only ever touches its own bank), so an
    // insertion costs at most log2(K) steps for every lane; a sorted list made the warp pay the longest shift of its 32 lanes
    // on every candidate (measured: 90 % of the kernel's instructions).
    // The heap is ordered by (squared distance, original index) exactly, but an entry stores only the fp32 image of the
    // distance next to the position: rounding to float is monotone, so two different keys order their entries the way the
    // doubles do, and EQUAL keys (a few in ten million comparisons) are settled by recomputing both distances from the point
    // records — the same expression, the same bits — and then by the index. 8 bytes per slot instead of 12 (shared memory is
    // what limits the occupancy of this kernel: 24 warps per SM instead of 16), one 64-bit load and one integer compare per
    // step instead of a double load, a double compare with a tie branch and a second array (the sifts were 73 % of the launch).
    auto key_of = [](double d) { return __float_as_uint(__double2float_rn(d)); };       // d >= 0: the bit pattern orders like the value
    auto exact_at = [&](int slot, double& d, int& idx) {
        double x, y, z; long long id;
        load_p4d(&T.pts[s_h[warp][slot][lane].y], x, y, z, id);
        const double dx = qx - x, dy = qy - y, dz = qz - z;
        d = dx * dx + dy * dy + dz * dz; idx = (int)id;
    };
    // (d, idx) of a new element against heap slot `slot`: is the new one smaller?
    auto new_less = [&](uint32_t key, double d, int idx, int slot, uint32_t skey) {
        if (key != skey) return key < skey;
        double sd; int si; exact_at(slot, sd, si);
        return d < sd || (d == sd && idx < si);
    };
    auto slot_less_new = [&](int slot, uint32_t skey, uint32_t key, double d, int idx) {     // heap slot smaller than the new element?
        if (key != skey) return skey < key;
        double sd; int si; exact_at(slot, sd, si);
        return sd < d || (sd == d && si < idx);
    };
    auto slot_less = [&](int a, uint32_t ka, int b, uint32_t kb) {     // heap slot a smaller than heap slot b?
        if (ka != kb) return ka < kb;
        double da, db; int ia, ib; exact_at(a, da, ia); exact_at(b, db, ib);
        return da < db || (da == db && ia < ib);
    };
    auto sift_down = [&](int j, int n, uint32_t key, double d, int idx, uint32_t pos) {
        for (;;) {
            int c = 2 * j + 1;
            if (c >= n) break;
            uint2 ce = s_h[warp][c][lane];
            if (c + 1 < n) {
                const uint2 re = s_h[warp][c + 1][lane];
                if (slot_less(c, ce.x, c + 1, re.x)) { c++; ce = re; }
            }
            if (!new_less(key, d, idx, c, ce.x)) break;
            s_h[warp][j][lane] = ce;
            j = c;
        }
        s_h[warp][j][lane] = make_uint2(key, pos);
    };
    auto set_bound = [&]() {
        kkey = s_h[warp][0][lane].x;
        kd = kkey >= 0x7f800000u ? INFINITY : (double)__uint_as_float(kkey + 1u);
    };
    // One candidate into this lane's heap (fill, or replace the K-th best when it is smaller).
    auto offer = [&](uint32_t key, double d, int idx, uint32_t pos) {
        if (cnt < K) {                                   // still filling: sift up
            int j = cnt++;
            while (j > 0) {
                const int pj = (j - 1) >> 1;
                const uint2 pe = s_h[warp][pj][lane];
                if (!slot_less_new(pj, pe.x, key, d, idx)) break;     // the parent moves down while it is smaller than the new element
                s_h[warp][j][lane] = pe;
                j = pj;
            }
            s_h[warp][j][lane] = make_uint2(key, pos);
            if (cnt == K) set_bound();
        } else if (key <= kkey && new_less(key, d, idx, 0, kkey)) {    // replaces the current K-th best (the root): sift down
            sift_down(0, K, key, d, idx, pos);
            set_bound();
        }
    };
    // (Measured and dropped, 10 M points, 77.9 ms with the loop below: marking a leaf's candidates first and inserting them
    // afterwards, so that the warp runs as many sifts as its busiest lane needs — 86 ms, the busiest lane still has 13 marks per
    // leaf; parking the candidates in a per-lane queue and draining all lanes together — 94 / 97 / 118 ms for queues of
    // 8 / 16 / 32: the bound a lane prunes with goes stale, and the extra candidates and leaves cost more than the shared sifts save.)
    auto scan_leaf = [&](uint32_t leaf, bool mine) {
        const uint32_t p = leaf * 32u + lane;
        __syncwarp();
        if (p < T.n) { double x, y, z; long long id; load_p4d(&T.pts[p], x, y, z, id); s_cx[warp][lane] = x; s_cy[warp][lane] = y; s_cz[warp][lane] = z; s_ci[warp][lane] = (int)id; }
        __syncwarp();
        const int nc = (int)min(32u, T.n - leaf * 32u);
        if (mine) {
            for (int c = 0; c < nc; c++) {
                const double dx = qx - s_cx[warp][c], dy = qy - s_cy[warp][c], dz = qz - s_cz[warp][c];
                const double d = dx * dx + dy * dy + dz * dz;
                const uint32_t key = key_of(d);
                if (cnt < K || key <= kkey) offer(key, d, s_ci[warp][c], leaf * 32u + c);      // equal keys are decided exactly inside
            }
        }
    };
    auto group_bound = [&]() {
        double b = valid ? kd : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b = fmax(b, shfl_xor_d(full, b, o));
        return b;
    };

    const int top = T.levels - 1;
    if (top == 0) scan_leaf(0, valid);
    else {
        double dch[BVH_MAXL];
        uint32_t node[BVH_MAXL];
        int lv = top;
        node[lv] = 0;
        auto expand = [&](int l, uint32_t nd) {
            const uint32_t c = nd * 32u + lane;
            dch[l] = (c < T.count[l - 1]) ? box_box_dist2(T.box[l - 1] + 6 * (size_t)c, glo, ghi) : INFINITY;
        };
        expand(lv, 0);
        double bound = INFINITY;
        for (;;) {
            double dmin = dch[lv]; int jmin = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double od = shfl_xor_d(full, dmin, o);
                const int oj = __shfl_xor_sync(full, jmin, o);
                if (od < dmin || (od == dmin && oj < jmin)) { dmin = od; jmin = oj; }
            }
            if (dmin == INFINITY || dmin > bound) {
                if (++lv > top) break;
                continue;
            }
            if (lane == jmin) dch[lv] = INFINITY;
            const uint32_t child = node[lv] * 32u + (uint32_t)jmin;
            // exact test, one lane per query: a leaf that straddles a jump of the Morton curve has a box spanning both
            // sides, and the box-to-box bound alone would send the whole warp through everything in between
            const double dq = box_dist2(T.box[lv - 1] + 6 * (size_t)child, qx, qy, qz);
            const bool want = valid && (cnt < K || dq <= kd);
            if (!__any_sync(full, want)) continue;
            if (lv == 1) { scan_leaf(child, want); bound = group_bound(); }
            else { lv--; node[lv] = child; expand(lv, child); }
        }
    }
    if (!valid) return;
    const int found = cnt;
    // heap sort in place: the cumulants below are summed in ascending (distance, index), the order the oracle pins
    for (int end = found - 1; end > 0; end--) {
        const uint2 e = s_h[warp][end][lane];
        double d; int idx; exact_at(end, d, idx);
        s_h[warp][end][lane] = s_h[warp][0][lane];
        sift_down(0, end, e.x, d, idx, e.y);
    }
    double nv[3] = {0.0, 0.0, 1.0};
    if (found >= 3) {
        double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = 0; j < found; j++) {
            double x, y, z; long long id;
            load_p4d(&T.pts[s_h[warp][j][lane].y], x, y, z, id);
            c[0] += x; c[1] += y; c[2] += z;
            c[3] += x * x; c[4] += x * y; c[5] += x * z;
            c[6] += y * y; c[7] += y * z; c[8] += z * z;
        }
        const double fn = (double)found;
#pragma unroll
        for (int q = 0; q < 9; q++) c[q] /= fn;
        double C[9];
        C[0] = c[3] - c[0] * c[0]; C[4] = c[6] - c[1] * c[1]; C[8] = c[8] - c[2] * c[2];
        C[1] = C[3] = c[4] - c[0] * c[1]; C[2] = C[6] = c[5] - c[0] * c[2]; C[5] = C[7] = c[7] - c[1] * c[2];
        double w[3], V[9];
        jacobi3_f64(C, w, V);
        nv[0] = V[0]; nv[1] = V[3]; nv[2] = V[6];
        const double nn = sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
        if (nn == 0.0) { nv[0] = 0; nv[1] = 0; nv[2] = 1; }
        if (nv[2] < 0 || (nv[2] == 0 && (nv[1] < 0 || (nv[1] == 0 && nv[0] < 0)))) { nv[0] = -nv[0]; nv[1] = -nv[1]; nv[2] = -nv[2]; }
    }
    double* o = &nrm[3 * (size_t)(morton_out ? (long long)me : mid)];
    o[0] = nv[0]; o[1] = nv[1]; o[2] = nv[2];
}

// normals gathered in Morton order -> the cloud's own order
__global__ void __launch_bounds__(256) k_normals_unsort(const P4d* __restrict__ pts, uint32_t n, const double* __restrict__ nrm_m, double* __restrict__ nrm) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y, z; long long id;
    load_p4d(&pts[i], x, y, z, id);
    nrm[3 * (size_t)id] = nrm_m[3 * (size_t)i]; nrm[3 * (size_t)id + 1] = nrm_m[3 * (size_t)i + 1]; nrm[3 * (size_t)id + 2] = nrm_m[3 * (size_t)i + 2];
}

__global__ void __launch_bounds__(256) k_fill_normals(double* __restrict__ nrm, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { nrm[3 * (size_t)i] = 0.0; nrm[3 * (size_t)i + 1] = 0.0; nrm[3 * (size_t)i + 2] = 1.0; }
}

// stable sort of (64-bit key, input index) by the low `bits` bits: one or two 32-bit LSD radix sorts.
// lin: keys in input order; ka = low words, va = 0..n-1 on entry. Returns the sorted index sequence.
static int sort_by_u64(const unsigned long long* lin, size_t n, int bits, uint32_t* ka, uint32_t* va, uint32_t* kb, uint32_t* vb,
                       void* scratch, cudaStream_t s, uint32_t** order) {
    uint32_t *ks, *vs;
    B2_CHECK(radix_sort_pairs(ka, va, kb, vb, n, std::min(bits, 32), scratch, s, &ks, &vs));
    if (bits > 32) {
        uint32_t* k2 = (ks == ka) ? kb : ka;
        uint32_t* v2 = (vs == va) ? vb : va;
        k_vds_key_hi<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(lin, vs, (uint32_t)n, ks); count_launch();
        B2_CUDA(cudaGetLastError());
        B2_CHECK(radix_sort_pairs(ks, vs, k2, v2, n, bits - 32, scratch, s, &ks, &vs));
    }
    *order = vs;
    return B2_OK;
}

int BvhIndex::build(const double* xyz, size_t n, DevBuf& work, cudaStream_t s) {
    dev = BvhDev{};
    if (n == 0) return B2_OK;
    if (n > 0x7fffffffull) { set_error("bvh: too many points"); return B2_ERR_ARG; }
    double mn[3], mx[3];
    B2_CHECK(bbox_f64(xyz, n, work, s, mn, mx));
    if (!(mn[0] <= mx[0])) return B2_OK;
    const double emax = std::max(mx[0] - mn[0], std::max(mx[1] - mn[1], mx[2] - mn[2]));
    const double scale = emax > 0 ? 2097151.0 / emax : 0.0;
    const unsigned nblk = (unsigned)((n + 255) / 256);
    size_t box_doubles = 0;
    for (uint32_t m = (uint32_t)((n + 31) / 32);; m = (m + 31) / 32) { box_doubles += 6 * (size_t)m; if (m == 1) break; }
    const size_t nal = (n + 63) & ~(size_t)63;
    const size_t need = nal * 8 + 4 * nal * 4 + sort_tmp_bytes(n) + 1024;
    B2_CHECK(work.reserve(need));
    B2_CHECK(pts.reserve(n * sizeof(P4d)));
    B2_CHECK(boxes.reserve(box_doubles * 8 + 64));
    unsigned long long* lin = work.as<unsigned long long>();
    uint32_t* ka = reinterpret_cast<uint32_t*>(lin + nal);
    uint32_t* va = ka + nal; uint32_t* kb = va + nal; uint32_t* vb = kb + nal;
    char* scratch = reinterpret_cast<char*>(vb + nal);
    unsigned int* d_valid = reinterpret_cast<unsigned int*>(scratch + sort_tmp_bytes(n));
    B2_CUDA(cudaMemsetAsync(d_valid, 0, 4, s));
    k_morton_key<<<nblk, 256, 0, s>>>(xyz, (uint32_t)n, mn[0], mn[1], mn[2], scale, lin, ka, va, d_valid); count_launch();
    B2_CUDA(cudaGetLastError());
    uint32_t* order = nullptr;
    B2_CHECK(sort_by_u64(lin, n, 64, ka, va, kb, vb, scratch, s, &order));
    k_celld_gather<<<nblk, 256, 0, s>>>(xyz, (uint32_t)n, order, pts.as<P4d>()); count_launch();
    uint32_t n_valid = 0;
    B2_CUDA(cudaMemcpyAsync(&n_valid, d_valid, 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    if (!n_valid) return B2_OK;
    dev.pts = pts.as<P4d>(); dev.n = n_valid; dev.levels = 0;
    double* bp = boxes.as<double>();
    for (uint32_t m = (n_valid + 31) / 32;; m = (m + 31) / 32) {
        dev.count[dev.levels] = m; dev.box[dev.levels] = bp; bp += 6 * (size_t)m; dev.levels++;
        if (m == 1 || dev.levels == BVH_MAXL) break;
    }
    for (int l = dev.levels; l < BVH_MAXL; l++) { dev.count[l] = 0; dev.box[l] = nullptr; }
    k_bvh_leaf_boxes<<<(dev.count[0] + 7) / 8, 256, 0, s>>>(dev.pts, n_valid, const_cast<double*>(dev.box[0]), dev.count[0]); count_launch();
    for (int l = 1; l < dev.levels; l++) {
        k_bvh_node_boxes<<<(dev.count[l] + 7) / 8, 256, 0, s>>>(dev.box[l - 1], dev.count[l - 1], const_cast<double*>(dev.box[l]), dev.count[l]); count_launch();
    }
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

int estimate_normals_knn(b2_cloud_s* c, int knn, b2_comm_s* comm) {
    if (knn < 1 || knn > NRM_MAXK) { set_error("estimate_normals: knn must be in [1, %d]", NRM_MAXK); return B2_ERR_ARG; }
    cudaStream_t s = c->stream;
    c->has_normals = false;
    if (c->n == 0) { c->has_normals = true; return B2_OK; }
    const size_t n = c->n;
    B2_CHECK(c->nrm.reserve(n * 24));
    k_fill_normals<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(c->nrm.as<double>(), (uint32_t)n); count_launch();
    BvhIndex& bvh = c->bvh;              // (a previous call's kernel has finished: bvh.build synchronises the stream for its bounding box)
    DevBuf& gathered = c->gathered;
    int st = bvh.build(c->xyz.as<double>(), n, c->work, s);
    cudaError_t e = cudaSuccess;
    if (st == B2_OK && bvh.dev.n && comm && comm->world > 1) {
        // Sharded set-up (config C5): every rank holds the whole cloud and the same BVH (the neighbourhood of a point may reach
        // anywhere), but serves only its contiguous range of leaves; the ranges are exchanged with one all-gather over NVLink
        // (24 B per point) and put back into the cloud's order. The 30-NN search is ~95 % of estimate_normals.
        const uint32_t L = bvh.dev.count[0];
        const uint32_t per = (L + (uint32_t)comm->world - 1) / (uint32_t)comm->world;
        const uint32_t lb = std::min(L, per * (uint32_t)comm->rank), le = std::min(L, lb + per);
        st = gathered.reserve((size_t)per * comm->world * 32 * 24);
        double* gm = gathered.as<double>();
        if (st == B2_OK && le > lb) {
            k_normals<<<(le - lb + NRM_WARPS - 1) / NRM_WARPS, NRM_WARPS * 32, 0, s>>>(bvh.dev, knn, gm, lb, le, 1); count_launch();
            e = cudaGetLastError();
        }
        if (st == B2_OK && e == cudaSuccess) st = comm_allgather_f64(comm, gm + (size_t)per * comm->rank * 96, gm, (size_t)per * 96, s);
        if (st == B2_OK && e == cudaSuccess) {
            k_normals_unsort<<<(bvh.dev.n + 255) / 256, 256, 0, s>>>(bvh.dev.pts, bvh.dev.n, gm, c->nrm.as<double>()); count_launch();
            e = cudaGetLastError();
        }
    } else if (st == B2_OK && bvh.dev.n) {
        k_normals<<<(bvh.dev.count[0] + NRM_WARPS - 1) / NRM_WARPS, NRM_WARPS * 32, 0, s>>>(bvh.dev, knn, c->nrm.as<double>(), 0u, bvh.dev.count[0], 0); count_launch();
        e = cudaGetLastError();
    }
    // no synchronisation here: the search is left running (see b2_cloud_s); a fault shows up at the consumer's synchronisation
    if (st != B2_OK) return st;
    if (e != cudaSuccess) { set_error("estimate_normals: %s", cudaGetErrorString(e)); return B2_ERR_CUDA; }
    c->has_normals = true;
    return B2_OK;
}

// ------------------------------------------------------------------------------------------------ transform / convert
struct Rigid { double R[9]; double t[3]; };
__global__ void __launch_bounds__(256) k_transform_d(double* __restrict__ xyz, double* __restrict__ nrm, uint32_t n, Rigid T) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* p = &xyz[3 * (size_t)i];
    const double x = p[0], y = p[1], z = p[2];
    p[0] = T.R[0] * x + T.R[1] * y + T.R[2] * z + T.t[0];
    p[1] = T.R[3] * x + T.R[4] * y + T.R[5] * z + T.t[1];
    p[2] = T.R[6] * x + T.R[7] * y + T.R[8] * z + T.t[2];
    if (nrm) {
        double* q = &nrm[3 * (size_t)i];
        const double a = q[0], b = q[1], c = q[2];
        q[0] = T.R[0] * a + T.R[1] * b + T.R[2] * c;
        q[1] = T.R[3] * a + T.R[4] * b + T.R[5] * c;
        q[2] = T.R[6] * a + T.R[7] * b + T.R[8] * c;
    }
}
__global__ void __launch_bounds__(256) k_f32_to_f64(const unsigned char* __restrict__ raw, size_t stride, uint32_t n, double* __restrict__ xyz) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    xyz[3 * (size_t)i] = (double)p[0]; xyz[3 * (size_t)i + 1] = (double)p[1]; xyz[3 * (size_t)i + 2] = (double)p[2];
}

}  // namespace b2

using namespace b2;

extern "C" {

int b2_cloud_create(b2_cloud_t* out) {
    if (!out) return B2_ERR_ARG;
    *out = nullptr;
    b2_cloud_s* c = new b2_cloud_s();
    if (cudaGetDevice(&c->device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("b2_cloud_create: %s", cudaGetErrorString(cudaGetLastError()));
        delete c; return B2_ERR_CUDA;
    }
    *out = c;
    return B2_OK;
}

int b2_cloud_destroy(b2_cloud_t c) {
    if (!c) return B2_OK;
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->xyz.release(); c->nrm.release(); c->work.release(); c->pin.release(); c->bvh.release(); c->gathered.release();
    if (c->ev0) { cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return B2_OK;
}

int b2_cloud_set_points(b2_cloud_t c, const double* xyz, size_t n) {
    B2_NVTX("b2_cloud_set_points");
    if (!c || (n && !xyz)) return B2_ERR_ARG;
    c->n = n; c->has_normals = false;
    if (!n) return B2_OK;
    B2_CHECK(c->xyz.reserve(n * 24));
    B2_CUDA(cudaMemcpyAsync(c->xyz.p, xyz, n * 24, cudaMemcpyHostToDevice, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return B2_OK;
}

int b2_cloud_set_points_f32(b2_cloud_t c, const void* base, size_t stride, size_t n) {
    if (!c || (n && !base) || stride < 12 || (stride & 3)) return B2_ERR_ARG;
    c->n = n; c->has_normals = false;
    if (!n) return B2_OK;
    if (n > 0x7fffffffull) return B2_ERR_ARG;
    B2_CHECK(c->xyz.reserve(n * 24));
    B2_CHECK(c->work.reserve(n * stride));
    B2_CUDA(cudaMemcpyAsync(c->work.p, base, n * stride, cudaMemcpyHostToDevice, c->stream));
    k_f32_to_f64<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->work.as<unsigned char>(), stride, (uint32_t)n, c->xyz.as<double>()); count_launch();
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return B2_OK;
}

int b2_cloud_size(b2_cloud_t c, size_t* n, int* has_normals) {
    if (!c) return B2_ERR_ARG;
    if (n) *n = c->n;
    if (has_normals) *has_normals = c->has_normals ? 1 : 0;
    return B2_OK;
}

int b2_cloud_get_points(b2_cloud_t c, double* xyz) {
    if (!c || (c->n && !xyz)) return B2_ERR_ARG;
    if (!c->n) return B2_OK;
    B2_CUDA(cudaMemcpyAsync(xyz, c->xyz.p, c->n * 24, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return B2_OK;
}

int b2_cloud_get_normals(b2_cloud_t c, double* nrm) {
    if (!c || (c->n && !nrm)) return B2_ERR_ARG;
    if (!c->has_normals) { set_error("b2_cloud_get_normals: the cloud has no normals"); return B2_ERR_STATE; }
    if (!c->n) return B2_OK;
    B2_CUDA(cudaMemcpyAsync(nrm, c->nrm.p, c->n * 24, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return B2_OK;
}

int b2_cloud_set_normals(b2_cloud_t c, const double* nrm) {
    if (!c || (c->n && !nrm)) return B2_ERR_ARG;
    if (c->n) {
        B2_CHECK(c->nrm.reserve(c->n * 24));
        B2_CUDA(cudaMemcpyAsync(c->nrm.p, nrm, c->n * 24, cudaMemcpyHostToDevice, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
    }
    c->has_normals = true;
    return B2_OK;
}

int b2_cloud_voxel_down_sample(b2_cloud_t c, double voxel_size, b2_cloud_t* out, int32_t* voxel_rank_of_point) {
    B2_NVTX("b2_cloud_voxel_down_sample");
    if (!c || !out) return B2_ERR_ARG;
    *out = nullptr;
    if (!(voxel_size > 0.0)) { set_error("voxel_down_sample: voxel_size <= 0"); return B2_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    b2_cloud_t o = nullptr;
    B2_CHECK(b2_cloud_create(&o));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, c->stream);
    const int st = voxel_down_sample(c, voxel_size, o, voxel_rank_of_point);
    cudaEventRecord(e1, c->stream); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&c->last_ms, e0, e1);
    c->ms_pending = false;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (st != B2_OK) { b2_cloud_destroy(o); return st; }
    *out = o;
    return B2_OK;
}

int b2_cloud_estimate_normals(b2_cloud_t c, int knn) {
    B2_NVTX("b2_cloud_estimate_normals");
    if (!c) return B2_ERR_ARG;
    B2_CUDA(cudaSetDevice(c->device));
    if (!c->ev0) { B2_CUDA(cudaEventCreate(&c->ev0)); B2_CUDA(cudaEventCreate(&c->ev1)); }
    B2_CUDA(cudaEventRecord(c->ev0, c->stream));
    const int st = estimate_normals_knn(c, knn, nullptr);
    B2_CUDA(cudaEventRecord(c->ev1, c->stream));
    c->ms_pending = true;                  // b2_cloud_last_gpu_ms waits for the events; this call does not
    return st;
}

// The same with the work split over the ranks of `comm` (every rank calls it on an identical cloud): each rank estimates the
// normals of 1/world of the points, one NCCL all-gather hands everybody all of them. Results are bit-identical to the
// single-GPU call (same kernel, same BVH).
int b2_cloud_estimate_normals_sharded(b2_cloud_t c, int knn, b2_comm_t comm) {
    B2_NVTX("b2_cloud_estimate_normals_sharded");
    if (!c) return B2_ERR_ARG;
    B2_CUDA(cudaSetDevice(c->device));
    if (!c->ev0) { B2_CUDA(cudaEventCreate(&c->ev0)); B2_CUDA(cudaEventCreate(&c->ev1)); }
    B2_CUDA(cudaEventRecord(c->ev0, c->stream));
    const int st = estimate_normals_knn(c, knn, comm);
    B2_CUDA(cudaEventRecord(c->ev1, c->stream));
    c->ms_pending = true;                  // b2_cloud_last_gpu_ms waits for the events; this call does not
    return st;
}

// b2_cloud_set_points when every rank of `comm` holds the same host array: each rank uploads 1/world of the rows over its own
// PCIe link and the slices are all-gathered over NVLink (world x less host-to-device traffic per rank).
int b2_cloud_set_points_sharded(b2_cloud_t c, const double* xyz, size_t n, b2_comm_t comm) {
    B2_NVTX("b2_cloud_set_points_sharded");
    if (!comm || comm->world <= 1) return b2_cloud_set_points(c, xyz, n);
    if (!c || (n && !xyz)) return B2_ERR_ARG;
    B2_CUDA(cudaSetDevice(c->device));
    c->n = n; c->has_normals = false;
    if (!n) return B2_OK;
    const size_t per = (n + comm->world - 1) / comm->world;
    B2_CHECK(c->xyz.reserve(per * comm->world * 24));
    const size_t b = std::min(n, per * (size_t)comm->rank), e = std::min(n, b + per);
    double* d = c->xyz.as<double>();
    if (e > b) B2_CUDA(cudaMemcpyAsync(d + 3 * b, xyz + 3 * b, (e - b) * 24, cudaMemcpyHostToDevice, c->stream));
    B2_CHECK(comm_allgather_f64(comm, d + 3 * per * comm->rank, d, 3 * per, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return B2_OK;
}

int b2_cloud_transform(b2_cloud_t c, const double T[16]) {
    if (!c || !T) return B2_ERR_ARG;
    if (!c->n) return B2_OK;
    Rigid r;
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) r.R[i * 3 + j] = T[i * 4 + j]; r.t[i] = T[i * 4 + 3]; }
    k_transform_d<<<(unsigned)((c->n + 255) / 256), 256, 0, c->stream>>>(c->xyz.as<double>(), c->has_normals ? c->nrm.as<double>() : nullptr,
                                                                         (uint32_t)c->n, r); count_launch();
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return B2_OK;
}

int b2_cloud_last_gpu_ms(b2_cloud_t c, float* ms) {
    if (!c || !ms) return B2_ERR_ARG;
    if (c->ms_pending) {
        B2_CUDA(cudaEventSynchronize(c->ev1));
        B2_CUDA(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
        c->ms_pending = false;
    }
    *ms = c->last_ms;
    return B2_OK;
}

}  // extern "C"
