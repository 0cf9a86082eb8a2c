// VoxelGrid downsampling on the device. Replaces pcl::VoxelGrid<PointT>::filter at
//   liosam_ws/src/LIO-SAM/src/featureExtraction.cpp:233-234, mapOptmization.cpp:719,879,928,932,960,965,
//   Calibration_Tookit/multi_lidar/.../multi_lidar_calibrator.cpp:113-121, heading_ws/src/src/PointCloudProcessing.cpp:23-30.
//
// Pipeline on one stream: bbox reduce -> (host: PCL's leaf/extent arithmetic on 6 floats) -> voxel key per point
// (float mul + floor, no FMA: the index must be bit-exact) -> stable LSD radix sort of (key, input index) ->
// run heads -> exclusive scans -> one thread per voxel sums its run in ascending input index (the float sum order
// the oracle pins) -> centroids in ascending voxel index, already in the caller's stride.
#include "b2_common.cuh"
#include <cmath>
#include <climits>
#include <algorithm>

namespace b2 {

__global__ void k_vx_bbox_init(uint32_t* bb) {
    if (threadIdx.x < 3) bb[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) bb[threadIdx.x] = 0u;
}

// V4: 16-byte records at a 16-byte aligned base (PointXYZI as the local map and the Python mirror hold it): one 128-bit load per
// point, four points in flight per thread, instead of three dependent-looking scalar loads (the kernel was 30 us of a 190 us filter
// at a million points: latency, not bandwidth)
template <bool V4>
__global__ void __launch_bounds__(256) k_vx_bbox(const unsigned char* __restrict__ raw, size_t stride, size_t n, uint32_t* __restrict__ bb) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    auto take = [&](float x, float y, float z) {
        if (isfinite(x) && isfinite(y) && isfinite(z)) {
            mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
            mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
        }
    };
    const size_t step = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if constexpr (V4) {
        const float4* q = reinterpret_cast<const float4*>(raw);
        for (; i + 3 * step < n; i += 4 * step) {
            const float4 a = __ldg(q + i), b = __ldg(q + i + step), c = __ldg(q + i + 2 * step), d = __ldg(q + i + 3 * step);
            take(a.x, a.y, a.z); take(b.x, b.y, b.z); take(c.x, c.y, c.z); take(d.x, d.y, d.z);
        }
        for (; i < n; i += step) { const float4 a = __ldg(q + i); take(a.x, a.y, a.z); }
    } else {
        for (; i < n; i += step) {
            const float* p = reinterpret_cast<const float*>(raw + i * stride);
            take(p[0], p[1], p[2]);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    // six atomics per CTA, not per warp (thousands of warps queueing on six words: 43 us of a 266 us filter at a million points)
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

struct VoxelGeom { float inv[3]; int min_b[3]; int mul[3]; uint32_t invalid_key; };

__global__ void __launch_bounds__(256) k_vx_key(const unsigned char* __restrict__ raw, size_t stride, uint32_t n, VoxelGeom g,
                                                uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, int32_t* __restrict__ vop) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x, y, z;
    if (stride == 16 && (reinterpret_cast<uintptr_t>(raw) & 15u) == 0) {          // uniform: one 128-bit load
        const float4 v = __ldg(reinterpret_cast<const float4*>(raw) + i);
        x = v.x; y = v.y; z = v.z;
    } else {
        const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
        x = p[0]; y = p[1]; z = p[2];
    }
    uint32_t key = g.invalid_key;
    int32_t v = -1;
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
        int i0 = (int)(floorf(x * g.inv[0]) - (float)g.min_b[0]);
        int i1 = (int)(floorf(y * g.inv[1]) - (float)g.min_b[1]);
        int i2 = (int)(floorf(z * g.inv[2]) - (float)g.min_b[2]);
        v = i0 * g.mul[0] + i1 * g.mul[1] + i2 * g.mul[2];
        key = (uint32_t)v;
    }
    keys[i] = key; vals[i] = i;
    if (vop) vop[i] = v;
}

// flags[i] = 1 where a new voxel run starts (invalid keys never start a run); flags[n] = 0
__global__ void __launch_bounds__(256) k_vx_heads(const uint32_t* __restrict__ keys, uint32_t n, uint32_t invalid_key, uint32_t* __restrict__ flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint32_t f = 0;
    if (i < n) { uint32_t k = keys[i]; f = (k != invalid_key) && (i == 0 || keys[i - 1] != k); }
    flags[i] = f;
}

// after the scan: segid[i] = run index of element i (meaningful at heads), segid[n] = number of runs.
// seg_start[run] = first element of the run; seg_start[nruns] = end of the valid (finite) elements.
__global__ void __launch_bounds__(256) k_vx_seg_start(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ segid, uint32_t n,
                                                      uint32_t invalid_key, uint32_t* __restrict__ seg_start) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { if (keys[n - 1] != invalid_key) seg_start[segid[n]] = n; return; }
    const uint32_t k = keys[i];
    const bool head = (i == 0) || (keys[i - 1] != k);
    if (!head) return;
    if (k == invalid_key) seg_start[segid[n]] = i;
    else seg_start[segid[i]] = i;
}

// keep[s] = 1 when run s holds at least min_pts points (setMinimumPointsNumberPerVoxel); zero past the last run
__global__ void __launch_bounds__(256) k_vx_keep(const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ nseg_p, uint32_t n,
                                                 uint32_t min_pts, uint32_t* __restrict__ keep) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n) return;
    uint32_t f = 0;
    if (s < *nseg_p) f = (seg_start[s + 1] - seg_start[s]) >= min_pts ? 1u : 0u;
    keep[s] = f;
}

// one thread per voxel: float sums in ascending input index, then the mean, written as a full output record
__global__ void __launch_bounds__(128) k_vx_centroid(const unsigned char* __restrict__ raw, size_t stride, int ioff, int n_fields,
                                                     const uint32_t* __restrict__ order, const uint32_t* __restrict__ keys,
                                                     const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ nseg_p,
                                                     const uint32_t* __restrict__ keep_scan, uint32_t n, uint32_t out_cap,
                                                     unsigned char* __restrict__ out, size_t ostride, int ooff, int32_t* __restrict__ out_key) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || s >= *nseg_p) return;
    const uint32_t rank = keep_scan[s];
    if (keep_scan[s + 1] == rank || rank >= out_cap) return;
    const uint32_t b = seg_start[s], e = seg_start[s + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (uint32_t i = b; i < e; i++) {
        const unsigned char* p = raw + (size_t)order[i] * stride;
        const float* f = reinterpret_cast<const float*>(p);
        sx += f[0]; sy += f[1]; sz += f[2];
        if (n_fields == 4) si += *reinterpret_cast<const float*>(p + ioff);
    }
    const float cnt = (float)(e - b);
    float* of = reinterpret_cast<float*>(out + (size_t)rank * ostride);
    const int words = (int)(ostride >> 2);
    for (int w = 3; w < words; w++) of[w] = 0.f;
    of[0] = sx / cnt; of[1] = sy / cnt; of[2] = sz / cnt;
    if (words >= 4 && !(n_fields == 4 && ooff == 12)) of[3] = 1.0f;          // PCL's homogeneous pad word
    if (n_fields == 4) of[ooff >> 2] = si / cnt;
    if (out_key) out_key[rank] = (int32_t)keys[b];
}


}  // namespace b2

using namespace b2;

// a filter in flight between voxel_filter_dev_begin and voxel_filter_dev_end (stage: 1 box queued, 2 count queued, 3 finished)
struct VoxelJob {
    const unsigned char* d_in = nullptr; size_t in_stride = 0, n = 0, out_stride = 0, out_capacity = 0; int n_fields = 0; int32_t* d_vop = nullptr;
    int stage = 0, refused = 0; uint32_t m = 0;
};
struct b2_voxel_s {
    VoxelJob job;
    float leaf[3] = {0.f, 0.f, 0.f};
    unsigned min_pts = 0;
    cudaStream_t stream = nullptr;
    DevBuf raw, work, out, small, vop;
    int device = b2::current_device();      // the device the handle was created on
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // device span of the last filter, upload and download excluded (b2_voxel_last_gpu_ms)
    bool timed = false;
    PinBuf pin;
};

extern "C" {

int b2_voxel_create(b2_voxel_t* out) {
    if (!out) { set_error("b2_voxel_create: null out"); return B2_ERR_ARG; }
    b2_voxel_s* h = new b2_voxel_s();
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("b2_voxel_create: %s", cudaGetErrorString(e)); delete h; return B2_ERR_CUDA; }
    *out = h;
    return B2_OK;
}

int b2_voxel_destroy(b2_voxel_t h) {
    if (h && h->ev0) { cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); h->ev0 = h->ev1 = nullptr; }
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h) return B2_ERR_ARG;
    h->raw.release(); h->work.release(); h->out.release(); h->small.release(); h->vop.release(); h->pin.release();
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return B2_OK;
}

int b2_voxel_set_leaf_size(b2_voxel_t h, float lx, float ly, float lz) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !(lx > 0.f) || !(ly > 0.f) || !(lz > 0.f)) { set_error("b2_voxel_set_leaf_size: leaf must be > 0"); return B2_ERR_ARG; }
    h->leaf[0] = lx; h->leaf[1] = ly; h->leaf[2] = lz;
    return B2_OK;
}

int b2_voxel_set_min_points_per_voxel(b2_voxel_t h, unsigned min_points) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h) return B2_ERR_ARG;
    h->min_pts = min_points;
    return B2_OK;
}

}  // extern "C"

namespace b2 {

// The filter proper, device in -> device out (h->out; *m_out voxels). d_in must stay valid until the stream has drained.
// refused = 1: PCL's "leaf size is too small" case, the output is a copy of the input. d_vop (optional, device, n ints).
// The filter in three stages, so that a caller with two filters to run (the local map's corner and surf clouds, each on its own
// handle and stream) can interleave them from one host thread: begin = bounding box queued; middle = wait for the box, PCL's
// geometry on the host, everything up to the centroids queued; end = wait for the voxel count. voxel_filter_dev is the three in a row.
int voxel_filter_dev_begin(b2_voxel_s* h, const unsigned char* d_in, size_t in_stride, size_t n, int n_fields, size_t out_stride,
                           size_t out_capacity, int32_t* d_vop) {
    VoxelJob& j = h->job;
    j = VoxelJob{};
    j.d_in = d_in; j.in_stride = in_stride; j.n = n; j.n_fields = n_fields; j.out_stride = out_stride; j.out_capacity = out_capacity; j.d_vop = d_vop;
    if (n == 0) { j.stage = 3; return B2_OK; }
    cudaStream_t s = h->stream;
    B2_CHECK(h->small.reserve(256));
    uint32_t* bb = h->small.as<uint32_t>();
    k_vx_bbox_init<<<1, 32, 0, s>>>(bb); count_launch();
    const int nbb = (int)std::min<size_t>((n + 255) / 256, (size_t)device_sm_count() * 4);
    if (in_stride == 16 && (reinterpret_cast<uintptr_t>(d_in) & 15u) == 0) { k_vx_bbox<true><<<nbb, 256, 0, s>>>(d_in, in_stride, n, bb); count_launch(); }
    else { k_vx_bbox<false><<<nbb, 256, 0, s>>>(d_in, in_stride, n, bb); count_launch(); }
    B2_CUDA(cudaGetLastError());
    B2_CHECK(h->pin.reserve(64));
    B2_CUDA(cudaMemcpyAsync(h->pin.as<uint32_t>(), bb, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    j.stage = 1;
    return B2_OK;
}

int voxel_filter_dev_middle(b2_voxel_s* h) {
    VoxelJob& j = h->job;
    if (j.stage != 1) return B2_OK;
    const unsigned char* d_in = j.d_in; const size_t in_stride = j.in_stride, n = j.n, out_stride = j.out_stride, out_capacity = j.out_capacity;
    const int n_fields = j.n_fields; int32_t* d_vop = j.d_vop;
    cudaStream_t s = h->stream;
    const int ioff = (int)B2_INTENSITY_OFFSET(in_stride), ooff = (int)B2_INTENSITY_OFFSET(out_stride);
    uint32_t* hbb = h->pin.as<uint32_t>();
    B2_CUDA(cudaStreamSynchronize(s));
    j.stage = 3;
    auto unflip = [](uint32_t u) { uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; float f; memcpy(&f, &v, 4); return f; };
    float mn[3], mx[3];
    for (int d = 0; d < 3; d++) { mn[d] = unflip(hbb[d]); mx[d] = unflip(hbb[3 + d]); }
    if (!(mn[0] <= mx[0])) return B2_OK;            // no finite point at all
    // PCL applyFilter arithmetic on the host, in float exactly as written there
    VoxelGeom g;
    int64_t dxyz[3];
    int max_b[3], div_b[3];
    for (int d = 0; d < 3; d++) {
        g.inv[d] = 1.0f / h->leaf[d];
        dxyz[d] = (int64_t)((mx[d] - mn[d]) * g.inv[d]) + 1;
    }
    if (dxyz[0] * dxyz[1] * dxyz[2] > (int64_t)INT32_MAX) {
        // "Leaf size is too small for the input dataset": output = input
        j.refused = 1;
        if (out_capacity < n) { set_error("b2_voxel_filter: refused (index overflow) and out_capacity < n"); return B2_ERR_CAPACITY; }
        B2_CHECK(h->out.reserve(n * out_stride));
        if (in_stride == out_stride) B2_CUDA(cudaMemcpyAsync(h->out.p, d_in, n * in_stride, cudaMemcpyDeviceToDevice, s));
        else {
            B2_CUDA(cudaMemsetAsync(h->out.p, 0, n * out_stride, s));
            B2_CUDA(cudaMemcpy2DAsync(h->out.p, out_stride, d_in, in_stride, 12, n, cudaMemcpyDeviceToDevice, s));
            if (n_fields == 4) B2_CUDA(cudaMemcpy2DAsync(h->out.as<unsigned char>() + ooff, out_stride, d_in + ioff, in_stride, 4, n, cudaMemcpyDeviceToDevice, s));
        }
        if (d_vop) B2_CUDA(cudaMemsetAsync(d_vop, 0xff, n * sizeof(int32_t), s));
        j.m = (uint32_t)n;
        return B2_OK;
    }
    uint64_t ncells = 1;
    for (int d = 0; d < 3; d++) {
        g.min_b[d] = (int)std::floor(mn[d] * g.inv[d]);
        max_b[d] = (int)std::floor(mx[d] * g.inv[d]);
        div_b[d] = max_b[d] - g.min_b[d] + 1;
        ncells *= (uint64_t)div_b[d];
    }
    g.mul[0] = 1; g.mul[1] = div_b[0]; g.mul[2] = div_b[0] * div_b[1];
    if (ncells > 0xfffffffeull) ncells = 0xfffffffeull;
    g.invalid_key = (uint32_t)ncells;               // sorts after every real voxel
    int bits = 1;
    while (bits < 32 && ((uint64_t)1 << bits) <= ncells) bits++;

    const size_t nal = (n + 64) & ~(size_t)63;       // room for n+1 entries
    const size_t scratch_bytes = std::max(sort_tmp_bytes(n), scan_tmp_bytes(n + 1)) + 1024;
    B2_CHECK(h->work.reserve(7 * nal * sizeof(uint32_t) + scratch_bytes));
    uint32_t* ka = h->work.as<uint32_t>();
    uint32_t* va = ka + nal; uint32_t* kb = va + nal; uint32_t* vb = kb + nal;
    uint32_t* segid = vb + nal; uint32_t* seg_start = segid + nal; uint32_t* keep = seg_start + nal;
    char* scratch = reinterpret_cast<char*>(keep + nal);
    const uint32_t n32 = (uint32_t)n;
    const unsigned nblk = (unsigned)((n + 255) / 256), nblk1 = (unsigned)((n + 1 + 255) / 256);
    k_vx_key<<<nblk, 256, 0, s>>>(d_in, in_stride, n32, g, ka, va, d_vop); count_launch();
    B2_CUDA(cudaGetLastError());
    uint32_t *ks, *vs;
    B2_CHECK(radix_sort_pairs(ka, va, kb, vb, n, bits, scratch, s, &ks, &vs));
    k_vx_heads<<<nblk1, 256, 0, s>>>(ks, n32, g.invalid_key, segid); count_launch();
    B2_CUDA(cudaGetLastError());
    B2_CHECK(exclusive_scan_u32(segid, n + 1, scratch, s));
    k_vx_seg_start<<<nblk1, 256, 0, s>>>(ks, segid, n32, g.invalid_key, seg_start); count_launch();
    k_vx_keep<<<nblk1, 256, 0, s>>>(seg_start, segid + n, n32, h->min_pts, keep); count_launch();
    B2_CUDA(cudaGetLastError());
    B2_CHECK(exclusive_scan_u32(keep, n + 1, scratch, s));
    const size_t cap = std::min(n, out_capacity);
    B2_CHECK(h->out.reserve(std::max<size_t>(cap, 1) * out_stride));
    k_vx_centroid<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_in, in_stride, ioff, n_fields, vs, ks, seg_start, segid + n,
                                                              keep, n32, (uint32_t)cap, h->out.as<unsigned char>(), out_stride, ooff, nullptr); count_launch();
    B2_CUDA(cudaGetLastError());
    uint32_t* hm = h->pin.as<uint32_t>() + 8;
    B2_CUDA(cudaMemcpyAsync(hm, keep + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    j.stage = 2;
    return B2_OK;
}

int voxel_filter_dev_end(b2_voxel_s* h, uint32_t* m_out, int* refused) {
    VoxelJob& j = h->job;
    *m_out = 0;
    if (refused) *refused = j.refused;
    if (j.stage == 2) {
        B2_CUDA(cudaStreamSynchronize(h->stream));
        const uint32_t* hm = h->pin.as<uint32_t>() + 8;
        j.stage = 3;
        if ((size_t)*hm > j.out_capacity) { set_error("b2_voxel_filter: %u voxels but out_capacity %zu", *hm, j.out_capacity); return B2_ERR_CAPACITY; }
        j.m = *hm;
    }
    *m_out = j.m;
    return B2_OK;
}

int voxel_filter_dev(b2_voxel_s* h, const unsigned char* d_in, size_t in_stride, size_t n, int n_fields, size_t out_stride,
                     size_t out_capacity, uint32_t* m_out, int* refused, int32_t* d_vop) {
    *m_out = 0;
    if (refused) *refused = 0;
    B2_CHECK(voxel_filter_dev_begin(h, d_in, in_stride, n, n_fields, out_stride, out_capacity, d_vop));
    B2_CHECK(voxel_filter_dev_middle(h));
    return voxel_filter_dev_end(h, m_out, refused);
}

const void* voxel_out_dev(b2_voxel_s* h) { return h->out.p; }
cudaStream_t voxel_stream(b2_voxel_s* h) { return h->stream; }

// host cloud in, downsampled cloud left on the device (packed xyzi); the caller synchronises voxel_stream(h)
int voxel_filter_host_to_dev(b2_voxel_s* h, const void* in, size_t in_stride, size_t n, uint32_t* m_out) {
    *m_out = 0;
    if (!(h->leaf[0] > 0.f)) { set_error("VoxelGrid: leaf size not set"); return B2_ERR_STATE; }
    if (n == 0) return B2_OK;
    B2_CHECK(h->raw.reserve(n * in_stride));
    B2_CUDA(cudaMemcpyAsync(h->raw.p, in, n * in_stride, cudaMemcpyHostToDevice, h->stream));
    int refused = 0;
    return voxel_filter_dev(h, h->raw.as<unsigned char>(), in_stride, n, 4, 16, n, m_out, &refused, nullptr);
}

}  // namespace b2

extern "C" {

int b2_voxel_filter(b2_voxel_t h, const void* in, size_t in_stride, size_t n, int n_fields, void* out, size_t out_stride,
                    size_t out_capacity, size_t* n_out, int* refused, int32_t* voxel_of_point) {
    B2_NVTX("b2_voxel_filter");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !n_out || (n && (!in || !out)) || (n_fields != 3 && n_fields != 4) || in_stride < (size_t)(n_fields * 4) ||
        out_stride < (size_t)(n_fields * 4) || (in_stride & 3) || (out_stride & 3) || n > 0x7ffffff0ull) {
        set_error("b2_voxel_filter: bad argument"); return B2_ERR_ARG;
    }
    if (!(h->leaf[0] > 0.f)) { set_error("b2_voxel_filter: leaf size not set"); return B2_ERR_STATE; }
    *n_out = 0;
    if (refused) *refused = 0;
    if (n == 0) return B2_OK;                       // PCL: empty input -> empty output
    cudaStream_t s = h->stream;
    B2_CHECK(h->raw.reserve(n * in_stride));
    B2_CUDA(cudaMemcpyAsync(h->raw.p, in, n * in_stride, cudaMemcpyHostToDevice, s));
    int32_t* d_vop = nullptr;
    if (voxel_of_point) { B2_CHECK(h->vop.reserve(n * sizeof(int32_t))); d_vop = h->vop.as<int32_t>(); }
    uint32_t m = 0;
    if (!h->ev0) { B2_CUDA(cudaEventCreate(&h->ev0)); B2_CUDA(cudaEventCreate(&h->ev1)); }
    B2_CUDA(cudaEventRecord(h->ev0, s));
    B2_CHECK(voxel_filter_dev(h, h->raw.as<unsigned char>(), in_stride, n, n_fields, out_stride, out_capacity, &m, refused, d_vop));
    B2_CUDA(cudaEventRecord(h->ev1, s));
    h->timed = true;
    if (voxel_of_point) B2_CUDA(cudaMemcpyAsync(voxel_of_point, d_vop, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (m) B2_CUDA(cudaMemcpyAsync(out, h->out.p, (size_t)m * out_stride, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    *n_out = m;
    return B2_OK;
}

// device time of the last b2_voxel_filter between the end of the upload and the start of the download (CUDA events)
int b2_voxel_last_gpu_ms(b2_voxel_t h, float* ms) {
    if (!h || !ms) return B2_ERR_ARG;
    b2::DeviceScope device_scope_(h->device);
    if (!h->timed) { set_error("b2_voxel_last_gpu_ms: no filter has run"); return B2_ERR_STATE; }
    B2_CUDA(cudaEventSynchronize(h->ev1));
    B2_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return B2_OK;
}

}  // extern "C"
