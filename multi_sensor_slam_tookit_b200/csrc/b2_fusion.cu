// Multi-lidar fusion front end on the device. Replaces, in PointClouds_Fusion,
//   fusion_pointclouds/src/fusion_pointcloud/src/fusion_pointclouds.cpp
//     :62-73    pcl::transformPointCloud(*pc_local_k, *pc_trans_k, trans_cpck_to_ppc.matrix())   (Eigen::Isometry3d, double)
//     :80-89    pc_fusion_local = pc_local_1 (+ pc_trans_4) (+ pc_trans_3) + pc_trans_2           (concatenation order)
//     :93-108   passthroughFiter (external bounds: keep min <= v <= max on x, then z, then y; float limits) and
//               conditionFiter (internal bounds: pcl::ConditionOr of GT/LT comparisons — keep what lies OUTSIDE the box)
//   lidar_fusion/src/src/lidar_fusion.cpp:239-252 (per-point float transform of one cloud), :333-334 (cloud1 + cloud2)
// One kernel per input cloud writes the transformed points straight into the fused buffer at the cloud's offset; one
// kernel marks the survivors of both filters, a prefix sum gives their slots, one kernel compacts in order. The fused
// cloud stays in HBM for whoever consumes it next (b2_fusion_device_cloud) or is copied out in the caller's stride.
#include "b2_common.cuh"
#include <vector>
#include <algorithm>

namespace b2 {

struct FuseT { double m[12]; int identity; };
struct FuseBounds { float emin[3], emax[3]; double imin[3], imax[3]; int ext, inn; };

// pcl::transformPointCloud with a double matrix: sums in double, stored as float; the intensity rides along
__global__ void __launch_bounds__(256) k_fuse_transform(const unsigned char* __restrict__ raw, size_t stride, int ioff, uint32_t n, FuseT T, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char* p = raw + (size_t)i * stride;
    const float* f = reinterpret_cast<const float*>(p);
    float4 o;
    if (T.identity) { o.x = f[0]; o.y = f[1]; o.z = f[2]; }
    else {
        const double x = (double)f[0], y = (double)f[1], z = (double)f[2];
        o.x = (float)(T.m[0] * x + T.m[1] * y + T.m[2] * z + T.m[3]);
        o.y = (float)(T.m[4] * x + T.m[5] * y + T.m[6] * z + T.m[7]);
        o.z = (float)(T.m[8] * x + T.m[9] * y + T.m[10] * z + T.m[11]);
    }
    o.w = *reinterpret_cast<const float*>(p + ioff);
    out[i] = o;
}

__global__ void __launch_bounds__(256) k_fuse_mark(const float4* __restrict__ pts, uint32_t n, FuseBounds B, uint32_t* __restrict__ keep) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint32_t k = 0;
    if (i < n) {
        const float4 p = pts[i];
        const float v[3] = {p.x, p.y, p.z};
        bool ok = true;
        if (B.ext) {
            // pcl::PassThrough: non-finite points go; keep min <= v <= max (float limits)
            ok = isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
#pragma unroll
            for (int d = 0; d < 3; d++) ok = ok && !(v[d] < B.emin[d] || v[d] > B.emax[d]);
        }
        if (ok && B.inn) {
            // pcl::ConditionalRemoval with a ConditionOr of FieldComparison GT / LT (double compares): keep if any holds
            bool any = false;
#pragma unroll
            for (int d = 0; d < 3; d++) any = any || ((double)v[d] > B.imax[d]) || ((double)v[d] < B.imin[d]);
            ok = any;
        }
        k = ok ? 1u : 0u;
    }
    keep[i] = k;
}

__global__ void __launch_bounds__(256) k_fuse_compact(const float4* __restrict__ pts, uint32_t n, const uint32_t* __restrict__ slot, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (slot[i + 1] != slot[i]) out[slot[i]] = pts[i];
}

__global__ void __launch_bounds__(256) k_fuse_unpack(const float4* __restrict__ in, uint32_t n, unsigned char* __restrict__ out, size_t stride, int ioff) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    float* o = reinterpret_cast<float*>(out + (size_t)i * stride);
    o[0] = p.x; o[1] = p.y; o[2] = p.z;
    if (stride >= 32) o[3] = 1.0f;
    *reinterpret_cast<float*>(out + (size_t)i * stride + ioff) = p.w;
}

}  // namespace b2

using namespace b2;

struct FuseInput { DevBuf raw; size_t stride = 0, n = 0; FuseT T{}; };

struct b2_fusion_s {
    cudaStream_t stream = nullptr;
    std::vector<FuseInput*> inputs;
    FuseBounds bounds{};
    DevBuf fused, kept, work, out_raw;
    uint32_t n_fused = 0, n_out = 0;
    bool ran = false;
    float last_ms = 0.f;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

extern "C" {

int b2_fusion_create(b2_fusion_t* out) {
    if (!out) return B2_ERR_ARG;
    *out = nullptr;
    b2_fusion_s* h = new b2_fusion_s();
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&h->e0) != cudaSuccess || cudaEventCreate(&h->e1) != cudaSuccess) {
        set_error("b2_fusion_create: %s", cudaGetErrorString(cudaGetLastError())); delete h; return B2_ERR_CUDA;
    }
    *out = h;
    return B2_OK;
}

int b2_fusion_clear(b2_fusion_t h) {
    if (!h) return B2_ERR_ARG;
    cudaStreamSynchronize(h->stream);
    for (FuseInput* in : h->inputs) { in->raw.release(); delete in; }
    h->inputs.clear();
    return B2_OK;
}

int b2_fusion_destroy(b2_fusion_t h) {
    if (!h) return B2_OK;
    b2_fusion_clear(h);
    h->fused.release(); h->kept.release(); h->work.release(); h->out_raw.release();
    if (h->e0) cudaEventDestroy(h->e0);
    if (h->e1) cudaEventDestroy(h->e1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return B2_OK;
}

/* clouds are fused in the order they are added; T (row-major 4x4 double) may be NULL: the cloud is taken as it is */
int b2_fusion_add_cloud(b2_fusion_t h, const void* pts, size_t stride, size_t n, const double T[16]) {
    if (!h || (n && !pts) || stride < 16 || (stride & 3) || n > 0x7fffffffull) { set_error("b2_fusion_add_cloud: bad argument"); return B2_ERR_ARG; }
    FuseInput* in = new FuseInput();
    in->stride = stride; in->n = n;
    in->T.identity = T ? 0 : 1;
    if (T) for (int i = 0; i < 12; i++) in->T.m[i] = T[i];
    if (n) {
        int st = in->raw.reserve(n * stride);
        if (st != B2_OK) { delete in; return st; }
        if (cudaMemcpyAsync(in->raw.p, pts, n * stride, cudaMemcpyHostToDevice, h->stream) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess) {
            set_error("b2_fusion_add_cloud: %s", cudaGetErrorString(cudaGetLastError())); in->raw.release(); delete in; return B2_ERR_CUDA;
        }
    }
    h->inputs.push_back(in);
    return B2_OK;
}

int b2_fusion_set_external_bounds(b2_fusion_t h, int enabled, const double min_xyz[3], const double max_xyz[3]) {
    if (!h || (enabled && (!min_xyz || !max_xyz))) return B2_ERR_ARG;
    h->bounds.ext = enabled ? 1 : 0;
    if (enabled) for (int d = 0; d < 3; d++) { h->bounds.emin[d] = (float)min_xyz[d]; h->bounds.emax[d] = (float)max_xyz[d]; }     // setFilterLimits takes floats
    return B2_OK;
}

int b2_fusion_set_internal_bounds(b2_fusion_t h, int enabled, const double min_xyz[3], const double max_xyz[3]) {
    if (!h || (enabled && (!min_xyz || !max_xyz))) return B2_ERR_ARG;
    h->bounds.inn = enabled ? 1 : 0;
    if (enabled) for (int d = 0; d < 3; d++) { h->bounds.imin[d] = min_xyz[d]; h->bounds.imax[d] = max_xyz[d]; }
    return B2_OK;
}

int b2_fusion_run(b2_fusion_t h, size_t* n_fused, size_t* n_out) {
    B2_NVTX("b2_fusion_run");
    if (!h) return B2_ERR_ARG;
    cudaStream_t s = h->stream;
    cudaEventRecord(h->e0, s);
    uint64_t total = 0;
    for (FuseInput* in : h->inputs) total += in->n;
    if (total > 0x7ffffff0ull) { set_error("b2_fusion_run: fused cloud too large"); return B2_ERR_TOO_LARGE; }
    const uint32_t n = (uint32_t)total;
    h->n_fused = n; h->n_out = 0; h->ran = false;
    B2_CHECK(h->fused.reserve(std::max<size_t>(n, 1) * sizeof(float4)));
    uint32_t off = 0;
    for (FuseInput* in : h->inputs) {
        if (!in->n) continue;
        k_fuse_transform<<<(unsigned)((in->n + 255) / 256), 256, 0, s>>>(in->raw.as<unsigned char>(), in->stride, (int)B2_INTENSITY_OFFSET(in->stride), (uint32_t)in->n,
                                                                      in->T, h->fused.as<float4>() + off); count_launch();
        off += (uint32_t)in->n;
    }
    B2_CUDA(cudaGetLastError());
    const float4* result = h->fused.as<float4>();
    uint32_t m = n;
    if (n && (h->bounds.ext || h->bounds.inn)) {
        const size_t np1 = (size_t)n + 1;
        B2_CHECK(h->work.reserve(((np1 + 63) & ~(size_t)63) * 4 + scan_tmp_bytes(np1) + 256));
        uint32_t* keep = h->work.as<uint32_t>();
        char* scratch = reinterpret_cast<char*>(keep + ((np1 + 63) & ~(size_t)63));
        k_fuse_mark<<<(unsigned)((np1 + 255) / 256), 256, 0, s>>>(h->fused.as<float4>(), n, h->bounds, keep); count_launch();
        B2_CHECK(exclusive_scan_u32(keep, np1, scratch, s));
        B2_CHECK(h->kept.reserve((size_t)n * sizeof(float4)));
        k_fuse_compact<<<(n + 255) / 256, 256, 0, s>>>(h->fused.as<float4>(), n, keep, h->kept.as<float4>()); count_launch();
        B2_CUDA(cudaGetLastError());
        B2_CUDA(cudaMemcpyAsync(&m, keep + n, 4, cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaStreamSynchronize(s));
        result = h->kept.as<float4>();
    }
    (void)result;
    h->n_out = m;
    cudaEventRecord(h->e1, s); cudaEventSynchronize(h->e1);
    cudaEventElapsedTime(&h->last_ms, h->e0, h->e1);
    h->ran = true;
    if (n_fused) *n_fused = n;
    if (n_out) *n_out = m;
    return B2_OK;
}

/* the fused (and filtered) cloud: packed xyzi float4 in device memory, valid until the next run / clear */
int b2_fusion_device_cloud(b2_fusion_t h, const void** d_xyzi, size_t* n) {
    if (!h || !d_xyzi || !n) return B2_ERR_ARG;
    if (!h->ran) { set_error("b2_fusion_device_cloud: run first"); return B2_ERR_STATE; }
    *d_xyzi = (h->bounds.ext || h->bounds.inn) && h->n_fused ? h->kept.p : h->fused.p;
    *n = h->n_out;
    return B2_OK;
}

int b2_fusion_get(b2_fusion_t h, void* out, size_t stride, size_t capacity, size_t* n) {
    if (!h || !n || stride < 16 || (stride & 3)) return B2_ERR_ARG;
    if (!h->ran) { set_error("b2_fusion_get: run first"); return B2_ERR_STATE; }
    *n = h->n_out;
    if (!out || !h->n_out) return B2_OK;
    if (capacity < h->n_out) { set_error("b2_fusion_get: capacity %zu < %u", capacity, h->n_out); return B2_ERR_CAPACITY; }
    const float4* src = static_cast<const float4*>((h->bounds.ext || h->bounds.inn) ? h->kept.p : h->fused.p);
    cudaStream_t s = h->stream;
    if (stride == 16) { B2_CUDA(cudaMemcpyAsync(out, src, (size_t)h->n_out * 16, cudaMemcpyDeviceToHost, s)); }
    else {
        B2_CHECK(h->out_raw.reserve((size_t)h->n_out * stride));
        B2_CUDA(cudaMemsetAsync(h->out_raw.p, 0, (size_t)h->n_out * stride, s));
        k_fuse_unpack<<<(h->n_out + 255) / 256, 256, 0, s>>>(src, h->n_out, h->out_raw.as<unsigned char>(), stride, (int)B2_INTENSITY_OFFSET(stride)); count_launch();
        B2_CUDA(cudaMemcpyAsync(out, h->out_raw.p, (size_t)h->n_out * stride, cudaMemcpyDeviceToHost, s));
    }
    B2_CUDA(cudaStreamSynchronize(s));
    return B2_OK;
}

int b2_fusion_last_gpu_ms(b2_fusion_t h, float* ms) {
    if (!h || !ms) return B2_ERR_ARG;
    *ms = h->last_ms;
    return B2_OK;
}

}  // extern "C"
