// Local-map assembly on the device: the step right before the scan-to-map loop.
// Replaces mapOptimization::extractCloud and the containers it works on:
//   liosam_ws/src/LIO-SAM/src/mapOptmization.cpp
//     :1512-1524  cornerCloudKeyFrames.push_back / surfCloudKeyFrames.push_back (saveKeyFramesAndFactor)
//     :899-938    extractCloud: per key frame transformPointCloud(cloud, pose) through the laserCloudMapContainer cache,
//                 `+=` concatenation in visiting order, downSizeFilterCorner / downSizeFilterSurf, cache cleared above 1000 entries
//     :1591       laserCloudMapContainer.clear() after a loop closure corrected the poses
//     :1289-1290  kdtree{Corner,Surf}FromMap->setInputCloud(laserCloud{Corner,Surf}FromMapDS)
// Key-frame clouds are uploaded once and stay in HBM (packed xyzi, 16 B per point); a transformed copy is cached per key
// frame exactly like laserCloudMapContainer; one gather kernel concatenates the visited key frames in the reference's
// order (the VoxelGrid centroid sums depend on it), the two VoxelGrid filters run device-to-device, and the scan-to-map
// handle indexes the result without the map ever crossing PCIe. Transform arithmetic is pointAssociateToMap's
// (x' = m00 x + m01 y + m02 z + m03, float, no FMA), the matrix pcl::getTransformation's, evaluated on the host.
#include "b2_common.cuh"
#include <vector>
#include <algorithm>

namespace b2 {

struct LmSegment { const float4* src; uint32_t begin; uint32_t n; };      // one key frame's slice of the concatenation

__global__ void __launch_bounds__(256) k_lm_pack(const unsigned char* __restrict__ raw, size_t stride, int ioff, uint32_t n, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char* p = raw + (size_t)i * stride;
    const float* f = reinterpret_cast<const float*>(p);
    out[i] = make_float4(f[0], f[1], f[2], *reinterpret_cast<const float*>(p + ioff));
}

struct LmAffine { float m[12]; };

__global__ void __launch_bounds__(256) k_lm_transform(const float4* __restrict__ in, uint32_t n, LmAffine A, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    float4 o;
    o.x = A.m[0] * p.x + A.m[1] * p.y + A.m[2] * p.z + A.m[3];
    o.y = A.m[4] * p.x + A.m[5] * p.y + A.m[6] * p.z + A.m[7];
    o.z = A.m[8] * p.x + A.m[9] * p.y + A.m[10] * p.z + A.m[11];
    o.w = p.w;
    out[i] = o;
}

// concatenation: output element j belongs to the segment found by binary search over the segment begins
__global__ void __launch_bounds__(256) k_lm_concat(const LmSegment* __restrict__ seg, int n_seg, uint32_t total, float4* __restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total) return;
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (seg[mid].begin <= j) lo = mid; else hi = mid - 1;
    }
    out[j] = seg[lo].src[j - seg[lo].begin];
}

__global__ void __launch_bounds__(256) k_lm_unpack(const float4* __restrict__ in, uint32_t n, unsigned char* __restrict__ out, size_t stride, int ioff) {
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

struct LmKeyFrame {
    DevBuf corner, surf;            // as saved (sensor frame)
    DevBuf t_corner, t_surf;        // laserCloudMapContainer entry (map frame)
    uint32_t n_corner = 0, n_surf = 0;
    float pose[6] = {0, 0, 0, 0, 0, 0};
    bool cached = false;
};

struct b2_localmap_s {
    cudaStream_t stream = nullptr;
    b2_voxel_t vox_corner = nullptr, vox_surf = nullptr;
    int device = b2::current_device();      // the device the handle was created on
    std::vector<LmKeyFrame*> keys;
    size_t n_cached = 0;
    DevBuf raw, cat_corner, cat_surf, seg;
    PinBuf pin;
    uint32_t n_cat_corner = 0, n_cat_surf = 0, n_ds_corner = 0, n_ds_surf = 0;
    bool extracted = false;
    float last_ms = 0.f;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

static int lm_upload(b2_localmap_s* h, const void* pts, size_t stride, size_t n, DevBuf& dst) {
    if (!n) return B2_OK;
    B2_CHECK(h->raw.reserve(n * stride));
    B2_CHECK(dst.reserve(n * sizeof(float4)));
    B2_CUDA(cudaMemcpyAsync(h->raw.p, pts, n * stride, cudaMemcpyHostToDevice, h->stream));
    k_lm_pack<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->raw.as<unsigned char>(), stride, (int)B2_INTENSITY_OFFSET(stride), (uint32_t)n, dst.as<float4>()); count_launch();
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaStreamSynchronize(h->stream));          // raw is reused by the next upload
    return B2_OK;
}

static void lm_drop_cache(b2_localmap_s* h) {
    for (LmKeyFrame* k : h->keys) if (k->cached) { k->t_corner.release(); k->t_surf.release(); k->cached = false; }
    h->n_cached = 0;
}

extern "C" {

int b2_localmap_create(b2_localmap_t* out, float mapping_corner_leaf_size, float mapping_surf_leaf_size) {
    if (!out || !(mapping_corner_leaf_size > 0.f) || !(mapping_surf_leaf_size > 0.f)) return B2_ERR_ARG;
    *out = nullptr;
    b2_localmap_s* h = new b2_localmap_s();
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&h->e0) != cudaSuccess ||
        cudaEventCreate(&h->e1) != cudaSuccess) {
        set_error("b2_localmap_create: %s", cudaGetErrorString(cudaGetLastError())); delete h; return B2_ERR_CUDA;
    }
    int st = b2_voxel_create(&h->vox_corner);
    if (st == B2_OK) st = b2_voxel_create(&h->vox_surf);
    if (st == B2_OK) st = b2_voxel_set_leaf_size(h->vox_corner, mapping_corner_leaf_size, mapping_corner_leaf_size, mapping_corner_leaf_size);
    if (st == B2_OK) st = b2_voxel_set_leaf_size(h->vox_surf, mapping_surf_leaf_size, mapping_surf_leaf_size, mapping_surf_leaf_size);
    if (st != B2_OK) { b2_localmap_destroy(h); return st; }
    *out = h;
    return B2_OK;
}

int b2_localmap_destroy(b2_localmap_t h) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h) return B2_OK;
    for (LmKeyFrame* k : h->keys) { k->corner.release(); k->surf.release(); k->t_corner.release(); k->t_surf.release(); delete k; }
    h->raw.release(); h->cat_corner.release(); h->cat_surf.release(); h->seg.release(); h->pin.release();
    if (h->vox_corner) b2_voxel_destroy(h->vox_corner);
    if (h->vox_surf) b2_voxel_destroy(h->vox_surf);
    if (h->e0) cudaEventDestroy(h->e0);
    if (h->e1) cudaEventDestroy(h->e1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return B2_OK;
}

int b2_localmap_add_keyframe(b2_localmap_t h, const void* corner, size_t corner_stride, size_t n_corner,
                             const void* surf, size_t surf_stride, size_t n_surf, const float pose6[6], int* key_index) {
    B2_NVTX("b2_localmap_add_keyframe");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !pose6 || (n_corner && !corner) || (n_surf && !surf) || corner_stride < 16 || surf_stride < 16 || (corner_stride & 3) || (surf_stride & 3) ||
        n_corner > 0x7fffffffull || n_surf > 0x7fffffffull) { set_error("b2_localmap_add_keyframe: bad argument"); return B2_ERR_ARG; }
    LmKeyFrame* k = new LmKeyFrame();
    int st = lm_upload(h, corner, corner_stride, n_corner, k->corner);
    if (st == B2_OK) st = lm_upload(h, surf, surf_stride, n_surf, k->surf);
    if (st != B2_OK) { k->corner.release(); k->surf.release(); delete k; return st; }
    k->n_corner = (uint32_t)n_corner; k->n_surf = (uint32_t)n_surf;
    memcpy(k->pose, pose6, sizeof(k->pose));
    h->keys.push_back(k);
    if (key_index) *key_index = (int)h->keys.size() - 1;
    return B2_OK;
}

int b2_localmap_num_keyframes(b2_localmap_t h, int* n) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !n) return B2_ERR_ARG;
    *n = (int)h->keys.size();
    return B2_OK;
}

/* correctPoses(): a corrected pose only matters once the cache is dropped, as in the reference (:1591) */
int b2_localmap_set_pose(b2_localmap_t h, int key_index, const float pose6[6]) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !pose6 || key_index < 0 || key_index >= (int)h->keys.size()) return B2_ERR_ARG;
    memcpy(h->keys[key_index]->pose, pose6, 24);
    return B2_OK;
}

int b2_localmap_clear_cache(b2_localmap_t h) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h) return B2_ERR_ARG;
    cudaStreamSynchronize(h->stream);
    lm_drop_cache(h);
    return B2_OK;
}

/* extractCloud(cloudToExtract) for the key frames the caller kept after the distance test of :905 */
int b2_localmap_extract(b2_localmap_t h, const int32_t* key_indices, int n_keys, size_t* n_corner_ds, size_t* n_surf_ds) {
    B2_NVTX("b2_localmap_extract");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || n_keys < 0 || (n_keys && !key_indices)) return B2_ERR_ARG;
    for (int i = 0; i < n_keys; i++)
        if (key_indices[i] < 0 || key_indices[i] >= (int)h->keys.size()) { set_error("b2_localmap_extract: key index %d out of range", key_indices[i]); return B2_ERR_ARG; }
    cudaStream_t s = h->stream;
    cudaEventRecord(h->e0, s);
    h->extracted = false;
    // 1. transformed clouds, through the cache
    uint64_t tot_c = 0, tot_s = 0;
    for (int i = 0; i < n_keys; i++) {
        LmKeyFrame* k = h->keys[key_indices[i]];
        if (!k->cached) {
            LmAffine A;
            pose_to_affine_host(k->pose, A.m);
            B2_CHECK(k->t_corner.reserve(std::max<size_t>(k->n_corner, 1) * sizeof(float4)));
            B2_CHECK(k->t_surf.reserve(std::max<size_t>(k->n_surf, 1) * sizeof(float4)));
            if (k->n_corner) { k_lm_transform<<<(k->n_corner + 255) / 256, 256, 0, s>>>(k->corner.as<float4>(), k->n_corner, A, k->t_corner.as<float4>()); count_launch(); }
            if (k->n_surf) { k_lm_transform<<<(k->n_surf + 255) / 256, 256, 0, s>>>(k->surf.as<float4>(), k->n_surf, A, k->t_surf.as<float4>()); count_launch(); }
            k->cached = true; h->n_cached++;
        }
        tot_c += k->n_corner; tot_s += k->n_surf;
    }
    B2_CUDA(cudaGetLastError());
    if (tot_c > 0x7ffffff0ull || tot_s > 0x7ffffff0ull) { set_error("b2_localmap_extract: map too large"); return B2_ERR_TOO_LARGE; }
    // 2. concatenation in visiting order (a key frame listed twice is added twice, as `+=` would)
    B2_CHECK(h->pin.reserve((size_t)std::max(n_keys, 1) * 2 * sizeof(LmSegment)));
    B2_CHECK(h->seg.reserve((size_t)std::max(n_keys, 1) * 2 * sizeof(LmSegment)));
    LmSegment* hs = h->pin.as<LmSegment>();
    uint32_t oc = 0, os = 0;
    for (int i = 0; i < n_keys; i++) {
        LmKeyFrame* k = h->keys[key_indices[i]];
        hs[i] = LmSegment{k->t_corner.as<float4>(), oc, k->n_corner};
        hs[n_keys + i] = LmSegment{k->t_surf.as<float4>(), os, k->n_surf};
        oc += k->n_corner; os += k->n_surf;
    }
    h->n_cat_corner = oc; h->n_cat_surf = os;
    B2_CHECK(h->cat_corner.reserve(std::max<size_t>(oc, 1) * sizeof(float4)));
    B2_CHECK(h->cat_surf.reserve(std::max<size_t>(os, 1) * sizeof(float4)));
    if (n_keys) {
        B2_CUDA(cudaMemcpyAsync(h->seg.p, hs, (size_t)n_keys * 2 * sizeof(LmSegment), cudaMemcpyHostToDevice, s));
        if (oc) { k_lm_concat<<<(oc + 255) / 256, 256, 0, s>>>(h->seg.as<LmSegment>(), n_keys, oc, h->cat_corner.as<float4>()); count_launch(); }
        if (os) { k_lm_concat<<<(os + 255) / 256, 256, 0, s>>>(h->seg.as<LmSegment>() + n_keys, n_keys, os, h->cat_surf.as<float4>()); count_launch(); }
        B2_CUDA(cudaGetLastError());
    }
    B2_CUDA(cudaStreamSynchronize(s));
    // 3. downSizeFilterCorner / downSizeFilterSurf, device to device (each on its own handle's stream)
    // The two filters are independent handles with their own streams, and each has two host waits inside (its bounding box, its
    // voxel count): run one after the other they cost the sum. Their stages are interleaved instead — both boxes queued, then both
    // pipelines, then both counts — so the waits overlap and the two streams run side by side (round 2 first used a host thread
    // for the corner filter: creating and joining it cost more than it hid).
    uint32_t mc = 0, ms = 0;
    int refused = 0, refused_c = 0;
    B2_CHECK(voxel_filter_dev_begin(h->vox_corner, h->cat_corner.as<unsigned char>(), 16, oc, 4, 16, std::max<size_t>(oc, 1), nullptr));
    B2_CHECK(voxel_filter_dev_begin(h->vox_surf, h->cat_surf.as<unsigned char>(), 16, os, 4, 16, std::max<size_t>(os, 1), nullptr));
    B2_CHECK(voxel_filter_dev_middle(h->vox_corner));
    B2_CHECK(voxel_filter_dev_middle(h->vox_surf));
    B2_CHECK(voxel_filter_dev_end(h->vox_corner, &mc, &refused_c));
    B2_CHECK(voxel_filter_dev_end(h->vox_surf, &ms, &refused));
    // (a refused filter — PCL's "leaf size too small" — leaves its copy queued on the filter's stream)
    if (cudaStreamSynchronize(voxel_stream(h->vox_corner)) != cudaSuccess || cudaStreamSynchronize(voxel_stream(h->vox_surf)) != cudaSuccess) {
        set_error("b2_localmap_extract: %s", cudaGetErrorString(cudaGetLastError())); return B2_ERR_CUDA;
    }
    h->n_ds_corner = mc; h->n_ds_surf = ms;
    // 4. "clear map cache if too large" (:936-937)
    if (h->n_cached > 1000) lm_drop_cache(h);
    cudaEventRecord(h->e1, s); cudaEventSynchronize(h->e1);
    cudaEventElapsedTime(&h->last_ms, h->e0, h->e1);
    h->extracted = true;
    if (n_corner_ds) *n_corner_ds = mc;
    if (n_surf_ds) *n_surf_ds = ms;
    return B2_OK;
}

/* which: 0 laserCloudCornerFromMap, 1 laserCloudSurfFromMap (before the filters), 2 / 3 the DS clouds after them */
int b2_localmap_get(b2_localmap_t h, int which, void* out, size_t stride, size_t capacity, size_t* n) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || which < 0 || which > 3 || !n || stride < 16 || (stride & 3)) return B2_ERR_ARG;
    if (!h->extracted) { set_error("b2_localmap_get: extract first"); return B2_ERR_STATE; }
    const uint32_t cnt = which == 0 ? h->n_cat_corner : which == 1 ? h->n_cat_surf : which == 2 ? h->n_ds_corner : h->n_ds_surf;
    const void* src = which == 0 ? h->cat_corner.p : which == 1 ? h->cat_surf.p : which == 2 ? voxel_out_dev(h->vox_corner) : voxel_out_dev(h->vox_surf);
    *n = cnt;
    if (!out || !cnt) return B2_OK;
    if (capacity < cnt) { set_error("b2_localmap_get: capacity %zu < %u", capacity, cnt); return B2_ERR_CAPACITY; }
    cudaStream_t s = h->stream;
    if (stride == 16) { B2_CUDA(cudaMemcpyAsync(out, src, (size_t)cnt * 16, cudaMemcpyDeviceToHost, s)); }
    else {
        B2_CHECK(h->raw.reserve((size_t)cnt * stride));
        B2_CUDA(cudaMemsetAsync(h->raw.p, 0, (size_t)cnt * stride, s));
        k_lm_unpack<<<(cnt + 255) / 256, 256, 0, s>>>(static_cast<const float4*>(src), cnt, h->raw.as<unsigned char>(), stride, (int)B2_INTENSITY_OFFSET(stride)); count_launch();
        B2_CUDA(cudaMemcpyAsync(out, h->raw.p, (size_t)cnt * stride, cudaMemcpyDeviceToHost, s));
    }
    B2_CUDA(cudaStreamSynchronize(s));
    return B2_OK;
}

/* kdtreeCornerFromMap->setInputCloud(laserCloudCornerFromMapDS); kdtreeSurfFromMap->setInputCloud(laserCloudSurfFromMapDS) (:1289-1290) */
int b2_s2m_set_map_from_localmap(b2_s2m_t s2m, b2_localmap_t h) {
    B2_NVTX("b2_s2m_set_map_from_localmap");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!s2m || !h) return B2_ERR_ARG;
    if (!h->extracted) { set_error("b2_s2m_set_map_from_localmap: extract first"); return B2_ERR_STATE; }
    return s2m_set_map_device(s2m, voxel_out_dev(h->vox_corner), h->n_ds_corner, voxel_out_dev(h->vox_surf), h->n_ds_surf);
}

int b2_localmap_last_gpu_ms(b2_localmap_t h, float* ms, size_t* n_cached) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h) return B2_ERR_ARG;
    if (ms) *ms = h->last_ms;
    if (n_cached) *n_cached = h->n_cached;
    return B2_OK;
}

}  // extern "C"
