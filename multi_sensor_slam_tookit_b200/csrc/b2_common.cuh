// Shared host/device helpers for libb2reg (sm_100a). Compiled with -fmad=false: every float expression that
// feeds a floor(), an ordering or a threshold must round exactly like the reference's baseline-x86-64 build
// (no FMA contraction, liosam_ws/src/LIO-SAM/CMakeLists.txt:4-6).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <nvtx3/nvToolsExt.h>
#include "../../include/b2reg.h"

namespace b2 {

void set_error(const char* fmt, ...);

#define B2_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            b2::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
            return B2_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define B2_CHECK(expr)                                                                             \
    do {                                                                                           \
        int _s = (expr);                                                                           \
        if (_s != B2_OK) return _s;                                                                \
    } while (0)

// Device memory comes from a per-process pool (b2_sort.cu): released blocks are kept and handed out again, because
// cudaMalloc/cudaFree of the multi-megabyte index buffers cost more than the kernels that use them when a caller
// builds an index per scan or per calibration pair. pool_free synchronises the device before a block can be reused.
void* pool_alloc(size_t bytes, size_t* cap_out);
void pool_free(void* p, size_t cap);

// Growable device buffer owned by a handle.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return B2_OK;
        if (p) { pool_free(p, cap); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4 + 256;
        p = pool_alloc(want, &cap);
        if (!p) { cap = 0; return B2_ERR_CUDA; }
        return B2_OK;
    }
    void release() { if (p) pool_free(p, cap); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Pinned host staging buffer (H2D / D2H of caller-owned pageable memory goes through it so copies are async).
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return B2_OK;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) { set_error("cudaMallocHost(%zu) -> %s", want, cudaGetErrorString(e)); return B2_ERR_CUDA; }
        cap = want;
        return B2_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

int device_sm_count();

// NVTX range per C-ABI call (header-only NVTX 3: a no-op function pointer unless a tool such as nsys / ncu --nvtx is attached)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
};
#define B2_NVTX(name) b2::NvtxRange nvtx_range_(name)

// Every handle records the device it was created on; its entry points (and its destroy, which may run on another host
// thread, e.g. from a Python finaliser) make that device current for the duration of the call.
struct DeviceScope {
    int prev = -1;
    explicit DeviceScope(int dev) {
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != dev && dev >= 0) { if (cudaSetDevice(dev) == cudaSuccess) prev = cur; }
    }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceScope(const DeviceScope&) = delete;
};
inline int current_device() { int d = 0; if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; } return d; }
void count_launch(int n = 1);     // bookkeeping for b2_kernel_launch_count()

// ---- scan / sort primitives (b2_sort.cu) -------------------------------------------------------
// exclusive prefix sum of n uint32 in place; d_tmp must hold scan_tmp_bytes(n)
size_t scan_tmp_bytes(size_t n);
int exclusive_scan_u32(uint32_t* d_data, size_t n, void* d_tmp, cudaStream_t s);
// stable LSD radix sort of (key, value) pairs on bits [0, key_bits); result returned in *keys_out/*vals_out which
// point at either the a or the b buffers. tmp must hold sort_tmp_bytes(n).
size_t sort_tmp_bytes(size_t n);
int radix_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, size_t n, int key_bits,
                     void* d_tmp, cudaStream_t s, uint32_t** keys_out, uint32_t** vals_out);

// ---- uniform-grid index (b2_grid.cu) -----------------------------------------------------------
struct GridDev {               // passed by value to kernels
    const float4* pts;         // cell-sorted points: x, y, z, __int_as_float(original index)
    const uint32_t* cell_start;// ncell + 1
    float ox, oy, oz, inv_h, h;
    int nx, ny, nz;
    int n;
    float max_d2;              // candidates at or beyond this squared distance are never reported (max_dist^2)
};
// GridDev as it lives in device memory when the device sizes the grid itself (GridIndex::build_async): the geometry, the
// bounding-box words the build reduces into, and a status (0 ok, B2_ERR_TOO_LARGE when the cells exceed the allocated table).
struct GridDevMem { GridDev g; uint32_t bb[6]; int status; uint32_t dirty[2]; uint32_t bar; long long tl[8]; };   // bar: arrivals at the build kernel's grid barriers   // dirty: entries in use in each half of the cell table
struct GridIndex {
    DevBuf pts, cell_start, cell_of, tmp, raw;
    DevBuf devmem;                                 // GridDevMem
    size_t cell_budget = 0;                        // cells the allocated cell_start can hold (device-sized builds)
    bool tables_clean_ = false; int table_active_ = 0; uint32_t bar_total_ = 0;   // the two halves of cell_start (build_async)
    PinBuf stage;
    GridDev dev{};
    float h = 1.0078125f;
    size_t n = 0;
    // build = begin (upload + bounding box, asynchronous) then finish (one 24-byte read-back, then the counting sort), so
    // that two indexes can be built side by side on two streams (corner and surf maps of one scan)
    int begin(const void* host_pts, size_t stride, size_t n, float max_dist, cudaStream_t s);
    // the same from points already in device memory (local-map assembly); d_pts must stay valid until finish() returns
    int begin_device(const void* d_pts, size_t stride, size_t n, float max_dist, cudaStream_t s);
    int finish(cudaStream_t s);
    int build(const void* host_pts, size_t stride, size_t n, float max_dist, cudaStream_t s) {
        int st = begin(host_pts, stride, n, max_dist, s);
        return st != B2_OK ? st : finish(s);
    }
    // Device-sized build, no host round trip: upload_async queues the copy of the caller's points (or adopts device points),
    // build_async queues ONE cooperative kernel that reduces the bounding box, derives the grid geometry from it with the
    // arithmetic of finish(), counting-sorts the points and leaves the GridDev in devmem. Consumers read dev_ptr(). If the
    // box needs more cells than the table holds the kernel only sets the status; rebuild_exact() then takes the
    // host-sized path (begin_device + finish), stores its GridDev in devmem and raises the budget.
    int upload_async(const void* host_pts, const void* dev_pts, size_t stride, size_t n, float max_dist, cudaStream_t s, bool copy_dev = false);
    int build_async(cudaStream_t s);
    int rebuild_exact(cudaStream_t s);
    const GridDevMem* dev_ptr() const { return devmem.as<GridDevMem>(); }
    void release() { pts.release(); cell_start.release(); cell_of.release(); tmp.release(); raw.release(); stage.release(); devmem.release(); ingest_flag_.release(); cell_budget = 0; bb_ready_ = false; }
    // developer timeline (B2_S2M_TIMELINE=1): events after the upload, after the bounding box + read-back, after the build
    cudaEvent_t tl_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool tl_on = false;
    void tl_rec(int i, cudaStream_t s) {
        if (!tl_on) return;
        if (!tl_ev[i]) cudaEventCreate(&tl_ev[i]);
        cudaEventRecord(tl_ev[i], s);
    }
    size_t stride_ = 0; float max_dist_ = 1.f;
    bool bb_ready_ = false;                       // the bounding-box words in tmp are already reset (by the last scatter)
    const unsigned char* src_ = nullptr;          // device points the build reads (raw.p after an upload)
    const unsigned char* pinned_src_ = nullptr;   // pinned host points the build kernel copies into raw itself (no DMA), or nullptr
    // the kernel writes ingest_seq_ into this pinned word once it has read the last of the caller's points: wait_ingest() is how a
    // caller-facing function keeps the promise that the caller's buffers are free when it returns
    PinBuf ingest_flag_;
    uint32_t ingest_seq_ = 0; bool ingest_wait_ = false;
    int wait_ingest() {
        if (!ingest_wait_) return B2_OK;
        ingest_wait_ = false;
        volatile uint32_t* f = ingest_flag_.as<volatile uint32_t>();
        for (unsigned long long spins = 0; *f != ingest_seq_; spins++) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            // the kernel was launched without an error; if the device has died since, do not spin forever
            if ((spins & 0xfffffull) == 0xfffffull && cudaPeekAtLastError() != cudaSuccess) { set_error("grid index: device error while reading the caller's points"); return B2_ERR_CUDA; }
        }
        return B2_OK;
    }
};

// ---- internal couplings between the modules (not part of the C ABI)
// VoxelGrid on device buffers: result left in the handle's output buffer (voxel_out_dev) on voxel_stream(h)
int voxel_filter_dev(b2_voxel_s* h, const unsigned char* d_in, size_t in_stride, size_t n, int n_fields, size_t out_stride,
                     size_t out_capacity, uint32_t* m_out, int* refused, int32_t* d_vop);
// the same in three stages (see b2_voxel.cu): two filters on two handles can be interleaved from one host thread
int voxel_filter_dev_begin(b2_voxel_s* h, const unsigned char* d_in, size_t in_stride, size_t n, int n_fields, size_t out_stride,
                           size_t out_capacity, int32_t* d_vop);
int voxel_filter_dev_middle(b2_voxel_s* h);
int voxel_filter_dev_end(b2_voxel_s* h, uint32_t* m_out, int* refused);
const void* voxel_out_dev(b2_voxel_s* h);
cudaStream_t voxel_stream(b2_voxel_s* h);
int voxel_filter_host_to_dev(b2_voxel_s* h, const void* in, size_t in_stride, size_t n, uint32_t* m_out);
// per-scan front end: cornerCloud / surfaceCloud of the last b2_scan_extract_features, packed xyzi in device memory
int scan_features_dev(b2_scan_s* h, const void** d_corner, size_t* n_corner, const void** d_surf, size_t* n_surf);
// scan-to-map: index the two map clouds from device memory (packed xyzi, 16-byte stride)
int s2m_set_map_device(b2_s2m_s* h, const void* d_corner, size_t n_corner, const void* d_surf, size_t n_surf);
// pcl::getTransformation(x, y, z, roll, pitch, yaw) in float, row-major 3x4 (pose6 = roll, pitch, yaw, x, y, z)
void pose_to_affine_host(const float pose6[6], float xf[12]);

}  // namespace b2

// ---- device-side helpers -----------------------------------------------------------------------
#ifdef __CUDACC__
namespace b2 {

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// order-preserving float <-> uint mapping for atomicMin/Max on floats
__device__ __forceinline__ uint32_t float_flip(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_unflip(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace b2
#endif
