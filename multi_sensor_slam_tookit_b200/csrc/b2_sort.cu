// Device-wide exclusive scan and stable LSD radix sort (key,value) used by the voxel grid and the grid index.
// Hand-written for sm_100a; all traffic is coalesced 128-bit where the data allows, scatter is staged through
// shared memory so the global writes of a tile land in per-digit runs.
#include "b2_common.cuh"
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace b2 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

static unsigned long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// ------------------------------------------------------------------------------------------------ device memory pool
struct PoolBlock { void* p; size_t cap; int dev; };
static std::mutex g_pool_mu;
static std::vector<PoolBlock> g_pool;
static size_t g_pool_bytes = 0;
constexpr size_t POOL_LIMIT = (size_t)48 << 30;

static void pool_trim_locked(int dev) {
    for (size_t i = 0; i < g_pool.size();) {
        if (dev < 0 || g_pool[i].dev == dev) { cudaFree(g_pool[i].p); g_pool_bytes -= g_pool[i].cap; g_pool[i] = g_pool.back(); g_pool.pop_back(); }
        else i++;
    }
}

void* pool_alloc(size_t bytes, size_t* cap_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        size_t best = (size_t)-1;
        for (size_t i = 0; i < g_pool.size(); i++) {
            const PoolBlock& b = g_pool[i];
            if (b.dev == dev && b.cap >= bytes && b.cap <= 2 * bytes + ((size_t)1 << 20) && (best == (size_t)-1 || b.cap < g_pool[best].cap)) best = i;
        }
        if (best != (size_t)-1) {
            PoolBlock b = g_pool[best];
            g_pool[best] = g_pool.back(); g_pool.pop_back();
            g_pool_bytes -= b.cap;
            *cap_out = b.cap;
            return b.p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        { std::lock_guard<std::mutex> lk(g_pool_mu); cudaDeviceSynchronize(); pool_trim_locked(dev); }
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return nullptr; }
    *cap_out = bytes;
    return p;
}

void pool_free(void* p, size_t cap) {
    if (!p) return;
    // the block belongs to the device it was allocated on, whatever device the calling thread has current (a handle may be
    // destroyed from another thread, or after b2_set_device): ask the pointer, synchronise and tag THAT device
    int dev = 0;
    cudaGetDevice(&dev);
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeDevice) dev = at.device; else cudaGetLastError();
    {
        DeviceScope scope(dev);
        cudaDeviceSynchronize();       // nothing in flight may still use the block when another handle picks it up
    }
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool_bytes + cap > POOL_LIMIT) { cudaFree(p); return; }
    g_pool.push_back(PoolBlock{p, cap, dev});
    g_pool_bytes += cap;
}

void pool_trim() {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    cudaDeviceSynchronize();
    pool_trim_locked(-1);
}

int device_sm_count() {
    static int sm = 0;
    if (sm == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm <= 0) sm = 148;
    }
    return sm;
}

// ------------------------------------------------------------------------------------------------ scan
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        uint32_t w = lane < nw ? warp_sums[lane] : 0u, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        if (lane < nw) warp_sums[lane] = winc - w;
        if (lane == nw - 1) warp_sums[32] = winc;
    }
    __syncthreads();
    block_total = warp_sums[32];
    uint32_t r = warp_sums[warp] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const uint32_t* __restrict__ d, size_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t ws[33];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) if (base + i < n) s += d[base + i];
    uint32_t tot;
    block_exclusive_scan(s, ws, tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(uint32_t* __restrict__ d, size_t n, const uint32_t* __restrict__ offs) {
    __shared__ uint32_t ws[33];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = (base + i < n) ? d[base + i] : 0u; s += v[i]; }
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(s, ws, tot) + (offs ? offs[blockIdx.x] : 0u);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) d[base + i] = ex; ex += v[i]; }
}

size_t scan_tmp_bytes(size_t n) {
    size_t total = 0;
    while (n > SCAN_TILE) { n = (n + SCAN_TILE - 1) / SCAN_TILE; total += ((n * sizeof(uint32_t)) + 255) & ~(size_t)255; }
    return total + 256;
}

int exclusive_scan_u32(uint32_t* d, size_t n, void* tmp, cudaStream_t s) {
    if (n == 0) return B2_OK;
    size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb == 1) {
        k_scan_apply<<<1, SCAN_THREADS, 0, s>>>(d, n, nullptr); count_launch();
        B2_CUDA(cudaGetLastError());
        return B2_OK;
    }
    uint32_t* sums = reinterpret_cast<uint32_t*>(tmp);
    k_scan_tile_sums<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(d, n, sums); count_launch();
    B2_CUDA(cudaGetLastError());
    char* next = reinterpret_cast<char*>(tmp) + (((nb * sizeof(uint32_t)) + 255) & ~(size_t)255);
    B2_CHECK(exclusive_scan_u32(sums, nb, next, s));
    k_scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(d, n, sums); count_launch();
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

// ------------------------------------------------------------------------------------------------ radix sort
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;      // 4096 keys per CTA
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_WARP_SPAN = RS_TILE / RS_WARPS;    // 512 consecutive keys per warp

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t nblk,
                                                        uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int i = 0; i < RS_ITEMS; i++) {
        size_t p = base + (size_t)i * RS_THREADS + threadIdx.x;
        if (p < n) atomicAdd(&h[(keys[p] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                           uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                           size_t n, int shift, uint32_t nblk, const uint32_t* __restrict__ hist) {
    __shared__ uint32_t wcnt[RS_WARPS][256];
    __shared__ uint32_t dstart[256];
    __shared__ uint32_t gbase[256];
    __shared__ uint32_t ws[33];
    __shared__ uint32_t skeys[RS_TILE];
    __shared__ uint32_t svals[RS_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const size_t tile_base = (size_t)blockIdx.x * RS_TILE;
    const size_t wbase = tile_base + (size_t)warp * RS_WARP_SPAN;
    uint32_t k[RS_ITEMS], v[RS_ITEMS], rk[RS_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        size_t p = wbase + (size_t)r * 32 + lane;
        bool valid = p < n;
        k[r] = valid ? keys[p] : 0xffffffffu;
        v[r] = valid ? vals[p] : 0u;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        size_t p = wbase + (size_t)r * 32 + lane;
        bool valid = p < n;
        uint32_t mask = __ballot_sync(0xffffffffu, valid);
        rk[r] = 0;
        if (valid) {
            uint32_t d = (k[r] >> shift) & 255u;
            uint32_t peers = __match_any_sync(mask, d);
            uint32_t prev = wcnt[warp][d];
            __syncwarp(mask);
            uint32_t r0 = __popc(peers & lt);
            if (r0 == 0) wcnt[warp][d] = prev + __popc(peers);
            __syncwarp(mask);
            rk[r] = prev + r0;
        }
    }
    __syncthreads();
    // per digit: turn per-warp counts into per-warp bases, total per digit
    uint32_t total = 0;
    {
        const int t = threadIdx.x;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) { uint32_t c = wcnt[w][t]; wcnt[w][t] = total; total += c; }
    }
    uint32_t blk_total;
    uint32_t ds = block_exclusive_scan(total, ws, blk_total);
    dstart[threadIdx.x] = ds;
    gbase[threadIdx.x] = hist[(size_t)threadIdx.x * nblk + blockIdx.x] - ds;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        size_t p = wbase + (size_t)r * 32 + lane;
        if (p < n) {
            uint32_t d = (k[r] >> shift) & 255u;
            uint32_t pos = dstart[d] + wcnt[warp][d] + rk[r];
            skeys[pos] = k[r]; svals[pos] = v[r];
        }
    }
    __syncthreads();
    const uint32_t count = (uint32_t)min((size_t)RS_TILE, n - tile_base);
#pragma unroll 4
    for (int j = 0; j < RS_ITEMS; j++) {
        uint32_t p = j * RS_THREADS + threadIdx.x;
        if (p < count) {
            uint32_t key = skeys[p];
            uint32_t d = (key >> shift) & 255u;
            uint32_t g = gbase[d] + p;
            keys_out[g] = key; vals_out[g] = svals[p];
        }
    }
}


// ---- single-pass radix passes ("onesweep"): one read of the keys gives the digit histograms of ALL passes; each pass is then ONE
// kernel in which a tile ranks its keys, publishes its per-digit counts, obtains the counts of the tiles before it by decoupled
// look-back (status word per (tile, digit): 2 flag bits + 30 count bits) and scatters. Against histogram + three scan launches +
// scatter per pass: a 24-bit sort is 5 launches instead of 15, and a pass moves 16 B per pair instead of 20 B + the scans.
constexpr uint32_t RS_FLAG_AGG = 1u << 30, RS_FLAG_PFX = 2u << 30, RS_CNT_MASK = (1u << 30) - 1u;
constexpr int RS_MAX_PASSES = 4;

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist_all(const uint32_t* __restrict__ keys, size_t n, int passes, uint32_t* __restrict__ ghist) {
    __shared__ uint32_t h[RS_MAX_PASSES][256];
    for (int i = threadIdx.x; i < RS_MAX_PASSES * 256; i += RS_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    for (size_t p = (size_t)blockIdx.x * RS_THREADS + threadIdx.x; p < n; p += (size_t)gridDim.x * RS_THREADS) {
        const uint32_t k = keys[p];
#pragma unroll
        for (int q = 0; q < RS_MAX_PASSES; q++) if (q < passes) atomicAdd(&h[q][(k >> (8 * q)) & 255u], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += RS_THREADS) { const uint32_t v = (&h[0][0])[i]; if (v) atomicAdd(&ghist[i], v); }
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_onesweep(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                            uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                            size_t n, int shift, const uint32_t* __restrict__ ghist /* this pass */,
                                                            uint32_t* __restrict__ status /* [nblk][256], zeroed */, uint32_t* __restrict__ ticket) {
    __shared__ uint32_t wcnt[RS_WARPS][256];
    __shared__ uint32_t dstart[256];
    __shared__ uint32_t gbase[256];
    __shared__ uint32_t ws[33];
    __shared__ uint32_t skeys[RS_TILE];
    __shared__ uint32_t svals[RS_TILE];
    __shared__ uint32_t s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);       // tiles are taken in order: every earlier tile is already running
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const size_t tile_base = (size_t)tile * RS_TILE;
    const size_t wbase = tile_base + (size_t)warp * RS_WARP_SPAN;
    uint32_t k[RS_ITEMS], v[RS_ITEMS], rk[RS_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const size_t p = wbase + (size_t)r * 32 + lane;
        const bool valid = p < n;
        k[r] = valid ? keys[p] : 0xffffffffu;
        v[r] = valid ? vals[p] : 0u;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const size_t p = wbase + (size_t)r * 32 + lane;
        const bool valid = p < n;
        const uint32_t mask = __ballot_sync(0xffffffffu, valid);
        rk[r] = 0;
        if (valid) {
            const uint32_t d = (k[r] >> shift) & 255u;
            const uint32_t peers = __match_any_sync(mask, d);
            const uint32_t prev = wcnt[warp][d];
            __syncwarp(mask);
            const uint32_t r0 = __popc(peers & lt);
            if (r0 == 0) wcnt[warp][d] = prev + __popc(peers);
            __syncwarp(mask);
            rk[r] = prev + r0;
        }
    }
    __syncthreads();
    // thread t owns digit t: per-warp counts -> per-warp bases, tile total of the digit
    uint32_t total = 0;
    {
        const int t = threadIdx.x;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) { const uint32_t c = wcnt[w][t]; wcnt[w][t] = total; total += c; }
    }
    // publish the tile's count of this digit, then add up the tiles before it
    volatile uint32_t* st = status;
    const int d = threadIdx.x;
    if (tile == 0) st[d] = RS_FLAG_PFX | total;
    else st[(size_t)tile * 256 + d] = RS_FLAG_AGG | total;
    uint32_t excl = 0;
    if (tile > 0) {
        for (uint32_t t = tile - 1;; t--) {
            uint32_t w;
            do { w = st[(size_t)t * 256 + d]; } while ((w >> 30) == 0u);
            excl += w & RS_CNT_MASK;
            if ((w >> 30) == 2u || t == 0) break;
        }
        st[(size_t)tile * 256 + d] = RS_FLAG_PFX | (excl + total);
    }
    // start of the digit in the output = exclusive prefix of the global histogram (256 bins: one block scan per tile)
    uint32_t hist_total;
    const uint32_t gstart = block_exclusive_scan(ghist[d], ws, hist_total);
    __syncthreads();
    uint32_t blk_total;
    const uint32_t ds = block_exclusive_scan(total, ws, blk_total);
    dstart[d] = ds;
    gbase[d] = gstart + excl - ds;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const size_t p = wbase + (size_t)r * 32 + lane;
        if (p < n) {
            const uint32_t dg = (k[r] >> shift) & 255u;
            const uint32_t pos = dstart[dg] + wcnt[warp][dg] + rk[r];
            skeys[pos] = k[r]; svals[pos] = v[r];
        }
    }
    __syncthreads();
    const uint32_t count = (uint32_t)min((size_t)RS_TILE, n - tile_base);
#pragma unroll 4
    for (int j = 0; j < RS_ITEMS; j++) {
        const uint32_t p = j * RS_THREADS + threadIdx.x;
        if (p < count) {
            const uint32_t key = skeys[p];
            const uint32_t g = gbase[(key >> shift) & 255u] + p;
            keys_out[g] = key; vals_out[g] = svals[p];
        }
    }
}

size_t sort_tmp_bytes(size_t n) {
    size_t nblk = (n + RS_TILE - 1) / RS_TILE;
    size_t hist = ((256 * nblk * sizeof(uint32_t)) + 255) & ~(size_t)255;
    const size_t classic = hist + scan_tmp_bytes(256 * nblk);
    const size_t onesweep = (RS_MAX_PASSES * 256 + 64) * sizeof(uint32_t) + RS_MAX_PASSES * hist;       // histograms + tickets, status per pass
    return classic > onesweep ? classic : onesweep;
}

int radix_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, size_t n, int key_bits,
                     void* tmp, cudaStream_t s, uint32_t** keys_out, uint32_t** vals_out) {
    *keys_out = keys_a; *vals_out = vals_a;
    if (n == 0) return B2_OK;
    const uint32_t nblk = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    if (key_bits < 1) key_bits = 1;
    static const bool classic = getenv("B2_SORT_CLASSIC") != nullptr;                    // A/B switch (same result either way)
    const int passes = (key_bits + 7) / 8;
    if (!classic && passes <= RS_MAX_PASSES && n < ((size_t)1 << 30)) {
        uint32_t* ghist = reinterpret_cast<uint32_t*>(tmp);                               // [passes][256]
        uint32_t* tickets = ghist + RS_MAX_PASSES * 256;                                  // [passes]
        const size_t status_words = (((size_t)256 * nblk * sizeof(uint32_t) + 255) & ~(size_t)255) / sizeof(uint32_t);
        uint32_t* status = tickets + 64;                                                  // [passes][nblk][256]
        B2_CUDA(cudaMemsetAsync(tmp, 0, (RS_MAX_PASSES * 256 + 64) * sizeof(uint32_t) + (size_t)passes * status_words * sizeof(uint32_t), s));
        const unsigned hb = (unsigned)std::min<size_t>((n + RS_THREADS * 8 - 1) / (RS_THREADS * 8), (size_t)device_sm_count() * 4);
        k_rs_hist_all<<<std::max(hb, 1u), RS_THREADS, 0, s>>>(keys_a, n, passes, ghist); count_launch();
        B2_CUDA(cudaGetLastError());
        uint32_t *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
        for (int q = 0; q < passes; q++) {
            k_rs_onesweep<<<nblk, RS_THREADS, 0, s>>>(ki, vi, ko, vo, n, 8 * q, ghist + 256 * q, status + (size_t)q * status_words, tickets + q); count_launch();
            B2_CUDA(cudaGetLastError());
            uint32_t* t = ki; ki = ko; ko = t;
            t = vi; vi = vo; vo = t;
        }
        *keys_out = ki; *vals_out = vi;
        return B2_OK;
    }
    uint32_t* hist = reinterpret_cast<uint32_t*>(tmp);
    char* scan_tmp = reinterpret_cast<char*>(tmp) + (((256 * (size_t)nblk * sizeof(uint32_t)) + 255) & ~(size_t)255);
    uint32_t *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
    for (int shift = 0; shift < key_bits; shift += 8) {
        k_rs_hist<<<nblk, RS_THREADS, 0, s>>>(ki, n, shift, nblk, hist); count_launch();
        B2_CUDA(cudaGetLastError());
        B2_CHECK(exclusive_scan_u32(hist, 256 * (size_t)nblk, scan_tmp, s));
        k_rs_scatter<<<nblk, RS_THREADS, 0, s>>>(ki, vi, ko, vo, n, shift, nblk, hist); count_launch();
        B2_CUDA(cudaGetLastError());
        uint32_t* t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    *keys_out = ki; *vals_out = vi;
    return B2_OK;
}

}  // namespace b2

extern "C" {
int b2_version(void) { return 100; }
unsigned long long b2_kernel_launch_count(void) { return b2::launches(); }
const char* b2_last_error(void) { return b2::get_error(); }
int b2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int b2_trim_memory(void) { b2::pool_trim(); return B2_OK; }
int b2_set_device(int ordinal) {
    B2_CUDA(cudaSetDevice(ordinal));
    return B2_OK;
}
}
