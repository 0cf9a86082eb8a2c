// Nearest-neighbour registration error and the yaw grid search built on it.
// Replaces, in SensorsCalibration's lidar2lidar auto-calibration,
//   Calibration_Tookit/SensorsCalibration/lidar2lidar/auto_calib/src/registration_icp.cpp
//     :51-52   pcl::KdTreeFLANN<pcl::PointXYZI> kdtree; kdtree.setInputCloud(tgt_ngcloud_)      -> b2_nnerr_set_target
//     :78-100  CalculateICPError: transform the source by T (double matrix, float result), 1-NN per point,
//              dist_sum += squared distance                                                       -> b2_nnerr_evaluate
//     :49-76   RegistrationByICP: 37 such evaluations over a shrinking yaw grid                  -> b2_nnerr_yaw_search
// One evaluation = one kernel: a warp serves 32 source points, each through the warp-wide exact 1-NN walk of the 32-ary
// BVH over the target (b2_bvh.cuh); per-warp partial sums are added on the host in warp order, so the error — and the
// argmin of the grid search — is reproducible. The grid search itself is 37 dependent evaluations and stays on the host.
#include "b2_cloud.cuh"
#include "b2_bvh.cuh"
#include <cmath>
#include <vector>

namespace b2 {

struct NnT { double m[12]; };

__global__ void __launch_bounds__(128) k_nnerr(const __grid_constant__ BvhDev T, const float* __restrict__ src, uint32_t n_src, NnT X, double* __restrict__ partial, uint32_t* __restrict__ found) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t base = warp * 32u;
    if (base >= n_src) return;
    float tx = 0, ty = 0, tz = 0;
    if (base + lane < n_src) {
        const size_t i = base + lane;
        const double x = (double)src[3 * i], y = (double)src[3 * i + 1], z = (double)src[3 * i + 2];
        // pcl::transformPointCloud with a double matrix: the sum runs in double, the point is stored as float
        tx = (float)(X.m[0] * x + X.m[1] * y + X.m[2] * z + X.m[3]);
        ty = (float)(X.m[4] * x + X.m[5] * y + X.m[6] * z + X.m[7]);
        tz = (float)(X.m[8] * x + X.m[9] * y + X.m[10] * z + X.m[11]);
    }
    const int nq = (int)min(32u, n_src - base);
    double sum = 0.0; uint32_t cnt = 0;
    for (int j = 0; j < nq; j++) {
        const double qx = (double)__shfl_sync(full, tx, j), qy = (double)__shfl_sync(full, ty, j), qz = (double)__shfl_sync(full, tz, j);
        double d2 = INFINITY;
        if (isfinite(qx) && isfinite(qy) && isfinite(qz)) d2 = bvh_nn1_warp(T, qx, qy, qz);
        if (d2 < INFINITY) { sum += d2; cnt++; }
    }
    if (lane == 0) { partial[warp] = sum; found[warp] = cnt; }
}

__global__ void __launch_bounds__(256) k_nnerr_pack(const unsigned char* __restrict__ raw, size_t stride, uint32_t n, float* __restrict__ xyz, double* __restrict__ wide) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    if (xyz) { xyz[3 * (size_t)i] = p[0]; xyz[3 * (size_t)i + 1] = p[1]; xyz[3 * (size_t)i + 2] = p[2]; }
    if (wide) { wide[3 * (size_t)i] = (double)p[0]; wide[3 * (size_t)i + 1] = (double)p[1]; wide[3 * (size_t)i + 2] = (double)p[2]; }
}

}  // namespace b2

using namespace b2;

struct b2_nnerr_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    BvhIndex bvh;
    DevBuf src, work, raw, partial, found;
    PinBuf pin;
    size_t n_src = 0, n_tgt = 0;
    bool have_tgt = false, have_src = false;
    int evaluations = 0;
    float last_ms = 0.f;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

static int nnerr_eval(b2_nnerr_s* h, const double T[16], double* dist_sum, size_t* n_found) {
    if (!h->have_tgt || !h->have_src) { set_error("nnerr: set_target and set_source first"); return B2_ERR_STATE; }
    *dist_sum = 0.0;
    if (n_found) *n_found = 0;
    h->evaluations++;
    if (!h->n_src || !h->bvh.dev.n) return B2_OK;
    const uint32_t warps = (uint32_t)((h->n_src + 31) / 32);
    B2_CHECK(h->partial.reserve((size_t)warps * 8));
    B2_CHECK(h->found.reserve((size_t)warps * 4));
    B2_CHECK(h->pin.reserve((size_t)warps * 12));
    NnT X;
    for (int i = 0; i < 12; i++) X.m[i] = T[i];
    k_nnerr<<<(warps + 3) / 4, 128, 0, h->stream>>>(h->bvh.dev, h->src.as<float>(), (uint32_t)h->n_src, X, h->partial.as<double>(), h->found.as<uint32_t>()); count_launch();
    B2_CUDA(cudaGetLastError());
    double* hp = h->pin.as<double>();
    uint32_t* hf = reinterpret_cast<uint32_t*>(hp + warps);
    B2_CUDA(cudaMemcpyAsync(hp, h->partial.p, (size_t)warps * 8, cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaMemcpyAsync(hf, h->found.p, (size_t)warps * 4, cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    double s = 0.0; size_t c = 0;
    for (uint32_t w = 0; w < warps; w++) { s += hp[w]; c += hf[w]; }
    *dist_sum = s;
    if (n_found) *n_found = c;
    return B2_OK;
}

static void nnerr_delta_t(float yaw, double D[16]) {
    // GetDeltaT(const float yaw) (:38-47): Rz(yaw * M_PI / 180) — as written in the reference
    const double a = (double)yaw * M_PI / 180.0;
    const double c = std::cos(a), s = std::sin(a);
    const double R[16] = {c, -s, 0, 0, s, c, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    memcpy(D, R, sizeof(R));
}
static void nnerr_mul4(const double* A, const double* B, double* C) {
    double t[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j]; t[i * 4 + j] = s; }
    memcpy(C, t, sizeof(t));
}

extern "C" {

int b2_nnerr_create(b2_nnerr_t* out) {
    if (!out) return B2_ERR_ARG;
    *out = nullptr;
    b2_nnerr_s* h = new b2_nnerr_s();
    if (cudaGetDevice(&h->device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->e0) != cudaSuccess || cudaEventCreate(&h->e1) != cudaSuccess) {
        set_error("b2_nnerr_create: %s", cudaGetErrorString(cudaGetLastError())); delete h; return B2_ERR_CUDA;
    }
    *out = h;
    return B2_OK;
}

int b2_nnerr_destroy(b2_nnerr_t h) {
    if (!h) return B2_OK;
    h->bvh.release(); h->src.release(); h->work.release(); h->raw.release(); h->partial.release(); h->found.release(); h->pin.release();
    if (h->e0) cudaEventDestroy(h->e0);
    if (h->e1) cudaEventDestroy(h->e1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return B2_OK;
}

int b2_nnerr_set_target(b2_nnerr_t h, const void* pts, size_t stride, size_t n) {
    if (!h || (n && !pts) || stride < 12 || (stride & 3) || n > 0x7fffffffull) return B2_ERR_ARG;
    h->have_tgt = false; h->n_tgt = n;
    B2_CUDA(cudaSetDevice(h->device));
    h->bvh.release();
    if (n) {
        DevBuf wide;
        B2_CHECK(h->raw.reserve(n * stride));
        B2_CHECK(wide.reserve(n * 24));
        B2_CUDA(cudaMemcpyAsync(h->raw.p, pts, n * stride, cudaMemcpyHostToDevice, h->stream));
        k_nnerr_pack<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->raw.as<unsigned char>(), stride, (uint32_t)n, nullptr, wide.as<double>()); count_launch();
        const int st = h->bvh.build(wide.as<double>(), n, h->work, h->stream);
        cudaStreamSynchronize(h->stream);
        wide.release();
        if (st != B2_OK) return st;
    }
    h->have_tgt = true;
    return B2_OK;
}

int b2_nnerr_set_source(b2_nnerr_t h, const void* pts, size_t stride, size_t n) {
    if (!h || (n && !pts) || stride < 12 || (stride & 3) || n > 0x7fffffffull) return B2_ERR_ARG;
    h->have_src = false; h->n_src = n;
    B2_CUDA(cudaSetDevice(h->device));
    if (n) {
        B2_CHECK(h->raw.reserve(n * stride));
        B2_CHECK(h->src.reserve(n * 12));
        B2_CUDA(cudaMemcpyAsync(h->raw.p, pts, n * stride, cudaMemcpyHostToDevice, h->stream));
        k_nnerr_pack<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->raw.as<unsigned char>(), stride, (uint32_t)n, h->src.as<float>(), nullptr); count_launch();
        B2_CUDA(cudaGetLastError());
        B2_CUDA(cudaStreamSynchronize(h->stream));
    }
    h->have_src = true;
    return B2_OK;
}

int b2_nnerr_evaluate(b2_nnerr_t h, const double T[16], double* dist_sum, size_t* n_found) {
    B2_NVTX("b2_nnerr_evaluate");
    if (!h || !T || !dist_sum) return B2_ERR_ARG;
    B2_CUDA(cudaSetDevice(h->device));
    return nnerr_eval(h, T, dist_sum, n_found);
}

int b2_nnerr_yaw_search(b2_nnerr_t h, const double init_guess[16], double T_out[16], double* best_yaw_out, double* min_error_out, int* evaluations) {
    B2_NVTX("b2_nnerr_yaw_search");
    if (!h || !init_guess || !T_out) return B2_ERR_ARG;
    B2_CUDA(cudaSetDevice(h->device));
    h->evaluations = 0;
    cudaEventRecord(h->e0, h->stream);
    auto error_at = [&](float yaw, double* err) {
        double D[16], T[16];
        nnerr_delta_t(yaw, D); nnerr_mul4(D, init_guess, T);
        return nnerr_eval(h, T, err, nullptr);
    };
    double cur_yaw = 0, min_error = 0;
    B2_CHECK(error_at((float)cur_yaw, &min_error));
    double best_yaw = cur_yaw;
    const float degree_2_radian = 0.017453293f;
    int iter_cnt = 0;
    double step = 5;
    int search_range = 10;
    while (iter_cnt < 5) {
        for (int delta = -search_range; delta < search_range; delta++) {
            const double yaw = cur_yaw + delta * step * degree_2_radian;
            double error = 0;
            B2_CHECK(error_at((float)yaw, &error));
            if (error < min_error) { min_error = error; best_yaw = yaw; }
        }
        search_range = (int)(search_range / 2 + 0.5);
        step /= 2;
        cur_yaw = best_yaw;
        iter_cnt++;
    }
    double D[16];
    nnerr_delta_t((float)best_yaw, D);
    nnerr_mul4(D, init_guess, T_out);
    cudaEventRecord(h->e1, h->stream); cudaEventSynchronize(h->e1);
    cudaEventElapsedTime(&h->last_ms, h->e0, h->e1);
    if (best_yaw_out) *best_yaw_out = best_yaw;
    if (min_error_out) *min_error_out = min_error;
    if (evaluations) *evaluations = h->evaluations;
    return B2_OK;
}

int b2_nnerr_last_gpu_ms(b2_nnerr_t h, float* ms) {
    if (!h || !ms) return B2_ERR_ARG;
    *ms = h->last_ms;
    return B2_OK;
}

}  // extern "C"
