// Point-to-point ICP on the device. Replaces pcl::IterativeClosestPoint<PointType, PointType> as the
// loop-closure thread of LIO-SAM uses it:
//   liosam_ws/src/LIO-SAM/src/mapOptmization.cpp
//     :559-565  setMaxCorrespondenceDistance / setMaximumIterations / setTransformationEpsilon / setEuclideanFitnessEpsilon /
//               setRANSACIterations(0)
//     :568-571  setInputSource / setInputTarget / align
//     :573      hasConverged(), getFitnessScore()        :580, :586  getFinalTransformation()
// Semantics: PCL 1.10 icp.hpp + DefaultConvergenceCriteria + TransformationEstimationSVD (Eigen::umeyama, no scaling).
// One iteration = one kernel over the working cloud: a warp serves 32 points, each through the warp-wide exact 1-NN walk
// of the 32-ary BVH over the target with FLANN's float distance (so the correspondences are the reference's, ties to the
// smaller index), then one lane per correspondence adds its 17 terms (n, sum p, sum q, sum q p^T, sum d^2) through the
// transposing warp reduction; per-warp partials are added on the host in warp order. The host turns the sums into the
// rigid transform (3x3 SVD in double), applies PCL's convergence test and launches the float update of the working
// cloud — the loop is a scalar recurrence with one reduction per step, like the reference's.
#include "b2_cloud.cuh"
#include "b2_bvh.cuh"
#include <cmath>
#include <cfloat>
#include <algorithm>
#include <vector>

namespace b2 {

constexpr int ICP_NSUM = 17;

// warp-wide exact nearest neighbour with float (FLANN L2_Simple) distances; returns d^2 and the position in T.pts
__device__ __forceinline__ void bvh_nn1f_warp(const BvhDev& T, float qx, float qy, float qz, float& best_d, uint32_t& best_pos) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float bd = INFINITY; long long bi = 0x7fffffffffffffffLL; uint32_t bp = 0xffffffffu;
    auto scan_leaf = [&](uint32_t leaf) {
        const uint32_t p = leaf * 32u + lane;
        float d = INFINITY; long long id = 0x7fffffffffffffffLL;
        if (p < T.n) {
            double x, y, z;
            load_p4d(&T.pts[p], x, y, z, id);
            const float dx = qx - (float)x, dy = qy - (float)y, dz = qz - (float)z;
            d = dx * dx;
            d = d + dy * dy;
            d = d + dz * dz;
        }
        uint32_t pp = p;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(full, d, o);
            const long long oi = ((long long)__shfl_xor_sync(full, (int)(id >> 32), o) << 32) | (unsigned)__shfl_xor_sync(full, (int)id, o);
            const uint32_t op = __shfl_xor_sync(full, pp, o);
            if (od < d || (od == d && oi < id)) { d = od; id = oi; pp = op; }
        }
        if (d < bd || (d == bd && id < bi)) { bd = d; bi = id; bp = pp; }
    };
    const int top = T.levels - 1;
    if (top == 0) { scan_leaf(0); best_d = bd; best_pos = bp; return; }
    double dch[BVH_MAXL];
    uint32_t node[BVH_MAXL];
    int lv = top;
    node[lv] = 0;
    const double dqx = (double)qx, dqy = (double)qy, dqz = (double)qz;
    auto expand = [&](int l, uint32_t nd) {
        const uint32_t c = nd * 32u + lane;
        dch[l] = (c < T.count[l - 1]) ? box_dist2(T.box[l - 1] + 6 * (size_t)c, dqx, dqy, dqz) : INFINITY;
    };
    expand(lv, 0);
    for (;;) {
        double dmin = dch[lv]; int jmin = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = shfl_xor_d(full, dmin, o);
            const int oj = __shfl_xor_sync(full, jmin, o);
            if (od < dmin || (od == dmin && oj < jmin)) { dmin = od; jmin = oj; }
        }
        // the box bound is exact in double; the point distances are rounded floats, so leave a relative slack
        if (dmin == INFINITY || dmin * (1.0 - 1e-6) > (double)bd) {
            if (++lv > top) break;
            continue;
        }
        if (lane == jmin) dch[lv] = INFINITY;
        const uint32_t child = node[lv] * 32u + (uint32_t)jmin;
        if (lv == 1) scan_leaf(child);
        else { lv--; node[lv] = child; expand(lv, child); }
    }
    best_d = bd; best_pos = bp;
}

struct IcpT { float m[12]; };

// correspondences + sums. apply = 1: the query is T * cur[i] (fitness pass on the original cloud); max_d2 < 0: no threshold
__global__ void __launch_bounds__(128) k_icp_correspond(BvhDev B, const float* __restrict__ cur, uint32_t n, IcpT X, int apply, double max_d2,
                                                        double* __restrict__ partial, int32_t* __restrict__ corr) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t base = warp * 32u;
    if (base >= n) return;
    float px = 0, py = 0, pz = 0;
    const bool valid = base + lane < n;
    if (valid) {
        const size_t i = base + lane;
        px = cur[3 * i]; py = cur[3 * i + 1]; pz = cur[3 * i + 2];
        if (apply) {
            const float x = px, y = py, z = pz;
            px = X.m[0] * x + X.m[1] * y + X.m[2] * z + X.m[3];
            py = X.m[4] * x + X.m[5] * y + X.m[6] * z + X.m[7];
            pz = X.m[8] * x + X.m[9] * y + X.m[10] * z + X.m[11];
        }
    }
    const int nq = (int)min(32u, n - base);
    float my_d = INFINITY; uint32_t my_pos = 0xffffffffu;
    for (int j = 0; j < nq; j++) {
        const float qx = __shfl_sync(full, px, j), qy = __shfl_sync(full, py, j), qz = __shfl_sync(full, pz, j);
        float d = INFINITY; uint32_t pos = 0xffffffffu;
        if (isfinite(qx) && isfinite(qy) && isfinite(qz)) bvh_nn1f_warp(B, qx, qy, qz, d, pos);
        if (lane == j) { my_d = d; my_pos = pos; }
    }
    double term[32];
#pragma unroll
    for (int i = 0; i < 32; i++) term[i] = 0.0;
    const bool hit = valid && my_pos != 0xffffffffu && !(max_d2 >= 0.0 && (double)my_d > max_d2);
    long long tidx = -1;
    if (hit) {
        double qx, qy, qz;
        load_p4d(&B.pts[my_pos], qx, qy, qz, tidx);
        const double p0 = (double)px, p1 = (double)py, p2 = (double)pz;
        term[0] = 1.0;
        term[1] = p0; term[2] = p1; term[3] = p2;
        term[4] = qx; term[5] = qy; term[6] = qz;
        term[7] = qx * p0; term[8] = qx * p1; term[9] = qx * p2;
        term[10] = qy * p0; term[11] = qy * p1; term[12] = qy * p2;
        term[13] = qz * p0; term[14] = qz * p1; term[15] = qz * p2;
        term[16] = (double)my_d;
    }
    if (corr && valid) corr[base + lane] = hit ? (int32_t)tidx : -1;
    __syncwarp();
    const double tot = warp_reduce_scatter32(term);
    if (lane < ICP_NSUM) partial[(size_t)warp * ICP_NSUM + lane] = tot;
}

__global__ void __launch_bounds__(256) k_icp_transform(float* __restrict__ cur, uint32_t n, IcpT X) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = cur[3 * (size_t)i], y = cur[3 * (size_t)i + 1], z = cur[3 * (size_t)i + 2];
    cur[3 * (size_t)i] = X.m[0] * x + X.m[1] * y + X.m[2] * z + X.m[3];
    cur[3 * (size_t)i + 1] = X.m[4] * x + X.m[5] * y + X.m[6] * z + X.m[7];
    cur[3 * (size_t)i + 2] = X.m[8] * x + X.m[9] * y + X.m[10] * z + X.m[11];
}

__global__ void __launch_bounds__(256) k_icp_pack(const unsigned char* __restrict__ raw, size_t stride, uint32_t n, float* __restrict__ xyz, double* __restrict__ wide) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    if (xyz) { xyz[3 * (size_t)i] = p[0]; xyz[3 * (size_t)i + 1] = p[1]; xyz[3 * (size_t)i + 2] = p[2]; }
    if (wide) { wide[3 * (size_t)i] = (double)p[0]; wide[3 * (size_t)i + 1] = (double)p[1]; wide[3 * (size_t)i + 2] = (double)p[2]; }
}

__global__ void __launch_bounds__(256) k_icp_unpack(const float* __restrict__ xyz, uint32_t n, unsigned char* __restrict__ out, size_t stride) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* o = reinterpret_cast<float*>(out + (size_t)i * stride);
    o[0] = xyz[3 * (size_t)i]; o[1] = xyz[3 * (size_t)i + 1]; o[2] = xyz[3 * (size_t)i + 2];
    if (stride >= 16) o[3] = 1.0f;
}

// symmetric 3x3 Jacobi (double), eigenvalues ascending, eigenvectors in columns
static void icp_jacobi3(const double A[9], double w[3], double V[9]) {
    double a[9];
    memcpy(a, A, sizeof(a));
    for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 50; sweep++) {
        const double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
        if (off < 1e-300) break;
        for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
            const double apq = a[p * 3 + q];
            if (apq == 0.0) continue;
            const double theta = (a[q * 3 + q] - a[p * 3 + p]) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 3; k++) { const double akp = a[k * 3 + p], akq = a[k * 3 + q]; a[k * 3 + p] = c * akp - s * akq; a[k * 3 + q] = s * akp + c * akq; }
            for (int k = 0; k < 3; k++) { const double apk = a[p * 3 + k], aqk = a[q * 3 + k]; a[p * 3 + k] = c * apk - s * aqk; a[q * 3 + k] = s * apk + c * aqk; }
            for (int k = 0; k < 3; k++) { const double vkp = V[k * 3 + p], vkq = V[k * 3 + q]; V[k * 3 + p] = c * vkp - s * vkq; V[k * 3 + q] = s * vkp + c * vkq; }
        }
    }
    w[0] = a[0]; w[1] = a[4]; w[2] = a[8];
    for (int i = 0; i < 2; i++) for (int j = i + 1; j < 3; j++) if (w[j] < w[i]) {
        std::swap(w[i], w[j]);
        for (int k = 0; k < 3; k++) std::swap(V[k * 3 + i], V[k * 3 + j]);
    }
}

// Eigen::umeyama(src, dst, false) from the sums of one iteration: R = U S V^T of the cross-covariance, t = mean_q - R mean_p
static void icp_umeyama(double n, const double sp[3], const double sq[3], const double sqp[9], float T[16]) {
    double mp[3], mq[3], S[9];
    for (int d = 0; d < 3; d++) { mp[d] = sp[d] / n; mq[d] = sq[d] / n; }
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) S[a * 3 + b] = sqp[a * 3 + b] / n - mq[a] * mp[b];
    double StS[9], w[3], V[9];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { double acc = 0; for (int k = 0; k < 3; k++) acc += S[k * 3 + a] * S[k * 3 + b]; StS[a * 3 + b] = acc; }
    icp_jacobi3(StS, w, V);
    double Vd[9], sv[3];
    for (int c = 0; c < 3; c++) { sv[c] = std::sqrt(std::max(w[2 - c], 0.0)); for (int r = 0; r < 3; r++) Vd[r * 3 + c] = V[r * 3 + (2 - c)]; }
    double U[9];
    for (int c = 0; c < 3; c++) {
        double col[3];
        for (int r = 0; r < 3; r++) col[r] = S[r * 3] * Vd[c] + S[r * 3 + 1] * Vd[3 + c] + S[r * 3 + 2] * Vd[6 + c];
        const double nrm = std::sqrt(col[0] * col[0] + col[1] * col[1] + col[2] * col[2]);
        if (c < 2 || nrm > 1e-12 * (sv[0] + 1e-300)) for (int r = 0; r < 3; r++) U[r * 3 + c] = nrm > 0 ? col[r] / nrm : (r == c ? 1.0 : 0.0);
        else {
            U[0 * 3 + 2] = U[1 * 3 + 0] * U[2 * 3 + 1] - U[2 * 3 + 0] * U[1 * 3 + 1];
            U[1 * 3 + 2] = U[2 * 3 + 0] * U[0 * 3 + 1] - U[0 * 3 + 0] * U[2 * 3 + 1];
            U[2 * 3 + 2] = U[0 * 3 + 0] * U[1 * 3 + 1] - U[1 * 3 + 0] * U[0 * 3 + 1];
        }
    }
    auto det3 = [](const double* M) { return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]); };
    const double sgn = det3(U) * det3(Vd) < 0 ? -1.0 : 1.0;
    double R[9];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) R[a * 3 + b] = U[a * 3] * Vd[b * 3] + U[a * 3 + 1] * Vd[b * 3 + 1] + sgn * U[a * 3 + 2] * Vd[b * 3 + 2];
    for (int i = 0; i < 16; i++) T[i] = 0.f;
    for (int a = 0; a < 3; a++) {
        for (int b = 0; b < 3; b++) T[a * 4 + b] = (float)R[a * 3 + b];
        T[a * 4 + 3] = (float)(mq[a] - (R[a * 3] * mp[0] + R[a * 3 + 1] * mp[1] + R[a * 3 + 2] * mp[2]));
    }
    T[15] = 1.f;
}

static void icp_mul4f(const float* A, const float* B, float* C) {
    float t[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { float s = 0; for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j]; t[i * 4 + j] = s; }
    memcpy(C, t, sizeof(t));
}

}  // namespace b2

using namespace b2;

struct b2_icp_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    // pcl::Registration / IterativeClosestPoint defaults
    double max_corr_dist = 1.3407807929942596e154;      // sqrt(DBL_MAX)
    int max_iterations = 10;
    double transformation_epsilon = 0.0;
    double euclidean_fitness_epsilon = -DBL_MAX;
    BvhIndex bvh;
    DevBuf src, cur, raw, work, partial, corr;
    PinBuf pin;
    size_t n_src = 0, n_tgt = 0;
    bool have_src = false, have_tgt = false;
    float final_T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    int nr_iterations = 0, launches = 0;
    bool converged = false;
    float last_ms = 0.f;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

// one pass over `pts`: returns the 17 sums (added in warp order)
static int icp_pass(b2_icp_s* h, const float* d_pts, const float T[16], int apply, double max_d2, double sums[ICP_NSUM], int32_t* d_corr) {
    for (int i = 0; i < ICP_NSUM; i++) sums[i] = 0.0;
    if (!h->n_src || !h->bvh.dev.n) return B2_OK;
    const uint32_t warps = (uint32_t)((h->n_src + 31) / 32);
    B2_CHECK(h->partial.reserve((size_t)warps * ICP_NSUM * 8));
    B2_CHECK(h->pin.reserve((size_t)warps * ICP_NSUM * 8));
    IcpT X;
    for (int i = 0; i < 12; i++) X.m[i] = T[i];
    k_icp_correspond<<<(warps + 3) / 4, 128, 0, h->stream>>>(h->bvh.dev, d_pts, (uint32_t)h->n_src, X, apply, max_d2, h->partial.as<double>(), d_corr); count_launch(); h->launches++;
    B2_CUDA(cudaGetLastError());
    double* hp = h->pin.as<double>();
    B2_CUDA(cudaMemcpyAsync(hp, h->partial.p, (size_t)warps * ICP_NSUM * 8, cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    for (uint32_t w = 0; w < warps; w++) for (int i = 0; i < ICP_NSUM; i++) sums[i] += hp[(size_t)w * ICP_NSUM + i];
    return B2_OK;
}

extern "C" {

int b2_icp_create(b2_icp_t* out) {
    if (!out) return B2_ERR_ARG;
    *out = nullptr;
    b2_icp_s* h = new b2_icp_s();
    if (cudaGetDevice(&h->device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->e0) != cudaSuccess || cudaEventCreate(&h->e1) != cudaSuccess) {
        set_error("b2_icp_create: %s", cudaGetErrorString(cudaGetLastError())); delete h; return B2_ERR_CUDA;
    }
    *out = h;
    return B2_OK;
}

int b2_icp_destroy(b2_icp_t h) {
    if (!h) return B2_OK;
    h->bvh.release(); h->src.release(); h->cur.release(); h->raw.release(); h->work.release(); h->partial.release(); h->corr.release(); h->pin.release();
    if (h->e0) cudaEventDestroy(h->e0);
    if (h->e1) cudaEventDestroy(h->e1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return B2_OK;
}

int b2_icp_set_max_correspondence_distance(b2_icp_t h, double d) { if (!h) return B2_ERR_ARG; h->max_corr_dist = d; return B2_OK; }
int b2_icp_set_maximum_iterations(b2_icp_t h, int n) { if (!h) return B2_ERR_ARG; h->max_iterations = n; return B2_OK; }
int b2_icp_set_transformation_epsilon(b2_icp_t h, double e) { if (!h) return B2_ERR_ARG; h->transformation_epsilon = e; return B2_OK; }
int b2_icp_set_euclidean_fitness_epsilon(b2_icp_t h, double e) { if (!h) return B2_ERR_ARG; h->euclidean_fitness_epsilon = e; return B2_OK; }
int b2_icp_set_ransac_iterations(b2_icp_t h, int n) {
    if (!h) return B2_ERR_ARG;
    if (n != 0) { set_error("b2_icp_set_ransac_iterations: only 0 is supported (the reference sets 0, mapOptmization.cpp:565)"); return B2_ERR_ARG; }
    return B2_OK;
}

int b2_icp_set_input_target(b2_icp_t h, const void* pts, size_t stride, size_t n) {
    if (!h || (n && !pts) || stride < 12 || (stride & 3) || n > 0x7fffffffull) return B2_ERR_ARG;
    h->have_tgt = false; h->n_tgt = n;
    B2_CUDA(cudaSetDevice(h->device));
    h->bvh.release();
    if (n) {
        DevBuf wide;
        B2_CHECK(h->raw.reserve(n * stride));
        B2_CHECK(wide.reserve(n * 24));
        B2_CUDA(cudaMemcpyAsync(h->raw.p, pts, n * stride, cudaMemcpyHostToDevice, h->stream));
        k_icp_pack<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->raw.as<unsigned char>(), stride, (uint32_t)n, nullptr, wide.as<double>()); count_launch();
        const int st = h->bvh.build(wide.as<double>(), n, h->work, h->stream);
        cudaStreamSynchronize(h->stream);
        wide.release();
        if (st != B2_OK) return st;
    }
    h->have_tgt = true;
    return B2_OK;
}

int b2_icp_set_input_source(b2_icp_t h, const void* pts, size_t stride, size_t n) {
    if (!h || (n && !pts) || stride < 12 || (stride & 3) || n > 0x7fffffffull) return B2_ERR_ARG;
    h->have_src = false; h->n_src = n;
    B2_CUDA(cudaSetDevice(h->device));
    if (n) {
        B2_CHECK(h->raw.reserve(n * stride));
        B2_CHECK(h->src.reserve(n * 12));
        B2_CUDA(cudaMemcpyAsync(h->raw.p, pts, n * stride, cudaMemcpyHostToDevice, h->stream));
        k_icp_pack<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->raw.as<unsigned char>(), stride, (uint32_t)n, h->src.as<float>(), nullptr); count_launch();
        B2_CUDA(cudaGetLastError());
        B2_CUDA(cudaStreamSynchronize(h->stream));
    }
    h->have_src = true;
    return B2_OK;
}

/* icp.hpp computeTransformation; guess may be NULL (identity). out_cloud (optional): the source under the final transformation. */
int b2_icp_align(b2_icp_t h, const float guess[16], void* out_cloud, size_t out_stride) {
    B2_NVTX("b2_icp_align");
    if (!h) return B2_ERR_ARG;
    if (!h->have_src || !h->have_tgt) { set_error("b2_icp_align: setInputSource and setInputTarget first"); return B2_ERR_STATE; }
    B2_CUDA(cudaSetDevice(h->device));
    static const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    cudaEventRecord(h->e0, h->stream);
    h->nr_iterations = 0; h->converged = false; h->launches = 0;
    memcpy(h->final_T, guess ? guess : I, sizeof(I));
    const uint32_t n = (uint32_t)h->n_src;
    B2_CHECK(h->cur.reserve(std::max<size_t>(n, 1) * 12));
    if (n) B2_CUDA(cudaMemcpyAsync(h->cur.p, h->src.p, (size_t)n * 12, cudaMemcpyDeviceToDevice, h->stream));
    if (n && memcmp(h->final_T, I, sizeof(I)) != 0) {
        IcpT G; for (int i = 0; i < 12; i++) G.m[i] = h->final_T[i];
        k_icp_transform<<<(n + 255) / 256, 256, 0, h->stream>>>(h->cur.as<float>(), n, G); count_launch(); h->launches++;
    }
    float tr[16];
    memcpy(tr, I, sizeof(I));
    double prev_mse = DBL_MAX;
    const double rotation_threshold = 1.0 - h->transformation_epsilon, translation_threshold = h->transformation_epsilon;
    const double mse_abs = 1e-12, mse_rel = h->euclidean_fitness_epsilon;
    const double max_d2 = h->max_corr_dist * h->max_corr_dist;
    bool conv = false;
    do {
        double s[ICP_NSUM];
        B2_CHECK(icp_pass(h, h->cur.as<float>(), I, 0, max_d2, s, nullptr));
        if (s[0] < 3) { conv = false; break; }                     // "Not enough correspondences found"
        icp_umeyama(s[0], &s[1], &s[4], &s[7], tr);
        IcpT X; for (int i = 0; i < 12; i++) X.m[i] = tr[i];
        k_icp_transform<<<(n + 255) / 256, 256, 0, h->stream>>>(h->cur.as<float>(), n, X); count_launch(); h->launches++;
        icp_mul4f(tr, h->final_T, h->final_T);
        ++h->nr_iterations;
        // DefaultConvergenceCriteria::hasConverged
        if (h->nr_iterations >= h->max_iterations) { conv = true; break; }
        const double cos_angle = 0.5 * ((double)tr[0] + (double)tr[5] + (double)tr[10] - 1);
        const double translation_sqr = (double)tr[3] * tr[3] + (double)tr[7] * tr[7] + (double)tr[11] * tr[11];
        if (cos_angle >= rotation_threshold && translation_sqr <= translation_threshold) { conv = true; break; }
        const double cur_mse = s[16] / s[0];
        if (std::fabs(cur_mse - prev_mse) < mse_abs) { conv = true; break; }
        if (std::fabs(cur_mse - prev_mse) / prev_mse < mse_rel) { conv = true; break; }
        prev_mse = cur_mse;
    } while (!conv);
    h->converged = conv;
    if (out_cloud && n) {
        if (out_stride < 12 || (out_stride & 3)) return B2_ERR_ARG;
        B2_CHECK(h->raw.reserve((size_t)n * out_stride));
        B2_CUDA(cudaMemsetAsync(h->raw.p, 0, (size_t)n * out_stride, h->stream));
        k_icp_unpack<<<(n + 255) / 256, 256, 0, h->stream>>>(h->cur.as<float>(), n, h->raw.as<unsigned char>(), out_stride); count_launch();
        B2_CUDA(cudaMemcpyAsync(out_cloud, h->raw.p, (size_t)n * out_stride, cudaMemcpyDeviceToHost, h->stream));
    }
    cudaEventRecord(h->e1, h->stream);
    B2_CUDA(cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&h->last_ms, h->e0, h->e1);
    return B2_OK;
}

int b2_icp_has_converged(b2_icp_t h, int* c) { if (!h || !c) return B2_ERR_ARG; *c = h->converged ? 1 : 0; return B2_OK; }
int b2_icp_get_final_transformation(b2_icp_t h, float T[16]) { if (!h || !T) return B2_ERR_ARG; memcpy(T, h->final_T, 64); return B2_OK; }
int b2_icp_get_final_num_iteration(b2_icp_t h, int* n) { if (!h || !n) return B2_ERR_ARG; *n = h->nr_iterations; return B2_OK; }

/* pcl::Registration::getFitnessScore(): mean squared distance of the source under the final transformation to its nearest target point */
int b2_icp_get_fitness_score(b2_icp_t h, double* score) {
    if (!h || !score) return B2_ERR_ARG;
    if (!h->have_src || !h->have_tgt) return B2_ERR_STATE;
    B2_CUDA(cudaSetDevice(h->device));
    double s[ICP_NSUM];
    B2_CHECK(icp_pass(h, h->src.as<float>(), h->final_T, 1, -1.0, s, nullptr));
    *score = s[0] > 0 ? s[16] / s[0] : DBL_MAX;
    return B2_OK;
}

int b2_icp_last_gpu_ms(b2_icp_t h, float* ms, int* launches) {
    if (!h) return B2_ERR_ARG;
    if (ms) *ms = h->last_ms;
    if (launches) *launches = h->launches;
    return B2_OK;
}

}  // extern "C"
