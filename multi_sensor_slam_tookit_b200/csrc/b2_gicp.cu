// Generalized ICP on the device, fp64. Replaces Open3D's registration_generalized_icp as Multi_LiCa calls it
//   Calibration_Tookit/Multi_LiCa/multi_lidar_calibrator/calibration/Calibration.py:331-340
//   (TransformationEstimationForGeneralizedICP(epsilon), ICPConvergenceCriteria(rel_fitness, rel_rmse, max_iteration))
// and is the kernel of the sharded map-to-map registration.
//
// One iteration = one launch of k_gicp_linearize (+ one all-reduce and a one-thread finalize kernel when the source
// is sharded over several GPUs):
//   search   one thread per source point of the shard: transform by the current T, exact nearest target point within
//            max_correspondence_distance (fine grid 3x3x3 block, certified by the ring bound; otherwise one pass over
//            the 3x3x3 block of a coarse grid whose cell edge is the search radius; ties to the smaller index);
//   terms    M = Ct + R Cs R^T, residual and Jacobian of the correspondence: 21 + 6 + 3 terms, reduced over the warp
//            with a shuffle tree as they are produced (lane i keeps the running total of term i);
//   epilogue per-CTA partial -> the last CTA adds the partials in CTA order (deterministic) and, on a single GPU,
//            solves the 6x6 system, updates T and evaluates Open3D's convergence rule, all on the device.
//
// Covariances. The reference never supplies covariances: Open3D derives them from the normals as R diag(eps,1,1) R^T
// with R the rotation of e1 onto the normal (identity when n.x < -0.99), i.e. C = I - (1-eps) m m^T with m the normal
// (or e1 in the special case). The device keeps m (24 B) instead of C (72 B). With u = m_target, v = R m_source,
// p = u + v, q = u - v:  M = 2I - (a/2)(pp^T + qq^T), a = 1 - eps, and since p is orthogonal to q
//   M^-1 = I/2 + a/(4 l1) pp^T + a/(4 l2) qq^T,  l1 = 2 - a|p|^2/2,  l2 = 2 - a|q|^2/2   (both >= 2 eps > 0).
// Open3D forms W = (M^-1)^(1/2), r = W d, J = W [-[vs]x | I]; only J^T J = A^T M^-1 A, J^T r = A^T M^-1 d and
// r^T r = d^T M^-1 d are ever used, so no matrix square root is needed.
#include "b2_cloud.cuh"
#include "b2_comm.cuh"
#include <cmath>
#include <algorithm>
#include <vector>
#include <cstdlib>

namespace b2 {

constexpr int GICP_THREADS = 256;
constexpr int GICP_WARPS = GICP_THREADS / 32;
constexpr int GICP_NSUM = 30;                // 21 JtJ upper + 6 Jtr + n_corr + sum d^2 + sum r^2
constexpr int GICP_HIST = 256;
constexpr uint32_t GICP_SHARD_CHUNKS = 128;  // 4096 consecutive (cell-sorted) source points per dealt block

struct GicpState {
    double T[16];
    double sums[GICP_NSUM];          // global sums of the last linearisation (after the all-reduce when sharded)
    double sums_local[GICP_NSUM];    // this GPU's sums
    double prev_fit, prev_rmse, fitness, rmse;
    double rel_fit, rel_rmse, n_src_total;
    int it, max_it, done, converged;
    unsigned ticket; int evals;
    double hist_fit[GICP_HIST], hist_rmse[GICP_HIST];
};

// Exchange area of the fused linearise + all-reduce (sharded registrations): every rank owns one, the peers map it through CUDA
// IPC and store their 30 local sums straight into it over NVLink, then a sequence number. Two slots alternate between
// evaluations (a rank can be at most one evaluation ahead of the slowest one: it cannot finish evaluation k + 1 before every
// peer has contributed to it, i.e. has finished reading evaluation k).
constexpr int GICP_MAX_WORLD = 16;
struct GicpPeerArea {
    double sums[2][GICP_MAX_WORLD][32];
    unsigned flag[2][GICP_MAX_WORLD];
};

struct GicpArgs {
    GridDDev tgt;                    // target index (fine: about two points per occupied cell)
    GridDDev tgtc;                   // the same points in cells of edge >= max_correspondence_distance
    const uint32_t* fine_pos_of;     // target original index -> position in the fine order
    int have_coarse;
    const double* tgt_m;             // effective normals of the target, fine-grid order
    const P4d* src;                  // source points, cell-sorted order, idx = original index
    const double* src_m;             // effective normals of the source, same order
    uint32_t n_src, n_local_chunks;  // valid source points; 32-point chunks this GPU owns
    int rank, world;                 // block-cyclic shard: blocks of GICP_SHARD_CHUNKS chunks dealt round-robin to the ranks
    double radius2, a;               // max_correspondence_distance^2, 1 - epsilon
    double* partials;                // [gridDim.x][GICP_NSUM]
    int32_t* corr;                   // optional: corr[source original index] = target original index or -1
    uint32_t* prev;                  // optional: prev[source position] = fine position of the last correspondence (seeds the search)
    GicpState* st;
    int mode;                        // 0: reduce into sums_local; 1: reduce into sums and finalize on the device;
                                     // 2: exchange the local sums with the peers inside the kernel, then finalize (every rank alike)
    GicpPeerArea* peers[GICP_MAX_WORLD];   // mode 2: exchange areas of all ranks (peers[rank] is this GPU's own)
    unsigned seq_base;               // mode 2: sequence numbers of this align are seq_base + evaluation + 1
    int xrank, xworld;               // mode 2: rank / size of the exchange (the shard deal above may be 0 / 1 for a sliced source)
};

// LDL^T solve of the symmetric 6x6 (Open3D: SolveLinearSystemPSD -> A.ldlt().solve(b)), no pivoting, oracle order
__device__ bool ldlt_solve6_d(const double A[36], const double b[6], double x[6]) {
    double L[36], D[6];
    for (int i = 0; i < 36; i++) L[i] = 0.0;
    for (int j = 0; j < 6; j++) {
        double d = A[j * 6 + j];
        for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
        D[j] = d;
        if (d == 0.0) return false;
        for (int i = j + 1; i < 6; i++) {
            double v = A[i * 6 + j];
            for (int k = 0; k < j; k++) v -= L[i * 6 + k] * L[j * 6 + k] * D[k];
            L[i * 6 + j] = v / d;
        }
    }
    double y[6];
    for (int i = 0; i < 6; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= L[i * 6 + k] * y[k]; y[i] = v; }
    for (int i = 0; i < 6; i++) y[i] /= D[i];
    for (int i = 5; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < 6; k++) v -= L[k * 6 + i] * x[k]; x[i] = v; }
    return true;
}

// fitness / inlier_rmse of the evaluation just reduced, Open3D's stopping rule, and the next T = [Rz Ry Rx | t] T
__device__ __noinline__ void gicp_finalize(GicpState* st) {
    const double* s = st->sums;
    const double fit = st->n_src_total > 0 ? s[27] / st->n_src_total : 0.0;
    const double rm = s[27] > 0 ? sqrt(s[28] / s[27]) : 0.0;
    st->fitness = fit; st->rmse = rm;
    if (st->evals < GICP_HIST) { st->hist_fit[st->evals] = fit; st->hist_rmse[st->evals] = rm; }
    st->evals++;
    if (st->it > 0 && fabs(st->prev_fit - fit) < st->rel_fit && fabs(st->prev_rmse - rm) < st->rel_rmse) { st->converged = 1; st->done = 1; return; }
    if (st->it >= st->max_it) { st->done = 1; return; }
    st->prev_fit = fit; st->prev_rmse = rm;
    double A[36], b[6], x[6];
    int q = 0;
    for (int r = 0; r < 6; r++) for (int c = r; c < 6; c++) { A[r * 6 + c] = s[q]; A[c * 6 + r] = s[q]; q++; }
    for (int r = 0; r < 6; r++) b[r] = -s[21 + r];
    if (!(s[27] < 1.0) && ldlt_solve6_d(A, b, x)) {
        const double ca = cos(x[0]), sa = sin(x[0]), cb = cos(x[1]), sb = sin(x[1]), cg = cos(x[2]), sg = sin(x[2]);
        double U[16];
        U[0] = cg * cb; U[1] = cg * sb * sa - sg * ca; U[2] = cg * sb * ca + sg * sa; U[3] = x[3];
        U[4] = sg * cb; U[5] = sg * sb * sa + cg * ca; U[6] = sg * sb * ca - cg * sa; U[7] = x[4];
        U[8] = -sb;     U[9] = cb * sa;                U[10] = cb * ca;               U[11] = x[5];
        U[12] = 0; U[13] = 0; U[14] = 0; U[15] = 1;
        double Tn[16];
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
            double acc = 0;
            for (int k = 0; k < 4; k++) acc += U[i * 4 + k] * st->T[k * 4 + j];
            Tn[i * 4 + j] = acc;
        }
        for (int i = 0; i < 16; i++) st->T[i] = Tn[i];
    }
    st->it++;
}

__global__ void k_gicp_finalize(GicpState* st) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && !st->done) gicp_finalize(st);
}

#ifndef GICP_MIN_BLOCKS
#define GICP_MIN_BLOCKS 4
#endif
#ifdef B2_NN1_STATS
__device__ unsigned long long g_nn1_stats[8];
#endif
__global__ void __launch_bounds__(GICP_THREADS, GICP_MIN_BLOCKS) k_gicp_linearize(GicpArgs A) {
    GicpState* st = A.st;
    if (st->done) return;
    __shared__ double s_red[GICP_WARPS][32];
    __shared__ bool s_last;
    __shared__ uint32_t s_tab[18][GICP_THREADS];       // per-thread row-range tables of the search (column = thread: no bank conflicts)
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // The pose is the same for every thread: kept in shared memory and read where it is used (broadcast loads). As 24 registers
    // per thread it was spilled around the search at the 64-register cap (local-memory traffic through L2 on every query).
    __shared__ double s_T[12];
    if (threadIdx.x < 12) s_T[threadIdx.x] = st->T[threadIdx.x < 9 ? (threadIdx.x / 3) * 4 + threadIdx.x % 3 : (threadIdx.x - 9) * 4 + 3];
    __syncthreads();
    const volatile double* R = s_T;          // R[0..8] row-major rotation
    const volatile double* t = s_T + 9;      // translation
    double totA = 0.0, totB = 0.0;   // lane l: running totals of terms (l & 15) and 16 + (l & 15)
    B2_STAT(NN1Stats stats = {};)
    for (uint32_t lc = blockIdx.x * GICP_WARPS + warp; lc < A.n_local_chunks; lc += gridDim.x * GICP_WARPS) {
        // local chunk -> global chunk of the block-cyclic deal (spatially mixed shards: every rank gets its share of the
        // points that need the coarse pass; contiguous slices left the slowest rank 40 % behind at 8 GPUs)
        const uint32_t chunk = ((lc / GICP_SHARD_CHUNKS) * (uint32_t)A.world + (uint32_t)A.rank) * GICP_SHARD_CHUNKS + lc % GICP_SHARD_CHUNKS;
        const uint32_t p = (chunk << 5) + lane;
        const bool valid = p < A.n_src;
        double px = 0, py = 0, pz = 0; long long sidx = -1;
        if (valid) load_p4d(&A.src[p], px, py, pz, sidx);
        const double vx = R[0] * px + R[1] * py + R[2] * pz + t[0];
        const double vy = R[3] * px + R[4] * py + R[5] * pz + t[1];
        const double vz = R[6] * px + R[7] * py + R[8] * pz + t[2];
        NN1 nn; nn.d2 = INFINITY; nn.pos = 0xffffffffu; nn.idx = -1;
        if (valid) nn = nn1_thread<GICP_THREADS>(A.tgt, A.tgtc, A.fine_pos_of, A.have_coarse != 0, vx, vy, vz, A.radius2, &s_tab[0][threadIdx.x], A.prev ? __ldcs(&A.prev[p]) : 0xffffffffu B2_STAT(, &stats));
        const bool hit = valid && nn.pos != 0xffffffffu;
        if (A.prev && valid) __stcs(&A.prev[p], nn.pos);
        if (A.corr && valid) A.corr[sidx] = hit ? (int32_t)nn.idx : -1;
        __syncwarp();
        if (!__any_sync(full, hit)) continue;
        // the 30 terms in two batches of 16 so that each batch lives in registers through its reduction:
        // lane l accumulates term (l & 15) of batch A in totA and term 16 + (l & 15) in totB
        double ta[16], tb[16];
#pragma unroll
        for (int i = 0; i < 16; i++) { ta[i] = 0.0; tb[i] = 0.0; }
        if (hit) {
            const double* um = &A.tgt_m[3 * (size_t)nn.pos];
            const double* sm = &A.src_m[3 * (size_t)p];
            const double u0 = um[0], u1 = um[1], u2 = um[2];
            const double m0 = sm[0], m1 = sm[1], m2 = sm[2];
            const double v0 = R[0] * m0 + R[1] * m1 + R[2] * m2;
            const double v1 = R[3] * m0 + R[4] * m1 + R[5] * m2;
            const double v2 = R[6] * m0 + R[7] * m1 + R[8] * m2;
            const double p0 = u0 + v0, p1 = u1 + v1, p2 = u2 + v2;
            const double q0 = u0 - v0, q1 = u1 - v1, q2 = u2 - v2;
            const double l1 = 2.0 - 0.5 * A.a * (p0 * p0 + p1 * p1 + p2 * p2);
            const double l2 = 2.0 - 0.5 * A.a * (q0 * q0 + q1 * q1 + q2 * q2);
            const double al = A.a / (4.0 * l1), be = A.a / (4.0 * l2);
            const double N00 = 0.5 + al * p0 * p0 + be * q0 * q0, N01 = al * p0 * p1 + be * q0 * q1, N02 = al * p0 * p2 + be * q0 * q2;
            const double N11 = 0.5 + al * p1 * p1 + be * q1 * q1, N12 = al * p1 * p2 + be * q1 * q2;
            const double N22 = 0.5 + al * p2 * p2 + be * q2 * q2;
            const double d0 = vx - nn.x, d1 = vy - nn.y, d2 = vz - nn.z;
            // B = N S, S = -[vs]x
            const double B00 = N02 * vy - N01 * vz, B01 = N00 * vz - N02 * vx, B02 = N01 * vx - N00 * vy;
            const double B10 = N12 * vy - N11 * vz, B11 = N01 * vz - N12 * vx, B12 = N11 * vx - N01 * vy;
            const double B20 = N22 * vy - N12 * vz, B21 = N02 * vz - N22 * vx, B22 = N12 * vx - N02 * vy;
            // top-left S^T B (symmetric), top-right S^T N = B^T, bottom-right N
            ta[0] = vy * B20 - vz * B10; ta[1] = vy * B21 - vz * B11; ta[2] = vy * B22 - vz * B12;
            ta[3] = B00; ta[4] = B10; ta[5] = B20;
            ta[6] = vz * B01 - vx * B21; ta[7] = vz * B02 - vx * B22;
            ta[8] = B01; ta[9] = B11; ta[10] = B21;
            ta[11] = vx * B12 - vy * B02;
            ta[12] = B02; ta[13] = B12; ta[14] = B22;
            ta[15] = N00; tb[0] = N01; tb[1] = N02; tb[2] = N11; tb[3] = N12; tb[4] = N22;
            const double g0 = N00 * d0 + N01 * d1 + N02 * d2, g1 = N01 * d0 + N11 * d1 + N12 * d2, g2 = N02 * d0 + N12 * d1 + N22 * d2;
            tb[5] = vy * g2 - vz * g1; tb[6] = vz * g0 - vx * g2; tb[7] = vx * g1 - vy * g0;
            tb[8] = g0; tb[9] = g1; tb[10] = g2;
            tb[11] = 1.0; tb[12] = nn.d2; tb[13] = d0 * g0 + d1 * g1 + d2 * g2;
        }
        __syncwarp();
        totA += warp_reduce_scatter16(ta);
        totB += warp_reduce_scatter16(tb);
    }
    B2_STAT(atomicAdd(&g_nn1_stats[0], stats.fine_cycles); atomicAdd(&g_nn1_stats[1], stats.coarse_cycles); atomicAdd(&g_nn1_stats[2], stats.coarse_queries);
            atomicAdd(&g_nn1_stats[3], stats.coarse_cands); atomicAdd(&g_nn1_stats[4], stats.coarse_exact); atomicAdd(&g_nn1_stats[5], stats.coarse_cells);
            atomicAdd(&g_nn1_stats[6], stats.unseeded);)
    // lane l < 16 holds term l in totA; lane l >= 16 holds term 16 + (l & 15) = l in totB: lane l -> term l
    const double tot = lane < 16 ? totA : totB;
    // epilogue: CTA partial, last CTA adds the partials in CTA order
    s_red[warp][lane] = tot;
    __syncthreads();
    if (threadIdx.x < GICP_NSUM) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < GICP_WARPS; w++) v += s_red[w][threadIdx.x];
        A.partials[(size_t)blockIdx.x * GICP_NSUM + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double* out = A.mode == 1 ? st->sums : st->sums_local;
    double v = 0.0;
    if (threadIdx.x < GICP_NSUM) {
        for (unsigned b = 0; b < gridDim.x; b++) v += __ldcg(&A.partials[(size_t)b * GICP_NSUM + threadIdx.x]);
        out[threadIdx.x] = v;
    }
    if (A.mode == 2) {
        // Fused exchange: this GPU's 30 sums go straight into every rank's exchange area (peer stores over NVLink, 240 B per
        // peer), then the sequence number of this evaluation; the epilogue waits for all ranks' numbers, adds the contributions
        // in rank order — the same bits on every rank — and runs the same finalisation as the single-GPU path. No NCCL launch
        // and no separate epilogue kernel per iteration.
        const int slot = st->evals & 1;
        const unsigned seq = A.seq_base + (unsigned)st->evals + 1u;
        if (threadIdx.x < GICP_NSUM)
            for (int p = 0; p < A.xworld; p++) A.peers[p]->sums[slot][A.xrank][threadIdx.x] = v;
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < A.xworld) {
            *reinterpret_cast<volatile unsigned*>(&A.peers[threadIdx.x]->flag[slot][A.xrank]) = seq;
            __threadfence_system();
            volatile unsigned* mine = reinterpret_cast<volatile unsigned*>(&A.peers[A.xrank]->flag[slot][threadIdx.x]);
            while (*mine != seq) __nanosleep(100);
            __threadfence_system();
        }
        __syncthreads();
        if (threadIdx.x < GICP_NSUM) {
            double t = 0.0;
            for (int r = 0; r < A.xworld; r++) t += *reinterpret_cast<volatile double*>(&A.peers[A.xrank]->sums[slot][r][threadIdx.x]);
            st->sums[threadIdx.x] = t;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        st->ticket = 0;
        if (A.mode != 0) gicp_finalize(st);
    }
}

// effective normal per point, gathered into the order of `order` (cell-sorted): e1 when n.x < -0.99 (Open3D's
// GetRotationFromE1ToX returns the identity there), the normal otherwise
__global__ void __launch_bounds__(256) k_gicp_eff_normals(const P4d* __restrict__ sorted, const double* __restrict__ nrm, uint32_t n,
                                                          double* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y, z; long long idx;
    load_p4d(&sorted[i], x, y, z, idx);
    double n0 = nrm[3 * (size_t)idx], n1 = nrm[3 * (size_t)idx + 1], n2 = nrm[3 * (size_t)idx + 2];
    if (n0 < -0.99) { n0 = 1.0; n1 = 0.0; n2 = 0.0; }
    out[3 * (size_t)i] = n0; out[3 * (size_t)i + 1] = n1; out[3 * (size_t)i + 2] = n2;
}

__global__ void __launch_bounds__(256) k_gicp_fine_pos(const P4d* __restrict__ sorted, uint32_t n, uint32_t* __restrict__ pos_of) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y, z; long long idx;
    load_p4d(&sorted[i], x, y, z, idx);
    pos_of[idx] = i;
}

__global__ void k_fill_i32(int32_t* p, uint32_t n, int32_t v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace b2

using namespace b2;

struct b2_gicp_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    b2_gicp_params prm{};
    GridD tgt_grid, tgt_coarse, src_grid;
    DevBuf tgt_m, src_m, partials, state, corr, tgt_xyz, fine_pos_of, prev, sub_xyz, sub_nrm;
    double coarse_for = -1.0;        // max_correspondence_distance the coarse grid was built for (< 0: none)
    bool have_coarse = false;
    int blocks_per_sm = GICP_MIN_BLOCKS;
    PinBuf pin;
    size_t n_tgt = 0, n_src = 0;
    uint32_t src_valid = 0;
    bool have_tgt = false, have_src = false;
    GicpPeerArea* peer_own = nullptr;                 // this GPU's exchange area (cudaMalloc: exported through CUDA IPC)
    GicpPeerArea* peer_ptr[GICP_MAX_WORLD] = {nullptr};
    int peer_world = 0, peer_rank = 0;                // > 0 once b2_gicp_set_peers succeeded: align exchanges inside the kernel
    unsigned peer_seq = 0;
    bool src_slice = false;          // src_grid holds only this rank's rows of the source (b2_gicp_set_source_slice)
    int rank = 0, world = 1;
    b2_comm_s* comm = nullptr;
    int grid_blocks = 1;
    float last_ms = 0.f; int last_launches = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<cudaEvent_t> ev;     // one per evaluation boundary of the last align (per-evaluation device times)
    int ev_used = 0;
};

static int gicp_valid_count(const GridD& g, cudaStream_t s, uint32_t* out) {
    const size_t ncell = (size_t)g.dev.nx * g.dev.ny * g.dev.nz;
    *out = 0;
    if (!g.n) return B2_OK;
    B2_CUDA(cudaMemcpyAsync(out, g.dev.cell_start + ncell, 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    return B2_OK;
}

// number of 32-point chunks the block-cyclic deal gives `rank` (chunks of full blocks + its part of the last round)
static uint32_t gicp_local_chunks(uint32_t n_valid, int rank, int world) {
    const uint32_t total = (n_valid + 31) / 32;
    const uint32_t S = GICP_SHARD_CHUNKS;
    const uint32_t round_chunks = S * (uint32_t)world;
    const uint32_t full_rounds = total / round_chunks;
    const uint32_t rem = total % round_chunks;
    uint32_t mine = full_rounds * S;
    const uint32_t lo = (uint32_t)rank * S;
    if (rem > lo) mine += std::min(S, rem - lo);
    return mine;
}

static void gicp_fill_args(b2_gicp_s* h, GicpArgs& a, int mode, int32_t* corr) {
    a.tgt = h->tgt_grid.dev;
    a.tgtc = h->have_coarse ? h->tgt_coarse.dev : h->tgt_grid.dev;
    a.fine_pos_of = h->fine_pos_of.as<uint32_t>();
    a.have_coarse = h->have_coarse ? 1 : 0;
    a.tgt_m = h->tgt_m.as<double>();
    a.src = h->src_grid.dev.pts;
    a.src_m = h->src_m.as<double>();
    a.n_src = h->src_valid;
    // a sliced source is all local: the deal over the ranks was made when the slice was taken
    a.rank = h->src_slice ? 0 : h->rank; a.world = h->src_slice ? 1 : h->world;
    a.n_local_chunks = gicp_local_chunks(h->src_valid, a.rank, a.world);
    a.radius2 = h->prm.max_correspondence_distance * h->prm.max_correspondence_distance;
    a.a = 1.0 - h->prm.epsilon;
    a.partials = h->partials.as<double>();
    a.corr = corr;
    a.prev = nullptr;
    a.st = h->state.as<GicpState>();
    a.mode = mode;
    for (int p = 0; p < GICP_MAX_WORLD; p++) a.peers[p] = h->peer_ptr[p];
    a.seq_base = h->peer_seq; a.xrank = h->peer_rank; a.xworld = h->peer_world;
    const uint32_t chunks = a.n_local_chunks;
    h->grid_blocks = (int)std::max<uint32_t>(1u, std::min<uint32_t>((chunks + GICP_WARPS - 1) / GICP_WARPS, (uint32_t)(device_sm_count() * h->blocks_per_sm)));
}

// The coarse grid serves queries whose nearest neighbour the fine 3x3x3 block cannot certify: cell edge =
// max_correspondence_distance * (1 + 2^-7), so its 3x3x3 block holds every point within the radius.
static int gicp_ensure_coarse(b2_gicp_s* h) {
    const double r = h->prm.max_correspondence_distance;
    if (h->coarse_for == r) return B2_OK;
    h->have_coarse = false;
    if (h->n_tgt && r > h->tgt_grid.dev.h * 0.999) {
        B2_CHECK(h->tgt_coarse.build(h->tgt_xyz.as<double>(), h->n_tgt, r * 1.0078125, 0.0, h->stream));
        B2_CHECK(h->tgt_coarse.build_cell_records(h->stream));
        h->have_coarse = true;
    } else h->tgt_coarse.release();
    h->coarse_for = r;
    return B2_OK;
}

static int gicp_prepare(b2_gicp_s* h) {
    if (!h->have_tgt || !h->have_src) { set_error("gicp: set_target and set_source first"); return B2_ERR_STATE; }
    B2_CHECK(h->partials.reserve((size_t)device_sm_count() * h->blocks_per_sm * GICP_NSUM * 8 + 256));
    B2_CHECK(gicp_ensure_coarse(h));
    B2_CHECK(h->state.reserve(sizeof(GicpState)));
    B2_CHECK(h->pin.reserve(sizeof(GicpState)));
    return B2_OK;
}

static int gicp_upload_state(b2_gicp_s* h, const double T[16], int max_it) {
    GicpState* hs = h->pin.as<GicpState>();
    memset(hs, 0, sizeof(GicpState));
    memcpy(hs->T, T, 16 * 8);
    hs->rel_fit = h->prm.relative_fitness; hs->rel_rmse = h->prm.relative_rmse;
    hs->n_src_total = (double)h->n_src;
    hs->max_it = max_it;
    B2_CUDA(cudaMemcpyAsync(h->state.p, hs, sizeof(GicpState), cudaMemcpyHostToDevice, h->stream));
    return B2_OK;
}

extern "C" {

void b2_gicp_default_params(b2_gicp_params* p) {
    if (!p) return;
    p->max_correspondence_distance = 1.0;   /* Multi_LiCa config/params.yaml:52 */
    p->epsilon = 0.005;                     /* :58 (Calibration class default 1e-4, Calibration.py:99) */
    p->relative_fitness = 1e-7;             /* :59 */
    p->relative_rmse = 1e-7;                /* :60 */
    p->max_iteration = 100;                 /* :61 */
}

int b2_gicp_create(b2_gicp_t* out, const b2_gicp_params* params) {
    if (!out) return B2_ERR_ARG;
    *out = nullptr;
    b2_gicp_s* h = new b2_gicp_s();
    if (params) h->prm = *params; else b2_gicp_default_params(&h->prm);
    if (cudaGetDevice(&h->device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->e0) != cudaSuccess || cudaEventCreate(&h->e1) != cudaSuccess) {
        set_error("b2_gicp_create: %s", cudaGetErrorString(cudaGetLastError()));
        delete h; return B2_ERR_CUDA;
    }
    *out = h;
    return B2_OK;
}

int b2_gicp_destroy(b2_gicp_t h) {
    if (!h) return B2_OK;
    h->tgt_grid.release(); h->tgt_coarse.release(); h->src_grid.release(); h->tgt_xyz.release(); h->fine_pos_of.release();
    h->sub_xyz.release(); h->sub_nrm.release(); h->tgt_m.release(); h->src_m.release(); h->partials.release(); h->state.release(); h->corr.release(); h->prev.release(); h->pin.release();
    for (int p = 0; p < GICP_MAX_WORLD; p++)
        if (h->peer_ptr[p] && h->peer_ptr[p] != h->peer_own) cudaIpcCloseMemHandle(h->peer_ptr[p]);
    if (h->peer_own) cudaFree(h->peer_own);
    if (h->e0) cudaEventDestroy(h->e0);
    if (h->e1) cudaEventDestroy(h->e1);
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return B2_OK;
}

int b2_gicp_set_params(b2_gicp_t h, const b2_gicp_params* p) {
    if (!h || !p) return B2_ERR_ARG;
    h->prm = *p;
    return B2_OK;
}

static int gicp_index_points(b2_gicp_s* h, const double* d_xyz, const double* d_nrm, size_t n, GridD& grid, DevBuf& m, double ppc, uint32_t* n_valid) {
    B2_CHECK(grid.build(d_xyz, n, 0.0, ppc, h->stream));
    B2_CHECK(gicp_valid_count(grid, h->stream, n_valid));
    B2_CHECK(m.reserve(std::max<size_t>(n, 1) * 24));
    if (*n_valid) {
        k_gicp_eff_normals<<<(*n_valid + 255) / 256, 256, 0, h->stream>>>(grid.dev.pts, d_nrm, *n_valid, m.as<double>()); count_launch();
        B2_CUDA(cudaGetLastError());
    }
    B2_CUDA(cudaStreamSynchronize(h->stream));
    return B2_OK;
}

static int gicp_set_cloud(b2_gicp_s* h, b2_cloud_s* c, GridD& grid, DevBuf& m, double ppc, uint32_t* n_valid, size_t begin = 0, size_t end = (size_t)-1) {
    end = std::min(end, c->n); begin = std::min(begin, end);
    if (!c->has_normals) {
        set_error("gicp: the cloud has no normals; call b2_cloud_estimate_normals first (Calibration.py:327-328 does)");
        return B2_ERR_STATE;
    }
    B2_CUDA(cudaSetDevice(h->device));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    // [begin, end): the rows this handle indexes (a rank's slice of a sharded source); grid indices are relative to `begin`
    return gicp_index_points(h, c->xyz.as<double>() + 3 * begin, c->nrm.as<double>() + 3 * begin, end - begin, grid, m, ppc, n_valid);
}

// this rank's blocks of GICP_SHARD_CHUNKS * 32 consecutive points of the cloud's Morton order (block b -> rank b mod world), with
// their normals: spatially dense pieces spread evenly over the scene
__global__ void __launch_bounds__(256) k_gicp_take_blocks(const P4d* __restrict__ morton, uint32_t n, const double* __restrict__ nrm, int rank, int world,
                                                          uint32_t m, double* __restrict__ xyz_out, double* __restrict__ nrm_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t B = GICP_SHARD_CHUNKS * 32u;
    const uint32_t pos = ((i / B) * (uint32_t)world + (uint32_t)rank) * B + (i % B);
    if (pos >= n) return;                       // (cannot happen: m counts only existing points)
    double x, y, z; long long id;
    load_p4d(&morton[pos], x, y, z, id);
    xyz_out[3 * (size_t)i] = x; xyz_out[3 * (size_t)i + 1] = y; xyz_out[3 * (size_t)i + 2] = z;
    nrm_out[3 * (size_t)i] = nrm[3 * (size_t)id]; nrm_out[3 * (size_t)i + 1] = nrm[3 * (size_t)id + 1]; nrm_out[3 * (size_t)i + 2] = nrm[3 * (size_t)id + 2];
}

int b2_gicp_set_target(b2_gicp_t h, b2_cloud_t target) {
    B2_NVTX("b2_gicp_set_target");
    if (!h || !target) return B2_ERR_ARG;
    h->have_tgt = false;
    uint32_t nv = 0;
    static const double tppc = [] { const char* e = getenv("B2_GICP_TARGET_PPC"); double v = e ? atof(e) : 2.0; return v > 0 ? v : 2.0; }();
    B2_CHECK(gicp_set_cloud(h, target, h->tgt_grid, h->tgt_m, tppc, &nv));
    h->n_tgt = target->n;
    h->coarse_for = -1.0; h->have_coarse = false;
    // the handle keeps its own copy of the points (the caller may destroy the cloud; the coarse grid is built lazily
    // for the search radius in force) and the original-index -> fine-position map the coarse pass needs
    B2_CHECK(h->tgt_xyz.reserve(std::max<size_t>(target->n, 1) * 24));
    B2_CHECK(h->fine_pos_of.reserve(std::max<size_t>(target->n, 1) * 4));
    if (target->n) {
        B2_CUDA(cudaMemcpyAsync(h->tgt_xyz.p, target->xyz.p, target->n * 24, cudaMemcpyDeviceToDevice, h->stream));
        B2_CUDA(cudaMemsetAsync(h->fine_pos_of.p, 0xff, target->n * 4, h->stream));
        if (nv) { k_gicp_fine_pos<<<(nv + 255) / 256, 256, 0, h->stream>>>(h->tgt_grid.dev.pts, nv, h->fine_pos_of.as<uint32_t>()); count_launch(); }
        B2_CUDA(cudaGetLastError());
        B2_CUDA(cudaStreamSynchronize(h->stream));
    }
    h->have_tgt = true;
    return B2_OK;
}

int b2_gicp_set_source(b2_gicp_t h, b2_cloud_t source) {
    B2_NVTX("b2_gicp_set_source");
    if (!h || !source) return B2_ERR_ARG;
    h->have_src = false;
    uint32_t nv = 0;
    static const double sppc = [] { const char* e = getenv("B2_GICP_TARGET_PPC"); double v = e ? atof(e) : 2.0; return v > 0 ? v : 2.0; }();
    B2_CHECK(gicp_set_cloud(h, source, h->src_grid, h->src_m, sppc, &nv));
    h->n_src = source->n;
    h->src_valid = nv;
    h->src_slice = false;
    h->have_src = true;
    return B2_OK;
}

// Sharded registration with a sharded set-up: this rank indexes only rows [begin, end) of the source (its own 1/world of the
// cloud: the sort of the source grid, the largest part of set_source, shrinks with the number of ranks) and processes all of
// them; the 30 sums are all-reduced as before and the fitness is taken over the whole cloud. For clouds in random order (the
// fused map of config C5) row slices are as well balanced as the block-cyclic deal of the cell-sorted source.
int b2_gicp_set_source_slice(b2_gicp_t h, b2_cloud_t source, size_t begin, size_t end) {
    B2_NVTX("b2_gicp_set_source_slice");
    if (!h || !source || begin > end) return B2_ERR_ARG;
    h->have_src = false;
    uint32_t nv = 0;
    static const double sppc = [] { const char* e = getenv("B2_GICP_TARGET_PPC"); double v = e ? atof(e) : 2.0; return v > 0 ? v : 2.0; }();
    B2_CHECK(gicp_set_cloud(h, source, h->src_grid, h->src_m, sppc, &nv, begin, end));
    h->n_src = source->n;                      // fitness = correspondences / all source points
    h->src_valid = nv;
    h->src_slice = true;
    h->have_src = true;
    return B2_OK;
}

// Sharded registration, spatially dense shards: the source is dealt to the ranks in blocks of 4096 consecutive points of its
// Morton order (the BVH estimate_normals left in the cloud), block b -> rank b mod world; this rank copies its blocks out and
// indexes only them. Dense blocks keep the target accesses of a warp inside a few cells (row slices of a randomly ordered cloud
// thin every shard out to 1/world of the density: measured 0.86 ms per steady evaluation at 8 GPUs against 0.63 ms), the
// round-robin deal keeps the ranks balanced, and the source sort shrinks with the number of ranks.
int b2_gicp_set_source_blocks(b2_gicp_t h, b2_cloud_t source, int rank, int world) {
    B2_NVTX("b2_gicp_set_source_blocks");
    if (!h || !source || world < 1 || rank < 0 || rank >= world) return B2_ERR_ARG;
    if (!source->has_normals || source->bvh.dev.n == 0 || source->bvh.dev.n > source->n) {
        set_error("b2_gicp_set_source_blocks: the cloud needs normals from b2_cloud_estimate_normals[_sharded] (its Morton order is reused)");
        return B2_ERR_STATE;
    }
    B2_CUDA(cudaSetDevice(h->device));
    B2_CUDA(cudaStreamSynchronize(source->stream));
    h->have_src = false;
    const uint32_t n = source->bvh.dev.n, B = GICP_SHARD_CHUNKS * 32u;
    const uint32_t nblocks = (n + B - 1) / B;
    uint64_t m = 0;
    for (uint32_t b = (uint32_t)rank; b < nblocks; b += (uint32_t)world) m += std::min<uint64_t>(B, (uint64_t)n - (uint64_t)b * B);
    B2_CHECK(h->sub_xyz.reserve(std::max<uint64_t>(m, 1) * 24));
    B2_CHECK(h->sub_nrm.reserve(std::max<uint64_t>(m, 1) * 24));
    if (m) {
        k_gicp_take_blocks<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(source->bvh.dev.pts, n, source->nrm.as<double>(), rank, world, (uint32_t)m,
                                                                               h->sub_xyz.as<double>(), h->sub_nrm.as<double>()); count_launch();
        B2_CUDA(cudaGetLastError());
    }
    uint32_t nv = 0;
    static const double sppc = [] { const char* e = getenv("B2_GICP_TARGET_PPC"); double v = e ? atof(e) : 2.0; return v > 0 ? v : 2.0; }();
    B2_CHECK(gicp_index_points(h, h->sub_xyz.as<double>(), h->sub_nrm.as<double>(), (size_t)m, h->src_grid, h->src_m, sppc, &nv));
    h->n_src = source->n;
    h->src_valid = nv;
    h->src_slice = true;
    h->have_src = true;
    return B2_OK;
}

int b2_gicp_set_shard(b2_gicp_t h, int rank, int world, b2_comm_t comm) {
    if (!h || world < 1 || rank < 0 || rank >= world) return B2_ERR_ARG;
    if (comm && (comm->rank != rank || comm->world != world)) { set_error("gicp: communicator rank/world mismatch"); return B2_ERR_ARG; }
    h->rank = rank; h->world = world; h->comm = world > 1 ? comm : nullptr;
    return B2_OK;
}

// Fused exchange set-up. b2_gicp_peer_handle allocates this rank's exchange area and returns its CUDA IPC handle (64 bytes); the
// host program gathers the handles of all ranks (as it distributes the NCCL id) and passes them to b2_gicp_set_peers, which
// maps the peers' areas. From then on b2_gicp_align exchanges the 30 sums inside k_gicp_linearize instead of calling NCCL.
int b2_gicp_peer_handle(b2_gicp_t h, unsigned char handle[64]) {
    if (!h || !handle) return B2_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    B2_CUDA(cudaSetDevice(h->device));
    if (!h->peer_own) {
        B2_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->peer_own), sizeof(GicpPeerArea)));
        B2_CUDA(cudaMemset(h->peer_own, 0, sizeof(GicpPeerArea)));
    }
    cudaIpcMemHandle_t ipc;
    B2_CUDA(cudaIpcGetMemHandle(&ipc, h->peer_own));
    memcpy(handle, &ipc, 64);
    return B2_OK;
}

int b2_gicp_set_peers(b2_gicp_t h, int rank, int world, const unsigned char* handles) {
    if (!h || !handles || world < 2 || world > GICP_MAX_WORLD || rank < 0 || rank >= world) { set_error("b2_gicp_set_peers: bad argument"); return B2_ERR_ARG; }
    if (!h->peer_own) { set_error("b2_gicp_set_peers: call b2_gicp_peer_handle first"); return B2_ERR_STATE; }
    B2_CUDA(cudaSetDevice(h->device));
    h->peer_world = 0;
    for (int p = 0; p < world; p++) {
        if (p == rank) { h->peer_ptr[p] = h->peer_own; continue; }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, handles + 64 * (size_t)p, 64);
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("b2_gicp_set_peers: rank %d cannot map rank %d's exchange area (%s); the NCCL all-reduce stays in use", rank, p, cudaGetErrorString(e));
            return B2_ERR_CUDA;
        }
        h->peer_ptr[p] = static_cast<GicpPeerArea*>(ptr);
    }
    h->peer_rank = rank; h->peer_world = world; h->peer_seq = 0;
    return B2_OK;
}

// The two calls above in one, for callers that have no side channel of their own: the 64-byte handles of all ranks are
// all-gathered over the registration's communicator (the one given to b2_gicp_set_shard; rank and world are its own).
int b2_gicp_exchange_setup(b2_gicp_t h) {
    if (!h) return B2_ERR_ARG;
    if (!h->comm || h->world < 2) { set_error("b2_gicp_exchange_setup: call b2_gicp_set_shard with a communicator first"); return B2_ERR_STATE; }
    if (h->world > GICP_MAX_WORLD) { set_error("b2_gicp_exchange_setup: at most %d ranks", GICP_MAX_WORLD); return B2_ERR_ARG; }
    unsigned char mine[64];
    B2_CHECK(b2_gicp_peer_handle(h, mine));
    DevBuf buf;
    B2_CHECK(buf.reserve((size_t)h->world * 64));
    unsigned char* d = buf.as<unsigned char>();
    B2_CUDA(cudaMemcpyAsync(d + 64 * (size_t)h->rank, mine, 64, cudaMemcpyHostToDevice, h->stream));
    B2_CHECK(comm_allgather_f64(h->comm, reinterpret_cast<const double*>(d + 64 * (size_t)h->rank), reinterpret_cast<double*>(d), 8, h->stream));
    std::vector<unsigned char> all((size_t)h->world * 64);
    B2_CUDA(cudaMemcpyAsync(all.data(), d, all.size(), cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    return b2_gicp_set_peers(h, h->rank, h->world, all.data());
}

int b2_gicp_linearize(b2_gicp_t h, const double T[16], double sums[30], int32_t* corr) {
    B2_NVTX("b2_gicp_linearize");
    if (!h || !T || !sums) return B2_ERR_ARG;
    B2_CUDA(cudaSetDevice(h->device));
    B2_CHECK(gicp_prepare(h));
    B2_CHECK(gicp_upload_state(h, T, 0));
    int32_t* d_corr = nullptr;
    if (corr && h->n_src) {
        B2_CHECK(h->corr.reserve(h->n_src * 4));
        d_corr = h->corr.as<int32_t>();
        k_fill_i32<<<(unsigned)((h->n_src + 255) / 256), 256, 0, h->stream>>>(d_corr, (uint32_t)h->n_src, -1); count_launch();
    }
    GicpArgs a;
    gicp_fill_args(h, a, 0, d_corr);
    k_gicp_linearize<<<h->grid_blocks, GICP_THREADS, 0, h->stream>>>(a); count_launch();
    B2_CUDA(cudaGetLastError());
    GicpState* ds = h->state.as<GicpState>();
    if (h->comm) B2_CHECK(comm_allreduce_sum_f64(h->comm, ds->sums_local, ds->sums_local, GICP_NSUM, h->stream));
    B2_CUDA(cudaMemcpyAsync(sums, ds->sums_local, GICP_NSUM * 8, cudaMemcpyDeviceToHost, h->stream));
    if (d_corr) B2_CUDA(cudaMemcpyAsync(corr, d_corr, h->n_src * 4, cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    return B2_OK;
}

int b2_gicp_align(b2_gicp_t h, const double init[16], double T_out[16], double* fitness, double* inlier_rmse,
                  int* iterations, int* converged) {
    B2_NVTX("b2_gicp_align");
    if (!h || !init || !T_out) return B2_ERR_ARG;
    if (h->world > 1 && !h->comm) { set_error("gicp: align on a sharded source needs a communicator (b2_gicp_set_shard)"); return B2_ERR_STATE; }
    B2_CUDA(cudaSetDevice(h->device));
    B2_CHECK(gicp_prepare(h));
    const int max_it = std::max(0, h->prm.max_iteration);
    B2_CHECK(gicp_upload_state(h, init, max_it));
    GicpArgs a;
    const bool fused = h->comm && h->peer_world == h->world && h->world > 1;
    gicp_fill_args(h, a, fused ? 2 : (h->comm ? 0 : 1), nullptr);
    if (fused) h->peer_seq += (unsigned)max_it + 2u;      // every rank advances alike (same parameters on all ranks)
    // the last correspondence of every source point seeds its next search (from the second evaluation on)
    if (h->src_valid && !getenv("B2_GICP_NO_SEED")) {
        const size_t slots = ((size_t)h->src_valid + 31) / 32 * 32;
        B2_CHECK(h->prev.reserve(slots * sizeof(uint32_t)));
        B2_CUDA(cudaMemsetAsync(h->prev.p, 0xff, slots * sizeof(uint32_t), h->stream));
        a.prev = h->prev.as<uint32_t>();
    }
    GicpState* ds = h->state.as<GicpState>();
    GicpState* hs = h->pin.as<GicpState>();
    // evaluations are enqueued in chunks; between chunks the host reads the small state back (one sync per chunk)
    const size_t per = std::max<size_t>(1, (size_t)a.n_local_chunks * 32);
    const int chunk = (int)std::min<size_t>(8, std::max<size_t>(1, 2000000 / per));
    int launched = 0, launches = 0;
    B2_CUDA(cudaEventRecord(h->e0, h->stream));
    const int total = max_it + 1;
    while ((int)h->ev.size() < std::min(total, GICP_HIST) + 1) { cudaEvent_t e; B2_CUDA(cudaEventCreate(&e)); h->ev.push_back(e); }
    h->ev_used = 0;
    B2_CUDA(cudaEventRecord(h->ev[0], h->stream));
    while (launched < total) {
        const int m = std::min(chunk, total - launched);
        for (int i = 0; i < m; i++) {
            k_gicp_linearize<<<h->grid_blocks, GICP_THREADS, 0, h->stream>>>(a); launches++;
#ifdef B2_NN1_STATS
            {
                unsigned long long z[8] = {}, v[8];
                cudaStreamSynchronize(h->stream); cudaMemcpyFromSymbol(v, g_nn1_stats, sizeof(v)); cudaMemcpyToSymbol(g_nn1_stats, z, sizeof(z));
                fprintf(stderr, "[nn1 stats] eval %d: fine cycles %.3g, coarse cycles %.3g (thread sums), coarse queries %llu, cells %llu, candidates %llu, exact %llu, unseeded %llu\n",
                        launched + i, (double)v[0], (double)v[1], v[2], v[5], v[3], v[4], v[6]);
            }
#endif
            if (h->ev_used + 1 < (int)h->ev.size()) { h->ev_used++; B2_CUDA(cudaEventRecord(h->ev[h->ev_used], h->stream)); }
            if (h->comm && !fused) {
                B2_CHECK(comm_allreduce_sum_f64(h->comm, ds->sums_local, ds->sums, GICP_NSUM, h->stream));
                k_gicp_finalize<<<1, 32, 0, h->stream>>>(ds); launches++;
            }
        }
        B2_CUDA(cudaGetLastError());
        launched += m;
        B2_CUDA(cudaMemcpyAsync(&hs->it, &ds->it, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        B2_CUDA(cudaStreamSynchronize(h->stream));
        if (hs->done) break;
    }
    B2_CUDA(cudaEventRecord(h->e1, h->stream));
    B2_CUDA(cudaMemcpyAsync(hs, ds, sizeof(GicpState), cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    count_launch(launches);
    cudaEventElapsedTime(&h->last_ms, h->e0, h->e1);
    h->last_launches = launches;
    memcpy(T_out, hs->T, 16 * 8);
    if (fitness) *fitness = hs->fitness;
    if (inlier_rmse) *inlier_rmse = hs->rmse;
    if (iterations) *iterations = hs->it;
    if (converged) *converged = hs->converged;
    return B2_OK;
}

int b2_gicp_get_history(b2_gicp_t h, double* fitness, double* inlier_rmse, int capacity, int* n_evaluations) {
    if (!h || capacity < 0) return B2_ERR_ARG;
    const GicpState* hs = h->pin.as<GicpState>();
    if (!hs) { set_error("gicp: no align call yet"); return B2_ERR_STATE; }
    const int n = std::min(std::min(hs->evals, GICP_HIST), capacity);
    for (int i = 0; i < n; i++) { if (fitness) fitness[i] = hs->hist_fit[i]; if (inlier_rmse) inlier_rmse[i] = hs->hist_rmse[i]; }
    if (n_evaluations) *n_evaluations = hs->evals;
    return B2_OK;
}

int b2_gicp_get_evaluation_ms(b2_gicp_t h, float* ms, int capacity, int* n_evaluations) {
    if (!h || capacity < 0) return B2_ERR_ARG;
    const GicpState* hs = h->pin.as<GicpState>();
    const int n = hs ? std::min(std::min(hs->evals, h->ev_used), capacity) : 0;
    for (int i = 0; i < n; i++) { ms[i] = 0.f; cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]); }
    if (n_evaluations) *n_evaluations = n;
    return B2_OK;
}

int b2_gicp_last_gpu_ms(b2_gicp_t h, float* ms, int* launches) {
    if (!h) return B2_ERR_ARG;
    if (ms) *ms = h->last_ms;
    if (launches) *launches = h->last_launches;
    return B2_OK;
}

int b2_gicp_index_info(b2_gicp_t h, double* target_cell_edge, double* target_points_per_cell, uint32_t* shard_points, uint32_t* shard_block_points) {
    if (!h || !h->have_tgt) return B2_ERR_STATE;
    if (target_cell_edge) *target_cell_edge = h->tgt_grid.dev.h;
    if (target_points_per_cell) *target_points_per_cell = h->tgt_grid.ppc;
    if (shard_points) {
        // points this rank evaluates (the last chunk of the cloud may be partial)
        const uint32_t total = (h->src_valid + 31) / 32;
        uint32_t pts = gicp_local_chunks(h->src_valid, h->rank, h->world) * 32u;
        const uint32_t last_block_owner = total ? ((total - 1) / GICP_SHARD_CHUNKS) % (uint32_t)h->world : 0u;
        if (total && (int)last_block_owner == h->rank) pts -= total * 32u - h->src_valid;
        *shard_points = pts;
    }
    if (shard_block_points) *shard_block_points = GICP_SHARD_CHUNKS * 32u;
    return B2_OK;
}

}  // extern "C"
