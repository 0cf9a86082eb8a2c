// Scan-to-map Levenberg-Marquardt on the device: ONE kernel per LM iteration, no host round trip in between.
//
// Replaces (liosam_ws/src/LIO-SAM/src/mapOptmization.cpp):
//   cornerOptimization :974-1064, surfOptimization :1066-1135, combineOptimizationCoeffs :1137-1156,
//   LMOptimization :1158-1280 and the loop of scan2MapOptimization :1282-1310.
//
// k_s2m_iteration, grid = (feature blocks, scans in the batch), 256 threads:
//   phase 1  32 features per CTA, 8 lanes per feature: transform (pointAssociateToMap :278-284), exact 5-NN in
//            the uniform grid (b2_grid.cuh), winners' coordinates parked in shared memory;
//   phase 2  warp 0, one lane per feature: line fit (3x3 Jacobi) or plane fit (5x3 pivoted QR), thresholds in the
//            reference's mixed float/double types, coeff, Euler-angle Jacobian row, 21+6 products + count in fp64,
//            xor-shuffle tree -> one 28-double partial per CTA;
//   epilogue the last CTA of each scan (atomic ticket) adds the partials in CTA order, rounds AtA/Atb to float as
//            cv::gemm does, Householder-QR solve, degeneracy projection on iteration 0, pose update, convergence
//            test, and prepares the next iteration's pose matrix and sines/cosines.
#include "b2_grid.cuh"
#include "b2_smallmath.cuh"
#include <cmath>
#include <vector>
#include <algorithm>
#include <mutex>

#ifndef B2_FIT_VARIANT
#define B2_FIT_VARIANT 1      // 0: rolled loops over per-thread arrays (small code); 1: fully unrolled, register resident
#endif
#ifndef B2_FIT_INLINE
#define B2_FIT_INLINE 0       // 1: the two fits are inlined into the kernel (kernel experiments)
#endif
#if B2_FIT_INLINE
#define B2_FIT_ATTR __forceinline__
#else
#define B2_FIT_ATTR __noinline__
#endif

namespace b2 {

constexpr int S2M_THREADS = 256;
// two shapes of the one kernel: <lanes per feature, rounds>
#ifndef S2M_LAT_LPF_V
#define S2M_LAT_LPF_V 16
#define S2M_LAT_ROUNDS_V 2
#endif
#ifndef S2M_LAT_ROUNDS_C_V
#define S2M_LAT_ROUNDS_C_V 1
#endif
// single scan: 32 surf features per CTA, many small CTAs (latency). Corner CTAs make S2M_LAT_ROUNDS_C passes only: their line
// fit (3x3 Jacobi, data-dependent rotation count) is the longest per-thread chain of the kernel — measured 23 k cycles against
// 13 k for the plane fit — so they get through their neighbour search in half the time and start fitting early.
constexpr int S2M_LAT_LPF = S2M_LAT_LPF_V, S2M_LAT_ROUNDS = S2M_LAT_ROUNDS_V, S2M_LAT_ROUNDS_C = S2M_LAT_ROUNDS_C_V;
#ifndef S2M_THR_LPF_V
#define S2M_THR_LPF_V 1
#define S2M_THR_ROUNDS_V 1
#endif
#ifndef S2M_THR_MINB
#define S2M_THR_MINB 4
#endif
constexpr int S2M_THR_LPF = S2M_THR_LPF_V, S2M_THR_ROUNDS = S2M_THR_ROUNDS_V;   // batch: 256 features per CTA, every warp busy in phase 2 (throughput)
constexpr int S2M_LAT_FPB = S2M_THREADS / S2M_LAT_LPF * S2M_LAT_ROUNDS;
constexpr int S2M_LAT_FPB_C = S2M_THREADS / S2M_LAT_LPF * S2M_LAT_ROUNDS_C;
constexpr int S2M_THR_FPB = S2M_THREADS / S2M_THR_LPF * S2M_THR_ROUNDS;
constexpr int S2M_NPART = 28;                       // 21 (AtA upper) + 6 (Atb) + 1 (count)

struct S2MState {                 // one per scan in the batch, lives in HBM
    float pose[6];                // roll pitch yaw x y z
    float xf[12];                 // pcl::getTransformation(x,y,z,roll,pitch,yaw), rows
    float trig[6];                // srx crx sry cry srz crz of LMOptimization :1170-1175
    float matP[36];
    float AtA[36], AtB[6], X[6];
    int degenerate;
    int done;                     // no further iterations for this scan
    int converged;
    int iters;                    // LM iterations executed
    int n_sel;                    // laserCloudSelNum of the last iteration
    int ran;                      // 0 when n_sel < min_correspondences
    unsigned ticket;
    int grid_status;              // status words of the two device-sized map grids (0, or B2_ERR_TOO_LARGE: rebuild on the host path)
    long long tl[8];              // %globaltimer stamps of a persistent solve: [0] first CTA in, [1 + k] end of iteration k (k < 6), [7] result written
};

struct S2MInit {                  // initial state of a single-scan solve, carried by the first launch's arguments
    float pose[6], xf[12], trig[6], matP[36];
    int degenerate;
};

struct S2MArgs {
    const GridDevMem* gcm; const GridDevMem* gsm; // corner / surf map grids (geometry in device memory: the device sized them)
    const float4* scan_c; const float4* scan_s;   // packed xyzi features, all scans concatenated
    const int* off_c; const int* off_s;           // batch+1 offsets
    S2MState* st;
    double* partial;                              // [batch][max_blocks][28]
    int max_blocks;
    int iter;                                     // iterCount for host-driven single iteration; -1 = use st->iters
    int device_driven;                            // epilogue prepares xf/trig for the next iteration
    int max_iters;                                // loop bound of scan2MapOptimization (device-driven runs)
    int min_corr; float eig_thr;
    int want_matP;                                // 0: iteration 0 may skip the 6x6 eigen-decomposition when the system is
                                                  //    certified non-degenerate (matP is then left untouched)
    int* done_count;                              // number of scans whose loop has ended (host polls it between chunks)
    uint32_t* nb_c; uint32_t* nb_s;               // [feature][5] grid positions of the 5 winners of the previous iteration
    int use_prev;                                 // 1: nb_* hold the previous iteration of this scan/map; -1: they do when st.iters > 0; 0: no
    int single;                                   // 1: one scan, its feature counts ride in the arguments (no offset table to upload or read)
    int single_nc, single_ns;
    int use_init;                                 // 1: first launch of a single-scan solve, state = init (no upload, no state loads)
    S2MInit init;
    S2MState* result; volatile int* result_seq; int seq; int chunk_last;   // mapped host memory: state + sequence flag (single scan)
    int persistent_iters;                         // > 0: ONE cooperative launch runs up to this many LM iterations (single scan); the
    int* iter_flag; int flag_base;                //      last CTA of an iteration publishes flag_base + k, the other CTAs wait for it
    unsigned long long* cand_count;               // optional: candidates loaded by the neighbour search, summed over the launch
    long long* prof;                              // optional per-CTA clock64 stamps (B2_S2M_PROF=1), 16 per CTA
    float* pose_hist; int hist_stride;            // optional [batch][max_iters][6]
    // optional per-feature introspection (single-scan parity runs)
    int32_t* dbg_idx_c; float* dbg_d2_c; float4* dbg_coeff_c; uint8_t* dbg_flag_c;
    int32_t* dbg_idx_s; float* dbg_d2_s; float4* dbg_coeff_s; uint8_t* dbg_flag_s;
};

// sin/cos of a float angle, rounded once from double: agrees with a correctly rounded sinf/cosf
__device__ __forceinline__ float sin_rn(float a) { return (float)sin((double)a); }
__device__ __forceinline__ float cos_rn(float a) { return (float)cos((double)a); }

__host__ __device__ inline void affine_from_trig(float x, float y, float z, float A, float B, float C, float D, float E, float F, float* t) {
    // A=cos yaw B=sin yaw C=cos pitch D=sin pitch E=cos roll F=sin roll (PCL getTransformation operation order)
    float DE = D * E, DF = D * F;
    t[0] = A * C;  t[1] = A * DF - B * E;  t[2]  = B * F + A * DE;  t[3]  = x;
    t[4] = B * C;  t[5] = A * E + B * DF;  t[6]  = B * DE - A * F;  t[7]  = y;
    t[8] = -D;     t[9] = C * F;           t[10] = C * E;           t[11] = z;
}

__device__ void prepare_pose_device(S2MState& s, const float* pose) {
    const float roll = pose[0], pitch = pose[1], yaw = pose[2];
    const float sr = sin_rn(roll), cr = cos_rn(roll), sp = sin_rn(pitch), cp = cos_rn(pitch), sy = sin_rn(yaw), cy = cos_rn(yaw);
    affine_from_trig(pose[3], pose[4], pose[5], cy, sy, cp, sp, cr, sr, s.xf);
    s.trig[0] = sp; s.trig[1] = cp;     // srx crx <- pitch
    s.trig[2] = sy; s.trig[3] = cy;     // sry cry <- yaw
    s.trig[4] = sr; s.trig[5] = cr;     // srz crz <- roll
}

__global__ void k_s2m_prepare(S2MState* st, int batch, const int* off_c, const int* off_s, int edge_min, int surf_min, int* not_enough, int* done_count) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    S2MState& s = st[b];
    prepare_pose_device(s, s.pose);
    s.done = 0; s.converged = 0; s.iters = 0; s.n_sel = 0; s.ran = 0; s.ticket = 0;
    const int nc = off_c[b + 1] - off_c[b], ns = off_s[b + 1] - off_s[b];
    const int ne = !(nc > edge_min && ns > surf_min);      // guard of scan2MapOptimization :1287
    if (ne) { s.done = 1; if (done_count) atomicAdd(done_count, 1); }
    if (not_enough) not_enough[b] = ne;
}

// ---- per-feature fits -------------------------------------------------------------------------------------------
// Returns true when the feature is kept; coeff = coeffSel entry.
__device__ __forceinline__ bool fit_line_impl(const float (&nx)[5], const float (&ny)[5], const float (&nz)[5],
                                         float x0, float y0, float z0, float4& coeff) {
    float cx = 0, cy = 0, cz = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) { cx += nx[j]; cy += ny[j]; cz += nz[j]; }
    cx /= 5; cy /= 5; cz /= 5;
    float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        float ax = nx[j] - cx, ay = ny[j] - cy, az = nz[j] - cz;
        a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
        a22 += ay * ay; a23 += ay * az;
        a33 += az * az;
    }
    a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
    float w[3], v[9];
#if B2_FIT_VARIANT == 1
    sym_eigen_jacobi3(a11, a12, a13, a22, a23, a33, w, v);
#else
    float m[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33};
    sym_eigen_jacobi_c<3>(m, w, v);
#endif
    if (!(w[0] > 3 * w[1])) return false;
    // two points on the line, 0.1 either side of the centroid (double literal: evaluated in double, narrowed)
    float x1 = cx + 0.1 * v[0], y1 = cy + 0.1 * v[1], z1 = cz + 0.1 * v[2];
    float x2 = cx - 0.1 * v[0], y2 = cy - 0.1 * v[1], z2 = cz - 0.1 * v[2];
    float mxy = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
    float mxz = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
    float myz = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
    float a012 = sqrtf(mxy * mxy + mxz * mxz + myz * myz);
    float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
    float la = ((y1 - y2) * mxy + (z1 - z2) * mxz) / a012 / l12;
    float lb = -((x1 - x2) * mxy - (z1 - z2) * myz) / a012 / l12;
    float lc = -((x1 - x2) * mxz + (y1 - y2) * myz) / a012 / l12;
    float ld2 = a012 / l12;
    float s = 1 - 0.9 * fabsf(ld2);
    coeff = make_float4(s * la, s * lb, s * lc, s * ld2);
    return s > 0.1;
}

__device__ __forceinline__ bool fit_plane_impl(const float (&nx)[5], const float (&ny)[5], const float (&nz)[5],
                                          float sx, float sy, float sz, float ox, float oy, float oz, float4& coeff) {
#if B2_FIT_VARIANT == 1
    float pa, pb, pc, pd = 1;
    plane_lsq_5x3(nx, ny, nz, pa, pb, pc);
#else
    float q[15], x3[3];
#pragma unroll
    for (int j = 0; j < 5; j++) { q[j * 3 + 0] = nx[j]; q[j * 3 + 1] = ny[j]; q[j * 3 + 2] = nz[j]; }
    plane_lsq_5x3_c(q, x3);
    float pa = x3[0], pb = x3[1], pc = x3[2], pd = 1;
#endif
    float ps = sqrtf(pa * pa + pb * pb + pc * pc);
    pa /= ps; pb /= ps; pc /= ps; pd /= ps;
#pragma unroll
    for (int j = 0; j < 5; j++)
        if (fabsf(pa * nx[j] + pb * ny[j] + pc * nz[j] + pd) > 0.2) return false;
    float pd2 = pa * sx + pb * sy + pc * sz + pd;
    float s = 1 - 0.9 * fabsf(pd2) / sqrtf(sqrtf(ox * ox + oy * oy + oz * oz));
    coeff = make_float4(s * pa, s * pb, s * pc, s * pd2);
    return s > 0.1;
}

// Out-of-line copies for the single-scan kernel (one warp walks the fits once per launch: code size is what it pays for); the
// batched fit kernel inlines them, so the neighbour coordinates stay in registers instead of going through local memory.
__device__ B2_FIT_ATTR bool fit_line(const float (&nx)[5], const float (&ny)[5], const float (&nz)[5], float x0, float y0, float z0, float4& coeff) {
    return fit_line_impl(nx, ny, nz, x0, y0, z0, coeff);
}
__device__ B2_FIT_ATTR bool fit_plane(const float (&nx)[5], const float (&ny)[5], const float (&nz)[5],
                                      float sx, float sy, float sz, float ox, float oy, float oz, float4& coeff) {
    return fit_plane_impl(nx, ny, nz, sx, sy, sz, ox, oy, oz, coeff);
}

// ---- epilogue: normal equations -> pose update (single thread) ------------------------------------------------------
__device__ __noinline__ void lm_epilogue(S2MState& s, const double* sums, int iterCount, const S2MArgs& a, int scan, long long* prof) {
    // fields another SM may have written in an earlier iteration of a persistent launch are read with ld.cg
    const int iters0 = __ldcg(&s.iters);
    int degen = __ldcg(&s.degenerate);
    const int K = (int)(sums[27] + 0.5);
    s.n_sel = K;
    if (K < a.min_corr) {
        // LMOptimization returns false before touching the pose (:1178-1180): every later iteration would repeat
        // this one exactly, so the loop is over; iters counts the iterations the reference would have spent.
        s.ran = 0;
        if (a.device_driven) {
            if (a.pose_hist)
                for (int it = iters0; it < a.max_iters; it++)
                    for (int i = 0; i < 6; i++) a.pose_hist[((size_t)scan * a.hist_stride + it) * 6 + i] = __ldcg(&s.pose[i]);
            s.iters = a.max_iters;
            s.done = 1;
            if (a.done_count) atomicAdd(a.done_count, 1);
        } else {
            s.iters = iters0 + 1;
        }
        return;
    }
    s.ran = 1;
    float AtA[36], AtB[6], X[6];
    {
        int q = 0;
        for (int r = 0; r < 6; r++)
            for (int c = r; c < 6; c++) { float f = (float)sums[q++]; AtA[r * 6 + c] = f; AtA[c * 6 + r] = f; }
        for (int r = 0; r < 6; r++) AtB[r] = (float)sums[21 + r];
    }
    for (int i = 0; i < 36; i++) s.AtA[i] = AtA[i];
    for (int i = 0; i < 6; i++) s.AtB[i] = AtB[i];
    {
        float w[36];
        for (int i = 0; i < 36; i++) w[i] = AtA[i];
        for (int i = 0; i < 6; i++) X[i] = AtB[i];
        if (!solve_householder<6>(w, X)) for (int i = 0; i < 6; i++) X[i] = 0.f;
    }
    if (prof) prof[8] = clock64();
    bool full_eigen = iterCount == 0;
    if (full_eigen && !a.want_matP) {
        // Degeneracy needs only "is the smallest eigenvalue below the threshold". LDL^T of (AtA - shift*I) in fp64 with
        // shift = threshold + 1e-4*trace has all-positive pivots iff lambda_min > shift; the margin is far above the
        // error of a float Jacobi sweep (~1e-6*||AtA||), so a certified system is non-degenerate for the reference's
        // cv::eigen as well. Anything closer to the threshold takes the literal path below.
        double Ld[36];
        double tr = 0.0;
        for (int i = 0; i < 6; i++) tr += (double)AtA[i * 6 + i];
        const double shift = (double)a.eig_thr + 1e-4 * tr;
        bool pd = true;
        for (int j = 0; j < 6 && pd; j++) {
            double d = (double)AtA[j * 6 + j] - shift;
            for (int q = 0; q < j; q++) d -= Ld[j * 6 + q] * Ld[j * 6 + q] * Ld[q * 6 + q];
            if (!(d > 0.0)) { pd = false; break; }
            Ld[j * 6 + j] = d;
            for (int i = j + 1; i < 6; i++) {
                double v = (double)AtA[i * 6 + j];
                for (int q = 0; q < j; q++) v -= Ld[i * 6 + q] * Ld[j * 6 + q] * Ld[q * 6 + q];
                Ld[i * 6 + j] = v / d;
            }
        }
        if (pd) { s.degenerate = 0; degen = 0; full_eigen = false; }
    }
    if (full_eigen) {
        float w[36], E[6], V[36], V2[36], Vi[36];
        for (int i = 0; i < 36; i++) w[i] = AtA[i];
        sym_eigen_jacobi_c<6>(w, E, V);
        for (int i = 0; i < 36; i++) V2[i] = V[i];
        int deg = 0;
        for (int i = 5; i >= 0; i--) {
            if (E[i] < a.eig_thr) { for (int j = 0; j < 6; j++) V2[i * 6 + j] = 0.f; deg = 1; }
            else break;
        }
        s.degenerate = deg; degen = deg;
        for (int i = 0; i < 36; i++) w[i] = V[i];
        invert_lu<6>(w, Vi);
        matmul_dacc<6, 6>(Vi, V2, s.matP);
    }
    if (prof) prof[9] = clock64();
    if (degen) {
        float X2[6], P[36];
        for (int i = 0; i < 6; i++) X2[i] = X[i];
        for (int i = 0; i < 36; i++) P[i] = __ldcg(&s.matP[i]);
        matmul_dacc<6, 1>(P, X2, X);
    }
    float pose[6];
    for (int i = 0; i < 6; i++) { s.X[i] = X[i]; pose[i] = __ldcg(&s.pose[i]) + X[i]; s.pose[i] = pose[i]; }
    const float d0 = X[0] * 57.29578f, d1 = X[1] * 57.29578f, d2 = X[2] * 57.29578f;
    const float t0 = X[3] * 100, t1 = X[4] * 100, t2 = X[5] * 100;
    const float deltaR = (float)sqrt((double)d0 * (double)d0 + (double)d1 * (double)d1 + (double)d2 * (double)d2);
    const float deltaT = (float)sqrt((double)t0 * (double)t0 + (double)t1 * (double)t1 + (double)t2 * (double)t2);
    const int conv = (deltaR < 0.05 && deltaT < 0.05) ? 1 : 0;
    s.converged = conv;
    if (a.pose_hist) {
        float* ph = a.pose_hist + ((size_t)scan * a.hist_stride + iters0) * 6;
        for (int i = 0; i < 6; i++) ph[i] = pose[i];
    }
    s.iters = iters0 + 1;
    if (prof) prof[10] = clock64();
    if (a.device_driven) {
        if (conv) { s.done = 1; if (a.done_count) atomicAdd(a.done_count, 1); }
        else prepare_pose_device(s, pose);
    }
}

// LPF lanes serve one feature in phase 1; a CTA of 256 threads makes ROUNDS passes, so it owns
// FPB = 256 / LPF * ROUNDS features and phase 2 runs on FPB threads (one warp for the latency shape <16,2>,
// all eight warps for the throughput shape <8,8>).
// Register budget: the latency shape must keep two CTAs on an SM (204 CTAs on 148 SMs: at 128 registers the occupancy
// calculator grants one, and the kernel ran as two waves — 38 us instead of 23), the throughput shape four.
#ifndef S2M_LAT_MAXREG
#define S2M_LAT_MAXREG 120
#endif
// PHASES: 3 = neighbour search and fits in one kernel (single scan, latency); the batched shape runs them as two launches with
// their own register budgets — 1 = search only (winners' grid positions to HBM, 20 B per feature), 2 = fits, Jacobian rows,
// reduction and epilogue from those positions. Fused at 64 registers the fits spilled ~440 MB of local memory per launch.
#ifndef B2_S2M_ARGS_QUAL
#define B2_S2M_ARGS_QUAL __grid_constant__
#endif
#ifndef S2M_THR_KNN_MAXREG
#define S2M_THR_KNN_MAXREG 64
#endif
#ifndef S2M_THR_FIT_MAXREG
#define S2M_THR_FIT_MAXREG 80
#endif
template <int LPF, int ROUNDS, int ROUNDS_C, int PHASES>
__global__ void __launch_bounds__(S2M_THREADS) __maxnreg__(PHASES == 1 ? S2M_THR_KNN_MAXREG : (PHASES == 2 ? S2M_THR_FIT_MAXREG : S2M_LAT_MAXREG))
k_s2m_iteration(const B2_S2M_ARGS_QUAL S2MArgs a) {   // __grid_constant__: lm_epilogue takes `a` by reference; without it every
                                                        // thread copies the 512-byte argument block to its local-memory stack first
    constexpr int FPR = S2M_THREADS / LPF;
    constexpr int FPB = FPR * ROUNDS;                 // features per surf CTA (and the size of the shared arrays)
    constexpr int FPB_C = FPR * ROUNDS_C;             // features per corner CTA
    static_assert(FPB <= S2M_THREADS && FPB % 32 == 0 && ROUNDS_C <= ROUNDS, "phase 2 maps one thread per feature in whole warps");
    constexpr int P2_WARPS = FPB / 32;
    // Programmatic dependent launch (single-scan shape): this grid may have been scheduled while the previous iteration was
    // still running; wait for it (and its writes) here, and let the next launch be scheduled behind us right away. Both are
    // no-ops for a launch without the attribute.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    const int scan = blockIdx.y;
    S2MState& st = a.st[scan];
    bool first = a.use_init != 0;                              // the state is in the launch arguments, not in memory yet
    if (first && a.persistent_iters > 0 && blockIdx.x == 0 && threadIdx.x == 0) { long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); st.tl[0] = gt; }
    if (!first && st.done) return;                             // uniform per CTA, written only by a previous launch
    if (first) {
        // a map that outgrew its device-sized cell table has no usable geometry: report it and let the host rebuild (the
        // host-driven and batched paths check the status before they launch)
        const int gst = a.gcm->status | a.gsm->status;
        if (gst != 0) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                st.grid_status = gst; st.done = 1; st.iters = 0; st.converged = 0;
                if (a.result) {
                    a.result->grid_status = gst; a.result->done = 1; a.result->iters = 0;
                    __threadfence_system();
                    *a.result_seq = a.seq;
                    __threadfence_system();
                }
            }
            return;
        }
    }
    const int c0 = a.single ? 0 : a.off_c[scan], nc = a.single ? a.single_nc : a.off_c[scan + 1] - c0;
    const int s0 = a.single ? 0 : a.off_s[scan], ns = a.single ? a.single_ns : a.off_s[scan + 1] - s0;
    const int nbc = (nc + FPB_C - 1) / FPB_C, nbs = (ns + FPB - 1) / FPB;
    const int nblk = nbc + nbs;
    if ((int)blockIdx.x >= nblk) return;
    const bool is_surf = (int)blockIdx.x >= nbc;
    const int fb = is_surf ? ((int)blockIdx.x - nbc) * FPB : (int)blockIdx.x * FPB_C;   // first feature of this CTA
    const int rounds = is_surf ? ROUNDS : ROUNDS_C;
    const int fpb = is_surf ? FPB : FPB_C;
    const int nfeat = is_surf ? ns : nc;
    const float4* scanp = is_surf ? (a.scan_s + s0) : (a.scan_c + c0);
    // the pose moves by millimetres between LM iterations: the previous winners bound this iteration's search radius
    uint32_t* nbp = (is_surf ? a.nb_s + (size_t)s0 * 5 : a.nb_c + (size_t)c0 * 5);
    const bool use_prev = LPF < 8 && (a.use_prev > 0 || (a.use_prev < 0 && !first && st.iters > 0));
    // Persistent form (single scan, cooperative launch, every CTA resident): the LM loop runs inside the kernel. An iteration
    // ends when its last CTA (atomic ticket) has solved the normal equations and published the iteration number; the other
    // CTAs wait for that number instead of for a kernel boundary (no launch gap, no empty launches after convergence, and no
    // guess at how many launches to enqueue). State written by another SM is read with ld.cg (L1 is not coherent).
    const int n_pit = a.persistent_iters > 0 ? a.persistent_iters : 1;

    __shared__ float4 s_nb[FPB][5];           // winners: x y z, original index bits
    __shared__ float s_d2[FPB][5];
    __shared__ float4 s_ori[FPB], s_sel[FPB]; // pointOri (xyzi), pointSel (xyz)
    __shared__ float s_xf[12], s_trig[6];
    __shared__ int s_last;
    __shared__ double s_wsum[P2_WARPS][S2M_NPART];
    __shared__ double s_red[S2M_THREADS / S2M_NPART][S2M_NPART];
    __shared__ double s_sum[S2M_NPART];

    long long* prof = a.prof ? a.prof + ((size_t)scan * a.max_blocks + blockIdx.x) * 16 : nullptr;
    if (prof && threadIdx.x == 0) { prof[0] = clock64(); long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); prof[6] = gt; }
#pragma unroll 1
  for (int pit = 0; pit < n_pit; pit++) {
    if (threadIdx.x < 12) s_xf[threadIdx.x] = first ? a.init.xf[threadIdx.x] : __ldcg(&st.xf[threadIdx.x]);
    else if (threadIdx.x < 18) s_trig[threadIdx.x - 12] = first ? a.init.trig[threadIdx.x - 12] : __ldcg(&st.trig[threadIdx.x - 12]);
    __syncthreads();

    // ---------------- phase 1: transform + 5-NN, LPF lanes per feature
    const GridDev g = is_surf ? a.gsm->g : a.gcm->g;          // (loaded per iteration: 14 registers that need not live across the loop)
    const int grp = threadIdx.x / LPF, sub = threadIdx.x & (LPF - 1);
    if constexpr (PHASES == 2) {
        // the search ran in its own launch: winners come back from HBM by position, distances are recomputed with the search's
        // own expression (same bits), everything lands in shared memory where the fused kernel would have put it
        const int slot = threadIdx.x, f = fb + slot;
        if (slot < fpb && f < nfeat) {
            const float4 po = __ldg(&scanp[f]);
            const float sx = s_xf[0] * po.x + s_xf[1] * po.y + s_xf[2] * po.z + s_xf[3];
            const float sy = s_xf[4] * po.x + s_xf[5] * po.y + s_xf[6] * po.z + s_xf[7];
            const float sz = s_xf[8] * po.x + s_xf[9] * po.y + s_xf[10] * po.z + s_xf[11];
            uint32_t pp[5];
#pragma unroll
            for (int j = 0; j < 5; j++) pp[j] = __ldcg(&nbp[(size_t)f * 5 + j]);
#pragma unroll
            for (int j = 0; j < 5; j++) {
                float4 c = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
                float d = INFINITY;
                if (pp[j] != 0xffffffffu) {
                    c = ldg4(&g.pts[pp[j]]);
                    const float dx = sx - c.x, dy = sy - c.y, dz = sz - c.z;
                    d = dx * dx;
                    d = d + dy * dy;
                    d = d + dz * dz;
                }
                s_nb[slot][j] = c;
                s_d2[slot][j] = d;
            }
            s_ori[slot] = po;
            s_sel[slot] = make_float4(sx, sy, sz, 0.f);
        }
    } else {
#pragma unroll 1
    for (int r = 0; r < rounds; r++) {
        const int slot = r * FPR + grp;
        const int f = fb + slot;
        const bool active = f < nfeat;
        float4 po = make_float4(0.f, 0.f, 0.f, 0.f);
        float sx = 0.f, sy = 0.f, sz = 0.f;
        if (active) {
            po = __ldg(&scanp[f]);
            sx = s_xf[0] * po.x + s_xf[1] * po.y + s_xf[2] * po.z + s_xf[3];
            sy = s_xf[4] * po.x + s_xf[5] * po.y + s_xf[6] * po.z + s_xf[7];
            sz = s_xf[8] * po.x + s_xf[9] * po.y + s_xf[10] * po.z + s_xf[11];
        }
        float bound2 = INFINITY;
        if (use_prev) {                                        // uniform per CTA
            float m = 0.f;
            for (int j = sub; j < 5; j += LPF) {
                const uint32_t pp = active ? __ldcg(&nbp[(size_t)f * 5 + j]) : 0xffffffffu;
                if (pp == 0xffffffffu) m = INFINITY;
                else {
                    const float4 c = ldg4(&g.pts[pp]);
                    const float dx = sx - c.x, dy = sy - c.y, dz = sz - c.z;
                    float d = dx * dx;                         // the search's own expression: these five pass its d <= bound test
                    d = d + dy * dy;
                    d = d + dz * dz;
                    m = fmaxf(m, d);
                }
            }
            __syncwarp();
#pragma unroll
            for (int o = LPF >> 1; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            bound2 = m;
        }
        unsigned long long key[5]; uint32_t pos[5];
        if constexpr (PHASES == 1) {
            uint32_t visited = 0;
            knn_group<5, LPF>(g, sx, sy, sz, active, bound2, key, pos, &visited);
            if (a.cand_count) {                    // statistics run (bench.py): one atomic per warp
                __syncwarp();
                const uint32_t w = __reduce_add_sync(0xffffffffu, visited);
                if ((threadIdx.x & 31) == 0 && w) atomicAdd(a.cand_count, (unsigned long long)w);
            }
        } else {
            knn_group<5, LPF>(g, sx, sy, sz, active, bound2, key, pos);
        }
        if (active) {
            if constexpr (LPF >= 8) {
                if (sub < 5) {
                    // lane j of the group fetches winner j (static indexing keeps key/pos in registers)
                    unsigned long long kk = key[0]; uint32_t pp = pos[0];
#pragma unroll
                    for (int j = 1; j < 5; j++) if (sub == j) { kk = key[j]; pp = pos[j]; }
                    float4 c = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
                    float d = INFINITY;
                    if (kk != KNN_EMPTY) { c = ldg4(&g.pts[pp]); d = __uint_as_float((uint32_t)(kk >> 32)); }
                    s_nb[slot][sub] = c;
                    s_d2[slot][sub] = d;
                } else if (sub == 5) {
                    s_ori[slot] = po;
                    s_sel[slot] = make_float4(sx, sy, sz, 0.f);
                }
            } else if (sub == 0) {
#pragma unroll
                for (int j = 0; j < 5; j++) {
                    if constexpr (PHASES == 3) {
                        float4 c = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
                        float d = INFINITY;
                        if (key[j] != KNN_EMPTY) { c = ldg4(&g.pts[pos[j]]); d = __uint_as_float((uint32_t)(key[j] >> 32)); }
                        s_nb[slot][j] = c;
                        s_d2[slot][j] = d;
                    }
                    __stcg(&nbp[(size_t)f * 5 + j], key[j] != KNN_EMPTY ? pos[j] : 0xffffffffu);
                }
                if constexpr (PHASES == 3) {
                    s_ori[slot] = po;
                    s_sel[slot] = make_float4(sx, sy, sz, 0.f);
                }
            }
        }
    }
    }
    if constexpr (PHASES == 1) return;               // search only: the fits run in the next launch
    __syncthreads();
    if (prof && threadIdx.x == 0) prof[1] = clock64();

    // ---------------- phase 2: one thread per feature
    if (threadIdx.x < ((fpb + 31) & ~31)) {           // whole warps (the reduction below shuffles with a full mask)
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int slot = threadIdx.x;
        const int ff = fb + slot;
        bool keep = false;
        float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
        float ox = 0.f, oy = 0.f, oz = 0.f;
        if (slot < fpb && ff < nfeat) {
            float nx[5], ny[5], nz[5];
#pragma unroll
            for (int j = 0; j < 5; j++) { const float4 c = s_nb[slot][j]; nx[j] = c.x; ny[j] = c.y; nz[j] = c.z; }
            const float4 po = s_ori[slot], ps = s_sel[slot];
            ox = po.x; oy = po.y; oz = po.z;
            const float d4 = s_d2[slot][4];
            if (d4 < 1.0) {
                if constexpr (PHASES == 2) {
                    if (is_surf) keep = fit_plane_impl(nx, ny, nz, ps.x, ps.y, ps.z, ox, oy, oz, coeff);
                    else keep = fit_line_impl(nx, ny, nz, ps.x, ps.y, ps.z, coeff);
                } else {
                    if (is_surf) keep = fit_plane(nx, ny, nz, ps.x, ps.y, ps.z, ox, oy, oz, coeff);
                    else keep = fit_line(nx, ny, nz, ps.x, ps.y, ps.z, coeff);
                }
            }
            int32_t* di = is_surf ? a.dbg_idx_s : a.dbg_idx_c;
            if (di) {
                float* dd = is_surf ? a.dbg_d2_s : a.dbg_d2_c;
                float4* dc = is_surf ? a.dbg_coeff_s : a.dbg_coeff_c;
                uint8_t* df = is_surf ? a.dbg_flag_s : a.dbg_flag_c;
#pragma unroll
                for (int j = 0; j < 5; j++) {
                    di[(size_t)ff * 5 + j] = __float_as_int(s_nb[slot][j].w);
                    dd[(size_t)ff * 5 + j] = s_d2[slot][j];
                }
                dc[ff] = coeff; df[ff] = keep ? 1 : 0;
            }
        }
        if (prof && threadIdx.x == 0) prof[4] = clock64();      // (overwritten by the last CTA's epilogue stamp)
        // Jacobian row (LMOptimization :1191-1222): camera-frame swap p' = (y, z, x), c' = (cy, cz, cx)
        float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};       // six columns of A and b
        if (keep) {
            const float srx = s_trig[0], crx = s_trig[1], sry = s_trig[2], cry = s_trig[3], srz = s_trig[4], crz = s_trig[5];
            const float px = oy, py = oz, pz = ox;
            const float cfx = coeff.y, cfy = coeff.z, cfz = coeff.x;
            float arx = (crx*sry*srz*px + crx*crz*sry*py - srx*sry*pz) * cfx
                      + (-srx*srz*px - crz*srx*py - crx*pz) * cfy
                      + (crx*cry*srz*px + crx*cry*crz*py - cry*srx*pz) * cfz;
            float ary = ((cry*srx*srz - crz*sry)*px
                      + (sry*srz + cry*crz*srx)*py + crx*cry*pz) * cfx
                      + ((-cry*crz - srx*sry*srz)*px
                      + (cry*srz - crz*srx*sry)*py - crx*sry*pz) * cfz;
            float arz = ((crz*srx*sry - cry*srz)*px + (-cry*crz-srx*sry*srz)*py)*cfx
                      + (crx*crz*px - crx*srz*py) * cfy
                      + ((sry*srz + cry*crz*srx)*px + (crz*sry-cry*srx*srz)*py)*cfz;
            row[0] = arz; row[1] = arx; row[2] = ary; row[3] = cfz; row[4] = cfx; row[5] = cfy; row[6] = -coeff.w;
        }
        if (prof && threadIdx.x == 0) prof[12] = clock64();
        // 21 + 6 products in fp64 (float x float is exact there) + the row count; the 28 xor-shuffle trees are independent,
        // unrolled so their latencies overlap
        __syncwarp();                    // the fits diverge per lane; reconverge before the full-mask shuffles
        double acc[S2M_NPART];
        {
            int q = 0;
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int c = r; c < 6; c++) acc[q++] = (double)row[r] * (double)row[c];
#pragma unroll
            for (int r = 0; r < 6; r++) acc[21 + r] = (double)row[r] * (double)row[6];
            acc[27] = keep ? 1.0 : 0.0;
        }
#pragma unroll
        for (int q = 0; q < S2M_NPART; q++) acc[q] = warp_sum(acc[q]);
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < S2M_NPART; q++) s_wsum[warp][q] = acc[q];
        }
    }
    __syncthreads();
    if (prof && threadIdx.x == 0) prof[2] = clock64();
    if (threadIdx.x < S2M_NPART) {
        double v = 0.0;
        const int p2w = (fpb + 31) >> 5;
#pragma unroll
        for (int w = 0; w < P2_WARPS; w++) if (w < p2w) v += s_wsum[w][threadIdx.x];
        __stcg(&a.partial[((size_t)scan * a.max_blocks + blockIdx.x) * S2M_NPART + threadIdx.x], v);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(&st.ticket, 1u);
        s_last = (t == (unsigned)(nblk - 1)) ? 1 : 0;
    }
    __syncthreads();
    if (prof && threadIdx.x == 0) { prof[3] = clock64(); long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); prof[7] = gt; }
    if (!s_last) {
        if (n_pit == 1) return;
    } else {

    // ---------------- epilogue: the last CTA of this scan adds the partials in a fixed (slice, CTA) order
    __threadfence();
    {
        // thread t owns term t % 28 of the CTAs congruent to t / 28 modulo 9; their partials are requested twelve at a time
        // (independent loads in flight: the sum of 191 x 28 doubles costs two L2 round trips instead of six)
        constexpr int NSL = S2M_THREADS / S2M_NPART, BATCH = 12;
        const int q = threadIdx.x % S2M_NPART, slice = threadIdx.x / S2M_NPART;
        if (slice < NSL) {
            const double* base = a.partial + (size_t)scan * a.max_blocks * S2M_NPART + q;
            double sum = 0.0;
            for (int b = slice; b < nblk; b += BATCH * NSL) {
                double v[BATCH];
#pragma unroll
                for (int k = 0; k < BATCH; k++) v[k] = (b + k * NSL < nblk) ? __ldcg(&base[(size_t)(b + k * NSL) * S2M_NPART]) : 0.0;
#pragma unroll
                for (int k = 0; k < BATCH; k++) sum += v[k];
            }
            s_red[slice][q] = sum;
        }
    }
    __syncthreads();
    if (threadIdx.x < S2M_NPART) {
        double v = 0.0;
#pragma unroll
        for (int sl = 0; sl < S2M_THREADS / S2M_NPART; sl++) v += s_red[sl][threadIdx.x];
        s_sum[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (prof) prof[4] = clock64();
        st.ticket = 0;
        if (first) {
#pragma unroll
            for (int i = 0; i < 6; i++) st.pose[i] = a.init.pose[i];
            for (int i = 0; i < 36; i++) st.matP[i] = a.init.matP[i];
            st.degenerate = a.init.degenerate;
            st.done = 0; st.converged = 0; st.iters = 0; st.n_sel = 0; st.ran = 0;
            st.grid_status = a.gcm->status | a.gsm->status;
        }
        const int iterCount = a.iter >= 0 ? a.iter : __ldcg(&st.iters);
        lm_epilogue(st, s_sum, iterCount, a, scan, prof);
        if (prof) { prof[5] = clock64(); long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); prof[7] = gt; }
        if (a.persistent_iters > 0 && pit < 6) { long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); st.tl[1 + pit] = gt; }
    }
    // single-scan solves: the state goes straight to mapped host memory when the loop ends or the chunk does, followed by the
    // sequence number the host spins on (no copy, no stream synchronisation on the critical path)
    if (a.result) {
        __syncthreads();
        if (__ldcg(&st.done) || (a.chunk_last && pit == n_pit - 1)) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(&st);
            uint32_t* dst = reinterpret_cast<uint32_t*>(a.result);
            for (int i = threadIdx.x; i < (int)(sizeof(S2MState) / 4); i += S2M_THREADS) dst[i] = __ldcg(&src[i]);
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0) { *a.result_seq = a.seq; __threadfence_system(); }
        }
    }
        if (n_pit > 1 && threadIdx.x == 0) {      // publish the iteration: everything written above is visible before the flag
            __threadfence();
            *reinterpret_cast<volatile int*>(a.iter_flag) = a.flag_base + pit + 1;
        }
    }
    if (n_pit == 1) return;
    if (threadIdx.x == 0 && !s_last) {
        const int target = a.flag_base + pit + 1;
        while (*reinterpret_cast<volatile int*>(a.iter_flag) - target < 0) __nanosleep(200);
        __threadfence();
    }
    __syncthreads();
    first = false;
    if (__ldcg(&st.done)) return;
  }
}

__global__ void __launch_bounds__(256) k_pack_xyzi(const unsigned char* __restrict__ raw, size_t stride, int ioff, uint32_t n, float4* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char* p = raw + (size_t)i * stride;
    const float* f = reinterpret_cast<const float*>(p);
    out[i] = make_float4(f[0], f[1], f[2], *reinterpret_cast<const float*>(p + ioff));
}

struct Xf12 { float v[12]; };
__global__ void __launch_bounds__(256) k_transform_cloud(const unsigned char* __restrict__ in, size_t istride, int ioff_in, uint32_t n,
                                                         const Xf12 xfm, unsigned char* __restrict__ out, size_t ostride, int ioff_out) {
    const float* xf = xfm.v;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char* p = in + (size_t)i * istride;
    const float* f = reinterpret_cast<const float*>(p);
    const float x = f[0], y = f[1], z = f[2];
    float* o = reinterpret_cast<float*>(out + (size_t)i * ostride);
    o[0] = xf[0] * x + xf[1] * y + xf[2] * z + xf[3];
    o[1] = xf[4] * x + xf[5] * y + xf[6] * z + xf[7];
    o[2] = xf[8] * x + xf[9] * y + xf[10] * z + xf[11];
    *reinterpret_cast<float*>(out + (size_t)i * ostride + ioff_out) = *reinterpret_cast<const float*>(p + ioff_in);
}

}  // namespace b2

using namespace b2;

struct b2_s2m_s {
    b2_s2m_params prm;
    cudaStream_t stream = nullptr, stream2 = nullptr, stream_up = nullptr;   // solve + corner index; surf index; scan uploads
    int device = b2::current_device();      // the device the handle was created on
    cudaEvent_t ev_map = nullptr, ev_up = nullptr, ev_surf = nullptr;
    int last_iters = 3;                            // iterations the previous single-scan solve needed (sizes the first chunk)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_idx = nullptr;                  // start of the last b2_s2m_rebuild_map_index (device span of "index build + solve")
    bool idx_timed = false;
    GridIndex gc, gs;
    DevBuf raw_c, raw_s, scan_c, scan_s, off_c, off_s, state, partial, hist, ne, nb_c, nb_s;
    bool nb_valid = false;            // nb_* were written by an iteration on the current map and scan (host-driven b2_s2m_iterate)
    DevBuf dbg_idx_c, dbg_d2_c, dbg_coeff_c, dbg_flag_c, dbg_idx_s, dbg_d2_s, dbg_coeff_s, dbg_flag_s;
    PinBuf pin;
    float4* scan_s_alias = nullptr;    // surf features inside scan_c's allocation (both clouds uploaded with one copy)
    PinBuf stage; size_t stage_used = 0; bool stage_all = true;   // pinned staging of the scan features (set_scan)
    S2MState* h_result = nullptr;      // mapped pinned memory the last CTA writes the state into (single-scan solves) ...
    volatile int* h_seq = nullptr;     // ... followed by the sequence number the host spins on
    int seq = 0;
    bool last_ms_valid = true;
    DevBuf cand; bool count_candidates = false; unsigned long long last_candidates = 0;   // b2_s2m_count_candidates
    DevBuf flag; int flag_next = 0;    // iteration counter of persistent single-scan solves (device int, zeroed once)
    int persistent_ok = -1;            // -1 unknown, 0 cooperative launch unavailable, else CTAs that can be co-resident
    bool grid_checked = false;         // the status words of the current map grids have been read (check_grid_status)
    void* state_zeroed = nullptr;      // the allocation of `state` whose tickets are known to be zero
    bool have_map = false, have_scan = false;
    bool use_pdl = getenv("B2_S2M_NO_PDL") == nullptr;
    bool map_pending = false, scan_pending = false;   // set_map / set_scan left work on the streams that nothing has waited for yet
    int batch = 0;
    int max_blocks = 0;
    int max_feat_c = 0, max_feat_s = 0;   // largest per-scan feature counts in the batch
    size_t n_c = 0, n_s = 0;          // total features over the batch
    std::vector<int> h_off_c, h_off_s, h_off;
    int degenerate = 0;               // persistent members (:136,:234)
    float matP[36] = {0};
    float last_ms = 0.f; int last_launches = 0;
    // developer timeline (B2_S2M_TIMELINE=1): host clock at the entry / exit of the three calls of a step
    bool tl_on = getenv("B2_S2M_TIMELINE") != nullptr;
    double tl_host[8] = {0};
    cudaEvent_t tl_scan = nullptr;
    int tl_count = 0;
};

static std::mutex g_persist_mu;      // one persistent (barrier-synchronised) solve at a time per process: two could starve each other of SM slots

static double host_us() {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec * 1e6 + (double)ts.tv_nsec * 1e-3;
}

static void host_prepare_pose(const float pose[6], float xf[12], float trig[6]) {
    // exactly what the reference evaluates on the CPU: float sin/cos from the C library
    const float roll = pose[0], pitch = pose[1], yaw = pose[2];
    const float sr = std::sin(roll), cr = std::cos(roll), sp = std::sin(pitch), cp = std::cos(pitch), sy = std::sin(yaw), cy = std::cos(yaw);
    affine_from_trig(pose[3], pose[4], pose[5], cy, sy, cp, sp, cr, sr, xf);
    trig[0] = sp; trig[1] = cp; trig[2] = sy; trig[3] = cy; trig[4] = sr; trig[5] = cr;
}

// set_map and set_scan return while their kernels are still queued. A second call before anything waited for the first
// (no solve in between) would reuse or reallocate buffers those kernels still read: drain first. The usual sequence
// set_map, set_scan, solve never takes this path (solve ends with a synchronisation).
static int drain_pending(b2_s2m_s* h) {
    B2_CUDA(cudaStreamSynchronize(h->stream2));
    B2_CUDA(cudaStreamSynchronize(h->stream_up));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    h->map_pending = h->scan_pending = false;
    return B2_OK;
}

// Paths that do not carry the grid status back with their result (host-driven single iterations, batches) ask for it once
// per map: two words, one synchronisation. A map that outgrew its device-sized table is rebuilt on the host-sized path.
static int check_grid_status(b2_s2m_s* h) {
    if (h->grid_checked) return B2_OK;
    int st[2] = {0, 0};
    B2_CUDA(cudaMemcpyAsync(&st[0], &h->gc.dev_ptr()->status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaMemcpyAsync(&st[1], &h->gs.dev_ptr()->status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    if (st[0] != 0) B2_CHECK(h->gc.rebuild_exact(h->stream));
    if (st[1] != 0) B2_CHECK(h->gs.rebuild_exact(h->stream));
    if (st[0] != 0 || st[1] != 0) B2_CUDA(cudaStreamSynchronize(h->stream));
    h->grid_checked = true;
    return B2_OK;
}

// host cloud -> device, on the upload stream (the copy does not queue behind the index builds set_map left on h->stream).
// Clouds that are already x y z intensity, 16 bytes apart, go straight into place; others land in `raw` and are repacked by
// pack_points once h->stream has been made to wait for the copies.
// Small clouds (the features of one scan, ~100 KB) are first copied into the handle's own pinned buffer: set_scan can then
// return without waiting for the DMA (the wait, not the copy, was 20 us of the step).
constexpr size_t S2M_STAGE_MAX = (size_t)512 << 10;
static int upload_points(b2_s2m_s* h, DevBuf& raw, DevBuf& packed, const void* pts, size_t stride, size_t n) {
    if (n == 0) return B2_OK;
    B2_CHECK(packed.reserve(n * sizeof(float4)));
    if (h->stage_used + n * stride <= S2M_STAGE_MAX) {
        char* st = h->stage.as<char>() + h->stage_used;
        memcpy(st, pts, n * stride);
        h->stage_used += (n * stride + 255) & ~(size_t)255;
        pts = st;
    } else {
        h->stage_all = false;
    }
    if (stride == sizeof(float4)) { B2_CUDA(cudaMemcpyAsync(packed.p, pts, n * sizeof(float4), cudaMemcpyHostToDevice, h->stream_up)); return B2_OK; }
    B2_CHECK(raw.reserve(n * stride));
    B2_CUDA(cudaMemcpyAsync(raw.p, pts, n * stride, cudaMemcpyHostToDevice, h->stream_up));
    return B2_OK;
}
static int pack_points(b2_s2m_s* h, DevBuf& raw, DevBuf& packed, size_t stride, size_t n) {
    if (n == 0 || stride == sizeof(float4)) return B2_OK;
    k_pack_xyzi<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(raw.as<unsigned char>(), stride, (int)B2_INTENSITY_OFFSET(stride), (uint32_t)n, packed.as<float4>()); count_launch();
    B2_CUDA(cudaGetLastError());
    return B2_OK;
}

// One LM iteration for every scan of the batch. Shape: a single scan wants many small CTAs (latency), a batch wants
// CTAs whose second phase keeps all eight warps busy (throughput).
static int launch_iteration(b2_s2m_s* h, const S2MArgs& a, int batch) {
    if (batch <= 2) {
        // the previous winners' bound costs two dependent loads before the search starts: a loss on the latency shape
        S2MArgs b = a; b.use_prev = 0;
        dim3 grid((unsigned)h->max_blocks, (unsigned)batch);
        // iteration k+1 depends on iteration k's last CTA: no overlap to gain, but the launch latency between the two
        // (3-4 us of a 26 us kernel) goes away when the next grid is already resident and waiting
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(S2M_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = h->use_pdl ? 1 : 0;
        const cudaError_t le = cudaLaunchKernelEx(&cfg, k_s2m_iteration<S2M_LAT_LPF, S2M_LAT_ROUNDS, S2M_LAT_ROUNDS_C, 3>, b);
        if (le != cudaSuccess) { set_error("k_s2m_iteration launch -> %s", cudaGetErrorString(le)); return B2_ERR_CUDA; }
    } else {
        const int nb = (h->max_feat_c + S2M_THR_FPB - 1) / S2M_THR_FPB + (h->max_feat_s + S2M_THR_FPB - 1) / S2M_THR_FPB;
        dim3 grid((unsigned)std::max(nb, 1), (unsigned)batch);
        k_s2m_iteration<S2M_THR_LPF, S2M_THR_ROUNDS, S2M_THR_ROUNDS, 1><<<grid, S2M_THREADS, 0, h->stream>>>(a);
        k_s2m_iteration<S2M_THR_LPF, S2M_THR_ROUNDS, S2M_THR_ROUNDS, 2><<<grid, S2M_THREADS, 0, h->stream>>>(a);
        B2_CUDA(cudaGetLastError());
        count_launch();
    }
    count_launch();
    return B2_OK;
}

static S2MArgs make_args(b2_s2m_s* h, int iter, int device_driven, bool debug, float* pose_hist, int hist_stride) {
    S2MArgs a{};
    a.max_iters = hist_stride > 0 ? hist_stride : h->prm.max_iterations;
    a.gcm = h->gc.dev_ptr(); a.gsm = h->gs.dev_ptr();
    a.scan_c = h->scan_c.as<float4>(); a.scan_s = h->scan_s_alias ? h->scan_s_alias : h->scan_s.as<float4>();
    a.off_c = h->off_c.as<int>(); a.off_s = h->off_c.as<int>() + (h->batch + 1);
    a.single = (h->batch == 1 && h->h_off_c[0] == 0 && h->h_off_s[0] == 0) ? 1 : 0;
    if (a.single) { a.single_nc = h->h_off_c[1] - h->h_off_c[0]; a.single_ns = h->h_off_s[1] - h->h_off_s[0]; }
    a.st = h->state.as<S2MState>();
    a.partial = h->partial.as<double>();
    a.max_blocks = h->max_blocks;
    a.iter = iter; a.device_driven = device_driven;
    a.min_corr = h->prm.min_correspondences; a.eig_thr = h->prm.degenerate_eigen_threshold;
    a.pose_hist = pose_hist; a.hist_stride = hist_stride;
    a.want_matP = 1; a.done_count = nullptr;
    a.nb_c = h->nb_c.as<uint32_t>(); a.nb_s = h->nb_s.as<uint32_t>();
    a.use_prev = device_driven ? -1 : 0;
    a.cand_count = (h->count_candidates && h->cand.p) ? h->cand.as<unsigned long long>() : nullptr;
    if (debug) {
        a.dbg_idx_c = h->dbg_idx_c.as<int32_t>(); a.dbg_d2_c = h->dbg_d2_c.as<float>(); a.dbg_coeff_c = h->dbg_coeff_c.as<float4>(); a.dbg_flag_c = h->dbg_flag_c.as<uint8_t>();
        a.dbg_idx_s = h->dbg_idx_s.as<int32_t>(); a.dbg_d2_s = h->dbg_d2_s.as<float>(); a.dbg_coeff_s = h->dbg_coeff_s.as<float4>(); a.dbg_flag_s = h->dbg_flag_s.as<uint8_t>();
    }
    return a;
}

// offsets, CTA counts and per-scan buffers for the features already sitting in scan_c / scan_s
static int set_scan_finish(b2_s2m_s* h, int batch, bool host_upload = false) {
    const int32_t* coff = h->h_off_c.data();
    const int32_t* soff = h->h_off_s.data();
    // both offset tables in one buffer, one copy (pageable source: staged before cudaMemcpyAsync returns)
    h->h_off.assign(h->h_off_c.begin(), h->h_off_c.end());
    h->h_off.insert(h->h_off.end(), h->h_off_s.begin(), h->h_off_s.end());
    B2_CHECK(h->off_c.reserve(2 * (size_t)(batch + 1) * sizeof(int)));
    if (batch > 1 || coff[0] != 0 || soff[0] != 0) B2_CUDA(cudaMemcpyAsync(h->off_c.p, h->h_off.data(), 2 * (size_t)(batch + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    int mb = 1;
    h->max_feat_c = h->max_feat_s = 0;
    for (int b = 0; b < batch; b++) {
        h->max_feat_c = std::max(h->max_feat_c, coff[b + 1] - coff[b]);
        h->max_feat_s = std::max(h->max_feat_s, soff[b + 1] - soff[b]);
        int nb = (coff[b + 1] - coff[b] + S2M_LAT_FPB_C - 1) / S2M_LAT_FPB_C + (soff[b + 1] - soff[b] + S2M_LAT_FPB - 1) / S2M_LAT_FPB;
        mb = std::max(mb, nb);          // sized for the small-CTA shape, the larger count of the two
    }
    h->max_blocks = mb;
    h->batch = batch;
    B2_CHECK(h->state.reserve((size_t)batch * sizeof(S2MState) + (size_t)(batch + 1) * sizeof(int)));
    if (h->state.p != h->state_zeroed) {
        // single-scan solves never upload the state: the per-scan ticket must be zero when the first CTA arrives, and every
        // launch leaves it zero again
        B2_CUDA(cudaMemsetAsync(h->state.p, 0, h->state.cap, h->stream));
        h->state_zeroed = h->state.p;
    }
    B2_CHECK(h->partial.reserve((size_t)batch * mb * S2M_NPART * sizeof(double)));
    B2_CHECK(h->ne.reserve((size_t)(batch + 1) * sizeof(int)));
    B2_CHECK(h->nb_c.reserve(std::max<size_t>(h->n_c, 1) * 5 * sizeof(uint32_t)));
    B2_CHECK(h->nb_s.reserve(std::max<size_t>(h->n_s, 1) * 5 * sizeof(uint32_t)));
    h->nb_valid = false;
    if (batch == 1) {
        const size_t nc = std::max<size_t>(h->n_c, 1), ns = std::max<size_t>(h->n_s, 1);
        B2_CHECK(h->dbg_idx_c.reserve(nc * 5 * 4)); B2_CHECK(h->dbg_d2_c.reserve(nc * 5 * 4));
        B2_CHECK(h->dbg_coeff_c.reserve(nc * 16)); B2_CHECK(h->dbg_flag_c.reserve(nc));
        B2_CHECK(h->dbg_idx_s.reserve(ns * 5 * 4)); B2_CHECK(h->dbg_d2_s.reserve(ns * 5 * 4));
        B2_CHECK(h->dbg_coeff_s.reserve(ns * 16)); B2_CHECK(h->dbg_flag_s.reserve(ns));
    }
    // the caller may free its buffers on return: wait for the uploads only (the pack kernels and whatever set_map left on
    // h->stream run on; the offset tables above come from pageable memory, which cudaMemcpyAsync stages before returning)
    if (host_upload) { if (!h->stage_all) B2_CUDA(cudaStreamSynchronize(h->stream_up)); h->scan_pending = true; }
    else { B2_CUDA(cudaStreamSynchronize(h->stream)); h->map_pending = h->scan_pending = false; }
    h->have_scan = true;
    return B2_OK;
}


static int set_scan_common(b2_s2m_s* h, int batch, const void* corner, size_t cstride, const int32_t* coff,
                           const void* surf, size_t sstride, const int32_t* soff) {
    if (batch < 1 || batch > h->prm.max_batch) { set_error("set_scan: batch %d outside [1, %d]", batch, h->prm.max_batch); return B2_ERR_ARG; }
    if (h->scan_pending) B2_CHECK(drain_pending(h));
    h->h_off_c.assign(coff, coff + batch + 1);
    h->h_off_s.assign(soff, soff + batch + 1);
    for (int b = 0; b < batch; b++)
        if (coff[b + 1] < coff[b] || soff[b + 1] < soff[b]) { set_error("set_scan: offsets must be non-decreasing"); return B2_ERR_ARG; }
    h->n_c = (size_t)coff[batch]; h->n_s = (size_t)soff[batch];
    if ((h->n_c && !corner) || (h->n_s && !surf)) { set_error("set_scan: null feature array"); return B2_ERR_ARG; }
    B2_CHECK(h->stage.reserve(S2M_STAGE_MAX));
    h->stage_used = 0; h->stage_all = true;
    h->scan_s_alias = nullptr;
    const size_t off_s = (h->n_c * sizeof(float4) + 255) & ~(size_t)255;
    if (cstride == sizeof(float4) && sstride == sizeof(float4) && off_s + h->n_s * sizeof(float4) <= S2M_STAGE_MAX && h->n_c + h->n_s > 0) {
        // both clouds packed: staged side by side and moved with ONE copy (each small DMA costs ~7 us before its first byte,
        // and the first iteration waits for the features)
        B2_CHECK(h->scan_c.reserve(off_s + h->n_s * sizeof(float4)));
        char* st = h->stage.as<char>();
        if (h->n_c) memcpy(st, corner, h->n_c * sizeof(float4));
        if (h->n_s) memcpy(st + off_s, surf, h->n_s * sizeof(float4));
        B2_CUDA(cudaMemcpyAsync(h->scan_c.p, st, off_s + h->n_s * sizeof(float4), cudaMemcpyHostToDevice, h->stream_up));
        h->scan_s_alias = reinterpret_cast<float4*>(h->scan_c.as<char>() + off_s);
    } else {
    B2_CHECK(upload_points(h, h->raw_c, h->scan_c, corner, cstride, h->n_c));
    B2_CHECK(upload_points(h, h->raw_s, h->scan_s, surf, sstride, h->n_s));
    }
    B2_CUDA(cudaEventRecord(h->ev_up, h->stream_up));
    if (h->tl_on) { if (!h->tl_scan) cudaEventCreate(&h->tl_scan); cudaEventRecord(h->tl_scan, h->stream_up); }
    B2_CUDA(cudaStreamWaitEvent(h->stream, h->ev_up, 0));
    if (!h->scan_s_alias) {
        B2_CHECK(pack_points(h, h->raw_c, h->scan_c, cstride, h->n_c));
        B2_CHECK(pack_points(h, h->raw_s, h->scan_s, sstride, h->n_s));
    }
    return set_scan_finish(h, batch, true);
}

namespace b2 {

void pose_to_affine_host(const float pose6[6], float xf[12]) {
    float trig[6];
    host_prepare_pose(pose6, xf, trig);
}

// laserCloudCornerLastDS / laserCloudSurfLastDS handed over in device memory (packed xyzi); copies are ordered on h->stream
int s2m_set_scan_device(b2_s2m_s* h, const void* d_corner, size_t n_corner, const void* d_surf, size_t n_surf) {
    if (!h || n_corner > 0x7fffffff || n_surf > 0x7fffffff) return B2_ERR_ARG;
    if (h->scan_pending) B2_CHECK(drain_pending(h));
    h->h_off_c.assign({0, (int32_t)n_corner});
    h->h_off_s.assign({0, (int32_t)n_surf});
    h->n_c = n_corner; h->n_s = n_surf;
    h->scan_s_alias = nullptr;
    if (n_corner) { B2_CHECK(h->scan_c.reserve(n_corner * sizeof(float4))); B2_CUDA(cudaMemcpyAsync(h->scan_c.p, d_corner, n_corner * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream)); }
    if (n_surf) { B2_CHECK(h->scan_s.reserve(n_surf * sizeof(float4))); B2_CUDA(cudaMemcpyAsync(h->scan_s.p, d_surf, n_surf * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream)); }
    return set_scan_finish(h, 1);
}

// kdtreeCornerFromMap->setInputCloud / kdtreeSurfFromMap->setInputCloud on clouds that already live in device memory
int s2m_set_map_device(b2_s2m_s* h, const void* d_corner, size_t n_corner, const void* d_surf, size_t n_surf) {
    if (!h) return B2_ERR_ARG;
    if (h->map_pending) B2_CHECK(drain_pending(h));
    B2_CHECK(h->gc.upload_async(nullptr, d_corner, 16, n_corner, h->prm.knn_max_dist, h->stream, true));
    B2_CHECK(h->gs.upload_async(nullptr, d_surf, 16, n_surf, h->prm.knn_max_dist, h->stream2, true));
    B2_CHECK(h->gc.build_async(h->stream));
    B2_CHECK(h->gs.build_async(h->stream2));
    B2_CUDA(cudaEventRecord(h->ev_surf, h->stream2));
    B2_CUDA(cudaStreamWaitEvent(h->stream, h->ev_surf, 0));
    h->map_pending = true; h->grid_checked = false;
    h->have_map = true; h->nb_valid = false;
    return B2_OK;
}

}  // namespace b2

extern "C" {

void b2_s2m_default_params(b2_s2m_params* p) {
    if (!p) return;
    p->edge_feature_min_valid_num = 10;
    p->surf_feature_min_valid_num = 100;
    p->max_iterations = 30;
    p->min_correspondences = 50;
    p->knn_max_dist = 1.0f;
    p->degenerate_eigen_threshold = 100.f;
    p->max_batch = 1;
}

int b2_s2m_create(b2_s2m_t* out, const b2_s2m_params* params) {
    if (!out) { set_error("b2_s2m_create: null out"); return B2_ERR_ARG; }
    b2_s2m_s* h = new b2_s2m_s();
    if (params) h->prm = *params; else b2_s2m_default_params(&h->prm);
    if (h->prm.max_batch < 1 || h->prm.max_iterations < 1 || !(h->prm.knn_max_dist > 0.f)) { delete h; set_error("b2_s2m_create: bad params"); return B2_ERR_ARG; }
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream_up, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_map, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_up, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_surf, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev_idx);
    if (e != cudaSuccess) { set_error("b2_s2m_create: %s", cudaGetErrorString(e)); delete h; return B2_ERR_CUDA; }
    *out = h;
    return B2_OK;
}

int b2_s2m_destroy(b2_s2m_t h) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h) return B2_ERR_ARG;
    if (h->map_pending || h->scan_pending) drain_pending(h);     // kernels of a set_map / set_scan nobody waited for
    h->gc.release(); h->gs.release();
    h->flag.release(); h->cand.release();
    DevBuf* bufs[] = {&h->raw_c, &h->raw_s, &h->scan_c, &h->scan_s, &h->off_c, &h->off_s, &h->state, &h->partial, &h->hist, &h->ne, &h->nb_c, &h->nb_s,
                      &h->dbg_idx_c, &h->dbg_d2_c, &h->dbg_coeff_c, &h->dbg_flag_c, &h->dbg_idx_s, &h->dbg_d2_s, &h->dbg_coeff_s, &h->dbg_flag_s};
    for (DevBuf* b : bufs) b->release();
    h->pin.release(); h->stage.release();
    if (h->h_result) { cudaStreamSynchronize(h->stream); cudaFreeHost(h->h_result); }
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_idx) cudaEventDestroy(h->ev_idx);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream_up) cudaStreamDestroy(h->stream_up);
    if (h->ev_map) cudaEventDestroy(h->ev_map);
    if (h->ev_up) cudaEventDestroy(h->ev_up);
    if (h->ev_surf) cudaEventDestroy(h->ev_surf);
    delete h;
    return B2_OK;
}

int b2_s2m_set_map(b2_s2m_t h, const void* corner, size_t cstride, size_t n_corner, const void* surf, size_t sstride, size_t n_surf) {
    B2_NVTX("b2_s2m_set_map");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || (n_corner && !corner) || (n_surf && !surf) || cstride < 12 || sstride < 12 || (cstride & 3) || (sstride & 3)) {
        set_error("b2_s2m_set_map: bad argument"); return B2_ERR_ARG;
    }
    if (h->map_pending) B2_CHECK(drain_pending(h));
    h->gc.tl_on = h->gs.tl_on = h->tl_on;
    h->tl_host[0] = host_us();
    // Both uploads are queued first (the small cloud on the solve stream, the large one on the second stream), each followed
    // by its device-sized index build: no bounding box comes back to the host, the builds run while the caller goes on to
    // set_scan and solve. The only wait here is for the two copies, after which the caller's buffers are free.
    B2_CHECK(h->gc.upload_async(corner, nullptr, cstride, n_corner, h->prm.knn_max_dist, h->stream));
    B2_CHECK(h->gs.upload_async(surf, nullptr, sstride, n_surf, h->prm.knn_max_dist, h->stream2));
    B2_CUDA(cudaEventRecord(h->ev_up, h->stream));
    B2_CUDA(cudaEventRecord(h->ev_map, h->stream2));
    static const bool builds_after_copies = getenv("B2_S2M_BUILDS_AFTER_COPIES") != nullptr;      // kernel experiments
    if (builds_after_copies) B2_CUDA(cudaStreamWaitEvent(h->stream, h->ev_map, 0));
    B2_CHECK(h->gc.build_async(h->stream));
    B2_CHECK(h->gs.build_async(h->stream2));
    B2_CUDA(cudaEventRecord(h->ev_surf, h->stream2));
    B2_CUDA(cudaStreamWaitEvent(h->stream, h->ev_surf, 0));      // everything that uses the indexes is enqueued on h->stream
    // (busy wait: cudaEventSynchronize wakes up 5-9 us after the copy has landed, on a step of ~170 us)
    for (cudaEvent_t ev : {h->ev_up, h->ev_map}) {
        cudaError_t q;
        while ((q = cudaEventQuery(ev)) == cudaErrorNotReady) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        if (q != cudaSuccess) { set_error("b2_s2m_set_map: %s", cudaGetErrorString(q)); return B2_ERR_CUDA; }
    }
    B2_CHECK(h->gc.wait_ingest()); B2_CHECK(h->gs.wait_ingest());      // pinned maps are read by the build kernels themselves: wait for their last read
    h->map_pending = true;
    h->have_map = true; h->nb_valid = false; h->grid_checked = false;
    h->tl_host[1] = host_us();
    return B2_OK;
}

// kdtreeCornerFromMap->setInputCloud / kdtreeSurfFromMap->setInputCloud once more on the clouds the last b2_s2m_set_map (or
// b2_s2m_set_map_from_localmap) left in device memory: the reference rebuilds both kd-trees for every scan
// (mapOptmization.cpp:1289-1290), and this is that step with the inputs already resident in HBM.
int b2_s2m_rebuild_map_index(b2_s2m_t h) {
    B2_NVTX("b2_s2m_rebuild_map_index");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h) { set_error("b2_s2m_rebuild_map_index: null handle"); return B2_ERR_ARG; }
    if (!h->have_map) { set_error("b2_s2m_rebuild_map_index: no map set"); return B2_ERR_STATE; }
    if (h->map_pending || h->scan_pending) B2_CHECK(drain_pending(h));
    const void* dc = h->gc.src_; const void* ds = h->gs.src_;
    const size_t nc = h->gc.n, ns = h->gs.n, sc = h->gc.stride_, ss = h->gs.stride_;
    B2_CUDA(cudaEventRecord(h->ev_idx, h->stream));
    B2_CUDA(cudaStreamWaitEvent(h->stream2, h->ev_idx, 0));
    B2_CHECK(h->gc.upload_async(nullptr, dc, sc, nc, h->prm.knn_max_dist, h->stream));
    B2_CHECK(h->gs.upload_async(nullptr, ds, ss, ns, h->prm.knn_max_dist, h->stream2));
    B2_CHECK(h->gc.build_async(h->stream));
    B2_CHECK(h->gs.build_async(h->stream2));
    B2_CUDA(cudaEventRecord(h->ev_surf, h->stream2));
    B2_CUDA(cudaStreamWaitEvent(h->stream, h->ev_surf, 0));
    h->grid_checked = false;
    h->map_pending = true; h->nb_valid = false; h->idx_timed = true;
    return B2_OK;
}

int b2_s2m_last_step_gpu_ms(b2_s2m_t h, float* ms) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !ms) return B2_ERR_ARG;
    if (!h->idx_timed) { set_error("b2_s2m_last_step_gpu_ms: call b2_s2m_rebuild_map_index and a solve first"); return B2_ERR_STATE; }
    B2_CUDA(cudaEventSynchronize(h->ev1));
    B2_CUDA(cudaEventElapsedTime(ms, h->ev_idx, h->ev1));
    return B2_OK;
}

int b2_s2m_set_scan(b2_s2m_t h, const void* corner, size_t cstride, size_t n_corner, const void* surf, size_t sstride, size_t n_surf) {
    B2_NVTX("b2_s2m_set_scan");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || cstride < 16 || sstride < 16 || (cstride & 3) || (sstride & 3) || n_corner > 0x7fffffff || n_surf > 0x7fffffff) {
        set_error("b2_s2m_set_scan: bad argument"); return B2_ERR_ARG;
    }
    int32_t co[2] = {0, (int32_t)n_corner}, so[2] = {0, (int32_t)n_surf};
    h->tl_host[2] = host_us();
    const int st = set_scan_common(h, 1, corner, cstride, co, surf, sstride, so);
    h->tl_host[3] = host_us();
    return st;
}

int b2_s2m_set_scan_batch(b2_s2m_t h, int batch, const void* corner, size_t cstride, const int32_t* coff,
                          const void* surf, size_t sstride, const int32_t* soff) {
    B2_NVTX("b2_s2m_set_scan_batch");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !coff || !soff || cstride < 16 || sstride < 16 || (cstride & 3) || (sstride & 3)) { set_error("b2_s2m_set_scan_batch: bad argument"); return B2_ERR_ARG; }
    return set_scan_common(h, batch, corner, cstride, coff, surf, sstride, soff);
}

// mapOptimization::laserCloudInfoHandler's two fromROSMsg calls (mapOptmization.cpp:245-246) + downsampleCurrentScan (:940-958)
int b2_s2m_set_scan_downsampled(b2_s2m_t h, b2_voxel_t ds_corner, const void* corner, size_t cstride, size_t n_corner,
                                b2_voxel_t ds_surf, const void* surf, size_t sstride, size_t n_surf, size_t* n_corner_ds, size_t* n_surf_ds) {
    B2_NVTX("b2_s2m_set_scan_downsampled");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !ds_corner || !ds_surf || ds_corner == ds_surf || (n_corner && !corner) || (n_surf && !surf) || cstride < 16 || sstride < 16 ||
        (cstride & 3) || (sstride & 3) || n_corner > 0x7ffffff0ull || n_surf > 0x7ffffff0ull) {
        set_error("b2_s2m_set_scan_downsampled: bad argument"); return B2_ERR_ARG;
    }
    if (h->prm.max_batch < 1) return B2_ERR_STATE;
    uint32_t mc = 0, ms = 0;
    B2_CHECK(voxel_filter_host_to_dev(ds_corner, corner, cstride, n_corner, &mc));
    B2_CHECK(voxel_filter_host_to_dev(ds_surf, surf, sstride, n_surf, &ms));
    B2_CUDA(cudaStreamSynchronize(voxel_stream(ds_corner)));
    B2_CUDA(cudaStreamSynchronize(voxel_stream(ds_surf)));
    if (n_corner_ds) *n_corner_ds = mc;
    if (n_surf_ds) *n_surf_ds = ms;
    return s2m_set_scan_device(h, voxel_out_dev(ds_corner), mc, voxel_out_dev(ds_surf), ms);
}

// the same hand-off when featureExtraction and mapOptimization share the process: cornerCloud / surfaceCloud stay in HBM
int b2_s2m_set_scan_from_front_end(b2_s2m_t h, b2_scan_t scan, b2_voxel_t ds_corner, b2_voxel_t ds_surf, size_t* n_corner_ds, size_t* n_surf_ds) {
    B2_NVTX("b2_s2m_set_scan_from_front_end");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !scan || !ds_corner || !ds_surf || ds_corner == ds_surf) { set_error("b2_s2m_set_scan_from_front_end: bad argument"); return B2_ERR_ARG; }
    const void *dc = nullptr, *dsf = nullptr; size_t nc = 0, ns = 0;
    B2_CHECK(scan_features_dev(scan, &dc, &nc, &dsf, &ns));
    uint32_t mc = 0, ms = 0; int refused = 0;
    B2_CHECK(voxel_filter_dev(ds_corner, reinterpret_cast<const unsigned char*>(dc), 16, nc, 4, 16, std::max<size_t>(nc, 1), &mc, &refused, nullptr));
    B2_CHECK(voxel_filter_dev(ds_surf, reinterpret_cast<const unsigned char*>(dsf), 16, ns, 4, 16, std::max<size_t>(ns, 1), &ms, &refused, nullptr));
    B2_CUDA(cudaStreamSynchronize(voxel_stream(ds_corner)));
    B2_CUDA(cudaStreamSynchronize(voxel_stream(ds_surf)));
    if (n_corner_ds) *n_corner_ds = mc;
    if (n_surf_ds) *n_surf_ds = ms;
    return s2m_set_scan_device(h, voxel_out_dev(ds_corner), mc, voxel_out_dev(ds_surf), ms);
}

int b2_s2m_get_scan(b2_s2m_t h, int which, float* xyzi, size_t capacity, size_t* n) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !n || (which != 0 && which != 1)) { set_error("b2_s2m_get_scan: bad argument"); return B2_ERR_ARG; }
    if (!h->have_scan) { set_error("b2_s2m_get_scan: no scan set"); return B2_ERR_STATE; }
    const size_t cnt = which ? h->n_s : h->n_c;
    *n = cnt;
    if (!xyzi) return B2_OK;
    if (capacity < cnt) { set_error("b2_s2m_get_scan: %zu points, capacity %zu", cnt, capacity); return B2_ERR_CAPACITY; }
    const void* d_surf = h->scan_s_alias ? static_cast<const void*>(h->scan_s_alias) : h->scan_s.p;
    if (cnt) B2_CUDA(cudaMemcpyAsync(xyzi, which ? d_surf : h->scan_c.p, cnt * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    return B2_OK;
}

int b2_s2m_set_state(b2_s2m_t h, int degenerate, const float matP[36]) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h) return B2_ERR_ARG;
    h->degenerate = degenerate ? 1 : 0;
    if (matP) memcpy(h->matP, matP, sizeof(h->matP));
    return B2_OK;
}

int b2_s2m_iterate(b2_s2m_t h, float pose[6], int iter, int* n_sel, int* ran, int* converged, int* degenerate, float matP[36]) {
    B2_NVTX("b2_s2m_iterate");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !pose || iter < 0) { set_error("b2_s2m_iterate: bad argument"); return B2_ERR_ARG; }
    if (!h->have_map || !h->have_scan || h->batch != 1) { set_error("b2_s2m_iterate: set_map and set_scan (single scan) first"); return B2_ERR_STATE; }
    B2_CHECK(check_grid_status(h));
    B2_CHECK(h->pin.reserve(sizeof(S2MState)));
    S2MState* hs = h->pin.as<S2MState>();
    memset(hs, 0, sizeof(S2MState));
    memcpy(hs->pose, pose, 24);
    host_prepare_pose(pose, hs->xf, hs->trig);
    memcpy(hs->matP, h->matP, sizeof(h->matP));
    hs->degenerate = h->degenerate;
    B2_CUDA(cudaMemcpyAsync(h->state.p, hs, sizeof(S2MState), cudaMemcpyHostToDevice, h->stream));
    S2MArgs a = make_args(h, iter, 0, true, nullptr, 0);
    a.want_matP = matP != nullptr;
    a.use_prev = h->nb_valid ? 1 : 0;
    h->nb_valid = true;
    const bool prof = getenv("B2_S2M_PROF") != nullptr;
    if (prof) {
        B2_CHECK(h->hist.reserve((size_t)h->max_blocks * 16 * sizeof(long long)));
        B2_CUDA(cudaMemsetAsync(h->hist.p, 0, (size_t)h->max_blocks * 16 * sizeof(long long), h->stream));
        a.prof = h->hist.as<long long>();
    }
    B2_CHECK(launch_iteration(h, a, 1));
    if (prof) {
        std::vector<long long> t((size_t)h->max_blocks * 16);
        B2_CUDA(cudaMemcpyAsync(t.data(), h->hist.p, t.size() * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
        B2_CUDA(cudaStreamSynchronize(h->stream));
        long long g0 = -1, g1 = 0;
        const int nbc = (h->h_off_c[1] - h->h_off_c[0] + S2M_LAT_FPB_C - 1) / S2M_LAT_FPB_C;
        struct Acc { double p1 = 0, p2 = 0, fit = 0, row = 0, sums = 0, tail = 0, tot = 0, mx1 = 0, mx2 = 0, mxt = 0; int n = 0; } acc[2];
        for (int b = 0; b < h->max_blocks; b++) {
            const long long* r = &t[(size_t)b * 16];
            if (!r[0]) continue;
            if (g0 < 0 || r[6] < g0) g0 = r[6];
            g1 = std::max(g1, r[7]);
            if (r[5]) {
                fprintf(stderr, "[b2 prof] last CTA %d: reduce %lld cyc, epilogue %lld cyc (convert + QR solve %lld, degeneracy %lld, update %lld, next pose %lld)\n", b, r[4] - r[3], r[5] - r[4],
                        r[8] - r[4], r[9] - r[8], r[10] - r[9], r[5] - r[10]);
                continue;
            }
            Acc& A = acc[b >= nbc ? 1 : 0];
            A.n++;
            A.p1 += (double)(r[1] - r[0]); A.p2 += (double)(r[2] - r[1]); A.fit += (double)(r[4] - r[1]); A.row += (double)(r[12] - r[4]);
            A.sums += (double)(r[2] - r[12]); A.tail += (double)(r[3] - r[2]); A.tot += (double)(r[3] - r[0]);
            A.mx1 = std::max(A.mx1, (double)(r[1] - r[0])); A.mx2 = std::max(A.mx2, (double)(r[2] - r[1])); A.mxt = std::max(A.mxt, (double)(r[3] - r[0]));
        }
        for (int k = 0; k < 2; k++) {
            const Acc& A = acc[k]; const double n = std::max(A.n, 1);
            fprintf(stderr, "[b2 prof] iter %d %s: %d CTAs, mean cycles phase1 %.0f (max %.0f) phase2 %.0f (max %.0f: fits %.0f row %.0f sums %.0f) store+ticket %.0f | CTA total %.0f (max %.0f)\n",
                    iter, k ? "surf" : "corner", A.n, A.p1 / n, A.mx1, A.p2 / n, A.mx2, A.fit / n, A.row / n, A.sums / n, A.tail / n, A.tot / n, A.mxt);
        }
        {
            std::vector<double> st_us, en_us;
            for (int b = 0; b < h->max_blocks; b++) {
                const long long* r = &t[(size_t)b * 16];
                if (!r[0]) continue;
                st_us.push_back((double)(r[6] - g0) * 1e-3); en_us.push_back((double)(r[7] - g0) * 1e-3);
            }
            std::sort(st_us.begin(), st_us.end()); std::sort(en_us.begin(), en_us.end());
            int occ = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_s2m_iteration<S2M_LAT_LPF, S2M_LAT_ROUNDS, S2M_LAT_ROUNDS_C, 3>, S2M_THREADS, 0);
            const size_t m = st_us.size();
            fprintf(stderr, "[b2 prof] iter %d: CTA starts (us) p50 %.2f p90 %.2f max %.2f | ends p50 %.2f p90 %.2f max %.2f | CTAs/SM by occupancy %d\n",
                    iter, st_us[m / 2], st_us[m * 9 / 10], st_us[m - 1], en_us[m / 2], en_us[m * 9 / 10], en_us[m - 1], occ);
        }
        fprintf(stderr, "[b2 prof] iter %d: first start -> last end %.2f us\n", iter, (double)(g1 - g0) * 1e-3);
    }
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaMemcpyAsync(hs, h->state.p, sizeof(S2MState), cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    h->map_pending = h->scan_pending = false;
    memcpy(pose, hs->pose, 24);
    h->degenerate = hs->degenerate;
    memcpy(h->matP, hs->matP, sizeof(h->matP));
    if (n_sel) *n_sel = hs->n_sel;
    if (ran) *ran = hs->ran;
    if (converged) *converged = hs->ran ? hs->converged : 0;
    if (degenerate) *degenerate = hs->degenerate;
    if (matP) memcpy(matP, hs->matP, sizeof(h->matP));
    return B2_OK;
}

// Single-scan solve (the latency path of the C1 workload). No upload and no read-back copy: the initial state rides in the
// first launch's arguments, and the last CTA of the iteration that ends the loop (or the chunk) writes the state into mapped
// host memory, then a sequence number the host spins on. If the device-sized map grids did not fit their tables the status
// comes back with the state; the grids are then rebuilt on the host-sized path and the solve is repeated.
static int run_solve_single(b2_s2m_s* h, float* pose, int max_iterations, int* iters_done, int* converged, int* degenerate,
                            float* matP_out, int* not_enough, float* pose_history, bool want_matP) {
    if (max_iterations < 1) max_iterations = h->prm.max_iterations;
    const int nc = h->h_off_c[1] - h->h_off_c[0], ns = h->h_off_s[1] - h->h_off_s[0];
    if (!(nc > h->prm.edge_feature_min_valid_num && ns > h->prm.surf_feature_min_valid_num)) {
        // guard of scan2MapOptimization :1287: nothing runs, the pose stays
        if (iters_done) *iters_done = 0;
        if (converged) *converged = 0;
        if (degenerate) *degenerate = h->degenerate;
        if (not_enough) *not_enough = 1;
        if (matP_out) memcpy(matP_out, h->matP, sizeof(h->matP));
        h->last_ms = 0.f; h->last_launches = 0; h->last_ms_valid = true;
        return B2_OK;
    }
    if (!h->h_result) {
        void* p = nullptr;
        B2_CUDA(cudaHostAlloc(&p, sizeof(S2MState) + 64, cudaHostAllocMapped));
        memset(p, 0, sizeof(S2MState) + 64);
        h->h_result = static_cast<S2MState*>(p);
        h->h_seq = reinterpret_cast<volatile int*>(static_cast<char*>(p) + sizeof(S2MState));
    }
    S2MState* d_result = nullptr;
    B2_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_result), h->h_result, 0));
    volatile int* d_seq = reinterpret_cast<volatile int*>(reinterpret_cast<char*>(d_result) + sizeof(S2MState));
    float* d_hist = nullptr;
    if (pose_history) {
        B2_CHECK(h->hist.reserve((size_t)max_iterations * 6 * sizeof(float)));
        d_hist = h->hist.as<float>();
        B2_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)max_iterations * 6 * sizeof(float), h->stream));
    }
    for (int attempt = 0; attempt < 2; attempt++) {
        S2MArgs a = make_args(h, -1, 1, false, d_hist, max_iterations);
        a.want_matP = want_matP ? 1 : 0;
        a.result = d_result; a.result_seq = d_seq;
        memcpy(a.init.pose, pose, 24);
        host_prepare_pose(pose, a.init.xf, a.init.trig);     // float sin/cos from the C library, exactly as the reference evaluates them
        memcpy(a.init.matP, h->matP, sizeof(h->matP));
        a.init.degenerate = h->degenerate;
        B2_CUDA(cudaEventRecord(h->ev0, h->stream));
        int launched = 0, n_launch = 0;
        const S2MState* r = h->h_result;
        if (h->persistent_ok < 0) {
            int per_sm = 0;
            static const bool off = getenv("B2_S2M_NO_PERSISTENT") != nullptr;

            if (!off && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_s2m_iteration<S2M_LAT_LPF, S2M_LAT_ROUNDS, S2M_LAT_ROUNDS_C, 3>, S2M_THREADS, 0) == cudaSuccess)
                h->persistent_ok = per_sm * device_sm_count();
            else { cudaGetLastError(); h->persistent_ok = 0; }
        }
        auto wait_seq = [&](int seq) -> int {
            // spin on the sequence number; a failed launch or a device fault shows up as a finished stream without it
            for (unsigned spins = 0; *h->h_seq != seq; spins++) {
                if ((spins & 0xfffu) == 0xfffu) {
                    const cudaError_t q = cudaStreamQuery(h->stream);
                    if (q != cudaErrorNotReady && *h->h_seq != seq) {
                        if (q == cudaSuccess) { set_error("scan-to-map solve: the stream drained without a result"); return B2_ERR_CUDA; }
                        set_error("scan-to-map solve: %s", cudaGetErrorString(q)); return B2_ERR_CUDA;
                    }
                }
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            __atomic_thread_fence(__ATOMIC_ACQUIRE);
            return B2_OK;
        };
        if (h->persistent_ok >= h->max_blocks) {
            // the whole LM loop in ONE launch (every CTA resident, iterations separated by a flag in L2)
            std::lock_guard<std::mutex> persist_lock(g_persist_mu);
            if (!h->flag.p) { B2_CHECK(h->flag.reserve(64)); B2_CUDA(cudaMemsetAsync(h->flag.p, 0, 64, h->stream)); h->flag_next = 0; }
            a.use_init = 1; a.persistent_iters = max_iterations; a.iter_flag = h->flag.as<int>(); a.flag_base = h->flag_next;
            h->flag_next += max_iterations;
            const int seq = ++h->seq;
            a.seq = seq; a.chunk_last = 1;
            // a plain launch (a cooperative one starts ~10 us later on B200): every CTA is resident because max_blocks fits the
            // occupancy and g_persist_mu keeps a second persistent solve of this process off the device until this one is done
            k_s2m_iteration<S2M_LAT_LPF, S2M_LAT_ROUNDS, S2M_LAT_ROUNDS_C, 3><<<dim3((unsigned)h->max_blocks, 1u), S2M_THREADS, 0, h->stream>>>(a);
            count_launch();
            B2_CUDA(cudaGetLastError());
            n_launch = 1; launched = max_iterations;
            B2_CUDA(cudaEventRecord(h->ev1, h->stream));
            B2_CHECK(wait_seq(seq));
        } else {
        // Iterations are enqueued in chunks; a finished scan's CTAs return at once, so an over-long chunk costs only empty
        // launches. First chunk: the previous solve's iteration count plus one (consecutive scans of a trajectory behave alike).
        const int first = std::min(max_iterations, std::max(4, h->last_iters + 1));
        static const int chunk_plan[] = {0, 4, 6, 8, 8};
        for (int c = 0; launched < max_iterations; c++) {
            const int chunk = std::min(c == 0 ? first : chunk_plan[std::min(c, 4)], max_iterations - launched);
            const int seq = ++h->seq;
            for (int it = 0; it < chunk; it++) {
                a.use_init = (launched + it == 0) ? 1 : 0;
                a.seq = seq; a.chunk_last = (it == chunk - 1) ? 1 : 0;
                B2_CHECK(launch_iteration(h, a, 1));
            }
            launched += chunk; n_launch += chunk;
            B2_CUDA(cudaEventRecord(h->ev1, h->stream));
            B2_CHECK(wait_seq(seq));
            if (r->done) break;
        }
        }
        h->last_launches = n_launch;
        h->last_ms_valid = false;                      // ev0 / ev1 are read on demand (b2_s2m_last_gpu_ms)
        h->map_pending = h->scan_pending = false;      // every kernel that reads the maps or the scan has finished
        if (r->grid_status != 0 && attempt == 0) {
            // a map outgrew its device-sized cell table: host-sized rebuild (raises the budget), then once more
            B2_CHECK(h->gc.rebuild_exact(h->stream));
            B2_CHECK(h->gs.rebuild_exact(h->stream));
            B2_CUDA(cudaStreamSynchronize(h->stream));
            continue;
        }
        if (r->grid_status != 0) { set_error("scan-to-map solve: map index does not fit (status %d)", r->grid_status); return B2_ERR_TOO_LARGE; }
        break;
    }
    const S2MState* r = h->h_result;
    if (pose_history) {
        B2_CUDA(cudaMemcpyAsync(pose_history, d_hist, (size_t)max_iterations * 6 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        B2_CUDA(cudaStreamSynchronize(h->stream));
    }
    h->nb_valid = true;
    h->last_iters = r->iters;
    memcpy(pose, r->pose, 24);
    if (iters_done) *iters_done = r->iters;
    if (converged) *converged = r->converged;
    if (degenerate) *degenerate = r->degenerate;
    if (not_enough) *not_enough = 0;
    h->degenerate = r->degenerate;
    memcpy(h->matP, r->matP, sizeof(h->matP));
    if (matP_out) memcpy(matP_out, r->matP, sizeof(h->matP));
    return B2_OK;
}

static int run_solve(b2_s2m_s* h, float* poses, int max_iterations, int* iters_done, int* converged, int* degenerate,
                     float* matP_out, int* not_enough, float* pose_history, bool want_matP) {
    const int B = h->batch;
    h->tl_host[4] = host_us();
    if (B == 1) {
        const int st = run_solve_single(h, poses, max_iterations, iters_done, converged, degenerate, matP_out, not_enough, pose_history, want_matP);
        h->tl_host[5] = host_us();
        if (st == B2_OK && h->tl_on && h->gc.tl_ev[1] && h->gs.tl_ev[3] && h->gc.tl_ev[3] && h->tl_scan && ++h->tl_count % 50 == 0) {
            cudaEventSynchronize(h->ev1);
            auto rel = [&](cudaEvent_t e) { float ms = 0.f; cudaEventElapsedTime(&ms, h->gc.tl_ev[0], e); return ms * 1e3f; };
            const double t0 = h->tl_host[0];
            fprintf(stderr, "[b2 timeline us] host: set_map %.1f..%.1f set_scan %.1f..%.1f solve %.1f..%.1f | device (0 = corner upload queued): "
                            "corner up %.1f build %.1f | surf start %.1f up %.1f build %.1f | scan up %.1f | solve ev0 %.1f ev1 %.1f\n",
                    0.0, h->tl_host[1] - t0, h->tl_host[2] - t0, h->tl_host[3] - t0, h->tl_host[4] - t0, h->tl_host[5] - t0,
                    rel(h->gc.tl_ev[1]), rel(h->gc.tl_ev[3]), rel(h->gs.tl_ev[0]), rel(h->gs.tl_ev[1]), rel(h->gs.tl_ev[3]), rel(h->tl_scan),
                    rel(h->ev0), rel(h->ev1));
            {
                const S2MState* r = h->h_result;
                fprintf(stderr, "[b2 timeline us] solve kernel (device clock): first CTA -> end of iteration 1 / 2 / 3: %.1f / %.1f / %.1f; events ev0 -> ev1 %.1f\n",
                        (r->tl[1] - r->tl[0]) * 1e-3, (r->tl[2] - r->tl[0]) * 1e-3, (r->tl[3] - r->tl[0]) * 1e-3, rel(h->ev1) - rel(h->ev0));
            }
            GridDevMem gm[2];
            cudaMemcpy(&gm[0], h->gc.dev_ptr(), sizeof(GridDevMem), cudaMemcpyDeviceToHost);
            cudaMemcpy(&gm[1], h->gs.dev_ptr(), sizeof(GridDevMem), cudaMemcpyDeviceToHost);
            for (int k = 0; k < 2; k++)
                fprintf(stderr, "[b2 timeline us] %s build phases: bbox %.1f geometry+count %.1f chunk sums %.1f scan %.1f scatter %.1f zero-other %.1f (grid %d x %d x %d)\n", k ? "surf" : "corner",
                        (gm[k].tl[1] - gm[k].tl[0]) * 1e-3, (gm[k].tl[2] - gm[k].tl[1]) * 1e-3, (gm[k].tl[3] - gm[k].tl[2]) * 1e-3, (gm[k].tl[4] - gm[k].tl[3]) * 1e-3,
                        (gm[k].tl[5] - gm[k].tl[4]) * 1e-3, (gm[k].tl[6] - gm[k].tl[5]) * 1e-3, gm[k].g.nx, gm[k].g.ny, gm[k].g.nz);
        }
        return st;
    }
    if (max_iterations < 1) max_iterations = h->prm.max_iterations;
    B2_CHECK(check_grid_status(h));
    B2_CHECK(h->pin.reserve((size_t)B * sizeof(S2MState) + (size_t)(B + 1) * sizeof(int) + 64));
    S2MState* hs = h->pin.as<S2MState>();
    memset(hs, 0, (size_t)B * sizeof(S2MState));
    for (int b = 0; b < B; b++) {
        memcpy(hs[b].pose, poses + (size_t)b * 6, 24);
        memcpy(hs[b].matP, h->matP, sizeof(h->matP));
        hs[b].degenerate = h->degenerate;
    }
    float* d_hist = nullptr;
    if (pose_history) {
        B2_CHECK(h->hist.reserve((size_t)B * max_iterations * 6 * sizeof(float)));
        d_hist = h->hist.as<float>();
    }
    B2_CUDA(cudaEventRecord(h->ev0, h->stream));
    // the not-enough flags and the scans-finished counter live right behind the states: zeroed by the same upload, read back
    // by the same copy
    const size_t state_bytes = (size_t)B * sizeof(S2MState), tail_bytes = (size_t)(B + 1) * sizeof(int);
    B2_CHECK(h->state.reserve(state_bytes + tail_bytes));
    int* h_tail = reinterpret_cast<int*>(reinterpret_cast<char*>(hs) + state_bytes);
    memset(h_tail, 0, tail_bytes);
    B2_CUDA(cudaMemcpyAsync(h->state.p, hs, state_bytes + tail_bytes, cudaMemcpyHostToDevice, h->stream));
    if (d_hist) B2_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)B * max_iterations * 6 * sizeof(float), h->stream));
    int* d_ne = reinterpret_cast<int*>(h->state.as<char>() + state_bytes);
    int* d_done = d_ne + B;
    int n_launch = 0;
    {
        k_s2m_prepare<<<(B + 127) / 128, 128, 0, h->stream>>>(h->state.as<S2MState>(), B, h->off_c.as<int>(), h->off_c.as<int>() + (B + 1),
                                                              h->prm.edge_feature_min_valid_num, h->prm.surf_feature_min_valid_num, d_ne, d_done); count_launch();
        B2_CUDA(cudaGetLastError());
        n_launch = 1;
    }
    if (h->count_candidates) {
        B2_CHECK(h->cand.reserve(64));
        B2_CUDA(cudaMemsetAsync(h->cand.p, 0, 64, h->stream));
    }
    S2MArgs a = make_args(h, -1, 1, false, d_hist, max_iterations);
    a.want_matP = want_matP ? 1 : 0;
    a.done_count = d_done;
    // Iterations are enqueued in chunks; after each chunk the host reads the scans-finished counter together with the
    // whole (small) result state in one batch of copies and one synchronisation, so a solve that ends inside its first
    // chunk (the usual 3-4 iterations) costs a single host round trip. A finished scan's CTAs return at once, so an
    // over-long chunk costs only empty launches; the first chunk is the previous solve's iteration count plus one.
    int* h_ne = reinterpret_cast<int*>(hs + B);
    int* h_done = h_ne + B;
    int launched = 0;
    const int first = std::min(max_iterations, std::max(4, h->last_iters + 1));
    static const int chunk_plan[] = {0, 4, 6, 8, 8};
    for (int c = 0; launched < max_iterations; c++) {
        const int chunk = std::min(c == 0 ? first : chunk_plan[std::min(c, 4)], max_iterations - launched);
        for (int it = 0; it < chunk; it++) B2_CHECK(launch_iteration(h, a, B));
        launched += chunk; n_launch += chunk;
        B2_CUDA(cudaGetLastError());
        B2_CUDA(cudaEventRecord(h->ev1, h->stream));
        B2_CUDA(cudaMemcpyAsync(hs, h->state.p, state_bytes + tail_bytes, cudaMemcpyDeviceToHost, h->stream));   // states, not-enough flags, scans finished
        B2_CUDA(cudaStreamSynchronize(h->stream));
        if (*h_done >= B) break;
    }
    if (pose_history) {
        B2_CUDA(cudaMemcpyAsync(pose_history, d_hist, (size_t)B * max_iterations * 6 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        B2_CUDA(cudaStreamSynchronize(h->stream));
    }
    B2_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    h->last_ms_valid = true;
    h->last_launches = n_launch;
    if (h->count_candidates) {
        B2_CUDA(cudaMemcpyAsync(&h->last_candidates, h->cand.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        B2_CUDA(cudaStreamSynchronize(h->stream));
    }
    h->map_pending = h->scan_pending = false;      // h->stream has drained, and it had waited for the other two
    h->nb_valid = true;
    for (int b = 0; b < B; b++) {
        if (!h_ne[b]) memcpy(poses + (size_t)b * 6, hs[b].pose, 24);
        if (iters_done) iters_done[b] = hs[b].iters;
        if (converged) converged[b] = hs[b].converged;
        if (degenerate) degenerate[b] = hs[b].degenerate;
        if (not_enough) not_enough[b] = h_ne[b];
    }
    if (matP_out) memcpy(matP_out, hs[0].matP, sizeof(h->matP));
    return B2_OK;
}

int b2_s2m_solve(b2_s2m_t h, float pose[6], int max_iterations, int* iters_done, int* converged, int* degenerate,
                 float matP[36], int* not_enough_features, float* pose_history) {
    B2_NVTX("b2_s2m_solve");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !pose) { set_error("b2_s2m_solve: bad argument"); return B2_ERR_ARG; }
    if (!h->have_map || !h->have_scan || h->batch != 1) { set_error("b2_s2m_solve: set_map and set_scan (single scan) first"); return B2_ERR_STATE; }
    return run_solve(h, pose, max_iterations, iters_done, converged, degenerate, matP, not_enough_features, pose_history, matP != nullptr);
}

int b2_s2m_solve_batch(b2_s2m_t h, float* poses, int max_iterations, int* iters_done, int* converged, int* degenerate) {
    B2_NVTX("b2_s2m_solve_batch");
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !poses) { set_error("b2_s2m_solve_batch: bad argument"); return B2_ERR_ARG; }
    if (!h->have_map || !h->have_scan) { set_error("b2_s2m_solve_batch: set_map and set_scan_batch first"); return B2_ERR_STATE; }
    return run_solve(h, poses, max_iterations, iters_done, converged, degenerate, nullptr, nullptr, nullptr, false);
}

int b2_s2m_get_pass(b2_s2m_t h, int which, int32_t* knn_idx, float* knn_d2, float* coeff, uint8_t* flag) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !h->have_scan || h->batch != 1) { set_error("b2_s2m_get_pass: single-scan state required"); return B2_ERR_STATE; }
    const size_t n = which ? h->n_s : h->n_c;
    if (n == 0) return B2_OK;
    DevBuf& di = which ? h->dbg_idx_s : h->dbg_idx_c; DevBuf& dd = which ? h->dbg_d2_s : h->dbg_d2_c;
    DevBuf& dc = which ? h->dbg_coeff_s : h->dbg_coeff_c; DevBuf& df = which ? h->dbg_flag_s : h->dbg_flag_c;
    if (knn_idx) B2_CUDA(cudaMemcpyAsync(knn_idx, di.p, n * 5 * 4, cudaMemcpyDeviceToHost, h->stream));
    if (knn_d2) B2_CUDA(cudaMemcpyAsync(knn_d2, dd.p, n * 5 * 4, cudaMemcpyDeviceToHost, h->stream));
    if (coeff) B2_CUDA(cudaMemcpyAsync(coeff, dc.p, n * 16, cudaMemcpyDeviceToHost, h->stream));
    if (flag) B2_CUDA(cudaMemcpyAsync(flag, df.p, n, cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    return B2_OK;
}

int b2_s2m_get_normal_equations(b2_s2m_t h, float AtA[36], float AtB[6], float X[6]) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h || !h->have_scan) { set_error("b2_s2m_get_normal_equations: no scan"); return B2_ERR_STATE; }
    S2MState hs;
    B2_CUDA(cudaMemcpyAsync(&hs, h->state.p, sizeof(S2MState), cudaMemcpyDeviceToHost, h->stream));
    B2_CUDA(cudaStreamSynchronize(h->stream));
    if (AtA) memcpy(AtA, hs.AtA, sizeof(hs.AtA));
    if (AtB) memcpy(AtB, hs.AtB, sizeof(hs.AtB));
    if (X) memcpy(X, hs.X, sizeof(hs.X));
    return B2_OK;
}

// Statistics for the roofline numerator: with `enable` the batched solves count every candidate their neighbour search loads
// (one atomic per warp; costs a few percent, so bench.py switches it on for one untimed solve only).
int b2_s2m_count_candidates(b2_s2m_t h, int enable, unsigned long long* last_count) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h) return B2_ERR_ARG;
    h->count_candidates = enable != 0;
    if (last_count) *last_count = h->last_candidates;
    return B2_OK;
}

int b2_s2m_last_gpu_ms(b2_s2m_t h, float* ms, int* launches) {
    b2::DeviceScope device_scope_(h ? h->device : -1);
    if (!h) return B2_ERR_ARG;
    if (!h->last_ms_valid) {
        B2_CUDA(cudaEventSynchronize(h->ev1));
        B2_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
        h->last_ms_valid = true;
    }
    if (ms) *ms = h->last_ms;
    if (launches) *launches = h->last_launches;
    return B2_OK;
}

int b2_transform_cloud(const void* in, size_t in_stride, size_t n, const float pose6[6], void* out, size_t out_stride) {
    B2_NVTX("b2_transform_cloud");
    if ((n && (!in || !out)) || !pose6 || in_stride < 16 || out_stride < 16 || (in_stride & 3) || (out_stride & 3)) { set_error("b2_transform_cloud: bad argument"); return B2_ERR_ARG; }
    if (n == 0) return B2_OK;
    Xf12 xf; float trig[6];
    host_prepare_pose(pose6, xf.v, trig);
    // pooled buffers and a stream of its own (no cudaMalloc / cudaFree and no legacy-stream copies per call)
    size_t cap_in = 0, cap_out = 0;
    void* d_in = pool_alloc(n * in_stride, &cap_in);
    void* d_out = d_in ? pool_alloc(n * out_stride, &cap_out) : nullptr;
    cudaStream_t st = nullptr;
    cudaError_t e = (d_in && d_out) ? cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) : cudaErrorMemoryAllocation;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, in, n * in_stride, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_out, 0, n * out_stride, st);
    if (e == cudaSuccess) {
        k_transform_cloud<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(static_cast<const unsigned char*>(d_in), in_stride, (int)B2_INTENSITY_OFFSET(in_stride), (uint32_t)n, xf,
                                                                      static_cast<unsigned char*>(d_out), out_stride, (int)B2_INTENSITY_OFFSET(out_stride)); count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, n * out_stride, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (st) cudaStreamDestroy(st);
    if (d_in) pool_free(d_in, cap_in);
    if (d_out) pool_free(d_out, cap_out);
    if (e != cudaSuccess) { set_error("b2_transform_cloud: %s", cudaGetErrorString(e)); cudaGetLastError(); return B2_ERR_CUDA; }
    return B2_OK;
}

}  // extern "C"
