// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing in the product path may include, link or call this.
//
// CPU restatement of LIO-SAM's per-scan front end:
//   liosam_ws/src/LIO-SAM/src/imageProjection.cpp
//     imuDeskewInfo :305-362 (+ imuRPY2rosRPY, utility.h:322-334),
//     findRotation :446-471, findPosition :473-487 (returns 0), deskewPoint :489-519,
//     projectPointCloud :521-572, cloudExtraction :574-598
//   liosam_ws/src/LIO-SAM/src/featureExtraction.cpp
//     calculateSmoothness :81-101, markOccludedPoints :103-139, extractFeatures :141-238
// Expression types follow SURVEY.md Appendix A. Raw input points are PointXYZIRT, 32 B AoS:
// x@0 y@4 z@8 intensity@16 ring(uint16)@20 time(float)@24 (imageProjection.cpp:4-15).
// Oracle definitions where the reference is undefined: curvature / picked /
// label are zeroed per frame, cloudSmoothness[i] = {0, i} outside [5, size-5), and suppression
// stops at the array ends.
// parity unpinned by reference tests (none exist); libm atan2f/sinf/cosf are glibc's here.
#include "o_math.h"
#include <vector>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <algorithm>

extern "C" int o_voxel_grid(const float* in, int n, float lx, float ly, float lz, unsigned min_points_per_voxel,
                            float* out, int* voxel_of_point, int* refused, int* out_voxel_idx);

namespace {

struct ImuTable { const double* t; const double* rx; const double* ry; const double* rz; int cur; };

void find_rotation(const ImuTable& im, double pointTime, float* rx, float* ry, float* rz) {
    *rx = 0; *ry = 0; *rz = 0;
    int front = 0;
    while (front < im.cur) {
        if (pointTime < im.t[front]) break;
        ++front;
    }
    if (pointTime > im.t[front] || front == 0) {
        *rx = im.rx[front]; *ry = im.ry[front]; *rz = im.rz[front];
    } else {
        int back = front - 1;
        double ratioFront = (pointTime - im.t[back]) / (im.t[front] - im.t[back]);
        double ratioBack = (im.t[front] - pointTime) / (im.t[front] - im.t[back]);
        *rx = im.rx[front] * ratioFront + im.rx[back] * ratioBack;
        *ry = im.ry[front] * ratioFront + im.ry[back] * ratioBack;
        *rz = im.rz[front] * ratioFront + im.rz[back] * ratioBack;
    }
}

// Eigen::Affine3f::inverse() for a general affine: 3x3 cofactor inverse, translation = -inv * t
void affine_inverse(const float a[12], float o[12]) {
    float m00 = a[0], m01 = a[1], m02 = a[2], m10 = a[4], m11 = a[5], m12 = a[6], m20 = a[8], m21 = a[9], m22 = a[10];
    float c00 = m11 * m22 - m12 * m21, c10 = m12 * m20 - m10 * m22, c20 = m10 * m21 - m11 * m20;
    float det = c00 * m00 + c10 * m01 + c20 * m02;
    float invdet = 1.0f / det;
    o[0] = c00 * invdet;                     o[1] = (m02 * m21 - m01 * m22) * invdet; o[2]  = (m01 * m12 - m02 * m11) * invdet;
    o[4] = c10 * invdet;                     o[5] = (m00 * m22 - m02 * m20) * invdet; o[6]  = (m02 * m10 - m00 * m12) * invdet;
    o[8] = c20 * invdet;                     o[9] = (m01 * m20 - m00 * m21) * invdet; o[10] = (m00 * m11 - m01 * m10) * invdet;
    o[3]  = -(o[0] * a[3] + o[1] * a[7] + o[2] * a[11]);
    o[7]  = -(o[4] * a[3] + o[5] * a[7] + o[6] * a[11]);
    o[11] = -(o[8] * a[3] + o[9] * a[7] + o[10] * a[11]);
}

void affine_mul(const float a[12], const float b[12], float o[12]) {
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++)
            o[r * 4 + c] = a[r * 4 + 0] * b[0 * 4 + c] + a[r * 4 + 1] * b[1 * 4 + c] + a[r * 4 + 2] * b[2 * 4 + c];
        o[r * 4 + 3] = a[r * 4 + 0] * b[3] + a[r * 4 + 1] * b[7] + a[r * 4 + 2] * b[11] + a[r * 4 + 3];
    }
}

struct Smooth { float value; size_t ind; };

}  // namespace

extern "C" {

// projectPointCloud + cloudExtraction.
// raw: n points, 32 B stride. imu table: n_imu entries (imuPointerCur = n_imu - 1), pass n_imu <= 1 for
// "imuAvailable == false" (no deskew). deskew_flag: 1 when the cloud has a time field, else -1.
// Outputs: range_mat [N_SCAN*H] (FLT_MAX = empty), full_cloud [N_SCAN*H*4] (untouched cells = NaN),
// winner [N_SCAN*H] input index that owns each cell (-1 empty), and the compacted arrays of cloudExtraction.
// returns the extracted point count.
int o_project(const unsigned char* raw, int n, int N_SCAN, int H, int downsampleRate,
              float lidarMinRange, float lidarMaxRange,
              const double* imuTime, const double* imuRotX, const double* imuRotY, const double* imuRotZ, int n_imu,
              double timeScanCur, int deskew_flag,
              float* range_mat, float* full_cloud, int* winner,
              float* extracted, int* pointColInd, float* pointRange, int* startRingIndex, int* endRingIndex) {
    const size_t cells = (size_t)N_SCAN * H;
    for (size_t i = 0; i < cells; i++) { range_mat[i] = FLT_MAX; winner[i] = -1; }
    for (size_t i = 0; i < cells * 4; i++) full_cloud[i] = std::nanf("");
    ImuTable im{imuTime, imuRotX, imuRotY, imuRotZ, n_imu - 1};
    const bool imuAvailable = (n_imu - 1) > 0;
    bool firstPointFlag = true;
    float transStartInverse[12];
    const float ang_res_x = 360.0 / float(H);
    for (int i = 0; i < n; i++) {
        const unsigned char* rp = raw + (size_t)i * 32;
        float px, py, pz, pi, ptime; uint16_t ring;
        std::memcpy(&px, rp + 0, 4); std::memcpy(&py, rp + 4, 4); std::memcpy(&pz, rp + 8, 4);
        std::memcpy(&pi, rp + 16, 4); std::memcpy(&ring, rp + 20, 2); std::memcpy(&ptime, rp + 24, 4);
        float range = std::sqrt(px * px + py * py + pz * pz);
        if (range < lidarMinRange || range > lidarMaxRange) continue;
        int rowIdn = ring;
        if (rowIdn < 0 || rowIdn >= N_SCAN) continue;
        if (rowIdn % downsampleRate != 0) continue;
        int columnIdn = -1;
        float horizonAngle = std::atan2(px, py) * 180 / M_PI;
        columnIdn = -std::round((horizonAngle - 90.0) / ang_res_x) + H / 2;
        if (columnIdn >= H) columnIdn -= H;
        if (columnIdn < 0 || columnIdn >= H) continue;
        if (range_mat[(size_t)rowIdn * H + columnIdn] != FLT_MAX) continue;
        float nx = px, ny = py, nz = pz;
        if (!(deskew_flag == -1 || !imuAvailable)) {
            double pointTime = timeScanCur + (double)ptime;
            float rx, ry, rz;
            find_rotation(im, pointTime, &rx, &ry, &rz);
            float tf[12];
            orc::pcl_get_transformation(0.f, 0.f, 0.f, rx, ry, rz, tf);
            if (firstPointFlag) { affine_inverse(tf, transStartInverse); firstPointFlag = false; }
            float bt[12];
            affine_mul(transStartInverse, tf, bt);
            nx = bt[0] * px + bt[1] * py + bt[2] * pz + bt[3];
            ny = bt[4] * px + bt[5] * py + bt[6] * pz + bt[7];
            nz = bt[8] * px + bt[9] * py + bt[10] * pz + bt[11];
        }
        size_t index = (size_t)columnIdn + (size_t)rowIdn * H;
        range_mat[index] = range;
        winner[index] = i;
        full_cloud[index * 4 + 0] = nx; full_cloud[index * 4 + 1] = ny; full_cloud[index * 4 + 2] = nz; full_cloud[index * 4 + 3] = pi;
    }
    int count = 0;
    for (int i = 0; i < N_SCAN; ++i) {
        startRingIndex[i] = count - 1 + 5;
        for (int j = 0; j < H; ++j) {
            size_t c = (size_t)i * H + j;
            if (range_mat[c] != FLT_MAX) {
                pointColInd[count] = j;
                pointRange[count] = range_mat[c];
                std::memcpy(&extracted[(size_t)count * 4], &full_cloud[c * 4], 16);
                ++count;
            }
        }
        endRingIndex[i] = count - 1 - 5;
    }
    return count;
}

// calculateSmoothness + markOccludedPoints (+ the state they leave behind)
void o_curvature_masks(const float* pointRange, const int* pointColInd, int cloudSize,
                       float* curvature, int* picked, int* label) {
    for (int i = 0; i < cloudSize; i++) { curvature[i] = 0.f; picked[i] = 0; label[i] = 0; }
    for (int i = 5; i < cloudSize - 5; i++) {
        float diffRange = pointRange[i-5] + pointRange[i-4]
                        + pointRange[i-3] + pointRange[i-2]
                        + pointRange[i-1] - pointRange[i] * 10
                        + pointRange[i+1] + pointRange[i+2]
                        + pointRange[i+3] + pointRange[i+4]
                        + pointRange[i+5];
        curvature[i] = diffRange * diffRange;
    }
    for (int i = 5; i < cloudSize - 6; ++i) {
        float depth1 = pointRange[i];
        float depth2 = pointRange[i+1];
        int columnDiff = std::abs(int(pointColInd[i+1] - pointColInd[i]));
        if (columnDiff < 10) {
            if (depth1 - depth2 > 0.3) {
                picked[i - 5] = 1; picked[i - 4] = 1; picked[i - 3] = 1;
                picked[i - 2] = 1; picked[i - 1] = 1; picked[i] = 1;
            } else if (depth2 - depth1 > 0.3) {
                picked[i + 1] = 1; picked[i + 2] = 1; picked[i + 3] = 1;
                picked[i + 4] = 1; picked[i + 5] = 1; picked[i + 6] = 1;
            }
        }
        float diff1 = std::abs(float(pointRange[i-1] - pointRange[i]));
        float diff2 = std::abs(float(pointRange[i+1] - pointRange[i]));
        if (diff1 > 0.02 * pointRange[i] && diff2 > 0.02 * pointRange[i])
            picked[i] = 1;
    }
}

// extractFeatures. curvature/picked/label come from o_curvature_masks (picked and label are updated in place).
// stable_sort: 1 = order by (value, index) [the pinned mode], 0 = std::sort on value only (libstdc++ order).
// Outputs: corner_idx (capacity N_SCAN*6*20) in push order; surf_idx (capacity cloudSize) in push order with
// surf_ring_start[N_SCAN+1] marking each ring's slice; surf_ds (capacity cloudSize*4) the concatenated per-ring
// VoxelGrid outputs. returns n_corner; *n_surf_cand, *n_surf_ds filled.
int o_extract_features(const float* extracted, const int* pointColInd, int cloudSize,
                       const int* startRingIndex, const int* endRingIndex, int N_SCAN,
                       float edgeThreshold, float surfThreshold, float surfLeaf, int stable_sort,
                       const float* curvature, int* picked, int* label,
                       int* corner_idx, int* surf_idx, int* surf_ring_start, float* surf_ds,
                       int* n_surf_cand, int* n_surf_ds) {
    std::vector<Smooth> sm(cloudSize);
    for (int i = 0; i < cloudSize; i++) {
        sm[i].ind = (size_t)i;
        sm[i].value = (i >= 5 && i < cloudSize - 5) ? curvature[i] : 0.f;
    }
    auto in_range = [&](int k) { return k >= 0 && k < cloudSize; };
    int ncorner = 0, nsurf = 0, nds = 0;
    std::vector<float> scanbuf, dsbuf;
    for (int i = 0; i < N_SCAN; i++) {
        surf_ring_start[i] = nsurf;
        scanbuf.clear();
        for (int j = 0; j < 6; j++) {
            int sp = (startRingIndex[i] * (6 - j) + endRingIndex[i] * j) / 6;
            int ep = (startRingIndex[i] * (5 - j) + endRingIndex[i] * (j + 1)) / 6 - 1;
            if (sp >= ep) continue;
            if (stable_sort)
                std::sort(sm.begin() + sp, sm.begin() + ep, [](const Smooth& l, const Smooth& r) {
                    return l.value < r.value || (l.value == r.value && l.ind < r.ind); });
            else
                std::sort(sm.begin() + sp, sm.begin() + ep, [](const Smooth& l, const Smooth& r) { return l.value < r.value; });
            int largestPickedNum = 0;
            for (int k = ep; k >= sp; k--) {
                int ind = (int)sm[k].ind;
                if (picked[ind] == 0 && curvature[ind] > edgeThreshold) {
                    largestPickedNum++;
                    if (largestPickedNum <= 20) {
                        label[ind] = 1;
                        corner_idx[ncorner++] = ind;
                    } else {
                        break;
                    }
                    picked[ind] = 1;
                    for (int l = 1; l <= 5; l++) {
                        if (!in_range(ind + l)) break;
                        int columnDiff = std::abs(int(pointColInd[ind + l] - pointColInd[ind + l - 1]));
                        if (columnDiff > 10) break;
                        picked[ind + l] = 1;
                    }
                    for (int l = -1; l >= -5; l--) {
                        if (!in_range(ind + l)) break;
                        int columnDiff = std::abs(int(pointColInd[ind + l] - pointColInd[ind + l + 1]));
                        if (columnDiff > 10) break;
                        picked[ind + l] = 1;
                    }
                }
            }
            for (int k = sp; k <= ep; k++) {
                int ind = (int)sm[k].ind;
                if (picked[ind] == 0 && curvature[ind] < surfThreshold) {
                    label[ind] = -1;
                    picked[ind] = 1;
                    for (int l = 1; l <= 5; l++) {
                        if (!in_range(ind + l)) break;
                        int columnDiff = std::abs(int(pointColInd[ind + l] - pointColInd[ind + l - 1]));
                        if (columnDiff > 10) break;
                        picked[ind + l] = 1;
                    }
                    for (int l = -1; l >= -5; l--) {
                        if (!in_range(ind + l)) break;
                        int columnDiff = std::abs(int(pointColInd[ind + l] - pointColInd[ind + l + 1]));
                        if (columnDiff > 10) break;
                        picked[ind + l] = 1;
                    }
                }
            }
            for (int k = sp; k <= ep; k++) {
                if (label[k] <= 0) {
                    surf_idx[nsurf++] = k;
                    scanbuf.insert(scanbuf.end(), &extracted[(size_t)k * 4], &extracted[(size_t)k * 4] + 4);
                }
            }
        }
        int m = (int)scanbuf.size() / 4;
        if (m > 0) {
            dsbuf.resize((size_t)m * 4);
            int refused = 0;
            int got = o_voxel_grid(scanbuf.data(), m, surfLeaf, surfLeaf, surfLeaf, 0, dsbuf.data(), nullptr, &refused, nullptr);
            std::memcpy(&surf_ds[(size_t)nds * 4], dsbuf.data(), (size_t)got * 16);
            nds += got;
        }
    }
    surf_ring_start[N_SCAN] = nsurf;
    *n_surf_cand = nsurf; *n_surf_ds = nds;
    return ncorner;
}

// imuDeskewInfo (imageProjection.cpp:305-362), with imuRPY2rosRPY / imuAngular2rosAngular (utility.h:304-334).
// The queue is given as arrays in arrival order: stamp[n] (header.stamp.toSec()), orientation quaternion xyzw[n*4],
// angular velocity[n*3] — already passed through imuConverter by imuHandler (:203-204). tf::Matrix3x3(q).getRPY is
// restated from tf's LinearMath (Matrix3x3::setRotation + getEulerYPR, solution 1), all in double.
// Returns what the function leaves behind: *n_popped queue entries dropped from the front (:309-315), the table
// imuTime / imuRot{X,Y,Z}[0 .. *n_table) where *n_table - 1 == imuPointerCur after the final decrement (:356),
// *imu_available (:361), and the roll / pitch / yaw of the last message at or before timeScanCur narrowed to the
// float32 fields of cloud_info (:329-330; untouched when no message qualifies).
// The reference's arrays hold queueLength = 2000 entries (:45) and are not bounds-checked; `capacity` plays that role
// here and -1 is returned instead of writing past it.
int o_imu_deskew_info(const double* stamp, const double* quat_xyzw, const double* gyro, int n,
                      double timeScanCur, double timeScanEnd,
                      double* imuTime, double* imuRotX, double* imuRotY, double* imuRotZ, int capacity,
                      int* n_table, int* n_popped, int* imu_available, float* rpy_init) {
    *imu_available = 0;                                                    // cloudInfo.imuAvailable = false
    int front = 0;
    while (front < n) {                                                    // while (!imuQueue.empty())
        if (stamp[front] < timeScanCur - 0.01) ++front;                    //   pop_front()
        else break;
    }
    *n_popped = front;
    *n_table = 0;
    if (front == n) return 0;                                              // if (imuQueue.empty()) return;
    int imuPointerCur = 0;
    for (int i = front; i < n; ++i) {
        double currentImuTime = stamp[i];
        if (currentImuTime <= timeScanCur) {
            const double x = quat_xyzw[4 * i + 0], y = quat_xyzw[4 * i + 1], z = quat_xyzw[4 * i + 2], w = quat_xyzw[4 * i + 3];
            double d = x * x + y * y + z * z + w * w;
            double s = 2.0 / d;
            double xs = x * s, ys = y * s, zs = z * s;
            double wx = w * xs, wy = w * ys, wz = w * zs;
            double xx = x * xs, xy = x * ys, xz = x * zs;
            double yy = y * ys, yz = y * zs, zz = z * zs;
            double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
            double roll, pitch, yaw;
            if (std::fabs(m20) >= 1) {
                yaw = 0;
                double delta = std::atan2(m21, m22);
                if (m20 < 0) { pitch = M_PI / 2.0; roll = delta; }
                else         { pitch = -M_PI / 2.0; roll = delta; }
            } else {
                pitch = -std::asin(m20);
                roll = std::atan2(m21 / std::cos(pitch), m22 / std::cos(pitch));
                yaw = std::atan2(m10 / std::cos(pitch), m00 / std::cos(pitch));
            }
            rpy_init[0] = (float)roll; rpy_init[1] = (float)pitch; rpy_init[2] = (float)yaw;
        }
        if (currentImuTime > timeScanEnd + 0.01) break;
        if (imuPointerCur >= capacity) return -1;
        if (imuPointerCur == 0) {
            imuRotX[0] = 0; imuRotY[0] = 0; imuRotZ[0] = 0;
            imuTime[0] = currentImuTime;
            ++imuPointerCur;
            continue;
        }
        double angular_x = gyro[3 * i + 0], angular_y = gyro[3 * i + 1], angular_z = gyro[3 * i + 2];
        double timeDiff = currentImuTime - imuTime[imuPointerCur - 1];
        imuRotX[imuPointerCur] = imuRotX[imuPointerCur - 1] + angular_x * timeDiff;
        imuRotY[imuPointerCur] = imuRotY[imuPointerCur - 1] + angular_y * timeDiff;
        imuRotZ[imuPointerCur] = imuRotZ[imuPointerCur - 1] + angular_z * timeDiff;
        imuTime[imuPointerCur] = currentImuTime;
        ++imuPointerCur;
    }
    --imuPointerCur;
    *n_table = imuPointerCur + 1;
    if (imuPointerCur <= 0) return 0;
    *imu_available = 1;
    return 0;
}

}  // extern "C"
