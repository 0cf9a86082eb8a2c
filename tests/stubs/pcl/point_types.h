// Minimal stand-ins for <pcl/point_types.h> / Eigen, ONLY so that tests/test_shim_compiles.py can syntax-check
// include/b2reg_pcl_shim.hpp against the C ABI in an image without PCL and Eigen. Not used by the product.
#pragma once
#include <cstdint>
#include <type_traits>
namespace Eigen {
enum { ColMajor = 0, RowMajor = 1 };
template <typename T, int R, int C, int O = ColMajor>
struct Matrix {
    T v[R * C];
    Matrix() : v{} {}
    template <int O2> Matrix(const Matrix<T, R, C, O2>& o) {
        for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) (*this)(r, c) = o(r, c);
    }
    T& operator()(int r, int c) { return O == RowMajor ? v[r * C + c] : v[c * R + r]; }
    const T& operator()(int r, int c) const { return O == RowMajor ? v[r * C + c] : v[c * R + r]; }
    T* data() { return v; }
    const T* data() const { return v; }
    static Matrix Identity() { Matrix m; for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = T(1); return m; }
};
typedef Matrix<float, 4, 4> Matrix4f;
}  // namespace Eigen
namespace pcl {
struct PointXYZ { float x, y, z, pad; };
struct PointXYZI { float x, y, z, pad0, intensity, pad1[3]; };
static_assert(sizeof(PointXYZI) == 32 && sizeof(PointXYZ) == 16, "PCL record sizes");
}  // namespace pcl
