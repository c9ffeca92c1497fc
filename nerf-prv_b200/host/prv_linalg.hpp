// prv_linalg.hpp -- the small slice of Eigen's fixed-size double API that the PRV_simulation hot path
// uses (Matrix4d, Vector3d, Vector4d), with the arithmetic order of Eigen 3.3 (README.md:7 of the
// reference pins Eigen 3.3.9) so that poses computed here are the poses the reference computes:
//   * M*M and M*v: coefficient-based product, inner index ascending, acc = a_k*b_k + acc (no FMA);
//   * inverse(): cofactor expansion (compute_inverse_size4 scalar path), det = col(0).row(0) of the
//     adjugate summed pairwise, then element-wise division;
//   * Vector3d::squaredNorm(): x*x + (y*y + z*z) (fixed-size unrolled reduction);
//   * normalized(): v / sqrt(squaredNorm()) when squaredNorm() > 0, else v.
// Must be compiled with FP contraction off (-ffp-contract=off / nvcc -Xcompiler -ffp-contract=off).
#pragma once
#include <cmath>
#include <cstddef>

namespace prv {

struct Vector3d {
    double v[3];
    Vector3d() : v{0, 0, 0} {}
    Vector3d(double x, double y, double z) : v{x, y, z} {}
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    Vector3d operator-(const Vector3d& o) const { return Vector3d(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vector3d operator+(const Vector3d& o) const { return Vector3d(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Vector3d operator*(double s) const { return Vector3d(v[0] * s, v[1] * s, v[2] * s); }
    Vector3d operator/(double s) const { return Vector3d(v[0] / s, v[1] / s, v[2] / s); }
    double squaredNorm() const { return v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]); }
    double norm() const { return std::sqrt(squaredNorm()); }
    Vector3d normalized() const {
        const double n = squaredNorm();
        return n > 0.0 ? (*this) / std::sqrt(n) : *this;
    }
    Vector3d cross(const Vector3d& o) const {
        return Vector3d(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
    }
};

struct Vector4d {
    double v[4];
    Vector4d() : v{0, 0, 0, 0} {}
    Vector4d(double x, double y, double z, double w) : v{x, y, z, w} {}
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
};

struct Matrix4d {
    double m[4][4];
    Matrix4d() {
        for (auto& r : m)
            for (double& e : r) e = 0.0;
    }
    static Matrix4d Identity() {
        Matrix4d r;
        for (int i = 0; i < 4; i++) r.m[i][i] = 1.0;
        return r;
    }
    static Matrix4d FromRowMajor(const double* p) {
        Matrix4d r;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) r.m[i][j] = p[4 * i + j];
        return r;
    }
    void toRowMajor(double* p) const {
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) p[4 * i + j] = m[i][j];
    }
    double& operator()(int i, int j) { return m[i][j]; }
    double operator()(int i, int j) const { return m[i][j]; }

    Matrix4d operator*(const Matrix4d& o) const {
        Matrix4d r;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
                double acc = m[i][0] * o.m[0][j];
                for (int k = 1; k < 4; k++) acc = m[i][k] * o.m[k][j] + acc;
                r.m[i][j] = acc;
            }
        return r;
    }
    Vector4d operator*(const Vector4d& x) const {
        Vector4d r;
        for (int i = 0; i < 4; i++) {
            double acc = m[i][0] * x.v[0];
            for (int k = 1; k < 4; k++) acc = m[i][k] * x.v[k] + acc;
            r.v[i] = acc;
        }
        return r;
    }

    Matrix4d inverse() const {
        // adj(j,i) = (-1)^(i+j) * minor(i,j), minor via the cyclic 3x3 expansion Eigen uses
        auto d3 = [this](int r1, int r2, int r3, int c1, int c2, int c3) {
            return m[r1][c1] * (m[r2][c2] * m[r3][c3] - m[r2][c3] * m[r3][c2]);
        };
        Matrix4d adj;
        for (int i = 0; i < 4; i++) {
            const int r1 = (i + 1) & 3, r2 = (i + 2) & 3, r3 = (i + 3) & 3;
            for (int j = 0; j < 4; j++) {
                const int c1 = (j + 1) & 3, c2 = (j + 2) & 3, c3 = (j + 3) & 3;
                const double cof = d3(r1, r2, r3, c1, c2, c3) + d3(r2, r3, r1, c1, c2, c3) + d3(r3, r1, r2, c1, c2, c3);
                adj.m[j][i] = ((i + j) % 2) ? -cof : cof;
            }
        }
        const double det = (m[0][0] * adj.m[0][0] + m[1][0] * adj.m[0][1]) + (m[2][0] * adj.m[0][2] + m[3][0] * adj.m[0][3]);
        for (auto& r : adj.m)
            for (double& e : r) e = e / det;
        return adj;
    }
};

// Rotation about +Z as Eigen produces it from AngleAxisd(0,X)*AngleAxisd(0,Y)*AngleAxisd(angle,Z)
// (a quaternion product -> Quaternion::toRotationMatrix(); View_Space.hpp:104-107 of the reference).
inline void rotation_about_z(double angle, double out[3][3]) {
    const double w = std::cos(0.5 * angle), z = std::sin(0.5 * angle);
    const double tz = 2.0 * z, twz = tz * w, tzz = tz * z;
    const double txx = 0.0, tyy = 0.0, txy = 0.0;
    out[0][0] = 1.0 - (tyy + tzz); out[0][1] = txy - twz;         out[0][2] = 0.0;
    out[1][0] = txy + twz;         out[1][1] = 1.0 - (txx + tzz); out[1][2] = 0.0;
    out[2][0] = 0.0;               out[2][1] = 0.0;               out[2][2] = 1.0 - (txx + tyy);
}

}  // namespace prv
