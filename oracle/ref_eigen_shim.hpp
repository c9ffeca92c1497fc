// ref_eigen_shim.hpp -- TEST INFRASTRUCTURE.  The slice of the Eigen 3.3 API that the reference's `class View`
// (PRV_simulation/View_Space.hpp:40-199) uses, so that this class can be compiled FROM WHERE IT LIES without Eigen
// (oracle/Makefile target `ref`) and the oracle's / host mirror's restatement of its LOGIC (look-at frame, 71-step roll
// search with its tie rule, final pose) can be pinned against the reference's own code.
// The ARITHMETIC inside the shim is ours (Eigen itself is absent): fixed-size products / inverse / norms come from
// nerf-prv_b200/host/prv_linalg.hpp (the documented Eigen 3.3 evaluation order), AngleAxis -> Quaternion -> matrix is
// written here from Eigen's generic (non-SIMD) formulas.  So this pins the reading of the reference, not Eigen's rounding.
#pragma once
#include <cmath>

#include "../nerf-prv_b200/host/prv_linalg.hpp"

namespace Eigen {

struct Vector3d : prv::Vector3d {
    Vector3d() {}
    Vector3d(double x, double y, double z) : prv::Vector3d(x, y, z) {}
    Vector3d(const prv::Vector3d& o) : prv::Vector3d(o) {}
    static Vector3d UnitX() { return Vector3d(1, 0, 0); }
    static Vector3d UnitY() { return Vector3d(0, 1, 0); }
    static Vector3d UnitZ() { return Vector3d(0, 0, 1); }
};

struct Vector4d : prv::Vector4d {
    Vector4d() {}
    Vector4d(double x, double y, double z, double w) : prv::Vector4d(x, y, z, w) {}
    Vector4d(const prv::Vector4d& o) : prv::Vector4d(o) {}
    Vector4d eval() const { return *this; }
};

struct Matrix4d : prv::Matrix4d {
    Matrix4d() {}
    Matrix4d(int, int) {}  // Eigen leaves a dynamic-style (rows, cols) construction uninitialised; the reference sets all 16
    Matrix4d(const prv::Matrix4d& o) : prv::Matrix4d(o) {}
    static Matrix4d Identity(int = 4, int = 4) { return Matrix4d(prv::Matrix4d::Identity()); }
    Matrix4d inverse() const { return Matrix4d(prv::Matrix4d::inverse()); }
    Matrix4d eval() const { return *this; }
    Matrix4d operator*(const Matrix4d& o) const { return Matrix4d(prv::Matrix4d::operator*(o)); }
    Vector4d operator*(const Vector4d& x) const { return Vector4d(prv::Matrix4d::operator*(x)); }
};

struct Quaterniond {
    double w, x, y, z;
    // Eigen/src/Geometry/Quaternion.h, generic quat_product
    Quaterniond operator*(const Quaterniond& b) const {
        const Quaterniond& a = *this;
        return Quaterniond{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                           a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
    }
};

struct AngleAxisd {
    double angle;
    Vector3d axis;
    AngleAxisd(double a, const Vector3d& ax) : angle(a), axis(ax) {}
    // QuaternionBase::operator=(const AngleAxis&): ha = 0.5 * angle; w = cos(ha); vec = sin(ha) * axis
    Quaterniond quat() const {
        const double ha = 0.5 * angle, s = std::sin(ha);
        return Quaterniond{std::cos(ha), s * axis(0), s * axis(1), s * axis(2)};
    }
    Quaterniond operator*(const AngleAxisd& o) const { return quat() * o.quat(); }
};
inline Quaterniond operator*(const Quaterniond& q, const AngleAxisd& a) { return q * a.quat(); }

struct Matrix3d {
    double m[3][3];
    Matrix3d() : m{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}} {}
    static Matrix3d Identity(int = 3, int = 3) {
        Matrix3d r;
        r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.0;
        return r;
    }
    // QuaternionBase::toRotationMatrix()
    Matrix3d& operator=(const Quaterniond& q) {
        const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
        const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
        const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
        const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
        m[0][0] = 1.0 - (tyy + tzz); m[0][1] = txy - twz;         m[0][2] = txz + twy;
        m[1][0] = txy + twz;         m[1][1] = 1.0 - (txx + tzz); m[1][2] = tyz - twx;
        m[2][0] = txz - twy;         m[2][1] = tyz + twx;         m[2][2] = 1.0 - (txx + tyy);
        return *this;
    }
    double& operator()(int i, int j) { return m[i][j]; }
    double operator()(int i, int j) const { return m[i][j]; }
    Matrix3d eval() const { return *this; }
    Vector3d eulerAngles(int, int, int) const { return Vector3d(); }  // the reference computes it and never uses it (View_Space.hpp:132)
};

}  // namespace Eigen
