#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <geometry_msgs/TransformStamped.h>
#include <geometry_msgs/Pose.h>
namespace tf2 {
struct TransformException : public std::runtime_error { using std::runtime_error::runtime_error; };
struct Vector3 {
  double v[3] = {0, 0, 0};
  Vector3() = default;
  Vector3(double x, double y, double z) : v{x, y, z} {}
  double getX() const { return v[0]; } double getY() const { return v[1]; } double getZ() const { return v[2]; }
  double operator[](int i) const { return v[i]; }
};
class Quaternion {
 public:
  double x_ = 0, y_ = 0, z_ = 0, w_ = 1;
  Quaternion() = default;
  Quaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {}
  // Standard ZYX (yaw-pitch-roll) half-angle composition.
  void setRPY(double roll, double pitch, double yaw) {
    const double hr = roll * 0.5, hp = pitch * 0.5, hy = yaw * 0.5;
    const double cr = std::cos(hr), sr = std::sin(hr), cp = std::cos(hp), sp = std::sin(hp), cy = std::cos(hy), sy = std::sin(hy);
    x_ = sr * cp * cy - cr * sp * sy;
    y_ = cr * sp * cy + sr * cp * sy;
    z_ = cr * cp * sy - sr * sp * cy;
    w_ = cr * cp * cy + sr * sp * sy;
  }
  double x() const { return x_; } double y() const { return y_; } double z() const { return z_; } double w() const { return w_; }
};
class Matrix3x3 {
 public:
  Vector3 r[3];
  Matrix3x3() = default;
  explicit Matrix3x3(const Quaternion& q) {
    const double x = q.x_, y = q.y_, z = q.z_, w = q.w_;
    r[0] = Vector3(1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w));
    r[1] = Vector3(2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w));
    r[2] = Vector3(2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y));
  }
  const Vector3& operator[](int i) const { return r[i]; }
  // Roll-pitch-yaw of a rotation matrix (ZYX convention), the non-degenerate branch of tf2's getEulerYPR.
  void getRPY(double& roll, double& pitch, double& yaw) const {
    if (std::fabs(r[2][0]) >= 1.0) {
      yaw = 0.0;
      const double delta = std::atan2(r[2][1], r[2][2]);
      if (r[2][0] < 0) { pitch = M_PI / 2.0; roll = delta; } else { pitch = -M_PI / 2.0; roll = delta; }
    } else {
      pitch = -std::asin(r[2][0]);
      roll = std::atan2(r[2][1] / std::cos(pitch), r[2][2] / std::cos(pitch));
      yaw = std::atan2(r[1][0] / std::cos(pitch), r[0][0] / std::cos(pitch));
    }
  }
};
class Transform {
 public:
  Matrix3x3 basis; Vector3 origin;
  const Matrix3x3& getBasis() const { return basis; }
  const Vector3& getOrigin() const { return origin; }
  // Inverse rigid transform: R^T, -R^T t.
  Transform inverse() const {
    Transform o;
    for (int i = 0; i < 3; ++i) o.basis.r[i] = Vector3(basis.r[0][i], basis.r[1][i], basis.r[2][i]);
    o.origin = Vector3(-(o.basis.r[0][0] * origin[0] + o.basis.r[0][1] * origin[1] + o.basis.r[0][2] * origin[2]),
                       -(o.basis.r[1][0] * origin[0] + o.basis.r[1][1] * origin[1] + o.basis.r[1][2] * origin[2]),
                       -(o.basis.r[2][0] * origin[0] + o.basis.r[2][1] * origin[1] + o.basis.r[2][2] * origin[2]));
    return o;
  }
};
inline void convert(const Quaternion& q, geometry_msgs::Quaternion& out) { out.x = q.x_; out.y = q.y_; out.z = q.z_; out.w = q.w_; }
inline void convert(const geometry_msgs::Quaternion& q, Quaternion& out) { out = Quaternion(q.x, q.y, q.z, q.w); }
inline void convert(const geometry_msgs::Transform& t, Transform& out) {
  out.basis = Matrix3x3(Quaternion(t.rotation.x, t.rotation.y, t.rotation.z, t.rotation.w));
  out.origin = Vector3(t.translation.x, t.translation.y, t.translation.z);
}
// Not on any tested path (util.cpp:50-82 only needs these to link): rotation is dropped, translation kept.
inline void convert(const Transform& t, geometry_msgs::Transform& out) {
  out.translation.x = t.origin[0]; out.translation.y = t.origin[1]; out.translation.z = t.origin[2];
  out.rotation = geometry_msgs::Quaternion();
}
inline void doTransform(const geometry_msgs::Point& in, geometry_msgs::Point& out, const geometry_msgs::TransformStamped& t) {
  Transform tt;
  convert(t.transform, tt);
  out.x = tt.basis[0][0] * in.x + tt.basis[0][1] * in.y + tt.basis[0][2] * in.z + tt.origin[0];
  out.y = tt.basis[1][0] * in.x + tt.basis[1][1] * in.y + tt.basis[1][2] * in.z + tt.origin[1];
  out.z = tt.basis[2][0] * in.x + tt.basis[2][1] * in.y + tt.basis[2][2] * in.z + tt.origin[2];
}
}  // namespace tf2
