// host_motion.cpp — the scalar half of the motion update: mean / standard deviation of the six normal distributions
// (x y z roll pitch yaw) every ParticleCloud::motionUpdate variant of the reference sets up before it walks the particles
// (src/particle_cloud.cpp:153-462), and the odometry integration of the reference pose that gates the sensor update
// (:180-182, :352-354; src/mcl_3d.cpp:353). Pure host code (a dozen flops per update); the per-particle work is the
// k_motion_apply kernel. Operand types follow the reference variant by variant (fp64 for the odometry message fields,
// fp32 = FLOAT_T for everything else) so that the sigmas are bit-identical. Built with -ffp-contract=off.
#include "../../include/tsdfloc.h"

#include <cmath>

namespace
{

// sigma_k = a_i * d^2 + a_j * theta^2 in the arithmetic type T of the variant
template <typename T>
void pair_sigmas(const float a[12], T d2, T t2, double sigma[6])
{
  for (int k = 0; k < 6; ++k) sigma[k] = static_cast<double>(a[2 * k] * d2 + a[2 * k + 1] * t2);
}

}  // namespace

extern "C" int tsdfloc_motion_model(int variant, const double in[4], float time_diff, const float a[12], double mean[6], double sigma[6],
                                    float ref_pose[6])
{
  if (!in || !a || !mean || !sigma) return TSDFLOC_E_BAD_ARG;
  for (int k = 0; k < 6; ++k) mean[k] = 0.0;
  switch (variant)
  {
    case TSDFLOC_MOTION_ODOM:
    {
      // nav_msgs::Odometry fields are float64: `auto linear_velocity = odom.twist.twist.linear.x` (:167-168)
      const double v = in[0], w = in[1];
      if (ref_pose)
      {
        const float x = static_cast<float>(ref_pose[0] + v * time_diff * std::cos(ref_pose[5] + (w / 2 * time_diff)));
        const float y = static_cast<float>(ref_pose[1] + v * time_diff * std::sin(ref_pose[5] + (w / 2 * time_diff)));
        const float yaw = static_cast<float>(ref_pose[5] + w * time_diff);
        ref_pose[0] = x;
        ref_pose[1] = y;
        ref_pose[5] = yaw;
      }
      const double d = v * time_diff, theta = w * time_diff;
      pair_sigmas<double>(a, d * d, theta * theta, sigma);
      mean[0] = d;
      mean[5] = theta;
      return TSDFLOC_OK;
    }
    case TSDFLOC_MOTION_IMU:
    case TSDFLOC_MOTION_NOISE:
    {
      // ImuAccumulator::Data / lin_scale, ang_scale are FLOAT_T (:337-338, :388): fp32 arithmetic
      const float v = static_cast<float>(in[0]), w = static_cast<float>(in[1]);
      if (variant == TSDFLOC_MOTION_IMU && ref_pose)
      {
        const float heading = ref_pose[5] + (w / 2 * time_diff);
        const float step = v * time_diff;
        const float x = static_cast<float>(ref_pose[0] + step * std::cos(static_cast<double>(heading)));
        const float y = static_cast<float>(ref_pose[1] + step * std::sin(static_cast<double>(heading)));
        const float yaw = ref_pose[5] + w * time_diff;
        ref_pose[0] = x;
        ref_pose[1] = y;
        ref_pose[5] = yaw;
      }
      const float d = v * time_diff, theta = w * time_diff;
      pair_sigmas<float>(a, d * d, theta * theta, sigma);
      if (variant == TSDFLOC_MOTION_IMU)
      {
        mean[0] = d;
        mean[5] = theta;
      }
      return TSDFLOC_OK;
    }
    case TSDFLOC_MOTION_NOISE_IMU:
    {
      // :436-456: translation noise from lin_scale * dt, rotation noise around the IMU's accumulated angle deltas
      const float d = static_cast<float>(in[0]) * time_diff;
      const float roll = static_cast<float>(in[1]), pitch = static_cast<float>(in[2]), theta = static_cast<float>(in[3]);
      pair_sigmas<float>(a, d * d, theta * theta, sigma);
      sigma[3] = static_cast<double>(a[7] * (roll * roll));
      sigma[4] = static_cast<double>(a[9] * (pitch * pitch));
      sigma[5] = static_cast<double>(a[11] * (theta * theta));
      mean[3] = roll;
      mean[4] = pitch;
      mean[5] = theta;
      return TSDFLOC_OK;
    }
    default:
      return TSDFLOC_E_BAD_ARG;
  }
}
