// oracle/pc_harness.cpp — C ABI over the UNMODIFIED reference ParticleCloud motion update (test infrastructure only).
//
// Compiled together with /root/reference/src/particle_cloud.cpp and src/util/util.cpp (taken by path, never copied) against
// the stub ROS headers into oracle/_ref/libtsdf_ref_pc.so. This TU is built with -fno-access-control so that it can seed the
// cloud's private std::mt19937 and step the stub clock (ros::Time::now()) by hand; the reference sources themselves are
// compiled as they are.
//
// The reference draws its six motion samples per particle from std::normal_distribution<> objects fed by one mt19937
// (particle_cloud.cpp:212-217 / apply_model :498-506). libstdc++'s distributions are deterministic, so the same draws are
// reproduced here by constructing the same six distributions (mean, sigma supplied by the caller) on an equally seeded
// generator and drawing in the same order — that is how a test obtains the noise to inject into the GPU kernel.
#include <cstdint>
#include <cstring>
#include <random>
#include <string>

#include <tsdf_localization/particle_cloud.h>

using namespace tsdf_localization;

extern "C"
{

// variant: 0 motionUpdate(lin_scale, ang_scale) (:388-420); 1 motionUpdate(odom) (:153-331); 2 motionUpdate(imu) (:333-386);
// 3 motionUpdate(lin_scale, imu) (:422-462). in[]: see tsdfloc_motion_model. The clock goes from 100.0 to 100.0 + dt.
// ref_pose_io: in/out 6 floats. Returns 0.
int ref_pc_motion_update(int variant, const double in[4], double dt, const float a[12], uint32_t seed, float* particles, uint64_t n,
                         float ref_pose_io[6])
{
  ParticleCloud pc;
  pc.m_particles.resize(n);
  std::memcpy(static_cast<void*>(pc.m_particles.data()), particles, n * sizeof(Particle));
  pc.m_generator_ptr.reset(new std::mt19937(seed));
  pc.a_1_ = a[0]; pc.a_2_ = a[1]; pc.a_3_ = a[2]; pc.a_4_ = a[3]; pc.a_5_ = a[4]; pc.a_6_ = a[5];
  pc.a_7_ = a[6]; pc.a_8_ = a[7]; pc.a_9_ = a[8]; pc.a_10_ = a[9]; pc.a_11_ = a[10]; pc.a_12_ = a[11];
  for (int k = 0; k < 6; ++k) pc.ref_pose[k] = ref_pose_io[k];
  ros::tsdf_stub_clock() = 100.0;
  pc.m_last_time = ros::Time::now();
  ros::tsdf_stub_clock() = 100.0 + dt;
  switch (variant)
  {
    case 0:
      pc.motionUpdate(static_cast<FLOAT_T>(in[0]), static_cast<FLOAT_T>(in[1]));
      break;
    case 1:
    {
      nav_msgs::Odometry odom;
      odom.twist.twist.linear.x = in[0];
      odom.twist.twist.angular.z = in[1];
      pc.motionUpdate(odom);
      break;
    }
    case 2:
    {
      ImuAccumulator::Data d;
      d.linear_vel = static_cast<FLOAT_T>(in[0]);
      d.angular_yaw = static_cast<FLOAT_T>(in[1]);
      pc.motionUpdate(d);
      break;
    }
    case 3:
    {
      ImuAccumulator::Data d;
      d.delta_roll = static_cast<FLOAT_T>(in[1]);
      d.delta_pitch = static_cast<FLOAT_T>(in[2]);
      d.delta_yaw = static_cast<FLOAT_T>(in[3]);
      pc.motionUpdate(static_cast<FLOAT_T>(in[0]), d);
      break;
    }
    default:
      ros::tsdf_stub_clock() = -1.0;
      return 1;
  }
  ros::tsdf_stub_clock() = -1.0;
  std::memcpy(particles, static_cast<void*>(pc.m_particles.data()), n * sizeof(Particle));
  for (int k = 0; k < 6; ++k) ref_pose_io[k] = pc.ref_pose[k];
  return 0;
}

// The draws the reference makes for n particles: six std::normal_distribution<>{mean[k], sigma[k]} on mt19937(seed), per
// particle in the order x y z roll pitch yaw. out: n x 6 doubles.
void ref_pc_draws(uint32_t seed, const double mean[6], const double sigma[6], uint64_t n, double* out)
{
  std::mt19937 gen(seed);
  std::normal_distribution<> d0{mean[0], sigma[0]}, d1{mean[1], sigma[1]}, d2{mean[2], sigma[2]}, d3{mean[3], sigma[3]},
      d4{mean[4], sigma[4]}, d5{mean[5], sigma[5]};
  for (uint64_t i = 0; i < n; ++i)
  {
    out[6 * i + 0] = d0(gen);
    out[6 * i + 1] = d1(gen);
    out[6 * i + 2] = d2(gen);
    out[6 * i + 3] = d3(gen);
    out[6 * i + 4] = d4(gen);
    out[6 * i + 5] = d5(gen);
  }
}

}  // extern "C"
