// r2ik_host.h -- host-side derivation of the per-arm constants (SymbolicIK.__init__ and the
// loop-invariant parts of get_joints / make_elbow_projection), shared by libr2ik.so
// (r2ik_create) and the test-only host harness.  Runs once per handle; not the hot path.
#pragma once

#include "r2ik_device.cuh"

namespace r2ik {

inline double deg2rad(double d) { return d * (kPi / 180.0); }

// sik:26-83 + utl:26-43 + sik:728-737 + sik:653-672
inline void derive_constants(const R2ikArmConfig &c, ArmConst &A, R2ikArmConstants &pub) {
  const double L1 = c.upper_arm_size, L2 = c.forearm_size;
  for (int k = 0; k < 3; ++k) A.s[k] = c.shoulder_position[k];
  A.L1 = L1; A.L2 = L2; A.L12 = L1 + L2;
  A.L1sq = L1 * L1; A.L2sq = L2 * L2;
  A.wo[0] = -c.tip_position[0]; A.wo[1] = c.tip_position[1]; A.wo[2] = c.tip_position[2];
  A.to[0] = -c.tip_position[0]; A.to[1] = c.tip_position[1]; A.to[2] = 0.0;
  A.tip_z = c.tip_position[2];
  double gripper = sqrt(c.tip_position[0] * c.tip_position[0] + c.tip_position[1] * c.tip_position[1] +
                        c.tip_position[2] * c.tip_position[2]);
  A.max_arm_length = L1 + L2 + gripper;
  A.d_min = sqrt(L1 * L1 + L2 * L2 - 2 * L1 * L2 * cos(deg2rad(180 - c.elbow_limit_deg)));
  A.proj_margin = c.projection_margin;
  A.backward_limit = c.backward_limit;
  A.nvm = c.normal_vector_margin;
  A.rL = sin(deg2rad(c.wrist_limit_deg)) * L2;
  A.rLsq = A.rL * A.rL;
  A.hL = sqrt(L2 * L2 - A.rL * A.rL);
  A.sing_coeff = c.singularity_limit_coeff;
  A.sing_offset = c.singularity_offset;
  A.elbow_limit = deg2rad(c.elbow_limit_deg);
  A.side = (double)c.side;

  // shoulder frame: M_torso_shoulder = R("xyz", rad(offset)) * R("xyz", [0, pi/2, 0])
  double ox = deg2rad(c.shoulder_orientation_deg[0]), oy = deg2rad(c.shoulder_orientation_deg[1]),
         oz = deg2rad(c.shoulder_orientation_deg[2]);
  Quat qo = quat_from_euler(0, 1, 2, false, ox, oy, oz);
  Quat q90 = quat_from_euler(0, 1, 2, false, 0.0, kHalfPi, 0.0);
  Quat qs = quat_mul(qo, q90);
  double qn = sqrt(qs.x * qs.x + qs.y * qs.y + qs.z * qs.z + qs.w * qs.w);  // Rotation.__mul__ normalises
  qs.x /= qn; qs.y /= qn; qs.z /= qn; qs.w /= qn;
  double Mts[9];
  quat_to_matrix(qs, Mts);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A.Mst[3 * i + j] = Mts[3 * j + i];
  for (int i = 0; i < 3; ++i)
    A.Pst[i] = (-A.Mst[3 * i]) * A.s[0] + (-A.Mst[3 * i + 1]) * A.s[1] + (-A.Mst[3 * i + 2]) * A.s[2];

  // singularity positions: T_torso_shoulder = [R(offset) | s] applied to points on the local -y axis
  double Ro[9];
  quat_to_matrix(qo, Ro);
  double ey = -L1 * c.side, wy = -(L1 + L2) * c.side;
  double ws[3];
  for (int i = 0; i < 3; ++i) {
    A.es[i] = Ro[3 * i] * 0.0 + Ro[3 * i + 1] * ey + Ro[3 * i + 2] * 0.0 + A.s[i];
    ws[i] = Ro[3 * i] * 0.0 + Ro[3 * i + 1] * wy + Ro[3 * i + 2] * 0.0 + A.s[i];
  }

  // singularity-limit plane (sik:653-672)
  double alpha = atan2(-c.singularity_limit_coeff, 1.0);
  double Ml[9];
  rot_from_euler_xyz(0.0, alpha, 0.0, Ml);
  double P[3];
  for (int i = 0; i < 3; ++i) P[i] = Ml[3 * i + 2] * (-c.singularity_offset) + A.es[i];
  // n1 = T_limits [1,0,0,1], n2 = T_limits [0,1,0,1]; v1 = n1 - P, v2 = n2 - P
  double v1[3], v2[3];
  for (int i = 0; i < 3; ++i) {
    v1[i] = (Ml[3 * i] + P[i]) - P[i];
    v2[i] = (Ml[3 * i + 1] + P[i]) - P[i];
  }
  double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
  double n3 = sqrt(v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2]);
  for (int i = 0; i < 3; ++i) { v3[i] /= n3; A.plP[i] = P[i]; A.plV[i] = v3[i]; }
  double dist = (A.s[0] - P[0]) * v3[0] + (A.s[1] - P[1]) * v3[1] + (A.s[2] - P[2]) * v3[2];
  double ds2 = 0.0;
  for (int i = 0; i < 3; ++i) {
    A.plC[i] = A.s[i] - dist * v3[i];
    ds2 += (A.s[i] - A.plC[i]) * (A.s[i] - A.plC[i]);
  }
  double dsn = sqrt(ds2);
  A.plRho = sqrt(L1 * L1 - dsn * dsn);

  pub.gripper_size = gripper;
  pub.max_arm_length = A.max_arm_length;
  pub.shoulder_wrist_min_distance = A.d_min;
  for (int i = 0; i < 3; ++i) {
    pub.elbow_singularity_position[i] = A.es[i];
    pub.wrist_singularity_position[i] = ws[i];
  }
}

}  // namespace r2ik
