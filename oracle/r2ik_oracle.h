/*
 * r2ik_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, libm) of the Reachy2 symbolic IK hot path of
 * pollen-robotics/reachy2_symbolic_ik, used only as the parity checker for the
 * CUDA library and as the `cpu_baseline` leg of bench.py.  Nothing under
 * reachy2_symbolic_ik_b200/ may include, link or call it.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the reference checkout, src/reachy2_symbolic_ik/...).  Part of the reference
 * arithmetic lives in un-vendored third-party code: scipy.spatial.transform.
 * Rotation (reference pins scipy == 1.8.0, setup.cfg:18-20; this container has
 * scipy 1.18.1) and numpy (linalg.lstsq/norm, isclose, linspace).  Their
 * published algorithms are restated here (scipy _rotation_xp.py of 1.18.1).
 *
 * Parity pinning: tests/test_oracle_golden.py checks this file against
 * tests/golden/ *.npz, which were produced by running the UNMODIFIED reference
 * (tests/golden/gen_golden.py, numpy 2.3.5 / scipy 1.18.1).
 */
#ifndef R2IK_ORACLE_H
#define R2IK_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* state codes shared with the product by convention */
enum {
  ORC_STATE_REACHABLE = 0,         /* "reachable"                 symbolic_ik.py:234 */
  ORC_STATE_POSE_OUT_OF_REACH = 1, /* "Pose out of reach"         symbolic_ik.py:300 */
  ORC_STATE_BACKWARD_POSE = 2,     /* "Backward pose"             symbolic_ik.py:306 */
  ORC_STATE_WRIST_OUT_OF_RANGE = 3,/* "wrist out of range"        symbolic_ik.py:159 */
  ORC_STATE_LIMITED_BY_WRIST = 4,  /* "limited by wrist"          symbolic_ik.py:262 */
  ORC_STATE_SHOULD_NOT_HAPPEN = 5, /* "out of reach - should not happen" :281 */
  ORC_STATE_LIMITED_BY_SHOULDER = 6, /* "limited by shoulder"     control_ik.py:363,452 */
  ORC_STATE_EMPTY = 7,             /* ""  (continuous success)    control_ik.py:297 */
  ORC_STATE_EMERGENCY = 8,         /* emergency string            utils.py:544-566,584-586 */
  ORC_STATE_INVALID_ROTATION = 9   /* scipy from_matrix ValueError (det <= 0) */
};

enum { ORC_POSE_EULER6 = 0, ORC_POSE_MAT4 = 1 };

/* emergency reason bits */
enum {
  ORC_EMG_SHOULDER_PITCH = 1,
  ORC_EMG_ELBOW_YAW = 2,
  ORC_EMG_WRIST_YAW = 4,
  ORC_EMG_DISCONTINUITY = 8
};

typedef struct {
  double shoulder_position[3];
  double shoulder_orientation_deg[3];
  double upper_arm_size;
  double forearm_size;
  double tip_position[3];
  double elbow_limit_deg;         /* 127   */
  double wrist_limit_deg;         /* 42.5  */
  double projection_margin;       /* 1e-8  */
  double backward_limit;          /* 0.02  */
  double normal_vector_margin;    /* 1e-7  */
  double singularity_offset;      /* 0.03 (SymbolicIK) / -1.01 (ControlIK non-DVT) */
  double singularity_limit_coeff; /* 1.0   */
  int32_t side;                   /* +1 r_arm, -1 l_arm */
  int32_t pad_;
} orc_arm_config;

typedef struct {
  double previous_theta;
  double previous_sol[7];
  int32_t has_previous_sol; /* 0 => first call / timeout: re-init from current joints */
  int32_t init;             /* ControlIK.init */
  int32_t emergency_stop;   /* ControlIK.emergency_stop */
  int32_t emergency_bits;   /* reasons accumulated in ControlIK.emergency_state */
} orc_ctl_state;

typedef struct {
  double preferred_theta;      /* per-call preferred_theta (already mirrored for l_arm) */
  double preferred_theta_ctor; /* ControlIK.preferred_theta[arm] */
  double interval_limit[2];    /* already mirrored for l_arm */
  double d_theta_max;          /* 0.01 */
  double orbita3d_max_angle;   /* deg2rad(42.5) */
  int32_t nb_search_points;    /* 20 (discrete) */
  int32_t nb_search_points_continuous; /* 10 */
} orc_ctl_params;

/* --- single-pose entry points (used by tests) ------------------------------ */

/* SymbolicIK.is_reachable + theta_to_joints(theta) on a fresh solver.
 * pose: 6 doubles (x,y,z,roll,pitch,yaw) if kind==EULER6, 16 doubles row-major if MAT4.
 * theta: NULL => theta_interval[0].  Outputs NaN-filled when unreachable. */
int orc_symik_solve(const orc_arm_config *cfg, int pose_kind, const double *pose,
                    const double *theta, const double prev_joints[7],
                    uint8_t *reachable, uint8_t *state, double interval[2],
                    double joints[7], double elbow[3]);

/* --- batch entry points (OpenMP over poses / trajectories) ----------------- */

void orc_symik_batch(const orc_arm_config *cfg, int pose_kind, const double *poses,
                     const double *theta /* nullable, n */, int64_t n,
                     uint8_t *reachable, uint8_t *state, double *interval /* n*2 */,
                     double *joints /* n*7 */, double *elbow /* n*3 */);

/* is_reachable_no_limits + get_joints(theta, previous_joints) (symbolic_ik.py:85-119, 697-863).
 * prev_joints: nullable n*7 (zeros = the reference default); projected: nullable n, 1 where
 * make_elbow_projection fired (the reference then returns a 3-vector elbow, :714 / :863). */
void orc_symik_no_limits_batch(const orc_arm_config *cfg, int pose_kind, const double *poses,
                               const double *theta /* n */, const double *prev_joints, int64_t n,
                               double *joints, double *elbow, uint8_t *projected);

/* get_elbow_position for K thetas per pose after is_reachable (no_limits = 0) or after
 * is_reachable_no_limits (no_limits = 1) (symbolic_ik.py:684-695; circle stored at :114-116 / :197).
 * projected: nullable n*K, 1 where get_joints(theta) would take the elbow-projection branch. */
void orc_elbow_positions_batch(const orc_arm_config *cfg, int pose_kind, const double *poses,
                               const double *thetas /* n*K */, int K, int64_t n, int no_limits,
                               double *elbows /* n*K*3 */, uint8_t *projected);

/* Result of one scalar call sequence (orc_symik_scalar); same layout as R2ikScalarResult of include/r2ik.h. */
typedef struct orc_scalar_result {
  double interval[2];               /* theta_interval; NaN when the call failed                           */
  double joints[7];                 /* get_joints(theta, previous_joints)                                 */
  double elbow[3];                  /* elbow returned by get_joints (after the projection, if it fired)   */
  double elbow_on_circle[3];        /* get_elbow_position(theta)                                          */
  double goal_position_solved[3];   /* self.goal_pose[0] / self.wrist_position after is_reachable[_no_limits] */
  double wrist_position_solved[3];
  double goal_position[3];          /* ... after get_joints                                               */
  double wrist_position[3];
  int32_t reachable, state, projected, reserved;
} orc_scalar_result;

void orc_symik_scalar(const orc_arm_config *cfg, const double *goal_pose6, int no_limits,
                      const double *theta_opt /* nullable: theta_interval[0] */,
                      const double *prev_joints /* nullable 7 */, orc_scalar_result *out);

/* ControlIK.symbolic_inverse_kinematics(name, M, "discrete") (control_ik.py:162-274,409-462).
 * prev_joints: ControlIK.previous_sol[arm]; current_joints: per-call current_joints
 * (both 7 doubles, broadcast to every pose).  emergency: emergency bits raised. */
void orc_ctl_discrete_batch(const orc_arm_config *cfg, const orc_ctl_params *par,
                            const double *M /* n*16 */, int64_t n,
                            const double prev_joints[7], const double current_joints[7],
                            double *joints /* n*7 */, uint8_t *reachable, uint8_t *state,
                            uint8_t *emergency);

/* ControlIK.symbolic_inverse_kinematics(name, M, "continuous") over T trajectories of
 * W waypoints (control_ik.py:276-407).  current_joints: T*7, current_pose: T*16
 * (used at (re)initialisation).  st: T states (in/out). */
void orc_ctl_continuous_batch(const orc_arm_config *cfg, const orc_ctl_params *par,
                              const double *M /* T*W*16 */, int64_t T, int32_t W,
                              const double *current_joints, const double *current_pose,
                              orc_ctl_state *st,
                              double *joints /* T*W*7 */, uint8_t *reachable, uint8_t *state);

/* Grid sweep of is_reachable: counts[v] = #orientations in [ori_begin,ori_end) reachable at
 * voxel v; voxel (ix,iy,iz) centre = origin + (ix,iy,iz)*step, v = (ix*dims[1]+iy)*dims[2]+iz. */
void orc_reach_map(const orc_arm_config *cfg, const double origin[3], const double step[3],
                   const int32_t dims[3], const double *orientations_euler /* n_ori*3 */,
                   int32_t ori_begin, int32_t ori_end, uint32_t *counts);

/* --- helpers exported for unit tests against numpy/scipy ------------------- */
void orc_euler_xyz_from_matrix(const double m9[9], double euler[3], int *status);
void orc_matrix_from_euler_xyz(const double euler[3], double m9[9]);
void orc_limit_orbita3d_joints(const double in3[3], double max_angle, double out3[3]);
double orc_angle_diff(double a, double b);
double orc_pymod(double a, double m);
void orc_limit_theta_to_interval(double theta, double previous_theta, const double interval[2], double *out);
void orc_rotation_matrix_from_vector(const double v[3], double m9[9]);
void orc_interval_limit(int side, int low_elbow, double out[2]);
int orc_max_threads(void);
void orc_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
