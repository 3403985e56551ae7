// K3, joints + finish pass (k_cont_joints_finish): the per-waypoint get_joints and the per-trajectory unwrap /
// continuity / emergency scan of ControlIK's continuous mode in ONE kernel, every joint row written once.
//
// A trajectory is a recursion over its waypoints through two thin strands only -- the rate-limited elbow angle
// (previous_theta) and the unwrap / continuity / emergency chain (previous_sol), r2ik_control.cuh "Continuous mode, cut at
// its data dependences".  k_cont_targets / k_cont_thetas produce the code and the theta of every waypoint; this kernel
// does the rest.  A block owns 16 trajectories and walks them 8 waypoints at a time; per tile of 16 x 8 waypoints its 128
// threads run
//   C  one thread / waypoint    elbow circle + get_joints(theta) + Orbita3D limit            -> tile.j (raw joints)
//   D  one thread / trajectory  unwrap / clamp / continuity / emergency scan over the 8 waypoints: the statements of the
//                               serial kernel on rows in shared memory, controller state in shared memory between tiles
//                                                                                           -> tile.j, tile.reach, tile.state
// separated by block barriers, and then store the tile with row-contiguous requests (8 waypoints x 7 joints = 448 B per
// trajectory).  Against the separate raw-joints and finish kernels this replaces, the joints are written once instead of
// written (with 32 sectors per request), re-read and re-written: 112 B per waypoint less HBM traffic and a 2.9 ms scan
// kernel gone (profiles/r1_s43_continuous_ncu_full.txt).  The scan is a latency chain (8 steps per tile); the other
// resident blocks' phase C fills the pipes meanwhile, and the next tile's matrices are fetched into registers before phase D.
//
// Why not also the first two phases in here (one read of the pose tensor): tried first (profiles/r2_experiments.md, "K3
// tiled v1") -- the theta scan is a 1 000-step recursion per trajectory whose chain latency (~20 x the per-waypoint issue
// time) can only be hidden with all 65 536 trajectories in flight at once; inside a tile loop it serialises every block
// behind it: 22.5 ms against 10.2 ms.
//
// A waypoint whose straight-line get_joints hits a degenerate input (exact singularities that need previous_sol, never
// on physical data; the ABI's test_force_serial_mod sends ordinary waypoints down that route) is redone inside phase D
// with the literal routines, in order, from previous_sol as the scan has it at that point.
//
// Kept in a header of its own so that tests/hostsim can compile this exact kernel body for the host (one host thread per
// CUDA thread of a block, barriers emulated); r2ik_kernels.cu includes it in place.
#pragma once

#define R2IK_TILE_T 16     // trajectories per block
#define R2IK_TILE_W 8      // waypoints per tile
#define R2IK_TILE_BLOCK (R2IK_TILE_T * R2IK_TILE_W)
#ifndef R2IK_TILE_MINBLOCKS
#define R2IK_TILE_MINBLOCKS 4
#endif

struct ContTile {
  double theta[R2IK_TILE_T][R2IK_TILE_W];        // phase A: target theta; phase B: the rate-limited theta
  double j[R2IK_TILE_T][R2IK_TILE_W][7];         // phase C: raw joints; phase D: final joints
  uint8_t code[R2IK_TILE_T][R2IK_TILE_W];        // R2IK_WP_* (| R2IK_WP_SERIAL)
  uint8_t state[R2IK_TILE_T][R2IK_TILE_W];       // R2IK_STATE_*
  uint8_t reach[R2IK_TILE_T][R2IK_TILE_W];
};

__device__ __forceinline__ void tile_load_pose(const double *__restrict__ M, double m[16]) {
#if defined(__CUDA_ARCH__)
  load_mat4(M, m);
#else
  for (int k = 0; k < 12; ++k) m[k] = M[k];
  m[12] = 0.0; m[13] = 0.0; m[14] = 0.0; m[15] = 1.0;
#endif
}

// The serial get_joints of one waypoint (ctl:369-393 with previous_sol[0], [2]): out of line, the rare route.
#if defined(__CUDACC__)
__device__ __noinline__
#else
static
#endif
void tile_serial_joints(const r2ik::ArmConst &A, const R2ikCtlParams &par, const double *M, int kind, double theta, double prev0,
                        double prev2, double *j) {
  using namespace r2ik;
  double m[16];
  tile_load_pose(M, m);
  Solve S;
  double pos[3] = {m[3], m[7], m[11]};
  rotation_from_mat4(m, true, S.R);
  if (kind != R2IK_WP_UNREACHABLE) is_reachable_R<false>(A, pos, S);
  double jj[7];
  cont_raw_joints(A, par, kind, pos, S, theta, prev0, prev2, jj);
  for (int q = 0; q < 7; ++q) j[q] = jj[q];
}

// One waypoint of the finish scan on the joints j (in / out); returns the flag to store in `reachable` and sets
// `st_out` when the waypoint's state changes (the latched answer).  ctl:205-210, 306-313, 393-405.
__device__ __forceinline__ uint8_t tile_finish_waypoint(const r2ik::ArmConst &A, const R2ikCtlParams &par, const double *__restrict__ M,
                                                        const double *__restrict__ current_joints, int c, double theta,
                                                        R2ikTrajState &cs, double *j, uint8_t &st_out) {
  using namespace r2ik;
  if (cs.emergency_stop) {
#pragma unroll
    for (int q = 0; q < 7; ++q) j[q] = cs.previous_sol[q];
    st_out = R2IK_STATE_EMERGENCY;
    return 0;
  }
  const int kind = c & 0x7f;
  if (kind == R2IK_WP_INVALID) return 0;                     // joints are NaN, state is INVALID_ROTATION already
  if (!cs.has_previous_sol) {
#pragma unroll
    for (int q = 0; q < 7; ++q) cs.previous_sol[q] = current_joints[q];
    cs.has_previous_sol = 1;
    cs.init = 1;
  }
  cs.previous_theta = theta;
  double jj[7];
  if (c & R2IK_WP_SERIAL) tile_serial_joints(A, par, M, kind, theta, cs.previous_sol[0], cs.previous_sol[2], jj);
  else {
#pragma unroll
    for (int q = 0; q < 7; ++q) jj[q] = j[q];
  }
  cont_finish(cs, jj);
#pragma unroll
  for (int q = 0; q < 7; ++q) j[q] = jj[q];
  return kind == R2IK_WP_TARGET ? 1 : 0;
}

__global__ void __launch_bounds__(R2IK_TILE_BLOCK, R2IK_TILE_MINBLOCKS)
k_cont_joints_finish(const __grid_constant__ r2ik::ArmConst A, const __grid_constant__ R2ikCtlParams par,
                     const double *__restrict__ M, int64_t T, int W, const double *__restrict__ current_joints,
                     R2ikTrajState *__restrict__ states, const double *__restrict__ ws, double *__restrict__ joints,
                     uint8_t *__restrict__ reachable, uint8_t *__restrict__ state, int force_serial_mod) {
  using namespace r2ik;
  constexpr int TT = R2IK_TILE_T, WW = R2IK_TILE_W;
  __shared__ ContTile tile;
  __shared__ R2ikTrajState s_cs[TT];                     // the controller states of the block's trajectories (phase D)
  const int tid = threadIdx.x;
  const int tt = tid / WW, g = tid % WW;                 // phase C: waypoint g of trajectory tt
  const int64_t t0 = (int64_t)blockIdx.x * TT;
  const int64_t t = t0 + tt;
  const bool t_ok = t < T;
  if (tid < TT && t0 + tid < T) s_cs[tid] = states[t0 + tid];

  double m[16];
  double theta_c = 0.0;
  int c = R2IK_WP_INVALID;
  if (t_ok && g < W) {
    const size_t k0 = (size_t)t * W + g;
    tile_load_pose(M + 16 * k0, m);
    theta_c = ws[k0];
    c = reachable[k0];
  }
  for (int w0 = 0; w0 < W; w0 += WW) {
    const int nw = W - w0 < WW ? W - w0 : WW;
    const bool wp_ok = t_ok && g < nw;
    // ---- phase C: one thread / waypoint
    if (wp_ok) {
      double j[7];
      bool serial = false;
      if (c == R2IK_WP_INVALID) {
#pragma unroll
        for (int q = 0; q < 7; ++q) j[q] = NAN;
      } else {
        Solve S;
        double pos[3] = {m[3], m[7], m[11]};
        rotation_from_mat4(m, true, S.R);
        if (c == R2IK_WP_UNREACHABLE) is_reachable_R<true>(A, pos, S);       // ctl:369
        else circle_of_reachable(A, pos, S);   // reachable (phase 1 decided): the elbow circle is all get_joints needs
        double st, ct, E[3];
        sincos_any(theta_c, st, ct);
        // straight-line get_joints only; a degenerate input (exact singularity: needs previous_sol) is left to phase D
        serial = !get_joints_impl<false>(A, S, ct, st, 0.0, 0.0, j, E);
        // test hook (ABI parameter test_force_serial_mod = m > 0): every m-th waypoint takes the serial route although it
        // does not need it, so that the route is exercised on ordinary data
        if (force_serial_mod > 0 && ((size_t)t * W + w0 + g) % (size_t)force_serial_mod == 0) serial = true;
        if (!serial) limit_orbita3d_wrist(j, par.orbita3d_max_angle);
      }
      tile.code[tt][g] = (uint8_t)(serial ? (c | R2IK_WP_SERIAL) : c);
      tile.theta[tt][g] = theta_c;
      tile.state[tt][g] = 0xff;                             // "unchanged"
#pragma unroll
      for (int q = 0; q < 7; ++q) tile.j[tt][g][q] = j[q];
    }
    // the next tile's inputs travel while the scan and the stores run
    if (t_ok && w0 + WW + g < W) {
      const size_t kn = (size_t)t * W + w0 + WW + g;
      tile_load_pose(M + 16 * kn, m);
      theta_c = ws[kn];
      c = reachable[kn];
    }
    __syncthreads();
    // ---- phase D: one thread / trajectory, the statements of the serial kernel on the tile's rows
    if (tid < TT && t0 + tid < T) {
      R2ikTrajState cs = s_cs[tid];
      const int64_t td = t0 + tid;
      for (int e = 0; e < nw; ++e) {
        uint8_t st_new = 0xff;
        tile.reach[tid][e] = tile_finish_waypoint(A, par, M + 16 * ((size_t)td * W + w0 + e), current_joints + 7 * td,
                                                  tile.code[tid][e], tile.theta[tid][e], cs, &tile.j[tid][e][0], st_new);
        if (st_new != 0xff) tile.state[tid][e] = st_new;
      }
      s_cs[tid] = cs;
    }
    __syncthreads();
    // ---- store the tile: per trajectory nw x 7 doubles and nw flag bytes are contiguous in global memory; the 8
    // lanes of a trajectory's group write 64 contiguous bytes per request
    if (t_ok) {
      const int row = nw * 7;
      double *dst = joints + ((size_t)t * W + w0) * 7;
      const double *src = &tile.j[tt][0][0];
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const int o = g + WW * i;
        if (o < row) dst[o] = src[o];
      }
      if (g < nw) {
        reachable[(size_t)t * W + w0 + g] = tile.reach[tt][g];
        if (tile.state[tt][g] != 0xff) state[(size_t)t * W + w0 + g] = tile.state[tt][g];     // only the latched answer overwrites
      }
    }
    __syncthreads();   // the next tile's phase C writes the tile
  }
  if (tid < TT && t0 + tid < T) states[t0 + tid] = s_cs[tid];
}
