// K3 finish scan, lane-parallel form (k_cont_finish_lanes).  Kept in a header of its own so that tests/hostsim can
// compile this exact kernel body for the host, with the warp votes emulated by 32 host threads per warp; r2ik_kernels.cu
// includes it in place.  Needs r2ik_control.cuh (R2ikTrajState, waypoint codes, pymod) and, on the device, nothing else.
#pragma once

// Lane-parallel form of the finish scan: 8 lanes per trajectory, lane q < 7 owns joint q (its previous_sol[q] stays in
// a register), the per-joint statements of allow_multiturn / multiturn_safety_check / continuity_check (utl:493-589) run
// in parallel over the lanes and their "any joint" conditions are ballots within the 8-lane group.  Against one thread
// per trajectory this is 8x the threads (the scan is a latency chain: 14 warps / SM could not hide it), a 7x shorter
// chain per waypoint, and one 56-byte row per group and request instead of 7 requests of 32 sectors.  A trajectory
// that meets a waypoint whose get_joints needs the serial route (R2IK_WP_SERIAL: exact singularities, out-of-range
// magnitudes) stops there, keeps its state, and is finished by k_cont_finish_direct<fixup>.
#define R2IK_FIN8_BLOCK 128
#ifndef R2IK_FIN_LANES
#define R2IK_FIN_LANES 4   // lanes per trajectory of the finish scan (2, 4 or 8; a build-time choice)
#endif
// Constants of the scan as a kernel parameter: constant-bank operands of DADD / DSETP instead of 64-bit immediates
// that cost a UMOV pair per use (23 of the first version's 206 instructions per waypoint).
struct ScanConst { double pi, two_pi, four_pi, eight_pi, lim; };

// utl:486-490 angle_diff for the scan: pymod_2pi's exact subtraction ladder on constant-bank operands; arguments
// outside [-2 pi, 8 pi) (never for joints within +-6 pi) take the generic routine.
__device__ __forceinline__ double angle_diff_scan(const ScanConst &K, double a, double b) {
  double x = (a - b) + K.pi;
  if (!(x >= -K.two_pi && x < K.eight_pi)) return pymod(x, kTwoPi) - kPi;
  double r = x < 0.0 ? x + K.two_pi : x;
  r = r >= K.four_pi ? r - K.four_pi : r;
  r = r >= K.two_pi ? r - K.two_pi : r;
  return r - K.pi;
}

// G lanes per trajectory (G = 2, 4, 8): lane g owns joints g, g + G, ... < 7; the last lane of a group also writes the
// flags.  Fewer lanes per trajectory share the per-waypoint overhead (loop, code / theta loads, votes, addresses) among
// more trajectories per warp, more lanes shorten the chain and raise the thread count.
template <int G>
__global__ void __launch_bounds__(R2IK_FIN8_BLOCK)
k_cont_finish_lanes(const __grid_constant__ ScanConst K, int64_t T, int W, const double *__restrict__ current_joints,
                    R2ikTrajState *__restrict__ states, const double *__restrict__ ws, double *__restrict__ joints,
                    uint8_t *__restrict__ reachable, uint8_t *__restrict__ state) {
  constexpr int NJ = (7 + G - 1) / G;                   // joints per lane (the last slot may be empty)
  constexpr unsigned GM = (1u << G) - 1u;
  const int lane = threadIdx.x & 31, g = lane & (G - 1);
  const unsigned gshift = (unsigned)(lane & ~(G - 1));
  const int64_t t_raw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const bool live = t_raw < T;
  const int64_t t = live ? t_raw : T - 1;               // idle groups of the last warp read a valid trajectory, write nothing
  const bool flagger = g == G - 1;
  bool has[NJ];
  int qj[NJ];
  double prev[NJ];
  const R2ikTrajState *cs = states + t;
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    has[i] = g + i * G < 7;
    qj[i] = has[i] ? g + i * G : 6;
    prev[i] = has[i] ? cs->previous_sol[qj[i]] : 0.0;
  }
  double previous_theta = cs->previous_theta;
  int has_prev = cs->has_previous_sol, init = cs->init, emergency_stop = cs->emergency_stop, emergency_bits = cs->emergency_bits;
  bool stopped = false;
  const size_t base = (size_t)t * W;
  double *pj = joints + 7 * base;
  uint8_t *pc = reachable + base;
  const double *pth = ws + base;
  for (int w = 0; w < W; ++w, pj += 7) {
    const int c = pc[w];
    const double theta = pth[w];
    const int kind = c & 0x7f;
    double j[NJ], m[NJ];
    bool hit = false, viol = false;
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      j[i] = has[i] ? pj[qj[i]] : 0.0;
      // utl:493-505 allow_multiturn; utl:535-568 clamp of joints 0 / 2 / 6; ctl:395-400 continuity against previous_sol
      // (max step 0.5 rad for joints 0-3, 1 rad for joints 4-6, utl:571-589)
      m[i] = prev[i] + angle_diff_scan(K, j[i], prev[i]);
      const bool clampq = has[i] && (qj[i] == 0 || qj[i] == 2 || qj[i] == 6);
      hit = hit || (clampq && (m[i] > K.lim || m[i] < -K.lim));
      viol = viol || (has[i] && fabs(angle_diff_scan(K, m[i], prev[i])) > (qj[i] < 4 ? 0.5 : 1.0));
    }
    // The ordinary waypoint -- valid code, no serial route, state initialised and not latched, no clamp, continuous --
    // is recognised for the whole warp at once; everything else takes the full statement order below.
    const bool ordinary = live && !stopped && !emergency_stop && has_prev && !init && kind != R2IK_WP_INVALID &&
                          !(c & R2IK_WP_SERIAL) && !hit && !viol;
    if (__all_sync(0xffffffffu, ordinary || !live)) {
      if (live) {
        previous_theta = theta;
#pragma unroll
        for (int i = 0; i < NJ; ++i)
          if (has[i]) { prev[i] = m[i]; pj[qj[i]] = m[i]; }
        if (flagger) pc[w] = kind == R2IK_WP_TARGET ? 1 : 0;
      }
      continue;
    }
    stopped = stopped || (!emergency_stop && kind != R2IK_WP_INVALID && (c & R2IK_WP_SERIAL));
    const bool emg = emergency_stop != 0;
    const bool work = live && !stopped && !emg && kind != R2IK_WP_INVALID;
    if (work && !has_prev) {                            // ctl:306-313
#pragma unroll
      for (int i = 0; i < NJ; ++i) prev[i] = has[i] ? current_joints[7 * t + qj[i]] : 0.0;
      has_prev = 1; init = 1;
    }
    if (work) previous_theta = theta;
    double nj[NJ];
    int bits = 0;
    bool viol2 = false;
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      nj[i] = prev[i] + angle_diff_scan(K, j[i], prev[i]);
      bool h = false;
      if (has[i] && (qj[i] == 0 || qj[i] == 2 || qj[i] == 6)) {
        if (nj[i] > K.lim) { nj[i] = K.lim; h = true; }
        if (nj[i] < -K.lim) { nj[i] = -K.lim; h = true; }
      }
      const unsigned hits = (__ballot_sync(0xffffffffu, work && h) >> gshift) & GM;   // bit g' <=> joint g' + i G
      if (0 / G == i && (hits >> (0 % G) & 1u)) bits |= R2IK_EMG_SHOULDER_PITCH;
      if (2 / G == i && (hits >> (2 % G) & 1u)) bits |= R2IK_EMG_ELBOW_YAW;
      if (6 / G == i && (hits >> (6 % G) & 1u)) bits |= R2IK_EMG_WRIST_YAW;
    }
    if (bits) { emergency_stop = 1; emergency_bits |= bits; }
#pragma unroll
    for (int i = 0; i < NJ; ++i)
      viol2 = viol2 || (has[i] && fabs(angle_diff_scan(K, nj[i], prev[i])) > (qj[i] < 4 ? 0.5 : 1.0));
    const bool disc = ((__ballot_sync(0xffffffffu, work && !init && viol2) >> gshift) & GM) != 0;
    if (disc) {
#pragma unroll
      for (int i = 0; i < NJ; ++i) nj[i] = prev[i];
      emergency_stop = 1; emergency_bits |= R2IK_EMG_DISCONTINUITY;
    }
    if (work) {
      init = 0;
      if (!emergency_stop) {
#pragma unroll
        for (int i = 0; i < NJ; ++i) prev[i] = nj[i];
      }
    }
    if (live && !stopped) {
      if (emg) {                                        // ctl:205-210: latched, answers with the last solution
#pragma unroll
        for (int i = 0; i < NJ; ++i)
          if (has[i]) pj[qj[i]] = prev[i];
        if (flagger) { pc[w] = 0; state[base + w] = R2IK_STATE_EMERGENCY; }
      } else {
        if (work) {
#pragma unroll
          for (int i = 0; i < NJ; ++i)
            if (has[i]) pj[qj[i]] = nj[i];
        }
        if (flagger) pc[w] = (work && kind == R2IK_WP_TARGET) ? 1 : 0;
      }
    }
  }
  if (live) {
    R2ikTrajState *o = states + t;
#pragma unroll
    for (int i = 0; i < NJ; ++i)
      if (has[i]) o->previous_sol[qj[i]] = prev[i];
    if (flagger) {
      o->previous_theta = previous_theta;
      o->has_previous_sol = has_prev; o->init = init; o->emergency_stop = emergency_stop; o->emergency_bits = emergency_bits;
    }
  }
}
