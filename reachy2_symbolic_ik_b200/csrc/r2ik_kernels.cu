// r2ik_kernels.cu -- sm_100a kernels and the C ABI of libr2ik.so (include/r2ik.h).
//
//   K1 k_symik_solve     one thread / pose   : is_reachable + get_joints
//   K1b k_symik_no_limits, k_elbow_positions : is_reachable_no_limits, get_elbow_position
//   K2 k_ctl_discrete    one lane / pose, warp-cooperative K-sample elbow search
//   K3 k_ctl_continuous  one thread / trajectory, sequential over waypoints
//   K4 k_reach_map       one thread / voxel, orientation table staged in shared memory
//
// The work is scalar FP64 (sincos / atan2 / sqrt / FMA chains): it runs on the FP64 pipe, not on
// tensor cores; HBM traffic is a few hundred bytes per pose.  There is no host implementation
// of any entry point in this library.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "r2ik_control.cuh"
#include "r2ik_host.h"

using namespace r2ik;

#define R2IK_BLOCK 128
#ifndef R2IK_K1_MINBLOCKS
#define R2IK_K1_MINBLOCKS 4   // resident blocks / SM the register allocation of K1 is held to
#endif

// ---------------------------------------------------------------------------------------
// vectorised global memory helpers
// ---------------------------------------------------------------------------------------
// 128-bit read-only load: pose buffers must be 16-byte aligned (checked by the C entry points)
__device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }

// 3x4 top of a row-major 4x4 (the last row is never read): 6 x 128-bit loads
__device__ __forceinline__ void load_mat4(const double *__restrict__ M, double m[16]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double2 a = ldg2(M + 4 * r), b = ldg2(M + 4 * r + 2);
    m[4 * r] = a.x; m[4 * r + 1] = a.y; m[4 * r + 2] = b.x; m[4 * r + 3] = b.y;
  }
  m[12] = 0.0; m[13] = 0.0; m[14] = 0.0; m[15] = 1.0;
}

// Goal position + goal rotation matrix of pose i.  false <=> invalid rotation (det <= 0).
template <int KIND>
__device__ __forceinline__ bool load_pose(const double *__restrict__ poses, int64_t i, bool snap, double pos[3], double R[9]) {
  if (KIND == R2IK_POSE_EULER6) {
    const double *p = poses + 6 * i;
    double2 a = ldg2(p), b = ldg2(p + 2), c = ldg2(p + 4);
    pos[0] = a.x; pos[1] = a.y; pos[2] = b.x;
    rot_from_euler_xyz(b.y, c.x, c.y, R);
    return true;
  } else {
    double m[16];
    load_mat4(poses + 16 * i, m);
    pos[0] = m[3]; pos[1] = m[7]; pos[2] = m[11];
    return rotation_from_mat4(m, snap, R);
  }
}

__device__ __forceinline__ void store_nan(double *p, int n) {
  for (int k = 0; k < n; ++k) p[k] = NAN;
}

// ---------------------------------------------------------------------------------------
// K1: SymbolicIK.is_reachable + theta_to_joints
// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(R2IK_BLOCK, R2IK_K1_MINBLOCKS)
k_symik_solve(const __grid_constant__ ArmConst A, const double *__restrict__ poses, const double *__restrict__ theta,
              const double *__restrict__ prev_joints, int64_t n, uint8_t *__restrict__ reachable,
              uint8_t *__restrict__ state, double *__restrict__ interval, double *__restrict__ joints,
              double *__restrict__ elbow) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
#ifdef R2IK_K1_PREFETCH
  {  // pull the pose a later wave of blocks will read into L2
    const int64_t pf = i + (int64_t)R2IK_K1_PREFETCH * R2IK_BLOCK;
    if (pf < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(poses + (KIND == R2IK_POSE_MAT4 ? 16 : 6) * pf));
  }
#endif
  double prev0 = 0.0, prev2 = 0.0;
  if (prev_joints) { prev0 = prev_joints[0]; prev2 = prev_joints[2]; }
  double pos[3];
  Solve S;
  Reach rc;
  if (load_pose<KIND>(poses, i, false, pos, S.R)) {
    rc = is_reachable_R<false>(A, pos, S);
  } else {
    rc.state = R2IK_STATE_INVALID_ROTATION; rc.i0 = NAN; rc.i1 = NAN;
  }
  bool ok = rc.state == R2IK_STATE_REACHABLE;
  reachable[i] = ok ? 1 : 0;
  state[i] = (uint8_t)rc.state;
  if (interval) { interval[2 * i] = rc.i0; interval[2 * i + 1] = rc.i1; }
  if (!joints && !elbow) return;
  double j[7], E[3];
  if (ok) {
    double ct = rc.c0, st = rc.s0;   // theta_interval[0]: its cos / sin come with the interval
    if (theta) sincos(theta[i], &st, &ct);
    get_joints_cs(A, S, ct, st, prev0, prev2, j, E);
  } else {
#pragma unroll
    for (int k = 0; k < 7; ++k) j[k] = NAN;
    E[0] = NAN; E[1] = NAN; E[2] = NAN;
  }
  if (joints) {
#pragma unroll
    for (int k = 0; k < 7; ++k) joints[7 * i + k] = j[k];
  }
  if (elbow) { elbow[3 * i] = E[0]; elbow[3 * i + 1] = E[1]; elbow[3 * i + 2] = E[2]; }
}

// ---------------------------------------------------------------------------------------
// K1, streaming form (MAT4 poses, all outputs, theta_interval[0]): persistent warps, each
// owning tiles of 32 consecutive poses.
//   in : one TMA bulk copy (cp.async.bulk, 4 KB contiguous) per tile into the warp's shared
//        buffer, completion on the warp's mbarrier; the copy of the warp's NEXT tile is issued
//        as soon as the current poses are in registers, so it lands during the solve;
//   out: results are staged in shared memory and leave as five contiguous TMA bulk stores per
//        tile (joints 1792 B, interval 512 B, elbow 768 B, flags 2 x 32 B).
// No thread waits on a global load or drains global stores; the only synchronisation is
// __syncwarp (no block barrier, so warps with early-out poses do not wait for their neighbours).
// The same kernel runs with pinned HOST pointers (UVA): bulk copies are the PCIe-friendly
// access pattern, which is the zero-copy end-to-end path of SymbolicIK.is_reachable_batch_host.
// ---------------------------------------------------------------------------------------
#define R2IK_TILE 32
#define R2IK_STREAM_WARPS (R2IK_BLOCK / 32)

struct __align__(128) R2ikWarpStage {
  double in[R2IK_TILE * 16];       // 4096 B  poses of the tile, row-major 4x4
  double joints[R2IK_TILE * 7];    // 1792 B
  double interval[R2IK_TILE * 2];  //  512 B
  double elbow[R2IK_TILE * 3];     //  768 B
  uint8_t reach[R2IK_TILE];        //   32 B
  uint8_t state[R2IK_TILE];        //   32 B
  unsigned long long bar;          // mbarrier of the input buffer
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Tile scheduler: sched[0] = next tile to hand out, sched[1] = warps that have finished.  Warps
// draw tiles with an atomic (one draw ahead, so its latency is hidden by the solve); the last warp
// to finish zeroes both words, so the slot is ready for the next launch on the same stream.
__device__ __forceinline__ int draw_tile(unsigned *sched, int lane) {
  unsigned t = 0;
  if (lane == 0) t = atomicAdd(sched, 1u);
  return (int)__shfl_sync(0xffffffffu, t, 0);
}

__global__ void __launch_bounds__(R2IK_BLOCK, R2IK_K1_MINBLOCKS)
k_symik_solve_stream(const __grid_constant__ ArmConst A, const double *__restrict__ poses, int64_t n, int n_tiles,
                     unsigned *__restrict__ sched, uint8_t *__restrict__ reachable, uint8_t *__restrict__ state,
                     double *__restrict__ interval, double *__restrict__ joints, double *__restrict__ elbow) {
  __shared__ R2ikWarpStage stage[R2IK_STREAM_WARPS];
  const int lane = threadIdx.x & 31;
  R2ikWarpStage &W = stage[threadIdx.x >> 5];
  const uint32_t bar = smem_u32(&W.bar), in_s = smem_u32(W.in);
  const int last_m = (int)(n - (int64_t)(n_tiles - 1) * R2IK_TILE);   // poses in the last tile (1..32)

  int tile = draw_tile(sched, lane);
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_async_smem();   // make the initialised barrier visible to the async proxy
    if (tile < n_tiles) {
      const uint32_t bytes = (uint32_t)(tile == n_tiles - 1 ? last_m : R2IK_TILE) * 128u;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(in_s, poses + (size_t)tile * (R2IK_TILE * 16), bytes, bar);
    }
  }
  __syncwarp();
  uint32_t parity = 0;
#pragma unroll 1
  while (tile < n_tiles) {
    const int next = draw_tile(sched, lane);   // consumed after the solve: the atomic's latency is hidden
    mbar_wait(bar, parity);
    parity ^= 1u;
    double mm[16];
    {
      const double2 *src = reinterpret_cast<const double2 *>(W.in + lane * 16);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        double2 a = src[2 * r], b = src[2 * r + 1];
        mm[4 * r] = a.x; mm[4 * r + 1] = a.y; mm[4 * r + 2] = b.x; mm[4 * r + 3] = b.y;
      }
      mm[12] = 0.0; mm[13] = 0.0; mm[14] = 0.0; mm[15] = 1.0;
    }
    __syncwarp();   // every lane holds its pose: the input buffer is free for the next tile
    if (lane == 0 && next < n_tiles) {
      const uint32_t bytes = (uint32_t)(next == n_tiles - 1 ? last_m : R2IK_TILE) * 128u;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(in_s, poses + (size_t)next * (R2IK_TILE * 16), bytes, bar);
    }
    const int m = tile == n_tiles - 1 ? last_m : R2IK_TILE;
    // ---- solve
    Reach rc;
    double j[7], E[3];
#pragma unroll
    for (int k = 0; k < 7; ++k) j[k] = NAN;
    E[0] = NAN; E[1] = NAN; E[2] = NAN;
    rc.state = R2IK_STATE_INVALID_ROTATION; rc.i0 = NAN; rc.i1 = NAN;
    if (lane < m) {
      Solve S;
      double pos[3] = {mm[3], mm[7], mm[11]};
      if (rotation_from_mat4(mm, false, S.R)) rc = is_reachable_R<false>(A, pos, S);
      if (rc.state == R2IK_STATE_REACHABLE) get_joints_cs(A, S, rc.c0, rc.s0, 0.0, 0.0, j, E);
    }
    const bool ok = rc.state == R2IK_STATE_REACHABLE;
    // ---- results
    if (m == R2IK_TILE) {
      if (lane == 0) bulk_wait_read0();   // the previous tile's bulk stores have read the staging area
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 7; ++k) W.joints[lane * 7 + k] = j[k];
      *reinterpret_cast<double2 *>(W.interval + lane * 2) = make_double2(rc.i0, rc.i1);
      W.elbow[lane * 3] = E[0]; W.elbow[lane * 3 + 1] = E[1]; W.elbow[lane * 3 + 2] = E[2];
      W.reach[lane] = ok ? 1 : 0;
      W.state[lane] = (uint8_t)rc.state;
      fence_async_smem();   // generic-proxy writes -> visible to the bulk-copy engine
      __syncwarp();
      if (lane == 0) {
        const size_t base = (size_t)tile * R2IK_TILE;
        bulk_s2g(joints + base * 7, smem_u32(W.joints), R2IK_TILE * 56);
        bulk_s2g(interval + base * 2, smem_u32(W.interval), R2IK_TILE * 16);
        bulk_s2g(elbow + base * 3, smem_u32(W.elbow), R2IK_TILE * 24);
        bulk_s2g(reachable + base, smem_u32(W.reach), R2IK_TILE);
        bulk_s2g(state + base, smem_u32(W.state), R2IK_TILE);
        bulk_commit();
      }
    } else if (lane < m) {   // ragged last tile: plain stores
      const size_t i = (size_t)tile * R2IK_TILE + lane;
#pragma unroll
      for (int k = 0; k < 7; ++k) joints[7 * i + k] = j[k];
      interval[2 * i] = rc.i0; interval[2 * i + 1] = rc.i1;
      elbow[3 * i] = E[0]; elbow[3 * i + 1] = E[1]; elbow[3 * i + 2] = E[2];
      reachable[i] = ok ? 1 : 0;
      state[i] = (uint8_t)rc.state;
    }
    tile = next;
  }
  if (lane == 0) {
    bulk_wait_read0();   // shared memory must outlive the bulk stores that read it
    const unsigned total = gridDim.x * R2IK_STREAM_WARPS;
    if (atomicAdd(sched + 1, 1u) == total - 1) {   // last warp out: re-arm the slot
      sched[0] = 0u;
      sched[1] = 0u;
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(R2IK_BLOCK)
k_symik_no_limits(const __grid_constant__ ArmConst A, const double *__restrict__ poses, const double *__restrict__ theta,
                  int64_t n, double *__restrict__ joints, double *__restrict__ elbow) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double pos[3], j[7], E[3];
  Solve S;
  bool ok = load_pose<KIND>(poses, i, false, pos, S.R);
  if (ok) ok = is_reachable_R<true>(A, pos, S).state == R2IK_STATE_REACHABLE;
  if (ok) {
    get_joints(A, S, theta[i], 0.0, 0.0, j, E);
  } else {
    for (int k = 0; k < 7; ++k) j[k] = NAN;
    E[0] = NAN; E[1] = NAN; E[2] = NAN;
  }
  for (int k = 0; k < 7; ++k) joints[7 * i + k] = j[k];
  if (elbow) { elbow[3 * i] = E[0]; elbow[3 * i + 1] = E[1]; elbow[3 * i + 2] = E[2]; }
}

template <int KIND>
__global__ void __launch_bounds__(R2IK_BLOCK)
k_elbow_positions(const __grid_constant__ ArmConst A, const double *__restrict__ poses, const double *__restrict__ thetas,
                  int K, int64_t n, double *__restrict__ elbows) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double pos[3];
  Solve S;
  bool ok = load_pose<KIND>(poses, i, false, pos, S.R);
  if (ok) ok = is_reachable_R<false>(A, pos, S).state == R2IK_STATE_REACHABLE;
  for (int k = 0; k < K; ++k) {
    double E[3] = {NAN, NAN, NAN};
    if (ok) elbow_position(S, thetas[(size_t)i * K + k], E);
    double *o = elbows + ((size_t)i * K + k) * 3;
    o[0] = E[0]; o[1] = E[1]; o[2] = E[2];
  }
}

// ---------------------------------------------------------------------------------------
// K2: ControlIK discrete mode.  Each lane solves its own pose (is_reachable, preferred-theta
// shortcut); poses that need the K-sample search are then served one at a time by the whole
// warp: the circle is broadcast with shuffles, lane l evaluates samples l, l+32, ..., and a
// shuffle arg-min with lowest-index tie-break reproduces the reference's strict-< scan
// (utl:381-390).  Each lane finally runs get_joints + safety_checks for its own pose.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__global__ void __launch_bounds__(R2IK_BLOCK)
k_ctl_discrete(const __grid_constant__ ArmConst A, const __grid_constant__ R2ikCtlParams par,
               const double *__restrict__ M, int64_t n, const double *__restrict__ prev_joints,
               const double *__restrict__ current_joints, double *__restrict__ joints, uint8_t *__restrict__ reachable,
               uint8_t *__restrict__ state, uint8_t *__restrict__ emergency) {
  const int lane = threadIdx.x & 31;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < n;
  double prev[7], cur[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) { prev[k] = prev_joints[k]; cur[k] = current_joints[k]; }

  Solve S;
  int st = R2IK_STATE_INVALID_ROTATION;
  bool found = false, need_search = false, valid_pose = false;
  double theta = 0.0, start = 0.0, stop = 0.0;
  if (active) {
    double pos[3];
    valid_pose = load_pose<R2IK_POSE_MAT4>(M, i, true, pos, S.R);
    if (valid_pose) {
      Reach rc = is_reachable_R<false>(A, pos, S);
      st = rc.state;
      if (st == R2IK_STATE_REACHABLE) {
        if (preferred_theta_works(A, S, rc.i0, rc.i1, par.preferred_theta)) {
          theta = par.preferred_theta; found = true;
        } else {
          need_search = true;
          search_range(rc.i0, rc.i1, start, stop);
        }
      }
    }
  }
  const int nb = par.nb_search_points;
  unsigned pending = __ballot_sync(0xffffffffu, need_search);
  while (pending) {
    const int src = __ffs(pending) - 1;
    pending &= pending - 1;
    Solve B;  // only the fields elbow_position reads
    B.c[0] = shfl_d(S.c[0], src); B.c[1] = shfl_d(S.c[1], src); B.c[2] = shfl_d(S.c[2], src);
    B.a1[0] = shfl_d(S.a1[0], src); B.a1[1] = shfl_d(S.a1[1], src); B.a1[2] = shfl_d(S.a1[2], src);
    B.a2[0] = shfl_d(S.a2[0], src); B.a2[1] = shfl_d(S.a2[1], src); B.a2[2] = shfl_d(S.a2[2], src);
    B.r = shfl_d(S.r, src);
    const double b_start = shfl_d(start, src), b_stop = shfl_d(stop, src);
    double best = INFINITY;
    int best_k = 0x7fffffff;
    for (int k = lane; k < nb; k += 32) {
      double th = linspace_at(b_start, b_stop, nb, k);
      double cost = sample_cost(A, B, th, par.preferred_theta);
      if (cost < best) { best = cost; best_k = k; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      double ob = __shfl_xor_sync(0xffffffffu, best, off);
      int ok = __shfl_xor_sync(0xffffffffu, best_k, off);
      if (ob < best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
    }
    if (lane == src) {
      found = best < INFINITY;
      if (found) theta = linspace_at(start, stop, nb, best_k);
      else st = R2IK_STATE_LIMITED_BY_SHOULDER;
    }
  }
  if (!active) return;
  double j[7];
  int bits = 0;
  if (valid_pose) {
    bits = discrete_finish(A, par, S, found, theta, prev, cur, j);
  } else {
#pragma unroll
    for (int k = 0; k < 7; ++k) j[k] = NAN;
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) joints[7 * i + k] = j[k];
  reachable[i] = found ? 1 : 0;
  state[i] = (uint8_t)st;
  if (emergency) emergency[i] = (uint8_t)bits;
}

// ---------------------------------------------------------------------------------------
// K3: ControlIK continuous mode.  A trajectory is a sequential recursion over its waypoints
// (previous_theta / previous_sol / init / emergency latch), so one thread owns one trajectory
// and keeps the controller state in registers; parallelism comes from the trajectories.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(R2IK_BLOCK)
k_ctl_continuous(const __grid_constant__ ArmConst A, const __grid_constant__ R2ikCtlParams par,
                 const double *__restrict__ M, int64_t T, int W, const double *__restrict__ current_joints,
                 const double *__restrict__ current_pose, R2ikTrajState *__restrict__ states,
                 double *__restrict__ joints, uint8_t *__restrict__ reachable, uint8_t *__restrict__ state) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  R2ikTrajState cs = states[t];
  double cj[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) cj[k] = current_joints[7 * t + k];
  for (int w = 0; w < W; ++w) {
    size_t k = (size_t)t * W + w;
    double m[16], cp[16];
    load_mat4(M + 16 * k, m);
    if (!cs.has_previous_sol && !cs.emergency_stop) load_mat4(current_pose + 16 * t, cp);
    double j[7];
    uint8_t r, s;
    continuous_step(A, par, m, cj, cp, cs, j, r, s);
#pragma unroll
    for (int q = 0; q < 7; ++q) joints[7 * k + q] = j[q];
    reachable[k] = r;
    state[k] = s;
  }
  states[t] = cs;
}

// ---------------------------------------------------------------------------------------
// K4: workspace reachability map.  One thread per voxel; the rotation-dependent data of the
// orientation slice (goal rotation matrix) is computed once per block into shared memory.
// Voxels outside the reach sphere or behind the torso plane leave before the orientation loop
// (those two states do not depend on the orientation, sik:284-307).
// ---------------------------------------------------------------------------------------
#define R2IK_ORI_CHUNK 64

__global__ void __launch_bounds__(R2IK_BLOCK)
k_reach_map(const __grid_constant__ ArmConst A, double ox, double oy, double oz, double sx, double sy, double sz,
            int d0, int d1, int d2, const double *__restrict__ ori_euler, int ori_begin, int ori_end,
            uint32_t *__restrict__ counts) {
  __shared__ double sR[R2IK_ORI_CHUNK][9];
  const int64_t nv = (int64_t)d0 * d1 * d2;
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_grid = v < nv;
  double px = 0, py = 0, pz = 0;
  bool live = false;
  if (in_grid) {
    int iz = (int)(v % d2);
    int iy = (int)((v / d2) % d1);
    int ix = (int)(v / ((int64_t)d2 * d1));
    px = ox + ix * sx; py = oy + iy * sy; pz = oz + iz * sz;
    live = reach_prechecks(A, px, py, pz) < 0;
  }
  uint32_t count = 0;
  for (int base = ori_begin; base < ori_end; base += R2IK_ORI_CHUNK) {
    int m = min(R2IK_ORI_CHUNK, ori_end - base);
    __syncthreads();
    for (int o = threadIdx.x; o < m; o += blockDim.x) {
      const double *e = ori_euler + 3 * (size_t)(base + o);
      double R[9];
      rot_from_euler_xyz(e[0], e[1], e[2], R);
#pragma unroll
      for (int k = 0; k < 9; ++k) sR[o][k] = R[k];
    }
    __syncthreads();
    if (live) {
      for (int o = 0; o < m; ++o) {
        Solve S;
        S.p[0] = px; S.p[1] = py; S.p[2] = pz;
#pragma unroll
        for (int k = 0; k < 9; ++k) S.R[k] = sR[o][k];
        Reach rc = solve_core<false, true>(A, S);
        count += (rc.state == R2IK_STATE_REACHABLE) ? 1u : 0u;
      }
    }
  }
  if (in_grid) counts[v] = count;
}

// ---------------------------------------------------------------------------------------
// FK: tip pose of the 7-joint arm chain (synthetic FK-sampled workloads, round-trip checks).
// A = A * [R|t] on 3x4 affine transforms held in registers; joint rotations by Rodrigues.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void affine_mul(double A[12], const double B[12]) {
  double o[12];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double v = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c];
      if (c == 3) v += A[4 * r + 3];
      o[4 * r + c] = v;
    }
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) A[k] = o[k];
}

__global__ void __launch_bounds__(R2IK_BLOCK)
k_fk(const __grid_constant__ R2ikFkChain ch, const double *__restrict__ joints, int64_t n, double *__restrict__ M) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    affine_mul(T, ch.fixed[k]);
    double s, c;
    sincos(joints[7 * i + k], &s, &c);
    const double ax = ch.axis[k][0], ay = ch.axis[k][1], az = ch.axis[k][2], v = 1.0 - c;
    // R = I + sin q K + (1 - cos q) K^2
    double Rj[12] = {1.0 - v * (ay * ay + az * az), -s * az + v * ax * ay, s * ay + v * ax * az, 0.0,
                     s * az + v * ax * ay, 1.0 - v * (ax * ax + az * az), -s * ax + v * ay * az, 0.0,
                     -s * ay + v * ax * az, s * ax + v * ay * az, 1.0 - v * (ax * ax + ay * ay), 0.0};
    affine_mul(T, Rj);
  }
  affine_mul(T, ch.fixed[7]);
  double2 *o = reinterpret_cast<double2 *>(M + 16 * i);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    o[2 * r] = make_double2(T[4 * r], T[4 * r + 1]);
    o[2 * r + 1] = make_double2(T[4 * r + 2], T[4 * r + 3]);
  }
  o[6] = make_double2(0.0, 0.0);
  o[7] = make_double2(0.0, 1.0);
}

// ---------------------------------------------------------------------------------------
// FP64 FMA peak probe (roofline denominator for bench.py; MEASURED_PEAKS.json has no FP64 entry)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_probe(int iters, double seed, double *sink) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) sink[0] = s;  // never true: keeps the chains alive
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
struct r2ik_context {
  int device;
  R2ikArmConfig cfg;
  ArmConst A;
  R2ikArmConstants pub;
  int stream_blocks;   // persistent grid of k_symik_solve_stream: SMs x resident blocks (0: kernel unavailable)
  int force_generic;   // R2IK_K1_GENERIC=1: always use the one-thread-per-pose kernel (A/B timing)
  // tile-scheduler slots of k_symik_solve_stream, one per CUDA stream that has launched it (launches on
  // one stream are ordered, so a slot is never shared by two running kernels); 2 x u32 each, self-resetting
  std::mutex sched_mutex;
  std::vector<std::pair<cudaStream_t, unsigned *>> sched_slots;
};

static unsigned *sched_slot_for(r2ik_context *h, cudaStream_t s) {
  std::lock_guard<std::mutex> lock(h->sched_mutex);
  for (auto &e : h->sched_slots)
    if (e.first == s) return e.second;
  unsigned *p = nullptr;
  if (cudaMalloc(&p, 2 * sizeof(unsigned)) != cudaSuccess) return nullptr;
  if (cudaMemset(p, 0, 2 * sizeof(unsigned)) != cudaSuccess) { cudaFree(p); return nullptr; }
  h->sched_slots.emplace_back(s, p);
  return p;
}

static thread_local char g_err[256] = "";

static int fail_arg(int code, const char *msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return code;
}
static int fail_cuda(cudaError_t e, const char *where) {
  snprintf(g_err, sizeof g_err, "%s: %s", where, cudaGetErrorString(e));
  return -(int)e;
}
#define R2IK_CUDA(call, where)                        \
  do {                                                \
    cudaError_t e_ = (call);                          \
    if (e_ != cudaSuccess) return fail_cuda(e_, where); \
  } while (0)

static inline bool misaligned16(const void *p) { return ((uintptr_t)p & 15) != 0; }
static inline unsigned blocks_for(int64_t n) { return (unsigned)((n + R2IK_BLOCK - 1) / R2IK_BLOCK); }

extern "C" {

int r2ik_abi_version(void) { return R2IK_ABI_VERSION; }
const char *r2ik_last_error(void) { return g_err; }

int r2ik_create(const R2ikArmConfig *cfg, int device, r2ik_handle *out) {
  if (!cfg || !out) return fail_arg(R2IK_ERR_NULL, "r2ik_create: null argument");
  if (cfg->side != 1 && cfg->side != -1) return fail_arg(R2IK_ERR_ARM, "r2ik_create: side must be +1 (r_arm) or -1 (l_arm)");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount");
  if (count == 0 || device < 0 || device >= count) return fail_arg(R2IK_ERR_NO_DEVICE, "r2ik_create: no such CUDA device");
  r2ik_context *h = new (std::nothrow) r2ik_context;
  if (!h) return fail_arg(R2IK_ERR_ARG, "r2ik_create: out of host memory");
  h->device = device;
  h->cfg = *cfg;
  derive_constants(*cfg, h->A, h->pub);
  {
    // size the persistent grid of the streaming K1 from the occupancy the driver reports
    h->stream_blocks = 0;
    const char *g = getenv("R2IK_K1_GENERIC");
    h->force_generic = (g && g[0] == '1') ? 1 : 0;
    int sms = 0, per_sm = 0;
    if (cudaSetDevice(device) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_symik_solve_stream, R2IK_BLOCK, 0) == cudaSuccess)
      h->stream_blocks = sms * per_sm;
    (void)cudaGetLastError();
  }
  *out = h;
  return 0;
}

int r2ik_destroy(r2ik_handle h) {
  if (h) {
    if (!h->sched_slots.empty() && cudaSetDevice(h->device) == cudaSuccess)
      for (auto &e : h->sched_slots) cudaFree(e.second);
    (void)cudaGetLastError();
  }
  delete h;
  return 0;
}

int r2ik_get_constants(r2ik_handle h, R2ikArmConstants *out) {
  if (!h || !out) return fail_arg(R2IK_ERR_NULL, "r2ik_get_constants: null argument");
  *out = h->pub;
  return 0;
}

int r2ik_interval_limit(int side, int low_elbow, double *out) {
  if (!out) return fail_arg(R2IK_ERR_NULL, "r2ik_interval_limit: null argument");
  // ctl:225-252
  double il0 = low_elbow ? -4 * kPi / 5 : 3 * kPi / 4;
  double il1 = low_elbow ? 0.0 : -2 * kPi / 6;
  if (side < 0) {
    double a = -kPi - il1, b = -kPi - il0;
    il0 = a; il1 = b;
    if (il0 < -kPi) il0 = pymod(il0, kTwoPi);
    if (il1 < -kPi) il1 = pymod(il1, kTwoPi);
    if (il0 > kPi) il0 = pymod(il0, -kTwoPi);
    if (il1 > kPi) il1 = pymod(il1, -kTwoPi);
  }
  out[0] = il0; out[1] = il1;
  return 0;
}

int r2ik_symik_solve_f64(r2ik_handle h, int pose_kind, const double *poses, const double *theta, const double *prev_joints,
                         int64_t n, uint8_t *reachable, uint8_t *state, double *interval, double *joints, double *elbow,
                         void *stream) {
  if (n < 0 || (pose_kind != R2IK_POSE_EULER6 && pose_kind != R2IK_POSE_MAT4))
    return fail_arg(R2IK_ERR_ARG, "r2ik_symik_solve_f64: bad n or pose_kind");
  if (!h) return fail_arg(R2IK_ERR_NULL, "r2ik_symik_solve_f64: null handle");
  if (n == 0) return 0;  // empty batch: nothing to read or write
  if (!poses || !reachable || !state) return fail_arg(R2IK_ERR_NULL, "r2ik_symik_solve_f64: null argument");
  if (misaligned16(poses)) return fail_arg(R2IK_ERR_ARG, "r2ik_symik_solve_f64: poses must be 16-byte aligned");
  R2IK_CUDA(cudaSetDevice(h->device), "cudaSetDevice");
  cudaStream_t s = (cudaStream_t)stream;
  // Streaming form: the common full-output call on 16-byte aligned buffers (TMA bulk copies).
  const uintptr_t align_or = (uintptr_t)reachable | (uintptr_t)state | (uintptr_t)interval | (uintptr_t)joints | (uintptr_t)elbow;
  if (pose_kind == R2IK_POSE_MAT4 && !theta && !prev_joints && interval && joints && elbow && (align_or & 15) == 0 &&
      h->stream_blocks > 0 && !h->force_generic) {
    const int64_t n_tiles = (n + R2IK_TILE - 1) / R2IK_TILE;
    unsigned *sched = n_tiles < (int64_t)1 << 30 ? sched_slot_for(h, s) : nullptr;
    if (sched) {
      const int64_t want = (n_tiles + R2IK_STREAM_WARPS - 1) / R2IK_STREAM_WARPS;
      const unsigned blocks = (unsigned)(want < (int64_t)h->stream_blocks ? want : (int64_t)h->stream_blocks);
      k_symik_solve_stream<<<blocks, R2IK_BLOCK, 0, s>>>(h->A, poses, n, (int)n_tiles, sched, reachable, state, interval, joints, elbow);
      R2IK_CUDA(cudaGetLastError(), "k_symik_solve_stream launch");
      return 0;
    }
  }
  if (pose_kind == R2IK_POSE_MAT4)
    k_symik_solve<R2IK_POSE_MAT4><<<blocks_for(n), R2IK_BLOCK, 0, s>>>(h->A, poses, theta, prev_joints, n, reachable, state, interval, joints, elbow);
  else
    k_symik_solve<R2IK_POSE_EULER6><<<blocks_for(n), R2IK_BLOCK, 0, s>>>(h->A, poses, theta, prev_joints, n, reachable, state, interval, joints, elbow);
  R2IK_CUDA(cudaGetLastError(), "k_symik_solve launch");
  return 0;
}

int r2ik_symik_no_limits_f64(r2ik_handle h, int pose_kind, const double *poses, const double *theta, int64_t n,
                             double *joints, double *elbow, void *stream) {
  if (n < 0 || (pose_kind != R2IK_POSE_EULER6 && pose_kind != R2IK_POSE_MAT4))
    return fail_arg(R2IK_ERR_ARG, "r2ik_symik_no_limits_f64: bad n or pose_kind");
  if (!h) return fail_arg(R2IK_ERR_NULL, "r2ik_symik_no_limits_f64: null handle");
  if (n == 0) return 0;
  if (!poses || !theta || !joints) return fail_arg(R2IK_ERR_NULL, "r2ik_symik_no_limits_f64: null argument");
  if (misaligned16(poses)) return fail_arg(R2IK_ERR_ARG, "r2ik_symik_no_limits_f64: poses must be 16-byte aligned");
  R2IK_CUDA(cudaSetDevice(h->device), "cudaSetDevice");
  cudaStream_t s = (cudaStream_t)stream;
  if (pose_kind == R2IK_POSE_MAT4)
    k_symik_no_limits<R2IK_POSE_MAT4><<<blocks_for(n), R2IK_BLOCK, 0, s>>>(h->A, poses, theta, n, joints, elbow);
  else
    k_symik_no_limits<R2IK_POSE_EULER6><<<blocks_for(n), R2IK_BLOCK, 0, s>>>(h->A, poses, theta, n, joints, elbow);
  R2IK_CUDA(cudaGetLastError(), "k_symik_no_limits launch");
  return 0;
}

int r2ik_elbow_positions_f64(r2ik_handle h, int pose_kind, const double *poses, const double *thetas, int32_t K, int64_t n,
                             double *elbows, void *stream) {
  if (n < 0 || K < 0 || (pose_kind != R2IK_POSE_EULER6 && pose_kind != R2IK_POSE_MAT4))
    return fail_arg(R2IK_ERR_ARG, "r2ik_elbow_positions_f64: bad n, K or pose_kind");
  if (!h) return fail_arg(R2IK_ERR_NULL, "r2ik_elbow_positions_f64: null handle");
  if (n == 0 || K == 0) return 0;
  if (!poses || !thetas || !elbows) return fail_arg(R2IK_ERR_NULL, "r2ik_elbow_positions_f64: null argument");
  if (misaligned16(poses)) return fail_arg(R2IK_ERR_ARG, "r2ik_elbow_positions_f64: poses must be 16-byte aligned");
  R2IK_CUDA(cudaSetDevice(h->device), "cudaSetDevice");
  cudaStream_t s = (cudaStream_t)stream;
  if (pose_kind == R2IK_POSE_MAT4)
    k_elbow_positions<R2IK_POSE_MAT4><<<blocks_for(n), R2IK_BLOCK, 0, s>>>(h->A, poses, thetas, K, n, elbows);
  else
    k_elbow_positions<R2IK_POSE_EULER6><<<blocks_for(n), R2IK_BLOCK, 0, s>>>(h->A, poses, thetas, K, n, elbows);
  R2IK_CUDA(cudaGetLastError(), "k_elbow_positions launch");
  return 0;
}

int r2ik_ctl_discrete_f64(r2ik_handle h, const R2ikCtlParams *par, const double *M, int64_t n, const double *prev_joints,
                          const double *current_joints, double *joints, uint8_t *reachable, uint8_t *state,
                          uint8_t *emergency, void *stream) {
  if (!h || !par) return fail_arg(R2IK_ERR_NULL, "r2ik_ctl_discrete_f64: null handle or parameters");
  if (n < 0 || par->nb_search_points < 2) return fail_arg(R2IK_ERR_ARG, "r2ik_ctl_discrete_f64: bad n or nb_search_points");
  if (n == 0) return 0;
  if (!M || !prev_joints || !current_joints || !joints || !reachable || !state)
    return fail_arg(R2IK_ERR_NULL, "r2ik_ctl_discrete_f64: null argument");
  if (misaligned16(M)) return fail_arg(R2IK_ERR_ARG, "r2ik_ctl_discrete_f64: M must be 16-byte aligned");
  R2IK_CUDA(cudaSetDevice(h->device), "cudaSetDevice");
  k_ctl_discrete<<<blocks_for(n), R2IK_BLOCK, 0, (cudaStream_t)stream>>>(h->A, *par, M, n, prev_joints, current_joints, joints,
                                                                        reachable, state, emergency);
  R2IK_CUDA(cudaGetLastError(), "k_ctl_discrete launch");
  return 0;
}

int r2ik_ctl_continuous_f64(r2ik_handle h, const R2ikCtlParams *par, const double *M, int64_t T, int32_t W,
                            const double *current_joints, const double *current_pose, R2ikTrajState *st, double *joints,
                            uint8_t *reachable, uint8_t *state, void *stream) {
  if (!h || !par) return fail_arg(R2IK_ERR_NULL, "r2ik_ctl_continuous_f64: null handle or parameters");
  if (T < 0 || W < 0 || par->nb_search_points_continuous < 2)
    return fail_arg(R2IK_ERR_ARG, "r2ik_ctl_continuous_f64: bad T, W or nb_search_points_continuous");
  if (T == 0 || W == 0) return 0;
  if (!M || !current_joints || !current_pose || !st || !joints || !reachable || !state)
    return fail_arg(R2IK_ERR_NULL, "r2ik_ctl_continuous_f64: null argument");
  if (misaligned16(M) || misaligned16(current_pose))
    return fail_arg(R2IK_ERR_ARG, "r2ik_ctl_continuous_f64: M and current_pose must be 16-byte aligned");
  R2IK_CUDA(cudaSetDevice(h->device), "cudaSetDevice");
  k_ctl_continuous<<<blocks_for(T), R2IK_BLOCK, 0, (cudaStream_t)stream>>>(h->A, *par, M, T, W, current_joints, current_pose, st,
                                                                          joints, reachable, state);
  R2IK_CUDA(cudaGetLastError(), "k_ctl_continuous launch");
  return 0;
}

int r2ik_reach_map_u32(r2ik_handle h, const double *origin, const double *step, const int32_t *dims,
                       const double *orientations_euler, int32_t ori_begin, int32_t ori_end, uint32_t *counts, void *stream) {
  if (!h || !origin || !step || !dims || !orientations_euler || !counts)
    return fail_arg(R2IK_ERR_NULL, "r2ik_reach_map_u32: null argument");
  if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || ori_begin < 0 || ori_end < ori_begin)
    return fail_arg(R2IK_ERR_ARG, "r2ik_reach_map_u32: bad dims or orientation range");
  int64_t nv = (int64_t)dims[0] * dims[1] * dims[2];
  R2IK_CUDA(cudaSetDevice(h->device), "cudaSetDevice");
  k_reach_map<<<blocks_for(nv), R2IK_BLOCK, 0, (cudaStream_t)stream>>>(h->A, origin[0], origin[1], origin[2], step[0], step[1],
                                                                      step[2], dims[0], dims[1], dims[2], orientations_euler,
                                                                      ori_begin, ori_end, counts);
  R2IK_CUDA(cudaGetLastError(), "k_reach_map launch");
  return 0;
}

int r2ik_fk_f64(const R2ikFkChain *chain, int device, const double *joints, int64_t n, double *M, void *stream) {
  if (!chain) return fail_arg(R2IK_ERR_NULL, "r2ik_fk_f64: null chain");
  if (n < 0) return fail_arg(R2IK_ERR_ARG, "r2ik_fk_f64: bad n");
  if (n == 0) return 0;
  if (!joints || !M) return fail_arg(R2IK_ERR_NULL, "r2ik_fk_f64: null argument");
  R2IK_CUDA(cudaSetDevice(device), "cudaSetDevice");
  k_fk<<<blocks_for(n), R2IK_BLOCK, 0, (cudaStream_t)stream>>>(*chain, joints, n, M);
  R2IK_CUDA(cudaGetLastError(), "k_fk launch");
  return 0;
}

int r2ik_dfma_probe(int device, int32_t iters, double *out_ms, double *out_flop, void *stream) {
  if (!out_ms || !out_flop) return fail_arg(R2IK_ERR_NULL, "r2ik_dfma_probe: null argument");
  R2IK_CUDA(cudaSetDevice(device), "cudaSetDevice");
  cudaDeviceProp prop;
  R2IK_CUDA(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
  cudaStream_t s = (cudaStream_t)stream;
  double *sink = nullptr;
  R2IK_CUDA(cudaMalloc(&sink, sizeof(double)), "cudaMalloc");
  const int threads = 256, blocks = prop.multiProcessorCount * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma_probe<<<blocks, threads, 0, s>>>(iters / 8 + 1, 1.0, sink);  // warm-up
  cudaEventRecord(e0, s);
  k_dfma_probe<<<blocks, threads, 0, s>>>(iters, 1.0, sink);
  cudaEventRecord(e1, s);
  cudaError_t e = cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(sink);
  if (e != cudaSuccess) return fail_cuda(e, "k_dfma_probe");
  *out_ms = (double)ms;
  *out_flop = 2.0 * 8.0 * (double)iters * (double)threads * (double)blocks;
  return 0;
}

}  // extern "C"
