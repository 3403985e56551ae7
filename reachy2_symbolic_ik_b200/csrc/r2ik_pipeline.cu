// r2ik_pipeline.cu -- the host-buffer pipelines of libr2ik.so (include/r2ik.h, "host pipelines").
//
// A batch that lives in HOST memory is cut into chunks that flow  H2D copy -> kernel -> D2H copy  on three CUDA streams
// chained by events, so that both PCIe directions and the kernel overlap.  The kernels are 10-40x faster than the link
// (K1: 1.5e10 poses/s against 4e8-9e8 poses/s of PCIe 5 x16), so what matters here is that the two copy engines never
// wait for the host: the chunk loop is native code (a handful of runtime calls per chunk, ~20 us) instead of an
// interpreter loop (measured: 74 % of the plain-copy bound from Python, profiles/r2_experiments.md).
//
// All device staging buffers, streams and events belong to the pipeline object (created once, reused by every call);
// the launch entries themselves still allocate nothing.  Calls enqueue and return; r2ik_pipeline_wait blocks until the
// results are in host memory.  Host buffers should be pinned (cudaHostAlloc / torch pin_memory) -- pageable memory
// works but makes the copies synchronous.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/r2ik.h"

namespace {

thread_local char g_perr[256] = "";

int pfail(int code, const char *msg) {
  snprintf(g_perr, sizeof g_perr, "%s", msg);
  return code;
}
int pfail_cuda(cudaError_t e, const char *where) {
  snprintf(g_perr, sizeof g_perr, "%s: %s", where, cudaGetErrorString(e));
  return -(int)e;
}
#define P_CUDA(call, where)                              \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return pfail_cuda(e_, where); \
  } while (0)

// Every entry runs on its handle's device and leaves the calling thread's current device as it found it (the caller --
// torch, or any other runtime user in the process -- keeps its own notion of the current device).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
    else if (err == cudaSuccess) prev = -1;     // nothing to restore
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct Slot {
  void *poses = nullptr;      // chunk x 16 doubles (also holds float poses and n x 6 layouts)
  uint8_t *reach = nullptr, *state = nullptr, *aux = nullptr;   // aux: emergency bits (discrete mode)
  void *interval = nullptr, *joints = nullptr, *elbow = nullptr;
  uint32_t *esc = nullptr, *n_esc = nullptr;                     // FP32 path scratch
  cudaEvent_t h2d_done = nullptr, k_done = nullptr, d2h_done[3] = {nullptr, nullptr, nullptr};
  bool used = false;
};

// Is `p` pinned host memory the device can address (cudaHostAlloc / cudaHostRegister under unified addressing)?  Then a
// kernel can write the one-byte outputs (state, reachable, emergency) straight into it: a D2H copy of a few hundred KB
// costs ~25 us of fixed latency on the copy stream (scripts/experiments/exp_r2_copy_patterns.py), more than its payload.
bool device_can_write(const void *p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost && a.devicePointer != nullptr;
}

}  // namespace

struct r2ik_pipeline {
  r2ik_handle h = nullptr;
  int device = 0;
  int64_t chunk = 0;
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out[3] = {nullptr, nullptr, nullptr};   // D2H: joints | interval | elbow + bytes
  std::vector<Slot> slots;
  double *prev = nullptr, *cur = nullptr;   // discrete mode: previous_sol / current_joints (7 doubles each)
  int64_t next_chunk = 0;                   // slots are used round-robin across calls
};

extern "C" {

const char *r2ik_pipeline_last_error(void) { return g_perr; }

int r2ik_pipeline_destroy(r2ik_pipeline *p) {
  if (!p) return 0;
  DeviceGuard guard_(p->device);
  for (cudaStream_t q : p->s_out)
    if (q) cudaStreamSynchronize(q);
  if (p->s_k) cudaStreamSynchronize(p->s_k);
  for (Slot &s : p->slots) {
    cudaFree(s.poses); cudaFree(s.reach); cudaFree(s.state); cudaFree(s.aux); cudaFree(s.interval); cudaFree(s.joints);
    cudaFree(s.elbow); cudaFree(s.esc); cudaFree(s.n_esc);
    if (s.h2d_done) cudaEventDestroy(s.h2d_done);
    if (s.k_done) cudaEventDestroy(s.k_done);
    for (cudaEvent_t ev : s.d2h_done)
      if (ev) cudaEventDestroy(ev);
  }
  cudaFree(p->prev); cudaFree(p->cur);
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_k) cudaStreamDestroy(p->s_k);
  for (cudaStream_t q : p->s_out)
    if (q) cudaStreamDestroy(q);
  delete p;
  return 0;
}

int r2ik_pipeline_create(r2ik_handle h, int device, int64_t chunk_poses, int32_t n_slots, r2ik_pipeline **out) {
  if (!h || !out) return pfail(R2IK_ERR_NULL, "r2ik_pipeline_create: null argument");
  if (chunk_poses <= 0 || chunk_poses > (1LL << 28) || n_slots < 2 || n_slots > 16)
    return pfail(R2IK_ERR_ARG, "r2ik_pipeline_create: chunk_poses must be 1 .. 2^28 and n_slots 2 .. 16");
  r2ik_pipeline *p = new (std::nothrow) r2ik_pipeline;
  if (!p) return pfail(R2IK_ERR_ARG, "r2ik_pipeline_create: out of host memory");
  p->h = h; p->device = device; p->chunk = chunk_poses;
  DeviceGuard guard_(device);
  cudaError_t e = guard_.err;
  auto check = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  check(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
  check(cudaStreamCreateWithFlags(&p->s_k, cudaStreamNonBlocking));
  for (cudaStream_t &q : p->s_out) check(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
  check(cudaMalloc(&p->prev, 7 * sizeof(double)));
  check(cudaMalloc(&p->cur, 7 * sizeof(double)));
  p->slots.resize((size_t)n_slots);
  const size_t c = (size_t)chunk_poses;
  for (Slot &s : p->slots) {
    check(cudaMalloc(&s.poses, c * 16 * sizeof(double)));
    check(cudaMalloc(&s.reach, c)); check(cudaMalloc(&s.state, c)); check(cudaMalloc(&s.aux, c));
    check(cudaMalloc(&s.interval, c * 2 * sizeof(double)));
    check(cudaMalloc(&s.joints, c * 7 * sizeof(double)));
    check(cudaMalloc(&s.elbow, c * 3 * sizeof(double)));
    check(cudaMalloc(&s.esc, c * sizeof(uint32_t))); check(cudaMalloc(&s.n_esc, sizeof(uint32_t)));
    check(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
    check(cudaEventCreateWithFlags(&s.k_done, cudaEventDisableTiming));
    for (cudaEvent_t &ev : s.d2h_done) check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }
  if (e != cudaSuccess) {
    r2ik_pipeline_destroy(p);
    return pfail_cuda(e, "r2ik_pipeline_create");
  }
  *out = p;
  return 0;
}

int r2ik_pipeline_wait(r2ik_pipeline *p) {
  if (!p) return pfail(R2IK_ERR_NULL, "r2ik_pipeline_wait: null pipeline");
  DeviceGuard guard_(p->device);
  P_CUDA(guard_.err, "cudaSetDevice");
  for (cudaStream_t q : p->s_out) P_CUDA(cudaStreamSynchronize(q), "cudaStreamSynchronize");
  P_CUDA(cudaStreamSynchronize(p->s_k), "cudaStreamSynchronize");     // the kernels write the byte outputs themselves
  return 0;
}

}  // extern "C"

namespace {

enum Mode { SYMIK_F64, SYMIK_F32, DISCRETE_F64 };

// One chunk through the three streams.  `esz` = bytes per pose element (8 / 4), `k` = elements per pose (16 / 6).
int run_chunks(r2ik_pipeline *p, Mode mode, int pose_kind, const void *poses_host, int64_t n, const R2ikCtlParams *par,
               uint8_t *reachable, uint8_t *state, void *interval, void *joints, void *elbow, uint8_t *aux) {
  const size_t esz = mode == SYMIK_F32 ? 4 : 8;
  const size_t k = (mode == DISCRETE_F64 || pose_kind == R2IK_POSE_MAT4) ? 16 : (pose_kind == R2IK_POSE_MAT34 ? 12 : 6);
  const char *src = static_cast<const char *>(poses_host);
  DeviceGuard guard_(p->device);
  P_CUDA(guard_.err, "cudaSetDevice");
  const bool direct_state = device_can_write(state), direct_reach = device_can_write(reachable), direct_aux = device_can_write(aux);
  // Chunk sizes.  The link is the bottleneck and every copy costs ~10-25 us of fixed latency on its stream
  // (scripts/experiments/exp_r2_copy_patterns.py, exp_r2_pipe_timeline.py), so the bulk moves in full-size chunks.  What
  // nothing overlaps is the first H2D copy and the last D2H copy.  When the inbound bytes dominate (4x4 matrices in), the
  // H2D stream is the critical one and the tail is cut into halving chunks so that the last, exposed D2H copy is short;
  // when the outbound bytes dominate (goal poses in, joints out) the D2H stream is critical and ends the call anyway --
  // a ramp at the front would have to grow by less than the ratio of the two copy rates (~1.3) to keep it fed, which
  // costs more in per-copy latency than it saves (measured, profiles/r2_experiments.md).
  std::vector<int64_t> sizes;
  {
    const size_t bytes_in = k * esz;
    const size_t bytes_out = (interval ? 2 * esz : 0) + (joints ? 7 * esz : 0) + (elbow ? 3 * esz : 0) + 1 + (reachable ? 1 : 0) + (aux ? 1 : 0);
    std::vector<int64_t> tail;
    int64_t left = n;
    if (bytes_in >= bytes_out)
      for (int64_t c = p->chunk / 8; c >= 2048 && c < p->chunk && left > 4 * p->chunk; c *= 2) { tail.push_back(c); left -= c; }
    for (; left > 0; left -= p->chunk) sizes.push_back(left < p->chunk ? left : p->chunk);
    sizes.insert(sizes.end(), tail.rbegin(), tail.rend());
  }
  int64_t lo = 0;
  for (size_t ci = 0; ci < sizes.size(); lo += sizes[ci], ++ci) {
    const int64_t m = sizes[ci];
    Slot &s = p->slots[(size_t)(p->next_chunk++ % (int64_t)p->slots.size())];
    // the kernel that read this slot's poses has finished
    if (s.used) P_CUDA(cudaStreamWaitEvent(p->s_in, s.k_done, 0), "cudaStreamWaitEvent");
    P_CUDA(cudaMemcpyAsync(s.poses, src + (size_t)lo * k * esz, (size_t)m * k * esz, cudaMemcpyHostToDevice, p->s_in), "H2D poses");
    P_CUDA(cudaEventRecord(s.h2d_done, p->s_in), "cudaEventRecord");
    P_CUDA(cudaStreamWaitEvent(p->s_k, s.h2d_done, 0), "cudaStreamWaitEvent");
    // this slot's previous results have left
    if (s.used)
      for (cudaEvent_t ev : s.d2h_done) P_CUDA(cudaStreamWaitEvent(p->s_k, ev, 0), "cudaStreamWaitEvent");
    const size_t o = (size_t)lo, mm = (size_t)m;
    // one-byte outputs: written by the kernel straight into the caller's pinned buffers when the device can address them
    uint8_t *k_state = direct_state ? state + o : s.state;
    uint8_t *k_reach = !reachable ? nullptr : (direct_reach ? reachable + o : s.reach);
    uint8_t *k_aux = !aux ? s.aux : (direct_aux ? aux + o : s.aux);
    int rc = 0;
    if (mode == SYMIK_F64)
      rc = r2ik_symik_solve_f64(p->h, pose_kind, static_cast<const double *>(s.poses), nullptr, nullptr, m, k_reach, k_state,
                                interval ? static_cast<double *>(s.interval) : nullptr,
                                joints ? static_cast<double *>(s.joints) : nullptr, elbow ? static_cast<double *>(s.elbow) : nullptr, p->s_k);
    else if (mode == SYMIK_F32)
      rc = r2ik_symik_solve_f32(p->h, pose_kind, static_cast<const float *>(s.poses), nullptr, nullptr, m, k_reach, k_state,
                                interval ? static_cast<float *>(s.interval) : nullptr, joints ? static_cast<float *>(s.joints) : nullptr,
                                elbow ? static_cast<float *>(s.elbow) : nullptr, s.esc, s.n_esc, p->s_k);
    else
      rc = r2ik_ctl_discrete_f64(p->h, par, static_cast<const double *>(s.poses), m, p->prev, p->cur, static_cast<double *>(s.joints),
                                 k_reach, k_state, k_aux, p->s_k);
    if (rc != 0) {
      snprintf(g_perr, sizeof g_perr, "pipeline launch: %s", r2ik_last_error());
      return rc;
    }
    P_CUDA(cudaEventRecord(s.k_done, p->s_k), "cudaEventRecord");
    for (cudaStream_t q : p->s_out) P_CUDA(cudaStreamWaitEvent(q, s.k_done, 0), "cudaStreamWaitEvent");
    // the three float outputs leave on three streams: their fixed per-copy latencies overlap instead of adding up
    if (joints) P_CUDA(cudaMemcpyAsync(static_cast<char *>(joints) + o * 7 * esz, s.joints, mm * 7 * esz, cudaMemcpyDeviceToHost, p->s_out[0]), "D2H joints");
    if (interval) P_CUDA(cudaMemcpyAsync(static_cast<char *>(interval) + o * 2 * esz, s.interval, mm * 2 * esz, cudaMemcpyDeviceToHost, p->s_out[1]), "D2H interval");
    if (elbow) P_CUDA(cudaMemcpyAsync(static_cast<char *>(elbow) + o * 3 * esz, s.elbow, mm * 3 * esz, cudaMemcpyDeviceToHost, p->s_out[2]), "D2H elbow");
    if (!direct_state) P_CUDA(cudaMemcpyAsync(state + o, s.state, mm, cudaMemcpyDeviceToHost, p->s_out[2]), "D2H state");
    if (reachable && !direct_reach) P_CUDA(cudaMemcpyAsync(reachable + o, s.reach, mm, cudaMemcpyDeviceToHost, p->s_out[1]), "D2H reachable");
    if (aux && !direct_aux) P_CUDA(cudaMemcpyAsync(aux + o, s.aux, mm, cudaMemcpyDeviceToHost, p->s_out[1]), "D2H emergency");
    for (int q = 0; q < 3; ++q) P_CUDA(cudaEventRecord(s.d2h_done[q], p->s_out[q]), "cudaEventRecord");
    s.used = true;
  }
  return 0;
}

}  // namespace

extern "C" {

int r2ik_pipeline_symik_f64(r2ik_pipeline *p, int pose_kind, const double *poses_host, int64_t n, uint8_t *reachable,
                            uint8_t *state, double *interval, double *joints, double *elbow) {
  if (!p) return pfail(R2IK_ERR_NULL, "r2ik_pipeline_symik_f64: null pipeline");
  if (n < 0 || (pose_kind != R2IK_POSE_EULER6 && pose_kind != R2IK_POSE_MAT4 && pose_kind != R2IK_POSE_MAT34))
    return pfail(R2IK_ERR_ARG, "r2ik_pipeline_symik_f64: bad n or pose_kind");
  if (n == 0) return 0;
  if (!poses_host || !state) return pfail(R2IK_ERR_NULL, "r2ik_pipeline_symik_f64: null argument");
  return run_chunks(p, SYMIK_F64, pose_kind, poses_host, n, nullptr, reachable, state, interval, joints, elbow, nullptr);
}

int r2ik_pipeline_symik_f32(r2ik_pipeline *p, int pose_kind, const float *poses_host, int64_t n, uint8_t *reachable,
                            uint8_t *state, float *interval, float *joints, float *elbow) {
  if (!p) return pfail(R2IK_ERR_NULL, "r2ik_pipeline_symik_f32: null pipeline");
  if (n < 0 || (pose_kind != R2IK_POSE_EULER6 && pose_kind != R2IK_POSE_MAT4)) return pfail(R2IK_ERR_ARG, "r2ik_pipeline_symik_f32: bad n or pose_kind");
  if (n == 0) return 0;
  if (!poses_host || !state || !reachable) return pfail(R2IK_ERR_NULL, "r2ik_pipeline_symik_f32: null argument (the FP32 kernel always writes `reachable`)");
  return run_chunks(p, SYMIK_F32, pose_kind, poses_host, n, nullptr, reachable, state, interval, joints, elbow, nullptr);
}

int r2ik_pipeline_ctl_discrete_f64(r2ik_pipeline *p, const R2ikCtlParams *par, const double *M_host, int64_t n,
                                   const double *prev_joints_host, const double *current_joints_host, double *joints,
                                   uint8_t *reachable, uint8_t *state, uint8_t *emergency) {
  if (!p || !par) return pfail(R2IK_ERR_NULL, "r2ik_pipeline_ctl_discrete_f64: null pipeline or parameters");
  if (n < 0) return pfail(R2IK_ERR_ARG, "r2ik_pipeline_ctl_discrete_f64: bad n");
  if (n == 0) return 0;
  if (!M_host || !prev_joints_host || !current_joints_host || !joints || !reachable || !state)
    return pfail(R2IK_ERR_NULL, "r2ik_pipeline_ctl_discrete_f64: null argument");
  DeviceGuard guard_(p->device);
  P_CUDA(guard_.err, "cudaSetDevice");
  // the two 7-vectors ride the kernel stream, ordered before this call's first launch (and after the previous call's last)
  P_CUDA(cudaMemcpyAsync(p->prev, prev_joints_host, 7 * sizeof(double), cudaMemcpyHostToDevice, p->s_k), "H2D previous_sol");
  P_CUDA(cudaMemcpyAsync(p->cur, current_joints_host, 7 * sizeof(double), cudaMemcpyHostToDevice, p->s_k), "H2D current_joints");
  return run_chunks(p, DISCRETE_F64, R2IK_POSE_MAT4, M_host, n, par, reachable, state, nullptr, joints, nullptr, emergency);
}

}  // extern "C"
