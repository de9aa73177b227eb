// mfg_step.cu - reset / step / random-action kernels: one thread advances one environment (mfg_core.cuh).
//
//   k_reset           Factory.reset        environment/rules.py:163-199 (SpawnAgents, SpawnEntity), Philox draws
//   k_step            Gamestate.tick + check_done + reward fold   utils/states.py:170-226, factory.py:222-259
//   k_random_actions  action_space.sample() of random_testrun.py:44-56
//
// State is struct-of-arrays with the env index fastest, so every per-field access of a warp is one coalesced
// transaction group; the path is integer / byte work bounded by HBM latency and bandwidth, no tensor cores.
#include "mfg_internal.hpp"

using namespace mfg;

template <int AMAX>
__global__ void __launch_bounds__(128) k_reset(const MfgSpec* __restrict__ sp, Tables tb, State st, const uint8_t* mask) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= st.N) return;
  if (mask && !mask[e]) return;
  uint32_t episode = mask ? st.episode[e] + 1 : 0;
  env_reset<AMAX>(*sp, tb, st, e, episode);
}

// k_step: CTA = 128 envs.  Every thread first copies the integer / byte fields of ITS env from the field-major global
// buffer into shared memory (row after row: fully coalesced, all loads independent => deep memory-level parallelism),
// runs the whole step against that copy (tile look-ups, slot scans and rule hooks hit shared memory instead of
// dependent HBM round trips), then writes the fields back.  f64 fields (dirt amounts, battery, returns) stay in HBM:
// they are touched by few actions / rules only.  No block-level synchronisation is needed: a thread only ever
// touches its own column.
constexpr int STEP_ENVS = 128;

template <int AMAX>
__global__ void __launch_bounds__(STEP_ENVS) k_step(const MfgSpec* __restrict__ spp, Tables tb, State st, StepIO io) {
  extern __shared__ __align__(16) unsigned char stage[];
  const MfgSpec& sp = *spp;
  const int el = threadIdx.x;
  const int64_t eg = (int64_t)blockIdx.x * STEP_ENVS + el;
  const bool live = eg < st.N;
  const size_t Ng = (size_t)st.N;
  State ss = st;
  ss.N = STEP_ENVS;
  size_t off = 0;
#define F(type, name, rows_expr)                                                                   \
  if constexpr (!std::is_same<type, double>::value) {                                              \
    const int rows = (int)(rows_expr);                                                             \
    type* s_ = reinterpret_cast<type*>(stage + off);                                               \
    ss.name = s_;                                                                                  \
    if (live) for (int r = 0; r < rows; ++r) s_[r * STEP_ENVS + el] = st.name[(size_t)r * Ng + eg]; \
    off += ((size_t)rows * STEP_ENVS * sizeof(type) + 15) & ~(size_t)15;                           \
  }
  MFG_STATE_FIELDS(F)
#undef F
  if (!live) return;
  env_step<AMAX>(sp, tb, ss, el, io, eg, (int64_t)Ng);
#define F(type, name, rows_expr)                                                                   \
  if constexpr (!std::is_same<type, double>::value) {                                              \
    const int rows = (int)(rows_expr);                                                             \
    for (int r = 0; r < rows; ++r) st.name[(size_t)r * Ng + eg] = ss.name[r * STEP_ENVS + el];     \
  }
  MFG_STATE_FIELDS(F)
#undef F
}

__global__ void __launch_bounds__(256) k_random_actions(const MfgSpec* __restrict__ sp, int64_t N, int64_t env_id_offset,
                                                        int32_t* actions, uint64_t seed, uint32_t step_index) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  Philox rng;
  rng.init(seed, (uint64_t)(env_id_offset + e), RS_ACTIONS, 0, step_index);
  const int A = sp->n_agents;
  for (int i = 0; i < A; ++i) actions[(size_t)e * A + i] = (int32_t)rng.below((uint32_t)sp->n_actions[i]);
}

namespace mfg {

template <typename Fn>
static void dispatch_amax(int A, Fn fn) {
  if (A <= 1) fn(std::integral_constant<int, 1>());
  else if (A <= 2) fn(std::integral_constant<int, 2>());
  else if (A <= 4) fn(std::integral_constant<int, 4>());
  else if (A <= 8) fn(std::integral_constant<int, 8>());
  else fn(std::integral_constant<int, 16>());
}

cudaError_t launch_reset(MfgHandle* h, const uint8_t* d_mask, cudaStream_t s) {
  const int threads = 128;
  const unsigned blocks = (unsigned)((h->N + threads - 1) / threads);
  dispatch_amax(h->sp.n_agents, [&](auto amax) {
    k_reset<decltype(amax)::value><<<blocks, threads, 0, s>>>(h->d_sp, h->tb, h->st, d_mask);
  });
  return cudaGetLastError();
}

size_t step_stage_bytes(const MfgSpec& sp) {
  size_t off = 0;
#define F(type, name, rows_expr) \
  if (!std::is_same<type, double>::value) off += ((size_t)(rows_expr) * STEP_ENVS * sizeof(type) + 15) & ~(size_t)15;
  MFG_STATE_FIELDS(F)
#undef F
  return off;
}

cudaError_t launch_step(MfgHandle* h, const StepIO& io, cudaStream_t s) {
  const unsigned blocks = (unsigned)((h->N + STEP_ENVS - 1) / STEP_ENVS);
  const size_t smem = step_stage_bytes(h->sp);
  cudaError_t err = cudaSuccess;
  dispatch_amax(h->sp.n_agents, [&](auto amax) {
    auto kern = k_step<decltype(amax)::value>;
    if (smem > 48 * 1024) err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err == cudaSuccess) kern<<<blocks, STEP_ENVS, smem, s>>>(h->d_sp, h->tb, h->st, io);
  });
  return err != cudaSuccess ? err : cudaGetLastError();
}

cudaError_t launch_random_actions(MfgHandle* h, int32_t* d_actions, uint64_t seed, uint32_t step_index, cudaStream_t s) {
  const int threads = 256;
  const unsigned blocks = (unsigned)((h->N + threads - 1) / threads);
  k_random_actions<<<blocks, threads, 0, s>>>(h->d_sp, h->N, h->tb.env_id_offset, d_actions, seed, step_index);
  return cudaGetLastError();
}

}  // namespace mfg
