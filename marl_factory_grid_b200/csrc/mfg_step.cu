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

template <int AMAX>
__global__ void __launch_bounds__(128) k_step(const MfgSpec* __restrict__ sp, Tables tb, State st, StepIO io) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= st.N) return;
  env_step<AMAX>(*sp, tb, st, e, io);
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

cudaError_t launch_step(MfgHandle* h, const StepIO& io, cudaStream_t s) {
  const int threads = 128;
  const unsigned blocks = (unsigned)((h->N + threads - 1) / threads);
  dispatch_amax(h->sp.n_agents, [&](auto amax) {
    k_step<decltype(amax)::value><<<blocks, threads, 0, s>>>(h->d_sp, h->tb, h->st, io);
  });
  return cudaGetLastError();
}

cudaError_t launch_random_actions(MfgHandle* h, int32_t* d_actions, uint64_t seed, uint32_t step_index, cudaStream_t s) {
  const int threads = 256;
  const unsigned blocks = (unsigned)((h->N + threads - 1) / threads);
  k_random_actions<<<blocks, threads, 0, s>>>(h->d_sp, h->N, h->tb.env_id_offset, d_actions, seed, step_index);
  return cudaGetLastError();
}

}  // namespace mfg
