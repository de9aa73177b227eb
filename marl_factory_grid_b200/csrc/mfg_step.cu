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
__global__ void __launch_bounds__(128) k_reset(const MfgSpec* __restrict__ sp, Tables tb, State st, const uint8_t* mask, int first) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= st.N) return;
  if (mask && !mask[e]) return;
  // episode = third word of the Philox counter: every reset of an env draws a fresh layout; episode 0 is the first
  // full reset of a handle (or the first after the "reseed" option)
  uint32_t episode = first ? 0u : field_at(st, st.episode, 0, e) + 1;
  env_reset<AMAX>(*sp, tb, st, e, episode);
}

constexpr int STEP_ENVS = ENV_BLOCK;
constexpr int STEP_ENVS_R = ENV_BLOCK;   // threads (= private columns) per CTA of the packed reset kernel

// deferred auto-reset: one thread per FINISHED env (ids appended by k_step), so the long Philox spawn path runs in
// fully packed warps instead of dragging 1-2 lanes of almost every step warp through it.  Each thread gathers its
// env's integer column into a private column of a shared-memory block image (independent loads), re-spawns against
// that copy (hundreds of dependent look-ups stay on chip) and scatters the column back.
template <int AMAX>
__global__ void __launch_bounds__(STEP_ENVS_R) k_reset_list(const MfgSpec* __restrict__ sp, Tables tb, State st, ColTab ct,
                                                            const uint32_t* __restrict__ list, const uint32_t* __restrict__ count) {
  extern __shared__ __align__(128) unsigned char stage[];
  const uint32_t n = *count;
  const int t = threadIdx.x;
  const State ss = staged_view(st, stage);
  const MfgSpec& spr = *sp;
  for (uint32_t base = blockIdx.x * STEP_ENVS_R; base < n; base += gridDim.x * STEP_ENVS_R) {
    const int n_here = (int)(n - base < (uint32_t)STEP_ENVS_R ? n - base : (uint32_t)STEP_ENVS_R);
    gather_columns(st, stage, ct, list, base, n_here);          // all threads: (row, env) pairs, loads in flight together
    __syncthreads();
    if (t < n_here) {
      const int64_t e = list[base + t];
      env_reset_inl<AMAX>(spr, tb, ss, t, field_at(ss, ss.episode, 0, t) + 1, e);     // inlined: State / Tables stay in registers
    }
    __syncthreads();
    scatter_columns(st, stage, ct, list, base, n_here);
    __syncthreads();
  }
}

// k_step: CTA = one or two 128-env state blocks.  A block's integer / byte fields are ONE contiguous slab in HBM (blocked
// layout, see State): TMA bulk copies (cp.async.bulk, completion on an mbarrier) stage it in shared memory, the whole
// step runs against that copy (tile look-ups, slot scans and rule hooks hit shared memory instead of dependent HBM
// round trips), and bulk stores write it back.  f64 fields (dirt amounts, battery, returns) stay in HBM: few actions /
// rules touch them.  Launch shapes: see the comment on the kernel below.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MFG_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MFG_DONE_%=;\n"
      "bra MFG_WAIT_%=;\n"
      "MFG_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}

// `so` = the state view of the shared-memory image: integer field pointers hold byte OFFSETS into the image (see
// stage_ref), f64 field pointers are the global ones.  Built by launch_step_kernel.
// SYNC = CTA barriers at the step's convergent points (see env_step): the four warps of the block stay on the same
// instruction-cache lines instead of each streaming the 7 k-instruction program from L2 on its own.
// SPLIT = the dirt-uid rows [cut, cut + gap) of the block stay in HBM (so.dirt_uid is then the global slab): they are only
// touched by the uid listing of faithful mode and by create / compact, and without their 10 KB a sixth CTA fits on the SM.
// NB = state blocks per CTA (128 threads and one image each): larger lockstep groups share more instruction fetches and wait
// longer at the barriers.
template <int AMAX, bool FLAGS, bool SYNC, bool SPLIT, int NB>
__global__ void __launch_bounds__(STEP_ENVS * NB, (SPLIT ? 6 : 5) / NB)
k_step(const __grid_constant__ HotSpec<AMAX, SPLIT> hs, const MfgSpec* __restrict__ full, const __grid_constant__ Tables tb,
       const __grid_constant__ State st, const __grid_constant__ State so, const __grid_constant__ StepIO io,
       const uint32_t cut, const uint32_t gap, const int n_blocks) {
  extern __shared__ __align__(128) unsigned char stage[];
  __shared__ __align__(8) unsigned long long bar;
  const int t = threadIdx.x, g = NB == 1 ? 0 : t >> 7, el = t & (STEP_ENVS - 1);
  const int b0 = blockIdx.x * NB, nb_here = n_blocks - b0 < NB ? n_blocks - b0 : NB;        // state blocks of this CTA
  const int64_t eg = (int64_t)(b0 + g) * STEP_ENVS + el;
  char* gblock = st.base_i + (size_t)b0 * st.blk_i;
  const uint32_t bytes = (uint32_t)st.blk_i - (SPLIT ? gap : 0u);          // image size
  const uint32_t tail = bytes - cut;                                        // bytes after the gap

  if (t == 0) {
    mbar_init(&bar, 1);
    mbar_expect_tx(&bar, bytes * nb_here);
    for (int k = 0; k < nb_here; ++k) {
      unsigned char* img = stage + (size_t)k * bytes;
      const char* src = gblock + (size_t)k * st.blk_i;
      if (SPLIT) {
        if (cut) bulk_g2s(img, src, cut, &bar);
        if (tail) bulk_g2s(img + cut, src + cut + gap, tail, &bar);
      } else {
        bulk_g2s(img, src, bytes, &bar);
      }
    }
  }
  // level tables the step touches with divergent indices: wall map, tile -> door map, door positions.  With most of
  // the SM's unified cache carved out as shared memory they would otherwise be L2 round trips.
  const int HW = hs.H * hs.W, HW4 = (HW + 3) >> 2;
  uint32_t* s_wall = reinterpret_cast<uint32_t*>(stage + (size_t)NB * bytes);
  uint32_t* s_dmap = s_wall + HW4;
  uint16_t* s_dpos = reinterpret_cast<uint16_t*>(s_dmap + HW4);
  {   // 32-bit copies (build_tables pads both tables to a multiple of 4 bytes)
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(tb.wall);
    const uint32_t* gd = reinterpret_cast<const uint32_t*>(tb.door_map);
    for (int i = t; i < HW4; i += STEP_ENVS * NB) { s_wall[i] = gw[i]; s_dmap[i] = gd[i]; }
  }
  if (t < hs.n_doors) s_dpos[t] = tb.door_pos[t];
  Tables tbs = tb;                  // staged tables: offsets relative to the thread's image, like the integer fields of `so`
  const uintptr_t toff = (uintptr_t)(NB - g) * bytes;
  tbs.wall = reinterpret_cast<const uint8_t*>(toff);
  tbs.door_map = reinterpret_cast<const uint8_t*>(toff + (uintptr_t)HW4 * 4);
  tbs.door_pos = reinterpret_cast<const uint16_t*>(toff + (uintptr_t)HW4 * 8);
  uint32_t sbase;
  asm volatile("mov.u32 %0, %1;\n" : "=r"(sbase) : "r"(smem_u32(stage) + (uint32_t)g * bytes));      // opaque: kept in ONE register, never re-derived

  __syncthreads();                  // barrier init + table copies visible
  mbar_wait(&bar, 0);

  if (eg < st.N) env_step<AMAX, HotSpec<AMAX, SPLIT>, FLAGS, SYNC>(hs, *full, tbs, so, el, io, eg, sbase);
  else if (SYNC) { for (int k = step_sync_points(hs); k > 0; --k) step_sync<true>(); }

  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  if (t == 0) {
    for (int k = 0; k < nb_here; ++k) {
      const unsigned char* img = stage + (size_t)k * bytes;
      char* dst = gblock + (size_t)k * st.blk_i;
      if (SPLIT) {
        if (cut) bulk_s2g(dst, img, cut);
        if (tail) bulk_s2g(dst + cut + gap, img + cut, tail);
      } else {
        bulk_s2g(dst, img, bytes);
      }
    }
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");     // shared memory must outlive the copy
  }
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
    k_reset<decltype(amax)::value><<<blocks, threads, 0, s>>>(h->d_sp, h->tb, h->st, d_mask, (!d_mask && !h->ever_reset) ? 1 : 0);
  });
  return cudaGetLastError();
}

// k_step alone (the deferred auto-reset list is filled but not consumed)
cudaError_t launch_step_kernel(MfgHandle* h, const StepIO& io, cudaStream_t s) {
  const unsigned blocks = (unsigned)((h->N + STEP_ENVS - 1) / STEP_ENVS);
  const size_t hw4 = ((size_t)h->sp.H * h->sp.W + 3) / 4 * 4;
  cudaError_t err = cudaSuccess;
  dispatch_amax(h->sp.n_agents, [&](auto amax) {
    constexpr int AMAX = decltype(amax)::value;
    if (io.auto_reset && io.reset_list) err = cudaMemsetAsync(io.reset_count, 0, sizeof(uint32_t), s);
    // the in-kernel re-spawn (auto-reset without the deferred list) works on the whole block image
    const bool split = h->step_kernel == 1 && !(io.auto_reset && !io.reset_list);
    // the rows left in HBM: all dirt-uid rows but the first few (as many as the largest small class has members - the uid
    // listing never looks further for a class member's uid, see Env::find_listed)
    const uint32_t uid_off = (uint32_t)(reinterpret_cast<char*>(h->st.dirt_uid) - h->st.base_i), row_b = ENV_BLOCK * (uint32_t)sizeof(uint16_t);
    int head = 0;
    for (int n : {h->sp.n_items, h->sp.n_pods, h->sp.n_dest, h->sp.n_dropoff, h->sp.n_machines, h->sp.n_maint}) head = n > head ? n : head;
    const int uid_rows = h->sp.has_dirt ? h->sp.dirt_slots : 0;
    if (head > uid_rows) head = uid_rows;
    const uint32_t cut = uid_off + (uint32_t)head * row_b;
    const uint32_t gap = (uint32_t)(uid_rows - head) * row_b;
    State so = h->st;                 // integer fields: offsets into the image
    so.N = STEP_ENVS;
#define F(type, name, rows_expr)                                                                              \
  if (!std::is_same<type, double>::value) {                                                                    \
    size_t off = (size_t)(reinterpret_cast<char*>(h->st.name) - h->st.base_i);                                 \
    if (split && off > cut) off -= gap;                                                                        \
    so.name = reinterpret_cast<type*>(off);                                                                    \
  }
    MFG_STATE_FIELDS(F)
#undef F
    if (split) so.dirt_uid = h->st.dirt_uid;       // stays global (Env::uid_at)
    const int n_blocks = (int)blocks;
    auto go = [&](auto kern, auto hs, int nb) {
      const size_t smem = (size_t)nb * (h->st.blk_i - (split ? gap : 0)) + 2 * hw4 + 2 * MFG_MAX_DOORS + 16;
      fill_hot_spec(h->sp, hs);
      hs.uid_head_off = uid_off;
      hs.uid_head_rows = head;
      if (err == cudaSuccess && smem > 48 * 1024) err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err == cudaSuccess) kern<<<(unsigned)((n_blocks + nb - 1) / nb), STEP_ENVS * nb, smem, s>>>(hs, h->d_sp, h->tb, h->st, so, io, cut, gap, n_blocks);
    };
    // step_kernel: 1 = barriers + split image (default); 2 = barriers, whole image; 0 = neither (the round-2 baseline)
    // blocks per CTA: two when three such CTAs still fit on the SM (same 24 warps, half as many lockstep groups: 0.56 -> 0.49 ms
    // on cfg4; three blocks: 0.51), else one
    const size_t img = h->st.blk_i - (split ? gap : 0), extra = 2 * hw4 + 2 * MFG_MAX_DOORS + 16 + 1024;
    const int nb = h->step_blocks ? h->step_blocks : (3 * (2 * img + extra) <= (size_t)228 * 1024 ? 2 : 1);
    if (split && nb == 2) { if (io.flags) go(k_step<AMAX, true, true, true, 2>, HotSpec<AMAX, true>(), 2); else go(k_step<AMAX, false, true, true, 2>, HotSpec<AMAX, true>(), 2); }
    else if (split && nb == 3) { if (io.flags) go(k_step<AMAX, true, true, true, 3>, HotSpec<AMAX, true>(), 3); else go(k_step<AMAX, false, true, true, 3>, HotSpec<AMAX, true>(), 3); }
    else if (split) { if (io.flags) go(k_step<AMAX, true, true, true, 1>, HotSpec<AMAX, true>(), 1); else go(k_step<AMAX, false, true, true, 1>, HotSpec<AMAX, true>(), 1); }
    else if (h->step_kernel != 0) { if (io.flags) go(k_step<AMAX, true, true, false, 1>, HotSpec<AMAX, false>(), 1); else go(k_step<AMAX, false, true, false, 1>, HotSpec<AMAX, false>(), 1); }
    else { if (io.flags) go(k_step<AMAX, true, false, false, 1>, HotSpec<AMAX, false>(), 1); else go(k_step<AMAX, false, false, false, 1>, HotSpec<AMAX, false>(), 1); }
  });
  return err != cudaSuccess ? err : cudaGetLastError();
}

// packed re-spawn of the envs k_step appended to the reset list
cudaError_t launch_reset_list(MfgHandle* h, const StepIO& io, cudaStream_t s) {
  if (!(io.auto_reset && io.reset_list)) return cudaSuccess;
  const unsigned blocks = (unsigned)((h->N + STEP_ENVS - 1) / STEP_ENVS);
  cudaError_t err = cudaSuccess;
  dispatch_amax(h->sp.n_agents, [&](auto amax) {
    constexpr int AMAX = decltype(amax)::value;
    // ~0.2 % of the envs finish per step (a handful of busy CTAs); the grid-stride loop covers larger lists
    const unsigned rblocks = blocks < 1184 ? blocks : 1184;      // CTAs beyond the list exit at once; a max-steps boundary can list most envs
    auto rk = k_reset_list<AMAX>;
    if (h->st.blk_i > 48 * 1024) err = cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->st.blk_i);
    if (err == cudaSuccess) rk<<<rblocks, STEP_ENVS_R, h->st.blk_i, s>>>(h->d_sp, h->tb, h->st, ColTab{h->d_row_tab, h->n_row_tab}, io.reset_list, io.reset_count);
  });
  return err != cudaSuccess ? err : cudaGetLastError();
}

cudaError_t launch_step(MfgHandle* h, const StepIO& io, cudaStream_t s) {
  cudaError_t err = launch_step_kernel(h, io, s);
  return err != cudaSuccess ? err : launch_reset_list(h, io, s);
}

cudaError_t launch_random_actions(MfgHandle* h, int32_t* d_actions, uint64_t seed, uint32_t step_index, cudaStream_t s) {
  const int threads = 256;
  const unsigned blocks = (unsigned)((h->N + threads - 1) / threads);
  k_random_actions<<<blocks, threads, 0, s>>>(h->d_sp, h->N, h->tb.env_id_offset, d_actions, seed, step_index);
  return cudaGetLastError();
}

}  // namespace mfg
