// mfg_obs.cu - observation kernels: OBSBuilder.build_for_all + RayCaster.visible_entities
// (marl_factory_grid/utils/observation_builder.py:98-235, utils/ray_caster.py:66-199).
//
//   k_obs_tiled<R, FAITHFUL>   CTA = one 128-env state block (its positional prefix staged by one TMA bulk copy), 8 autonomous
//                warps; both parity modes.  A list mode observes the envs of a device-side id list (re-spawned envs).
//       phase 1  lane = (env, agent), 32 / A_pad envs per warp pass: window visibility from a per-tile table keyed by the
//                closed doors of the window (<= 4 doors; else the 49-bit light-block mask + ray march over the constexpr
//                window-ray trie: straight-line bit tests), one pass over the listed entities -> visibility masks per class, then every visible in-window entity
//                becomes an 8-byte "sprite" (plane-cell index, kind, value) in shared memory; walls stay a 49-bit mask.
//                Faithful mode adds the uid de-duplication of `set(visible_entities)`: candidate mask of the radius-D box,
//                uid conflict masks, first-visit ranks on demand (first_visit_rank).
//       phase 2  the same warp, one env at a time, `ppp` planes (a part of the tile) at a time: zero the part buffer in
//                shared memory, set the wall cells (lane = a third of a wall plane), write the scalar channels (lane = list
//                entry) and the sprites (lanes that hit the same cell are grouped by a warp match: integer stacks first,
//                fractional encodings last => the f64 sums of the reference are reproduced with one rounding), then ONE TMA
//                bulk store (cp.async.bulk shared -> global) of the 16-byte aligned part.  The observation write is
//                87-93 % of the algorithmic bytes of an env-step.  Identity: 64 registers / 53 KB => 4 CTAs per SM.
//   k_obs_block   exact per-agent path over whole state blocks (slab staged by one bulk copy): full observability and specs
//                the tiled kernel cannot take; first-visit ranks from a per-(tile, closed-door subset) table when built.
//   k_obs_redo    exact per-agent path over a device-side env list (sprite / conflict list overflow), columns staged in
//                shared memory.
//   k_obs_direct  the same exact path on plain global state, one thread per (env, agent) (obs_kernel option 3).
#include <utility>
#include "mfg_internal.hpp"
#include "mfg_rays_gen.h"

using namespace mfg;

template <int AMAX>
__global__ void __launch_bounds__(128) k_obs_direct(const MfgSpec* __restrict__ sp, Tables tb, State st, float* obs,
                                                    int total_channels, int stride) {
  // agent-major thread mapping: consecutive threads = consecutive envs of the same agent (coalesced state loads)
  extern __shared__ __align__(16) float s_planes[];
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int A = sp->n_agents;
  const int DD = obs_plane_cells(*sp);
  const bool live = t < st.N * A;
  const int a = live ? (int)(t / st.N) : 0;
  const int64_t e = live ? t - (int64_t)a * st.N : 0;
  float* out = obs + ((size_t)e * total_channels + sp->ch_offset[a]) * DD;
  if (stride == 0) {                   // planes too large for shared memory: accumulate in the output tensor itself
    if (live) obs_agent_direct<AMAX>(*sp, tb, st, e, a, out);
    return;
  }
  // small planes (full observability on the small shipped levels): every thread composes its agent's planes in its own
  // shared-memory slice (the accumulation is a read-modify-write per contribution - in global memory that was a dependent
  // HBM round trip each), then the CTA copies the slices out with coalesced stores
  float* mine = s_planes + (size_t)threadIdx.x * stride;
  const int nfl = sp->n_channels[a] * DD;
  if (live) obs_agent_direct<AMAX>(*sp, tb, st, e, a, mine);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < (int)blockDim.x; j += (int)(blockDim.x >> 5)) {      // slice j belongs to thread j of the CTA
    const int64_t tj = (int64_t)blockIdx.x * blockDim.x + j;
    if (tj >= st.N * A) break;
    const int aj = (int)(tj / st.N);
    const int64_t ej = tj - (int64_t)aj * st.N;
    float* dst = obs + ((size_t)ej * total_channels + sp->ch_offset[aj]) * DD;
    const float* src = s_planes + (size_t)j * stride;
    const int n = sp->n_channels[aj] * DD;
    for (int i = lane; i < n; i += 32) dst[i] = src[i];
  }
  (void)nfl;
}

// ---------------------------------------------------------------------------------------------------------------
// ray march on a (2R+1)^2-bit light-block mask (ray_caster.py:81-103): straight-line code from the constexpr trie
// ---------------------------------------------------------------------------------------------------------------
template <int R, int n>
__device__ __forceinline__ void trie_step(const unsigned long long B, unsigned long long& vis, unsigned long long& cont) {
  using T = RayTrie<R>;
  constexpr int p = T::parent(n), c = T::cell(n), da = T::da(n), db = T::db(n);
  const bool reach = p < 0 ? true : (((cont >> p) & 1ull) != 0);
  const bool hits = (B & (1ull << c)) != 0;
  bool diag = false;
  if constexpr (da != 255) diag = ((B & (1ull << da)) != 0) && ((B & (1ull << db)) != 0);
  if (reach && !diag) vis |= 1ull << c;
  if (reach && !hits && !diag) cont |= 1ull << n;
}
template <int R, int... I>
__device__ __forceinline__ void march_impl(const unsigned long long B, unsigned long long& vis, unsigned long long& cont,
                                           std::integer_sequence<int, I...>) {
  (trie_step<R, I>(B, vis, cont), ...);
}
template <int R>
__device__ __forceinline__ unsigned long long march(const unsigned long long B) {
  constexpr int D = 2 * R + 1, centre = R * D + R;
  unsigned long long vis = 1ull << centre, cont = 0ull;
  if (B & (1ull << centre)) return vis;           // inside a closed door: only the own tile (origin cell stops every ray)
  march_impl<R>(B, vis, cont, std::make_integer_sequence<int, RayTrie<R>::N>());
  return vis;
}

// faithful mode: first-visit rank (= trie node id, the reference's visit order) of one cell of the radius-D box, or
// 0x7FFF when no ray reaches it.  b0..b3 = light-block mask of the (2D+1)^2 box.  Only the trie paths that end in this
// cell are walked: a node is visited iff every ancestor is visited and passable and the node is not diagonally occluded
// (ray_caster.py:81-103).
template <int R>
__device__ __noinline__ int first_visit_rank(unsigned long long b0, unsigned long long b1, unsigned long long b2,
                                             unsigned long long b3, int bx, int by) {
  using T = FullTrie<R>;
  constexpr int D = 2 * R + 1, BW = 2 * D + 1, origin = D * BW + D;
  if (bx < -D || bx > D || by < -D || by > D) return 0x7FFF;
  if (bx == 0 && by == 0) return -1;                                      // the own tile comes first
  auto blk = [&](int ci) -> bool {
    const int w = ci >> 6;
    const unsigned long long word = w == 0 ? b0 : w == 1 ? b1 : w == 2 ? b2 : b3;
    return ((word >> (ci & 63)) & 1ull) != 0;
  };
  if (blk(origin)) return 0x7FFF;                                         // inside a closed door: only the own tile
  const int ci = (bx + D) * BW + (by + D);
  for (int i = T::cell_off(ci); i < T::cell_off(ci + 1); ++i) {
    const int n = T::cell_node(i);
    uint32_t rec = T::node(n);
    int da = (rec >> 16) & 255, db = rec >> 24;
    bool ok = !(da != 255 && blk(da) && blk(db));
    int p = rec & 255;
    while (ok && p != 255) {
      rec = T::node(p);
      da = (rec >> 16) & 255; db = rec >> 24;
      ok = !blk((rec >> 8) & 255) && !(da != 255 && blk(da) && blk(db));
      p = rec & 255;
    }
    if (ok) return n;
  }
  return 0x7FFF;
}

// ---------------------------------------------------------------------------------------------------------------
// sprites
// ---------------------------------------------------------------------------------------------------------------
enum { SK_INT = 0, SK_STORE = 1, SK_DOOR = 2, SK_DIRT = 3 };
struct Sprite { uint32_t w; float val; };          // w = index | kind << 16 | aux << 24

// positions of one env inside the staged block prefix: slot-major slabs of 128 envs (see MFG_STATE_FIELDS order)
struct BlkPos {
  const uint16_t* base;
  int eb;                                   // env index inside the 128-env block
  __device__ __forceinline__ uint16_t operator[](int s) const { return base[s * ENV_BLOCK + eb]; }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store_tile(float* dst, const float* src_smem, uint32_t bytes) {
#if defined(MFG_ZFILL_EVICT_LAST)
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\n" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes), "l"(pol) : "memory");
#else
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
#endif
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}

// Envs whose sprite list overflowed (any mode) or that have more uid conflicts than the packed list holds (faithful mode)
// are appended to a redo list; this kernel rewrites their observations with the exact per-agent path of mfg_core.cuh.
constexpr int REDO_ENVS = 32;       // envs per CTA round of k_obs_redo (columns of the block image in use)
template <int AMAX>
__global__ void __launch_bounds__(128) k_obs_redo(const MfgSpec* __restrict__ sp, Tables tb, State st, ColTab ct, float* obs,
                                                  int total_channels, const uint32_t* __restrict__ list,
                                                  const uint32_t* __restrict__ count) {
  extern __shared__ __align__(128) unsigned char stage[];
  const int A = sp->n_agents;
  const int DD = obs_plane_cells(*sp);
  const uint32_t n = *count;
  const State ss = staged_view(st, stage);
  for (uint32_t base = blockIdx.x * REDO_ENVS; base < n; base += gridDim.x * REDO_ENVS) {
    const int n_here = (int)(n - base < (uint32_t)REDO_ENVS ? n - base : (uint32_t)REDO_ENVS);
    gather_columns(st, stage, ct, list, base, n_here);          // the envs' integer columns -> shared-memory image
    __syncthreads();
    for (int i = threadIdx.x; i < n_here * A; i += blockDim.x) {
      const int j = i / A, a = i - j * A;
      const int64_t e = list[base + j];
      obs_agent_direct<AMAX>(*sp, tb, ss, j, a, obs + ((size_t)e * total_channels + sp->ch_offset[a]) * DD, e);
    }
    __syncthreads();
  }
}

// Exact per-agent path over whole state blocks: CTA = one 128-env block, its integer slab staged in shared memory by one
// bulk copy (like k_step), one thread per (env, agent) of the block.  The product path of full observability and of specs
// the tiled kernel cannot take; with the rank table (Tables::rank_tab) the ray walk is a look-up.
template <int AMAX>
__global__ void __launch_bounds__(256) k_obs_block(const MfgSpec* __restrict__ sp, Tables tb, State st, float* obs, int total_channels);

__device__ __forceinline__ void mbar_init1(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_wait0(unsigned long long* bar) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MFG_OWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@p bra MFG_ODONE_%=;\n"
      "bra MFG_OWAIT_%=;\n"
      "MFG_ODONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)) : "memory");
}

template <int AMAX>
__global__ void __launch_bounds__(256) k_obs_block(const MfgSpec* __restrict__ sp, Tables tb, State st, float* obs, int total_channels) {
  extern __shared__ __align__(128) unsigned char stage[];
  __shared__ __align__(8) unsigned long long bar;
  const int A = sp->n_agents;
  const int DD = obs_plane_cells(*sp);
  const int64_t blk0 = (int64_t)blockIdx.x * ENV_BLOCK;
  const int n_live = (int)(st.N - blk0 < ENV_BLOCK ? st.N - blk0 : ENV_BLOCK);
  if (threadIdx.x == 0) {
    mbar_init1(&bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"((uint32_t)st.blk_i) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(stage)),
                 "l"(st.base_i + (size_t)blockIdx.x * st.blk_i), "r"((uint32_t)st.blk_i), "r"(smem_u32(&bar)) : "memory");
  }
  const State ss = staged_view(st, stage);
  // the block's observations are one contiguous run of the output tensor: cleared here with coalesced stores (a per-thread
  // clear of its own planes wrote every sector four times)
  {
    float* o0 = obs + (size_t)blk0 * total_channels * DD;
    const size_t nfl = (size_t)n_live * total_channels * DD;
    if ((reinterpret_cast<uintptr_t>(o0) & 15) == 0) {
      float4* o4 = reinterpret_cast<float4*>(o0);
      for (size_t i = threadIdx.x; i < (nfl >> 2); i += blockDim.x) o4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (size_t i = (nfl & ~(size_t)3) + threadIdx.x; i < nfl; i += blockDim.x) o0[i] = 0.f;
    } else {
      for (size_t i = threadIdx.x; i < nfl; i += blockDim.x) o0[i] = 0.f;
    }
  }
  __syncthreads();
  mbar_wait0(&bar);
  // agent-major inside the block: consecutive threads = consecutive envs of the same agent
  for (int i = threadIdx.x; i < n_live * A; i += blockDim.x) {
    const int a = i / n_live, j = i - a * n_live;
    const int64_t e = blk0 + j;
    obs_agent_direct<AMAX>(*sp, tb, ss, j, a, obs + ((size_t)e * total_channels + sp->ch_offset[a]) * DD, e, false);
  }
}

// CTA = one 128-env state block.  Every warp works on its own sub-groups of EPW = 32 / APAD envs (APAD = agent count
// rounded up to a power of two): lane = (env, agent) in phase 1, the same warp expands the tiles of those envs in phase 2,
// so after the prefix has landed no CTA-wide barrier is needed and warps in different phases overlap on the SM.
#ifndef MFG_OBS_CTAS_F
#define MFG_OBS_CTAS_F 2          // resident CTAs per SM the faithful kernel is compiled for (register cap 128 / 80)
#endif
template <int R, bool FAITHFUL>
__global__ void __launch_bounds__(256, FAITHFUL ? MFG_OBS_CTAS_F : 4) k_obs_tiled(const MfgSpec* __restrict__ sp, Tables tb, State st, ObsSlots sl, WallPlanes wp,
                                                    float* __restrict__ obs, int total_channels, int cap, int apad_log2,
                                                    int bulk, int ppp, uint32_t* __restrict__ redo,
                                                    const uint8_t* __restrict__ skip, ObsList ol,
                                                    const uint32_t* __restrict__ prog, int prog_words) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  constexpr int D = 2 * R + 1, DD = D * D;
  const int A = sp->n_agents;
  const int NW = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int EPW = 32 >> apad_log2;                                 // envs per warp pass
  const int64_t blk0 = (int64_t)blockIdx.x * ENV_BLOCK;
  // list mode (ol.ids != null): the CTA's 128 "envs" are entries of a device-side id list (the envs re-spawned in this step);
  // their columns are gathered into the same shared-memory image the block mode fills with one bulk copy
  const bool lmode = ol.ids != nullptr;
  const uint32_t n_listed = lmode ? *ol.count : 0u;
  const int tile_floats = total_channels * DD;                     // one env's packed observation
  const int part_floats = ppp * DD;                                // the tile is composed `ppp` planes at a time (bulk: a multiple of 4 planes = whole 16-byte vectors)

  // shared memory carve-up (every region start stays 16-byte aligned):
  //   prefix | per warp: part buffer, sprite lists, counters | candidate masks (faithful) | channel masks | x / H, y / W
  unsigned char* s_blk = smem_raw;                                                          // staged block prefix
  const size_t part_bytes = ((size_t)part_floats * 4 + 15) & ~(size_t)15;
  const size_t per_warp = part_bytes + (size_t)EPW * cap * 8 + (((size_t)EPW * 4 + 15) & ~(size_t)15);
  unsigned char* wbase = smem_raw + sl.prefix_bytes + (size_t)warp * per_warp;
  float* tile = reinterpret_cast<float*>(wbase);                                            // one part of one env's tile
  Sprite* s_spr = reinterpret_cast<Sprite*>(wbase + part_bytes);                            // [EPW][cap]
  int* s_cnt = reinterpret_cast<int*>(wbase + part_bytes + (size_t)EPW * cap * 8);          // [EPW]
  unsigned char* misc = smem_raw + sl.prefix_bytes + (size_t)NW * per_warp;
  uint32_t* s_cm = reinterpret_cast<uint32_t*>(misc) + (size_t)warp * 256 + lane;      // [NW][8][32] 32-bit words: word w of this lane = s_cm[w * 32]
  // the per-spec constant "observation program" (built once by plan_obs, see ObsProg): per-agent channel offsets and scalar
  // channels, the wall planes, term -> channel masks, the GlobalPosition encodings x / H, y / W (entity/util.py:56-66)
  uint32_t* s_prog = reinterpret_cast<uint32_t*>(misc + (FAITHFUL ? (size_t)NW * 1024 : 0));
  const int* s_coff = reinterpret_cast<const int*>(s_prog + ObsProg::COFF);
  const uint32_t* s_hasbat = s_prog + ObsProg::HASBAT;
  const uint32_t* s_scl = s_prog + ObsProg::SCL;              // scalar channels of all agents: packed plane | kind << 12 | agent << 16
  const uint32_t* s_wplane = s_prog + ObsProg::WPLANE;        // wall planes: packed channel index | agent << 10, ascending
  const uint32_t* s_chm = s_prog + ObsProg::CHM;              // [A][MFG_N_TERMS] term -> channel bits
  const float* s_gx = reinterpret_cast<const float*>(s_chm + ((A * MFG_N_TERMS + 3) & ~3));
  const float* s_gy = s_gx + ((sp->H + 3) & ~3);
  // ---- stage the positional prefix of this block: one TMA bulk copy (dirt/item/.../agent positions, door + dest masks)
  // block mode: one pass; list mode: grid-stride over chunks of 128 listed envs
  for (uint32_t lbase = blockIdx.x * ENV_BLOCK;; lbase += gridDim.x * ENV_BLOCK) {
  int n_live;
  if (lmode) {
    if (lbase >= n_listed) break;
    n_live = (int)(n_listed - lbase < (uint32_t)ENV_BLOCK ? n_listed - lbase : (uint32_t)ENV_BLOCK);
  } else {
    n_live = (int)(st.N - blk0 < ENV_BLOCK ? st.N - blk0 : ENV_BLOCK);
  }
  auto env_of = [&](int eb) -> int64_t { return lmode ? (int64_t)ol.ids[lbase + eb] : blk0 + eb; };
  if (lmode) {
    gather_columns(st, s_blk, ColTab{ol.rows, ol.n_rows}, ol.ids, lbase, n_live);
  } else if (threadIdx.x == 0) {
    mbar_init1(&bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"((uint32_t)sl.prefix_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(s_blk)),
                 "l"(st.base_i + (size_t)blockIdx.x * st.blk_i), "r"((uint32_t)sl.prefix_bytes), "r"(smem_u32(&bar)) : "memory");
  }
  for (int i = threadIdx.x; i < prog_words; i += blockDim.x) s_prog[i] = prog[i];
  const int spW = sp->W, n_doors = sp->n_doors, n_dest = sp->n_dest, has_dirt = sp->has_dirt, n_walls = sp->n_walls;
  __syncthreads();
  const int n_scl = (int)s_prog[ObsProg::N_SCL], n_wp = (int)s_prog[ObsProg::N_WP];
  if (!lmode) mbar_wait0(&bar);
  const uint16_t* blk16 = reinterpret_cast<const uint16_t*>(s_blk);
  const unsigned long long* blk_dopen = reinterpret_cast<const unsigned long long*>(s_blk + sl.off_dopen);
  const uint32_t* blk_reached = reinterpret_cast<const uint32_t*>(s_blk + sl.off_reached);
  const unsigned long long* blk_dlisted = reinterpret_cast<const unsigned long long*>(s_blk + sl.off_door_listed);
  const unsigned long long* blk_dirtlisted = reinterpret_cast<const unsigned long long*>(s_blk + sl.off_dirt_listed);
  const uint16_t* blk_dirt_uid = reinterpret_cast<const uint16_t*>(s_blk + sl.off_dirt_uid);
  const int lo[6] = {sl.item0, sl.pod0, sl.dest0, sl.drop0, sl.mach0, sl.maint0};
  const int hi[6] = {sl.pod0, sl.dest0, sl.drop0, sl.mach0, sl.maint0, sl.agent0};

  for (int sub = warp; sub * EPW < ENV_BLOCK; sub += NW) {
    if (sub * EPW >= n_live) break;
    if (lane < EPW) s_cnt[lane] = 0;
    __syncwarp();

    // ---------------- phase 1: lane = (env el, agent a) ---------------------------------------------------------
    const int el = lane >> apad_log2, a = lane & ((1 << apad_log2) - 1);
    unsigned long long wv = 0ull;
    uint32_t skip_mask = 0u;
    bool lane_on = false;             // this lane computed an (env, agent) pair in phase 1
    float batv = 0.f;                 // Battery channel value of this (env, agent) lane
    uint32_t axy = 0u;                // agent position of this lane (GlobalPosition channel)          // bit (el << apad_log2): env el of this pass is skipped
    {
      const int eb = sub * EPW + el;
      const bool in_range = eb < n_live;
      const int64_t e = in_range ? env_of(eb) : 0;
      // skipped envs (being re-spawned concurrently) are rewritten later; one flag load per lane, shared through a ballot
      const bool lane_skip = skip != nullptr && in_range && skip[e] != 0;
      skip_mask = __ballot_sync(0xffffffffu, lane_skip);
      // faithful mode: the phase-1 results phase 2 needs (wall mask, battery level, position) are handed over through the
      // lane's candidate-mask slots (dead once the lane's classification is done) instead of registers that would have to
      // live across the whole pass (80-register budget for three resident CTAs per SM)
      if (FAITHFUL && !(a < A && in_range && !lane_skip)) { s_cm[0] = 0u; s_cm[32] = 0u; s_cm[64] = 0u; s_cm[96] = 0u; }
      if (a < A && in_range && !lane_skip) {
        const BlkPos pos{blk16, eb};
        const unsigned long long dopen = n_doors ? blk_dopen[eb] : 0ull;
        const uint32_t reached = n_dest ? blk_reached[eb] : 0u;
        // listing bits (uid-equality artefact): everything is listed in identity mode
        const unsigned long long dlisted = (FAITHFUL && n_doors) ? blk_dlisted[eb] : ~0ull;
        const unsigned long long dirtlisted = has_dirt ? blk_dirtlisted[eb] : 0ull;
        uint32_t glisted[6] = {~0u, ~0u, ~0u, ~0u, ~0u, ~0u};
        if (FAITHFUL) {
#pragma unroll
          for (int g = 0; g < 6; ++g)
            if (sl.off_listed[g] >= 0) glisted[g] = reinterpret_cast<const uint32_t*>(s_blk + sl.off_listed[g])[eb];
        }
        const uint16_t p = pos[sl.agent0 + a];
        const int ax = px(p), ay = py(p);
        axy = p;
        const int tile_id = ax * spW + ay;
        const unsigned long long W49 = tb.wall_win[tile_id];
        // static walls-only visibility of the radius-D box (faithful conflict filter); requested early, used late
        unsigned long long sb0 = 0ull, sb1 = 0ull, sb2 = 0ull, sb3 = 0ull;
        if (FAITHFUL) {
          const ulonglong2* sbp = reinterpret_cast<const ulonglong2*>(tb.vis_box) + (size_t)tile_id * 2;
          const ulonglong2 u0 = __ldg(sbp), u1 = __ldg(sbp + 1);
          sb0 = u0.x; sb1 = u0.y; sb2 = u1.x; sb3 = u1.y;
        }
        // doors inside the radius-D box of this tile (static table): usually 0..3 of them
        const unsigned long long dnear = n_doors ? (tb.door_near[tile_id] & dlisted) : 0ull;
        // window visibility: a table look-up keyed by (tile, closed listed doors among the <= 4 doors of the window); the
        // ray march itself only runs on tiles whose window holds more doors
        unsigned long long vis;
#ifndef MFG_OBS_VISTAB_I
#define MFG_OBS_VISTAB_I 1        // identity mode: 1 = visibility table as well, 0 = always march (the kernel is issue-bound at 4 CTAs per SM)
#endif
        const uint32_t dwin = (FAITHFUL || MFG_OBS_VISTAB_I) ? tb.door_win[tile_id] : (7u << 24);
        if ((FAITHFUL || MFG_OBS_VISTAB_I) && (dwin >> 24) <= 4u) {
          const unsigned long long closed = dnear & ~dopen;
          const uint32_t sub = ((uint32_t)(closed >> (dwin & 63u)) & 1u) | (((uint32_t)(closed >> ((dwin >> 6) & 63u)) & 1u) << 1) |
                               (((uint32_t)(closed >> ((dwin >> 12) & 63u)) & 1u) << 2) | (((uint32_t)(closed >> ((dwin >> 18) & 63u)) & 1u) << 3);
          vis = tb.vis_tab[(size_t)tile_id * 16 + (sub & ((1u << (dwin >> 24)) - 1u))];
        } else {
          unsigned long long B = W49;
          for (unsigned long long m = dnear & ~dopen; m; m &= m - 1) {
            const uint16_t q = tb.door_pos[__ffsll((long long)m) - 1];
            const int dx = px(q) - ax + R, dy = py(q) - ay + R;
            if ((unsigned)dx < (unsigned)D && (unsigned)dy < (unsigned)D) B |= 1ull << (dx * D + dy);
          }
          vis = march<R>(B);
        }
        wv = W49 & vis;

        // ---- one pass over the listed entities: which are visible inside the window (`*_w` masks), and (faithful
        // mode) which uids have two candidates.  Among the visible LISTED entities that share a uid
        // only the one the full radius-D rays visit first survives `set(visible_entities)`.  An entity can only take
        // part if it is visible: exact `vis` inside the window, the static walls-only visibility (a superset) on the
        // ring between window and radius D.  Ranks are derived only for the uids that really have two candidates.
        constexpr int BW = 2 * D + 1;
        if (FAITHFUL) {
          // candidate mask of the radius-D box: static walls-only visibility on the ring, the exact `vis` inside the window
          unsigned long long cm[4] = {sb0, sb1, sb2, sb3};
#pragma unroll
          for (int r = 0; r < D; ++r) {
            constexpr unsigned long long ROW = (1ull << D) - 1ull;
            const int pos = (r + D - R) * BW + (D - R), w = pos >> 6, sh = pos & 63;      // compile-time after unrolling
            const unsigned long long row = (vis >> (r * D)) & ROW;
            cm[w] = (cm[w] & ~(ROW << sh)) | (row << sh);
            if (sh + D > 64) cm[w + 1] = (cm[w + 1] & ~(ROW >> (64 - sh))) | (row >> (64 - sh));
          }
#pragma unroll
          for (int w = 0; w < 4; ++w) { s_cm[(2 * w) * 32] = (uint32_t)cm[w]; s_cm[(2 * w + 1) * 32] = (uint32_t)(cm[w] >> 32); }
        }
        // returns 0 = not a candidate, 1 = candidate on the ring, 3 = visible inside the window (branch-free)
        auto classify = [&](uint16_t q) -> int {
          const int bx = px(q) - ax, by = py(q) - ay;
          const bool inwin = (unsigned)(bx + R) < (unsigned)D && (unsigned)(by + R) < (unsigned)D;
          if (!FAITHFUL) return (inwin && ((vis >> ((bx + R) * D + by + R)) & 1ull)) ? 3 : 0;
          const bool inbox = (unsigned)(bx + D) < (unsigned)BW && (unsigned)(by + D) < (unsigned)BW;
          const int ci = inbox ? (bx + D) * BW + by + D : 0;
          const uint32_t bit = (s_cm[(ci >> 5) * 32] >> (ci & 31)) & 1u;
          return inbox ? (int)(bit | ((inwin ? bit : 0u) << 1)) : 0;
        };
        unsigned long long door_w = 0ull, dirt_w = 0ull, grp_w = 0ull;
        const unsigned long long wc64 = FAITHFUL ? tb.wall_cand64[tile_id] : 0ull;   // walls with uid < 64 that may be visible
        // uids seen so far / seen twice / seen inside the window.  A shared uid only matters when one of its holders is
        // inside the window (nothing outside the window is drawn), which prunes most ring-ring pairs.
        unsigned long long seen = wc64, dup = 0ull, win_uids = FAITHFUL ? tb.wall_win64[tile_id] : 0ull;
        for (unsigned long long m = dnear; m; m &= m - 1) {
          const int d = __ffsll((long long)m) - 1;
          const int c = classify(tb.door_pos[d]);
          const unsigned long long bit = c ? 1ull << d : 0ull, wbit = c == 3 ? bit : 0ull;
          dup |= seen & bit; seen |= bit;
          door_w |= wbit; win_uids |= wbit;
        }
        {
          // small groups hold at most 32 members: 32-bit uid masks, merged below
          uint32_t seen32 = (uint32_t)seen, dup32 = 0u, win32 = 0u;
#pragma unroll
          for (int g = 0; g < 6; ++g)
            for (int s = lo[g]; s < hi[g]; ++s) {
              const uint16_t q = pos[s];
              const int c = (q != NO_POS && ((glisted[g] >> (s - lo[g])) & 1)) ? classify(q) : 0;
              const uint32_t bit = c ? 1u << (s - lo[g]) : 0u;
              dup32 |= seen32 & bit; seen32 |= bit;
              if (c == 3) { grp_w |= 1ull << (s - sl.item0); win32 |= bit; }
            }
          seen |= seen32; dup |= dup32; win_uids |= win32;
        }
        // dirt piles last (their uids are unique among dirt piles): conflicting ones go to a 12-entry packed list
        unsigned long long dlist = 0ull, dlist2 = 0ull, dlist3 = 0ull;          // 12 packed entries: uid | slot << 10
        int ndl = 0;
        bool overflow = false;
        {
          const uint32_t wrng = FAITHFUL ? tb.wall_cand_rng[tile_id] : 0xFFFFu;     // [lo, hi] of the candidate wall uids >= 64
          // listed piles (identity mode: the live ones); the 64-bit slot mask is walked as two 32-bit words
#pragma unroll
          for (int half = 0; half < 2; ++half) {
          // branch-free walk: bit 1 of the class = visible inside the window, bit 0 = candidate (window or ring)
          uint32_t cand = 0u, dw = 0u;
          for (uint32_t dm = half ? (uint32_t)(dirtlisted >> 32) : (uint32_t)dirtlisted; dm; dm &= dm - 1) {
            const int kk = __ffs(dm) - 1;
            const uint32_t c = (uint32_t)classify(pos[kk + 32 * half]);
            dw |= ((c >> 1) & 1u) << kk;
            cand |= (c & 1u) << kk;
          }
          dirt_w |= (unsigned long long)dw << (32 * half);
          // the visible piles' f64 amounts are read in phase 2 (HBM, uncoalesced): start the fetches now
          for (uint32_t m = dw; m; m &= m - 1)
            asm volatile("prefetch.global.L2 [%0];\n" ::"l"(&field_at(st, st.dirt_amt, __ffs(m) - 1 + 32 * half, e)));
          // faithful: the uid bookkeeping of the candidates runs as a second, short loop over the lane's candidate slots only
          // (~1 of ~20 listed piles per lane) instead of as a divergent tail of every iteration of the walk above.  Pile
          // uids are unique among the piles, so the order of the two loops does not matter to `seen / win_uids`.
          if (FAITHFUL)
          for (uint32_t cm2 = cand; cm2; cm2 &= cm2 - 1) {
            const int k = __ffs(cm2) - 1 + 32 * half;
            const int c = ((dirt_w >> k) & 1ull) ? 3 : 1;
            {
              const uint32_t uid = blk_dirt_uid[k * ENV_BLOCK + eb];
              bool cf;
              if (uid < 64) {
                // another holder of the uid exists, and one of the two is inside the window (the final pruning, applied
                // before the pile takes a slot of the packed list: on small levels most piles share a uid with some wall)
                cf = ((seen >> uid) & 1ull) != 0 && (c == 3 || ((win_uids >> uid) & 1ull) != 0);
                if (c == 3) win_uids |= 1ull << uid;
              } else {
                cf = uid >= (wrng & 0xFFFFu) && uid <= (wrng >> 16);
                if (cf && c != 3) {                 // pile on the ring: only matters if the wall of that uid is inside the window
                  const uint16_t w = tb.wall_pos[uid];
                  const int wx = px(w) - ax, wy = py(w) - ay;
                  cf = wx >= -R && wx <= R && wy >= -R && wy <= R;
                }
              }
              if (cf) {
                if (uid < 64) dup |= 1ull << uid;
                if (ndl < 12 && uid < 1024) {
                  const unsigned long long en = (unsigned long long)(uid | ((uint32_t)k << 10)) << (16 * (ndl & 3));
                  if (ndl < 4) dlist |= en; else if (ndl < 8) dlist2 |= en; else dlist3 |= en;
                  ++ndl;
                }
                else overflow = true;
              }
            }
          }
          }
        }
        if (FAITHFUL) {
          if (overflow) {
            atomicAdd(s_cnt + el, cap + 1);           // too many conflicts: the exact per-agent path redoes this env
          } else if ((dup &= win_uids) != 0ull || ndl != 0) {
            // true light-block mask of the radius-D box: walls from the per-tile table + closed listed doors
            const unsigned long long* wb = reinterpret_cast<const unsigned long long*>(tb.wall_box) + (size_t)tile_id * 4;
            unsigned long long b0 = wb[0], b1 = wb[1], b2 = wb[2], b3 = wb[3];
            for (unsigned long long m = dnear & ~dopen; m; m &= m - 1) {
              const uint16_t q = tb.door_pos[__ffsll((long long)m) - 1];
              const int bi = (px(q) - ax + D) * BW + (py(q) - ay + D), w = bi >> 6;
              const unsigned long long bit = 1ull << (bi & 63);
              b0 |= w == 0 ? bit : 0ull; b1 |= w == 1 ? bit : 0ull; b2 |= w == 2 ? bit : 0ull; b3 |= w == 3 ? bit : 0ull;
            }
            auto rk = [&](uint16_t q) -> int { return first_visit_rank<R>(b0, b1, b2, b3, px(q) - ax, py(q) - ay); };
            auto drop_wall = [&](uint16_t w) {
              const int wx = px(w) - ax + R, wy = py(w) - ay + R;
              if ((unsigned)wx < (unsigned)D && (unsigned)wy < (unsigned)D) wv &= ~(1ull << (wx * D + wy));
            };
            constexpr int INF = 0x7FFF;
            for (unsigned long long m = dup; m; m &= m - 1) {
              const int u = __ffsll((long long)m) - 1;
              const bool has_wall = u < n_walls && ((wc64 >> u) & 1ull);
              const int r_wall = has_wall ? rk(tb.wall_pos[u]) : INF;
              const int r_door = (u < n_doors && ((dlisted >> u) & 1)) ? rk(tb.door_pos[u]) : INF;
              int r_g[6];
#pragma unroll
              for (int g = 0; g < 6; ++g) {
                r_g[g] = INF;
                if (u < hi[g] - lo[g] && ((glisted[g] >> u) & 1)) {
                  const uint16_t q = pos[lo[g] + u];
                  if (q != NO_POS) r_g[g] = rk(q);
                }
              }
              int r_dirt = INF, k_dirt = -1;
              for (int i = 0; i < ndl; ++i) {
                const uint32_t en = (uint32_t)((i < 4 ? dlist : i < 8 ? dlist2 : dlist3) >> (16 * (i & 3))) & 0xFFFFu;
                if ((int)(en & 1023u) == u) { k_dirt = (int)(en >> 10); r_dirt = rk(pos[k_dirt]); }
              }
              int best = r_wall < r_door ? r_wall : r_door;
              best = r_dirt < best ? r_dirt : best;
#pragma unroll
              for (int g = 0; g < 6; ++g) best = r_g[g] < best ? r_g[g] : best;
              if (has_wall && r_wall > best) drop_wall(tb.wall_pos[u]);
              if (r_door > best) door_w &= ~(1ull << u);
              if (k_dirt >= 0 && r_dirt > best) dirt_w &= ~(1ull << k_dirt);
#pragma unroll
              for (int g = 0; g < 6; ++g)
                if (r_g[g] > best && u < hi[g] - lo[g]) grp_w &= ~(1ull << (lo[g] + u - sl.item0));
            }
            // dirt piles with uid >= 64 can only meet the wall of that uid
            for (int i = 0; i < ndl; ++i) {
              const uint32_t en = (uint32_t)((i < 4 ? dlist : i < 8 ? dlist2 : dlist3) >> (16 * (i & 3))) & 0xFFFFu;
              const int uid = (int)(en & 1023u), k = (int)(en >> 10);
              if (uid < 64 || uid >= n_walls) continue;
              const uint16_t w = tb.wall_pos[uid];
              const int rw = rk(w), rd = rk(pos[k]);
              if (rw < rd) dirt_w &= ~(1ull << k);
              else if (rd < rw) drop_wall(w);
            }
          }
        }

        // ---- emission: every surviving visible entity becomes one sprite per channel that shows its group
        const uint32_t* chm = s_chm + a * MFG_N_TERMS;
        const int coff = s_coff[a];
        Sprite* spr = s_spr + (size_t)el * cap;
        int* cnt = s_cnt + el;
        auto put = [&](uint32_t w, float val) {      // the A agent-lanes of an env append to one list
          const int slot = atomicAdd(cnt, 1);
          if (slot < cap) spr[slot] = Sprite{w, val};
        };
        auto emit = [&](uint32_t m, int cell, uint32_t kind, uint32_t aux, float val) {
          while (m) {
            const int c = __ffs(m) - 1;
            m &= m - 1;
            put((uint32_t)((coff + c) * DD + cell) | (kind << 16) | (aux << 24), val);
          }
        };
        auto wcell = [&](uint16_t q) -> int { return (px(q) - ax + R) * D + (py(q) - ay + R); };   // q is inside the window
        // agents (each agent plane is 1.0 at the agent's cell; stacks add up in Combined planes)
        for (int j = 0; j < A; ++j) {
          const uint32_t m = chm[MFG_G_AGENT0 + j];
          if (!m) continue;
          const uint16_t q = pos[sl.agent0 + j];
          const int bx = px(q) - ax + R, by = py(q) - ay + R;          // agents have string ids: only window visibility matters
          if ((unsigned)bx < (unsigned)D && (unsigned)by < (unsigned)D && ((vis >> (bx * D + by)) & 1ull)) emit(m, bx * D + by, SK_INT, 0, 1.0f);
        }
        // small groups
        {
          const int term[6] = {MFG_G_ITEMS, MFG_G_PODS, MFG_G_DEST, MFG_G_DROPOFF, MFG_G_MACHINES, MFG_G_MAINT};
#pragma unroll
          for (int g = 0; g < 6; ++g) {
            const uint32_t cm = chm[term[g]];
            if (!cm || hi[g] == lo[g]) continue;
            uint32_t m = (uint32_t)(grp_w >> (lo[g] - sl.item0)) & (uint32_t)((1ull << (hi[g] - lo[g])) - 1ull);
            if (g == 2) m &= ~reached;                                   // a reached destination encodes as 0
            for (; m; m &= m - 1) {
              const int s = lo[g] + __ffs(m) - 1;
              emit(cm, wcell(pos[s]), SK_INT, 0, g == 4 ? (float)ENC_MACHINE : 1.0f);
            }
          }
        }
        // doors
        if (chm[MFG_G_DOORS]) {
          for (unsigned long long m = door_w; m; m &= m - 1) {
            const int d = __ffsll((long long)m) - 1;
            emit(chm[MFG_G_DOORS], wcell(tb.door_pos[d]), SK_DOOR, (uint32_t)((dopen >> d) & 1), 0.f);
          }
        }
        // dirt piles
        if (chm[MFG_G_DIRT]) {
          for (unsigned long long m = dirt_w; m; m &= m - 1) {
            const int k = __ffsll((long long)m) - 1;
            emit(chm[MFG_G_DIRT], wcell(pos[k]), SK_DIRT, (uint32_t)k, 0.f);
          }
        }
        // the battery level (f64 in HBM) of this (env, agent) lane: Battery channel, written in phase 2
        if (s_hasbat[a]) batv = (float)field_at(st, st.bat, a, e);
        lane_on = true;
        if (FAITHFUL) { s_cm[0] = (uint32_t)wv; s_cm[32] = (uint32_t)(wv >> 32); s_cm[64] = __float_as_uint(batv); s_cm[96] = axy; }
      }
    }
    __syncwarp();
    // ---------------- phase 2: the warp composes its EPW tiles in shared memory, `ppp` planes at a time ---------------
    const uint32_t wv_lo = (uint32_t)wv, wv_hi = (uint32_t)(wv >> 32);
    for (int elx = 0; elx < EPW; ++elx) {
      const int eb = sub * EPW + elx;
      if (eb >= n_live) break;
      if ((skip_mask >> (elx << apad_log2)) & 1u) continue;     // being re-spawned concurrently: the list-mode launch writes it
      const int64_t e = env_of(eb);
      float* te = obs + (size_t)e * tile_floats;
      const int cnt = s_cnt[elx];
      if (cnt > cap) {                      // sprite list overflowed (or too many uid conflicts): k_obs_redo rewrites this env
        if (lane == 0) redo[1 + atomicAdd(redo, 1u)] = (uint32_t)(e);
        continue;
      }
      // sprites.  Value of a cell = (float)((double)integer stack + fractional encoding): the integer-valued sprites of a cell
      // add up exactly, a door / dirt encoding is added last with one rounding each (like the reference's f64 sums).
      // One sprite per lane (cnt <= 32): lanes that hit the same cell are grouped by a warp match and the group's first lane
      // carries the final value; the dirt amounts (f64, HBM, uncoalesced) are in flight while the first part is cleared.
      uint32_t idx = 0xFFFFFFFFu;           // cell this lane writes (tile-relative), none by default
      float out = 0.f;
      double d0 = 0.0;
      Sprite s0{0u, 0.f};
      uint32_t k0 = 0xFF;
      const bool on = lane < cnt && cnt <= 32;
      if (on) {
        s0 = s_spr[(size_t)elx * cap + lane];
        k0 = (s0.w >> 16) & 0xFF;
        if (k0 == SK_DIRT) d0 = field_at(st, st.dirt_amt, (int)(s0.w >> 24), e);
      }
      const uint32_t act = __ballot_sync(0xffffffffu, on);
      // walls: lane L handles one third (17 / 17 / 15 window cells) of wall plane L / 3 and sets its cells one by one - a
      // handful of iterations for all planes at once (lane = window cell took two predicated stores per plane)
      const int src0 = elx << apad_log2;
      const uint32_t* s_res = s_cm - lane;                     // faithful: [8][32], words 0 / 1 = wall mask, 2 = battery level, 3 = position
      int w_pl = -1, w_base = 0;                                // this lane's wall plane (packed channel index) / first cell
      uint32_t w_bits = 0u;                                     // visible walls among the cells [w_base, w_base + 17)
      {
        const int j = lane / 3, t = lane - 3 * j;
        const uint32_t rec = j < n_wp ? s_wplane[j] : 0u;
        unsigned long long m;
        if (FAITHFUL) m = (unsigned long long)s_res[src0 + (int)(rec >> 10)] | ((unsigned long long)s_res[32 + src0 + (int)(rec >> 10)] << 32);
        else m = (unsigned long long)__shfl_sync(0xffffffffu, wv_lo, src0 + (int)(rec >> 10)) |
                 ((unsigned long long)__shfl_sync(0xffffffffu, wv_hi, src0 + (int)(rec >> 10)) << 32);
        if (j < n_wp) { w_pl = (int)(rec & 1023u); w_base = 17 * t; w_bits = (uint32_t)(m >> w_base) & 0x1FFFFu; }
      }
      // scalar channels (observation_builder.py:205-218): lane i holds entry i of the list: battery level / (x / H, y / W) at
      // the first cells of the plane; the values sit in the (env, agent) lanes of phase 1
      int sc_pl = -1;
      float sc_v0 = 0.f, sc_v1 = 0.f;
      bool sc_two = false;
      if (n_scl) {
        const uint32_t rec = lane < n_scl ? s_scl[lane] : 0u;
        float bv;
        uint32_t pp;
        if (FAITHFUL) {
          bv = __uint_as_float(s_res[64 + src0 + (int)(rec >> 16)]); pp = s_res[96 + src0 + (int)(rec >> 16)];
        } else {
          bv = __shfl_sync(0xffffffffu, batv, src0 + (int)(rec >> 16));
          pp = __shfl_sync(0xffffffffu, axy, src0 + (int)(rec >> 16));
        }
        const float gx = s_gx[pp >> 8], gy = s_gy[pp & 255u];
        sc_two = ((rec >> 12) & 15u) != MFG_CH_BATTERY;
        sc_v0 = sc_two ? gx : bv;
        sc_v1 = gy;
        sc_pl = lane < n_scl ? (int)(rec & 4095u) : -1;
      }
      bool resolved = false;
      int f_lo = 0;
      for (int p0 = 0; p0 < total_channels; p0 += ppp, f_lo += part_floats) {
        const int p1 = p0 + ppp < total_channels ? p0 + ppp : total_channels;
        const int nfl = (p1 - p0) * DD;
        if (bulk) {      // the bulk store that last read the part buffer must have finished reading shared memory
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
          __syncwarp();
        }
        {
          float4* t4 = reinterpret_cast<float4*>(tile);
          const int n4 = (nfl + 3) >> 2;
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          int i = lane;
#pragma unroll 1
          for (; i + 96 < n4; i += 128) { t4[i] = z; t4[i + 32] = z; t4[i + 64] = z; t4[i + 96] = z; }
#pragma unroll 1
          for (; i < n4; i += 32) t4[i] = z;
        }
        if (!resolved) {       // (after the clear has been issued: the dirt amounts had time to arrive)
          resolved = true;
          if (on) {
            idx = s0.w & 0xFFFFu;
            const uint32_t grp = __match_any_sync(act, idx);
            if (grp == (1u << lane)) {
              out = k0 == SK_INT ? s0.val : k0 == SK_DOOR ? (float)(0.0 + ((s0.w >> 24) ? ENC_DOOR_OPEN : ENC_DOOR_CLOSED)) : (float)(0.0 + d0);
            } else {
              const int isum = __reduce_add_sync(grp, k0 == SK_INT ? (int)s0.val : 0);
              const uint32_t doors = __ballot_sync(grp, k0 == SK_DOOR) & grp, dirts = __ballot_sync(grp, k0 == SK_DIRT) & grp;
              const uint32_t w_door = __shfl_sync(grp, s0.w, doors ? __ffs(doors) - 1 : lane);
              const double d_dirt = __shfl_sync(grp, d0, dirts ? __ffs(dirts) - 1 : lane);
              float f = (float)isum;
              if (doors) f = (float)((double)f + ((w_door >> 24) ? ENC_DOOR_OPEN : ENC_DOOR_CLOSED));
              if (dirts) f = (float)((double)f + d_dirt);
              out = f;
              if ((uint32_t)(__ffs(grp) - 1) != (uint32_t)lane) idx = 0xFFFFFFFFu;       // the group's first lane writes
            }
          }
        }
        __syncwarp();
        // wall planes of this part: 1.0 where a visible wall is.  Nothing else can be on a wall cell, so every store has a
        // unique writer.
        if ((unsigned)(w_pl - p0) < (unsigned)(p1 - p0)) {
          float* cells = tile + (w_pl - p0) * DD + w_base;
          for (uint32_t m = w_bits; m; m &= m - 1) cells[__ffs(m) - 1] = 1.0f;
        }
        for (int base = 32; base < 3 * n_wp; base += 32) {        // more than 10 wall planes: the remaining ones, round by round
          const int L = base + lane, j = L / 3, t = L - 3 * j;
          const bool valid = L < 3 * n_wp;
          const uint32_t rec = valid ? s_wplane[j] : 0u;
          unsigned long long m;
          if (FAITHFUL) m = (unsigned long long)s_res[src0 + (int)(rec >> 10)] | ((unsigned long long)s_res[32 + src0 + (int)(rec >> 10)] << 32);
          else m = (unsigned long long)__shfl_sync(0xffffffffu, wv_lo, src0 + (int)(rec >> 10)) |
                   ((unsigned long long)__shfl_sync(0xffffffffu, wv_hi, src0 + (int)(rec >> 10)) << 32);
          const int pl = (int)(rec & 1023u);
          if (valid && (unsigned)(pl - p0) < (unsigned)(p1 - p0)) {
            float* cells = tile + (pl - p0) * DD + 17 * t;
            for (uint32_t b = (uint32_t)(m >> (17 * t)) & 0x1FFFFu; b; b &= b - 1) cells[__ffs(b) - 1] = 1.0f;
          }
        }
        if ((unsigned)(sc_pl - p0) < (unsigned)(p1 - p0)) {
          float* cells = tile + (sc_pl - p0) * DD;
          cells[0] = sc_v0;
          if (sc_two) cells[1] = sc_v1;
        }
        if (n_scl > 32) {           // more scalar channels than lanes (many agents): the remaining entries, round by round
          for (int base = 32; base < n_scl; base += 32) {
            const uint32_t rec = base + lane < n_scl ? s_scl[base + lane] : 0u;
            float bv;
            uint32_t pp;
            if (FAITHFUL) {
              bv = __uint_as_float(s_res[64 + src0 + (int)(rec >> 16)]); pp = s_res[96 + src0 + (int)(rec >> 16)];
            } else {
              bv = __shfl_sync(0xffffffffu, batv, src0 + (int)(rec >> 16));
              pp = __shfl_sync(0xffffffffu, axy, src0 + (int)(rec >> 16));
            }
            const int pl = (int)(rec & 4095u);
            if (base + lane < n_scl && (unsigned)(pl - p0) < (unsigned)(p1 - p0)) {
              float* cells = tile + (pl - p0) * DD;
              if (((rec >> 12) & 15u) == MFG_CH_BATTERY) cells[0] = bv;
              else { cells[0] = s_gx[pp >> 8]; cells[1] = s_gy[pp & 255u]; }
            }
          }
        }
        if (idx - (uint32_t)f_lo < (uint32_t)nfl) tile[idx - f_lo] = out;
        if (cnt > 32) {
          // rare (large sprite capacities only): rounds of 32 sprites, one pass per kind, read-modify-write in the part buffer
          // (ordered by the warp barriers); same arithmetic as above
          __syncwarp();
          for (int kind = SK_INT; kind <= SK_DIRT; ++kind) {
            if (kind == SK_STORE) continue;
            for (int base = 0; base < cnt; base += 32) {
              const int i = base + lane;
              Sprite s1{0u, 0.f};
              if (i < cnt) s1 = s_spr[(size_t)elx * cap + i];
              const uint32_t ix = s1.w & 0xFFFFu;
              const bool on1 = i < cnt && (int)((s1.w >> 16) & 0xFF) == kind && ix - (uint32_t)f_lo < (uint32_t)nfl;
              const uint32_t act1 = __ballot_sync(0xffffffffu, on1);
              if (on1) {
                const uint32_t grp = __match_any_sync(act1, ix);
                const float cur = tile[ix - f_lo];
                float f;
                if (kind == SK_INT) f = cur + (float)__reduce_add_sync(grp, (int)s1.val);
                else if (kind == SK_DOOR) f = (float)((double)cur + ((s1.w >> 24) ? ENC_DOOR_OPEN : ENC_DOOR_CLOSED));
                else f = (float)((double)cur + field_at(st, st.dirt_amt, (int)(s1.w >> 24), e));
                if ((uint32_t)(__ffs(grp) - 1) == (uint32_t)lane) tile[ix - f_lo] = f;
              }
              __syncwarp();
            }
          }
        }
        // ---- stream the part out
        if (bulk) {
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to the async proxy
          __syncwarp();
          if (lane == 0) bulk_store_tile(te + f_lo, tile, (uint32_t)(nfl * sizeof(float)));
        } else {
          __syncwarp();
          for (int i = lane; i < nfl; i += 32) te[f_lo + i] = tile[i];
          __syncwarp();
        }
      }
    }
    __syncwarp();
  }
  if (!lmode) break;
  __syncthreads();                   // every warp is done with the image before the next chunk is staged
  }
  if (bulk && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // smem must outlive the copies
}

// ---------------------------------------------------------------------------------------------------------------
// host side: planning + launch
// ---------------------------------------------------------------------------------------------------------------
namespace mfg {

static void build_window_rays(const MfgSpec& sp, WindowRays& wr) {
  const int r = sp.pomdp_r, D = 2 * r + 1;
  std::vector<std::vector<std::pair<int, int>>> kept;
  for (int k = 0; k < sp.n_rays; ++k) {
    std::vector<std::pair<int, int>> pref;
    for (int s = 1; s < sp.ray_len[k]; ++s) {
      int dx = sp.ray_dx[k][s], dy = sp.ray_dy[k][s];
      if (dx < -r || dx > r || dy < -r || dy > r) break;
      pref.emplace_back(dx, dy);
    }
    if (pref.empty()) continue;
    bool dup = false;
    for (auto& q : kept) dup |= q == pref;
    if (!dup) kept.push_back(pref);
  }
  std::vector<std::vector<std::pair<int, int>>> fin;
  for (auto& p : kept) {
    bool is_prefix = false;
    for (auto& q : kept)
      if (q.size() > p.size() && std::equal(p.begin(), p.end(), q.begin())) is_prefix = true;
    if (!is_prefix) fin.push_back(p);
  }
  memset(&wr, 0, sizeof(wr));
  wr.n = (int)fin.size();
  for (int k = 0; k < wr.n && k < MAX_WRAYS; ++k) {
    wr.len[k] = (int)fin[k].size() > MAX_WLEN ? MAX_WLEN : (int)fin[k].size();
    int px_ = 0, py_ = 0;
    for (int s = 0; s < wr.len[k]; ++s) {
      int dx = fin[k][s].first, dy = fin[k][s].second;
      int cx = dx - px_, cy = dy - py_;
      wr.cell[k][s] = (uint8_t)((dx + r) * D + (dy + r));
      if (cx != 0 && cy != 0) {
        wr.da[k][s] = (uint8_t)((dx + r) * D + (dy - cy + r));
        wr.db[k][s] = (uint8_t)((dx - cx + r) * D + (dy + r));
      } else {
        wr.da[k][s] = wr.db[k][s] = 255;
      }
      px_ = dx; py_ = dy;
    }
  }
}

// the generated constexpr trie must describe exactly the rays derived from the spec's full ray table
template <int R>
static bool trie_matches(const WindowRays& wr) {
  using T = RayTrie<R>;
  // every run-time ray must be a root-to-node path of the trie with identical cells / diagonal neighbours ...
  int used[T::N] = {0};
  for (int k = 0; k < wr.n; ++k) {
    int parent = -1;
    for (int s = 0; s < wr.len[k]; ++s) {
      int found = -1;
      for (int n = 0; n < T::N; ++n)
        if (T::parent(n) == parent && T::cell(n) == wr.cell[k][s] && T::da(n) == wr.da[k][s] && T::db(n) == wr.db[k][s]) found = n;
      if (found < 0) return false;
      used[found] = 1;
      parent = found;
    }
  }
  // ... and the trie must not contain anything else
  for (int n = 0; n < T::N; ++n) if (!used[n]) return false;
  return wr.n > 0;
}

// the generated full-radius trie must describe exactly the spec's rays (origin cell excluded), in the same order
template <int R>
static bool full_trie_matches(const MfgSpec& sp) {
  using T = FullTrie<R>;
  const int D = 2 * R + 1, BW = 2 * D + 1;
  std::vector<int> used(T::N, 0);
  int next_new = 0;
  for (int k = 0; k < sp.n_rays; ++k) {
    int parent = -1, pxo = 0, pyo = 0;
    for (int s = 1; s < sp.ray_len[k]; ++s) {
      const int dx = sp.ray_dx[k][s], dy = sp.ray_dy[k][s];
      if (dx < -D || dx > D || dy < -D || dy > D) return false;
      const int cx = dx - pxo, cy = dy - pyo;
      const int cell = (dx + D) * BW + (dy + D);
      const int da = (cx && cy) ? (dx + D) * BW + (dy - cy + D) : 255, db = (cx && cy) ? (dx - cx + D) * BW + (dy + D) : 255;
      const int wc = (dx >= -R && dx <= R && dy >= -R && dy <= R) ? (dx + R) * D + (dy + R) : -1;
      int found = -1;
      for (int n = 0; n < T::N; ++n)
        if (T::parent(n) == parent && T::cell(n) == cell && T::da(n) == da && T::db(n) == db && T::wcell(n) == wc) found = n;
      if (found < 0) return false;
      if (!used[found]) { if (found != next_new) return false; ++next_new; }     // node ids follow first-visit order
      used[found] = 1;
      parent = found; pxo = dx; pyo = dy;
    }
  }
  for (int n = 0; n < T::N; ++n) {
    if (!used[n]) return false;
    bool listed = false;
    for (int i = T::cell_off(T::cell(n)); i < T::cell_off(T::cell(n) + 1); ++i) listed |= T::cell_node(i) == n;
    if (!listed) return false;
  }
  return true;
}

// static (walls-only) visibility tables of the faithful mode: per tile, the box cells reached by the full rays when only
// walls block light, and which wall uids lie on such cells (walls are static, so this part of the conflict filter is free)
template <int R>
static void build_vis_tables_r(const MfgSpec& sp, HostTables& t) {
  using T = FullTrie<R>;
  constexpr int D = 2 * R + 1, BW = 2 * D + 1;
  const int H = sp.H, W = sp.W;
  t.vis_box.assign((size_t)H * W * 4, 0);
  t.wall_cand64.assign((size_t)H * W, 0);
  t.wall_cand_rng.assign((size_t)H * W, 0x0000FFFFu);
  t.wall_win64.assign((size_t)H * W, 0);
  std::vector<char> cont(T::N);
  for (int x = 0; x < H; ++x)
    for (int y = 0; y < W; ++y) {
      uint64_t* vb = &t.vis_box[((size_t)x * W + y) * 4];
      auto blk = [&](int ci) -> bool {
        const int xx = x + ci / BW - D, yy = y + ci % BW - D;
        return xx >= 0 && yy >= 0 && xx < H && yy < W && t.wall[(size_t)xx * W + yy];
      };
      vb[(D * BW + D) >> 6] |= 1ull << ((D * BW + D) & 63);
      if (!t.wall[(size_t)x * W + y]) {
        for (int n = 0; n < T::N; ++n) {          // parents precede their children (node ids = first-visit order)
          const int p = T::parent(n), c = T::cell(n), da = T::da(n), db = T::db(n);
          const bool reach = p < 0 ? true : cont[p] != 0;
          const bool diag = da != 255 && blk(da) && blk(db);
          if (reach && !diag) vb[c >> 6] |= 1ull << (c & 63);
          cont[n] = reach && !diag && !blk(c);
        }
      }
      uint32_t lo = 0xFFFF, hi = 0;
      for (int dx = -D; dx <= D; ++dx)
        for (int dy = -D; dy <= D; ++dy) {
          const int xx = x + dx, yy = y + dy, ci = (dx + D) * BW + (dy + D);
          if (xx < 0 || yy < 0 || xx >= H || yy >= W || !t.wall[(size_t)xx * W + yy] || !((vb[ci >> 6] >> (ci & 63)) & 1)) continue;
          const uint32_t u = t.wall_uid[(size_t)xx * W + yy];
          if (u < 64) {
            t.wall_cand64[(size_t)x * W + y] |= 1ull << u;
            if (dx >= -R && dx <= R && dy >= -R && dy <= R) t.wall_win64[(size_t)x * W + y] |= 1ull << u;
          }
          else { lo = u < lo ? u : lo; hi = u > hi ? u : hi; }
        }
      t.wall_cand_rng[(size_t)x * W + y] = lo | (hi << 16);
    }
}

void build_vis_tables(const MfgSpec& sp, HostTables& t) {
  if (!sp.faithful) return;
  switch (sp.pomdp_r) {
    case 1: build_vis_tables_r<1>(sp, t); break;
    case 2: build_vis_tables_r<2>(sp, t); break;
    case 3: build_vis_tables_r<3>(sp, t); break;
    default: break;
  }
}

// window visibility table: vis_tab[tile][subset] = march(wall mask of the window | closed doors of the subset), for the (at most
// 4) doors inside the tile's window; door_win[tile] names them.  Replaces the per-agent ray march of the tiled kernel by one
// 8-byte load (the march stays as the fall-back for tiles whose window holds more than 4 doors).
template <int R>
static uint64_t march_host(uint64_t B) {
  using T = RayTrie<R>;
  constexpr int D = 2 * R + 1, centre = R * D + R;
  uint64_t vis = 1ull << centre, cont = 0;
  if (B & (1ull << centre)) return vis;
  for (int n = 0; n < T::N; ++n) {
    const int p = T::parent(n), c = T::cell(n), da = T::da(n), db = T::db(n);
    const bool reach = p < 0 ? true : ((cont >> p) & 1ull) != 0;
    const bool hits = (B >> c) & 1ull;
    const bool diag = da != 255 && ((B >> da) & 1ull) && ((B >> db) & 1ull);
    if (reach && !diag) vis |= 1ull << c;
    if (reach && !hits && !diag) cont |= 1ull << n;
  }
  return vis;
}
template <int R>
static void build_win_vis_r(const MfgSpec& sp, HostTables& t) {
  constexpr int D = 2 * R + 1;
  const int H = sp.H, W = sp.W;
  for (int x = 0; x < H; ++x)
    for (int y = 0; y < W; ++y) {
      const size_t tile = (size_t)x * W + y;
      int idx[4], cell[4], n = 0;
      bool many = false;
      for (int d = 0; d < sp.n_doors; ++d) {
        const int dx = px(sp.door_pos[d]) - x + R, dy = py(sp.door_pos[d]) - y + R;
        if (dx < 0 || dy < 0 || dx >= D || dy >= D) continue;
        if (n == 4) { many = true; break; }
        idx[n] = d; cell[n] = dx * D + dy; ++n;
      }
      if (many) { t.door_win[tile] = 7u << 24; continue; }
      uint32_t rec = (uint32_t)n << 24;
      for (int k = 0; k < n; ++k) rec |= (uint32_t)idx[k] << (6 * k);
      t.door_win[tile] = rec;
      for (int sub = 0; sub < (1 << n); ++sub) {
        uint64_t B = t.wall_win[tile];
        for (int k = 0; k < n; ++k) if ((sub >> k) & 1) B |= 1ull << cell[k];
        t.vis_tab[tile * 16 + sub] = march_host<R>(B);
      }
    }
}
void build_win_vis_tables(const MfgSpec& sp, HostTables& t) {
  t.door_win.assign((size_t)sp.H * sp.W, 7u << 24);
  t.vis_tab.assign((size_t)sp.H * sp.W * 16, 0);
  switch (sp.pomdp_r) {
    case 1: build_win_vis_r<1>(sp, t); break;
    case 2: build_win_vis_r<2>(sp, t); break;
    case 3: build_win_vis_r<3>(sp, t); break;
    default: break;
  }
}

void plan_obs(MfgHandle* h) {
  const MfgSpec& sp = h->sp;
  ObsPlan& p = h->plan;
  ObsSlots& sl = p.slots;
  sl.dirt0 = 0;
  sl.item0 = sp.has_dirt ? sp.dirt_slots : 0;
  sl.pod0 = sl.item0 + sp.n_items;
  sl.dest0 = sl.pod0 + sp.n_pods;
  sl.drop0 = sl.dest0 + sp.n_dest;
  sl.mach0 = sl.drop0 + sp.n_dropoff;
  sl.maint0 = sl.mach0 + sp.n_machines;
  sl.agent0 = sl.maint0 + sp.n_maint;
  sl.total = sl.agent0 + sp.n_agents;
  // the slots are exactly the first rows of every state block (MFG_STATE_FIELDS order), followed by the bit masks
  {
    Layout L = compute_layout(sp, h->N);
    auto off_of = [&](const char* name) -> int {
      for (auto& f : L.fields) if (!strcmp(f.name, name)) return f.rows ? (int)f.offset : -1;
      return -1;
    };
    auto end_of = [&](const char* name) -> int {
      for (auto& f : L.fields) if (!strcmp(f.name, name)) return (int)(f.offset + (size_t)f.rows * ENV_BLOCK * f.elem_size);
      return 0;
    };
    sl.off_dopen = off_of("door_open");
    sl.off_reached = off_of("dest_reached");
    sl.off_door_listed = off_of("door_listed");
    sl.off_dirt_listed = off_of("dirt_listed");
    sl.off_dirt_uid = off_of("dirt_uid");
    const char* ln[6] = {"item_listed", "pod_listed", "dest_listed", "drop_listed", "mach_listed", "maint_listed"};
    for (int g = 0; g < 6; ++g) sl.off_listed[g] = off_of(ln[g]);
    sl.prefix_bytes = sp.faithful ? end_of("dirt_uid") : (sp.has_dirt ? end_of("dirt_listed") : end_of("dest_reached"));
    const bool order_ok = (!sp.has_dirt || off_of("dirt_pos") == 0) && off_of("apos") == sl.agent0 * ENV_BLOCK * 2 &&
                          end_of("apos") <= sl.prefix_bytes && end_of("door_open") <= sl.prefix_bytes;
    if (!order_ok || sl.prefix_bytes % 16) { p.ok = false; return; }
    for (int* o : {&sl.off_dopen, &sl.off_reached, &sl.off_door_listed, &sl.off_dirt_listed, &sl.off_dirt_uid})
      if (*o < 0) *o = 0;                       // absent fields are never dereferenced; keep the pointers in range
  }
  const int tcdd = h->total_channels * h->DD;
  p.ge = 1;
  p.apad_log2 = 0;
  while ((1 << p.apad_log2) < sp.n_agents) ++p.apad_log2;
  const int epw = 32 >> p.apad_log2;                            // envs per warp pass
  p.cap = 8 * sp.n_agents < 16 ? 16 : 8 * sp.n_agents;        // sprite slots per env (overflow => generic slow path)
  p.cap_max = p.cap;
  // the constant table the kernel copies into shared memory (ObsProg): everything its prologue used to derive from the spec
  const size_t prog_words = (size_t)ObsProg::CHM + (size_t)((sp.n_agents * MFG_N_TERMS + 3) & ~3) + (size_t)((sp.H + 3) & ~3) +
                            (size_t)((sp.W + 3) & ~3);
  {
    p.prog.assign(prog_words, 0u);
    uint32_t n_scl = 0;
    for (int a = 0; a < sp.n_agents; ++a) {
      p.prog[ObsProg::COFF + a] = (uint32_t)sp.ch_offset[a];
      int na = 0;
      for (int c = 0; c < sp.n_channels[a]; ++c) {
        const int kind = sp.ch_kind[a][c];
        if (kind == MFG_CH_BATTERY) p.prog[ObsProg::HASBAT + a] = 1u;
        if ((kind == MFG_CH_BATTERY || kind == MFG_CH_GLOBALPOS) && na < 4) {      // scalar channels: packed plane | kind << 12 | agent << 16
          p.prog[ObsProg::SCL + n_scl++] = (uint32_t)(sp.ch_offset[a] + c) | ((uint32_t)kind << 12) | ((uint32_t)a << 16);
          ++na;
        }
      }
      for (int t = 0; t < MFG_N_TERMS; ++t) p.prog[ObsProg::CHM + a * MFG_N_TERMS + t] = sp.term_chmask[a][t];
    }
    p.prog[ObsProg::N_SCL] = n_scl;
    float* gx = reinterpret_cast<float*>(&p.prog[ObsProg::CHM + ((sp.n_agents * MFG_N_TERMS + 3) & ~3)]);
    float* gy = gx + ((sp.H + 3) & ~3);
    for (int i = 0; i < sp.H; ++i) gx[i] = (float)((double)i / (double)sp.H);      // entity/util.py:56-66
    for (int i = 0; i < sp.W; ++i) gy[i] = (float)((double)i / (double)sp.W);
  }
  // bulk (TMA) stores need parts that are whole numbers of 16-byte vectors: 4 planes of (2r+1)^2 floats are, so the tile is
  // composed in parts of `ppp` = 4k planes; with a channel count that is not a multiple of 4 the parts are copied out by the
  // warp itself
  p.bulk = (h->total_channels % 4 == 0) ? 1 : 0;
  auto smem_for = [&](int nw, int ppp) {
    size_t part = (((size_t)ppp * h->DD * 4) + 15) & ~(size_t)15;
    size_t per_warp = part + (size_t)epw * p.cap_max * 8 + (((size_t)epw * 4 + 15) & ~(size_t)15);   // part buffer, sprite lists, counters
    size_t misc = (sp.faithful ? (size_t)nw * 1024 : 0) + prog_words * 4;     // candidate masks, the constant table (ObsProg)
    return (size_t)sl.prefix_bytes + 32 + (size_t)nw * per_warp + misc;
  };
  // as many warps as there are sub-groups in a block, at most 8; resident CTAs per SM aimed at: 3 (faithful: 80 registers) or
  // 4 (identity: 64 registers) => (228 KB - 1 KB reserve per CTA) / CTAs, minus the static shared memory
  p.nw = ENV_BLOCK / epw < 8 ? ENV_BLOCK / epw : 8;
  const size_t budget = (size_t)233472 / (sp.faithful ? MFG_OBS_CTAS_F : 4) - 1024 - 512;
  p.ppp = (h->total_channels + 3) / 4 * 4;                     // whole tile first
  while (p.ppp > 4 && smem_for(p.nw, p.ppp) > budget) p.ppp -= 4;
  while (p.nw > 1 && smem_for(p.nw, p.ppp) > budget) p.nw >>= 1;
  if (p.ppp > h->total_channels) p.ppp = h->total_channels;
  p.smem = smem_for(p.nw, p.ppp);
  p.nbuf = 1;
  WindowRays wr;
  build_window_rays(sp, wr);
  bool trie_ok = sp.pomdp_r == 1 ? trie_matches<1>(wr) : sp.pomdp_r == 2 ? trie_matches<2>(wr)
               : sp.pomdp_r == 3 ? trie_matches<3>(wr) : false;
  // wall planes: (agent, packed channel) of every channel that contains Walls
  p.walls.n = 0;
  bool walls_fit = true;
  for (int a = 0; a < sp.n_agents; ++a)
    for (int c = 0; c < sp.n_channels[a]; ++c)
      if ((sp.term_chmask[a][MFG_G_WALLS] >> c) & 1) {
        if (p.walls.n >= MAX_WALL_PLANES) { walls_fit = false; continue; }
        p.walls.agent[p.walls.n] = (uint8_t)a;
        p.walls.plane[p.walls.n] = (uint16_t)(sp.ch_offset[a] + c);
        p.walls.n++;
      }
  p.prog[ObsProg::N_WP] = (uint32_t)p.walls.n;
  for (int w = 0; w < p.walls.n; ++w) p.prog[ObsProg::WPLANE + w] = (uint32_t)p.walls.plane[w] | ((uint32_t)p.walls.agent[w] << 10);
  bool full_ok = !sp.faithful || (sp.pomdp_r == 1 ? full_trie_matches<1>(sp) : sp.pomdp_r == 2 ? full_trie_matches<2>(sp)
                                  : sp.pomdp_r == 3 ? full_trie_matches<3>(sp) : false);
  if (sl.agent0 - sl.item0 > 64 || sp.n_walls > 0xFFFE) full_ok = false;   // 64-bit visibility masks
  p.ok = trie_ok && full_ok && walls_fit && p.smem <= 200 * 1024 && tcdd <= 0xFFFF;
}

// exact per-agent observation of the envs in a device-side list (k_obs_redo)
cudaError_t launch_obs_list(MfgHandle* h, float* d_obs, cudaStream_t s, const uint32_t* d_list, const uint32_t* d_count) {
  const unsigned blocks = (unsigned)((h->N + REDO_ENVS - 1) / REDO_ENVS);
  const unsigned rblocks = blocks < 96u ? blocks : 96u;
  const size_t smem = h->st.blk_i;
  const ColTab ct{h->d_row_tab, h->n_row_tab};
  cudaError_t err = cudaSuccess;
  if (h->sp.n_agents <= 4) {
    if (smem > 48 * 1024) err = cudaFuncSetAttribute(k_obs_redo<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err == cudaSuccess) k_obs_redo<4><<<rblocks, 128, smem, s>>>(h->d_sp, h->tb, h->st, ct, d_obs, h->total_channels, d_list, d_count);
  } else {
    if (smem > 48 * 1024) err = cudaFuncSetAttribute(k_obs_redo<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err == cudaSuccess) k_obs_redo<16><<<rblocks, 128, smem, s>>>(h->d_sp, h->tb, h->st, ct, d_obs, h->total_channels, d_list, d_count);
  }
  return err != cudaSuccess ? err : cudaGetLastError();
}

template <int R, bool FAITHFUL>
static cudaError_t launch_tiled_f(MfgHandle* h, float* d_obs, cudaStream_t s, const uint8_t* skip, const ObsList& ol) {
  auto kern = k_obs_tiled<R, FAITHFUL>;
  const ObsPlan& p = h->plan;
  if (p.smem > 40 * 1024) {         // (static shared memory counts against the 48 KB default limit too)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
  }
  if (!h->d_redo) {                  // two lists (block-mode launch, list-mode launch), each: [0] = count, [1..N] = env ids
    cudaError_t e = cudaMalloc(&h->d_redo, 2 * ((size_t)h->N + 1) * sizeof(uint32_t));
    if (e != cudaSuccess) return e;
  }
  if (!h->d_obs_prog) {
    cudaError_t e = cudaMalloc(&h->d_obs_prog, p.prog.size() * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_obs_prog, p.prog.data(), p.prog.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
  }
  uint32_t* redo = h->d_redo + (ol.ids ? (size_t)h->N + 1 : 0);
  cudaError_t e = cudaMemsetAsync(redo, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  unsigned blocks = (unsigned)((h->N + ENV_BLOCK - 1) / ENV_BLOCK);
  if (ol.ids && blocks > 592u) blocks = 592u;        // list mode: grid-stride over the list inside the kernel (idle CTAs exit at once)
  kern<<<blocks, p.nw * 32, p.smem, s>>>(h->d_sp, h->tb, h->st, p.slots, p.walls, d_obs, h->total_channels, p.cap, p.apad_log2,
                                         (h->obs_store != 0 && p.bulk) ? 1 : 0, p.ppp, redo, skip, ol, h->d_obs_prog, (int)p.prog.size());
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  return launch_obs_list(h, d_obs, s, redo + 1, redo);
}

template <int R>
static cudaError_t launch_tiled_r(MfgHandle* h, float* d_obs, cudaStream_t s, const uint8_t* skip, const ObsList& ol) {
  return h->sp.faithful ? launch_tiled_f<R, true>(h, d_obs, s, skip, ol) : launch_tiled_f<R, false>(h, d_obs, s, skip, ol);
}

// tiled observation of the envs in a device-side list (list mode of k_obs_tiled)
cudaError_t launch_obs_tiled_list(MfgHandle* h, float* d_obs, cudaStream_t s, const uint32_t* d_list, const uint32_t* d_count) {
  ObsList ol{d_list, d_count, h->d_row_tab, 0};
  // the staged prefix = the first rows of the block image
  for (int i = 0; i < h->n_row_tab && (h->row_tab_host[i] & 0x0FFFFFFFu) < (uint32_t)h->plan.slots.prefix_bytes; ++i) ol.n_rows = i + 1;
  switch (h->sp.pomdp_r) {
    case 1: return launch_tiled_r<1>(h, d_obs, s, nullptr, ol);
    case 2: return launch_tiled_r<2>(h, d_obs, s, nullptr, ol);
    case 3: return launch_tiled_r<3>(h, d_obs, s, nullptr, ol);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_obs_tiled(MfgHandle* h, float* d_obs, cudaStream_t s, const uint8_t* skip) {
  const ObsList ol{nullptr, nullptr, nullptr, 0};
  switch (h->sp.pomdp_r) {
    case 1: return launch_tiled_r<1>(h, d_obs, s, skip, ol);
    case 2: return launch_tiled_r<2>(h, d_obs, s, skip, ol);
    case 3: return launch_tiled_r<3>(h, d_obs, s, skip, ol);
    default: return cudaErrorInvalidValue;
  }
}

// resident CTAs per SM of the tiled kernel as planned (diagnostics: mfg_get_info "obs_ctas_per_sm")
template <int R>
static int ctas_per_sm_r(const MfgHandle* h) {
  int n = 0;
  const ObsPlan& p = h->plan;
  if (h->sp.faithful) {
    cudaFuncSetAttribute(k_obs_tiled<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_obs_tiled<R, true>, p.nw * 32, p.smem);
  } else {
    cudaFuncSetAttribute(k_obs_tiled<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_obs_tiled<R, false>, p.nw * 32, p.smem);
  }
  return n;
}
int obs_ctas_per_sm(const MfgHandle* h) {
  if (!h->plan.ok) return 0;
  return h->sp.pomdp_r == 1 ? ctas_per_sm_r<1>(h) : h->sp.pomdp_r == 2 ? ctas_per_sm_r<2>(h) : ctas_per_sm_r<3>(h);
}

cudaError_t launch_obs_direct(MfgHandle* h, float* d_obs, cudaStream_t s) {
  const int threads = 128;
  const int64_t total = h->N * h->sp.n_agents;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  const int A = h->sp.n_agents;
  // per-thread shared-memory slice of the agent's planes when it is small (odd stride: no bank conflicts between threads)
  int cmax = 0;
  for (int a = 0; a < A; ++a) cmax = h->sp.n_channels[a] > cmax ? h->sp.n_channels[a] : cmax;
  int stride = (cmax * h->DD) | 1;
  size_t smem = (size_t)stride * threads * sizeof(float);
  if (smem > 96 * 1024) { stride = 0; smem = 0; }
  cudaError_t err = cudaSuccess;
  if (h->obs_kernel != 3 && h->st.blk_i <= 200 * 1024 && h->st.blk_i % 16 == 0) {      // block-staged exact kernel (option 3: the plain one)
    const unsigned nblk = (unsigned)((h->N + ENV_BLOCK - 1) / ENV_BLOCK);
    auto gob = [&](auto kern) {
      if (h->st.blk_i > 48 * 1024) err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->st.blk_i);
      if (err == cudaSuccess) kern<<<nblk, 256, h->st.blk_i, s>>>(h->d_sp, h->tb, h->st, d_obs, h->total_channels);
    };
    if (A <= 1) gob(k_obs_block<1>);
    else if (A <= 2) gob(k_obs_block<2>);
    else if (A <= 4) gob(k_obs_block<4>);
    else if (A <= 8) gob(k_obs_block<8>);
    else gob(k_obs_block<16>);
    return err != cudaSuccess ? err : cudaGetLastError();
  }
  auto go = [&](auto kern) {
    if (smem > 48 * 1024) err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err == cudaSuccess) kern<<<blocks, threads, smem, s>>>(h->d_sp, h->tb, h->st, d_obs, h->total_channels, stride);
  };
  if (A <= 1) go(k_obs_direct<1>);
  else if (A <= 2) go(k_obs_direct<2>);
  else if (A <= 4) go(k_obs_direct<4>);
  else if (A <= 8) go(k_obs_direct<8>);
  else go(k_obs_direct<16>);
  return err != cudaSuccess ? err : cudaGetLastError();
}

}  // namespace mfg
