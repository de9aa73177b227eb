// mfg_core.cuh - per-environment semantics of the batched marl-factory-grid engine.
//
// One C++ restatement of the reference's hot path, written so that ONE THREAD advances ONE
// environment (struct-of-arrays state, env index fastest => coalesced across the warp):
//   env_step()        Gamestate.tick + check_done + reward fold   (utils/states.py:170-226, factory.py:189-259)
//   env_reset()       Factory.reset spawn rules                    (environment/rules.py:163-199, collection.py:102-130)
//   obs_agent_direct  OBSBuilder.build_for_agent + RayCaster        (utils/observation_builder.py:138-220, ray_caster.py:66-104)
// The normative behaviour is the reference's ACTUAL one, including the
// uid-equality artefact when spec.faithful != 0.
//
// The functions are __host__ __device__ so that tests/hostsim can compile the very same code with g++
// and replay reference traces without a GPU.  That host build is test-only; the product library
// (mfg_kernels.cu) only ever launches them as CUDA kernels.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <type_traits>
#include "../../include/mfg_b200.h"

#if defined(__CUDACC__)
#define MFG_HD __host__ __device__ __forceinline__
#define MFG_HDN __host__ __device__
#define MFG_HDNI __host__ __device__ __noinline__      // rare, large: kept out of the callers' instruction stream
// tiny run-time trip counts, inlined at dozens of call sites: unrolling them quadruples the kernel image (icache misses)
#define MFG_NOUNROLL _Pragma("unroll 1")
#define MFG_UNROLL4 _Pragma("unroll 4")
#define MFG_UNROLL _Pragma("unroll")
#else
#define MFG_HD inline
#define MFG_HDN inline
#define MFG_HDNI inline
#define MFG_UNROLL
#define MFG_NOUNROLL
#define MFG_UNROLL4
#endif

namespace mfg {

constexpr int DOOR_INTERVAL = 10;        // modules/doors/entitites.py:69 auto_close_interval
constexpr double DIRT_PILE_MAX = 5.0;    // modules/clean_up/entitites.py: DirtPile max_local_amount default
constexpr double CHARGE_RATE = 0.4;      // modules/batteries/entitites.py:98
constexpr double ENC_DOOR_OPEN = 0.4444, ENC_DOOR_CLOSED = 0.6666, ENC_MACHINE = 15.0;
constexpr uint16_t NO_POS = 0xFFFF;
constexpr int RESPAWN_TAPE_W = 8;
constexpr int ENV_BLOCK = 128;           // envs per state block (see State)

// entity classes with an integer uid (object.py:103-113)
enum { C_DOOR = 0, C_DIRT, C_ITEM, C_POD, C_DEST, C_DROP, C_MACH, C_MAINT, C_NONE };

// ------------------------------------------------------------------------------------------------
// state layout: every field is a [rows][N] array inside one caller-owned buffer
// ------------------------------------------------------------------------------------------------
// NOTE the order: the positional fields the observation kernel needs come FIRST and in its slot order (dirt piles,
// items, pods, destinations, drop-offs, machines, maintainers, agents, then the door / destination bit masks), so that
// they form one contiguous prefix of every block and can be staged with a single bulk copy.
#define MFG_STATE_FIELDS(F)                                                        \
  F(uint16_t, dirt_pos, sp.has_dirt ? sp.dirt_slots : 0)                           \
  F(uint16_t, item_pos, sp.n_items)                                                \
  F(uint16_t, pod_pos, sp.n_pods)                                                  \
  F(uint16_t, dest_pos, sp.n_dest)                                                 \
  F(uint16_t, drop_pos, sp.n_dropoff)                                              \
  F(uint16_t, mach_pos, sp.n_machines)                                             \
  F(uint16_t, maint_pos, sp.n_maint)                                               \
  F(uint16_t, apos, sp.n_agents)                                                   \
  F(uint64_t, door_open, sp.n_doors ? 1 : 0)                                       \
  F(uint32_t, dest_reached, sp.n_dest ? 1 : 0)                                     \
  F(uint64_t, dirt_listed, sp.has_dirt ? 1 : 0)   /* identity mode: = the live piles */ \
  /* ---- end of the identity-mode observation prefix; the faithful mode also needs listing bits and dirt uids */ \
  F(uint64_t, door_listed, sp.n_doors ? 1 : 0)                                     \
  F(uint32_t, item_listed, sp.n_items ? 1 : 0)                                     \
  F(uint32_t, pod_listed, sp.n_pods ? 1 : 0)                                       \
  F(uint32_t, dest_listed, sp.n_dest ? 1 : 0)                                      \
  F(uint32_t, drop_listed, sp.n_dropoff ? 1 : 0)                                   \
  F(uint32_t, mach_listed, sp.n_machines ? 1 : 0)                                  \
  F(uint32_t, maint_listed, sp.n_maint ? 1 : 0)                                    \
  F(uint16_t, dirt_uid, sp.has_dirt ? sp.dirt_slots : 0)                           \
  /* ---- end of the faithful-mode observation prefix */                           \
  F(uint16_t, step, 1)                                                             \
  F(uint32_t, episode, 1)                                                          \
  F(uint32_t, clock, 1)                                                            \
  F(uint32_t, astamp, sp.n_agents)                                                 \
  F(uint8_t, aflag, sp.n_agents)                                                   \
  F(uint8_t, finished, 1)      /* episode over and not yet re-spawned: statistics are counted once */ \
  F(double, bat, sp.has_batteries ? sp.n_agents : 0)                               \
  F(double, ep_ret, sp.n_agents)                                                   \
  F(uint8_t, door_timer, sp.n_doors)                                               \
  F(double, dirt_amt, sp.has_dirt ? sp.dirt_slots : 0)                             \
  F(uint8_t, dirt_end, sp.has_dirt ? 1 : 0)                                        \
  F(uint8_t, dirt_n, sp.has_dirt ? 1 : 0)                                          \
  F(uint16_t, dirt_next_uid, sp.has_dirt ? 1 : 0)                                  \
  F(int16_t, dirt_next_spawn, sp.has_dirt ? 1 : 0)                                 \
  F(uint16_t, maint_target, sp.n_maint)                                            \
  F(uint16_t, maint_rand, sp.n_maint)                                              \
  F(uint32_t, maint_remaining, sp.n_maint)                                         \
  F(uint8_t, maint_last, sp.n_maint)

// The buffer is BLOCKED: envs are grouped in blocks of ENV_BLOCK = 128; one block holds, back to back, every row of
// every integer / byte field as a [rows][128] slab (blk_i bytes per block, contiguous => one TMA bulk copy stages a
// CTA's whole working set), a second region holds the f64 fields the same way (blk_f bytes per block).  Within a slab
// the env index is the fastest one, so a warp's access to one field row is a single coalesced 32/64/128/256-byte segment.
struct State {
  int64_t N;       // live environments
  size_t blk_i;    // bytes per 128-env block of the integer region
  size_t blk_f;    // bytes per 128-env block of the f64 region
  char* base_i;    // start of block 0 of the integer region
#define F(type, name, rows) type* name;
  MFG_STATE_FIELDS(F)
#undef F
};

// derived level tables (device memory, built by mfg_create)
struct Tables {
  const uint8_t* wall;         // [H*W]
  const uint8_t* door_map;     // [H*W] door index or 0xFF
  const uint16_t* floor_pos;   // [F] pos16
  const uint16_t* floor_index; // [H*W] floor index or 0xFFFF
  const uint16_t* wall_uid;    // [H*W] row-major wall index or 0xFFFF
  const uint16_t* wall_pos;    // [n_walls] pos16
  const uint16_t* door_pos;    // [ND] pos16
  const uint8_t* nexthop;      // [F*F] or null
  const uint64_t* wall_win;    // [H*W] (2r+1)^2-bit wall mask of the window centred on the tile (r <= 3)
  const uint64_t* door_near;   // [H*W] doors inside the radius-D box centred on the tile (bit = door index)
  const uint64_t* door_adj;    // [H*W] doors inside the 3x3 neighbourhood of the tile (DoorUse, doors/actions.py:18-34)
  const uint64_t* wall_box;    // [H*W][4] wall mask of the (2D+1)^2 box (D = 2r+1) centred on the tile, bit = (dx+D)*(2D+1)+(dy+D)
  // faithful observation mode (built by build_vis_tables, mfg_obs.cu): static walls-only visibility
  const uint64_t* vis_box;       // [H*W][4] box cells some full ray reaches when only walls block light (superset of the truth)
  const uint64_t* wall_cand64;   // [H*W] bit u: wall with uid u < 64 lies on such a cell
  const uint32_t* wall_cand_rng; // [H*W] lo | hi << 16: range of the wall uids >= 64 on such cells (lo > hi: none)
  const uint64_t* wall_win64;    // [H*W] bit u: wall with uid u < 64 lies inside the window of the tile
  // window visibility as a table (build_win_vis_tables, mfg_obs.cu): walls are static and a window holds few doors, so the ray
  // march result only depends on (tile, which of the window's doors are closed)
  const uint32_t* door_win;      // [H*W] up to 4 door indices inside the window (6 bits each) | count << 24 (7 = too many: march)
  const uint64_t* vis_tab;       // [H*W][16] visibility mask of the window for every closed-door subset
  // exact per-agent observation path, full observability with at most 4 doors (built by build_tables): first-visit ranks of
  // the ray-radius box for every (tile, closed listed door subset) - the ray walk becomes a table look-up
  const uint16_t* rank_tab;    // [H*W][1 << n_doors][(2R+1)^2] or null
  int64_t env_id_offset;
  unsigned long long* stats;   // [MFG_N_STATS]
};

// element (row, env e) of a field whose block-0 slab starts at `base`
template <typename T>
MFG_HD T& field_at(const State& st, T* base, int row, int64_t e) {
  const size_t blk = std::is_same<T, double>::value ? st.blk_f : st.blk_i;
  char* b = reinterpret_cast<char*>(base) + (size_t)(e >> 7) * blk;
  return reinterpret_cast<T*>(b)[row * ENV_BLOCK + (int)(e & (ENV_BLOCK - 1))];
}

MFG_HD int ctz64(uint64_t m) {            // index of the lowest set bit (m != 0)
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)m) - 1;
#else
  return __builtin_ctzll(m);
#endif
}
MFG_HD int px(uint16_t p) { return p >> 8; }
MFG_HD int py(uint16_t p) { return p & 255; }
MFG_HD uint16_t mkpos(int x, int y) { return (uint16_t)((x << 8) | y); }

MFG_HD int dir_dx(int d) { return d == 0 || d == 4 || d == 7 ? -1 : (d == 2 || d == 5 || d == 6 ? 1 : 0); }
MFG_HD int dir_dy(int d) { return d == 1 || d == 4 || d == 5 ? 1 : (d == 3 || d == 6 || d == 7 ? -1 : 0); }

// run-time index into a small register-resident array: a select chain instead of a local-memory array (lets the agent
// loop of env_step stay rolled - one copy of the action code in the instruction stream instead of AMAX)
template <int N, typename T>
MFG_HD T sel_get(const T (&a)[N], int i) {
  T r = a[0];
  MFG_UNROLL
  for (int j = 1; j < N; ++j) r = (j == i) ? a[j] : r;
  return r;
}
template <int N, typename T>
MFG_HD void sel_set(T (&a)[N], int i, T x) {
  MFG_UNROLL
  for (int j = 0; j < N; ++j) a[j] = (j == i) ? x : a[j];
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10, counter-based: key = seed ^ global env id, counter = (block, step, episode, stream)
// ------------------------------------------------------------------------------------------------
enum { RS_RESET = 0, RS_RESPAWN = 1, RS_ACTIONS = 2, RS_MAINT0 = 8 };

struct Philox {
  uint32_t k0, k1, c1, c2, c3, block;
  uint32_t out[4];
  int have;
  MFG_HD void init(uint64_t seed, uint64_t gid, uint32_t stream, uint32_t episode, uint32_t step) {
    uint64_t k = seed ^ (gid * 0x9E3779B97F4A7C15ull);
    k0 = (uint32_t)k; k1 = (uint32_t)(k >> 32);
    c1 = step; c2 = episode; c3 = stream; block = 0; have = 0;
  }
  static MFG_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
  }
  MFG_HD void refill() {
    uint32_t x0 = block, x1 = c1, x2 = c2, x3 = c3, a = k0, b = k1;
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, x0, hi0, lo0);
      mulhilo(0xCD9E8D57u, x2, hi1, lo1);
      uint32_t y0 = hi1 ^ x1 ^ a, y1 = lo1, y2 = hi0 ^ x3 ^ b, y3 = lo0;
      x0 = y0; x1 = y1; x2 = y2; x3 = y3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
    ++block; have = 4;
  }
  MFG_HD uint32_t next() {
    if (have == 0) refill();
    uint32_t v = have == 4 ? out[0] : have == 3 ? out[1] : have == 2 ? out[2] : out[3];
    --have;
    return v;
  }
  MFG_HD uint32_t below(uint32_t n) { return (uint32_t)(((uint64_t)next() * (uint64_t)n) >> 32); }
  MFG_HD double uniform(double lo, double hi) { return lo + (hi - lo) * ((double)next() * 2.3283064365386963e-10); }
};

// ------------------------------------------------------------------------------------------------
// per-environment context: hot fields live in registers, the rest is touched in the SoA buffers
// ------------------------------------------------------------------------------------------------
// Compact copy of the spec fields the step path reads, small enough to be passed BY VALUE as a kernel parameter
// (constant bank: uniform reads are broadcast, no L1/L2 round trip).  Member names equal MfgSpec's so that the same
// templated code runs against either.
// UIDG: the staged image the spec is used with leaves the dirt-uid rows in global memory (see k_step)
template <int AMAX, bool UIDG_ = false>
struct HotSpec {
  static constexpr bool UIDG = UIDG_;
  int32_t H, W, pomdp_r, n_agents, individual_rewards, faithful, n_floor, n_doors, n_walls, has_dirt, dirt_slots,
      dirt_quantity, has_batteries, has_globalpos, n_items, n_dropoff, n_pods, n_dest, n_machines, n_maint, n_rules, n_groups;
  double dirt_initial_amount, dirt_clean_amount, dirt_max_global, dirt_n_var, dirt_amount_var, battery_initial;
  uint64_t seed;
  int32_t n_actions[AMAX];
  int32_t agent_blocking[AMAX];
  int8_t act_opcode[AMAX][MFG_MAX_ACTIONS];
  int8_t act_dir[AMAX][MFG_MAX_ACTIONS];
  double act_valid[AMAX][MFG_MAX_ACTIONS];
  double act_fail[AMAX][MFG_MAX_ACTIONS];
  double act_aux[AMAX][MFG_MAX_ACTIONS];
  int32_t rule_op[MFG_MAX_RULES];
  double rule_param[MFG_MAX_RULES][MFG_RULE_NPARAM];
  int8_t dest_bound[MFG_MAX_SMALL];
  double act_cost[AMAX][MFG_MAX_ACTIONS + 1];
  // UIDG only: the first uid_head_rows dirt-uid rows ARE part of the image (at byte offset uid_head_off): the uid listing
  // asks for small uids (class members), and slot j always holds a uid >= j, so those queries never look further
  uint32_t uid_head_off;
  int32_t uid_head_rows;
};

template <int AMAX, bool UIDG>
inline void fill_hot_spec(const MfgSpec& s, HotSpec<AMAX, UIDG>& h) {
#define CP(x) h.x = s.x;
  CP(H) CP(W) CP(pomdp_r) CP(n_agents) CP(individual_rewards) CP(faithful) CP(n_floor) CP(n_doors) CP(n_walls) CP(has_dirt)
  CP(dirt_slots) CP(dirt_quantity) CP(has_batteries) CP(has_globalpos) CP(n_items) CP(n_dropoff) CP(n_pods) CP(n_dest)
  CP(n_machines) CP(n_maint) CP(n_rules) CP(n_groups) CP(dirt_initial_amount) CP(dirt_clean_amount) CP(dirt_max_global)
  CP(dirt_n_var) CP(dirt_amount_var) CP(battery_initial) CP(seed)
#undef CP
  for (int i = 0; i < AMAX; ++i) {
    const bool in = i < s.n_agents && i < MFG_MAX_AGENTS;
    h.n_actions[i] = in ? s.n_actions[i] : 0;
    h.agent_blocking[i] = in ? s.agent_blocking[i] : 0;
    for (int a = 0; a < MFG_MAX_ACTIONS; ++a) {
      h.act_opcode[i][a] = in ? (int8_t)s.act_opcode[i][a] : 0;
      h.act_dir[i][a] = in ? (int8_t)s.act_dir[i][a] : 0;
      h.act_valid[i][a] = in ? s.act_valid[i][a] : 0.0;
      h.act_fail[i][a] = in ? s.act_fail[i][a] : 0.0;
      h.act_aux[i][a] = in ? s.act_aux[i][a] : 0.0;
    }
    for (int a = 0; a <= MFG_MAX_ACTIONS; ++a) h.act_cost[i][a] = in ? s.act_cost[i][a] : 0.0;
  }
  for (int k = 0; k < MFG_MAX_SMALL; ++k) h.dest_bound[k] = (int8_t)s.dest_bound[k];
  for (int r = 0; r < MFG_MAX_RULES; ++r) {
    h.rule_op[r] = s.rule_op[r];
    for (int k = 0; k < MFG_RULE_NPARAM; ++k) h.rule_param[r][k] = s.rule_param[r][k];
  }
}

// k_step runs against a shared-memory image of the block (integer fields) and of the small level tables; telling the
// compiler so turns the generic loads / stores of that path into LDS / STS.
template <typename SpecT> struct spec_traits { static constexpr bool staged = false; static constexpr bool uid_global = false; };
template <int AMAX, bool UIDG> struct spec_traits<HotSpec<AMAX, UIDG>> { static constexpr bool staged = true; static constexpr bool uid_global = UIDG; };
// The staged view carries 32-bit OFFSETS in its integer field pointers (built on the host, so they sit in the constant
// bank): an access is `opaque shared base register + offset + index`, three instructions.  (With generic pointers into
// the shared image the compiler re-derived the 64-bit address from scratch at every one of the hundreds of access sites.)
template <typename T>
MFG_HD T& stage_ref(uint32_t sbase, const T* off, int idx) {
#if defined(__CUDA_ARCH__)
  const uint32_t a = sbase + (uint32_t)reinterpret_cast<uintptr_t>(off) + (uint32_t)idx * (uint32_t)sizeof(T);
  T* p = reinterpret_cast<T*>(__cvta_shared_to_generic(a));
  __builtin_assume(__isShared(p));
  return *p;
#else
  (void)sbase;
  return const_cast<T*>(off)[idx];
#endif
}

template <int AMAX, typename SpecT = MfgSpec>
struct Env {
  static constexpr bool STAGED = spec_traits<SpecT>::staged;
  static constexpr bool UIDG = spec_traits<SpecT>::uid_global;     // st.dirt_uid is the global slab (indexed by eg), not part of the image
  const SpecT& sp;
  const Tables& tb;
  const State& st;
  int64_t e;        // env index inside `st`'s integer region (== eg unless that region was staged into a per-CTA copy)
  int64_t eg;       // global env index: f64 fields (never staged), actions / reward / done, tape, Philox key
  uint32_t sbase;   // staged view only: shared-memory address of the block image
  int A;
  uint16_t apos[AMAX];
  uint64_t dopen, dlisted, dirt_listed;
  int dirt_end, dirt_n;

  MFG_HD Env(const SpecT& sp_, const Tables& tb_, const State& st_, int64_t e_, int64_t eg_ = -1, uint32_t sbase_ = 0)
      : sp(sp_), tb(tb_), st(st_), e(e_), eg(eg_ < 0 ? e_ : eg_), sbase(sbase_) {
    A = sp.n_agents;
    dopen = dlisted = dirt_listed = 0;
    dirt_end = dirt_n = 0;
  }
  template <typename T> MFG_HD T& at(T* base, int row) const {
    if constexpr (std::is_same<T, double>::value) return field_at(st, base, row, eg);
    else if constexpr (STAGED) return stage_ref(sbase, base, row * ENV_BLOCK + (int)e);
    else return field_at(st, base, row, e);
  }
  // small level tables (wall map, tile -> door map, door positions): staged copies in k_step
  template <typename T> MFG_HD T tbl(const T* base, int i) const {
    if constexpr (STAGED) return stage_ref(sbase, base, i);
    else return base[i];
  }

  MFG_HD void load() {
#pragma unroll
    for (int i = 0; i < AMAX; ++i) apos[i] = i < A ? at(st.apos, i) : NO_POS;
    if (sp.n_doors) { dopen = at(st.door_open, 0); dlisted = at(st.door_listed, 0); }
    if (sp.has_dirt) { dirt_listed = at(st.dirt_listed, 0); dirt_end = at(st.dirt_end, 0); dirt_n = at(st.dirt_n, 0); }
  }
  MFG_HD void store() {
#pragma unroll
    for (int i = 0; i < AMAX; ++i) if (i < A) at(st.apos, i) = apos[i];
    if (sp.n_doors) { at(st.door_open, 0) = dopen; at(st.door_listed, 0) = dlisted; }
    if (sp.has_dirt) { at(st.dirt_listed, 0) = dirt_listed; at(st.dirt_end, 0) = (uint8_t)dirt_end; at(st.dirt_n, 0) = (uint8_t)dirt_n; }
  }

  // ---------------------------------------------------------------- small-group access by class
  MFG_HD int cls_count(int c) const {
    return c == C_ITEM ? sp.n_items : c == C_POD ? sp.n_pods : c == C_DEST ? sp.n_dest : c == C_DROP ? sp.n_dropoff
         : c == C_MACH ? sp.n_machines : c == C_MAINT ? sp.n_maint : 0;
  }
  MFG_HD uint16_t* cls_pos(int c) const {
    return c == C_ITEM ? st.item_pos : c == C_POD ? st.pod_pos : c == C_DEST ? st.dest_pos : c == C_DROP ? st.drop_pos
         : c == C_MACH ? st.mach_pos : st.maint_pos;
  }
  MFG_HD uint32_t* cls_listed(int c) const {
    return c == C_ITEM ? st.item_listed : c == C_POD ? st.pod_listed : c == C_DEST ? st.dest_listed
         : c == C_DROP ? st.drop_listed : c == C_MACH ? st.mach_listed : st.maint_listed;
  }

  // ---------------------------------------------------------------- tile queries
  MFG_HD bool in_grid(int x, int y) const { return x >= 0 && y >= 0 && x < sp.H && y < sp.W; }
  // dirt uids: rarely touched (uid listing in faithful mode, create / compact), so k_step keeps their rows out of its image
  MFG_HD uint16_t& uid_at(int k) const {
    if constexpr (UIDG) {
      if (k < sp.uid_head_rows) return stage_ref(sbase, reinterpret_cast<const uint16_t*>((uintptr_t)sp.uid_head_off), k * ENV_BLOCK + (int)e);
      return field_at(st, st.dirt_uid, k, eg);
    } else {
      return at(st.dirt_uid, k);
    }
  }
  MFG_HD bool is_wall(int idx) const { return tbl(tb.wall, idx) != 0; }
  MFG_HD int door_idx(int idx) const {          // door index of a tile or -1
    const int d = tbl(tb.door_map, idx);
    return d == 0xFF ? -1 : d;
  }
  MFG_HD int door_at(int x, int y) const { return door_idx(x * sp.W + y); }
  MFG_HD bool closed_listed_door(int x, int y) const {
    if (!sp.n_doors) return false;
    int d = door_at(x, y);
    return d >= 0 && !((dopen >> d) & 1) && ((dlisted >> d) & 1);
  }
  MFG_HD int agents_at(uint16_t p) const {
    int n = 0;
#pragma unroll
    for (int i = 0; i < AMAX; ++i) n += (i < A && apos[i] == p) ? 1 : 0;
    return n;
  }
  MFG_HD int listed_maints_at(uint16_t p) const {
    int n = 0;
    if (sp.n_maint) {
      uint32_t l = at(st.maint_listed, 0);
      MFG_NOUNROLL
      for (int k = 0; k < sp.n_maint; ++k) n += (((l >> k) & 1) && at(st.maint_pos, k) == p) ? 1 : 0;
    }
    return n;
  }
  // states.py:259-270 check_pos_validity (negated): wall / off-grid / closed listed door / blocking agent
  MFG_HD bool blocked(int x, int y) const {
    if (!in_grid(x, y) || is_wall(x * sp.W + y)) return true;
    if (closed_listed_door(x, y)) return true;
    uint16_t p = mkpos(x, y);
#pragma unroll
    for (int i = 0; i < AMAX; ++i) if (i < A && sp.agent_blocking[i] && apos[i] == p) return true;
    return false;
  }
  // number of collidable LISTED entities on an in-grid tile: agents, maintainers, closed doors, walls
  MFG_HD int n_coll(int x, int y) const {
    uint16_t p = mkpos(x, y);
    return agents_at(p) + listed_maints_at(p) + (closed_listed_door(x, y) ? 1 : 0) + (is_wall(x * sp.W + y) ? 1 : 0);
  }
  MFG_HD bool is_free(int x, int y) const { return !blocked(x, y) && n_coll(x, y) == 0; }

  // ---------------------------------------------------------------- uid listing (objects.py:193-214)
  // own_dirt: the query is for the uid of a dirt pile itself.  Dirt uids are unique among the piles, so the pile part of
  // the search is then known without a scan: k >= 0 = only slot k can match (dirt_delete), -1 = a fresh uid matches no
  // pile (dirt_create); -2 = some other entity's uid: scan.
  MFG_HD bool find_listed(int uid, uint16_t p, int& cls, int& idx, int own_dirt = -2) const {
    if (uid < sp.n_doors && tbl(tb.door_pos, uid) == p && ((dlisted >> uid) & 1)) { cls = C_DOOR; idx = uid; return true; }
    if (own_dirt >= 0) {
      if ((dirt_listed >> own_dirt) & 1) { cls = C_DIRT; idx = own_dirt; return true; }
    } else if (own_dirt == -2 && sp.has_dirt && uid < (int)at(st.dirt_next_uid, 0)) {
      // slots are in creation order and uids only grow (deleted slots keep theirs, compaction keeps the order), so the
      // scan can stop at the first larger uid - the uids asked for (maintainers, items, doors ...) are small.  The uids
      // of an episode start at 0, so slot k holds a uid >= k: nothing beyond slot `uid` can match.
      MFG_NOUNROLL
      for (int k = 0; k < dirt_end && k <= uid; ++k) {
        const int du = uid_at(k);
        if (du > uid) break;
        if (du == uid && at(st.dirt_pos, k) == p && ((dirt_listed >> k) & 1)) { cls = C_DIRT; idx = k; return true; }
      }
    }
    MFG_UNROLL
    for (int c = C_ITEM; c <= C_MAINT; ++c) {
      if (uid < cls_count(c) && at(cls_pos(c), uid) == p && ((at(cls_listed(c), 0) >> uid) & 1)) { cls = c; idx = uid; return true; }
    }
    return false;
  }
  MFG_HD void set_listed(int cls, int idx, bool v) {
    if (cls == C_DOOR) dlisted = v ? (dlisted | (1ull << idx)) : (dlisted & ~(1ull << idx));
    else if (cls == C_DIRT) dirt_listed = v ? (dirt_listed | (1ull << idx)) : (dirt_listed & ~(1ull << idx));
    else { uint32_t& l = at(cls_listed(cls), 0); l = v ? (l | (1u << idx)) : (l & ~(1u << idx)); }
  }
  MFG_HD void l_add(int cls, int idx, int uid, uint16_t p, int own_dirt = -2) {
    int c2, i2;
    if (sp.faithful && find_listed(uid, p, c2, i2, own_dirt)) set_listed(cls, idx, false);
    else set_listed(cls, idx, true);
  }
  MFG_HD void l_del(int cls, int idx, int uid, uint16_t p, int own_dirt = -2) {
    if (sp.faithful) {
      int c2, i2;
      if (find_listed(uid, p, c2, i2, own_dirt)) set_listed(c2, i2, false);
    } else {
      set_listed(cls, idx, false);
    }
  }

  // ---------------------------------------------------------------- doors (doors/actions.py:18-34, entitites.py:97-140)
  MFG_HD bool toggle_near(uint16_t p) {
    // doors of the 3x3 neighbourhood (POS_MASK_8 + own tile) from the per-tile table, in door order
    uint64_t m = tb.door_adj[px(p) * sp.W + py(p)] & dlisted;
    const bool valid = m != 0;
    for (; m; m &= m - 1) {
      const int d = ctz64(m);
      if ((dopen >> d) & 1) dopen &= ~(1ull << d);
      else { dopen |= (1ull << d); at(st.door_timer, d) = DOOR_INTERVAL; }
    }
    return valid;
  }

  // ---------------------------------------------------------------- dirt (clean_up/groups.py:70-95, actions.py:19-36)
  // bit k = pred(position of slot k) for the slots below dirt_end: branch-free, four independent loads in flight (an
  // early-exit scan is one dependent shared-memory round trip per slot)
  template <typename Pred>
  MFG_HD uint64_t dirt_scan(Pred pred) const {
    uint32_t lo = 0u, hi = 0u;
    const int n_lo = dirt_end < 32 ? dirt_end : 32;
    MFG_UNROLL4
    for (int k = 0; k < n_lo; ++k) lo |= pred(at(st.dirt_pos, k)) ? (1u << k) : 0u;
    MFG_UNROLL4
    for (int k = 32; k < dirt_end; ++k) hi |= pred(at(st.dirt_pos, k)) ? (1u << (k - 32)) : 0u;
    return (uint64_t)lo | ((uint64_t)hi << 32);
  }
  MFG_HD int dirt_at(uint16_t p) const {        // first slot on tile p (tombstones hold NO_POS) or -1
    const uint64_t m = dirt_scan([&](uint16_t q) { return q == p; });
    return m ? ctz64(m) : -1;
  }
  MFG_HD double dirt_sum() const {
    // left-to-right f64 sum over the live piles (clean_up/groups.py:27-32).  The amounts live in HBM: fetch eight at a
    // time with independent loads, then add in order (a dependent load per pile made the respawning lane - and with it
    // the whole CTA - wait ~40 memory round trips).
    double s = 0.0;
    MFG_NOUNROLL
    for (int k0 = 0; k0 < dirt_end; k0 += 8) {
      double a[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int j = 0; j < 8; ++j) a[j] = (k0 + j < dirt_end) ? at(st.dirt_amt, k0 + j) : 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int j = 0; j < 8; ++j) if (k0 + j < dirt_end && at(st.dirt_pos, k0 + j) != NO_POS) s += a[j];
    }
    return s;
  }
  MFG_HD void dirt_compact() {
    int w = 0;
    uint64_t nl = 0;
    MFG_NOUNROLL
    for (int k = 0; k < dirt_end; ++k) {
      uint16_t p = at(st.dirt_pos, k);
      if (p == NO_POS) continue;
      if (w != k) { at(st.dirt_pos, w) = p; at(st.dirt_amt, w) = at(st.dirt_amt, k); uid_at(w) = uid_at(k); }
      if ((dirt_listed >> k) & 1) nl |= 1ull << w;
      ++w;
    }
    MFG_NOUNROLL
    for (int k = w; k < dirt_end; ++k) at(st.dirt_pos, k) = NO_POS;
    dirt_end = w; dirt_listed = nl;
  }
  MFG_HD void dirt_create(uint16_t p, double amount) {
    if (dirt_end == sp.dirt_slots) dirt_compact();
    uint16_t uid = at(st.dirt_next_uid, 0);
    at(st.dirt_next_uid, 0) = (uint16_t)(uid + 1);      // the reference's uid counter advances regardless
    if (dirt_end == sp.dirt_slots) {
#if defined(__CUDA_ARCH__)
      atomicAdd(&tb.stats[MFG_ST_DIRT_OVERFLOW], 1ull);
#else
      tb.stats[MFG_ST_DIRT_OVERFLOW] += 1;
#endif
      return;
    }
    int k = dirt_end++;
    at(st.dirt_pos, k) = p; at(st.dirt_amt, k) = amount; uid_at(k) = uid;
    ++dirt_n;
    l_add(C_DIRT, k, uid, p, -1);
  }
  MFG_HD void dirt_delete(int k) {
    l_del(C_DIRT, k, sp.faithful ? (int)uid_at(k) : 0, at(st.dirt_pos, k), k);
    dirt_listed &= ~(1ull << k);
    at(st.dirt_pos, k) = NO_POS;
    --dirt_n;
  }
  // trigger_spawn body after the random draws: tiles/amounts zipped
  template <typename TileFn, typename AmtFn>
  MFG_HD void dirt_spawn(int n, TileFn tile, AmtFn amount) {
    // global_amount (clean_up/groups.py:27-32) is a left-to-right f64 sum over the piles in creation order; a new pile
    // is appended at the end, so `sum + a` IS the recomputed sum bit for bit; only a top-up forces a re-scan
    double total = n > 0 ? dirt_sum() : 0.0;
    for (int j = 0; j < n; ++j) {
      if (total > sp.dirt_max_global) return;
      uint16_t p = tile(j);
      double a = amount(j);
      int k = dirt_at(p);
      if (k >= 0) { at(st.dirt_amt, k) = fmin(at(st.dirt_amt, k) + a, DIRT_PILE_MAX); total = dirt_sum(); }
      else { const int before = dirt_n; dirt_create(p, a); if (dirt_n > before) total += a; }
    }
  }

  // ---------------------------------------------------------------- moves (actions.py:77-100, states.py:240-257)
  MFG_HD bool try_move(uint16_t p, int d, bool mover_blocks, uint16_t& target) const {
    int x = px(p) + dir_dx(d), y = py(p) + dir_dy(d);
    if (blocked(x, y)) return false;
    if (mover_blocks && n_coll(x, y) >= 1) return false;      // is_occupied (global_entities.py:187-194)
    target = mkpos(x, y);
    return true;
  }

  // ---------------------------------------------------------------- free tile sampling (global_entities.py:111-121)
  MFG_HD uint16_t sample_free(Philox& rng, const uint16_t* taken, int n_taken, bool must_be_empty) const {
    for (int attempt = 0; attempt < 64 + 16 * sp.n_floor; ++attempt) {
      uint16_t p = tb.floor_pos[rng.below((uint32_t)sp.n_floor)];
      int x = px(p), y = py(p);
      bool ok = must_be_empty ? (agents_at(p) == 0 && (sp.n_doors == 0 || door_at(x, y) < 0)) : is_free(x, y);
      MFG_NOUNROLL
      for (int j = 0; ok && j < n_taken; ++j) ok = taken[j] != p;
      if (ok) return p;
    }
    return NO_POS;
  }
};

MFG_HD void stat_add(const Tables& tb, int idx, unsigned long long v) {
#if defined(__CUDA_ARCH__)
  atomicAdd(&tb.stats[idx], v);
#else
  tb.stats[idx] += v;
#endif
}
MFG_HD void stat_add_f64(const Tables& tb, int idx, double v) {
#if defined(__CUDA_ARCH__)
  atomicAdd(reinterpret_cast<double*>(&tb.stats[idx]), v);
#else
  *reinterpret_cast<double*>(&tb.stats[idx]) += v;
#endif
}

// ================================================================================================
// reset: Factory.reset with a fresh Factory: SpawnAgents, then the groups in Entities order
// ================================================================================================
template <int AMAX>
MFG_HD void env_reset_inl(const MfgSpec& sp, const Tables& tb, const State& st, int64_t e, uint32_t episode, int64_t eg = -1);
// out-of-line form: kept out of the instruction stream of kernels that only re-spawn on a rare path (k_step)
template <int AMAX>
MFG_HDNI void env_reset(const MfgSpec& sp, const Tables& tb, const State& st, int64_t e, uint32_t episode,
                       int64_t eg = -1) {
  env_reset_inl<AMAX>(sp, tb, st, e, episode, eg);
}
template <int AMAX>
MFG_HD void env_reset_inl(const MfgSpec& sp, const Tables& tb, const State& st, int64_t e, uint32_t episode, int64_t eg) {
  Env<AMAX> v(sp, tb, st, e, eg);
  const int A = v.A;
  Philox rng;
  rng.init(sp.seed, (uint64_t)(tb.env_id_offset + v.eg), RS_RESET, episode, 0);

  v.at(st.step, 0) = 0;
  v.at(st.episode, 0) = episode;
  v.at(st.clock, 0) = (uint32_t)A;
  v.at(st.finished, 0) = 0;
  for (int i = 0; i < A; ++i) {
    v.at(st.astamp, i) = (uint32_t)i;
    v.at(st.aflag, i) = 0;
    v.at(st.ep_ret, i) = 0.0;
    if (sp.has_batteries) v.at(st.bat, i) = sp.battery_initial;
    v.apos[i] = NO_POS;
  }
  // doors: closed, timer = interval, listed (doors/entitites.py:150-156 Door.reset)
  if (sp.n_doors) {
    v.dopen = 0;
    v.dlisted = sp.n_doors == 64 ? ~0ull : ((1ull << sp.n_doors) - 1);
    for (int d = 0; d < sp.n_doors; ++d) v.at(st.door_timer, d) = DOOR_INTERVAL;
  }
  if (sp.has_dirt) {
    for (int k = 0; k < sp.dirt_slots; ++k) v.at(st.dirt_pos, k) = NO_POS;
    v.dirt_listed = 0; v.dirt_end = 0; v.dirt_n = 0;
    v.at(st.dirt_next_uid, 0) = 0;
    int16_t next = -1;
    for (int r = 0; r < sp.n_rules; ++r) if (sp.rule_op[r] == MFG_R_RESPAWN_DIRT) next = (int16_t)sp.rule_param[r][0];
    v.at(st.dirt_next_spawn, 0) = next;              // clean_up/rules.py:47 (fresh rule object)
  }
  MFG_NOUNROLL
  for (int c = C_ITEM; c <= C_MAINT; ++c) {
    int n = v.cls_count(c);
    MFG_NOUNROLL
    for (int k = 0; k < n; ++k) v.at(v.cls_pos(c), k) = NO_POS;
    if (n) v.at(v.cls_listed(c), 0) = 0;
  }
  if (sp.n_dest) v.at(st.dest_reached, 0) = 0;
  for (int k = 0; k < sp.n_maint; ++k) {
    v.at(st.maint_target, k) = NO_POS; v.at(st.maint_rand, k) = NO_POS;
    v.at(st.maint_remaining, k) = 0; v.at(st.maint_last, k) = 0xFF;
  }

  // ---- SpawnAgents (rules.py:182-199): configured position if still empty, else a random EMPTY tile
  for (int i = 0; i < A; ++i) {
    uint16_t p = NO_POS;
    if (sp.agent_n_fixed[i] > 0) {
      for (int j = 0; j < sp.agent_n_fixed[i] && p == NO_POS; ++j) {
        uint16_t q = sp.agent_fixed_pos[i][j];
        if (v.agents_at(q) == 0 && (sp.n_doors == 0 || v.door_at(px(q), py(q)) < 0)) p = q;
      }
    } else {
      p = v.sample_free(rng, nullptr, 0, true);
    }
    if (p == NO_POS) { stat_add(tb, MFG_ST_SPAWN_FAIL, 1); p = tb.floor_pos[0]; }
    v.apos[i] = p;
  }

  // ---- SpawnEntity rules in Entities order (rules.py:163-167 -> collection.py:102-130)
  uint16_t chosen[MFG_MAX_DIRT];
  for (int g = 0; g < sp.n_groups; ++g) {
    int gid = sp.group_id[g];
    if (gid == MFG_SP_DIRT) {
      // clean_up/groups.py:70-95: count draw, positions, then `quantity` amount draws
      int q = sp.dirt_quantity;
      int n_new = (int)fabs((double)q + rng.uniform(-sp.dirt_n_var, sp.dirt_n_var));
      if (n_new > MFG_MAX_DIRT) n_new = MFG_MAX_DIRT;
      int got = 0;
      for (int j = 0; j < n_new; ++j) {
        uint16_t p = v.sample_free(rng, chosen, got, false);
        if (p == NO_POS) break;
        chosen[got++] = p;
      }
      double amounts[MFG_MAX_DIRT];
      int qa = q < MFG_MAX_DIRT ? q : MFG_MAX_DIRT;
      for (int j = 0; j < qa; ++j) amounts[j] = sp.dirt_initial_amount + rng.uniform(-sp.dirt_amount_var, sp.dirt_amount_var);
      int n = got < qa ? got : qa;
      v.dirt_spawn(n, [&](int j) { return chosen[j]; }, [&](int j) { return amounts[j]; });
    } else if (gid == MFG_SP_DEST && sp.dest_mode != MFG_DEST_FREE) {
      // bound destinations (modules/destinations/rules.py:95-162): one per agent / per dict entry
      for (int k = 0; k < sp.n_dest; ++k) {
        const int ag = sp.dest_bound[k];
        uint16_t p = NO_POS;
        if (sp.dest_mode == MFG_DEST_ON_AGENT) {
          p = v.apos[ag];                                   // "just below him": the agent's own tile
        } else {
          // shuffled candidate list, first tile that is not the agent's and holds no destination yet == a uniform draw
          // among the valid candidates
          const int nc = sp.dest_n_cand[k] > 0 ? sp.dest_n_cand[k] : sp.n_floor;
          uint16_t valid[MFG_MAX_FIXED];
          int nv = 0;
          if (sp.dest_n_cand[k] > 0) {
            for (int j = 0; j < nc; ++j) {
              const uint16_t q = sp.dest_cand[k][j];
              bool ok = q != v.apos[ag];
              for (int m = 0; ok && m < k; ++m) ok = v.at(st.dest_pos, m) != q;
              if (ok) valid[nv++] = q;
            }
            if (nv) p = valid[rng.below((uint32_t)nv)];
          } else {
            for (int attempt = 0; attempt < 64 + 16 * sp.n_floor && p == NO_POS; ++attempt) {
              const uint16_t q = tb.floor_pos[rng.below((uint32_t)sp.n_floor)];
              bool ok = q != v.apos[ag];
              for (int m = 0; ok && m < k; ++m) ok = v.at(st.dest_pos, m) != q;
              if (ok) p = q;
            }
          }
        }
        if (p == NO_POS) { stat_add(tb, MFG_ST_SPAWN_FAIL, 1); continue; }        // the reference exits here
        v.at(st.dest_pos, k) = p;
        v.l_add(C_DEST, k, k, p);
      }
    } else if (gid == MFG_SP_PODS || gid == MFG_SP_DEST || gid == MFG_SP_ITEMS || gid == MFG_SP_DROPOFF ||
               gid == MFG_SP_MACHINES || gid == MFG_SP_MAINT) {
      int c = gid == MFG_SP_PODS ? C_POD : gid == MFG_SP_DEST ? C_DEST : gid == MFG_SP_ITEMS ? C_ITEM
            : gid == MFG_SP_DROPOFF ? C_DROP : gid == MFG_SP_MACHINES ? C_MACH : C_MAINT;
      int n = sp.group_quantity[g];
      int got = 0;
      for (int j = 0; j < n; ++j) {
        uint16_t p = sp.group_n_fixed[g] > 0 ? sp.group_fixed_pos[g][j] : v.sample_free(rng, chosen, got, false);
        if (p == NO_POS) { stat_add(tb, MFG_ST_SPAWN_FAIL, 1); break; }
        chosen[got++] = p;
      }
      // all tiles are picked against the pre-spawn state (islice over the generator), then entities are created
      for (int j = 0; j < got; ++j) {
        v.at(v.cls_pos(c), j) = chosen[j];
        v.l_add(c, j, j, chosen[j]);
      }
    }
  }
  // ---- DoRandomInitialSteps.on_reset_post_spawn (environment/rules.py:341-355): a random free tile, one of its 4 floor
  // neighbours, the first agent standing there moves onto the free tile
  if (sp.random_initial_steps > 0) {
    uint32_t clock = v.at(st.clock, 0);
    for (int it = 0; it < sp.random_initial_steps; ++it) {
      const uint16_t fp = v.sample_free(rng, nullptr, 0, false);
      if (fp == NO_POS) break;
      uint16_t nb[4];
      int nn = 0;
      for (int d = 0; d < 4; ++d) {                         // POS_MASK_4, restricted to floor tiles
        const int x = px(fp) + dir_dx(d), y = py(fp) + dir_dy(d);
        if (v.in_grid(x, y) && !v.is_wall(x * sp.W + y)) nb[nn++] = mkpos(x, y);
      }
      if (!nn) continue;
      const uint16_t from = nb[rng.below((uint32_t)nn)];
      int who = -1;
      uint32_t best = 0;
      for (int i = 0; i < A; ++i)                           // Agents.by_pos: the agent that arrived first
        if (v.apos[i] == from) { const uint32_t s2 = v.at(st.astamp, i); if (who < 0 || s2 < best) { who = i; best = s2; } }
      if (who < 0) continue;                                // (the reference asserts an agent stands there)
      // Entity.move re-validates against the live state (entity.py:175-199)
      const int x = px(fp), y = py(fp);
      if (v.blocked(x, y) || (sp.agent_blocking[who] && v.n_coll(x, y) >= 1)) continue;
      v.apos[who] = fp;
      v.at(st.astamp, who) = clock++;
    }
    v.at(st.clock, 0) = clock;
  }
  v.store();
}

// ================================================================================================
// step
// ================================================================================================
struct StepIO {
  const int32_t* actions;     // [N][A]
  const uint8_t* maint_act;   // [N][NM] or null
  const int8_t* respawn_n;    // [N] or null
  const uint16_t* respawn_pos;// [N][8] or null
  float* reward;              // [N][A]
  uint8_t* done;              // [N]
  int auto_reset;             // 0 = never, 1 = re-spawn finished envs (inline, or deferred when reset_list is set)
  uint32_t* reset_list;       // deferred reset: finished env ids are appended here and re-spawned by k_reset_list, packed
  uint32_t* reset_count;      //                 (an in-line reset makes almost every warp run the long spawn path for 1-2 lanes)
  uint8_t* flags;             // optional [N][A + 1]: per agent MFG_FLAG_* bits of this step, then the done reason (0 = not done)
};

// maintainer policy when no tape is given (maintenance/entities.py:37-136), next hop from the BFS table
template <int AMAX, typename SpecT>
MFG_HD int maint_policy(Env<AMAX, SpecT>& v, int k, uint32_t step) {
  const SpecT& sp = v.sp; const State& st = v.st; const Tables& tb = v.tb;
  uint16_t p = v.at(st.maint_pos, k);
  int here = -1;
  for (int m = 0; m < sp.n_machines && here < 0; ++m) if (v.at(st.mach_pos, m) == p) here = m;
  if (here >= 0 && here != (int)v.at(st.maint_last, k)) { v.at(st.maint_last, k) = (uint8_t)here; return MFG_MAINT_MACHINE; }
  uint16_t target = v.at(st.maint_target, k);
  if (target == NO_POS || target == p) {
    Philox rng;
    rng.init(sp.seed, (uint64_t)(tb.env_id_offset + v.eg), RS_MAINT0 + k, v.at(st.episode, 0), step);
    for (int attempt = 0; attempt < 2; ++attempt) {
      uint32_t rem = v.at(st.maint_remaining, k);
      if (rem == 0) {
        if (attempt == 1) break;                       // reference would raise IndexError here
        v.at(st.maint_rand, k) = v.sample_free(rng, nullptr, 0, false);
        rem = (sp.n_machines >= 31 ? 0x7FFFFFFFu : ((1u << (sp.n_machines + 1)) - 1));
      }
      int cnt = 0;
      for (int b = 0; b <= sp.n_machines; ++b) cnt += (rem >> b) & 1;
      int pick = (int)rng.below((uint32_t)cnt), sel = 0;
      for (int b = 0; b <= sp.n_machines; ++b) if ((rem >> b) & 1) { if (pick-- == 0) { sel = b; break; } }
      v.at(st.maint_remaining, k) = rem & ~(1u << sel);
      target = sel < sp.n_machines ? v.at(st.mach_pos, sel) : v.at(st.maint_rand, k);
      if (target != NO_POS && target != p) break;
    }
    v.at(st.maint_target, k) = target;
    if (target == NO_POS || target == p) return MFG_MAINT_NOOP;
  }
  int fi = tb.floor_index[px(p) * sp.W + py(p)], fj = tb.floor_index[px(target) * sp.W + py(target)];
  int d = tb.nexthop ? tb.nexthop[(size_t)fi * sp.n_floor + fj] : 255;
  if (d > 7) return MFG_MAINT_NOOP;
  int nx = px(p) + dir_dx(d), ny = py(p) + dir_dy(d);
  int door = sp.n_doors ? v.door_at(nx, ny) : -1;
  if (door >= 0 && !((v.dopen >> door) & 1)) return MFG_MAINT_DOORUSE;     // Doors.by_pos: listing ignored
  if (v.n_coll(nx, ny) > 0) return MFG_MAINT_NOOP;
  return d;
}

// `sp` may be the full MfgSpec or its compact HotSpec copy; `full` (the MfgSpec) is only read by the in-kernel reset
// FLAGS: also write the per-agent result flags (StepIO.flags); a compile-time switch so that the plain step does not carry it
// SYNC (k_step only): CTA barriers at the convergent points of the step - the top of every agent iteration and of every
// rule-loop iteration - keep the warps of a CTA at the same place in the (long, mostly straight-line) program, so they
// share instruction-cache lines instead of each streaming the program from L2 on its own.  STEP_SYNC_POINTS = how many
// barriers one call executes (threads without an env execute the same number).
#ifndef MFG_STEP_SYNC_FINE
#define MFG_STEP_SYNC_FINE 1           // 1 = more barriers inside the agent iteration, DoorAutoClose and MoveMaintainers
#endif
// number of barriers one env_step<SYNC = true> call executes (threads of a partly filled block execute the same number)
template <typename SpecT>
MFG_HD int step_sync_points(const SpecT& sp) {
  int n = sp.n_agents + 3 * sp.n_rules + 1;
  if (MFG_STEP_SYNC_FINE)
    for (int r = 0; r < sp.n_rules; ++r)
      n += sp.rule_op[r] == MFG_R_DOOR_AUTO_CLOSE ? 2 : sp.rule_op[r] == MFG_R_MOVE_MAINTAINERS ? sp.n_maint : 0;
  return n;
}
template <bool SYNC>
MFG_HD void step_sync() {
#if defined(__CUDA_ARCH__)
  // (per-thread barrier, not the warp-aligned __syncthreads: lanes of a partly filled last warp arrive from elsewhere)
  if constexpr (SYNC) asm volatile("barrier.sync 0;\n" ::: "memory");
#endif
}
template <int AMAX, typename SpecT, bool FLAGS = true, bool SYNC = false>
MFG_HDN void env_step(const SpecT& sp, const MfgSpec& full, const Tables& tb, const State& st, int64_t e_local,
                      const StepIO& io, int64_t eg = -1, uint32_t sbase = 0) {
  Env<AMAX, SpecT> v(sp, tb, st, e_local, eg, sbase);
  const int64_t e = v.eg;            // index into the caller's actions / tape / reward / done buffers
  v.load();
  const int A = v.A;
  const int step = (int)v.at(st.step, 0) + 1;
  v.at(st.step, 0) = (uint16_t)step;

  double rew[AMAX];
#pragma unroll
  for (int i = 0; i < AMAX; ++i) rew[i] = 0.0;
  double glob = 0.0;
  uint32_t okmask = 0u, collmask = 0u, mcollmask = 0u, auxmask = 0u;          // per-agent result bits of this step (StepIO.flags)
  // The f64 fields live in HBM (never staged): fetch every battery level, episode return and action of this env up
  // front with independent loads (one round trip instead of a dependent one per use) and keep them in registers.
  double bat[AMAX], epr[AMAX];
  int act[AMAX];
  uint32_t skipmask = 0u;                  // paralysed agents of this step
#pragma unroll
  for (int i = 0; i < AMAX; ++i) {
    bat[i] = (i < A && sp.has_batteries) ? v.at(st.bat, i) : 1.0;
    epr[i] = i < A ? v.at(st.ep_ret, i) : 0.0;
    act[i] = i < A ? io.actions[(size_t)e * A + i] : 0;
  }

  // ---- agents act sequentially against the live state (states.py:189-198)
  // (rolled: ONE copy of the action code; the per-agent registers are reached through select chains)
  MFG_NOUNROLL
  for (int i = 0; i < A; ++i) {
    step_sync<SYNC>();
    if (v.at(st.aflag, i) & 1) { skipmask |= 1u << i; continue; }      // paralysed: skipped entirely
    int a = sel_get(act, i);
    if (a < 0 || a >= sp.n_actions[i]) a = 0;
    const int op = sp.act_opcode[i][a];
    const uint16_t p = sel_get(v.apos, i);
    bool ok = false;
    double r_extra = 0.0;
    bool use_extra = false;
    if (op == MFG_OP_MOVE) {
      uint16_t t;
      ok = v.try_move(p, sp.act_dir[i][a], sp.agent_blocking[i] != 0, t);
      if (ok) {
        sel_set(v.apos, i, t);
        uint32_t c = v.at(st.clock, 0);
        v.at(st.astamp, i) = c;
        v.at(st.clock, 0) = c + 1;
      }
      // actions.py:80-96 action_introduced_collision: the move failed, or the mover now shares its tile with a collidable
      if (FLAGS && io.flags && (!ok || v.n_coll(px(t), py(t)) > 1)) mcollmask |= 1u << i;
    } else if (op == MFG_OP_NOOP) {
      ok = true;
    } else if (op == MFG_OP_DOORUSE) {
      ok = v.toggle_near(p);
    } else if (op == MFG_OP_CLEAN) {                           // clean_up/actions.py:19-36 (global index => listed pile)
      int k = v.dirt_at(p);
      ok = k >= 0 && ((v.dirt_listed >> k) & 1);
      if (ok) {
        double nw = v.at(st.dirt_amt, k) - sp.dirt_clean_amount;
        if (nw <= 0) v.dirt_delete(k);
        else v.at(st.dirt_amt, k) = fmin(fmax(nw, 0.0), DIRT_PILE_MAX);
      }
    } else if (op == MFG_OP_ITEM) {                            // items/actions.py:41-63
      bool on_drop = false;
      MFG_NOUNROLL
      for (int k = 0; k < sp.n_dropoff; ++k) on_drop |= v.at(st.drop_pos, k) == p;
      if (on_drop) { use_extra = true; r_extra = sp.act_aux[i][a]; }
      else {
        int it = -1;
        MFG_NOUNROLL
        for (int k = 0; k < sp.n_items && it < 0; ++k) if (v.at(st.item_pos, k) == p) it = k;
        ok = it >= 0;
        if (ok) { v.l_del(C_ITEM, it, it, p); v.set_listed(C_ITEM, it, false); v.at(st.item_pos, it) = NO_POS; }
      }
    } else if (op == MFG_OP_CHARGE) {                          // batteries/actions.py:20-31, entitites.py:98-111
      bool on_pod = false;
      MFG_NOUNROLL
      for (int k = 0; k < sp.n_pods; ++k) on_pod |= v.at(st.pod_pos, k) == p;
      if (on_pod) {
        double b = sel_get(bat, i);
        if (!(b >= 1.0) && !(v.agents_at(p) > 1)) { sel_set(bat, i, fmin(1.0, CHARGE_RATE + b)); ok = true; }
      }
    } else if (op == MFG_OP_DEST) {                            // destinations/actions.py:17-24 (reference raises on a dest)
      ok = false;
    } else if (op == MFG_OP_MACHINE) {                         // machines/actions.py:19-25
      MFG_NOUNROLL
      for (int k = 0; k < sp.n_machines; ++k) ok |= v.at(st.mach_pos, k) == p;
    }
    const double r_act = use_extra ? r_extra : (ok ? sp.act_valid[i][a] : sp.act_fail[i][a]);
#pragma unroll
    for (int j = 0; j < AMAX; ++j) if (j == i) rew[j] += r_act;
    if (FLAGS && ok) okmask |= 1u << i;
    if (FLAGS && use_extra) auxmask |= 1u << i;
  }

  // ---- tick_step hooks in yaml order (states.py:56-61)
  for (int r = 0; r < sp.n_rules; ++r) {
    step_sync<SYNC>();
    const int op = sp.rule_op[r];
    const double* P = sp.rule_param[r];
    if (op == MFG_R_DOOR_AUTO_CLOSE) {
      // doors/entitites.py:108-122: len(global pos_dict[door.pos]) = agents + listed entities, 2-bit saturating counters
      uint64_t c0 = v.dlisted, c1 = 0;
      auto add = [&](uint16_t q) {
        if (q == NO_POS) return;
        int d = v.door_idx(px(q) * sp.W + py(q));
        if (d < 0) return;
        uint64_t b = 1ull << d, carry = c0 & b;
        c0 ^= b;
        uint64_t carry2 = c1 & carry;
        c1 ^= carry;
        c0 |= carry2; c1 |= carry2;
      };
#pragma unroll
      for (int i = 0; i < AMAX; ++i) if (i < A) add(v.apos[i]);
      step_sync<SYNC && MFG_STEP_SYNC_FINE>();
      if (sp.has_dirt) {
        // listed piles on door tiles: one branch-free pass marks the slots whose tile holds a door (a tombstone's NO_POS is
        // clamped into the map and masked out by the listing bits), the few hits are then counted
        const int last = sp.H * sp.W - 1;
        uint64_t hits = v.dirt_scan([&](uint16_t q) {
          const int idx = px(q) * sp.W + py(q);
          return v.tbl(tb.door_map, idx < last ? idx : last) != 0xFF;
        }) & v.dirt_listed;
        for (; hits; hits &= hits - 1) add(v.at(st.dirt_pos, ctz64(hits)));
      }
      MFG_UNROLL
      for (int c = C_ITEM; c <= C_MAINT; ++c) {
        int n = v.cls_count(c);
        if (!n) continue;
        uint32_t l = v.at(v.cls_listed(c), 0);
        MFG_NOUNROLL
        for (int k = 0; k < n; ++k) if ((l >> k) & 1) add(v.at(v.cls_pos(c), k));
      }
      // count n = c0 + 2 c1 (saturating at 3).  Only open doors (count down / close) and crowded tiles (n = 3: timer
      // reset) change state, so only those doors are visited.
      step_sync<SYNC && MFG_STEP_SYNC_FINE>();
      const uint64_t all = sp.n_doors >= 64 ? ~0ull : ((1ull << sp.n_doors) - 1ull);
      const uint64_t crowded = c0 & c1 & all;
      for (uint64_t m = (v.dopen | crowded) & all; m; m &= m - 1) {
        const int d = ctz64(m);
        if (!((crowded >> d) & 1)) {
          uint8_t t = v.at(st.door_timer, d);
          if (t) v.at(st.door_timer, d) = (uint8_t)(t - 1);
          else v.dopen &= ~(1ull << d);
        } else {
          v.at(st.door_timer, d) = DOOR_INTERVAL;
        }
      }
    } else if (op == MFG_R_MOVE_MAINTAINERS) {
      for (int k = 0; k < sp.n_maint; ++k) {
        int code = io.maint_act ? (int)io.maint_act[(size_t)e * sp.n_maint + k] : maint_policy<AMAX, SpecT>(v, k, (uint32_t)step);
        step_sync<SYNC && MFG_STEP_SYNC_FINE>();
        uint16_t p = v.at(st.maint_pos, k);
        if (code < 8) {
          uint16_t t;
          if (v.try_move(p, code, false, t)) {
            v.l_del(C_MAINT, k, k, p);
            v.set_listed(C_MAINT, k, false);
            v.at(st.maint_pos, k) = t;
            v.l_add(C_MAINT, k, k, t);
          }
        } else if (code == MFG_MAINT_DOORUSE) {
          v.toggle_near(p);
        }
      }
    } else if (op == MFG_R_RESPAWN_DIRT) {                      // clean_up/rules.py:49-59
      int16_t next = v.at(st.dirt_next_spawn, 0);
      if (next < 0) {
      } else if (next == 0) {
        const int n_resp = (int)P[1];
        const double amt = P[2];
        if (io.respawn_n) {
          int n = io.respawn_n[e];
          if (n > n_resp) n = n_resp;
          const uint16_t* tiles = io.respawn_pos + (size_t)e * RESPAWN_TAPE_W;
          v.dirt_spawn(n, [&](int j) { return tiles[j]; }, [&](int) { return amt; });
        } else {
          Philox rng;
          rng.init(sp.seed, (uint64_t)(tb.env_id_offset + e), RS_RESPAWN, v.at(st.episode, 0), (uint32_t)step);
          int n_new = (int)fabs((double)n_resp + rng.uniform(-sp.dirt_n_var, sp.dirt_n_var));
          if (n_new > MFG_MAX_DIRT) n_new = MFG_MAX_DIRT;
          uint16_t chosen[MFG_MAX_DIRT];
          int got = 0;
          for (int j = 0; j < n_new; ++j) {
            uint16_t q = v.sample_free(rng, chosen, got, false);
            if (q == NO_POS) break;
            chosen[got++] = q;
          }
          int n = got < n_resp ? got : n_resp;
          v.dirt_spawn(n, [&](int j) { return chosen[j]; }, [&](int) { return amt; });
        }
        v.at(st.dirt_next_spawn, 0) = (int16_t)P[0];
      } else {
        v.at(st.dirt_next_spawn, 0) = (int16_t)(next - 1);
      }
    } else if (op == MFG_R_BATTERY_DECHARGE || op == MFG_R_DONE_BATTERY) {   // batteries/rules.py:50-63
#pragma unroll
      for (int i = 0; i < AMAX; ++i) {
        if (i >= A) continue;
        double cost = P[0];
        if (P[5] != 0) {                      // per_action_costs as a dict: keyed by the class of the action taken this tick
          int a = act[i];
          if (a < 0 || a >= sp.n_actions[i]) a = 0;
          cost = sp.act_cost[i][((skipmask >> i) & 1u) ? sp.n_actions[i] : a];          // a paralysed agent's default state is a 'Noop'
        }
        if (bat[i] != 0) bat[i] = fmax(0.0, cost + bat[i]);
      }
    } else if (op == MFG_R_DEST_REACH_REWARD || op == MFG_R_DONE_DEST) {     // destinations/rules.py:34-54
      uint32_t reached = v.at(st.dest_reached, 0);
      for (int k = 0; k < sp.n_dest; ++k) {
        if ((reached >> k) & 1) continue;
        uint16_t q = v.at(st.dest_pos, k);
        int last = -1;
        uint32_t best = 0;
#pragma unroll
        for (int i = 0; i < AMAX; ++i) {
          if (i < A && v.apos[i] == q) {
            uint32_t s = v.at(st.astamp, i);
            if (last < 0 || s > best) { last = i; best = s; }
          }
        }
        // a bound destination (SpawnDestinationsPerAgent / OnAgent) is only reached by its own agent; the reward still goes
        // to the last agent listed on the tile (the loop variable of destinations/rules.py:40-50)
        const int bound = sp.dest_bound[k];
        if (last >= 0 && bound >= 0) {
          bool there = false;
#pragma unroll
          for (int i = 0; i < AMAX; ++i) there |= i < A && i == bound && v.apos[i] == q;
          if (!there) last = -1;
        }
        if (last >= 0) {
          reached |= 1u << k;
#pragma unroll
          for (int i = 0; i < AMAX; ++i) if (i == last) rew[i] += P[0];
        }
      }
      v.at(st.dest_reached, 0) = reached;
    }
  }

  // ---- tick_post_step hooks (states.py:70-75)
  int n_collisions = 0;
  for (int r = 0; r < sp.n_rules; ++r) {
    step_sync<SYNC>();
    const int op = sp.rule_op[r];
    const double* P = sp.rule_param[r];
    if (op == MFG_R_WATCH_COLLISIONS) {                          // rules.py:276-306, states.py:228-238
#pragma unroll
      for (int i = 0; i < AMAX; ++i) {
        if (i < A && v.n_coll(px(v.apos[i]), py(v.apos[i])) >= 2) { rew[i] += P[0]; ++n_collisions; if (FLAGS) collmask |= 1u << i; }
      }
    } else if (op == MFG_R_BATTERY_DECHARGE || op == MFG_R_DONE_BATTERY) {   // batteries/rules.py:66-87
#pragma unroll
      for (int i = 0; i < AMAX; ++i) {
        if (i >= A) continue;
        bool discharged = bat[i] == 0;
        uint8_t f = v.at(st.aflag, i);
        if (discharged) {
#pragma unroll
          for (int j = 0; j < AMAX; ++j) if (j == i) rew[j] += P[1];
          if (P[2] != 0) f |= 1;
        }
        if ((f & 1) && !discharged) f &= ~1;
        v.at(st.aflag, i) = f;
      }
    }
  }

  // ---- on_check_done hooks (states.py:216-226)
  bool done = false;
  int reason = -1;
  for (int r = 0; r < sp.n_rules; ++r) {
    step_sync<SYNC>();
    const int op = sp.rule_op[r];
    const double* P = sp.rule_param[r];
    bool fired = false;
    if (op == MFG_R_DONE_MAX_STEPS) {
      fired = (int)P[0] <= step;
      if (fired && reason < 0) reason = MFG_ST_DONE_MAX_STEPS;
    } else if (op == MFG_R_DONE_ALL_DIRT) {
      fired = v.dirt_n == 0 && step > 0;
      if (fired) { glob += P[0]; if (reason < 0) reason = MFG_ST_DONE_ALL_DIRT; }
    } else if (op == MFG_R_DONE_BATTERY) {
      bool any = false;
#pragma unroll
      for (int i = 0; i < AMAX; ++i) any |= i < A && bat[i] == 0;
      fired = P[4] != 0 && any;
      if (fired) { glob += P[3]; if (reason < 0) reason = MFG_ST_DONE_BATTERY; }
    } else if (op == MFG_R_DONE_DEST) {
      uint32_t reached = v.at(st.dest_reached, 0);
      uint32_t all = sp.n_dest >= 32 ? 0xFFFFFFFFu : ((1u << sp.n_dest) - 1);
      int cond = (int)P[1];
      fired = cond == 0 ? reached != 0 : reached == all;
      if (fired) { glob += P[2]; if (reason < 0) reason = MFG_ST_DONE_DEST; }
      else if (cond == 2) v.at(st.dest_reached, 0) = 0;
    } else if (op == MFG_R_DONE_MAINT_COLLISION) {               // maintenance/rules.py:32-40 (group-local positions)
#pragma unroll
      for (int i = 0; i < AMAX; ++i) {
        if (i >= A) continue;
        bool hit = false;
        MFG_NOUNROLL
        for (int k = 0; k < sp.n_maint; ++k) hit |= v.at(st.maint_pos, k) == v.apos[i];
        if (hit) { fired = true; rew[i] += -5.0; }
      }
      if (fired && reason < 0) reason = MFG_ST_DONE_MAINT;
    } else if (op == MFG_R_WATCH_COLLISIONS && P[1] != 0) {      // rules.py:308-325
      bool any = n_collisions > 0;
      if (!any && sp.n_maint) {
        uint32_t l = v.at(st.maint_listed, 0);
        for (int k = 0; k < sp.n_maint; ++k) {
          uint16_t q = v.at(st.maint_pos, k);
          any |= ((l >> k) & 1) && v.n_coll(px(q), py(q)) >= 2;
        }
      }
      fired = any;
      if (fired) { glob += P[2]; if (reason < 0) reason = MFG_ST_DONE_COLLISION; }
    }
    done |= fired;
  }

  // ---- reward fold (factory.py:222-259), individual rewards: agent's own results + the global ones.  The scalar form
  // (individual_rewards: false) does not exist: the reference raises at factory.py:217 (`sum(reward)` of a float) on the
  // first step, so mfg_create rejects such a spec.
  step_sync<SYNC>();
#pragma unroll
  for (int i = 0; i < AMAX; ++i) {
    if (i < A) {
      double r = rew[i] + glob;
      io.reward[(size_t)e * A + i] = (float)r;
      epr[i] += r;
      v.at(st.ep_ret, i) = epr[i];
    }
  }
  io.done[e] = done ? 1 : 0;
  if (FLAGS && io.flags) {
    uint8_t* f = io.flags + (size_t)e * (A + 1);
#pragma unroll
    for (int i = 0; i < AMAX; ++i)
      if (i < A) f[i] = (uint8_t)((((okmask >> i) & 1u) ? MFG_FLAG_VALID : 0) | (((skipmask >> i) & 1u) ? MFG_FLAG_SKIPPED : 0) |
                                  (((collmask >> i) & 1u) ? MFG_FLAG_COLLISION : 0) |
                                  (((mcollmask >> i) & 1u) ? MFG_FLAG_MOVE_COLLISION : 0) |
                                  (((auxmask >> i) & 1u) ? MFG_FLAG_AUX_REWARD : 0));
    f[A] = (uint8_t)(done ? (reason >= 0 ? reason : 255) : 0);
  }
  if (sp.has_batteries) {
#pragma unroll
    for (int i = 0; i < AMAX; ++i) if (i < A) v.at(st.bat, i) = bat[i];
  }
  v.store();

  // ---- episode statistics + optional in-kernel auto reset
  if (n_collisions) stat_add(tb, MFG_ST_COLLISIONS, (unsigned long long)n_collisions);
  // an env that is stepped on after its episode ended (auto_reset off, like the reference's un-reset Factory) keeps
  // ticking, but its episode is counted once
  const bool first_done = done && !v.at(st.finished, 0);
  if (done) v.at(st.finished, 0) = 1;
  if (first_done) {
    stat_add(tb, MFG_ST_EPISODES, 1);
    stat_add(tb, MFG_ST_STEPS, (unsigned long long)step);
    if (reason >= 0) stat_add(tb, reason, 1);
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < AMAX; ++i) if (i < A) { double x = epr[i]; tot += x; stat_add_f64(tb, MFG_ST_RETURN_AGENT0 + i, x); }
    stat_add_f64(tb, MFG_ST_RETURN_SUM, tot);
  }
  if (done) {
    if (io.auto_reset) {
#if defined(__CUDA_ARCH__)
      if (io.reset_list) io.reset_list[atomicAdd(io.reset_count, 1u)] = (uint32_t)v.eg;
      else
#endif
      {
        // copies: the call must not make the caller's State / Tables escape (they would then live in local memory and
        // every field access of the hot path would become a dependent local load)
        State st2 = st;
        Tables tb2 = tb;
#if defined(__CUDA_ARCH__)
        if constexpr (Env<AMAX, SpecT>::UIDG) return;     // split image: launched with the deferred reset list only
        if constexpr (Env<AMAX, SpecT>::STAGED) {       // offsets of the staged view -> generic pointers into the image
          char* g = reinterpret_cast<char*>(__cvta_shared_to_generic(v.sbase));
#define F(type, name, rows_expr) \
  if constexpr (!std::is_same<type, double>::value) st2.name = reinterpret_cast<type*>(g + reinterpret_cast<uintptr_t>(st.name));
          MFG_STATE_FIELDS(F)
#undef F
          st2.base_i = g;
          tb2.wall = reinterpret_cast<const uint8_t*>(g + reinterpret_cast<uintptr_t>(tb.wall));
          tb2.door_map = reinterpret_cast<const uint8_t*>(g + reinterpret_cast<uintptr_t>(tb.door_map));
          tb2.door_pos = reinterpret_cast<const uint16_t*>(g + reinterpret_cast<uintptr_t>(tb.door_pos));
        }
#endif
        env_reset<AMAX>(full, tb2, st2, v.e, v.at(st.episode, 0) + 1, v.eg);
      }
    }
  }
}

// ================================================================================================
// observation, one (env, agent) per thread, all parity modes (the tiled fast path lives in mfg_kernels.cu)
// ================================================================================================
// first-visit ranks of the cells of the ray-radius box.  POMDP: radius = window diameter <= 7 (225 cells); full
// observability (pomdp_r == 0): radius = min(H, W) <= MFG_MAX_RAY_LEN - 1 = 15 (961 cells).
typedef uint16_t rank_t;
constexpr int RANK_INF = 0xFFFF;
constexpr int RANK_RMAX = MFG_MAX_RAY_LEN - 1;
constexpr int RANK_CELLS = (2 * RANK_RMAX + 1) * (2 * RANK_RMAX + 1);

// observation geometry (observation_builder.py:51, 156-160, 244): egocentric (2r+1)^2 window, or the whole level
MFG_HD bool obs_full(const MfgSpec& sp) { return sp.pomdp_r == 0; }
MFG_HD int obs_plane_cells(const MfgSpec& sp) { return obs_full(sp) ? sp.H * sp.W : (2 * sp.pomdp_r + 1) * (2 * sp.pomdp_r + 1); }
MFG_HD int obs_ray_radius(const MfgSpec& sp) { return obs_full(sp) ? (sp.H < sp.W ? sp.H : sp.W) : 2 * sp.pomdp_r + 1; }

// ray walk (ray_caster.py:81-103): first-visit order of the cells of the radius box around (ax, ay); rays in the reference's
// order, a ray stops at a light-blocking cell or between two diagonal blockers
template <typename BlocksLight>
MFG_HD void ray_walk(const MfgSpec& sp, int ax, int ay, int R, int BW, BlocksLight blocks_light, rank_t* rank) {
  for (int i = 0; i < BW * BW; ++i) rank[i] = RANK_INF;
  int visit = 0;
  for (int ray = 0; ray < sp.n_rays; ++ray) {
    int pxr = ax, pyr = ay;
    for (int s = 0; s < sp.ray_len[ray]; ++s) {
      int dx = sp.ray_dx[ray][s], dy = sp.ray_dy[ray][s];
      int x = ax + dx, y = ay + dy;
      int cx = x - pxr, cy = y - pyr;
      bool hits = blocks_light(x, y);
      bool diag = (cx != 0 && cy != 0) && blocks_light(x, y - cy) && blocks_light(x - cx, y);
      if (!diag) {
        rank_t& rk = rank[(dx + R) * BW + (dy + R)];
        if (rk == RANK_INF) rk = (rank_t)visit++;
      }
      if (hits || diag) break;
      pxr = x; pyr = y;
    }
  }
}

template <int AMAX>
struct ObsCtx {
  Env<AMAX>& v;
  int a, ax, ay, r, D, R, BW;
  const rank_t* rank;                     // [(2R+1)^2] first-visit order of every cell of the radius box, RANK_INF = unseen
  MFG_HD ObsCtx(Env<AMAX>& v_, const rank_t* rank_) : v(v_), rank(rank_) {}
  MFG_HD bool blocks_light(int x, int y) const {
    if (!v.in_grid(x, y)) return false;
    return v.tbl(v.tb.wall, x * v.sp.W + y) || v.closed_listed_door(x, y);
  }
  MFG_HD int rank_of(uint16_t p) const {
    if (p == NO_POS) return RANK_INF;
    int dx = px(p) - ax, dy = py(p) - ay;
    if (dx < -R || dx > R || dy < -R || dy > R) return RANK_INF;
    return rank[(dx + R) * BW + (dy + R)];
  }
};

// Does some OTHER visible listed entity with the same uid precede (cls, idx) in first-visit order?
// (observation_builder.py:155 `set(visible_entities)` with uid equality)
template <int AMAX>
MFG_HD bool uid_shadowed(const ObsCtx<AMAX>& o, int cls, int idx, int uid, int my_rank) {
  const Env<AMAX>& v = o.v; const MfgSpec& sp = v.sp; const State& st = v.st; const Tables& tb = v.tb;
  if (uid < sp.n_walls && o.rank_of(tb.wall_pos[uid]) < my_rank) return true;      // walls are always listed
  if (uid < sp.n_doors && cls != C_DOOR && ((v.dlisted >> uid) & 1) && o.rank_of(v.tbl(tb.door_pos, uid)) < my_rank) return true;
  if (sp.has_dirt && uid < (int)v.at(st.dirt_next_uid, 0)) {
    for (int k = 0; k < v.dirt_end; ++k) {
      if (cls == C_DIRT && k == idx) continue;
      if (v.uid_at(k) == uid && ((v.dirt_listed >> k) & 1) && o.rank_of(v.at(st.dirt_pos, k)) < my_rank) return true;
    }
  }
  for (int c = C_ITEM; c <= C_MAINT; ++c) {
    if (c == cls || uid >= v.cls_count(c)) continue;
    if (((v.at(v.cls_listed(c), 0) >> uid) & 1) && o.rank_of(v.at(v.cls_pos(c), uid)) < my_rank) return true;
  }
  return false;
}

// What an entity contributes: an integer-valued encoding (stacks add up exactly), a door (open / closed encoding),
// a dirt pile (f64 amount of slot `aux`), or a directly stored scalar.
enum { OK_INT = 0, OK_STORE = 1, OK_DOOR = 2, OK_DIRT = 3 };

// Exact observation of one agent in every parity mode (observation_builder.py:138-220 + ray_caster.py:66-104):
// full radius rays in the reference's order, first-visit rank per cell, uid shadowing when spec.faithful.
// The result goes to a Sink:  sink.wall(cell)                       visible, unshadowed wall on a window cell
//                             sink.ent(chmask, cell, kind, aux, v)  entity contribution to every channel in chmask
//                             sink.scalar(channel, flat_index, v)   Battery / GlobalPosition values
// Call order: walls, agents, integer-valued groups, doors, dirt, scalars (fractional encodings last).
template <int AMAX, typename Sink>
MFG_HDN void obs_agent_exact(const MfgSpec& sp, const Tables& tb, const State& st, int64_t e, int a, rank_t* rank,
                             Sink& sink, int64_t eg = -1) {
  Env<AMAX> v(sp, tb, st, e, eg);      // eg: global env id when `st` is a staged image (f64 fields are never staged)
  v.load();
  ObsCtx<AMAX> o(v, rank);
  const bool full = obs_full(sp);
  const int r = sp.pomdp_r, D = 2 * r + 1, R = obs_ray_radius(sp), BW = 2 * R + 1;
  o.a = a; o.ax = px(v.apos[a]); o.ay = py(v.apos[a]); o.r = r; o.D = D; o.R = R; o.BW = BW;
  // ---- first-visit order of the cells of the radius box: from the per-(tile, closed door subset) table when there is one,
  // else the ray walk
  if (tb.rank_tab) {
    const uint32_t nsub = 1u << sp.n_doors;
    const uint32_t sub = (uint32_t)(~v.dopen & v.dlisted) & (nsub - 1u);
    o.rank = tb.rank_tab + ((size_t)(o.ax * sp.W + o.ay) * nsub + sub) * (size_t)(BW * BW);
  } else {
    ray_walk(sp, o.ax, o.ay, R, BW, [&](int x, int y) { return o.blocks_light(x, y); }, rank);
  }
  const rank_t* const rankv = o.rank;

  const uint32_t* chm = sp.term_chmask[a];
  auto in_window = [&](uint16_t p, int& cell) {
    if (full) { cell = px(p) * sp.W + py(p); return true; }          // the plane is the whole level, absolute coordinates
    int dx = px(p) - o.ax + r, dy = py(p) - o.ay + r;
    if (dx < 0 || dy < 0 || dx >= D || dy >= D) return false;
    cell = dx * D + dy;
    return true;
  };

  // ---- walls (uid = row-major wall index)
  int max_small = 0;
  for (int c = C_ITEM; c <= C_MAINT; ++c) max_small = v.cls_count(c) > max_small ? v.cls_count(c) : max_small;
  if (chm[MFG_G_WALLS]) {
    const int wr = full ? R : r;                 // walls can only be seen inside the ray radius
    for (int dx = -wr; dx <= wr; ++dx) for (int dy = -wr; dy <= wr; ++dy) {
      int x = o.ax + dx, y = o.ay + dy;
      if (!v.in_grid(x, y) || !v.tbl(tb.wall, x * sp.W + y)) continue;
      int rk = rankv[(dx + R) * BW + (dy + R)];
      if (rk == RANK_INF) continue;
      if (sp.faithful) {
        // walls are never shadowed by other walls; only a dynamic entity with the same uid seen earlier hides it
        int uid = tb.wall_uid[x * sp.W + y];
        bool sh = false;
        if (uid < sp.n_doors && ((v.dlisted >> uid) & 1) && o.rank_of(v.tbl(tb.door_pos, uid)) < rk) sh = true;
        if (!sh && sp.has_dirt && uid < (int)v.at(st.dirt_next_uid, 0))
          for (int k = 0; k < v.dirt_end && !sh; ++k)
            sh = v.uid_at(k) == uid && ((v.dirt_listed >> k) & 1) && o.rank_of(v.at(st.dirt_pos, k)) < rk;
        if (uid < max_small)                  // (most walls have a larger uid than any small group has members)
          for (int c = C_ITEM; c <= C_MAINT && !sh; ++c)
            sh = uid < v.cls_count(c) && ((v.at(v.cls_listed(c), 0) >> uid) & 1) && o.rank_of(v.at(v.cls_pos(c), uid)) < rk;
        if (sh) continue;
      }
      sink.wall(full ? x * sp.W + y : (dx + r) * D + (dy + r));
    }
  }
  // ---- agents (string identifiers: never shadowed)
  for (int j = 0; j < v.A; ++j) {
    int cell;
    if (!chm[MFG_G_AGENT0 + j] || !in_window(v.apos[j], cell)) continue;
    if (o.rank_of(v.apos[j]) == RANK_INF) continue;
    sink.ent(chm[MFG_G_AGENT0 + j], cell, OK_INT, 0, 1.0);
  }
  // ---- small groups with constant encodings
  const int term_of[8] = {MFG_G_DOORS, MFG_G_DIRT, MFG_G_ITEMS, MFG_G_PODS, MFG_G_DEST, MFG_G_DROPOFF, MFG_G_MACHINES, MFG_G_MAINT};
  for (int c = C_ITEM; c <= C_MAINT; ++c) {
    uint32_t m = chm[term_of[c]];
    int n = v.cls_count(c);
    if (!m || !n) continue;
    uint32_t l = v.at(v.cls_listed(c), 0);
    for (int k = 0; k < n; ++k) {
      uint16_t p = v.at(v.cls_pos(c), k);
      int cell;
      if (!((l >> k) & 1) || p == NO_POS || !in_window(p, cell)) continue;
      int rk = o.rank_of(p);
      if (rk == RANK_INF) continue;
      if (sp.faithful && uid_shadowed(o, c, k, k, rk)) continue;
      if (c == C_DEST && ((v.at(st.dest_reached, 0) >> k) & 1)) continue;         // a reached destination encodes as 0
      sink.ent(m, cell, OK_INT, 0, c == C_MACH ? ENC_MACHINE : 1.0);
    }
  }
  // ---- doors, then dirt (fractional encodings go last so that integer stacks are summed exactly first)
  if (chm[MFG_G_DOORS]) {
    for (int d = 0; d < sp.n_doors; ++d) {
      uint16_t p = v.tbl(tb.door_pos, d);
      int cell;
      if (!((v.dlisted >> d) & 1) || !in_window(p, cell)) continue;
      int rk = o.rank_of(p);
      if (rk == RANK_INF) continue;
      if (sp.faithful && uid_shadowed(o, C_DOOR, d, d, rk)) continue;
      const bool open = (v.dopen >> d) & 1;
      sink.ent(chm[MFG_G_DOORS], cell, OK_DOOR, open ? 1 : 0, open ? ENC_DOOR_OPEN : ENC_DOOR_CLOSED);
    }
  }
  if (sp.has_dirt && chm[MFG_G_DIRT]) {
    for (int k = 0; k < v.dirt_end; ++k) {
      uint16_t p = v.at(st.dirt_pos, k);
      int cell;
      if (p == NO_POS || !((v.dirt_listed >> k) & 1) || !in_window(p, cell)) continue;
      int rk = o.rank_of(p);
      if (rk == RANK_INF) continue;
      if (sp.faithful && uid_shadowed(o, C_DIRT, k, v.uid_at(k), rk)) continue;
      sink.ent(chm[MFG_G_DIRT], cell, OK_DIRT, k, v.at(st.dirt_amt, k));
    }
  }
  // ---- scalar channels (observation_builder.py:205-218, entity/util.py:56-66)
  const int C = sp.n_channels[a];
  for (int c = 0; c < C; ++c) {
    int kind = sp.ch_kind[a][c];
    if (kind == MFG_CH_BATTERY) sink.scalar(c, 0, (float)v.at(st.bat, a));
    else if (kind == MFG_CH_GLOBALPOS) {
      sink.scalar(c, 0, (float)((double)o.ax / (double)sp.H));
      sink.scalar(c, 1, (float)((double)o.ay / (double)sp.W));
    }
  }
}

// Sink that writes the agent's planes directly: out = [C_a][D*D] floats, contiguous
struct FloatSink {
  float* out;
  int DD;
  uint32_t wall_mask;
  MFG_HD void add(uint32_t mask, int cell, double val) {
    while (mask) {
      const int c = ctz64(mask);
      mask &= mask - 1;
      float& f = out[c * DD + cell];
      f = (float)((double)f + val);
    }
  }
  MFG_HD void wall(int cell) { add(wall_mask, cell, 1.0); }
  MFG_HD void ent(uint32_t mask, int cell, int, int, double val) { add(mask, cell, val); }
  MFG_HD void scalar(int c, int flat, float v) { out[c * DD + flat] = v; }
};

template <int AMAX>
MFG_HDN void obs_agent_direct(const MfgSpec& sp, const Tables& tb, const State& st, int64_t e, int a, float* out,
                              int64_t eg = -1, bool zero = true) {
  const int DD = obs_plane_cells(sp);
  const int C = sp.n_channels[a];
  if (zero) for (int i = 0; i < C * DD; ++i) out[i] = 0.0f;      // (false: the caller has cleared the planes)
  rank_t rank[RANK_CELLS];
  FloatSink sink{out, DD, sp.term_chmask[a][MFG_G_WALLS]};
  obs_agent_exact<AMAX>(sp, tb, st, e, a, rank, sink, eg);
}

}  // namespace mfg
