/*
 * mfg_b200.h - C ABI of the B200-native batched stepping engine for marl-factory-grid.
 *
 * The reference is pure Python and has no FFI; this header is the drop-in boundary a binding
 * (ctypes / cffi / pybind) attaches to.  Each entry point replaces one method of the reference's
 * gym-style environment, batched over N independent environments:
 *
 *   mfg_create / mfg_destroy   <- Factory.__init__ / close      marl_factory_grid/environment/factory.py:81-129
 *   mfg_reset                  <- Factory.reset                 factory.py:134-148 (+ rules.py:163-199 spawn rules)
 *   mfg_step                   <- Factory.step -> Gamestate.tick factory.py:189-220, utils/states.py:170-226
 *   mfg_observe                <- OBSBuilder.build_for_all      utils/observation_builder.py:98-235, ray_caster.py:66-199
 *   mfg_step_observe           <- step + observation in one call (what Factory.step returns)
 *   mfg_stats                  <- EnvMonitor-style episode statistics   utils/logging/envmonitor.py:28-56
 *   mfg_random_actions         <- the `action_space.sample()` loop of random_testrun.py:44-56
 *
 * Conventions
 *   - plain pointers and sizes only; every `void* stream` is a cudaStream_t (0 = default stream).
 *   - all pointers named d_* are DEVICE pointers owned by the caller (torch tensors on the host side);
 *     the handle owns only the constant spec tables and the small statistics vector.
 *   - calls are asynchronous on the given stream; they return 0 on success, a negative MFG_E_* code
 *     otherwise (never exceptions, never exit()); mfg_last_error() returns a thread-local message.
 *   - one host thread per handle; handles are independent (re-entrant across handles).
 *   - positions are packed as pos16 = (row << 8) | column; MFG_NO_POS marks "not on the map".
 */
#ifndef MFG_B200_H
#define MFG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFG_MAX_AGENTS 16
#define MFG_MAX_ACTIONS 32
#define MFG_MAX_DOORS 64
#define MFG_MAX_DIRT 64
#define MFG_MAX_RULES 32
#define MFG_MAX_CHANNELS 32
#define MFG_MAX_SMALL 32
#define MFG_MAX_GROUPS 16
#define MFG_MAX_FIXED 32
#define MFG_N_TERMS (9 + MFG_MAX_AGENTS)
#define MFG_RULE_NPARAM 6
#define MFG_MAX_RAYS 128
#define MFG_MAX_RAY_LEN 16
#define MFG_NO_POS 0xFFFFu
#define MFG_N_STATS 32

enum { MFG_OK = 0, MFG_E_INVALID = -1, MFG_E_CUDA = -2, MFG_E_NOMEM = -3, MFG_E_UNSUPPORTED = -4 };

/* action opcodes (environment/actions.py, modules/<m>/actions.py) */
enum { MFG_OP_NOOP = 0, MFG_OP_MOVE, MFG_OP_DOORUSE, MFG_OP_CLEAN, MFG_OP_ITEM, MFG_OP_CHARGE, MFG_OP_DEST, MFG_OP_MACHINE };
/* observation terms: positional groups, then one term per agent */
enum { MFG_G_WALLS = 0, MFG_G_DOORS, MFG_G_DIRT, MFG_G_ITEMS, MFG_G_DROPOFF, MFG_G_PODS, MFG_G_DEST, MFG_G_MACHINES,
       MFG_G_MAINT, MFG_G_AGENT0 };
/* spawn-program group ids (Entities yaml order) */
enum { MFG_SP_DOORS = 0, MFG_SP_DIRT, MFG_SP_BATTERIES, MFG_SP_PODS, MFG_SP_DEST, MFG_SP_ITEMS, MFG_SP_INVENTORIES,
       MFG_SP_DROPOFF, MFG_SP_MACHINES, MFG_SP_MAINT, MFG_SP_GLOBALPOS };
enum { MFG_CH_TERMS = 0, MFG_CH_ZERO, MFG_CH_BATTERY, MFG_CH_GLOBALPOS };
/* rule opcodes (environment/rules.py, modules/<m>/rules.py) */
enum { MFG_R_WATCH_COLLISIONS = 0, MFG_R_RESPAWN_DIRT, MFG_R_SMEAR_DIRT, MFG_R_DOOR_AUTO_CLOSE, MFG_R_DONE_ALL_DIRT,
       MFG_R_BATTERY_DECHARGE, MFG_R_DONE_BATTERY, MFG_R_DEST_REACH_REWARD, MFG_R_DONE_DEST, MFG_R_RESPAWN_ITEMS,
       MFG_R_MOVE_MAINTAINERS, MFG_R_DONE_MAINT_COLLISION, MFG_R_DONE_MAX_STEPS, MFG_R_RANDOM_INITIAL_STEPS };
/* how Destinations are spawned (modules/destinations/rules.py:95-162) */
enum { MFG_DEST_FREE = 0, MFG_DEST_ON_AGENT = 1, MFG_DEST_PER_AGENT = 2 };
/* maintainer tape codes: 0..7 move direction (Move8 order), then */
enum { MFG_MAINT_NOOP = 8, MFG_MAINT_DOORUSE = 9, MFG_MAINT_MACHINE = 10 };

/* per-agent result bits of one step (mfg_bind_step_flags): what the reference reports through `info` / `agent.state`
 * (utils/results.py:42-52, factory.py:222-239): the action was valid, the agent was paralysed and skipped, the agent stands on a
 * collision tile (WatchCollisions), the agent's move "introduced a collision" (actions.py:80-96: it failed, or ended on a tile
 * shared with another collidable entity), the action paid its auxiliary reward (ItemAction on a drop-off, items/actions.py:41-63) */
enum { MFG_FLAG_VALID = 1, MFG_FLAG_SKIPPED = 2, MFG_FLAG_COLLISION = 4, MFG_FLAG_MOVE_COLLISION = 8, MFG_FLAG_AUX_REWARD = 16 };

/* indices into the statistics vector returned by mfg_stats (int64 counters, doubles bit-cast for sums) */
enum { MFG_ST_EPISODES = 0, MFG_ST_STEPS, MFG_ST_DONE_MAX_STEPS, MFG_ST_DONE_ALL_DIRT, MFG_ST_DONE_BATTERY,
       MFG_ST_DONE_DEST, MFG_ST_DONE_MAINT, MFG_ST_DONE_COLLISION, MFG_ST_COLLISIONS, MFG_ST_DIRT_OVERFLOW,
       MFG_ST_SPAWN_FAIL, MFG_ST_RETURN_SUM /* f64 bits, sum over agents */, MFG_ST_RETURN_AGENT0 /* f64 bits, 16 slots */ };

/* Compiled environment description (host memory; copied by mfg_create).  Produced from the reference's
 * yaml + level .txt by the host-side config compiler. */
typedef struct MfgSpec {
  int32_t H, W, pomdp_r, n_agents;
  int32_t individual_rewards;     /* must be 1: the reference raises TypeError (factory.py:217) on the first step otherwise */
  int32_t faithful;               /* 1: reproduce the reference's uid-equality artefact */
  int32_t n_floor, n_doors, n_walls;
  int32_t has_dirt, dirt_slots, dirt_quantity;
  int32_t has_batteries, has_globalpos;
  int32_t n_items, n_dropoff, n_pods, n_dest, n_machines, n_maint;
  int32_t n_rules, n_groups, n_rays;
  int32_t reserved0;
  double dirt_initial_amount, dirt_clean_amount, dirt_max_global, dirt_n_var, dirt_amount_var;
  double battery_initial;
  uint64_t seed;
  /* agents: action tables (actions.py:18-36 valid/fail reward, ItemAction aux = failed drop-off reward) */
  int32_t n_actions[MFG_MAX_AGENTS];
  int32_t agent_blocking[MFG_MAX_AGENTS];
  int32_t agent_n_fixed[MFG_MAX_AGENTS];
  uint16_t agent_fixed_pos[MFG_MAX_AGENTS][MFG_MAX_FIXED];
  int32_t act_opcode[MFG_MAX_AGENTS][MFG_MAX_ACTIONS];
  int32_t act_dir[MFG_MAX_AGENTS][MFG_MAX_ACTIONS];
  double act_valid[MFG_MAX_AGENTS][MFG_MAX_ACTIONS];
  double act_fail[MFG_MAX_AGENTS][MFG_MAX_ACTIONS];
  double act_aux[MFG_MAX_AGENTS][MFG_MAX_ACTIONS];
  /* observation channel program (observation_builder.py:164-220, 237-277) */
  int32_t n_channels[MFG_MAX_AGENTS];
  int32_t ch_offset[MFG_MAX_AGENTS];                       /* first channel of the agent in the packed tensor */
  int32_t ch_kind[MFG_MAX_AGENTS][MFG_MAX_CHANNELS];
  uint32_t term_chmask[MFG_MAX_AGENTS][MFG_N_TERMS];       /* channels (bits) an entity of that term adds to */
  /* ordered rule program (yaml order) */
  int32_t rule_op[MFG_MAX_RULES];
  double rule_param[MFG_MAX_RULES][MFG_RULE_NPARAM];
  /* spawn program (Entities yaml order) */
  int32_t group_id[MFG_MAX_GROUPS];
  int32_t group_quantity[MFG_MAX_GROUPS];
  int32_t group_n_fixed[MFG_MAX_GROUPS];
  uint16_t group_fixed_pos[MFG_MAX_GROUPS][MFG_MAX_FIXED];
  /* full visibility rays, in the reference's ray order, origin cell first (ray_caster.py:34-49, 143-199) */
  int32_t ray_len[MFG_MAX_RAYS];
  int8_t ray_dx[MFG_MAX_RAYS][MFG_MAX_RAY_LEN];
  int8_t ray_dy[MFG_MAX_RAYS][MFG_MAX_RAY_LEN];
  /* bound destinations: SpawnDestinationOnAgent / SpawnDestinationsPerAgent (modules/destinations/rules.py:95-162) */
  int32_t dest_mode;                                       /* MFG_DEST_* */
  int32_t random_initial_steps;                            /* DoRandomInitialSteps.random_steps (environment/rules.py:328-355), 0 = rule absent */
  int32_t dest_bound[MFG_MAX_SMALL];                       /* agent a destination is bound to, or -1 */
  int32_t dest_n_cand[MFG_MAX_SMALL];                      /* per-agent mode: candidate tiles (0 = any floor tile) */
  uint16_t dest_cand[MFG_MAX_SMALL][MFG_MAX_FIXED];
  /* BatteryDecharge.per_action_costs given as a dict (batteries/rules.py:50-63): cost by the class name of the action taken;
   * used when rule_param[r][5] != 0, act_cost[a][n_actions[a]] = the 'Noop' entry (a paralysed agent's default state) */
  double act_cost[MFG_MAX_AGENTS][MFG_MAX_ACTIONS + 1];
  /* level tables (host pointers) */
  const uint8_t* walls;           /* [H*W] 1 = wall */
  const uint16_t* floor_pos;      /* [n_floor] pos16, row-major */
  const uint16_t* door_pos;       /* [n_doors] pos16, row-major */
  const uint8_t* nexthop;         /* [n_floor*n_floor] direction 0..7 from floor i towards floor j (255 = none), or NULL */
} MfgSpec;

/* Replay inputs for one step (device pointers, env-major).  NULL members fall back to the engine's Philox draws. */
typedef struct MfgTape {
  const uint8_t* d_maint_action;   /* [N][n_maint]  MFG_MAINT_* code per maintainer */
  const int8_t* d_respawn_n;       /* [N]           number of proposed tiles if RespawnDirt fires this step */
  const uint16_t* d_respawn_pos;   /* [N][8]        proposed free tiles (pos16) */
} MfgTape;

typedef struct MfgField {
  size_t offset;                   /* byte offset of the field's [rows][128] slab of env block 0 in the state buffer */
  int32_t rows;
  int32_t elem_size;               /* bytes per element */
  size_t block_bytes;              /* distance between the slabs of consecutive 128-env blocks */
} MfgField;                        /* element (row r, env e) lives at offset + (e / 128) * block_bytes + (r * 128 + e % 128) * elem_size */

typedef struct MfgHandle MfgHandle;

int mfg_create(const MfgSpec* spec, int64_t n_envs, int64_t env_id_offset, MfgHandle** out);
void mfg_destroy(MfgHandle* h);
const char* mfg_last_error(void);
const char* mfg_version(void);

/* State lives in ONE caller-owned device buffer of mfg_state_bytes(h) bytes:
 * BLOCKED struct-of-arrays: envs are grouped in blocks of 128; a block stores every row of every integer / byte field
 * back to back as [rows][128] slabs (one contiguous range => one TMA bulk copy stages it), a second region stores the
 * f64 fields the same way.  Within a slab the env index is the fastest one (coalesced per warp). */
size_t mfg_state_bytes(const MfgHandle* h);
int mfg_state_field(const MfgHandle* h, const char* name, MfgField* out);
int mfg_bind_state(MfgHandle* h, void* d_state);

/* Factory.reset: spawn every (masked) env from the engine's counter-based Philox streams. d_env_mask may be NULL. */
int mfg_reset(MfgHandle* h, const uint8_t* d_env_mask, void* stream);
/* Gamestate.tick + check_done + reward fold.  d_actions [N][A] int32, d_reward [N][A] float,
 * d_done [N] uint8.  auto_reset != 0 re-spawns finished envs in the same launch. */
int mfg_step(MfgHandle* h, const int32_t* d_actions, const MfgTape* tape, float* d_reward, uint8_t* d_done,
             int auto_reset, void* stream);
/* OBSBuilder.build_for_all (observation_builder.py:98-235): packed observation tensor [N][sum(C_a)][D][D] float32. */
int mfg_observe(MfgHandle* h, float* d_obs, void* stream);
/* Factory.step as one call (factory.py:189-220: tick, done, reward fold, observations).  With auto_reset != 0 the finished
 * envs are re-spawned and observed on an internal high-priority stream while the tiled observation kernel covers all
 * other envs on `stream`; the call joins before returning control to `stream`, so the caller sees plain stream order. */
int mfg_step_observe(MfgHandle* h, const int32_t* d_actions, const MfgTape* tape, float* d_reward, uint8_t* d_done,
                     float* d_obs, int auto_reset, void* stream);
/* uniform random actions in [0, n_actions[a]) from Philox (seed, global env id, step_index). */
int mfg_random_actions(MfgHandle* h, int32_t* d_actions, uint64_t seed, uint64_t step_index, void* stream);
/* Host-buffer convenience path (what a caller holding numpy arrays uses): H2D actions, step+observe, D2H results. */
int mfg_step_host(MfgHandle* h, const int32_t* h_actions, float* h_reward, uint8_t* h_done, float* h_obs,
                  int auto_reset, void* stream);
/* Optional per-step result flags: d_flags [N][A + 1] uint8, written by every following mfg_step / mfg_step_observe: per agent the
 * MFG_FLAG_* bits, then the done reason of the env (0 = not done, else the MFG_ST_DONE_* index of the first rule that fired,
 * 255 = other).  NULL switches the output off (default).  This is the batched form of the reference's per-step `info` dict. */
int mfg_bind_step_flags(MfgHandle* h, uint8_t* d_flags);
/* copies the MFG_N_STATS int64 statistics vector (device) into d_out; zero_after != 0 clears it afterwards */
int mfg_stats(MfgHandle* h, int64_t* d_out, int zero_after, void* stream);
/* options: "obs_kernel" (0 auto, 1 exact per-agent kernel over block-staged state, 2 tiled shared-memory kernel, 3 exact kernel on plain global state), "obs_store" (1 TMA bulk store),
 * "obs_cap" (sprite slots per env), "step_kernel" (k_step launch shape: 1 barriers at the convergent points + dirt uids left in HBM, 2 barriers
 * only, 0 neither), "step_blocks" (state blocks per k_step CTA, 0 auto), "defer_reset" (1 packed reset kernel), "overlap_reset" (1 side stream in
 * mfg_step_observe), "timing" (1: CUDA event pairs around the kernels, read with mfg_get_info "step_ns" / "obs_ns" /
 * "reset_ns").  info: "launches", "tiled_ok", "obs_smem", "obs_threads", "obs_ctas_per_sm", "obs_cap", "obs_cap_max". */
int mfg_set_option(MfgHandle* h, const char* name, int64_t value);
int64_t mfg_get_info(const MfgHandle* h, const char* name);

#ifdef __cplusplus
}
#endif
#endif /* MFG_B200_H */
