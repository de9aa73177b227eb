// mfg_host.hpp - host-side helpers shared by the CUDA library (mfg_abi.cu) and the test-only host
// build (tests/hostsim): state layout computation and the derived level tables.
#pragma once
#include <cstring>
#include <string>
#include <vector>
#include "mfg_core.cuh"

namespace mfg {

struct FieldInfo {
  const char* name;
  size_t offset;        // byte offset of the field's slab inside block 0 (relative to the buffer start)
  int rows;
  int elem_size;
  size_t block_bytes;   // distance between the slabs of consecutive 128-env blocks
};

struct Layout {
  std::vector<FieldInfo> fields;
  size_t blk_i = 0, blk_f = 0;      // bytes per block: integer region / f64 region
  size_t off_f = 0;                 // start of the f64 region
  size_t total = 0;
  int64_t n_blocks = 0;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// blocked layout (see State in mfg_core.cuh): [block][field][row][128 envs], integer region first, f64 region second
inline Layout compute_layout(const MfgSpec& sp, int64_t N) {
  Layout L;
  L.n_blocks = (N + ENV_BLOCK - 1) / ENV_BLOCK;
  size_t oi = 0, of = 0;
#define F(type, name, rows_expr)                                                                      \
  {                                                                                                   \
    const int rows = (int)(rows_expr);                                                                \
    const bool f64 = std::is_same<type, double>::value;                                               \
    L.fields.push_back(FieldInfo{#name, f64 ? of : oi, rows, (int)sizeof(type), 0});                  \
    (f64 ? of : oi) += (size_t)rows * ENV_BLOCK * sizeof(type);                                       \
  }
  MFG_STATE_FIELDS(F)
#undef F
  L.blk_i = oi;                      // every slab is a multiple of 128 bytes => blocks stay 128-byte aligned
  L.blk_f = of;
  L.off_f = align_up((size_t)L.n_blocks * L.blk_i, 256);
  L.total = align_up(L.off_f + (size_t)L.n_blocks * L.blk_f, 256);
  size_t i = 0;
#define F(type, name, rows_expr)                                                                      \
  {                                                                                                   \
    const bool f64 = std::is_same<type, double>::value;                                               \
    if (f64) L.fields[i].offset += L.off_f;                                                           \
    L.fields[i].block_bytes = f64 ? L.blk_f : L.blk_i;                                                \
    ++i;                                                                                              \
  }
  MFG_STATE_FIELDS(F)
#undef F
  return L;
}

inline void bind_state(const MfgSpec& sp, int64_t N, void* base, State& st) {
  Layout L = compute_layout(sp, N);
  st.N = N;
  st.blk_i = L.blk_i;
  st.blk_f = L.blk_f;
  st.base_i = static_cast<char*>(base);
  size_t i = 0;
  char* b = static_cast<char*>(base);
#define F(type, name, rows_expr) st.name = reinterpret_cast<type*>(b + L.fields[i++].offset);
  MFG_STATE_FIELDS(F)
#undef F
}

// derived tables, built on the host from the spec's level arrays
struct HostTables {
  std::vector<uint8_t> wall, door_map, nexthop;
  std::vector<uint16_t> floor_pos, floor_index, wall_uid, wall_pos, door_pos;
  std::vector<uint64_t> wall_win, wall_box, door_near, door_adj, vis_box, wall_cand64, wall_win64;
  std::vector<uint32_t> wall_cand_rng, door_win;
  std::vector<uint64_t> vis_tab;
  std::vector<uint16_t> rank_tab;
};

// light blocking for the rank tables: walls and the closed doors of the subset
struct RankTabLight {
  const uint8_t* wall; const uint8_t* door_map; int H, W; uint32_t sub;
  MFG_HD bool operator()(int xx, int yy) const {
    if (xx < 0 || yy < 0 || xx >= H || yy >= W) return false;
    if (wall[(size_t)xx * W + yy]) return true;
    const int d = door_map[(size_t)xx * W + yy];
    return d != 0xFF && ((sub >> d) & 1u);
  }
};

// first-visit rank tables of the exact observation path (Tables::rank_tab): full observability, at most 4 doors, bounded size.
// Built with the very ray walk the kernels use (ray_walk, mfg_core.cuh) over every (tile, closed door subset).
inline void build_rank_table(const MfgSpec& sp, HostTables& t) {
  t.rank_tab.clear();
  if (sp.pomdp_r != 0 || sp.n_doors > 4 || sp.n_rays <= 0) return;
  const int H = sp.H, W = sp.W, R = obs_ray_radius(sp), BW = 2 * R + 1;
  if (R > RANK_RMAX) return;
  const size_t nsub = (size_t)1 << sp.n_doors, cells = (size_t)BW * BW;
  if ((size_t)H * W * nsub * cells > ((size_t)32 << 20)) return;             // keep it well inside L2
  t.rank_tab.assign((size_t)H * W * nsub * cells, (uint16_t)RANK_INF);
  for (int x = 0; x < H; ++x)
    for (int y = 0; y < W; ++y) {
      if (t.wall[(size_t)x * W + y]) continue;
      for (size_t sub = 0; sub < nsub; ++sub) {
        const RankTabLight blocks{t.wall.data(), t.door_map.data(), H, W, (uint32_t)sub};
        ray_walk(sp, x, y, R, BW, blocks, &t.rank_tab[(((size_t)x * W + y) * nsub + sub) * cells]);
      }
    }
}

inline std::string build_tables(const MfgSpec& sp, HostTables& t) {
  const int H = sp.H, W = sp.W;
  if (H <= 0 || W <= 0 || H > 255 || W > 255) return "level shape must be within 1..255";
  if (!sp.walls || !sp.floor_pos) return "walls / floor_pos must not be NULL";
  if (sp.n_doors && !sp.door_pos) return "door_pos must not be NULL when n_doors > 0";
  t.wall.assign(sp.walls, sp.walls + (size_t)H * W);
  const size_t hw_pad = ((size_t)H * W + 3) / 4 * 4;       // k_step copies wall / door_map as 32-bit words
  t.floor_pos.assign(sp.floor_pos, sp.floor_pos + sp.n_floor);
  t.door_pos.assign(sp.door_pos, sp.door_pos + sp.n_doors);
  if (t.door_pos.empty()) t.door_pos.push_back(NO_POS);
  t.door_map.assign((size_t)H * W, 0xFF);
  for (int d = 0; d < sp.n_doors; ++d) {
    int x = px(sp.door_pos[d]), y = py(sp.door_pos[d]);
    if (x >= H || y >= W) return "door position outside the level";
    t.door_map[(size_t)x * W + y] = (uint8_t)d;
  }
  t.wall.resize(hw_pad, 0);
  t.door_map.resize(hw_pad, 0xFF);
  t.floor_index.assign((size_t)H * W, 0xFFFF);
  for (int f = 0; f < sp.n_floor; ++f) {
    int x = px(sp.floor_pos[f]), y = py(sp.floor_pos[f]);
    if (x >= H || y >= W || t.wall[(size_t)x * W + y]) return "floor_pos entry is not a floor tile";
    t.floor_index[(size_t)x * W + y] = (uint16_t)f;
  }
  t.wall_uid.assign((size_t)H * W, 0xFFFF);
  t.wall_pos.clear();
  for (int x = 0; x < H; ++x)
    for (int y = 0; y < W; ++y)
      if (t.wall[(size_t)x * W + y]) {
        t.wall_uid[(size_t)x * W + y] = (uint16_t)t.wall_pos.size();
        t.wall_pos.push_back(mkpos(x, y));
      }
  if ((int)t.wall_pos.size() != sp.n_walls) return "n_walls does not match the wall map";
  if (t.wall_pos.empty()) t.wall_pos.push_back(NO_POS);
  if (sp.nexthop) t.nexthop.assign(sp.nexthop, sp.nexthop + (size_t)sp.n_floor * sp.n_floor);
  // per-tile window wall masks (bit = dx_idx * D + dy_idx), used by the tiled observation kernel
  const int r = sp.pomdp_r, D = 2 * r + 1;
  t.wall_win.assign((size_t)H * W, 0);
  if (r >= 1 && r <= 3) {
    for (int x = 0; x < H; ++x)
      for (int y = 0; y < W; ++y) {
        uint64_t m = 0;
        for (int dx = -r; dx <= r; ++dx)
          for (int dy = -r; dy <= r; ++dy) {
            int xx = x + dx, yy = y + dy;
            if (xx >= 0 && yy >= 0 && xx < H && yy < W && t.wall[(size_t)xx * W + yy])
              m |= 1ull << ((dx + r) * D + (dy + r));
          }
        t.wall_win[(size_t)x * W + y] = m;
      }
  }
  // per-tile wall masks of the radius-D box (faithful observation mode: first-visit order over the full rays)
  t.wall_box.assign((size_t)H * W * 4, 0);
  if (r >= 1 && r <= 3) {
    const int BW = 2 * D + 1;
    for (int x = 0; x < H; ++x)
      for (int y = 0; y < W; ++y)
        for (int dx = -D; dx <= D; ++dx)
          for (int dy = -D; dy <= D; ++dy) {
            int xx = x + dx, yy = y + dy;
            if (xx >= 0 && yy >= 0 && xx < H && yy < W && t.wall[(size_t)xx * W + yy]) {
              const int b = (dx + D) * BW + (dy + D);
              t.wall_box[((size_t)x * W + y) * 4 + (b >> 6)] |= 1ull << (b & 63);
            }
          }
  }
  // per-tile mask of the doors inside the radius-D box (the observation kernels only look at those)
  t.door_near.assign((size_t)H * W, 0);
  t.door_adj.assign((size_t)H * W, 0);
  for (int x = 0; x < H; ++x)
    for (int y = 0; y < W; ++y)
      for (int d = 0; d < sp.n_doors; ++d) {
        const int dx = px(sp.door_pos[d]) - x, dy = py(sp.door_pos[d]) - y;
        if (dx >= -D && dx <= D && dy >= -D && dy <= D) t.door_near[(size_t)x * W + y] |= 1ull << d;
        if (dx >= -1 && dx <= 1 && dy >= -1 && dy <= 1) t.door_adj[(size_t)x * W + y] |= 1ull << d;
      }
  return "";
}

inline std::string validate_spec(const MfgSpec& sp) {
  if (sp.n_agents < 1 || sp.n_agents > MFG_MAX_AGENTS) return "n_agents out of range";
  if (!sp.individual_rewards)
    return "individual_rewards = 0 is not runnable: the reference raises TypeError at environment/factory.py:217 on the first step";
  if (sp.pomdp_r < 0 || sp.pomdp_r > 3) return "pomdp_r must be 0 (full observability) or 1..3";
  if (sp.pomdp_r == 0 && (sp.H < sp.W ? sp.H : sp.W) > MFG_MAX_RAY_LEN - 1)
    return "full observability needs min(H, W) <= 15 (ray length limit)";
  if (sp.n_doors < 0 || sp.n_doors > MFG_MAX_DOORS) return "n_doors out of range";
  if (sp.has_dirt && (sp.dirt_slots < 1 || sp.dirt_slots > MFG_MAX_DIRT)) return "dirt_slots out of range";
  if (sp.n_rules < 0 || sp.n_rules > MFG_MAX_RULES) return "n_rules out of range";
  if (sp.n_groups < 0 || sp.n_groups > MFG_MAX_GROUPS) return "n_groups out of range";
  if (sp.n_rays < 1 || sp.n_rays > MFG_MAX_RAYS) return "n_rays out of range";
  const int small[] = {sp.n_items, sp.n_dropoff, sp.n_pods, sp.n_dest, sp.n_machines, sp.n_maint};
  for (int n : small) if (n < 0 || n > MFG_MAX_SMALL) return "small group size out of range";
  for (int a = 0; a < sp.n_agents; ++a) {
    if (sp.n_actions[a] < 1 || sp.n_actions[a] > MFG_MAX_ACTIONS) return "n_actions out of range";
    if (sp.n_channels[a] < 1 || sp.n_channels[a] > MFG_MAX_CHANNELS) return "n_channels out of range";
  }
  for (int r = 0; r < sp.n_rays; ++r) if (sp.ray_len[r] < 1 || sp.ray_len[r] > MFG_MAX_RAY_LEN) return "ray_len out of range";
  if (sp.n_floor < sp.n_agents) return "fewer floor tiles than agents";
  if (sp.dest_mode < MFG_DEST_FREE || sp.dest_mode > MFG_DEST_PER_AGENT) return "dest_mode out of range";
  for (int k = 0; k < sp.n_dest; ++k) {
    if (sp.dest_bound[k] < -1 || sp.dest_bound[k] >= sp.n_agents) return "dest_bound out of range";
    if (sp.dest_mode != MFG_DEST_FREE && sp.dest_bound[k] < 0) return "bound destination spawn modes need dest_bound";
    if (sp.dest_n_cand[k] < 0 || sp.dest_n_cand[k] > MFG_MAX_FIXED) return "dest_n_cand out of range";
  }
  for (int r = 0; r < sp.n_rules; ++r)
    if (sp.rule_op[r] == MFG_R_DONE_MAX_STEPS && !(sp.rule_param[r][0] >= 0 && sp.rule_param[r][0] <= 65535))
      return "DoneAtMaxStepsReached.max_steps must be within 0..65535 (16-bit step counter)";
  return "";
}

}  // namespace mfg
