// hostsim.cpp - TEST-ONLY host build of the per-environment device functions (mfg_core.cuh).
//
// It compiles the exact step / reset / observation code the CUDA kernels run with g++, so that the
// reference traces can be replayed in the CPU-only build container.  It is never linked into, loaded
// by or shipped with the product library; the product has no CPU path.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../marl_factory_grid_b200/csrc/mfg_host.hpp"

using namespace mfg;

struct HsHandle {
  MfgSpec sp;
  HostTables ht;
  Tables tb;
  State st;
  int64_t N;
  std::vector<FieldInfo> fields;
  size_t bytes;
  std::vector<unsigned long long> stats;
  bool ever_reset = false;
  uint8_t* flags = nullptr;
};

template <typename Fn>
static void dispatch(int A, Fn fn) {
  if (A <= 1) fn(std::integral_constant<int, 1>());
  else if (A <= 2) fn(std::integral_constant<int, 2>());
  else if (A <= 4) fn(std::integral_constant<int, 4>());
  else if (A <= 8) fn(std::integral_constant<int, 8>());
  else fn(std::integral_constant<int, 16>());
}

extern "C" {

const char* hs_create(const MfgSpec* spec, int64_t n_envs, int64_t env_id_offset, HsHandle** out) {
  static std::string err;
  err = validate_spec(*spec);
  if (!err.empty()) return err.c_str();
  HsHandle* h = new HsHandle();
  h->sp = *spec;
  err = build_tables(*spec, h->ht);
  if (!err.empty()) { delete h; return err.c_str(); }
  h->N = n_envs;
  {
    Layout L = compute_layout(h->sp, n_envs);
    h->fields = L.fields;
    h->bytes = L.total;
  }
  h->stats.assign(MFG_N_STATS, 0);
  h->tb.wall = h->ht.wall.data(); h->tb.door_map = h->ht.door_map.data(); h->tb.floor_pos = h->ht.floor_pos.data();
  h->tb.floor_index = h->ht.floor_index.data(); h->tb.wall_uid = h->ht.wall_uid.data(); h->tb.wall_pos = h->ht.wall_pos.data();
  h->tb.door_pos = h->ht.door_pos.data(); h->tb.nexthop = h->ht.nexthop.empty() ? nullptr : h->ht.nexthop.data();
  h->tb.wall_win = h->ht.wall_win.data();
  h->tb.wall_box = h->ht.wall_box.data();
  h->tb.door_adj = h->ht.door_adj.data();
  h->tb.env_id_offset = env_id_offset; h->tb.stats = h->stats.data();
  h->sp.walls = nullptr; h->sp.floor_pos = nullptr; h->sp.door_pos = nullptr; h->sp.nexthop = nullptr;
  *out = h;
  return nullptr;
}
void hs_destroy(HsHandle* h) { delete h; }
size_t hs_state_bytes(const HsHandle* h) { return h->bytes; }
int hs_state_field(const HsHandle* h, const char* name, MfgField* out) {
  for (const auto& f : h->fields)
    if (std::string(f.name) == name) { out->offset = f.offset; out->rows = f.rows; out->elem_size = f.elem_size; out->block_bytes = f.block_bytes; return 0; }
  return -1;
}
void hs_bind_state(HsHandle* h, void* base) { bind_state(h->sp, h->N, base, h->st); }

void hs_reset(HsHandle* h, const uint8_t* mask) {
  dispatch(h->sp.n_agents, [&](auto amax) {
    for (int64_t e = 0; e < h->N; ++e)
      if (!mask || mask[e])
        env_reset<decltype(amax)::value>(h->sp, h->tb, h->st, e, (!mask && !h->ever_reset) ? 0u : field_at(h->st, h->st.episode, 0, e) + 1);
  });
  if (!mask) h->ever_reset = true;
}
void hs_step(HsHandle* h, const int32_t* actions, const uint8_t* maint_act, const int8_t* respawn_n,
             const uint16_t* respawn_pos, float* reward, uint8_t* done, int auto_reset) {
  StepIO io{actions, maint_act, respawn_n, respawn_pos, reward, done, auto_reset, nullptr, nullptr, h->flags};
  dispatch(h->sp.n_agents, [&](auto amax) {
    for (int64_t e = 0; e < h->N; ++e) env_step<decltype(amax)::value, MfgSpec>(h->sp, h->sp, h->tb, h->st, e, io);
  });
}
void hs_observe(HsHandle* h, float* obs) {
  const int DD = obs_plane_cells(h->sp);
  int total = 0;
  for (int a = 0; a < h->sp.n_agents; ++a) total += h->sp.n_channels[a];
  dispatch(h->sp.n_agents, [&](auto amax) {
    for (int64_t e = 0; e < h->N; ++e)
      for (int a = 0; a < h->sp.n_agents; ++a)
        obs_agent_direct<decltype(amax)::value>(h->sp, h->tb, h->st, e, a,
                                                obs + ((size_t)e * total + h->sp.ch_offset[a]) * DD);
  });
}
void hs_bind_flags(HsHandle* h, uint8_t* flags) { h->flags = flags; }
void hs_stats(HsHandle* h, int64_t* out) { for (int i = 0; i < MFG_N_STATS; ++i) out[i] = (int64_t)h->stats[i]; }
}
