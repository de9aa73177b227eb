// mfg_abi.cu - the C ABI of libmfg_b200.so (include/mfg_b200.h): handle management, argument checking, launches.
// Kernels live in mfg_step.cu (reset / step) and mfg_obs.cu (observations).
#include <cstdio>
#include <cstring>
#include "mfg_internal.hpp"

using namespace mfg;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(call)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess) return fail(MFG_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
  } while (0)

template <typename T>
static int upload(MfgHandle* h, const std::vector<T>& v, const T** out) {
  void* d = nullptr;
  size_t bytes = v.size() * sizeof(T);
  if (bytes == 0) { *out = nullptr; return MFG_OK; }
  CUDA_TRY(cudaMalloc(&d, bytes));
  h->dev_allocs.push_back(d);
  CUDA_TRY(cudaMemcpy(d, v.data(), bytes, cudaMemcpyHostToDevice));
  *out = static_cast<const T*>(d);
  return MFG_OK;
}

// live kernel timing: event pairs on the launching stream, read back (and released) by mfg_get_info
struct Timed {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>>* v; cudaStream_t s; cudaEvent_t b = nullptr, e = nullptr;
  Timed(MfgHandle* h, std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& vec, cudaStream_t st) : v(h->timing ? &vec : nullptr), s(st) {
    if (v && cudaEventCreate(&b) == cudaSuccess && cudaEventCreate(&e) == cudaSuccess) cudaEventRecord(b, s); else v = nullptr;
  }
  ~Timed() { if (v) { cudaEventRecord(e, s); v->emplace_back(b, e); } }
};
static int64_t drain_ns(std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& v) {
  double ms = 0.0;
  for (auto& p : v) {
    float t = 0.f;
    cudaEventSynchronize(p.second);
    if (cudaEventElapsedTime(&t, p.first, p.second) == cudaSuccess) ms += t;
    cudaEventDestroy(p.first); cudaEventDestroy(p.second);
  }
  v.clear();
  return (int64_t)(ms * 1e6);
}

extern "C" {

const char* mfg_last_error(void) { return g_err.c_str(); }
const char* mfg_version(void) { return "mfg_b200 0.1 (sm_100a)"; }

int mfg_create(const MfgSpec* spec, int64_t n_envs, int64_t env_id_offset, MfgHandle** out) {
  if (!spec || !out || n_envs <= 0) return fail(MFG_E_INVALID, "mfg_create: bad arguments");
  std::string err = validate_spec(*spec);
  if (!err.empty()) return fail(MFG_E_INVALID, "mfg_create: " + err);
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return fail(MFG_E_CUDA, "mfg_create: no CUDA device available (this engine has no CPU path)");
  HostTables ht;
  err = build_tables(*spec, ht);
  if (!err.empty()) return fail(MFG_E_INVALID, "mfg_create: " + err);

  build_vis_tables(*spec, ht);
  build_win_vis_tables(*spec, ht);
  build_rank_table(*spec, ht);

  MfgHandle* h = new MfgHandle();
  h->sp = *spec;
  h->sp.walls = nullptr; h->sp.floor_pos = nullptr; h->sp.door_pos = nullptr; h->sp.nexthop = nullptr;
  h->N = n_envs;
  int rc;
#define UP(field) if ((rc = upload(h, ht.field, &h->tb.field)) != MFG_OK) { mfg_destroy(h); return rc; }
  UP(wall) UP(door_map) UP(floor_pos) UP(floor_index) UP(wall_uid) UP(wall_pos) UP(door_pos) UP(nexthop) UP(wall_win) UP(wall_box) UP(door_near) UP(door_adj)
  UP(vis_box) UP(wall_cand64) UP(wall_cand_rng) UP(wall_win64) UP(door_win) UP(vis_tab) UP(rank_tab)
#undef UP
  h->tb.env_id_offset = env_id_offset;
  void* d = nullptr;
  if (const char* ev = getenv("MFG_STEP_KERNEL")) h->step_kernel = atoi(ev);      // development aid (A/B runs)
  if (const char* ev = getenv("MFG_STEP_BLOCKS")) h->step_blocks = atoi(ev);
  if (cudaMalloc(&d, sizeof(unsigned long long) * MFG_N_STATS) != cudaSuccess) { mfg_destroy(h); return fail(MFG_E_NOMEM, "stats alloc"); }
  h->dev_allocs.push_back(d);
  cudaMemset(d, 0, sizeof(unsigned long long) * MFG_N_STATS);
  h->tb.stats = static_cast<unsigned long long*>(d);
  if (cudaMalloc(&d, sizeof(MfgSpec)) != cudaSuccess) { mfg_destroy(h); return fail(MFG_E_NOMEM, "spec alloc"); }
  h->dev_allocs.push_back(d);
  h->d_sp = static_cast<MfgSpec*>(d);
  cudaMemcpy(h->d_sp, &h->sp, sizeof(MfgSpec), cudaMemcpyHostToDevice);
  {
    Layout L = compute_layout(h->sp, n_envs);
    h->fields = L.fields;
    h->state_bytes = L.total;
  }
  h->total_channels = 0;
  for (int a = 0; a < h->sp.n_agents; ++a) h->total_channels += h->sp.n_channels[a];
  h->DD = obs_plane_cells(h->sp);

  plan_obs(h);
  *out = h;
  return MFG_OK;
}

void mfg_destroy(MfgHandle* h) {
  if (!h) return;
  for (void* p : h->dev_allocs) cudaFree(p);
  if (h->d_actions) cudaFree(h->d_actions);
  if (h->d_reward) cudaFree(h->d_reward);
  if (h->d_done) cudaFree(h->d_done);
  if (h->d_obs) cudaFree(h->d_obs);
  if (h->d_reset_list) cudaFree(h->d_reset_list);
  if (h->d_reset_count) cudaFree(h->d_reset_count);
  if (h->d_redo) cudaFree(h->d_redo);
  if (h->d_obs_prog) cudaFree(h->d_obs_prog);
  if (h->d_row_tab) cudaFree(h->d_row_tab);
  drain_ns(h->t_step); drain_ns(h->t_obs); drain_ns(h->t_reset);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->side) cudaStreamDestroy(h->side);
  delete h;
}

size_t mfg_state_bytes(const MfgHandle* h) { return h ? h->state_bytes : 0; }

int mfg_state_field(const MfgHandle* h, const char* name, MfgField* out) {
  if (!h || !name || !out) return fail(MFG_E_INVALID, "mfg_state_field: bad arguments");
  for (const auto& f : h->fields)
    if (strcmp(f.name, name) == 0) { out->offset = f.offset; out->rows = f.rows; out->elem_size = f.elem_size; out->block_bytes = f.block_bytes; return MFG_OK; }
  return fail(MFG_E_INVALID, std::string("mfg_state_field: unknown field ") + name);
}

int mfg_bind_state(MfgHandle* h, void* d_state) {
  if (!h || !d_state) return fail(MFG_E_INVALID, "mfg_bind_state: bad arguments");
  if (reinterpret_cast<uintptr_t>(d_state) % 256) return fail(MFG_E_INVALID, "mfg_bind_state: buffer must be 256-byte aligned");
  bind_state(h->sp, h->N, d_state, h->st);
  if (!h->d_row_tab) {               // rows of the integer region (ColTab): offsets do not depend on the buffer address
    std::vector<uint32_t> rows;
    const MfgSpec& sp = h->sp;
#define F(type, name, rows_expr)                                                                                    \
    if (!std::is_same<type, double>::value) {                                                                        \
      const uint32_t off = (uint32_t)(reinterpret_cast<const char*>(h->st.name) - h->st.base_i);                     \
      const uint32_t lg = sizeof(type) == 1 ? 0u : sizeof(type) == 2 ? 1u : sizeof(type) == 4 ? 2u : 3u;             \
      for (int r = 0; r < (int)(rows_expr); ++r) rows.push_back((off + (uint32_t)(r * ENV_BLOCK * sizeof(type))) | (lg << 28)); \
    }
    MFG_STATE_FIELDS(F)
#undef F
    h->n_row_tab = (int)rows.size();
    h->row_tab_host = rows;
    CUDA_TRY(cudaMalloc(&h->d_row_tab, rows.size() * sizeof(uint32_t)));
    CUDA_TRY(cudaMemcpy(h->d_row_tab, rows.data(), rows.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  h->bound = true;
  return MFG_OK;
}

#define NEED_BOUND(h) if (!(h) || !(h)->bound) return fail(MFG_E_INVALID, "state buffer not bound (call mfg_bind_state)")

int mfg_reset(MfgHandle* h, const uint8_t* d_env_mask, void* stream) {
  NEED_BOUND(h);
  CUDA_TRY(launch_reset(h, d_env_mask, static_cast<cudaStream_t>(stream)));
  if (!d_env_mask) h->ever_reset = true;
  h->launches++;
  return MFG_OK;
}

static int step_impl(MfgHandle* h, const int32_t* d_actions, const MfgTape* tape, float* d_reward, uint8_t* d_done,
                     int auto_reset, void* stream, bool split_reset) {
  NEED_BOUND(h);
  if (!d_actions || !d_reward || !d_done) return fail(MFG_E_INVALID, "mfg_step: NULL buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (auto_reset && h->defer_reset && !h->d_reset_list) {
    CUDA_TRY(cudaMalloc(&h->d_reset_list, (size_t)h->N * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&h->d_reset_count, sizeof(uint32_t)));
  }
  const bool defer = auto_reset && h->defer_reset;
  StepIO io{d_actions, tape ? tape->d_maint_action : nullptr, tape ? tape->d_respawn_n : nullptr,
            tape ? tape->d_respawn_pos : nullptr, d_reward, d_done, auto_reset,
            defer ? h->d_reset_list : nullptr, defer ? h->d_reset_count : nullptr, h->d_flags};
  bool need_policy = false;
  for (int r = 0; r < h->sp.n_rules; ++r) need_policy |= h->sp.rule_op[r] == MFG_R_MOVE_MAINTAINERS;
  if (need_policy && !io.maint_act && !h->tb.nexthop)
    return fail(MFG_E_INVALID, "mfg_step: MoveMaintainers without a tape needs the next-hop table (MfgSpec.nexthop)");
  if (split_reset) {                 // mfg_step_observe runs the re-spawn itself (overlapped with the observation kernel)
    Timed t(h, h->t_step, s);
    CUDA_TRY(launch_step_kernel(h, io, s));
    h->launches++;
    return MFG_OK;
  }
  {
    Timed t(h, h->t_step, s);
    CUDA_TRY(launch_step_kernel(h, io, s));
  }
  {
    Timed t(h, h->t_reset, s);
    CUDA_TRY(launch_reset_list(h, io, s));
  }
  h->launches += defer ? 2 : 1;
  return MFG_OK;
}

int mfg_step(MfgHandle* h, const int32_t* d_actions, const MfgTape* tape, float* d_reward, uint8_t* d_done,
             int auto_reset, void* stream) {
  return step_impl(h, d_actions, tape, d_reward, d_done, auto_reset, stream, false);
}

int mfg_observe(MfgHandle* h, float* d_obs, void* stream) {
  NEED_BOUND(h);
  if (!d_obs) return fail(MFG_E_INVALID, "mfg_observe: NULL buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (h->obs_kernel == 2 && !h->plan.ok) return fail(MFG_E_UNSUPPORTED, "tiled observation kernel not available for this spec");
  const bool tiled = h->plan.ok && h->obs_kernel != 1 && h->obs_kernel != 3;
  Timed t(h, h->t_obs, s);
  CUDA_TRY(tiled ? launch_obs_tiled(h, d_obs, s) : launch_obs_direct(h, d_obs, s));
  h->launches += tiled ? 2 : 1;       // tiled kernel + its redo pass
  return MFG_OK;
}

int mfg_step_observe(MfgHandle* h, const int32_t* d_actions, const MfgTape* tape, float* d_reward, uint8_t* d_done,
                     float* d_obs, int auto_reset, void* stream) {
  NEED_BOUND(h);
  const bool tiled = h->plan.ok && h->obs_kernel != 1 && h->obs_kernel != 3;
  if (!(auto_reset && h->defer_reset && h->overlap_reset && tiled && d_obs)) {
    int rc = mfg_step(h, d_actions, tape, d_reward, d_done, auto_reset, stream);
    if (rc != MFG_OK) return rc;
    return mfg_observe(h, d_obs, stream);
  }
  // Overlapped form.  Finished envs are re-spawned by the packed reset kernel and then observed by the exact per-agent
  // kernel over the same list, both on a high-priority side stream, WHILE the tiled observation kernel (caller's
  // stream) covers every other env - it skips the finished ones via the done flags and does not write their tiles.
  // The side work is latency-bound (a few thousand busy threads); serialised it costs ~8 % of the step.
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!h->side) {
    int prio_lo = 0, prio_hi = 0;      // highest priority: the few CTAs of the side work must not queue behind the observation grid
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_hi));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  }
  int rc = step_impl(h, d_actions, tape, d_reward, d_done, auto_reset, stream, true);
  if (rc != MFG_OK) return rc;
  StepIO io{};
  io.auto_reset = 1; io.reset_list = h->d_reset_list; io.reset_count = h->d_reset_count;
  CUDA_TRY(cudaEventRecord(h->ev_fork, s));
  CUDA_TRY(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
  {
    Timed t(h, h->t_reset, h->side);
    CUDA_TRY(launch_reset_list(h, io, h->side));
  }
  CUDA_TRY(launch_obs_tiled_list(h, d_obs, h->side, h->d_reset_list, h->d_reset_count));   // tiled kernel in list mode (+ its redo pass)
  CUDA_TRY(cudaEventRecord(h->ev_join, h->side));
  {
    Timed t(h, h->t_obs, s);
    CUDA_TRY(launch_obs_tiled(h, d_obs, s, d_done));
  }
  CUDA_TRY(cudaStreamWaitEvent(s, h->ev_join, 0));
  h->launches += 5;                   // reset list, tiled observation + its redo pass, the same pair over the re-spawned envs
  return MFG_OK;
}

int mfg_random_actions(MfgHandle* h, int32_t* d_actions, uint64_t seed, uint64_t step_index, void* stream) {
  if (!h || !d_actions) return fail(MFG_E_INVALID, "mfg_random_actions: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUDA_TRY(launch_random_actions(h, d_actions, seed, (uint32_t)step_index, s));
  h->launches++;
  return MFG_OK;
}

int mfg_step_host(MfgHandle* h, const int32_t* h_actions, float* h_reward, uint8_t* h_done, float* h_obs,
                  int auto_reset, void* stream) {
  NEED_BOUND(h);
  if (!h_actions || !h_reward || !h_done) return fail(MFG_E_INVALID, "mfg_step_host: NULL buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int A = h->sp.n_agents, NR = A;
  const size_t obs_bytes = (size_t)h->N * h->total_channels * h->DD * sizeof(float);
  if (!h->d_actions) {
    CUDA_TRY(cudaMalloc(&h->d_actions, (size_t)h->N * A * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&h->d_reward, (size_t)h->N * NR * sizeof(float)));
    CUDA_TRY(cudaMalloc(&h->d_done, (size_t)h->N));
    CUDA_TRY(cudaMalloc(&h->d_obs, obs_bytes));
  }
  CUDA_TRY(cudaMemcpyAsync(h->d_actions, h_actions, (size_t)h->N * A * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  int rc = mfg_step_observe(h, h->d_actions, nullptr, h->d_reward, h->d_done, h->d_obs, auto_reset, stream);
  if (rc != MFG_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(h_reward, h->d_reward, (size_t)h->N * NR * sizeof(float), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h_done, h->d_done, (size_t)h->N, cudaMemcpyDeviceToHost, s));
  if (h_obs) CUDA_TRY(cudaMemcpyAsync(h_obs, h->d_obs, obs_bytes, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return MFG_OK;
}

int mfg_bind_step_flags(MfgHandle* h, uint8_t* d_flags) {
  if (!h) return fail(MFG_E_INVALID, "mfg_bind_step_flags: bad arguments");
  h->d_flags = d_flags;
  return MFG_OK;
}

int mfg_stats(MfgHandle* h, int64_t* d_out, int zero_after, void* stream) {
  if (!h || !d_out) return fail(MFG_E_INVALID, "mfg_stats: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUDA_TRY(cudaMemcpyAsync(d_out, h->tb.stats, sizeof(int64_t) * MFG_N_STATS, cudaMemcpyDeviceToDevice, s));
  if (zero_after) CUDA_TRY(cudaMemsetAsync(h->tb.stats, 0, sizeof(int64_t) * MFG_N_STATS, s));
  return MFG_OK;
}

int mfg_set_option(MfgHandle* h, const char* name, int64_t value) {
  if (!h || !name) return fail(MFG_E_INVALID, "mfg_set_option: bad arguments");
  if (strcmp(name, "obs_kernel") == 0) {
    if (value < 0 || value > 3) return fail(MFG_E_INVALID, "obs_kernel must be 0 (auto), 1 (exact, block-staged), 2 (tiled) or 3 (exact, plain)");
    if (value == 2 && !h->plan.ok) return fail(MFG_E_UNSUPPORTED, "tiled observation kernel not available for this spec");
    h->obs_kernel = (int)value;
    return MFG_OK;
  }
  if (strcmp(name, "defer_reset") == 0) {       // 1 = packed reset kernel after the step (default), 0 = in-line reset
    h->defer_reset = value != 0;
    return MFG_OK;
  }
  if (strcmp(name, "obs_cap") == 0) {          // sprite slots per (env, agent); small values force the overflow path (tests)
    if (value < 1 || value > h->plan.cap_max) return fail(MFG_E_INVALID, "obs_cap must be in 1..cap_max");
    h->plan.cap = (int)value;
    return MFG_OK;
  }
  if (strcmp(name, "overlap_reset") == 0) { h->overlap_reset = value != 0; return MFG_OK; }
  if (strcmp(name, "reseed") == 0) { h->ever_reset = false; return MFG_OK; }      // the next full reset starts again at episode 0
  if (strcmp(name, "timing") == 0) { h->timing = value != 0; return MFG_OK; }
  if (strcmp(name, "step_kernel") == 0) {      // 1 = barriers + dirt uids left in HBM (default), 2 = barriers only, 0 = neither
    if (value < 0 || value > 2) return fail(MFG_E_INVALID, "step_kernel must be 0, 1 or 2");
    h->step_kernel = (int)value;
    return MFG_OK;
  }
  if (strcmp(name, "step_blocks") == 0) {      // state blocks per k_step CTA: 0 = auto (default), 1..3
    if (value < 0 || value > 3) return fail(MFG_E_INVALID, "step_blocks must be in 0..3");
    h->step_blocks = (int)value;
    return MFG_OK;
  }
  if (strcmp(name, "obs_store") == 0) {        // 1 = TMA bulk store of the tile (default), 0 = LDS/STG loop
    h->obs_store = value != 0;
    return MFG_OK;
  }
  return fail(MFG_E_INVALID, std::string("mfg_set_option: unknown option ") + name);
}

int64_t mfg_get_info(const MfgHandle* h, const char* name) {
  if (!h || !name) return -1;
  if (strcmp(name, "launches") == 0) return h->launches;
  if (strcmp(name, "step_ns") == 0) return drain_ns(const_cast<MfgHandle*>(h)->t_step);      // sum since the last read (synchronises)
  if (strcmp(name, "obs_ns") == 0) return drain_ns(const_cast<MfgHandle*>(h)->t_obs);
  if (strcmp(name, "reset_ns") == 0) return drain_ns(const_cast<MfgHandle*>(h)->t_reset);
  if (strcmp(name, "tiled_ok") == 0) return h->plan.ok ? 1 : 0;
  if (strcmp(name, "obs_redo_count") == 0) {          // envs the last tiled launch handed to the exact path (synchronises)
    uint32_t n = 0;
    if (!h->d_redo || cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(&n, h->d_redo, sizeof(n), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return n;
  }
  if (strcmp(name, "obs_ctas_per_sm") == 0) return obs_ctas_per_sm(h);
  if (strcmp(name, "obs_smem") == 0) return (int64_t)h->plan.smem;
  if (strcmp(name, "obs_threads") == 0) return h->plan.nw * 32;
  if (strcmp(name, "total_channels") == 0) return h->total_channels;
  if (strcmp(name, "n_envs") == 0) return h->N;
  if (strcmp(name, "obs_cap_max") == 0) return h->plan.cap_max;
  return -1;
}

}  // extern "C"
