// mfg_kernels.cu - sm_100a kernels and the C ABI (include/mfg_b200.h) of the batched marl-factory-grid engine.
//
// Kernels (all HBM-bound integer / byte work; no tensor cores by design):
//   k_reset        one thread per env   Philox spawn                                  (mfg_core.cuh env_reset)
//   k_step         one thread per env   actions + rule hooks + done/reward (+ reset)  (mfg_core.cuh env_step)
//   k_obs_tiled    CTA = 32 envs        phase 1: one thread per (env, agent) marches the 28 window rays on a
//                                       49-bit light-block mask; phase 2: one warp per env rasterises the
//                                       channel planes in shared memory and streams them out with 128-bit stores
//   k_obs_direct   one thread per (env, agent), all parity modes incl. the faithful uid de-duplication
//   k_random_actions  uniform actions from Philox (benchmark / random rollouts)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "mfg_host.hpp"

using namespace mfg;

// --------------------------------------------------------------------------------------------------------------
// window ray table for the tiled observation kernel: built on the host from the full rays
// --------------------------------------------------------------------------------------------------------------
constexpr int MAX_WRAYS = 64;
constexpr int MAX_WLEN = 4;
struct WindowRays {
  int n;
  int len[MAX_WRAYS];
  uint8_t cell[MAX_WRAYS][MAX_WLEN];   // window cell index (dx + r) * D + (dy + r)
  uint8_t da[MAX_WRAYS][MAX_WLEN];     // the two orthogonal neighbours for the diagonal-occlusion test, 255 = straight step
  uint8_t db[MAX_WRAYS][MAX_WLEN];
};

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
  // drop prefixes of longer rays
  std::vector<std::vector<std::pair<int, int>>> fin;
  for (auto& p : kept) {
    bool is_prefix = false;
    for (auto& q : kept)
      if (q.size() > p.size() && std::equal(p.begin(), p.end(), q.begin())) is_prefix = true;
    if (!is_prefix) fin.push_back(p);
  }
  memset(&wr, 0, sizeof(wr));
  wr.n = (int)fin.size();
  for (int k = 0; k < wr.n; ++k) {
    wr.len[k] = (int)fin[k].size();
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

// --------------------------------------------------------------------------------------------------------------
// kernels
// --------------------------------------------------------------------------------------------------------------
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

template <int AMAX>
__global__ void __launch_bounds__(128) k_obs_direct(const MfgSpec* __restrict__ sp, Tables tb, State st, float* obs,
                                                    int total_channels) {
  // agent-major thread mapping: consecutive threads = consecutive envs of the same agent (coalesced state loads)
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int A = sp->n_agents;
  if (t >= st.N * A) return;
  int a = (int)(t / st.N);
  int64_t e = t - (int64_t)a * st.N;
  const int DD = (2 * sp->pomdp_r + 1) * (2 * sp->pomdp_r + 1);
  obs_agent_direct<AMAX>(*sp, tb, st, e, a, obs + ((size_t)e * total_channels + sp->ch_offset[a]) * DD);
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

// ---- tiled observation kernel (identity parity mode, pomdp_r <= 3) ---------------------------------------
constexpr int OBS_ENVS = 32;   // envs per CTA (one warp of env lanes per agent in phase 1)

struct ObsSlots {               // dynamic-position entity slots staged per env in shared memory
  int dirt0, item0, pod0, dest0, drop0, mach0, maint0, agent0, total, stride;
};

__device__ __forceinline__ uint16_t load_slot(const State& st, const ObsSlots& sl, int s, int64_t e) {
  const size_t N = (size_t)st.N;
  if (s < sl.item0) return st.dirt_pos[(size_t)(s - sl.dirt0) * N + e];
  if (s < sl.pod0) return st.item_pos[(size_t)(s - sl.item0) * N + e];
  if (s < sl.dest0) return st.pod_pos[(size_t)(s - sl.pod0) * N + e];
  if (s < sl.drop0) return st.dest_pos[(size_t)(s - sl.dest0) * N + e];
  if (s < sl.mach0) return st.drop_pos[(size_t)(s - sl.drop0) * N + e];
  if (s < sl.maint0) return st.mach_pos[(size_t)(s - sl.mach0) * N + e];
  if (s < sl.agent0) return st.maint_pos[(size_t)(s - sl.maint0) * N + e];
  return st.apos[(size_t)(s - sl.agent0) * N + e];
}

__device__ __forceinline__ void tile_add_int(float* tile, uint32_t mask, int coff, int cell, int DD, float v) {
  while (mask) {
    int c = __ffs(mask) - 1;
    mask &= mask - 1;
    atomicAdd(&tile[(coff + c) * DD + cell], v);
  }
}
__device__ __forceinline__ void tile_add_frac(float* tile, uint32_t mask, int coff, int cell, int DD, double v) {
  while (mask) {
    int c = __ffs(mask) - 1;
    mask &= mask - 1;
    float* f = &tile[(coff + c) * DD + cell];
    *f = (float)((double)*f + v);           // integer stacks were summed exactly before; one rounding like the reference
  }
}

template <int GE>   // GE = envs per output tile so that the tile is a whole number of 16-byte vectors
__global__ void k_obs_tiled(const MfgSpec* __restrict__ sp, Tables tb, State st, const WindowRays* __restrict__ wr,
                            ObsSlots sl, float* __restrict__ obs, int total_channels) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int A = sp->n_agents;
  const int r = sp->pomdp_r, D = 2 * r + 1, DD = D * D;
  const int NW = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t env0 = (int64_t)blockIdx.x * OBS_ENVS;
  const int tile_floats = GE * total_channels * DD;

  // shared memory carve-up
  float* tiles = reinterpret_cast<float*>(smem_raw);                                   // [NW][tile_floats]
  unsigned long long* s_vis = reinterpret_cast<unsigned long long*>(tiles + (size_t)NW * tile_floats);  // [32][A]
  unsigned long long* s_dopen = s_vis + OBS_ENVS * A;                                 // [32]
  uint32_t* s_reached = reinterpret_cast<uint32_t*>(s_dopen + OBS_ENVS);               // [32]
  uint16_t* s_pos = reinterpret_cast<uint16_t*>(s_reached + OBS_ENVS);                 // [32][stride]

  // ---------------- phase 1: one thread per (env = lane, agent = warp .. ) --------------------------------
  {
    const int64_t e = env0 + lane;
    const bool live = e < st.N;
    // stage the dynamic entity positions: warp w copies slots w, w+NW, ... (coalesced over the env lanes)
    for (int s = warp; s < sl.total; s += NW) s_pos[lane * sl.stride + s] = live ? load_slot(st, sl, s, e) : NO_POS;
    if (warp == 0) {
      s_dopen[lane] = (live && sp->n_doors) ? st.door_open[e] : 0ull;
      s_reached[lane] = (live && sp->n_dest) ? st.dest_reached[e] : 0u;
    }
    for (int a = warp; a < A; a += NW) {
      unsigned long long vis = 0ull;
      if (live) {
        const uint16_t p = st.apos[(size_t)a * st.N + e];
        const int ax = px(p), ay = py(p);
        unsigned long long B = tb.wall_win[ax * sp->W + ay];
        if (sp->n_doors) {
          const unsigned long long dopen = st.door_open[e];
          for (int d = 0; d < sp->n_doors; ++d) {
            const uint16_t q = tb.door_pos[d];
            const int dx = px(q) - ax + r, dy = py(q) - ay + r;
            if (!((dopen >> d) & 1) && dx >= 0 && dy >= 0 && dx < D && dy < D) B |= 1ull << (dx * D + dy);
          }
        }
        const int centre = r * D + r;
        vis = 1ull << centre;
        if (!((B >> centre) & 1)) {                       // an agent inside a closed door only sees its own tile
          for (int k = 0; k < wr->n; ++k) {
            const int len = wr->len[k];
            for (int s = 0; s < len; ++s) {
              const int cell = wr->cell[k][s], da = wr->da[k][s], db = wr->db[k][s];
              const bool hits = (B >> cell) & 1;
              const bool diag = da != 255 && ((B >> da) & 1) && ((B >> db) & 1);
              if (!diag) vis |= 1ull << cell;
              if (hits || diag) break;
            }
          }
        }
      }
      s_vis[lane * A + a] = vis;
    }
  }
  __syncthreads();

  // ---------------- phase 2: one warp per group of GE envs ------------------------------------------------
  float* tile = tiles + (size_t)warp * tile_floats;
  const int n_groups = OBS_ENVS / GE;
  for (int g = warp; g < n_groups; g += NW) {
    const int64_t eg = env0 + (int64_t)g * GE;
    if (eg >= st.N) break;
    // zero the tile
    {
      float4* t4 = reinterpret_cast<float4*>(tile);
      const int n4 = tile_floats >> 2;
      for (int i = lane; i < n4; i += 32) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    for (int ge = 0; ge < GE; ++ge) {
      const int el = g * GE + ge;                 // env index inside the CTA
      const int64_t e = env0 + el;
      if (e >= st.N) break;
      float* te = tile + (size_t)ge * total_channels * DD;
      const uint16_t* pos = s_pos + el * sl.stride;
      // -- integer-valued terms: walls, agents, items, pods, destinations, drop-offs, machines, maintainers
      for (int a = 0; a < A; ++a) {
        const uint32_t* chm = sp->term_chmask[a];
        const int coff = sp->ch_offset[a];
        const unsigned long long vis = s_vis[el * A + a];
        const uint16_t ap = pos[sl.agent0 + a];
        const int ax = px(ap) - r, ay = py(ap) - r;
        const uint32_t mw = chm[MFG_G_WALLS];
        if (mw) {
          const unsigned long long wv = tb.wall_win[px(ap) * sp->W + py(ap)] & vis;
          for (int cell = lane; cell < DD; cell += 32)
            if ((wv >> cell) & 1) tile_add_int(te, mw, coff, cell, DD, 1.0f);
        }
        for (int s = sl.item0 + lane; s < sl.total; s += 32) {
          const uint16_t q = pos[s];
          if (q == NO_POS) continue;
          int term; float enc = 1.0f;
          if (s < sl.pod0) term = MFG_G_ITEMS;
          else if (s < sl.dest0) term = MFG_G_PODS;
          else if (s < sl.drop0) { term = MFG_G_DEST; if ((s_reached[el] >> (s - sl.dest0)) & 1) continue; }
          else if (s < sl.mach0) term = MFG_G_DROPOFF;
          else if (s < sl.maint0) { term = MFG_G_MACHINES; enc = (float)ENC_MACHINE; }
          else if (s < sl.agent0) term = MFG_G_MAINT;
          else term = MFG_G_AGENT0 + (s - sl.agent0);
          const uint32_t m = chm[term];
          if (!m) continue;
          const int dx = px(q) - ax, dy = py(q) - ay;
          if (dx < 0 || dy < 0 || dx >= D || dy >= D) continue;
          const int cell = dx * D + dy;
          if ((vis >> cell) & 1) tile_add_int(te, m, coff, cell, DD, enc);
        }
      }
      __syncwarp();
      // -- doors (0.4444 open / 0.6666 closed)
      if (sp->n_doors) {
        const unsigned long long dopen = s_dopen[el];
        for (int a = 0; a < A; ++a) {
          const uint32_t m = sp->term_chmask[a][MFG_G_DOORS];
          if (!m) continue;
          const unsigned long long vis = s_vis[el * A + a];
          const uint16_t ap = pos[sl.agent0 + a];
          const int ax = px(ap) - r, ay = py(ap) - r;
          for (int d = lane; d < sp->n_doors; d += 32) {
            const uint16_t q = tb.door_pos[d];
            const int dx = px(q) - ax, dy = py(q) - ay;
            if (dx < 0 || dy < 0 || dx >= D || dy >= D) continue;
            const int cell = dx * D + dy;
            if ((vis >> cell) & 1)
              tile_add_frac(te, m, sp->ch_offset[a], cell, DD, ((dopen >> d) & 1) ? ENC_DOOR_OPEN : ENC_DOOR_CLOSED);
          }
        }
        __syncwarp();
      }
      // -- dirt piles (f64 amounts)
      if (sp->has_dirt) {
        for (int a = 0; a < A; ++a) {
          const uint32_t m = sp->term_chmask[a][MFG_G_DIRT];
          if (!m) continue;
          const unsigned long long vis = s_vis[el * A + a];
          const uint16_t ap = pos[sl.agent0 + a];
          const int ax = px(ap) - r, ay = py(ap) - r;
          for (int k = lane; k < sl.item0; k += 32) {
            const uint16_t q = pos[k];
            if (q == NO_POS) continue;
            const int dx = px(q) - ax, dy = py(q) - ay;
            if (dx < 0 || dy < 0 || dx >= D || dy >= D) continue;
            const int cell = dx * D + dy;
            if ((vis >> cell) & 1) tile_add_frac(te, m, sp->ch_offset[a], cell, DD, st.dirt_amt[(size_t)k * st.N + e]);
          }
        }
        __syncwarp();
      }
      // -- scalar channels: battery level / global position at flat index 0 (1)
      for (int a = lane; a < A; a += 32) {
        const int C = sp->n_channels[a], coff = sp->ch_offset[a];
        for (int c = 0; c < C; ++c) {
          const int kind = sp->ch_kind[a][c];
          if (kind == MFG_CH_BATTERY) te[(coff + c) * DD] = (float)st.bat[(size_t)a * st.N + e];
          else if (kind == MFG_CH_GLOBALPOS) {
            const uint16_t ap = pos[sl.agent0 + a];
            te[(coff + c) * DD] = (float)((double)px(ap) / (double)sp->H);
            te[(coff + c) * DD + 1] = (float)((double)py(ap) / (double)sp->W);
          }
        }
      }
    }
    __syncwarp();
    // stream the tile out: 128-bit, fully coalesced, write-once data -> streaming stores
    {
      const int ne = (int)((st.N - eg) < GE ? (st.N - eg) : GE);
      const int nfl = ne * total_channels * DD;
      float* dst = obs + (size_t)eg * total_channels * DD;
      if (ne == GE) {
        const float4* t4 = reinterpret_cast<const float4*>(tile);
        float4* d4 = reinterpret_cast<float4*>(dst);
        const int n4 = nfl >> 2;
        for (int i = lane; i < n4; i += 32) __stcs(&d4[i], t4[i]);
      } else {
        for (int i = lane; i < nfl; i += 32) dst[i] = tile[i];
      }
    }
    __syncwarp();
  }
}

// --------------------------------------------------------------------------------------------------------------
// C ABI
// --------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(call)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess) return fail(MFG_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
  } while (0)

struct MfgHandle {
  MfgSpec sp;                  // host copy (level pointers nulled)
  MfgSpec* d_sp = nullptr;
  WindowRays* d_wr = nullptr;
  Tables tb{};
  State st{};
  std::vector<FieldInfo> fields;
  std::vector<void*> dev_allocs;
  size_t state_bytes = 0;
  int64_t N = 0;
  int total_channels = 0, DD = 0;
  bool bound = false;
  int obs_kernel = 0;          // 0 = auto (tiled when possible), 1 = direct
  ObsSlots slots{};
  int obs_ge = 1;
  size_t obs_smem = 0;
  int obs_threads = 0;
  bool tiled_ok = false;
  // host-buffer path staging
  int32_t* d_actions = nullptr; float* d_reward = nullptr; uint8_t* d_done = nullptr; float* d_obs = nullptr;
  int64_t launches = 0;
};

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

template <typename Fn>
static int dispatch_amax(int A, Fn fn) {
  if (A <= 1) return fn(std::integral_constant<int, 1>());
  if (A <= 2) return fn(std::integral_constant<int, 2>());
  if (A <= 4) return fn(std::integral_constant<int, 4>());
  if (A <= 8) return fn(std::integral_constant<int, 8>());
  return fn(std::integral_constant<int, 16>());
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

  MfgHandle* h = new MfgHandle();
  h->sp = *spec;
  h->sp.walls = nullptr; h->sp.floor_pos = nullptr; h->sp.door_pos = nullptr; h->sp.nexthop = nullptr;
  h->N = n_envs;
  int rc;
#define UP(field) if ((rc = upload(h, ht.field, &h->tb.field)) != MFG_OK) { mfg_destroy(h); return rc; }
  UP(wall) UP(door_map) UP(floor_pos) UP(floor_index) UP(wall_uid) UP(wall_pos) UP(door_pos) UP(nexthop) UP(wall_win)
#undef UP
  h->tb.env_id_offset = env_id_offset;
  void* d = nullptr;
  if (cudaMalloc(&d, sizeof(unsigned long long) * MFG_N_STATS) != cudaSuccess) { mfg_destroy(h); return fail(MFG_E_NOMEM, "stats alloc"); }
  h->dev_allocs.push_back(d);
  cudaMemset(d, 0, sizeof(unsigned long long) * MFG_N_STATS);
  h->tb.stats = static_cast<unsigned long long*>(d);
  if (cudaMalloc(&d, sizeof(MfgSpec)) != cudaSuccess) { mfg_destroy(h); return fail(MFG_E_NOMEM, "spec alloc"); }
  h->dev_allocs.push_back(d);
  h->d_sp = static_cast<MfgSpec*>(d);
  cudaMemcpy(h->d_sp, &h->sp, sizeof(MfgSpec), cudaMemcpyHostToDevice);
  WindowRays wr;
  build_window_rays(h->sp, wr);
  if (cudaMalloc(&d, sizeof(WindowRays)) != cudaSuccess) { mfg_destroy(h); return fail(MFG_E_NOMEM, "ray alloc"); }
  h->dev_allocs.push_back(d);
  h->d_wr = static_cast<WindowRays*>(d);
  cudaMemcpy(h->d_wr, &wr, sizeof(WindowRays), cudaMemcpyHostToDevice);

  h->state_bytes = compute_layout(h->sp, n_envs, h->fields);
  h->total_channels = 0;
  for (int a = 0; a < h->sp.n_agents; ++a) h->total_channels += h->sp.n_channels[a];
  const int D = 2 * h->sp.pomdp_r + 1;
  h->DD = D * D;

  // tiled observation kernel configuration
  const MfgSpec& sp = h->sp;
  ObsSlots& sl = h->slots;
  sl.dirt0 = 0;
  sl.item0 = sp.has_dirt ? sp.dirt_slots : 0;
  sl.pod0 = sl.item0 + sp.n_items;
  sl.dest0 = sl.pod0 + sp.n_pods;
  sl.drop0 = sl.dest0 + sp.n_dest;
  sl.mach0 = sl.drop0 + sp.n_dropoff;
  sl.maint0 = sl.mach0 + sp.n_machines;
  sl.agent0 = sl.maint0 + sp.n_maint;
  sl.total = sl.agent0 + sp.n_agents;
  sl.stride = sl.total + ((sl.total & 1) ? 1 : 0);          // in uint16; make the 32-bit word stride odd
  if (((sl.stride / 2) & 1) == 0) sl.stride += 2;
  const int tcdd = h->total_channels * h->DD;
  h->obs_ge = (tcdd % 4 == 0) ? 1 : (tcdd % 2 == 0) ? 2 : 4;
  const int NW = sp.n_agents < 2 ? 2 : (sp.n_agents > 8 ? 8 : sp.n_agents);
  h->obs_threads = NW * 32;
  h->obs_smem = (size_t)NW * h->obs_ge * tcdd * sizeof(float) + (size_t)OBS_ENVS * sp.n_agents * 8 + OBS_ENVS * 8 +
                OBS_ENVS * 4 + (size_t)OBS_ENVS * sl.stride * 2 + 16;
  h->tiled_ok = !sp.faithful && h->obs_smem <= 200 * 1024 && wr.n > 0 && D * D <= 64;
  if (h->tiled_ok && h->obs_smem > 48 * 1024) {
    cudaError_t e1 = cudaFuncSetAttribute(k_obs_tiled<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->obs_smem);
    cudaError_t e2 = cudaFuncSetAttribute(k_obs_tiled<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->obs_smem);
    cudaError_t e4 = cudaFuncSetAttribute(k_obs_tiled<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->obs_smem);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e4 != cudaSuccess) h->tiled_ok = false;
  }
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
  delete h;
}

size_t mfg_state_bytes(const MfgHandle* h) { return h ? h->state_bytes : 0; }

int mfg_state_field(const MfgHandle* h, const char* name, MfgField* out) {
  if (!h || !name || !out) return fail(MFG_E_INVALID, "mfg_state_field: bad arguments");
  for (const auto& f : h->fields)
    if (strcmp(f.name, name) == 0) { out->offset = f.offset; out->rows = f.rows; out->elem_size = f.elem_size; return MFG_OK; }
  return fail(MFG_E_INVALID, std::string("mfg_state_field: unknown field ") + name);
}

int mfg_bind_state(MfgHandle* h, void* d_state) {
  if (!h || !d_state) return fail(MFG_E_INVALID, "mfg_bind_state: bad arguments");
  if (reinterpret_cast<uintptr_t>(d_state) % 256) return fail(MFG_E_INVALID, "mfg_bind_state: buffer must be 256-byte aligned");
  bind_state(h->sp, h->N, d_state, h->st);
  h->bound = true;
  return MFG_OK;
}

#define NEED_BOUND(h) if (!(h) || !(h)->bound) return fail(MFG_E_INVALID, "state buffer not bound (call mfg_bind_state)")

int mfg_reset(MfgHandle* h, const uint8_t* d_env_mask, void* stream) {
  NEED_BOUND(h);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int threads = 128;
  const unsigned blocks = (unsigned)((h->N + threads - 1) / threads);
  int rc = dispatch_amax(h->sp.n_agents, [&](auto amax) {
    k_reset<decltype(amax)::value><<<blocks, threads, 0, s>>>(h->d_sp, h->tb, h->st, d_env_mask);
    return MFG_OK;
  });
  (void)rc;
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return MFG_OK;
}

int mfg_step(MfgHandle* h, const int32_t* d_actions, const MfgTape* tape, float* d_reward, uint8_t* d_done,
             int auto_reset, void* stream) {
  NEED_BOUND(h);
  if (!d_actions || !d_reward || !d_done) return fail(MFG_E_INVALID, "mfg_step: NULL buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  StepIO io{d_actions, tape ? tape->d_maint_action : nullptr, tape ? tape->d_respawn_n : nullptr,
            tape ? tape->d_respawn_pos : nullptr, d_reward, d_done, auto_reset};
  bool need_policy = false;
  for (int r = 0; r < h->sp.n_rules; ++r) need_policy |= h->sp.rule_op[r] == MFG_R_MOVE_MAINTAINERS;
  if (need_policy && !io.maint_act && !h->tb.nexthop)
    return fail(MFG_E_INVALID, "mfg_step: MoveMaintainers without a tape needs the next-hop table (MfgSpec.nexthop)");
  const int threads = 128;
  const unsigned blocks = (unsigned)((h->N + threads - 1) / threads);
  dispatch_amax(h->sp.n_agents, [&](auto amax) {
    k_step<decltype(amax)::value><<<blocks, threads, 0, s>>>(h->d_sp, h->tb, h->st, io);
    return MFG_OK;
  });
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return MFG_OK;
}

int mfg_observe(MfgHandle* h, float* d_obs, void* stream) {
  NEED_BOUND(h);
  if (!d_obs) return fail(MFG_E_INVALID, "mfg_observe: NULL buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool tiled = h->tiled_ok && h->obs_kernel != 1;
  if (h->obs_kernel == 2 && !h->tiled_ok) return fail(MFG_E_UNSUPPORTED, "tiled observation kernel not available for this spec");
  if (tiled) {
    const unsigned blocks = (unsigned)((h->N + OBS_ENVS - 1) / OBS_ENVS);
    if (h->obs_ge == 1)
      k_obs_tiled<1><<<blocks, h->obs_threads, h->obs_smem, s>>>(h->d_sp, h->tb, h->st, h->d_wr, h->slots, d_obs, h->total_channels);
    else if (h->obs_ge == 2)
      k_obs_tiled<2><<<blocks, h->obs_threads, h->obs_smem, s>>>(h->d_sp, h->tb, h->st, h->d_wr, h->slots, d_obs, h->total_channels);
    else
      k_obs_tiled<4><<<blocks, h->obs_threads, h->obs_smem, s>>>(h->d_sp, h->tb, h->st, h->d_wr, h->slots, d_obs, h->total_channels);
  } else {
    const int threads = 128;
    const int64_t total = h->N * h->sp.n_agents;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    dispatch_amax(h->sp.n_agents, [&](auto amax) {
      k_obs_direct<decltype(amax)::value><<<blocks, threads, 0, s>>>(h->d_sp, h->tb, h->st, d_obs, h->total_channels);
      return MFG_OK;
    });
  }
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return MFG_OK;
}

int mfg_step_observe(MfgHandle* h, const int32_t* d_actions, const MfgTape* tape, float* d_reward, uint8_t* d_done,
                     float* d_obs, int auto_reset, void* stream) {
  int rc = mfg_step(h, d_actions, tape, d_reward, d_done, auto_reset, stream);
  if (rc != MFG_OK) return rc;
  return mfg_observe(h, d_obs, stream);
}

int mfg_random_actions(MfgHandle* h, int32_t* d_actions, uint64_t seed, uint64_t step_index, void* stream) {
  if (!h || !d_actions) return fail(MFG_E_INVALID, "mfg_random_actions: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int threads = 256;
  const unsigned blocks = (unsigned)((h->N + threads - 1) / threads);
  k_random_actions<<<blocks, threads, 0, s>>>(h->d_sp, h->N, h->tb.env_id_offset, d_actions, seed, (uint32_t)step_index);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return MFG_OK;
}

int mfg_step_host(MfgHandle* h, const int32_t* h_actions, float* h_reward, uint8_t* h_done, float* h_obs,
                  int auto_reset, void* stream) {
  NEED_BOUND(h);
  if (!h_actions || !h_reward || !h_done) return fail(MFG_E_INVALID, "mfg_step_host: NULL buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int A = h->sp.n_agents, NR = h->sp.individual_rewards ? A : 1;
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
    if (value < 0 || value > 2) return fail(MFG_E_INVALID, "obs_kernel must be 0 (auto), 1 (direct) or 2 (tiled)");
    if (value == 2 && !h->tiled_ok) return fail(MFG_E_UNSUPPORTED, "tiled observation kernel not available for this spec");
    h->obs_kernel = (int)value;
    return MFG_OK;
  }
  return fail(MFG_E_INVALID, std::string("mfg_set_option: unknown option ") + name);
}

int64_t mfg_get_info(const MfgHandle* h, const char* name) {
  if (!h || !name) return -1;
  if (strcmp(name, "launches") == 0) return h->launches;
  if (strcmp(name, "tiled_ok") == 0) return h->tiled_ok ? 1 : 0;
  if (strcmp(name, "obs_smem") == 0) return (int64_t)h->obs_smem;
  if (strcmp(name, "obs_threads") == 0) return h->obs_threads;
  if (strcmp(name, "total_channels") == 0) return h->total_channels;
  if (strcmp(name, "n_envs") == 0) return h->N;
  return -1;
}

}  // extern "C"
