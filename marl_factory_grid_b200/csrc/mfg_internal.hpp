// mfg_internal.hpp - handle layout and launch prototypes shared by the translation units of libmfg_b200.so
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <utility>
#include <vector>
#include "mfg_host.hpp"

namespace mfg {

constexpr int OBS_ENVS = 32;      // envs per CTA of the tiled observation kernel (one lane per env in phase 1)
constexpr int MAX_WRAYS = 64;
constexpr int MAX_WLEN = 4;

// run-time derived window rays (used to validate the generated constexpr tries and by nothing else)
struct WindowRays {
  int n;
  int len[MAX_WRAYS];
  uint8_t cell[MAX_WRAYS][MAX_WLEN];
  uint8_t da[MAX_WRAYS][MAX_WLEN];
  uint8_t db[MAX_WRAYS][MAX_WLEN];
};

// dynamic-position entity slots staged per env in shared memory by the observation kernel
struct ObsSlots {
  int dirt0, item0, pod0, dest0, drop0, mach0, maint0, agent0, total;
  int off_dopen, off_reached, prefix_bytes;    // layout of the staged block prefix (bytes)
  int off_door_listed, off_dirt_listed, off_dirt_uid, off_listed[6];   // faithful mode only (-1 / 0 = absent)
};

constexpr int MAX_WALL_PLANES = 48;
struct WallPlanes {               // channels that contain Walls, as (agent, plane index in the packed tensor)
  int n;
  uint16_t plane[MAX_WALL_PLANES];
  uint8_t agent[MAX_WALL_PLANES];
};

// word offsets of the constant table the tiled observation kernel copies into shared memory (built by plan_obs)
struct ObsProg {
  enum { N_SCL = 0, N_WP = 1, COFF = 2, HASBAT = COFF + MFG_MAX_AGENTS, SCL = HASBAT + MFG_MAX_AGENTS,
         WPLANE = SCL + MFG_MAX_AGENTS * 4, CHM = WPLANE + MAX_WALL_PLANES };
};

struct ObsPlan {
  bool ok = false;        // tiled kernel usable for this spec
  int ge = 1;             // (unused: one env per output tile)
  int bulk = 1;           // parts of 4k planes are whole numbers of 16-byte vectors: bulk (TMA) stores possible
  int ppp = 4;            // planes per part (the tile is composed in shared memory part by part)
  int nw = 4;             // warps per CTA
  int apad_log2 = 0;      // log2 of the agent count rounded up to a power of two (lanes per env in phase 1)
  int nbuf = 1;           // tile buffers per warp (2 = overlap the bulk store with the next env)
  int cap = 48;           // sprite slots per env (all agents share one list)
  int cap_max = 48;
  WallPlanes walls{};
  size_t smem = 0;
  ObsSlots slots{};
  std::vector<uint32_t> prog;        // ObsProg image
};

}  // namespace mfg

namespace mfg {
// Rows of the integer region of a state block: byte offset of the row's [128] slab inside the block | log2(element size) << 28.
// Lets a CTA copy the columns of arbitrary envs into a shared-memory block image (and back) with independent loads.
struct ColTab {
  const uint32_t* rows;
  int n;
};
// list mode of the tiled observation kernel: ids of the envs to observe, the prefix rows to gather
struct ObsList {
  const uint32_t* ids;
  const uint32_t* count;
  const uint32_t* rows;
  int n_rows;
};

#if defined(__CUDACC__)
// column j of the image <- env list[base + j], for j < n_here; all threads of the CTA take part
__device__ __forceinline__ void gather_columns(const State& st, unsigned char* stage, ColTab ct, const uint32_t* __restrict__ list,
                                               uint32_t base, int n_here) {
  const int total = ct.n * n_here;
#pragma unroll 4
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int r = idx / n_here, j = idx - r * n_here;
    const uint32_t rec = ct.rows[r], off = rec & 0x0FFFFFFFu, lg = rec >> 28;
    const int64_t e = list[base + j];
    const char* src = st.base_i + (size_t)(e >> 7) * st.blk_i + off + ((size_t)(e & (ENV_BLOCK - 1)) << lg);
    unsigned char* dst = stage + off + ((size_t)j << lg);
    if (lg == 0) *dst = __ldg(reinterpret_cast<const unsigned char*>(src));
    else if (lg == 1) *reinterpret_cast<uint16_t*>(dst) = __ldg(reinterpret_cast<const uint16_t*>(src));
    else if (lg == 2) *reinterpret_cast<uint32_t*>(dst) = __ldg(reinterpret_cast<const uint32_t*>(src));
    else *reinterpret_cast<unsigned long long*>(dst) = __ldg(reinterpret_cast<const unsigned long long*>(src));
  }
}
__device__ __forceinline__ void scatter_columns(const State& st, const unsigned char* stage, ColTab ct, const uint32_t* __restrict__ list,
                                                uint32_t base, int n_here) {
  const int total = ct.n * n_here;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int r = idx / n_here, j = idx - r * n_here;
    const uint32_t rec = ct.rows[r], off = rec & 0x0FFFFFFFu, lg = rec >> 28;
    const int64_t e = list[base + j];
    char* dst = st.base_i + (size_t)(e >> 7) * st.blk_i + off + ((size_t)(e & (ENV_BLOCK - 1)) << lg);
    const unsigned char* src = stage + off + ((size_t)j << lg);
    if (lg == 0) *reinterpret_cast<unsigned char*>(dst) = *src;
    else if (lg == 1) *reinterpret_cast<uint16_t*>(dst) = *reinterpret_cast<const uint16_t*>(src);
    else if (lg == 2) *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(src);
    else *reinterpret_cast<unsigned long long*>(dst) = *reinterpret_cast<const unsigned long long*>(src);
  }
}
// view of a shared-memory block image: same field offsets as `st`, block 0 == the image
__device__ __forceinline__ State staged_view(const State& st, unsigned char* stage) {
  State ss = st;
  ss.N = ENV_BLOCK;
  ss.base_i = reinterpret_cast<char*>(stage);
  const char* g0 = st.base_i;
#define F(type, name, rows_expr) \
  if constexpr (!std::is_same<type, double>::value) ss.name = reinterpret_cast<type*>(stage + (reinterpret_cast<const char*>(st.name) - g0));
  MFG_STATE_FIELDS(F)
#undef F
  return ss;
}
#endif
}  // namespace mfg

struct MfgHandle {
  MfgSpec sp;                  // host copy (level pointers nulled)
  MfgSpec* d_sp = nullptr;
  mfg::Tables tb{};
  mfg::State st{};
  std::vector<mfg::FieldInfo> fields;
  std::vector<void*> dev_allocs;
  size_t state_bytes = 0;
  int64_t N = 0;
  int total_channels = 0, DD = 0;
  bool bound = false;
  bool ever_reset = false;     // the first full (unmasked) reset seeds episode 0; later resets advance each env's episode counter
  int obs_kernel = 0;          // 0 = auto (tiled when possible), 1 = direct, 2 = tiled
  int obs_store = 1;           // 1 = TMA bulk store of the tile, 0 = LDS/STG loop
  mfg::ObsPlan plan;
  // host-buffer path staging
  int32_t* d_actions = nullptr; float* d_reward = nullptr; uint8_t* d_done = nullptr; float* d_obs = nullptr;
  int64_t launches = 0;
  uint8_t* d_flags = nullptr;         // optional per-step result flags (caller-owned, mfg_bind_step_flags)
  uint32_t* d_reset_list = nullptr;   // [N] ids of envs that finished in the current step
  uint32_t* d_reset_count = nullptr;
  uint32_t* d_row_tab = nullptr;      // ColTab rows (built at mfg_bind_state)
  int n_row_tab = 0;
  std::vector<uint32_t> row_tab_host;
  int step_blocks = 0;                // k_step: state blocks per CTA (0 = auto, 1..3; split image only)
  int step_kernel = 1;                // k_step: 1 = barriers at the convergent points + dirt uids left in HBM (6 CTAs per SM), 2 = barriers only, 0 = neither
  uint32_t* d_obs_prog = nullptr;     // device copy of plan.prog
  uint32_t* d_redo = nullptr;         // [1 + N] observation redo list: count, env ids (tiled kernel's rare exact path)
  int defer_reset = 1;
  // mfg_step_observe overlaps the packed re-spawn (side stream) with the observation kernel (caller's stream)
  int overlap_reset = 1;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // optional live kernel timing (mfg_set_option "timing"): event pairs recorded around the kernels, summed by mfg_get_info
  int timing = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> t_step, t_obs, t_reset;
};

namespace mfg {
// mfg_step.cu
cudaError_t launch_reset(MfgHandle* h, const uint8_t* d_mask, cudaStream_t s);
cudaError_t launch_step(MfgHandle* h, const StepIO& io, cudaStream_t s);
cudaError_t launch_step_kernel(MfgHandle* h, const StepIO& io, cudaStream_t s);
cudaError_t launch_reset_list(MfgHandle* h, const StepIO& io, cudaStream_t s);
cudaError_t launch_random_actions(MfgHandle* h, int32_t* d_actions, uint64_t seed, uint32_t step_index, cudaStream_t s);
// mfg_obs.cu
void plan_obs(MfgHandle* h);
int obs_ctas_per_sm(const MfgHandle* h);
void build_vis_tables(const MfgSpec& sp, HostTables& t);
void build_win_vis_tables(const MfgSpec& sp, HostTables& t);
cudaError_t launch_obs_direct(MfgHandle* h, float* d_obs, cudaStream_t s);
cudaError_t launch_obs_tiled(MfgHandle* h, float* d_obs, cudaStream_t s, const uint8_t* skip = nullptr);
cudaError_t launch_obs_list(MfgHandle* h, float* d_obs, cudaStream_t s, const uint32_t* d_list, const uint32_t* d_count);
cudaError_t launch_obs_tiled_list(MfgHandle* h, float* d_obs, cudaStream_t s, const uint32_t* d_list, const uint32_t* d_count);
}  // namespace mfg
