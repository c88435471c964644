// rcw_capi.cu — the C ABI of librcw_b200.so (include/rcw_b200.h): handle lifetime, HBM layout,
// host<->device plumbing and kernel launches.  No CPU implementation of the hot path lives here:
// without a CUDA device every entry point fails with RCW_ECUDA.

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost a pointer test unless a profiler is attached

#include <cuda_fp16.h>

#include "rcw_internal.h"

using namespace rcw;

// NVTX range around the enqueue work of an entry point (tracing)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------

static thread_local std::string t_last_error;

static int32_t fail(int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_last_error = buf;
    return code;
}

#define RCW_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            int32_t _c = (_e == cudaErrorMemoryAllocation) ? RCW_ENOMEM : RCW_ECUDA;           \
            cudaGetLastError();                                                                \
            return fail(_c, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                        __LINE__);                                                             \
        }                                                                                      \
    } while (0)

// Is the primary context of `dev` alive?  cudaGetDevice() answers 0 on a thread that never selected a device, and
// since CUDA 12 cudaSetDevice(0) CREATES that context (hundreds of MB and a slice of GPU 0 in every rank of a
// plain-C / Julia host that runs one process per GPU with cfg.device = rank and never calls cudaSetDevice
// itself).  The driver entry point is fetched through the runtime, so the library does not link libcuda.
static bool primary_context_active(int dev) {
    typedef int (*fn_t)(int /*CUdevice*/, unsigned int*, int*);
    static fn_t fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuDevicePrimaryCtxGetState", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            f = nullptr;
        }
        return reinterpret_cast<fn_t>(f);
    }();
    if (!fn) return true;   // cannot tell: behave as before
    unsigned int flags = 0;
    int active = 0;
    if (fn(dev, &flags, &active) != 0) return true;
    return active != 0;
}

// Selects the handle's device for the duration of a call.  The caller's device is restored only when it
// differs AND already has a context (see above): a guard never creates a context on a device the host did not use.
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    bool restore = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        restore = prev >= 0 && prev != dev && primary_context_active(prev);
        ok = (prev == dev && primary_context_active(dev)) || cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (restore) cudaSetDevice(prev);
    }
};

// constant-memory slots for direction tables, per device
static std::mutex g_slot_mutex;
static bool g_slot_used[64][kDirSlots];

static int acquire_dir_slot(int device) {
    if (device < 0 || device >= 64) return -1;
    std::lock_guard<std::mutex> lk(g_slot_mutex);
    for (int s = 0; s < kDirSlots; ++s)
        if (!g_slot_used[device][s]) {
            g_slot_used[device][s] = true;
            return s;
        }
    return -1;
}

static void release_dir_slot(int device, int slot) {
    if (device < 0 || device >= 64 || slot < 0) return;
    std::lock_guard<std::mutex> lk(g_slot_mutex);
    g_slot_used[device][slot] = false;
}

// ------------------------------------------------------------------------------------------
// the handle
// ------------------------------------------------------------------------------------------

constexpr int kActionRing = 4;
// env_kernel (one warp per env) is used for launches of at least this many envs: below it the grid of the
// item kernel (one warp per 32-ray group) fills the 148 SMs better
constexpr int64_t kEnvPerWarpMinEnvs = 2048;

struct rcw_batch {
    rcw_config cfg{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;      // second half-batch of the steps of a multi-step call (never visible to the caller)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_step[2] = {nullptr, nullptr}, ev_top[2] = {nullptr, nullptr};   // step / top view pipeline of multi-step calls
    bool top_pipeline = true;            // RCW_TOP_PIPELINE=0: top views in multi-step calls as two half-batches like the steps
    bool two_streams = true;             // RCW_TWO_STREAMS=0: every step is one launch on `stream`
    int64_t two_streams_min = 512;       // smallest batch that is split (below it a step is launch-latency bound anyway)
    int sm_count = 0;
    int bpp = 3;
    int gpe = 0;
    int wpr = 0;
    int map_words = 0;
    int dir_slot = -1;
    int ctas_per_sm = 0;          // 0: one CTA per 8 items; >0: persistent grid of sm_count * this
    bool bulk = false;            // renderer: TMA bulk stores of whole bands (true) or per-lane vector stores
    bool split = false;           // one env-step = front launch + paint launch (true) or one fused launch
    bool no_packed_actions = false;  // RCW_PACKED_ACTIONS=0: always stage host actions through a device copy
    PackedActions packed{};       // scratch for host actions delivered through the kernel parameters
    bool env_per_warp = false;    // narrow cameras (<= 128 rays): a warp owns a whole env (env_kernel) ...
    int64_t env_per_warp_min = 0; // ... in launches of at least this many envs
    bool occ4 = false;            // fused kernel variant compiled for 32 instead of 24 warps per SM (front-bound steps)
    uint32_t* d_col_info = nullptr;
    int pat_stride = 0;
    uint8_t* d_patterns = nullptr;
    // device memory
    float2* d_dirs = nullptr;
    float4* d_ray_table = nullptr;
    uint32_t* d_wall_map = nullptr;      // layers shared by the batch: [wall][extra 0..3][any], map_words each (see BitsMap);
                                         // the `any` layer sits right behind the last extra layer in use
    int n_extra = 0;                     // extra object layers (rcw_config.num_object_layers - 2)
    std::vector<uint32_t> h_layers;      // host copy of [wall][extra 0 .. n_extra-1], to rebuild `any` when one changes
    uint32_t* d_wall_maps_env = nullptr; // [num_envs][map_words], allocated by rcw_set_wall_maps
    bool per_env_maps = false;
    bool closed_border = true;           // every border tile of the active wall layer(s) is a wall
    bool room = true;                    // the shared wall layer is exactly the border of the map (SingleRoom's own map,
                                         // single_room.jl:57-60): kernels without a wall layer in shared memory (RoomMap)
    bool room_allowed = true;            // RCW_ROOM=0 keeps every launch on the bit-packed wall layer (tests, A/B)
    bool room_forced = false;            // RCW_ROOM=2: RoomMap kernels also where the store stream bounds the step
    uint8_t* d_col_table = nullptr;      // ready-made columns for env_kernel's table renderer (small columns only)
    StateRef st[2]{};
    int cur = 0;
    float* d_reward = nullptr;
    uint8_t* d_done = nullptr;
    float* d_ep_return = nullptr;
    uint32_t* d_ep_length = nullptr;
    DeviceStats* d_stats = nullptr;
    uint8_t* d_actions = nullptr;
    uint8_t* d_tape = nullptr;          // device copy of a host action tape (rcw_step_tape), grown on demand
    size_t tape_capacity = 0;
    uint8_t* d_obs = nullptr;
    size_t obs_env_stride = 0;    // bytes between consecutive envs (all ring positions of an env)
    size_t frame_stride = 0;      // bytes between consecutive ring positions of one env (rcw_config.frame_stack)
    int frame_stack = 1;
    int frame_newest = 0;         // ring position of the newest frame
    int64_t obs_window = 0;       // env slots of the observation buffer (num_envs unless cfg.obs_window_envs)
    int col_pitch = 0;            // bytes between consecutive columns of an observation (multiple of 32)
    size_t obs_bytes = 0;
    // rcw_reset scratch (host layouts and mask), allocated on first use
    int32_t* d_reset_goal = nullptr;
    int32_t* d_reset_player = nullptr;
    int32_t* d_reset_dir = nullptr;
    uint8_t* d_reset_mask = nullptr;
    // top view (update_top_view!): allocated when first drawn
    uint8_t* d_top = nullptr;
    size_t top_env_stride = 0;
    // pinned staging for host-side action arrays
    uint8_t* h_actions[kActionRing]{};
    cudaEvent_t h_actions_free[kActionRing]{};
    int ring = 0;
    DeviceStats* h_stats = nullptr;  // pinned
    uint8_t* h_reward_done = nullptr; // pinned staging: reward f32[E] followed by done u8[E]
    // result ring (rcw_config.result_ring = D): D slots of {reward f32[E], done u8[E]} in pinned, mapped host
    // memory that the step kernel of rcw_step_async writes through to; one event per slot marks its step done
    int result_ring = 0;
    uint8_t* h_results = nullptr;     // host address of slot 0
    uint8_t* d_results = nullptr;     // the same memory as the device sees it
    size_t result_slot_bytes = 0;     // 5 * E rounded up to 128
    cudaEvent_t result_ready[64]{};
    int64_t next_ticket = 0;
    int result_slot = -1;             // >= 0 while rcw_step_async enqueues its step
    bool device_actions_pending = false;  // a device-side action array was used since the last check
    bool expand_pending = false;          // rcw_expand_columns ran since the last check (caller-supplied words)
    // counters
    uint64_t step_index = 0;
    int64_t launches = 0;
    std::vector<void*> allocs;
};

template <typename T>
static cudaError_t dev_alloc(rcw_batch* b, T** out, size_t count, bool zero = true) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e != cudaSuccess) return e;
    b->allocs.push_back(p);
    *out = static_cast<T*>(p);
    if (zero) e = cudaMemsetAsync(p, 0, count * sizeof(T), b->stream);
    return e;
}

static int bytes_per_pixel(int fmt) {
    return fmt == RCW_OBS_RGB8 ? 3 : ((fmt == RCW_OBS_GRAY8 || fmt == RCW_OBS_GRAY8_HALF) ? 1 : (fmt == RCW_OBS_GRAY16F ? 2 : 4));
}

// columns / rows of one observation of a handle: the camera's, or half of each under the 2 x 2 box filter
static int obs_columns(const rcw_config& c) { return c.obs_format == RCW_OBS_GRAY8_HALF ? c.num_rays / 2 : c.num_rays; }
static int obs_rows(const rcw_config& c) {
    return c.obs_format == RCW_OBS_GRAY8_HALF ? c.height_camera_view_pu / 2 : c.height_camera_view_pu;
}

// BT.601 luma of a 0x00RRGGBB pixel (RCW_OBS_GRAY8), and the same luma / 255 as IEEE binary16 bits (RCW_OBS_GRAY16F)
static uint32_t luma_of(uint32_t col) {
    return (77u * ((col >> 16) & 255u) + 150u * ((col >> 8) & 255u) + 29u * (col & 255u) + 128u) >> 8;
}
static uint32_t luma_half_bits(uint32_t col) {
    volatile float v = (float)luma_of(col) / 255.0f;     // one binary32 division, then one rounding to binary16
    return (uint32_t)__half_as_ushort(__float2half_rn(v));
}

// bytes of one column of an observation: the pixels of the column, or its one RCW_OBS_COLUMNS word
static int column_bytes(const rcw_batch* b) {
    return b->cfg.obs_format == RCW_OBS_COLUMNS ? 4 : obs_rows(b->cfg) * b->bpp;
}

// pixel_fmt: the format the renderer fields are prepared for (-1: the handle's own; rcw_expand_columns paints
// a format of the caller's choice; a RCW_OBS_COLUMNS step paints nothing)
static void fill_frame_params(const rcw_batch* b, FrameParams& p, int pixel_fmt = -1) {
    const rcw_config& c = b->cfg;
    if (pixel_fmt < 0) pixel_fmt = c.obs_format == RCW_OBS_COLUMNS ? (int)RCW_OBS_RGB8 : c.obs_format;
    const int px_bpp = bytes_per_pixel(pixel_fmt);
    const int px_col_bytes = (pixel_fmt == RCW_OBS_GRAY8_HALF ? c.height_camera_view_pu / 2 : c.height_camera_view_pu) * px_bpp;
    const int px_col_pitch = (px_col_bytes + 31) & ~31;
    memset(&p, 0, sizeof(p));
    p.H = c.height_tile_map_tu;
    p.W = c.width_tile_map_tu;
    p.wpr = b->wpr;
    p.map_words = b->map_words;
    p.n_extra = b->per_env_maps ? 0 : b->n_extra;
    p.stage_words = p.n_extra ? (p.n_extra + 2) * b->map_words : b->map_words;
    p.n_colors = 4 + 2 * b->n_extra;
    for (int k = 0; k < b->n_extra; ++k) {
        if (c.layer_kind[k] == RCW_LAYER_TERMINAL) p.layer_terminal_mask |= 1u << k;
        p.layer_reward[k] = c.layer_reward[k];
    }
    p.N = c.num_directions;
    p.R = c.num_rays;
    p.P = c.height_camera_view_pu;
    p.gpe = b->gpe;
    p.px_bytes = px_bpp;
    p.col_bytes = px_col_bytes;
    p.col_pitch = px_col_pitch;
    p.dda_flags = c.dda_flags;
    p.closed_border = b->closed_border ? 1u : 0u;
    p.room = (b->room && b->room_allowed && !b->per_env_maps) ? 1u : 0u;
    p.col_table = (pixel_fmt == c.obs_format) ? b->d_col_table : nullptr;
    p.radius = c.player_radius_wu;
    p.incr = c.position_increment_wu;
    p.goal_reward = c.goal_reward;
    // single_room.jl:406 — camera_height_tile_wu * num_rays and 2 * semi_field_of_view_wu, each one
    // binary32 rounding (volatile keeps the host compiler from widening or folding differently)
    volatile float hl_num = c.camera_height_tile_wu * (float)c.num_rays;
    volatile float two_s = 2.0f * c.semi_field_of_view_wu;
    p.hl_num = hl_num;
    p.two_s = two_s;
    const int n_pal = 6 + 2 * b->n_extra;
    for (int i = 0; i < n_pal; ++i) {
        const uint32_t col = (i < 6 ? c.palette[i] : c.layer_palette[(i - 6) >> 1][(i - 6) & 1]) & 0x00FFFFFFu;
        if (pixel_fmt == RCW_OBS_GRAY8 || pixel_fmt == RCW_OBS_GRAY8_HALF) {
            // BT.601 luma of the reference pixel, replicated so the colour is "flat" for the renderer
            p.palette[i] = luma_of(col) * 0x00010101u;
        } else if (pixel_fmt == RCW_OBS_GRAY16F) {
            p.palette[i] = luma_half_bits(col) * 0x00010001u;   // two pixels per word
        } else {
            p.palette[i] = col;
        }
    }
    {   // what the renderer stores per column for each palette entry (see FrameParams::col_entry)
        auto flat = [&](uint32_t col) { return pixel_fmt != RCW_OBS_RGB8 || ((col ^ (col >> 8)) & 0xFFFFu) == 0; };
        const bool cf = flat(p.palette[RCW_COLOR_CEILING]) && flat(p.palette[RCW_COLOR_FLOOR]);
        for (int i = 0; i < n_pal; ++i) {
            const bool slow = !(cf && flat(p.palette[i]));
            const uint32_t col = p.palette[i];
            const uint32_t word = (pixel_fmt == RCW_OBS_XRGB32 || pixel_fmt == RCW_OBS_GRAY16F) ? col : (col & 0xFFu) * 0x01010101u;
            p.col_entry[i] = make_uint2(slow ? 0x80000000u : 0u, slow ? col : word);
        }
    }
    if (b->gpe >= 2) {   // libdivide's branch-free u32 divider for d = gpe
        const uint32_t d = (uint32_t)b->gpe;
        uint32_t k = 31;
        while (!(d >> k)) --k;                                   // floor(log2 d)
        if ((d & (d - 1)) == 0) {
            p.gpe_magic = 0;
            p.gpe_shift = k - 1;
        } else {
            const uint64_t num = 1ULL << (32 + k);
            uint64_t m = num / d;
            const uint64_t rem = num % d;
            m += m;
            if (2 * rem >= d) m += 1;
            p.gpe_magic = (uint32_t)(m + 1);                     // low 32 bits of the 33-bit multiplier
            p.gpe_shift = k;
        }
    }
    {   // renderer sweep units per column: mirror pairs of 64 bytes when col_bytes % 64 == 0, else 32-byte sectors
        const int col_bytes = px_col_bytes;
        const int U = (col_bytes & 63) == 0 ? (col_bytes >> 6) : (px_col_pitch >> 5);
        p.unit_inv16 = U < 32 ? (uint32_t)((65536 + U - 1) / U) : 0u;
        p.unit_adv_cl = 32 / U;
        p.unit_adv_u = 32 % U;
        const int NS = px_col_pitch >> 5;
        p.sec_inv16 = NS < 32 ? (uint32_t)((65536 + NS - 1) / NS) : 0u;
        p.sec_adv_cl = 32 / NS;
        p.sec_adv_u = 32 % NS;
    }
    p.dir_slot = b->dir_slot;
    p.dirs = b->d_dirs;
    p.ray_table = b->d_ray_table;
    p.wall_map = b->per_env_maps ? b->d_wall_maps_env : b->d_wall_map;
    p.map_env_stride = b->per_env_maps ? (uint32_t)b->map_words : 0u;
    p.patterns = b->d_patterns;
    p.pat_stride = b->pat_stride;
    p.col_info = b->d_col_info;
    p.col_info_stride = (uint32_t)c.num_rays;
    p.in = b->st[b->cur];
    p.out = b->st[b->cur ^ 1];
    p.actions = nullptr;
    p.reward = b->d_reward;
    p.done = b->d_done;
    if (b->result_slot >= 0) {
        uint8_t* slot = b->d_results + (size_t)b->result_slot * b->result_slot_bytes;
        p.host_reward = reinterpret_cast<float*>(slot);
        p.host_done = slot + (size_t)c.num_envs * 4;
    }
    p.ep_return = b->d_ep_return;
    p.ep_length = b->d_ep_length;
    p.stats = b->d_stats;
    p.obs = b->d_obs + (size_t)b->frame_newest * b->frame_stride;
    p.obs_env_stride = b->obs_env_stride;
    if (c.obs_format == RCW_OBS_COLUMNS) {   // the front stage's column words are the observation
        p.col_info = reinterpret_cast<uint32_t*>(p.obs);
        p.col_info_stride = (uint32_t)(b->obs_env_stride / 4);
    }
    p.obs_window = (uint32_t)b->obs_window;
    p.obs_slot0 = 0;
    p.num_envs = c.num_envs;
    p.env_first = 0;
    p.env_count = c.num_envs;
    p.env_id_offset = (uint64_t)c.env_id_offset;
    p.seed = c.seed;
    p.step_index = b->step_index;
    p.auto_reset = c.auto_reset;
}

static int grid_for(const rcw_batch* b, int64_t env_count) {
    const int64_t items = env_count * b->gpe;
    int64_t ctas = (items + kWarpsPerCta - 1) / kWarpsPerCta;
    if (b->ctas_per_sm > 0 && !b->per_env_maps) {   // per-env wall layers: one round per CTA
        const int64_t cap = (int64_t)b->sm_count * b->ctas_per_sm;
        if (ctas > cap) ctas = cap;
    }
    return (int)(ctas < 1 ? 1 : ctas);
}

// how a step / render launch of n envs is shaped
static LaunchShape shape_for(const rcw_batch* b, int64_t n) {
    LaunchShape sh{b->bulk, b->split, b->occ4, grid_for(b, n)};
    sh.env_per_warp = b->env_per_warp && n >= b->env_per_warp_min;
    // RoomMap kernels where act! + DDA bound the step: a warp owns a whole env (env_kernel), nothing is painted (column
    // words), or the item kernel's items are small (32 columns under 16 KB: 512x256 GRAY8 1.064 -> 1.133 of the copy peak,
    // 160x120 RGB8 1.026 -> 1.041).  Where the store stream bounds the item kernel the bit-packed variant is faster even
    // though it executes more instructions — default camera 0.2270 vs 0.2296 ms per 4096 envs, 256x192 RGB8 1.135 vs 1.123,
    // 64x64 maps at the default camera (config 5, 20 DDA iterations per item) 3.527 vs 3.611 ms per 65,536 envs, all A/B on
    // one box — so it stays there; RCW_ROOM=2 forces RoomMap.
    sh.room = b->room && b->room_allowed && !b->per_env_maps &&
              (sh.env_per_warp || b->cfg.obs_format == RCW_OBS_COLUMNS || 32 * b->col_pitch < 16384 || b->room_forced);
    sh.table = sh.env_per_warp && b->d_col_table != nullptr;
    return sh;
}

// update_top_view! for envs [env0, env0 + n) from state `st`, into the slots that start at slot0.
static int32_t enqueue_top_view(rcw_batch* b, const StateRef& st, int64_t env0, int64_t n, uint32_t slot0,
                                const uint8_t* d_mask = nullptr, cudaStream_t stream = nullptr) {
    if (!stream) stream = b->stream;
    const rcw_config& c = b->cfg;
    if (!b->d_top) {
        const size_t px = (size_t)c.height_tile_map_tu * c.pu_per_tu * (size_t)c.width_tile_map_tu * c.pu_per_tu;
        b->top_env_stride = (px * 4 + 127) & ~(size_t)127;
        RCW_CUDA(dev_alloc(b, &b->d_top, b->top_env_stride * (size_t)b->obs_window, false));
    }
    TopViewParams t;
    memset(&t, 0, sizeof(t));
    t.H = c.height_tile_map_tu;
    t.W = c.width_tile_map_tu;
    t.wpr = b->wpr;
    t.map_words = b->map_words;
    t.n_extra = b->per_env_maps ? 0 : b->n_extra;
    t.stage_words = t.n_extra ? (t.n_extra + 2) * b->map_words : b->map_words;
    for (int k = 0; k < b->n_extra; ++k) t.extra_color[k] = c.layer_top_color[k] & 0x00FFFFFFu;
    t.N = c.num_directions;
    t.R = c.num_rays;
    t.pu = c.pu_per_tu;
    t.Hp = t.H * t.pu;
    t.Wp = t.W * t.pu;
    t.dda_flags = c.dda_flags;
    t.closed_border = b->closed_border ? 1u : 0u;
    t.radius = c.player_radius_wu;
    for (int i = 0; i < 6; ++i) t.palette[i] = c.top_palette[i] & 0x00FFFFFFu;
    t.dir_slot = b->dir_slot;
    t.dirs = b->d_dirs;
    t.ray_table = b->d_ray_table;
    t.wall_map = b->per_env_maps ? b->d_wall_maps_env : b->d_wall_map;
    t.map_env_stride = b->per_env_maps ? (uint32_t)b->map_words : 0u;
    t.st = st;
    t.mask = d_mask;
    t.top = b->d_top;
    t.env_stride = b->top_env_stride;
    t.window = (uint32_t)b->obs_window;
    t.slot0 = slot0;
    t.env_first = env0;
    t.env_count = n;
    t.sm_count = (uint32_t)b->sm_count;
    t.room = (b->room && b->room_allowed && !b->per_env_maps) ? 1u : 0u;
    RCW_CUDA(launch_top_view(t, stream));
    b->launches += 1;
    return RCW_OK;
}

// One frame of the whole batch.  With an observation window the batch is rendered window by window
// (one launch each, every frame still written to HBM); all launches read state `cur` and write `cur ^ 1`.
// 2 bits per env into the kernel-parameter block (validated values 1..4)
static void pack_actions(const uint8_t* actions, int64_t n, PackedActions& pa) {
    int64_t k = 0;
    for (int64_t w = 0; k + 16 <= n; ++w, k += 16) {
        uint32_t v = 0;
        for (int j = 0; j < 16; ++j) v |= (uint32_t)(actions[k + j] - 1u) << (2 * j);
        pa.w[w] = v;
    }
    if (k < n) {
        uint32_t v = 0;
        for (int j = 0; k + j < n; ++j) v |= (uint32_t)(actions[k + j] - 1u) << (2 * j);
        pa.w[k >> 4] = v;
    }
}

// launches of at most kPackedActionEnvs envs on the fused path take host actions through the kernel parameters
static bool packs_actions(const rcw_batch* b, int64_t env_count) {
    return (b->cfg.obs_format == RCW_OBS_COLUMNS || (!b->split && !b->bulk)) && !b->no_packed_actions &&
           env_count <= kPackedActionEnvs;
}

// h_actions != nullptr: validated host actions of the whole batch (d_actions is then ignored)
// d_render_mask (kModeRender): redraw only the envs whose byte is nonzero
static int32_t enqueue_frame(rcw_batch* b, int mode, const uint8_t* d_actions, const uint8_t* h_actions = nullptr,
                             const uint8_t* d_render_mask = nullptr) {
    if (mode == kModeStep) b->frame_newest = (b->frame_newest + 1) % b->frame_stack;   // a step writes the next ring position
    FrameParams p;
    fill_frame_params(b, p);
    p.actions = h_actions ? nullptr : d_actions;
    p.render_mask = (mode == kModeRender && !b->split) ? d_render_mask : nullptr;   // (a split launch redraws every env)
    const int64_t E = b->cfg.num_envs;
    for (int64_t e0 = 0; e0 < E; e0 += b->obs_window) {
        p.env_first = e0;
        p.env_count = E - e0 < b->obs_window ? E - e0 : b->obs_window;
        p.obs_slot0 = 0;
        const LaunchShape sh = shape_for(b, p.env_count);
        if (h_actions) {
            pack_actions(h_actions + e0, p.env_count, b->packed);
            RCW_CUDA(launch_frame(p, mode, b->cfg.obs_format, sh, b->stream, &b->packed));
        } else {
            RCW_CUDA(launch_frame(p, mode, b->cfg.obs_format, sh, b->stream));
        }
        b->launches += b->split ? 2 : 1;
        // the reference's act!(env) / reset!(env) also redraw the top view (single_room.jl:329,337)
        if (b->cfg.top_view)
            if (int32_t rc = enqueue_top_view(b, mode == kModeStep ? p.out : p.in, e0, p.env_count, 0, p.render_mask))
                return rc;
    }
    if (mode == kModeStep) {
        b->cur ^= 1;
        b->step_index += 1;
    }
    return RCW_OK;
}

// act! for the envs [env0, env0 + n) only: the launch reads state `cur` and writes `cur ^ 1` like a
// full step, then the range is copied back so that `cur` stays the handle's one current buffer.
// h_actions != nullptr: validated host actions of the range, delivered through the kernel parameters
static int32_t enqueue_range_step(rcw_batch* b, const uint8_t* d_actions_env0, int64_t env0, int64_t n,
                                  const uint8_t* h_actions = nullptr) {
    FrameParams p;
    fill_frame_params(b, p);
    p.actions = h_actions ? nullptr : d_actions_env0 - env0;   // the kernel indexes actions by env
    p.env_first = env0;
    p.env_count = n;
    p.obs_slot0 = (uint32_t)(env0 % b->obs_window);
    const LaunchShape sh = shape_for(b, n);
    if (h_actions) {
        pack_actions(h_actions, n, b->packed);
        RCW_CUDA(launch_frame(p, kModeStep, b->cfg.obs_format, sh, b->stream, &b->packed));
    } else {
        RCW_CUDA(launch_frame(p, kModeStep, b->cfg.obs_format, sh, b->stream));
    }
    RCW_CUDA(launch_commit_range(b->st[b->cur ^ 1], b->st[b->cur], env0, n, b->stream));
    b->launches += (b->split ? 2 : 1) + 1;
    if (b->cfg.top_view) return enqueue_top_view(b, b->st[b->cur], env0, n, p.obs_slot0);
    return RCW_OK;
}

static int32_t check_handle(const rcw_batch* b) {
    if (!b) return fail(RCW_EINVAL, "null rcw_batch handle");
    return RCW_OK;
}

// blocks, then reports (and clears) a device-side invalid-action flag
static int32_t sync_and_check(rcw_batch* b, bool need_stats = false) {
    // the stats block is only fetched when somebody needs it: the caller, or a device-side action
    // array whose values could not be validated on the host
    const bool fetch = need_stats || b->device_actions_pending || b->expand_pending;
    if (fetch)
        RCW_CUDA(cudaMemcpyAsync(b->h_stats, b->d_stats, sizeof(DeviceStats), cudaMemcpyDeviceToHost,
                                 b->stream));
    RCW_CUDA(cudaStreamSynchronize(b->stream));
    if (fetch) b->device_actions_pending = b->expand_pending = false;
    if (fetch && b->h_stats->bad_action) {
        RCW_CUDA(cudaMemsetAsync(&b->d_stats->bad_action, 0, sizeof(int), b->stream));
        return fail(RCW_EACTION, "a device-side action array held a value outside 1..4; "
                                 "the affected envs were not stepped");
    }
    if (fetch && b->h_stats->bad_columns) {
        RCW_CUDA(cudaMemsetAsync(&b->d_stats->bad_columns, 0, sizeof(int), b->stream));
        return fail(RCW_EINVAL, "rcw_expand_columns read a column word outside the format (palette index not in "
                                "2..5 or more ceiling rows than half the column); the word was clamped");
    }
    return RCW_OK;
}

// Host layers -> device [wall][extra ...][any], and what the kernels may assume about them: `closed` (every border
// tile carries an object: rays cannot leave the map) and `room` (the wall layer is exactly the border and there are
// no extra layers: RoomMap kernels).  Blocks until the previous work on the stream has finished.
static int32_t upload_layers(rcw_batch* b) {
    const int H = b->cfg.height_tile_map_tu, W = b->cfg.width_tile_map_tu, mw = b->map_words, L = 1 + b->n_extra;
    std::vector<uint32_t> dev((size_t)(L + 1) * mw, 0u);
    std::copy(b->h_layers.begin(), b->h_layers.end(), dev.begin());
    uint32_t* any = dev.data() + (size_t)L * mw;
    for (int l = 0; l < L; ++l)
        for (int k = 0; k < mw; ++k) any[k] |= b->h_layers[(size_t)l * mw + k];
    bool closed = true, interior = false;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const bool obj = (any[(size_t)i * b->wpr + (j >> 5)] >> (j & 31)) & 1u;
            const bool border = i == 0 || i == H - 1 || j == 0 || j == W - 1;
            if (obj && !border) interior = true;
            if (!obj && border) closed = false;
        }
    RCW_CUDA(cudaStreamSynchronize(b->stream));
    // without extra layers only the wall layer is staged (the kernels read it as `any`)
    RCW_CUDA(cudaMemcpy(b->d_wall_map, dev.data(), sizeof(uint32_t) * (b->n_extra ? dev.size() : (size_t)mw), cudaMemcpyHostToDevice));
    b->per_env_maps = false;
    b->closed_border = closed;
    b->room = closed && !interior && b->n_extra == 0;
    return RCW_OK;
}

static void pack_tiles(const uint8_t* tiles, int H, int W, int wpr, uint32_t* words) {
    for (int j = 0; j < W; ++j)
        for (int i = 0; i < H; ++i)
            if (tiles[(size_t)j * H + i]) words[(size_t)i * wpr + (j >> 5)] |= 1u << (j & 31);
}

static void pack_border_walls(int H, int W, int wpr, std::vector<uint32_t>& words) {
    // tile_map[WALL, :, 1] = tile_map[WALL, :, W] = tile_map[WALL, 1, :] = tile_map[WALL, H, :] = true
    // (single_room.jl:57-60)
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j)
            if (i == 0 || i == H - 1 || j == 0 || j == W - 1) words[(size_t)i * wpr + (j >> 5)] |= 1u << (j & 31);
}

// the top view of one env is composed in shared memory (two bit planes + tile tables): bounded by the SM
static int32_t check_top_view_fits(const rcw_config& c) {
    const int64_t Hp = (int64_t)c.height_tile_map_tu * c.pu_per_tu, Wp = (int64_t)c.width_tile_map_tu * c.pu_per_tu;
    const int layers = c.num_object_layers > 2 ? c.num_object_layers : 1;     // staged: [wall][extras][any], or the wall layer alone
    const int map_words = layers * (((c.height_tile_map_tu * ((c.width_tile_map_tu + 31) / 32) + 3) / 4) * 4);
    if (Hp > 32767 || Wp > 32767 || Hp * Wp >= (1LL << 28) ||
        top_view_smem_bytes(c.height_tile_map_tu, c.width_tile_map_tu, c.num_rays, c.pu_per_tu, c.player_radius_wu, map_words) > 200 * 1024)
        return fail(RCW_ESIZE, "a top view of %lldx%lld pixels does not fit the renderer's shared memory; "
                    "lower pu_per_tu", (long long)Hp, (long long)Wp);
    return RCW_OK;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------

extern "C" {

int32_t rcw_version(void) { return RCW_ABI_VERSION; }

const char* rcw_last_error(void) { return t_last_error.c_str(); }

int32_t rcw_config_init(rcw_config* cfg) {
    if (!cfg) return fail(RCW_EINVAL, "cfg is null");
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = (uint32_t)sizeof(rcw_config);
    cfg->device = 0;
    cfg->num_envs = 1;
    cfg->env_id_offset = 0;
    cfg->height_tile_map_tu = 8;        // single_room.jl:44
    cfg->width_tile_map_tu = 16;        // :45
    cfg->num_directions = 128;          // :46
    cfg->num_rays = 512;                // :52
    cfg->height_camera_view_pu = 256;   // :271
    cfg->player_radius_wu = (float)(1.0 / 8.0);       // :47
    cfg->position_increment_wu = (float)(1.0 / 8.0);  // :48
    cfg->semi_field_of_view_wu = (float)(2.0 / 3.0);  // :51
    cfg->camera_height_tile_wu = 1.0f;  // :270
    cfg->goal_reward = 1.0f;            // :82
    cfg->obs_format = RCW_OBS_RGB8;
    cfg->auto_reset = 1;
    cfg->seed = 0;
    cfg->palette[RCW_COLOR_CEILING] = 0x00FFFFFFu;  // :292
    cfg->palette[RCW_COLOR_FLOOR] = 0x00404040u;    // :291
    cfg->palette[RCW_COLOR_WALL_1] = 0x00808080u;   // :293
    cfg->palette[RCW_COLOR_WALL_2] = 0x00c0c0c0u;   // :294
    cfg->palette[RCW_COLOR_GOAL_1] = 0x00800000u;   // :295
    cfg->palette[RCW_COLOR_GOAL_2] = 0x00c00000u;   // :296
    cfg->dda_flags = 0;
    cfg->obs_window_envs = 0;
    cfg->top_view = 0;
    cfg->pu_per_tu = 32;                                   // :269
    cfg->top_palette[RCW_TOP_COLOR_WALL] = 0x00FFFFFFu;    // tile_map_colors :288
    cfg->top_palette[RCW_TOP_COLOR_GOAL] = 0x00FF0000u;
    cfg->top_palette[RCW_TOP_COLOR_EMPTY] = 0x00000000u;
    cfg->top_palette[RCW_TOP_COLOR_BORDER] = 0x00ccccccu;  // :364-367
    cfg->top_palette[RCW_TOP_COLOR_RAY] = 0x00808080u;     // ray_color :289
    cfg->top_palette[RCW_TOP_COLOR_PLAYER] = 0x00c0c0c0u;  // player_color :290
    cfg->num_object_layers = 2;                            // NUM_OBJECTS :16
    return RCW_OK;
}

int32_t rcw_destroy(rcw_batch* b) {
    if (!b) return RCW_OK;
    DeviceGuard g(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    for (void* p : b->allocs) cudaFree(p);
    for (int i = 0; i < kActionRing; ++i) {
        if (b->h_actions[i]) cudaFreeHost(b->h_actions[i]);
        if (b->h_actions_free[i]) cudaEventDestroy(b->h_actions_free[i]);
    }
    if (b->h_stats) cudaFreeHost(b->h_stats);
    if (b->h_reward_done) cudaFreeHost(b->h_reward_done);
    if (b->h_results) cudaFreeHost(b->h_results);
    for (cudaEvent_t ev : b->result_ready)
        if (ev) cudaEventDestroy(ev);
    if (b->stream2) {
        cudaStreamSynchronize(b->stream2);
        cudaStreamDestroy(b->stream2);
    }
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    if (b->ev_join) cudaEventDestroy(b->ev_join);
    for (int k = 0; k < 2; ++k) {
        if (b->ev_step[k]) cudaEventDestroy(b->ev_step[k]);
        if (b->ev_top[k]) cudaEventDestroy(b->ev_top[k]);
    }
    if (b->stream) cudaStreamDestroy(b->stream);
    release_dir_slot(b->device, b->dir_slot);
    cudaGetLastError();
    delete b;
    return RCW_OK;
}

static int32_t create_impl(rcw_batch* b, const float* directions_wu) {
    const rcw_config& c = b->cfg;
    const int H = c.height_tile_map_tu, W = c.width_tile_map_tu, N = c.num_directions;
    const int R = c.num_rays, P = c.height_camera_view_pu;
    const int64_t E = c.num_envs;

    cudaDeviceProp prop;
    RCW_CUDA(cudaGetDeviceProperties(&prop, b->device));
    b->sm_count = prop.multiProcessorCount;
    RCW_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    // a second stream for the multi-step calls, whose steps run as two half-batches that overlap each other's
    // launch ramp and tail (enqueue_random_steps_two_streams)
    RCW_CUDA(cudaStreamCreateWithFlags(&b->stream2, cudaStreamNonBlocking));
    RCW_CUDA(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
    RCW_CUDA(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) {
        RCW_CUDA(cudaEventCreateWithFlags(&b->ev_step[k], cudaEventDisableTiming));
        RCW_CUDA(cudaEventCreateWithFlags(&b->ev_top[k], cudaEventDisableTiming));
    }
    if (const char* s = getenv("RCW_TOP_PIPELINE")) b->top_pipeline = atoi(s) != 0;
    if (const char* s = getenv("RCW_TWO_STREAMS")) b->two_streams = atoi(s) != 0;
    if (const char* s = getenv("RCW_TWO_STREAMS_MIN")) b->two_streams_min = atoll(s);
    if (const char* s = getenv("RCW_CTAS_PER_SM")) b->ctas_per_sm = atoi(s);
    else b->ctas_per_sm = -1;   // decided below, once the observation geometry is known

    // ---- direction table (single_room.jl:65-69) -------------------------------------------
    std::vector<float2> dirs((size_t)N);
    if (directions_wu) {
        for (int i = 0; i < N; ++i) dirs[i] = make_float2(directions_wu[2 * i], directions_wu[2 * i + 1]);
    } else {
        const double pi = 3.14159265358979323846;
        for (int i = 1; i <= N; ++i) {
            const double theta = (double)(i - 1) * 2 * pi / (double)N;
            dirs[i - 1] = make_float2((float)cos(theta), (float)sin(theta));
        }
    }
    RCW_CUDA(dev_alloc(b, &b->d_dirs, (size_t)N, false));
    RCW_CUDA(cudaMemcpyAsync(b->d_dirs, dirs.data(), sizeof(float2) * (size_t)N,
                             cudaMemcpyHostToDevice, b->stream));
    if (N <= kDirSlotEntries) {
        b->dir_slot = acquire_dir_slot(b->device);
        if (b->dir_slot >= 0) RCW_CUDA(upload_dir_slot(b->dir_slot, dirs.data(), N, b->stream));
    }
    RCW_CUDA(cudaStreamSynchronize(b->stream));  // dirs is a host temporary

    // ---- ray table ---------------------------------------------------------------------------
    RCW_CUDA(dev_alloc(b, &b->d_ray_table, (size_t)N * (size_t)R, false));
    RCW_CUDA(launch_build_ray_table(b->d_dirs, N, R, c.semi_field_of_view_wu, b->d_ray_table, b->stream));
    b->launches += 1;

    // ---- wall layer ---------------------------------------------------------------------------
    b->wpr = (W + 31) / 32;
    b->map_words = ((H * b->wpr + 3) / 4) * 4;
    b->h_layers.assign((size_t)(1 + b->n_extra) * b->map_words, 0u);   // [wall][extra layers, empty until rcw_set_layer]
    pack_border_walls(H, W, b->wpr, b->h_layers);
    RCW_CUDA(dev_alloc(b, &b->d_wall_map, (size_t)(b->n_extra + 2) * b->map_words, true));
    if (int32_t rc = upload_layers(b)) return rc;

    // ---- single-colour pattern buffers for the bulk renderer: palette entry k repeated as the
    //      observation's byte stream (RGB8: R,G,B,R,...; XRGB32: little-endian 0x00RRGGBB words) ----
    const int col_bytes = column_bytes(b);
    b->pat_stride = ((col_bytes < 3072 ? col_bytes : 3072) + 32 + 15) & ~15;
    // Renderer: lane-written whole sectors (default) or TMA bulk stores of the bands.  The bulk path
    // is kept as a measured alternative (profiles/): its per-lane UBLKCP issue serialises and its
    // 16-byte band edges leave partial sectors, so it is slower than the sector writer.
    b->bulk = false;
    if (const char* s = getenv("RCW_RENDER_PATH")) b->bulk = strcmp(s, "bulk") == 0 && c.obs_format != RCW_OBS_GRAY8 && c.obs_format != RCW_OBS_GRAY16F;
    if (const char* s = getenv("RCW_SPLIT")) b->split = atoi(s) != 0;
    if (c.obs_format == RCW_OBS_COLUMNS || c.obs_format == RCW_OBS_GRAY8_HALF) b->bulk = b->split = false;   // nothing / no full frame is painted
    if (b->n_extra) b->bulk = b->split = false;                        // the measured alternatives know the reference's two objects only
    if (b->split) RCW_CUDA(dev_alloc(b, &b->d_col_info, (size_t)E * (size_t)R));   // 4 B per column, two-launch path only
    {
        std::vector<uint8_t> pat((size_t)6 * b->pat_stride);
        for (int k = 0; k < 6; ++k)
            for (int j = 0; j < b->pat_stride; ++j) {
                const uint32_t col = c.palette[k] & 0x00FFFFFFu;
                pat[(size_t)k * b->pat_stride + j] =
                    b->bpp == 3 ? (uint8_t)(col >> (16 - 8 * (j % 3))) : (uint8_t)(col >> (8 * (j % 4)));
            }
        RCW_CUDA(dev_alloc(b, &b->d_patterns, pat.size(), false));
        RCW_CUDA(cudaMemcpyAsync(b->d_patterns, pat.data(), pat.size(), cudaMemcpyHostToDevice, b->stream));
        RCW_CUDA(cudaStreamSynchronize(b->stream));
    }

    // ---- state SoA -----------------------------------------------------------------------------
    for (int k = 0; k < 2; ++k) {
        RCW_CUDA(dev_alloc(b, &b->st[k].pos_x, (size_t)E));
        RCW_CUDA(dev_alloc(b, &b->st[k].pos_y, (size_t)E));
        RCW_CUDA(dev_alloc(b, &b->st[k].dir_au, (size_t)E));
        RCW_CUDA(dev_alloc(b, &b->st[k].goal, (size_t)E));
        RCW_CUDA(dev_alloc(b, &b->st[k].episode, (size_t)E));
    }
    {   // reward f32[E] and done u8[E] share one allocation so that both reach the host in one copy
        uint8_t* rd = nullptr;
        RCW_CUDA(dev_alloc(b, &rd, (size_t)E * 5));
        b->d_reward = reinterpret_cast<float*>(rd);
        b->d_done = rd + (size_t)E * 4;
    }
    RCW_CUDA(dev_alloc(b, &b->d_ep_return, (size_t)E));
    RCW_CUDA(dev_alloc(b, &b->d_ep_length, (size_t)E));
    RCW_CUDA(dev_alloc(b, &b->d_stats, 1));
    RCW_CUDA(dev_alloc(b, &b->d_actions, (size_t)E));
    for (int i = 0; i < kActionRing; ++i) {
        RCW_CUDA(cudaMallocHost((void**)&b->h_actions[i], (size_t)E));
        RCW_CUDA(cudaEventCreateWithFlags(&b->h_actions_free[i], cudaEventDisableTiming));
    }
    RCW_CUDA(cudaMallocHost((void**)&b->h_stats, sizeof(DeviceStats)));
    RCW_CUDA(cudaMallocHost((void**)&b->h_reward_done, (size_t)E * 5));
    b->result_ring = c.result_ring;
    if (b->result_ring > 0) {
        b->result_slot_bytes = ((size_t)E * 5 + 127) & ~(size_t)127;
        RCW_CUDA(cudaHostAlloc((void**)&b->h_results, b->result_slot_bytes * (size_t)b->result_ring,
                               cudaHostAllocMapped | cudaHostAllocPortable));
        memset(b->h_results, 0, b->result_slot_bytes * (size_t)b->result_ring);
        RCW_CUDA(cudaHostGetDevicePointer((void**)&b->d_results, b->h_results, 0));
        for (int i = 0; i < b->result_ring; ++i)
            RCW_CUDA(cudaEventCreateWithFlags(&b->result_ready[i], cudaEventDisableTiming));
    }

    // ---- observations ---------------------------------------------------------------------------
    // every column starts on a 32-byte sector, every env on a 128-byte line (see rcw_obs_layout)
    b->col_pitch = c.obs_format == RCW_OBS_COLUMNS ? 4 : (obs_rows(c) * b->bpp + 31) & ~31;
    b->frame_stack = c.frame_stack > 1 ? c.frame_stack : 1;
    b->frame_stride = (((size_t)obs_columns(c) * b->col_pitch) + 127) & ~(size_t)127;
    b->obs_env_stride = b->frame_stride * (size_t)b->frame_stack;
    // Grid shape: one CTA per 8 items and the hardware scheduler balances the tail (measured best once
    // the register budget below is chosen per geometry; RCW_CTAS_PER_SM > 0 caps the grid and loops).
    if (b->ctas_per_sm < 0) b->ctas_per_sm = 0;
    // Register budget.  When a warp's item is small (< 20 KB: small or medium frames, one-byte pixels) or
    // the map is large (long DDA walks), act! + DDA bound the step and 4 CTAs per SM hide their latency
    // better (+6..14 %); when the store stream bounds it (default camera), 3 CTAs per SM with more
    // registers are 2 % faster (profiles/README.md).
    b->occ4 = (32 * b->col_pitch < 20000) || ((int64_t)H * W >= 1024);
    if (const char* s = getenv("RCW_OCC")) b->occ4 = atoi(s) == 4;
    // Small items (32 columns of at most 10 KB: narrow or one-byte-per-pixel cameras) are bound by act! and the
    // DDA, not by the stores; a warp that owns the whole env needs no block barrier behind act! (env_kernel).
    // Measured at 65,536 envs (GB/s, item kernel -> env kernel): RGB8 64x64 5091 -> 6053, 84x84 5397 -> 6887,
    // 96x96 6224 -> 6971; GRAY8 84x84 2184 -> 2987, 128x128 3866 -> 5615, 160x120 3732 -> 5302, 256x192
    // 5657 -> 7132, 512x256 7031 -> 7017.  Store-bound items lose: RGB8 128x128 (12 KB) 6897 -> 6756,
    // 160x120 6701 -> 6113, 256x192 7366 -> 6305.
    b->env_per_warp = b->gpe <= 8 && 32 * b->col_pitch <= 10240;
    if (c.obs_format == RCW_OBS_COLUMNS) b->env_per_warp = true;   // nothing is painted: act! + DDA only, any width
    if (c.obs_format == RCW_OBS_GRAY8_HALF) b->env_per_warp = true;   // the box filter is composed by env_kernel only
    if (const char* s = getenv("RCW_ENV_PER_WARP")) b->env_per_warp = atoi(s) != 0;   // 1 forces it for any width
    if (const char* s = getenv("RCW_PACKED_ACTIONS")) b->no_packed_actions = atoi(s) == 0;
    b->env_per_warp_min = kEnvPerWarpMinEnvs;
    if (const char* s = getenv("RCW_ENV_PER_WARP_MIN")) b->env_per_warp_min = atoll(s);   // tests force the kernel on small batches
    if (c.obs_format == RCW_OBS_GRAY8_HALF) {   // whatever the switches say: only env_kernel composes the box filter
        b->env_per_warp = true;
        b->env_per_warp_min = 0;
    }
    b->obs_window = (c.obs_window_envs > 0 && c.obs_window_envs < E) ? c.obs_window_envs : E;
    b->obs_bytes = b->obs_env_stride * (size_t)b->obs_window;
    RCW_CUDA(dev_alloc(b, &b->d_obs, b->obs_bytes, /*zero=*/b->frame_stack > 1));   // older ring positions start black
    if (const char* s = getenv("RCW_ROOM")) {
        b->room_allowed = atoi(s) != 0;
        b->room_forced = atoi(s) == 2;
    }
    // ---- ready-made columns (env_kernel's table renderer) ------------------------------------------
    // update_camera_view! paints one of (P / 2 + 1) x 4 possible columns (rows of ceiling = rows of floor x
    // wall / goal colour per hit dimension, single_room.jl:417-439).  When they are small, all of them together
    // fit the L1 / L2 caches, and painting a column is copying it.  Built here with the renderer's own rules
    // (pixel bytes per format, the pitch padding behind the last row continues the floor).
    {
        const int n_colors = 4 + 2 * b->n_extra;    // wall, goal and extra-layer colours x hit dimension
        const size_t table_bytes = (size_t)(P / 2 + 1) * n_colors * (size_t)b->col_pitch;
        size_t limit = 64 * 1024;
        if (const char* s = getenv("RCW_COL_TABLE_KB")) limit = (size_t)atoll(s) * 1024;   // 0 disables the table
        if (b->env_per_warp && c.obs_format != RCW_OBS_COLUMNS && c.obs_format != RCW_OBS_GRAY8_HALF && table_bytes <= limit) {
            uint32_t pal[6 + 2 * RCW_MAX_EXTRA_LAYERS];
            for (int i = 0; i < 2 + n_colors; ++i) {
                const uint32_t col = (i < 6 ? c.palette[i] : c.layer_palette[(i - 6) >> 1][(i - 6) & 1]) & 0x00FFFFFFu;
                pal[i] = c.obs_format == RCW_OBS_GRAY8 ? luma_of(col) * 0x00010101u
                                                       : (c.obs_format == RCW_OBS_GRAY16F ? luma_half_bits(col) * 0x00010001u : col);
            }
            auto pixel_byte = [&](uint32_t col, int k) -> uint8_t {   // byte k of a pixel of colour 0x00RRGGBB
                if (c.obs_format == RCW_OBS_RGB8) return (uint8_t)(col >> (16 - 8 * k));
                if (c.obs_format == RCW_OBS_XRGB32 || c.obs_format == RCW_OBS_GRAY16F) return (uint8_t)(col >> (8 * k));
                return (uint8_t)col;
            };
            std::vector<uint8_t> tab(table_bytes);
            for (int pad = 0; pad <= P / 2; ++pad)
                for (int k = 0; k < n_colors; ++k) {
                    uint8_t* colp = tab.data() + ((size_t)pad * n_colors + k) * b->col_pitch;
                    for (int ob = 0; ob < b->col_pitch; ++ob) {
                        const int row = ob / b->bpp;
                        const uint32_t col = row < pad ? pal[RCW_COLOR_CEILING]
                                                       : (row < P - pad ? pal[RCW_COLOR_WALL_1 + k] : pal[RCW_COLOR_FLOOR]);
                        colp[ob] = pixel_byte(col, ob - row * b->bpp);
                    }
                }
            RCW_CUDA(dev_alloc(b, &b->d_col_table, table_bytes, false));
            RCW_CUDA(cudaMemcpyAsync(b->d_col_table, tab.data(), table_bytes, cudaMemcpyHostToDevice, b->stream));
            RCW_CUDA(cudaStreamSynchronize(b->stream));
        }
    }
    return RCW_OK;
}

int32_t rcw_create(const rcw_config* cfg, const float* directions_wu, rcw_batch** out) {
    if (!cfg || !out) return fail(RCW_EINVAL, "cfg/out is null");
    *out = nullptr;
    if (cfg->struct_size != sizeof(rcw_config))
        return fail(RCW_ESIZE, "rcw_config.struct_size is %u, this library expects %zu",
                    cfg->struct_size, sizeof(rcw_config));
    for (uint32_t r : cfg->reserved)
        if (r) return fail(RCW_EINVAL, "rcw_config.reserved must be zero");
    const int H = cfg->height_tile_map_tu, W = cfg->width_tile_map_tu;
    if (H < 3 || W < 3 || H > 32767 || W > 32767)
        return fail(RCW_EINVAL, "tile map must be between 3x3 and 32767x32767 (got %dx%d)", H, W);
    if (cfg->num_directions < 1 || cfg->num_directions > (1 << 20))
        return fail(RCW_EINVAL, "num_directions out of range: %d", cfg->num_directions);
    if (cfg->num_rays < 1 || cfg->height_camera_view_pu < 1)
        return fail(RCW_EINVAL, "num_rays and height_camera_view_pu must be positive");
    if (cfg->height_camera_view_pu > 32767)
        return fail(RCW_EINVAL, "height_camera_view_pu must be below 32768");
    if (cfg->obs_format == RCW_OBS_GRAY8_HALF && ((cfg->num_rays | cfg->height_camera_view_pu) & 1))
        return fail(RCW_EINVAL, "RCW_OBS_GRAY8_HALF needs an even num_rays and height_camera_view_pu (2 x 2 blocks)");
    if (cfg->obs_format != RCW_OBS_RGB8 && cfg->obs_format != RCW_OBS_XRGB32 && cfg->obs_format != RCW_OBS_GRAY8 &&
        cfg->obs_format != RCW_OBS_GRAY16F && cfg->obs_format != RCW_OBS_GRAY8_HALF &&
        cfg->obs_format != RCW_OBS_COLUMNS)
        return fail(RCW_EINVAL, "unknown obs_format %d", cfg->obs_format);
    if (cfg->num_envs < 1) return fail(RCW_EINVAL, "num_envs must be positive");
    if (!(cfg->player_radius_wu > 0.0f) || !(cfg->player_radius_wu < 0.5f))
        return fail(RCW_EINVAL, "player_radius_wu must be in (0, 0.5) (single_room.jl:47)");
    if (!(cfg->semi_field_of_view_wu > 0.0f)) return fail(RCW_EINVAL, "semi_field_of_view_wu must be positive");
    if (cfg->dda_flags & ~(uint32_t)(RCW_DDA_TIE_LE | RCW_DDA_DIST_POST))
        return fail(RCW_EINVAL, "unknown dda_flags 0x%x", cfg->dda_flags);
    if (cfg->obs_window_envs < 0) return fail(RCW_EINVAL, "obs_window_envs must be >= 0");
    if (cfg->frame_stack < 0 || cfg->frame_stack > 64) return fail(RCW_EINVAL, "frame_stack must be in 0..64");
    if (cfg->result_ring < 0 || cfg->result_ring > 64) return fail(RCW_EINVAL, "result_ring must be in 0..64");
    if (cfg->frame_stack > 1 && cfg->obs_window_envs > 0 && cfg->obs_window_envs < cfg->num_envs)
        return fail(RCW_EINVAL, "frame_stack cannot be combined with an observation window");
    if (cfg->pu_per_tu < 1 || cfg->pu_per_tu > 1024) return fail(RCW_EINVAL, "pu_per_tu must be in 1..1024");
    if (cfg->top_view != 0 && cfg->top_view != 1) return fail(RCW_EINVAL, "top_view must be 0 or 1");
    const int bpp = bytes_per_pixel(cfg->obs_format);   // RCW_OBS_COLUMNS: the 4 bytes of a column word
    const int gpe = (cfg->num_rays + 31) / 32;
    if ((int64_t)cfg->num_rays * cfg->height_camera_view_pu * bpp >= (1LL << 30))
        return fail(RCW_ESIZE, "one observation must be smaller than 1 GiB");
    if ((int64_t)cfg->num_directions * cfg->num_rays >= (1LL << 31))
        return fail(RCW_ESIZE, "num_directions * num_rays must be below 2^31 (the ray table is indexed with 32 bits)");
    if (cfg->num_envs * gpe >= (1LL << 31))
        return fail(RCW_ESIZE, "num_envs * ceil(num_rays/32) must be below 2^31 per handle");
    if (cfg->num_object_layers != 0 && (cfg->num_object_layers < 2 || cfg->num_object_layers > 2 + RCW_MAX_EXTRA_LAYERS))
        return fail(RCW_EINVAL, "num_object_layers must be 2..%d (got %d)", 2 + RCW_MAX_EXTRA_LAYERS, cfg->num_object_layers);
    const int n_extra = cfg->num_object_layers > 2 ? cfg->num_object_layers - 2 : 0;
    for (int k = 0; k < n_extra; ++k)
        if (cfg->layer_kind[k] != RCW_LAYER_BLOCKING && cfg->layer_kind[k] != RCW_LAYER_TERMINAL)
            return fail(RCW_EINVAL, "layer_kind[%d] must be RCW_LAYER_BLOCKING or RCW_LAYER_TERMINAL", k);
    if ((int64_t)((H * ((W + 31) / 32) + 3) / 4) * 16 * (n_extra ? n_extra + 2 : 1) > 200 * 1024)
        return fail(RCW_ESIZE, "bit-packed tile map (%d layers) does not fit in shared memory", n_extra ? n_extra + 2 : 1);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(RCW_ECUDA, "no CUDA device available; librcw_b200 has no CPU fallback");
    }
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(RCW_EINVAL, "device %d out of range (%d devices)", cfg->device, ndev);

    if (cfg->top_view)
        if (int32_t rc = check_top_view_fits(*cfg)) return rc;

    rcw_batch* b = new (std::nothrow) rcw_batch();
    if (!b) return fail(RCW_ENOMEM, "out of host memory");
    b->cfg = *cfg;
    b->device = cfg->device;
    b->bpp = bpp;
    b->gpe = gpe;
    b->n_extra = n_extra;
    DeviceGuard g(b->device);
    if (!g.ok) {
        delete b;
        return fail(RCW_ECUDA, "cudaSetDevice(%d) failed", cfg->device);
    }
    int32_t rc = create_impl(b, directions_wu);
    if (rc == RCW_OK) rc = rcw_reset(b, nullptr, nullptr, nullptr, nullptr);
    if (rc == RCW_OK) rc = sync_and_check(b);
    if (rc != RCW_OK) {
        std::string keep = t_last_error;
        rcw_destroy(b);
        t_last_error = keep;
        return rc;
    }
    *out = b;
    return RCW_OK;
}

int32_t rcw_set_wall_map(rcw_batch* b, const uint8_t* wall) {
    if (int32_t rc = check_handle(b)) return rc;
    if (!wall) return fail(RCW_EINVAL, "wall is null");
    DeviceGuard g(b->device);
    std::fill(b->h_layers.begin(), b->h_layers.begin() + b->map_words, 0u);
    pack_tiles(wall, b->cfg.height_tile_map_tu, b->cfg.width_tile_map_tu, b->wpr, b->h_layers.data());
    return upload_layers(b);
}

int32_t rcw_set_layer(rcw_batch* b, int32_t layer, const uint8_t* tiles) {
    if (int32_t rc = check_handle(b)) return rc;
    if (!tiles) return fail(RCW_EINVAL, "tiles is null");
    if (layer == 1) return rcw_set_wall_map(b, tiles);
    if (layer == 2) return fail(RCW_EINVAL, "layer 2 (GOAL) holds one tile per env: move it with rcw_reset / rcw_set_state");
    if (layer < 3 || layer > 2 + b->n_extra)
        return fail(RCW_EINVAL, "no object layer %d (rcw_config.num_object_layers = %d)", layer, 2 + b->n_extra);
    DeviceGuard g(b->device);
    uint32_t* words = b->h_layers.data() + (size_t)(layer - 2) * b->map_words;
    std::fill(words, words + b->map_words, 0u);
    pack_tiles(tiles, b->cfg.height_tile_map_tu, b->cfg.width_tile_map_tu, b->wpr, words);
    return upload_layers(b);
}

int32_t rcw_set_wall_maps(rcw_batch* b, const uint8_t* walls) {
    if (int32_t rc = check_handle(b)) return rc;
    if (!walls) return fail(RCW_EINVAL, "walls is null");
    DeviceGuard g(b->device);
    const int H = b->cfg.height_tile_map_tu, W = b->cfg.width_tile_map_tu;
    const size_t E = (size_t)b->cfg.num_envs, mw = (size_t)b->map_words, tiles = (size_t)H * W;
    if ((size_t)kWarpsPerCta * mw * 4 > 200 * 1024)
        return fail(RCW_ESIZE, "per-env tile maps of %dx%d do not fit in shared memory", H, W);
    if (b->n_extra) return fail(RCW_EINVAL, "per-env wall layers cannot be combined with extra object layers");
    bool closed = true;
    std::vector<uint32_t> words;
    try {
        words.assign(E * mw, 0u);
    } catch (const std::bad_alloc&) {
        return fail(RCW_ENOMEM, "out of host memory packing %zu wall layers", E);
    }
    for (size_t e = 0; e < E; ++e) {
        const uint8_t* w = walls + e * tiles;
        uint32_t* out = words.data() + e * mw;
        for (int j = 0; j < W; ++j)
            for (int i = 0; i < H; ++i) {
                const bool is_wall = w[(size_t)j * H + i] != 0;
                if (is_wall) out[(size_t)i * b->wpr + (j >> 5)] |= 1u << (j & 31);
                else if (i == 0 || i == H - 1 || j == 0 || j == W - 1) closed = false;
            }
    }
    RCW_CUDA(cudaStreamSynchronize(b->stream));
    if (!b->d_wall_maps_env) RCW_CUDA(dev_alloc(b, &b->d_wall_maps_env, E * mw, false));
    RCW_CUDA(cudaMemcpy(b->d_wall_maps_env, words.data(), sizeof(uint32_t) * words.size(),
                        cudaMemcpyHostToDevice));
    b->per_env_maps = true;
    b->closed_border = closed;
    b->room = false;
    return RCW_OK;
}

int32_t rcw_render_top_view(rcw_batch* b) {
    NvtxRange nvtx("rcw_render_top_view");
    if (int32_t rc = check_handle(b)) return rc;
    if (int32_t rc = check_top_view_fits(b->cfg)) return rc;
    DeviceGuard g(b->device);
    const int64_t E = b->cfg.num_envs;
    for (int64_t e0 = 0; e0 < E; e0 += b->obs_window)
        if (int32_t rc = enqueue_top_view(b, b->st[b->cur], e0, E - e0 < b->obs_window ? E - e0 : b->obs_window, 0))
            return rc;
    return RCW_OK;
}

int32_t rcw_top_view_device_ptr(rcw_batch* b, void** dptr, size_t* total_bytes, size_t* env_stride_bytes) {
    if (int32_t rc = check_handle(b)) return rc;
    if (dptr) *dptr = b->d_top;
    if (total_bytes) *total_bytes = b->d_top ? b->top_env_stride * (size_t)b->obs_window : 0;
    if (env_stride_bytes) *env_stride_bytes = b->top_env_stride;
    return RCW_OK;
}

int32_t rcw_copy_top_view(rcw_batch* b, int64_t env0, int64_t n, void* host) {
    if (int32_t rc = check_handle(b)) return rc;
    if (!host) return fail(RCW_EINVAL, "host is null");
    if (!b->d_top) return fail(RCW_EINVAL, "no top view has been drawn yet (rcw_render_top_view or rcw_config.top_view)");
    if (env0 < 0 || n < 1 || env0 + n > b->cfg.num_envs)
        return fail(RCW_ESIZE, "env range [%lld, %lld) outside 0..%lld", (long long)env0,
                    (long long)(env0 + n), (long long)b->cfg.num_envs);
    if (n > b->obs_window)
        return fail(RCW_ESIZE, "%lld envs requested, the window holds %lld", (long long)n, (long long)b->obs_window);
    DeviceGuard g(b->device);
    const rcw_config& c = b->cfg;
    const size_t bytes = (size_t)c.height_tile_map_tu * c.pu_per_tu * (size_t)c.width_tile_map_tu * c.pu_per_tu * 4;
    const int64_t slot0 = env0 % b->obs_window;
    if (slot0 + n > b->obs_window) {
        const int64_t n1 = b->obs_window - slot0;
        if (int32_t rc = rcw_copy_top_view(b, env0, n1, host)) return rc;
        return rcw_copy_top_view(b, env0 + n1, n - n1, static_cast<uint8_t*>(host) + (size_t)n1 * bytes);
    }
    RCW_CUDA(cudaMemcpy2DAsync(host, bytes, b->d_top + (size_t)slot0 * b->top_env_stride, b->top_env_stride, bytes,
                               (size_t)n, cudaMemcpyDeviceToHost, b->stream));
    return sync_and_check(b);
}

int32_t rcw_render(rcw_batch* b) {
    NvtxRange nvtx("rcw_render");
    if (int32_t rc = check_handle(b)) return rc;
    DeviceGuard g(b->device);
    return enqueue_frame(b, kModeRender, nullptr);
}

static bool is_pinned_host(const void* ptr) {
    if (!ptr) return false;
    cudaPointerAttributes attr;
    const cudaError_t pe = cudaPointerGetAttributes(&attr, ptr);
    if (pe != cudaSuccess) cudaGetLastError();
    return pe == cudaSuccess && attr.type == cudaMemoryTypeHost;
}

int32_t rcw_reset(rcw_batch* b, const int32_t* goal_ij, const int32_t* player_ij,
                  const int32_t* dir_au, const uint8_t* mask) {
    NvtxRange nvtx("rcw_reset");
    if (int32_t rc = check_handle(b)) return rc;
    const bool any = goal_ij || player_ij || dir_au;
    if (any && !(goal_ij && player_ij && dir_au))
        return fail(RCW_EINVAL, "goal_ij, player_ij and dir_au must be all given or all null");
    DeviceGuard g(b->device);
    const rcw_config& c = b->cfg;
    const int64_t E = c.num_envs;
    int32_t *d_goal = nullptr, *d_player = nullptr, *d_dir = nullptr;
    uint8_t* d_mask = nullptr;
    if (any) {
        for (int64_t e = 0; e < E; ++e) {
            if (mask && !mask[e]) continue;
            const int gi = goal_ij[2 * e], gj = goal_ij[2 * e + 1];
            const int pi = player_ij[2 * e], pj = player_ij[2 * e + 1];
            if (gi < 1 || gi > c.height_tile_map_tu || gj < 1 || gj > c.width_tile_map_tu ||
                pi < 1 || pi > c.height_tile_map_tu || pj < 1 || pj > c.width_tile_map_tu)
                return fail(RCW_EINVAL, "env %lld: tile outside the map", (long long)e);
            if (dir_au[e] < 0 || dir_au[e] >= c.num_directions)
                return fail(RCW_EINVAL, "env %lld: direction %d outside 0..%d", (long long)e,
                            dir_au[e], c.num_directions - 1);
        }
        // the scratch buffers live as long as the handle; copies from pageable host memory are staged before
        // cudaMemcpyAsync returns, copies and kernels are ordered by the handle's stream: no blocking here
        // (pinned / registered arrays: see below)
        if (!b->d_reset_goal) {
            RCW_CUDA(dev_alloc(b, &b->d_reset_goal, 2 * (size_t)E, false));
            RCW_CUDA(dev_alloc(b, &b->d_reset_player, 2 * (size_t)E, false));
            RCW_CUDA(dev_alloc(b, &b->d_reset_dir, (size_t)E, false));
        }
        d_goal = b->d_reset_goal;
        d_player = b->d_reset_player;
        d_dir = b->d_reset_dir;
        RCW_CUDA(cudaMemcpyAsync(d_goal, goal_ij, sizeof(int32_t) * 2 * (size_t)E, cudaMemcpyHostToDevice, b->stream));
        RCW_CUDA(cudaMemcpyAsync(d_player, player_ij, sizeof(int32_t) * 2 * (size_t)E, cudaMemcpyHostToDevice, b->stream));
        RCW_CUDA(cudaMemcpyAsync(d_dir, dir_au, sizeof(int32_t) * (size_t)E, cudaMemcpyHostToDevice, b->stream));
    }
    if (mask) {
        if (!b->d_reset_mask) RCW_CUDA(dev_alloc(b, &b->d_reset_mask, (size_t)E, false));
        d_mask = b->d_reset_mask;
        RCW_CUDA(cudaMemcpyAsync(d_mask, mask, (size_t)E, cudaMemcpyHostToDevice, b->stream));
    }
    // "staged before cudaMemcpyAsync returns" only holds for pageable memory: the DMA engine reads pinned or
    // registered host arrays (torch pinned tensors, cudaHostRegister) when the copy executes.  The header promises
    // that host pointers are only touched during the call, so wait for those copies here.
    if (is_pinned_host(goal_ij) || is_pinned_host(player_ij) || is_pinned_host(dir_au) || is_pinned_host(mask))
        RCW_CUDA(cudaStreamSynchronize(b->stream));
    ResetParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.H = c.height_tile_map_tu;
    rp.W = c.width_tile_map_tu;
    rp.wpr = b->wpr;
    rp.N = c.num_directions;
    rp.wall_map = b->per_env_maps ? b->d_wall_maps_env : b->d_wall_map;
    rp.map_env_stride = b->per_env_maps ? (uint32_t)b->map_words : 0u;
    rp.n_extra = b->per_env_maps ? 0 : b->n_extra;
    rp.map_words = b->map_words;
    rp.st = b->st[b->cur];
    rp.reward = b->d_reward;
    rp.done = b->d_done;
    rp.ep_return = b->d_ep_return;
    rp.ep_length = b->d_ep_length;
    rp.num_envs = E;
    rp.env_id_offset = (uint64_t)c.env_id_offset;
    rp.seed = c.seed;
    rp.goal_ij = d_goal;
    rp.player_ij = d_player;
    rp.dir_au = d_dir;
    rp.mask = d_mask;
    RCW_CUDA(launch_reset(rp, b->stream));
    b->launches += 1;
    // only the envs that were reset are redrawn (the others' observations are still those of their state)
    return enqueue_frame(b, kModeRender, nullptr, nullptr, d_mask);
}

// The actions of envs [env0, env0 + n) as a device pointer to the first of them.  A device array is used
// in place (its values are checked by the kernel); a host array is validated (the reference's @assert,
// single_room.jl:140) while it is staged into pinned memory, then copied to the handle's action buffer.
static int32_t stage_actions(rcw_batch* b, const uint8_t* actions, int64_t env0, int64_t n, const uint8_t** d_out) {
    cudaPointerAttributes attr;
    const cudaError_t pe = cudaPointerGetAttributes(&attr, actions);
    if (pe != cudaSuccess) cudaGetLastError();
    if (pe == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged)) {
        b->device_actions_pending = true;
        *d_out = actions;
        return RCW_OK;
    }
    const int slot = b->ring;
    RCW_CUDA(cudaEventSynchronize(b->h_actions_free[slot]));
    uint8_t* stage = b->h_actions[slot] + env0;
    uint32_t bad = 0;
    for (int64_t k = 0; k < n; ++k) {
        const uint8_t a = actions[k];
        bad |= (uint32_t)(a - 1u > 3u);
        stage[k] = a;
    }
    if (bad) {
        for (int64_t k = 0; k < n; ++k)
            if (actions[k] < 1 || actions[k] > 4)
                return fail(RCW_EACTION, "Invalid action: %d (env %lld); actions must be in 1..4",
                            (int)actions[k], (long long)(env0 + k));
    }
    RCW_CUDA(cudaMemcpyAsync(b->d_actions + env0, stage, (size_t)n, cudaMemcpyHostToDevice, b->stream));
    RCW_CUDA(cudaEventRecord(b->h_actions_free[slot], b->stream));
    b->ring = (slot + 1) % kActionRing;
    *d_out = b->d_actions + env0;
    return RCW_OK;
}

static bool is_device_pointer(const void* ptr) {
    cudaPointerAttributes attr;
    const cudaError_t pe = cudaPointerGetAttributes(&attr, ptr);
    if (pe != cudaSuccess) cudaGetLastError();
    return pe == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
}

// the reference's @assert (single_room.jl:140) for a host array
static int32_t validate_host_actions(const uint8_t* actions, int64_t env0, int64_t n) {
    uint32_t bad = 0;
    for (int64_t k = 0; k < n; ++k) bad |= (uint32_t)((uint8_t)(actions[k] - 1u) > 3u);
    if (bad)
        for (int64_t k = 0; k < n; ++k)
            if (actions[k] < 1 || actions[k] > 4)
                return fail(RCW_EACTION, "Invalid action: %d (env %lld); actions must be in 1..4", (int)actions[k],
                            (long long)(env0 + k));
    return RCW_OK;
}

int32_t rcw_step(rcw_batch* b, const uint8_t* actions) {
    NvtxRange nvtx("rcw_step");
    if (int32_t rc = check_handle(b)) return rc;
    if (!actions) return fail(RCW_EINVAL, "actions is null (use rcw_step_random for the random policy)");
    DeviceGuard g(b->device);
    const int64_t per_launch = b->cfg.num_envs < b->obs_window ? b->cfg.num_envs : b->obs_window;
    if (packs_actions(b, per_launch) && !is_device_pointer(actions)) {
        // host actions ride in the kernel parameters: no staging copy in front of the launch
        if (int32_t rc = validate_host_actions(actions, 0, b->cfg.num_envs)) return rc;
        return enqueue_frame(b, kModeStep, nullptr, actions);
    }
    const uint8_t* d_actions = nullptr;
    if (int32_t rc = stage_actions(b, actions, 0, b->cfg.num_envs, &d_actions)) return rc;
    return enqueue_frame(b, kModeStep, d_actions);
}

int32_t rcw_step_async(rcw_batch* b, const uint8_t* actions, int64_t* ticket) {
    NvtxRange nvtx("rcw_step_async");
    if (int32_t rc = check_handle(b)) return rc;
    if (!ticket) return fail(RCW_EINVAL, "ticket is null");
    if (b->result_ring < 1)
        return fail(RCW_EINVAL, "rcw_step_async needs a result ring (rcw_config.result_ring >= 1)");
    if (b->split || b->bulk)
        return fail(RCW_EINVAL, "rcw_step_async is not available on the RCW_SPLIT / RCW_RENDER_PATH=bulk variants");
    const int slot = (int)(b->next_ticket % b->result_ring);
    b->result_slot = slot;
    const int32_t rc = rcw_step(b, actions);
    b->result_slot = -1;
    if (rc) return rc;   // nothing was enqueued (null / invalid host actions): no ticket is consumed
    DeviceGuard g(b->device);
    RCW_CUDA(cudaEventRecord(b->result_ready[slot], b->stream));
    *ticket = b->next_ticket++;
    return RCW_OK;
}

int32_t rcw_wait(rcw_batch* b, int64_t ticket, const float** reward, const uint8_t** done) {
    if (int32_t rc = check_handle(b)) return rc;
    if (b->result_ring < 1) return fail(RCW_EINVAL, "no result ring (rcw_config.result_ring >= 1)");
    if (ticket < 0 || ticket >= b->next_ticket)
        return fail(RCW_EINVAL, "ticket %lld was not issued (next is %lld)", (long long)ticket, (long long)b->next_ticket);
    if (ticket + b->result_ring < b->next_ticket)
        return fail(RCW_EINVAL, "ticket %lld is too old: its slot of the %d-deep result ring was reused by step %lld",
                    (long long)ticket, b->result_ring, (long long)(ticket + b->result_ring));
    const int slot = (int)(ticket % b->result_ring);
    RCW_CUDA(cudaEventSynchronize(b->result_ready[slot]));
    const uint8_t* base = b->h_results + (size_t)slot * b->result_slot_bytes;
    if (reward) *reward = reinterpret_cast<const float*>(base);
    if (done) *done = base + (size_t)b->cfg.num_envs * 4;
    return RCW_OK;
}

int32_t rcw_step_range(rcw_batch* b, const uint8_t* actions, int64_t env0, int64_t n) {
    NvtxRange nvtx("rcw_step_range");
    if (int32_t rc = check_handle(b)) return rc;
    if (!actions) return fail(RCW_EINVAL, "actions is null");
    if (env0 < 0 || n < 1 || env0 + n > b->cfg.num_envs)
        return fail(RCW_ESIZE, "env range [%lld, %lld) outside 0..%lld", (long long)env0,
                    (long long)(env0 + n), (long long)b->cfg.num_envs);
    if (n > b->obs_window)
        return fail(RCW_ESIZE, "a range of %lld envs does not fit the observation window of %lld",
                    (long long)n, (long long)b->obs_window);
    if (b->frame_stack > 1)
        return fail(RCW_EINVAL, "rcw_step_range cannot be used with a frame ring (frame_stack = %d): the ring "
                    "position is shared by the batch", b->frame_stack);
    DeviceGuard g(b->device);
    if (packs_actions(b, n) && !is_device_pointer(actions)) {
        if (int32_t rc = validate_host_actions(actions, env0, n)) return rc;
        return enqueue_range_step(b, nullptr, env0, n, actions);
    }
    const uint8_t* d_actions = nullptr;
    if (int32_t rc = stage_actions(b, actions, env0, n, &d_actions)) return rc;
    return enqueue_range_step(b, d_actions, env0, n);
}

// n steps of the random policy as two half-batches, one per stream.  Envs are independent, the state is
// struct-of-arrays and every launch touches only its own envs' entries, so the two halves never meet: each stream
// simply runs its half's steps back to back, and the end of one half's launch (a few half-empty waves and ~3 us
// of launch gap) is filled by the other half's kernel.  Measured (tools/two_stream_probe.py, default camera): 0.0643
// -> 0.0541 ms per step at 1024 envs (+19 %), 0.2270 -> 0.2178 at 4096 (+4.2 %), +1 % at 16,384.  Only inside a
// multi-step call: between single rcw_step calls the caller may enqueue consumers of the observations on the
// handle's stream, and those must stay ordered with the next step.  fork: the side stream waits for everything
// enqueued on the handle's stream so far; join: the handle's stream waits for the side stream's last step.
// d_tape: nullptr = the random policy; otherwise device actions [n_steps][num_envs], step s reads row s
static int32_t enqueue_steps_two_streams(rcw_batch* b, int32_t n_steps, const uint8_t* d_tape = nullptr) {
    const int64_t E = b->cfg.num_envs;
    const int64_t half = (E / 2) & ~(int64_t)(kWarpsPerCta - 1);
    RCW_CUDA(cudaEventRecord(b->ev_fork, b->stream));
    RCW_CUDA(cudaStreamWaitEvent(b->stream2, b->ev_fork, 0));
    // With the top view redrawn in every step (rcw_config.top_view) the two streams are used as a two-stage pipeline
    // instead: the handle's stream runs the step kernels of the whole batch back to back, the side stream the top view
    // kernels, top view k behind step k (an event) and therefore under step k + 1 — the top view's ray walk is
    // issue-bound and stores little while it draws, the step kernel is a pure store stream, so each fills the other's
    // gaps.  State is double-buffered: step k + 2 overwrites what top view k reads, so it waits for it (another event).
    // Measured at 4096 envs: 0.577 ms per step (0.997 of the copy peak for 917,504 B per env-step) against 0.589 for two
    // half-batches and 0.631 for one stream; equal to the half-batches at 16,384 envs (1.02).  RCW_TOP_PIPELINE=0 = halves.
    const bool pipeline = b->cfg.top_view != 0 && b->top_pipeline;
    auto steps = [&]() -> int32_t {
        for (int32_t s = 0; s < n_steps; ++s) {
            b->frame_newest = (b->frame_newest + 1) % b->frame_stack;
            FrameParams p;
            fill_frame_params(b, p);
            p.actions = d_tape ? d_tape + (size_t)s * (size_t)E : nullptr;   // (the kernels index the actions by env)
            if (pipeline) {
                if (s >= 2) RCW_CUDA(cudaStreamWaitEvent(b->stream, b->ev_top[s & 1], 0));   // top view s - 2 has read `out`
                RCW_CUDA(launch_frame(p, kModeStep, b->cfg.obs_format, shape_for(b, E), b->stream));
                RCW_CUDA(cudaEventRecord(b->ev_step[s & 1], b->stream));
                RCW_CUDA(cudaStreamWaitEvent(b->stream2, b->ev_step[s & 1], 0));
                b->launches += 1;
                if (int32_t rc = enqueue_top_view(b, p.out, 0, E, 0, nullptr, b->stream2)) return rc;
                RCW_CUDA(cudaEventRecord(b->ev_top[s & 1], b->stream2));
                b->cur ^= 1;
                b->step_index += 1;
                continue;
            }
            for (int part = 0; part < 2; ++part) {
                cudaStream_t stream = part ? b->stream : b->stream2;
                p.env_first = part ? half : 0;
                p.env_count = part ? E - half : half;
                p.obs_slot0 = (uint32_t)p.env_first;
                RCW_CUDA(launch_frame(p, kModeStep, b->cfg.obs_format, shape_for(b, p.env_count), stream));
                b->launches += 1;
                if (b->cfg.top_view)
                    if (int32_t rc = enqueue_top_view(b, p.out, p.env_first, p.env_count, p.obs_slot0, nullptr, stream)) return rc;
            }
            b->cur ^= 1;
            b->step_index += 1;
        }
        return RCW_OK;
    };
    const int32_t rc = steps();
    // join whatever happened: the handle's stream must cover everything that was enqueued on the side stream
    cudaError_t e = cudaEventRecord(b->ev_join, b->stream2);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(b->stream, b->ev_join, 0);
    if (rc != RCW_OK) return rc;
    RCW_CUDA(e);
    return RCW_OK;
}

// multi-step calls of at least two steps on a batch that is rendered in one launch may run as two half-batches
static bool two_streams_apply(const rcw_batch* b, int32_t n_steps) {
    return n_steps >= 2 && b->two_streams && !b->split && !b->bulk && b->obs_window == b->cfg.num_envs &&
           b->cfg.num_envs >= b->two_streams_min && b->cfg.num_envs >= 2 * kWarpsPerCta;
}

int32_t rcw_step_tape(rcw_batch* b, const uint8_t* actions, int32_t n_steps) {
    NvtxRange nvtx("rcw_step_tape");
    if (int32_t rc = check_handle(b)) return rc;
    if (!actions) return fail(RCW_EINVAL, "actions is null (use rcw_step_random for the random policy)");
    if (n_steps < 0) return fail(RCW_EINVAL, "n_steps must be non-negative");
    if (n_steps == 0) return RCW_OK;
    DeviceGuard g(b->device);
    const size_t E = (size_t)b->cfg.num_envs, total = E * (size_t)n_steps;
    const uint8_t* d_tape = actions;
    if (is_device_pointer(actions)) {
        b->device_actions_pending = true;       // an invalid value is reported by the next blocking call
    } else {
        // the reference's @assert for every step of the tape, before anything is enqueued
        if (int32_t rc = validate_host_actions(actions, 0, (int64_t)total)) return rc;
        if (b->tape_capacity < total) {
            uint8_t* fresh = nullptr;
            RCW_CUDA(dev_alloc(b, &fresh, total, false));      // (an outgrown tape buffer stays allocated until rcw_destroy)
            b->d_tape = fresh;
            b->tape_capacity = total;
        }
        RCW_CUDA(cudaMemcpyAsync(b->d_tape, actions, total, cudaMemcpyHostToDevice, b->stream));
        if (is_pinned_host(actions)) RCW_CUDA(cudaStreamSynchronize(b->stream));   // host pointers are only touched during the call
        d_tape = b->d_tape;
    }
    if (two_streams_apply(b, n_steps)) {
        int32_t first = 0;
        if (b->cfg.top_view && !b->d_top) {    // (the top views are allocated by their first draw)
            if (int32_t rc = enqueue_frame(b, kModeStep, d_tape)) return rc;
            first = 1;
        }
        return enqueue_steps_two_streams(b, n_steps - first, d_tape + (size_t)first * E);
    }
    for (int32_t s = 0; s < n_steps; ++s)
        if (int32_t rc = enqueue_frame(b, kModeStep, d_tape + (size_t)s * E)) return rc;
    return RCW_OK;
}

int32_t rcw_step_random(rcw_batch* b, int32_t n_steps) {
    NvtxRange nvtx("rcw_step_random");
    if (int32_t rc = check_handle(b)) return rc;
    if (n_steps < 0) return fail(RCW_EINVAL, "n_steps must be non-negative");
    DeviceGuard g(b->device);
    if (two_streams_apply(b, n_steps)) {
        if (b->cfg.top_view && !b->d_top) {    // (the top views are allocated by their first draw)
            if (int32_t rc = enqueue_frame(b, kModeStep, nullptr)) return rc;
            --n_steps;
        }
        return enqueue_steps_two_streams(b, n_steps);
    }
    for (int32_t s = 0; s < n_steps; ++s)
        if (int32_t rc = enqueue_frame(b, kModeStep, nullptr)) return rc;
    return RCW_OK;
}

int32_t rcw_get_state(rcw_batch* b, float* pos_xy, int32_t* dir_au, int32_t* goal_ij, float* reward,
                      uint8_t* done) {
    if (int32_t rc = check_handle(b)) return rc;
    DeviceGuard g(b->device);
    const size_t E = (size_t)b->cfg.num_envs;
    const StateRef& s = b->st[b->cur];
    std::vector<float> px, py;
    std::vector<uint32_t> goal;
    if (pos_xy) {
        px.resize(E);
        py.resize(E);
        RCW_CUDA(cudaMemcpyAsync(px.data(), s.pos_x, sizeof(float) * E, cudaMemcpyDeviceToHost, b->stream));
        RCW_CUDA(cudaMemcpyAsync(py.data(), s.pos_y, sizeof(float) * E, cudaMemcpyDeviceToHost, b->stream));
    }
    if (dir_au) RCW_CUDA(cudaMemcpyAsync(dir_au, s.dir_au, sizeof(int32_t) * E, cudaMemcpyDeviceToHost, b->stream));
    if (goal_ij) {
        goal.resize(E);
        RCW_CUDA(cudaMemcpyAsync(goal.data(), s.goal, sizeof(uint32_t) * E, cudaMemcpyDeviceToHost, b->stream));
    }
    if (reward || done)
        RCW_CUDA(cudaMemcpyAsync(b->h_reward_done, b->d_reward, E * 5, cudaMemcpyDeviceToHost, b->stream));
    if (int32_t rc = sync_and_check(b)) return rc;
    if (reward) memcpy(reward, b->h_reward_done, sizeof(float) * E);
    if (done) memcpy(done, b->h_reward_done + E * 4, E);
    if (pos_xy)
        for (size_t e = 0; e < E; ++e) {
            pos_xy[2 * e] = px[e];
            pos_xy[2 * e + 1] = py[e];
        }
    if (goal_ij)
        for (size_t e = 0; e < E; ++e) {
            goal_ij[2 * e] = (int32_t)(goal[e] & 0xFFFFu);
            goal_ij[2 * e + 1] = (int32_t)(goal[e] >> 16);
        }
    return RCW_OK;
}

int32_t rcw_set_state(rcw_batch* b, const float* pos_xy, const int32_t* dir_au,
                      const int32_t* goal_ij, const float* reward, const uint8_t* done) {
    if (int32_t rc = check_handle(b)) return rc;
    DeviceGuard g(b->device);
    const rcw_config& c = b->cfg;
    const size_t E = (size_t)c.num_envs;
    const StateRef& s = b->st[b->cur];
    std::vector<float> px, py;
    std::vector<uint32_t> goal;
    if (dir_au)
        for (size_t e = 0; e < E; ++e)
            if (dir_au[e] < 0 || dir_au[e] >= c.num_directions)
                return fail(RCW_EINVAL, "env %zu: direction %d outside 0..%d", e, dir_au[e], c.num_directions - 1);
    if (goal_ij)
        for (size_t e = 0; e < E; ++e)
            if (goal_ij[2 * e] < 1 || goal_ij[2 * e] > c.height_tile_map_tu || goal_ij[2 * e + 1] < 1 ||
                goal_ij[2 * e + 1] > c.width_tile_map_tu)
                return fail(RCW_EINVAL, "env %zu: goal tile outside the map", e);
    if (pos_xy) {
        px.resize(E);
        py.resize(E);
        for (size_t e = 0; e < E; ++e) {
            const float x = pos_xy[2 * e], y = pos_xy[2 * e + 1];
            if (!(x >= 0.0f) || !(x < (float)c.height_tile_map_tu) || !(y >= 0.0f) || !(y < (float)c.width_tile_map_tu))
                return fail(RCW_EINVAL, "env %zu: position (%g, %g) outside the map", e, (double)x, (double)y);
            px[e] = x;
            py[e] = y;
        }
        RCW_CUDA(cudaMemcpyAsync(s.pos_x, px.data(), sizeof(float) * E, cudaMemcpyHostToDevice, b->stream));
        RCW_CUDA(cudaMemcpyAsync(s.pos_y, py.data(), sizeof(float) * E, cudaMemcpyHostToDevice, b->stream));
    }
    if (dir_au) RCW_CUDA(cudaMemcpyAsync(s.dir_au, dir_au, sizeof(int32_t) * E, cudaMemcpyHostToDevice, b->stream));
    if (goal_ij) {
        goal.resize(E);
        for (size_t e = 0; e < E; ++e)
            goal[e] = (uint32_t)goal_ij[2 * e] | ((uint32_t)goal_ij[2 * e + 1] << 16);
        RCW_CUDA(cudaMemcpyAsync(s.goal, goal.data(), sizeof(uint32_t) * E, cudaMemcpyHostToDevice, b->stream));
    }
    if (reward) RCW_CUDA(cudaMemcpyAsync(b->d_reward, reward, sizeof(float) * E, cudaMemcpyHostToDevice, b->stream));
    if (done) RCW_CUDA(cudaMemcpyAsync(b->d_done, done, E, cudaMemcpyHostToDevice, b->stream));
    RCW_CUDA(cudaStreamSynchronize(b->stream));  // px/py/goal are host temporaries
    return RCW_OK;
}

// ---- checkpoint / resume ---------------------------------------------------------------------
struct CheckpointHeader {
    uint32_t magic;      // 'RCWC'
    uint32_t version;
    int64_t num_envs;
    int64_t env_id_offset;
    uint64_t seed;
    uint64_t step_index;
    int32_t H, W, N, pad;
    unsigned long long episodes, sum_length;
    double sum_return;
};
constexpr uint32_t kCheckpointMagic = 0x43574352u, kCheckpointVersion = 1u;
// per env: pos_x, pos_y, dir_au, goal, episode, reward, ep_return, ep_length (4 bytes each) + done (1 byte)
constexpr size_t kCheckpointWordsPerEnv = 8;

static size_t checkpoint_bytes(const rcw_batch* b) {
    return sizeof(CheckpointHeader) + (size_t)b->cfg.num_envs * (kCheckpointWordsPerEnv * 4 + 1);
}

int32_t rcw_checkpoint_size(rcw_batch* b, size_t* bytes) {
    if (int32_t rc = check_handle(b)) return rc;
    if (bytes) *bytes = checkpoint_bytes(b);
    return RCW_OK;
}

int32_t rcw_save_checkpoint(rcw_batch* b, void* host, size_t bytes) {
    NvtxRange nvtx("rcw_save_checkpoint");
    if (int32_t rc = check_handle(b)) return rc;
    if (!host) return fail(RCW_EINVAL, "host is null");
    if (bytes < checkpoint_bytes(b))
        return fail(RCW_ESIZE, "checkpoint buffer of %zu bytes, %zu needed", bytes, checkpoint_bytes(b));
    DeviceGuard g(b->device);
    const size_t E = (size_t)b->cfg.num_envs;
    const StateRef& s = b->st[b->cur];
    uint8_t* out = static_cast<uint8_t*>(host) + sizeof(CheckpointHeader);
    const void* words[kCheckpointWordsPerEnv] = {s.pos_x, s.pos_y, s.dir_au, s.goal, s.episode,
                                                 b->d_reward, b->d_ep_return, b->d_ep_length};
    for (size_t k = 0; k < kCheckpointWordsPerEnv; ++k)
        RCW_CUDA(cudaMemcpyAsync(out + k * E * 4, words[k], E * 4, cudaMemcpyDeviceToHost, b->stream));
    RCW_CUDA(cudaMemcpyAsync(out + kCheckpointWordsPerEnv * E * 4, b->d_done, E, cudaMemcpyDeviceToHost, b->stream));
    if (int32_t rc = sync_and_check(b, /*need_stats=*/true)) return rc;
    CheckpointHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = kCheckpointMagic;
    h.version = kCheckpointVersion;
    h.num_envs = b->cfg.num_envs;
    h.env_id_offset = b->cfg.env_id_offset;
    h.seed = b->cfg.seed;
    h.step_index = b->step_index;
    h.H = b->cfg.height_tile_map_tu;
    h.W = b->cfg.width_tile_map_tu;
    h.N = b->cfg.num_directions;
    h.episodes = b->h_stats->episodes;
    h.sum_length = b->h_stats->sum_length;
    h.sum_return = b->h_stats->sum_return;
    memcpy(host, &h, sizeof(h));
    return RCW_OK;
}

int32_t rcw_load_checkpoint(rcw_batch* b, const void* host, size_t bytes) {
    NvtxRange nvtx("rcw_load_checkpoint");
    if (int32_t rc = check_handle(b)) return rc;
    if (!host) return fail(RCW_EINVAL, "host is null");
    if (bytes < sizeof(CheckpointHeader)) return fail(RCW_ESIZE, "checkpoint of %zu bytes has no header", bytes);
    CheckpointHeader h;
    memcpy(&h, host, sizeof(h));
    if (h.magic != kCheckpointMagic || h.version != kCheckpointVersion)
        return fail(RCW_EINVAL, "not a librcw_b200 checkpoint (magic 0x%x, version %u)", h.magic, h.version);
    const rcw_config& c = b->cfg;
    if (h.num_envs != c.num_envs || h.H != c.height_tile_map_tu || h.W != c.width_tile_map_tu ||
        h.N != c.num_directions)
        return fail(RCW_ESIZE, "checkpoint of %lld envs on %dx%d tiles / %d directions does not match this batch "
                    "(%lld envs, %dx%d / %d)", (long long)h.num_envs, h.H, h.W, h.N, (long long)c.num_envs,
                    c.height_tile_map_tu, c.width_tile_map_tu, c.num_directions);
    if (bytes < checkpoint_bytes(b))
        return fail(RCW_ESIZE, "checkpoint truncated: %zu bytes, %zu needed", bytes, checkpoint_bytes(b));
    DeviceGuard g(b->device);
    const size_t E = (size_t)c.num_envs;
    const StateRef& s = b->st[b->cur];
    const uint8_t* in = static_cast<const uint8_t*>(host) + sizeof(CheckpointHeader);
    {   // the same validation as rcw_set_state: a corrupt snapshot must not index outside the tables
        const float* px = reinterpret_cast<const float*>(in);
        const float* py = px + E;
        const int32_t* au = reinterpret_cast<const int32_t*>(in + 2 * E * 4);
        const uint32_t* goal = reinterpret_cast<const uint32_t*>(in + 3 * E * 4);
        for (size_t e = 0; e < E; ++e) {
            const int gi = (int)(goal[e] & 0xFFFFu), gj = (int)(goal[e] >> 16);
            if (!(px[e] >= 0.0f) || !(px[e] < (float)c.height_tile_map_tu) || !(py[e] >= 0.0f) ||
                !(py[e] < (float)c.width_tile_map_tu) || au[e] < 0 || au[e] >= c.num_directions || gi < 1 ||
                gi > c.height_tile_map_tu || gj < 1 || gj > c.width_tile_map_tu)
                return fail(RCW_EINVAL, "checkpoint holds an invalid state for env %zu", e);
        }
    }
    void* words[kCheckpointWordsPerEnv] = {s.pos_x, s.pos_y, s.dir_au, s.goal, s.episode,
                                           b->d_reward, b->d_ep_return, b->d_ep_length};
    for (size_t k = 0; k < kCheckpointWordsPerEnv; ++k)
        RCW_CUDA(cudaMemcpyAsync(words[k], in + k * E * 4, E * 4, cudaMemcpyHostToDevice, b->stream));
    RCW_CUDA(cudaMemcpyAsync(b->d_done, in + kCheckpointWordsPerEnv * E * 4, E, cudaMemcpyHostToDevice, b->stream));
    DeviceStats st;
    memset(&st, 0, sizeof(st));
    st.episodes = h.episodes;
    st.sum_length = h.sum_length;
    st.sum_return = h.sum_return;
    RCW_CUDA(cudaMemcpyAsync(b->d_stats, &st, sizeof(st), cudaMemcpyHostToDevice, b->stream));
    RCW_CUDA(cudaStreamSynchronize(b->stream));   // `host` and `st` are only valid during the call
    b->step_index = h.step_index;
    b->cfg.seed = h.seed;
    b->cfg.env_id_offset = h.env_id_offset;
    return enqueue_frame(b, kModeRender, nullptr);
}

int32_t rcw_get_rays(rcw_batch* b, int64_t env0, int64_t n, int32_t* hit_ij, int32_t* hit_dim,
                     float* dist, float* ray_dir) {
    if (int32_t rc = check_handle(b)) return rc;
    if (env0 < 0 || n < 1 || env0 + n > b->cfg.num_envs)
        return fail(RCW_ESIZE, "env range [%lld, %lld) outside 0..%lld", (long long)env0,
                    (long long)(env0 + n), (long long)b->cfg.num_envs);
    DeviceGuard g(b->device);
    const size_t cnt = (size_t)n * (size_t)b->cfg.num_rays;
    int32_t *d_hit = nullptr, *d_dim = nullptr;
    float *d_dist = nullptr, *d_dir = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_hit);
        cudaFree(d_dim);
        cudaFree(d_dist);
        cudaFree(d_dir);
    };
    if (cudaMalloc((void**)&d_hit, cnt * 8) != cudaSuccess || cudaMalloc((void**)&d_dim, cnt * 4) != cudaSuccess ||
        cudaMalloc((void**)&d_dist, cnt * 4) != cudaSuccess || cudaMalloc((void**)&d_dir, cnt * 8) != cudaSuccess) {
        cleanup();
        cudaGetLastError();
        return fail(RCW_ENOMEM, "device allocation for the ray dump failed");
    }
    FrameParams p;
    fill_frame_params(b, p);
    p.env_first = env0;
    p.env_count = n;
    p.dump_hit = d_hit;
    p.dump_dim = d_dim;
    p.dump_dist = d_dist;
    p.dump_dir = d_dir;
    LaunchShape sh{false, false, false, grid_for(b, n)};
    sh.room = p.room != 0;
    cudaError_t e = launch_frame(p, kModeRays, b->cfg.obs_format, sh, b->stream);
    b->launches += 1;
    if (e == cudaSuccess && hit_ij) e = cudaMemcpyAsync(hit_ij, d_hit, cnt * 8, cudaMemcpyDeviceToHost, b->stream);
    if (e == cudaSuccess && hit_dim) e = cudaMemcpyAsync(hit_dim, d_dim, cnt * 4, cudaMemcpyDeviceToHost, b->stream);
    if (e == cudaSuccess && dist) e = cudaMemcpyAsync(dist, d_dist, cnt * 4, cudaMemcpyDeviceToHost, b->stream);
    if (e == cudaSuccess && ray_dir) e = cudaMemcpyAsync(ray_dir, d_dir, cnt * 8, cudaMemcpyDeviceToHost, b->stream);
    cudaError_t se = cudaStreamSynchronize(b->stream);
    cleanup();
    if (e == cudaSuccess) e = se;
    RCW_CUDA(e);
    return RCW_OK;
}

int32_t rcw_obs_device_ptr(rcw_batch* b, void** dptr, size_t* total_bytes, size_t* env_stride_bytes) {
    if (int32_t rc = check_handle(b)) return rc;
    if (dptr) *dptr = b->d_obs;
    if (total_bytes) *total_bytes = b->obs_bytes;
    if (env_stride_bytes) *env_stride_bytes = b->obs_env_stride;
    return RCW_OK;
}

int32_t rcw_obs_layout(rcw_batch* b, size_t* env_stride_bytes, size_t* column_stride_bytes,
                       size_t* column_bytes, int32_t* bytes_per_pixel) {
    if (int32_t rc = check_handle(b)) return rc;
    if (env_stride_bytes) *env_stride_bytes = b->obs_env_stride;
    if (column_stride_bytes) *column_stride_bytes = (size_t)b->col_pitch;
    if (column_bytes) *column_bytes = (size_t)::column_bytes(b);
    if (bytes_per_pixel) *bytes_per_pixel = b->bpp;
    return RCW_OK;
}

int32_t rcw_obs_frames(rcw_batch* b, int32_t* frame_stack, int32_t* newest, size_t* frame_stride_bytes) {
    if (int32_t rc = check_handle(b)) return rc;
    if (frame_stack) *frame_stack = b->frame_stack;
    if (newest) *newest = b->frame_newest;
    if (frame_stride_bytes) *frame_stride_bytes = b->frame_stride;
    return RCW_OK;
}

int32_t rcw_copy_obs_frame(rcw_batch* b, int64_t env0, int64_t n, int32_t age, void* host) {
    NvtxRange nvtx("rcw_copy_obs");
    if (int32_t rc = check_handle(b)) return rc;
    if (!host) return fail(RCW_EINVAL, "host is null");
    if (env0 < 0 || n < 1 || env0 + n > b->cfg.num_envs)
        return fail(RCW_ESIZE, "env range [%lld, %lld) outside 0..%lld", (long long)env0,
                    (long long)(env0 + n), (long long)b->cfg.num_envs);
    if (n > b->obs_window)
        return fail(RCW_ESIZE, "%lld envs requested, the observation window holds %lld", (long long)n,
                    (long long)b->obs_window);
    if (age < 0 || age >= b->frame_stack)
        return fail(RCW_ESIZE, "frame age %d outside 0..%d", age, b->frame_stack - 1);
    DeviceGuard g(b->device);
    const size_t R = (size_t)obs_columns(b->cfg), col_bytes = (size_t)column_bytes(b);   // columns of one observation
    const int64_t slot0 = env0 % b->obs_window;
    if (slot0 + n > b->obs_window) {   // the range wraps around the window: two pieces
        const int64_t n1 = b->obs_window - slot0;
        if (int32_t rc = rcw_copy_obs_frame(b, env0, n1, age, host)) return rc;
        return rcw_copy_obs_frame(b, env0 + n1, n - n1, age, static_cast<uint8_t*>(host) + (size_t)n1 * R * col_bytes);
    }
    const int pos = (b->frame_newest - age + b->frame_stack) % b->frame_stack;
    const uint8_t* src = b->d_obs + (size_t)slot0 * b->obs_env_stride + (size_t)pos * b->frame_stride;
    uint8_t* dst = static_cast<uint8_t*>(host);
    if (col_bytes == (size_t)b->col_pitch && R * col_bytes == b->obs_env_stride) {
        RCW_CUDA(cudaMemcpyAsync(dst, src, R * col_bytes * (size_t)n, cudaMemcpyDeviceToHost, b->stream));
    } else if (col_bytes == (size_t)b->col_pitch) {
        // dense frames, envs apart (frame ring or padded env stride): one 2-D copy, a row per env
        RCW_CUDA(cudaMemcpy2DAsync(dst, R * col_bytes, src, b->obs_env_stride, R * col_bytes, (size_t)n,
                                   cudaMemcpyDeviceToHost, b->stream));
    } else if (R * (size_t)b->col_pitch == b->obs_env_stride) {
        // pitched columns, envs back to back: one 2-D copy over all columns
        RCW_CUDA(cudaMemcpy2DAsync(dst, col_bytes, src, (size_t)b->col_pitch, col_bytes, R * (size_t)n,
                                   cudaMemcpyDeviceToHost, b->stream));
    } else {
        for (int64_t e = 0; e < n; ++e)
            RCW_CUDA(cudaMemcpy2DAsync(dst + (size_t)e * R * col_bytes, col_bytes,
                                       src + (size_t)e * b->obs_env_stride, (size_t)b->col_pitch, col_bytes, R,
                                       cudaMemcpyDeviceToHost, b->stream));
    }
    return sync_and_check(b);
}

int32_t rcw_expanded_layout(rcw_batch* b, int32_t pixel_format, size_t* env_stride_bytes,
                            size_t* column_stride_bytes, size_t* column_bytes) {
    if (int32_t rc = check_handle(b)) return rc;
    if (pixel_format != RCW_OBS_RGB8 && pixel_format != RCW_OBS_XRGB32 && pixel_format != RCW_OBS_GRAY8 &&
        pixel_format != RCW_OBS_GRAY16F)
        return fail(RCW_EINVAL, "pixel_format must be RCW_OBS_RGB8, RCW_OBS_XRGB32, RCW_OBS_GRAY8 or RCW_OBS_GRAY16F");
    const size_t cb = (size_t)b->cfg.height_camera_view_pu * bytes_per_pixel(pixel_format);
    const size_t pitch = (cb + 31) & ~(size_t)31;
    if (column_bytes) *column_bytes = cb;
    if (column_stride_bytes) *column_stride_bytes = pitch;
    if (env_stride_bytes) *env_stride_bytes = ((size_t)b->cfg.num_rays * pitch + 127) & ~(size_t)127;
    return RCW_OK;
}

int32_t rcw_expand_columns(rcw_batch* b, const uint32_t* columns, size_t columns_env_stride_bytes, int64_t n,
                           int32_t pixel_format, void* dst) {
    NvtxRange nvtx("rcw_expand_columns");
    if (int32_t rc = check_handle(b)) return rc;
    size_t env_stride = 0;
    if (int32_t rc = rcw_expanded_layout(b, pixel_format, &env_stride, nullptr, nullptr)) return rc;
    if (!columns || !dst) return fail(RCW_EINVAL, "columns / dst is null");
    if (!is_device_pointer(columns) || !is_device_pointer(dst))
        return fail(RCW_EINVAL, "columns and dst must be device pointers");
    if (columns_env_stride_bytes == 0) columns_env_stride_bytes = (size_t)b->cfg.num_rays * 4;
    if (columns_env_stride_bytes % 4 || columns_env_stride_bytes < (size_t)b->cfg.num_rays * 4 ||
        columns_env_stride_bytes / 4 > 0xFFFFFFFFull)
        return fail(RCW_EINVAL, "columns_env_stride_bytes must be a multiple of 4 and at least num_rays * 4");
    if (n < 1 || n * b->gpe >= (1LL << 31) || n > 0xFFFFFFFFLL)
        return fail(RCW_ESIZE, "n = %lld out of range", (long long)n);
    if ((uintptr_t)dst % 32) return fail(RCW_EINVAL, "dst must be 32-byte aligned");
    DeviceGuard g(b->device);
    FrameParams p;
    fill_frame_params(b, p, pixel_format);
    p.col_info = const_cast<uint32_t*>(columns);
    p.col_info_stride = (uint32_t)(columns_env_stride_bytes / 4);
    p.obs = static_cast<uint8_t*>(dst);
    p.obs_env_stride = env_stride;
    p.obs_window = (uint32_t)n;
    p.obs_slot0 = 0;
    p.env_first = 0;
    p.env_count = n;
    RCW_CUDA(launch_expand_columns(p, pixel_format, grid_for(b, n), b->stream));
    b->launches += 1;
    b->expand_pending = true;   // the words are the caller's: a bad one is clamped and reported by the next blocking call
    return RCW_OK;
}

int32_t rcw_copy_obs(rcw_batch* b, int64_t env0, int64_t n, void* host) {
    return rcw_copy_obs_frame(b, env0, n, 0, host);
}

int32_t rcw_episode_stats(rcw_batch* b, int64_t* episodes, double* sum_return, int64_t* sum_length,
                          int32_t reset_counters) {
    if (int32_t rc = check_handle(b)) return rc;
    DeviceGuard g(b->device);
    if (int32_t rc = sync_and_check(b, /*need_stats=*/true)) return rc;
    if (episodes) *episodes = (int64_t)b->h_stats->episodes;
    if (sum_return) *sum_return = b->h_stats->sum_return;
    if (sum_length) *sum_length = (int64_t)b->h_stats->sum_length;
    if (reset_counters) RCW_CUDA(cudaMemsetAsync(b->d_stats, 0, sizeof(DeviceStats), b->stream));
    return RCW_OK;
}

int32_t rcw_launch_count(rcw_batch* b, int64_t* launches) {
    if (int32_t rc = check_handle(b)) return rc;
    if (launches) *launches = b->launches;
    return RCW_OK;
}

int32_t rcw_stream(rcw_batch* b, void** stream) {
    if (int32_t rc = check_handle(b)) return rc;
    if (stream) *stream = (void*)b->stream;
    return RCW_OK;
}

int32_t rcw_sync(rcw_batch* b) {
    if (int32_t rc = check_handle(b)) return rc;
    DeviceGuard g(b->device);
    return sync_and_check(b);
}

// ---- one process, several GPUs: independent env shards, one handle per device ----------
// Every enqueueing entry point returns as soon as its kernels are queued, so ONE host thread keeps all the
// GPUs busy: the sharded calls below just walk the handles.  No collective anywhere: the only cross-shard
// quantity, the episode totals, is three scalars summed on the host.

int32_t rcw_shard_envs(int64_t total_envs, int32_t n_shards, int32_t shard, int64_t* offset, int64_t* count) {
    if (n_shards < 1 || shard < 0 || shard >= n_shards) return fail(RCW_EINVAL, "bad shard %d of %d", shard, n_shards);
    if (total_envs < 0) return fail(RCW_EINVAL, "total_envs must be non-negative");
    const int64_t base = total_envs / n_shards, extra = total_envs % n_shards;
    if (count) *count = base + (shard < extra ? 1 : 0);
    if (offset) *offset = (int64_t)shard * base + (shard < extra ? shard : extra);
    return RCW_OK;
}

int32_t rcw_create_sharded(const rcw_config* cfg, const float* directions_wu, const int32_t* devices, int32_t n_shards,
                           rcw_batch** handles) {
    if (!cfg || !handles) return fail(RCW_EINVAL, "cfg/handles is null");
    if (n_shards < 1 || n_shards > 64) return fail(RCW_EINVAL, "n_shards must be 1..64 (got %d)", n_shards);
    for (int32_t k = 0; k < n_shards; ++k) handles[k] = nullptr;
    if (cfg->struct_size != sizeof(rcw_config))
        return fail(RCW_ESIZE, "rcw_config.struct_size is %u, this library expects %zu", cfg->struct_size, sizeof(rcw_config));
    if (cfg->num_envs < n_shards) return fail(RCW_EINVAL, "fewer envs (%lld) than shards (%d)", (long long)cfg->num_envs, n_shards);
    for (int32_t k = 0; k < n_shards; ++k) {
        rcw_config c = *cfg;
        int64_t off = 0, cnt = 0;
        rcw_shard_envs(cfg->num_envs, n_shards, k, &off, &cnt);
        c.num_envs = cnt;
        c.env_id_offset = cfg->env_id_offset + off;   // global env ids key the Philox streams: results do not depend on n_shards
        c.device = devices ? devices[k] : k;
        if (c.obs_window_envs > cnt) c.obs_window_envs = 0;
        const int32_t rc = rcw_create(&c, directions_wu, &handles[k]);
        if (rc != RCW_OK) {
            std::string keep = t_last_error;
            for (int32_t j = 0; j < k; ++j) {
                rcw_destroy(handles[j]);
                handles[j] = nullptr;
            }
            t_last_error = "shard " + std::to_string(k) + " (device " + std::to_string(c.device) + "): " + keep;
            return rc;
        }
    }
    return RCW_OK;
}

int32_t rcw_destroy_sharded(rcw_batch* const* handles, int32_t n_shards) {
    if (!handles) return RCW_OK;
    int32_t first = RCW_OK;
    for (int32_t k = 0; k < n_shards; ++k) {
        const int32_t rc = rcw_destroy(handles[k]);
        if (first == RCW_OK) first = rc;
    }
    return first;
}

static int32_t check_shards(rcw_batch* const* handles, int32_t n_shards) {
    if (!handles || n_shards < 1) return fail(RCW_EINVAL, "no handles");
    for (int32_t k = 0; k < n_shards; ++k)
        if (!handles[k]) return fail(RCW_EINVAL, "handle %d is null", k);
    return RCW_OK;
}

int32_t rcw_step_sharded(rcw_batch* const* handles, int32_t n_shards, const uint8_t* actions) {
    if (int32_t rc = check_shards(handles, n_shards)) return rc;
    if (!actions) return fail(RCW_EINVAL, "actions is null (use rcw_step_random_sharded for the random policy)");
    if (is_device_pointer(actions)) return fail(RCW_EINVAL, "rcw_step_sharded takes a HOST array of the whole batch");
    // like the reference's @assert, an invalid action anywhere means nothing is enqueued anywhere
    int64_t off = 0;
    for (int32_t k = 0; k < n_shards; ++k) {
        if (int32_t rc = validate_host_actions(actions + off, 0, handles[k]->cfg.num_envs)) return rc;
        off += handles[k]->cfg.num_envs;
    }
    off = 0;
    for (int32_t k = 0; k < n_shards; ++k) {
        if (int32_t rc = rcw_step(handles[k], actions + off)) return rc;
        off += handles[k]->cfg.num_envs;
    }
    return RCW_OK;
}

int32_t rcw_step_random_sharded(rcw_batch* const* handles, int32_t n_shards, int32_t n_steps) {
    if (int32_t rc = check_shards(handles, n_shards)) return rc;
    if (n_steps < 0) return fail(RCW_EINVAL, "n_steps must be non-negative");
    for (int32_t s = 0; s < n_steps; ++s)            // step-major: every GPU always has work queued
        for (int32_t k = 0; k < n_shards; ++k)
            if (int32_t rc = rcw_step_random(handles[k], 1)) return rc;
    return RCW_OK;
}

int32_t rcw_sync_sharded(rcw_batch* const* handles, int32_t n_shards) {
    if (int32_t rc = check_shards(handles, n_shards)) return rc;
    int32_t first = RCW_OK;
    std::string msg;
    for (int32_t k = 0; k < n_shards; ++k) {
        const int32_t rc = rcw_sync(handles[k]);
        if (rc != RCW_OK && first == RCW_OK) {
            first = rc;
            msg = t_last_error;
        }
    }
    if (first != RCW_OK) t_last_error = msg;
    return first;
}

int32_t rcw_reduce_episode_stats(rcw_batch* const* handles, int32_t n_shards, int64_t* episodes, double* sum_return,
                                 int64_t* sum_length, int32_t reset_counters) {
    if (int32_t rc = check_shards(handles, n_shards)) return rc;
    int64_t ep = 0, sl = 0;
    double sr = 0.0;
    for (int32_t k = 0; k < n_shards; ++k) {          // fixed order: the double sum is reproducible
        int64_t e = 0, l = 0;
        double r = 0.0;
        if (int32_t rc = rcw_episode_stats(handles[k], &e, &r, &l, reset_counters)) return rc;
        ep += e;
        sr += r;
        sl += l;
    }
    if (episodes) *episodes = ep;
    if (sum_return) *sum_return = sr;
    if (sum_length) *sum_length = sl;
    return RCW_OK;
}

}  // extern "C"
