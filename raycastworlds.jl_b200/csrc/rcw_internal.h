// rcw_internal.h — structures shared by the kernels and the C-ABI layer of librcw_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rcw_b200.h"

namespace rcw {

#ifndef RCW_WARPS_PER_CTA
#define RCW_WARPS_PER_CTA 4
#endif
#ifndef RCW_WARPS_PER_SM_LO
#define RCW_WARPS_PER_SM_LO 20   // store-bound variant: 20 warps/SM, 91 registers (16-20 measured best; 24: -0.7 %, 28: -2.4 %, 12: -4 %)
#endif
#ifndef RCW_WARPS_PER_SM_HI
#define RCW_WARPS_PER_SM_HI 32   // register budget of the front-bound variant (LaunchShape::occ4): <= 64 registers
#endif
#ifndef RCW_DDA_STEPS_PER_VOTE
#define RCW_DDA_STEPS_PER_VOTE 2
#endif
#ifndef RCW_PAIR_UNROLL
#define RCW_PAIR_UNROLL 2
#endif
#ifndef RCW_PASS1_UNROLL
#define RCW_PASS1_UNROLL 4
#endif
constexpr int kWarpsPerCta = RCW_WARPS_PER_CTA;  // one warp = one (env, 32-ray group) work item at a time;
                                                 // 4 measured best (8: -1.2 %, 16: -18 %, 2 and 1: same as 4)
constexpr int kThreadsPerCta = kWarpsPerCta * 32;
constexpr int kCtasPerSmLo = RCW_WARPS_PER_SM_LO / RCW_WARPS_PER_CTA;   // __launch_bounds__ minimum, store-bound variant
constexpr int kCtasPerSmHi = RCW_WARPS_PER_SM_HI / RCW_WARPS_PER_CTA;   // front-bound variant
constexpr int kPass1Unroll = RCW_PASS1_UNROLL;  // independent 16-byte stores kept in flight per warp
constexpr int kDdaStepsPerVote = RCW_DDA_STEPS_PER_VOTE;
constexpr int kPairUnroll = RCW_PAIR_UNROLL;    // mirror-pair renderer: two stores per unrolled iteration
constexpr int kDirSlots = 8;              // constant-memory slots for direction tables
constexpr int kDirSlotEntries = 512;      // float2 entries per slot (4 KB)

enum FrameMode : int {
    kModeStep = 0,    // act! (+ auto-reset) -> cast_rays! -> update_camera_view!
    kModeRender = 1,  // cast_rays! -> update_camera_view! from the current state
    kModeRays = 2     // cast_rays! only, results dumped to global arrays (parity / debug)
};

// Totals over finished episodes + sticky error flags, one per handle, in device memory.
struct DeviceStats {
    unsigned long long episodes;
    unsigned long long sum_length;
    double sum_return;
    int bad_action;   // set when a device-side action array held a value outside 1..4
    int bad_columns;  // set when rcw_expand_columns read a column word with a palette index outside
                      // RCW_COLOR_WALL_1..RCW_COLOR_GOAL_2 or more ceiling rows than half the column (clamped)
};

// Struct-of-arrays environment state in HBM.  The fields every warp of an env reads
// (position, direction, goal, episode) are double-buffered: a step reads buffer `in`
// and the env's group-0 warp writes buffer `out`, so the 16 warps that render one env
// never race with the one that advances it.
struct StateRef {
    float* pos_x;
    float* pos_y;
    int32_t* dir_au;
    uint32_t* goal;      // (i | j << 16), 1-based tile
    uint32_t* episode;   // Philox episode counter of the env
};

struct FrameParams {
    // geometry
    int32_t H, W;            // tiles
    int32_t wpr;             // 32-bit words per bit-packed map row (bit j-1 of row i-1)
    int32_t map_words;       // H * wpr rounded up to a multiple of 4 (16-byte TMA granule): one layer
    int32_t stage_words;     // words staged in shared memory per map: map_words, or (n_extra + 2) * map_words with extra
                             // object layers ([wall][extra ...][any], see BitsMap)
    int32_t n_extra;         // object layers beyond WALL and GOAL (rcw_config.num_object_layers - 2)
    uint32_t layer_terminal_mask;   // bit k: extra layer k ends the episode like GOAL (else it blocks like WALL)
    float layer_reward[RCW_MAX_EXTRA_LAYERS];
    int32_t n_colors;        // wall / goal / extra-layer colours a column can have: 4 + 2 * n_extra (col_table rows per pad)
    int32_t N;               // num_directions
    int32_t R;               // num_rays == observation width
    int32_t P;               // height_camera_view_pu
    int32_t gpe;             // 32-ray groups per env = ceil(R / 32)
    int32_t px_bytes;        // bytes per pixel of the format being painted (1, 2, 3 or 4)
    int32_t col_bytes;       // P * bytes per pixel
    int32_t col_pitch;       // bytes between consecutive columns: col_bytes rounded up to a multiple of 32
    uint32_t dda_flags;
    uint32_t closed_border;  // every border tile of (every env's) wall layer is a wall: rays cannot leave the map
    // scalars of the reference constructor
    float radius, incr, goal_reward;
    float hl_num;            // camera_height_tile_wu * Float32(num_rays)   (single_room.jl:406)
    float two_s;             // 2 * semi_field_of_view_wu
    uint32_t palette[6 + 2 * RCW_MAX_EXTRA_LAYERS];   // RCW_COLOR_*, then extra layer k at 6 + 2 k (dim 1), 7 + 2 k (dim 2)
    uint2 col_entry[6 + 2 * RCW_MAX_EXTRA_LAYERS];    // renderer's {slow << 31, colour word} of a column painted with palette entry k: the
                             // word is the byte-replicated colour when entry k, ceiling and floor all have R == G == B
                             // (or the format has whole-word pixels); otherwise the column takes the phase-rotated path
    uint32_t gpe_magic;      // item / gpe for any 32-bit item (gpe >= 2), branch-free round-up method (libdivide):
    uint32_t gpe_shift;      //   t = umulhi(gpe_magic, n); q = (((n - t) >> 1) + t) >> gpe_shift
    // the renderer sweeps units of 32 (mirror pairs: 64) bytes, U per column (render_span): lane L starts in
    // column (L * unit_inv16) >> 16 and advances by (unit_adv_cl columns, unit_adv_u units) per iteration
    uint32_t unit_inv16;     // U < 32 ? ceil(65536 / U) : 0
    int32_t unit_adv_cl;     // 32 / U
    int32_t unit_adv_u;      // 32 % U
    // the same for U = col_pitch / 32 sectors per column (table renderer of env_kernel)
    uint32_t sec_inv16;
    int32_t sec_adv_cl, sec_adv_u;
    const uint8_t* col_table;// [P / 2 + 1 rows of ceiling][4 colours WALL_1..GOAL_2][col_pitch] ready-made pitched columns
                             // (nullptr: none — the table is only built for small columns, LaunchShape::table)
    uint32_t room;           // the wall layer is exactly the border of the map (and H, W <= 32767): RoomMap kernels
    // tables
    int32_t dir_slot;        // >= 0: directions live in constant memory slot; < 0: use `dirs`
    const float2* dirs;      // [N] unit vectors (global copy, also the source of the ray table)
    const float4* ray_table; // [N][R] {ray_x, ray_y, |1/ray_x|, |1/ray_y|}
    const uint32_t* wall_map;// [map_words] shared wall layer, bit-packed, or [num_envs][map_words] per env
    uint32_t map_env_stride; // words between the wall layers of consecutive envs; 0 = one layer shared by all
    const uint8_t* patterns; // [6][pat_stride] single-colour byte runs, one per palette entry (bulk renderer)
    int32_t pat_stride;      // bytes, multiple of 16: min(col_bytes, 3072) + 32 rounded up
    uint32_t* col_info;      // [env slots][col_info_stride] pad | palette index << 16, column order: the
                             // RCW_OBS_COLUMNS observation itself, or the hand-over of a split launch
    uint32_t col_info_stride;// words between consecutive env slots of col_info (>= R)
    // state
    StateRef in, out;
    const uint8_t* render_mask;  // kModeRender only: nonzero = redraw this env (nullptr: all) — masked resets
    const uint8_t* actions;  // device, 1..4 per env; nullptr => random policy
    float* reward;
    uint8_t* done;
    float* host_reward;      // kModeStep, optional: slot of the pinned, mapped result ring (rcw_step_async) that the
    uint8_t* host_done;      //   env's writer lane also stores reward / done to — no copy behind the kernel
    float* ep_return;
    uint32_t* ep_length;
    DeviceStats* stats;
    uint8_t* obs;
    size_t obs_env_stride;   // bytes, multiple of 16
    uint32_t obs_window;     // env slots in the observation buffer (== num_envs unless rcw_config.obs_window_envs)
    uint32_t obs_slot0;      // slot of env_first; env_first + k lives in slot (obs_slot0 + k) mod obs_window
    int64_t num_envs;
    int64_t env_first;       // first env processed by this launch (kModeRays: dump window)
    int64_t env_count;       // envs processed by this launch
    uint64_t env_id_offset;
    uint64_t seed;
    uint64_t step_index;
    int32_t auto_reset;
    // kModeRays outputs, indexed relative to env_first
    int32_t* dump_hit;       // [n][R][2]
    int32_t* dump_dim;       // [n][R]
    float* dump_dist;        // [n][R]
    float* dump_dir;         // [n][R][2]
};

// Host-supplied actions of one launch, two bits per env (action - 1), passed by value in the kernel parameter
// space (CUDA 12.1+: up to 32,764 bytes of parameters): no pinned staging, no host-to-device copy in front of
// the kernel, the warps read their two bits with one constant-bank load.  Env k of the launch (relative to
// FrameParams::env_first) is bits [2 (k mod 16), 2 (k mod 16) + 2) of w[k / 16].
constexpr int kPackedActionEnvs = 32768;
struct PackedActions {
    uint32_t w[kPackedActionEnvs / 16];
};

struct ResetParams {
    int32_t H, W, wpr, N;
    int32_t n_extra, map_words;   // extra object layers behind the wall layer (see BitsMap)
    const uint32_t* wall_map;
    uint32_t map_env_stride;   // 0 = shared wall layer
    StateRef st;
    float* reward;
    uint8_t* done;
    float* ep_return;
    uint32_t* ep_length;
    int64_t num_envs;
    uint64_t env_id_offset;
    uint64_t seed;
    const int32_t* goal_ij;   // device copies of the host layout, or nullptr => Philox
    const int32_t* player_ij;
    const int32_t* dir_au;
    const uint8_t* mask;      // nullptr => all
};

// update_top_view! for envs [env_first, env_first + env_count) (rcw_topview.cuh)
struct TopViewParams {
    int32_t H, W, wpr, map_words;
    int32_t N, R;
    int32_t pu;              // pu_per_tu
    int32_t Hp, Wp;          // H * pu, W * pu
    uint32_t dda_flags;
    uint32_t closed_border;
    float radius;            // player_radius_wu
    uint32_t palette[6];     // RCW_TOP_COLOR_*
    int32_t n_extra;         // extra object layers, staged behind the wall layer
    int32_t stage_words;     // (n_extra ? n_extra + 2 : 1) * map_words
    uint32_t extra_color[RCW_MAX_EXTRA_LAYERS];   // tile_map_colors[3 + k]
    int32_t dir_slot;
    const float2* dirs;
    const float4* ray_table;
    const uint32_t* wall_map;
    uint32_t map_env_stride;
    StateRef st;             // the state to draw
    const uint8_t* mask;     // nonzero = draw this env (nullptr: all)
    uint8_t* top;            // [window][env_stride] bytes; one env = uint32 [Wp][Hp], row fastest
    size_t env_stride;       // bytes, multiple of 128
    uint32_t window, slot0;  // env_first + k lives in slot (slot0 + k) mod window
    int64_t env_first, env_count;
    uint32_t sm_count;       // SMs of the device
    uint32_t room;           // the wall layer is exactly the border of the map: RoomMap kernel, nothing staged
};

// kernel launchers (rcw_kernels.cu)
cudaError_t launch_build_ray_table(const float2* dirs, int N, int R, float sfov, float4* table,
                                   cudaStream_t s);
// how a frame launch is shaped (decided per handle in rcw_capi.cu)
struct LaunchShape {
    bool bulk;    // renderer variant: per-lane TMA bulk stores (measured alternative)
    bool split;   // two launches (front, paint) instead of the fused kernel (measured alternative)
    bool occ4;    // fused kernel compiled for 4 CTAs per SM (steps bound by act! / DDA rather than by stores)
    int ctas;     // grid size
    bool env_per_warp = false;   // small items: env_kernel, one warp = one env
    bool room = false;           // wall layer == border of the map: kernels without a map in shared memory (RoomMap)
    bool table = false;          // env_kernel copies ready-made columns out of FrameParams::col_table
};
// packed != nullptr (step mode, fused path, env_count <= kPackedActionEnvs): the actions ride in the parameters
cudaError_t launch_frame(const FrameParams& p, int mode, int obs_format, const LaunchShape& sh, cudaStream_t s,
                         const PackedActions* packed = nullptr);
// RCW_OBS_COLUMNS -> pixels: p.col_info (column words of p.env_count envs) painted into p.obs in pixel_format
cudaError_t launch_expand_columns(const FrameParams& p, int pixel_format, int ctas, cudaStream_t s);
cudaError_t launch_reset(const ResetParams& p, cudaStream_t s);
// state `from` -> state `to` for envs [env0, env0 + n): makes a range step visible in the buffer it read
cudaError_t launch_commit_range(const StateRef& from, const StateRef& to, int64_t env0, int64_t n, cudaStream_t s);
size_t top_view_smem_bytes(int H, int W, int R, int pu, float radius, int map_words);
cudaError_t launch_top_view(const TopViewParams& p, cudaStream_t s);
cudaError_t upload_dir_slot(int slot, const float2* host_dirs, int n, cudaStream_t s);

}  // namespace rcw
