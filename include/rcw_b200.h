/*
 * rcw_b200.h — C ABI of librcw_b200.so: the B200-native batched SingleRoom engine.
 *
 * This is the drop-in boundary for the hot path of RayCastWorlds.jl
 *   act!(world)  ->  cast_rays!(world)  ->  update_camera_view!(env)
 * (reference: src/single_room.jl:333-340, top view excluded).  The reference has no
 * FFI of its own (it is pure Julia); the seam it offers is multiple dispatch on
 * `AbstractGame` (src/RayCastWorlds.jl:5-14).  A Julia host keeps that API and calls
 * these entry points with `ccall` (see INTEGRATION.md and
 * raycastworlds.jl_b200/julia/BatchedRayCastWorlds.jl).  Each entry point below cites
 * the reference function it replaces.
 *
 * Conventions
 *   - every call returns RCW_OK (0) or a negative rcw_status; nothing throws or aborts;
 *     the message of the last failure on the calling thread is rcw_last_error().
 *   - all host pointers are caller-owned and only touched during the call.
 *   - tile indices are 1-based (i in 1..height_tu, j in 1..width_tu), angles are
 *     0..num_directions-1, exactly as in the reference (README.md:75-83).
 *   - one handle = one device + one stream; calls on one handle must be serialised by
 *     the caller; different handles may be driven from different host threads.
 *   - rcw_step*, rcw_reset, rcw_render only enqueue work; rcw_sync / rcw_get_* /
 *     rcw_copy_obs block until the handle's stream has drained.
 *   - there is NO CPU fallback: without a CUDA device rcw_create fails with RCW_ECUDA.
 */
#ifndef RCW_B200_H
#define RCW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCW_ABI_VERSION 4

typedef enum rcw_status {
    RCW_OK      = 0,
    RCW_EINVAL  = -1, /* bad argument / bad config                                   */
    RCW_EACTION = -2, /* action outside 1..4  (reference: @assert, single_room.jl:140) */
    RCW_ECUDA   = -3, /* CUDA runtime error (or no device)                            */
    RCW_ENOMEM  = -4, /* host or device allocation failed                             */
    RCW_ESIZE   = -5  /* struct_size / range mismatch                                 */
} rcw_status;

/* Observation formats.  Both are "column-major" like the reference's
 * camera_view::Array{UInt32}(height_px, num_rays) (single_room.jl:300): the pixel row is
 * the fastest index after the channel, so one ray's column is contiguous. */
typedef enum rcw_obs_format {
    RCW_OBS_RGB8   = 0, /* uint8  [num_envs][num_rays columns][height_px][3]  (R,G,B bytes of the reference pixel) */
    RCW_OBS_XRGB32 = 1, /* uint32 [num_envs][num_rays columns][height_px]     (bit-identical to the reference's UInt32 pixels) */
    RCW_OBS_GRAY8  = 2, /* uint8  [num_envs][num_rays columns][height_px]     learner-facing: BT.601 luma of the reference
                           pixel, (77 R + 150 G + 29 B + 128) >> 8; a third of the RGB8 write */
    RCW_OBS_COLUMNS = 3, /* uint32 [num_envs][num_rays columns]: the camera view BEFORE it is expanded into pixels —
                           what update_camera_view! decides per ray (single_room.jl:404-439): word = pad | cid << 16,
                           pad = rows of ceiling = rows of floor (0: the whole column has the wall colour; the wall
                           band has height_px - 2 pad rows), cid = RCW_COLOR_WALL_1 .. RCW_COLOR_GOAL_2.  Lossless
                           (rcw_expand_columns reproduces the pixel image bit for bit) at 4 bytes per column instead
                           of height_px * 3: 2 KB instead of 393 KB per default frame, for replay buffers that
                           rasterise only the frames they sample.  NOT rendered pixels: steps in
                           this format are bound by act! and the DDA, not by HBM writes. */
    RCW_OBS_GRAY16F = 4, /* float16 [num_envs][num_rays columns][height_px]: the GRAY8 luma divided by 255 and rounded to IEEE
                           binary16 — the normalised frame a convolutional learner takes, without a conversion kernel
                           behind the renderer (single_room.jl:576 is the consumer hand-off) */
    RCW_OBS_GRAY8_HALF = 5 /* uint8 [num_envs][num_rays / 2 columns][height_px / 2]: the GRAY8 frame under a 2 x 2 box filter,
                           (a + b + c + d + 2) >> 2 over two adjacent columns x two adjacent rows, computed from the two
                           rays' column decisions without ever writing the full-resolution frame: a quarter of the GRAY8
                           bytes with every ray still contributing (anti-aliased, unlike a camera of half the rays).
                           num_rays and height_px must be even.   */
    /* dense on the host (rcw_copy_obs); on the device columns may be pitched, see rcw_obs_layout */
} rcw_obs_format;

/* Indices into rcw_config.palette (reference values: single_room.jl:291-296). */
enum {
    RCW_COLOR_CEILING = 0, /* 0x00FFFFFF */
    RCW_COLOR_FLOOR   = 1, /* 0x00404040 */
    RCW_COLOR_WALL_1  = 2, /* 0x00808080  wall hit across dimension 1 */
    RCW_COLOR_WALL_2  = 3, /* 0x00c0c0c0 */
    RCW_COLOR_GOAL_1  = 4, /* 0x00800000 */
    RCW_COLOR_GOAL_2  = 5  /* 0x00c00000 */
};

/* Object layers beyond the reference's NUM_OBJECTS = 2 (WALL = 1, GOAL = 2; single_room.jl:16-18): object 3 + k is
 * "extra layer k" (SURVEY.md 8(f) N2).  Every object stops rays (:209, any over the layers); a column is painted with
 * the colours of the first object on the hit tile (:417-429 generalised); the top view shows findfirst over the
 * layers (:355-360); the player is never placed on an object (utils.jl:27); and a layer either refuses the move like
 * WALL or ends the episode with its own reward like GOAL (:162-176; terminal layers are checked in object order
 * after GOAL).  In RCW_OBS_COLUMNS words extra layer k has the colour ids 6 + 2 k (hit across dimension 1) and
 * 7 + 2 k. */
#define RCW_MAX_EXTRA_LAYERS 4
enum {
    RCW_LAYER_BLOCKING = 0, /* like WALL: the move is refused, reward 0                  */
    RCW_LAYER_TERMINAL = 1  /* like GOAL: reward = layer_reward[k], done, no move        */
};

/* Indices into rcw_config.top_palette: the colours of the top view (single_room.jl:288-290, 364-367). */
enum {
    RCW_TOP_COLOR_WALL   = 0, /* 0x00FFFFFF  tile_map_colors[WALL]            */
    RCW_TOP_COLOR_GOAL   = 1, /* 0x00FF0000  tile_map_colors[GOAL]            */
    RCW_TOP_COLOR_EMPTY  = 2, /* 0x00000000  tile_map_colors[end]: no object  */
    RCW_TOP_COLOR_BORDER = 3, /* 0x00cccccc  one-pixel border of every tile   */
    RCW_TOP_COLOR_RAY    = 4, /* 0x00808080  ray_color                        */
    RCW_TOP_COLOR_PLAYER = 5  /* 0x00c0c0c0  player_color                     */
};

/* Switches for the behaviour of RayCaster.cast_ray that the reference tree does not pin
 * (RayCaster.jl 0.1 is not vendored; see DESIGN.md "Unpinned decisions" D1/D2). */
enum {
    RCW_DDA_TIE_LE    = 1u << 0, /* D1: advance along dimension 1 when side_x <= side_y (default: strict <) */
    RCW_DDA_DIST_POST = 1u << 1  /* D2: distance = side - delta after the loop (default: side before the increment) */
};

/* Keyword arguments of SingleRoom(...) (single_room.jl:258-272) plus the batch fields. */
typedef struct rcw_config {
    uint32_t struct_size;            /* = sizeof(rcw_config); checked by rcw_create            */
    int32_t  device;                 /* CUDA device ordinal                                    */
    int64_t  num_envs;               /* environments owned by this handle                      */
    int64_t  env_id_offset;          /* global id of env 0 (multi-GPU shards); keys the RNG    */
    int32_t  height_tile_map_tu;     /* default 8                                              */
    int32_t  width_tile_map_tu;      /* default 16                                             */
    int32_t  num_directions;         /* default 128                                            */
    int32_t  num_rays;               /* default 512                                            */
    int32_t  height_camera_view_pu;  /* default 256                                            */
    float    player_radius_wu;       /* default 1/8                                            */
    float    position_increment_wu;  /* default 1/8                                            */
    float    semi_field_of_view_wu;  /* default 2/3 (rounded to f32)                           */
    float    camera_height_tile_wu;  /* default 1                                              */
    float    goal_reward;            /* default 1                                              */
    int32_t  obs_format;             /* rcw_obs_format, default RCW_OBS_RGB8                   */
    int32_t  auto_reset;             /* 1: a terminated env is re-drawn inside the same step   */
    uint64_t seed;                   /* Philox4x32-10 key                                      */
    uint32_t palette[6];             /* 0x00RRGGBB, indices RCW_COLOR_*                        */
    uint32_t dda_flags;              /* RCW_DDA_* (0 = default contract)                       */
    int32_t  obs_window_envs;        /* 0: the observation buffer holds every env (default).  K > 0: it holds K
                                        env slots and env e is rendered into slot e mod K, so a batch whose
                                        observations exceed HBM (2^20 default-camera envs = 412 GB) can still be
                                        stepped: rcw_step* renders the batch window by window (every frame is
                                        still written to HBM), rcw_step_range steps one window for a learner
                                        that consumes it before the next one is rendered.                  */
    int32_t  top_view;               /* 1: update_top_view! runs inside every rcw_step* / rcw_reset / rcw_render, the
                                        reference's act!(env) sequence (single_room.jl:333-340); 0 (default for a
                                        batch): only on rcw_render_top_view — the top view is a debug picture and
                                        doubles the HBM writes of a step                                       */
    int32_t  pu_per_tu;              /* pixels per tile of the top view, default 32 (single_room.jl:269)       */
    uint32_t top_palette[6];         /* 0x00RRGGBB, indices RCW_TOP_COLOR_*                                    */
    int32_t  frame_stack;            /* 0 or 1 (default): one frame per env.  K > 1 (up to 64): the observation buffer
                                        keeps the K most recent frames of every env in a ring — [env][K][columns][rows]
                                        — and every rcw_step / rcw_step_random writes the next ring position, so a
                                        learner that stacks frames reads them in place instead of copying K - 1 old
                                        frames per step.  rcw_obs_frames tells which position is
                                        the newest.  Resets and rcw_render overwrite the newest frame; not combinable
                                        with obs_window_envs / rcw_step_range.                                     */
    int32_t  result_ring;            /* 0 (default): none.  D in 1..64: rcw_step_async is available — the step kernel
                                        also writes every env's reward and done straight into slot (ticket mod D) of a
                                        pinned host ring (mapped memory: no copy is enqueued behind the kernel), and
                                        rcw_wait(ticket) hands out the slot once that step has finished.  With D >= 2 a
                                        host loop can enqueue step k + 1 before it reads the results of step k, so the
                                        device never idles while the host wakes up.                               */
    int32_t  num_object_layers;      /* NUM_OBJECTS (single_room.jl:16): 2 = WALL, GOAL (default; 0 means 2) ... 6.  Objects
                                        3 .. num_object_layers are extra static layers, empty until rcw_set_layer    */
    int32_t  layer_kind[RCW_MAX_EXTRA_LAYERS];       /* RCW_LAYER_BLOCKING / RCW_LAYER_TERMINAL per extra layer      */
    float    layer_reward[RCW_MAX_EXTRA_LAYERS];     /* reward of a terminal extra layer (goal_reward's counterpart) */
    uint32_t layer_palette[RCW_MAX_EXTRA_LAYERS][2]; /* camera-view colours 0x00RRGGBB when hit across dimension 1 / 2 */
    uint32_t layer_top_color[RCW_MAX_EXTRA_LAYERS];  /* top-view tile colour, tile_map_colors[3 + k]                 */
    uint32_t reserved[3];            /* must be zero                                           */
} rcw_config;

typedef struct rcw_batch rcw_batch; /* opaque */

/* ---- lifetime ------------------------------------------------------------------------- */

int32_t rcw_version(void);

/* Fill cfg with the reference's defaults (single_room.jl:43-52,258-272,288-296). */
int32_t rcw_config_init(rcw_config* cfg);

/* Replaces SingleRoom(...) / SingleRoomWorld(...) (single_room.jl:42-108,258-324) for a batch.
 * directions_wu: [num_directions][2] float32 unit vectors; pass the Julia host's own table so
 * the reference's cos/sin values are used (single_room.jl:65-69).  NULL => computed here as
 * (float)cos(theta), (float)sin(theta), theta = i*2*pi/num_directions in double.
 * The new batch is reset with on-device Philox draws and rendered once. */
int32_t rcw_create(const rcw_config* cfg, const float* directions_wu, rcw_batch** out);
int32_t rcw_destroy(rcw_batch* b);

/* Replace the wall layer (tile_map[WALL,:,:], single_room.jl:55-60) shared by every env of the
 * batch.  wall: [width_tu][height_tu] bytes, Julia column-major (i fastest), nonzero = wall.
 * Does not re-render; follow with rcw_reset or rcw_render. */
int32_t rcw_set_wall_map(rcw_batch* b, const uint8_t* wall);

/* tile_map[layer, :, :] = tiles for object `layer` (1-based like the reference's constants): 1 = WALL — the same as
 * rcw_set_wall_map — or 3 .. num_object_layers, an extra object layer shared by the batch.  Layer 2 (GOAL) is not a
 * map: it holds one tile per env, the goal position (rcw_reset / rcw_set_state).  tiles: [width_tu][height_tu] bytes,
 * Julia column-major (i fastest), nonzero = object present.  Not combinable with per-env wall layers.  Does not
 * re-render; follow with rcw_reset or rcw_render. */
int32_t rcw_set_layer(rcw_batch* b, int32_t layer, const uint8_t* tiles);

/* One wall layer per env: walls is [num_envs][width_tu][height_tu] bytes (each env as above).
 * Each env's layer is staged into shared memory by its own TMA bulk copy.  rcw_set_wall_map
 * switches back to a layer shared by the batch.  Does not re-render. */
int32_t rcw_set_wall_maps(rcw_batch* b, const uint8_t* walls);

/* ---- the reference's generic functions -------------------------------------------------- */

/* reset!(env) (single_room.jl:110-137,326-331) for the envs whose mask byte is nonzero
 * (mask NULL => all).  goal_ij/player_ij: [num_envs][2] int32 1-based tiles, dir_au:
 * [num_envs] int32; the player is placed at the tile centre (i-0.5, j-0.5).  All three NULL =>
 * the layout is drawn on the device (uniform interior goal, uniform empty player tile by
 * rejection, uniform direction — same draw order as the reference).  Sets reward=0,
 * done=false, then casts and renders the envs that were reset (the observations of the others still
 * belong to their unchanged state).  Does not block for pageable host arrays (they are staged before the call
 * returns); pinned / registered host arrays are waited for, so the caller may reuse them at once either way. */
int32_t rcw_reset(rcw_batch* b, const int32_t* goal_ij, const int32_t* player_ij,
                  const int32_t* dir_au, const uint8_t* mask);

/* act!(env, action) (single_room.jl:139-191,333-340) for every env: actions[e] in 1..4
 * (1 forward, 2 backward, 3 turn left, 4 turn right; single_room.jl:486).  `actions` may be a
 * host pointer or a device pointer.  A host array holding a value outside 1..4 => RCW_EACTION
 * and nothing is enqueued; for a device array the offending env is left untouched and the
 * error is reported by the next blocking call. */
int32_t rcw_step(rcw_batch* b, const uint8_t* actions);

/* rcw_step whose rewards and terminations also land in host memory without a copy (rcw_config.result_ring = D >= 1).
 * *ticket numbers the steps enqueued this way (0, 1, 2, ...).  Only enqueues; on an error (RCW_EACTION for a host
 * array, ...) nothing is enqueued and no ticket is consumed. */
int32_t rcw_step_async(rcw_batch* b, const uint8_t* actions, int64_t* ticket);

/* Blocks until the step of `ticket` has finished (later steps may still be running), then returns pointers to its
 * reward [num_envs] f32 and done [num_envs] u8 inside the pinned result ring — what rcw_get_state would have
 * returned right after that step (terminal reward / done survive the same-step auto-reset).  The memory is
 * read-only for the caller and stays valid until D further rcw_step_async calls have been made, when the slot is
 * written again; a ticket older than that => RCW_EINVAL.  An env whose device-side action was outside 1..4 keeps
 * its previous reward / done; that error is reported by the next blocking call other than rcw_wait.
 * Either pointer may be NULL. */
int32_t rcw_wait(rcw_batch* b, int64_t ticket, const float** reward, const uint8_t** done);

/* rcw_step for the envs [env0, env0 + n) only; the other envs keep their state, reward and done.
 * actions: [n] (actions[k] belongs to env env0 + k), host or device pointer, 1..4 as in rcw_step.
 * With an observation window (rcw_config.obs_window_envs = K) n must not exceed K; the range's
 * observations are then in slots (env0 + k) mod K — this is how a learner walks a batch whose
 * observations do not fit in HBM: step a window, consume it, step the next one. */
int32_t rcw_step_range(rcw_batch* b, const uint8_t* actions, int64_t env0, int64_t n);

/* n_steps of rcw_step with a uniform random policy drawn on the device (Philox keyed by
 * seed / global env id / step index); the benchmark path.  A call of two or more steps may overlap the launches of
 * consecutive steps on a second, internal stream (two half-batches, or — with rcw_config.top_view — step kernels and
 * top view kernels as a two-stage pipeline); the internal stream is joined before the call returns, so everything
 * the caller enqueues on rcw_stream afterwards is ordered behind all n_steps. */
int32_t rcw_step_random(rcw_batch* b, int32_t n_steps);

/* n_steps of rcw_step driven by an action tape: actions [n_steps][num_envs] (host or device pointer), step s takes row
 * s.  For action sequences known in advance — replays, evaluation of open-loop plans, action repeat.  A host tape is
 * validated as a whole before anything is enqueued (RCW_EACTION, the reference's @assert); a device tape like a device
 * array of rcw_step.  Like rcw_step_random, a multi-step call lets the launches of consecutive steps overlap (the
 * batch runs as two half-batches on two streams, joined before the call returns its stream to the caller), which
 * single rcw_step calls cannot do because the caller may order work of its own between them.  The observation
 * buffer holds the last step's frames (all of them with a frame ring). */
int32_t rcw_step_tape(rcw_batch* b, const uint8_t* actions, int32_t n_steps);

/* cast_rays!(world) + update_camera_view!(env) (single_room.jl:195-231,374-444) from the
 * current state, without acting. */
int32_t rcw_render(rcw_batch* b);

/* ---- state access (parity injection, checkpoint/resume) -------------------------------- */

/* Any pointer may be NULL (skipped).  pos_xy [num_envs][2] f32 world units, dir_au [num_envs],
 * goal_ij [num_envs][2] 1-based, reward [num_envs] f32, done [num_envs] u8.
 * rcw_set_state does not re-render; follow with rcw_render. */
int32_t rcw_get_state(rcw_batch* b, float* pos_xy, int32_t* dir_au, int32_t* goal_ij,
                      float* reward, uint8_t* done);
int32_t rcw_set_state(rcw_batch* b, const float* pos_xy, const int32_t* dir_au,
                      const int32_t* goal_ij, const float* reward, const uint8_t* done);

/* Exact snapshot of the dynamic state of a batch (checkpoint / resume): positions, directions, goals,
 * rewards, terminations, the per-env Philox episode counters, the running episode returns / lengths, the
 * episode totals and the step index that keys the random policy.  A handle created with the same rcw_config
 * that loads the snapshot continues bit-identically to the run that saved it (tested); the Philox key
 * (seed) and env_id_offset of the saved run are restored with it.  Wall layers and the rest of the
 * configuration are not part of the snapshot: the caller supplies them as it did originally.
 * rcw_checkpoint_size: bytes needed; rcw_save_checkpoint blocks and fills `host`; rcw_load_checkpoint
 * validates the header against the handle (RCW_ESIZE on a mismatch), restores the state and re-renders. */
int32_t rcw_checkpoint_size(rcw_batch* b, size_t* bytes);
int32_t rcw_save_checkpoint(rcw_batch* b, void* host, size_t bytes);
int32_t rcw_load_checkpoint(rcw_batch* b, const void* host, size_t bytes);

/* world.ray_stop_position_tu / ray_hit_dimension / ray_distance_wu / ray_directions_wu
 * (single_room.jl:29-31,39) for envs [env0, env0+n): hit_ij [n][num_rays][2] 1-based,
 * hit_dim [n][num_rays], dist [n][num_rays], ray_dir [n][num_rays][2].  Debug / parity only:
 * runs the ray-cast kernel in dump mode from the current state. */
int32_t rcw_get_rays(rcw_batch* b, int64_t env0, int64_t n, int32_t* hit_ij, int32_t* hit_dim,
                     float* dist, float* ray_dir);

/* ---- observations (RLBase.state, single_room.jl:576) ------------------------------------ */

/* Borrowed device pointer to the whole observation buffer (layout: rcw_obs_format), valid
 * until the next rcw_step* / rcw_reset / rcw_render / rcw_destroy — the same aliasing rule as
 * the reference, whose `state` returns the live camera_view array.
 * env_stride_bytes: distance between consecutive envs (see rcw_obs_layout).
 * total_bytes covers obs_window_envs env slots when a window is configured, num_envs otherwise. */
int32_t rcw_obs_device_ptr(rcw_batch* b, void** dptr, size_t* total_bytes, size_t* env_stride_bytes);

/* Device layout of the observation buffer.  (RCW_OBS_COLUMNS: one env = num_rays uint32 words, column_bytes =
 * column_stride_bytes = bytes_per_pixel = 4.)  One env = num_rays columns; one column = height_px
 * pixels (column_bytes) followed by padding up to column_stride_bytes, a multiple of 32, so that
 * every column starts on a 32-byte sector and the renderer only ever writes whole sectors
 * (column_stride_bytes == column_bytes whenever column_bytes is a multiple of 32, e.g. the default
 * 256 px).  env_stride_bytes is a multiple of 128.  rcw_copy_obs removes the padding. */
int32_t rcw_obs_layout(rcw_batch* b, size_t* env_stride_bytes, size_t* column_stride_bytes,
                       size_t* column_bytes, int32_t* bytes_per_pixel);

/* Frame ring of the observation buffer (rcw_config.frame_stack): number of ring positions K, the position
 * the newest frame is in, and the distance in bytes between consecutive positions of one env (the frame with
 * age a, 0 = newest, is at position (newest - a) mod K; env_stride_bytes of rcw_obs_layout spans all K). */
int32_t rcw_obs_frames(rcw_batch* b, int32_t* frame_stack, int32_t* newest, size_t* frame_stride_bytes);

/* rcw_copy_obs for the frame of the given age (0 = newest ... frame_stack - 1 = oldest). */
int32_t rcw_copy_obs_frame(rcw_batch* b, int64_t env0, int64_t n, int32_t age, void* host);

/* Blocking copy of the observations of envs [env0, env0+n) to host memory, densely packed
 * (n * num_rays * height_px * bytes_per_pixel).  With an observation window n must not exceed it and
 * the copy returns what the slots (env0 + k) mod K hold — the caller knows which envs it rendered last. */
int32_t rcw_copy_obs(rcw_batch* b, int64_t env0, int64_t n, void* host);

/* Rasterise camera views stored as RCW_OBS_COLUMNS words: the second half of update_camera_view!
 * (single_room.jl:413-441) as a pure store stream.  columns: DEVICE pointer to n envs' words,
 * columns_env_stride_bytes apart (0 => dense, num_rays * 4) — the handle's own observations
 * (rcw_obs_device_ptr of a RCW_OBS_COLUMNS handle) or records a learner kept in its replay buffer; the handle only
 * supplies the geometry (num_rays, height_px) and the palette, so any handle of that geometry can expand them.
 * dst: DEVICE pointer to n images in pixel_format (RGB8 / XRGB32 / GRAY8) laid out as rcw_expanded_layout says
 * (columns pitched to 32 bytes, envs to 128; dense whenever height_px * bytes_per_pixel is a multiple of 32 and
 * num_rays * that a multiple of 128, e.g. the default camera).  Only enqueues (on the handle's stream).
 * The words are validated on the device: a palette index outside RCW_COLOR_WALL_1..RCW_COLOR_GOAL_2 or a pad
 * above height_px / 2 (a stale or uninitialised replay row) is clamped into range — nothing is written outside
 * the word's own column — and the next blocking call on the handle returns RCW_EINVAL once. */
int32_t rcw_expand_columns(rcw_batch* b, const uint32_t* columns, size_t columns_env_stride_bytes, int64_t n,
                           int32_t pixel_format, void* dst);
int32_t rcw_expanded_layout(rcw_batch* b, int32_t pixel_format, size_t* env_stride_bytes,
                            size_t* column_stride_bytes, size_t* column_bytes);

/* ---- top view (single_room.jl:342-372, 446-483) ------------------------ */

/* update_top_view!(env) from the current state of every env: tile grid with borders, the num_rays ray
 * segments, the player's circle, as the reference's top_view::Array{UInt32}(height_tu * pu_per_tu,
 * width_tu * pu_per_tu) (single_room.jl:302).  Runs automatically after every step / reset / render
 * when rcw_config.top_view = 1.  The shapes are SimpleDraw.jl 0.3's (not vendored): Bresenham line,
 * midpoint circle, clipped to the image — unpinned against Julia like the DDA. */
int32_t rcw_render_top_view(rcw_batch* b);

/* Borrowed device pointer to the top views: uint32 [env slots][width_tu * pu columns][height_tu * pu rows],
 * the pixel row fastest like the reference's column-major array; env_stride_bytes between envs (a multiple
 * of 128).  NULL until the first top view was drawn.  Same aliasing rule and window as the observations. */
int32_t rcw_top_view_device_ptr(rcw_batch* b, void** dptr, size_t* total_bytes, size_t* env_stride_bytes);

/* Blocking copy of the top views of envs [env0, env0+n) to host memory, densely packed uint32. */
int32_t rcw_copy_top_view(rcw_batch* b, int64_t env0, int64_t n, void* host);

/* ---- bookkeeping ------------------------------------------------------------------------ */

/* Totals over finished episodes since creation (or the last call with reset_counters != 0). */
int32_t rcw_episode_stats(rcw_batch* b, int64_t* episodes, double* sum_return,
                          int64_t* sum_length, int32_t reset_counters);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches claim). */
int32_t rcw_launch_count(rcw_batch* b, int64_t* launches);

/* The CUDA stream of the handle (a cudaStream_t) so a host can record events on it. */
int32_t rcw_stream(rcw_batch* b, void** stream);

int32_t rcw_sync(rcw_batch* b);
const char* rcw_last_error(void);

/* ---- one process, several GPUs (envs are independent, single_room.jl:241-256) ----------------
 * The batch is cut into contiguous blocks of global env ids, one handle (device, stream, observation buffer) per
 * block.  Global env ids key the Philox streams, so every trajectory is bit-identical whatever the number of shards
 * (and to the one-process-per-GPU layout, which is the same handles in different processes).  rcw_step* only
 * enqueue, so one host thread keeps every GPU busy through these calls; alternatively each handle may be driven
 * from its own host thread.  There is no collective on the step path; the only cross-shard quantity is the
 * episode totals, three scalars summed on the host. */

/* Block of shard `shard` of `n_shards`: the first total_envs % n_shards shards own one env more. */
int32_t rcw_shard_envs(int64_t total_envs, int32_t n_shards, int32_t shard, int64_t* offset, int64_t* count);

/* rcw_create for every shard: cfg->num_envs is the TOTAL, cfg->env_id_offset the global id of its first env;
 * shard k runs on devices[k] (devices NULL: device k).  handles: [n_shards], filled on success; on failure
 * nothing is left allocated and every entry is NULL. */
int32_t rcw_create_sharded(const rcw_config* cfg, const float* directions_wu, const int32_t* devices, int32_t n_shards,
                           rcw_batch** handles);
int32_t rcw_destroy_sharded(rcw_batch* const* handles, int32_t n_shards);

/* rcw_step on every shard; actions: HOST array of the whole batch, [total envs] in global env order.  A value
 * outside 1..4 anywhere => RCW_EACTION and nothing is enqueued on any shard (the reference's @assert). */
int32_t rcw_step_sharded(rcw_batch* const* handles, int32_t n_shards, const uint8_t* actions);
int32_t rcw_step_random_sharded(rcw_batch* const* handles, int32_t n_shards, int32_t n_steps);
int32_t rcw_sync_sharded(rcw_batch* const* handles, int32_t n_shards);

/* rcw_episode_stats summed over the shards (in handle order, so the double sum is reproducible). */
int32_t rcw_reduce_episode_stats(rcw_batch* const* handles, int32_t n_shards, int64_t* episodes, double* sum_return,
                                 int64_t* sum_length, int32_t reset_counters);

#ifdef __cplusplus
}
#endif
#endif /* RCW_B200_H */
