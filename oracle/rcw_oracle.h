/*
 * rcw_oracle.h — CPU oracle for the SingleRoom hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (raycastworlds.jl_b200/, librcw_b200.so) never links,
 * imports or calls it.
 *
 * PARITY UNPINNED: the reference (pure Julia) cannot run in this image and its DDA lives in
 * the un-vendored package RayCaster.jl 0.1 (Project.toml:10,17; call site
 * src/single_room.jl:223).  Part A below restates the reference line by line; the DDA follows
 * the reconstructed contract of SURVEY.md §8(a) a10 with the open decisions D1/D2 exposed as
 * switches.  Part B is the CPU statement of the *new* batched semantics (Philox draws,
 * auto-reset) that the reference does not have.
 */
#ifndef RCW_ORACLE_H
#define RCW_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_EXTRA_LAYERS 4

typedef struct orc_config {
    int32_t H, W;          /* height_tile_map_tu, width_tile_map_tu  (single_room.jl:44-45) */
    int32_t N;             /* num_directions                         (single_room.jl:46)    */
    int32_t R;             /* num_rays                               (single_room.jl:52)    */
    int32_t P;             /* height_camera_view_pu                  (single_room.jl:271)   */
    float radius;          /* player_radius_wu                       (single_room.jl:47)    */
    float incr;            /* position_increment_wu                  (single_room.jl:48)    */
    float sfov;            /* semi_field_of_view_wu                  (single_room.jl:51)    */
    float cam_h;           /* camera_height_tile_wu                  (single_room.jl:270)   */
    float goal_reward;     /* goal_reward                            (single_room.jl:82)    */
    uint32_t palette[6];   /* ceiling, floor, wall1, wall2, goal1, goal2 (single_room.jl:291-296) */
    int32_t tie_le;        /* D1: 1 => tx <= ty advances dimension 1 */
    int32_t dist_post;     /* D2: 1 => distance = side - delta after the loop */
    int32_t pu_per_tu;     /* pu_per_tu of the top view                (single_room.jl:269)  */
    uint32_t top_palette[6]; /* tile_map_colors (wall, goal, empty :288), tile border (:364-367), ray_color (:289),
                                player_color (:290) */
    /* NUM_OBJECTS > 2 (single_room.jl:16-18).  Object k = 3 + index lives in its own layer of
     * tile_map; it stops rays like every object (:209 any over the layers), is painted with its own colour pair
     * (:417-429 generalised to the first object on the hit tile), shown in the top view by findfirst (:355-360),
     * keeps the player from being placed on it (utils.jl:27), and either blocks a move like WALL or ends the
     * episode with a reward like GOAL (:162-176). */
    int32_t num_layers;                          /* NUM_OBJECTS, 2..6 (default 2) */
    int32_t layer_kind[ORC_MAX_EXTRA_LAYERS];    /* 0: blocking like WALL, 1: terminal like GOAL */
    float layer_reward[ORC_MAX_EXTRA_LAYERS];    /* reward of a terminal layer */
    uint32_t layer_palette[2 * ORC_MAX_EXTRA_LAYERS]; /* camera colours [k][hit across dimension 1, 2] */
    uint32_t layer_top_color[ORC_MAX_EXTRA_LAYERS];   /* tile_map_colors[3 + k] */
} orc_config;

typedef struct orc_world orc_world;

/* ---- Part A: restatement of the reference ------------------------------------------- */
void orc_config_default(orc_config* c);
orc_world* orc_create(const orc_config* c, const float* directions /* [N][2] or NULL */);
void orc_destroy(orc_world* w);

void orc_directions(int32_t N, float* out /* [N][2] */);

/* direct field access, like mutating the Julia struct */
void orc_set_wall_map(orc_world* w, const uint8_t* wall /* [W][H], i fastest */);
int32_t orc_set_layer(orc_world* w, int32_t layer /* 1 = WALL, 3.. = extra objects */, const uint8_t* tiles /* [W][H] */);
void orc_set_state(orc_world* w, float x, float y, int32_t au, int32_t gi, int32_t gj,
                   float reward, int32_t done);
void orc_get_state(const orc_world* w, float* xy, int32_t* au, int32_t* goal_ij, float* reward,
                   int32_t* done);
/* place per reset!: goal tile, player at centre of tile, direction; reward 0, done false */
void orc_reset_to(orc_world* w, int32_t gi, int32_t gj, int32_t pi, int32_t pj, int32_t au);

int32_t orc_is_player_colliding(const orc_world* w, int32_t layer /*1 wall, 2 goal*/, float x,
                                float y);
int32_t orc_act(orc_world* w, int32_t action);   /* returns 0, or -2 for an invalid action */
void orc_cast_rays(orc_world* w);
void orc_update_camera_view(orc_world* w);
void orc_camera_columns(const orc_world* w, uint32_t* out);   /* [R] pad | palette index << 16, column order */
int32_t orc_step(orc_world* w, int32_t action);  /* act -> cast_rays -> update_camera_view */
/* update_top_view!(env) (single_room.jl:446-483) from the rays of the last orc_cast_rays.  The drawing
 * primitives come from the un-vendored package SimpleDraw.jl 0.3 (Project.toml) and are restated from their
 * published algorithms: UNPINNED like the DDA (see rcw_oracle.c). */
void orc_update_top_view(orc_world* w);
const uint32_t* orc_top_view(const orc_world* w); /* [W*pu columns][H*pu rows], row fastest (Array{UInt32}(H*pu, W*pu)) */

/* one ray, exposed for unit tests: returns hit tile (1-based), dim, dist */
void orc_cast_ray(const orc_world* w, float x, float y, float dx, float dy, int32_t* i_hit,
                  int32_t* j_hit, int32_t* dim, float* dist);

const int32_t* orc_ray_stop(const orc_world* w);   /* [R][2] (i, j) 1-based */
const int32_t* orc_ray_dim(const orc_world* w);    /* [R] */
const float* orc_ray_dist(const orc_world* w);     /* [R] */
const float* orc_ray_dir(const orc_world* w);      /* [R][2] */
const uint32_t* orc_camera_view(const orc_world* w); /* [R columns][P rows], row fastest */
void orc_wall_heights(const orc_world* w, int32_t* height_line_pu /* [R], per ray */);
void orc_obs_rgb8(const orc_world* w, uint8_t* out /* [R][P][3] */);

/* ---- Part B: batched semantics of the new engine (no reference counterpart) ----------- */
/* the camera view in the engine's byte formats (fmt 2 = RGB8 [R][P][3], 3 = GRAY8 [R][P]), written directly */
void orc_update_camera_view_bytes(orc_world* w, int32_t fmt);
const uint8_t* orc_frame_bytes(const orc_world* w);
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* layout draw of episode `episode` of global env `env_id` */
void orc_draw_layout(const orc_world* w, uint64_t seed, uint64_t env_id, uint32_t episode,
                     int32_t* goal_ij, int32_t* player_ij, int32_t* au);
/* action of global env `env_id` at global step `step` under the random policy (1..4) */
int32_t orc_draw_action(uint64_t seed, uint64_t env_id, uint64_t step);

typedef struct orc_batch orc_batch;
orc_batch* orc_batch_create(const orc_config* c, const float* directions, int64_t num_envs,
                            int64_t env_id_offset, uint64_t seed, int32_t auto_reset);
void orc_batch_destroy(orc_batch* b);
orc_world* orc_batch_world(orc_batch* b, int64_t e);
void orc_batch_reset(orc_batch* b);                       /* Philox layouts, episode += 1 */
/* one step of every env; actions NULL => random policy.  threads >= 1 (pthreads over envs). */
int32_t orc_batch_step(orc_batch* b, const uint8_t* actions, int32_t threads);
/* n random-policy steps with `threads` pthreads, each env stepped independently
 * (env-major loop: the analogue of Threads.@threads over envs).  render: 0 skips the camera view, 1 = the
 * reference's UInt32 pixels, 2 = RGB8 bytes, 3 = GRAY8 bytes written directly (orc_update_camera_view_bytes). */
void orc_batch_rollout(orc_batch* b, int32_t n_steps, int32_t threads, int32_t render);
void orc_batch_episode_stats(const orc_batch* b, int64_t* episodes, double* sum_return,
                             int64_t* sum_length);
void orc_batch_get_reward_done(const orc_batch* b, float* reward, uint8_t* done);
uint64_t orc_batch_step_index(const orc_batch* b);

#ifdef __cplusplus
}
#endif
#endif
