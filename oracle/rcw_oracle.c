/*
 * rcw_oracle.c — CPU oracle for the SingleRoom hot path.  TEST INFRASTRUCTURE ONLY
 * (see rcw_oracle.h for who may load it).  PARITY UNPINNED versus Julia: see the header.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
 * All world arithmetic is IEEE binary32 with one rounding per operation (x86-64 SSE2, no FMA
 * contraction), which is what Julia emits for Float32 `+ - * /` and `sqrt`.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 */
#include "rcw_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

struct orc_world {
    orc_config cfg;
    uint8_t* wall;       /* tile_map[WALL, :, :]  idx = (i-1) + H*(j-1)   (single_room.jl:55-60) */
    uint8_t* extra[ORC_MAX_EXTRA_LAYERS]; /* tile_map[3 + k, :, :]: object layers beyond the reference's NUM_OBJECTS = 2
                                             (single_room.jl:16-18), same indexing; NULL above cfg.num_layers */
    float* directions;   /* directions_wu [N][2]                        (single_room.jl:65-69) */
    float pos[2];        /* player_position_wu */
    int32_t au;          /* player_direction_au */
    int32_t goal[2];     /* goal_position (tile_map[GOAL, gi, gj] is the only set bit of that layer) */
    float reward;
    int32_t done;
    int32_t* ray_stop;   /* ray_stop_position_tu, here [R][2] */
    int32_t* ray_dim;    /* ray_hit_dimension */
    float* ray_dist;     /* ray_distance_wu */
    float* ray_dir;      /* ray_directions_wu [R][2] */
    uint32_t* camera;    /* camera_view [R columns][P rows], row fastest (Array{UInt32}(P,R)) */
    uint32_t* top;       /* top_view [W*pu columns][H*pu rows], row fastest (Array{UInt32}(H*pu, W*pu), :302) */
    uint8_t* frame8;     /* Part B: the camera view written directly as RGB8 / GRAY8 bytes (bench arm), lazily allocated */
};

/* ------------------------------------------------------------------------------------- */
/* Part A — the reference                                                                */
/* ------------------------------------------------------------------------------------- */

/* defaults: single_room.jl:43-52, 258-272, 291-296 */
void orc_config_default(orc_config* c) {
    memset(c, 0, sizeof(*c));
    c->H = 8;
    c->W = 16;
    c->N = 128;
    c->R = 512;
    c->P = 256;
    c->radius = (float)(1.0 / 8.0);
    c->incr = (float)(1.0 / 8.0);
    c->sfov = (float)(2.0 / 3.0);
    c->cam_h = 1.0f;
    c->goal_reward = 1.0f;
    c->palette[0] = 0x00FFFFFFu; /* ceiling_color  :292 */
    c->palette[1] = 0x00404040u; /* floor_color    :291 */
    c->palette[2] = 0x00808080u; /* wall_dim_1     :293 */
    c->palette[3] = 0x00c0c0c0u; /* wall_dim_2     :294 */
    c->palette[4] = 0x00800000u; /* goal_dim_1     :295 */
    c->palette[5] = 0x00c00000u; /* goal_dim_2     :296 */
    c->pu_per_tu = 32;               /* :269 */
    c->top_palette[0] = 0x00FFFFFFu; /* tile_map_colors[WALL]  :288 */
    c->top_palette[1] = 0x00FF0000u; /* tile_map_colors[GOAL]  :288 */
    c->top_palette[2] = 0x00000000u; /* tile_map_colors[end] (no object) :288 */
    c->top_palette[3] = 0x00ccccccu; /* tile border            :365-368 */
    c->top_palette[4] = 0x00808080u; /* ray_color              :289 */
    c->top_palette[5] = 0x00c0c0c0u; /* player_color           :290 */
    c->num_layers = 2;               /* NUM_OBJECTS = 2: WALL, GOAL (:16-18) */
}

/* single_room.jl:65-69 — theta in Float64, cos/sin in Float64, convert to T */
void orc_directions(int32_t N, float* out) {
    const double pi = 3.14159265358979323846;
    for (int32_t i = 1; i <= N; ++i) {
        double theta = (double)(i - 1) * 2 * pi / (double)N; /* (i - 1) * 2 * pi / num_directions, left to right */
        out[2 * (i - 1) + 0] = (float)cos(theta);
        out[2 * (i - 1) + 1] = (float)sin(theta);
    }
}

orc_world* orc_create(const orc_config* c, const float* directions) {
    orc_world* w = (orc_world*)calloc(1, sizeof(orc_world));
    if (!w) return NULL;
    w->cfg = *c;
    const int H = c->H, W = c->W, N = c->N, R = c->R, P = c->P;
    w->wall = (uint8_t*)calloc((size_t)H * W, 1);
    if (w->cfg.num_layers < 2) w->cfg.num_layers = 2;
    if (w->cfg.num_layers > 2 + ORC_MAX_EXTRA_LAYERS) w->cfg.num_layers = 2 + ORC_MAX_EXTRA_LAYERS;
    for (int k = 0; k < w->cfg.num_layers - 2; ++k) w->extra[k] = (uint8_t*)calloc((size_t)H * W, 1);
    w->directions = (float*)malloc(sizeof(float) * 2 * (size_t)N);
    w->ray_stop = (int32_t*)calloc((size_t)R * 2, sizeof(int32_t));
    w->ray_dim = (int32_t*)calloc((size_t)R, sizeof(int32_t));
    w->ray_dist = (float*)calloc((size_t)R, sizeof(float));
    w->ray_dir = (float*)calloc((size_t)R * 2, sizeof(float));
    w->camera = (uint32_t*)calloc((size_t)R * P, sizeof(uint32_t));
    /* border walls: single_room.jl:57-60 */
    for (int i = 1; i <= H; ++i) {
        w->wall[(i - 1) + H * 0] = 1;
        w->wall[(i - 1) + H * (W - 1)] = 1;
    }
    for (int j = 1; j <= W; ++j) {
        w->wall[0 + H * (j - 1)] = 1;
        w->wall[(H - 1) + H * (j - 1)] = 1;
    }
    if (directions)
        memcpy(w->directions, directions, sizeof(float) * 2 * (size_t)N);
    else
        orc_directions(N, w->directions);
    /* a defined initial state (the reference draws one with its RNG, :62-74,105) */
    w->goal[0] = 2;
    w->goal[1] = 2;
    w->pos[0] = (float)(H - 1) - 0.5f;
    w->pos[1] = (float)(W - 1) - 0.5f;
    w->au = 0;
    return w;
}

void orc_destroy(orc_world* w) {
    if (!w) return;
    free(w->wall);
    for (int k = 0; k < ORC_MAX_EXTRA_LAYERS; ++k) free(w->extra[k]);
    free(w->directions);
    free(w->ray_stop);
    free(w->ray_dim);
    free(w->ray_dist);
    free(w->ray_dir);
    free(w->camera);
    free(w->top);
    free(w->frame8);
    free(w);
}

void orc_set_wall_map(orc_world* w, const uint8_t* wall) {
    for (int k = 0; k < w->cfg.H * w->cfg.W; ++k) w->wall[k] = wall[k] ? 1 : 0;
}

/* tile_map[layer, :, :] = tiles for layer 1 (WALL) or an extra object layer 3..num_layers; the GOAL layer (2) holds
 * exactly one tile, the goal position, and is moved with orc_set_state / orc_reset_to. */
int32_t orc_set_layer(orc_world* w, int32_t layer, const uint8_t* tiles) {
    if (layer == 1) {
        orc_set_wall_map(w, tiles);
        return 0;
    }
    if (layer < 3 || layer > w->cfg.num_layers) return -1;
    for (int k = 0; k < w->cfg.H * w->cfg.W; ++k) w->extra[layer - 3][k] = tiles[k] ? 1 : 0;
    return 0;
}

void orc_set_state(orc_world* w, float x, float y, int32_t au, int32_t gi, int32_t gj,
                   float reward, int32_t done) {
    w->pos[0] = x;
    w->pos[1] = y;
    w->au = au;
    w->goal[0] = gi;
    w->goal[1] = gj;
    w->reward = reward;
    w->done = done;
}

void orc_get_state(const orc_world* w, float* xy, int32_t* au, int32_t* goal_ij, float* reward,
                   int32_t* done) {
    if (xy) {
        xy[0] = w->pos[0];
        xy[1] = w->pos[1];
    }
    if (au) *au = w->au;
    if (goal_ij) {
        goal_ij[0] = w->goal[0];
        goal_ij[1] = w->goal[1];
    }
    if (reward) *reward = w->reward;
    if (done) *done = w->done;
}

/* single_room.jl:118-132 with the RNG draws supplied by the caller:
 * goal tile set, player at the tile centre convert(T, i - 0.5), direction, reward 0, done false */
void orc_reset_to(orc_world* w, int32_t gi, int32_t gj, int32_t pi, int32_t pj, int32_t au) {
    w->goal[0] = gi;
    w->goal[1] = gj;
    w->pos[0] = (float)((double)pi - 0.5);
    w->pos[1] = (float)((double)pj - 0.5);
    w->au = au;
    w->reward = 0.0f;
    w->done = 0;
}

/* tile_map[layer, i, j] with the new engine's rule for F6: outside the map = empty */
static int layer_at(const orc_world* w, int layer, int i, int j) {
    if (i < 1 || i > w->cfg.H || j < 1 || j > w->cfg.W) return 0;
    if (layer == 1) return w->wall[(i - 1) + w->cfg.H * (j - 1)];
    if (layer == 2) return (i == w->goal[0] && j == w->goal[1]);
    if (layer > w->cfg.num_layers) return 0;
    return w->extra[layer - 3][(i - 1) + w->cfg.H * (j - 1)];
}

/* findfirst(@view tile_map[:, i, j]) (single_room.jl:355): the lowest object index on the tile, 0 = none.  The camera
 * view asks `tile_map[WALL, i, j] ? wall colours : goal colours` (:417-429), which is the same question for
 * NUM_OBJECTS = 2; outside the map (open host map) counts as WALL. */
static int first_object_at(const orc_world* w, int i, int j) {
    if (i < 1 || i > w->cfg.H || j < 1 || j > w->cfg.W) return 1;
    for (int layer = 1; layer <= w->cfg.num_layers; ++layer)
        if (layer_at(w, layer, i, j)) return layer;
    return 0;
}

/* camera-view colour of object `layer` hit across dimension `dim` (:417-429, palette :293-296) */
static uint32_t object_color(const orc_world* w, int layer, int dim) {
    if (layer <= 1) return dim == 1 ? w->cfg.palette[2] : w->cfg.palette[3];
    if (layer == 2) return dim == 1 ? w->cfg.palette[4] : w->cfg.palette[5];
    return w->cfg.layer_palette[2 * (layer - 3) + (dim == 1 ? 0 : 1)];
}

/* utils.jl:5 — wu_to_tu(x) = floor(Int, x) + 1 */
static int wu_to_tu(float x) { return (int)floorf(x) + 1; }

/* collision_detection.jl:9-19 — clamp to the square, squared distance < r*r (strict) */
static int is_colliding(float half_side, float radius, float px, float py) {
    float qx = px < -half_side ? -half_side : (px > half_side ? half_side : px);
    float qy = py < -half_side ? -half_side : (py > half_side ? half_side : py);
    float vx = px - qx;
    float vy = py - qy;
    float s = vx * vx + vy * vy; /* sum(vec .^ 2) */
    float rr = radius * radius;
    return s < rr;
}

/* collision_detection.jl:21-42 — 3x3 tiles around the player tile, j outer, i inner */
int32_t orc_is_player_colliding(const orc_world* w, int32_t layer, float x, float y) {
    const float half = 0.5f;
    int ip = wu_to_tu(x);
    int jp = wu_to_tu(y);
    for (int j = jp - 1; j <= jp + 1; ++j) {
        for (int i = ip - 1; i <= ip + 1; ++i) {
            float tx = (float)i - half; /* i - convert(T, 0.5)  :33 */
            float ty = (float)j - half;
            if (layer_at(w, layer, i, j) && is_colliding(half, w->cfg.radius, x - tx, y - ty))
                return 1;
        }
    }
    return 0;
}

/* single_room.jl:139-191 */
int32_t orc_act(orc_world* w, int32_t action) {
    if (action < 1 || action > 4) return -2; /* @assert :140 */
    if (action <= 2) {
        const float dx = w->directions[2 * w->au + 0]; /* directions_wu[au + 1] :153 */
        const float dy = w->directions[2 * w->au + 1];
        float nx, ny;
        float sx = w->cfg.incr * dx; /* utils.jl:16-17: incr * dir, then + / - */
        float sy = w->cfg.incr * dy;
        if (action == 1) {
            nx = w->pos[0] + sx;
            ny = w->pos[1] + sy;
        } else {
            nx = w->pos[0] - sx;
            ny = w->pos[1] - sy;
        }
        int hit_goal = orc_is_player_colliding(w, 2, nx, ny); /* :162 */
        int hit_wall = orc_is_player_colliding(w, 1, nx, ny); /* :163 */
        float goal_reward = w->cfg.goal_reward;
        /* object layers beyond WALL and GOAL behave like one of the two: a terminal layer like GOAL (its own reward,
         * done, no move — checked in object order after GOAL), a blocking layer like WALL */
        for (int layer = 3; layer <= w->cfg.num_layers; ++layer) {
            if (!orc_is_player_colliding(w, layer, nx, ny)) continue;
            if (w->cfg.layer_kind[layer - 3] == 1) {
                if (!hit_goal) {
                    hit_goal = 1;
                    goal_reward = w->cfg.layer_reward[layer - 3];
                }
            } else {
                hit_wall = 1;
            }
        }
        if (hit_goal || hit_wall) {
            if (hit_goal) { /* :166-168 */
                w->reward = goal_reward;
                w->done = 1;
            } else { /* :170-171 */
                w->reward = 0.0f;
                w->done = 0;
            }
        } else { /* :174-176 */
            w->pos[0] = nx;
            w->pos[1] = ny;
            w->reward = 0.0f;
            w->done = 0;
        }
    } else {
        /* utils.jl:13-14 — floored mod */
        int N = w->cfg.N;
        int nau = (action == 3) ? (w->au + 1) % N : ((w->au - 1) % N + N) % N;
        w->au = nau;
        w->reward = 0.0f;
        w->done = 0;
    }
    return 0;
}

/* obstacle_map = any(tile_map, dims = 1)  (single_room.jl:209): wall OR goal.
 * Outside the map counts as an obstacle so an open host-supplied map cannot run away
 * (the reference would throw a BoundsError there). */
static int obstacle_at(const orc_world* w, int i, int j) {
    if (i < 1 || i > w->cfg.H || j < 1 || j > w->cfg.W) return 1;
    return first_object_at(w, i, j) != 0;
}

/* [EXT] RayCaster.cast_ray, contract reconstructed in SURVEY.md §8(a) a10 (call site
 * single_room.jl:223).  D1 = tie_le, D2 = dist_post, D4 start tile = floor + 1,
 * D5 start tile already an obstacle => (dim 0, dist 0), D6 all arithmetic in binary32. */
void orc_cast_ray(const orc_world* w, float x, float y, float dx, float dy, int32_t* i_hit,
                  int32_t* j_hit, int32_t* dim_out, float* dist_out) {
    int i = wu_to_tu(x);
    int j = wu_to_tu(y);
    float ddx = fabsf(1.0f / dx);
    float ddy = fabsf(1.0f / dy);
    int sx, sy;
    float tx, ty;
    if (dx < 0.0f) {
        sx = -1;
        tx = (x - (float)(i - 1)) * ddx;
    } else {
        sx = 1;
        tx = ((float)i - x) * ddx;
    }
    if (dy < 0.0f) {
        sy = -1;
        ty = (y - (float)(j - 1)) * ddy;
    } else {
        sy = 1;
        ty = ((float)j - y) * ddy;
    }
    int dim = 0;
    float d = 0.0f;
    while (!obstacle_at(w, i, j)) {
        int take_x = w->cfg.tie_le ? (tx <= ty) : (tx < ty);
        if (take_x) {
            d = tx;
            tx = tx + ddx;
            i += sx;
            dim = 1;
        } else {
            d = ty;
            ty = ty + ddy;
            j += sy;
            dim = 2;
        }
    }
    if (w->cfg.dist_post && dim != 0) d = (dim == 1) ? (tx - ddx) : (ty - ddy);
    *i_hit = i;
    *j_hit = j;
    *dim_out = dim;
    *dist_out = d;
}

/* single_room.jl:193-231 */
void orc_cast_rays(orc_world* w) {
    const int R = w->cfg.R;
    const float s = w->cfg.sfov;
    const float d0 = w->directions[2 * w->au + 0];
    const float d1 = w->directions[2 * w->au + 1];
    const float c0 = d1;  /* rotate_minus_90: (v[2], -v[1])  :193 */
    const float c1 = -d0;
    /* first = dir + s * cam, last = dir - s * cam  (:216-217), f32 mul then add */
    const float f0 = d0 + s * c0, f1 = d1 + s * c1;
    const float l0 = d0 - s * c0, l1 = d1 - s * c1;
    /* LinRange length: lendiv = max(len - 1, 1) [EXT Base] */
    const int lendiv = (R - 1) > 1 ? (R - 1) : 1;
    for (int i = 1; i <= R; ++i) {
        /* [EXT Base.lerpi]: t = j/d in Float64; T((1 - t)*a + t*b) evaluated in Float64 */
        double t = (double)(i - 1) / (double)lendiv;
        float u0 = (float)((1.0 - t) * (double)f0 + t * (double)l0);
        float u1 = (float)((1.0 - t) * (double)f1 + t * (double)l1);
        /* [EXT StaticArrays.normalize]: inv(norm(v)) * v, norm = sqrt(v1^2 + v2^2) in f32 */
        float n = sqrtf(u0 * u0 + u1 * u1);
        float q = 1.0f / n;
        float r0 = q * u0;
        float r1 = q * u1;
        w->ray_dir[2 * (i - 1) + 0] = r0;
        w->ray_dir[2 * (i - 1) + 1] = r1;
        int32_t ih, jh, dim;
        float dist;
        orc_cast_ray(w, w->pos[0], w->pos[1], r0, r1, &ih, &jh, &dim, &dist); /* :223 */
        w->ray_stop[2 * (i - 1) + 0] = ih;
        w->ray_stop[2 * (i - 1) + 1] = jh;
        w->ray_dim[i - 1] = dim;
        w->ray_dist[i - 1] = dist;
    }
}

/* height of the wall line of ray index i0 (0-based), single_room.jl:404-411 */
static int height_line_pu(const orc_world* w, int i0) {
    const int R = w->cfg.R, P = w->cfg.P;
    const float d0 = w->directions[2 * w->au + 0];
    const float d1 = w->directions[2 * w->au + 1];
    const float r0 = w->ray_dir[2 * i0 + 0], r1 = w->ray_dir[2 * i0 + 1];
    float dot = d0 * r0 + d1 * r1;                 /* sum(dir .* ray) */
    float proj = w->ray_dist[i0] * dot;            /* :404 */
    float num = w->cfg.cam_h * (float)R;           /* cam_h * num_rays */
    float den = (2.0f * w->cfg.sfov) * proj;       /* *(2, s, p) = (2*s)*p */
    float hl = num / den;                          /* :406 */
    if (!isfinite(hl)) return P;                   /* :409-410 */
    if (hl >= (float)P) return P;  /* floor(Int, hl) >= P: same picture as P (:433); avoids int overflow */
    int h = (int)floorf(hl);                       /* :408 */
    if (h < 0) h = 0; /* unreachable in SingleRoom (dot > 0, dist >= 0); defined here, the reference would throw */
    return h;
}

void orc_wall_heights(const orc_world* w, int32_t* out) {
    for (int i0 = 0; i0 < w->cfg.R; ++i0) out[i0] = height_line_pu(w, i0);
}

/* single_room.jl:374-444 */
void orc_update_camera_view(orc_world* w) {
    const int R = w->cfg.R, P = w->cfg.P, H = w->cfg.H;
    const uint32_t ceiling = w->cfg.palette[0], floorc = w->cfg.palette[1];
    for (int i = 1; i <= R; ++i) {
        int h = height_line_pu(w, i - 1);
        int dim = w->ray_dim[i - 1];
        int ih = w->ray_stop[2 * (i - 1) + 0], jh = w->ray_stop[2 * (i - 1) + 1];
        /* :417-428: wall colours if tile_map[WALL, i, j], else the goal's — the first object on the hit tile (outside the
         * map, for an open host map, is painted as wall) */
        const uint32_t color = object_color(w, first_object_at(w, ih, jh), dim);
        (void)H;
        int k = R - i + 1; /* :431 */
        uint32_t* col = w->camera + (size_t)(k - 1) * P;
        if (h >= P - 1) { /* :433-434 */
            for (int p = 0; p < P; ++p) col[p] = color;
        } else { /* :436-439 */
            int pad = (P - h) / 2;
            for (int p = 0; p < pad; ++p) col[p] = ceiling;
            for (int p = pad; p < P - pad; ++p) col[p] = color;
            for (int p = P - pad; p < P; ++p) col[p] = floorc;
        }
    }
}

/* The camera view before it is expanded into pixels (the batched engine's RCW_OBS_COLUMNS format, no reference
 * counterpart): per image column k = R - i + 1 (:431) the two decisions update_camera_view! takes for ray i —
 * the rows of ceiling = rows of floor, pad = (P - h) / 2 (:436), 0 when the whole column has the wall colour
 * (h >= P - 1, :433), and the colour as an index into the palette (2/3: wall hit across dimension 1 / 2, 4/5: goal;
 * :417-428).  word = pad | index << 16. */
void orc_camera_columns(const orc_world* w, uint32_t* out) {
    const int R = w->cfg.R, P = w->cfg.P, H = w->cfg.H;
    for (int i = 1; i <= R; ++i) {
        const int h = height_line_pu(w, i - 1);
        const int dim = w->ray_dim[i - 1];
        const int ih = w->ray_stop[2 * (i - 1) + 0], jh = w->ray_stop[2 * (i - 1) + 1];
        int layer = first_object_at(w, ih, jh);
        if (layer < 1) layer = 1;
        (void)H;
        const uint32_t cid = 2u * (uint32_t)layer + (dim == 1 ? 0u : 1u);   /* 2/3 wall, 4/5 goal, 6/7 object 3, ... */
        const uint32_t pad = h >= P - 1 ? 0u : (uint32_t)((P - h) / 2);
        out[R - i] = pad | (cid << 16);
    }
}

/* ---- top view (single_room.jl:342-372, 446-483) ------------------------------------------------
 * The shapes are drawn by SimpleDraw.jl 0.3 [EXT, not vendored, no Manifest]: FilledRectangle(position,
 * height, width), Line(point1, point2), Circle(position, diameter).  Restated here from the algorithms
 * that package documents — Bresenham's line over all octants and the midpoint circle — with every pixel
 * write bounds-checked.  UNPINNED: the exact error-term conventions of SimpleDraw are not checkable here. */

static int wu_to_pu(float x_wu, int pu_per_wu) { return (int)floorf(x_wu * (float)pu_per_wu) + 1; } /* utils.jl:6 */

static void put_pixel(orc_world* w, int i, int j, uint32_t color) { /* 1-based, clipped */
    const int Hp = w->cfg.H * w->cfg.pu_per_tu, Wp = w->cfg.W * w->cfg.pu_per_tu;
    if (i >= 1 && i <= Hp && j >= 1 && j <= Wp) w->top[(size_t)(i - 1) + (size_t)Hp * (j - 1)] = color;
}

/* [EXT SimpleDraw] Line: Bresenham, all octants, both end points drawn */
static void draw_line(orc_world* w, int i1, int j1, int i2, int j2, uint32_t color) {
    const int di = abs(i2 - i1), dj = -abs(j2 - j1);
    const int si = i1 < i2 ? 1 : -1, sj = j1 < j2 ? 1 : -1;
    int err = di + dj, i = i1, j = j1;
    for (;;) {
        put_pixel(w, i, j, color);
        if (i == i2 && j == j2) break;
        const int e2 = 2 * err;
        if (e2 >= dj) {
            err += dj;
            i += si;
        }
        if (e2 <= di) {
            err += di;
            j += sj;
        }
    }
}

/* [EXT SimpleDraw] Circle(position = top-left of the bounding box, odd diameter 2r + 1): midpoint circle of
 * radius r about the centre pixel, eight-way symmetric */
static void draw_circle(orc_world* w, int ic, int jc, int r, uint32_t color) {
    int a = 0, b = r, d = 1 - r;
    while (a <= b) {
        put_pixel(w, ic + a, jc + b, color);
        put_pixel(w, ic - a, jc + b, color);
        put_pixel(w, ic + a, jc - b, color);
        put_pixel(w, ic - a, jc - b, color);
        put_pixel(w, ic + b, jc + a, color);
        put_pixel(w, ic - b, jc + a, color);
        put_pixel(w, ic + b, jc - a, color);
        put_pixel(w, ic - b, jc - a, color);
        if (d < 0) {
            d += 2 * a + 3;
        } else {
            d += 2 * (a - b) + 5;
            b -= 1;
        }
        a += 1;
    }
}

/* draw_tile_map! single_room.jl:342-372 */
static void draw_tile_map(orc_world* w) {
    const int H = w->cfg.H, W = w->cfg.W, pu = w->cfg.pu_per_tu;
    const uint32_t border = w->cfg.top_palette[3];
    for (int j = 1; j <= W; ++j)
        for (int i = 1; i <= H; ++i) {
            const int it = (i - 1) * pu + 1, jt = (j - 1) * pu + 1; /* :350-351 */
            /* findfirst over the layers WALL = 1, GOAL = 2 (:355-360) */
            uint32_t color = w->cfg.top_palette[2];
            const int object = first_object_at(w, i, j);
            if (object == 1) color = w->cfg.top_palette[0];
            else if (object == 2) color = w->cfg.top_palette[1];
            else if (object >= 3) color = w->cfg.layer_top_color[object - 3];   /* tile_map_colors[object] */
            for (int b = 0; b < pu; ++b)
                for (int a = 0; a < pu; ++a) put_pixel(w, it + a, jt + b, color); /* :353,362 */
            for (int b = 0; b < pu; ++b) {
                put_pixel(w, it, jt + b, border);          /* :364 */
                put_pixel(w, it + pu - 1, jt + b, border); /* :365 */
                put_pixel(w, it + b, jt, border);          /* :366 */
                put_pixel(w, it + b, jt + pu - 1, border); /* :367 */
            }
        }
}

/* update_top_view! single_room.jl:446-483 */
void orc_update_top_view(orc_world* w) {
    const int pu = w->cfg.pu_per_tu, R = w->cfg.R;
    if (!w->top) w->top = (uint32_t*)calloc((size_t)w->cfg.H * pu * w->cfg.W * pu, sizeof(uint32_t));
    const int ip = wu_to_pu(w->pos[0], pu), jp = wu_to_pu(w->pos[1], pu); /* :469 */
    const int rp = wu_to_pu(w->cfg.radius, pu);                           /* :470 */
    draw_tile_map(w);                                                     /* :472 */
    for (int i = 0; i < R; ++i) {                                         /* :474-478 */
        const float sx = w->pos[0] + w->ray_dist[i] * w->ray_dir[2 * i + 0];
        const float sy = w->pos[1] + w->ray_dist[i] * w->ray_dir[2 * i + 1];
        draw_line(w, ip, jp, wu_to_pu(sx, pu), wu_to_pu(sy, pu), w->cfg.top_palette[4]);
    }
    /* Circle(Point(i - r, j - r), 2r + 1) (:480): centre (ip, jp), radius rp */
    draw_circle(w, ip, jp, rp, w->cfg.top_palette[5]);
}

const uint32_t* orc_top_view(const orc_world* w) { return w->top; }

/* act!(env) without the top view: single_room.jl:333-340 */
int32_t orc_step(orc_world* w, int32_t action) {
    int32_t rc = orc_act(w, action);
    if (rc) return rc;
    orc_cast_rays(w);
    orc_update_camera_view(w);
    return 0;
}

const int32_t* orc_ray_stop(const orc_world* w) { return w->ray_stop; }
const int32_t* orc_ray_dim(const orc_world* w) { return w->ray_dim; }
const float* orc_ray_dist(const orc_world* w) { return w->ray_dist; }
const float* orc_ray_dir(const orc_world* w) { return w->ray_dir; }
const uint32_t* orc_camera_view(const orc_world* w) { return w->camera; }

/* RGB8 view of the reference pixel 0x00RRGGBB: bytes R, G, B */
void orc_obs_rgb8(const orc_world* w, uint8_t* out) {
    size_t n = (size_t)w->cfg.R * w->cfg.P;
    for (size_t k = 0; k < n; ++k) {
        uint32_t c = w->camera[k];
        out[3 * k + 0] = (uint8_t)(c >> 16);
        out[3 * k + 1] = (uint8_t)(c >> 8);
        out[3 * k + 2] = (uint8_t)c;
    }
}

/* ------------------------------------------------------------------------------------- */
/* Part B — batched semantics of the new engine (DESIGN.md "Batched semantics")           */
/* ------------------------------------------------------------------------------------- */

/* update_camera_view! (single_room.jl:374-444) writing the engine's byte formats directly instead of UInt32
 * pixels, so that the CPU arm of bench.py stores the same bytes per frame as the GPU arm: fmt 2 = RGB8 (bytes
 * R, G, B of the reference pixel, 3 per pixel), fmt 3 = GRAY8 (BT.601 luma, 1 per pixel).  Same bands, same
 * column order; the result equals orc_obs_rgb8 / the luma of orc_camera_view (tested). */
static void fill_pixels(uint8_t* dst, int n_px, uint32_t color, int fmt) {
    if (fmt == 3) {
        const uint32_t y = (77u * ((color >> 16) & 255u) + 150u * ((color >> 8) & 255u) + 29u * (color & 255u) + 128u) >> 8;
        memset(dst, (int)y, (size_t)n_px);
        return;
    }
    const uint8_t r = (uint8_t)(color >> 16), g = (uint8_t)(color >> 8), b = (uint8_t)color;
    if (r == g && g == b) {
        memset(dst, r, (size_t)n_px * 3);
        return;
    }
    for (int k = 0; k < n_px; ++k) {
        dst[3 * k + 0] = r;
        dst[3 * k + 1] = g;
        dst[3 * k + 2] = b;
    }
}

void orc_update_camera_view_bytes(orc_world* w, int32_t fmt) {
    const int R = w->cfg.R, P = w->cfg.P, H = w->cfg.H;
    const int bpp = fmt == 3 ? 1 : 3;
    if (!w->frame8) w->frame8 = (uint8_t*)calloc((size_t)R * P * 3, 1);
    const uint32_t ceiling = w->cfg.palette[0], floorc = w->cfg.palette[1];
    for (int i = 1; i <= R; ++i) {
        const int h = height_line_pu(w, i - 1);
        const int dim = w->ray_dim[i - 1];
        const int ih = w->ray_stop[2 * (i - 1) + 0], jh = w->ray_stop[2 * (i - 1) + 1];
        const uint32_t color = object_color(w, first_object_at(w, ih, jh), dim);
        (void)H;
        uint8_t* col = w->frame8 + (size_t)(R - i) * P * bpp; /* k = R - i + 1 (:431) */
        if (h >= P - 1) {
            fill_pixels(col, P, color, fmt);
        } else {
            const int pad = (P - h) / 2;
            fill_pixels(col, pad, ceiling, fmt);
            fill_pixels(col + (size_t)pad * bpp, P - 2 * pad, color, fmt);
            fill_pixels(col + (size_t)(P - pad) * bpp, pad, floorc, fmt);
        }
    }
}

const uint8_t* orc_frame_bytes(const orc_world* w) { return w->frame8; }

/* Philox4x32-10, Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

#define ORC_STREAM_RESET 0x52u
#define ORC_STREAM_ACTION 0x41u

static uint32_t uniform_below(uint32_t u, uint32_t n) { return (uint32_t)(((uint64_t)u * n) >> 32); }

/* Draw order of reset! (single_room.jl:120,124,128): goal_i, goal_j, player tile (one draw per
 * try, rejection while any object on the tile, utils.jl:23-37,52-58), direction. */
void orc_draw_layout(const orc_world* w, uint64_t seed, uint64_t env_id, uint32_t episode,
                     int32_t* goal_ij, int32_t* player_ij, int32_t* au) {
    const int H = w->cfg.H, W = w->cfg.W, N = w->cfg.N;
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(uint32_t)env_id, (uint32_t)(env_id >> 32), episode, ORC_STREAM_RESET << 24};
    uint32_t u[4];
    orc_philox4x32_10(ctr, key, u);
    int gi = 2 + (int)uniform_below(u[0], (uint32_t)(H - 2)); /* rand(2 : H-1) */
    int gj = 2 + (int)uniform_below(u[1], (uint32_t)(W - 2));
    *au = (int)uniform_below(u[2], (uint32_t)N);               /* rand(0 : N-1) */
    long long max_tries = 1024LL * H * W;                      /* utils.jl:55 */
    if (max_tries > (1LL << 22)) max_tries = 1LL << 22;
    uint32_t draw = u[3];
    int pi = 1, pj = 1;
    for (long long t = 0;; ++t) {
        uint32_t lin = uniform_below(draw, (uint32_t)(H * W)); /* CartesianIndices((1:H, 1:W)), i fastest */
        pi = (int)(lin % (uint32_t)H) + 1;
        pj = (int)(lin / (uint32_t)H) + 1;
        int occupied = (pi == gi && pj == gj);      /* any(tile_map[:, pos]) with the new goal already placed (utils.jl:27) */
        for (int layer = 1; layer <= w->cfg.num_layers && !occupied; ++layer)
            if (layer != 2) occupied = layer_at(w, layer, pi, pj);
        if (!occupied || t == max_tries) break;
        /* next draw: try t+1 uses word (t % 4) of Philox call 1 + t/4 */
        if ((t & 3) == 0) {
            ctr[3] = (ORC_STREAM_RESET << 24) | (uint32_t)(1 + t / 4);
            orc_philox4x32_10(ctr, key, u);
        }
        draw = u[t & 3];
    }
    goal_ij[0] = gi;
    goal_ij[1] = gj;
    player_ij[0] = pi;
    player_ij[1] = pj;
}

int32_t orc_draw_action(uint64_t seed, uint64_t env_id, uint64_t step) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(uint32_t)env_id, (uint32_t)(env_id >> 32), (uint32_t)step,
                       (ORC_STREAM_ACTION << 24) | ((uint32_t)(step >> 32) & 0xFFFFFFu)};
    uint32_t u[4];
    orc_philox4x32_10(ctr, key, u);
    return (int32_t)(u[0] & 3u) + 1;
}

struct orc_batch {
    orc_config cfg;
    int64_t num_envs, env_id_offset;
    uint64_t seed;
    int32_t auto_reset;
    orc_world** worlds;
    uint32_t* episode;
    float* reward;     /* outputs of the last step (terminal values survive the auto-reset) */
    uint8_t* done;
    float* ep_return;
    uint32_t* ep_length;
    int64_t episodes;
    double sum_return;
    int64_t sum_length;
    uint64_t step_index;
};

orc_batch* orc_batch_create(const orc_config* c, const float* directions, int64_t num_envs,
                            int64_t env_id_offset, uint64_t seed, int32_t auto_reset) {
    orc_batch* b = (orc_batch*)calloc(1, sizeof(orc_batch));
    b->cfg = *c;
    b->num_envs = num_envs;
    b->env_id_offset = env_id_offset;
    b->seed = seed;
    b->auto_reset = auto_reset;
    b->worlds = (orc_world**)calloc((size_t)num_envs, sizeof(orc_world*));
    b->episode = (uint32_t*)calloc((size_t)num_envs, sizeof(uint32_t));
    b->reward = (float*)calloc((size_t)num_envs, sizeof(float));
    b->done = (uint8_t*)calloc((size_t)num_envs, 1);
    b->ep_return = (float*)calloc((size_t)num_envs, sizeof(float));
    b->ep_length = (uint32_t*)calloc((size_t)num_envs, sizeof(uint32_t));
    for (int64_t e = 0; e < num_envs; ++e) b->worlds[e] = orc_create(c, directions);
    orc_batch_reset(b);
    return b;
}

void orc_batch_destroy(orc_batch* b) {
    if (!b) return;
    for (int64_t e = 0; e < b->num_envs; ++e) orc_destroy(b->worlds[e]);
    free(b->worlds);
    free(b->episode);
    free(b->reward);
    free(b->done);
    free(b->ep_return);
    free(b->ep_length);
    free(b);
}

orc_world* orc_batch_world(orc_batch* b, int64_t e) { return b->worlds[e]; }

static void batch_new_episode(orc_batch* b, int64_t e) {
    orc_world* w = b->worlds[e];
    int32_t g[2], p[2], au;
    b->episode[e] += 1;
    orc_draw_layout(w, b->seed, (uint64_t)(b->env_id_offset + e), b->episode[e], g, p, &au);
    orc_reset_to(w, g[0], g[1], p[0], p[1], au);
    b->ep_return[e] = 0.0f;
    b->ep_length[e] = 0;
}

void orc_batch_reset(orc_batch* b) {
    for (int64_t e = 0; e < b->num_envs; ++e) {
        batch_new_episode(b, e);
        b->reward[e] = 0.0f;
        b->done[e] = 0;
        orc_cast_rays(b->worlds[e]);
        orc_update_camera_view(b->worlds[e]);
    }
}

typedef struct {
    int64_t episodes;
    double sum_return;
    int64_t sum_length;
} stats_acc;

/* One env-step: act, then (auto_reset) a terminated env is re-drawn inside the same step, then
 * cast + render.  reward/done keep the values act! produced. */
static int32_t batch_step_env(orc_batch* b, int64_t e, int32_t action, uint64_t step, int render,
                              stats_acc* acc) {
    orc_world* w = b->worlds[e];
    if (action == 0) action = orc_draw_action(b->seed, (uint64_t)(b->env_id_offset + e), step);
    int32_t rc = orc_act(w, action);
    if (rc) return rc;
    b->reward[e] = w->reward;
    b->done[e] = (uint8_t)w->done;
    b->ep_return[e] += w->reward;
    b->ep_length[e] += 1;
    if (w->done) {
        acc->episodes += 1;
        acc->sum_return += (double)b->ep_return[e];
        acc->sum_length += (int64_t)b->ep_length[e];
        if (b->auto_reset) {
            batch_new_episode(b, e);
        } else {
            b->ep_return[e] = 0.0f;
            b->ep_length[e] = 0;
        }
    }
    orc_cast_rays(w);
    if (render == 1) orc_update_camera_view(w);                 /* the reference's UInt32 pixels */
    else if (render >= 2) orc_update_camera_view_bytes(w, render); /* 2: RGB8, 3: GRAY8 (the engine's formats) */
    return 0;
}

typedef struct {
    orc_batch* b;
    const uint8_t* actions;
    int64_t e0, e1;
    int32_t n_steps;
    int render;
    int32_t rc;
    stats_acc acc;
} worker_arg;

static void* worker_main(void* p) {
    worker_arg* a = (worker_arg*)p;
    a->rc = 0;
    memset(&a->acc, 0, sizeof(a->acc));
    for (int64_t e = a->e0; e < a->e1; ++e) {
        for (int32_t s = 0; s < a->n_steps; ++s) {
            int32_t act = a->actions ? (int32_t)a->actions[e] : 0;
            if (a->actions && (act < 1 || act > 4)) {
                a->rc = -2;
                continue;
            }
            int32_t rc = batch_step_env(a->b, e, act, a->b->step_index + (uint64_t)s, a->render, &a->acc);
            if (rc) a->rc = rc;
        }
    }
    return NULL;
}

static int32_t run_workers(orc_batch* b, const uint8_t* actions, int32_t n_steps, int32_t threads,
                           int render) {
    if (threads < 1) threads = 1;
    if ((int64_t)threads > b->num_envs) threads = (int32_t)(b->num_envs > 0 ? b->num_envs : 1);
    worker_arg* args = (worker_arg*)calloc((size_t)threads, sizeof(worker_arg));
    pthread_t* tids = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    int64_t per = (b->num_envs + threads - 1) / threads;
    for (int32_t t = 0; t < threads; ++t) {
        args[t].b = b;
        args[t].actions = actions;
        args[t].e0 = t * per < b->num_envs ? t * per : b->num_envs;
        args[t].e1 = (t + 1) * per < b->num_envs ? (t + 1) * per : b->num_envs;
        args[t].n_steps = n_steps;
        args[t].render = render;
        if (threads == 1)
            worker_main(&args[t]);
        else
            pthread_create(&tids[t], NULL, worker_main, &args[t]);
    }
    int32_t rc = 0;
    for (int32_t t = 0; t < threads; ++t) {
        if (threads > 1) pthread_join(tids[t], NULL);
        if (args[t].rc) rc = args[t].rc;
        b->episodes += args[t].acc.episodes;
        b->sum_return += args[t].acc.sum_return;
        b->sum_length += args[t].acc.sum_length;
    }
    b->step_index += (uint64_t)n_steps;
    free(args);
    free(tids);
    return rc;
}

int32_t orc_batch_step(orc_batch* b, const uint8_t* actions, int32_t threads) {
    if (actions)
        for (int64_t e = 0; e < b->num_envs; ++e)
            if (actions[e] < 1 || actions[e] > 4) return -2; /* nothing happens, like the @assert */
    return run_workers(b, actions, 1, threads, 1);
}

void orc_batch_rollout(orc_batch* b, int32_t n_steps, int32_t threads, int32_t render) {
    run_workers(b, NULL, n_steps, threads, render);
}

void orc_batch_episode_stats(const orc_batch* b, int64_t* episodes, double* sum_return,
                             int64_t* sum_length) {
    if (episodes) *episodes = b->episodes;
    if (sum_return) *sum_return = b->sum_return;
    if (sum_length) *sum_length = b->sum_length;
}

void orc_batch_get_reward_done(const orc_batch* b, float* reward, uint8_t* done) {
    if (reward) memcpy(reward, b->reward, sizeof(float) * (size_t)b->num_envs);
    if (done) memcpy(done, b->done, (size_t)b->num_envs);
}

uint64_t orc_batch_step_index(const orc_batch* b) { return b->step_index; }
