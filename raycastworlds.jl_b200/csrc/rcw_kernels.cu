// rcw_kernels.cu — sm_100a kernels of the batched SingleRoom engine.
//
// One launch of frame_kernel<kModeStep> is one env-step of the whole batch:
//   act!(world) (+ auto-reset)  ->  cast_rays!(world)  ->  update_camera_view!(env)
// (reference: src/single_room.jl:139-191, 195-231, 374-444; collision_detection.jl:9-42).
//
// Mapping.  A work item is (env, group of 32 consecutive rays) and belongs to one warp:
// lane <-> ray for the DDA, then all 32 lanes stream the group's 32 observation columns,
// which are contiguous in memory (ray i paints column R-i+1, single_room.jl:431), with
// 16-byte stores.  The eight warps of a CTA take eight consecutive items, so a CTA writes one
// contiguous span of the observation buffer.  The path is bound by HBM writes (393 KB per
// env-step at the default resolution versus a few hundred bytes of state), so everything else
// is arranged to keep the store stream dense: the shared wall layer is staged once per CTA into
// shared memory with a TMA bulk copy, the per-(direction, ray) table {ray, |1/ray|} is read with
// one coalesced 16-byte load per lane from an L2-resident table, and act!/auto-reset are
// recomputed by every warp of an env from double-buffered state instead of synchronising.
//
// Arithmetic.  Every binary32 operation that the reference performs is written with an explicit
// round-to-nearest intrinsic (__fmul_rn, __fadd_rn, __fdiv_rn, __fsqrt_rn), which the compiler
// never contracts into FMAs, so positions, hit tiles, hit sides and distances are bit-identical
// to the CPU restatement (the file is also compiled with -fmad=false).

#include "rcw_internal.h"

namespace rcw {

__constant__ float2 c_dirs[kDirSlots][kDirSlotEntries];

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

// TMA 1-D bulk copy global -> shared (SASS: UBLKCP); completion is counted on the mbarrier.
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                              uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

// Philox4x32-10 (Salmon et al., SC'11).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

constexpr uint32_t kStreamReset = 0x52u;
constexpr uint32_t kStreamAction = 0x41u;

__device__ __forceinline__ uint32_t uniform_below(uint32_t u, uint32_t n) { return __umulhi(u, n); }

// wall layer probe, 0-based tile, caller guarantees the tile is inside the map
__device__ __forceinline__ bool wall_bit(const uint32_t* map, int wpr, int i0, int j0) {
    return (map[i0 * wpr + (j0 >> 5)] >> (j0 & 31)) & 1u;
}

// Uniform random policy: action in 1..4 for (global env id, global step index).
__device__ __forceinline__ int draw_action(uint64_t seed, uint64_t env_id, uint64_t step) {
    const uint4 u = philox4x32_10(
        make_uint4((uint32_t)env_id, (uint32_t)(env_id >> 32), (uint32_t)step,
                   (kStreamAction << 24) | ((uint32_t)(step >> 32) & 0xFFFFFFu)),
        make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return (int)(u.x & 3u) + 1;
}

// Layout of a new episode, in the draw order of reset! (single_room.jl:120,124,128):
// goal_i in 2..H-1, goal_j in 2..W-1, player tile uniform over all tiles with rejection while a
// wall or the goal is on it (utils.jl:23-37,52-58), direction in 0..N-1.  All outputs 1-based.
__device__ inline void draw_layout(const uint32_t* map, int H, int W, int wpr, int N, uint64_t seed,
                                   uint64_t env_id, uint32_t episode, int& gi, int& gj, int& pi,
                                   int& pj, int& au) {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    uint4 ctr = make_uint4((uint32_t)env_id, (uint32_t)(env_id >> 32), episode, kStreamReset << 24);
    uint4 u = philox4x32_10(ctr, key);
    gi = 2 + (int)uniform_below(u.x, (uint32_t)(H - 2));
    gj = 2 + (int)uniform_below(u.y, (uint32_t)(W - 2));
    au = (int)uniform_below(u.z, (uint32_t)N);
    long long max_tries = 1024LL * H * W;  // utils.jl:55
    if (max_tries > (1LL << 22)) max_tries = 1LL << 22;
    uint32_t draw = u.w;
    for (long long t = 0;; ++t) {
        const uint32_t lin = uniform_below(draw, (uint32_t)(H * W));  // CartesianIndices, i fastest
        pi = (int)(lin % (uint32_t)H) + 1;
        pj = (int)(lin / (uint32_t)H) + 1;
        const bool occupied = wall_bit(map, wpr, pi - 1, pj - 1) || (pi == gi && pj == gj);
        if (!occupied || t == max_tries) break;
        const int word = (int)(t & 3);
        if (word == 0) {
            ctr.w = (kStreamReset << 24) | (uint32_t)(1 + t / 4);
            u = philox4x32_10(ctr, key);
        }
        draw = word == 0 ? u.x : word == 1 ? u.y : word == 2 ? u.z : u.w;
    }
}

// collision_detection.jl:9-19,33-35: circle of `radius` at (x, y) against the unit tile (i, j)
// (1-based): clamp the offset to the square, squared distance strictly below radius^2.
__device__ __forceinline__ bool circle_hits_tile(float x, float y, int i, int j, float radius) {
    const float px = __fsub_rn(x, __fsub_rn((float)i, 0.5f));
    const float py = __fsub_rn(y, __fsub_rn((float)j, 0.5f));
    const float qx = px < -0.5f ? -0.5f : (px > 0.5f ? 0.5f : px);
    const float qy = py < -0.5f ? -0.5f : (py > 0.5f ? 0.5f : py);
    const float vx = __fsub_rn(px, qx), vy = __fsub_rn(py, qy);
    return __fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)) < __fmul_rn(radius, radius);
}

// ------------------------------------------------------------------------------------------
// pixel formats
// ------------------------------------------------------------------------------------------

template <int FMT>
struct PixelFormat;

template <>
struct PixelFormat<RCW_OBS_RGB8> {
    static constexpr int kBpp = 3;
    // byte k of a pixel of colour 0x00RRGGBB: R, G, B
    __device__ static __forceinline__ uint32_t byte_of(uint32_t c, int k) {
        return (c >> (16 - 8 * k)) & 0xFFu;
    }
    // 16 bytes of a single-colour run that starts at byte `ob` of the column
    __device__ static __forceinline__ uint4 run16(uint32_t c, int ob) {
        const int ph = ob % 3;
        const uint32_t p0 = __byte_perm(c, 0, 0x2012);  // R G B R
        const uint32_t p1 = __byte_perm(c, 0, 0x1201);  // G B R G
        const uint32_t p2 = __byte_perm(c, 0, 0x0120);  // B R G B
        const uint32_t sh = 8u * (uint32_t)ph;
        const uint32_t w0 = __funnelshift_r(p0, p1, sh);
        const uint32_t w1 = __funnelshift_r(p1, p2, sh);
        const uint32_t w2 = __funnelshift_r(p2, p0, sh);
        return make_uint4(w0, w1, w2, w0);
    }
};

template <>
struct PixelFormat<RCW_OBS_XRGB32> {
    static constexpr int kBpp = 4;
    __device__ static __forceinline__ uint32_t byte_of(uint32_t c, int k) {
        return (c >> (8 * k)) & 0xFFu;  // little-endian UInt32 0x00RRGGBB
    }
    __device__ static __forceinline__ uint4 run16(uint32_t c, int ob) {
        const uint32_t sh = 8u * (uint32_t)(ob & 3);
        const uint32_t w = __funnelshift_r(c, c, sh);
        return make_uint4(w, w, w, w);
    }
};

// One observation column: rows [0, pad) ceiling, [pad, P - pad) wall colour, [P - pad, P) floor
// (single_room.jl:433-439; a full-height column is pad = 0).  In bytes: b1 = bpp * pad,
// b2 = col_bytes - b1.
struct ColumnBands {
    int b1, b2;
    uint32_t wall, ceiling, floor;
    // 0 ceiling, 1 wall, 2 floor, 3 mixed — for the 16 bytes starting at column offset ob
    __device__ __forceinline__ int classify16(int ob) const {
        if (ob + 16 <= b1) return 0;
        if (ob >= b2) return 2;
        if (ob >= b1 && ob + 16 <= b2) return 1;
        return 3;
    }
    __device__ __forceinline__ uint32_t color_at(int ob) const {
        return ob < b1 ? ceiling : (ob < b2 ? wall : floor);
    }
};

template <int FMT>
__device__ __forceinline__ uint32_t column_byte(const ColumnBands& cb, int ob) {
    constexpr int bpp = PixelFormat<FMT>::kBpp;
    const int px = ob / bpp;
    return PixelFormat<FMT>::byte_of(cb.color_at(px * bpp), ob - px * bpp);
}

__device__ __forceinline__ void store_stream16(uint8_t* p, uint4 v) {
    __stcs(reinterpret_cast<uint4*>(p), v);
}

// ------------------------------------------------------------------------------------------
// the frame kernel
// ------------------------------------------------------------------------------------------

template <int MODE, int FMT>
__global__ void __launch_bounds__(kThreadsPerCta)
frame_kernel(const __grid_constant__ FrameParams p) {
    extern __shared__ __align__(16) uint32_t s_map[];   // bit-packed wall layer
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint2 s_col[kWarpsPerCta][32];            // per column of the warp: {pad, colour}

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    // ---- stage the wall layer: one TMA bulk copy per CTA, completion on an mbarrier ----------
    if (threadIdx.x == 0) mbar_init(&s_mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)p.map_words * 4u;
        mbar_arrive_expect_tx(&s_mbar, bytes);
        bulk_copy_g2s(s_map, p.wall_map, bytes, &s_mbar);
    }
    mbar_wait(&s_mbar, 0);

    constexpr int bpp = PixelFormat<FMT>::kBpp;
    const int H = p.H, W = p.W, wpr = p.wpr, R = p.R, P = p.P, CB = p.col_bytes;
    const uint32_t n_items = (uint32_t)(p.env_count * p.gpe);
    const uint32_t item_stride = gridDim.x * kWarpsPerCta;

    for (uint32_t item = blockIdx.x * kWarpsPerCta + warp; item < n_items; item += item_stride) {
        const uint32_t env_rel = item / (uint32_t)p.gpe;
        const int g = (int)(item - env_rel * (uint32_t)p.gpe);
        const int64_t env = p.env_first + env_rel;

        // ---- state (identical in all lanes) --------------------------------------------------
        float x = __ldg(p.in.pos_x + env);
        float y = __ldg(p.in.pos_y + env);
        int au = __ldg(p.in.dir_au + env);
        uint32_t goal = __ldg(p.in.goal + env);
        int gi = (int)(goal & 0xFFFFu), gj = (int)(goal >> 16);

        if (MODE == kModeStep) {
            uint32_t episode = __ldg(p.in.episode + env);
            const uint64_t env_id = p.env_id_offset + (uint64_t)env;
            const int a = p.actions ? (int)__ldg(p.actions + env)
                                    : draw_action(p.seed, env_id, p.step_index);
            const bool valid = (a >= 1) && (a <= 4);
            float reward = 0.0f;
            bool done = false;
            if (valid) {
                if (a <= 2) {
                    // move_forward / move_backward (utils.jl:16-17): pos +- incr * dir
                    const float2 d = p.dir_slot >= 0 ? c_dirs[p.dir_slot][au] : __ldg(p.dirs + au);
                    const float sx = __fmul_rn(p.incr, d.x), sy = __fmul_rn(p.incr, d.y);
                    const float nx = (a == 1) ? __fadd_rn(x, sx) : __fsub_rn(x, sx);
                    const float ny = (a == 1) ? __fadd_rn(y, sy) : __fsub_rn(y, sy);
                    // is_player_colliding (collision_detection.jl:21-42): lanes 0..8 probe the
                    // 3x3 tiles around wu_to_tu(candidate); tiles outside the map are empty.
                    const int ti = __float2int_rd(nx) + (lane % 3);        // = ip - 1 + di, 1-based
                    const int tj = __float2int_rd(ny) + ((lane / 3) % 3);
                    bool hit_goal = false, hit_wall = false;
                    if (lane < 9 && ti >= 1 && ti <= H && tj >= 1 && tj <= W) {
                        const bool is_goal = (ti == gi) && (tj == gj);
                        const bool is_wall = wall_bit(s_map, wpr, ti - 1, tj - 1);
                        if (is_goal || is_wall) {
                            const bool c = circle_hits_tile(nx, ny, ti, tj, p.radius);
                            hit_goal = is_goal && c;
                            hit_wall = is_wall && c;
                        }
                    }
                    const bool any_goal = __any_sync(0xFFFFFFFFu, hit_goal);
                    const bool any_wall = __any_sync(0xFFFFFFFFu, hit_wall);
                    if (any_goal) {           // single_room.jl:166-168 — reward, done, no move
                        reward = p.goal_reward;
                        done = true;
                    } else if (!any_wall) {   // :174-176
                        x = nx;
                        y = ny;
                    }
                } else {
                    // turn_left / turn_right (utils.jl:13-14): floored mod
                    au = (a == 3) ? (au + 1 == p.N ? 0 : au + 1) : (au == 0 ? p.N - 1 : au - 1);
                }
            }
            // bookkeeping of the finished step, then (auto-reset) the next episode's layout
            float ep_return = 0.0f;
            uint32_t ep_length = 0;
            const bool writer = (g == 0) && (lane == 0);
            if (writer && valid) {
                ep_return = __fadd_rn(p.ep_return[env], reward);
                ep_length = p.ep_length[env] + 1u;
            }
            if (done) {
                if (writer) {
                    atomicAdd(&p.stats->episodes, 1ULL);
                    atomicAdd(&p.stats->sum_length, (unsigned long long)ep_length);
                    atomicAdd(&p.stats->sum_return, (double)ep_return);
                    ep_return = 0.0f;
                    ep_length = 0u;
                }
                if (p.auto_reset) {
                    episode += 1u;
                    int pi, pj;
                    draw_layout(s_map, H, W, wpr, p.N, p.seed, env_id, episode, gi, gj, pi, pj, au);
                    x = __fsub_rn((float)pi, 0.5f);   // tile centre (single_room.jl:125)
                    y = __fsub_rn((float)pj, 0.5f);
                }
            }
            if (writer) {
                p.out.pos_x[env] = x;
                p.out.pos_y[env] = y;
                p.out.dir_au[env] = au;
                p.out.goal[env] = (uint32_t)gi | ((uint32_t)gj << 16);
                p.out.episode[env] = episode;
                if (valid) {
                    p.reward[env] = reward;
                    p.done[env] = done ? 1 : 0;
                    p.ep_return[env] = ep_return;
                    p.ep_length[env] = ep_length;
                } else {
                    atomicExch(&p.stats->bad_action, 1);
                }
            }
        }

        // ---- cast_rays! (single_room.jl:195-231): lane <-> ray ----------------------------------
        const int r0 = g * 32;
        const int ncols = min(32, R - r0);
        const int ray = r0 + lane;
        const bool active = lane < ncols;
        const float2 dir = p.dir_slot >= 0 ? c_dirs[p.dir_slot][au] : __ldg(p.dirs + au);
        float4 rt = make_float4(1.0f, 0.0f, 1.0f, 1.0f);
        if (active) rt = __ldg(p.ray_table + (size_t)au * (size_t)R + ray);

        // RayCaster.cast_ray contract: tiles here are 0-based
        int ti = __float2int_rd(x), tj = __float2int_rd(y);
        const int si = rt.x < 0.0f ? -1 : 1, sj = rt.y < 0.0f ? -1 : 1;
        float tx = rt.x < 0.0f ? __fmul_rn(__fsub_rn(x, (float)ti), rt.z)
                               : __fmul_rn(__fsub_rn((float)(ti + 1), x), rt.z);
        float ty = rt.y < 0.0f ? __fmul_rn(__fsub_rn(y, (float)tj), rt.w)
                               : __fmul_rn(__fsub_rn((float)(tj + 1), y), rt.w);
        int dim = 0;
        float dist = 0.0f;
        bool hit_wall_tile = true;   // outside the map is painted as wall
        const bool tie_le = (p.dda_flags & RCW_DDA_TIE_LE) != 0;
        bool walking = active;
        // the warp leaves the loop together when the last lane has hit (ballot early-exit)
        while (__any_sync(0xFFFFFFFFu, walking)) {
            if (walking) {
                const bool inside = ((unsigned)ti < (unsigned)H) && ((unsigned)tj < (unsigned)W);
                const bool is_goal = (ti + 1 == gi) && (tj + 1 == gj);
                const bool is_wall = inside ? wall_bit(s_map, wpr, ti, tj) : true;
                if (is_wall || is_goal) {
                    hit_wall_tile = is_wall;
                    walking = false;
                } else {
                    const bool take_x = tie_le ? (tx <= ty) : (tx < ty);
                    if (take_x) {
                        dist = tx;
                        tx = __fadd_rn(tx, rt.z);
                        ti += si;
                        dim = 1;
                    } else {
                        dist = ty;
                        ty = __fadd_rn(ty, rt.w);
                        tj += sj;
                        dim = 2;
                    }
                }
            }
        }
        if ((p.dda_flags & RCW_DDA_DIST_POST) && dim != 0)
            dist = (dim == 1) ? __fsub_rn(tx, rt.z) : __fsub_rn(ty, rt.w);

        if (MODE == kModeRays) {
            if (active) {
                const size_t k = (size_t)env_rel * (size_t)R + (size_t)ray;
                p.dump_hit[2 * k + 0] = ti + 1;
                p.dump_hit[2 * k + 1] = tj + 1;
                p.dump_dim[k] = dim;
                p.dump_dist[k] = dist;
                p.dump_dir[2 * k + 0] = rt.x;
                p.dump_dir[2 * k + 1] = rt.y;
            }
            continue;
        }

        // ---- update_camera_view! (single_room.jl:374-444) ---------------------------------------
        // height of the wall line of this lane's ray (:404-411)
        {
            const float dot = __fadd_rn(__fmul_rn(dir.x, rt.x), __fmul_rn(dir.y, rt.y));
            const float proj = __fmul_rn(dist, dot);
            const float hl = __fdiv_rn(p.hl_num, __fmul_rn(p.two_s, proj));
            int h = P;                                   // non-finite => full height (:409-410)
            if (isfinite(hl) && hl < (float)P) h = max(__float2int_rd(hl), 0);
            const int pad = (h >= P - 1) ? 0 : ((P - h) >> 1);   // :433-436
            const uint32_t color = hit_wall_tile ? (dim == 1 ? p.palette[RCW_COLOR_WALL_1] : p.palette[RCW_COLOR_WALL_2])
                                                 : (dim == 1 ? p.palette[RCW_COLOR_GOAL_1] : p.palette[RCW_COLOR_GOAL_2]);
            // ray r paints column R-1-r (0-based); within the warp's span that is ncols-1-lane
            if (active) s_col[warp][ncols - 1 - lane] = make_uint2((uint32_t)pad, color);
        }
        __syncwarp();

        uint8_t* const env_obs = p.obs + (size_t)env * p.obs_env_stride;
        const int B0 = (R - r0 - ncols) * CB;          // byte span of the warp's columns in the env image
        const int B1 = B0 + ncols * CB;
        ColumnBands cb;
        cb.ceiling = p.palette[RCW_COLOR_CEILING];
        cb.floor = p.palette[RCW_COLOR_FLOOR];

        // pass 1: every aligned 16-byte vector of the span that lies inside one band of one column
        {
            const int v_hi = B1 >> 4;
            int v = ((B0 + 15) >> 4) + lane;
            int cl = 0, ob = 0;
            if (v < v_hi) {
                const int rel = (v << 4) - B0;
                cl = rel / CB;
                ob = rel - cl * CB;
            }
            for (; v < v_hi; v += 32) {
                if (ob + 16 <= CB) {
                    const uint2 info = s_col[warp][cl];
                    cb.b1 = (int)info.x * bpp;
                    cb.b2 = CB - cb.b1;
                    cb.wall = info.y;
                    const int cls = cb.classify16(ob);
                    if (cls != 3) {
                        const uint32_t c = cls == 0 ? cb.ceiling : (cls == 1 ? cb.wall : cb.floor);
                        store_stream16(env_obs + ((size_t)v << 4), PixelFormat<FMT>::run16(c, ob));
                    }
                }
                ob += 512;
                while (ob >= CB) {
                    ob -= CB;
                    ++cl;
                }
            }
        }
        // pass 2: lane <-> column; the (at most two) vectors that straddle a band boundary, and the
        // unaligned head / tail bytes of the column when col_bytes is not a multiple of 16
        if (active) {
            const uint2 info = s_col[warp][lane];
            cb.b1 = (int)info.x * bpp;
            cb.b2 = CB - cb.b1;
            cb.wall = info.y;
            const int S = B0 + lane * CB, E = S + CB;
            const int head_end = min(E, (S + 15) & ~15);
            const int tail_start = max(head_end, E & ~15);
            for (int b = S; b < head_end; ++b) env_obs[b] = (uint8_t)column_byte<FMT>(cb, b - S);
            for (int b = tail_start; b < E; ++b) env_obs[b] = (uint8_t)column_byte<FMT>(cb, b - S);
            const int va = (S + cb.b1) & ~15, vb = (S + cb.b2) & ~15;
#pragma unroll 1
            for (int k = 0; k < 2; ++k) {
                const int vo = k == 0 ? va : vb;
                if (k == 1 && vb == va) break;
                const int ob = vo - S;
                if (ob < 0 || ob + 16 > CB || cb.classify16(ob) != 3) continue;
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t acc = 0;
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        acc |= column_byte<FMT>(cb, ob + 4 * q + t) << (8 * t);
                    w[q] = acc;
                }
                store_stream16(env_obs + vo, make_uint4(w[0], w[1], w[2], w[3]));
            }
        }
        __syncwarp();   // s_col is rewritten by the next item
    }
}

// ------------------------------------------------------------------------------------------
// reset kernel: one thread per env (rare path)
// ------------------------------------------------------------------------------------------

__global__ void reset_kernel(const ResetParams p) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.num_envs) return;
    if (p.mask && !p.mask[env]) return;
    int gi, gj, pi, pj, au;
    uint32_t episode = p.st.episode[env];
    if (p.goal_ij) {
        gi = p.goal_ij[2 * env + 0];
        gj = p.goal_ij[2 * env + 1];
        pi = p.player_ij[2 * env + 0];
        pj = p.player_ij[2 * env + 1];
        au = p.dir_au[env];
    } else {
        episode += 1u;
        draw_layout(p.wall_map, p.H, p.W, p.wpr, p.N, p.seed, p.env_id_offset + (uint64_t)env,
                    episode, gi, gj, pi, pj, au);
    }
    p.st.pos_x[env] = __fsub_rn((float)pi, 0.5f);
    p.st.pos_y[env] = __fsub_rn((float)pj, 0.5f);
    p.st.dir_au[env] = au;
    p.st.goal[env] = (uint32_t)gi | ((uint32_t)gj << 16);
    p.st.episode[env] = episode;
    p.reward[env] = 0.0f;
    p.done[env] = 0;
    p.ep_return[env] = 0.0f;
    p.ep_length[env] = 0u;
}

// ------------------------------------------------------------------------------------------
// ray table: {ray_x, ray_y, |1/ray_x|, |1/ray_y|} for every (direction, ray)
//   single_room.jl:193,214-221 + Base.lerpi (Float64) + StaticArrays.normalize (inv(norm) * v)
// ------------------------------------------------------------------------------------------

__global__ void build_ray_table_kernel(const float2* __restrict__ dirs, int N, int R, float s,
                                       float4* __restrict__ table) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * R) return;
    const int au = idx / R, i0 = idx - au * R;
    const float2 d = dirs[au];
    const float c0 = d.y, c1 = -d.x;                              // rotate_minus_90 (:193)
    const float f0 = __fadd_rn(d.x, __fmul_rn(s, c0)), f1 = __fadd_rn(d.y, __fmul_rn(s, c1));
    const float l0 = __fsub_rn(d.x, __fmul_rn(s, c0)), l1 = __fsub_rn(d.y, __fmul_rn(s, c1));
    const int lendiv = max(R - 1, 1);
    const double t = __ddiv_rn((double)i0, (double)lendiv);      // lerpi: t = j / d in Float64
    const double omt = __dsub_rn(1.0, t);
    const float u0 = __double2float_rn(__dadd_rn(__dmul_rn(omt, (double)f0), __dmul_rn(t, (double)l0)));
    const float u1 = __double2float_rn(__dadd_rn(__dmul_rn(omt, (double)f1), __dmul_rn(t, (double)l1)));
    const float n = __fsqrt_rn(__fadd_rn(__fmul_rn(u0, u0), __fmul_rn(u1, u1)));
    const float q = __fdiv_rn(1.0f, n);
    const float rx = __fmul_rn(q, u0), ry = __fmul_rn(q, u1);
    table[idx] = make_float4(rx, ry, fabsf(__fdiv_rn(1.0f, rx)), fabsf(__fdiv_rn(1.0f, ry)));
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------

cudaError_t launch_build_ray_table(const float2* dirs, int N, int R, float sfov, float4* table,
                                   cudaStream_t s) {
    const int n = N * R;
    build_ray_table_kernel<<<(n + 255) / 256, 256, 0, s>>>(dirs, N, R, sfov, table);
    return cudaGetLastError();
}

template <int MODE, int FMT>
static cudaError_t launch_frame_t(const FrameParams& p, int ctas, cudaStream_t s) {
    const size_t smem = (size_t)p.map_words * 4;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(frame_kernel<MODE, FMT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    frame_kernel<MODE, FMT><<<ctas, kThreadsPerCta, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_frame(const FrameParams& p, int mode, int obs_format, int ctas, cudaStream_t s) {
    const bool rgb = obs_format == RCW_OBS_RGB8;
    switch (mode) {
        case kModeStep:
            return rgb ? launch_frame_t<kModeStep, RCW_OBS_RGB8>(p, ctas, s)
                       : launch_frame_t<kModeStep, RCW_OBS_XRGB32>(p, ctas, s);
        case kModeRender:
            return rgb ? launch_frame_t<kModeRender, RCW_OBS_RGB8>(p, ctas, s)
                       : launch_frame_t<kModeRender, RCW_OBS_XRGB32>(p, ctas, s);
        case kModeRays:
            return launch_frame_t<kModeRays, RCW_OBS_RGB8>(p, ctas, s);
        default:
            return cudaErrorInvalidValue;
    }
}

int frame_kernel_max_ctas_per_sm(int obs_format, int map_bytes) {
    int n = 0;
    cudaError_t e = obs_format == RCW_OBS_RGB8
                        ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                              &n, frame_kernel<kModeStep, RCW_OBS_RGB8>, kThreadsPerCta, (size_t)map_bytes)
                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                              &n, frame_kernel<kModeStep, RCW_OBS_XRGB32>, kThreadsPerCta, (size_t)map_bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

cudaError_t launch_reset(const ResetParams& p, cudaStream_t s) {
    const int64_t blocks = (p.num_envs + 255) / 256;
    reset_kernel<<<(unsigned)blocks, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t upload_dir_slot(int slot, const float2* host_dirs, int n, cudaStream_t s) {
    return cudaMemcpyToSymbolAsync(c_dirs, host_dirs, sizeof(float2) * (size_t)n,
                                   sizeof(float2) * (size_t)slot * kDirSlotEntries,
                                   cudaMemcpyHostToDevice, s);
}

}  // namespace rcw
