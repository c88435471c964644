// rcw_kernels.cu — sm_100a kernels of the batched SingleRoom engine.
//
// One launch of frame_kernel<kModeStep> is one env-step of the whole batch:
//   act!(world) (+ auto-reset)  ->  cast_rays!(world)  ->  update_camera_view!(env)
// (reference: src/single_room.jl:139-191, 195-231, 374-444; collision_detection.jl:9-42).
//
// Mapping.  A work item is (env, group of 32 consecutive rays) and belongs to one warp:
// lane <-> ray for the DDA, then all 32 lanes stream the group's 32 observation columns, which
// are contiguous in memory (ray i paints column R-i+1, single_room.jl:431), in whole 32-byte
// sectors with 256-bit stores.  The four warps of a CTA take four consecutive items, so a CTA
// writes one contiguous span of the observation buffer.  The path is bound by HBM writes (393 KB
// per env-step at the default resolution versus a few hundred bytes of state), so everything else
// is arranged to keep the store stream dense and complete: the wall layer is staged into shared
// memory with TMA bulk copies (once per CTA when the batch shares it, once per env otherwise),
// the per-(direction, ray) table {ray, |1/ray|} is read with one coalesced 16-byte load per lane
// from an L2-resident table (prefetched for the three directions the env can face next), act! /
// auto-reset run once per env of a CTA and reach the other warps through shared memory, and the
// state they read is double-buffered so that the warps of an env in other CTAs never race.
// When the items are small (narrow or one-byte-per-pixel cameras) the step is bound by act! and the DDA
// instead, and env_kernel gives a whole env to one warp (no block barrier behind act!).  Host-supplied
// actions of up to 32,768 envs ride in the kernel parameters (frame_kernel_pa / env_kernel_pa).
// rcw_topview.cuh (included below) draws the reference's top view.
//
// Arithmetic.  Every binary32 operation that the reference performs is written with an explicit
// round-to-nearest intrinsic (__fmul_rn, __fadd_rn, __fdiv_rn, __fsqrt_rn), which the compiler
// never contracts into FMAs, so positions, hit tiles, hit sides and distances are bit-identical
// to the CPU restatement (the file is also compiled with -fmad=false).

#include <type_traits>

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

// TMA 1-D bulk copy shared -> global (SASS: UBLKCP), tracked by the thread's bulk async-group.
__device__ __forceinline__ void bulk_copy_s2g(void* dst_gmem, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(src_smem), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// all bulk stores of this thread have finished READING shared memory (the CTA may exit)
__device__ __forceinline__ void bulk_wait_group_read0() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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

// The wall layer as the kernels see it (0-based tiles; the caller guarantees the tile is inside the map).
// BitsMap: bit-packed words, bit j0 of row i0 — any map, staged in shared memory (or read from global memory
// by the reset kernel).  RoomMap: the border tiles are walls and nothing else is — the map every SingleRoom of
// the reference has (single_room.jl:57-60); it needs no memory at all, and its DDA counts down to the border
// instead of probing (dda_walk_room).  The host picks the view per launch (FrameParams::room).
// Object layers beyond WALL and GOAL (NUM_OBJECTS > 2, single_room.jl:16-18) are further bit-packed
// layers behind the wall layer, `stride` words apart: [wall][extra 0] ... [extra n-1][any = OR of all of them], staged
// together.  Rays stop at any object (:209), so the DDA probes the `any` layer — which IS the wall layer when there
// are no extra objects — and only the tile a ray stopped on is looked up layer by layer.
struct BitsMap {
    const uint32_t* w;      // the wall layer
    const uint32_t* any;    // wall | every extra layer (== w without extra layers)
    int wpr;
    int n_extra;
    int stride;             // words between consecutive layers
    __device__ __forceinline__ static bool bit(const uint32_t* m, int wpr, int i0, int j0) {
        return (m[i0 * wpr + (j0 >> 5)] >> (j0 & 31)) & 1u;
    }
    __device__ __forceinline__ bool wall(int i0, int j0) const { return bit(w, wpr, i0, j0); }
    __device__ __forceinline__ bool obstacle(int i0, int j0) const { return bit(any, wpr, i0, j0); }
    __device__ __forceinline__ bool extra(int k, int i0, int j0) const { return bit(w + (k + 1) * stride, wpr, i0, j0); }
};
struct RoomMap {
    int H1, W1;   // H - 1, W - 1
    static constexpr int n_extra = 0;
    __device__ __forceinline__ bool wall(int i0, int j0) const {
        return ((unsigned)(i0 - 1) >= (unsigned)(H1 - 1)) | ((unsigned)(j0 - 1) >= (unsigned)(W1 - 1));
    }
    __device__ __forceinline__ bool obstacle(int i0, int j0) const { return wall(i0, j0); }
    __device__ __forceinline__ bool extra(int, int, int) const { return false; }
};

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
template <class Map>
__device__ inline void draw_layout(const Map& map, int H, int W, int N, uint64_t seed,
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
        const bool occupied = map.obstacle(pi - 1, pj - 1) || (pi == gi && pj == gj);   // any(tile_map[:, pos]) (utils.jl:27)
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
    // R == G == B: the byte stream of a run is the same byte repeated, whatever the phase
    __device__ static __forceinline__ bool is_flat(uint32_t c) {
        return ((c ^ (c >> 8)) & 0xFFFFu) == 0;
    }
    __device__ static __forceinline__ uint32_t flat_word(uint32_t c) { return (c & 0xFFu) * 0x01010101u; }
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
    __device__ static __forceinline__ bool is_flat(uint32_t) { return true; }   // one pixel = one word
    __device__ static __forceinline__ uint32_t flat_word(uint32_t c) { return c; }
    __device__ static __forceinline__ uint4 run16(uint32_t c, int ob) {
        const uint32_t sh = 8u * (uint32_t)(ob & 3);
        const uint32_t w = __funnelshift_r(c, c, sh);
        return make_uint4(w, w, w, w);
    }
};

// one byte per pixel; the palette handed to the kernel already holds the luma in all three bytes
template <>
struct PixelFormat<RCW_OBS_GRAY8> {
    static constexpr int kBpp = 1;
    __device__ static __forceinline__ uint32_t byte_of(uint32_t c, int) { return c & 0xFFu; }
    __device__ static __forceinline__ bool is_flat(uint32_t) { return true; }
    __device__ static __forceinline__ uint32_t flat_word(uint32_t c) { return (c & 0xFFu) * 0x01010101u; }
    __device__ static __forceinline__ uint4 run16(uint32_t c, int) {
        const uint32_t w = (c & 0xFFu) * 0x01010101u;
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

// Observation stores: written once, read later by another kernel (the learner), never re-read
// here -> streaming (evict-first) by default.  RCW_STORE_POLICY / RCW_EXP are build-time A/B knobs.
#ifndef RCW_EXP
#define RCW_EXP 0   // development experiments (profiles/README.md): 1 = no act / DDA (store phase only), 2 = no stores
#endif
#ifndef RCW_STORE_POLICY
#define RCW_STORE_POLICY 0
#endif
#ifndef RCW_TABLE_COPY
#define RCW_TABLE_COPY 1   // env_kernel's table renderer: 1 = lane <-> consecutive sectors of the span (coalesced stores,
                           // shipped), 0 = lane <-> column (fewer instructions, but every warp store scatters 32 sectors
                           // col_pitch apart: 5-25 % slower, profiles/README.md)
#endif
__device__ __forceinline__ void store_stream16(uint8_t* p, uint4 v) {
#if RCW_EXP == 2
    if (v.x == 0x12345u) __stcs(reinterpret_cast<uint4*>(p), v);   // experiment: compute only
#elif RCW_STORE_POLICY == 0
    __stcs(reinterpret_cast<uint4*>(p), v);
#elif RCW_STORE_POLICY == 1
    *reinterpret_cast<uint4*>(p) = v;
#elif RCW_STORE_POLICY == 2
    __stwt(reinterpret_cast<uint4*>(p), v);
#else
    __stcg(reinterpret_cast<uint4*>(p), v);
#endif
}

// mask of the bytes of a little-endian word whose index is below d (d <= 0: none, d >= 4: all)
__device__ __forceinline__ uint32_t low_bytes_mask(int d) {
    uint32_t r;
    // shl.b32 clamps shift amounts above 31 to 32 (result 0), which a C shift does not promise
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(0xFFFFFFFFu), "r"(8u * (uint32_t)max(d, 0)));
    return ~r;
}

// The 16 bytes at column offset ob when they straddle a band boundary: the three single-colour
// runs merged under byte masks (ceiling below b1, wall colour below b2, floor above).
template <int FMT>
__device__ __forceinline__ uint4 compose16(const ColumnBands& cb, int ob) {
    const uint4 c = PixelFormat<FMT>::run16(cb.ceiling, ob);
    const uint4 w = PixelFormat<FMT>::run16(cb.wall, ob);
    const uint4 f = PixelFormat<FMT>::run16(cb.floor, ob);
    const uint32_t cw[4] = {c.x, c.y, c.z, c.w}, ww[4] = {w.x, w.y, w.z, w.w}, fw[4] = {f.x, f.y, f.z, f.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t m1 = low_bytes_mask(cb.b1 - ob - 4 * k);
        const uint32_t m2 = low_bytes_mask(cb.b2 - ob - 4 * k);
        const uint32_t wf = (ww[k] & m2) | (fw[k] & ~m2);
        o[k] = (cw[k] & m1) | (wf & ~m1);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// The parts of one column (env-relative bytes [S, S + CB)) that are not whole single-colour
// 16-byte vectors: the (at most two) aligned vectors that straddle a band boundary, and — only
// when col_bytes is not a multiple of 16 — the unaligned head / tail bytes.
template <int FMT>
__device__ __forceinline__ void column_edges(uint8_t* env_obs, int S, int CB, const ColumnBands& cb) {
    const int E = S + CB;
    if ((S | CB) & 15) {
        const int head_end = min(E, (S + 15) & ~15);
        const int tail_start = max(head_end, E & ~15);
        for (int b = S; b < head_end; ++b) env_obs[b] = (uint8_t)column_byte<FMT>(cb, b - S);
        for (int b = tail_start; b < E; ++b) env_obs[b] = (uint8_t)column_byte<FMT>(cb, b - S);
    }
    const int va = (S + cb.b1) & ~15, vb = (S + cb.b2) & ~15;
    const int oa = va - S, ob = vb - S;
    if (oa >= 0 && oa + 16 <= CB && cb.classify16(oa) == 3)
        store_stream16(env_obs + va, compose16<FMT>(cb, oa));
    if (vb != va && ob >= 0 && ob + 16 <= CB && cb.classify16(ob) == 3)
        store_stream16(env_obs + vb, compose16<FMT>(cb, ob));
}

constexpr int kPatChunk = 3072;   // bytes per bulk store: a multiple of 16 (TMA) and of 3 and 4 (pixel phase)

// One band [lo, hi) (env-relative bytes) of a column starting at S: the 16-byte aligned interior is
// one TMA bulk store (chunked) out of the band colour's pattern buffer, read at the offset that
// has the same pixel phase as the destination.
template <int FMT>
__device__ __forceinline__ void band_bulk(uint8_t* env_obs, int S, int lo, int hi, uint32_t s_pat_color) {
    const int alo = (lo + 15) & ~15, ahi = hi & ~15;
    if (ahi <= alo) return;
    const int ob = alo - S;
    // RGB8: 16 = 1 (mod 3), so source offset 16 * (ob mod 3) is 16-byte aligned and in phase;
    // XRGB32: columns start on pixel boundaries, every 16-byte aligned offset is in phase
    const uint32_t src = s_pat_color + (FMT == RCW_OBS_RGB8 ? 16u * (uint32_t)(ob % 3) : 0u);
    for (int off = alo; off < ahi; off += kPatChunk)
        bulk_copy_s2g(env_obs + off, src, (uint32_t)min(kPatChunk, ahi - off));
}

template <int FMT>
__device__ __forceinline__ void column_bulk(uint8_t* env_obs, int S, int CB, const ColumnBands& cb, int cid,
                                            uint32_t s_pat, int pat_stride) {
    band_bulk<FMT>(env_obs, S, S, S + cb.b1, s_pat + (uint32_t)(RCW_COLOR_CEILING * pat_stride));
    band_bulk<FMT>(env_obs, S, S + cb.b1, S + cb.b2, s_pat + (uint32_t)(cid * pat_stride));
    band_bulk<FMT>(env_obs, S, S + cb.b2, S + CB, s_pat + (uint32_t)(RCW_COLOR_FLOOR * pat_stride));
}

// ------------------------------------------------------------------------------------------
// act! for one env, executed by a whole warp (state identical in all lanes on entry and exit)
// ------------------------------------------------------------------------------------------

struct EnvPose {
    float x, y;
    int au;
    uint32_t goal;   // (i | j << 16), 1-based
};

// what act! needs from HBM; loaded early so the latency overlaps the CTA prologue
struct EnvInputs {
    float x, y;
    int au;
    uint32_t goal, episode;
    int action;
};

// `pa`: the launch's actions packed into the kernel parameters (or nullptr); `env_rel`: env index in the launch
__device__ __forceinline__ EnvInputs load_env_inputs(const FrameParams& p, int64_t env, const PackedActions* pa,
                                                     uint32_t env_rel) {
    EnvInputs in;
    in.x = __ldg(p.in.pos_x + env);
    in.y = __ldg(p.in.pos_y + env);
    in.au = __ldg(p.in.dir_au + env);
    in.goal = __ldg(p.in.goal + env);
    in.episode = __ldg(p.in.episode + env);
    if (pa) in.action = (int)((pa->w[env_rel >> 4] >> ((env_rel & 15u) * 2u)) & 3u) + 1;
    else in.action = p.actions ? (int)__ldg(p.actions + env) : -1;   // -1: random policy
    return in;
}

// single_room.jl:139-191 (+ same-step auto-reset and episode bookkeeping of the batched engine).
// `writer` is true in the one warp of the whole grid that owns the env's persistent state.
template <class Map>
__device__ __forceinline__ EnvPose act_env(const FrameParams& p, const Map& map, int64_t env,
                                           const EnvInputs& in, bool writer, int lane) {
    const int H = p.H, W = p.W;
    float x = in.x, y = in.y;
    int au = in.au;
    uint32_t episode = in.episode;
    int gi = (int)(in.goal & 0xFFFFu), gj = (int)(in.goal >> 16);
    const uint64_t env_id = p.env_id_offset + (uint64_t)env;
    const int a = in.action >= 0 ? in.action : draw_action(p.seed, env_id, p.step_index);
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
            // is_player_colliding (collision_detection.jl:21-42): lanes 0..8 probe the 3x3 tiles
            // around wu_to_tu(candidate); tiles outside the map are empty.
            const int ti = __float2int_rd(nx) + (lane % 3);        // = ip - 1 + di, 1-based
            const int tj = __float2int_rd(ny) + ((lane / 3) % 3);
            bool hit_goal = false, hit_wall = false;
            uint32_t hit_extra = 0u;          // bit k: the circle touches a tile of extra object layer k
            if (lane < 9 && ti >= 1 && ti <= H && tj >= 1 && tj <= W) {
                const bool is_goal = (ti == gi) && (tj == gj);
                const bool is_obstacle = map.obstacle(ti - 1, tj - 1);
                if (is_goal || is_obstacle) {
                    const bool c = circle_hits_tile(nx, ny, ti, tj, p.radius);
                    hit_goal = is_goal && c;
                    if (map.n_extra == 0) {
                        hit_wall = is_obstacle && c;
                    } else if (is_obstacle && c) {
                        hit_wall = map.wall(ti - 1, tj - 1);
                        for (int k = 0; k < map.n_extra; ++k) hit_extra |= map.extra(k, ti - 1, tj - 1) ? (1u << k) : 0u;
                    }
                }
            }
            bool any_goal = __any_sync(0xFFFFFFFFu, hit_goal);
            bool any_wall = __any_sync(0xFFFFFFFFu, hit_wall);
            float goal_reward = p.goal_reward;
            if (map.n_extra != 0) {
                // extra objects act like one of the reference's two: a terminal layer like GOAL (its own reward, done,
                // no move; checked in object order after GOAL), a blocking layer like WALL
                const uint32_t touched = __reduce_or_sync(0xFFFFFFFFu, hit_extra);
                const uint32_t terminal = touched & p.layer_terminal_mask;
                if (!any_goal && terminal) {
                    any_goal = true;
                    goal_reward = p.layer_reward[__ffs(terminal) - 1];
                }
                any_wall |= (touched & ~p.layer_terminal_mask) != 0u;
            }
            if (any_goal) {           // single_room.jl:166-168 — reward, done, no move
                reward = goal_reward;
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
    const bool w0 = writer && lane == 0;
    if (w0 && valid) {
        ep_return = __fadd_rn(p.ep_return[env], reward);
        ep_length = p.ep_length[env] + 1u;
    }
    if (done) {
        if (w0) {
            atomicAdd(&p.stats->episodes, 1ULL);
            atomicAdd(&p.stats->sum_length, (unsigned long long)ep_length);
            atomicAdd(&p.stats->sum_return, (double)ep_return);
            ep_return = 0.0f;
            ep_length = 0u;
        }
        if (p.auto_reset) {
            episode += 1u;
            int pi, pj;
            draw_layout(map, H, W, p.N, p.seed, env_id, episode, gi, gj, pi, pj, au);
            x = __fsub_rn((float)pi, 0.5f);   // tile centre (single_room.jl:125)
            y = __fsub_rn((float)pj, 0.5f);
        }
    }
    EnvPose pose;
    pose.x = x;
    pose.y = y;
    pose.au = au;
    pose.goal = (uint32_t)gi | ((uint32_t)gj << 16);
    if (w0) {
        p.out.pos_x[env] = x;
        p.out.pos_y[env] = y;
        p.out.dir_au[env] = au;
        p.out.goal[env] = pose.goal;
        p.out.episode[env] = episode;
        if (valid) {
            p.reward[env] = reward;
            p.done[env] = done ? 1 : 0;
            p.ep_return[env] = ep_return;
            p.ep_length[env] = ep_length;
            if (p.host_reward) {   // write-through to the result ring in host memory (posted PCIe writes)
                p.host_reward[env] = reward;
                p.host_done[env] = done ? 1 : 0;
            }
        } else {
            atomicExch(&p.stats->bad_action, 1);
            if (p.host_reward) {   // the env was not stepped: its previous reward / done stand
                p.host_reward[env] = p.reward[env];
                p.host_done[env] = p.done[env];
            }
        }
    }
    return pose;
}

// ------------------------------------------------------------------------------------------
// the frame kernel
// ------------------------------------------------------------------------------------------

// What one lane knows about its ray's column after cast_rays! + the height computation.
struct ColumnShade {
    int pad;         // rows of ceiling (= rows of floor); 0 = full-height column
    int cid;         // palette index of the wall / goal colour
};

// What RayCaster.cast_ray returns for one ray (tiles 0-based here), plus which layer stopped it.
struct RayHit {
    int ti, tj;      // hit tile
    int dim;         // 1: crossed along dimension 1 (i), 2: along dimension 2 (j), 0: started inside an obstacle
    float dist;      // Euclidean distance along the unit ray
    int object;      // first object on the hit tile, findfirst(tile_map[:, i, j]) - 1: 0 wall (outside the map counts
                     // as wall), 1 goal, 2 + k extra object layer k
};

// The DDA walk itself, branch-free per step.  TIE_LE: decision D1 (advance along dimension 1 on a tie).
// CLOSED: every border tile of the wall layer is a wall, so a ray can never leave the map and the probe
// needs no bounds test (the default SingleRoom map; checked on the host for supplied maps).
// A lane that has hit stays on its obstacle tile; the warp leaves together once no lane is still walking.
template <bool TIE_LE, bool CLOSED, class Map>
__device__ __forceinline__ void dda_walk(const Map& map, int H, int W, int gi0, int gj0, float dx,
                                         float dy, int si, int sj, float& tx, float& ty, int& ti, int& tj, int& dim,
                                         float& dist) {
    // the tile the ray stands on is an obstacle: wall (outside the map counts as wall) or this env's goal
    auto probe = [&]() {
        bool wall;
        if (CLOSED) {
            wall = map.obstacle(ti, tj);
        } else {
            const bool inside = ((unsigned)ti < (unsigned)H) & ((unsigned)tj < (unsigned)W);
            wall = !inside | map.obstacle(inside ? ti : 0, inside ? tj : 0);
        }
        return wall | ((ti == gi0) & (tj == gj0));
    };
    bool stop = probe();
#pragma unroll 1
    while (__any_sync(0xFFFFFFFFu, !stop)) {
        // kDdaStepsPerVote DDA steps per warp vote: a stopped lane just probes its own tile again
#pragma unroll
        for (int u = 0; u < kDdaStepsPerVote; ++u) {
            const bool cmp = TIE_LE ? (tx <= ty) : (tx < ty);
            const bool mx = !stop & cmp, my = !stop & !cmp;
            dist = mx ? tx : (my ? ty : dist);
            tx = mx ? __fadd_rn(tx, dx) : tx;
            ty = my ? __fadd_rn(ty, dy) : ty;
            ti += mx ? si : 0;
            tj += my ? sj : 0;
            dim = mx ? 1 : (my ? 2 : dim);
            stop = probe();
        }
    }
}

// The same walk through a room (RoomMap: walls on the border only), from a start tile inside the map.  The side
// distances, the comparison, the order of the steps and therefore tile, dimension and distance are those of
// dda_walk; what changes is the probe.  A ray that walks away from an interior tile can only meet the border
// it walks towards, so instead of looking tiles up the lane counts the steps left to that border in each
// dimension — both counts in one register, (ci << 16) | cj, decremented by 0x10000 or 1 — and stops when either
// reaches zero (one subtract and one mask test: a half that is zero borrows) or when the pair equals the
// goal's (one compare).  No shared memory, no address arithmetic: 11 instructions per step instead of ~30.
template <bool TIE_LE>
__device__ __forceinline__ void dda_walk_room(const RoomMap& map, int gi0, int gj0, float dx, float dy, int si,
                                              int sj, float& tx, float& ty, int& ti, int& tj, int& dim, float& dist) {
    // steps left until the border tile in the direction of travel (H, W <= 32767: FrameParams::room)
    const int ci0 = si > 0 ? map.H1 - ti : ti, cj0 = sj > 0 ? map.W1 - tj : tj;
    const int gci = si > 0 ? map.H1 - gi0 : gi0, gcj = sj > 0 ? map.W1 - gj0 : gj0;
    // a goal outside 0..32767 in either dimension can never be met: give it a pair no count reaches
    const bool goal_in = ((unsigned)gci < 0x8000u) & ((unsigned)gcj < 0x8000u);
    const uint32_t gcnt = goal_in ? (((uint32_t)gci << 16) | (uint32_t)gcj) : 0xFFFFFFFFu;
    // A start tile on the border is a wall whichever way the ray points; such a lane never walks.  Its count is
    // zeroed so that the stop test below holds for it in every step (the tile is restored after the loop).
    const bool start_wall = map.wall(ti, tj);
    uint32_t cnt = start_wall ? 0u : (((uint32_t)ci0 << 16) | (uint32_t)cj0);
    const uint32_t stop0 = (start_wall | (cnt == gcnt)) ? 1u : 0u;
    uint32_t last_x = 0u;
    // The loop in PTX, so that a step stays the twelve predicated instructions it is (the compiler turns the C++
    // form into branches and register moves, 17-19 per step).  Per step, for a lane that has not stopped:
    //   px = tx < ty (<= with D1), py = !px;  dist = px ? tx : ty;  @px { tx += dx; count -= 0x10000 }
    //   @py { ty += dy; count -= 1 };  stop = a half of the count reached zero (it borrows) | count == goal's.
    // A stopped lane keeps its count, so its stop test keeps holding; the warp votes every kDdaStepsPerVote steps.
#define RCW_ROOM_STEP(CMP)                                  \
    "setp." CMP ".and.f32 px|py, %0, %1, !ps;\n\t"          \
    "@!ps selp.f32 %3, %0, %1, px;\n\t"                     \
    "@!ps selp.u32 %4, 1, 0, px;\n\t"                       \
    "@px add.rn.f32 %0, %0, %5;\n\t"                        \
    "@py add.rn.f32 %1, %1, %6;\n\t"                        \
    "@px sub.u32 %2, %2, 0x10000;\n\t"                      \
    "@py sub.u32 %2, %2, 1;\n\t"                            \
    "sub.u32 t, %2, 0x00010001;\n\t"                        \
    "and.b32 t, t, 0x80008000;\n\t"                         \
    "setp.ne.u32 ps, t, 0;\n\t"                             \
    "setp.eq.or.u32 ps, %2, %7, ps;\n\t"
#define RCW_ROOM_LOOP(CMP)                                                                              \
    asm volatile("{\n\t"                                                                                \
                 ".reg .pred ps, px, py, pgo;\n\t"                                                       \
                 ".reg .u32 t;\n\t"                                                                      \
                 "setp.ne.u32 ps, %8, 0;\n"                                                              \
                 "RCW_ROOM_LOOP_TOP:\n\t"                                                                \
                 "vote.sync.any.pred pgo, !ps, 0xffffffff;\n\t"                                          \
                 "@!pgo bra RCW_ROOM_LOOP_END;\n\t"                                                      \
                 RCW_ROOM_STEP(CMP) RCW_ROOM_STEP(CMP)                                                  \
                 "bra RCW_ROOM_LOOP_TOP;\n"                                                              \
                 "RCW_ROOM_LOOP_END:\n\t"                                                                \
                 "}"                                                                                    \
                 : "+f"(tx), "+f"(ty), "+r"(cnt), "+f"(dist), "+r"(last_x)                              \
                 : "f"(dx), "f"(dy), "r"(gcnt), "r"(stop0))
    static_assert(kDdaStepsPerVote == 2, "RCW_ROOM_LOOP unrolls two steps per vote");
    if (TIE_LE) RCW_ROOM_LOOP("le");
    else RCW_ROOM_LOOP("lt");
#undef RCW_ROOM_LOOP
#undef RCW_ROOM_STEP
    if (!start_wall) {
        const int ci = (int)(cnt >> 16), cj = (int)(cnt & 0xFFFFu);
        if (ci != ci0 || cj != cj0) dim = last_x ? 1 : 2;
        ti = si > 0 ? map.H1 - ci : ci;
        tj = sj > 0 ? map.W1 - cj : cj;
    }
}

// RayCaster.cast_ray contract for the ray rt = {ray_x, ray_y, |1/ray_x|, |1/ray_y|} from (x, y).
// Must be called by all 32 lanes of a warp (lane <-> ray).  `closed`: see dda_walk.
template <class Map>
__device__ __forceinline__ RayHit dda_cast(const Map& map, int H, int W, uint32_t dda_flags,
                                           bool closed, float x, float y, int gi0, int gj0, const float4 rt, int lane) {
    int ti = __float2int_rd(x), tj = __float2int_rd(y);
    const int si = rt.x < 0.0f ? -1 : 1, sj = rt.y < 0.0f ? -1 : 1;
    float tx = rt.x < 0.0f ? __fmul_rn(__fsub_rn(x, (float)ti), rt.z)
                           : __fmul_rn(__fsub_rn((float)(ti + 1), x), rt.z);
    float ty = rt.y < 0.0f ? __fmul_rn(__fsub_rn(y, (float)tj), rt.w)
                           : __fmul_rn(__fsub_rn((float)(tj + 1), y), rt.w);
    int dim = 0;
    float dist = 0.0f;
#if RCW_EXP == 1
    dist = 1.0f + 0.01f * (float)lane; dim = 1 + (lane & 1);
#else
    (void)lane;
    const bool tie_le = (dda_flags & RCW_DDA_TIE_LE) != 0;
    // The unchecked walk is only safe from inside the map.  A player can be outside it: injected onto a border
    // tile and stepped outwards, or carried over the border wall by an increment larger than a tile (the
    // collision test looks at the candidate position only, collision_detection.jl:21-42).  Same pose in all lanes.
    const bool start_inside = ((unsigned)ti < (unsigned)H) & ((unsigned)tj < (unsigned)W);
    if constexpr (std::is_same<Map, RoomMap>::value) {
        if (start_inside) {
            if (tie_le) dda_walk_room<true>(map, gi0, gj0, rt.z, rt.w, si, sj, tx, ty, ti, tj, dim, dist);
            else dda_walk_room<false>(map, gi0, gj0, rt.z, rt.w, si, sj, tx, ty, ti, tj, dim, dist);
        }   // from outside the map every tile counts as wall: the ray stops where it stands (dim 0, dist 0)
    } else {
        if (closed && start_inside) {
            if (tie_le) dda_walk<true, true>(map, H, W, gi0, gj0, rt.z, rt.w, si, sj, tx, ty, ti, tj, dim, dist);
            else dda_walk<false, true>(map, H, W, gi0, gj0, rt.z, rt.w, si, sj, tx, ty, ti, tj, dim, dist);
        } else {
            if (tie_le) dda_walk<true, false>(map, H, W, gi0, gj0, rt.z, rt.w, si, sj, tx, ty, ti, tj, dim, dist);
            else dda_walk<false, false>(map, H, W, gi0, gj0, rt.z, rt.w, si, sj, tx, ty, ti, tj, dim, dist);
        }
    }
#endif
    if ((dda_flags & RCW_DDA_DIST_POST) && dim != 0)
        dist = (dim == 1) ? __fsub_rn(tx, rt.z) : __fsub_rn(ty, rt.w);
    RayHit h;
    h.ti = ti;
    h.tj = tj;
    h.dim = dim;
    h.dist = dist;
    // which object stopped the ray (single_room.jl:417-429 asks tile_map[WALL, i, j], else GOAL; with more objects:
    // the first one on the tile): a wall (outside the map counts as wall), this env's goal, an extra layer
    const bool inside = ((unsigned)ti < (unsigned)H) & ((unsigned)tj < (unsigned)W);
    const int i0 = inside ? ti : 0, j0 = inside ? tj : 0;
    h.object = 0;
    if (inside & !map.wall(i0, j0)) {
        h.object = 1;
        if (map.n_extra != 0 && !((ti == gi0) & (tj == gj0))) {
            h.object = 0;                       // (a stopped ray stands on an object; none found = treated as wall)
            for (int k = map.n_extra - 1; k >= 0; --k) h.object = map.extra(k, i0, j0) ? 2 + k : h.object;
        }
    }
    return h;
}

// cast_rays! for the 32 rays [32 g, 32 g + 32) of an env (single_room.jl:195-231) and the height /
// colour of every ray's column (:404-429).  lane <-> ray.  Optionally dumps the ray results.
template <int MODE, class Map>
__device__ __forceinline__ ColumnShade cast_and_shade(const FrameParams& p, const Map& map,
                                                      const EnvPose& pose, const float4 rt, int g, int lane,
                                                      uint32_t env_rel) {
    const int R = p.R, P = p.P;
    const int au = pose.au;
    const int gi0 = (int)(pose.goal & 0xFFFFu) - 1, gj0 = (int)(pose.goal >> 16) - 1;
    const int ray = g * 32 + lane;
    const float2 dir = p.dir_slot >= 0 ? c_dirs[p.dir_slot][au] : __ldg(p.dirs + au);
    const RayHit hit = dda_cast(map, p.H, p.W, p.dda_flags, p.closed_border != 0, pose.x, pose.y, gi0, gj0, rt, lane);
    const int ti = hit.ti, tj = hit.tj, dim = hit.dim;
    const float dist = hit.dist;

    if (MODE == kModeRays) {
        if (ray < R) {
            const size_t k = (size_t)env_rel * (size_t)R + (size_t)ray;
            p.dump_hit[2 * k + 0] = ti + 1;
            p.dump_hit[2 * k + 1] = tj + 1;
            p.dump_dim[k] = dim;
            p.dump_dist[k] = dist;
            p.dump_dir[2 * k + 0] = rt.x;
            p.dump_dir[2 * k + 1] = rt.y;
        }
    }
    // update_camera_view!: height of the wall line (:404-411), padding (:433-436), colour (:417-429)
    const float dot = __fadd_rn(__fmul_rn(dir.x, rt.x), __fmul_rn(dir.y, rt.y));
    const float proj = __fmul_rn(dist, dot);
    const float hl = __fdiv_rn(p.hl_num, __fmul_rn(p.two_s, proj));
    int h = P;                                   // non-finite => full height (:409-410)
    if (isfinite(hl) && hl < (float)P) h = max(__float2int_rd(hl), 0);
    ColumnShade cs;
    cs.pad = (h >= P - 1) ? 0 : ((P - h) >> 1);
    cs.cid = RCW_COLOR_WALL_1 + 2 * hit.object + (dim == 1 ? 0 : 1);   // 2/3 wall, 4/5 goal, 6 + 2 k / 7 + 2 k extra layer k
    return cs;
}

// 32 bytes = one L2 / DRAM sector, written by one lane with one 256-bit store (sm_100: STG.256).
// The observation stream must be written in whole sectors: a 16-byte store that leaves the other
// half of its sector for later makes L2 fetch the sector from DRAM to merge it, which cost 35 % of
// the step time when the renderer skipped the vectors that straddle a band boundary (profiles/).
__device__ __forceinline__ void store_stream32(uint8_t* p, uint4 lo, uint4 hi) {
#if RCW_EXP == 2
    if (lo.x != 0x12345u) return;   // experiment: compute only
#endif
    // Write-once stream: keep it out of L1 and mark the L2 lines evict-first, so that they drain to
    // HBM behind the writer instead of aging in the 126 MB L2 (+5 % over .cs, profiles/README.md).
#if RCW_STORE_POLICY == 0
#define RCW_ST256 "st.global.L1::no_allocate.L2::evict_first.v8.b32"
#elif RCW_STORE_POLICY == 1
#define RCW_ST256 "st.global.cs.v8.b32"
#elif RCW_STORE_POLICY == 2
#define RCW_ST256 "st.global.v8.b32"
#else
#define RCW_ST256 "st.global.L1::no_allocate.v8.b32"
#endif
    asm volatile(RCW_ST256 " [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(lo.x), "r"(lo.y),
                 "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                 : "memory");
}

// A sector with one band boundary n bytes in, when both bands are "flat" words (one byte repeated,
// or whole 4-byte pixels with n a multiple of 4): bytes [0, n) from word a, [n, 32) from word b.
__device__ __forceinline__ void store_split32(uint8_t* dst, uint32_t a, uint32_t b, int n) {
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t m = low_bytes_mask(n - 4 * k);
        w[k] = (a & m) | (b & ~m);
    }
    store_stream32(dst, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
}

// update_camera_view! stores (single_room.jl:431-439) for the 32 columns of one (env, ray group).
// On entry s_col[k] = {b1 | slow << 31, colour word} for column k of the span (k = 0 is the lowest
// address; ray r paints column R-1-r).  `item_slow`: some column of the span has a colour whose
// bytes differ (RGB8 only), so single-colour runs need the phase rotation of run16.
template <int FMT>
__device__ __forceinline__ void render_span(const FrameParams& p, const uint2* colinfo, uint8_t* env_obs,
                                            int B0, int ncols, int lane, bool item_slow) {
    const int CB = p.col_bytes;
    const uint32_t ceil_c = p.palette[RCW_COLOR_CEILING], floor_c = p.palette[RCW_COLOR_FLOOR];
    uint8_t* const span = env_obs + B0;
    ColumnBands cb;
    cb.ceiling = ceil_c;
    cb.floor = floor_c;

    if ((CB & 63) == 0) {
        // ---- whole sectors, mirror pairs -------------------------------------------------------
        // A column is symmetric: ceiling rows [0, pad) <-> floor rows [P-pad, P).  The sector at
        // column offset ob and its mirror at CB-32-ob are classified by one comparison against b1
        // (ceiling/floor, wall/wall, or both straddling a band boundary).  All 32 lanes sweep the
        // sector pairs of the span; a lane writes two sectors per iteration.  The sectors that
        // straddle a boundary (at most two per column) are composed afterwards, lane <-> column.
        const int HS = CB >> 6;                      // sector pairs per column
        const int n_sp = ncols * HS;
        const int n_iter = lane < n_sp ? ((n_sp - lane + 31) >> 5) : 0;
        int cl = (int)(((uint32_t)lane * p.unit_inv16) >> 16), hs = lane - cl * HS;
        const int adv_cl = p.unit_adv_cl, adv_hs = p.unit_adv_u;
        if (!item_slow) {
            const uint32_t ceil_w = PixelFormat<FMT>::flat_word(ceil_c), floor_w = PixelFormat<FMT>::flat_word(floor_c);
#pragma unroll kPairUnroll
            for (int it = 0; it < n_iter; ++it) {
                const uint2 info = colinfo[cl];
                const int b1 = (int)info.x, ob = hs << 5;
                const bool in_ceil = ob + 32 <= b1;
                uint8_t* const top = span + cl * CB + ob;
                if (in_ceil | (ob >= b1)) {
                    const uint32_t wt = in_ceil ? ceil_w : info.y, wb = in_ceil ? floor_w : info.y;
                    const uint4 vt = make_uint4(wt, wt, wt, wt), vb = make_uint4(wb, wb, wb, wb);
                    store_stream32(top, vt, vt);
                    store_stream32(top + (CB - 32 - 2 * ob), vb, vb);
                }
                cl += adv_cl;
                hs += adv_hs;
                if (hs >= HS) {
                    hs -= HS;
                    ++cl;
                }
            }
        } else {
#pragma unroll 1
            for (int it = 0; it < n_iter; ++it) {
                const uint2 info = colinfo[cl];
                const int b1 = (int)(info.x & 0x7FFFFFFFu), ob = hs << 5, mb = CB - 32 - ob;
                const bool in_ceil = ob + 32 <= b1;
                uint8_t* const top = span + cl * CB + ob;
                if (in_ceil | (ob >= b1)) {
                    const uint32_t ct = in_ceil ? ceil_c : info.y, cbm = in_ceil ? floor_c : info.y;
                    store_stream32(top, PixelFormat<FMT>::run16(ct, ob), PixelFormat<FMT>::run16(ct, ob + 16));
                    store_stream32(span + cl * CB + mb, PixelFormat<FMT>::run16(cbm, mb),
                                   PixelFormat<FMT>::run16(cbm, mb + 16));
                }
                cl += adv_cl;
                hs += adv_hs;
                if (hs >= HS) {
                    hs -= HS;
                    ++cl;
                }
            }
        }
        // lane <-> column: the sectors that contain a band boundary
        if (lane < ncols) {
            const uint2 info = colinfo[lane];
            cb.b1 = (int)(info.x & 0x7FFFFFFFu);
            cb.b2 = CB - cb.b1;
            cb.wall = FMT == RCW_OBS_RGB8 ? (info.y & 0x00FFFFFFu) : info.y;   // (whole-word formats keep all four bytes: GRAY16F)
            uint8_t* const col = span + lane * CB;
            if (cb.b1 & 31) {
                const int sa = cb.b1 & ~31, sb = cb.b2 & ~31;
                if (!item_slow) {
                    store_split32(col + sa, PixelFormat<FMT>::flat_word(ceil_c), info.y, cb.b1 - sa);
                    store_split32(col + sb, info.y, PixelFormat<FMT>::flat_word(floor_c), cb.b2 - sb);
                } else {
                    store_stream32(col + sa, compose16<FMT>(cb, sa), compose16<FMT>(cb, sa + 16));
                    store_stream32(col + sb, compose16<FMT>(cb, sb), compose16<FMT>(cb, sb + 16));
                }
            }
        }
        return;
    }

    // ---- any column size: whole sectors of a pitched column ---------------------------------------
    // Columns whose byte length is not a multiple of 32 are laid out with a pitch rounded up to 32
    // bytes (rcw_obs_layout), so that every column starts on a sector boundary and the renderer never
    // writes part of a sector.  A sector is ceiling, wall colour or floor (the padding behind the last
    // row counts as floor), or it contains a band boundary and is composed, lane <-> column, below.
    const int CP = p.col_pitch;
    const int NS = CP >> 5;                          // sectors per column
    const int n_sec = ncols * NS;
    const int n_iter = lane < n_sec ? ((n_sec - lane + 31) >> 5) : 0;
    int cl = (int)(((uint32_t)lane * p.unit_inv16) >> 16), sc = lane - cl * NS;
    const int adv_cl = p.unit_adv_cl, adv_sc = p.unit_adv_u;
    // the sectors of a pitched span are consecutive in memory: lane L writes sectors L, L + 32, ...
    uint8_t* dst = span + (lane << 5);
    if (!item_slow) {
        const uint32_t ceil_w = PixelFormat<FMT>::flat_word(ceil_c), floor_w = PixelFormat<FMT>::flat_word(floor_c);
#pragma unroll 2
        for (int it = 0; it < n_iter; ++it) {
            const uint2 info = colinfo[cl];
            const int b1 = (int)info.x, b2 = CB - b1, ob = sc << 5, oe = ob + 32;
            const bool in_ceil = oe <= b1, in_floor = ob >= b2, in_wall = (ob >= b1) & (oe <= b2);
            if (in_ceil | in_floor | in_wall) {
                const uint32_t w = in_ceil ? ceil_w : (in_floor ? floor_w : info.y);
                const uint4 v = make_uint4(w, w, w, w);
                store_stream32(dst, v, v);
            }
            dst += 1024;
            cl += adv_cl;
            sc += adv_sc;
            if (sc >= NS) {
                sc -= NS;
                ++cl;
            }
        }
    } else {
#pragma unroll 1
        for (int it = 0; it < n_iter; ++it) {
            const uint2 info = colinfo[cl];
            const int b1 = (int)(info.x & 0x7FFFFFFFu), b2 = CB - b1, ob = sc << 5;
            const bool in_ceil = ob + 32 <= b1, in_floor = ob >= b2, in_wall = (ob >= b1) & (ob + 32 <= b2);
            if (in_ceil | in_floor | in_wall) {
                const uint32_t c = in_ceil ? ceil_c : (in_floor ? floor_c : (info.y & 0x00FFFFFFu));
                store_stream32(dst, PixelFormat<FMT>::run16(c, ob), PixelFormat<FMT>::run16(c, ob + 16));
            }
            dst += 1024;
            cl += adv_cl;
            sc += adv_sc;
            if (sc >= NS) {
                sc -= NS;
                ++cl;
            }
        }
    }
    if (lane < ncols) {
        const uint2 info = colinfo[lane];
        cb.b1 = (int)(info.x & 0x7FFFFFFFu);
        cb.b2 = CB - cb.b1;
        cb.wall = FMT == RCW_OBS_RGB8 ? (info.y & 0x00FFFFFFu) : info.y;   // (whole-word formats keep all four bytes: GRAY16F)
        uint8_t* const col = span + lane * CP;
        const int sa = cb.b1 & ~31, sb = cb.b2 & ~31;
        if (!item_slow && sa != sb) {
            if (cb.b1 & 31) store_split32(col + sa, PixelFormat<FMT>::flat_word(ceil_c), info.y, cb.b1 - sa);
            if (cb.b2 & 31) store_split32(col + sb, info.y, PixelFormat<FMT>::flat_word(floor_c), cb.b2 - sb);
        } else {
            if (cb.b1 & 31) store_stream32(col + sa, compose16<FMT>(cb, sa), compose16<FMT>(cb, sa + 16));
            if ((cb.b2 & 31) && (sb != sa || !(cb.b1 & 31)))
                store_stream32(col + sb, compose16<FMT>(cb, sb), compose16<FMT>(cb, sb + 16));
        }
    }
}

// s_col entry of a column from its shade.  Single-colour vectors need no phase rotation when R == G == B
// (RGB8) or always (XRGB32, GRAY8); otherwise the column takes the slow (funnel-shift) path.  Both cases
// are worked out per palette entry on the host (FrameParams::col_entry).
template <int FMT>
__device__ __forceinline__ uint2 column_entry(const FrameParams& p, int pad, int cid, bool& slow) {
    const uint2 e = p.col_entry[cid];
    slow = FMT == RCW_OBS_RGB8 && (e.x != 0u);
    // bytes per pixel from the launch, not from FMT: RCW_OBS_GRAY16F (two-byte pixels, one 16-bit pattern repeated)
    // runs on the XRGB32 instantiations — to the renderer both are whole 32-bit words of one value
    return make_uint2((uint32_t)(pad * p.px_bytes) | e.x, e.y);
}

// item / gpe for the 32-bit work-item index (exact for every n: FrameParams::gpe_magic / gpe_shift)
__device__ __forceinline__ uint32_t div_gpe(const FrameParams& p, uint32_t n) {
    if (p.gpe == 1) return n;
    const uint32_t t = __umulhi(p.gpe_magic, n);
    return (((n - t) >> 1) + t) >> p.gpe_shift;
}

// STAGE: what one launch does for every (env, group of 32 rays) item, one warp per item
//   kStageFused : act! -> cast_rays! -> update_camera_view!           (one kernel per env-step)
//   kStageFront : act! -> cast_rays! -> {pad, colour} of every column written to p.col_info
//   kStagePaint : p.col_info -> update_camera_view! stores            (pure store stream)
// Front + Paint is the same step split in two launches: the latency-bound part runs without a
// saturated store queue in front of its loads, and the store stream runs without waiting on it.
// BULK (fused / paint): bands written as TMA bulk stores out of single-colour pattern buffers in
// shared memory (measured alternative, slower); otherwise the lanes write whole 32-byte sectors.
enum : int { kStageFused = 0, kStageFront = 1, kStagePaint = 2 };

// OCC = CTAs per SM the register allocation aims for.  20 warps per SM (91 registers) is best when the step
// is bound by the store stream (default camera, RGB8 / XRGB32); 32 warps per SM (<= 64 registers) is 6-14 %
// faster when act! and the DDA bound it (small frames, one-byte pixels, large maps) — profiles/README.md.
// ROOM: the wall layer is the border of the map and nothing else (RoomMap): nothing is staged in shared memory.
template <bool ROOM>
struct MapOf {
    using type = BitsMap;
    __device__ static __forceinline__ BitsMap make(const FrameParams& p, const uint32_t* words) {
        return BitsMap{words, p.n_extra ? words + (p.n_extra + 1) * p.map_words : words, p.wpr, p.n_extra, p.map_words};
    }
};
template <>
struct MapOf<true> {
    using type = RoomMap;
    __device__ static __forceinline__ RoomMap make(const FrameParams& p, const uint32_t*) { return RoomMap{p.H - 1, p.W - 1}; }
};

template <int MODE, int FMT, bool BULK, int STAGE, bool ROOM = false>
__device__ __forceinline__ void frame_body(const FrameParams& p, const PackedActions* pa) {
    extern __shared__ __align__(128) uint32_t s_dyn[];  // [pattern buffers (BULK)] [bit-packed wall layer]
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ __align__(8) uint64_t s_mbar_env[kWarpsPerCta];   // per-env wall layers: one per env slot
    __shared__ uint2 s_col[kWarpsPerCta][32];            // per column of the warp: {b1 | slow << 31, colour word}
    __shared__ EnvPose s_env[2][kWarpsPerCta];           // poses after act!, one slot per env of the round

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr bool kPaints = STAGE != kStageFront && MODE != kModeRays;
    constexpr bool kCasts = STAGE != kStagePaint;
    constexpr bool kStagesMap = kCasts && !ROOM;         // the DDA / act! read a bit-packed wall layer from shared memory

    // ---- stage the wall layer (and the pattern buffers): TMA bulk copies, completion on mbarriers.
    // A wall layer shared by the batch is copied once per CTA here; per-env wall layers
    // (map_env_stride != 0) are copied per env of the round by that env's first warp, below.
    const uint32_t pat_bytes = (BULK && kPaints) ? 6u * (uint32_t)p.pat_stride : 0u;
    uint32_t* const s_map = s_dyn + pat_bytes / 4;
    const bool per_env_maps = kStagesMap && p.map_env_stride != 0;
    const uint32_t map_bytes = (uint32_t)p.stage_words * 4u;   // every layer of the map (one without extra objects)
    const uint32_t shared_bytes = (kStagesMap && !per_env_maps) ? map_bytes : 0u;
    if (kStagesMap || pat_bytes) {
        if (threadIdx.x == 0) {
            mbar_init(&s_mbar, 1);
            if (per_env_maps)
                for (int k = 0; k < kWarpsPerCta; ++k) mbar_init(&s_mbar_env[k], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0 && shared_bytes + pat_bytes) {
            mbar_arrive_expect_tx(&s_mbar, shared_bytes + pat_bytes);
            if (shared_bytes) bulk_copy_g2s(s_map, p.wall_map, shared_bytes, &s_mbar);
            if (pat_bytes) bulk_copy_g2s(s_dyn, p.patterns, pat_bytes, &s_mbar);
        }
    }
    bool staged = (shared_bytes + pat_bytes) == 0;   // the wait is deferred until shared memory is first needed
    const uint32_t s_pat = smem_u32(s_dyn);

    const int R = p.R, CB = p.col_bytes;
    const uint32_t gpe = (uint32_t)p.gpe;
    const uint32_t n_items = (uint32_t)p.env_count * gpe;
    const uint32_t round_stride = gridDim.x * kWarpsPerCta;

    uint32_t parity = 0;
    for (uint32_t base = blockIdx.x * kWarpsPerCta; base < n_items; base += round_stride, parity ^= 1u) {
        const uint32_t item = base + warp;
        const bool item_ok = item < n_items;
        const uint32_t env_rel = div_gpe(p, item_ok ? item : n_items - 1);
        const int g = (int)(item - env_rel * gpe);
        const int64_t env = p.env_first + env_rel;
        const int r0 = g * 32;
        const int ncols = min(32, R - r0);
        const int col0 = R - r0 - ncols;               // first column of the span (ray r paints column R-1-r)

        // a masked render (after a masked reset) redraws only the chosen envs; the decision is per env, so
        // all warps of an env — including the one that stages its wall layer — agree
        if (MODE == kModeRender && STAGE != kStagePaint && p.render_mask && !__ldg(p.render_mask + env)) continue;

        // the env's slot in the observation buffer (a window of obs_window envs; the whole batch by default)
        uint32_t obs_slot = p.obs_slot0 + env_rel;
        if (obs_slot >= p.obs_window) obs_slot -= p.obs_window;

        ColumnShade cs;
        cs.pad = 0;
        cs.cid = RCW_COLOR_WALL_1;
        if (kCasts) {
            EnvPose pose;
            float4 rt;
            // lanes past the last ray shadow the last ray (same walk, nothing stored)
            const float4* const rt_lane = p.ray_table + min(r0 + lane, R - 1);
            const uint32_t slot = env_rel - div_gpe(p, base);    // envs of this round, in order
            const bool leader = item_ok && (warp == 0 || g == 0);
            const uint32_t* my_map = s_map;
            if (per_env_maps) {
                // this env's wall layer -> its slot (one round per CTA in this mode, see grid_for)
                my_map = s_map + slot * (uint32_t)p.stage_words;
                if (leader && lane == 0) {
                    mbar_arrive_expect_tx(&s_mbar_env[slot], map_bytes);
                    bulk_copy_g2s(const_cast<uint32_t*>(my_map), p.wall_map + (size_t)env * p.map_env_stride,
                                  map_bytes, &s_mbar_env[slot]);
                }
            }
            if (MODE == kModeStep) {
#if RCW_EXP != 1
                // ---- act!: once per env of the round, by the first warp of the CTA that works on it.
                // Loads are issued before the prologue wait; every warp meanwhile fetches the ray-table
                // rows of the three directions the env can face after this step (turn right / keep /
                // turn left), so the DDA does not start with a dependent L2 round trip.
                EnvInputs in;
                if (leader) in = load_env_inputs(p, env, pa, env_rel);
                const int au_in = __ldg(p.in.dir_au + env);
                const int au_m = au_in == 0 ? p.N - 1 : au_in - 1, au_p = au_in + 1 == p.N ? 0 : au_in + 1;
                const float4 rt_m = __ldg(rt_lane + (size_t)au_m * (size_t)R);
                const float4 rt_0 = __ldg(rt_lane + (size_t)au_in * (size_t)R);
                const float4 rt_p = __ldg(rt_lane + (size_t)au_p * (size_t)R);
                if (!staged) {
                    mbar_wait(&s_mbar, 0);
                    staged = true;
                }
                if (per_env_maps) mbar_wait(&s_mbar_env[slot], 0);
                if (leader) {
                    pose = act_env(p, MapOf<ROOM>::make(p, my_map), env, in, /*writer=*/g == 0, lane);
                    if (lane == 0) s_env[parity][slot] = pose;
                }
                __syncthreads();
                pose = s_env[parity][slot];
                if (pose.au == au_in) rt = rt_0;
                else if (pose.au == au_m) rt = rt_m;
                else if (pose.au == au_p) rt = rt_p;
                else rt = __ldg(rt_lane + (size_t)pose.au * (size_t)R);   // auto-reset drew a new direction
#else
                pose.x = 2.5f; pose.y = 2.5f; pose.au = 3; pose.goal = 0x00050005u;
                rt = __ldg(rt_lane + (size_t)pose.au * (size_t)R);
#endif
            } else {
                pose.x = __ldg(p.in.pos_x + env);
                pose.y = __ldg(p.in.pos_y + env);
                pose.au = __ldg(p.in.dir_au + env);
                pose.goal = __ldg(p.in.goal + env);
                rt = __ldg(rt_lane + (size_t)pose.au * (size_t)R);
                if (!staged) {
                    mbar_wait(&s_mbar, 0);
                    staged = true;
                }
                if (per_env_maps) mbar_wait(&s_mbar_env[slot], 0);
            }
            if (!item_ok) continue;   // (no block barrier below this point)
            cs = cast_and_shade<MODE>(p, MapOf<ROOM>::make(p, my_map), pose, rt, g, lane, env_rel);
            if (MODE == kModeRays) continue;
            if (STAGE == kStageFront) {
                // column order, so the paint launch reads its 32 columns with one coalesced load
                if (lane < ncols)
                    p.col_info[(size_t)obs_slot * p.col_info_stride + (size_t)(col0 + ncols - 1 - lane)] =
                        (uint32_t)cs.pad | ((uint32_t)cs.cid << 16);
                continue;
            }
        } else {
            if (!item_ok) continue;
            if (!staged) {
                mbar_wait(&s_mbar, 0);
                staged = true;
            }
        }

        uint8_t* const env_obs = p.obs + (size_t)obs_slot * p.obs_env_stride;
        const int B0 = col0 * p.col_pitch;             // byte span of the warp's columns in the env image
        int my_col = ncols - 1 - lane;                 // span column of this lane's ray
        if (STAGE == kStagePaint) {
            my_col = lane;                             // the info array is already in column order
            if (lane < ncols) {
                const uint32_t info = __ldg(p.col_info + (size_t)obs_slot * p.col_info_stride + (size_t)(col0 + lane));
                cs.pad = (int)(info & 0xFFFFu);
                cs.cid = (int)(info >> 16);
                // the words may come from a caller's replay buffer: a stale or uninitialised row must not index
                // outside the palette or paint outside its column — clamp, and leave a sticky flag for the host
                const int max_cid = RCW_COLOR_GOAL_2 + 2 * p.n_extra;
                const bool bad = (cs.cid < RCW_COLOR_WALL_1) | (cs.cid > max_cid) | (cs.pad > (p.P >> 1));
                if (bad) {
                    cs.cid = min(max(cs.cid, (int)RCW_COLOR_WALL_1), max_cid);
                    cs.pad = min(cs.pad, p.P >> 1);
                    if (p.stats) atomicExch(&p.stats->bad_columns, 1);
                }
            }
        }

        if (BULK) {
            // lane <-> one column: three TMA bulk stores out of the single-colour pattern buffers
            // + the boundary vectors
            if (lane < ncols) {
                ColumnBands cb;
                cb.ceiling = p.palette[RCW_COLOR_CEILING];
                cb.floor = p.palette[RCW_COLOR_FLOOR];
                cb.b1 = cs.pad * PixelFormat<FMT>::kBpp;
                cb.b2 = CB - cb.b1;
                cb.wall = p.palette[cs.cid];
                const int S = B0 + my_col * p.col_pitch;
                column_bulk<FMT>(env_obs, S, CB, cb, cs.cid, s_pat, p.pat_stride);
                column_edges<FMT>(env_obs, S, CB, cb);
            }
            bulk_commit_group();
            continue;
        }
        bool slow = false;
        if (lane < ncols) s_col[warp][my_col] = column_entry<FMT>(p, cs.pad, cs.cid, slow);
        __syncwarp();
        const bool item_slow = __any_sync(0xFFFFFFFFu, slow);
        render_span<FMT>(p, s_col[warp], env_obs, B0, ncols, lane, item_slow);
        __syncwarp();   // s_col is rewritten in the next round
    }
    if (BULK && kPaints) bulk_wait_group_read0();   // shared memory must outlive the TMA reads
    // a thread that never needed the staged wall layer (all its items skipped or out of range) still waits for
    // the bulk copy: shared memory must not be released while the copy is in flight
    if (!staged) mbar_wait(&s_mbar, 0);
}

// OCC = CTAs per SM the register allocation aims for (see above).
template <int MODE, int FMT, bool BULK, int STAGE, int OCC, bool ROOM = false>
__global__ void __launch_bounds__(kThreadsPerCta, OCC)
frame_kernel(const __grid_constant__ FrameParams p) {
    frame_body<MODE, FMT, BULK, STAGE, ROOM>(p, nullptr);
}

// the step (fused, or the front stage alone: RCW_OBS_COLUMNS) with host-supplied actions packed into the parameters
template <int FMT, int OCC, int STAGE = kStageFused, bool ROOM = false>
__global__ void __launch_bounds__(kThreadsPerCta, OCC)
frame_kernel_pa(const __grid_constant__ FrameParams p, const __grid_constant__ PackedActions a) {
    frame_body<kModeStep, FMT, false, STAGE, ROOM>(p, &a);
}

// ------------------------------------------------------------------------------------------
// small items: one warp = one env
// ------------------------------------------------------------------------------------------
// When an item (32 columns) is small — few rows or one byte per pixel — the step is bound by act! and the DDA,
// not by the stores; the CTA barrier behind the env's leader warp then costs a fifth of the time
// (profiles/README.md).  Here a warp owns a whole env: act! once, then its groups one after the other, the
// ray-table row of the next group in flight while the current one is cast and painted.  No block barrier,
// no pose exchange through shared memory, no item -> (env, group) arithmetic.
// OUT: what a warp does with the 32 columns of a ray group.
//   kOutPaint : render_span, as in frame_kernel (any geometry)
//   kOutWords : the observation is the column words themselves (RCW_OBS_COLUMNS): nothing is painted
//   kOutTable : small columns: every possible column (rows of ceiling x colour) is kept ready-made in a table of
//               (P / 2 + 1) x 4 pitched columns (FrameParams::col_table, built on the host with the renderer's
//               rules, L1 / L2 resident), and painting a column is copying its col_pitch bytes, whole sectors,
//               256 bits per load and store — about 12 instead of 200 instructions per ray group.  The format
//               only decides what is in the table, so one instantiation serves RGB8, XRGB32 and GRAY8.
// ROOM: see MapOf — no wall layer in shared memory, no TMA, no mbarrier, no CTA barrier at all.
//   kOutHalf  : RCW_OBS_GRAY8_HALF — the GRAY8 frame under a 2 x 2 box filter, composed from the column decisions of
//               the two rays of every output column; the full-resolution frame is never written
enum : int { kOutPaint = 0, kOutWords = 1, kOutTable = 2, kOutHalf = 3 };

__device__ __forceinline__ void load_nc32(const uint8_t* p, uint4& lo, uint4& hi) {
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
}

template <int MODE, int FMT, int OUT = kOutPaint, bool ROOM = false>
__device__ __forceinline__ void env_body(const FrameParams& p, const PackedActions* pa) {
    extern __shared__ __align__(128) uint32_t s_dyn[];   // wall layer(s): one shared, or one slot per warp
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ __align__(8) uint64_t s_mbar_env[kWarpsPerCta];
    __shared__ uint2 s_col[OUT == kOutPaint ? kWarpsPerCta : 1][32];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool per_env_maps = !ROOM && p.map_env_stride != 0;
    const uint32_t map_bytes = (uint32_t)p.stage_words * 4u;   // every layer of the map (one without extra objects)
    if (!ROOM) {
        if (threadIdx.x == 0) {
            mbar_init(&s_mbar, 1);
            if (per_env_maps)
                for (int k = 0; k < kWarpsPerCta; ++k) mbar_init(&s_mbar_env[k], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0 && !per_env_maps) {
            mbar_arrive_expect_tx(&s_mbar, map_bytes);
            bulk_copy_g2s(s_dyn, p.wall_map, map_bytes, &s_mbar);
        }
    }
    // (no block barrier below this point)  A warp without work still waits for the CTA's bulk copy, so that shared
    // memory is not released while the copy is in flight.
    const uint32_t env_rel = blockIdx.x * kWarpsPerCta + warp;
    const int64_t env = p.env_first + env_rel;
    if (env_rel >= (uint32_t)p.env_count ||
        (MODE == kModeRender && p.render_mask && !__ldg(p.render_mask + env))) {       // out of range / masked render
        if (!ROOM && !per_env_maps) mbar_wait(&s_mbar, 0);
        return;
    }
    const int R = p.R;
    const uint32_t* my_map = s_dyn;
    if (per_env_maps) {
        my_map = s_dyn + (uint32_t)warp * (uint32_t)p.stage_words;
        if (lane == 0) {
            mbar_arrive_expect_tx(&s_mbar_env[warp], map_bytes);
            bulk_copy_g2s(const_cast<uint32_t*>(my_map), p.wall_map + (size_t)env * p.map_env_stride, map_bytes,
                          &s_mbar_env[warp]);
        }
    }
    const typename MapOf<ROOM>::type map = MapOf<ROOM>::make(p, my_map);
    const float4* const rt_lane0 = p.ray_table + min(lane, R - 1);     // group 0
    EnvPose pose;
    float4 rt;
    if (MODE == kModeStep) {
        // loads first; the ray-table rows of the three directions the env can face after this step
        const EnvInputs in = load_env_inputs(p, env, pa, env_rel);
        const int au_m = in.au == 0 ? p.N - 1 : in.au - 1, au_p = in.au + 1 == p.N ? 0 : in.au + 1;
        const float4 rt_m = __ldg(rt_lane0 + (size_t)au_m * (size_t)R);
        const float4 rt_0 = __ldg(rt_lane0 + (size_t)in.au * (size_t)R);
        const float4 rt_p = __ldg(rt_lane0 + (size_t)au_p * (size_t)R);
        if (!ROOM) {
            if (per_env_maps) mbar_wait(&s_mbar_env[warp], 0);
            else mbar_wait(&s_mbar, 0);
        }
        pose = act_env(p, map, env, in, /*writer=*/true, lane);
        if (pose.au == in.au) rt = rt_0;
        else if (pose.au == au_m) rt = rt_m;
        else if (pose.au == au_p) rt = rt_p;
        else rt = __ldg(rt_lane0 + (size_t)pose.au * (size_t)R);       // auto-reset drew a new direction
    } else {
        pose.x = __ldg(p.in.pos_x + env);
        pose.y = __ldg(p.in.pos_y + env);
        pose.au = __ldg(p.in.dir_au + env);
        pose.goal = __ldg(p.in.goal + env);
        rt = __ldg(rt_lane0 + (size_t)pose.au * (size_t)R);
        if (!ROOM) {
            if (per_env_maps) mbar_wait(&s_mbar_env[warp], 0);
            else mbar_wait(&s_mbar, 0);
        }
    }
    uint32_t obs_slot = p.obs_slot0 + env_rel;
    if (obs_slot >= p.obs_window) obs_slot -= p.obs_window;
    uint8_t* const env_obs = p.obs + (size_t)obs_slot * p.obs_env_stride;
    const float4* const rt_row = p.ray_table + (size_t)pose.au * (size_t)R;
    const int gpe = p.gpe;
    for (int g = 0; g < gpe; ++g) {
        // the next group's row (behind the last group: the last ray's entry once more, unused)
        const float4 rt_next = __ldg(rt_row + min((g + 1) * 32 + lane, R - 1));
        const ColumnShade cs = cast_and_shade<MODE>(p, map, pose, rt, g, lane, env_rel);
        const int r0 = g * 32;
        const int ncols = min(32, R - r0);
        const int col0 = R - r0 - ncols;               // ray r paints column R-1-r
        if (OUT == kOutWords) {
            if (lane < ncols)
                p.col_info[(size_t)obs_slot * p.col_info_stride + (size_t)(col0 + ncols - 1 - lane)] =
                    (uint32_t)cs.pad | ((uint32_t)cs.cid << 16);
            rt = rt_next;
            continue;
        }
        if (OUT == kOutTable) {
#if RCW_TABLE_COPY == 0
            // lane <-> column: copy the column's col_pitch bytes from its table entry, sector by sector
            const int CP = p.col_pitch;
            if (lane < ncols) {
                const uint8_t* src = p.col_table + (uint32_t)(cs.pad * p.n_colors + (cs.cid - RCW_COLOR_WALL_1)) * (uint32_t)CP;
                uint8_t* dst = env_obs + (size_t)(col0 + ncols - 1 - lane) * CP;
#pragma unroll 3
                for (int o = 0; o < CP; o += 32) {
                    uint4 lo, hi;
                    load_nc32(src + o, lo, hi);
                    store_stream32(dst + o, lo, hi);
                }
            }
#else
            // The span's sectors are consecutive in memory: lane L copies sectors L, L + 32, ... — sector s is
            // sector (s mod NS) of the table entry of column s / NS, whose ray sits in lane ncols - 1 - s / NS.
            const int CP = p.col_pitch, NS = CP >> 5;
            const uint32_t entry = (uint32_t)(cs.pad * p.n_colors + (cs.cid - RCW_COLOR_WALL_1)) * (uint32_t)CP;
            const int n_sec = ncols * NS;
            int cl = (int)(((uint32_t)lane * p.sec_inv16) >> 16), sc = lane - cl * NS;
            const int adv_cl = p.sec_adv_cl, adv_sc = p.sec_adv_u;
            uint8_t* dst = env_obs + (size_t)col0 * CP + (lane << 5);
#pragma unroll 2
            for (int s = lane; s < ((n_sec + 31) & ~31); s += 32) {   // (uniform trip count: the shuffle needs all lanes)
                const uint32_t e = __shfl_sync(0xFFFFFFFFu, entry, max(ncols - 1 - cl, 0));
                if (s < n_sec) {
                    uint4 lo, hi;
                    load_nc32(p.col_table + e + (uint32_t)(sc << 5), lo, hi);
                    store_stream32(dst, lo, hi);
                }
                dst += 1024;
                cl += adv_cl;
                sc += adv_sc;
                if (sc >= NS) {
                    sc -= NS;
                    ++cl;
                }
            }
#endif
            rt = rt_next;
            continue;
        }
        if (OUT == kOutHalf) {
            // Output column c covers image columns 2 c and 2 c + 1, i.e. (ray i paints column R - i, :431, and R is even)
            // the rays of one even / odd lane pair; output row r covers rows 2 r and 2 r + 1.  A column of the full frame
            // is ceiling for rows < pad, the hit object's luma for pad <= row < P - pad, floor below (:433-439), so a
            // 2 x 2 block is uniform almost everywhere: lane L takes the 32-byte sectors L, L + 32, ... of the span's
            // consecutive output columns, fetches the two rays' {pad, luma} with two shuffles, and only where a block
            // row straddles a band boundary of either ray is the box filter evaluated pixel by pixel.
            const int P = p.P, CP = p.col_pitch, NS = CP >> 5;        // pitch / sectors of an OUTPUT column (P / 2 rows)
            const uint32_t mine = (uint32_t)cs.pad | ((p.palette[cs.cid] & 0xFFu) << 16);
            const uint32_t C = p.palette[RCW_COLOR_CEILING] & 0xFFu, F = p.palette[RCW_COLOR_FLOOR] & 0xFFu;
            const int n_sec = (ncols >> 1) * NS;
            int cl = (int)(((uint32_t)lane * p.sec_inv16) >> 16), sc = lane - cl * NS;
            const int adv_cl = p.sec_adv_cl, adv_sc = p.sec_adv_u;
            uint8_t* dst = env_obs + (size_t)(col0 >> 1) * CP + (lane << 5);
#pragma unroll 1
            for (int s = lane; s < ((n_sec + 31) & ~31); s += 32) {   // (uniform trip count: the shuffles need all lanes)
                const int la = max(ncols - 2 - 2 * cl, 0);            // output column cl of the span <- rays of lanes la, la + 1
                const uint32_t ia = __shfl_sync(0xFFFFFFFFu, mine, la), ib = __shfl_sync(0xFFFFFFFFu, mine, la + 1);
                if (s < n_sec) {
                    const int pa = (int)(ia & 0xFFFFu), pb = (int)(ib & 0xFFFFu);
                    const uint32_t Wa = ia >> 16, Wb = ib >> 16;
                    // Output rows, P2 = P / 2 of them.  Ray a's two columns of a block row are both ceiling above row
                    // pa >> 1 and both object from row (pa + 1) >> 1 on (a block row in between, for an odd pad, mixes
                    // them), mirrored at the bottom.  So the output column is ceiling above T0, the two objects' mean
                    // between T1 and M0, floor from M1 on, and only the few rows of [T0, T1) and [M0, M1) — as many as
                    // the two rays' pads differ, about one — need the filter evaluated: the sector is written as the
                    // three flat bands and those rows are then stored over it, byte by byte (same lane, program order).
                    const int P2 = P >> 1;
                    const int T0 = min(pa, pb) >> 1, T1 = (max(pa, pb) + 1) >> 1, M0 = P2 - T1, M1 = P2 - T0;
                    const uint32_t mid = ((2u * Wa + 2u * Wb + 2u) >> 2) * 0x01010101u;
                    const uint32_t Cw = C * 0x01010101u, Fw = F * 0x01010101u;
                    const int q0 = sc << 5;                           // first output row of this sector
                    uint32_t w[8];
                    if (q0 + 32 <= T0 || q0 >= M1 || (q0 >= T0 && q0 + 32 <= M1)) {
                        const uint32_t v = q0 + 32 <= T0 ? Cw : (q0 >= M1 ? Fw : mid);
#pragma unroll
                        for (int k = 0; k < 8; ++k) w[k] = v;
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint32_t m1 = low_bytes_mask(T0 - q0 - 4 * k), m2 = low_bytes_mask(M1 - q0 - 4 * k);
                            w[k] = (Cw & m1) | (((mid & m2) | (Fw & ~m2)) & ~m1);
                        }
                    }
                    store_stream32(dst, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
                    auto pixel = [&](int pad, uint32_t W, int row) { return row < pad ? C : (row < P - pad ? W : F); };
                    auto patch = [&](int lo, int hi) {                // output rows [lo, hi) of this sector, filtered
                        for (int r = max(lo, q0); r < min(hi, q0 + 32); ++r) {
                            const uint32_t sum = pixel(pa, Wa, 2 * r) + pixel(pa, Wa, 2 * r + 1) + pixel(pb, Wb, 2 * r) +
                                                 pixel(pb, Wb, 2 * r + 1);
                            dst[r - q0] = (uint8_t)((sum + 2u) >> 2);
                        }
                    };
                    patch(T0, T1);
                    patch(M0, M1);
                }
                dst += 1024;
                cl += adv_cl;
                sc += adv_sc;
                if (sc >= NS) {
                    sc -= NS;
                    ++cl;
                }
            }
            rt = rt_next;
            continue;
        }
        bool slow = false;
        if (lane < ncols) s_col[OUT == kOutPaint ? warp : 0][ncols - 1 - lane] = column_entry<FMT>(p, cs.pad, cs.cid, slow);
        __syncwarp();
        const bool item_slow = __any_sync(0xFFFFFFFFu, slow);
        render_span<FMT>(p, s_col[OUT == kOutPaint ? warp : 0], env_obs, col0 * p.col_pitch, ncols, lane, item_slow);
        __syncwarp();                                  // s_col is rewritten by the next group
        rt = rt_next;
    }
}

template <int MODE, int FMT, int OUT = kOutPaint, bool ROOM = false>
__global__ void __launch_bounds__(kThreadsPerCta, kCtasPerSmHi)
env_kernel(const __grid_constant__ FrameParams p) {
    env_body<MODE, FMT, OUT, ROOM>(p, nullptr);
}

template <int FMT, int OUT = kOutPaint, bool ROOM = false>
__global__ void __launch_bounds__(kThreadsPerCta, kCtasPerSmHi)
env_kernel_pa(const __grid_constant__ FrameParams p, const __grid_constant__ PackedActions a) {
    env_body<kModeStep, FMT, OUT, ROOM>(p, &a);
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
        const uint32_t* const words = p.wall_map + (size_t)env * p.map_env_stride;
        draw_layout(BitsMap{words, p.n_extra ? words + (p.n_extra + 1) * p.map_words : words, p.wpr, p.n_extra, p.map_words},
                    p.H, p.W, p.N, p.seed,
                    p.env_id_offset + (uint64_t)env, episode, gi, gj, pi, pj, au);
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

// After a step of the env range [env0, env0 + n) only (rcw_step_range): copy the range's new state
// back into the buffer the step read, so that the handle's current state stays in one buffer.
__global__ void commit_range_kernel(const StateRef from, const StateRef to, int64_t env0, int64_t n) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t env = env0 + k;
    to.pos_x[env] = from.pos_x[env];
    to.pos_y[env] = from.pos_y[env];
    to.dir_au[env] = from.dir_au[env];
    to.goal[env] = from.goal[env];
    to.episode[env] = from.episode[env];
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
    const int n = N * R;   // rcw_create rejects N * R >= 2^31
    build_ray_table_kernel<<<(n + 255) / 256, 256, 0, s>>>(dirs, N, R, sfov, table);
    return cudaGetLastError();
}

template <int MODE, int FMT, bool BULK, int STAGE, int OCC, bool ROOM = false>
static cudaError_t launch_frame_t(const FrameParams& p, int ctas, cudaStream_t s) {
    const bool casts = STAGE != kStagePaint, paints = STAGE != kStageFront && MODE != kModeRays;
    const size_t map_slots = p.map_env_stride ? kWarpsPerCta : 1;   // per-env wall layers: one slot per env of a round
    const size_t smem = ((casts && !ROOM) ? map_slots * (size_t)p.stage_words * 4 : 0) + ((BULK && paints) ? 6 * (size_t)p.pat_stride : 0);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(frame_kernel<MODE, FMT, BULK, STAGE, OCC, ROOM>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    frame_kernel<MODE, FMT, BULK, STAGE, OCC, ROOM><<<ctas, kThreadsPerCta, smem, s>>>(p);
    return cudaGetLastError();
}

static size_t env_smem(const FrameParams& p, bool room) {
    return room ? 0 : (p.map_env_stride ? kWarpsPerCta : 1) * (size_t)p.stage_words * 4;
}

template <int MODE, int FMT, int OUT, bool ROOM>
static cudaError_t launch_env_t(const FrameParams& p, cudaStream_t s) {
    const size_t smem = env_smem(p, ROOM);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(env_kernel<MODE, FMT, OUT, ROOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const unsigned ctas = (unsigned)((p.env_count + kWarpsPerCta - 1) / kWarpsPerCta);
    env_kernel<MODE, FMT, OUT, ROOM><<<ctas, kThreadsPerCta, smem, s>>>(p);
    return cudaGetLastError();
}

template <int FMT, int OUT, bool ROOM>
static cudaError_t launch_env_pa_t(const FrameParams& p, const PackedActions& pa, cudaStream_t s) {
    const size_t smem = env_smem(p, ROOM);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(env_kernel_pa<FMT, OUT, ROOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const unsigned ctas = (unsigned)((p.env_count + kWarpsPerCta - 1) / kWarpsPerCta);
    env_kernel_pa<FMT, OUT, ROOM><<<ctas, kThreadsPerCta, smem, s>>>(p, pa);
    return cudaGetLastError();
}

template <int FMT, int OCC, int STAGE, bool ROOM>
static cudaError_t launch_frame_pa_t(const FrameParams& p, const PackedActions& pa, int ctas, cudaStream_t s) {
    const size_t smem = env_smem(p, ROOM);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(frame_kernel_pa<FMT, OCC, STAGE, ROOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    frame_kernel_pa<FMT, OCC, STAGE, ROOM><<<ctas, kThreadsPerCta, smem, s>>>(p, pa);
    return cudaGetLastError();
}

constexpr int kOcc = kCtasPerSmLo;

// env_kernel: MODE in {step, render}; OUT / FMT from the handle's format and whether it has a column table
template <int MODE, bool ROOM>
static cudaError_t launch_env_m(const FrameParams& p, int obs_format, const LaunchShape& sh, cudaStream_t s,
                                const PackedActions* packed) {
    // (a packed launch is always a step: `packed` is only passed with MODE == kModeStep)
#define RCW_ENV_LAUNCH(FMT, OUT)                                                        \
    return packed ? launch_env_pa_t<FMT, OUT, ROOM>(p, *packed, s) : launch_env_t<MODE, FMT, OUT, ROOM>(p, s)
    if (obs_format == RCW_OBS_COLUMNS) { RCW_ENV_LAUNCH(RCW_OBS_RGB8, kOutWords); }
    if (obs_format == RCW_OBS_GRAY8_HALF) { RCW_ENV_LAUNCH(RCW_OBS_GRAY8, kOutHalf); }
    if (sh.table) { RCW_ENV_LAUNCH(RCW_OBS_GRAY8, kOutTable); }
    if (obs_format == RCW_OBS_GRAY8) { RCW_ENV_LAUNCH(RCW_OBS_GRAY8, kOutPaint); }
    if (obs_format == RCW_OBS_RGB8) { RCW_ENV_LAUNCH(RCW_OBS_RGB8, kOutPaint); }
    RCW_ENV_LAUNCH(RCW_OBS_XRGB32, kOutPaint);   // XRGB32 and GRAY16F (see column_entry)
#undef RCW_ENV_LAUNCH
}

// frame_kernel, the shipped path (fused stage or the front stage alone, lane-written sectors): both register budgets
template <int MODE, int FMT, int STAGE, bool ROOM>
static cudaError_t launch_item(const FrameParams& p, const LaunchShape& sh, cudaStream_t s, const PackedActions* packed) {
    const bool hi = sh.occ4 || STAGE == kStageFront;   // the front stage alone is bound by act! + DDA
    if (packed)
        return hi ? launch_frame_pa_t<FMT, kCtasPerSmHi, STAGE, ROOM>(p, *packed, sh.ctas, s)
                  : launch_frame_pa_t<FMT, kOcc, STAGE, ROOM>(p, *packed, sh.ctas, s);
    return hi ? launch_frame_t<MODE, FMT, false, STAGE, kCtasPerSmHi, ROOM>(p, sh.ctas, s)
              : launch_frame_t<MODE, FMT, false, STAGE, kOcc, ROOM>(p, sh.ctas, s);
}

template <int MODE, bool ROOM>
static cudaError_t launch_shipped(const FrameParams& p, int obs_format, const LaunchShape& sh, cudaStream_t s,
                                  const PackedActions* packed) {
    if (sh.env_per_warp) return launch_env_m<MODE, ROOM>(p, obs_format, sh, s, packed);
    if (obs_format == RCW_OBS_GRAY8_HALF) return cudaErrorInvalidValue;   // (only env_kernel composes the box filter)
    if (obs_format == RCW_OBS_COLUMNS) return launch_item<MODE, RCW_OBS_RGB8, kStageFront, ROOM>(p, sh, s, packed);
    if (obs_format == RCW_OBS_GRAY8) return launch_item<MODE, RCW_OBS_GRAY8, kStageFused, ROOM>(p, sh, s, packed);
    if (obs_format == RCW_OBS_RGB8) return launch_item<MODE, RCW_OBS_RGB8, kStageFused, ROOM>(p, sh, s, packed);
    return launch_item<MODE, RCW_OBS_XRGB32, kStageFused, ROOM>(p, sh, s, packed);
}

// measured alternatives (RCW_RENDER_PATH=bulk, RCW_SPLIT=1) and the paint stage of rcw_expand_columns: bit-packed maps only
template <int MODE, int STAGE>
static cudaError_t launch_frame_alt(const FrameParams& p, int obs_format, const LaunchShape& sh, cudaStream_t s) {
    if (obs_format == RCW_OBS_GRAY8) return launch_frame_t<MODE, RCW_OBS_GRAY8, false, STAGE, kOcc>(p, sh.ctas, s);
    if (obs_format == RCW_OBS_RGB8)
        return sh.bulk ? launch_frame_t<MODE, RCW_OBS_RGB8, true, STAGE, kOcc>(p, sh.ctas, s)
                       : launch_frame_t<MODE, RCW_OBS_RGB8, false, STAGE, kOcc>(p, sh.ctas, s);
    return sh.bulk ? launch_frame_t<MODE, RCW_OBS_XRGB32, true, STAGE, kOcc>(p, sh.ctas, s)
                   : launch_frame_t<MODE, RCW_OBS_XRGB32, false, STAGE, kOcc>(p, sh.ctas, s);
}

// split = false: one fused launch.  split = true: two launches (front, then paint) through p.col_info.
cudaError_t launch_frame(const FrameParams& p, int mode, int obs_format, const LaunchShape& sh, cudaStream_t s,
                         const PackedActions* packed) {
    if (packed && (mode != kModeStep || p.env_count > kPackedActionEnvs)) return cudaErrorInvalidValue;
    if (mode == kModeRays)
        return sh.room ? launch_frame_t<kModeRays, RCW_OBS_RGB8, false, kStageFused, kOcc, true>(p, sh.ctas, s)
                       : launch_frame_t<kModeRays, RCW_OBS_RGB8, false, kStageFused, kOcc, false>(p, sh.ctas, s);
    if (mode != kModeStep && mode != kModeRender) return cudaErrorInvalidValue;
    const bool alt = obs_format != RCW_OBS_COLUMNS && (sh.split || sh.bulk);
    if (!alt) {
        if (mode == kModeStep)
            return sh.room ? launch_shipped<kModeStep, true>(p, obs_format, sh, s, packed)
                           : launch_shipped<kModeStep, false>(p, obs_format, sh, s, packed);
        return sh.room ? launch_shipped<kModeRender, true>(p, obs_format, sh, s, nullptr)
                       : launch_shipped<kModeRender, false>(p, obs_format, sh, s, nullptr);
    }
    if (packed) return cudaErrorInvalidValue;
    if (!sh.split)
        return mode == kModeStep ? launch_frame_alt<kModeStep, kStageFused>(p, obs_format, sh, s)
                                 : launch_frame_alt<kModeRender, kStageFused>(p, obs_format, sh, s);
    cudaError_t e = mode == kModeStep
                        ? launch_frame_t<kModeStep, RCW_OBS_RGB8, false, kStageFront, kOcc>(p, sh.ctas, s)
                        : launch_frame_t<kModeRender, RCW_OBS_RGB8, false, kStageFront, kOcc>(p, sh.ctas, s);
    if (e != cudaSuccess) return e;
    return launch_frame_alt<kModeRender, kStagePaint>(p, obs_format, sh, s);
}

cudaError_t launch_expand_columns(const FrameParams& p, int pixel_format, int ctas, cudaStream_t s) {
    LaunchShape sh{false, true, false, ctas};
    if (pixel_format != RCW_OBS_RGB8 && pixel_format != RCW_OBS_XRGB32 && pixel_format != RCW_OBS_GRAY8 &&
        pixel_format != RCW_OBS_GRAY16F)
        return cudaErrorInvalidValue;
    return launch_frame_alt<kModeRender, kStagePaint>(p, pixel_format, sh, s);
}

cudaError_t launch_reset(const ResetParams& p, cudaStream_t s) {
    const int64_t blocks = (p.num_envs + 255) / 256;
    reset_kernel<<<(unsigned)blocks, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_commit_range(const StateRef& from, const StateRef& to, int64_t env0, int64_t n, cudaStream_t s) {
    commit_range_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(from, to, env0, n);
    return cudaGetLastError();
}

#include "rcw_topview.cuh"

cudaError_t upload_dir_slot(int slot, const float2* host_dirs, int n, cudaStream_t s) {
    return cudaMemcpyToSymbolAsync(c_dirs, host_dirs, sizeof(float2) * (size_t)n,
                                   sizeof(float2) * (size_t)slot * kDirSlotEntries,
                                   cudaMemcpyHostToDevice, s);
}

}  // namespace rcw
