// rcw_topview.cuh — update_top_view!(env) for a batch (reference: src/single_room.jl:342-372, 446-483).
// Included by rcw_kernels.cu inside namespace rcw (it shares dda_cast, the TMA helpers and c_dirs).
//
// The reference redraws, per env and per step, the tile grid (filled 32x32 squares with a one-pixel
// border), the 512 ray segments from the player to where each ray stopped, and the player's circle, into
// top_view::Array{UInt32}(H * pu, W * pu) (:302).  The drawing primitives belong to SimpleDraw.jl 0.3, which
// is not vendored: they are restated from the algorithms that package documents (Bresenham line over all
// octants, midpoint circle), every pixel bounds-checked — UNPINNED like the DDA.
//
// One CTA draws one env.  The image is never read back from HBM: the ray segments are rasterised into a bit
// plane in shared memory (one bit per pixel, atomicOr), the circle into a small bitmap of its bounding box,
// then the CTA streams the whole image out once, in whole 32-byte sectors, composing every pixel as
// circle > ray > tile border > tile colour — the draw order of the reference (:472, :474-478, :480).
// 512 KB per env at the defaults: HBM-write bound.

#ifndef RCW_TOP_WALK
#define RCW_TOP_WALK 0      // 0: whole segments in lock step (shipped); 1: chunks of 32 pixels with a closed-form start
#endif
#ifndef RCW_TOP_STAGGER_NS
#define RCW_TOP_STAGGER_NS 0
#endif
constexpr int kTopThreads = 256;
constexpr int kTopChunkLog = 5;          // a ray segment is drawn in chunks of 32 pixels, one lane each

// utils.jl:6 — wu_to_pu(x_wu, pu_per_wu) = floor(Int, x_wu * pu_per_wu) + 1 (Float32 product)
__device__ __forceinline__ int wu_to_pu(float x_wu, float pu) { return __float2int_rd(__fmul_rn(x_wu, pu)) + 1; }

// Bit planes: pixel (i0, j0) (0-based) is bit j0 * SB + i0, SB = Hp rounded up to a multiple of 32, plus 32:
// every image column starts on a word of its own and consecutive columns start in consecutive banks, so a
// warp whose 32 rays sit in 32 different columns at the same row does not collide on one bank.
__host__ __device__ __forceinline__ uint32_t plane_col_bits(int Hp) { return (((uint32_t)Hp + 31u) & ~31u) + 32u; }

__device__ __forceinline__ void plane_set(uint32_t* plane, int i, int j, int Hp, int Wp, uint32_t SB) {
    if (i >= 1 && i <= Hp && j >= 1 && j <= Wp) {
        const uint32_t idx = (uint32_t)(j - 1) * SB + (uint32_t)(i - 1);
        atomicOr(plane + (idx >> 5), 1u << (idx & 31u));
    }
}

// The player's circle lives in a bitmap of its bounding box: D = 2 rp + 1 columns of D bits (rp = radius in pixels),
// each column in CW = ceil(D / 32) words; local pixel (ci, cj) = (i - ip + rp, j - jp + rp).
__host__ __device__ __forceinline__ uint32_t circle_words(int rp) {
    const uint32_t D = 2u * (uint32_t)rp + 1u;
    return D * ((D + 31u) >> 5);
}

// ROOM: the wall layer is exactly the border of the map (RoomMap): no layer is staged, the DDA counts down to the border.
template <bool ROOM>
__global__ void __launch_bounds__(kTopThreads, 6) top_view_kernel(const __grid_constant__ TopViewParams p) {
    // [wall layer][ray plane][palette 8 x u32][line list R x int2][chunk starts R x u32][tile colour (W + 1) x H u32]
    // [circle bitmap][row info u16][column info u16][row-sector info u16][column offset u16]
    extern __shared__ __align__(128) uint32_t s_top[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ unsigned long long s_counts;                    // distinct segments | chunks << 32
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, W = p.W, Hp = p.Hp, Wp = p.Wp, pu = p.pu, R = p.R;
    const float fpu = (float)pu;
    const int rp = wu_to_pu(p.radius, fpu);                    // :470, the same for every env
    const uint32_t SB = plane_col_bits(Hp);                    // bits per plane column
    const uint32_t plane_words = ((uint32_t)Wp * (SB >> 5) + 3u) & ~3u;   // the plane is cleared as uint4
    uint32_t* const s_map = s_top;
    uint32_t* const s_ray = s_map + (ROOM ? 0 : p.stage_words);
    uint32_t* const s_pal = s_ray + plane_words;
    int2* const s_line = reinterpret_cast<int2*>(s_pal + 8);                       // [R] end pixels of the distinct segments
    uint32_t* const s_cstart = reinterpret_cast<uint32_t*>(s_line + R);            // [R] first chunk of every distinct segment
    uint32_t* const s_tilec = s_cstart + R;                                        // [W + 1][H] tile colour; column W = border colour
    uint32_t* const s_circ = s_tilec + (size_t)(W + 1) * H;                        // [2 rp + 1][CW] the circle's bounding box
    uint16_t* const s_row = reinterpret_cast<uint16_t*>(s_circ + circle_words(rp));   // [Hp] tile row | border << 15
    uint16_t* const s_colinfo = s_row + ((Hp + 1) & ~1);                           // [Wp] tile column | border << 15
    uint16_t* const s_rowsec = s_colinfo + ((Wp + 1) & ~1);                        // [Hp / 8] tile row | first px border << 14 | last << 15
    uint16_t* const s_coloff = s_rowsec + (((Hp >> 3) + 2) & ~1);                  // [Wp] H * tile column, or H * W for a border column

    const uint32_t env_rel = blockIdx.x;
    const int64_t env = p.env_first + env_rel;
    if (p.mask && !__ldg(p.mask + env)) return;               // masked redraw: the whole CTA leaves
#if RCW_TOP_STAGGER_NS > 0
    // experiment: the CTAs that start together on an SM (first wave) begin a sixth of a CTA lifetime apart
    if (blockIdx.x < 6u * p.sm_count) __nanosleep(((blockIdx.x / p.sm_count) % 6u) * RCW_TOP_STAGGER_NS);
#endif

    // ---- this env's wall layer: one TMA bulk copy; planes cleared and tables built meanwhile
    if (tid == 0) {
        s_counts = 0ULL;
        if (!ROOM) {
            mbar_init(&s_bar, 1);
            mbar_arrive_expect_tx(&s_bar, (uint32_t)p.stage_words * 4u);
            bulk_copy_g2s(s_map, p.wall_map + (size_t)env * p.map_env_stride, (uint32_t)p.stage_words * 4u, &s_bar);
        }
    }
    {
        uint4* const z = reinterpret_cast<uint4*>(s_ray);       // 16-byte aligned: map_words is a multiple of 4
        for (uint32_t k = tid; k < plane_words >> 2; k += kTopThreads) z[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    const float x = __ldg(p.st.pos_x + env), y = __ldg(p.st.pos_y + env);
    const int au = __ldg(p.st.dir_au + env);
    const uint32_t goal = __ldg(p.st.goal + env);
    const int gi0 = (int)(goal & 0xFFFFu) - 1, gj0 = (int)(goal >> 16) - 1;
    if (tid < 6) s_pal[tid] = p.palette[tid];
    for (int i = tid; i < Hp; i += kTopThreads) {
        const int t = i / pu, r = i - t * pu;
        s_row[i] = (uint16_t)(t | ((r == 0 || r == pu - 1) ? 0x8000 : 0));
    }
    for (int j = tid; j < Wp; j += kTopThreads) {
        const int t = j / pu, r = j - t * pu;
        const bool border = r == 0 || r == pu - 1;
        s_colinfo[j] = (uint16_t)(t | (border ? 0x8000 : 0));
        s_coloff[j] = (uint16_t)(H * (border ? W : t));
    }
    for (int q = tid; q < (Hp >> 3); q += kTopThreads) {        // used when pu % 8 == 0: a sector lies in one tile
        const int i = q << 3, t = i / pu, r = i - t * pu;
        s_rowsec[q] = (uint16_t)(t | (r == 0 ? 0x4000 : 0) | (r + 7 == pu - 1 ? 0x8000 : 0));
    }
    // ---- the player (:480): [EXT SimpleDraw] Circle(Point(i - r, j - r), 2r + 1) = midpoint circle of radius r,
    //      drawn by one thread of the last warp while the others build the tables
    const uint32_t CW = (2u * (uint32_t)rp + 32u) >> 5;        // words per column of the circle bitmap
    if (tid == kTopThreads - 1) {
        for (uint32_t k = 0; k < circle_words(rp); ++k) s_circ[k] = 0u;
        auto set = [&](int da, int db) {                        // pixel (ip + da, jp + db)
            const uint32_t ci = (uint32_t)(da + rp), cj = (uint32_t)(db + rp);
            s_circ[cj * CW + (ci >> 5)] |= 1u << (ci & 31u);
        };
        int a = 0, b = rp, d = 1 - rp;
        while (a <= b) {
            set(a, b), set(-a, b), set(a, -b), set(-a, -b);
            set(b, a), set(-b, a), set(b, -a), set(-b, -a);
            if (d < 0) {
                d += 2 * a + 3;
            } else {
                d += 2 * (a - b) + 5;
                b -= 1;
            }
            a += 1;
        }
    }
    __syncthreads();          // mbarrier initialised, plane cleared
    if (!ROOM) mbar_wait(&s_bar, 0);

    // ---- draw_tile_map! colour of every tile: findfirst over the layers WALL, GOAL (:355-360), then the extra objects
    typename std::conditional<ROOM, RoomMap, BitsMap>::type map;
    if constexpr (ROOM) map = RoomMap{H - 1, W - 1};
    else map = BitsMap{s_map, p.n_extra ? s_map + (p.n_extra + 1) * p.map_words : s_map, p.wpr, p.n_extra, p.map_words};
    for (int t = tid; t < H * (W + 1); t += kTopThreads) {
        const int j0 = t / H, i0 = t - j0 * H;
        int code = RCW_TOP_COLOR_BORDER;                      // pseudo tile column W: what a border column of the image shows
        if (j0 < W)
            code = map.wall(i0, j0) ? RCW_TOP_COLOR_WALL
                                    : ((i0 == gi0 && j0 == gj0) ? RCW_TOP_COLOR_GOAL : RCW_TOP_COLOR_EMPTY);
        uint32_t colour = p.palette[code];
        if (code == RCW_TOP_COLOR_EMPTY && j0 < W)        // tile_map_colors[findfirst(...)] over the extra objects
            for (int k = map.n_extra - 1; k >= 0; --k) colour = map.extra(k, i0, j0) ? p.extra_color[k] : colour;
        s_tilec[t] = colour;
    }

    // ---- the ray segments (:474-478), first half: lane <-> ray, where each ray stops, as a pixel
    const int ip = wu_to_pu(x, fpu), jp = wu_to_pu(y, fpu);           // :469
    const int groups = (R + 31) >> 5;
    for (int g = warp; g < groups; g += kTopThreads / 32) {
        const int ray = g * 32 + lane;
        const float4 rt = __ldg(p.ray_table + (size_t)au * (size_t)R + (size_t)min(ray, R - 1));
        const RayHit hit = dda_cast(map, H, W, p.dda_flags, p.closed_border != 0, x, y, gi0, gj0, rt, lane);
        // player_position_wu + ray_distance_wu[i] * ray_direction_wu (:476), one rounding per operation
        const int i2 = wu_to_pu(__fadd_rn(x, __fmul_rn(hit.dist, rt.x)), fpu);
        const int j2 = wu_to_pu(__fadd_rn(y, __fmul_rn(hit.dist, rt.y)), fpu);
#if RCW_TOP_WALK >= 1   // chunked walk (1) / lock-step first ring + chunked remainder (2): measured alternatives, profiles/README.md
        // All segments start at the player's pixel, and neighbouring rays stop a fraction of a pixel apart (0.2 - 0.6
        // px at the usual distances): a ray whose stop pixel equals its lower neighbour's draws exactly the same
        // pixels and is dropped here.  The distinct segments (about 4 in 10 at the defaults) are appended to one
        // list per env — their order only decides which lane sets a shared bit, not the picture — together with the
        // index of their first chunk: a segment of n steps (n + 1 pixels) is drawn as n / 32 + 1 chunks of up to 32
        // pixels, and the chunks of all segments are numbered consecutively in list order.
        const int i2_below = __shfl_up_sync(0xFFFFFFFFu, i2, 1), j2_below = __shfl_up_sync(0xFFFFFFFFu, j2, 1);
        const bool distinct = (ray < R) & ((lane == 0) | (i2 != i2_below) | (j2 != j2_below));
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, distinct);
#if RCW_TOP_WALK == 2
        // hybrid: the first 32 steps of every segment are walked in lock step (below); chunks cover steps 32 ..
        const int n_steps = max(abs(i2 - ip), abs(j2 - jp));
        const uint32_t nch = (distinct && n_steps >= 32) ? ((uint32_t)(n_steps - 32) >> kTopChunkLog) + 1u : 0u;
#else
        const uint32_t nch = distinct ? ((uint32_t)max(abs(i2 - ip), abs(j2 - jp)) >> kTopChunkLog) + 1u : 0u;
#endif
        uint32_t incl = nch;                                    // inclusive prefix sum of the chunk counts
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += up;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        unsigned long long base = 0ULL;
        if (lane == 0) base = atomicAdd(&s_counts, (unsigned long long)__popc(ballot) | ((unsigned long long)total << 32));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (distinct) {
            const uint32_t slot = (uint32_t)base + (uint32_t)__popc(ballot & ((1u << lane) - 1u));
            s_line[slot] = make_int2(i2, j2);
            s_cstart[slot] = (uint32_t)(base >> 32) + incl - nch;
        }
    }
    __syncthreads();          // the list is complete (and every tile colour written)

    // ---- second half: lane <-> chunk.  [EXT SimpleDraw] Line(point1, point2): Bresenham, all octants, both end
    //      points drawn; the walk reaches the end after exactly n = max(|di|, |dj|) steps, the major axis advancing in
    //      every one of them (the all-octant form `e2 = 2 err; if e2 >= dj ...; if e2 <= di ...` reduces to that), so
    //      one decision variable suffices: with M / m the larger / smaller of |di|, |dj| and F = 2 err' - M (err' = err
    //      for |di| >= |dj|, -err otherwise; F starts at M - 2 m), the minor axis steps iff F <= 0, and
    //      F += step ? 2 (M - m) : -2 m — the same pixels, ties included, in both octant families.  That recurrence
    //      has a closed form: after k steps the minor axis has advanced s_k = floor((2 m k + M) / (2 M)) times and
    //      F_k = M - 2 m (k + 1) + 2 M s_k (induction on k: s_{k+1} = s_k + [F_k <= 0]).  So a segment need not be
    //      walked by one lane from its start: chunk c of a segment begins at step 32 c with one integer division, and
    //      every lane walks at most 32 pixels — no warp waits for its longest line (the lock-step walk of whole
    //      segments spent 1480 warp-iterations per env on 780 iterations' worth of pixels), and the CTA reaches the
    //      barrier in front of the stream-out together.  Chunks are dealt to the lanes with a stride of 8, so the lanes
    //      of a warp draw in different segments or far apart in the same one and rarely meet in a plane word.
    const int n_lines = (int)(uint32_t)s_counts;
    const uint32_t n_chunks = (uint32_t)(s_counts >> 32);
#if RCW_TOP_WALK == 2
    // ---- the first ring: steps 0 .. 31 of every segment, lane <-> segment in lock step.  Next to the player the 225
    //      segments of a default env draw 7.2 K pixels onto 600: neighbouring lanes (neighbouring rays) are on the same
    //      pixel most of the time, and a lane whose pixel equals its lower neighbour's leaves the atomic to it.
    for (int l0 = warp * 32; l0 < n_lines; l0 += kTopThreads) {
        const bool have1 = l0 + lane < n_lines;
        const int2 end = s_line[min(l0 + lane, n_lines - 1)];
        const int i2 = end.x, j2 = end.y;
        const int di = abs(i2 - ip), dj = abs(j2 - jp);
        const int si = ip < i2 ? 1 : -1, sj = jp < j2 ? 1 : -1;
        const int M = max(di, dj), m = min(di, dj);
        const int n = have1 ? min(M, 31) : -1;
        const int n_max = __reduce_max_sync(0xFFFFFFFFu, n);
        const int n_below = __shfl_up_sync(0xFFFFFFFFu, n, 1);
        const bool clip = (ip < 1) | (ip > Hp) | (jp < 1) | (jp > Wp) | (i2 < 1) | (i2 > Hp) | (j2 < 1) | (j2 > Wp);
        const bool i_major = di >= dj;
        const int f_stay = -2 * m, f_step = 2 * (M - m);
        int F = M - 2 * m;
        if (!__any_sync(0xFFFFFFFFu, clip && have1)) {
            const int bit_si = si, bit_sj = sj * (int)SB;
            const int adv_major = i_major ? bit_si : bit_sj, adv_both = bit_si + bit_sj;
            const int n_dup = lane > 0 ? n_below : -1;
            uint32_t idx = (uint32_t)(jp - 1) * SB + (uint32_t)(ip - 1);
#pragma unroll 2
            for (int k = 0; k <= n_max; ++k) {
                const uint32_t below = __shfl_up_sync(0xFFFFFFFFu, idx, 1);
                const bool dup = (k <= n_dup) & (below == idx);
                if ((k <= n) & !dup) atomicOr(s_ray + (idx >> 5), 1u << (idx & 31u));
                const bool step = F <= 0;
                F += step ? f_step : f_stay;
                idx += (uint32_t)(step ? adv_both : adv_major);
            }
        } else if (have1) {
            int i = ip, j = jp;
            for (int k = 0; k <= n; ++k) {
                plane_set(s_ray, i, j, Hp, Wp, SB);
                const bool step = F <= 0;
                F += step ? f_step : f_stay;
                i += (i_major | step) ? si : 0;
                j += (!i_major | step) ? sj : 0;
            }
        }
    }
    constexpr int kFirstChunkStep = 32;
    // chunks: thread t takes chunks t T .. t T + T - 1 (T = rounds): the lanes of a warp are T chunks apart, mostly in
    // different segments, and only the CTA's last few threads ever idle
    const uint32_t n_rounds = (n_chunks + kTopThreads - 1) / kTopThreads;
#else
    constexpr int kFirstChunkStep = 0;
#endif
    int search0 = 1;
    while (search0 < n_lines) search0 <<= 1;                    // first probe distance of the segment search
#if RCW_TOP_WALK == 2
    for (uint32_t t = 0; t < n_rounds; ++t) {
        const uint32_t g = (uint32_t)tid * n_rounds + t;
        if (!__any_sync(0xFFFFFFFFu, g < n_chunks)) break;
        const bool have = g < n_chunks;
#else
    for (uint32_t g0 = 0; g0 < n_chunks; g0 += kTopThreads) {
        const uint32_t g = g0 + (uint32_t)(lane * (kTopThreads / 32) + warp);
        const bool have = g < n_chunks;
#endif
        // the segment this chunk belongs to: the last one whose first chunk is <= g
        int seg = 0;
        for (int d = search0 >> 1; d > 0; d >>= 1)
            if (seg + d < n_lines && s_cstart[seg + d] <= g) seg += d;
        const int2 end = s_line[seg];
        const int i2 = end.x, j2 = end.y;
        const int di = abs(i2 - ip), dj = abs(j2 - jp);
        const int si = ip < i2 ? 1 : -1, sj = jp < j2 ? 1 : -1;
        const bool i_major = di >= dj;
        const int M = max(di, dj), m = min(di, dj);
        const int k0 = have ? kFirstChunkStep + (int)((g - s_cstart[seg]) << kTopChunkLog) : 0;
        const int n_px = have ? min(1 << kTopChunkLog, M - k0 + 1) : 0;       // pixels of this chunk
        const int s0 = M > 0 ? (int)((2u * (uint32_t)m * (uint32_t)k0 + (uint32_t)M) / (2u * (uint32_t)M)) : 0;
        int F = M - 2 * m * (k0 + 1) + 2 * M * s0;
        // both end points inside the image => every pixel of the line is (it stays in their bounding box)
        const bool clip = (ip < 1) | (ip > Hp) | (jp < 1) | (jp > Wp) | (i2 < 1) | (i2 > Hp) | (j2 < 1) | (j2 > Wp);
        const int f_stay = -2 * m, f_step = 2 * (M - m);
        if (!__any_sync(0xFFFFFFFFu, clip && have)) {
            const int bit_si = si, bit_sj = sj * (int)SB;       // bit index steps of the two axes
            const int adv_major = i_major ? bit_si : bit_sj, adv_both = bit_si + bit_sj;
            uint32_t idx = (uint32_t)(jp - 1) * SB + (uint32_t)(ip - 1) + (uint32_t)(k0 * adv_major + s0 * (adv_both - adv_major));
#pragma unroll 4
            for (int k = 0; k < (1 << kTopChunkLog); ++k) {
                if (k < n_px) atomicOr(s_ray + (idx >> 5), 1u << (idx & 31u));
                const bool step = F <= 0;
                F += step ? f_step : f_stay;
                idx += (uint32_t)(step ? adv_both : adv_major);
            }
        } else {
            int i = ip + si * (i_major ? k0 : s0), j = jp + sj * (i_major ? s0 : k0);
            for (int k = 0; k < n_px; ++k) {
                plane_set(s_ray, i, j, Hp, Wp, SB);
                const bool step = F <= 0;
                F += step ? f_step : f_stay;
                i += (i_major | step) ? si : 0;
                j += (!i_major | step) ? sj : 0;
            }
        }
    }

#else
        // All segments start at the player's pixel, and neighbouring rays stop a fraction of a pixel apart (0.2 - 0.6
        // px at the usual distances): a ray whose stop pixel equals its lower neighbour's draws exactly the same
        // pixels and is dropped here.  The distinct segments (about 4 in 10 at the defaults) are appended to one
        // list per env — their order only decides which lane sets a shared bit, not the picture.
        const int i2_below = __shfl_up_sync(0xFFFFFFFFu, i2, 1), j2_below = __shfl_up_sync(0xFFFFFFFFu, j2, 1);
        const bool distinct = (ray < R) & ((lane == 0) | (i2 != i2_below) | (j2 != j2_below));
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, distinct);
        uint32_t base = 0u;
        if (lane == 0) base = (uint32_t)atomicAdd(&s_counts, (unsigned long long)__popc(ballot));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (distinct) s_line[base + (uint32_t)__popc(ballot & ((1u << lane) - 1u))] = make_int2(i2, j2);
    }
    __syncthreads();          // the list is complete (and every tile colour written)

    // ---- second half: lane <-> distinct segment.  [EXT SimpleDraw] Line(point1, point2): Bresenham, all octants,
    //      both end points drawn; the walk reaches the end after exactly max(|di|, |dj|) steps, the major axis
    //      advancing in every one of them (the all-octant form `e2 = 2 err; if e2 >= dj ...; if e2 <= di ...`
    //      reduces to that), so one decision variable suffices: with M / m the larger / smaller of |di|, |dj| and
    //      F = 2 err' - M (err' = err for |di| >= |dj|, -err otherwise; F starts at M - 2 m), the minor axis steps
    //      iff F <= 0, and F += step ? 2 (M - m) : -2 m — the same pixels, ties included, in both octant families.
    //      The warp walks its 32 segments in lock step up to the longest one; a lane whose pixel equals its lower
    //      neighbour's (neighbouring segments share their first pixels) leaves the shared-memory atomic to it.
    const int n_lines = (int)(uint32_t)s_counts;
    for (int l0 = warp * 32; l0 < n_lines; l0 += kTopThreads) {
        const bool have = l0 + lane < n_lines;
        const int2 end = s_line[min(l0 + lane, n_lines - 1)];
        const int i2 = end.x, j2 = end.y;
        const int di = abs(i2 - ip), dj = abs(j2 - jp);
        const int si = ip < i2 ? 1 : -1, sj = jp < j2 ? 1 : -1;
        const int n = have ? max(di, dj) : -1;                    // steps of this lane's segment; -1: none
        const int n_max = __reduce_max_sync(0xFFFFFFFFu, n);
        const int n_below = __shfl_up_sync(0xFFFFFFFFu, n, 1);
        // both end points inside the image => every pixel of the line is (it stays in their bounding box)
        const bool clip = (ip < 1) | (ip > Hp) | (jp < 1) | (jp > Wp) | (i2 < 1) | (i2 > Hp) | (j2 < 1) | (j2 > Wp);
        if (!__any_sync(0xFFFFFFFFu, clip && have)) {
            const bool i_major = di >= dj;
            const int M = max(di, dj), m = min(di, dj);
            const int bit_si = si, bit_sj = sj * (int)SB;       // bit index steps of the two axes
            const int adv_major = i_major ? bit_si : bit_sj, adv_both = bit_si + bit_sj;
            const int f_stay = -2 * m, f_step = 2 * (M - m);
            const int n_dup = lane > 0 ? n_below : -1;          // the lower neighbour still draws while k <= n_dup
            int F = M - 2 * m;
            uint32_t idx = (uint32_t)(jp - 1) * SB + (uint32_t)(ip - 1);
#pragma unroll 2
            for (int k = 0; k <= n_max; ++k) {
                const uint32_t below = __shfl_up_sync(0xFFFFFFFFu, idx, 1);
                const bool dup = (k <= n_dup) & (below == idx);
                if ((k <= n) & !dup) atomicOr(s_ray + (idx >> 5), 1u << (idx & 31u));
                const bool step = F <= 0;
                F += step ? f_step : f_stay;
                idx += (uint32_t)(step ? adv_both : adv_major);
            }
        } else if (have) {
            int i = ip, j = jp, err = di - dj;
            for (int k = n;; --k) {
                plane_set(s_ray, i, j, Hp, Wp, SB);
                if (k == 0) break;
                const int e2 = 2 * err;
                if (e2 >= -dj) {
                    err -= dj;
                    i += si;
                }
                if (e2 <= di) {
                    err += di;
                    j += sj;
                }
            }
        }
    }

#endif
    __syncthreads();

    // ---- stream the image out: one 32-byte sector (8 consecutive pixels of the column-major image) per
    //      lane and iteration, every pixel composed as circle > ray > tile border > tile colour
    uint32_t slot = p.slot0 + env_rel;
    if (slot >= p.window) slot -= p.window;
    uint8_t* const img = p.top + (size_t)slot * p.env_stride;
    const uint8_t* const ray_bytes = reinterpret_cast<const uint8_t*>(s_ray);
    const uint32_t border_c = s_pal[RCW_TOP_COLOR_BORDER], ray_c = s_pal[RCW_TOP_COLOR_RAY],
                   player_c = s_pal[RCW_TOP_COLOR_PLAYER];
    const uint32_t L = (uint32_t)Hp * (uint32_t)Wp;          // pixels
    const uint32_t n_sec = (L + 7u) >> 3;
    const int ci0 = ip - 1 - rp, cj0 = jp - 1 - rp;          // 0-based image position of the circle bitmap's corner
    const uint32_t D = 2u * (uint32_t)rp + 1u;

    // sector s of any image size, pixel by pixel through the row / column tables, circle included
    auto store_sector_by_pixel = [&](uint32_t s) {
        const uint32_t idx0 = s << 3;
        uint32_t j0 = idx0 / (uint32_t)Hp, i0 = idx0 - j0 * (uint32_t)Hp;
        uint32_t px[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t jc = min(j0, (uint32_t)Wp - 1u);
            const uint32_t rinfo = s_row[i0], cinfo = s_colinfo[jc];
            uint32_t c = ((rinfo | cinfo) & 0x8000u) ? border_c : s_tilec[(rinfo & 0x7FFFu) + (uint32_t)H * (cinfo & 0x7FFFu)];
            const uint32_t bit = jc * SB + i0;
            c = ((s_ray[bit >> 5] >> (bit & 31u)) & 1u) ? ray_c : c;
            const uint32_t ci = i0 - (uint32_t)ci0, cj = jc - (uint32_t)cj0;
            if (ci < D && cj < D && ((s_circ[cj * CW + (ci >> 5)] >> (ci & 31u)) & 1u)) c = player_c;
            px[k] = (idx0 + (uint32_t)k < L) ? c : 0u;    // bytes behind the last pixel are padding
            if (++i0 == (uint32_t)Hp) {                    // next column of the image
                i0 = 0;
                ++j0;
            }
        }
        store_stream32(img + ((size_t)s << 5), make_uint4(px[0], px[1], px[2], px[3]),
                       make_uint4(px[4], px[5], px[6], px[7]));
    };

    if (((Hp | pu) & 7) != 0) {
        for (uint32_t s = tid; s < n_sec; s += kTopThreads) store_sector_by_pixel(s);
        return;
    }

    // pu and Hp multiples of 8: a sector lies inside one tile of one image column — one colour, borders only at
    // its two ends.  The sweep leaves the circle out; the few sectors it touches are rewritten afterwards.
    {
        const uint32_t SPC = (uint32_t)Hp >> 3, SBy = SB >> 3;   // sectors / plane bytes per column
        uint32_t j0 = (uint32_t)tid / SPC, q = (uint32_t)tid - j0 * SPC;
        const uint32_t adv_j = kTopThreads / SPC, adv_q = kTopThreads - adv_j * SPC;
        if (adv_q == 0) {
            // the CTA covers whole columns per sweep (256 % (Hp / 8) == 0, e.g. the default 256 rows): a thread
            // stays on its rows, so the tile row and the border ends of its sectors are fixed
            const uint32_t rs = s_rowsec[q];
            const uint32_t* const tile_row = s_tilec + (rs & 0x3FFFu);
            const bool border_first = (rs & 0x4000u) != 0u, border_last = (rs & 0x8000u) != 0u;
            const uint8_t* rayp = ray_bytes + q + j0 * SBy;
            uint8_t* dst = img + ((size_t)tid << 5);
#pragma unroll 2
            for (; j0 < (uint32_t)Wp; j0 += adv_j) {
                const uint32_t base = tile_row[s_coloff[j0]];
                const uint32_t rb = *rayp;
                uint32_t px[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) px[k] = base;
                if (border_first) px[0] = border_c;
                if (border_last) px[7] = border_c;
                if (rb) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) px[k] = ((rb >> k) & 1u) ? ray_c : px[k];
                }
                store_stream32(dst, make_uint4(px[0], px[1], px[2], px[3]), make_uint4(px[4], px[5], px[6], px[7]));
                rayp += adv_j * SBy;
                dst += (size_t)kTopThreads << 5;
            }
        } else {
#pragma unroll 2
            for (uint32_t s = tid; s < n_sec; s += kTopThreads) {
                const uint32_t rs = s_rowsec[q];
                const uint32_t base = s_tilec[(rs & 0x3FFFu) + s_coloff[j0]];
                const uint32_t rb = ray_bytes[j0 * SBy + q];
                uint32_t px[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) px[k] = base;
                if (rs & 0x4000u) px[0] = border_c;
                if (rs & 0x8000u) px[7] = border_c;
                if (rb) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) px[k] = ((rb >> k) & 1u) ? ray_c : px[k];
                }
                store_stream32(img + ((size_t)s << 5), make_uint4(px[0], px[1], px[2], px[3]),
                               make_uint4(px[4], px[5], px[6], px[7]));
                j0 += adv_j;
                q += adv_q;
                if (q >= SPC) {
                    q -= SPC;
                    ++j0;
                }
            }
        }
    }
    // the circle: the sectors under its bounding box once more, now complete.  The barrier orders the two stores
    // of such a sector (both by this CTA), so the later one is what memory keeps.
    __syncthreads();
    {
        const int j_lo = max(cj0, 0), j_hi = min(cj0 + 2 * rp, Wp - 1);
        const int i_lo = max(ci0, 0), i_hi = min(ci0 + 2 * rp, Hp - 1);
        if (j_lo <= j_hi && i_lo <= i_hi) {
            const uint32_t q_lo = (uint32_t)i_lo >> 3, nq = ((uint32_t)i_hi >> 3) - q_lo + 1u;
            const uint32_t n_fix = (uint32_t)(j_hi - j_lo + 1) * nq;
            for (uint32_t t = tid; t < n_fix; t += kTopThreads) {
                const uint32_t c = t / nq, k = t - c * nq;
                store_sector_by_pixel(((uint32_t)j_lo + c) * ((uint32_t)Hp >> 3) + q_lo + k);
            }
        }
    }
}

// map_words: words staged for the map (every object layer)
size_t top_view_smem_bytes(int H, int W, int R, int pu, float radius, int map_words) {
    const size_t Hp = (size_t)H * pu, Wp = (size_t)W * pu;
    const size_t plane_words = (Wp * (plane_col_bits((int)Hp) >> 5) + 3) & ~(size_t)3;
    const int rp = (int)floorf(radius * (float)pu) + 1;          // wu_to_pu(radius, pu)
    return (size_t)map_words * 4 + plane_words * 4 + 8 * 4 + (size_t)R * 12 + (size_t)(W + 1) * H * 4 +
           (size_t)circle_words(rp) * 4 + 2 * ((Hp + 1) & ~(size_t)1) + 2 * ((Wp + 1) & ~(size_t)1) +
           2 * (((Hp >> 3) + 2) & ~(size_t)1) + 2 * ((Wp + 1) & ~(size_t)1);
}

cudaError_t launch_top_view(const TopViewParams& p, cudaStream_t s) {
    const size_t smem = top_view_smem_bytes(p.H, p.W, p.R, p.pu, p.radius, p.room ? 0 : p.stage_words);
    if (smem > 48 * 1024) {
        cudaError_t e = p.room ? cudaFuncSetAttribute(top_view_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                               : cudaFuncSetAttribute(top_view_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (p.room) top_view_kernel<true><<<(unsigned)p.env_count, kTopThreads, smem, s>>>(p);
    else top_view_kernel<false><<<(unsigned)p.env_count, kTopThreads, smem, s>>>(p);
    return cudaGetLastError();
}
