# BatchedRayCastWorlds.jl — the reference-side binding of librcw_b200.so.
#
# UNTESTED IN THIS IMAGE (no julia binary here or on the GPU box).  It is the binding a
# RayCastWorlds.jl maintainer adds next to src/single_room.jl: the host stays in Julia and keeps the
# package's API (AbstractGame, reset!, act!, get_action_names, RLBaseEnv); the hot path
#   act!(world) -> cast_rays!(world) -> update_camera_view!(env)      (src/single_room.jl:333-340)
# runs in hand-written sm_100a kernels behind the C ABI of include/rcw_b200.h.
#
# Usage
#   include("BatchedRayCastWorlds.jl"); import .BatchedRayCastWorlds as BRCW
#   env = BRCW.BatchedSingleRoom(num_envs = 4096)            # same keywords as SingleRoom(...)
#   RCW.reset!(env); RCW.act!(env, rand(UInt8(1):UInt8(4), 4096))
#   obs_ptr, nbytes, stride = BRCW.obs_device_ptr(env)       # wrap with CUDA.unsafe_wrap if CUDA.jl is loaded
#   r, d = BRCW.reward_done(env)

module BatchedRayCastWorlds

import RayCastWorlds as RCW
import ReinforcementLearningBase as RLBase

const LIB = get(ENV, "RCW_B200_LIB", joinpath(@__DIR__, "..", "lib", "librcw_b200.so"))

const RCW_OK = Int32(0)
const RCW_EACTION = Int32(-2)
const RCW_OBS_RGB8 = Int32(0)
const RCW_OBS_XRGB32 = Int32(1)
const RCW_OBS_GRAY8 = Int32(2)      # BT.601 luma of the reference pixel, one byte per pixel
const RCW_OBS_COLUMNS = Int32(3)    # one UInt32 per ray column: pad | colour id << 16 (see expand_columns)
const RCW_OBS_GRAY8_HALF = Int32(5) # UInt8 [num_envs][num_rays / 2][height_px / 2]: the GRAY8 frame under a 2 x 2 box filter
const RCW_OBS_GRAY16F = Int32(4)   # Float16 [num_envs][num_rays][height_px]: GRAY8 luma / 255 (normalised learner frames)
const RCW_ABI_VERSION = Int32(4)
const RCW_MAX_EXTRA_LAYERS = 4
const RCW_LAYER_BLOCKING = Int32(0)   # an extra object layer that refuses the move like WALL
const RCW_LAYER_TERMINAL = Int32(1)   # ... that ends the episode with its own reward like GOAL

version() = ccall((:rcw_version, LIB), Int32, ())

# mirrors `struct rcw_config` (include/rcw_b200.h); isbits, same field order and C layout
struct RcwConfig
    struct_size::UInt32
    device::Int32
    num_envs::Int64
    env_id_offset::Int64
    height_tile_map_tu::Int32
    width_tile_map_tu::Int32
    num_directions::Int32
    num_rays::Int32
    height_camera_view_pu::Int32
    player_radius_wu::Float32
    position_increment_wu::Float32
    semi_field_of_view_wu::Float32
    camera_height_tile_wu::Float32
    goal_reward::Float32
    obs_format::Int32
    auto_reset::Int32
    seed::UInt64
    palette::NTuple{6, UInt32}
    dda_flags::UInt32
    obs_window_envs::Int32
    top_view::Int32
    pu_per_tu::Int32
    top_palette::NTuple{6, UInt32}
    frame_stack::Int32
    result_ring::Int32
    num_object_layers::Int32
    layer_kind::NTuple{4, Int32}
    layer_reward::NTuple{4, Float32}
    layer_palette::NTuple{8, UInt32}      # [k][hit across dimension 1, 2], row-major like the C array
    layer_top_color::NTuple{4, UInt32}
    reserved::NTuple{3, UInt32}
end

last_error() = unsafe_string(ccall((:rcw_last_error, LIB), Cstring, ()))

function check(rc::Int32)
    rc == RCW_OK && return nothing
    msg = last_error()
    # RCW_EACTION is the reference's `@assert action in Base.OneTo(NUM_ACTIONS)` (single_room.jl:140)
    rc == RCW_EACTION && throw(AssertionError(msg))
    error("librcw_b200 error $(rc): $(msg)")
end

mutable struct BatchedSingleRoom <: RCW.AbstractGame
    handle::Ptr{Cvoid}
    num_envs::Int
    height_tile_map_tu::Int
    width_tile_map_tu::Int
    num_rays::Int
    height_camera_view_pu::Int
    obs_format::Int32
    goal_reward::Float32
    reward::Vector{Float32}     # host mirrors filled by reward_done!
    done::Vector{UInt8}
end

# rcw_config + the reference's own direction table from the keyword arguments of SingleRoom(...) (single_room.jl:258-272)
function build_config(;
        num_envs = 1,
        device = 0,
        T = Float32,
        height_tile_map_tu = 8,
        width_tile_map_tu = 16,
        num_directions = 128,
        player_radius_wu = convert(T, 1 / 8),
        position_increment_wu = convert(T, 1 / 8),
        semi_field_of_view_wu = convert(T, 2 / 3),
        num_rays = 512,
        camera_height_tile_wu = convert(T, 1),
        height_camera_view_pu = 256,
        goal_reward = one(Float32),
        obs_format = RCW_OBS_RGB8,
        auto_reset = true,
        seed = 0,
        env_id_offset = 0,
        obs_window_envs = 0,
        top_view = false,
        pu_per_tu = 32,
        frame_stack = 1,
        result_ring = 0,
        # object layers beyond WALL and GOAL (NUM_OBJECTS > 2, single_room.jl:16-18): one entry per extra layer
        num_object_layers = 2,
        layer_kind = Int32[],            # RCW_LAYER_BLOCKING / RCW_LAYER_TERMINAL
        layer_reward = Float32[],
        layer_palette = NTuple{2, UInt32}[],   # (colour when hit across dimension 1, dimension 2)
        layer_top_color = UInt32[],
    )
    T === Float32 || error("librcw_b200 computes in Float32 (the reference default, single_room.jl:43)")
    version() == RCW_ABI_VERSION || error("librcw_b200 ABI $(version()), this binding was written for $(RCW_ABI_VERSION)")

    # the reference's own direction table (single_room.jl:65-69), so that cos/sin come from Julia
    directions = Matrix{Float32}(undef, 2, num_directions)
    for i in 1:num_directions
        theta = (i - 1) * 2 * pi / num_directions
        directions[1, i] = convert(Float32, cos(theta))
        directions[2, i] = convert(Float32, sin(theta))
    end

    palette = (0x00FFFFFF, 0x00404040, 0x00808080, 0x00c0c0c0, 0x00800000, 0x00c00000)  # single_room.jl:291-296
    # tile_map_colors, tile border, ray_color, player_color  (single_room.jl:288-290, 364-367)
    top_palette = (0x00FFFFFF, 0x00FF0000, 0x00000000, 0x00cccccc, 0x00808080, 0x00c0c0c0)
    cfg = Ref(RcwConfig(UInt32(sizeof(RcwConfig)), Int32(device), Int64(num_envs), Int64(env_id_offset),
                        Int32(height_tile_map_tu), Int32(width_tile_map_tu), Int32(num_directions),
                        Int32(num_rays), Int32(height_camera_view_pu), Float32(player_radius_wu),
                        Float32(position_increment_wu), Float32(semi_field_of_view_wu),
                        Float32(camera_height_tile_wu), Float32(goal_reward), Int32(obs_format),
                        Int32(auto_reset), UInt64(seed), palette, UInt32(0), Int32(obs_window_envs),
                        Int32(top_view), Int32(pu_per_tu), top_palette, Int32(frame_stack),
                        Int32(result_ring), Int32(num_object_layers),
                        ntuple(k -> k <= length(layer_kind) ? Int32(layer_kind[k]) : Int32(0), 4),
                        ntuple(k -> k <= length(layer_reward) ? Float32(layer_reward[k]) : 0f0, 4),
                        ntuple(k -> (k + 1) ÷ 2 <= length(layer_palette) ? UInt32(layer_palette[(k + 1) ÷ 2][2 - k % 2]) : UInt32(0), 8),
                        ntuple(k -> k <= length(layer_top_color) ? UInt32(layer_top_color[k]) : UInt32(0), 4),
                        ntuple(_ -> UInt32(0), 3)))
    return cfg, directions
end

# wrap a handle made by rcw_create / rcw_create_sharded (the wrapper owns it and destroys it when collected)
function wrap_handle(handle::Ptr{Cvoid}, cfg::RcwConfig, num_envs::Integer)
    env = BatchedSingleRoom(handle, Int(num_envs), Int(cfg.height_tile_map_tu), Int(cfg.width_tile_map_tu), Int(cfg.num_rays),
                            Int(cfg.height_camera_view_pu), cfg.obs_format,
                            cfg.goal_reward, zeros(Float32, num_envs), zeros(UInt8, num_envs))
    finalizer(e -> (e.handle != C_NULL && ccall((:rcw_destroy, LIB), Int32, (Ptr{Cvoid},), e.handle); e.handle = C_NULL), env)
    return env
end

function BatchedSingleRoom(; kw...)
    cfg, directions = build_config(; kw...)
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve directions begin
        check(ccall((:rcw_create, LIB), Int32, (Ref{RcwConfig}, Ptr{Float32}, Ref{Ptr{Cvoid}}),
                    cfg, directions, handle))
    end
    return wrap_handle(handle[], cfg[], cfg[].num_envs)
end

# ---- one process, several GPUs (rcw_b200.h: independent env shards, one handle per device, no collective) -------
# shard is 0-based like the C ABI; returns (offset, count) of the shard's block of global env ids
function shard_envs(total_envs::Integer, n_shards::Integer, shard::Integer)
    off = Ref{Int64}(0); cnt = Ref{Int64}(0)
    check(ccall((:rcw_shard_envs, LIB), Int32, (Int64, Int32, Int32, Ref{Int64}, Ref{Int64}),
                Int64(total_envs), Int32(n_shards), Int32(shard), off, cnt))
    return off[], cnt[]
end

struct ShardedSingleRoom <: RCW.AbstractGame
    shards::Vector{BatchedSingleRoom}
    handles::Vector{Ptr{Cvoid}}
    num_envs::Int
end

# num_envs is the TOTAL; shard k runs on devices[k]; every other keyword as for BatchedSingleRoom
function ShardedSingleRoom(; devices = [0], kw...)
    cfg, directions = build_config(; kw...)
    n = length(devices)
    dev = convert(Vector{Int32}, devices)
    handles = fill(C_NULL, n)
    GC.@preserve directions dev handles begin
        check(ccall((:rcw_create_sharded, LIB), Int32, (Ref{RcwConfig}, Ptr{Float32}, Ptr{Int32}, Int32, Ptr{Ptr{Cvoid}}),
                    cfg, directions, dev, Int32(n), handles))
    end
    shards = [wrap_handle(handles[k + 1], cfg[], shard_envs(cfg[].num_envs, n, k)[2]) for k in 0:(n - 1)]
    return ShardedSingleRoom(shards, handles, Int(cfg[].num_envs))
end

# reset!(env): every shard draws new layouts on its device (keyed by global env id)
function RCW.reset!(env::ShardedSingleRoom)
    foreach(RCW.reset!, env.shards)
    return nothing
end

# act!(env, actions): actions of the whole batch in global env order (host vector)
function RCW.act!(env::ShardedSingleRoom, actions::AbstractVector{<:Integer})
    length(actions) == env.num_envs || throw(DimensionMismatch("one action per env"))
    a = convert(Vector{UInt8}, actions)
    GC.@preserve a begin
        check(ccall((:rcw_step_sharded, LIB), Int32, (Ptr{Ptr{Cvoid}}, Int32, Ptr{UInt8}), env.handles, Int32(length(env.handles)), a))
    end
    return nothing
end

step_random!(env::ShardedSingleRoom, n_steps::Integer = 1) =
    check(ccall((:rcw_step_random_sharded, LIB), Int32, (Ptr{Ptr{Cvoid}}, Int32, Int32), env.handles, Int32(length(env.handles)), Int32(n_steps)))

sync(env::ShardedSingleRoom) =
    check(ccall((:rcw_sync_sharded, LIB), Int32, (Ptr{Ptr{Cvoid}}, Int32), env.handles, Int32(length(env.handles))))

function episode_stats(env::ShardedSingleRoom; reset_counters = false)
    ep = Ref{Int64}(0); sr = Ref{Float64}(0); sl = Ref{Int64}(0)
    check(ccall((:rcw_reduce_episode_stats, LIB), Int32, (Ptr{Ptr{Cvoid}}, Int32, Ref{Int64}, Ref{Float64}, Ref{Int64}, Int32),
                env.handles, Int32(length(env.handles)), ep, sr, sl, Int32(reset_counters)))
    return ep[], sr[], sl[]
end

# destroys every shard now instead of at collection (the wrappers' finalizers then find C_NULL)
function close!(env::ShardedSingleRoom)
    check(ccall((:rcw_destroy_sharded, LIB), Int32, (Ptr{Ptr{Cvoid}}, Int32), env.handles, Int32(length(env.handles))))
    for s in env.shards
        s.handle = C_NULL
    end
    fill!(env.handles, C_NULL)
    return nothing
end

# reset!(env) — src/single_room.jl:326-331 (layouts drawn on the device)
function RCW.reset!(env::BatchedSingleRoom)
    check(ccall((:rcw_reset, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{UInt8}),
                env.handle, C_NULL, C_NULL, C_NULL, C_NULL))
    return nothing
end

# reset!(env; goal, player, direction) with host-supplied layouts: 2 x num_envs Int32 matrices (1-based tiles)
function reset_to!(env::BatchedSingleRoom, goal_ij::Matrix{Int32}, player_ij::Matrix{Int32}, dir_au::Vector{Int32})
    GC.@preserve goal_ij player_ij dir_au begin
        check(ccall((:rcw_reset, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{UInt8}),
                    env.handle, goal_ij, player_ij, dir_au, C_NULL))
    end
    return nothing
end

# act!(env, actions) — src/single_room.jl:333-340 for every env; actions[e] in 1:4
function RCW.act!(env::BatchedSingleRoom, actions::Vector{UInt8})
    length(actions) == env.num_envs || throw(DimensionMismatch("one action per env"))
    GC.@preserve actions begin
        check(ccall((:rcw_step, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}), env.handle, actions))
    end
    return nothing
end
RCW.act!(env::BatchedSingleRoom, actions::AbstractVector{<:Integer}) = RCW.act!(env, convert(Vector{UInt8}, actions))

# act! whose rewards / terminations also land in the pinned result ring (`result_ring = D >= 1`); returns the
# step's ticket.  With D >= 2, enqueue step k + 1 before `wait_results(env, k)`: the device never idles.
function act_async!(env::BatchedSingleRoom, actions::Vector{UInt8})
    length(actions) == env.num_envs || throw(DimensionMismatch("one action per env"))
    ticket = Ref{Int64}(0)
    GC.@preserve actions begin
        check(ccall((:rcw_step_async, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Ref{Int64}), env.handle, actions, ticket))
    end
    return ticket[]
end

# blocks until the step of `ticket` has finished; (reward, done) are views of the pinned ring (not owned by Julia),
# valid until `result_ring` further act_async! calls
function wait_results(env::BatchedSingleRoom, ticket::Integer)
    r = Ref{Ptr{Float32}}(C_NULL)
    d = Ref{Ptr{UInt8}}(C_NULL)
    check(ccall((:rcw_wait, LIB), Int32, (Ptr{Cvoid}, Int64, Ref{Ptr{Float32}}, Ref{Ptr{UInt8}}),
                env.handle, Int64(ticket), r, d))
    return unsafe_wrap(Array, r[], env.num_envs; own = false), unsafe_wrap(Array, d[], env.num_envs; own = false)
end

# act! for the envs env0+1 : env0+length(actions) only (0-based env0); with an observation window
# (`obs_window_envs`) this is how a learner walks a batch whose observations do not fit in HBM
function act_range!(env::BatchedSingleRoom, actions::Vector{UInt8}, env0::Integer)
    GC.@preserve actions begin
        check(ccall((:rcw_step_range, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64),
                    env.handle, actions, Int64(env0), Int64(length(actions))))
    end
    return nothing
end

# random policy on the device (benchmark path)
step_random!(env::BatchedSingleRoom, n_steps = 1) =
    check(ccall((:rcw_step_random, LIB), Int32, (Ptr{Cvoid}, Int32), env.handle, Int32(n_steps)))

# size(actions, 2) steps driven by an action tape: actions[e, s] is the action of env e in step s (num_envs x n_steps,
# column-major = the C ABI's [n_steps][num_envs]); a multi-step call, consecutive steps may overlap on the device
function act_tape!(env::BatchedSingleRoom, actions::Matrix{UInt8})
    size(actions, 1) == env.num_envs || throw(DimensionMismatch("one row per env"))
    GC.@preserve actions begin
        check(ccall((:rcw_step_tape, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32), env.handle, actions, Int32(size(actions, 2))))
    end
    return nothing
end

# cast_rays! + update_camera_view! — src/single_room.jl:195-231, 374-444
function RCW.cast_rays!(env::BatchedSingleRoom)
    check(ccall((:rcw_render, LIB), Int32, (Ptr{Cvoid},), env.handle))
    return nothing
end
RCW.update_camera_view!(env::BatchedSingleRoom) = nothing   # fused into cast_rays! above

# update_top_view! — src/single_room.jl:446-483, for every env (UInt32 images on the device)
function RCW.update_top_view!(env::BatchedSingleRoom)
    check(ccall((:rcw_render_top_view, LIB), Int32, (Ptr{Cvoid},), env.handle))
    return nothing
end

RCW.get_action_names(env::BatchedSingleRoom) = (:MOVE_FORWARD, :MOVE_BACKWARD, :TURN_LEFT, :TURN_RIGHT)  # :486

function reward_done!(env::BatchedSingleRoom)
    r, d = env.reward, env.done
    GC.@preserve r d begin
        check(ccall((:rcw_get_state, LIB), Int32,
                    (Ptr{Cvoid}, Ptr{Float32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Ptr{UInt8}),
                    env.handle, C_NULL, C_NULL, C_NULL, r, d))
    end
    return r, d
end

# borrowed device pointer of the observations (valid until the next act!/reset!), RLBase.state below
function obs_device_ptr(env::BatchedSingleRoom)
    ptr = Ref{Ptr{Cvoid}}(C_NULL); total = Ref{Csize_t}(0); stride = Ref{Csize_t}(0)
    check(ccall((:rcw_obs_device_ptr, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ref{Csize_t}, Ref{Csize_t}),
                env.handle, ptr, total, stride))
    return ptr[], Int(total[]), Int(stride[])
end

# dense host array for n envs in the handle's observation format:
# RGB8 UInt8[3, P, R, n] | XRGB32 UInt32[P, R, n] (the reference's camera_view, :300) | GRAY8 UInt8[P, R, n] | GRAY16F Float16[P, R, n] | COLUMNS UInt32[R, n]
function _host_obs(env::BatchedSingleRoom, n, fmt = env.obs_format)
    P, R = env.height_camera_view_pu, env.num_rays
    fmt == RCW_OBS_RGB8 && return Array{UInt8}(undef, 3, P, R, n)
    fmt == RCW_OBS_XRGB32 && return Array{UInt32}(undef, P, R, n)
    fmt == RCW_OBS_GRAY8 && return Array{UInt8}(undef, P, R, n)
    fmt == RCW_OBS_GRAY16F && return Array{Float16}(undef, P, R, n)
    fmt == RCW_OBS_GRAY8_HALF && return Array{UInt8}(undef, P ÷ 2, R ÷ 2, n)
    return Array{UInt32}(undef, R, n)
end

# host copy of the observations of envs e0:e0+n-1 (1-based)
function copy_obs(env::BatchedSingleRoom, e0 = 1, n = env.num_envs)
    out = _host_obs(env, n)
    GC.@preserve out begin
        check(ccall((:rcw_copy_obs, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}),
                    env.handle, Int64(e0 - 1), Int64(n), out))
    end
    return out
end

# the same for the frame of the given age of a frame ring (`frame_stack = K`): 0 = newest ... K - 1 = oldest
function copy_obs_frame(env::BatchedSingleRoom, age::Integer, e0 = 1, n = env.num_envs)
    out = _host_obs(env, n)
    GC.@preserve out begin
        check(ccall((:rcw_copy_obs_frame, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int32, Ptr{Cvoid}),
                    env.handle, Int64(e0 - 1), Int64(n), Int32(age), out))
    end
    return out
end

# (frame_stack K, newest ring position (0-based), bytes between two positions of one env): the frame with
# age a is at position mod(newest - a, K) of every env of the device buffer
function obs_frames(env::BatchedSingleRoom)
    k = Ref{Int32}(0); newest = Ref{Int32}(0); stride = Ref{Csize_t}(0)
    check(ccall((:rcw_obs_frames, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}, Ref{Csize_t}),
                env.handle, k, newest, stride))
    return Int(k[]), Int(newest[]), Int(stride[])
end

# device layout of one env of the observation buffer:
# (env_stride_bytes, column_stride_bytes, column_bytes, bytes_per_pixel); columns are pitched to 32 bytes
function obs_layout(env::BatchedSingleRoom)
    es = Ref{Csize_t}(0); cs = Ref{Csize_t}(0); cb = Ref{Csize_t}(0); bpp = Ref{Int32}(0)
    check(ccall((:rcw_obs_layout, LIB), Int32, (Ptr{Cvoid}, Ref{Csize_t}, Ref{Csize_t}, Ref{Csize_t}, Ref{Int32}),
                env.handle, es, cs, cb, bpp))
    return Int(es[]), Int(cs[]), Int(cb[]), Int(bpp[])
end

# Column words -> pixels on the device (the second half of update_camera_view!, single_room.jl:413-441).
# `columns`: device pointer to n envs' UInt32 words (this handle's own observations when it was created with
# obs_format = RCW_OBS_COLUMNS, or records a replay buffer kept), `columns_env_stride` bytes apart (0 = dense);
# `dst`: device pointer to n images in `pixel_format`, laid out as `expanded_layout` says.  Only enqueues.
function expand_columns!(env::BatchedSingleRoom, dst::Ptr{Cvoid}, columns::Ptr{Cvoid}, n::Integer;
                         pixel_format = RCW_OBS_RGB8, columns_env_stride = 0)
    check(ccall((:rcw_expand_columns, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Int64, Int32, Ptr{Cvoid}),
                env.handle, columns, Csize_t(columns_env_stride), Int64(n), Int32(pixel_format), dst))
    return nothing
end

# (env_stride_bytes, column_stride_bytes, column_bytes) of the images expand_columns! writes
function expanded_layout(env::BatchedSingleRoom, pixel_format = RCW_OBS_RGB8)
    es = Ref{Csize_t}(0); cs = Ref{Csize_t}(0); cb = Ref{Csize_t}(0)
    check(ccall((:rcw_expanded_layout, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Csize_t}, Ref{Csize_t}, Ref{Csize_t}),
                env.handle, Int32(pixel_format), es, cs, cb))
    return Int(es[]), Int(cs[]), Int(cb[])
end

# host copy of the top views of envs e0:e0+n-1 (1-based): UInt32[H * pu, W * pu, n], the reference's top_view layout
function copy_top_view(env::BatchedSingleRoom, height_pu::Integer, width_pu::Integer, e0 = 1, n = env.num_envs)
    out = Array{UInt32}(undef, height_pu, width_pu, n)
    GC.@preserve out begin
        check(ccall((:rcw_copy_top_view, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}),
                    env.handle, Int64(e0 - 1), Int64(n), out))
    end
    return out
end

# borrowed device pointer of the top views: UInt32 [env slots][W * pu columns][H * pu rows]
function top_view_device_ptr(env::BatchedSingleRoom)
    ptr = Ref{Ptr{Cvoid}}(C_NULL); total = Ref{Csize_t}(0); stride = Ref{Csize_t}(0)
    check(ccall((:rcw_top_view_device_ptr, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ref{Csize_t}, Ref{Csize_t}),
                env.handle, ptr, total, stride))
    return ptr[], Int(total[]), Int(stride[])
end

# ---- world fields (parity injection / inspection): the reference reads and writes them as
#      world.player_position_wu, .player_direction_au, .goal_position, .reward, .done (single_room.jl:21-40)

# (pos 2 x n Float32 [wu], dir n Int32 [au, 0-based], goal 2 x n Int32 [1-based tile], reward n Float32, done n UInt8)
function get_state(env::BatchedSingleRoom)
    n = env.num_envs
    pos = Matrix{Float32}(undef, 2, n); dir = Vector{Int32}(undef, n); goal = Matrix{Int32}(undef, 2, n)
    r = Vector{Float32}(undef, n); d = Vector{UInt8}(undef, n)
    GC.@preserve pos dir goal r d begin
        check(ccall((:rcw_get_state, LIB), Int32,
                    (Ptr{Cvoid}, Ptr{Float32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Ptr{UInt8}),
                    env.handle, pos, dir, goal, r, d))
    end
    return (pos = pos, dir = dir, goal = goal, reward = r, done = d)
end

_ptr_or_null(::Nothing, T) = Ptr{T}(C_NULL)
_ptr_or_null(a::Array{T}, ::Type{T}) where {T} = pointer(a)

# overwrite any subset of the fields (`nothing` = keep); does not re-render: follow with RCW.cast_rays!(env)
function set_state!(env::BatchedSingleRoom; pos = nothing, dir = nothing, goal = nothing, reward = nothing, done = nothing)
    GC.@preserve pos dir goal reward done begin
        check(ccall((:rcw_set_state, LIB), Int32,
                    (Ptr{Cvoid}, Ptr{Float32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Ptr{UInt8}),
                    env.handle, _ptr_or_null(pos, Float32), _ptr_or_null(dir, Int32), _ptr_or_null(goal, Int32),
                    _ptr_or_null(reward, Float32), _ptr_or_null(done, UInt8)))
    end
    return nothing
end

# world.ray_stop_position_tu (2 x R x n), ray_hit_dimension (R x n), ray_distance_wu (R x n), ray_directions_wu
# (2 x R x n) of envs e0:e0+n-1 (1-based), single_room.jl:29-31,39 — debug / parity only
function get_rays(env::BatchedSingleRoom, e0 = 1, n = env.num_envs)
    R = env.num_rays
    hit = Array{Int32}(undef, 2, R, n); dim = Matrix{Int32}(undef, R, n)
    dist = Matrix{Float32}(undef, R, n); ray = Array{Float32}(undef, 2, R, n)
    GC.@preserve hit dim dist ray begin
        check(ccall((:rcw_get_rays, LIB), Int32,
                    (Ptr{Cvoid}, Int64, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Ptr{Float32}),
                    env.handle, Int64(e0 - 1), Int64(n), hit, dim, dist, ray))
    end
    return (hit = hit, dim = dim, dist = dist, ray = ray)
end

# tile_map[WALL, :, :] (single_room.jl:55-60) supplied by the host.  `wall`: H x W (one layer shared by the
# batch) or H x W x num_envs (one per env), nonzero / true = wall; Julia's column-major order is the ABI's.
# Does not re-render: follow with RCW.reset!(env) or RCW.cast_rays!(env).
function set_wall_map!(env::BatchedSingleRoom, wall::AbstractArray{<:Union{Bool, Integer}})
    H, W = env.height_tile_map_tu, env.width_tile_map_tu
    bytes = convert(Array{UInt8}, wall .!= 0)
    if size(bytes) == (H, W)
        GC.@preserve bytes check(ccall((:rcw_set_wall_map, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}), env.handle, bytes))
    elseif size(bytes) == (H, W, env.num_envs)
        GC.@preserve bytes check(ccall((:rcw_set_wall_maps, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}), env.handle, bytes))
    else
        throw(DimensionMismatch("wall layer must be $(H) x $(W) or $(H) x $(W) x $(env.num_envs)"))
    end
    return nothing
end
# tile_map[layer, :, :] = tiles for every env: layer 1 = WALL, 3 .. num_object_layers = extra object layers
# (tiles: height_tu x width_tu Bool matrix, column-major like the reference's BitArray slices)
function set_layer!(env::BatchedSingleRoom, layer::Integer, tiles::AbstractMatrix{Bool})
    size(tiles) == (env.height_tile_map_tu, env.width_tile_map_tu) || throw(DimensionMismatch("tiles must be height_tu x width_tu"))
    bytes = convert(Array{UInt8}, tiles)
    GC.@preserve bytes check(ccall((:rcw_set_layer, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt8}), env.handle, Int32(layer), bytes))
    return nothing
end


# reset!(env) for the envs whose mask entry is true only (the others keep state and observation)
function reset_masked!(env::BatchedSingleRoom, mask::AbstractVector{Bool})
    length(mask) == env.num_envs || throw(DimensionMismatch("one mask entry per env"))
    m = convert(Vector{UInt8}, mask)
    GC.@preserve m begin
        check(ccall((:rcw_reset, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{UInt8}),
                    env.handle, C_NULL, C_NULL, C_NULL, m))
    end
    return nothing
end

# blocks until everything enqueued on the handle has finished (and reports a deferred RCW_EACTION)
sync(env::BatchedSingleRoom) = check(ccall((:rcw_sync, LIB), Int32, (Ptr{Cvoid},), env.handle))

# the handle's cudaStream_t, e.g. to build a CUDA.CuStream for work that must follow the step
function cuda_stream(env::BatchedSingleRoom)
    s = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rcw_stream, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), env.handle, s))
    return s[]
end

function launch_count(env::BatchedSingleRoom)
    n = Ref{Int64}(0)
    check(ccall((:rcw_launch_count, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), env.handle, n))
    return n[]
end

# exact snapshot of the dynamic state (checkpoint / resume)
function save_checkpoint(env::BatchedSingleRoom)
    n = Ref{Csize_t}(0)
    check(ccall((:rcw_checkpoint_size, LIB), Int32, (Ptr{Cvoid}, Ref{Csize_t}), env.handle, n))
    buf = Vector{UInt8}(undef, n[])
    GC.@preserve buf begin
        check(ccall((:rcw_save_checkpoint, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t), env.handle, buf, n[]))
    end
    return buf
end

function load_checkpoint!(env::BatchedSingleRoom, buf::Vector{UInt8})
    GC.@preserve buf begin
        check(ccall((:rcw_load_checkpoint, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t), env.handle, buf, length(buf)))
    end
    return nothing
end

function episode_stats(env::BatchedSingleRoom; reset_counters = false)
    ep = Ref{Int64}(0); sr = Ref{Float64}(0); sl = Ref{Int64}(0)
    check(ccall((:rcw_episode_stats, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Float64}, Ref{Int64}, Int32),
                env.handle, ep, sr, sl, Int32(reset_counters)))
    return ep[], sr[], sl[]
end

# RLBase API — same shape as src/single_room.jl:574-584, vector-valued for the batch
RLBase.StateStyle(env::RCW.RLBaseEnv{E}) where {E <: BatchedSingleRoom} = RLBase.Observation{Any}()
RLBase.state_space(env::RCW.RLBaseEnv{E}, ::RLBase.Observation) where {E <: BatchedSingleRoom} = nothing
RLBase.state(env::RCW.RLBaseEnv{E}, ::RLBase.Observation) where {E <: BatchedSingleRoom} = obs_device_ptr(env.env)
RLBase.reset!(env::RCW.RLBaseEnv{E}) where {E <: BatchedSingleRoom} = RCW.reset!(env.env)
RLBase.action_space(env::RCW.RLBaseEnv{E}) where {E <: BatchedSingleRoom} = Base.OneTo(4)
(env::RCW.RLBaseEnv{E})(actions) where {E <: BatchedSingleRoom} = RCW.act!(env.env, actions)
RLBase.reward(env::RCW.RLBaseEnv{E}) where {E <: BatchedSingleRoom} = reward_done!(env.env)[1]
RLBase.is_terminated(env::RCW.RLBaseEnv{E}) where {E <: BatchedSingleRoom} = reward_done!(env.env)[2] .!= 0

end # module
