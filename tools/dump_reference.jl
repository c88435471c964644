# dump_reference.jl — run the REAL reference (RayCastWorlds.jl + RayCaster.jl 0.1) over the states of
# tests/golden/singleroom_golden.npz and write what it computes, so that the parity claims of this
# repo can be pinned against Julia.  CANNOT RUN IN THIS IMAGE (no julia); run it on any machine with
#   julia --project -e 'import Pkg; Pkg.add(["RayCastWorlds", "NPZ"])'
#   julia --project tools/dump_reference.jl tests/golden/singleroom_golden.npz tests/golden/julia_reference.npz
# then `python tools/resolve_julia_pin.py` (or `pytest tests/test_julia_pin.py`) evaluates the oracle under all four
# settings of the two unpinned cast_ray decisions (D1 tie rule, D2 distance form), says which one reproduces this
# dump and — if it is not the engine's default — exactly which default to change.
# Until someone does, DESIGN.md and the oracle header say "parity unpinned".

import NPZ
import RayCastWorlds as RCW
import StaticArrays as SA

const SR = RCW.SingleRoomModule

function world_for(case)
    case == "A" && return SR.SingleRoom()
    case == "B" && return SR.SingleRoom(height_tile_map_tu = 64, width_tile_map_tu = 64, num_directions = 256,
                                        num_rays = 128, height_camera_view_pu = 96, pu_per_tu = 4)
    case == "C" && return SR.SingleRoom(height_tile_map_tu = 5, width_tile_map_tu = 7, num_directions = 36,
                                        num_rays = 45, height_camera_view_pu = 51, player_radius_wu = 0.2f0,
                                        position_increment_wu = 0.3f0, semi_field_of_view_wu = 0.5f0,
                                        camera_height_tile_wu = 0.8f0, pu_per_tu = 7)
    case == "T" && return SR.SingleRoom(height_tile_map_tu = 7, width_tile_map_tu = 7, num_directions = 8, num_rays = 33,
                                        height_camera_view_pu = 40, pu_per_tu = 4)   # exact ties: resolves D1
    error("unknown case")
end

function set_state!(env, x, y, au, gi, gj)
    w = env.world
    w.tile_map[SR.GOAL, w.goal_position] = false          # the struct is mutable (single_room.jl:21)
    w.goal_position = CartesianIndex(Int(gi), Int(gj))
    w.tile_map[SR.GOAL, w.goal_position] = true
    w.player_position_wu = SA.SVector(Float32(x), Float32(y))
    w.player_direction_au = Int(au)
    w.reward = 0f0
    w.done = false
end

function main(in_path, out_path)
    g = NPZ.npzread(in_path)
    out = Dict{String, Any}()
    for case in ("A", "B", "C", "T")
        env = world_for(case)
        states, au, goal = g["$(case)_states"], g["$(case)_au"], g["$(case)_goal"]
        n, R = size(states, 1), length(env.world.ray_directions_wu)
        hit = zeros(Int32, n, R, 2); dim = zeros(Int32, n, R); dist = zeros(Float32, n, R)
        rdir = zeros(Float32, n, R, 2); img = zeros(UInt32, n, R, size(env.camera_view, 1))
        top = zeros(UInt32, n, size(env.top_view, 2), size(env.top_view, 1))   # update_top_view! (SimpleDraw 0.3 shapes)
        for k in 1:n
            set_state!(env, states[k, 1], states[k, 2], au[k], goal[k, 1], goal[k, 2])
            RCW.cast_rays!(env.world)
            RCW.update_camera_view!(env)
            RCW.update_top_view!(env)
            for i in 1:R
                hit[k, i, 1] = env.world.ray_stop_position_tu[1, i]
                hit[k, i, 2] = env.world.ray_stop_position_tu[2, i]
                dim[k, i] = env.world.ray_hit_dimension[i]
                dist[k, i] = env.world.ray_distance_wu[i]
                rdir[k, i, 1], rdir[k, i, 2] = env.world.ray_directions_wu[i]
            end
            img[k, :, :] = permutedims(env.camera_view)     # [column, row]
            top[k, :, :] = permutedims(env.top_view)
        end
        out["$(case)_hit"] = hit; out["$(case)_dim"] = dim; out["$(case)_dist"] = dist
        out["$(case)_ray_dir"] = rdir; out["$(case)_image"] = img; out["$(case)_top"] = top
        # act! trajectories
        haskey(g, "$(case)_act_init") || continue
        init, actions = g["$(case)_act_init"], g["$(case)_act_actions"]
        ne, T = size(actions)
        pos = zeros(Float32, ne, T, 2); dir = zeros(Int32, ne, T); rew = zeros(Float32, ne, T); done = zeros(UInt8, ne, T)
        for e in 1:ne
            gi, gj, pi, pj, a0 = init[e, :]
            set_state!(env, pi - 0.5f0, pj - 0.5f0, a0, gi, gj)
            for t in 1:T
                RCW.act!(env.world, Int(actions[e, t]))
                pos[e, t, 1], pos[e, t, 2] = env.world.player_position_wu
                dir[e, t] = env.world.player_direction_au
                rew[e, t] = env.world.reward
                done[e, t] = env.world.done
            end
        end
        out["$(case)_act_pos"] = pos; out["$(case)_act_au"] = dir
        out["$(case)_act_reward"] = rew; out["$(case)_act_done"] = done
    end
    NPZ.npzwrite(out_path, out)
    println("wrote ", out_path)
end

main(ARGS[1], ARGS[2])
