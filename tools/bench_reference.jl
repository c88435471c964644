# bench_reference.jl — the reference's own CPU path, Threads.@threads over independent envs, on the
# workload of bench.py (N default SingleRoom envs, random policy, camera view rendered every step).
# CANNOT RUN IN THIS IMAGE (no julia).  On a machine with Julia:
#   julia -t auto --project tools/bench_reference.jl 4096 20
# prints env-steps/s with and without the top-view redraw that RCW.act!(env) also performs
# (src/single_room.jl:337); bench.py's --impl reference arm times the C port of the same loop.

import RayCastWorlds as RCW
const SR = RCW.SingleRoomModule

function step_no_top_view!(env, a)
    RCW.act!(env.world, a)
    RCW.cast_rays!(env.world)
    RCW.update_camera_view!(env)
end

function main(n, steps)
    envs = [SR.SingleRoom() for _ in 1:n]
    for (name, f) in (("act! (with top view)", RCW.act!), ("act + cast + camera view", step_no_top_view!))
        for s in 1:2; Threads.@threads for e in 1:n; f(envs[e], rand(1:4)); end; end
        t = @elapsed for s in 1:steps
            Threads.@threads for e in 1:n
                f(envs[e], rand(1:4))
            end
        end
        println(name, ": ", round(n * steps / t), " env-steps/s on ", Threads.nthreads(), " threads")
    end
end

main(parse(Int, get(ARGS, 1, "4096")), parse(Int, get(ARGS, 2, "20")))
