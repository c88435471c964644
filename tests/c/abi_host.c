/* A plain C host of librcw_b200.so: what a foreign-language binding (Julia ccall, cgo, JNI) does,
 * written in C so that it can be compiled and run here.  Built by tests/test_c_host.py with gcc,
 * linked against the library only (no CUDA headers, no Python).
 *
 * usage: abi_host <num_envs> <steps> <seed>
 * prints one line per checkpoint: step index, sum of x, sum of y, sum of directions, finished episodes,
 * and a checksum of the observations of env 0, which the test compares with the oracle. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rcw_b200.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int32_t rc_ = (call);                                                    \
        if (rc_ != RCW_OK) {                                                     \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, rcw_last_error());     \
            return 1;                                                            \
        }                                                                        \
    } while (0)

static uint32_t fnv1a(const uint8_t* p, size_t n) {
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 16777619u;
    return h;
}

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 16;
    const int steps = argc > 2 ? atoi(argv[2]) : 50;
    const uint64_t seed = argc > 3 ? strtoull(argv[3], NULL, 10) : 1;

    if (rcw_version() != RCW_ABI_VERSION) return 2;
    rcw_config cfg;
    CHECK(rcw_config_init(&cfg));
    cfg.num_envs = n;
    cfg.seed = seed;
    cfg.num_rays = 96;
    cfg.height_camera_view_pu = 64;
    cfg.result_ring = 2;   /* rewards / terminations arrive in pinned host memory, read one step behind */

    /* a struct from a different ABI revision is refused, not misread */
    rcw_config bad = cfg;
    bad.struct_size = 12;
    rcw_batch* b = NULL;
    if (rcw_create(&bad, NULL, &b) != RCW_ESIZE || b != NULL) return 3;

    CHECK(rcw_create(&cfg, NULL, &b));
    size_t env_stride, col_stride, col_bytes;
    int32_t bpp;
    CHECK(rcw_obs_layout(b, &env_stride, &col_stride, &col_bytes, &bpp));
    const size_t dense = (size_t)cfg.num_rays * col_bytes;
    uint8_t* obs = (uint8_t*)malloc(dense);
    uint8_t* actions = (uint8_t*)malloc((size_t)n);
    float* pos = (float*)malloc(sizeof(float) * 2 * (size_t)n);
    int32_t* dir = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    uint8_t* done = (uint8_t*)malloc((size_t)n);

    /* an invalid action is refused and nothing happens (the reference's @assert) */
    memset(actions, 1, (size_t)n);
    actions[n - 1] = 7;
    if (rcw_step(b, actions) != RCW_EACTION) return 4;

    double ring_reward = 0;
    long ring_done = 0;
    for (int s = 1; s <= steps; ++s) {
        for (int64_t e = 0; e < n; ++e) actions[e] = (uint8_t)(1 + ((e + s) % 7 == 0 ? 2 : 0) + ((e * 31 + s) % 11 == 0 ? 1 : 0));
        /* enqueue step s, then read the results of step s - 1 while step s runs */
        int64_t ticket;
        CHECK(rcw_step_async(b, actions, &ticket));
        if (ticket != s - 1) return 6;
        if (ticket > 0) {
            const float* r;
            const uint8_t* d;
            CHECK(rcw_wait(b, ticket - 1, &r, &d));
            for (int64_t e = 0; e < n; ++e) {
                ring_reward += r[e];
                ring_done += d[e];
            }
        }
        if (s % 10 == 0 || s == steps) {
            CHECK(rcw_get_state(b, pos, dir, NULL, NULL, done));
            CHECK(rcw_copy_obs(b, 0, 1, obs));
            double sx = 0, sy = 0;
            long sd = 0;
            for (int64_t e = 0; e < n; ++e) {
                sx += pos[2 * e];
                sy += pos[2 * e + 1];
                sd += dir[e];
            }
            int64_t episodes, sum_length;
            double sum_return;
            CHECK(rcw_episode_stats(b, &episodes, &sum_return, &sum_length, 0));
            printf("%d %.6f %.6f %ld %lld %u\n", s, sx, sy, sd, (long long)episodes, fnv1a(obs, dense));
        }
    }
    {
        const float* r;
        const uint8_t* d;
        CHECK(rcw_wait(b, steps - 1, &r, &d));
        for (int64_t e = 0; e < n; ++e) {
            ring_reward += r[e];
            ring_done += d[e];
        }
        if (rcw_wait(b, steps, &r, &d) != RCW_EINVAL) return 7;   /* not issued */
        printf("ring %.6f %ld\n", ring_reward, ring_done);
    }
    int64_t launches;
    CHECK(rcw_launch_count(b, &launches));
    if (launches < steps) return 5;

    /* One process, several handles (one per GPU in production; here the three shards share device 0): the same batch
     * cut into contiguous blocks of global env ids, stepped with the same action stream, must end in the same state,
     * with the same episode totals summed over the shards.  Shard 1's first env is compared pixel for pixel too. */
    {
        enum { SHARDS = 3 };
        int32_t devices[SHARDS] = {0, 0, 0};
        rcw_batch* shard[SHARDS];
        rcw_config scfg = cfg;
        scfg.result_ring = 0;
        CHECK(rcw_create_sharded(&scfg, NULL, devices, SHARDS, shard));
        for (int s = 1; s <= steps; ++s) {
            for (int64_t e = 0; e < n; ++e) actions[e] = (uint8_t)(1 + ((e + s) % 7 == 0 ? 2 : 0) + ((e * 31 + s) % 11 == 0 ? 1 : 0));
            CHECK(rcw_step_sharded(shard, SHARDS, actions));
        }
        CHECK(rcw_sync_sharded(shard, SHARDS));
        int64_t ep1, len1, ep2, len2;
        double ret1, ret2;
        CHECK(rcw_episode_stats(b, &ep1, &ret1, &len1, 0));
        CHECK(rcw_reduce_episode_stats(shard, SHARDS, &ep2, &ret2, &len2, 0));
        if (ep1 != ep2 || len1 != len2 || ret1 != ret2) return 8;
        float* spos = (float*)malloc(sizeof(float) * 2 * (size_t)n);
        CHECK(rcw_get_state(b, pos, NULL, NULL, NULL, NULL));
        int64_t off = 0, cnt = 0, off1 = 0;
        for (int k = 0; k < SHARDS; ++k) {
            CHECK(rcw_shard_envs(n, SHARDS, k, &off, &cnt));
            if (k == 1) off1 = off;
            CHECK(rcw_get_state(shard[k], spos + 2 * off, NULL, NULL, NULL, NULL));
        }
        if (memcmp(pos, spos, sizeof(float) * 2 * (size_t)n) != 0) return 9;
        uint8_t* obs1 = (uint8_t*)malloc(dense);
        CHECK(rcw_copy_obs(b, off1, 1, obs));
        CHECK(rcw_copy_obs(shard[1], 0, 1, obs1));
        if (memcmp(obs, obs1, dense) != 0) return 10;
        /* an invalid action anywhere: nothing is enqueued on any shard */
        actions[n - 1] = 9;
        if (rcw_step_sharded(shard, SHARDS, actions) != RCW_EACTION) return 11;
        CHECK(rcw_destroy_sharded(shard, SHARDS));
        free(spos);
        free(obs1);
        printf("sharded %lld %.6f %lld\n", (long long)ep2, ret2, (long long)len2);
    }
    CHECK(rcw_destroy(b));
    free(obs);
    free(actions);
    free(pos);
    free(dir);
    free(done);
    return 0;
}
