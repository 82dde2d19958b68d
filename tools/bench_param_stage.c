/* Times the host parameter stage (b200mix_calc_voice) on one core: sources per second.
 * build: gcc -O2 -I include tools/bench_param_stage.c -L openal-soft_b200 -lb200mix -lm -o /tmp/bps
 * run:   LD_LIBRARY_PATH=openal-soft_b200 /tmp/bps */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "b200mix.h"

int main(void)
{
    enum { N = 200000 };
    b200mix_listener_props lp; memset(&lp, 0, sizeof(lp));
    lp.struct_size = sizeof(lp); lp.orient_at[2] = -1.0f; lp.orient_up[1] = 1.0f;
    lp.gain = 1.0f; lp.gain_boost = 1.0f; lp.meters_per_unit = 1.0f; lp.air_absorption_gain_hf = 0.99426f;
    lp.doppler_factor = 1.0f; lp.doppler_velocity = 1.0f; lp.speed_of_sound = 343.3f; lp.distance_model = 2;
    b200mix_listener_params lis;
    if(b200mix_calc_listener_params(&lp, &lis)) return 1;
    static float scale[4] = {1, 1, 1, 1}; static uint32_t index[4] = {0, 1, 2, 3};
    b200mix_voice_env env; memset(&env, 0, sizeof(env));
    env.struct_size = sizeof(env); env.device_rate = 48000; env.num_sends = 1; env.render_mode = 2;
    env.wet_stride = 4; env.dry.channels = 4; env.dry.scale = scale; env.dry.index = index;
    env.wet[0] = env.dry;
    b200mix_source_props *sp = calloc(N, sizeof(*sp));
    srand(1);
    for(int i = 0;i < N;++i)
    {
        b200mix_source_props *p = &sp[i];
        p->struct_size = sizeof(*p); p->pitch = 1.0f; p->gain = 1.0f; p->max_gain = 1.0f;
        p->inner_angle = 360.0f; p->outer_angle = 360.0f; p->ref_distance = 1.0f; p->max_distance = 1e9f;
        p->rolloff_factor = 1.0f; p->distance_model = 2; p->doppler_factor = 1.0f;
        for(int k = 0;k < 3;++k) p->position[k] = (rand()/(float)RAND_MAX - 0.5f)*40.0f;
        p->direct.gain = 1.0f; p->direct.gain_hf = 1.0f; p->direct.gain_lf = 1.0f;
        p->direct.hf_reference = 5000.0f; p->direct.lf_reference = 250.0f;
        p->sends[0].gain = 1.0f; p->sends[0].gain_hf = 1.0f; p->sends[0].gain_lf = 1.0f;
        p->sends[0].hf_reference = 5000.0f; p->sends[0].lf_reference = 250.0f; p->sends[0].active = 1;
        p->sends[0].slot_decay_time = 1.49f; p->sends[0].slot_air_absorption_gain_hf = 0.994f;
        p->air_absorption_factor = i & 1 ? 1.0f : 0.0f; p->dry_gain_hf_auto = p->wet_gain_auto = p->wet_gain_hf_auto = 1;
    }
    struct timespec t0, t1;
    double acc = 0.0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for(int i = 0;i < N;++i)
    {
        b200mix_voice_params vp; memset(&vp, 0, sizeof(vp));
        float dir[4], dry[4], send[4];
        b200mix_voice_filter f[1 + B200MIX_MAX_SENDS];
        if(b200mix_calc_voice(&sp[i], &lis, &env, 48000, &vp, dir, dry, send, f) != 0) return 2;
        acc += dir[0] + vp.hrtf_gain + f[0].lowpass[0];
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double s = (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec)*1e-9;
    printf("b200mix_calc_voice: %d sources in %.3f s = %.0f ns/source, %.2f M sources/s on one core (checksum %g)\n",
        N, s, s/N*1e9, N/s*1e-6, acc);

    /* the threaded batch form */
    uint32_t *rates = malloc(N*sizeof(*rates));
    b200mix_voice_params *vps = calloc(N, sizeof(*vps));
    float *dirs = malloc((size_t)N*4*sizeof(float)), *dry = malloc((size_t)N*4*sizeof(float));
    float *snd = malloc((size_t)N*4*sizeof(float));
    b200mix_voice_filter *fl = malloc((size_t)N*2*sizeof(*fl));
    for(int i = 0;i < N;++i) rates[i] = 48000;
    for(unsigned th = 1;th <= 16;th *= 2)
    {
        clock_gettime(CLOCK_MONOTONIC, &t0);
        if(b200mix_calc_voices(N, sp, &lis, &env, rates, vps, dirs, dry, snd, fl, th) != 0) return 3;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double sb = (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec)*1e-9;
        printf("b200mix_calc_voices, %2u thread(s): %.2f M sources/s\n", th, N/sb*1e-6);
    }
    free(rates); free(vps); free(dirs); free(dry); free(snd); free(fl);
    free(sp);
    return 0;
}
