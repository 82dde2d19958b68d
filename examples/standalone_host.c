/* A host that uses ONLY libb200mix (no OpenAL Soft objects): an HRTF device set up from the .mhr
 * data set, one looping mono source orbiting the listener, the library's own parameter stage per
 * update (b200mix_calc_listener_params / b200mix_calc_voice), the HRIR blend on the GPU
 * (b200mix_voices_update_dirs), a 1024-frame render per update.
 *
 * build: gcc -O2 -I include examples/standalone_host.c -L openal-soft_b200 -lb200mix -lm -o standalone_host
 * run:   LD_LIBRARY_PATH=openal-soft_b200 ./standalone_host "openal-soft_b200/data/Default HRTF.mhr"
 * Without a CUDA device b200mix_create fails (there is no CPU path) and the program says so. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b200mix.h"

#define CHECK(call) do { int rc_ = (call); if(rc_ != B200MIX_OK) { \
    fprintf(stderr, "%s -> %d (%s)\n", #call, rc_, b200mix_last_error(dev)); return 1; } } while(0)

int main(int argc, char **argv)
{
    b200mix_device *dev = NULL;
    if(argc < 2) { fprintf(stderr, "usage: %s <file.mhr>\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if(!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    void *mhr = malloc((size_t)sz);
    if(fread(mhr, 1, (size_t)sz, f) != (size_t)sz) return 2;
    fclose(f);

    b200mix_hrtf *hrtf = NULL;
    if(b200mix_hrtf_load(mhr, (size_t)sz, &hrtf) != B200MIX_OK) { fprintf(stderr, "not a MinPHR03 file\n"); return 2; }
    uint32_t rate = 0, ir = 0, count = 0;
    b200mix_hrtf_info(hrtf, &rate, &ir, &count);

    /* the HRTF decoder of a first-order device, from the data set alone */
    static float dec[4*128*2]; float hf[4], sc1; uint32_t dec_ir = 0;
    if(b200mix_hrtf_build_decoder(hrtf, 1, ir, &dec_ir, dec, hf, &sc1) != 4) return 2;
    const float sc[4] = {sc1, sc1, sc1, sc1};

    b200mix_device_desc d; memset(&d, 0, sizeof(d));
    d.struct_size = sizeof(d); d.cuda_device = -1; d.sample_rate = rate;
    d.dry_channels = 4; d.real_channels = 2; d.num_sends = 0; d.wet_channels = 0; d.ir_size = ir;
    d.post_process = B200MIX_POST_HRTF; d.real_left = 0; d.real_right = 1;
    d.max_voices = 1; d.max_buffers = 1; d.max_slots = 0;
    int rc = b200mix_create(&d, &dev);
    if(rc != B200MIX_OK)
    {
        printf("b200mix_create -> %d: %s\n(the mixer has no CPU path; run this on a CUDA machine)\n", rc,
            b200mix_last_error(NULL));
        b200mix_hrtf_free(hrtf); free(mhr);
        return rc == B200MIX_ERR_CUDA ? 0 : 1;
    }
    CHECK(b200mix_set_hrtf_decoder(dev, 4, dec_ir, dec, hf, sc));
    CHECK(b200mix_hrtf_attach(dev, hrtf));

    /* one second of a 440 Hz tone as the source's buffer */
    static int16_t pcm[48000];
    for(int i = 0;i < 48000;++i) pcm[i] = (int16_t)(8000.0*sin(2.0*3.14159265358979*440.0*i/48000.0));
    CHECK(b200mix_buffer_data(dev, 0, B200MIX_FMT_I16, 1, 48000, pcm, sizeof(pcm)));

    b200mix_listener_props lp; memset(&lp, 0, sizeof(lp));
    lp.struct_size = sizeof(lp); lp.orient_at[2] = -1.0f; lp.orient_up[1] = 1.0f;
    lp.gain = 1.0f; lp.gain_boost = 1.0f; lp.meters_per_unit = 1.0f; lp.air_absorption_gain_hf = 0.99426f;
    lp.doppler_factor = 1.0f; lp.doppler_velocity = 1.0f; lp.speed_of_sound = 343.3f;
    lp.distance_model = 2;                       /* InverseClamped */
    b200mix_listener_params lis;
    CHECK(b200mix_calc_listener_params(&lp, &lis));

    b200mix_source_props sp; memset(&sp, 0, sizeof(sp));
    sp.struct_size = sizeof(sp); sp.pitch = 1.0f; sp.gain = 1.0f; sp.max_gain = 1.0f;
    sp.inner_angle = 360.0f; sp.outer_angle = 360.0f; sp.ref_distance = 1.0f; sp.max_distance = 1e9f;
    sp.rolloff_factor = 1.0f; sp.distance_model = 2; sp.doppler_factor = 1.0f;
    sp.direct.gain = 1.0f; sp.direct.gain_hf = 1.0f; sp.direct.gain_lf = 1.0f;
    sp.direct.hf_reference = 5000.0f; sp.direct.lf_reference = 250.0f;

    b200mix_voice_env env; memset(&env, 0, sizeof(env));
    env.struct_size = sizeof(env); env.device_rate = rate; env.num_sends = 0; env.render_mode = 2;

    static float left[1024], right[1024];
    float *outs[2] = {left, right};
    for(int u = 0;u < 8;++u)
    {
        const float ang = 0.4f*(float)u;
        sp.position[0] = 2.0f*sinf(ang); sp.position[1] = 0.0f; sp.position[2] = -2.0f*cosf(ang);
        b200mix_voice_params vp; memset(&vp, 0, sizeof(vp));
        vp.voice = 0; vp.buffer = 0; vp.resampler = B200MIX_RESAMPLER_BSINC24;
        vp.flags = B200MIX_VF_PLAYING | B200MIX_VF_STATIC | B200MIX_VF_LOOPING | (u == 0 ? B200MIX_VF_RESET : 0u);
        vp.loop_start = 0; vp.loop_end = 48000;
        for(int s = 0;s < (int)B200MIX_MAX_SENDS;++s) vp.send_slot[s] = B200MIX_NO_SLOT;
        float dir[4];
        b200mix_voice_filter filt[1 + B200MIX_MAX_SENDS];
        CHECK(b200mix_calc_voice(&sp, &lis, &env, 48000, &vp, dir, NULL, NULL, filt));
        CHECK(b200mix_voices_update_dirs(dev, 1, &vp, dir, NULL, NULL));
        CHECK(b200mix_render(dev, 1024, outs, NULL));
        double el = 0.0, er = 0.0;
        for(int i = 0;i < 1024;++i) { el += left[i]*left[i]; er += right[i]*right[i]; }
        printf("update %d: azimuth %+5.1f deg, rms L %.4f R %.4f\n", u, ang*57.29578f, sqrt(el/1024), sqrt(er/1024));
    }
    b200mix_destroy(dev);
    b200mix_hrtf_free(hrtf); free(mhr);
    return 0;
}
