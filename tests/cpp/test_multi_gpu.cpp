// tests/cpp/test_multi_gpu.cpp — a C++ host (no Python, no torch) drives the library through include/hxr.h alone:
// parse cornell_box.hexray with the library's front-end, render it path traced on ONE GPU and on a context that owns
// SEVERAL GPUs, and require the two frames to agree (Philox counters make the N-GPU frame equal to the 1-GPU frame up to
// FP32 summation order). This is the replacement of the reference's ThreadPool fork-join (src/threading.cpp:54-97) seen
// from the host program's side (src/main.cpp:530-569).
//   usage: test_multi_gpu <scene.hexray> <n_gpus> [spp] [size]      (n_gpus contexts may share a device: "0,0")
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/hxr.h"

static int die(const char* what, hxr_ctx* ctx)
{
    fprintf(stderr, "FAIL %s: %s\n", what, hxr_last_error(ctx));
    return 1;
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s scene.hexray device,device[,...] [spp] [size]\n", argv[0]); return 2; }
    std::vector<int32_t> devices;
    for (char* tok = strtok(argv[2], ","); tok; tok = strtok(nullptr, ",")) devices.push_back(atoi(tok));
    const int spp = argc > 3 ? atoi(argv[3]) : 64, size = argc > 4 ? atoi(argv[4]) : 256;
    hxr_scene_file* sf = nullptr;
    if (hxr_scene_load(argv[1], &sf) != HXR_OK) return die("hxr_scene_load", nullptr);
    hxr_camera cam;
    if (hxr_scene_file_camera(sf, &cam) != HXR_OK) return die("camera", nullptr);
    hxr_render_params p;
    memset(&p, 0, sizeof p);
    p.width = p.height = size;
    p.mode = HXR_MODE_MONTECARLO;
    p.spp = spp;
    p.want_aa = -1;
    p.max_depth = -1;
    p.seed = 5;
    const size_t n = (size_t)size * size * 3;
    std::vector<float> one(n), many(n);
    hxr_stats s1, sN;
    for (int pass = 0; pass < 2; pass++) {
        hxr_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.device = devices[0];
        if (pass == 1) { cfg.n_devices = (int32_t)devices.size(); cfg.devices = devices.data(); }
        hxr_ctx* ctx = nullptr;
        if (hxr_create(&cfg, &ctx) != HXR_OK) return die("hxr_create", nullptr);
        if (hxr_upload_scene(ctx, hxr_scene_file_scene(sf)) != HXR_OK) return die("hxr_upload_scene", ctx);
        if (hxr_set_camera(ctx, &cam) != HXR_OK) return die("hxr_set_camera", ctx);
        if (hxr_render(ctx, &p, pass ? many.data() : one.data(), pass ? &sN : &s1) != HXR_OK) return die("hxr_render", ctx);
        if (pass) printf("reduce backend: %s\n", hxr_reduce_backend(ctx));
        hxr_destroy(ctx);
    }
    hxr_scene_file_free(sf);
    double maxd = 0, mean = 0;
    for (size_t i = 0; i < n; i++) {
        maxd = std::fmax(maxd, std::fabs((double)one[i] - many[i]));
        mean += one[i];
    }
    mean /= (double)n;
    printf("1 GPU: %.2f ms, %llu rays | %u GPUs: %.2f ms (reduce %.3f ms), %llu rays | mean %.4f max |diff| %.3g\n", s1.render_ms,
           (unsigned long long)(s1.rays_closest + s1.rays_shadow), sN.n_devices, sN.render_ms, sN.reduce_ms,
           (unsigned long long)(sN.rays_closest + sN.rays_shadow), mean, maxd);
    if (sN.n_devices != devices.size()) { fprintf(stderr, "FAIL: the frame was not rendered by %zu GPUs\n", devices.size()); return 1; }
    if (s1.rays_closest != sN.rays_closest) { fprintf(stderr, "FAIL: closest-hit ray counts differ\n"); return 1; }
    if (!(mean > 0.01)) { fprintf(stderr, "FAIL: black frame\n"); return 1; }
    if (!(maxd < 2e-4 * std::fmax(1.0, mean * 4))) { fprintf(stderr, "FAIL: frames differ\n"); return 1; }
    printf("OK\n");
    return 0;
}
