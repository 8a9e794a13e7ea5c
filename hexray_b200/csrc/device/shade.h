// shade.h — textures, bump mapping, environment lookup, camera rays, light sampling
// (device functions; pure, no queue access).
//
//   image_filtered      Bitmap::getFilteredPixel           src/bitmap.cpp:66-81
//   tex_sample          CheckerTexture / BitmapTexture / Fresnel::sample
//                                                           src/shading.cpp:50-55, 147-157, 326-343
//   tex_modify_normal   BumpTexture / Bumps::modifyNormal   src/shading.cpp:345-356, 363-377
//   env_lookup          CubemapEnvironment::getEnvironment  src/environment.cpp:54-81
//   camera_ray          Camera::getScreenRay / getDOFScreenRay  src/camera.cpp:65-86
//   light_nth_sample    PointLight / RectLight::getNthSample    src/lights.h:89-93, src/lights.cpp:37-51
//   light_solid_angle   RectLight::getSolidAngle                src/lights.cpp:90-102
#pragma once
#include "scene_dev.h"
#include "rng.h"

namespace hxr {

HXR_HD f3 image_filtered(const DImage& im, float x, float y)
{
    if (!im.rgb || !im.w || !im.h || x < 0 || x >= im.w || y < 0 || y >= im.h) return mkc(0, 0, 0);
    int tx = (int)floorf(x);
    int ty = (int)floorf(y);
    if (tx < 0 || ty < 0) return mkc(0, 0, 0);
    int tx_next = (tx + 1) % im.w;
    int ty_next = (ty + 1) % im.h;
    float p = x - tx;
    float q = y - ty;
    return ldc(im.rgb + 3 * (ty * im.w + tx)) * ((1.0f - p) * (1.0f - q)) +
           ldc(im.rgb + 3 * (ty * im.w + tx_next)) * (p * (1.0f - q)) +
           ldc(im.rgb + 3 * (ty_next * im.w + tx)) * ((1.0f - p) * q) +
           ldc(im.rgb + 3 * (ty_next * im.w + tx_next)) * (p * q);
}

HXR_HD float fresnel_schlick(float NdotI, float ior)
{
    float t = (1.0f - ior) / (1.0f + ior);
    float f = (float)((double)t * (double)t);  // sqr() takes and returns double (src/util.h:37)
    float x = 1 - NdotI;
    return f + (1 - f) * powf(x, 5.0f);
}

HXR_HD f3 tex_sample(const DScene& sc, int ti, const d3& rayDir, const Hit& info)
{
    const hxr_texture& T = sc.textures[ti];
    switch (T.type) {
        case HXR_TEX_CHECKER: {
            int u1 = (int)floor(info.u / T.scaling);
            int v1 = (int)floor(info.v / T.scaling);
            return ((u1 + v1) % 2 == 0) ? ldc(T.color1) : ldc(T.color2);
        }
        case HXR_TEX_BITMAP: {
            const DImage& im = sc.images[T.image];
            float u = (float)(info.u / T.scaling);
            float v = (float)(info.v / T.scaling);
            u -= floorf(u);
            v -= floorf(v);
            u *= im.w;
            v *= im.h;
            return image_filtered(im, u, v);
        }
        case HXR_TEX_FRESNEL: {
            float eta = (float)T.ior;
            float NdotI = (float)dot(rayDir, info.norm);
            if (NdotI > 0) eta = 1 / eta; else NdotI = -NdotI;
            float fr = fresnel_schlick(NdotI, eta);
            return mkc(fr, fr, fr);
        }
        default: return mkc(0, 0, 0);  // BumpTexture / Bumps sample black
    }
}

HXR_HD void tex_modify_normal(const DScene& sc, int ti, Hit& info)
{
    const hxr_texture& T = sc.textures[ti];
    if (T.type == HXR_TEX_BUMP) {
        const DImage& im = sc.images[T.image];
        float x = (float)fmod(info.u * T.scaling * im.w, (double)im.w);
        float y = (float)fmod(info.v * T.scaling * im.h, (double)im.h);
        f3 bump = image_filtered(im, x, y);
        // bump.r * strength: float * double -> double, stored to float
        float dx = (float)(bump.r * T.strength);
        float dy = (float)(bump.g * T.strength);
        info.norm = normalize_m(info.norm + (info.dNdx * (double)dx + info.dNdy * (double)dy));
    } else if (T.type == HXR_TEX_BUMPS) {
        const float strength = (float)T.strength;
        if (strength > 0) {
            const float freqX[3] = {0.5f, 1.21f, 1.9f}, freqZ[3] = {0.4f, 1.13f, 1.81f};
            const float fm = 0.2f;
            const float intensityX[3] = {0.1f, 0.08f, 0.05f}, intensityZ[3] = {0.1f, 0.08f, 0.05f};
            double dx = 0, dy = 0;
            for (int i = 0; i < 3; i++) {
                dx += sin((double)(fm * freqX[i]) * info.u) * intensityX[i] * strength;
                dy += sin((double)(fm * freqZ[i]) * info.v) * intensityZ[i] * strength;
            }
            info.norm = normalize_m(info.norm + (dx * info.dNdx + dy * info.dNdy));
        }
    }
}

HXR_HD f3 env_side(const DImage& im, double x, double y)
{
    return image_filtered(im, (float)((x + 1) * 0.5 * (im.w - 1)), (float)((y + 1) * 0.5 * (im.h - 1)));
}

HXR_HD f3 env_lookup(const DScene& sc, const d3& dir)
{
    if (!sc.has_env) return mkc(0, 0, 0);
    int dim = max_dimension(dir);
    bool positive = comp(dir, dim) > 0;
    d3 s = div3(dir, fabs(comp(dir, dim)));
    int caseNum = (positive ? 3 : 0) + dim;
    const DImage& side = sc.images[sc.env_images[caseNum]];
    switch (caseNum) {
        case 0: return env_side(side, s.z, -s.y);
        case 1: return env_side(side, s.x, -s.z);
        case 2: return env_side(side, s.x, s.y);
        case 3: return env_side(side, -s.z, -s.y);
        case 4: return env_side(side, s.x, s.z);
        default: return env_side(side, s.x, -s.y);
    }
}

HXR_HD Ray camera_ray(const hxr_camera& c, double W, double H, double x, double y, double u, double v, double stereoOffset)
{
    Ray r;
    r.depth = 0;
    r.flags = 0;
    const d3 pos = ld3(c.pos), tl = ld3(c.top_left), tr = ld3(c.top_right), bl = ld3(c.bottom_left);
    r.o = pos;
    d3 through = tl + (tr - tl) * (x / W) + (bl - tl) * (y / H);
    r.d = normalize_m(through - r.o);
    if (stereoOffset != 0) r.o = r.o + ld3(c.right) * (stereoOffset * c.stereo_separation);
    if (c.dof) {
        double M = c.focal_plane_dist / dot(ld3(c.front), r.d);
        d3 T = r.o + r.d * M;
        r.o = pos + (u * c.aperture_size) * ld3(c.right) + (v * c.aperture_size) * ld3(c.up);
        r.d = normalize_f(T - r.o);
    }
    return r;
}

HXR_HD int light_num_samples(const hxr_light& L) { return L.type == HXR_LIGHT_RECT ? L.xsubd * L.ysubd : 1; }

HXR_HD void light_nth_sample(const hxr_light& L, int idx, const d3& shadePos, Rng& rng, d3& samplePos, f3& color)
{
    if (L.type == HXR_LIGHT_POINT) {
        samplePos = ld3(L.pos);
        color = ldc(L.color);
        return;
    }
    double lx = ((idx % L.xsubd) + rng.rand_double()) / L.xsubd;
    double ly = ((idx / L.xsubd) + rng.rand_double()) / L.ysubd;
    samplePos = mul_vm(mk3(lx - 0.5, -1e-6, ly - 0.5), L.T.m) + ld3(L.T.offset);
    d3 sl = mul_vm(shadePos - ld3(L.T.offset), L.T.inv);
    if (sl.y < 0) {
        // Color * double -> Color * float(multiplier), then / double -> operator/(Color, float)
        color = divc(ldc(L.color) * (float)(-sl.y), (float)length(sl));
    } else {
        color = mkc(0, 0, 0);
    }
}

HXR_HD double light_solid_angle(const hxr_light& L, const d3& p)
{
    if (L.type != HXR_LIGHT_RECT) return 0;
    d3 dirW = normalize_f(mul_vm(mk3(0, -1, 0), L.T.m));
    d3 posW = ld3(L.T.offset);
    d3 lightToP = p - posW;
    double cosTerm = dot(lightToP, dirW);
    if (cosTerm < 0) return 0;
    double d = length(lightToP);
    cosTerm /= d;
    return L.area * cosTerm / ((1 + d) * (1 + d));
}

// Light::getColor() = color * power (src/lights.h:36)
HXR_HD f3 light_color_power(const hxr_light& L) { return ldc(L.color) * L.power; }

}  // namespace hxr
