/*
 * synth.c -- deterministic synthetic G-buffer sequence.
 *
 * Stands in for the reference's path tracer (shaders/ptRaygen.rgen:48-89), which is out of
 * scope: it emits exactly the planes that producer writes, in the same conventions, so the
 * denoising modules see what they would see behind the real renderer:
 *   depth        r32f   Euclidean camera->hit distance            (ptRaygen.rgen:49)
 *   normal       rg32f  (theta = acos n.z, phi = atan2(n.y, n.x)) (ptRaygen.rgen:53-56)
 *   material     rgba8  0
 *   albedo       rgba8  diffuse colour                             (ptRaygen.rgen:48)
 *   illumination rgba32f demodulated 1-spp radiance, min(., 1e3)   (ptRaygen.rgen:81-88)
 *   miss pixels: position 1e10*(1,1,1), normal (1,1,1), albedo 0   (ptMiss.rmiss:10-15)
 * Camera: vsg::Perspective(60, W/H, .1, 1000) and vsg::LookAt((0,-3,1),(0,0,1),(0,0,1))
 * (source/VulkanPBRT.cpp:314-316), orbiting 0.25 deg/frame with a 0.01/frame dolly.
 * Matrices are built in double like vsg does (external/vsg/include/vsg/maths/transform.h:
 * 148-183) and narrowed to float; primary rays use the float matrices with the formula of
 * shaders/camera.glsl:7-15 so the accumulator's reconstruction is self-consistent.
 *
 * Input tooling, not an oracle and not a fallback: nothing here denoises anything.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define SYNTH_API __attribute__((visibility("default")))

typedef struct {
    float view[16];
    float inv_view[16];
    float proj[16];
    float inv_proj[16];
} vkpbrt_synth_camera_t;

/* ---- double 4x4 helpers, column-major m[col*4+row] ---- */
static void dmul(const double* a, const double* b, double* r)
{
    for (int c = 0; c < 4; ++c)
        for (int i = 0; i < 4; ++i) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += a[k * 4 + i] * b[c * 4 + k];
            r[c * 4 + i] = s;
        }
}
static void dinverse(const double* m, double* inv)
{
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            a[r][c] = m[c * 4 + r];
            a[r][c + 4] = (r == c) ? 1.0 : 0.0;
        }
    for (int i = 0; i < 4; ++i) {
        int p = i;
        for (int r = i + 1; r < 4; ++r)
            if (fabs(a[r][i]) > fabs(a[p][i])) p = r;
        if (p != i)
            for (int c = 0; c < 8; ++c) { double t = a[i][c]; a[i][c] = a[p][c]; a[p][c] = t; }
        double d = 1.0 / a[i][i];
        for (int c = 0; c < 8; ++c) a[i][c] *= d;
        for (int r = 0; r < 4; ++r)
            if (r != i) {
                double f = a[r][i];
                for (int c = 0; c < 8; ++c) a[r][c] -= f * a[i][c];
            }
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) inv[c * 4 + r] = a[r][c + 4];
}

SYNTH_API void vkpbrt_synth_camera(int W, int H, int frame, vkpbrt_synth_camera_t* out)
{
    const double pi = 3.14159265358979323846;
    /* vsg::perspective(radians(60), aspect, .1, 1000) */
    double fovy = 60.0 * pi / 180.0, aspect = (double)W / (double)H, zn = 0.1, zf = 1000.0;
    double f = 1.0 / tan(fovy * 0.5), r = 1.0 / (zn - zf);
    double P[16] = {f / aspect, 0, 0, 0, 0, -f, 0, 0, 0, 0, zf * r, -1, 0, 0, (zf * zn) * r, 0};
    /* orbit + dolly */
    double ang = 0.25 * pi / 180.0 * (double)frame;
    double rad = 3.0 - 0.01 * (double)frame;
    if (rad < 1.6) rad = 1.6;
    double eye[3] = {rad * sin(ang), -rad * cos(ang), 1.0}, ctr[3] = {0, 0, 1}, up[3] = {0, 0, 1};
    /* vsg::lookAt */
    double fw[3] = {ctr[0] - eye[0], ctr[1] - eye[1], ctr[2] - eye[2]};
    double l = sqrt(fw[0] * fw[0] + fw[1] * fw[1] + fw[2] * fw[2]);
    for (int i = 0; i < 3; ++i) fw[i] /= l;
    double s[3] = {fw[1] * up[2] - fw[2] * up[1], fw[2] * up[0] - fw[0] * up[2], fw[0] * up[1] - fw[1] * up[0]};
    l = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    for (int i = 0; i < 3; ++i) s[i] /= l;
    double u[3] = {s[1] * fw[2] - s[2] * fw[1], s[2] * fw[0] - s[0] * fw[2], s[0] * fw[1] - s[1] * fw[0]};
    l = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    for (int i = 0; i < 3; ++i) u[i] /= l;
    double R[16] = {s[0], u[0], -fw[0], 0, s[1], u[1], -fw[1], 0, s[2], u[2], -fw[2], 0, 0, 0, 0, 1};
    double T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, -eye[0], -eye[1], -eye[2], 1};
    double V[16], Vi[16], Pi[16];
    dmul(R, T, V);
    dinverse(V, Vi);
    dinverse(P, Pi);
    for (int i = 0; i < 16; ++i) {
        out->view[i] = (float)V[i];
        out->inv_view[i] = (float)Vi[i];
        out->proj[i] = (float)P[i];
        out->inv_proj[i] = (float)Pi[i];
    }
}

/* ---- scene ---- */
typedef struct { float c[3]; float r; float alb[3]; } sphere_t;
static const sphere_t spheres[5] = {
    {{-1.6f, 0.4f, 0.6f}, 0.6f, {0.85f, 0.25f, 0.2f}},
    {{0.0f, 0.0f, 0.8f}, 0.8f, {0.9f, 0.9f, 0.9f}},
    {{1.5f, 0.6f, 0.5f}, 0.5f, {0.2f, 0.4f, 0.85f}},
    {{0.9f, -1.1f, 0.3f}, 0.3f, {0.25f, 0.8f, 0.3f}},
    {{-0.8f, -1.2f, 0.25f}, 0.25f, {0.9f, 0.8f, 0.2f}},
};
static const float light_pos[3] = {2.5f, -2.0f, 5.0f};

static inline uint32_t hash_u32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
static inline float rnd01(uint32_t x, uint32_t y, uint32_t frame, uint32_t seed, uint32_t dim)
{
    uint32_t h = hash_u32(x + 0x9e3779b9u * hash_u32(y + 0x85ebca6bu * hash_u32(frame + 0xc2b2ae35u * hash_u32(seed + dim))));
    return (float)(h >> 8) * (1.0f / 16777216.0f);
}

/* nearest hit: returns t (>0) or -1; fills normal + albedo + id */
static float trace(const float* o, const float* d, float* n, float* alb, int* id)
{
    float best = 1e30f;
    *id = -1;
    /* ground plane z = 0 */
    if (d[2] < -1e-6f) {
        float t = -o[2] / d[2];
        if (t > 1e-4f && t < best) {
            float px = o[0] + t * d[0], py = o[1] + t * d[1];
            if (fabsf(px) < 12.0f && py > -12.0f && py < 6.0f) {
                best = t; *id = 0;
                n[0] = 0; n[1] = 0; n[2] = 1;
                int chk = ((int)floorf(px * 1.25f) + (int)floorf(py * 1.25f)) & 1;
                float a = chk ? 0.8f : 0.35f;
                alb[0] = a; alb[1] = a * 0.95f; alb[2] = a * 0.9f;
            }
        }
    }
    /* back wall y = 6, 0 <= z <= 4 */
    if (d[1] > 1e-6f) {
        float t = (6.0f - o[1]) / d[1];
        if (t > 1e-4f && t < best) {
            float px = o[0] + t * d[0], pz = o[2] + t * d[2];
            if (fabsf(px) < 12.0f && pz >= 0.0f && pz <= 4.0f) {
                best = t; *id = 1;
                n[0] = 0; n[1] = -1; n[2] = 0;
                int row = (int)floorf(pz * 2.0f);
                int brick = ((int)floorf(px * 1.0f + (row & 1) * 0.5f) + row) & 1;
                alb[0] = brick ? 0.7f : 0.55f; alb[1] = brick ? 0.35f : 0.3f; alb[2] = brick ? 0.3f : 0.25f;
            }
        }
    }
    for (int i = 0; i < 5; ++i) {
        float oc[3] = {o[0] - spheres[i].c[0], o[1] - spheres[i].c[1], o[2] - spheres[i].c[2]};
        float b = oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2];
        float c = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - spheres[i].r * spheres[i].r;
        float disc = b * b - c;
        if (disc <= 0) continue;
        float sq = sqrtf(disc);
        float t = -b - sq;
        if (t <= 1e-4f) t = -b + sq;
        if (t > 1e-4f && t < best) {
            best = t; *id = 2 + i;
            float inv = 1.0f / spheres[i].r;
            for (int k = 0; k < 3; ++k) n[k] = (o[k] + t * d[k] - spheres[i].c[k]) * inv;
            float stripe = (((int)floorf((o[2] + t * d[2]) * 6.0f)) & 1) ? 1.0f : 0.8f;
            for (int k = 0; k < 3; ++k) alb[k] = spheres[i].alb[k] * stripe;
        }
    }
    return (*id < 0) ? -1.0f : best;
}

static int occluded(const float* p, const float* l, float dist)
{
    for (int i = 0; i < 5; ++i) {
        float oc[3] = {p[0] - spheres[i].c[0], p[1] - spheres[i].c[1], p[2] - spheres[i].c[2]};
        float b = oc[0] * l[0] + oc[1] * l[1] + oc[2] * l[2];
        float c = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - spheres[i].r * spheres[i].r;
        float disc = b * b - c;
        if (disc <= 0) continue;
        float t = -b - sqrtf(disc);
        if (t > 1e-3f && t < dist) return 1;
    }
    return 0;
}

/* Renders image rows [row0, row1) of frame `frame` of a W x H sequence.  Plane pointers address
 * the FULL planes ([H][W]...), only the requested rows are written, so band-sharded ranks can
 * render their band plus apron locally. */
SYNTH_API void vkpbrt_synth_frame(int W, int H, int frame, uint32_t seed, int row0, int row1, float* depth,
                                  float* normal, uint8_t* albedo, uint8_t* material, float* illum)
{
    vkpbrt_synth_camera_t cam;
    vkpbrt_synth_camera(W, H, frame, &cam);
    const float* iv = cam.inv_view;
    const float* ip = cam.inv_proj;
    const float org[3] = {iv[12], iv[13], iv[14]};
    if (row0 < 0) row0 = 0;
    if (row1 > H) row1 = H;
#pragma omp parallel for schedule(dynamic, 8)
    for (int y = row0; y < row1; ++y)
        for (int x = 0; x < W; ++x) {
            size_t pix = (size_t)y * W + x;
            /* camera.glsl:7-15 */
            float cx = (((float)x + 0.5f) / (float)W) * 2.0f - 1.0f;
            float cy = (((float)y + 0.5f) / (float)H) * 2.0f - 1.0f;
            float vd[3];
            for (int i = 0; i < 3; ++i) vd[i] = ip[0 + i] * cx + ip[4 + i] * cy + ip[8 + i] + ip[12 + i];
            float il = 1.0f / sqrtf(vd[0] * vd[0] + vd[1] * vd[1] + vd[2] * vd[2]);
            for (int i = 0; i < 3; ++i) vd[i] *= il;
            float d[3];
            for (int i = 0; i < 3; ++i) d[i] = iv[0 + i] * vd[0] + iv[4 + i] * vd[1] + iv[8 + i] * vd[2];
            float n[3], alb[3];
            int id;
            float t = trace(org, d, n, alb, &id);
            if (material) memset(material + 4 * pix, 0, 4);
            if (t < 0) {
                /* ptMiss.rmiss:10-15 */
                float px = 1.0e10f - org[0], py = 1.0e10f - org[1], pz = 1.0e10f - org[2];
                depth[pix] = sqrtf(px * px + py * py + pz * pz);
                normal[2 * pix + 0] = acosf(1.0f);
                normal[2 * pix + 1] = atan2f(1.0f, 1.0f);
                albedo[4 * pix + 0] = albedo[4 * pix + 1] = albedo[4 * pix + 2] = 0;
                albedo[4 * pix + 3] = 255;
                illum[4 * pix + 0] = illum[4 * pix + 1] = illum[4 * pix + 2] = 0.0f;
                illum[4 * pix + 3] = 1.0f;
                continue;
            }
            float p[3] = {org[0] + t * d[0], org[1] + t * d[1], org[2] + t * d[2]};
            float dx = p[0] - org[0], dy = p[1] - org[1], dz = p[2] - org[2];
            depth[pix] = sqrtf(dx * dx + dy * dy + dz * dz);
            float nz = n[2] > 1.0f ? 1.0f : (n[2] < -1.0f ? -1.0f : n[2]);
            normal[2 * pix + 0] = acosf(nz);
            normal[2 * pix + 1] = atan2f(n[1], n[0]);
            for (int k = 0; k < 3; ++k) albedo[4 * pix + k] = (uint8_t)(alb[k] * 255.0f + 0.5f);
            albedo[4 * pix + 3] = 255;
            /* smooth analytic irradiance from one point light + sky ambient, shadowed by the spheres */
            float l[3] = {light_pos[0] - p[0], light_pos[1] - p[1], light_pos[2] - p[2]};
            float ld = sqrtf(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
            for (int k = 0; k < 3; ++k) l[k] /= ld;
            float ndl = n[0] * l[0] + n[1] * l[1] + n[2] * l[2];
            float ps[3] = {p[0] + 1e-3f * n[0], p[1] + 1e-3f * n[1], p[2] + 1e-3f * n[2]};
            float direct = (ndl > 0 && !occluded(ps, l, ld)) ? 30.0f * ndl / (ld * ld) : 0.0f;
            float ambient = 0.15f + 0.1f * n[2];
            /* 1-spp Monte-Carlo noise: the light sample is taken with probability q and weighted
             * 1/q; the ambient term gets a uniform-sphere style multiplicative estimator. */
            float u0 = rnd01((uint32_t)x, (uint32_t)y, (uint32_t)frame, seed, 0u);
            float u1 = rnd01((uint32_t)x, (uint32_t)y, (uint32_t)frame, seed, 1u);
            const float q = 0.35f;
            float dsample = (u0 < q) ? direct / q : 0.0f;
            float asample = ambient * 2.0f * u1;
            float e[3] = {dsample * 1.0f + asample * 0.8f, dsample * 0.95f + asample * 0.9f, dsample * 0.85f + asample * 1.0f};
            for (int k = 0; k < 3; ++k) illum[4 * pix + k] = e[k] < 1e3f ? e[k] : 1e3f;
            illum[4 * pix + 3] = 1.0f;
        }
}

/* The noise-free expectation of the illumination (same scene); handy as a quality reference
 * in examples, never used by parity tests. */
SYNTH_API uint32_t vkpbrt_synth_version(void) { return 1u; }
